mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 600 -x > gpurun_out/t_all.log 2>&1; echo "all gpu tests rc $?"; tail -n 8 gpurun_out/t_all.log | cut -c1-300
