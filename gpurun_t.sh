mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hdr.py -k "exposure or gradient" -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/t_hdr.log 2>&1; echo "hdr tests rc $?"; tail -n 25 gpurun_out/t_hdr.log | cut -c1-600
