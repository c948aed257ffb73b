mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -k "halo" -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/t_tc.log 2>&1; echo "halo tests rc $?"; tail -n 12 gpurun_out/t_tc.log | cut -c1-700
