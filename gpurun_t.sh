mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensorcore.py -k "conv" -m gpu -q --tb=short -p no:cacheprovider --timeout 600 > gpurun_out/t_c.log 2>&1; echo "conv tests rc $?"; tail -n 12 gpurun_out/t_c.log | cut -c1-600
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_sched.py -m gpu -q --tb=short -p no:cacheprovider --timeout 600 > gpurun_out/t_u.log 2>&1; echo "unet tests rc $?"; tail -n 25 gpurun_out/t_u.log | cut -c1-600
