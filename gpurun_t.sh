mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -k "gemm or conv" -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/t_tc.log 2>&1; echo "tc tests rc $?"; tail -n 8 gpurun_out/t_tc.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_unet.py tests/test_gpu_pipeline.py -m gpu -q --tb=short -p no:cacheprovider --timeout 300 -x > gpurun_out/t_up.log 2>&1; echo "unet/pipe tests rc $?"; tail -n 4 gpurun_out/t_up.log | cut -c1-300
timeout 600 python profiles/layer_times.py > gpurun_out/layer_times_r01m.txt 2>&1; echo "layer rc $?"; head -14 gpurun_out/layer_times_r01m.txt
