mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -k "conv or gemm" -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/t_tc.log 2>&1; echo "tc tests rc $?"; tail -n 14 gpurun_out/t_tc.log | cut -c1-400
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_pipeline.py -m gpu -q --tb=short -p no:cacheprovider --timeout 600 -x > gpurun_out/t_up.log 2>&1; echo "unet/pipe tests rc $?"; tail -n 8 gpurun_out/t_up.log | cut -c1-400
timeout 600 python profiles/layer_times.py > gpurun_out/layer_times_r01n.txt 2>&1; echo "layer rc $?"; head -12 gpurun_out/layer_times_r01n.txt
