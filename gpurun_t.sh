mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -k "attention" -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/t_attn.log 2>&1; echo "attn tests rc $?"; tail -n 12 gpurun_out/t_attn.log | cut -c1-500
timeout 600 python profiles/layer_times.py > gpurun_out/layer_times_r01y.txt 2>&1; echo "layer rc $?"; head -1 gpurun_out/layer_times_r01y.txt; grep -E "attn" gpurun_out/layer_times_r01y.txt | head -8
timeout 600 python -m pytest tests/test_gpu_unet.py -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/t_u.log 2>&1; echo "unet tests rc $?"; tail -n 8 gpurun_out/t_u.log | cut -c1-500
python tests/diag_unet.py 2>&1 | tail -6
