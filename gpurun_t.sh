mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -k "attention" -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/t_attn.log 2>&1; echo "attn tests rc $?"; tail -n 5 gpurun_out/t_attn.log | cut -c1-400
timeout 600 python profiles/layer_times.py > gpurun_out/layer_times_r01ac.txt 2>&1; echo "layer rc $?"; head -1 gpurun_out/layer_times_r01ac.txt; grep -E "attn.*Nk=77" gpurun_out/layer_times_r01ac.txt | head -8
