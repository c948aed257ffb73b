mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -k "gemm or conv" -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/t_tc.log 2>&1; echo "tc tests rc $?"; tail -n 6 gpurun_out/t_tc.log | cut -c1-600
timeout 600 python profiles/layer_times.py > gpurun_out/layer_times_r01z.txt 2>&1; echo "layer rc $?"; head -1 gpurun_out/layer_times_r01z.txt; grep -E "f32out" gpurun_out/layer_times_r01z.txt | head -8
