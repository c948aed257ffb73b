mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -k "conv or gemm or groupnorm" -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/t_tc.log 2>&1; echo "tc tests rc $?"; tail -n 5 gpurun_out/t_tc.log | cut -c1-400
timeout 600 python profiles/layer_times.py > gpurun_out/layer_times_r01q.txt 2>&1; echo "layer rc $?"; head -1 gpurun_out/layer_times_r01q.txt; grep -E "geglu" gpurun_out/layer_times_r01q.txt | head -8
python profiles/bench_gn.py 2>&1 | head -6
