mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -k "conv" -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/t_tc.log 2>&1; echo "conv tests rc $?"; tail -n 8 gpurun_out/t_tc.log | cut -c1-600
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_pipeline.py -m gpu -q --tb=short -p no:cacheprovider --timeout 600 -x > gpurun_out/t_up.log 2>&1; echo "unet/pipe tests rc $?"; tail -n 8 gpurun_out/t_up.log | cut -c1-400
timeout 600 python profiles/layer_times.py > gpurun_out/layer_times_r01s.txt 2>&1; echo "layer rc $?"; head -1 gpurun_out/layer_times_r01s.txt; grep -E "conv3 \(8, 16, 16\)|conv3 \(8, 64, 64\) K=2880|conv3 \(8, 32" gpurun_out/layer_times_r01s.txt | head
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r01_i.json 2> gpurun_out/bench_err.log; echo "bench rc $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r01_i.json')); print(d['value'], 'img/s', d['ms_per_denoise_step'], 'ms/step', d['roofline'])"
