mkdir -p gpurun_out
for sel in "tests/test_gpu_tensorcore.py" "tests/test_gpu_unet.py" "tests/test_gpu_pipeline.py"; do
  name=$(echo "$sel" | tr ' /' '__' | tr -d "'")
  eval timeout 900 python -m pytest $sel -m gpu -q --tb=short -p no:cacheprovider --timeout 600 > gpurun_out/t_$name.log 2>&1
  echo "== $sel -> rc $?"; tail -n 15 gpurun_out/t_$name.log | cut -c1-300 | grep -v "^$"
done
python profiles/layer_times.py > gpurun_out/layer_times_r01e.txt 2>&1; echo "layer rc $?"; head -3 gpurun_out/layer_times_r01e.txt; grep -E "attn B=16 Nq=4096 Nk=4096|groupnorm|layernorm" gpurun_out/layer_times_r01e.txt | head -12
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_b.json 2> gpurun_out/bench_err.log; echo "bench rc $?"; tail -3 gpurun_out/bench_err.log; python -c "
import json; d=json.load(open('gpurun_out/bench_r01_b.json')); print({k:d[k] for k in ['value','ms_per_step','ms_per_denoise_step','gpu_launches','clocks']}); print(d['e2e']['value'], d['roofline']['achieved'], d['kernels'])"
