mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 600 -x > gpurun_out/t_all.log 2>&1; echo "all gpu tests rc $?"; tail -n 12 gpurun_out/t_all.log | cut -c1-300
for i in 1 2; do python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_d$i.json 2> gpurun_out/bench_err.log; python -c "
import json; d=json.load(open('gpurun_out/bench_r01_d$i.json')); print({k:d[k] for k in ['value','ms_per_step','ms_per_denoise_step','ms_tail_vae_x2_plus_eq1']}, d['clocks'])"; done
