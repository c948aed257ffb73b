mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 600 -x > gpurun_out/t_all.log 2>&1; echo "all gpu tests rc $?"; tail -n 5 gpurun_out/t_all.log | cut -c1-300
( time python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r01_e.json 2> gpurun_out/bench_err.log ) 2>&1 | grep real; echo "bench rc $?"; tail -3 gpurun_out/bench_err.log; python -c "
import json; d=json.load(open('gpurun_out/bench_r01_e.json')); print({k:d[k] for k in ['value','ms_per_step','ms_per_denoise_step','ms_tail_vae_x2_plus_eq1']}); print(d['cpu_baseline'])"
( time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_ref.json 2> gpurun_out/bench_ref_err.log ) 2>&1 | grep real; tail -2 gpurun_out/bench_ref_err.log; cut -c1-400 gpurun_out/bench_r01_ref.json
