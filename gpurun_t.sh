mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 900 > gpurun_out/t_all.log 2>&1; echo "all gpu tests rc $?"; tail -n 5 gpurun_out/t_all.log | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_r01_m.json 2> gpurun_out/bench_err.log; echo "bench rc $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r01_m.json')); print(d['value'], 'img/s', d['ms_per_denoise_step'], 'ms/step; e2e', d['e2e']['value'], 'tail', d['ms_tail_vae_x2_plus_eq1'], 'roofline', d['roofline']['achieved'], d['roofline']['frac'], d['clocks'])"
timeout 600 python profiles/layer_times.py > gpurun_out/layer_times_final.txt 2>&1; head -1 gpurun_out/layer_times_final.txt
