mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_pipeline.py -m gpu -q --tb=short -p no:cacheprovider --timeout 600 -x > gpurun_out/t_up.log 2>&1; echo "unet/pipe tests rc $?"; tail -n 5 gpurun_out/t_up.log | cut -c1-400
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_l.json 2> gpurun_out/bench_err.log; echo "bench rc $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r01_l.json')); print(d['value'], 'img/s', d['ms_per_denoise_step'], 'ms/step; tail', d['ms_tail_vae_x2_plus_eq1'], 'roofline', d['roofline']['achieved'])"
