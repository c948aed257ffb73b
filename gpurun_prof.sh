mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:gemm_kernel|attn_kernel|gn_|layernorm_kernel|sched_kernel|hdr_kernel|splitk_finalize|pack_unet|nchw|softmax_rows|timestep_embedding|silu_kernel' -c 3000 --csv --log-file gpurun_out/launches_bench_r01.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu bench rc $?"
wc -l gpurun_out/launches_bench_r01.csv
