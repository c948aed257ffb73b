mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc $?"; tail -3 gpurun_out/smoke.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_first.json 2> gpurun_out/bench_err.log; echo "bench rc $?"; tail -5 gpurun_out/bench_err.log; cat gpurun_out/bench_r01_first.json
