mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 900 -x > gpurun_out/t_all.log 2>&1; echo "all gpu tests rc $?"; tail -n 6 gpurun_out/t_all.log | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01_g.json 2> gpurun_out/bench_err.log; echo "bench rc $?"; cut -c1-900 gpurun_out/bench_r01_g.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r01_ref_g.json 2> gpurun_out/bench_ref_err.log; echo "ref bench rc $?"; cut -c1-600 gpurun_out/bench_r01_ref_g.json
