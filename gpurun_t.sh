mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sched.py tests/test_gpu_pipeline.py -m gpu -q --tb=short -p no:cacheprovider --timeout 600 > gpurun_out/t_sp.log 2>&1; echo "sched/pipe tests rc $?"; tail -n 12 gpurun_out/t_sp.log | cut -c1-400
