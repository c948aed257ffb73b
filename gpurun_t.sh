mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py -k "sdr_to_hdr" -m gpu -q --tb=short -p no:cacheprovider --timeout 600 > gpurun_out/t_p.log 2>&1; echo "pipe tests rc $?"; tail -n 30 gpurun_out/t_p.log | cut -c1-800
