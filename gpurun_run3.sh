mkdir -p gpurun_out
python tests/diag_unet.py 2>&1 | grep -E "run-to-run|nondeterministic|batch3|hw=" | head
for sel in "tests/test_gpu_tensorcore.py -k 'groupnorm or layernorm or gemm'" "tests/test_gpu_unet.py" "tests/test_gpu_pipeline.py"; do
  name=$(echo "$sel" | tr ' /' '__' | tr -d "'")
  eval timeout 900 python -m pytest $sel -m gpu -q --tb=short -p no:cacheprovider --timeout 600 > gpurun_out/t_$name.log 2>&1
  echo "== $sel -> rc $?"; tail -n 30 gpurun_out/t_$name.log | cut -c1-300 | grep -v "^$"
done
