mkdir -p gpurun_out
for sel in "tests/test_gpu_tensorcore.py -k 'gemm or conv'" "tests/test_gpu_unet.py" "tests/test_gpu_pipeline.py"; do
  name=$(echo "$sel" | tr ' /' '__' | tr -d "'")
  eval timeout 600 python -m pytest $sel -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/t_$name.log 2>&1
  echo "== $sel -> rc $?"; tail -n 15 gpurun_out/t_$name.log | cut -c1-300 | grep -v "^$"
done
timeout 300 python profiles/prof_gemm_small.py
timeout 600 python profiles/layer_times.py > gpurun_out/layer_times_r01g.txt 2>&1; echo "layer rc $?"; head -24 gpurun_out/layer_times_r01g.txt
