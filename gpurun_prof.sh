mkdir -p gpurun_out
python profiles/prof_gemm_small.py > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 3 -o gpurun_out/prof_gemm_small_r01b python profiles/prof_gemm_small.py > gpurun_out/ncu.log 2>&1; echo "ncu rc $?"; cat gpurun_out/plain.log; tail -2 gpurun_out/ncu.log
