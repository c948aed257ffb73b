mkdir -p gpurun_out
python profiles/prof_epi_ncu.py > gpurun_out/plain_epi.log 2>&1 || { echo "plain failed"; exit 1; }
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/prof_epi_r01 -f python profiles/prof_epi_ncu.py > gpurun_out/ncu_epi.log 2>&1; echo "ncu rc $?"
ls -la gpurun_out/prof_epi_r01.ncu-rep
