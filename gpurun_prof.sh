mkdir -p gpurun_out
python profiles/prof_attn.py > gpurun_out/plain_attn.log 2>&1 || { echo "plain attn failed"; exit 1; }
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/prof_attn_r01b -f python profiles/prof_attn.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc $?"
python profiles/prof_gemm_small.py > gpurun_out/plain_gs.log 2>&1 || { echo "plain gemm failed"; exit 1; }
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_kernel --launch-skip 8 --launch-count 1 -o gpurun_out/prof_geglu_r01 -f python profiles/prof_gemm_small.py > gpurun_out/ncu_gs.log 2>&1; echo "ncu geglu rc $?"
ls -la gpurun_out/*.ncu-rep
