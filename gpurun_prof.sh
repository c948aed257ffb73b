mkdir -p gpurun_out
python profiles/prof_attn.py > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:attn_kernel -s 2 -c 1 -o gpurun_out/prof_attn_r01 python profiles/prof_attn.py > gpurun_out/ncu.log 2>&1; echo "ncu rc $?"; tail -2 gpurun_out/ncu.log
