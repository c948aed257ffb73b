mkdir -p gpurun_out
python profiles/prof_smallm.py > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,launch__grid_size --clock-control none -s 3 -c 6 --csv --log-file gpurun_out/smallm.csv python profiles/prof_smallm.py > gpurun_out/ncu.log 2>&1; echo "ncu rc $?"; grep -E "gemm_kernel|splitk" gpurun_out/smallm.csv | cut -d, -f5,13,15 | cut -c1-160
