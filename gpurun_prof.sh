mkdir -p gpurun_out
python profiles/prof_step.py > gpurun_out/plain_step.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_step.log; exit 1; }
tail -1 gpurun_out/plain_step.log
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1200 --csv --log-file gpurun_out/launches_step_r01_final.csv python profiles/prof_step.py > gpurun_out/ncu_step.log 2>&1; echo "ncu rc $?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none --profile-from-start off -c 1200 --csv --log-file gpurun_out/launches_step_r01_final_warm.csv python profiles/prof_step.py > gpurun_out/ncu_step2.log 2>&1; echo "ncu warm rc $?"
python profiles/prof_gemm.py > gpurun_out/plain_gemm.log 2>&1; tail -2 gpurun_out/plain_gemm.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_kernel --launch-skip 4 --launch-count 1 -o gpurun_out/prof_conv_halo_r01 -f python profiles/prof_gemm.py > gpurun_out/ncu_gemm.log 2>&1; echo "ncu conv rc $?"
ls -la gpurun_out/*.ncu-rep | tail -3
