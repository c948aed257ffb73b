mkdir -p gpurun_out
python profiles/prof_step.py > gpurun_out/plain_step.log 2>&1; tail -1 gpurun_out/plain_step.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none --profile-from-start off -c 1200 --csv --log-file gpurun_out/launches_step_r01d_warm.csv python profiles/prof_step.py > gpurun_out/ncu_step2.log 2>&1; echo "ncu warm rc $?"
