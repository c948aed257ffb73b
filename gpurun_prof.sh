mkdir -p gpurun_out
python profiles/prof_gn.py 0,3 1 > gpurun_out/plain_gn.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_gn.log; exit 1; }
timeout 600 ncu --set full --clock-control none --profile-from-start off --launch-count 4 -k regex:gn_ -o gpurun_out/prof_gn_r01 -f python profiles/prof_gn.py 0,3 1 > gpurun_out/ncu_gn.log 2>&1; echo "ncu rc $?"
ls -la gpurun_out/prof_gn_r01.ncu-rep
