mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_r01_2gpu.json 2> gpurun_out/bench2_err.log; echo "bench2 rc $?"; tail -c 1500 gpurun_out/bench_r01_2gpu.json; tail -3 gpurun_out/bench2_err.log | cut -c1-300
