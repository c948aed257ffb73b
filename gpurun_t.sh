mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 profiles/bench_cfg_pair.py > gpurun_out/pair_bench.log 2>&1; echo "rc $?"; grep "B=" gpurun_out/pair_bench.log
