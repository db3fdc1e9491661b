cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29517 tests/mgpu_worker.py > gpurun_out/r6_mgpu2.log 2>&1; echo "rc=$?" >> gpurun_out/r6_mgpu2.log; grep "mgpu\]" gpurun_out/r6_mgpu2.log | tail -6; tail -2 gpurun_out/r6_mgpu2.log
timeout 300 $TR --master-port 29518 tools/solve_dist.py --matrix powerlaw --rows 4000000 --top-base 2 --top-step 0.01 --real-arith pairs --fast-real-schur > gpurun_out/r6_pl2.json 2> gpurun_out/r6_pl2.err; cut -c 1-900 gpurun_out/r6_pl2.json | tail -2
