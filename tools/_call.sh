set -x
mkdir -p gpurun_out
cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "spmv_window" > gpurun_out/r3_t_win.log 2>&1; tail -3 gpurun_out/r3_t_win.log
timeout 900 $TR --master-port 29517 tests/mgpu_worker.py > gpurun_out/r3_mgpu2.log 2>&1; echo "rc=$?" >> gpurun_out/r3_mgpu2.log; grep "mgpu\]" gpurun_out/r3_mgpu2.log | tail -30; tail -3 gpurun_out/r3_mgpu2.log
timeout 200 $TR --master-port 29531 tools/comm_bench.py > gpurun_out/r3_comm2.json 2> gpurun_out/r3_comm2.err; cat gpurun_out/r3_comm2.json
timeout 300 python tools/event_cost.py > gpurun_out/r3_evcost1.json 2> gpurun_out/r3_evcost1.err; cat gpurun_out/r3_evcost1.json
timeout 300 $TR --master-port 29532 tools/event_cost.py > gpurun_out/r3_evcost2.json 2> gpurun_out/r3_evcost2.err; cat gpurun_out/r3_evcost2.json
timeout 600 $TR --master-port 29533 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r3_bench_g2.json 2> gpurun_out/r3_bench_g2.err; echo "rc=$?" >> gpurun_out/r3_bench_g2.err; tail -3 gpurun_out/r3_bench_g2.err; cut -c 1-300 gpurun_out/r3_bench_g2.json
