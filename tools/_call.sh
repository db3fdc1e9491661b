set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29531 tools/comm_bench.py > gpurun_out/r2_comm8.json 2> gpurun_out/r2_comm8.err; cat gpurun_out/r2_comm8.json
timeout 400 $TR --master-port 29532 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r2_bench_g8.json 2> gpurun_out/r2_bench_g8.err; echo "rc=$?" >> gpurun_out/r2_bench_g8.err; cut -c 1-300 gpurun_out/r2_bench_g8.json
timeout 300 $TR --master-port 29533 tools/solve_dist.py --matrix mark --grid 4000 --nev 20 --max-dim 60 --real-arith pairs --fast-real-schur > gpurun_out/r2_cfg3_8gpu_pairs.json 2> gpurun_out/r2_cfg3_8gpu_pairs.err; cut -c 1-400 gpurun_out/r2_cfg3_8gpu_pairs.json
timeout 400 $TR --master-port 29534 tools/solve_dist.py --matrix powerlaw --rows 100000000 --top-base 2 --top-step 0.01 --real-arith pairs --fast-real-schur > gpurun_out/r2_cfg4_8gpu.json 2> gpurun_out/r2_cfg4_8gpu.err; cut -c 1-500 gpurun_out/r2_cfg4_8gpu.json
timeout 300 $TR --master-port 29535 tools/stepbench_dist.py --rows 100000000 --cmax 100 > gpurun_out/r2_step8_1e8.json 2> gpurun_out/r2_step8_1e8.err; cut -c 1-300 gpurun_out/r2_step8_1e8.json
timeout 300 $TR --master-port 29536 tools/stepbench_dist.py --rows 200000000 --cmax 100 > gpurun_out/r2_step8_2e8.json 2> gpurun_out/r2_step8_2e8.err; cut -c 1-300 gpurun_out/r2_step8_2e8.json
