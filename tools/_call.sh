set -x
mkdir -p gpurun_out
cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29532 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r5_bench_g8.json 2> gpurun_out/r5_bench_g8.err; echo "rc=$?" >> gpurun_out/r5_bench_g8.err; cut -c 1-200 gpurun_out/r5_bench_g8.json
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR2 --master-port 29534 bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e --no-converged > gpurun_out/r5_bench_g2.json 2> gpurun_out/r5_bench_g2.err; echo "rc=$?" >> gpurun_out/r5_bench_g2.err; cut -c 1-200 gpurun_out/r5_bench_g2.json
timeout 100 python bench.py --steps 8 --warmup 3 --no-e2e --no-converged --no-cpu --no-parity > gpurun_out/r5_bench_g1.json 2> gpurun_out/r5_bench_g1.err; cut -c 1-200 gpurun_out/r5_bench_g1.json
