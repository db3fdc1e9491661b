set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2_gpus2.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/mgpu_worker.py > gpurun_out/r2_mgpu2.log 2>&1; echo "rc=$?" >> gpurun_out/r2_mgpu2.log
grep "mgpu\]" gpurun_out/r2_mgpu2.log | tail -30; tail -5 gpurun_out/r2_mgpu2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2_bench_g2.json 2> gpurun_out/r2_bench_g2.err; echo "rc=$?" >> gpurun_out/r2_bench_g2.err
tail -5 gpurun_out/r2_bench_g2.err; cut -c 1-600 gpurun_out/r2_bench_g2.json
