set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_gpu.txt
timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -s > gpurun_out/r2_t_round2.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t_round2.log
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_partial_schur.py -x -q > gpurun_out/r2_t_old.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t_old.log
timeout 600 python tools/spmv_sweep.py --matrix lap2d --size 4096 "spmv_variant=2,spmv_threads=128" "spmv_variant=1" "" "spmv_threads=128" "spmv_stages=2" "spmv_stages=4" "spmv_tile=640" "spmv_tile=2560" "spmv_tile=2560,spmv_stages=2" "spmv_bps=2" "spmv_bps=3" "spmv_threads=128,spmv_tile=640,spmv_stages=4" > gpurun_out/r2_spmv_sweep_lap.log 2>&1
timeout 300 python tools/spmv_sweep.py --matrix mark --size 4000 --modes real "spmv_variant=2,spmv_threads=128" "" "spmv_threads=128" "spmv_stages=4" "spmv_tile=2048" > gpurun_out/r2_spmv_sweep_mark.log 2>&1
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "rc=$?" >> gpurun_out/r2_bench1.err
tail -c 1500 gpurun_out/r2_t_round2.log
tail -c 600 gpurun_out/r2_t_old.log
cat gpurun_out/r2_spmv_sweep_lap.log
