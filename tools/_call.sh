set -x
mkdir -p gpurun_out
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_round2.py -x -q -m gpu -k "ortho or fused or arnoldi" > gpurun_out/r3_t_ortho.log 2>&1; tail -3 gpurun_out/r3_t_ortho.log
timeout 600 python tools/sweep.py --grid 1448 --cycles 6 "" "ortho_variant=1" > gpurun_out/r3_sweep_small2.log 2>&1; cat gpurun_out/r3_sweep_small2.log
