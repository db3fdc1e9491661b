cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_round2.py -x -q -m gpu -k "ortho or fused or arnoldi" > gpurun_out/r6_t_ortho.log 2>&1; tail -2 gpurun_out/r6_t_ortho.log
timeout 300 python tools/sweep.py --grid 4096 --cycles 4 "" "ortho_variant=2" "ortho_variant=3" > gpurun_out/r6_sweep_fused.log 2>&1; cat gpurun_out/r6_sweep_fused.log
timeout 300 python tools/sweep.py --grid 4096 --cycles 3 --complex-storage "" > gpurun_out/r6_sweep_fused_c.log 2>&1; cat gpurun_out/r6_sweep_fused_c.log
