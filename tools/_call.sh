cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_partial_schur.py -x -q -m gpu -k "spmv_window or powerlaw" > gpurun_out/r8_t_win.log 2>&1; tail -2 gpurun_out/r8_t_win.log
timeout 300 python tools/spmv_sweep.py --matrix powerlaw --size 10000000 --modes real "" "spmv_tile=256" "spmv_variant=4" > gpurun_out/r8_spmv.log 2>&1; cat gpurun_out/r8_spmv.log
