cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "spmv_window" > gpurun_out/r7_t_win.log 2>&1; tail -3 gpurun_out/r7_t_win.log
