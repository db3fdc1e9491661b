cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r7_t_all.log 2>&1; tail -3 gpurun_out/r7_t_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r7_smoke.log 2>&1; tail -1 gpurun_out/r7_smoke.log
