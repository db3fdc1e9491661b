cd /root/repo
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r4_t_all2.log 2>&1; tail -3 gpurun_out/r4_t_all2.log
AB200_TRACE_DESTROY=1 timeout 600 python bench.py --steps 8 --no-cpu --no-converged --no-parity > gpurun_out/r4_trace.json 2> gpurun_out/r4_trace.err; grep "free V" gpurun_out/r4_trace.err | tail -5; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r4_trace.json') if l.startswith('{')][-1]); print(d['value'], d['e2e']['value'], d['e2e']['host_phases_s'])"
