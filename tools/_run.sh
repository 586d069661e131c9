timeout 300 python -m pytest tests/test_gpu_env.py -x -q -k "host_session or step_host" 2>&1 | tail -3
timeout 200 python tools/trace_session.py 4096 300 2>&1 | tail -13
timeout 300 python bench.py --steps 300 --warmup 10 --no-sweep --no-cpu-baseline --no-configs > gpurun_out/r2_v8.json 2> gpurun_out/r2_v8.err
tail -3 gpurun_out/r2_v8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_v8.json'))
print('value %.4g ms %.5f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))
e=d['e2e']; print('e2e', e['value'], e['us_per_step'], e['mode'], e['host_us'])
l=d['e2e_launch_per_step']; print('launch/step', l['value'], l['us_per_step'], l['host_us'], 'b2b', l['back_to_back']['us_per_step'], 'pipe', l['pipelined']['us_per_step'])
PY
