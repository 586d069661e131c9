set -u
python -m pytest tests/test_gpu_env.py -x -q -k "single_pass or fused_obs or ragged or plan or boundary" 2>&1 | tail -4
B="python bench.py --steps 300 --warmup 10 --no-sweep --no-cpu-baseline"
$B > gpurun_out/r2_v1.json 2> gpurun_out/r2_v1.err
SAFELIFE_B200_PDL=0 $B --no-configs > gpurun_out/r2_v2.json 2> gpurun_out/r2_v2.err
SAFELIFE_B200_EPC=8 $B --no-configs > gpurun_out/r2_v3.json 2> gpurun_out/r2_v3.err
SAFELIFE_B200_EPC=8 SAFELIFE_B200_PDL=0 $B --no-configs > gpurun_out/r2_v4.json 2> gpurun_out/r2_v4.err
python - <<'PY'
import json
for f,name in (('v1','pdl+epc'),('v2','epc only'),('v3','pdl only'),('v4','neither')):
    try:
        d=json.load(open('gpurun_out/r2_%s.json'%f))
        print(name, 'value %.4g ms %.5f frac %.3f | flushed %s | obs %.5f (%.3f) flushed %s | b2b %.5f graph %s | e2e %.4g' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['single_step_flushed']['ms_per_step'], d['step_with_obs']['ms'], d['step_with_obs']['frac'], d['step_with_obs']['single_step_flushed_ms'], d['back_to_back']['ms_per_step'], d['graph_replay'].get('ms_per_step'), d['e2e']['value']), d['roofline']['kernel'])
        for c in d.get('configs', []):
            print('  cfg', c['config'], 'ms %.5f frac %.3f flushed %s parity %s' % (c['ms_per_step'], c['roofline']['frac'], c['single_step_flushed_ms'], c['parity_spot_check']['warp_kernels_equal_generic_kernel']), c['kernel'], c.get('episode_end',{}).get('board_steps_per_s'))
    except Exception as e:
        print(name, 'FAILED', e)
PY
tail -3 gpurun_out/r2_v1.err
python tools/trace_step.py 4096 2>&1 | tail -9
