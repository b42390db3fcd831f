timeout 100 python tools/one_eval.py kitti 128 2>&1 | tail -4
for b in 3 2 1 8; do
DSLAM_EVAL_CTAS_PER_SM=$b timeout 150 python bench.py --streams 512 --no-cpu-baseline --no-scan-context > gpurun_out/q_$b.json 2> gpurun_out/q_$b.err
python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/q_$b.json').read()); print('ctas/sm=$b value %.0f e2e %.0f frac %.3f us %.1f' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_us']), d['host_ms_per_step'])
except Exception as e:
    print('b=$b FAILED', open('gpurun_out/q_$b.err').read()[-300:])"
done
