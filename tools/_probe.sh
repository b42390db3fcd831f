python -m pytest tests/test_gpu_pyramid.py -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline --no-scan-context 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('S=128 value %.0f e2e %.0f frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac']))"
