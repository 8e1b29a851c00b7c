# e2e stability: the bench's e2e value and the slowest timed step of N runs on one box (LCD_BENCH_VERBOSE timelines)
for i in $(seq 1 ${1:-5}); do
  LCD_BENCH_VERBOSE=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-whole-program 2>/tmp/e2e.err | python -c "
import json,sys,re
steps=[float(m.group(1)) for m in re.finditer(r'\[e2e step \d+\] ([0-9.]+) ms', open('/tmp/e2e.err').read())]
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('run $i: value %.1f e2e %.1f; timed steps (last 5): %s' % (d['value'], d['e2e']['value'], steps[-5:]))
"
done
