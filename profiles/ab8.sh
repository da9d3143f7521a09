#!/bin/bash
run() { echo "== $*"; env "$@" python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms']
print('ms/frame %.2f  extend %.2f shade %.2f shadow %.2f' % (d['ms_per_step'], s['ms_extend'], s['ms_shade'], s['ms_shadow']))"; }
run RT795_DUMMY=1
run RT795_TRAVERSAL_FLAGS=11
run RT795_TRAVERSAL_FLAGS=10
