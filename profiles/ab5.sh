#!/bin/bash
run() { echo "== $*"; env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms']
print('ms/frame %.2f  extend %.2f shade %.2f shadow %.2f raygen %.2f' % (d['ms_per_step'], s['ms_extend'], s['ms_shade'], s['ms_shadow'], s['ms_raygen']))"; }
run RT795_QUEUE_CAP=4194304
run RT795_QUEUE_CAP=8388608
run RT795_QUEUE_CAP=16777216
run RT795_QUEUE_CAP=2097152
run RT795_QUEUE_CAP=1048576
