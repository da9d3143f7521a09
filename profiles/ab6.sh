#!/bin/bash
run() { echo "== $*"; env "$@" python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/frame %.2f e2e %.2f' % (d['ms_per_step'], d['e2e']['ms_per_step']))"; }
run RT795_PIPES=2
run RT795_PIPES=3
run RT795_PIPES=4
run RT795_PIPES=2 RT795_QUEUE_CAP=4194304
run RT795_PIPES=4 RT795_QUEUE_CAP=4194304
run RT795_PIPES=3 RT795_QUEUE_CAP=16777216
