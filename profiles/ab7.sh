#!/bin/bash
run() { echo "== $*"; env "$@" python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/frame %.2f e2e %.2f' % (d['ms_per_step'], d['e2e']['ms_per_step']))"; }
run RT795_DUMMY=1
run RT795_FULL_GRID=1
run RT795_FULL_GRID=1 RT795_TRAVERSAL_FLAGS=2
run RT795_TRAVERSAL=classic
run RT795_PIPES=3 RT795_QUEUE_CAP=2097152
