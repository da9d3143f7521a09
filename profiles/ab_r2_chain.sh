#!/bin/bash
# A/B: chained state transitions in the traversal loop (new, in-tree) vs the previous build (lib_ab/prechain)
out=gpurun_out/r2o_chain.txt; : > $out
run() { name=$1; cfg=$2; shift 2; env "$@" python bench.py --config $cfg $( [ $cfg = c5 ] && echo --spp 32 ) --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms']; print('$cfg', '$name', round(d['ms_per_step'],3), 'extend', round(s['ms_extend'],2), 'shade', round(s['ms_shade'],2), 'shadow', round(s['ms_shadow'],2))" >> $out; }
python -m pytest tests/test_gpu_parity.py tests/test_gpu_exact.py -x -q 2>&1 | tail -2 >> $out
for cfg in c4 c2 c5 c3; do
run chained $cfg RT795_NOP=1
run previous $cfg RT795_CORE_SO=raytracer-795_b200/lib_ab/prechain/librt795.so
done
run chained c4 RT795_NOP=1
