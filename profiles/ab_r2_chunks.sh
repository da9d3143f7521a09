#!/bin/bash
# A/B: CTA-chunked ray cursor (TraceIO.flags bits 16..23 = chunk / 256) on C4, then C2.
out=gpurun_out/r2l_chunks.txt; : > $out
run() { name=$1; cfg=$2; shift 2; env "$@" python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms']; print('$cfg', '$name', round(d['ms_per_step'],3), 'extend', round(s['ms_extend'],2), 'shade', round(s['ms_shade'],2), 'shadow', round(s['ms_shadow'],2))" >> $out; }
run base c4 RT795_NOP=1
for c in 2 4 8 16 32; do run chunk$((c*256)) c4 RT795_FLAGS_CLOSEST=$((c*65536+2)) RT795_FLAGS_OCCLUSION=$((c*65536+3)); done
run base c4 RT795_NOP=1
run base c2 RT795_NOP=1
for c in 4 16; do run chunk$((c*256)) c2 RT795_FLAGS_CLOSEST=$((c*65536+2)) RT795_FLAGS_OCCLUSION=$((c*65536+3)); done
python -m pytest tests/test_gpu_parity.py -x -q -k "c1 or c2 or c4" 2>&1 | tail -2 >> $out
RT795_FLAGS_CLOSEST=$((8*65536+2)) RT795_FLAGS_OCCLUSION=$((8*65536+3)) python -m pytest tests/test_gpu_parity.py -x -q -k "c1 or c2 or c4 or edge" 2>&1 | tail -2 >> $out
