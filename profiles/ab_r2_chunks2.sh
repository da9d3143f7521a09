#!/bin/bash
out=gpurun_out/r2m_chunks.txt; : > $out
run() { name=$1; cfg=$2; shift 2; env "$@" python bench.py --config $cfg $( [ $cfg = c5 ] && echo --spp 32 ) --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms']; print('$cfg', '$name', round(d['ms_per_step'],3), 'extend', round(s['ms_extend'],2), 'shade', round(s['ms_shade'],2), 'shadow', round(s['ms_shadow'],2))" >> $out; }
run base c4 RT795_NOP=1
for pair in "0 1" "0 2" "0 4" "1 1" "1 2" "2 2" "1 4" "2 4"; do set -- $pair; run "closest$(( $1*256 ))_occl$(( $2*256 ))" c4 RT795_FLAGS_CLOSEST=$(( $1*65536+2 )) RT795_FLAGS_OCCLUSION=$(( $2*65536+3 )); done
run base c4 RT795_NOP=1
for pair in "0 2" "2 2" "1 1"; do set -- $pair; run "closest$(( $1*256 ))_occl$(( $2*256 ))" c5 RT795_FLAGS_CLOSEST=$(( $1*65536+2 )) RT795_FLAGS_OCCLUSION=$(( $2*65536+3 )); done
run base c5 RT795_NOP=1
run base c3 RT795_NOP=1
run closest512_occl512 c3 RT795_FLAGS_CLOSEST=$(( 2*65536+2 )) RT795_FLAGS_OCCLUSION=$(( 2*65536+3 ))
