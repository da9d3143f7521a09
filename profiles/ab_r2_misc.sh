out=gpurun_out/r2u_misc.txt; : > $out
run() { name=$1; shift; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms']; print('$name', round(d['ms_per_step'],3), 'extend', round(s['ms_extend'],2), 'shade', round(s['ms_shade'],2), 'shadow', round(s['ms_shadow'],2))" >> $out; }
run base RT795_NOP=1
run pipes3 RT795_PIPES=3
run pipes4 RT795_PIPES=4
run cap16 RT795_QUEUE_CAP=16777216
run cap4 RT795_QUEUE_CAP=4194304
run cap4_pipes4 RT795_QUEUE_CAP=4194304 RT795_PIPES=4
run cap2_pipes4 RT795_QUEUE_CAP=2097152 RT795_PIPES=4
run smem RT795_SMEM_TLAS=1
run base RT795_NOP=1
