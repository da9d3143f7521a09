out=gpurun_out/r2z_qnodes.txt; : > $out
run() { name=$1; cfg=$2; shift 2; env "$@" python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms']; print('$name $cfg', round(d['ms_per_step'],3), 'extend', round(s['ms_extend'],2), 'shade', round(s['ms_shade'],2), 'shadow', round(s['ms_shadow'],2))" >> $out; }
run base c4 RT795_NOP=1
run qnodes c4 RT795_QNODES=1
run base c4 RT795_NOP=1
