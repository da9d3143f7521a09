out=gpurun_out/r2y_ldg256.txt; : > $out
run() { name=$1; cfg=$2; shift 2; env "$@" python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms']; print('$name $cfg', round(d['ms_per_step'],3), 'extend', round(s['ms_extend'],2), 'shade', round(s['ms_shade'],2), 'shadow', round(s['ms_shadow'],2))" >> $out; }
V=$PWD/raytracer-795_b200/lib_ab
for v in "$@"; do
run base c4 RT795_NOP=1
run $v c4 RT795_CORE_SO=$V/$v/librt795.so
done
run base c4 RT795_NOP=1
for v in "$@"; do
run base c2 RT795_NOP=1
run $v c2 RT795_CORE_SO=$V/$v/librt795.so
done
