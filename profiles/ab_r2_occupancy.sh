out=gpurun_out/r2j_minb.txt; : > $out
run() { name=$1; shift; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms']; print('$name', round(d['ms_per_step'],3), 'extend', round(s['ms_extend'],2), 'shade', round(s['ms_shade'],2), 'shadow', round(s['ms_shadow'],2))" >> $out; }
run base RT795_NOP=1
for v in any8 closest8 both8 any6 shade6 shade4; do run $v RT795_CORE_SO=raytracer-795_b200/lib_ab/$v/librt795.so; done
run thr28 RT795_FLAGS_OCCLUSION=7171 RT795_FLAGS_CLOSEST=7170
run thr14 RT795_FLAGS_OCCLUSION=3587 RT795_FLAGS_CLOSEST=3586
run base2 RT795_NOP=1
