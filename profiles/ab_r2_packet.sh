out=gpurun_out/r2v_packet.txt; : > $out
run() { name=$1; shift; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms']; print('$name', round(d['ms_per_step'],3), 'extend', round(s['ms_extend'],2), 'shade', round(s['ms_shade'],2), 'shadow', round(s['ms_shadow'],2))" >> $out; }
run base RT795_PACKET=0
run packet RT795_PACKET=1
run base RT795_PACKET=0
run packet RT795_PACKET=1
for w in c1 c2 c3 c5; do
for p in 0 1; do
RT795_PACKET=$p python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$w packet=$p', round(d['ms_per_step'],3))" >> $out
done; done
RT795_PACKET=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_exact.py -x -q -m gpu 2>&1 | tail -8 >> $out
