out=gpurun_out/r2v_packet2.txt; : > $out
for w in c1 c2 c3 c5; do
for p in 0 1; do
RT795_PACKET=$p python bench.py --config $w --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms']; print('$w packet=$p', round(d['ms_per_step'],3), 'extend', round(s['ms_extend'],2))" >> $out
done; done
