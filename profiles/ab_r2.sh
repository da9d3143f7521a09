#!/bin/bash
# A/B of the round-2 driver / kernel switches on one GPU box.  Usage: profiles/ab_r2.sh <tag> [config]
# Each line of gpurun_out/<tag>_ab.txt: the switch set, ms/frame, Mrays/s, stage times (one extra frame with per-stage events).
tag=${1:-ab}; cfg=${2:-c4}
out=gpurun_out/${tag}_ab.txt; : > $out
run() {
  env "$@" python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/_ab.json 2> gpurun_out/_ab.err || { echo "$* FAILED" >> $out; tail -3 gpurun_out/_ab.err >> $out; return; }
  python - "$*" >> $out <<'PY'
import json, sys
d = json.load(open("gpurun_out/_ab.json"))
s = d["roofline"]["stage_ms"]
print(f"{sys.argv[1]:55s} {d['ms_per_step']:8.3f} ms  {d['value']:9.1f} Mrays/s  e2e {d['e2e']['ms_per_step']:8.3f}  extend {s['ms_extend']:.2f} shade {s['ms_shade']:.2f} shadow {s['ms_shadow']:.2f} raygen {s['ms_raygen']:.2f}")
PY
}
if [ "$cfg" = c2 ]; then
run RT795_NOP=1
run RT795_HINTS=0
run RT795_DRIVER=chunked
run RT795_PIPES=3
run RT795_QUEUE_CAP=8388608
else
run RT795_NOP=1
run RT795_SMEM_TLAS=1
run RT795_FUSED_RAYGEN=0
run RT795_PIPES=1
fi
cat $out
