#!/bin/bash
# A/B of the SAH leaf rule with the HOST builder (bench --bvh host): max leaf size and node/triangle cost ratio.
out=gpurun_out/r2s_sah.txt; : > $out
run() { name=$1; shift; env "$@" python bench.py --bvh host --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms']; r=d['roofline']; o=r['other_traversal_kernel']['issue']
print('$name', round(d['ms_per_step'],3), 'extend', round(s['ms_extend'],2), 'shade', round(s['ms_shade'],2), 'shadow', round(s['ms_shadow'],2), 'K4 visits/tests', round(r['node_visits_per_ray'],2), round(r['prim_tests_per_ray'],2), 'K2', round(o['node_visits_per_ray'],2), round(o['prim_tests_per_ray'],2), 'bvh MB', round(d['config']['scene_bytes']['bvh']/1e6,1))" >> $out; }
run base RT795_NOP=1
run leaf8_ct1 RT795_SAH_MAX_LEAF=8
run leaf8_ct2 RT795_SAH_MAX_LEAF=8 RT795_SAH_TRAVERSAL_COST=2
run leaf8_ct4 RT795_SAH_MAX_LEAF=8 RT795_SAH_TRAVERSAL_COST=4
run leaf4_ct2 RT795_SAH_TRAVERSAL_COST=2
run leaf4_ct05 RT795_SAH_TRAVERSAL_COST=0.5
run leaf2 RT795_SAH_MAX_LEAF=2
run leaf6_ct2 RT795_SAH_MAX_LEAF=6 RT795_SAH_TRAVERSAL_COST=2
