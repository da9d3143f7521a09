#!/bin/bash
# PLOC search radius vs tree quality (frame time) and build time on C4
for r in 4 8 16 32 64; do
  echo "== radius $r"
  RT795_PLOC_RADIUS=$r python profiles/bvh_build.py c4 3 | python -c "
import json,sys
d=json.load(sys.stdin)['builders']['device_ploc']
print({k: (round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k in ('ms_bvh_device','ms_bvh_build','ploc_iterations','nodes4','ms_frame','closest_node_visits_per_ray','shadow_node_visits_per_ray','closest_tri_tests_per_ray')})"
done
