#!/bin/bash
# PLOC -> SAH hand-over point (clusters left when the host binned-SAH builds the top) vs tree quality (frame time) and build time on C4
for t in 1 256 4096 32768 131072; do
  echo "== top $t"
  RT795_PLOC_TOP=$t python profiles/bvh_build.py c4 3 | python -c "
import json,sys
d=json.load(sys.stdin)['builders']['device_ploc']
print({k: (round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k in ('ms_bvh_device','ms_bvh_build','ploc_iterations','nodes4','ms_frame','closest_node_visits_per_ray','shadow_node_visits_per_ray','closest_tri_tests_per_ray')})"
done
