#!/bin/bash
# Wall time of the drop-in CLI (process start -> image file written) on C4, cold and with the scene blob cache.
python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import __graft_entry__ as ge, bench
rt = ge.load_package()
print(bench.make_scene(rt, "c4", bench.WORKLOADS["c4"][0], "full"))
PY
XML=$(python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import __graft_entry__ as ge, bench
rt = ge.load_package()
print(bench.make_scene(rt, "c4", bench.WORKLOADS["c4"][0], "full"))
PY
)
cd "$(dirname "$XML")"
rm -f "$XML.rt795blob"
B=/root/repo/raytracer-795_b200/bin/raytracer
for args in "--bvh host" "--bvh auto" "--bvh auto --cache" "--bvh auto --cache" "--bvh auto --cache --f32"; do
  s=$(date +%s.%N); $B "$XML" $args > /dev/null 2> /tmp/cli.err; rc=$?; e=$(date +%s.%N)
  echo "raytracer c4.xml $args: rc=$rc wall $(python -c "print(f'{$e-$s:.3f}')") s "; grep -v "^\[" /tmp/cli.err | sed "s/^/    /"
done
ls -la | head -12
