#!/bin/bash
# pipelines x wave size on the full C4 frame
for cfg in "" "RT795_PIPES=3" "RT795_PIPES=4" "RT795_PIPES=4 RT795_QUEUE_CAP=2097152" "RT795_PIPES=3 RT795_QUEUE_CAP=2097152" "RT795_QUEUE_CAP=8388608" "RT795_PIPES=1"; do
  echo "== $cfg: $(env $cfg python profiles/tile_share.py c4 1 5 2>&1 | grep 'timing=0' | cut -c1-90)"
done
