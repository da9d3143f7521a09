#!/bin/bash
# A/B of wave size and pipeline count for one rank's 1/8 share of C4.
for cfg in "" "RT795_PIPES=4" "RT795_QUEUE_CAP=2097152" "RT795_QUEUE_CAP=2097152 RT795_PIPES=4" "RT795_QUEUE_CAP=1048576 RT795_PIPES=4" "RT795_PIPES=1"; do
  echo "== $cfg"
  env $cfg python profiles/tile_share.py c4 8 6
done
env python profiles/tile_share.py c4 1 4
