#!/bin/bash
# TraceIO.flags: bit 0 dynamic fetch, bit 1 postponed leaves, bits 8..15 fetch threshold (0 -> 20)
t() { echo $(( $1 + ($2 << 8) )); }
for cfg in "" "RT795_FLAGS_OCCLUSION=$(t 3 12)" "RT795_FLAGS_OCCLUSION=$(t 3 16)" "RT795_FLAGS_OCCLUSION=$(t 3 24)" "RT795_FLAGS_OCCLUSION=$(t 3 28)" "RT795_FLAGS_OCCLUSION=2" "RT795_FLAGS_OCCLUSION=1" \
           "RT795_FLAGS_CLOSEST=$(t 3 12)" "RT795_FLAGS_CLOSEST=$(t 3 20)" "RT795_FLAGS_CLOSEST=$(t 3 28)" "RT795_FLAGS_CLOSEST=0"; do
  echo "== $cfg: $(env $cfg python profiles/tile_share.py c4 1 4 2>&1 | grep 'timing' | sed 's/c4 share 1\/1 //; s/host call best [0-9.]* ms, waves 32, launches 129, //' | tr '\n' ' ' | cut -c1-170)"
done
