#!/bin/bash
for cfg in "RT795_AFFINE_TRIS=0" "RT795_AFFINE_TRIS=1" "RT795_AFFINE_TRIS=0" "RT795_AFFINE_TRIS=1"; do
  echo "== $cfg: $(env $cfg python profiles/tile_share.py c4 1 4 2>&1 | grep 'timing' | sed 's/c4 share 1\/1 //; s/host call best [0-9.]* ms, waves 32, launches 129, //' | tr '\n' ' ' | cut -c1-170)"
done
