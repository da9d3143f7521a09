#!/bin/bash
for v in envrej envdir envrej envdir; do
  RT795_CORE_SO=raytracer-795_b200/lib_ab/$v/librt795.so python profiles/ab_variant.py c5 3 32 2>&1 | grep -v '^\[' | tail -1 | cut -c1-220
done
