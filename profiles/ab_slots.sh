#!/bin/bash
for v in "$@"; do
  RT795_CORE_SO=raytracer-795_b200/lib_ab/$v/librt795.so python profiles/ab_variant.py c4 5 2>&1 | grep -v '^\[' | tail -1
done
