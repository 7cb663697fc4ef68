#!/bin/bash
# on the GPU box: bench every variants/lib_<k>.so (kernel times only)
cd "$(dirname "$0")/.."
for so in variants/lib_*.so; do
  cp $so micro-regime-propogation_b200/csrc/libmrp_fg.so
  echo "== $so $(cat ${so%.so}.txt)"
  if [ -n "$AB_TEST" ]; then timeout 300 python -m pytest $AB_TEST -m gpu -x -q 2>&1 | tail -1; fi
  timeout 200 python bench.py --no-e2e --no-cpu-baseline ${AB_ARGS} 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(d['ms_per_step'], d['roofline']['kernels_ms'])"
done
