#!/bin/bash
# on the GPU box: kernel times against the segment length (rows per CTA of the windows / book kernels)
cd "$(dirname "$0")/.."
for s in ${SEGS:-16384 14080 12288 11264 9472 8192}; do
  echo -n "segment_rows $s: "
  MRP_FG_SEGMENT_ROWS=$s timeout 200 python bench.py --no-e2e --no-cpu-baseline ${AB_ARGS} 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(d['ms_per_step'], d['roofline']['kernels_ms'])"
done
