#!/bin/bash
# A/B of kernel variants on one GPU box.  usage (here):   tools/ab.sh "-DX=0" "-DX=1" ...   builds variants/lib_<k>.so
# then run on the box:   bash tools/ab_run.sh
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
k=0
for f in "$@"; do
  MRP_FG_NVCC_EXTRA="$f" python -c "import importlib; b=importlib.import_module('micro-regime-propogation_b200.build'); b.build_cuda(force=True)"
  cp micro-regime-propogation_b200/csrc/libmrp_fg.so variants/lib_$k.so
  echo "$f" > variants/lib_$k.txt
  k=$((k+1))
done
# leave the tree with the DEFAULT library (the flags stamp next to the .so makes build() notice a leftover variant anyway)
python -c "import importlib; b=importlib.import_module('micro-regime-propogation_b200.build'); b.build_cuda()"
