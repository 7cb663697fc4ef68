for t in 1000 1500 2000 3000; do
  echo "tile $t"; MRP_FG_NORM_TILE=$t python bench.py --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(d['ms_per_step'], d['roofline']['kernels_ms'])"
done
