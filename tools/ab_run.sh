#!/bin/bash
# on the GPU box: bench every variants/lib_<k>.so (kernel times only)
cd "$(dirname "$0")/.."
for so in variants/lib_*.so; do
  cp $so micro-regime-propogation_b200/csrc/libmrp_fg.so
  echo "== $so $(cat ${so%.so}.txt)"
  if [ -n "$AB_TEST" ]; then timeout 300 python -m pytest $AB_TEST -m gpu -x -q 2>&1 | tail -1; fi
  timeout 200 python bench.py --no-e2e --no-cpu-baseline ${AB_ARGS} 2>gpurun_out/ab_err.log | python -c "import json,sys; d=json.load(sys.stdin); print(d['ms_per_step'], d['roofline']['kernels_ms'])"
done
# AB_NCU=1: instruction-cache metrics of k_windows for every variant (8 M rows)
if [ -n "$AB_NCU" ]; then
for so in variants/lib_*.so; do
  cp $so micro-regime-propogation_b200/csrc/libmrp_fg.so
  echo "== ncu $so $(cat ${so%.so}.txt)"
  timeout 300 ncu --metrics gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,sm__icc_request_hit_rate.pct,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:${AB_KERNEL:-k_windows} --launch-skip 3 -c 1 --csv python bench.py --rows 8000000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | grep -E "^\"[0-9]" | awk -F'","' '{printf "%s=%s  ", $(NF-2), $NF} END{print ""}' | sed 's/"//g'
done
fi
