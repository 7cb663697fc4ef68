#!/bin/bash
# on the GPU box: round-2 (second session) evidence — tests, bench lines, ncu launch list, one full ncu capture per kernel at 50 M rows
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; tail -1 gpurun_out/r2b_tests.log
timeout 900 python bench.py > gpurun_out/r2b_bench_50Mrows.json 2> gpurun_out/r2b_bench.err; cut -c1-200 gpurun_out/r2b_bench_50Mrows.json
timeout 600 python bench.py --impl reference > gpurun_out/r2b_bench_reference_arm.json 2> gpurun_out/r2b_bench_ref.err; cut -c1-200 gpurun_out/r2b_bench_reference_arm.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/r2b_launches_50Mrows.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2b_launches.log 2>&1
timeout 300 python bench.py --days 32 --rows 46600 --no-e2e --no-cpu-baseline > gpurun_out/r2b_bench_multidate_32x46600.json 2>/dev/null; cut -c1-120 gpurun_out/r2b_bench_multidate_32x46600.json
timeout 300 python bench.py --days 32 --rows 1500000 --no-e2e --no-cpu-baseline > gpurun_out/r2b_bench_multidate_32x1500000.json 2>/dev/null; cut -c1-120 gpurun_out/r2b_bench_multidate_32x1500000.json
timeout 300 python bench.py --levels 50 --rows 20000000 --no-e2e --no-cpu-baseline > gpurun_out/r2b_bench_deep_book_L50_20Mrows.json 2>/dev/null; cut -c1-120 gpurun_out/r2b_bench_deep_book_L50_20Mrows.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_book|k_windows|k_normalize_runs" -s 9 -c 3 -f -o gpurun_out/r2b_50M python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2b_ncu.log 2>&1; tail -2 gpurun_out/r2b_ncu.log | cut -c1-100
