#!/bin/bash
cd "$(dirname "$0")/.."
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/final_launches.log 2>&1
timeout 300 python bench.py --days 32 --rows 46600 --no-e2e --no-cpu-baseline > gpurun_out/final_multidate_32x46600.json 2>/dev/null; cut -c1-160 gpurun_out/final_multidate_32x46600.json
timeout 300 python bench.py --days 32 --rows 1500000 --no-e2e --no-cpu-baseline > gpurun_out/final_multidate_32x1500000.json 2>/dev/null; cut -c1-160 gpurun_out/final_multidate_32x1500000.json
timeout 300 python bench.py --levels 50 --rows 20000000 --no-e2e --no-cpu-baseline > gpurun_out/final_deep_book.json 2>/dev/null; cut -c1-160 gpurun_out/final_deep_book.json
