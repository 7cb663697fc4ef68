#!/bin/bash
# on the GPU box: the round's final evidence (tests, bench lines, ncu launch list, one full ncu capture per kernel)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; tail -1 gpurun_out/final_tests.log
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; cut -c1-200 gpurun_out/final_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; cut -c1-200 gpurun_out/final_bench_ref.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/final_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_ --launch-skip 12 -c 4 -f -o gpurun_out/prof_r1_final2 python bench.py --rows 8000000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/final_ncu.log 2>&1; tail -2 gpurun_out/final_ncu.log | cut -c1-100
