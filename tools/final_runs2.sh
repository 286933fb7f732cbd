#!/bin/bash
mkdir -p gpurun_out/final2
timeout -s KILL 120 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/final2/smoke.log
SFGPU_STREAM_CHECK=1 timeout -s KILL 300 python -m pytest tests -m gpu -x -q > gpurun_out/final2/pytest_gpu.log 2>&1; tail -2 gpurun_out/final2/pytest_gpu.log
timeout -s KILL 200 python bench.py --steps 20 --warmup 5 2>gpurun_out/final2/bench_n1.err > gpurun_out/final2/bench_n1.json; cut -c1-250 gpurun_out/final2/bench_n1.json
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 300 --csv --log-file gpurun_out/final2/launches.csv python bench.py --steps 8 --warmup 4 --no-cpu-baseline > /dev/null 2>&1
for f in 0 8; do timeout -s KILL 300 python bench.py --workload c --steps 12 --warmup 4 --no-cpu-baseline --step-flags $f > gpurun_out/final2/bench_c_flags$f.json 2>/dev/null; cut -c1-200 gpurun_out/final2/bench_c_flags$f.json; done
