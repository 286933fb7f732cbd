#!/bin/bash
# round-1 closing measurements on one B200: tests, default bench, reference arm, ncu launch list, streaming path, configs C and E
mkdir -p gpurun_out/final
SFGPU_STREAM_CHECK=1 timeout -s KILL 300 python -m pytest tests -m gpu -x -q > gpurun_out/final/pytest_gpu.log 2>&1; tail -2 gpurun_out/final/pytest_gpu.log
timeout -s KILL 200 python bench.py --steps 20 --warmup 5 2>gpurun_out/final/bench_n1.err > gpurun_out/final/bench_n1.json; cut -c1-250 gpurun_out/final/bench_n1.json
timeout -s KILL 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/bench_ref.json 2>/dev/null; cut -c1-200 gpurun_out/final/bench_ref.json
timeout -s KILL 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --step-flags 8 > gpurun_out/final/bench_stream_n1.json 2>/dev/null; cut -c1-250 gpurun_out/final/bench_stream_n1.json
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 300 --csv --log-file gpurun_out/final/launches.csv python bench.py --steps 8 --warmup 4 --no-cpu-baseline > /dev/null 2>&1
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 300 --csv --log-file gpurun_out/final/launches_stream.csv python bench.py --steps 8 --warmup 4 --no-cpu-baseline --step-flags 8 > /dev/null 2>&1
for w in c e; do for f in 0 8; do timeout -s KILL 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --step-flags $f > gpurun_out/final/bench_${w}_flags$f.json 2>/dev/null; cut -c1-200 gpurun_out/final/bench_${w}_flags$f.json; done; done
