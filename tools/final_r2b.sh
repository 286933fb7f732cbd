#!/bin/bash
# closing measurements after the fused counting pass (the parity + segment files ran green on this build just before: profiles/r2_final_pytest_gpu.log)
mkdir -p gpurun_out/final_r2
timeout -s KILL 200 python -m pytest tests -m gpu -x -q --timeout 150 --ignore=tests/test_gpu_parity.py --ignore=tests/test_gpu_segments.py > gpurun_out/final_r2/pytest_gpu_rest.log 2>&1; tail -2 gpurun_out/final_r2/pytest_gpu_rest.log
timeout -s KILL 200 python bench.py 2>gpurun_out/final_r2/bench_n1.err > gpurun_out/final_r2/bench_n1.json; cut -c1-200 gpurun_out/final_r2/bench_n1.json
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:k_fast_step -s 6 -c 3 -f -o gpurun_out/final_r2/prof python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/final_r2/prof.log 2>&1
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/final_r2/launches.csv python bench.py --steps 12 --warmup 6 --no-cpu-baseline > /dev/null 2>&1
timeout -s KILL 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_r2/smoke.log 2>&1; tail -2 gpurun_out/final_r2/smoke.log
for w in c e; do timeout -s KILL 200 python bench.py --workload $w --no-cpu-baseline > gpurun_out/final_r2/bench_${w}_n1.json 2>/dev/null; cut -c1-160 gpurun_out/final_r2/bench_${w}_n1.json; done
