#!/bin/bash
# round-2 closing measurements on one B200: tests, default bench (+ cpu baseline), reference arm, ncu capture + launch list of the same command
mkdir -p gpurun_out/final_r2
timeout -s KILL 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/final_r2/pytest_gpu.log 2>&1; tail -2 gpurun_out/final_r2/pytest_gpu.log
grep -q " passed" gpurun_out/final_r2/pytest_gpu.log && ! grep -q " failed\| error" gpurun_out/final_r2/pytest_gpu.log || { echo "GPU tests not green: stopping"; tail -30 gpurun_out/final_r2/pytest_gpu.log; exit 1; }
timeout -s KILL 300 python bench.py 2>gpurun_out/final_r2/bench_n1.err > gpurun_out/final_r2/bench_n1.json; cut -c1-200 gpurun_out/final_r2/bench_n1.json
timeout -s KILL 300 python bench.py --impl reference > gpurun_out/final_r2/bench_ref.json 2>/dev/null; cut -c1-200 gpurun_out/final_r2/bench_ref.json
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_fast_step -s 6 -c 3 -f -o gpurun_out/final_r2/prof python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/final_r2/prof.log 2>&1
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/final_r2/launches.csv python bench.py --steps 12 --warmup 6 --no-cpu-baseline > /dev/null 2>&1
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_r2/smoke.log 2>&1; tail -2 gpurun_out/final_r2/smoke.log
for w in c e; do timeout -s KILL 600 python bench.py --workload $w --no-cpu-baseline > gpurun_out/final_r2/bench_${w}_n1.json 2>/dev/null; cut -c1-160 gpurun_out/final_r2/bench_${w}_n1.json; done
SFGPU_HALO=2 timeout -s KILL 300 python bench.py --no-cpu-baseline 2>/dev/null > gpurun_out/final_r2/bench_n1_halo2.json; cut -c1-160 gpurun_out/final_r2/bench_n1_halo2.json
