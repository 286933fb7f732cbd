#!/bin/bash
# one GPU iteration: smoke (bounded), parity tests on the current library, then bench.py over every library variant in build_variants/
mkdir -p gpurun_out/iter
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/iter/smoke.log 2>&1 || { echo "SMOKE FAILED/HUNG"; tail -5 gpurun_out/iter/smoke.log; exit 1; }
tail -2 gpurun_out/iter/smoke.log
timeout -s KILL 600 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/iter/pytest_gpu.log 2>&1; tail -3 gpurun_out/iter/pytest_gpu.log
bash tools/variants.sh 2>&1 | tee gpurun_out/iter/variants.log
