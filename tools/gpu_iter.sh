#!/bin/bash
# one GPU iteration: parity tests on the current library, then bench.py over every library variant in build_variants/
mkdir -p gpurun_out/iter
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/iter/pytest_gpu.log 2>&1; tail -3 gpurun_out/iter/pytest_gpu.log
bash tools/variants.sh 2>&1 | tee gpurun_out/iter/variants.log
