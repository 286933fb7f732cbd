#!/bin/bash
# runs bench.py once per library variant in build_variants/ (SFGPU_LIB_PATH) and prints ms/step and kernel ms
for so in build_variants/lib_*.so; do
  out=$(SFGPU_LIB_PATH=$PWD/$so timeout -s KILL 100 python bench.py --steps 20 --warmup 5 --no-cpu-baseline "$@" 2>/dev/null)
  echo "$so $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms/step %.3f kernel %.3f e2e_ms %.3f fallback %.4f' % (d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['e2e']['ms_per_step'], d.get('untiled_deposit_fraction',0)))" 2>&1 | tail -1)"
done
