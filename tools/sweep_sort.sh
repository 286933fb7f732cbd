#!/bin/bash
# bench.py over sort intervals for one library ($1)
for se in ${SE:-1 2 3}; do
  out=$(SFGPU_SORT_EVERY=$se SFGPU_LIB_PATH=$PWD/$1 timeout -s KILL 100 python bench.py --steps 12 --warmup 4 --no-cpu-baseline 2>/dev/null)
  echo "$1 sort_every=$se $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms/step %.3f kernel %.3f e2e_ms %.3f fallback %.4f' % (d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['e2e']['ms_per_step'], d.get('untiled_deposit_fraction',0)))" 2>&1 | tail -1)"
done
