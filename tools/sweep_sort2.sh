run() { out=$(env "$@" timeout -s KILL 100 python bench.py --steps 60 --warmup 6 --no-cpu-baseline 2>/dev/null); echo "$* $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms/step %.3f kernel %.3f e2e_ms %.3f fallback %.5f halo %d' % (d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['e2e']['ms_per_step'], d.get('untiled_deposit_fraction',0), d.get('tile_halo',0)))")"; }
run SFGPU_SORT_EVERY=3
run SFGPU_SORT_EVERY=4
run SFGPU_SORT_EVERY=4 SFGPU_SORT_PREDICT=2.5
run SFGPU_SORT_EVERY=5 SFGPU_SORT_PREDICT=3
run SFGPU_SORT_EVERY=2
