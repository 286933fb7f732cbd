import sys,json
for l in sys.stdin.read().strip().splitlines():
    if l.startswith('{'):
        d=json.loads(l); print("value %.2f G/s" % (d["value"]/1e9), "ms/step %.3f" % d["ms_per_step"], "kernel ms %.3f" % d["roofline"]["kernel_ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "fallback %.4f" % d.get("untiled_deposit_fraction",0), "e2e %.2f G/s" % (d["e2e"]["value"]/1e9))
    else: print(l[:300])
