import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from starfish_b200 import KineticMaterial, Particles, synthetic as S
wl = S.config_b(ni=int(os.environ.get("NI", "129")), nj=int(os.environ.get("NI", "129")))
m = wl.mesh
n = int(os.environ.get("NP", str(1 << 20)))
km = KineticMaterial("O+", wl.charge, wl.mass, [m], m.domain_type, capacity_hint=n)
km.dt = wl.dt
km.addParticles(m, Particles(n, **wl.particles(0, n)), wl.dt)
for it in range(8):
    km.step_raw(wl.dt)
    tot, ker, _ = km.lastStepTiming()
    print(it, "np", km.getNp(), "fallback", km.lastStepFallback(), "frac %.5f" % (km.lastStepFallback() / n), "kernel ms %.3f" % ker, "kind", km.lastStepKernel())
km.close()
