#!/usr/bin/env python
"""Injection of N particles per call: host-sampled particles through sfgpu_inject (pageable numpy arrays, the upload is part of
the call) against the device-side UniformSource (sfgpu_source_uniform).  Wall clock around the C-ABI calls, median of 7."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from starfish_b200 import KineticMaterial, Particles, synthetic as S  # noqa: E402
from starfish_b200.domain import DomainType, LinearSpline  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
m = S.make_mesh(513, 513, DomainType.XY, 1e-3, "open")
sp = LinearSpline([(0.0, 0.5), (0.0, 0.3), (0.0, 0.01)])
with KineticMaterial("O+", S.QE, 16 * S.AMU, [m], DomainType.XY, capacity_hint=16 * n) as km:
    km.dt = 1e-7
    th, td, state = [], [], 0x5DEECE66D
    for k in range(7):
        y = np.linspace(0.01, 0.5, n)
        p = Particles(n, x=np.full(n, 7e-10), y=y, z=np.zeros(n), u=np.full(n, 7000.0), v=np.zeros(n), w=np.zeros(n), mpw=np.full(n, 1e3))
        km.sync()
        t0 = time.perf_counter()
        km.addParticles(m, p, 1e-7)
        km.sync()
        th.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        added, state = km.sampleUniformSource(sp, 7000.0, n, state, dt=1e-7, mpw=1e3)
        km.sync()
        td.append(time.perf_counter() - t0)
    print("n=%d  host-sampled sfgpu_inject %.3f ms (%.1f M particles/s)   device UniformSource %.3f ms (%.1f M particles/s)" % (
        n, 1e3 * np.median(th), n / np.median(th) / 1e6, 1e3 * np.median(td), n / np.median(td) / 1e6))
