"""Single-caller multi-GPU mirror: one Python (Java) thread drives every GPU of the box through ``sfgpu_multi_*``
(include/sfgpu.h): meshes and fields replicated, particles partitioned by index, the deposit all-reduced inside the step,
results read back once from rank 0."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .kinetic_material import NFIELDS, Particles, SfgpuError, _ptr


class MultiGpuKineticMaterial:
    def __init__(self, name, charge, mass, mesh, n_gpus, device_ids=None, capacity_hint=0, step_flags=0):
        self.lib = _lib.load()
        self.name, self.charge, self.mass, self.mesh = name, float(charge), float(mass), mesh
        self.step_flags = int(step_flags)
        self._g = C.c_void_p()
        ids = None if device_ids is None else np.ascontiguousarray(device_ids, np.int32)
        rc = self.lib.sfgpu_multi_create(int(n_gpus), None if ids is None else ids.ctypes.data_as(_lib.c_int32_p), int(mesh.domain_type), C.byref(self._g))
        if rc:
            raise SfgpuError(rc, self.lib.sfgpu_last_error(None).decode())
        m = mesh
        self._keep = [np.ascontiguousarray(b, np.int8) for b in m.bc] + [np.ascontiguousarray(b, np.int32) for b in m.nbr]
        bc = (C.c_void_p * 4)(*[_ptr(a) for a in self._keep[:4]])
        nbr = (C.c_void_p * 4)(*[_ptr(a) for a in self._keep[4:]])
        has_seg, node_vol = np.ascontiguousarray(m.has_seg, np.uint8), np.ascontiguousarray(m.node_vol, np.float64)
        x0, dh = np.ascontiguousarray(m.x0, np.float64), np.ascontiguousarray(m.dh, np.float64)
        mid, sp = C.c_int32(-1), C.c_int32(-1)
        self._check(self.lib.sfgpu_multi_mesh_add(self._g, m.ni, m.nj, x0.ctypes.data_as(_lib.c_double_p), dh.ctypes.data_as(_lib.c_double_p), bc, nbr,
                                                  _ptr(has_seg), _ptr(node_vol), C.byref(mid)))
        self._check(self.lib.sfgpu_multi_species_add(self._g, self.charge, self.mass, int(capacity_hint), C.byref(sp)))
        self._sp = sp.value
        self.setFields()
        self.dt = 0.0

    def _check(self, rc):
        if rc:
            raise SfgpuError(rc, self.lib.sfgpu_multi_last_error(self._g).decode())

    @property
    def n_gpus(self):
        return int(self.lib.sfgpu_multi_size(self._g))

    def setFields(self):
        m = self.mesh
        efi, efj = np.ascontiguousarray(m.efi, np.float64), np.ascontiguousarray(m.efj, np.float64)
        self._check(self.lib.sfgpu_multi_set_fields(self._g, 0, _ptr(efi), _ptr(efj), None, None))

    def addParticles(self, p: Particles, dt=None):
        added = C.c_int64()
        v = p.view()
        self._check(self.lib.sfgpu_multi_inject(self._g, self._sp, 0, C.byref(v), float(self.dt if dt is None else dt), _lib.INJECT_REWIND, C.byref(added)))
        return added.value

    def updateFields(self, dt=None):
        self._check(self.lib.sfgpu_multi_step(self._g, self._sp, float(self.dt if dt is None else dt), self.step_flags))
        sums = (C.c_double * 5)()
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self.lib.sfgpu_multi_get_sums(self._g, self._sp, sums, C.byref(a), C.byref(b), C.byref(c)))
        self.sums5 = np.array(sums[:])
        self.np_alive, self.n_exited, self.n_slow = a.value, b.value, c.value

    def deposit(self):
        m = self.mesh
        dep = np.empty((NFIELDS, m.ni, m.nj))
        ptrs = (C.c_void_p * NFIELDS)(*[dep[f].ctypes.data for f in range(NFIELDS)])
        self._check(self.lib.sfgpu_multi_get_deposit(self._g, self._sp, 0, ptrs))
        return dep

    def density(self):
        nd = np.empty((self.mesh.ni, self.mesh.nj))
        self._check(self.lib.sfgpu_multi_get_moments(self._g, self._sp, 0, _ptr(nd), None, None, None))
        return nd

    def getParticles(self):
        """All particles of all GPUs (per-rank sfgpu_download through sfgpu_multi_ctx), concatenated."""
        parts = []
        for r in range(self.n_gpus):
            ctx = C.c_void_p(self.lib.sfgpu_multi_ctx(self._g, r))
            n = C.c_int64()
            rc = self.lib.sfgpu_np(ctx, self._sp, 0, C.byref(n))
            if rc:
                raise SfgpuError(rc, self.lib.sfgpu_last_error(ctx).decode())
            p = Particles.empty(n.value)
            if n.value:
                v = p.view()
                rc = self.lib.sfgpu_download(ctx, self._sp, 0, 0, C.byref(v))
                if rc:
                    raise SfgpuError(rc, self.lib.sfgpu_last_error(ctx).decode())
            parts.append(p)
        keys = ("x", "y", "z", "u", "v", "w", "mpw", "li", "lj", "dt", "id", "born_it")
        return Particles(sum(p.n for p in parts), **{k: np.concatenate([getattr(p, k) for p in parts]) for k in keys})

    def close(self):
        if self._g:
            self.lib.sfgpu_multi_destroy(self._g)
            self._g = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
