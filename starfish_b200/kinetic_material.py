"""Host-side mirror of ``starfish.core.materials.KineticMaterial`` for the GPU path.

Same member names and meaning as the Java class (KineticMaterial.java): ``updateFields()`` :117,
``addParticle`` :759/:810/:826, ``getNp()`` :1297, ``getDen/getU/getV/getW`` (Material.java:718-721), the
velocity-moment running sums ``count-sum``.. ``mpc-sum`` (:95-110, :1570-1595), ``clearSamples()`` :1509,
``mass_sum / momentum_sum / energy_sum`` (:247-259).  All particle arithmetic happens in
``libstarfish_gpu.so``; this class only moves arrays across the C ABI and does the mesh-sized
bookkeeping the Java subclass would do (adding the per-step increments into the running sums).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import FIELD_NAMES, NFIELDS, Particles as _CParticles, SlowExtra as _CSlowExtra
from .domain import DomainType, UniformMesh

_PKEYS = ("x", "y", "z", "u", "v", "w", "mpw", "li", "lj", "dt")


class SfgpuError(RuntimeError):
    """Raised for every non-zero return of the C ABI (the Java glue maps this to Log.error)."""

    def __init__(self, code, msg):
        super().__init__(f"sfgpu error {code}: {msg}")
        self.code = code


class Particles:
    """SoA particle arrays on the host (the fields of KineticMaterial.Particle, KM:1207-1217)."""

    def __init__(self, n=0, **arrays):
        self.n = int(n)
        for k in _PKEYS:
            a = arrays.get(k)
            setattr(self, k, None if a is None else np.ascontiguousarray(a, dtype=np.float64))
        for k in ("id", "born_it"):
            a = arrays.get(k)
            setattr(self, k, None if a is None else np.ascontiguousarray(a, dtype=np.int32))
        for k in _PKEYS + ("id", "born_it"):
            a = getattr(self, k)
            if a is not None and a.shape != (self.n,):
                raise ValueError(f"Particles.{k} has shape {a.shape}, expected ({self.n},)")

    @classmethod
    def empty(cls, n):
        return cls(n, **{k: np.empty(n) for k in _PKEYS}, id=np.empty(n, np.int32), born_it=np.empty(n, np.int32))

    def view(self):
        v = _CParticles()
        v.n = self.n
        for k in _PKEYS:
            a = getattr(self, k)
            setattr(v, k, None if a is None else a.ctypes.data_as(_lib.c_double_p))
        for k in ("id", "born_it"):
            a = getattr(self, k)
            setattr(v, k, None if a is None else a.ctypes.data_as(_lib.c_int32_p))
        return v

    def sorted_by_id(self):
        o = np.argsort(self.id, kind="stable")
        return Particles(self.n, **{k: getattr(self, k)[o] for k in _PKEYS + ("id", "born_it")})

    def take(self, sel):
        d = {k: (None if getattr(self, k) is None else getattr(self, k)[sel]) for k in _PKEYS + ("id", "born_it")}
        return Particles(len(d["x"]), **d)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


_SAMPLE_NAMES = ("count-sum", "u-sum", "v-sum", "w-sum", "uu-sum", "vv-sum", "ww-sum", "mpc-sum")


class _Fields(dict):
    """Per-mesh result fields.  ``nd`` is refreshed by every updateFields() (the field solver reads it every step,
    SolverModule.java:144-171); the mean velocities u/v/w and the running velocity-moment sums (KM:1570-1595) stay on
    the device and are fetched when somebody reads them, like Java's computeFields() every 10 steps and the output writers."""

    def __init__(self, km, k):
        super().__init__()
        self._km, self._k = km, k

    def __getitem__(self, name):
        if name in _SAMPLE_NAMES:
            self._km._fetch_samples(self._k)
        elif name in ("u", "v", "w"):
            self._km._fetch_velocities(self._k)
        return super().__getitem__(name)


class _LazyDeposit:
    """last_deposit[k]: the raw per-step sums [8][ni][nj] of mesh k, downloaded on first access after a step."""

    def __init__(self, km):
        self._km = km

    def __getitem__(self, k):
        return self._km._fetch_deposit(k)

    def __len__(self):
        return len(self._km.meshes)


class KineticMaterial:
    """One kinetic species on one GPU.  ``meshes`` is the ordered mesh list (Starfish.getMeshList())."""

    def __init__(self, name, charge, mass, meshes, domain_type=DomainType.XY, spwt=1.0, device=0,
                 capacity_hint=0, step_flags=0, device_segments=True):
        self.lib = _lib.load()
        self.name = name
        self.charge, self.mass, self.spwt0 = float(charge), float(mass), float(spwt)
        self.q_over_m = self.charge / self.mass  # Material.java:711
        self.domain_type = DomainType(domain_type)
        self.meshes = list(meshes) if not isinstance(meshes, UniformMesh) else [meshes]
        self.step_flags = int(step_flags)
        self._ctx = C.c_void_p()
        rc = self.lib.sfgpu_create(int(device), int(self.domain_type), C.byref(self._ctx))
        if rc:
            raise SfgpuError(rc, self.lib.sfgpu_last_error(None).decode())
        for k, m in enumerate(self.meshes):
            if m.domain_type != self.domain_type:
                raise ValueError("mesh domain type differs from the material's")
            m.index = k
            bc = (C.c_void_p * 4)(*[_ptr(np.ascontiguousarray(b, np.int8)) for b in m.bc])
            self._keep = [np.ascontiguousarray(b, np.int8) for b in m.bc] + [np.ascontiguousarray(b, np.int32) for b in m.nbr]
            bc = (C.c_void_p * 4)(*[_ptr(a) for a in self._keep[:4]])
            nbr = (C.c_void_p * 4)(*[_ptr(a) for a in self._keep[4:]])
            has_seg = np.ascontiguousarray(m.has_seg, np.uint8)
            node_vol = np.ascontiguousarray(m.node_vol, np.float64)
            mid = C.c_int32(-1)
            x0 = np.ascontiguousarray(m.x0, np.float64)
            dh = np.ascontiguousarray(m.dh, np.float64)
            self._check(self.lib.sfgpu_mesh_add(self._ctx, m.ni, m.nj, x0.ctypes.data_as(_lib.c_double_p),
                                                dh.ctypes.data_as(_lib.c_double_p), bc, nbr, _ptr(has_seg), _ptr(node_vol),
                                                C.byref(mid)))
            assert mid.value == k
            # deterministic surface hits on the device (domain.set_boundaries built the node -> segment table)
            sg = getattr(m, "segments", None)
            if device_segments and sg is not None and len(sg["x1"]):
                ip = lambda a: np.ascontiguousarray(a, np.int32).ctypes.data_as(_lib.c_int32_p)
                dp = lambda a: np.ascontiguousarray(a, np.float64).ctypes.data_as(_lib.c_double_p)
                self._check(self.lib.sfgpu_mesh_set_segments(self._ctx, k, len(sg["x1"]), dp(sg["x1"]), dp(sg["y1"]), dp(sg["x2"]), dp(sg["y2"]),
                                                             ip(sg["kind"]), ip(sg["sink"]), ip(m.seg_offs), ip(m.seg_ids)))
        sp = C.c_int32(-1)
        self._check(self.lib.sfgpu_species_add(self._ctx, self.charge, self.mass, int(capacity_hint), C.byref(sp)))
        self._sp = sp.value
        for m in self.meshes:
            self.setFields(m)
        # per-mesh result fields, double[ni][nj] like Field2D.data, in page-locked memory (direct DMA, no staging)
        self._pinned = []
        self.fields = []
        for k, m in enumerate(self.meshes):
            f = _Fields(self, k)
            for name in ("nd", "u", "v", "w") + _SAMPLE_NAMES:
                dict.__setitem__(f, name, self.hostArray((m.ni, m.nj)))
            self.fields.append(f)
        self._dep = [self.hostArray((NFIELDS, m.ni, m.nj)) for m in self.meshes]
        self._dep_step = [-1] * len(self.meshes)
        self._samp_step = [-1] * len(self.meshes)
        self._vel_step = [-1] * len(self.meshes)
        self.download_fields = True  # False on the ranks of a multi-GPU run that do not feed the host solver (read back once)
        self._step_no = 0
        self.last_deposit = _LazyDeposit(self)
        self.mass_sum = 0.0
        self.momentum_sum = np.zeros(3)
        self.energy_sum = 0.0
        self.n_exited = 0
        self.n_slow = 0
        self.dt = 0.0
        self.slow_path_handler = None  # callable(km, Particles, extra dict) -> Particles of survivors (the Java ProcessBoundary)

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc):
        if rc:
            raise SfgpuError(rc, self.lib.sfgpu_last_error(self._ctx).decode())

    def hostArray(self, shape):
        """Zeroed float64 array in page-locked host memory (sfgpu_host_alloc): what a Java host would hold as a
        direct buffer for its Field2D mirrors.  Lives until close()."""
        n = int(np.prod(shape))
        p = C.c_void_p()
        rc = self.lib.sfgpu_host_alloc(n * 8, C.byref(p))
        if rc:
            raise SfgpuError(rc, self.lib.sfgpu_last_error(None).decode())
        self._pinned.append(p)
        a = np.ctypeslib.as_array((C.c_double * n).from_address(p.value)).reshape(shape)
        a[...] = 0.0
        return a

    @property
    def num_samples(self):
        n = C.c_int64()
        self._check(self.lib.sfgpu_get_samples(self._ctx, self._sp, 0, None, C.byref(n)))
        return n.value

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self.lib.sfgpu_destroy(self._ctx)
            self._ctx = C.c_void_p()
            for p in getattr(self, "_pinned", []):
                self.lib.sfgpu_host_free(p)
            self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ------------------------------------------------------------------ inputs
    def setFields(self, mesh):
        """Upload efi/efj(/bfi/bfj) of a mesh: what MeshData reads each step (KM:1319-1322)."""
        efi = np.ascontiguousarray(mesh.efi, np.float64)
        efj = np.ascontiguousarray(mesh.efj, np.float64)
        bfi = None if mesh.bfi is None else np.ascontiguousarray(mesh.bfi, np.float64)
        bfj = None if mesh.bfj is None else np.ascontiguousarray(mesh.bfj, np.float64)
        self._check(self.lib.sfgpu_set_fields(self._ctx, mesh.index, _ptr(efi), _ptr(efj), _ptr(bfi), _ptr(bfj)))

    def addParticles(self, mesh, parts: Particles, dt=None, rewind=True, flags=0):
        """Bulk ``addParticle(MeshData, Particle)`` (KM:759-802) in array order."""
        if rewind:
            flags |= _lib.INJECT_REWIND
        dt = self.dt if dt is None else dt
        added = C.c_int64(0)
        v = parts.view()
        self._check(self.lib.sfgpu_inject(self._ctx, self._sp, mesh.index, C.byref(v), float(dt), int(flags), C.byref(added)))
        return added.value

    def sampleUniformSource(self, spline, v_drift, num_mp, rng_state, dt=None, mpw=None, born_it=0, cold_beam=False):
        """``Source.sampleKinetic`` over ``UniformSource.sampleParticle`` (Source.java:167-198, UniformSource.java:56-72) sampled
        on the device with the draws of ``java.util.Random``; ``cold_beam``: ``ColdBeamSource.sampleParticle``.  ``spline``: a ``domain.LinearSpline``;
        ``rng_state``: the 48-bit internal state of the Random behind ``Starfish.rnd()``.  Returns (added, new state)."""
        dt = self.dt if dt is None else dt
        ptr = lambda a: a.ctypes.data_as(_lib.c_double_p)
        sp = _lib.Spline(int(spline.n_seg), *[ptr(getattr(spline, k)) for k in ("x1", "y1", "x2", "y2", "nx", "ny", "area", "cum_area")],
                         float(spline.spline_area))
        st, added = C.c_uint64(int(rng_state)), C.c_int64(0)
        self._check(self.lib.sfgpu_source_uniform(self._ctx, self._sp, C.byref(sp), _lib.SOURCE_COLD_BEAM if cold_beam else 0, float(v_drift), float(self.spwt0 if mpw is None else mpw),
                                                  int(born_it), int(num_mp), float(dt), C.byref(st), C.byref(added)))
        return added.value, int(st.value)

    def addParticle(self, pos, vel, mesh=None, mpw=None):
        """``addParticle(double[] pos, double[] vel)`` (KM:826-835): one particle of weight spwt0."""
        if mesh is None:  # DomainModule.getMesh(pos), DomainModule.java:106-117
            mesh = next((m for m in self.meshes if m.containsPos(np.asarray(pos[:2]))), None)
            if mesh is None:
                return False
        one = lambda v: np.array([v], dtype=np.float64)
        p = Particles(1, x=one(pos[0]), y=one(pos[1]), z=one(pos[2]), u=one(vel[0]), v=one(vel[1]), w=one(vel[2]),
                      mpw=one(self.spwt0 if mpw is None else mpw))
        return self.addParticles(mesh, p) == 1

    # ------------------------------------------------------------------ the hot path
    def updateFields(self, dt=None):
        """``KineticMaterial.updateFields()`` (KM:117-163) on the GPU."""
        dt = self.dt if dt is None else float(dt)
        defer = self.slow_path_handler is not None
        self._check(self.lib.sfgpu_step(self._ctx, self._sp, dt, self.step_flags | (_lib.STEP_DEFER_FINISH if defer else 0)))
        if defer:
            slow, extra = self.takeSlowPath()
            if slow.n:
                for mesh_index, survivors in self.slow_path_handler(self, slow, extra):
                    if survivors.n:
                        self.addParticles(self.meshes[mesh_index], survivors, rewind=False, flags=_lib.INJECT_DEPOSIT_NOW)
            self._check(self.lib.sfgpu_finish_step(self._ctx, self._sp))
        self._collect()

    def _collect(self):
        sums = (C.c_double * 5)()
        np_alive, n_exit, n_slow = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self.lib.sfgpu_get_sums(self._ctx, self._sp, sums, C.byref(np_alive), C.byref(n_exit), C.byref(n_slow)))
        # KM:252-258
        self.mass_sum = sums[0] * self.mass
        self.momentum_sum = np.array([sums[1], sums[2], sums[3]]) * self.mass
        self.energy_sum = sums[4] * self.mass
        self.n_exited, self.n_slow = n_exit.value, n_slow.value
        self._step_no += 1
        for k in range(len(self.meshes)):
            f = self.fields[k]
            # nd of updateFields(MeshData), KM:168-197: what the field solver reads every step; u,v,w follow on first access
            if self.download_fields:
                self._check(self.lib.sfgpu_get_moments(self._ctx, self._sp, k, C.c_void_p(dict.__getitem__(f, "nd").ctypes.data), None, None, None))

    def _fetch_velocities(self, k):
        if self._vel_step[k] != self._step_no:
            f = self.fields[k]
            self._check(self.lib.sfgpu_get_moments(self._ctx, self._sp, k, None, *[C.c_void_p(dict.__getitem__(f, q).ctypes.data) for q in ("u", "v", "w")]))
            self._vel_step[k] = self._step_no

    def _fetch_deposit(self, k):
        if self._dep_step[k] != self._step_no:
            dep = self._dep[k]
            ptrs = (C.c_void_p * NFIELDS)(*[dep[f].ctypes.data for f in range(NFIELDS)])
            self._check(self.lib.sfgpu_get_deposit(self._ctx, self._sp, k, ptrs))
            self._dep_step[k] = self._step_no
        return self._dep[k]

    def _fetch_samples(self, k):
        if self._samp_step[k] != self._step_no:
            f = self.fields[k]
            ptrs = (C.c_void_p * NFIELDS)(*[dict.__getitem__(f, name).ctypes.data for name in _SAMPLE_NAMES])
            self._check(self.lib.sfgpu_get_samples(self._ctx, self._sp, k, ptrs, None))
            self._samp_step[k] = self._step_no

    def clearSamples(self):  # KM:1509-1528
        self._check(self.lib.sfgpu_clear_samples(self._ctx, self._sp))
        self._samp_step = [-1] * len(self.meshes)

    # ------------------------------------------------------------------ outputs
    def getNp(self, mesh=None):  # KM:1297 / :1385
        n = C.c_int64()
        self._check(self.lib.sfgpu_np(self._ctx, self._sp, -1 if mesh is None else mesh.index, C.byref(n)))
        return n.value

    def getDen(self, mesh):
        return self.fields[mesh.index]["nd"]

    def getU(self, mesh):
        return self.fields[mesh.index]["u"]

    def getV(self, mesh):
        return self.fields[mesh.index]["v"]

    def getW(self, mesh):
        return self.fields[mesh.index]["w"]

    def getMassSum(self):
        return self.mass_sum

    def getMomentumSum(self):
        return self.momentum_sum

    def getEnergySum(self):
        return self.energy_sum

    def getParticles(self, mesh) -> Particles:
        """All particles of a mesh (what ``getIterator(mesh)`` walks, KM:271), in device order."""
        n = self.getNp(mesh)
        out = Particles.empty(n)
        v = out.view()
        self._check(self.lib.sfgpu_download(self._ctx, self._sp, mesh.index, 0, C.byref(v)))
        return out

    def setParticles(self, mesh, parts: Particles, first=0):
        """Write particles back after a host-side mutation (MCC/DSMC/chemistry edit Particle fields in place)."""
        v = parts.view()
        self._check(self.lib.sfgpu_upload(self._ctx, self._sp, mesh.index, int(first), C.byref(v)))

    def saveRestartParticles(self, mesh) -> bytes:
        """Particle section of saveRestartData for one mesh (KM:907-924): the exact DataOutputStream bytes."""
        need = C.c_int64()
        self._check(self.lib.sfgpu_restart_save(self._ctx, self._sp, mesh.index, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        self._check(self.lib.sfgpu_restart_save(self._ctx, self._sp, mesh.index, buf, need.value, None))
        return buf.raw

    def loadRestartParticles(self, mesh, data: bytes, dt=None):
        """loadRestartData's particle loop (KM:961-979) from the stream bytes; returns (bytes consumed, particles added)."""
        used, added = C.c_int64(), C.c_int64()
        dt = self.dt if dt is None else dt
        self._check(self.lib.sfgpu_restart_load(self._ctx, self._sp, mesh.index, data, len(data), float(dt), C.byref(used), C.byref(added)))
        return used.value, added.value

    def cellLists(self, mesh):
        """Per-cell particle lists (what DSMC.java:194-252 / KM:1150-1179 build on the host): after this call
        ``getParticles(mesh)`` is in cell order and particles ``first[i, j] .. first[i, j] + count[i, j]`` are the ones of cell (i, j).
        Returns (first, count, n_sorted); the exceptional records follow at [n_sorted, np)."""
        first = np.zeros((mesh.ni - 1, mesh.nj - 1), np.int64)
        count = np.zeros((mesh.ni - 1, mesh.nj - 1), np.int32)
        ns = C.c_int64()
        self._check(self.lib.sfgpu_cell_lists(self._ctx, self._sp, mesh.index, first.ctypes.data_as(_lib.c_int64_p), count.ctypes.data_as(_lib.c_int32_p), C.byref(ns)))
        return first, count, ns.value

    def takeSurfaceHits(self):
        """Surface hits of the last step (KM:586-602): dict of arrays mesh, seg, t, u, v, w, mpw, alive; also sets n_absorbed."""
        n, na = C.c_int64(), C.c_int64()
        self._check(self.lib.sfgpu_take_surface_hits(self._ctx, self._sp, 0, None, None, None, None, None, None, None, None, C.byref(n), C.byref(na)))
        self.n_absorbed = na.value
        k = n.value
        out = dict(mesh=np.empty(k, np.int32), seg=np.empty(k, np.int32), t=np.empty(k), u=np.empty(k), v=np.empty(k), w=np.empty(k), mpw=np.empty(k),
                   alive=np.empty(k, np.int8))
        if k:
            d = lambda a: a.ctypes.data_as(_lib.c_double_p)
            self._check(self.lib.sfgpu_take_surface_hits(self._ctx, self._sp, k, out["mesh"].ctypes.data_as(_lib.c_int32_p), out["seg"].ctypes.data_as(_lib.c_int32_p),
                                                         d(out["t"]), d(out["u"]), d(out["v"]), d(out["w"]), d(out["mpw"]), out["alive"].ctypes.data_as(C.POINTER(C.c_int8)),
                                                         C.byref(n), C.byref(na)))
        return out

    def takeSlowPath(self):
        n_slow = C.c_int64()
        self._check(self.lib.sfgpu_get_sums(self._ctx, self._sp, None, None, None, C.byref(n_slow)))
        n = n_slow.value
        out = Particles.empty(n)
        extra = {k: np.empty(n) for k in ("old_x", "old_y", "old_li", "old_lj")}
        extra.update(bounces=np.empty(n, np.int32), mesh=np.empty(n, np.int32))
        ex = _CSlowExtra()
        for k in ("old_x", "old_y", "old_li", "old_lj"):
            setattr(ex, k, extra[k].ctypes.data_as(_lib.c_double_p))
        ex.bounces = extra["bounces"].ctypes.data_as(_lib.c_int32_p)
        ex.mesh = extra["mesh"].ctypes.data_as(_lib.c_int32_p)
        got = C.c_int64()
        v = out.view()
        self._check(self.lib.sfgpu_take_slowpath(self._ctx, self._sp, n, C.byref(v), C.byref(ex), C.byref(got)))
        assert got.value == n
        return out, extra

    # ------------------------------------------------------------------ multi GPU + measurement
    @staticmethod
    def commUniqueId() -> bytes:
        lib = _lib.load()
        buf = C.create_string_buffer(128)
        rc = lib.sfgpu_comm_unique_id(buf)
        if rc:
            raise SfgpuError(rc, lib.sfgpu_last_error(None).decode())
        return buf.raw

    def commInit(self, nranks, rank, unique_id: bytes):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(self.lib.sfgpu_comm_init(self._ctx, int(nranks), int(rank), buf))

    def step_raw(self, dt, flags=None):
        """sfgpu_step only (no result download): the device-resident hot path, used by bench.py."""
        self._check(self.lib.sfgpu_step(self._ctx, self._sp, float(dt), self.step_flags if flags is None else int(flags)))

    def lastStepTiming(self):
        tot, ker, n = C.c_float(), C.c_float(), C.c_int32()
        self._check(self.lib.sfgpu_last_step_timing(self._ctx, C.byref(tot), C.byref(ker), C.byref(n)))
        return tot.value, ker.value, n.value

    def stepStats(self):
        """(particles alive, particles that left through open faces in the last step, step-kernel ms, step kernel kind, out-of-tile
        deposits) of the last step with as few library calls as possible (bench.py calls this inside its timed loop)."""
        a, b = C.c_int64(), C.c_int64()
        self._check(self.lib.sfgpu_get_sums(self._ctx, self._sp, None, C.byref(a), C.byref(b), None))
        ker, k, f = C.c_float(), C.c_int32(), C.c_int64()
        self._check(self.lib.sfgpu_last_step_timing(self._ctx, None, C.byref(ker), None))
        self._check(self.lib.sfgpu_last_step_kernel(self._ctx, C.byref(k)))
        self._check(self.lib.sfgpu_last_step_counters(self._ctx, C.byref(f)))
        return a.value, b.value, ker.value, k.value, f.value

    def lastStepKernel(self):
        """Which step kernel ran last: 0 tiled in-place, 1 streaming (moves, deposits and re-sorts), 2 generic."""
        k = C.c_int32()
        self._check(self.lib.sfgpu_last_step_kernel(self._ctx, C.byref(k)))
        return k.value

    def sync(self):
        self._check(self.lib.sfgpu_sync(self._ctx))

    def n_exited_last(self):
        """Particles that left through OPEN faces in the last step (they were pushed in it)."""
        n = C.c_int64()
        self._check(self.lib.sfgpu_get_sums(self._ctx, self._sp, None, None, C.byref(n), None))
        return n.value

    def lastStepFallback(self):
        n = C.c_int64()
        self._check(self.lib.sfgpu_last_step_counters(self._ctx, C.byref(n)))
        return n.value

    def sort(self):
        """Explicit cell sort + compaction of the device store (sfgpu_sort)."""
        self._check(self.lib.sfgpu_sort(self._ctx, self._sp))

    def setSortInterval(self, steps):
        self._check(self.lib.sfgpu_set_sort_interval(self._ctx, int(steps)))

    def setTileHalo(self, halo):
        """0: automatic (narrow tile until deposits miss it), 1 / 2: fixed halo of the tiled kernel's accumulation tile (sfgpu_set_tile_halo)."""
        self._check(self.lib.sfgpu_set_tile_halo(self._ctx, int(halo)))

    def tileHalo(self):
        """(halo the next tiled step runs with, automatic?)"""
        h, a = C.c_int32(), C.c_int32()
        self._check(self.lib.sfgpu_get_tile_halo(self._ctx, C.byref(h), C.byref(a)))
        return h.value, bool(a.value)

    def timerStart(self):
        self._check(self.lib.sfgpu_timer_start(self._ctx))

    def timerStop(self):
        """Device milliseconds since timerStart(), CUDA events on the stream the kernels run on."""
        ms = C.c_float()
        self._check(self.lib.sfgpu_timer_stop(self._ctx, C.byref(ms)))
        return ms.value

    def launchCount(self):
        n = C.c_int64()
        self._check(self.lib.sfgpu_launch_count(self._ctx, C.byref(n)))
        return n.value
