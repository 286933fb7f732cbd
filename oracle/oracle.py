"""ctypes driver of the CPU ORACLE (oracle/sf_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module; the product package (starfish_b200) never does.  PARITY UNPINNED by the reference (it ships no
tests or golden vectors and no JVM exists in the image): the oracle is pinned by the analytic
known-answer tests in tests/test_oracle_kat.py and by the independent pure-Python restatement in
tests/pyref.py.

OracleKM reproduces the driver logic of KineticMaterial.updateFields() (KineticMaterial.java:117-163)
around the per-particle C functions: moveParticles(false), the <=10 transfer sweeps (KM:131-142),
updateFields(MeshData) (KM:168-197) and updateSamples (KM:1570-1595).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsf_oracle.so")

ALIVE, REMOVED, DEAD, SLOW, TRANSFER, ABSORBED = 0, 1, 2, 3, 4, 5
_PK = ("x", "y", "z", "u", "v", "w", "mpw", "li", "lj", "dt")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class _Segment(C.Structure):
    _fields_ = [("x1", C.c_double), ("y1", C.c_double), ("x2", C.c_double), ("y2", C.c_double), ("kind", C.c_int32), ("sink", C.c_int32)]


class _Hits(C.Structure):
    _fields_ = [("cap", C.c_int64), ("n", C.c_int64), ("seg", _ip), ("t", _dp), ("u", _dp), ("v", _dp), ("w", _dp), ("mpw", _dp),
                ("alive", C.POINTER(C.c_int8))]


class _Mesh(C.Structure):
    _fields_ = [("ni", C.c_int32), ("nj", C.c_int32), ("x0", C.c_double * 2), ("dh", C.c_double * 2),
                ("domain_type", C.c_int32), ("bc", C.c_void_p * 4), ("nbr", C.c_void_p * 4), ("has_seg", C.c_void_p),
                ("seg_offs", C.c_void_p), ("seg_ids", C.c_void_p), ("segs", C.c_void_p), ("hits", C.c_void_p),
                ("efi", C.c_void_p), ("efj", C.c_void_p), ("bfi", C.c_void_p), ("bfj", C.c_void_p)]


class _Parts(C.Structure):
    _fields_ = [("n", C.c_int64)] + [(k, _dp) for k in _PK]


class _Spline(C.Structure):
    _fields_ = [("n_seg", C.c_int32)] + [(k, _dp) for k in ("x1", "y1", "x2", "y2", "nx", "ny", "area", "cum_area")] + [("spline_area", C.c_double)]


class _MoveOut(C.Structure):
    _fields_ = [("status", C.POINTER(C.c_int8)), ("xfer_mask", _ip), ("xfer_mesh", _ip), ("xfer_li", _dp), ("xfer_lj", _dp),
                ("old_x", _dp), ("old_y", _dp), ("old_li", _dp), ("old_lj", _dp), ("bounces", _ip)]


_lib = None


def build():
    """Compile the oracle with gcc (oracle/Makefile).  Building the checker is not using it."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    lib.sfo_gather.restype = C.c_double
    lib.sfo_gather.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
    lib.sfo_gather_safe.restype = C.c_double
    lib.sfo_gather_safe.argtypes = lib.sfo_gather.argtypes
    lib.sfo_scatter.restype = None
    lib.sfo_scatter.argtypes = [C.c_void_p, C.POINTER(_Mesh), C.c_double, C.c_double, C.c_double]
    lib.sfo_xtol.restype = None
    lib.sfo_xtol.argtypes = [C.POINTER(_Mesh), C.c_double, C.c_double, _dp, _dp]
    lib.sfo_contains_pos.restype = C.c_int
    lib.sfo_contains_pos.argtypes = [C.POINTER(_Mesh), C.c_double, C.c_double]
    lib.sfo_boris.restype = None
    lib.sfo_boris.argtypes = [C.c_double, C.c_double, _dp, _dp, _dp]
    lib.sfo_mirror.restype = None
    lib.sfo_mirror.argtypes = [_dp, _dp]
    lib.sfo_move.restype = None
    lib.sfo_move.argtypes = [C.POINTER(_Mesh), C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.POINTER(_Parts),
                             C.c_int64, C.c_int64, C.POINTER(_MoveOut), _dp]
    lib.sfo_move_mt.restype = None
    lib.sfo_move_mt.argtypes = [C.POINTER(_Mesh), C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.POINTER(_Parts),
                                C.POINTER(_MoveOut), _dp, C.c_int]
    lib.sfo_deposit.restype = None
    lib.sfo_deposit.argtypes = [C.POINTER(_Mesh), C.POINTER(_Parts), _dp, _dp, _dp, _dp]
    lib.sfo_divide_by_field.restype = None
    lib.sfo_divide_by_field.argtypes = [_dp, _dp, C.c_int64]
    lib.sfo_scale_by_vol.restype = None
    lib.sfo_scale_by_vol.argtypes = [_dp, _dp, C.c_int64]
    lib.sfo_sample.restype = None
    lib.sfo_sample.argtypes = [C.POINTER(_Mesh), C.POINTER(_Parts)] + [_dp] * 8
    lib.sfo_add_particles.restype = None
    lib.sfo_add_particles.argtypes = [C.POINTER(_Mesh), C.c_double, C.c_double, C.c_int, C.POINTER(_Parts), C.c_int64, C.c_int64]
    lib.sfo_java_seed.restype = C.c_uint64
    lib.sfo_java_seed.argtypes = [C.c_int64]
    lib.sfo_java_next_int.restype = C.c_int32
    lib.sfo_java_next_int.argtypes = [C.POINTER(C.c_uint64)]
    lib.sfo_java_next_double.restype = C.c_double
    lib.sfo_java_next_double.argtypes = [C.POINTER(C.c_uint64)]
    lib.sfo_uniform_source.restype = None
    lib.sfo_uniform_source.argtypes = [C.POINTER(_Spline), C.c_int, C.c_double, C.c_double, C.c_int64, C.POINTER(C.c_uint64), C.POINTER(_Mesh), C.c_int,
                                       _dp, _dp, _dp, _dp, _dp, _dp, _ip]
    _lib = lib
    return lib


def java_seed(seed):
    """java.util.Random scrambled internal state for ``new Random(seed)``."""
    return int(load().sfo_java_seed(int(seed)))


def _d(a):
    return a.ctypes.data_as(_dp)


class MeshSet:
    """C view of a list of mesh objects (duck typed: ni, nj, x0, dh, domain_type, bc[4], nbr[4], has_seg,
    efi, efj, bfi, bfj -- starfish_b200.domain.UniformMesh fits)."""

    def __init__(self, meshes, use_segments=True, hit_cap=1 << 20):
        self.meshes = list(meshes)
        self.keep = []
        self.use_segments, self.hit_cap = use_segments, hit_cap
        self.hits = {}
        self.arr = (_Mesh * len(self.meshes))()
        self.refresh()

    def take_hits(self, k):
        """Surface hits recorded for mesh k since the last call (KM:586-602), as arrays."""
        if k not in self.hits:
            return None
        h, hb = self.hits[k]
        n = int(min(h.n, h.cap))
        out = {key: v[:n].copy() for key, v in hb.items()}
        h.n = 0
        return out

    def refresh(self):
        self.keep = []
        self.hits = {}
        for k, m in enumerate(self.meshes):
            c = self.arr[k]
            c.ni, c.nj = int(m.ni), int(m.nj)
            c.x0[0], c.x0[1] = float(m.x0[0]), float(m.x0[1])
            c.dh[0], c.dh[1] = float(m.dh[0]), float(m.dh[1])
            c.domain_type = int(m.domain_type)
            for f in range(4):
                b = np.ascontiguousarray(m.bc[f], np.int8)
                n = np.ascontiguousarray(m.nbr[f], np.int32)
                self.keep += [b, n]
                c.bc[f] = b.ctypes.data
                c.nbr[f] = n.ctypes.data
            hs = np.ascontiguousarray(m.has_seg, np.uint8)
            self.keep.append(hs)
            c.has_seg = hs.ctypes.data if hs.any() else None
            # segment tables (starfish_b200.domain.set_boundaries): with them the oracle runs the segment part of
            # ProcessBoundary itself (KM:504-603), without them such particles come back as SLOW
            c.seg_offs = c.seg_ids = c.segs = c.hits = None
            sg = getattr(m, "segments", None)
            if sg is not None and self.use_segments and hs.any():
                arr = (_Segment * len(sg["x1"]))()
                for q in range(len(sg["x1"])):
                    arr[q] = _Segment(float(sg["x1"][q]), float(sg["y1"][q]), float(sg["x2"][q]), float(sg["y2"][q]), int(sg["kind"][q]), int(sg["sink"][q]))
                offs, ids = np.ascontiguousarray(m.seg_offs, np.int32), np.ascontiguousarray(m.seg_ids, np.int32)
                cap = self.hit_cap
                hb = dict(seg=np.zeros(cap, np.int32), t=np.zeros(cap), u=np.zeros(cap), v=np.zeros(cap), w=np.zeros(cap), mpw=np.zeros(cap),
                          alive=np.zeros(cap, np.int8))
                h = _Hits(cap, 0, hb["seg"].ctypes.data_as(_ip), _d(hb["t"]), _d(hb["u"]), _d(hb["v"]), _d(hb["w"]), _d(hb["mpw"]),
                          hb["alive"].ctypes.data_as(C.POINTER(C.c_int8)))
                self.keep += [arr, offs, ids, hb, h]
                self.hits[k] = (h, hb)
                c.seg_offs, c.seg_ids, c.segs, c.hits = offs.ctypes.data, ids.ctypes.data, C.addressof(arr), C.addressof(h)
            for name in ("efi", "efj", "bfi", "bfj"):
                a = getattr(m, name, None)
                if a is None:
                    setattr(c, name, None)
                else:
                    a = np.ascontiguousarray(a, np.float64)
                    self.keep.append(a)
                    setattr(c, name, a.ctypes.data)


def empty_parts(n=0):
    p = {k: np.zeros(n) for k in _PK}
    p["id"] = np.zeros(n, np.int32)
    p["born_it"] = np.zeros(n, np.int32)
    return p


def _cparts(p):
    c = _Parts()
    c.n = len(p["x"])
    for k in _PK:
        assert p[k].dtype == np.float64 and p[k].flags.c_contiguous and len(p[k]) == c.n, k
        setattr(c, k, _d(p[k]))
    return c


def _concat(a, b):
    return {k: np.concatenate([a[k], b[k]]) for k in a}


def _take(p, sel):
    return {k: np.ascontiguousarray(v[sel]) for k, v in p.items()}


def move(ms: MeshSet, mesh_id, qm, charge, dt, transfer, p, threads=1):
    """sfo_move over all particles of `p` in place; returns (status, aux dict, sums5)."""
    lib = load()
    n = len(p["x"])
    out = _MoveOut()
    aux = dict(status=np.zeros(n, np.int8), xfer_mask=np.zeros(n, np.int32), xfer_mesh=np.full(2 * n, -1, np.int32),
               xfer_li=np.zeros(2 * n), xfer_lj=np.zeros(2 * n), old_x=np.zeros(n), old_y=np.zeros(n), old_li=np.zeros(n),
               old_lj=np.zeros(n), bounces=np.zeros(n, np.int32))
    out.status = aux["status"].ctypes.data_as(C.POINTER(C.c_int8))
    for k in ("xfer_mask", "xfer_mesh", "bounces"):
        setattr(out, k, aux[k].ctypes.data_as(_ip))
    for k in ("xfer_li", "xfer_lj", "old_x", "old_y", "old_li", "old_lj"):
        setattr(out, k, _d(aux[k]))
    sums = np.zeros(5)
    cp = _cparts(p)
    if threads > 1:
        lib.sfo_move_mt(ms.arr, mesh_id, qm, charge, dt, int(transfer), C.byref(cp), C.byref(out), _d(sums), threads)
    else:
        lib.sfo_move(ms.arr, mesh_id, qm, charge, dt, int(transfer), C.byref(cp), 0, n, C.byref(out), _d(sums))
    return aux["status"], aux, sums


class OracleKM:
    """KineticMaterial on the CPU oracle: same observable state as the Java class after updateFields()."""

    FIELDS = ("count-sum", "u-sum", "v-sum", "w-sum", "uu-sum", "vv-sum", "ww-sum", "mpc-sum")

    def __init__(self, charge, mass, meshes, threads=1, use_segments=True):
        self.lib = load()
        self.charge, self.mass = float(charge), float(mass)
        self.qm = self.charge / self.mass  # Material.java:711
        self.meshes = list(meshes)
        self.ms = MeshSet(self.meshes, use_segments=use_segments)
        self.n_absorbed = 0
        self.hits = []  # per step: surface hits per mesh (dict of arrays) in the order the mover met them
        self.threads = threads
        self.parts = [empty_parts() for _ in self.meshes]
        self.transfer = [empty_parts() for _ in self.meshes]
        self.id_counter = 0
        z = lambda m: np.zeros((m.ni, m.nj))
        self.fields = [{k: z(m) for k in ("nd", "u", "v", "w") + self.FIELDS} for m in self.meshes]
        self.raw = [None] * len(self.meshes)  # raw per-step deposit [8][ni][nj], order of SFGPU_F_*
        self.num_samples = 0
        self.mass_sum, self.momentum_sum, self.energy_sum = 0.0, np.zeros(3), 0.0
        self.sums5 = np.zeros(5)
        self.n_exited = 0
        self.n_removed = 0
        self.slow = []  # (mesh_id, parts, aux) handed to the slow path

    def refresh_fields(self):
        self.ms.refresh()

    # KM:759-802 + MeshData.addParticle KM:1356-1361
    def sampleUniformSource(self, spline, v_drift, num_mp, dt, rng_state, mpw, born_it=0, cold_beam=False):
        """Source.sampleKinetic over UniformSource.sampleParticle: returns (particles added, new RNG state).
        Particles keep the order they were sampled in; ids count up over the accepted ones (KM:797)."""
        n = int(num_mp)
        x, y, z, u, v, w = (np.zeros(n) for _ in range(6))
        mesh_of = np.zeros(n, np.int32)
        sp = _Spline(spline.n_seg, *[_d(getattr(spline, k)) for k in ("x1", "y1", "x2", "y2", "nx", "ny", "area", "cum_area")], spline.spline_area)
        st = C.c_uint64(int(rng_state))
        self.lib.sfo_uniform_source(C.byref(sp), int(bool(cold_beam)), float(v_drift), float(dt), n, C.byref(st), self.ms.arr, len(self.meshes),
                                    _d(x), _d(y), _d(z), _d(u), _d(v), _d(w), mesh_of.ctypes.data_as(_ip))
        acc = mesh_of >= 0
        ids = (self.id_counter + np.cumsum(acc) - 1).astype(np.int32)
        added = 0
        for k in range(len(self.meshes)):
            sel = mesh_of == k
            if not sel.any():
                continue
            arr = dict(x=x[sel], y=y[sel], z=z[sel], u=u[sel], v=v[sel], w=w[sel], mpw=np.full(int(sel.sum()), float(mpw)))
            self.addParticles(k, arr, dt, ids=ids[sel], born_it=np.full(int(sel.sum()), int(born_it), np.int32))
            added += int(sel.sum())
        self.id_counter += int(acc.sum())
        return added, int(st.value)

    def addParticles(self, mesh_id, arrays, dt, rewind=True, ids=None, born_it=None, transfer=False):
        n = len(arrays["x"])
        p = empty_parts(n)
        for k in _PK:
            if arrays.get(k) is not None:
                p[k][:] = arrays[k]
        compute_lc = arrays.get("li") is None
        cp = _cparts(p)
        if transfer:
            assert not rewind
        if rewind:
            self.lib.sfo_add_particles(self.ms.arr[mesh_id], self.qm, dt, int(compute_lc), C.byref(cp), 0, n)
        elif compute_lc:
            m = self.meshes[mesh_id]
            p["li"] = (p["x"] - m.x0[0]) / m.dh[0]
            p["lj"] = (p["y"] - m.x0[1]) / m.dh[1]
            p["li"][p["li"] >= m.ni] = m.ni - 1
            p["lj"][p["lj"] >= m.nj] = m.nj - 1
        if ids is None:
            p["id"] = (self.id_counter + np.arange(n)).astype(np.int32)
            self.id_counter += n
        else:
            p["id"] = np.asarray(ids, np.int32).copy()
        if born_it is not None:
            p["born_it"] = np.asarray(born_it, np.int32).copy()
        ok = np.isfinite(p["u"]) & np.isfinite(p["v"]) & np.isfinite(p["w"])
        p = _take(p, ok)
        if transfer:
            self.transfer[mesh_id] = _concat(self.transfer[mesh_id], p)
        else:
            self.parts[mesh_id] = _concat(self.parts[mesh_id], p)
        return int(ok.sum())

    def _hand_off(self, p, st, aux):
        sel = np.nonzero(st == TRANSFER)[0]
        for q in sel:
            for k in range(2):
                if aux["xfer_mask"][q] & (1 << k):
                    nb = int(aux["xfer_mesh"][2 * q + k])
                    one = _take(p, [q])
                    one["li"][0] = aux["xfer_li"][2 * q + k]
                    one["lj"][0] = aux["xfer_lj"][2 * q + k]
                    self.transfer[nb] = _concat(self.transfer[nb], one)

    def _slow(self, mesh_id, p, st, aux):
        sel = st == SLOW
        if sel.any():
            self.slow.append((mesh_id, _take(p, sel), {k: v[sel] for k, v in aux.items() if len(v) == len(sel)}))

    def updateFields(self, dt):
        self.slow = []
        self.n_exited = 0
        self.n_removed = 0
        self.n_absorbed = 0
        sums = np.zeros(5)
        # moveParticles(false), KM:126
        for k in range(len(self.meshes)):
            p = self.parts[k]
            st, aux, s5 = move(self.ms, k, self.qm, self.charge, dt, False, p, self.threads)
            sums += s5
            self._hand_off(p, st, aux)
            self._slow(k, p, st, aux)
            self.n_exited += int((st == DEAD).sum())
            self.n_removed += int((st == REMOVED).sum())
            self.n_absorbed += int((st == ABSORBED).sum())
            self.parts[k] = _take(p, st == ALIVE)
        self.sums5 = sums
        self.mass_sum = sums[0] * self.mass  # KM:252-258
        self.momentum_sum = sums[1:4] * self.mass
        self.energy_sum = sums[4] * self.mass
        # transfer sweeps, KM:131-142
        for _loop in range(10):
            for k in range(len(self.meshes)):
                tp = self.transfer[k]
                if len(tp["x"]) == 0:
                    continue
                self.transfer[k] = empty_parts()
                st, aux, _ = move(self.ms, k, self.qm, self.charge, dt, True, tp, 1)
                self._hand_off(tp, st, aux)
                self._slow(k, tp, st, aux)
                self.n_exited += int((st == DEAD).sum())
                self.n_removed += int((st == REMOVED).sum())
                self.n_absorbed += int((st == ABSORBED).sum())
                ok = (st == ALIVE) & np.isfinite(tp["u"]) & np.isfinite(tp["v"]) & np.isfinite(tp["w"])
                self.parts[k] = _concat(self.parts[k], _take(tp, ok))
            if sum(len(t["x"]) for t in self.transfer) == 0:
                break
        self.hits = [self.ms.take_hits(k) for k in range(len(self.meshes))]
        self.deposit()

    def deposit(self):
        """updateFields(MeshData) KM:168-197 + updateSamples KM:1570-1595 for every mesh."""
        for k, m in enumerate(self.meshes):
            p = self.parts[k]
            cp = _cparts(p)
            raw = np.zeros((8, m.ni, m.nj))
            self.lib.sfo_deposit(self.ms.arr[k], C.byref(cp), _d(raw[0]), _d(raw[1]), _d(raw[2]), _d(raw[3]))
            # second pass into fresh arrays: u,v,w,uu,vv,ww,count,mpc
            s = np.zeros((8, m.ni, m.nj))
            self.lib.sfo_sample(self.ms.arr[k], C.byref(cp), _d(s[0]), _d(s[1]), _d(s[2]), _d(s[3]), _d(s[4]), _d(s[5]),
                                _d(s[6]), _d(s[7]))
            raw[4], raw[5], raw[6], raw[7] = s[4], s[5], s[6], s[7]
            self.raw[k] = raw
            self.sample_pass = s
            f = self.fields[k]
            f["count-sum"] += s[0]
            f["u-sum"] += s[1]
            f["v-sum"] += s[2]
            f["w-sum"] += s[3]
            f["uu-sum"] += s[4]
            f["vv-sum"] += s[5]
            f["ww-sum"] += s[6]
            f["mpc-sum"] += s[7]
            nd, u, v, w = raw[0].copy(), raw[1].copy(), raw[2].copy(), raw[3].copy()
            nn = nd.size
            for a in (u, v, w):
                self.lib.sfo_divide_by_field(_d(a), _d(nd), nn)
            vol = np.ascontiguousarray(m.node_vol, np.float64)
            self.lib.sfo_scale_by_vol(_d(nd), _d(vol), nn)
            f["nd"], f["u"], f["v"], f["w"] = nd, u, v, w
        self.num_samples += 1

    def getNp(self, mesh_id=None):
        if mesh_id is None:
            return sum(len(p["x"]) for p in self.parts)
        return len(self.parts[mesh_id]["x"])

    def sorted_parts(self, mesh_id):
        p = self.parts[mesh_id]
        o = np.argsort(p["id"], kind="stable")
        return _take(p, o)
