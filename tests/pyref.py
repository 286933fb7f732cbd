"""Independent pure-Python restatement of the Java hot path (test infrastructure).

Written from the Java sources separately from oracle/sf_oracle.c and structured like the Java (Particle
objects with pos[3]/vel[3]/lc[2], Field2D-like gather/scatter on nested lists), so that a transcription slip
in one restatement shows up as a disagreement between the two.  Python floats are IEEE doubles and CPython
never fuses a*b+c, which is exactly Java's arithmetic.  Small N only.

Cited Java (src/starfish/core/): materials/KineticMaterial.java (KM), domain/Field2D.java (F2D),
domain/UniformMesh.java (UM), domain/Mesh.java (MESH), common/Vec.java.
"""
import math

RIGHT, TOP, LEFT, BOTTOM = 0, 1, 2, 3  # MESH:107-118
OPEN, DIRICHLET, NEUMANN, PERIODIC, SYMMETRY, MESH_BC, SINK, CIRCUIT = -1, 0, 1, 2, 3, 4, 5, 6  # MESH:140-155
XY, RZ, ZR = 0, 1, 2
FLT_EPS = 1e-7


def jint(d):
    """Java (int)double."""
    if d != d:
        return 0
    if d >= 2147483647.0:
        return 2147483647
    if d <= -2147483648.0:
        return -2147483648
    return int(d)


class IndexOutOfBounds(Exception):
    pass


class Field:
    """Field2D: data[i][j] (F2D:44)."""

    def __init__(self, mesh, data=None):
        self.mesh = mesh
        self.ni, self.nj = mesh.ni, mesh.nj
        self.data = [[0.0] * self.nj for _ in range(self.ni)] if data is None else [[float(v) for v in row] for row in data]

    def _at(self, i, j):
        if i < 0 or j < 0 or i >= self.ni or j >= self.nj:
            raise IndexOutOfBounds()
        return self.data[i][j]

    def gather(self, lc):  # F2D:300-350
        try:
            fi, fj = lc
            i, j = jint(fi), jint(fj)
            di, dj = fi - i, fj - j
            v = (1 - di) * (1 - dj) * self._at(i, j)
            v += di * (1 - dj) * self._at(i + 1, j)
            v += di * dj * self._at(i + 1, j + 1)
            v += (1 - di) * dj * self._at(i, j + 1)
            return v
        except IndexOutOfBounds:
            return self.gather_safe(lc)

    def gather_safe(self, lc):  # F2D:371-390
        fi, fj = lc
        i, j = jint(fi), jint(fj)
        di, dj = fi - i, fj - j
        if i < 0:
            i, di = 0, 0.0
        if j < 0:
            j, dj = 0, 0.0
        if i >= self.ni - 1:
            i, di = self.ni - 1, 0.0
        if j >= self.nj - 1:
            j, dj = self.nj - 1, 0.0
        v = (1 - di) * (1 - dj) * self.data[i][j]
        if di > 0:
            v += di * (1 - dj) * self.data[i + 1][j]
        if di > 0 and dj > 0:
            v += di * dj * self.data[i + 1][j + 1]
        if dj > 0:
            v += (1 - di) * dj * self.data[i][j + 1]
        return v

    def scatter(self, lc, val):  # F2D:244-295
        fi, fj = lc
        i, j = jint(fi), jint(fj)
        di, dj = fi - i, fj - j
        if i < 0 or j < 0 or i >= self.ni - 1 or j >= self.nj - 1:
            return
        m = self.mesh
        if m.domain_type == RZ:
            rp, rm, r = m.R(i + 1, fj), m.R(i, fj), m.R(fi, fj)
            di = 1 - (0.5 * (rp - r) * (2 * rp + 3 * rm - r) / (rp * rp - rm * rm))
        elif m.domain_type == ZR:
            rp, rm, r = m.R(fi, j + 1), m.R(fi, j), m.R(fi, fj)
            di = fi - i
            dj = 1 - (0.5 * (rp - r) * (2 * rp + 3 * rm - r) / (rp * rp - rm * rm))
        self.data[i][j] += (1 - di) * (1 - dj) * val
        self.data[i + 1][j] += di * (1 - dj) * val
        self.data[i + 1][j + 1] += di * dj * val
        self.data[i][j + 1] += (1 - di) * dj * val


class Mesh:
    """UniformMesh (UM:33-43) + the Mesh members the path reads."""

    def __init__(self, ni, nj, x0, dh, domain_type=XY):
        self.ni, self.nj = ni, nj
        self.x0 = [float(x0[0]), float(x0[1])]
        self.dh = [float(dh[0]), float(dh[1])]
        self.xd = [self.x0[0] + (ni - 1) * self.dh[0], self.x0[1] + (nj - 1) * self.dh[1]]  # UM:131-135
        self.domain_type = domain_type
        self.bc = {RIGHT: [OPEN] * nj, LEFT: [OPEN] * nj, TOP: [OPEN] * ni, BOTTOM: [OPEN] * ni}
        self.nbr = {f: [[None, None] for _ in range(nj if f in (RIGHT, LEFT) else ni)] for f in (RIGHT, TOP, LEFT, BOTTOM)}
        self.has_seg = [[0] * nj for _ in range(ni)]
        self.Efi, self.Efj, self.Bfi, self.Bfj = Field(self), Field(self), Field(self), Field(self)

    def pos(self, lc):  # UM:139-145
        return [self.x0[0] + lc[0] * self.dh[0], self.x0[1] + lc[1] * self.dh[1]]

    def XtoL(self, x):  # UM:154-161
        return [(x[0] - self.x0[0]) / self.dh[0], (x[1] - self.x0[1]) / self.dh[1]]

    def R(self, i, j):  # MESH:824-831
        if self.domain_type == RZ:
            return self.x0[0] + i * self.dh[0]
        if self.domain_type == ZR:
            return self.x0[1] + j * self.dh[1]
        return 1.0

    def containsPos(self, x):  # MESH:1476-1483
        lc = self.XtoL(x)
        return not (lc[0] < -FLT_EPS or lc[1] < -FLT_EPS or lc[0] > (self.ni - 1 + FLT_EPS) or lc[1] > (self.nj - 1 + FLT_EPS))

    def faceNormal(self, face):  # UM:174-187
        return {LEFT: [1.0, 0.0, 0.0], RIGHT: [-1.0, 0.0, 0.0], BOTTOM: [0.0, 1.0, 0.0], TOP: [0.0, -1.0, 0.0]}[face]


class Particle:  # KM:1207-1282
    def __init__(self, pos, vel, mpw, pid=0):
        self.pos, self.vel = [float(v) for v in pos], [float(v) for v in vel]
        self.lc = None
        self.mpw, self.dt, self.id = float(mpw), 0.0, pid

    def copy(self):
        q = Particle(self.pos, self.vel, self.mpw, self.id)
        q.lc, q.dt = list(self.lc), self.dt
        return q


def mirror(vec, r):  # Vec.java:406-418
    t_mag = 0.0
    for k in range(3):
        t_mag += vec[k] * r[k]
    t = [r[k] * t_mag for k in range(3)]
    n = [vec[k] - t[k] for k in range(3)]
    t = [t[k] * -1 for k in range(3)]
    return [t[k] + n[k] for k in range(3)]


def cross3(a, b):  # Vec.java:318-325
    return [a[1] * b[2] - a[2] * b[1], -a[0] * b[2] + a[2] * b[0], a[0] * b[1] - a[1] * b[0]]


class KM:
    """KineticMaterial restricted to the hot path, serial (one ParticleMover per mesh)."""

    def __init__(self, charge, mass, meshes):
        self.charge, self.mass = charge, mass
        self.q_over_m = charge / mass
        self.meshes = meshes
        self.particles = [[] for _ in meshes]
        self.transfer = [[] for _ in meshes]
        self.slow = []
        self.id_counter = 0
        self.n_absorbed, self.hits = 0, []
        self.sums = [0.0] * 5
        self.n_exited = 0

    def boris(self, part, E, B):  # KM:847-893
        qm, dt = self.q_over_m, part.dt
        t = [qm * B[k] * 0.5 * dt for k in range(3)]
        t_mag2 = t[0] * t[0] + t[1] * t[1] + t[2] * t[2]
        s = [2 * t[k] / (1 + t_mag2) for k in range(3)]
        v_minus = [part.vel[k] + qm * E[k] * 0.5 * dt for k in range(3)]
        c = cross3(v_minus, t)
        v_prime = [v_minus[k] + c[k] for k in range(3)]
        c = cross3(v_prime, s)
        v_plus = [v_minus[k] + c[k] for k in range(3)]
        part.vel = [v_plus[k] + qm * E[k] * 0.5 * dt for k in range(3)]

    def kick(self, mesh, part):  # KM:336-353 / :782-794
        ef = [mesh.Efi.gather(part.lc), mesh.Efj.gather(part.lc), 0.0]
        bf = [mesh.Bfi.gather(part.lc), mesh.Bfj.gather(part.lc), 0.0]
        if bf[0] == 0 and bf[1] == 0:
            part.vel[0] += self.q_over_m * ef[0] * part.dt
            part.vel[1] += self.q_over_m * ef[1] * part.dt
        else:
            self.boris(part, ef, bf)

    def addParticle(self, m, part, dt):  # KM:759-802
        mesh = self.meshes[m]
        if part.lc is None:
            part.lc = mesh.XtoL(part.pos)
            if part.lc[0] >= mesh.ni:
                part.lc[0] = mesh.ni - 1
            if part.lc[1] >= mesh.nj:
                part.lc[1] = mesh.nj - 1
        part.dt = -0.5 * dt
        self.kick(mesh, part)
        part.dt = 0.0
        part.id = self.id_counter
        self.id_counter += 1
        if all(math.isfinite(v) for v in part.vel):  # KM:1357-1361
            self.particles[m].append(part)

    def bbox_segments(self, mesh, lc, lc_old):  # KM:482-518 reduced to "any segment node in the box"
        def mn(a, b):
            return float("nan") if (a != a or b != b) else min(a, b)

        def mx(a, b):
            return float("nan") if (a != a or b != b) else max(a, b)
        i_min, i_max = jint(mn(lc[0], lc_old[0])), jint(mx(lc[0], lc_old[0]))
        j_min, j_max = jint(mn(lc[1], lc_old[1])), jint(mx(lc[1], lc_old[1]))
        i_min, j_min = max(i_min, 0), max(j_min, 0)
        if i_max >= mesh.ni:
            i_max = mesh.ni - 1
        if j_max >= mesh.nj:
            j_max = mesh.nj - 1
        for i in range(i_min, i_max + 1):
            for j in range(j_min, j_max + 1):
                if mesh.has_seg[i][j]:
                    return True
        return False

    def segment_hit(self, mesh, part, old, lc_old, dt0):
        """Segment part of ProcessBoundary, KM:482-603, for LinearSegments whose surface outcome is deterministic.
        Returns None (no hit), True (hit, alive) or False (hit, removed)."""
        def mn(a, b):
            return float("nan") if (a != a or b != b) else min(a, b)

        def mx(a, b):
            return float("nan") if (a != a or b != b) else max(a, b)
        i_min, i_max = jint(mn(part.lc[0], lc_old[0])), jint(mx(part.lc[0], lc_old[0]))
        j_min, j_max = jint(mn(part.lc[1], lc_old[1])), jint(mx(part.lc[1], lc_old[1]))
        i_min, j_min = max(i_min, 0), max(j_min, 0)
        i_max, j_max = min(i_max, mesh.ni - 1), min(j_max, mesh.nj - 1)
        segments = []  # Set<Segment>, KM:505-518 (insertion order; the order only matters for exact ties)
        for i in range(i_min, i_max + 1):
            for j in range(j_min, j_max + 1):
                for sid in mesh.node_segments[i][j]:
                    if sid not in segments:
                        segments.append(sid)
        tp_min, tsurf_min, seg_min = 2.0, 0.0, None
        for sid in segments:
            seg = mesh.segments[sid]
            t = seg.intersect(old, part.pos)
            t_part = t[1]
            if t_part > 0:
                n = seg.normal
                acos = (n[0] * part.vel[0] + n[1] * part.vel[1]) / math.sqrt(part.vel[0] * part.vel[0] + part.vel[1] * part.vel[1]) \
                    if (part.vel[0] != 0 or part.vel[1] != 0) else float("nan")
                if t_part < FLT_EPS and acos > 0:
                    continue
                if t_part < tp_min:
                    tp_min, tsurf_min, seg_min = t_part, t[0], seg
        if seg_min is None:
            return None
        tp_min *= 0.9999
        part.pos[0] = old[0] + tp_min * (part.pos[0] - old[0])
        part.pos[1] = old[1] + tp_min * (part.pos[1] - old[1])
        part.lc = mesh.XtoL(part.pos)
        part.dt = dt0 * (1 - tp_min)
        if part.lc[0] < 0 and part.lc[0] > -FLT_EPS:
            part.lc[0] = 0.0
        if part.lc[1] < 0 and part.lc[1] > -FLT_EPS:
            part.lc[1] = 0.0
        alive = seg_min.kind != 0  # Material.performSurfaceInteraction (Material.java:279-300): no handler listed / ABSORB kill, NONE keeps
        if seg_min.kind == 2:  # SurfaceImpactSpecular without a species change, SurfaceInteraction.java:104-149: vel += n * (|vel_xy| * sqrt 2)
            n = seg_min.normal
            mag = math.sqrt(part.vel[0] * part.vel[0] + part.vel[1] * part.vel[1]) * math.sqrt(2)
            part.vel[0] += n[0] * mag
            part.vel[1] += n[1] * mag
        if seg_min.sink:
            alive = False
        self.hits.append((seg_min.sid, tsurf_min, list(part.vel), part.mpw, alive))
        return alive

    def process_boundary(self, m, part, old, lc_old):
        """KM:471-750.  Returns 'alive', 'dead', 'slow', 'absorbed' or 'transfer'."""
        mesh = self.meshes[m]
        near = self.bbox_segments(mesh, part.lc, lc_old)
        if near and not getattr(mesh, "segments", None):
            return "slow"
        dt0 = part.dt
        part.dt = 0.0
        if near:
            if self.segment_hit(mesh, part, old, lc_old, dt0) is False:
                self.n_absorbed += 1
                return "absorbed"
        ni, nj = mesh.ni, mesh.nj
        if part.lc[0] < 0 or part.lc[1] < 0 or part.lc[0] >= ni - 1 or part.lc[1] >= nj - 1:
            t_right = t_top = t_left = t_bottom = 99.0
            if part.lc[0] >= ni - 1:
                t_right = (ni - 1.0 - lc_old[0]) / (part.lc[0] - lc_old[0])
            if part.lc[1] >= nj - 1:
                t_top = (nj - 1.0 - lc_old[1]) / (part.lc[1] - lc_old[1])
            if part.lc[0] < 0:
                t_left = lc_old[0] / (lc_old[0] - part.lc[0])
            if part.lc[1] < 0:
                t_bottom = lc_old[1] / (lc_old[1] - part.lc[1])
            face, t = RIGHT, t_right
            if t_top < t:
                face, t = TOP, t_top
            if t_left < t:
                face, t = LEFT, t_left
            if t_bottom < t:
                face, t = BOTTOM, t_bottom
            part.lc[0] = lc_old[0] + t * (part.lc[0] - lc_old[0])
            part.lc[1] = lc_old[1] + t * (part.lc[1] - lc_old[1])
            if part.lc[0] < 0:
                part.lc[0] = 0.0
            elif part.lc[0] > ni - 1:
                part.lc[0] = float(ni - 1)
            if part.lc[1] < 0:
                part.lc[1] = 0.0
            elif part.lc[1] > nj - 1:
                part.lc[1] = float(nj - 1)
            x = mesh.pos(part.lc)
            part.pos[0], part.pos[1] = x[0], x[1]
            part.dt = dt0 * (1 - t)
            i, j = jint(part.lc[0]), jint(part.lc[1])
            if face == TOP:
                j += 1
            if face == RIGHT:
                i += 1
            i, j = max(i, 0), max(j, 0)
            if i >= ni - 1:
                i = ni - 1
            if j >= nj - 1:
                j = nj - 1
            typ = mesh.bc[face][j] if face in (LEFT, RIGHT) else mesh.bc[face][i]
            if typ == OPEN:
                return "dead"
            if typ == SYMMETRY:
                part.vel = mirror(part.vel, mesh.faceNormal(face))
                return "alive"
            if typ == PERIODIC:
                if face == LEFT:
                    part.pos[0] += (mesh.xd[0] - mesh.x0[0])
                elif face == RIGHT:
                    part.pos[0] -= (mesh.xd[0] - mesh.x0[0])
                elif face == BOTTOM:
                    part.pos[1] += (mesh.xd[1] - mesh.x0[1])
                else:
                    part.pos[1] -= (mesh.xd[1] - mesh.x0[1])
                return "alive"
            if typ == MESH_BC:
                index = jint(part.lc[1]) if face in (LEFT, RIGHT) else jint(part.lc[0])
                for k in range(2):
                    nb = mesh.nbr[face][index][k]
                    if nb is not None and self.meshes[nb].containsPos(part.pos):
                        part.lc = self.meshes[nb].XtoL(part.pos)
                        self.transfer[nb].append(part.copy())
                return "transfer"
            if typ == CIRCUIT and self.charge < 0:
                return "slow"
            return "dead"
        return "alive"

    def mover(self, m, plist, dt, particle_transfer):  # KM:298-422
        mesh = self.meshes[m]
        keep = []
        N = P0 = P1 = P2 = E = 0.0
        for part in plist:
            if part.mpw <= 0:
                continue
            if not particle_transfer:
                part.dt += dt
                self.kick(mesh, part)
            bounces, alive = 0, True
            while part.dt > 0 and bounces < 10:
                bounces += 1
                old = [part.pos[0], part.pos[1]]
                old_lc = [part.lc[0], part.lc[1]]
                part.pos[0] += part.vel[0] * part.dt
                part.pos[1] += part.vel[1] * part.dt
                if mesh.domain_type == RZ:  # KM:424-442
                    A = part.vel[2] * part.dt
                    B = part.pos[0]
                    R = math.sqrt(A * A + B * B)
                    cos, sin = B / R, A / R
                    part.pos[2] -= math.asin(sin)
                    v1, v2 = part.vel[0], part.vel[2]
                    part.pos[0] = R
                    part.vel[0] = cos * v1 + sin * v2
                    part.vel[2] = -sin * v1 + cos * v2
                elif mesh.domain_type == ZR:  # KM:444-462
                    A = part.vel[2] * part.dt
                    B = part.pos[1]
                    R = math.sqrt(A * A + B * B)
                    cos, sin = B / R, A / R
                    part.pos[2] += math.acos(cos)
                    v1, v2 = part.vel[1], part.vel[2]
                    part.pos[1] = R
                    part.vel[1] = cos * v1 + sin * v2
                    part.vel[2] = -sin * v1 + cos * v2
                else:
                    part.pos[2] += part.vel[2] * part.dt
                part.lc = mesh.XtoL(part.pos)
                res = self.process_boundary(m, part, old, old_lc)
                if res != "alive":
                    alive = False
                    if res == "dead":
                        self.n_exited += 1
                    if res == "slow":
                        self.slow.append((m, part, old, old_lc, bounces))
                    break
            if alive:
                N += part.mpw
                P0 += part.mpw * part.vel[0]
                P1 += part.mpw * part.vel[1]
                P2 += part.mpw * part.vel[2]
                E += part.mpw * math.sqrt(part.vel[0] * part.vel[0] + part.vel[1] * part.vel[1] + part.vel[2] * part.vel[2])
                keep.append(part)
        return keep, [N, P0, P1, P2, E]

    def finish_slow(self, m, part, old, old_lc, bounces):
        """What the Java host does with a particle handed over by sfgpu_take_slowpath: ProcessBoundary on the
        pre-ProcessBoundary state (KM:387), then the remaining sub-steps of the mover loop (KM:360-398).
        Returns True if the particle is still alive in mesh m."""
        mesh = self.meshes[m]
        res = self.process_boundary(m, part, old, old_lc)
        while res == "alive" and part.dt > 0 and bounces < 10:
            bounces += 1
            old = [part.pos[0], part.pos[1]]
            old_lc = [part.lc[0], part.lc[1]]
            part.pos[0] += part.vel[0] * part.dt
            part.pos[1] += part.vel[1] * part.dt
            if mesh.domain_type == XY:
                part.pos[2] += part.vel[2] * part.dt
            else:
                raise NotImplementedError("finish_slow: XY only in this test helper")
            part.lc = mesh.XtoL(part.pos)
            res = self.process_boundary(m, part, old, old_lc)
        if res == "dead":
            self.n_exited += 1
        return res == "alive"

    def updateFields(self, dt):  # KM:117-163
        self.slow, self.n_exited = [], 0
        self.n_absorbed, self.hits = 0, []
        self.sums = [0.0] * 5
        for m in range(len(self.meshes)):
            self.particles[m], s = self.mover(m, self.particles[m], dt, False)
            for k in range(5):
                self.sums[k] += s[k]
        for _ in range(10):
            for m in range(len(self.meshes)):
                if not self.transfer[m]:
                    continue
                tp = self.transfer[m]
                self.transfer[m] = []
                keep, _s = self.mover(m, tp, dt, True)
                self.particles[m] += [p for p in keep if all(math.isfinite(v) for v in p.vel)]
            if not any(self.transfer):
                break
        # KM:168-188 and KM:1580-1594, raw sums in the order den,u,v,w,uu,vv,ww,mpc
        self.raw = []
        for m, mesh in enumerate(self.meshes):
            F = [Field(mesh) for _ in range(8)]
            for part in self.particles[m]:
                F[0].scatter(part.lc, part.mpw)
                F[1].scatter(part.lc, part.vel[0] * part.mpw)
                F[2].scatter(part.lc, part.vel[1] * part.mpw)
                F[3].scatter(part.lc, part.vel[2] * part.mpw)
                F[4].scatter(part.lc, part.mpw * part.vel[0] * part.vel[0])
                F[5].scatter(part.lc, part.mpw * part.vel[1] * part.vel[1])
                F[6].scatter(part.lc, part.mpw * part.vel[2] * part.vel[2])
                ci, cj = jint(part.lc[0]), jint(part.lc[1])
                if 0 <= ci < mesh.ni and 0 <= cj < mesh.nj:
                    F[7].data[ci][cj] += 1
            self.raw.append([f.data for f in F])


class WallSegment:  # boundaries/LinearSegment.java:21-47, :113-179 as ProcessBoundary uses it
    def __init__(self, sid, x1, y1, x2, y2, kind=0, sink=False):
        self.sid, self.x1, self.x2 = sid, [x1, y1], [x2, y2]
        self.kind, self.sink = kind, sink
        dx, dy = x2 - x1, y2 - y1
        length = math.sqrt(dx * dx + dy * dy)
        dx /= length
        dy /= length
        self.normal = [-dy, dx, 0.0]

    @staticmethod
    def infinite_line_intersect(p1, p2, p3, p4):
        x1, x2, x3, x4 = p1[0], p2[0], p3[0], p4[0]
        y1, y2, y3, y4 = p1[1], p2[1], p3[1], p4[1]
        den = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4)
        if den == 0:
            return None
        return [((x1 * y2 - y1 * x2) * (x3 - x4) - (x1 - x2) * (x3 * y4 - y3 * x4)) / den,
                ((x1 * y2 - y1 * x2) * (y3 - y4) - (y1 - y2) * (x3 * y4 - y3 * x4)) / den]

    def intersect(self, p3, p4):
        p1, p2 = self.x1, self.x2
        xp = self.infinite_line_intersect(p1, p2, p3, p4)
        if xp is None:
            return [-1.0, -1.0]
        t = [0.0, 0.0]
        if abs(p2[0] - p1[0]) > 1e-6:
            t[0] = (xp[0] - p1[0]) / (p2[0] - p1[0])
        else:
            t[0] = (xp[1] - p1[1]) / (p2[1] - p1[1])
        if t[0] < -FLT_EPS or t[0] > (1 + FLT_EPS):
            return [-1.0, -1.0]
        if abs(p4[0] - p3[0]) > 1e-6:
            t[1] = (xp[0] - p3[0]) / (p4[0] - p3[0])
        else:
            t[1] = (xp[1] - p3[1]) / (p4[1] - p3[1])
        if t[1] < -FLT_EPS or t[1] > (1 + FLT_EPS):
            return [-1.0, -1.0]
        return [min(max(t[0], 0.0), 1.0), min(max(t[1], 0.0), 1.0)]


# ---------------------------------------------------------------------------------------------
# UniformSource on a Boundary of linear segments (XY), written from the Java, independently of the C oracle
# ---------------------------------------------------------------------------------------------
class JavaRandom:
    """java.util.Random (the JDK's documented 48-bit LCG).  Starfish.rnd() = random.nextDouble(), Starfish.java:244."""

    MASK = (1 << 48) - 1

    def __init__(self, seed=None, state=None):
        self.state = ((seed ^ 0x5DEECE66D) & self.MASK) if state is None else int(state)

    def next(self, bits):
        self.state = (self.state * 0x5DEECE66D + 0xB) & self.MASK
        v = self.state >> (48 - bits)
        return v - (1 << 32) if v >= (1 << 31) else v  # (int)

    def nextInt(self):
        return self.next(32)

    def nextDouble(self):
        return ((self.next(26) << 27) + self.next(27)) * (1.0 / (1 << 53))


class LinearSegment:  # boundaries/LinearSegment.java:20-101
    domain_type = XY  # Starfish.getDomainType()

    def area_to(self, t):  # LinearSegment.area(t), :50-80
        if self.domain_type == XY:
            return t * self.length
        pos = self.pos(t)
        if self.domain_type == RZ:
            r1, z1, r2, z2 = self.x1[0], self.x1[1], pos[0], pos[1]
        else:
            r1, z1, r2, z2 = self.x1[1], self.x1[0], pos[1], pos[0]
        dr, dz = r1 - r2, z1 - z2
        A = math.pi * (r1 + r2) * math.sqrt(dr * dr + dz * dz)
        return -A if A < 0 else A

    def __init__(self, x1, x2):
        self.x1, self.x2 = [float(x1[0]), float(x1[1])], [float(x2[0]), float(x2[1])]
        dx, dy = self.x2[0] - self.x1[0], self.x2[1] - self.x1[1]
        self.length = math.sqrt(dx * dx + dy * dy)
        dx /= self.length
        dy /= self.length
        self.normal = [-dy, dx, 0.0]
        self.area = self.area_to(1.0)  # Segment.area = area(1)

    def pos(self, t):
        return [self.x1[0] + t * (self.x2[0] - self.x1[0]), self.x1[1] + t * (self.x2[1] - self.x1[1]), 0.0]


class Spline:  # boundaries/Spline.java (linear segments)
    def __init__(self, points, domain_type=XY):
        self.domain_type = domain_type
        seg_cls = type("LinearSegment%d" % domain_type, (LinearSegment,), {"domain_type": domain_type})
        self.segments = [seg_cls(points[k], points[k + 1]) for k in range(len(points) - 1)]
        self.cum_area = [0.0]
        for s in self.segments:
            self.cum_area.append(self.cum_area[-1] + s.area)
        self.spline_area = self.cum_area[-1]

    @staticmethod
    def binarySearch(vec, val):  # Vec.java:529-546
        if val < vec[0]:
            return -1
        if val > vec[-1]:
            return len(vec)
        i1, i2 = 0, len(vec)
        while True:
            i_mid = int(0.5 * (i1 + i2))
            if val < vec[i_mid]:
                i2 = i_mid
            elif val > vec[i_mid]:
                i1 = i_mid
            else:
                return i_mid
            if i2 - i1 <= 1:
                return i1

    def randomT(self, rnd):  # Spline.java:582-641
        A1 = rnd.nextDouble() * self.spline_area
        i = self.binarySearch(self.cum_area, A1)
        seg = self.segments[i]
        seg_area = seg.area
        frac = (A1 - self.cum_area[i]) / seg_area
        if self.domain_type != XY:  # search the t that sweeps the wanted area (:594-637)
            max_steps, tol = 10, 1e-6
            x, f = [0.0] * max_steps, [0.0] * max_steps
            f_goal = frac * seg_area
            x[0] = frac
            f[0] = seg.area_to(x[0])
            diff = abs(f[0] - f_goal) / seg_area
            k = 1
            x[1] = x[0] + (f_goal - f[0])
            if diff > tol:
                f[1] = seg.area_to(x[1])
            while diff > tol and k < max_steps - 1:
                x[k + 1] = (x[k] - x[k - 1]) * (f_goal - f[k - 1]) / (f[k] - f[k - 1]) + x[k - 1]
                f[k + 1] = seg.area_to(x[k + 1])
                diff = abs(f[k + 1] - f_goal) / seg_area
                k += 1
            frac = x[k]
        return i + frac

    def pos(self, t):  # Spline.java:700-707
        si = jint(t)
        seg_t = t - si
        if si > len(self.segments) - 1:
            si, seg_t = len(self.segments) - 1, 1.0
        return self.segments[si].pos(seg_t)

    def normal(self, t):  # Spline.java:947-954
        si = min(jint(t), len(self.segments) - 1)
        return list(self.segments[si].normal)


def uniform_source_sample(km, spline, v_drift, num_mp, dt, rnd, spwt0, born_it=0, cold_beam=False):
    """Source.sampleKinetic (Source.java:167-198) with UniformSource.sampleParticle (sources/UniformSource.java:56-72) and
    KineticMaterial.addParticle(Particle) (KM:810-818) + DomainModule.getMesh (DomainModule.java:106-117)."""
    count = 0
    while num_mp > 0:
        t = spline.randomT(rnd)
        x, n = spline.pos(t), spline.normal(t)
        vel = [n[0] * v_drift, n[1] * v_drift, 0.0] if cold_beam else [n[k] * v_drift for k in range(3)]  # ColdBeamSource.java:56-76 / UniformSource.java:56-72
        part = Particle([x[0], x[1], 0.0], vel, spwt0)
        part.born_it = born_it
        num_mp -= 1
        for k in range(3):
            part.pos[k] += part.vel[k] * 1e-6 * dt
        mesh = None
        for m, mm in enumerate(km.meshes):  # containsPosStrict, UM:164-171
            if part.pos[0] >= mm.x0[0] and part.pos[0] < mm.xd[0] and part.pos[1] >= mm.x0[1] and part.pos[1] < mm.xd[1]:
                mesh = m
                break
        if mesh is None:
            for m, mm in enumerate(km.meshes):
                if mm.containsPos(part.pos):
                    mesh = m
                    break
        if mesh is None:
            continue
        km.addParticle(mesh, part, dt)
        count += 1
    return count
