"""Host-side mirror of the mesh state the hot path reads.

Mirrors ``starfish.core.domain.UniformMesh`` (UniformMesh.java:33-43, :139-161) and the parts of
``Mesh`` the particle path touches: face boundary types (Mesh.java:107-118, :140-155, :215),
mesh neighbours (Mesh.java:160-165), node volumes (Mesh.java:998-1020) and ``containsPos``
(Mesh.java:1476-1483).  Pure data + numpy: no particle arithmetic happens here.
"""
from __future__ import annotations

import enum

import numpy as np

FLT_EPS = 1e-7  # Constants.java:26


class DomainType(enum.IntEnum):  # DomainModule.java:28
    XY = 0
    RZ = 1
    ZR = 2


class Face(enum.IntEnum):  # Mesh.java:107-118
    RIGHT = 0
    TOP = 1
    LEFT = 2
    BOTTOM = 3


class DomainBoundaryType(enum.IntEnum):  # Mesh.java:140-155
    OPEN = -1
    DIRICHLET = 0
    NEUMANN = 1
    PERIODIC = 2
    SYMMETRY = 3
    MESH = 4
    SINK = 5
    CIRCUIT = 6


class UniformMesh:
    """Rectilinear mesh with uniform spacing: ``pos(i,j) = x0 + (i,j)*dh`` (UniformMesh.java:139-145)."""

    def __init__(self, ni, nj, x0, dh, domain_type=DomainType.XY, name="mesh"):
        self.ni, self.nj = int(ni), int(nj)
        self.x0 = np.array(x0, dtype=np.float64)
        self.dh = np.array(dh, dtype=np.float64)
        self.domain_type = DomainType(domain_type)
        self.name = name
        # xd = x0 + (n-1)*dh, UniformMesh.java:131-135
        self.xd = np.array([self.x0[0] + (self.ni - 1) * self.dh[0], self.x0[1] + (self.nj - 1) * self.dh[1]])
        # per-face per-node boundary type, default OPEN (Mesh.java:162)
        self.bc = [np.full(self._face_len(f), int(DomainBoundaryType.OPEN), dtype=np.int8) for f in range(4)]
        # per-face per-node two neighbour mesh ids, -1 = none (Mesh.java:160-165)
        self.nbr = [np.full((self._face_len(f), 2), -1, dtype=np.int32) for f in range(4)]
        self.has_seg = np.zeros((self.ni, self.nj), dtype=np.uint8)
        self.efi = np.zeros((self.ni, self.nj))
        self.efj = np.zeros((self.ni, self.nj))
        self.bfi = None
        self.bfj = None
        self.node_vol = self._node_volumes()
        self.index = -1  # set when attached to a material

    def _face_len(self, f):
        return self.nj if f in (Face.RIGHT, Face.LEFT) else self.ni

    # Mesh.setMeshBCType, Mesh.java:191-207
    def setMeshBCType(self, face, bc_type):
        self.bc[int(face)][:] = int(bc_type)

    def boundaryType(self, face, index):  # Mesh.java:215-217
        return DomainBoundaryType(int(self.bc[int(face)][index]))

    def setNeighbor(self, face, index, slot, mesh_index):
        self.nbr[int(face)][index, slot] = mesh_index
        self.bc[int(face)][index] = int(DomainBoundaryType.MESH)

    def XtoL(self, x):  # UniformMesh.java:154-161
        x = np.asarray(x, dtype=np.float64)
        return np.stack([(x[..., 0] - self.x0[0]) / self.dh[0], (x[..., 1] - self.x0[1]) / self.dh[1]], axis=-1)

    def pos(self, lc):  # UniformMesh.java:139-145
        lc = np.asarray(lc, dtype=np.float64)
        return np.stack([self.x0[0] + lc[..., 0] * self.dh[0], self.x0[1] + lc[..., 1] * self.dh[1]], axis=-1)

    def containsPos(self, x):  # Mesh.java:1476-1483
        lc = self.XtoL(x)
        return ~((lc[..., 0] < -FLT_EPS) | (lc[..., 1] < -FLT_EPS) | (lc[..., 0] > self.ni - 1 + FLT_EPS)
                 | (lc[..., 1] > self.nj - 1 + FLT_EPS))

    def _node_volumes(self):
        """Node control volumes in the spirit of Mesh.nodeVol (Mesh.java:998-1020): half cells on mesh
        edges, XY depth 1, axisymmetric volumes 2*pi*r*area with the r +- 0.25*dr shift on the edge
        rows.  Mesh setup stays in Java; this array is an INPUT of the path (Field2D.scaleByVol)."""
        wi = np.ones(self.ni)
        wi[0] = wi[-1] = 0.5
        wj = np.ones(self.nj)
        wj[0] = wj[-1] = 0.5
        area = np.outer(wi * self.dh[0], wj * self.dh[1])
        if self.domain_type == DomainType.XY:
            return area
        if self.domain_type == DomainType.RZ:
            ii = np.arange(self.ni, dtype=np.float64)
            ii[0] += 0.25
            ii[-1] -= 0.25
            r = (self.x0[0] + ii * self.dh[0])[:, None]
        else:
            jj = np.arange(self.nj, dtype=np.float64)
            jj[0] += 0.25
            jj[-1] -= 0.25
            r = (self.x0[1] + jj * self.dh[1])[None, :]
        return 2 * np.pi * area * r


def set_mesh_neighbors(meshes):
    """``Mesh.setMeshNeighbors`` (Mesh.java:405-468) for a list of uniform meshes: a boundary node whose position lies
    inside another mesh (``containsPos``, FLT_EPS tolerance) becomes a MESH boundary with up to two neighbours
    (``addMeshToBoundary``, Mesh.java:476-489).  Interior face nodes first, then the corners, which only join a face that
    is already a MESH face next to them -- the rule that keeps a mesh stacked on top of another from getting a MESH
    boundary on its side face.  Mesh setup stays in Java; this mirror only builds inputs for the tests and examples."""
    def add(mesh, face, index, other):
        row = mesh.nbr[int(face)][index]
        mesh.bc[int(face)][index] = int(DomainBoundaryType.MESH)
        if row[0] < 0:
            row[0] = other
        else:
            row[1] = other

    for me in meshes:
        ni, nj = me.ni, me.nj
        pos = lambda i, j: np.array([me.x0[0] + i * me.dh[0], me.x0[1] + j * me.dh[1]])
        for k, other in enumerate(meshes):
            if other is me:
                continue
            for j in range(1, nj - 1):
                if other.containsPos(pos(0, j)):
                    add(me, Face.LEFT, j, k)
                if other.containsPos(pos(ni - 1, j)):
                    add(me, Face.RIGHT, j, k)
            for i in range(1, ni - 1):
                if other.containsPos(pos(i, 0)):
                    add(me, Face.BOTTOM, i, k)
                if other.containsPos(pos(i, nj - 1)):
                    add(me, Face.TOP, i, k)
        is_mesh = lambda face, idx: me.bc[int(face)][idx] == int(DomainBoundaryType.MESH)
        for k, other in enumerate(meshes):
            if other is me:
                continue
            if is_mesh(Face.LEFT, 1) and other.containsPos(pos(0, 0)):
                add(me, Face.LEFT, 0, k)
            if is_mesh(Face.LEFT, nj - 2) and other.containsPos(pos(0, nj - 1)):
                add(me, Face.LEFT, nj - 1, k)
            if is_mesh(Face.RIGHT, 1) and other.containsPos(pos(ni - 1, 0)):
                add(me, Face.RIGHT, 0, k)
            if is_mesh(Face.RIGHT, nj - 2) and other.containsPos(pos(ni - 1, nj - 1)):
                add(me, Face.RIGHT, nj - 1, k)
            if is_mesh(Face.TOP, 1) and other.containsPos(pos(0, nj - 1)):
                add(me, Face.TOP, 0, k)
            if is_mesh(Face.TOP, ni - 2) and other.containsPos(pos(ni - 1, nj - 1)):
                add(me, Face.TOP, ni - 1, k)
            if is_mesh(Face.BOTTOM, 1) and other.containsPos(pos(0, 0)):
                add(me, Face.BOTTOM, 0, k)
            if is_mesh(Face.BOTTOM, ni - 2) and other.containsPos(pos(ni - 1, 0)):
                add(me, Face.BOTTOM, ni - 1, k)


class LinearSpline:
    """A Boundary made of linear segments as the reference's ``Spline`` holds it for an XY domain: per segment the end
    points, ``LinearSegment.normal`` = (-dy, dx, 0)/length (LinearSegment.java:20-45) and ``area`` = length
    (LinearSegment.area, XY), plus ``cum_area`` and ``spline_area`` (Spline.java).  Geometry set-up stays in Java; this
    mirror only builds the INPUT of the source sampling for the tests and examples.  (pos(1) = x1 + 1*(x2 - x1) is used for the
    axisymmetric area, like LinearSegment.area(1).)"""

    def __init__(self, points, domain_type=DomainType.XY):
        pts = np.asarray(points, np.float64)
        assert pts.ndim == 2 and pts.shape[1] == 2 and len(pts) >= 2
        self.x1, self.y1 = np.ascontiguousarray(pts[:-1, 0]), np.ascontiguousarray(pts[:-1, 1])
        self.x2, self.y2 = np.ascontiguousarray(pts[1:, 0]), np.ascontiguousarray(pts[1:, 1])
        dx, dy = self.x2 - self.x1, self.y2 - self.y1
        length = np.sqrt(dx * dx + dy * dy)
        dx, dy = dx / length, dy / length
        self.nx, self.ny = np.ascontiguousarray(-dy), np.ascontiguousarray(dx)
        if domain_type == DomainType.XY:
            self.area = np.ascontiguousarray(length)  # LinearSegment.area(1) = length
        else:  # lateral area of the conical frustum swept by the segment, LinearSegment.java:62-79
            px, py = self.x1 + 1.0 * (self.x2 - self.x1), self.y1 + 1.0 * (self.y2 - self.y1)  # pos(1), not x2: same rounding as the Java
            r1, z1, r2, z2 = (self.x1, self.y1, px, py) if domain_type == DomainType.RZ else (self.y1, self.x1, py, px)
            dr, dz = r1 - r2, z1 - z2
            self.area = np.ascontiguousarray(np.abs(np.pi * (r1 + r2) * np.sqrt(dr * dr + dz * dz)))
        self.cum_area = np.zeros(len(length) + 1)
        for k in range(len(length)):  # sequential sum, like the Java loop
            self.cum_area[k + 1] = self.cum_area[k] + self.area[k]
        self.spline_area = float(self.cum_area[-1])
        self.n_seg = len(length)


# ---------------------------------------------------------------------------------------------------------------------
# solid boundaries: which nodes own which segments (host-side set-up, stays Java in production)
# ---------------------------------------------------------------------------------------------------------------------
def _jint(d):
    """Java (int)double: toward zero, NaN -> 0, saturating."""
    if d != d:
        return 0
    return int(max(min(d, 2147483647.0), -2147483648.0))


def _segment_intersect(x1, y1, x2, y2, p3, p4):
    """LinearSegment.intersect(p3, p4), LinearSegment.java:113-179 -> (t_segment, t_p3p4) or (-1, -1)."""
    x3, y3, x4, y4 = p3[0], p3[1], p4[0], p4[1]
    den = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4)
    if den == 0:
        return -1.0, -1.0
    xp0 = ((x1 * y2 - y1 * x2) * (x3 - x4) - (x1 - x2) * (x3 * y4 - y3 * x4)) / den
    xp1 = ((x1 * y2 - y1 * x2) * (y3 - y4) - (y1 - y2) * (x3 * y4 - y3 * x4)) / den
    t0 = (xp0 - x1) / (x2 - x1) if abs(x2 - x1) > 1e-6 else (xp1 - y1) / (y2 - y1)
    if t0 < -FLT_EPS or t0 > 1 + FLT_EPS:
        return -1.0, -1.0
    t1 = (xp0 - x3) / (x4 - x3) if abs(x4 - x3) > 1e-6 else (xp1 - y3) / (y4 - y3)
    if t1 < -FLT_EPS or t1 > 1 + FLT_EPS:
        return -1.0, -1.0
    return min(max(t0, 0.0), 1.0), min(max(t1, 0.0), 1.0)


class SolidBoundary:
    """A ``Boundary`` of type DIRICHLET ("solid") or SINK made of linear segments (boundaries.xml ``<path>M .. L ..</path>``),
    with the outcome ``Material.performSurfaceInteraction`` (Material.java:279-300) has for the species that hits it:
    ``kind`` 0 = the particle dies (no interaction listed for the pair, or ABSORB), 1 = it lives on unchanged (NONE), 2 = SPECULAR
    without a species change (SurfaceInteraction.java:104-149: ``vel += normal * |vel_xy| * sqrt(2)``, alive)."""

    def __init__(self, name, points, kind=0, sink=False):
        pts = np.asarray(points, np.float64)
        self.name, self.kind, self.sink = name, int(kind), bool(sink)
        self.x1, self.y1 = np.ascontiguousarray(pts[:-1, 0]), np.ascontiguousarray(pts[:-1, 1])
        self.x2, self.y2 = np.ascontiguousarray(pts[1:, 0]), np.ascontiguousarray(pts[1:, 1])
        self.n_seg = len(self.x1)


def set_boundaries(mesh, boundaries):
    """``Mesh.setNodeControlVolumes`` (Mesh.java:1215-1290) for the DIRICHLET / SINK boundaries: every node whose +-1.01-cell box
    is touched by a segment owns it (``node.segments``).  Fills ``mesh.has_seg`` (KM:508-518) and the segment tables the device
    and the oracle take: ``mesh.segments`` = dict(x1,y1,x2,y2,kind,sink) over all segments and the node CSR ``seg_offs``/``seg_ids``."""
    ni, nj = mesh.ni, mesh.nj
    owners = [[[] for _ in range(nj)] for _ in range(ni)]
    seg = dict(x1=[], y1=[], x2=[], y2=[], kind=[], sink=[], boundary=[], index=[])
    for b_id, b in enumerate(boundaries):
        for s in range(b.n_seg):
            sid = len(seg["x1"])
            x1, y1, x2, y2 = float(b.x1[s]), float(b.y1[s]), float(b.x2[s]), float(b.y2[s])
            for k, v in zip(("x1", "y1", "x2", "y2", "kind", "sink", "boundary", "index"), (x1, y1, x2, y2, b.kind, int(b.sink), b_id, s)):
                seg[k].append(v)
            lo = mesh.XtoL(np.array([min(x1, x2), min(y1, y2)]))  # Segment.getBox + Mesh.XtoI
            hi = mesh.XtoL(np.array([max(x1, x2), max(y1, y2)]))
            lcm = [_jint(lo[0]) - 1, _jint(lo[1]) - 1]
            lcp = [_jint(hi[0]) + 2, _jint(hi[1]) + 2]
            lcm = [max(lcm[0], 0), max(lcm[1], 0)]
            lcp = [min(lcp[0], ni - 1), min(lcp[1], nj - 1)]
            for j in range(lcm[1], lcp[1] + 1):
                for i in range(lcm[0], lcp[0] + 1):
                    bsize = 1.01
                    b1 = mesh.pos(np.array([i - bsize, j - bsize]))
                    b2 = mesh.pos(np.array([i + bsize, j + bsize]))
                    inbox = lambda px, py: b1[0] <= px <= b2[0] and b1[1] <= py <= b2[1]
                    hit = inbox(x1, y1) or inbox(x2, y2)
                    if not hit:  # Segment.segmentInBox: cuts of the four box faces, Segment.java:246-290
                        faces = (((b1[0], b1[1]), (b2[0], b1[1])), ((b2[0], b1[1]), (b2[0], b2[1])), ((b2[0], b2[1]), (b1[0], b2[1])),
                                 ((b1[0], b2[1]), (b1[0], b1[1])))
                        hit = any(_segment_intersect(x1, y1, x2, y2, f1, f2)[0] >= 0 for f1, f2 in faces)
                    if hit and sid not in owners[i][j]:
                        owners[i][j].append(sid)
    mesh.has_seg = np.zeros((ni, nj), dtype=np.uint8)
    offs, ids = [0], []
    for i in range(ni):
        for j in range(nj):
            if owners[i][j]:
                mesh.has_seg[i, j] = 1
            ids += owners[i][j]
            offs.append(len(ids))
    mesh.segments = {k: np.ascontiguousarray(v, np.float64 if k in ("x1", "y1", "x2", "y2") else np.int32) for k, v in seg.items()}
    mesh.seg_offs = np.ascontiguousarray(offs, np.int32)
    mesh.seg_ids = np.ascontiguousarray(ids, np.int32)
    mesh.node_segments = owners
    return mesh
