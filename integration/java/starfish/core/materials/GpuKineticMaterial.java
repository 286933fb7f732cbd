/*
 * GpuKineticMaterial -- KineticMaterial whose per-step move + deposit runs on a B200 through libstarfish_gpu.so.
 *
 * Drop-in: everything else in Starfish (sources, solvers, interactions, output, restart) keeps working against the
 * public KineticMaterial / Material API.  Declared in package starfish.core.materials because ProcessBoundary
 * (KineticMaterial.java:471) and the mover sums (Material.java:204-206) are package-private; the project is not modular,
 * so a class-path entry suffices.  What runs where:
 *   GPU   moveParticles + transfer sweeps (KM:126-142), updateFields(MeshData) (KM:168-197), updateSamples (KM:1570-1595),
 *         addParticle's XtoL / clamp / -0.5dt rewind (KM:759-802)
 *   Java  ProcessBoundary for the few particles whose sub-step touches a DIRICHLET / SINK segment node or a CIRCUIT face
 *         (handed back UNMOVED by sfgpu_take_slowpath), computeFields (KM:1606-1671), updateBoundaries (Material.java:650-657)
 * Only UniformMesh domains are accepted (BASELINE north_star); other mesh types must keep type="kinetic".
 */
package starfish.core.materials;

import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.nio.DoubleBuffer;
import java.lang.reflect.Field;
import java.util.ArrayList;
import java.util.HashMap;
import java.util.Iterator;

import org.w3c.dom.Element;

import starfish.core.boundaries.Boundary;
import starfish.core.boundaries.Boundary.BoundaryType;
import starfish.core.boundaries.Segment;
import starfish.core.common.Starfish;
import starfish.core.common.Starfish.Log;
import starfish.core.domain.DomainModule.DomainType;
import starfish.core.domain.Field2D;
import starfish.core.domain.Mesh;
import starfish.core.domain.Mesh.DomainBoundaryType;
import starfish.core.domain.Mesh.Face;
import starfish.core.domain.Mesh.MeshBoundaryData;
import starfish.core.domain.UniformMesh;
import starfish.interactions.MaterialInteraction;
import starfish.interactions.SurfaceInteraction;

public class GpuKineticMaterial extends KineticMaterial {
    private long ctx = 0;
    private int sp = -1;
    private int nMesh = 0;
    private Mesh[] meshes;
    private boolean needsSlowPath = false;
    /* page-locked planes per mesh, double[ni][nj] flattened as i*nj + j */
    private ByteBuffer[] efi, efj, bfi, bfj, nd, u, v, w;
    private ByteBuffer[][] samples; /* [mesh][8]: count,u,v,w,uu,vv,ww,mpc sums */
    /* particles added since the last step, per mesh (KM:759: sources call addParticle one particle at a time) */
    private ArrayList<ArrayList<Particle>> pending = new ArrayList<>();
    private boolean gpu_first_time = true;
    /* surfaces handled on the device (sfgpu_mesh_set_segments): per mesh the segments in table order and their SFGPU_SURFACE_* outcome; null = host path */
    private Segment[][] devSeg;
    private int[][] devKind;
    private boolean deviceSurfaces = false;

    public GpuKineticMaterial(String name, Element element) {
        super(name, element);
        Log.log("> kinetic path: libstarfish_gpu.so (B200)");
    }

    private void check(long rc, String what) {
        if (rc < 0)
            Log.error("sfgpu " + what + ": " + SfgpuJni.lastError(ctx)); /* exits (console) or throws (GUI), LoggerModule.java:170-176 */
    }

    private static ByteBuffer plane(int n) {
        ByteBuffer b = SfgpuJni.hostAlloc(8L * n);
        if (b == null)
            Log.error("sfgpu_host_alloc failed: " + SfgpuJni.lastError(0));
        return b.order(ByteOrder.nativeOrder());
    }

    private static void flatten(double[][] src, ByteBuffer dst) {
        DoubleBuffer d = dst.asDoubleBuffer();
        for (int i = 0; i < src.length; i++)
            d.put(src[i]); /* rows in order: index i*nj + j */
    }

    private static void unflatten(ByteBuffer src, double[][] dst) {
        DoubleBuffer d = src.asDoubleBuffer();
        for (int i = 0; i < dst.length; i++)
            d.get(dst[i]);
    }

    /**
     * What Material.performSurfaceInteraction (Material.java:279-300) does to THIS material on a segment, if that is deterministic:
     * SURFACE_REMOVE (nothing listed for the pair, or ABSORB), SURFACE_NONE (NONE, or a boundary without material, KM:586), SURFACE_SPECULAR
     * (SPECULAR without a species change); -1 when the outcome needs the Java code (models that draw random numbers, emission hooks,
     * several handlers with probabilities, curved segments).  MaterialInteraction keeps its handler package-private: read by reflection.
     */
    private int surfaceOutcome(Segment seg) {
        if (!seg.getClass().getSimpleName().equals("LinearSegment"))
            return -1;
        Material target = seg.getBoundary().getMaterial(0.5);
        if (target == null)
            return SfgpuJni.SURFACE_NONE;
        if (target.target_interactions != null && !target.target_interactions.getInteractionList(mat_index).isEmpty())
            return -1; /* sputtering / emission hooks run first, Material.java:281-286 */
        if (target.source_interactions == null)
            return SfgpuJni.SURFACE_REMOVE;
        ArrayList<MaterialInteraction> list = target.source_interactions.getInteractionList(mat_index);
        if (list.isEmpty())
            return SfgpuJni.SURFACE_REMOVE; /* Material.java:291-295 */
        if (list.size() != 1 || list.get(0).getProbability() != 1.0)
            return -1;
        try {
            MaterialInteraction mi = list.get(0);
            Field hf = MaterialInteraction.class.getDeclaredField("surface_impact_handler");
            hf.setAccessible(true);
            Object handler = hf.get(mi);
            if (handler == SurfaceInteraction.SurfaceImpactAbsorb)
                return SfgpuJni.SURFACE_REMOVE;
            if (handler == SurfaceInteraction.SurfaceImpactNone)
                return SfgpuJni.SURFACE_NONE;
            if (handler == SurfaceInteraction.SurfaceImpactSpecular) {
                Field sf = MaterialInteraction.class.getDeclaredField("source_mat");
                Field pf = MaterialInteraction.class.getDeclaredField("product_mat");
                Field kf = MaterialInteraction.class.getDeclaredField("product_km_mat");
                sf.setAccessible(true); pf.setAccessible(true); kf.setAccessible(true);
                boolean speciesChange = sf.get(mi) != pf.get(mi) && kf.get(mi) != null; /* SurfaceInteraction.java:128 */
                return speciesChange ? -1 : SfgpuJni.SURFACE_SPECULAR;
            }
        } catch (ReflectiveOperationException | SecurityException ex) {
            Log.warning("sfgpu: cannot inspect the surface model of " + seg.getBoundary().getName() + " (" + ex + "): host path");
        }
        return -1;
    }

    @Override
    public void init() {
        super.init(); /* KM:83-112: MeshData per mesh, sample fields */
        ArrayList<Mesh> list = Starfish.getMeshList();
        nMesh = list.size();
        meshes = list.toArray(new Mesh[0]);
        DomainType dom = Starfish.getDomainType();
        int domain = dom == DomainType.XY ? 0 : (dom == DomainType.RZ ? 1 : 2); /* SFGPU_XY / RZ / ZR */
        ctx = SfgpuJni.create(0, domain);
        if (ctx == 0)
            Log.error("sfgpu_create: " + SfgpuJni.lastError(0));
        efi = new ByteBuffer[nMesh]; efj = new ByteBuffer[nMesh]; bfi = new ByteBuffer[nMesh]; bfj = new ByteBuffer[nMesh];
        nd = new ByteBuffer[nMesh]; u = new ByteBuffer[nMesh]; v = new ByteBuffer[nMesh]; w = new ByteBuffer[nMesh];
        samples = new ByteBuffer[nMesh][SfgpuJni.NFIELDS];
        devSeg = new Segment[nMesh][];
        devKind = new int[nMesh][];
        for (int m = 0; m < nMesh; m++) {
            if (!(meshes[m] instanceof UniformMesh))
                Log.error("type=\"kinetic_gpu\" needs uniform meshes; mesh " + meshes[m].getName() + " is not");
            UniformMesh mesh = (UniformMesh) meshes[m];
            final int ni = mesh.ni, nj = mesh.nj;
            /* per-face DomainBoundaryType.value() per node (MESH:140-155, :215) and neighbour mesh indices (MESH:160-165) */
            byte[][] bc = new byte[4][];
            int[][] nbr = new int[4][];
            for (Face face : Face.values()) {
                int n = (face == Face.LEFT || face == Face.RIGHT) ? nj : ni;
                bc[face.val()] = new byte[n];
                nbr[face.val()] = new int[2 * n];
                for (int k = 0; k < n; k++) {
                    DomainBoundaryType type = mesh.boundaryType(face, k);
                    bc[face.val()][k] = (byte) type.value();
                    if (type == DomainBoundaryType.CIRCUIT)
                        needsSlowPath = true;
                    MeshBoundaryData bd = mesh.boundaryData(face, k);
                    for (int q = 0; q < 2; q++)
                        nbr[face.val()][2 * k + q] = bd.neighbor[q] == null ? -1 : list.indexOf(bd.neighbor[q]);
                }
            }
            /* nodes that own a DIRICHLET or SINK segment (KM:508-518), and node.segments as a CSR over the distinct segments of the mesh */
            byte[] hasSeg = new byte[ni * nj];
            ArrayList<Segment> segList = new ArrayList<>();
            HashMap<Segment, Integer> segIndex = new HashMap<>();
            int[] nodeOffs = new int[ni * nj + 1];
            ArrayList<Integer> nodeIds = new ArrayList<>();
            for (int i = 0; i < ni; i++)
                for (int j = 0; j < nj; j++) {
                    nodeOffs[i * nj + j] = nodeIds.size();
                    for (Segment seg : mesh.getNode(i, j).segments)
                        if (seg.getBoundaryType() == BoundaryType.DIRICHLET || seg.getBoundaryType() == BoundaryType.SINK) {
                            hasSeg[i * nj + j] = 1;
                            Integer k = segIndex.get(seg);
                            if (k == null) {
                                k = segList.size();
                                segIndex.put(seg, k);
                                segList.add(seg);
                            }
                            if (!nodeIds.subList(nodeOffs[i * nj + j], nodeIds.size()).contains(k))
                                nodeIds.add(k);
                        }
                }
            nodeOffs[ni * nj] = nodeIds.size();
            /* surfaces on the device when every segment of the mesh has a deterministic outcome for this material (sfgpu.h: all or none per mesh) */
            int[] kinds = new int[segList.size()];
            boolean onDevice = !segList.isEmpty();
            for (int k = 0; k < kinds.length && onDevice; k++) {
                kinds[k] = surfaceOutcome(segList.get(k));
                onDevice = kinds[k] >= 0;
            }
            if (!segList.isEmpty() && !onDevice)
                needsSlowPath = true;
            double[] nodeVol = new double[ni * nj];
            double[][] nv = Starfish.getFieldCollection("NodeVol").getField(mesh).getData();
            for (int i = 0; i < ni; i++)
                System.arraycopy(nv[i], 0, nodeVol, i * nj, nj);
            int id = SfgpuJni.meshAdd(ctx, ni, nj, mesh.x0, mesh.dh, bc, nbr, hasSeg, nodeVol);
            check(id, "mesh_add");
            if (id != m)
                Log.error("sfgpu mesh ids must follow Starfish.getMeshList()");
            if (onDevice) {
                int ns = segList.size();
                double[][] xy = new double[4][ns];
                int[] sink = new int[ns], ids = new int[nodeIds.size()];
                for (int k = 0; k < ns; k++) {
                    double[] p1 = segList.get(k).firstPoint(), p2 = segList.get(k).lastPoint();
                    xy[0][k] = p1[0]; xy[1][k] = p1[1]; xy[2][k] = p2[0]; xy[3][k] = p2[1];
                    sink[k] = segList.get(k).getBoundaryType() == BoundaryType.SINK ? 1 : 0;
                }
                for (int k = 0; k < ids.length; k++)
                    ids[k] = nodeIds.get(k);
                check(SfgpuJni.meshSetSegments(ctx, m, ns, xy, kinds, sink, nodeOffs, ids), "mesh_set_segments");
                devSeg[m] = segList.toArray(new Segment[0]);
                devKind[m] = kinds;
                deviceSurfaces = true;
                Log.log("> mesh " + mesh.getName() + ": " + ns + " surface segments handled on the GPU");
            }
            efi[m] = plane(ni * nj); efj[m] = plane(ni * nj); bfi[m] = plane(ni * nj); bfj[m] = plane(ni * nj);
            nd[m] = plane(ni * nj); u[m] = plane(ni * nj); v[m] = plane(ni * nj); w[m] = plane(ni * nj);
            for (int f = 0; f < SfgpuJni.NFIELDS; f++)
                samples[m][f] = plane(ni * nj);
            pending.add(new ArrayList<Particle>());
        }
        sp = SfgpuJni.speciesAdd(ctx, charge, mass, 0);
        check(sp, "species_add");
    }

    /** KM:759-802.  The XtoL / plus-edge clamp / -0.5dt rewind / id assignment happen on the device when the batch is flushed. */
    @Override
    public boolean addParticle(MeshData md, Particle part) {
        pending.get(md.mesh.getIndex()).add(part);
        return true;
    }

    private static double[][] soa(int n, boolean withLc) {
        double[][] a = new double[10][];
        for (int k = 0; k < 7; k++)
            a[k] = new double[n];
        if (withLc)
            for (int k = 7; k < 10; k++)
                a[k] = new double[n];
        return a;
    }

    private static void put(double[][] a, int k, Particle p, boolean withLc) {
        a[0][k] = p.pos[0]; a[1][k] = p.pos[1]; a[2][k] = p.pos[2];
        a[3][k] = p.vel[0]; a[4][k] = p.vel[1]; a[5][k] = p.vel[2];
        a[6][k] = p.mpw;
        if (withLc) {
            a[7][k] = p.lc[0]; a[8][k] = p.lc[1]; a[9][k] = p.dt;
        }
    }

    /** particles with lc == null (the sources' case) go as one batch; the rare caller-supplied lc (restart) as a second one */
    private void flushPending(int m) {
        ArrayList<Particle> list = pending.get(m);
        if (list.isEmpty())
            return;
        ArrayList<Particle> plain = new ArrayList<>(), withLc = new ArrayList<>();
        for (Particle p : list)
            (p.lc == null ? plain : withLc).add(p);
        for (int pass = 0; pass < 2; pass++) {
            ArrayList<Particle> src = pass == 0 ? plain : withLc;
            if (src.isEmpty())
                continue;
            double[][] a = soa(src.size(), pass == 1);
            int[] born = new int[src.size()];
            for (int k = 0; k < src.size(); k++) {
                put(a, k, src.get(k), pass == 1);
                if (pass == 1)
                    a[9][k] = 0; /* addParticle overwrites dt (KM:777, :795) */
                born[k] = src.get(k).born_it;
            }
            check(SfgpuJni.inject(ctx, sp, m, src.size(), a, null /* id = part_id_counter++ on the device, KM:797 */, born, Starfish.getDt(),
                    SfgpuJni.INJECT_REWIND), "inject");
        }
        list.clear();
    }

    @Override
    public void updateFields() { /* KM:117-163 */
        if (Starfish.steady_state() && !steady_state) {
            clearSamples();
            steady_state = true;
        }
        final double dt = Starfish.getDt();
        for (int m = 0; m < nMesh; m++) {
            flushPending(m);
            MeshData md = mesh_data[m];
            flatten(md.Efi.getData(), efi[m]);
            flatten(md.Efj.getData(), efj[m]);
            flatten(md.Bfi.getData(), bfi[m]);
            flatten(md.Bfj.getData(), bfj[m]);
            check(SfgpuJni.setFields(ctx, m, efi[m], efj[m], bfi[m], bfj[m]), "set_fields"); /* KM:1319-1322 */
        }
        /* move + transfer sweeps + deposit + running sums */
        check(SfgpuJni.step(ctx, sp, dt, needsSlowPath ? SfgpuJni.STEP_DEFER_FINISH : 0), "step");
        if (needsSlowPath) {
            finishSlowPath(dt);
            check(SfgpuJni.finishStep(ctx, sp), "finish_step");
        }
        if (deviceSurfaces)
            applySurfaceHits();
        double[] s5 = new double[5];
        long[] counts = new long[3];
        check(SfgpuJni.getSums(ctx, sp, s5, counts), "get_sums");
        mass_sum = s5[0] * mass; /* KM:252-258 */
        momentum_sum[0] = s5[1] * mass; momentum_sum[1] = s5[2] * mass; momentum_sum[2] = s5[3] * mass;
        energy_sum = s5[4] * mass;
        for (int m = 0; m < nMesh; m++) { /* KM:168-197: nd, u, v, w */
            check(SfgpuJni.getMoments(ctx, sp, m, nd[m], u[m], v[m], w[m]), "get_moments");
            unflatten(nd[m], getDen(meshes[m]).getData());
            unflatten(u[m], getU(meshes[m]).getData());
            unflatten(v[m], getV(meshes[m]).getData());
            unflatten(w[m], getW(meshes[m]).getData());
        }
        /* updateGasProperties (KM:1553-1563): the running sums live on the device; computeFields needs them every 10 steps */
        if (gpu_first_time || Starfish.getIt() % 10 == 0) {
            long[] ns = new long[1];
            final String[] names = { "count-sum", "u-sum", "v-sum", "w-sum", "uu-sum", "vv-sum", "ww-sum", "mpc-sum" };
            for (int m = 0; m < nMesh; m++) {
                check(SfgpuJni.getSamples(ctx, sp, m, samples[m], ns), "get_samples");
                for (int f = 0; f < SfgpuJni.NFIELDS; f++)
                    unflatten(samples[m][f], field_manager2d.get(meshes[m], names[f]).getData());
            }
            num_samples = (int) ns[0];
            computeFields(); /* unchanged Java, KM:1606-1671 */
        }
        updateBoundaries(); /* Material.java:650-657 */
        gpu_first_time = false;
        first_time = false;
    }

    /** what ProcessBoundary does with a surface hit besides moving the particle (KM:586-602), for the hits the device processed in this step */
    private void applySurfaceHits() {
        long[] out2 = new long[2];
        check(SfgpuJni.takeSurfaceHits(ctx, sp, 0, null, null, null, null, out2), "take_surface_hits");
        int n = (int) out2[0];
        if (n == 0)
            return;
        int[] hm = new int[n], hs = new int[n];
        double[][] h = new double[5][n]; /* t, u, v, w, mpw */
        byte[] alive = new byte[n];
        long got = SfgpuJni.takeSurfaceHits(ctx, sp, n, hm, hs, h, alive, out2);
        check(got, "take_surface_hits");
        double[] vel = new double[3];
        for (int k = 0; k < got; k++) {
            Segment seg = devSeg[hm[k]][hs[k]];
            Boundary boundary = seg.getBoundary();
            double boundary_t = seg.id() + h[0][k]; /* KM:583 */
            vel[0] = h[1][k]; vel[1] = h[2][k]; vel[2] = h[3][k];
            if (devKind[hm[k]][hs[k]] == SfgpuJni.SURFACE_REMOVE)
                Starfish.source_module.boundary_charge += h[4][k] * charge; /* KM:589-591: only what performSurfaceInteraction removed, not the SINK */
            addSurfaceMomentum(boundary, boundary_t, vel, h[4][k]); /* KM:596 */
            if (alive[k] == 0)
                addSurfaceMassDeposit(boundary, boundary_t, h[4][k]); /* KM:598-602 */
        }
    }

    /** the unchanged Java surface handling for the particles the device handed back in their pre-ProcessBoundary state */
    private void finishSlowPath(double dt) {
        long[] counts = new long[3];
        check(SfgpuJni.getSums(ctx, sp, null, counts), "get_sums");
        int n = (int) counts[2];
        if (n == 0)
            return;
        double[][] a = soa(n, true);
        double[][] ex = new double[4][n];
        int[] id = new int[n], born = new int[n], bounces = new int[n], mesh = new int[n];
        long got = SfgpuJni.takeSlowpath(ctx, sp, n, a, id, born, ex, bounces, mesh);
        check(got, "take_slowpath");
        ArrayList<ArrayList<Particle>> survivors = new ArrayList<>();
        for (int m = 0; m < nMesh; m++)
            survivors.add(new ArrayList<Particle>());
        final int max_bounces = 10; /* KM:300 */
        double[] old = new double[3], old_lc = new double[2];
        for (int k = 0; k < got; k++) {
            Particle part = new Particle(this);
            part.pos[0] = a[0][k]; part.pos[1] = a[1][k]; part.pos[2] = a[2][k];
            part.vel[0] = a[3][k]; part.vel[1] = a[4][k]; part.vel[2] = a[5][k];
            part.mpw = a[6][k];
            part.lc = new double[] { a[7][k], a[8][k] };
            part.dt = a[9][k];
            part.id = id[k];
            part.born_it = born[k];
            Mesh m = meshes[mesh[k]];
            old[0] = ex[0][k]; old[1] = ex[1][k];
            old_lc[0] = ex[2][k]; old_lc[1] = ex[3][k];
            boolean alive = ProcessBoundary(part, m, old, old_lc); /* KM:471-750 */
            int b = bounces[k];
            while (alive && part.dt > 0 && b++ < max_bounces) { /* the remaining sub-steps of KM:360-398 */
                old[0] = part.pos[0]; old[1] = part.pos[1];
                old_lc[0] = part.lc[0]; old_lc[1] = part.lc[1];
                part.pos[0] += part.vel[0] * part.dt;
                part.pos[1] += part.vel[1] * part.dt;
                switch (Starfish.getDomainType()) {
                case RZ:
                    rotate(part, 0);
                    break;
                case ZR:
                    rotate(part, 1);
                    break;
                default:
                    part.pos[2] += part.vel[2] * part.dt;
                    break;
                }
                part.lc = m.XtoL(part.pos);
                alive = ProcessBoundary(part, m, old, old_lc);
            }
            if (alive)
                survivors.get(mesh[k]).add(part);
        }
        for (int m = 0; m < nMesh; m++) {
            ArrayList<Particle> s = survivors.get(m);
            if (s.isEmpty())
                continue;
            double[][] sa = soa(s.size(), true);
            int[] sid = new int[s.size()], sborn = new int[s.size()];
            for (int k = 0; k < s.size(); k++) {
                put(sa, k, s.get(k), true);
                sid[k] = s.get(k).id;
                sborn[k] = s.get(k).born_it;
            }
            check(SfgpuJni.inject(ctx, sp, m, s.size(), sa, sid, sborn, dt, SfgpuJni.INJECT_DEPOSIT_NOW), "inject(survivors)");
        }
    }

    /** rotateToRZ (r = 0) / rotateToZR (r = 1), KM:424-462 (private in ParticleMover) */
    private static void rotate(Particle part, int r) {
        double A = part.vel[2] * part.dt;
        double B = part.pos[r];
        double R = Math.sqrt(A * A + B * B);
        double cos = B / R;
        double sin = A / R;
        if (r == 0)
            part.pos[2] -= Math.asin(sin);
        else
            part.pos[2] += Math.acos(cos);
        part.pos[r] = R;
        double v1 = part.vel[r];
        double v2 = part.vel[2];
        part.vel[r] = cos * v1 + sin * v2;
        part.vel[2] = -sin * v1 + cos * v2;
    }

    @Override
    public void clearSamples() { /* KM:1509-1528 */
        super.clearSamples();
        if (ctx != 0)
            check(SfgpuJni.clearSamples(ctx, sp), "clear_samples");
    }

    @Override
    public long getNp() { /* KM:1297 */
        long n = SfgpuJni.np(ctx, sp, -1);
        for (ArrayList<Particle> l : pending)
            n += l.size();
        return n;
    }

    /**
     * getIterator(mesh), KM:271: a materialised view for output / diagnostics (SampleVDFModule.java:118, VTKWriter.java:1018-1027).
     * Consumers that mutate particles in place (MCC.java:167-216, DSMC.java:210-232) call writeBack() afterwards.
     */
    @Override
    public Iterator<Particle> getIterator(Mesh mesh) {
        return materialise(mesh.getIndex()).iterator();
    }

    private ArrayList<Particle> lastView;
    private int lastViewMesh = -1;

    private ArrayList<Particle> materialise(int m) {
        flushPending(m);
        int n = (int) SfgpuJni.np(ctx, sp, m);
        double[][] a = soa(n, true);
        int[] id = new int[n], born = new int[n];
        if (n > 0)
            check(SfgpuJni.download(ctx, sp, m, 0, n, a, id, born), "download");
        ArrayList<Particle> out = new ArrayList<>(n);
        for (int k = 0; k < n; k++) {
            Particle p = new Particle(this);
            p.pos[0] = a[0][k]; p.pos[1] = a[1][k]; p.pos[2] = a[2][k];
            p.vel[0] = a[3][k]; p.vel[1] = a[4][k]; p.vel[2] = a[5][k];
            p.mpw = a[6][k];
            p.lc = new double[] { a[7][k], a[8][k] };
            p.dt = a[9][k];
            p.id = id[k];
            p.born_it = born[k];
            out.add(p);
        }
        lastView = out;
        lastViewMesh = m;
        return out;
    }

    /** write the particles of the last getIterator(mesh) view back after a host-side mutation (collisions, chemistry) */
    public void writeBack() {
        if (lastView == null)
            return;
        int n = lastView.size();
        double[][] a = soa(n, true);
        int[] id = new int[n], born = new int[n];
        for (int k = 0; k < n; k++) {
            put(a, k, lastView.get(k), true);
            id[k] = lastView.get(k).id;
            born[k] = lastView.get(k).born_it;
        }
        if (n > 0)
            check(SfgpuJni.upload(ctx, sp, lastViewMesh, 0, n, a, id, born), "upload");
        lastView = null;
    }

    /** KM:904-941: num_samples, then per mesh the particle records (packed big-endian on the device) and the fields */
    @Override
    public void saveRestartData(java.io.DataOutputStream out) throws java.io.IOException {
        out.writeInt(num_samples);
        for (int m = 0; m < nMesh; m++) {
            flushPending(m);
            long need = SfgpuJni.restartSave(ctx, sp, m, null);
            check(need, "restart_save");
            byte[] buf = new byte[(int) need];
            check(SfgpuJni.restartSave(ctx, sp, m, buf), "restart_save");
            out.write(buf);
            Mesh mesh = meshes[m];
            getDen(mesh).binaryWrite(out); getDenAve(mesh).binaryWrite(out); getT(mesh).binaryWrite(out);
            getU(mesh).binaryWrite(out); getV(mesh).binaryWrite(out); getW(mesh).binaryWrite(out);
            getUAve(mesh).binaryWrite(out); getVAve(mesh).binaryWrite(out); getWAve(mesh).binaryWrite(out);
        }
    }

    public void close() {
        if (ctx != 0)
            SfgpuJni.destroy(ctx);
        ctx = 0;
    }
}
