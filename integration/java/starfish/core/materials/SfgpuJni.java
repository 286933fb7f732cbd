/*
 * SfgpuJni -- JNI face of libstarfish_gpu.so (include/sfgpu.h); the native side is integration/jni/sfgpu_jni.c.
 * JDK 11 compatible (the reference's CI / Docker images); on JDK 22+ the same C symbols bind through
 * java.lang.foreign without this adapter (INTEGRATION.md 2a).
 *
 * Conventions: every int-returning method returns 0 or a negative SFGPU_E* code (message: lastError);
 * field planes are page-locked direct buffers from hostAlloc(), laid out like double[ni][nj]: index i*nj + j;
 * particle arrays are plain Java arrays (pinned with GetPrimitiveArrayCritical for the duration of the call).
 */
package starfish.core.materials;

import java.nio.ByteBuffer;

final class SfgpuJni {
    static {
        System.loadLibrary("sfgpu_jni"); // links libstarfish_gpu.so
    }

    private SfgpuJni() {
    }

    /* sfgpu_step / sfgpu_inject flags, include/sfgpu.h */
    static final int INJECT_REWIND = 1, INJECT_DEPOSIT_NOW = 2, INJECT_TRANSFER = 4;
    static final int STEP_GENERIC = 1, STEP_DEFER_FINISH = 2, STEP_INPLACE = 4, STEP_STREAM = 8;
    static final int NFIELDS = 8;

    /** sfgpu_create: context handle, or 0 (see lastError(0)) */
    static native long create(int device, int domainType);

    static native void destroy(long ctx);

    static native String lastError(long ctx);

    /** sfgpu_host_alloc wrapped in a direct buffer (native byte order) */
    static native ByteBuffer hostAlloc(long bytes);

    static native void hostFree(ByteBuffer buffer);

    /** sfgpu_mesh_add: mesh id >= 0 or a negative error.  bc/nbr: RIGHT, TOP, LEFT, BOTTOM (Mesh.Face.val()). */
    static native int meshAdd(long ctx, int ni, int nj, double[] x0, double[] dh, byte[][] bc, int[][] nbr, byte[] hasSeg, double[] nodeVol);

    static native int setFields(long ctx, int mesh, ByteBuffer efi, ByteBuffer efj, ByteBuffer bfi, ByteBuffer bfj);

    /** sfgpu_species_add: species id >= 0 or a negative error */
    static native int speciesAdd(long ctx, double charge, double mass, long capacityHint);

    /** sfgpu_inject: soa = {x,y,z,u,v,w,mpw,li,lj,dt} (li, lj, dt entries may be null); accepted count or a negative error */
    static native long inject(long ctx, int sp, int mesh, int n, double[][] soa, int[] id, int[] bornIt, double dtStep, int flags);

    static native int step(long ctx, int sp, double dt, int flags);

    static native int finishStep(long ctx, int sp);

    static native int getMoments(long ctx, int sp, int mesh, ByteBuffer nd, ByteBuffer u, ByteBuffer v, ByteBuffer w);

    /** sfgpu_get_samples: out[0..7] = count,u,v,w,uu,vv,ww,mpc sums (entries may be null); numSamples[0] */
    static native int getSamples(long ctx, int sp, int mesh, ByteBuffer[] out, long[] numSamples);

    static native int clearSamples(long ctx, int sp);

    /** sfgpu_get_sums: sums5 = N,Px,Py,Pz,E; counts3 = np_alive, n_exited, n_slow */
    static native int getSums(long ctx, int sp, double[] sums5, long[] counts3);

    static native long np(long ctx, int sp, int mesh);

    /** sfgpu_take_slowpath: soa = {x,y,z,u,v,w,mpw,li,lj,dt}, extra = {old_x,old_y,old_li,old_lj}; returns the number copied */
    static native long takeSlowpath(long ctx, int sp, int max, double[][] soa, int[] id, int[] bornIt, double[][] extra, int[] bounces, int[] mesh);

    /** sfgpu_download / sfgpu_upload of particles [first, first+n) of a mesh */
    static native int download(long ctx, int sp, int mesh, long first, int n, double[][] soa, int[] id, int[] bornIt);

    static native int upload(long ctx, int sp, int mesh, long first, int n, double[][] soa, int[] id, int[] bornIt);

    /** sfgpu_restart_save into a byte array (null: size query); bytes needed / written or a negative error */
    static native long restartSave(long ctx, int sp, int mesh, byte[] buf);

    /** sfgpu_restart_load; out2 = {bytes_used, n_loaded} */
    static native int restartLoad(long ctx, int sp, int mesh, byte[] buf, double dtStep, long[] out2);

    static native int sort(long ctx, int sp);

    /* per-segment outcome of a surface hit, include/sfgpu.h SFGPU_SURFACE_* */
    static final int SURFACE_REMOVE = 0, SURFACE_NONE = 1, SURFACE_SPECULAR = 2;

    /** sfgpu_mesh_set_segments: xy = {x1, y1, x2, y2}; kind = SURFACE_*; sink != 0 for SINK boundaries; CSR of node.segments over nodes i*nj + j.  nSeg = 0 clears. */
    static native int meshSetSegments(long ctx, int mesh, int nSeg, double[][] xy, int[] kind, int[] sink, int[] nodeOffs, int[] nodeIds);

    /** sfgpu_take_surface_hits: tuvwm = {t, u, v, w, mpw}; out2 = {hits of the step, particles removed}; returns the number copied (max = 0: counts only) */
    static native long takeSurfaceHits(long ctx, int sp, int max, int[] mesh, int[] seg, double[][] tuvwm, byte[] alive, long[] out2);

    static native int setSortInterval(long ctx, int steps);

    /** sfgpu_set_tile_halo: 0 automatic, 1 / 2 fixed */
    static native int setTileHalo(long ctx, int halo);
}
