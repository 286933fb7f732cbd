/*
 * sfgpu_jni.c -- JNI adapter between starfish.core.materials.SfgpuJni and the C ABI of include/sfgpu.h.
 * One function per native method; no state, no copies beyond what JNI itself needs (direct buffers for the
 * page-locked field planes, GetPrimitiveArrayCritical for the particle arrays).
 *
 *   gcc -shared -fPIC -O2 -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../include sfgpu_jni.c \
 *       -L../../starfish_b200 -l:libstarfish_gpu.so -o libsfgpu_jni.so
 */
#include <jni.h>
#include <stdint.h>
#include <string.h>

#include "sfgpu.h"

#define CTX(h) ((sfgpu_ctx *)(intptr_t)(h))
#define FN(name) Java_starfish_core_materials_SfgpuJni_##name

static void *direct(JNIEnv *e, jobject buf) { return buf ? (*e)->GetDirectBufferAddress(e, buf) : NULL; }

JNIEXPORT jlong JNICALL FN(create)(JNIEnv *e, jclass c, jint device, jint domain)
{
    sfgpu_ctx *ctx = NULL;
    return sfgpu_create(device, domain, &ctx) == SFGPU_OK ? (jlong)(intptr_t)ctx : 0;
}

JNIEXPORT void JNICALL FN(destroy)(JNIEnv *e, jclass c, jlong ctx) { sfgpu_destroy(CTX(ctx)); }

JNIEXPORT jstring JNICALL FN(lastError)(JNIEnv *e, jclass c, jlong ctx) { return (*e)->NewStringUTF(e, sfgpu_last_error(CTX(ctx))); }

JNIEXPORT jobject JNICALL FN(hostAlloc)(JNIEnv *e, jclass c, jlong bytes)
{
    void *p = NULL;
    if (sfgpu_host_alloc((size_t)bytes, &p) != SFGPU_OK) return NULL;
    return (*e)->NewDirectByteBuffer(e, p, bytes);
}

JNIEXPORT void JNICALL FN(hostFree)(JNIEnv *e, jclass c, jobject buf) { sfgpu_host_free(direct(e, buf)); }

JNIEXPORT jint JNICALL FN(meshAdd)(JNIEnv *e, jclass c, jlong ctx, jint ni, jint nj, jdoubleArray x0, jdoubleArray dh, jobjectArray bc,
                                   jobjectArray nbr, jbyteArray hasSeg, jdoubleArray nodeVol)
{
    jdouble x0v[2], dhv[2];
    (*e)->GetDoubleArrayRegion(e, x0, 0, 2, x0v);
    (*e)->GetDoubleArrayRegion(e, dh, 0, 2, dhv);
    jbyteArray bca[4];
    jintArray nba[4];
    const int8_t *bcp[4];
    const int32_t *nbp[4];
    for (int f = 0; f < 4; f++) {
        bca[f] = (jbyteArray)(*e)->GetObjectArrayElement(e, bc, f);
        nba[f] = nbr ? (jintArray)(*e)->GetObjectArrayElement(e, nbr, f) : NULL;
        bcp[f] = (const int8_t *)(*e)->GetByteArrayElements(e, bca[f], NULL);
        nbp[f] = nba[f] ? (const int32_t *)(*e)->GetIntArrayElements(e, nba[f], NULL) : NULL;
    }
    jbyte *seg = hasSeg ? (*e)->GetByteArrayElements(e, hasSeg, NULL) : NULL;
    jdouble *vol = nodeVol ? (*e)->GetDoubleArrayElements(e, nodeVol, NULL) : NULL;
    int32_t id = -1;
    int rc = sfgpu_mesh_add(CTX(ctx), ni, nj, x0v, dhv, bcp, nbp, (const uint8_t *)seg, vol, &id);
    for (int f = 0; f < 4; f++) {
        (*e)->ReleaseByteArrayElements(e, bca[f], (jbyte *)bcp[f], JNI_ABORT);
        if (nba[f]) (*e)->ReleaseIntArrayElements(e, nba[f], (jint *)nbp[f], JNI_ABORT);
    }
    if (seg) (*e)->ReleaseByteArrayElements(e, hasSeg, seg, JNI_ABORT);
    if (vol) (*e)->ReleaseDoubleArrayElements(e, nodeVol, vol, JNI_ABORT);
    return rc == SFGPU_OK ? id : rc;
}

JNIEXPORT jint JNICALL FN(setFields)(JNIEnv *e, jclass c, jlong ctx, jint mesh, jobject efi, jobject efj, jobject bfi, jobject bfj)
{
    return sfgpu_set_fields(CTX(ctx), mesh, direct(e, efi), direct(e, efj), direct(e, bfi), direct(e, bfj));
}

JNIEXPORT jint JNICALL FN(speciesAdd)(JNIEnv *e, jclass c, jlong ctx, jdouble charge, jdouble mass, jlong hint)
{
    int32_t sp = -1;
    int rc = sfgpu_species_add(CTX(ctx), charge, mass, hint, &sp);
    return rc == SFGPU_OK ? sp : rc;
}

/* pins soa[0..9] (null rows stay NULL) + id + born_it into an sfgpu_particles view */
typedef struct {
    jdoubleArray rows[10];
    jdouble *ptr[10];
    jint *id, *born;
} pinned_particles;

static void pin(JNIEnv *e, jobjectArray soa, jintArray id, jintArray born, jlong n, pinned_particles *pp, sfgpu_particles *p)
{
    memset(pp, 0, sizeof *pp);
    memset(p, 0, sizeof *p);
    for (int k = 0; k < 10; k++) {
        pp->rows[k] = (jdoubleArray)(*e)->GetObjectArrayElement(e, soa, k);
        pp->ptr[k] = pp->rows[k] ? (*e)->GetDoubleArrayElements(e, pp->rows[k], NULL) : NULL;
    }
    pp->id = id ? (*e)->GetIntArrayElements(e, id, NULL) : NULL;
    pp->born = born ? (*e)->GetIntArrayElements(e, born, NULL) : NULL;
    p->n = n;
    p->x = pp->ptr[0]; p->y = pp->ptr[1]; p->z = pp->ptr[2]; p->u = pp->ptr[3]; p->v = pp->ptr[4]; p->w = pp->ptr[5];
    p->mpw = pp->ptr[6]; p->li = pp->ptr[7]; p->lj = pp->ptr[8]; p->dt = pp->ptr[9];
    p->id = (int32_t *)pp->id;
    p->born_it = (int32_t *)pp->born;
}

static void unpin(JNIEnv *e, jintArray id, jintArray born, pinned_particles *pp, jint mode)
{
    for (int k = 0; k < 10; k++)
        if (pp->rows[k]) (*e)->ReleaseDoubleArrayElements(e, pp->rows[k], pp->ptr[k], mode);
    if (pp->id) (*e)->ReleaseIntArrayElements(e, id, pp->id, mode);
    if (pp->born) (*e)->ReleaseIntArrayElements(e, born, pp->born, mode);
}

JNIEXPORT jlong JNICALL FN(inject)(JNIEnv *e, jclass c, jlong ctx, jint sp, jint mesh, jint n, jobjectArray soa, jintArray id, jintArray born,
                                   jdouble dtStep, jint flags)
{
    pinned_particles pp;
    sfgpu_particles p;
    pin(e, soa, id, born, n, &pp, &p);
    int64_t added = 0;
    int rc = sfgpu_inject(CTX(ctx), sp, mesh, &p, dtStep, (uint32_t)flags, &added);
    unpin(e, id, born, &pp, JNI_ABORT);
    return rc == SFGPU_OK ? added : rc;
}

JNIEXPORT jint JNICALL FN(step)(JNIEnv *e, jclass c, jlong ctx, jint sp, jdouble dt, jint flags) { return sfgpu_step(CTX(ctx), sp, dt, (uint32_t)flags); }

JNIEXPORT jint JNICALL FN(finishStep)(JNIEnv *e, jclass c, jlong ctx, jint sp) { return sfgpu_finish_step(CTX(ctx), sp); }

JNIEXPORT jint JNICALL FN(getMoments)(JNIEnv *e, jclass c, jlong ctx, jint sp, jint mesh, jobject nd, jobject u, jobject v, jobject w)
{
    return sfgpu_get_moments(CTX(ctx), sp, mesh, direct(e, nd), direct(e, u), direct(e, v), direct(e, w));
}

JNIEXPORT jint JNICALL FN(getSamples)(JNIEnv *e, jclass c, jlong ctx, jint sp, jint mesh, jobjectArray out, jlongArray numSamples)
{
    double *planes[SFGPU_NFIELDS];
    for (int f = 0; f < SFGPU_NFIELDS; f++) planes[f] = out ? direct(e, (*e)->GetObjectArrayElement(e, out, f)) : NULL;
    int64_t ns = 0;
    int rc = sfgpu_get_samples(CTX(ctx), sp, mesh, out ? planes : NULL, &ns);
    if (numSamples) {
        jlong v = ns;
        (*e)->SetLongArrayRegion(e, numSamples, 0, 1, &v);
    }
    return rc;
}

JNIEXPORT jint JNICALL FN(clearSamples)(JNIEnv *e, jclass c, jlong ctx, jint sp) { return sfgpu_clear_samples(CTX(ctx), sp); }

JNIEXPORT jint JNICALL FN(getSums)(JNIEnv *e, jclass c, jlong ctx, jint sp, jdoubleArray sums5, jlongArray counts3)
{
    double s[5];
    int64_t np = 0, nx = 0, ns = 0;
    int rc = sfgpu_get_sums(CTX(ctx), sp, s, &np, &nx, &ns);
    if (sums5) (*e)->SetDoubleArrayRegion(e, sums5, 0, 5, s);
    if (counts3) {
        jlong v[3] = {np, nx, ns};
        (*e)->SetLongArrayRegion(e, counts3, 0, 3, v);
    }
    return rc;
}

JNIEXPORT jlong JNICALL FN(np)(JNIEnv *e, jclass c, jlong ctx, jint sp, jint mesh)
{
    int64_t n = 0;
    int rc = sfgpu_np(CTX(ctx), sp, mesh, &n);
    return rc == SFGPU_OK ? n : rc;
}

JNIEXPORT jlong JNICALL FN(takeSlowpath)(JNIEnv *e, jclass c, jlong ctx, jint sp, jint max, jobjectArray soa, jintArray id, jintArray born,
                                         jobjectArray extra, jintArray bounces, jintArray mesh)
{
    pinned_particles pp;
    sfgpu_particles p;
    pin(e, soa, id, born, max, &pp, &p);
    jdoubleArray er[4];
    jdouble *ep[4];
    for (int k = 0; k < 4; k++) {
        er[k] = (jdoubleArray)(*e)->GetObjectArrayElement(e, extra, k);
        ep[k] = (*e)->GetDoubleArrayElements(e, er[k], NULL);
    }
    jint *bp = (*e)->GetIntArrayElements(e, bounces, NULL), *mp = (*e)->GetIntArrayElements(e, mesh, NULL);
    sfgpu_slow_extra x;
    x.old_x = ep[0]; x.old_y = ep[1]; x.old_li = ep[2]; x.old_lj = ep[3];
    x.bounces = (int32_t *)bp;
    x.mesh = (int32_t *)mp;
    int64_t n = 0;
    int rc = sfgpu_take_slowpath(CTX(ctx), sp, max, &p, &x, &n);
    for (int k = 0; k < 4; k++) (*e)->ReleaseDoubleArrayElements(e, er[k], ep[k], 0);
    (*e)->ReleaseIntArrayElements(e, bounces, bp, 0);
    (*e)->ReleaseIntArrayElements(e, mesh, mp, 0);
    unpin(e, id, born, &pp, 0);
    return rc == SFGPU_OK ? n : rc;
}

JNIEXPORT jint JNICALL FN(download)(JNIEnv *e, jclass c, jlong ctx, jint sp, jint mesh, jlong first, jint n, jobjectArray soa, jintArray id, jintArray born)
{
    pinned_particles pp;
    sfgpu_particles p;
    pin(e, soa, id, born, n, &pp, &p);
    int rc = sfgpu_download(CTX(ctx), sp, mesh, first, &p);
    unpin(e, id, born, &pp, 0);
    return rc;
}

JNIEXPORT jint JNICALL FN(upload)(JNIEnv *e, jclass c, jlong ctx, jint sp, jint mesh, jlong first, jint n, jobjectArray soa, jintArray id, jintArray born)
{
    pinned_particles pp;
    sfgpu_particles p;
    pin(e, soa, id, born, n, &pp, &p);
    int rc = sfgpu_upload(CTX(ctx), sp, mesh, first, &p);
    unpin(e, id, born, &pp, JNI_ABORT);
    return rc;
}

JNIEXPORT jlong JNICALL FN(restartSave)(JNIEnv *e, jclass c, jlong ctx, jint sp, jint mesh, jbyteArray buf)
{
    int64_t need = 0;
    jbyte *p = buf ? (*e)->GetByteArrayElements(e, buf, NULL) : NULL;
    int rc = sfgpu_restart_save(CTX(ctx), sp, mesh, p, buf ? (*e)->GetArrayLength(e, buf) : 0, &need);
    if (p) (*e)->ReleaseByteArrayElements(e, buf, p, 0);
    return rc == SFGPU_OK ? need : rc;
}

JNIEXPORT jint JNICALL FN(restartLoad)(JNIEnv *e, jclass c, jlong ctx, jint sp, jint mesh, jbyteArray buf, jdouble dtStep, jlongArray out2)
{
    int64_t used = 0, loaded = 0;
    jbyte *p = (*e)->GetByteArrayElements(e, buf, NULL);
    int rc = sfgpu_restart_load(CTX(ctx), sp, mesh, p, (*e)->GetArrayLength(e, buf), dtStep, &used, &loaded);
    (*e)->ReleaseByteArrayElements(e, buf, p, JNI_ABORT);
    if (out2) {
        jlong v[2] = {used, loaded};
        (*e)->SetLongArrayRegion(e, out2, 0, 2, v);
    }
    return rc;
}

JNIEXPORT jint JNICALL FN(sort)(JNIEnv *e, jclass c, jlong ctx, jint sp) { return sfgpu_sort(CTX(ctx), sp); }

/* sfgpu_mesh_set_segments: xy = {x1, y1, x2, y2} (one double[] each), kind / sink per segment, CSR of node.segments over nodes i*nj + j */
JNIEXPORT jint JNICALL FN(meshSetSegments)(JNIEnv *e, jclass c, jlong ctx, jint mesh, jint nSeg, jobjectArray xy, jintArray kind, jintArray sink,
                                           jintArray nodeOffs, jintArray nodeIds)
{
    if (nSeg == 0) return sfgpu_mesh_set_segments(CTX(ctx), mesh, 0, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL);
    jdoubleArray a[4];
    jdouble *p[4];
    for (int k = 0; k < 4; k++) {
        a[k] = (jdoubleArray)(*e)->GetObjectArrayElement(e, xy, k);
        p[k] = (*e)->GetDoubleArrayElements(e, a[k], NULL);
    }
    jint *kd = (*e)->GetIntArrayElements(e, kind, NULL);
    jint *sk = sink ? (*e)->GetIntArrayElements(e, sink, NULL) : NULL;
    jint *no = (*e)->GetIntArrayElements(e, nodeOffs, NULL);
    jint *nd = nodeIds ? (*e)->GetIntArrayElements(e, nodeIds, NULL) : NULL;
    int rc = sfgpu_mesh_set_segments(CTX(ctx), mesh, nSeg, p[0], p[1], p[2], p[3], (const int32_t *)kd, (const int32_t *)sk, (const int32_t *)no, (const int32_t *)nd);
    for (int k = 0; k < 4; k++) (*e)->ReleaseDoubleArrayElements(e, a[k], p[k], JNI_ABORT);
    (*e)->ReleaseIntArrayElements(e, kind, kd, JNI_ABORT);
    if (sk) (*e)->ReleaseIntArrayElements(e, sink, sk, JNI_ABORT);
    (*e)->ReleaseIntArrayElements(e, nodeOffs, no, JNI_ABORT);
    if (nd) (*e)->ReleaseIntArrayElements(e, nodeIds, nd, JNI_ABORT);
    return rc;
}

/* sfgpu_take_surface_hits: tuvwm = {t, u, v, w, mpw}; out2 = {hits of the step, particles the surfaces removed}; returns the number copied */
JNIEXPORT jlong JNICALL FN(takeSurfaceHits)(JNIEnv *e, jclass c, jlong ctx, jint sp, jint max, jintArray mesh, jintArray seg, jobjectArray tuvwm,
                                            jbyteArray alive, jlongArray out2)
{
    int64_t n = 0, n_abs = 0;
    int rc;
    if (max <= 0 || !tuvwm) {
        rc = sfgpu_take_surface_hits(CTX(ctx), sp, 0, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, &n, &n_abs);
    } else {
        jdoubleArray a[5];
        jdouble *p[5];
        for (int k = 0; k < 5; k++) {
            a[k] = (jdoubleArray)(*e)->GetObjectArrayElement(e, tuvwm, k);
            p[k] = (*e)->GetDoubleArrayElements(e, a[k], NULL);
        }
        jint *ms = (*e)->GetIntArrayElements(e, mesh, NULL);
        jint *sg = (*e)->GetIntArrayElements(e, seg, NULL);
        jbyte *al = (*e)->GetByteArrayElements(e, alive, NULL);
        rc = sfgpu_take_surface_hits(CTX(ctx), sp, max, (int32_t *)ms, (int32_t *)sg, p[0], p[1], p[2], p[3], p[4], (int8_t *)al, &n, &n_abs);
        for (int k = 0; k < 5; k++) (*e)->ReleaseDoubleArrayElements(e, a[k], p[k], 0);
        (*e)->ReleaseIntArrayElements(e, mesh, ms, 0);
        (*e)->ReleaseIntArrayElements(e, seg, sg, 0);
        (*e)->ReleaseByteArrayElements(e, alive, al, 0);
    }
    if (out2) {
        jlong v[2] = {n, n_abs};
        (*e)->SetLongArrayRegion(e, out2, 0, 2, v);
    }
    if (rc != SFGPU_OK) return rc;
    return n < max ? n : max;
}

JNIEXPORT jint JNICALL FN(setSortInterval)(JNIEnv *e, jclass c, jlong ctx, jint steps) { return sfgpu_set_sort_interval(CTX(ctx), steps); }

JNIEXPORT jint JNICALL FN(setTileHalo)(JNIEnv *e, jclass c, jlong ctx, jint halo) { return sfgpu_set_tile_halo(CTX(ctx), halo); }
