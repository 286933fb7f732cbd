// sf_multi.cuh -- single-caller multi-GPU entry (SURVEY 8b "sfgpu_create(cfg{n_gpus, device_ids[]})", 8e).
//
// Starfish's main loop is ONE thread (Starfish.java:77-121).  A group owns one context per GPU and one worker thread per
// context; every sfgpu_multi_* call fans out to the workers, which run the ordinary per-context entry points concurrently
// (their blocking synchronisations and the NCCL all-reduce of the deposit then overlap across GPUs instead of deadlocking a
// single thread), and returns when all of them are done.  Particles are partitioned by index, meshes and fields replicated,
// results read back once from rank 0 (after the all-reduce every rank holds the same deposit).
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

struct sfgpu_multi {
    std::vector<sfgpu_ctx *> ctx;
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::function<int(int)> job;
    uint64_t gen = 0;
    int pending = 0;
    bool quit = false;
    std::vector<int> rc;
    std::vector<int32_t> id_counter; // per species: part_id_counter of the whole population (KM:80), ids stay unique across GPUs
    std::string err;
};

static void multi_worker(sfgpu_multi *g, int r)
{
    uint64_t seen = 0;
    for (;;) {
        std::function<int(int)> f;
        {
            std::unique_lock<std::mutex> lk(g->mu);
            g->cv_job.wait(lk, [&] { return g->quit || g->gen != seen; });
            if (g->quit) return;
            seen = g->gen;
            f = g->job;
        }
        const int rc = f(r);
        {
            std::lock_guard<std::mutex> lk(g->mu);
            g->rc[r] = rc;
            if (--g->pending == 0) g->cv_done.notify_all();
        }
    }
}

// f(rank) on every worker at once; first failure wins, its message becomes the group's
static int multi_run(sfgpu_multi *g, std::function<int(int)> f)
{
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->job = std::move(f);
        g->pending = (int)g->workers.size();
        g->gen++;
    }
    g->cv_job.notify_all();
    std::unique_lock<std::mutex> lk(g->mu);
    g->cv_done.wait(lk, [&] { return g->pending == 0; });
    for (size_t r = 0; r < g->rc.size(); r++)
        if (g->rc[r]) {
            g->err = g->ctx[r] ? g->ctx[r]->err : std::string("rank failed before its context existed");
            g_last_error = g->err;
            return g->rc[r];
        }
    return 0;
}

extern "C" int sfgpu_multi_create(int32_t n, const int32_t *device_ids, int domain_type, sfgpu_multi **out)
{
    if (!out || n < 1 || n > 64) return fail(nullptr, SFGPU_EINVAL, "sfgpu_multi_create: bad arguments");
    *out = nullptr;
    sfgpu_multi *g = new (std::nothrow) sfgpu_multi();
    if (!g) return fail(nullptr, SFGPU_ENOMEM, "host allocation failed");
    g->ctx.assign(n, nullptr);
    g->rc.assign(n, 0);
    for (int r = 0; r < n; r++) g->workers.emplace_back(multi_worker, g, r);
    std::vector<std::string> errs(n);
    int rc = multi_run(g, [&](int r) {
        int e = sfgpu_create(device_ids ? device_ids[r] : r, domain_type, &g->ctx[r]);
        if (e) errs[r] = g_last_error; // (thread local)
        return e;
    });
    if (rc) {
        for (auto &s : errs)
            if (!s.empty()) g->err = s;
    }
    if (!rc && n > 1) {
        char id[128];
        rc = sfgpu_comm_unique_id(id);
        if (rc) g->err = g_last_error;
        if (!rc) rc = multi_run(g, [&](int r) { return sfgpu_comm_init(g->ctx[r], n, r, id); }); // concurrent: ncclCommInitRank meets its peers
    }
    if (rc) {
        const std::string why = g->err;
        extern void sfgpu_multi_destroy(sfgpu_multi *);
        sfgpu_multi_destroy(g);
        return fail(nullptr, rc, "sfgpu_multi_create: %s", why.c_str());
    }
    *out = g;
    return 0;
}

extern "C" void sfgpu_multi_destroy(sfgpu_multi *g)
{
    if (!g) return;
    multi_run(g, [&](int r) {
        if (g->ctx[r]) sfgpu_destroy(g->ctx[r]);
        g->ctx[r] = nullptr;
        return 0;
    });
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->quit = true;
    }
    g->cv_job.notify_all();
    for (auto &t : g->workers) t.join();
    delete g;
}

extern "C" int32_t sfgpu_multi_size(sfgpu_multi *g) { return g ? (int32_t)g->ctx.size() : 0; }
extern "C" sfgpu_ctx *sfgpu_multi_ctx(sfgpu_multi *g, int32_t rank) { return (g && rank >= 0 && rank < (int)g->ctx.size()) ? g->ctx[rank] : nullptr; }
extern "C" const char *sfgpu_multi_last_error(sfgpu_multi *g) { return g ? g->err.c_str() : g_last_error.c_str(); }

extern "C" int sfgpu_multi_mesh_add(sfgpu_multi *g, int32_t ni, int32_t nj, const double x0[2], const double dh[2], const int8_t *const bc[4],
                                    const int32_t *const nbr[4], const uint8_t *has_seg, const double *node_vol, int32_t *mesh_id)
{
    if (!g || !mesh_id) return fail(nullptr, SFGPU_EINVAL, "null group");
    std::vector<int32_t> ids(g->ctx.size(), -1);
    int rc = multi_run(g, [&](int r) { return sfgpu_mesh_add(g->ctx[r], ni, nj, x0, dh, bc, nbr, has_seg, node_vol, &ids[r]); });
    *mesh_id = ids[0];
    return rc;
}

extern "C" int sfgpu_multi_mesh_set_segments(sfgpu_multi *g, int32_t mesh_id, int32_t n_seg, const double *x1, const double *y1, const double *x2,
                                             const double *y2, const int32_t *kind, const int32_t *sink, const int32_t *node_offs, const int32_t *node_ids)
{
    if (!g) return fail(nullptr, SFGPU_EINVAL, "null group");
    return multi_run(g, [&](int r) { return sfgpu_mesh_set_segments(g->ctx[r], mesh_id, n_seg, x1, y1, x2, y2, kind, sink, node_offs, node_ids); });
}

extern "C" int sfgpu_multi_set_fields(sfgpu_multi *g, int32_t mesh_id, const double *efi, const double *efj, const double *bfi, const double *bfj)
{
    if (!g) return fail(nullptr, SFGPU_EINVAL, "null group");
    return multi_run(g, [&](int r) { return sfgpu_set_fields(g->ctx[r], mesh_id, efi, efj, bfi, bfj); });
}

extern "C" int sfgpu_multi_species_add(sfgpu_multi *g, double charge, double mass, int64_t capacity_hint, int32_t *sp)
{
    if (!g || !sp) return fail(nullptr, SFGPU_EINVAL, "null group");
    const int n = (int)g->ctx.size();
    std::vector<int32_t> ids(n, -1);
    int rc = multi_run(g, [&](int r) { return sfgpu_species_add(g->ctx[r], charge, mass, (capacity_hint + n - 1) / n, &ids[r]); });
    *sp = ids[0];
    if (!rc && (int)g->id_counter.size() <= *sp) g->id_counter.resize(*sp + 1, 0);
    return rc;
}

// particles [first_r, first_r + count_r) of the batch go to rank r (contiguous, balanced); ids are assigned here when the caller
// leaves them to the library, so that they stay unique over the whole population
extern "C" int sfgpu_multi_inject(sfgpu_multi *g, int32_t sp, int32_t mesh_id, const sfgpu_particles *p, double dt_step, uint32_t flags, int64_t *n_added)
{
    if (!g || !p) return fail(nullptr, SFGPU_EINVAL, "null group / particles");
    if (sp < 0 || sp >= (int)g->id_counter.size()) return fail(nullptr, SFGPU_EINVAL, "bad species id %d", sp);
    const int n = (int)g->ctx.size();
    std::vector<int32_t> ids;
    const int32_t *idp = p->id;
    if (!idp) {
        ids.resize((size_t)p->n);
        for (int64_t k = 0; k < p->n; k++) ids[(size_t)k] = g->id_counter[sp] + (int32_t)k;
        g->id_counter[sp] += (int32_t)p->n;
        idp = ids.data();
    }
    std::vector<int64_t> added(n, 0);
    int rc = multi_run(g, [&](int r) {
        const int64_t base = p->n / n, extra = p->n % n;
        const int64_t first = r * base + (r < extra ? r : extra), count = base + (r < extra ? 1 : 0);
        if (count == 0) return 0;
        sfgpu_particles q = *p;
        q.n = count;
        double **src[10] = {&q.x, &q.y, &q.z, &q.u, &q.v, &q.w, &q.mpw, &q.li, &q.lj, &q.dt};
        for (auto a : src)
            if (*a) *a += first;
        q.id = const_cast<int32_t *>(idp) + first;
        if (q.born_it) q.born_it += first;
        return sfgpu_inject(g->ctx[r], sp, mesh_id, &q, dt_step, flags, &added[r]);
    });
    if (n_added) {
        *n_added = 0;
        for (auto a : added) *n_added += a;
    }
    return rc;
}

// KineticMaterial.updateFields() on all GPUs at once; the deposit and the mover sums are all-reduced inside (NCCL)
extern "C" int sfgpu_multi_step(sfgpu_multi *g, int32_t sp, double dt, uint32_t flags)
{
    if (!g) return fail(nullptr, SFGPU_EINVAL, "null group");
    return multi_run(g, [&](int r) { return sfgpu_step(g->ctx[r], sp, dt, flags); });
}
extern "C" int sfgpu_multi_finish_step(sfgpu_multi *g, int32_t sp)
{
    if (!g) return fail(nullptr, SFGPU_EINVAL, "null group");
    return multi_run(g, [&](int r) { return sfgpu_finish_step(g->ctx[r], sp); });
}

// results: once, from rank 0 (every rank holds the same summed deposit and running sums)
extern "C" int sfgpu_multi_get_moments(sfgpu_multi *g, int32_t sp, int32_t mesh_id, double *nd, double *u, double *v, double *w)
{
    if (!g) return fail(nullptr, SFGPU_EINVAL, "null group");
    return multi_run(g, [&](int r) { return r == 0 ? sfgpu_get_moments(g->ctx[0], sp, mesh_id, nd, u, v, w) : 0; });
}
extern "C" int sfgpu_multi_get_deposit(sfgpu_multi *g, int32_t sp, int32_t mesh_id, double *const out[SFGPU_NFIELDS])
{
    if (!g) return fail(nullptr, SFGPU_EINVAL, "null group");
    return multi_run(g, [&](int r) { return r == 0 ? sfgpu_get_deposit(g->ctx[0], sp, mesh_id, out) : 0; });
}
extern "C" int sfgpu_multi_get_samples(sfgpu_multi *g, int32_t sp, int32_t mesh_id, double *const out[SFGPU_NFIELDS], int64_t *num_samples)
{
    if (!g) return fail(nullptr, SFGPU_EINVAL, "null group");
    return multi_run(g, [&](int r) { return r == 0 ? sfgpu_get_samples(g->ctx[0], sp, mesh_id, out, num_samples) : 0; });
}
extern "C" int sfgpu_multi_clear_samples(sfgpu_multi *g, int32_t sp)
{
    if (!g) return fail(nullptr, SFGPU_EINVAL, "null group");
    return multi_run(g, [&](int r) { return sfgpu_clear_samples(g->ctx[r], sp); });
}
// sums5: the all-reduced mover sums; counts: summed over the GPUs
extern "C" int sfgpu_multi_get_sums(sfgpu_multi *g, int32_t sp, double sums5[5], int64_t *np_alive, int64_t *n_exited, int64_t *n_slow)
{
    if (!g) return fail(nullptr, SFGPU_EINVAL, "null group");
    const int n = (int)g->ctx.size();
    std::vector<int64_t> a(n, 0), b(n, 0), c(n, 0);
    int rc = multi_run(g, [&](int r) { return sfgpu_get_sums(g->ctx[r], sp, r == 0 ? sums5 : nullptr, &a[r], &b[r], &c[r]); });
    int64_t sa = 0, sb = 0, sc = 0;
    for (int r = 0; r < n; r++) { sa += a[r]; sb += b[r]; sc += c[r]; }
    if (np_alive) *np_alive = sa;
    if (n_exited) *n_exited = sb;
    if (n_slow) *n_slow = sc;
    return rc;
}
