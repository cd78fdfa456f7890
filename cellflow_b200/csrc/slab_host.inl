// slab_host.inl — host side of the multi-GPU slab decomposition (included by engine.cu).
// One process per GPU; ring neighbours exchange migrants and one ghost layer per step with
// ncclSend/ncclRecv (or plain device copies when world == 1, which exercises the same code
// path on a single GPU).

static void slab_free(cf_sim* s) {
    for (int d = 0; d < 2; d++) {
        cudaFree(s->send_mig[d]);
        cudaFree(s->recv_mig[d]);
        cudaFree(s->send_halo[d]);
        cudaFree(s->recv_halo[d]);
        cudaFree(s->akeys[d]);
        cudaFree(s->avals[d]);
        cudaFree(s->gkeys[d]);
        s->send_mig[d] = s->recv_mig[d] = s->send_halo[d] = s->recv_halo[d] = nullptr;
        s->akeys[d] = s->avals[d] = s->gkeys[d] = nullptr;
    }
    cudaFree(s->d_slab_counts);
    s->d_slab_counts = nullptr;
    if (s->h_slab_counts) cudaFreeHost(s->h_slab_counts);
    s->h_slab_counts = nullptr;
    if (s->comm && nccl_api().ok) nccl_api().CommDestroy(s->comm);
    s->comm = nullptr;
}

static float slab_bound(const cf_sim* s, int r) {
    return (float)((double)s->params.canvasWidth * (double)r / (double)s->world);
}

#define NCCLCHK(call)                                                                         \
    do {                                                                                      \
        ncclResult_t r_ = (call);                                                             \
        if (r_ != 0)                                                                          \
            return fail(CF_ERR_NCCL, "%s failed: %s", #call, nccl_api().GetErrorString(r_)); \
    } while (0)

extern "C" int cf_nccl_unique_id(void* id128) {
    ARG(id128);
    NcclApi& api = nccl_api();
    if (!api.ok) return fail(CF_ERR_NCCL, "NCCL unavailable: %s", api.why);
    ncclUniqueId id;
    NCCLCHK(api.GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return CF_OK;
}

extern "C" int cf_comm_init(cf_sim* s, int rank, int world, const void* id128, int capacity) {
    ARG(s && world >= 1 && rank >= 0 && rank < world && capacity >= 1);
    if (int rc = set_device(s)) return rc;
    CU(cudaStreamSynchronize(s->stream));
    slab_free(s);
    if (world > 1) {
        ARG(id128 != nullptr);
        NcclApi& api = nccl_api();
        if (!api.ok) return fail(CF_ERR_NCCL, "NCCL unavailable: %s", api.why);
        ncclUniqueId id;
        memcpy(&id, id128, 128);
        NCCLCHK(api.CommInitRank(&s->comm, world, id, rank));
    }
    s->rank = rank;
    s->world = world;
    s->slab = true;
    s->cap_own = capacity;
    if (s->cap_halo <= 0) s->cap_halo = std::max(32768, capacity / 6);
    if (s->cap_mig <= 0) s->cap_mig = std::max(16384, capacity / 20);
    s->n = 0;
    if (int rc = alloc_particle_buffers(s, s->cap_own + 2 * s->cap_halo)) return rc;
    s->base = s->cap_halo;
    for (int d = 0; d < 2; d++) {
        CU(cudaMalloc(&s->send_mig[d], slab_mig_bytes(s->cap_mig)));
        CU(cudaMalloc(&s->recv_mig[d], slab_mig_bytes(s->cap_mig)));
        CU(cudaMalloc(&s->send_halo[d], slab_halo_bytes(s->cap_halo)));
        CU(cudaMalloc(&s->recv_halo[d], slab_halo_bytes(s->cap_halo)));
        CU(cudaMalloc(&s->akeys[d], sizeof(uint32_t) * 2 * (size_t)s->cap_mig));
        CU(cudaMalloc(&s->avals[d], sizeof(uint32_t) * 2 * (size_t)s->cap_mig));
        CU(cudaMalloc(&s->gkeys[d], sizeof(uint32_t) * (size_t)s->cap_halo));
        CU(cudaMemsetAsync(s->recv_mig[d], 0, slab_mig_bytes(s->cap_mig), s->stream));
        CU(cudaMemsetAsync(s->recv_halo[d], 0, slab_halo_bytes(s->cap_halo), s->stream));
    }
    CU(cudaMalloc(&s->d_slab_counts, 8 * sizeof(int)));
    CU(cudaMemsetAsync(s->d_slab_counts, 0, 8 * sizeof(int), s->stream));
    CU(cudaMallocHost(&s->h_slab_counts, 8 * sizeof(int)));
    CU(cudaStreamSynchronize(s->stream));
    return CF_OK;
}

// Ring exchange of two fixed-size messages.  What I send right arrives at my right neighbour
// "from left", and vice versa.  Receives are posted right-then-left so that with world == 2
// (both neighbours are the same peer) the peer's in-order sends (left, right) match.
static int slab_exchange(cf_sim* s, char* send_left, char* send_right, char* recv_from_left,
                         char* recv_from_right, size_t bytes) {
    if (s->world == 1) {
        CU(cudaMemcpyAsync(recv_from_left, send_right, bytes, cudaMemcpyDeviceToDevice, s->stream));
        CU(cudaMemcpyAsync(recv_from_right, send_left, bytes, cudaMemcpyDeviceToDevice, s->stream));
        return 0;
    }
    NcclApi& api = nccl_api();
    int left = (s->rank - 1 + s->world) % s->world, right = (s->rank + 1) % s->world;
    NCCLCHK(api.GroupStart());
    NCCLCHK(api.Send(send_left, bytes, ncclChar_, left, s->comm, s->stream));
    NCCLCHK(api.Send(send_right, bytes, ncclChar_, right, s->comm, s->stream));
    NCCLCHK(api.Recv(recv_from_right, bytes, ncclChar_, right, s->comm, s->stream));
    NCCLCHK(api.Recv(recv_from_left, bytes, ncclChar_, left, s->comm, s->stream));
    NCCLCHK(api.GroupEnd());
    return 0;
}

// Elements actually shipped per message.  Buffers are allocated for cap_halo / cap_mig, but a
// fixed-size message of that capacity would move ~30 MB per step and rank; both ends of a link
// instead derive the same tighter bound from numbers every rank knows (global count, world size,
// owned layers): 2x the mean ghost layer + 8192, mean/64 + 16384 migrants.  A rank that would
// exceed it fails with CF_ERR_CAPACITY (raise it with the halo_slack / migrant_slack options).
static int slab_halo_msg_cap(const cf_sim* s) {
    if (s->n_total <= 0) return s->cap_halo; // global count unknown: full (identical) capacity
    double mean = (double)s->n_total / s->world / std::max(s->nxl, 1);
    long long cap = (long long)(s->halo_slack * mean) + 8192;
    return (int)std::min<long long>(cap, s->cap_halo);
}
static int slab_mig_msg_cap(const cf_sim* s) {
    if (s->n_total <= 0) return s->cap_mig;
    double mean = (double)s->n_total / s->world;
    long long cap = (long long)(s->mig_slack * mean / 64.0) + 16384;
    return (int)std::min<long long>(cap, s->cap_mig);
}

static int slab_check_flags(cf_sim* s) {
    int f = s->h_slab_counts[3];
    if (f == 1) return fail(CF_ERR_STATE, "a particle moved further than one slab width in one step");
    if (f == 2)
        return fail(CF_ERR_CAPACITY, "a ghost layer exceeded the halo message capacity (%d): raise halo_slack",
                    slab_halo_msg_cap(s));
    return 0;
}

// Host-side wall-clock breakdown of the slab cell-list build (CF_SLAB_DEBUG=1 prints it at
// cf_destroy): where a rank waits — its own GPU (sync after the sort) or its neighbours.
struct SlabHostTimes {
    double sort_sync = 0, mig_sync = 0, enqueue = 0;
    double gpu_mig = 0, gpu_mid = 0, gpu_halo = 0; // device time: migrant exchange, merge/reorder, halo exchange
    long long calls = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};
static thread_local SlabHostTimes g_slab_times;
static double wall_now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

// Cell-list build in slab mode: classify + sort, migrate, merge, reorder, bounds, ghost exchange.
static int ensure_sorted_slab(cf_sim* s, cudaEvent_t ev_x0, cudaEvent_t ev_x1) {
    if (s->sorted_valid) return 0;
    const double t_begin = wall_now();
    const int cur = s->cur, nxt = cur ^ 1;
    const int B = s->base;
    int n = s->n;
    const long long KC = (long long)s->ncell * CF_KEY_SUB;
    s->geom.class_stride = (uint32_t)KC;
    int src = 0;
    int n_stay = 0, n_left = 0, n_right = 0;
    if (n > 0) {
        LAUNCH(s, slab_key_kernel, div_up(n, 256), 256, 0, s->pos[cur] + B, s->keys[0], s->vals[0], n, s->sc, s->geom,
               s->d_slab_counts + 3);
        if (int rc = radix_sort_pairs(s, s->keys, s->vals, n, 3 * KC, &src)) return rc;
        LAUNCH(s, slab_class_counts_kernel, 1, 32, 0, s->keys[src], n, (uint32_t)KC, s->d_slab_counts);
    } else {
        CU(cudaMemsetAsync(s->d_slab_counts, 0, 3 * sizeof(int), s->stream));
    }
    const int mcap = slab_mig_msg_cap(s), hcap = slab_halo_msg_cap(s);
    if (ev_x0) CU(cudaEventRecord(ev_x0, s->stream));
    const bool dbg = getenv("CF_SLAB_DEBUG") != nullptr;
    if (dbg && !g_slab_times.ev[0])
        for (int i = 0; i < 4; i++) cudaEventCreate(&g_slab_times.ev[i]);
    if (dbg) cudaEventRecord(g_slab_times.ev[0], s->stream);
    // ---- migrants: packed with device-side counts, so nothing waits for the host here ----
    LAUNCH(s, slab_pack_migrants_kernel, div_up(2 * mcap, 256), 256, 0, s->vals[src], s->pos[cur] + B, s->vel[cur] + B,
           s->id[cur] + B, s->d_slab_counts, s->send_mig[0], s->send_mig[1], mcap);
    if (int rc = slab_exchange(s, s->send_mig[0], s->send_mig[1], s->recv_mig[0], s->recv_mig[1],
                               slab_mig_bytes(mcap)))
        return rc;
    // the only host synchronisation of the cell-list build: class counts + arrival counts
    CU(cudaMemcpyAsync(s->h_slab_counts, s->d_slab_counts, 4 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaMemcpyAsync(&s->h_slab_counts[6], mig_count(s->recv_mig[0], mcap), sizeof(int),
                       cudaMemcpyDeviceToHost, s->stream));
    CU(cudaMemcpyAsync(&s->h_slab_counts[7], mig_count(s->recv_mig[1], mcap), sizeof(int),
                       cudaMemcpyDeviceToHost, s->stream));
    if (dbg) cudaEventRecord(g_slab_times.ev[1], s->stream);
    const double t_s1 = wall_now();
    CU(cudaStreamSynchronize(s->stream));
    if (g_slab_times.calls >= 5) g_slab_times.mig_sync += wall_now() - t_s1;
    if (int rc = slab_check_flags(s)) return rc;
    n_stay = s->h_slab_counts[0], n_left = s->h_slab_counts[1], n_right = s->h_slab_counts[2];
    if (n_left > mcap || n_right > mcap)
        return fail(CF_ERR_CAPACITY, "%d/%d migrants exceed the migrant message capacity (%d): raise migrant_slack",
                    n_left, n_right, mcap);
    const int n_al = s->h_slab_counts[6], n_ar = s->h_slab_counts[7], n_a = n_al + n_ar;
    const int n_new = n_stay + n_a;
    if (n + n_a > s->cap_own || n_new > s->cap_own)
        return fail(CF_ERR_CAPACITY, "rank %d would own %d particles, capacity %d", s->rank, n_new, s->cap_own);

    uint32_t* fkeys = s->keys[src];
    uint32_t* fvals = s->vals[src];
    if (n_a > 0) {
        LAUNCH(s, slab_unpack_arrivals_kernel, div_up(n_a, 256), 256, 0, s->recv_mig[0], s->recv_mig[1], n_al, n_ar,
               mcap, s->pos[cur] + B, s->vel[cur] + B, s->id[cur] + B, n, s->akeys[0], s->avals[0], s->sc);
        int asrc = 0;
        if (int rc = radix_sort_pairs(s, s->akeys, s->avals, n_a, KC, &asrc)) return rc;
        LAUNCH(s, slab_merge_kernel, div_up(n_new, 256), 256, 0, s->keys[src], s->vals[src], n_stay, s->akeys[asrc],
               s->avals[asrc], n_a, s->keys[src ^ 1], s->vals[src ^ 1]);
        fkeys = s->keys[src ^ 1];
        fvals = s->vals[src ^ 1];
        src ^= 1;
    }
    // ---- reorder into the other buffer, cell bounds of the owned layers ----
    LAUNCH(s, reorder_bounds_kernel, div_up(std::max(n_new, s->ncell + 1), 256), 256, 0, fvals, fkeys, s->pos[cur] + B,
           s->vel[cur] + B, s->id[cur] + B, s->pos[nxt] + B, s->vel[nxt] + B, s->id[nxt] + B, n_new, nullptr,
           s->cell_start, s->ncell, B, nullptr);
    if (src != 0) std::swap(s->keys[0], s->keys[1]), std::swap(s->vals[0], s->vals[1]);
    s->cur = nxt;
    s->n = n_new;

    // ---- ghost layers ----
    if (dbg) cudaEventRecord(g_slab_times.ev[2], s->stream);
    const int layer_cells = s->sc.dims[1] * s->sc.dims[2];
    LAUNCH(s, slab_pack_halo_kernel, div_up(hcap, 256), 256, 0, s->pos[nxt], s->id[nxt], s->cell_start,
           layer_cells, s->nxl, s->send_halo[0], s->send_halo[1], hcap, s->d_slab_counts + 3);
    if (int rc = slab_exchange(s, s->send_halo[0], s->send_halo[1], s->recv_halo[0], s->recv_halo[1],
                               slab_halo_bytes(hcap)))
        return rc;
    LAUNCH(s, slab_unpack_ghosts_kernel, div_up(hcap, 256), 256, 0, s->recv_halo[0], s->recv_halo[1], hcap,
           s->pos[nxt], s->id[nxt], B, n_new, s->gkeys[0], s->gkeys[1], s->sc);
    LAUNCH(s, slab_ghost_bounds_kernel, div_up(2 * layer_cells + 1, 256), 256, 0, s->gkeys[0], s->gkeys[1],
           s->recv_halo[0], s->recv_halo[1], hcap, s->cell_start, layer_cells, s->ncell, B, n_new, s->d_slab_counts + 4);
    if (ev_x1) CU(cudaEventRecord(ev_x1, s->stream));
    s->sorted_valid = true;
    CU(cudaGetLastError());
    if (dbg) {
        cudaEventRecord(g_slab_times.ev[3], s->stream);
        cudaEventSynchronize(g_slab_times.ev[3]);
        float a = 0, b = 0, c = 0;
        cudaEventElapsedTime(&a, g_slab_times.ev[0], g_slab_times.ev[1]);
        cudaEventElapsedTime(&b, g_slab_times.ev[1], g_slab_times.ev[2]);
        cudaEventElapsedTime(&c, g_slab_times.ev[2], g_slab_times.ev[3]);
        if (g_slab_times.calls >= 5) g_slab_times.gpu_mig += a, g_slab_times.gpu_mid += b, g_slab_times.gpu_halo += c;
    }
    if (g_slab_times.calls >= 5) g_slab_times.enqueue += wall_now() - t_begin;
    g_slab_times.calls++;
    return 0;
}

// Global initial condition in slab mode: every rank generates all n_total particles (counter-
// based generator, any rank can regenerate any particle) and keeps, in id order, those whose x
// lies in its slab.
static int slab_init_particles(cf_sim* s, long long n_total, uint64_t seed, int mode) {
    ARG(n_total >= 0 && n_total < (1ll << 31));
    const int N = (int)n_total;
    s->n_total = n_total;
    s->geom.x_lo = slab_bound(s, s->rank);
    s->geom.x_hi = slab_bound(s, s->rank + 1);
    s->geom.W = s->params.canvasWidth;
    s->geom.slab_w = s->params.canvasWidth / (float)s->world;
    float4 *tp = nullptr, *tv = nullptr, *tf = nullptr;
    int* ti = nullptr;
    uint32_t *k[2] = {nullptr, nullptr}, *v[2] = {nullptr, nullptr};
    size_t c = (size_t)std::max(N, 1);
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaMalloc(&tp, c * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc(&tv, c * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc(&tf, c * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc(&ti, c * sizeof(int));
    for (int b = 0; b < 2 && e == cudaSuccess; b++) {
        e = cudaMalloc(&k[b], c * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc(&v[b], c * sizeof(uint32_t));
    }
    int rc = 0;
    int mine = 0;
    if (e != cudaSuccess) rc = fail(CF_ERR_CUDA, "slab init: %s", cudaGetErrorString(e));
    if (rc == 0 && N > 0) {
        LAUNCH(s, init_particles_kernel, div_up(N, 256), 256, 0, tp, tv, tf, ti, N, 0, s->T, seed, mode,
               s->params.canvasWidth, s->params.canvasHeight, s->params.canvasDepth);
        LAUNCH(s, slab_init_class_kernel, div_up(N, 256), 256, 0, tp, k[0], v[0], N, s->geom);
        int src = 0;
        rc = radix_sort_pairs(s, k, v, N, 2, &src);
        if (rc == 0) {
            LAUNCH(s, slab_class_counts_kernel, 1, 32, 0, k[src], N, 1u, s->d_slab_counts);
            if (cudaMemcpyAsync(s->h_slab_counts, s->d_slab_counts, 3 * sizeof(int), cudaMemcpyDeviceToHost,
                                s->stream) != cudaSuccess ||
                cudaStreamSynchronize(s->stream) != cudaSuccess)
                rc = fail(CF_ERR_CUDA, "slab init sync failed");
        }
        if (rc == 0) {
            mine = s->h_slab_counts[0];
            if (mine > s->cap_own) rc = fail(CF_ERR_CAPACITY, "rank %d owns %d particles, capacity %d", s->rank, mine, s->cap_own);
        }
        if (rc == 0 && mine > 0)
            LAUNCH(s, reorder_kernel, div_up(mine, 256), 256, 0, v[src], tp, tv, ti, opos(s), ovel(s), oid(s), mine);
        if (rc == 0 && mine > 0) cudaMemsetAsync(ofrc(s), 0, sizeof(float4) * (size_t)mine, s->stream);
    }
    cudaStreamSynchronize(s->stream);
    cudaFree(tp), cudaFree(tv), cudaFree(tf), cudaFree(ti);
    for (int b = 0; b < 2; b++) cudaFree(k[b]), cudaFree(v[b]);
    if (rc) return rc;
    s->n = mine;
    s->sorted_valid = false;
    return CF_OK;
}
