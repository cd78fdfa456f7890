// slab_host.inl — host side of the multi-GPU slab decomposition (included by engine.cu).
// One process per GPU.  Ring neighbours exchange migrants and one ghost layer per step by storing into each
// other's mailboxes (CUDA IPC mapped device memory, NVLink 5 / NVSwitch) from inside the producing kernels —
// no NCCL call, no host synchronisation and no host-visible count anywhere on the step path (device side:
// kernels_slab.cuh).  world == 1 maps the rank's own mailbox as both neighbours, which runs the identical
// protocol on a single GPU.

static void slab_free(cf_sim* s) {
    for (int d = 0; d < 2; d++) {
        if (s->peer_box[d] && s->peer_box[d] != s->mailbox && !(d == 1 && s->peer_box[1] == s->peer_box[0]))
            cudaIpcCloseMemHandle(s->peer_box[d]);
        cudaFree(s->gkeys[d]);
        s->gkeys[d] = nullptr;
    }
    s->peer_box[0] = s->peer_box[1] = nullptr;
    cudaFree(s->mailbox);
    s->mailbox = nullptr;
    cudaFree(s->d_slab);
    s->d_slab = nullptr;
    if (s->h_slab) cudaFreeHost(s->h_slab);
    s->h_slab = nullptr;
    s->connected = false;
}

// Slab bound r of `world`: the user's bounds (cf_slab_set_bounds, identical on every rank) or the uniform split.
static float slab_bound(const cf_sim* s, int r) {
    if (!s->bounds.empty()) return s->bounds[(size_t)std::max(0, std::min(r, s->world))];
    return (float)((double)s->params.canvasWidth * (double)r / (double)s->world);
}
static float slab_width(const cf_sim* s, int r) {
    r = ((r % s->world) + s->world) % s->world;
    return slab_bound(s, r + 1) - slab_bound(s, r);
}

static void slab_update_geom(cf_sim* s) {
    s->geom.x_lo = slab_bound(s, s->rank);
    s->geom.x_hi = slab_bound(s, s->rank + 1);
    s->geom.W = s->params.canvasWidth;
    s->geom.w_own = slab_width(s, s->rank);
    s->geom.w_left = slab_width(s, s->rank - 1);
    s->geom.w_right = slab_width(s, s->rank + 1);
}

// Blocks of 256 threads for a grid-stride walk over the owned particles: sized for the expected count (the
// global count over the ranks, else the capacity), capped at a few waves of the device.
static int slab_grid(const cf_sim* s) {
    const long long expect = s->n_total > 0 ? std::min<long long>(s->cap_own, s->n_total / s->world + 1) : s->cap_own;
    // (every block of an emitting kernel ends with a fence + a ticket atomic on one word: a few blocks per SM, not
    //  one per 256 particles — 3,900 tickets cost 25 us of the slab integrate at 1 M particles)
    return (int)std::max<long long>(1, std::min<long long>((expect + 255) / 256, (long long)s->sm_count * 8));
}

static SlabPeers slab_peers(const cf_sim* s) {
    SlabPeers P;
    P.left = s->peer_box[0];
    P.right = s->peer_box[1];
    P.mail = s->mail;
    P.status = s->d_slab;
    return P;
}

// The owned count and the error word live on the device; this is where the host learns them (called where
// the API synchronises anyway: cf_sync, downloads, stats, graph results).
static int slab_refresh(cf_sim* s) {
    if (!s->slab) return 0;
    CU(cudaMemcpyAsync(s->h_slab, s->d_slab, SLAB_WORDS * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    s->n = std::max(0, std::min(s->h_slab[SLAB_NCUR], s->cap_own));
    const int e = s->h_slab[SLAB_ERR];
    if (e == 0) return 0;
    cudaMemsetAsync(s->d_slab + SLAB_ERR, 0, sizeof(int), s->stream); // reported once
    if (e & SLAB_ERR_TIMEOUT)
        return fail(CF_ERR_STATE, "rank %d: a neighbour's message did not arrive within %.0f ms (a rank stopped stepping?)",
                    s->rank, s->wait_timeout_ms);
    if (e & SLAB_ERR_FAR) return fail(CF_ERR_STATE, "rank %d: a particle moved further than the neighbouring slab in one step", s->rank);
    if (e & SLAB_ERR_HALO)
        return fail(CF_ERR_CAPACITY, "rank %d: a boundary layer exceeded the halo capacity (%d): raise halo_capacity", s->rank,
                    s->mail.cap_halo);
    if (e & SLAB_ERR_MIG)
        return fail(CF_ERR_CAPACITY, "rank %d: leavers exceeded the migrant capacity (%d): raise migrant_capacity", s->rank,
                    s->mail.cap_mig);
    return fail(CF_ERR_CAPACITY, "rank %d would own more particles than its capacity %d", s->rank, s->cap_own);
}

static int slab_set_owned_count(cf_sim* s, int n) {
    s->n = n;
    s->h_slab[SLAB_NCUR] = n;
    CU(cudaMemcpyAsync(s->d_slab + SLAB_NCUR, &s->h_slab[SLAB_NCUR], sizeof(int), cudaMemcpyHostToDevice, s->stream));
    return 0;
}

extern "C" int cf_comm_init(cf_sim* s, int rank, int world, int capacity) {
    ARG(s && world >= 1 && rank >= 0 && rank < world && capacity >= 1);
    if (int rc = set_device(s)) return rc;
    CU(cudaStreamSynchronize(s->stream));
    slab_free(s);
    s->rank = rank;
    s->world = world;
    s->slab = true;
    s->cap_own = capacity;
    // a boundary layer can be the whole slab (one or two x layers per rank at 8-way strong scaling)
    if (s->cap_halo <= 0) s->cap_halo = std::max(32768, capacity);
    if (s->cap_mig <= 0) s->cap_mig = std::max(16384, capacity / 8);
    s->mail.cap_halo = s->cap_halo;
    s->mail.cap_mig = s->cap_mig;
    s->n = 0;
    s->bounds.clear();
    if (int rc = alloc_particle_buffers(s, s->cap_own + 2 * s->cap_halo)) return rc;
    s->base = s->cap_halo;
    for (int d = 0; d < 2; d++) CU(cudaMalloc(&s->gkeys[d], sizeof(uint32_t) * (size_t)s->cap_halo));
    CU(cudaMalloc(&s->mailbox, s->mail.bytes()));
    CU(cudaMemsetAsync(s->mailbox, 0, s->mail.bytes(), s->stream));
    CU(cudaMalloc(&s->d_slab, SLAB_WORDS * sizeof(int)));
    CU(cudaMemsetAsync(s->d_slab, 0, SLAB_WORDS * sizeof(int), s->stream));
    CU(cudaMallocHost(&s->h_slab, SLAB_WORDS * sizeof(int)));
    memset(s->h_slab, 0, SLAB_WORDS * sizeof(int));
    s->seq_mig = s->seq_halo = 0;
    s->mig_sent = false;
    CU(cudaStreamSynchronize(s->stream));
    if (world == 1) { // loop-back: my own mailbox is both neighbours'
        s->peer_box[0] = s->peer_box[1] = s->mailbox;
        s->connected = true;
    }
    return CF_OK;
}

extern "C" int cf_comm_mailbox_handle(cf_sim* s, void* handle64) {
    ARG(s && handle64 && s->slab && s->mailbox);
    if (int rc = set_device(s)) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == CF_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, s->mailbox));
    memcpy(handle64, &h, sizeof(h));
    return CF_OK;
}

extern "C" int cf_comm_connect(cf_sim* s, const void* left_handle64, const void* right_handle64) {
    ARG(s && s->slab && s->mailbox);
    if (int rc = set_device(s)) return rc;
    if (s->world == 1) return CF_OK;
    ARG(left_handle64 && right_handle64);
    const void* hs[2] = {left_handle64, right_handle64};
    for (int d = 0; d < 2; d++) {
        if (d == 1 && s->world == 2) { // both neighbours are the same rank: one mapping
            s->peer_box[1] = s->peer_box[0];
            break;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, hs[d], sizeof(h));
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
            return fail(CF_ERR_CUDA, "cudaIpcOpenMemHandle (%s neighbour) failed: %s — the ranks of a slab run need peer access "
                        "(NVLink / NVSwitch or PCIe P2P) between their GPUs", d ? "right" : "left", cudaGetErrorString(e));
        s->peer_box[d] = (char*)p;
    }
    s->connected = true;
    return CF_OK;
}

static int slab_drain(cf_sim* s);
extern "C" int cf_slab_set_bounds(cf_sim* s, const float* bounds, int count) {
    ARG(s && s->slab && bounds && count == s->world + 1);
    for (int r = 0; r < s->world; r++) ARG(bounds[r + 1] > bounds[r]);
    ARG(bounds[0] == 0.0f);
    if (int rc = set_device(s)) return rc;
    if (int rc = slab_drain(s)) return rc; // a pending migrant exchange belongs to the old bounds (collective)
    s->bounds.assign(bounds, bounds + count);
    s->sorted_valid = false, s->state_gen++;
    return CF_OK;
}

// Cell-list build in slab mode, all on the device: (emit migrants,) wait + append arrivals, one class sort,
// reorder + bounds (new owned count), halo pack -> neighbours' mailboxes, wait + unpack ghosts, ghost bounds.
static int ensure_sorted_slab(cf_sim* s, cudaEvent_t* ev_x) {
    if (s->sorted_valid) return 0;
    if (!s->connected) return fail(CF_ERR_STATE, "slab mode: cf_comm_connect has not been called");
    const int cur = s->cur, nxt = cur ^ 1;
    const int B = s->base;
    const int NU = s->cap_own; // launch bound of everything that walks the owned particles
    const int gs = slab_grid(s); // grid-stride kernels: blocks for the expected count, not for the capacity
    const long long KC = (long long)s->ncell * CF_KEY_SUB;
    s->geom.class_stride = (uint32_t)KC;
    const SlabPeers P = slab_peers(s);
    const unsigned long long tmo = (unsigned long long)(s->wait_timeout_ms * 1e6);
    // ---- migrants: emitted by the previous step's integrate kernel, or here after an upload / spawn / move ----
    if (!s->mig_sent) {
        s->seq_mig++;
        LAUNCH(s, slab_emit_migrants_kernel, gs, 256, 0, s->pos[cur] + B, s->vel[cur] + B, s->id[cur] + B, NU,
               s->geom, P, s->seq_mig);
    }
    s->mig_sent = false;
    if (ev_x) CU(cudaEventRecord(ev_x[0], s->stream));
    const int mpar = s->seq_mig & 1;
    LAUNCH(s, slab_wait_kernel, 1, 32, 0, s->mail.flag_mig(s->mailbox, 0), s->mail.flag_mig(s->mailbox, 1), s->seq_mig,
           s->d_slab, tmo);
    LAUNCH(s, slab_unpack_arrivals_kernel, 1, 1024, 0, s->mailbox, s->mail, mpar, s->pos[cur] + B, s->vel[cur] + B,
           s->id[cur] + B, s->cap_own, s->d_slab);
    if (ev_x) CU(cudaEventRecord(ev_x[1], s->stream));
    // ---- one sort: stayers + arrivals by cell, leavers behind them ----
    int src = 0;
    if (int rc = radix_sort_run<SlabKeyFn, true>(s, SlabKeyFn{s->pos[cur] + B, s->sc, s->geom, s->d_slab}, s->keys, s->vals, NU,
                                                 s->d_slab + SLAB_NTMP, 2 * KC, &src))
        return rc;
    LAUNCH(s, reorder_bounds_kernel, div_up(std::max(NU, s->ncell + 1), 256), 256, 0, s->vals[src], s->keys[src],
           s->pos[cur] + B, s->vel[cur] + B, s->id[cur] + B, s->pos[nxt] + B, s->vel[nxt] + B, s->id[nxt] + B, NU,
           s->d_slab + SLAB_NTMP, s->cell_start, s->ncell, B, s->d_slab + SLAB_NCUR);
    if (src != 0) std::swap(s->keys[0], s->keys[1]), std::swap(s->vals[0], s->vals[1]);
    s->cur = nxt, s->state_gen++;
    // ---- ghost layers ----
    if (ev_x) CU(cudaEventRecord(ev_x[2], s->stream));
    const int layer_cells = s->sc.dims[1] * s->sc.dims[2];
    s->seq_halo++;
    const int hpar = s->seq_halo & 1;
    LAUNCH(s, slab_pack_halo_kernel, std::max(1, gs / 4), 256, 0, s->pos[nxt], s->id[nxt], s->cell_start, layer_cells,
           s->nxl, P, s->seq_halo);
    LAUNCH(s, slab_wait_kernel, 1, 32, 0, s->mail.flag_halo(s->mailbox, 0), s->mail.flag_halo(s->mailbox, 1), s->seq_halo,
           s->d_slab, tmo);
    LAUNCH(s, slab_unpack_ghosts_kernel, std::max(1, gs / 4), 256, 0, s->mailbox, s->mail, hpar, s->pos[nxt], s->id[nxt],
           B, s->d_slab, s->gkeys[0], s->gkeys[1], s->sc);
    LAUNCH(s, slab_ghost_bounds_kernel, div_up(2 * layer_cells + 1, 256), 256, 0, s->gkeys[0], s->gkeys[1], s->mailbox, s->mail,
           hpar, s->cell_start, layer_cells, s->ncell, B, s->d_slab);
    if (ev_x) CU(cudaEventRecord(ev_x[3], s->stream));
    s->sorted_valid = true;
    CU(cudaGetLastError());
    return 0;
}

// A pending migrant exchange (sent by the last step's integrate) must be consumed before the state is
// replaced or shifted: every rank does this collectively, like the build itself.
static int prepare_step_const(cf_sim* s);
static int slab_drain(cf_sim* s) {
    if (!s->slab || !s->mig_sent) return 0;
    if (int rc = prepare_step_const(s)) return rc;
    return ensure_sorted_slab(s, nullptr);
}

// Global initial condition in slab mode: every rank generates all n_total particles (counter-
// based generator, any rank can regenerate any particle) and keeps, in id order, those whose x
// lies in its slab.
static int slab_init_particles(cf_sim* s, long long n_total, uint64_t seed, int mode) {
    ARG(n_total >= 0 && n_total < (1ll << 31));
    if (int rc = slab_drain(s)) return rc;
    const int N = (int)n_total;
    s->n_total = n_total;
    slab_update_geom(s);
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
            LAUNCH(s, slab_count_mine_kernel, 1, 32, 0, k[src], N, s->d_slab + SLAB_NTMP);
            if (cudaMemcpyAsync(&s->h_slab[SLAB_NTMP], s->d_slab + SLAB_NTMP, sizeof(int), cudaMemcpyDeviceToHost,
                                s->stream) != cudaSuccess ||
                cudaStreamSynchronize(s->stream) != cudaSuccess)
                rc = fail(CF_ERR_CUDA, "slab init sync failed");
        }
        if (rc == 0) {
            mine = s->h_slab[SLAB_NTMP];
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
    if (int rc2 = slab_set_owned_count(s, mine)) return rc2;
    s->sorted_valid = false, s->state_gen++;
    return CF_OK;
}
