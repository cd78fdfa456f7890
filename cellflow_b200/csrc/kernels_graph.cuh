// kernels_graph.cuh — same-type proximity graph over the cell list, warp-aggregated compaction.
//
// Rule (generateProximityGraphKernel, ParticleSimulation.cu:212-246), per particle i:
//   candidates = the first 2*maxConn particles j > i IN ORIGINAL INDEX ORDER with the same type
//   and plain (non-wrapped) d2 = fma(dz,dz,fma(dx,dx,dy*dy)) < dist^2; stable insertion sort by
//   d2; keep the first maxConn.
// The reference scans all j > i; here candidates come from the 27 cells (clamped, not periodic,
// because the rule does not wrap) of a dedicated (type, cell) list with cell edge >= dist, the
// 2*maxConn smallest original ids are kept in a sorted per-thread list, and the same stable sort
// by d2 selects the edges.  The edge SET is
// identical; the reference's own edge ORDER is nondeterministic (one atomicAdd per thread).
#pragma once
#include "cf_device.cuh"

#define CF_GRAPH_K 32 // 2 * CF_MAX_GRAPH_CONN candidates, the reference's nearby[32] (.cu:210)

// The graph has its own cell list: only same-type pairs within `dist` matter, so particles are
// keyed by (type, cell of edge >= dist) — 27 cells then hold ~27 * n_type * dist^3 / V candidates
// instead of every particle of 27 force cells of all types (25x fewer at BASELINE config 5).
struct GraphGrid {
    float org[3];   // lower corner of the gridded region
    float inv[3];   // cells per unit length
    int dims[3];
    int ncell;      // dims product
    int T;
    float x_min, x_max; // particles outside [x_min, x_max) (seam ghosts) are not gridded
};

__device__ __forceinline__ int graph_coord(float x, float org, float inv, int n) {
    int c = (int)__fmul_rn(__fsub_rn(x, org), inv);
    return c < 0 ? 0 : (c > n - 1 ? n - 1 : c);
}

// key = type * ncell + cell for every slot in [first, first + n); excluded slots get the end key.
// Slots that take part are [lo, hi): the owned range [own_first, own_first + own_n), extended by
// the low / high ghost layer when it is a real neighbour (not the periodic seam).  The ghost
// extents are read from the cell list on the device (cell_start[0], cell_start[ncell]), so the
// host never has to wait for the ghost counts.
// Slab mode: the slot count (d_n) and the owned count (d_own_n) are read on the device.
__global__ void graph_key_kernel(const float4* __restrict__ pos4, int first_host, const int* __restrict__ d_first,
                                 int n_upper, const int* __restrict__ d_n,
                                 GraphGrid g, const int* __restrict__ cell_start, int ncell, int own_first, int own_n_host,
                                 const int* __restrict__ d_own_n, int use_lo_ghost, int use_hi_ghost,
                                 uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int n = d_n ? min(*d_n, n_upper) : n_upper;
    const int own_n = d_own_n ? *d_own_n : own_n_host;
    const int first = d_first ? *d_first : first_host;
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int slot = first + k;
    const int lo = use_lo_ghost ? cell_start[0] : own_first;
    const int hi = use_hi_ghost ? cell_start[ncell] : own_first + own_n;
    uint32_t key = (uint32_t)g.T * (uint32_t)g.ncell; // "not gridded"
    if (slot >= lo && slot < hi) {
        float4 p = pos4[slot];
        uint32_t t = __float_as_uint(p.w);
        if (t < (uint32_t)g.T && p.x >= g.x_min && p.x < g.x_max) {
            int cx = graph_coord(p.x, g.org[0], g.inv[0], g.dims[0]);
            int cy = graph_coord(p.y, g.org[1], g.inv[1], g.dims[1]);
            int cz = graph_coord(p.z, g.org[2], g.inv[2], g.dims[2]);
            key = t * (uint32_t)g.ncell + (uint32_t)((cx * g.dims[1] + cy) * g.dims[2] + cz);
        }
    }
    keys[k] = key;
    vals[k] = (uint32_t)slot;
}

// Positions in graph order with the original id in .w, so the candidate loop is one 16-byte load.
__global__ void graph_gather_kernel(const uint32_t* __restrict__ gvals, const float4* __restrict__ pos4,
                                    const int* __restrict__ id, int n_upper, const int* __restrict__ d_n,
                                    float4* __restrict__ gpos) {
    const int n = d_n ? min(*d_n, n_upper) : n_upper;
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    uint32_t slot = gvals[q];
    float4 p = pos4[slot];
    gpos[q] = make_float4(p.x, p.y, p.z, __int_as_float(id[slot]));
}

// gstart[k] = first graph-order position whose key is >= k, k in [0, nkeys].
__global__ void graph_bounds_kernel(const uint32_t* __restrict__ skeys, int n_upper, const int* __restrict__ d_n,
                                    int* __restrict__ gstart, int nkeys) {
    const int n = d_n ? min(*d_n, n_upper) : n_upper;
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > nkeys) return;
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (skeys[mid] < (uint32_t)k) lo = mid + 1; else hi = mid;
    }
    gstart[k] = lo;
}

// One thread per graph-order position q (threads of a warp share a cell: candidate loads are
// warp-uniform broadcasts).  Only owned particles (slot in [own_first, own_first + n_own)) emit.
// The per-thread candidate lists live in shared memory, laid out [entry][thread]: dynamic
// indexing costs one conflict-free LDS/STS instead of a scattered local-memory transaction
// per lane (the first version kept them in local memory and spent most of its time there).
// The lists are UNORDERED: a candidate is appended, or replaces the largest id kept once the list is
// full (that id and its position are found by one scan of the list whenever it changes while full), and
// the final stable sort by d2 breaks ties by id (index order = id order).  The id-sorted insertion of
// round 1 ran its shared-memory shift loop at 2 active lanes per instruction for 40 % of the kernel's
// instructions (ncu, profiles/r02_graph_blocks_c5-settings-2M.txt): graph phase 1.09 -> 0.83 ms at
// BASELINE config 5.
// The lists are sized for the actual 2*maxConn (dynamic shared memory: 7.5 KB per CTA at the
// default maxConn = 5, so 32 CTAs = all 64 warps of an SM are resident), and the candidate loop
// keeps CF_GRAPH_BATCH independent 16-byte loads in flight: the kernel is bound by the latency
// of those (L2-resident) loads, not by arithmetic.
#define CF_GRAPH_THREADS 64
#ifndef CF_GRAPH_APPEND
#define CF_GRAPH_APPEND 1 // unordered candidate lists (0: id-sorted insertion, the round-1 form)
#endif
#define CF_GRAPH_BATCH 4
struct GraphList {
    int* id;
    int* q;
    float* d2;
    __device__ __forceinline__ int& I(int k) { return id[k * CF_GRAPH_THREADS]; }
    __device__ __forceinline__ int& Q(int k) { return q[k * CF_GRAPH_THREADS]; }
    __device__ __forceinline__ float& D(int k) { return d2[k * CF_GRAPH_THREADS]; }
};

__host__ __device__ constexpr size_t graph_list_bytes(int max_conn) { // [3][2*maxConn][threads]: id, position, d2
    return (size_t)3 * 2 * max_conn * CF_GRAPH_THREADS * sizeof(int);
}

__global__ void __launch_bounds__(CF_GRAPH_THREADS)
graph_kernel(const float4* __restrict__ gpos, const uint32_t* __restrict__ gvals, const uint32_t* __restrict__ gkeys,
             const int* __restrict__ gstart, int nq_upper, const int* __restrict__ d_nq, int own_first, int n_own_host,
             const int* __restrict__ d_own_n, GraphGrid g, float dist2, int max_conn, int2* __restrict__ edges,
             int2* __restrict__ edge_slots, int capacity, int* __restrict__ edge_count) {
    extern __shared__ int s_lists[];
    const int nq = d_nq ? min(*d_nq, nq_upper) : nq_upper;
    const int n_own = d_own_n ? *d_own_n : n_own_host; // [3][K][CF_GRAPH_THREADS]: id, graph position, d2
    const int K = 2 * max_conn;
    GraphList L{s_lists + threadIdx.x, s_lists + K * CF_GRAPH_THREADS + threadIdx.x,
                reinterpret_cast<float*>(s_lists + 2 * K * CF_GRAPH_THREADS) + threadIdx.x};
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    int ncand = 0, my_id = 0, my_slot = 0;
#if CF_GRAPH_APPEND
    int maxid = 0x7fffffff, maxpos = 0;
#endif
    bool active = false;
    if (q < nq) {
        uint32_t key = gkeys[q];
        my_slot = (int)gvals[q];
        active = key < (uint32_t)g.T * (uint32_t)g.ncell && my_slot >= own_first && my_slot < own_first + n_own;
        if (active) {
            float4 p = gpos[q];
            my_id = __float_as_int(p.w);
            int t = (int)(key / (uint32_t)g.ncell);
            int cell = (int)(key - (uint32_t)t * (uint32_t)g.ncell);
            int cz = cell % g.dims[2], cy = (cell / g.dims[2]) % g.dims[1], cx = cell / (g.dims[2] * g.dims[1]);
            int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dims[0] - 1);
            int y0 = max(cy - 1, 0), y1 = min(cy + 1, g.dims[1] - 1);
            int z0 = max(cz - 1, 0), z1 = min(cz + 1, g.dims[2] - 1);
            const int tbase = t * g.ncell;
            for (int x = x0; x <= x1; x++)
                for (int y = y0; y <= y1; y++) {
                    int row = tbase + (x * g.dims[1] + y) * g.dims[2];
                    // z-adjacent cells of one type are contiguous: one range per (x, y)
                    int j0 = gstart[row + z0], j1 = gstart[row + z1 + 1];
                    for (int jb = j0; jb < j1; jb += CF_GRAPH_BATCH) {
                        float4 ob[CF_GRAPH_BATCH];
#pragma unroll
                        for (int u = 0; u < CF_GRAPH_BATCH; u++) // independent loads; id -1 is never a candidate
                            ob[u] = jb + u < j1 ? gpos[jb + u] : make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
#pragma unroll
                        for (int u = 0; u < CF_GRAPH_BATCH; u++) { // in index order, as the rule scans
                            const float4 o = ob[u];
                            const int j = jb + u;
                            int jid = __float_as_int(o.w);
                            if (jid <= my_id) continue;
                            float dx = __fsub_rn(o.x, p.x), dy = __fsub_rn(o.y, p.y), dz = __fsub_rn(o.z, p.z);
                            float d2 = cf_dist2(dx, dy, dz);
                            if (!(d2 < dist2)) continue;
#if CF_GRAPH_APPEND
                            // unordered list: append, or replace the largest id kept (maxid at maxpos) once it is full
                            if (jid > maxid) continue;
                            const int pos = ncand < K ? ncand++ : maxpos;
                            L.I(pos) = jid;
                            L.Q(pos) = j;
                            L.D(pos) = d2;
                            if (ncand == K) { // full: the largest id kept and where it sits
                                maxid = L.I(0), maxpos = 0;
                                for (int e = 1; e < K; e++) {
                                    const int v = L.I(e);
                                    if (v > maxid) maxid = v, maxpos = e;
                                }
                            }
#else
                            if (ncand == K && jid > L.I(K - 1)) continue;
                            // insert into the id-sorted candidate list (drop the largest id when full)
                            int pos = ncand < K ? ncand : K - 1;
                            while (pos > 0 && L.I(pos - 1) > jid) {
                                L.I(pos) = L.I(pos - 1);
                                L.Q(pos) = L.Q(pos - 1);
                                L.D(pos) = L.D(pos - 1);
                                pos--;
                            }
                            L.I(pos) = jid;
                            L.Q(pos) = j;
                            L.D(pos) = d2;
                            if (ncand < K) ncand++;
#endif
                        }
                    }
                }
            // stable insertion sort by d2 (.cu:235-243); the list is in index order, as the
            // reference's scan would have produced it
            for (int a = 1; a < ncand; a++) {
                float kd = L.D(a);
                int ki = L.I(a), kq = L.Q(a);
                int b = a - 1;
#if CF_GRAPH_APPEND
                while (b >= 0 && (L.D(b) > kd || (L.D(b) == kd && L.I(b) > ki))) { // ties: index order = id order
#else
                while (b >= 0 && L.D(b) > kd) {
#endif
                    L.D(b + 1) = L.D(b);
                    L.I(b + 1) = L.I(b);
                    L.Q(b + 1) = L.Q(b);
                    b--;
                }
                L.D(b + 1) = kd;
                L.I(b + 1) = ki;
                L.Q(b + 1) = kq;
            }
        }
    }
    int w = active ? min(ncand, max_conn) : 0;
    // warp-aggregated compaction: one atomicAdd per warp reserves space for all its edges
    int lane = threadIdx.x & 31;
    int incl = w;
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(edge_count, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    int off = base + incl - w;
    for (int a = 0; a < w; a++) {
        int e = off + a;
        if (e < capacity) {
            edges[e] = make_int2(my_id, L.I(a));
            edge_slots[e] = make_int2(my_slot, (int)gvals[L.Q(a)]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Dense states: ONE WARP PER PARTICLE.  With hundreds of candidates per particle (the reference's
// spawn cube, or any clustered state) the thread-per-particle kernel above runs its insertion
// path on almost every candidate for the whole warp (lanes accept different candidates), and its
// per-thread serial scan is long.  Here the 32 lanes test 32 consecutive candidates at once
// (coalesced loads), and the 2*maxConn smallest ids are kept in a list DISTRIBUTED OVER THE LANES
// (one entry per lane, unordered; the largest id kept is tracked with a warp reduction): inserting one
// accepted candidate is a shuffle and a select, warp-uniform, no shared memory.  Same rule, same edge set.
// ---------------------------------------------------------------------------------------------
#define CF_GRAPHW_WARPS 4
__global__ void __launch_bounds__(CF_GRAPHW_WARPS * 32)
graph_warp_kernel(const float4* __restrict__ gpos, const uint32_t* __restrict__ gvals, const uint32_t* __restrict__ gkeys,
                  const int* __restrict__ gstart, int nq_upper, const int* __restrict__ d_nq, int own_first,
                  int n_own_host, const int* __restrict__ d_own_n, GraphGrid g, float dist2, int max_conn,
                  int2* __restrict__ edges, int2* __restrict__ edge_slots, int capacity, int* __restrict__ edge_count) {
    const int nq = d_nq ? min(*d_nq, nq_upper) : nq_upper;
    const int n_own = d_own_n ? *d_own_n : n_own_host;
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * CF_GRAPHW_WARPS + (threadIdx.x >> 5); // warp-uniform
    if (q >= nq) return;
    const uint32_t key = gkeys[q];
    const int my_slot = (int)gvals[q];
    if (!(key < (uint32_t)g.T * (uint32_t)g.ncell && my_slot >= own_first && my_slot < own_first + n_own)) return;
    const int K = 2 * max_conn; // <= 32: one list entry per lane
    const float4 p = gpos[q];
    const int my_id = __float_as_int(p.w);
    // (type, cell): the type needs the one integer division; the cell coordinates are recomputed from the position
    // with the key kernel's own expression (same floats, same result) instead of three more divisions
    const int t = (int)(key / (uint32_t)g.ncell);
    const int cx = graph_coord(p.x, g.org[0], g.inv[0], g.dims[0]);
    const int cy = graph_coord(p.y, g.org[1], g.inv[1], g.dims[1]);
    const int cz = graph_coord(p.z, g.org[2], g.inv[2], g.dims[2]);
    const int z0 = max(cz - 1, 0), z1 = min(cz + 1, g.dims[2] - 1);
    const int tbase = t * g.ncell;
    // the <= 9 candidate ranges (one per (x, y) row; z-adjacent cells of a type are contiguous):
    // lane r < 9 fetches the bounds of row r, so the 18 loads are two instructions
    int r_j0 = 0, r_j1 = 0;
    if (lane < 9) {
        const int x = cx + lane / 3 - 1, y = cy + lane % 3 - 1;
        if (x >= 0 && x < g.dims[0] && y >= 0 && y < g.dims[1]) {
            const int row = tbase + (x * g.dims[1] + y) * g.dims[2];
            r_j0 = gstart[row + z0];
            r_j1 = gstart[row + z1 + 1];
        }
    }
    // distributed list, UNORDERED: lane r < cnt holds one candidate (id, graph position).  While the list is not
    // full a candidate goes to lane cnt; once it is full it replaces the largest id (lane kl), and the new largest
    // id kth is one REDUX.MAX away — no shifting of a sorted list (round 1: a ballot, a popc and three shuffles up
    // per insertion; 40 % of the kernel's instructions at the spawn cube of BASELINE config 2)
    int l_id = 0x7fffffff, l_q = 0;
    int cnt = 0, kth = 0x7fffffff, kl = 0; // kth = largest id kept when the list is full, held by lane kl
    // chunks of 32 candidates over all rows; the next chunk is loaded before the current one is
    // processed (the kernel is bound by the latency of these loads and of the insertions)
    int row = -1, c0 = 0, end = 0;
    bool have;
#define GW_ADVANCE()                                       \
    do {                                                   \
        c0 += 32;                                          \
        have = true;                                       \
        while (c0 >= end) {                                \
            if (++row >= 9) { have = false; break; }       \
            c0 = __shfl_sync(0xffffffffu, r_j0, row);      \
            end = __shfl_sync(0xffffffffu, r_j1, row);     \
        }                                                  \
    } while (0)
    GW_ADVANCE();
    float4 nxt = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
    if (have && c0 + lane < end) nxt = gpos[c0 + lane];
    while (have) {
        const float4 o = nxt;
        const int base = c0;
        GW_ADVANCE();
        nxt = make_float4(0.f, 0.f, 0.f, __int_as_float(-1)); // id -1 is never a candidate
        if (have && c0 + lane < end) nxt = gpos[c0 + lane];
        const int jid = __float_as_int(o.w);
        const float dx = __fsub_rn(o.x, p.x), dy = __fsub_rn(o.y, p.y), dz = __fsub_rn(o.z, p.z);
        const float d2 = cf_dist2(dx, dy, dz);
        unsigned acc = __ballot_sync(0xffffffffu, jid > my_id && d2 < dist2 && jid < kth);
        while (acc) {
            const int src = __ffs(acc) - 1;
            acc &= acc - 1;
            const int c_id = __shfl_sync(0xffffffffu, jid, src);
            if (c_id >= kth) continue; // the list filled up meanwhile (warp-uniform)
            if (lane == (cnt < K ? cnt : kl)) l_id = c_id, l_q = base + src;
            cnt = min(cnt + 1, K);
            if (cnt == K) { // full: the largest id kept and its lane
                kth = __reduce_max_sync(0xffffffffu, lane < K ? l_id : (int)0x80000000);
                kl = __ffs(__ballot_sync(0xffffffffu, lane < K && l_id == kth)) - 1;
            }
        }
    }
#undef GW_ADVANCE
    if (cnt == 0) return;
    // distances of the kept candidates (same expression as the test above: bit-identical), then the stable sort by
    // d2 (.cu:235-243) as a rank: entries with smaller d2, or equal d2 and smaller id (the reference's list is in
    // index order = id order)
    float l_d2 = 0.f;
    if (lane < cnt) {
        const float4 o = gpos[l_q];
        const float dx = __fsub_rn(o.x, p.x), dy = __fsub_rn(o.y, p.y), dz = __fsub_rn(o.z, p.z);
        l_d2 = cf_dist2(dx, dy, dz);
    }
    int rank = 0;
    for (int b = 0; b < cnt; b++) {
        const float bd = __shfl_sync(0xffffffffu, l_d2, b);
        const int bi = __shfl_sync(0xffffffffu, l_id, b);
        rank += (bd < l_d2 || (bd == l_d2 && bi < l_id)) ? 1 : 0;
    }
    const int w = min(cnt, max_conn);
    int base = 0;
    if (lane == 0) base = atomicAdd(edge_count, w);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (lane < cnt && rank < w) {
        const int e = base + rank;
        if (e < capacity) {
            edges[e] = make_int2(my_id, l_id);
            edge_slots[e] = make_int2(my_slot, (int)gvals[l_q]);
        }
    }
}

// Sum over keys of (particles of the key)^2: divided by the particle count it is the mean number
// of same-(type, cell) companions of a particle, which decides between the two graph kernels.
__global__ void graph_occupancy_kernel(const int* __restrict__ gstart, int nkeys, unsigned long long* __restrict__ sumsq) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = 0;
    if (k < nkeys) {
        const unsigned long long c = (unsigned long long)(gstart[k + 1] - gstart[k]);
        v = c * c;
    }
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(sumsq, v);
}

// Reference VBO layout (.cu:255-275): per edge 2 vertices x (pos xyz + colour of i's type).
__global__ void graph_vertices_kernel(const int2* __restrict__ edge_slots, int ne,
                                      const float4* __restrict__ pos4, const float* __restrict__ colors,
                                      int num_types, float* __restrict__ out) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    int2 sl = edge_slots[e];
    float4 a = pos4[sl.x], b = pos4[sl.y];
    uint32_t t = __float_as_uint(a.w) % (uint32_t)num_types; // .cu:253
    float r = colors[3 * t], g = colors[3 * t + 1], bl = colors[3 * t + 2];
    float* v = out + 12 * (size_t)e;
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = r, v[4] = g, v[5] = bl;
    v[6] = b.x, v[7] = b.y, v[8] = b.z, v[9] = r, v[10] = g, v[11] = bl;
}
