// kernels_tile.cuh — tiled pair-force kernel for dense cells, GENERATION 3 (kept selectable with
// cf_set_option("force_kernel", 2) and used for per-type radii with short sub-runs; the default dense
// kernel is generation 4, kernels_tile4.cuh, which reuses the packed-math helpers and the tile list
// builder defined here).
//
// Work decomposition
//   tile  = up to TK_TI (128) particles of ONE cell ("i" side), owned by one WARP: every lane
//           keeps TK_IPT = 4 i-particles (4 "layers" of 32) and their force/count accumulators in
//           registers.  Tiles come from a device-built list, fetched with an atomic counter by a
//           persistent grid (clustered states make cells wildly uneven: work-based scheduling).
//   j side = the 27 neighbour cells as <= 18 contiguous runs of the sorted array (cells that are
//           adjacent in z are adjacent in memory), streamed through a warp-private double buffer
//           in shared memory, 64 particles per chunk, SoA (xs, ys, zs, table row[, half-radius]).
//
// Why it looks like this (measured on B200, profiles/r01_pipe_microbench.txt and
// profiles/r01_force_kernel_history.md)
//   * shared->register bandwidth is one 32-bit word per lane per clock per SM: an LDS.128
//     broadcast costs ~4 SM-cycles.  With one i per lane the j stream alone would need 2x the
//     cycles of the arithmetic, hence 4 i per lane (each loaded j is used 4x32 times).
//   * FFMA2/FADD2/FMUL2 (fma.rn.f32x2 ...) run at the same lane rate as the scalar forms but
//     take half the issue slots, which lets compares and LDS issue for free beside a busy FMA
//     pipe.  j particles are processed in packed pairs (x0,x1).
//   * ~85-93% of tested pairs are out of range, and a per-lane "collect accept bits, then drain"
//     scheme ran its drain loop at 7 of 32 lanes.  Instead the sort puts particles in Morton
//     order inside each cell, so that a layer (32 i) and a quad (4 j) are compact blobs, and ONE
//     warp vote per (layer, quad) decides whether the force terms are evaluated at all.
//
// Exactness: displacement = (jx + (-px)) [+ s], s in {-W, 0, +W} per run.  For a periodic axis
// with >= 4 cells the reference's two-sided wrap test (.cu:97-98) is decided by which neighbour
// cell j lives in, and adding +-W is exact (Sterbenz), so this is bit-identical to the
// reference's d; d2 = fma(dz,dz,fma(dx,dx,dy*dy)) as compiled; accept <=> d2 < cut2[ti][tj].
#pragma once
#include "cf_device.cuh"
#include "kernels_force.cuh"

#define TK_IPT 4               // i particles per lane (register tile)
#define TK_TI (32 * TK_IPT)    // i particles per tile (one warp)
#define TK_THREADS 128
#define TK_MAX_RUNS 18
#define TK_FAR 1.0e30f

typedef unsigned long long u64;

__device__ __forceinline__ u64 tk_pack(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void tk_unpack(u64 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 tk_add2(u64 a, u64 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 tk_mul2(u64 a, u64 b) {
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 tk_fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// Pair tests executed by a 27-cell stencil pass: sum over cells of n_cell * (particles in the
// distinct neighbour cells).  Used for the tested-pairs figure of cf_get_stats.
__global__ void count_tests_kernel(const int* __restrict__ cell_start, int ncell, StepConst c,
                                   unsigned long long* out) {
    int cell = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long local = 0;
    if (cell < ncell) {
        int nc = cell_start[cell + 1] - cell_start[cell];
        if (nc > 0) {
            int cz = cell % c.dims[2];
            int cy = (cell / c.dims[2]) % c.dims[1];
            int cx = cell / (c.dims[2] * c.dims[1]);
            if (cx < c.x_off || cx >= c.x_off + c.x_cells) nc = 0; // ghost layers hold no i-particles
            int xs[3], ys[3], zs[3];
            int kx = cf_axis_cells(cx, c.dims[0], c.periodic_x != 0, xs);
            int ky = cf_axis_cells(cy, c.dims[1], true, ys);
            int kz = cf_axis_cells(cz, c.dims[2], true, zs);
            unsigned long long m = 0;
            for (int a = 0; a < kx; a++)
                for (int b = 0; b < ky; b++)
                    for (int q = 0; q < kz; q++) {
                        int o = (xs[a] * c.dims[1] + ys[b]) * c.dims[2] + zs[q];
                        m += (unsigned long long)(cell_start[o + 1] - cell_start[o]);
                    }
            local = (unsigned long long)nc * m;
        }
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, local);
}

// Tile list: one thread per cell appends ceil(n_cell / TK_TI) tiles; ctrl[0] = number of tiles, ctrl[1] = fetch
// counter.  Two passes (part 0: the full 128-particle tiles, part 1: the partly filled last tile of every cell),
// so the persistent grid works through the expensive tiles first and finishes on the cheap ones — longest-
// processing-time-first; with one arbitrary order the warps that drew a full tile last left the others idle for
// up to a tile's duration (~1.5 ms at the bench default: 10 % of the SM-cycles, profiles/r02_*).  Results do not
// depend on the order.
__global__ void build_tiles_kernel(const int* __restrict__ cell_start, int ncell, int first_cell_x,
                                   int last_cell_x, int cells_per_x, int2* __restrict__ tiles,
                                   int* __restrict__ ctrl, int part, int tile_i) {
    int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= ncell) return;
    int cx = cell / cells_per_x;
    if (cx < first_cell_x || cx > last_cell_x) return; // ghost layers are never i-cells
    int n = cell_start[cell + 1] - cell_start[cell];
    if (n <= 0) return;
    const int nfull = n / tile_i; // particles per tile: 128 (generation 3, generation 4) or 32 * T4_IPT
    if (part == 0) {
        if (nfull == 0) return;
        int base = atomicAdd(&ctrl[0], nfull);
        for (int k = 0; k < nfull; k++) tiles[base + k] = make_int2(cell, k);
    } else if (n > nfull * tile_i) {
        int base = atomicAdd(&ctrl[0], 1);
        tiles[base] = make_int2(cell, nfull);
    }
}

// ---------------------------------------------------------------------------------------------
// Warp-per-tile pair-force kernel.
//
// A warp owns one tile (<= 128 particles of one cell: 4 register-resident "layers" of 32) and
// streams the neighbour runs through its private double-buffered shared-memory stage, 64 j per
// chunk; there is no block-level barrier anywhere, so every resident warp always has work
// whatever the cell occupancy (profiles/r01: a CTA-per-cell version spent 72% of its stall
// samples at __syncthreads).
//
// Inner loop, per quad of 4 staged j (3 broadcast LDS.128) and per layer k:
//   packed displacement + squared distance for the 4 pairs (12 FADD2/FMUL2/FFMA2),
//   4 compares, ONE warp vote: if no lane of the layer accepts any of the 4 j the quad is done;
//   otherwise the force terms of the 4 pairs are evaluated for all lanes from the displacements
//   still in registers (rejected pairs contribute an exact 0).  Particles are Morton-ordered inside
//   cells, so a layer and a quad are compact blobs and most quads are decided for the whole warp
//   at once: no accept masks, no per-lane drain loop, no gathers.
// ---------------------------------------------------------------------------------------------
#define TKW_WARPS 4            // warps per CTA (independent of each other)
#define TKW_JC 64              // staged j per chunk and warp

template <bool UNIFORM>
struct TkwShared {
    float x[TKW_WARPS][2][TKW_JC];
    float y[TKW_WARPS][2][TKW_JC];
    float z[TKW_WARPS][2][TKW_JC];
    float h[TKW_WARPS][2][TKW_JC];      // conservative half radius of j (non-uniform radii only)
    int t[TKW_WARPS][2][TKW_JC];        // byte offset of j's row in the transposed pair table
};

// Force terms of one packed pair of j against one i (lane).  d* are the packed displacements,
// d2 the packed squared distances, acc* packed partial sums (lo/hi are added at the end).
template <bool UNIFORM>
__device__ __forceinline__ void tkw_force_pair(u64 dx, u64 dy, u64 dz, u64 d2, bool ok0, bool ok1,
                                               const char* tab_i, int t0, int t1, float c2u, float bu,
                                               float repulsion, float attraction, float nk_log2e,
                                               u64& ax, u64& ay, u64& az, int& cnt) {
    const u64 x2 = tk_add2(d2, tk_pack(0.0001f, 0.0001f));
    float x0, x1;
    tk_unpack(x2, x0, x1);
    u64 s;
    if (UNIFORM) {
        // s = fv * (rep * e / dist - att / Reff), e = exp2(c2 * x); packed except the two MUFUs
        float fv0 = *reinterpret_cast<const float*>(tab_i + t0);
        float fv1 = *reinterpret_cast<const float*>(tab_i + t1);
        float t0f, t1f;
        tk_unpack(tk_mul2(x2, tk_pack(c2u, c2u)), t0f, t1f);
        const u64 er = tk_mul2(tk_pack(cf_ex2(t0f), cf_ex2(t1f)), tk_pack(cf_rsqrt(x0), cf_rsqrt(x1)));
        const u64 u = tk_fma2(er, tk_pack(repulsion, repulsion), tk_pack(-bu, -bu));
        s = tk_mul2(u, tk_pack(ok0 ? fv0 : 0.f, ok1 ? fv1 : 0.f)); // rejected pairs: exact 0
    } else {
        // table entry = (fv, 1/Reff, cut2, -): the exact accept test happens here
        float4 p0 = *reinterpret_cast<const float4*>(tab_i + t0);
        float4 p1 = *reinterpret_cast<const float4*>(tab_i + t1);
        float a0, a1;
        tk_unpack(d2, a0, a1);
        ok0 = a0 < p0.z;
        ok1 = a1 < p1.z;
        const u64 inv = tk_pack(p0.y, p1.y);
        float t0f, t1f;
        tk_unpack(tk_mul2(x2, tk_mul2(tk_mul2(inv, inv), tk_pack(nk_log2e, nk_log2e))), t0f, t1f);
        const u64 er = tk_mul2(tk_pack(cf_ex2(t0f), cf_ex2(t1f)), tk_pack(cf_rsqrt(x0), cf_rsqrt(x1)));
        const u64 u = tk_fma2(er, tk_pack(repulsion, repulsion), tk_mul2(inv, tk_pack(-attraction, -attraction)));
        s = tk_mul2(u, tk_pack(ok0 ? p0.x : 0.f, ok1 ? p1.x : 0.f));
    }
    cnt += (ok0 ? 1 : 0) + (ok1 ? 1 : 0);
    ax = tk_fma2(s, dx, ax);
    ay = tk_fma2(s, dy, ay);
    az = tk_fma2(s, dz, az);
}

template <bool UNIFORM, bool WRAP, int K>
__device__ __forceinline__ void tkw_chunk(const float* __restrict__ xs, const float* __restrict__ ys,
                                          const float* __restrict__ zs, const float* __restrict__ hs,
                                          const int* __restrict__ ts, int nquads, const float (&npx)[TK_IPT],
                                          const float (&npy)[TK_IPT], const float (&npz)[TK_IPT],
                                          const float (&hi)[TK_IPT], const char* const (&tab_i)[TK_IPT],
                                          float sx, float sy, float sz, float cut, float c2u, float bu,
                                          const StepConst& c, u64 (&ax)[TK_IPT], u64 (&ay)[TK_IPT],
                                          u64 (&az)[TK_IPT], int (&cnt)[TK_IPT]) {
    const u64 sx2 = tk_pack(sx, sx), sy2 = tk_pack(sy, sy), sz2 = tk_pack(sz, sz);
#pragma unroll 1
    for (int q = 0; q < nquads; q++) {
        const float4 X = *reinterpret_cast<const float4*>(xs + 4 * q);
        const float4 Y = *reinterpret_cast<const float4*>(ys + 4 * q);
        const float4 Z = *reinterpret_cast<const float4*>(zs + 4 * q);
        const int4 Tq = *reinterpret_cast<const int4*>(ts + 4 * q);
        float4 H = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!UNIFORM) H = *reinterpret_cast<const float4*>(hs + 4 * q);
        const u64 xa = tk_pack(X.x, X.y), xb = tk_pack(X.z, X.w);
        const u64 ya = tk_pack(Y.x, Y.y), yb = tk_pack(Y.z, Y.w);
        const u64 za = tk_pack(Z.x, Z.y), zb = tk_pack(Z.z, Z.w);
#pragma unroll
        for (int k = 0; k < K; k++) {
            const u64 px = tk_pack(npx[k], npx[k]), py = tk_pack(npy[k], npy[k]), pz = tk_pack(npz[k], npz[k]);
            u64 dxa = tk_add2(xa, px), dxb = tk_add2(xb, px);
            u64 dya = tk_add2(ya, py), dyb = tk_add2(yb, py);
            u64 dza = tk_add2(za, pz), dzb = tk_add2(zb, pz);
            if (WRAP) {
                dxa = tk_add2(dxa, sx2), dxb = tk_add2(dxb, sx2);
                dya = tk_add2(dya, sy2), dyb = tk_add2(dyb, sy2);
                dza = tk_add2(dza, sz2), dzb = tk_add2(dzb, sz2);
            }
            const u64 d2a = tk_fma2(dza, dza, tk_fma2(dxa, dxa, tk_mul2(dya, dya)));
            const u64 d2b = tk_fma2(dzb, dzb, tk_fma2(dxb, dxb, tk_mul2(dyb, dyb)));
            float a0, a1, b0, b1;
            tk_unpack(d2a, a0, a1);
            tk_unpack(d2b, b0, b1);
            bool o0, o1, o2, o3;
            if (UNIFORM) {
                o0 = a0 < cut, o1 = a1 < cut, o2 = b0 < cut, o3 = b1 < cut;
            } else { // conservative bound (h_i + h_j)^2 >= cut2[ti][tj]; exact test in the force path
                float t0 = H.x + hi[k], t1 = H.y + hi[k], t2 = H.z + hi[k], t3 = H.w + hi[k];
                o0 = a0 < t0 * t0, o1 = a1 < t1 * t1, o2 = b0 < t2 * t2, o3 = b1 < t3 * t3;
            }
            if (__any_sync(0xffffffffu, o0 || o1 || o2 || o3)) {
                tkw_force_pair<UNIFORM>(dxa, dya, dza, d2a, o0, o1, tab_i[k], Tq.x, Tq.y, c2u, bu, c.repulsion,
                                        c.attraction, c.nk_log2e, ax[k], ay[k], az[k], cnt[k]);
                tkw_force_pair<UNIFORM>(dxb, dyb, dzb, d2b, o2, o3, tab_i[k], Tq.z, Tq.w, c2u, bu, c.repulsion,
                                        c.attraction, c.nk_log2e, ax[k], ay[k], az[k], cnt[k]);
            }
        }
    }
}

template <bool UNIFORM>
__global__ void __launch_bounds__(TKW_WARPS * 32, 4)
force_tile_kernel(const float4* __restrict__ pos4, const int* __restrict__ cell_start,
                  const int2* __restrict__ tiles, int* __restrict__ ctrl, float4* __restrict__ frc4,
                  StepConst c, const DeviceTables* __restrict__ tables, float radius_half_scale,
                  const float* __restrict__ half_radius) {
    __shared__ __align__(16) TkwShared<UNIFORM> sm;
    // transposed pair table: entry (tj, ti) so that lanes (different ti) read neighbouring words
    __shared__ __align__(16) float s_tab[CF_TT_MAX * (UNIFORM ? 1 : 4)];
    __shared__ float s_half[CF_T_MAX];
    (void)radius_half_scale;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = c.T;
    for (int i = tid; i < T * T; i += blockDim.x) {
        int ti = i / T, tj = i % T; // tables are [ti][tj]
        if (UNIFORM) {
            s_tab[tj * T + ti] = tables->force[i];
        } else {
            float* e = &s_tab[(tj * T + ti) * 4];
            e[0] = tables->force[i];
            e[1] = tables->inv_reff[i];
            e[2] = tables->cut2[i];
            e[3] = 0.f;
        }
    }
    if (tid < T) s_half[tid] = half_radius[tid];
    __syncthreads(); // the only block-level barrier: tables are read-only afterwards
    const int entry = UNIFORM ? 4 : 16;       // bytes per table entry
    const int row_bytes = T * entry;          // one tj row
    const int ntiles = ctrl[0];
    const int ny = c.dims[1], nz = c.dims[2];
    const float cut = c.cut2_uniform;
    const float c2u = c.nk_log2e * c.inv_reff_uniform * c.inv_reff_uniform;
    const float bu = c.attraction * c.inv_reff_uniform;
    float* const wx = &sm.x[warp][0][0];
    float* const wy = &sm.y[warp][0][0];
    float* const wz = &sm.z[warp][0][0];
    float* const wh = &sm.h[warp][0][0];
    int* const wt = &sm.t[warp][0][0];

    for (;;) {
        int tile = 0;
        if (lane == 0) tile = atomicAdd(&ctrl[1], 1);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= ntiles) break;
        const int2 tl = tiles[tile];
        const int cell = tl.x;
        const int cz = cell % nz, cy = (cell / nz) % ny, cx = cell / (nz * ny);
        const int i_begin = cell_start[cell] + tl.y * TK_TI;
        const int ni = min(cell_start[cell + 1] - i_begin, TK_TI);

        // ---- neighbour runs: lane r < 18 holds run r = (row r/2 of the 3x3 (x,y) rows, z segment r%2)
        int r_j0 = 0, r_j1 = 0;
        float r_sx = 0.f, r_sy = 0.f, r_sz = 0.f;
        if (lane < TK_MAX_RUNS) {
            int rho = lane >> 1, seg = lane & 1;
            int x = cx + rho / 3 - 1, y = cy + rho % 3 - 1;
            bool valid = true;
            if (c.periodic_x) {
                if (x < 0) { x = c.dims[0] - 1; r_sx = -c.W[0]; } else if (x >= c.dims[0]) { x = 0; r_sx = c.W[0]; }
            } else { // slab mode: i-cells are layers 1..dims-2, so x stays inside [0, dims-1]
                if (x < 0 || x >= c.dims[0]) valid = false;
                else if (x == 0) r_sx = c.gshift_lo;
                else if (x == c.dims[0] - 1) r_sx = c.gshift_hi;
            }
            if (y < 0) { y = ny - 1; r_sy = -c.W[1]; } else if (y >= ny) { y = 0; r_sy = c.W[1]; }
            int z0 = 0, z1 = 0;
            if (seg == 0) {
                z0 = max(cz - 1, 0);
                z1 = min(cz + 1, nz - 1);
            } else if (cz == 0) {
                z0 = z1 = nz - 1;
                r_sz = -c.W[2];
            } else if (cz == nz - 1) {
                z0 = z1 = 0;
                r_sz = c.W[2];
            } else {
                valid = false;
            }
            if (valid) {
                int row = (x * ny + y) * nz;
                r_j0 = cell_start[row + z0];
                r_j1 = cell_start[row + z1 + 1];
            }
        }

        // ---- my i particles: layer k holds slots i_begin + 32k + lane ----
        float npx[TK_IPT], npy[TK_IPT], npz[TK_IPT], hi[TK_IPT];
        const char* tab_i[TK_IPT];
        u64 ax[TK_IPT], ay[TK_IPT], az[TK_IPT];
        int cnt[TK_IPT], self_ok[TK_IPT];
        const int kmax = (ni + 31) >> 5;
#pragma unroll
        for (int k = 0; k < TK_IPT; k++) {
            int il = k * 32 + lane;
            bool v = il < ni;
            float4 p = v ? pos4[i_begin + il] : make_float4(-TK_FAR, -TK_FAR, -TK_FAR, 0.f);
            int ti = v ? (int)__float_as_uint(p.w) : 0;
            npx[k] = -p.x, npy[k] = -p.y, npz[k] = -p.z;
            hi[k] = v ? s_half[ti] : 0.f;
            tab_i[k] = reinterpret_cast<const char*>(s_tab) + ti * entry;
            ax[k] = ay[k] = az[k] = tk_pack(0.f, 0.f);
            cnt[k] = 0;
            self_ok[k] = (UNIFORM ? cut : s_tab[(ti * T + ti) * 4 + 2]) > 0.f ? 1 : 0;
        }

        // ---- stream the runs ----
        int buf = 0;
        for (int r = 0; r < TK_MAX_RUNS; r++) {
            const int j0 = __shfl_sync(0xffffffffu, r_j0, r), j1 = __shfl_sync(0xffffffffu, r_j1, r);
            if (j1 <= j0) continue;
            const float sx = __shfl_sync(0xffffffffu, r_sx, r), sy = __shfl_sync(0xffffffffu, r_sy, r),
                        sz = __shfl_sync(0xffffffffu, r_sz, r);
            const bool wrap = (sx != 0.f) || (sy != 0.f) || (sz != 0.f);
            // prefetch the first chunk of the run
            float4 q0 = make_float4(TK_FAR, TK_FAR, TK_FAR, 0.f), q1 = q0;
            if (j0 + lane < j1) q0 = pos4[j0 + lane];
            if (j0 + 32 + lane < j1) q1 = pos4[j0 + 32 + lane];
            for (int off = j0; off < j1; off += TKW_JC) {
                // publish the prefetched chunk
                float* bx = wx + buf * TKW_JC;
                float* by = wy + buf * TKW_JC;
                float* bz = wz + buf * TKW_JC;
                float* bh = wh + buf * TKW_JC;
                int* bt = wt + buf * TKW_JC;
                {
                    int t0 = (int)__float_as_uint(q0.w), t1 = (int)__float_as_uint(q1.w);
                    bx[lane] = q0.x, by[lane] = q0.y, bz[lane] = q0.z, bt[lane] = t0 * row_bytes;
                    bx[lane + 32] = q1.x, by[lane + 32] = q1.y, bz[lane + 32] = q1.z, bt[lane + 32] = t1 * row_bytes;
                    if (!UNIFORM) {
                        bh[lane] = off + lane < j1 ? s_half[t0] : 0.f;
                        bh[lane + 32] = off + 32 + lane < j1 ? s_half[t1] : 0.f;
                    }
                }
                __syncwarp();
                const int cntj = min(TKW_JC, j1 - off);
                // prefetch the next chunk while this one is processed
                const int noff = off + TKW_JC;
                q0 = make_float4(TK_FAR, TK_FAR, TK_FAR, 0.f);
                q1 = q0;
                if (noff + lane < j1) q0 = pos4[noff + lane];
                if (noff + 32 + lane < j1) q1 = pos4[noff + 32 + lane];
                const int nquads = (cntj + 3) >> 2;
#define TKW_CALL(WR, KK)                                                                                  \
    tkw_chunk<UNIFORM, WR, KK>(bx, by, bz, bh, bt, nquads, npx, npy, npz, hi, tab_i, sx, sy, sz, cut, c2u, bu, \
                               c, ax, ay, az, cnt)
                if (wrap) {
                    switch (kmax) {
                        case 1: TKW_CALL(true, 1); break;
                        case 2: TKW_CALL(true, 2); break;
                        case 3: TKW_CALL(true, 3); break;
                        default: TKW_CALL(true, 4); break;
                    }
                } else {
                    switch (kmax) {
                        case 1: TKW_CALL(false, 1); break;
                        case 2: TKW_CALL(false, 2); break;
                        case 3: TKW_CALL(false, 3); break;
                        default: TKW_CALL(false, 4); break;
                    }
                }
#undef TKW_CALL
                buf ^= 1; // the other buffer was last read one chunk ago by this same warp
            }
        }

        // ---- write back: the particle itself was tested too (d = 0, force term exactly 0) ----
#pragma unroll
        for (int k = 0; k < TK_IPT; k++) {
            int il = k * 32 + lane;
            if (il < ni) {
                float x0, x1, y0, y1, z0, z1;
                tk_unpack(ax[k], x0, x1);
                tk_unpack(ay[k], y0, y1);
                tk_unpack(az[k], z0, z1);
                frc4[i_begin + il] = make_float4(x0 + x1, y0 + y1, z0 + z1, __int_as_float(cnt[k] - self_ok[k]));
            }
        }
    }
}

// The tile kernels decide the minimum-image wrap per neighbour cell, which is only equivalent to
// the reference's per-pair test when every periodic axis has at least 4 cells; they pay off when
// the cells that hold the particles hold enough of them to fill warps.  `occupancy` is the number
// of particles a particle shares its cell with, averaged over PARTICLES (sum n_c^2 / n): for a
// uniform state that is n / ncell, for a clustered one it is what matters (most cells are empty,
// most particles sit in full ones).
static inline bool tile_kernel_applicable(double occupancy) { return occupancy >= 48.0; }

// sum over every `stride`-th cell of (particles in the cell)^2 (the host scales by `stride`: clusters
// span hundreds of cells, so a strided sample estimates the sum well and costs 1/stride on the
// huge, mostly empty grids of sparse states).  One launch, no memset, no copy: acc[0] accumulates,
// acc[1] is a ticket; the block that takes the last ticket publishes the total to `host_out`
// (pinned, mapped host memory) and resets both for the next step.
__global__ void cell_occupancy_kernel(const int* __restrict__ cell_start, int ncell, int stride,
                                      unsigned long long* __restrict__ acc, unsigned long long* host_out) {
    const long long c = (long long)(blockIdx.x * blockDim.x + threadIdx.x) * stride;
    unsigned long long v = 0;
    if (c < ncell) {
        const unsigned long long k = (unsigned long long)(cell_start[c + 1] - cell_start[c]);
        v = k * k;
    }
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) {
        atomicAdd(&acc[0], v);
        __threadfence(); // the contribution is visible before this block's ticket is taken
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long ticket = atomicAdd(&acc[1], 1ull);
        if (ticket == (unsigned long long)gridDim.x - 1ull) { // every other block has contributed
            __threadfence();
            const unsigned long long total = atomicExch(&acc[0], 0ull);
            acc[1] = 0ull;
            *reinterpret_cast<volatile unsigned long long*>(host_out) = total;
            __threadfence_system();
        }
    }
}
