// kernels_tile.cuh — tiled pair-force kernel for dense cells (the FP32-bound hot kernel).
//
// Work decomposition
//   tile  = up to TK_TI (512) particles of ONE cell ("i" side), handled by one 128-thread CTA;
//           every thread keeps TK_IPT = 4 i-particles and their force/count accumulators in
//           registers.  Tiles come from a device-built list, fetched with an atomic counter by a
//           persistent grid (clustered states make cells wildly uneven: work-based scheduling).
//   j side = the 27 neighbour cells as <= 18 contiguous runs of the sorted array (cells that are
//           adjacent in z are adjacent in memory), streamed through double-buffered shared
//           memory in chunks of 128 particles, SoA (xs, ys, zs, type[, half-radius]).
//
// Why it looks like this (measured on B200, profiles/r01_pipe_microbench.txt)
//   * shared->register bandwidth is one 32-bit word per lane per clock per SM: an LDS.128
//     broadcast costs ~4 SM-cycles.  With one i per lane the j stream alone would need 2x the
//     cycles of the arithmetic, hence 4 i per lane (each loaded j is used 4x32 times).
//   * FFMA2/FADD2/FMUL2 (fma.rn.f32x2 ...) run at the same lane rate as the scalar forms but
//     take half the issue slots, which lets compares, mask updates and LDS issue for free
//     beside a saturated FMA pipe.  j particles are processed in packed pairs (x0,x1).
//   * ~85% of tested pairs are out of range.  The test phase only records accept bits
//     (one 64-bit mask per i per 64 j); the force terms are evaluated afterwards for the set
//     bits only, so the expensive path never runs predicated-off for rejected pairs.
//
// Exactness: displacement = (jx + (-px)) [+ s], s in {-W, 0, +W} per run.  For a periodic axis
// with >= 4 cells the reference's two-sided wrap test (.cu:97-98) is decided by which neighbour
// cell j lives in, and adding +-W is exact (Sterbenz), so this is bit-identical to the
// reference's d; d2 = fma(dz,dz,fma(dx,dx,dy*dy)) as compiled; accept <=> d2 < cut2[ti][tj].
#pragma once
#include "cf_device.cuh"
#include "kernels_force.cuh"

#define TK_THREADS 128
#define TK_IPT 4
#define TK_TI (TK_THREADS * TK_IPT)
#define TK_JC 128   // j particles per staged chunk
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

struct TileRun {
    int j0, j1;       // slot range of the run
    float sx, sy, sz; // exact minimum-image shift of the run
    int wrap;         // any shift non-zero
};

// Per type pair: A = repulsion * fv, B = attraction * fv / Reff, c2 = -k log2(e) / Reff^2, cut2.
// s = fv * net / dist = A * e / dist - B  with  e = exp2(c2 * (d2 + 1e-4)).
struct TilePairConst {
    float A, B, c2, cut2;
};

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

// Tile list: one thread per cell appends ceil(n_cell / TK_TI) tiles.  Tile order is arbitrary
// (results do not depend on it); ctrl[0] = number of tiles, ctrl[1] = fetch counter.
__global__ void build_tiles_kernel(const int* __restrict__ cell_start, int ncell, int first_cell_x,
                                   int last_cell_x, int cells_per_x, int2* __restrict__ tiles,
                                   int* __restrict__ ctrl) {
    int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= ncell) return;
    int cx = cell / cells_per_x;
    if (cx < first_cell_x || cx > last_cell_x) return; // ghost layers are never i-cells
    int n = cell_start[cell + 1] - cell_start[cell];
    if (n <= 0) return;
    int nt = (n + TK_TI - 1) / TK_TI;
    int base = atomicAdd(&ctrl[0], nt);
    for (int k = 0; k < nt; k++) tiles[base + k] = make_int2(cell, k);
}

// One 32-j word of the test phase for K i-particles per lane: 8 quads of staged j, each loaded
// once (3 LDS.128) and tested against all K register-resident i in packed pairs -> K x 32 bits.
template <int K, bool WRAP, bool UNIFORM>
__device__ __forceinline__ void tk_test_word(const float* __restrict__ xq, const float* __restrict__ yq,
                                             const float* __restrict__ zq, const float* __restrict__ hq,
                                             const u64 (&npx)[TK_IPT], const u64 (&npy)[TK_IPT],
                                             const u64 (&npz)[TK_IPT], const u64 (&hi2)[TK_IPT], u64 sx,
                                             u64 sy, u64 sz, float cut, unsigned (&m)[TK_IPT]) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const float4 X = *reinterpret_cast<const float4*>(xq + 4 * q);
        const float4 Y = *reinterpret_cast<const float4*>(yq + 4 * q);
        const float4 Z = *reinterpret_cast<const float4*>(zq + 4 * q);
        u64 xa = tk_pack(X.x, X.y), xb = tk_pack(X.z, X.w);
        u64 ya = tk_pack(Y.x, Y.y), yb = tk_pack(Y.z, Y.w);
        u64 za = tk_pack(Z.x, Z.y), zb = tk_pack(Z.z, Z.w);
        u64 ha = 0, hb = 0;
        if (!UNIFORM) {
            const float4 H = *reinterpret_cast<const float4*>(hq + 4 * q);
            ha = tk_pack(H.x, H.y), hb = tk_pack(H.z, H.w);
        }
        if (WRAP) { // shift the j side once per quad: (jx + s) is NOT what the reference rounds,
                    // so the shift is applied after the subtraction below, per k
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            u64 dxa = tk_add2(xa, npx[k]), dxb = tk_add2(xb, npx[k]);
            u64 dya = tk_add2(ya, npy[k]), dyb = tk_add2(yb, npy[k]);
            u64 dza = tk_add2(za, npz[k]), dzb = tk_add2(zb, npz[k]);
            if (WRAP) {
                dxa = tk_add2(dxa, sx), dxb = tk_add2(dxb, sx);
                dya = tk_add2(dya, sy), dyb = tk_add2(dyb, sy);
                dza = tk_add2(dza, sz), dzb = tk_add2(dzb, sz);
            }
            u64 d2a = tk_fma2(dza, dza, tk_fma2(dxa, dxa, tk_mul2(dya, dya)));
            u64 d2b = tk_fma2(dzb, dzb, tk_fma2(dxb, dxb, tk_mul2(dyb, dyb)));
            float a0, a1, b0, b1;
            tk_unpack(d2a, a0, a1);
            tk_unpack(d2b, b0, b1);
            float t0 = cut, t1 = cut, t2 = cut, t3 = cut;
            if (!UNIFORM) {
                // conservative per-pair bound (h_i + h_j)^2 >= cut2[ti][tj]; exact test in the force phase
                u64 ta = tk_add2(ha, hi2[k]), tb = tk_add2(hb, hi2[k]);
                ta = tk_mul2(ta, ta), tb = tk_mul2(tb, tb);
                tk_unpack(ta, t0, t1);
                tk_unpack(tb, t2, t3);
            }
            unsigned mk = m[k];
            if (a0 < t0) mk |= 1u << (4 * q);
            if (a1 < t1) mk |= 1u << (4 * q + 1);
            if (b0 < t2) mk |= 1u << (4 * q + 2);
            if (b1 < t3) mk |= 1u << (4 * q + 3);
            m[k] = mk;
        }
    }
}

// Test phase of one 64-j block for K i-layers: fills mask[k] (bit b = staged j o + b).
template <int K, bool UNIFORM>
__device__ __forceinline__ void tk_test_block(const float* xs, const float* ys, const float* zs, const float* hs,
                                              int nwords, bool wrap, const u64 (&npx)[TK_IPT],
                                              const u64 (&npy)[TK_IPT], const u64 (&npz)[TK_IPT],
                                              const u64 (&hi2)[TK_IPT], u64 sx, u64 sy, u64 sz, float cut,
                                              u64 (&mask)[TK_IPT]) {
#pragma unroll
    for (int k = 0; k < TK_IPT; k++) mask[k] = 0ull;
#pragma unroll 1
    for (int w = 0; w < nwords; w++) {
        unsigned m[TK_IPT];
#pragma unroll
        for (int k = 0; k < TK_IPT; k++) m[k] = 0u;
        const int o = 32 * w;
        if (wrap)
            tk_test_word<K, true, UNIFORM>(xs + o, ys + o, zs + o, hs + o, npx, npy, npz, hi2, sx, sy, sz, cut, m);
        else
            tk_test_word<K, false, UNIFORM>(xs + o, ys + o, zs + o, hs + o, npx, npy, npz, hi2, sx, sy, sz, cut, m);
#pragma unroll
        for (int k = 0; k < K; k++) mask[k] |= (u64)m[k] << o;
    }
}

template <bool UNIFORM>
__global__ void __launch_bounds__(TK_THREADS, 4)
force_tile_kernel(const float4* __restrict__ pos4, const int* __restrict__ cell_start,
                  const int2* __restrict__ tiles, int* __restrict__ ctrl, float4* __restrict__ frc4,
                  StepConst c, const DeviceTables* __restrict__ tables, float radius_half_scale,
                  const float* __restrict__ half_radius) {
    __shared__ __align__(16) float s_x[2][TK_JC];
    __shared__ __align__(16) float s_y[2][TK_JC];
    __shared__ __align__(16) float s_z[2][TK_JC];
    __shared__ __align__(16) float s_h[2][TK_JC];
    __shared__ uint32_t s_t[2][TK_JC];
    __shared__ TilePairConst s_pc[CF_TT_MAX];
    __shared__ float s_half[CF_T_MAX];
    __shared__ TileRun s_runs[TK_MAX_RUNS];
    __shared__ int s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = c.T;
    for (int i = tid; i < T * T; i += TK_THREADS) {
        float inv = tables->inv_reff[i];
        TilePairConst pc;
        pc.A = c.repulsion * tables->force[i];
        pc.B = c.attraction * tables->force[i] * inv;
        pc.c2 = c.nk_log2e * inv * inv;
        pc.cut2 = tables->cut2[i];
        s_pc[i] = pc;
    }
    if (tid < T) s_half[tid] = half_radius[tid];
    (void)radius_half_scale;
    const int ntiles = ctrl[0];
    const int ny = c.dims[1], nz = c.dims[2];

    for (;;) {
        __syncthreads(); // previous tile fully done (smem reuse), tables visible
        if (tid == 0) s_tile = atomicAdd(&ctrl[1], 1);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= ntiles) break;
        const int2 tl = tiles[tile];
        const int cell = tl.x;
        const int cz = cell % nz, cy = (cell / nz) % ny, cx = cell / (nz * ny);
        const int i_begin = cell_start[cell] + tl.y * TK_TI;
        const int i_end = min(cell_start[cell + 1], i_begin + TK_TI);
        const int ni = i_end - i_begin;

        // ---- neighbour runs: 9 (x, y) rows x {main z segment, wrapped z segment} ----------
        if (tid < TK_MAX_RUNS) {
            int rho = tid >> 1, seg = tid & 1;
            int ddx = rho / 3 - 1, ddy = rho % 3 - 1;
            int x = cx + ddx, y = cy + ddy;
            float sx = 0.f, sy = 0.f, sz = 0.f;
            bool valid = true;
            if (c.periodic_x) {
                if (x < 0) { x = c.dims[0] - 1; sx = -c.W[0]; } else if (x >= c.dims[0]) { x = 0; sx = c.W[0]; }
            } else { // slab mode: i-cells are layers 1..dims-2, so x stays inside [0, dims-1]
                if (x < 0 || x >= c.dims[0]) valid = false;
                else if (x == 0) sx = c.gshift_lo;
                else if (x == c.dims[0] - 1) sx = c.gshift_hi;
            }
            if (y < 0) { y = ny - 1; sy = -c.W[1]; } else if (y >= ny) { y = 0; sy = c.W[1]; }
            int z0, z1;
            if (seg == 0) {
                z0 = max(cz - 1, 0);
                z1 = min(cz + 1, nz - 1);
            } else if (cz == 0) {
                z0 = z1 = nz - 1;
                sz = -c.W[2];
            } else if (cz == nz - 1) {
                z0 = z1 = 0;
                sz = c.W[2];
            } else {
                valid = false;
                z0 = z1 = 0;
            }
            TileRun r;
            r.j0 = r.j1 = 0;
            if (valid) {
                int row = (x * ny + y) * nz;
                r.j0 = cell_start[row + z0];
                r.j1 = cell_start[row + z1 + 1];
            }
            r.sx = sx, r.sy = sy, r.sz = sz;
            r.wrap = (sx != 0.f || sy != 0.f || sz != 0.f) ? 1 : 0;
            s_runs[tid] = r;
        }

        // ---- my i particles: slot = i_begin + warp*128 + k*32 + lane (fills whole warps first) ----
        u64 npx[TK_IPT], npy[TK_IPT], npz[TK_IPT], hi2[TK_IPT];
        float fx[TK_IPT], fy[TK_IPT], fz[TK_IPT];
        const float cut = c.cut2_uniform;
        int cnt[TK_IPT], ti[TK_IPT];
        const int warp_first = warp * (32 * TK_IPT);
        int kmax = 0; // i-layers this warp actually holds (warp-uniform)
#pragma unroll
        for (int k = 0; k < TK_IPT; k++) {
            int il = warp_first + k * 32 + lane;
            bool v = il < ni;
            float4 p = v ? pos4[i_begin + il] : make_float4(-TK_FAR, -TK_FAR, -TK_FAR, 0.f);
            ti[k] = v ? (int)__float_as_uint(p.w) : 0;
            npx[k] = tk_pack(-p.x, -p.x);
            npy[k] = tk_pack(-p.y, -p.y);
            npz[k] = tk_pack(-p.z, -p.z);
            float h = v ? s_half[ti[k]] : 0.f;
            hi2[k] = tk_pack(h, h);
            fx[k] = fy[k] = fz[k] = 0.f;
            cnt[k] = 0;
            if (warp_first + k * 32 < ni) kmax = k + 1;
        }
        __syncthreads(); // runs visible (s_half was synced before the loop's first barrier)

        // ---- chunk stream over the runs, double-buffered ------------------------------------
        int run = 0, off = 0;
        // advance to the first non-empty run
        while (run < TK_MAX_RUNS && s_runs[run].j1 - s_runs[run].j0 <= 0) run++;
        // stage the first chunk
        auto stage = [&](int buf, int r, int o) {
            int j = s_runs[r].j0 + o + tid;
            bool v = j < s_runs[r].j1;
            float4 q = v ? pos4[j] : make_float4(TK_FAR, TK_FAR, TK_FAR, 0.f);
            s_x[buf][tid] = q.x;
            s_y[buf][tid] = q.y;
            s_z[buf][tid] = q.z;
            uint32_t tj = v ? __float_as_uint(q.w) : 0u;
            s_t[buf][tid] = tj;
            if (!UNIFORM) s_h[buf][tid] = v ? s_half[tj] : 0.f;
        };
        int buf = 0;
        if (run < TK_MAX_RUNS) stage(0, run, 0);
        __syncthreads();
        while (run < TK_MAX_RUNS) {
            const TileRun R = s_runs[run];
            const int cntj = min(TK_JC, R.j1 - R.j0 - off);
            // next chunk coordinates (uniform)
            int nrun = run, noff = off + TK_JC;
            if (noff >= R.j1 - R.j0) {
                nrun = run + 1;
                noff = 0;
                while (nrun < TK_MAX_RUNS && s_runs[nrun].j1 - s_runs[nrun].j0 <= 0) nrun++;
            }
            // prefetch the next chunk's particle into registers (latency hidden by the compute)
            float4 nq = make_float4(TK_FAR, TK_FAR, TK_FAR, 0.f);
            bool nv = false;
            if (nrun < TK_MAX_RUNS) {
                int j = s_runs[nrun].j0 + noff + tid;
                nv = j < s_runs[nrun].j1;
                if (nv) nq = pos4[j];
            }

            if (kmax > 0) {
                const u64 sx2 = tk_pack(R.sx, R.sx), sy2 = tk_pack(R.sy, R.sy), sz2 = tk_pack(R.sz, R.sz);
                const int nblk = (cntj + 63) >> 6; // 64-j blocks in this chunk
                for (int blk = 0; blk < nblk; blk++) {
                    const int o = blk * 64;
                    u64 mask[TK_IPT];
                    // ---------------- test phase ----------------
                    const int nwords = cntj > o + 32 ? 2 : 1;
                    const float* xs = &s_x[buf][o];
                    const float* ys = &s_y[buf][o];
                    const float* zs = &s_z[buf][o];
                    const float* hs = &s_h[buf][o];
                    switch (kmax) { // warp-uniform
                        case 1: tk_test_block<1, UNIFORM>(xs, ys, zs, hs, nwords, R.wrap != 0, npx, npy, npz, hi2, sx2, sy2, sz2, cut, mask); break;
                        case 2: tk_test_block<2, UNIFORM>(xs, ys, zs, hs, nwords, R.wrap != 0, npx, npy, npz, hi2, sx2, sy2, sz2, cut, mask); break;
                        case 3: tk_test_block<3, UNIFORM>(xs, ys, zs, hs, nwords, R.wrap != 0, npx, npy, npz, hi2, sx2, sy2, sz2, cut, mask); break;
                        default: tk_test_block<4, UNIFORM>(xs, ys, zs, hs, nwords, R.wrap != 0, npx, npy, npz, hi2, sx2, sy2, sz2, cut, mask); break;
                    }
                    if (UNIFORM) {
#pragma unroll
                        for (int k = 0; k < TK_IPT; k++) cnt[k] += __popcll(mask[k]);
                    }
                    // ---------------- force phase: set bits only ----------------
#pragma unroll
                    for (int k = 0; k < TK_IPT; k++) {
                        if (k < kmax) {
                            u64 m = mask[k];
                            float px, py, pz, dummy;
                            tk_unpack(npx[k], px, dummy);
                            tk_unpack(npy[k], py, dummy);
                            tk_unpack(npz[k], pz, dummy);
                            const TilePairConst* row = &s_pc[ti[k] * T];
                            while (__any_sync(0xffffffffu, m != 0ull)) {
                                if (m != 0ull) {
                                    int b = __ffsll((long long)m) - 1;
                                    m &= m - 1ull;
                                    int j = o + b;
                                    float dx = __fadd_rn(__fadd_rn(s_x[buf][j], px), R.sx);
                                    float dy = __fadd_rn(__fadd_rn(s_y[buf][j], py), R.sy);
                                    float dz = __fadd_rn(__fadd_rn(s_z[buf][j], pz), R.sz);
                                    float d2 = cf_dist2(dx, dy, dz);
                                    TilePairConst pc = row[s_t[buf][j]];
                                    bool ok = UNIFORM ? true : (d2 < pc.cut2);
                                    if (ok) {
                                        if (!UNIFORM) cnt[k]++;
                                        float x = __fadd_rn(d2, 0.0001f);
                                        float rinv = cf_rsqrt(x);
                                        float e = cf_ex2(x * pc.c2);
                                        float s = fmaf(e * pc.A, rinv, -pc.B);
                                        fx[k] = fmaf(s, dx, fx[k]);
                                        fy[k] = fmaf(s, dy, fy[k]);
                                        fz[k] = fmaf(s, dz, fz[k]);
                                    }
                                }
                            }
                        }
                    }
                }
            }

            // publish the prefetched chunk into the other buffer
            if (nrun < TK_MAX_RUNS) {
                int nb = buf ^ 1;
                s_x[nb][tid] = nq.x;
                s_y[nb][tid] = nq.y;
                s_z[nb][tid] = nq.z;
                uint32_t tj = nv ? __float_as_uint(nq.w) : 0u;
                s_t[nb][tid] = tj;
                if (!UNIFORM) s_h[nb][tid] = nv ? s_half[tj] : 0.f;
            }
            __syncthreads();
            buf ^= 1;
            run = nrun;
            off = noff;
        }

        // ---- write back: the particle itself was tested too (d = 0, force term exactly 0) ----
#pragma unroll
        for (int k = 0; k < TK_IPT; k++) {
            int il = warp_first + k * 32 + lane;
            if (il < ni) {
                int self = s_pc[ti[k] * T + ti[k]].cut2 > 0.f ? 1 : 0;
                frc4[i_begin + il] = make_float4(fx[k], fy[k], fz[k], __int_as_float(cnt[k] - self));
            }
        }
    }
}

// The tile kernel decides the minimum-image wrap per neighbour cell, which is only equivalent to
// the reference's per-pair test when every periodic axis has at least 4 cells; it pays off when
// cells hold enough particles to fill warps.
static inline bool tile_kernel_applicable(const StepConst&, int n, int ncell) {
    return (double)n / (double)ncell >= 48.0;
}
