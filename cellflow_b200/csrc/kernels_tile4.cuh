// kernels_tile4.cuh — pair-force kernel, generation 4 (the FP32-bound hot kernel).
//
// Same decomposition as kernels_tile.cuh (one WARP owns a tile of 32 x IPT particles of one cell as
// IPT register-resident layers of 32; the 27 neighbour cells are streamed as <= 18 contiguous runs
// through a warp-private double buffer, 128 j per chunk; one warp vote per (layer, quad of 4 j)
// gates the force terms), with what the round-1 ncu captures asked for
// (profiles/r01_force_kernel_history.md):
//
//   * BOX PREFILTER.  The exact test of a (layer, quad) block costs 12 packed FP32 instructions
//     = 24 FMA-pipe cycles per SM sub-partition, and 50-70% of the blocks are dead.  A chunk is
//     128 j, one quad of 4 consecutive j per lane: each lane takes the bounding box of its
//     own quad (no shuffles) and tests it against the bounding box of each layer's 32 i (one
//     ballot per layer and chunk).  The quad loop walks the quads with a live layer only (dead quads cost
//     nothing); uniform radii (MODE 0) also skip the layers of a live quad whose own box is far,
//     per-type radii (MODE 1) test all of them (the extra branch cost more than it saved there).
//   * COMPACTED STAGING (STAGE 4, the default since the middle of round 2).  The prefilter runs on the
//     registers the chunk was loaded into, and only the quads it lets through are stored to shared
//     memory, packed to the front of the warp's buffer (rank = popc of the live mask below the lane).
//     A dead chunk stores nothing; the block loop of a live one is a running shared address — no bit
//     scan (BREV + FLO are XU-pipe instructions, the pipe the MUFU of the force terms need), 25
//     instead of 30 instructions per exact-tested block.
//   * PREDICATED ACCUMULATION.  The three FFMA and the count of a pair run under the predicate of its
//     accept test (T4AccScalar::add_if): 36 instead of 42 instructions per live block.
//   * TYPE-HOMOGENEOUS j RUNS for per-type radii (MODE 1).  The j stream comes from a second
//     copy of the positions sorted by (xy row, type, z cell, Hilbert): inside a sub-run every j
//     has the same type, so cut2 / force value / 1/Reff of a pair depend on the lane only and
//     are fetched once per sub-run instead of once per pair; the vote test is the exact accept
//     test.  The v3 kernel needed a float4 table gather and a second compare per evaluated pair.
//   * LEANER FORCE PATH.  A rejected, padded or overflowed pair touches neither the force nor the count
//     (predicates; -DT4_PRED=0: integer accept masks, set.lt.s32.f32 + AND on the bits of s), so its
//     NaN/Inf can never leak; no generic-pointer arithmetic on shared memory.
//   * REGISTERS: 64 / 8 CTAs per SM (1 layer per tile, per-type radii), 80 / 6 (2 layers, uniform radius; round 1: 4
//     layers, 96 registers, 5 CTAs): scalar force accumulators, layer boxes in shared memory, the
//     next chunk prefetched into L1 by a hint instead of into registers (and across run
//     boundaries); one code path for every tile occupancy (empty layers have empty boxes and
//     positions at 1e30, so they are never live).
//   * TAIL.  The tiles of the last, partly filled round of the persistent grid are handed out as
//     sub-tiles of 2 or 1 layers when that shortens the round.
//
// Exactness is unchanged: displacement = (jx + (-px)) [+ s], s in {-W, 0, +W} per run (exact,
// see kernels_tile.cuh), d2 = fma(dz,dz, fma(dx,dx, dy*dy)), accept <=> d2 < cut2[ti][tj].
// The box prefilter is conservative (margin 1e-4 relative + 0.01 absolute on the squared gap).
#pragma once
#include "kernels_tile.cuh"

// Layers of 32 i per tile (= i particles per lane) and resident CTAs per SM, per mode.  Round 1 used 4 layers
// (96 registers, 5 CTAs) to amortise the shared-memory reads of a quad over 4 blocks; with the box prefilter dead
// quads are never read (LSU 12-19 % busy), and the kernel is latency / issue bound at 5 warps per scheduler, so
// fewer layers = fewer registers = more resident warps wins (A/B profiles/r02_force_kernel_history.md):
//   per-type radii (MODE 1): 1 layer, 64 registers, 8 CTAs;  uniform radius (MODE 0): 2 layers, 80 registers, 6 CTAs.
#ifndef T4_IPT_MODE0
#define T4_IPT_MODE0 2
#endif
#ifndef T4_IPT_MODE1
#define T4_IPT_MODE1 1
#endif
__host__ __device__ constexpr int t4_ipt(int mode) { return mode ? T4_IPT_MODE1 : T4_IPT_MODE0; }
__host__ __device__ constexpr int t4_minb(int ipt) { return ipt <= 1 ? 8 : (ipt == 2 ? 6 : 5); }
#ifndef T4_PRED
#define T4_PRED 1 // accepted pairs accumulated under predicates (0: integer accept masks, the round-1 form)
#endif
#ifndef T4_UNROLL
#define T4_UNROLL 1
#endif
#ifndef T4_WARPS
#define T4_WARPS 4
#endif
#define T4_JC 128 // staged j per chunk and warp: one quad of 4 consecutive j per lane
#define T4_MAXSUB (TK_MAX_RUNS * CF_T_MAX)
#define T4_INF __int_as_float(0x7f800000)

template <int MODE, int IPT>
struct T4Shared {
    float x[T4_WARPS][2][T4_JC];
    float y[T4_WARPS][2][T4_JC];
    float z[T4_WARPS][2][T4_JC];
    int t[T4_WARPS][2][T4_JC];                       // MODE 0: byte offset of j's row in s_fv
    int2 sub[T4_WARPS][MODE ? T4_MAXSUB : TK_MAX_RUNS]; // [j0, j1) of every (sub-)run
    float4 cst[MODE ? T4_WARPS : 1][IPT][32];          // MODE 1: (c2, A, B, cut2) of (layer, lane) for the current sub-run
    float4 box[T4_WARPS][IPT][2];                      // (lo.xyz, prefilter threshold), (hi.xyz, -) of every layer
    unsigned long long bar[T4_WARPS][2];             // STAGE 1: one mbarrier per (warp, stage buffer)
};

// ---- bulk-copy staging (STAGE 1): cp.async.bulk global -> shared, completion on an mbarrier (SASS: UBLKCP) ----
__device__ __forceinline__ void t4_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void t4_mbar_expect(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void t4_bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ bool t4_mbar_try(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ int t4_setlt(float a, float b) { // 0xffffffff if a < b else 0
    int r;
    asm("set.lt.s32.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float t4_and(float v, int m) { return __int_as_float(__float_as_int(v) & m); }

// order-preserving float <-> uint map (for REDUX.MIN/MAX on floats)
__device__ __forceinline__ unsigned t4_f2o(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float t4_o2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ float t4_warp_min(float f) { return t4_o2f(__reduce_min_sync(0xffffffffu, t4_f2o(f))); }
__device__ __forceinline__ float t4_warp_max(float f) { return t4_o2f(__reduce_max_sync(0xffffffffu, t4_f2o(f))); }

__device__ __forceinline__ float t4_lds(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// Force terms of the 4 pairs of one (layer, quad) block for one lane.
//   s = fv * (rep * e / dist - att / Reff),  e = exp2(c2 * x),  x = d2 + 1e-4,  1/dist = rsqrt(x)
// MODE 0: c2, nb = -att/Reff are kernel constants, fv comes from the transposed table in shared
// memory; MODE 1: c2, A = fv * rep, B = -fv * att / Reff are per (lane, layer) constants of the
// current sub-run.  Rejected pairs: the bits of s are ANDed with the accept mask.
// Force accumulators of one (lane, layer): scalar (3 registers); the packed products of a block are
// folded in with 12 FFMA (same FMA-pipe cycles as 6 packed FFMA2 into 6 registers).
struct T4AccScalar {
    float x, y, z;
    __device__ __forceinline__ void zero() { x = y = z = 0.f; }
    __device__ __forceinline__ void add(u64 sa, u64 sb, u64 dxa, u64 dya, u64 dza, u64 dxb, u64 dyb, u64 dzb) {
        float s0, s1, s2, s3, a0, a1, b0, b1;
        tk_unpack(sa, s0, s1);
        tk_unpack(sb, s2, s3);
        tk_unpack(dxa, a0, a1);
        tk_unpack(dxb, b0, b1);
        x = fmaf(s3, b1, fmaf(s2, b0, fmaf(s1, a1, fmaf(s0, a0, x))));
        tk_unpack(dya, a0, a1);
        tk_unpack(dyb, b0, b1);
        y = fmaf(s3, b1, fmaf(s2, b0, fmaf(s1, a1, fmaf(s0, a0, y))));
        tk_unpack(dza, a0, a1);
        tk_unpack(dzb, b0, b1);
        z = fmaf(s3, b1, fmaf(s2, b0, fmaf(s1, a1, fmaf(s0, a0, z))));
    }
    // one pair under its accept predicate d2 < thr: force term and neighbour count (a rejected, padded or
    // overflowed pair touches nothing)
    __device__ __forceinline__ void add_if(float d2, float thr, float s, float dx, float dy, float dz, int& cnt) {
        asm("{\n\t"
            ".reg .pred p;\n\t"
            "setp.lt.f32 p, %4, %5;\n\t"
            "@p fma.rn.f32 %0, %6, %7, %0;\n\t"
            "@p fma.rn.f32 %1, %6, %8, %1;\n\t"
            "@p fma.rn.f32 %2, %6, %9, %2;\n\t"
            "@p add.s32 %3, %3, 1;\n\t"
            "}"
            : "+f"(x), "+f"(y), "+f"(z), "+r"(cnt)
            : "f"(d2), "f"(thr), "f"(s), "f"(dx), "f"(dy), "f"(dz));
    }
    __device__ __forceinline__ float3 sum() const { return make_float3(x, y, z); }
};

template <int MODE>
__device__ __forceinline__ void t4_live(u64 dxa, u64 dya, u64 dza, u64 d2a, u64 dxb, u64 dyb, u64 dzb, u64 d2b,
                                        float a0, float a1, float b0, float b1, float thr, float c2, float pa,
                                        float pb, unsigned fv_base, int4 Tq, T4AccScalar& acc, int& cnt) {
    const u64 eps = tk_pack(0.0001f, 0.0001f);
    const u64 xa = tk_add2(d2a, eps), xb = tk_add2(d2b, eps);
    const u64 c22 = tk_pack(c2, c2);
    float ta0, ta1, tb0, tb1, xa0, xa1, xb0, xb1;
    tk_unpack(tk_mul2(xa, c22), ta0, ta1);
    tk_unpack(tk_mul2(xb, c22), tb0, tb1);
    tk_unpack(xa, xa0, xa1);
    tk_unpack(xb, xb0, xb1);
    const u64 era = tk_mul2(tk_pack(cf_ex2(ta0), cf_ex2(ta1)), tk_pack(cf_rsqrt(xa0), cf_rsqrt(xa1)));
    const u64 erb = tk_mul2(tk_pack(cf_ex2(tb0), cf_ex2(tb1)), tk_pack(cf_rsqrt(xb0), cf_rsqrt(xb1)));
    u64 sa = tk_fma2(era, tk_pack(pa, pa), tk_pack(pb, pb));
    u64 sb = tk_fma2(erb, tk_pack(pa, pa), tk_pack(pb, pb));
    if (MODE == 0) {
        const float f0 = t4_lds(fv_base + Tq.x), f1 = t4_lds(fv_base + Tq.y);
        const float f2 = t4_lds(fv_base + Tq.z), f3 = t4_lds(fv_base + Tq.w);
        sa = tk_mul2(sa, tk_pack(f0, f1));
        sb = tk_mul2(sb, tk_pack(f2, f3));
    }
    float s0, s1, s2, s3;
    tk_unpack(sa, s0, s1);
    tk_unpack(sb, s2, s3);
#if T4_PRED
    // accepted pairs only: the three FFMA and the count of a pair run under the predicate of its accept test
    // (4 FSETP + 12 @P FFMA + 4 @P IADD; the mask form below needs 4 SEL + 4 LOP3 + 2 IADD3 more per block)
    float x0, x1, x2, x3;
    tk_unpack(dxa, x0, x1), tk_unpack(dxb, x2, x3);
    float y0, y1, y2, y3;
    tk_unpack(dya, y0, y1), tk_unpack(dyb, y2, y3);
    float z0, z1, z2, z3;
    tk_unpack(dza, z0, z1), tk_unpack(dzb, z2, z3);
    acc.add_if(a0, thr, s0, x0, y0, z0, cnt);
    acc.add_if(a1, thr, s1, x1, y1, z1, cnt);
    acc.add_if(b0, thr, s2, x2, y2, z2, cnt);
    acc.add_if(b1, thr, s3, x3, y3, z3, cnt);
#else
    const int m0 = t4_setlt(a0, thr), m1 = t4_setlt(a1, thr), m2 = t4_setlt(b0, thr), m3 = t4_setlt(b1, thr);
    sa = tk_pack(t4_and(s0, m0), t4_and(s1, m1));
    sb = tk_pack(t4_and(s2, m2), t4_and(s3, m3));
    cnt -= (m0 + m1) + (m2 + m3);
    acc.add(sa, sb, dxa, dya, dza, dxb, dyb, dzb);
#endif
}

// All live (layer, quad) blocks of one staged chunk.  MODE 0 takes the uniform threshold and derives
// the table address from the packed types; MODE 1 reads (c2, A, B, cut2) of (layer, lane) from
// shared memory per block.
template <int MODE, bool WRAP, bool COUNT, int IPT>
__device__ __forceinline__ void t4_chunk(unsigned sbase, const unsigned (&live)[IPT], const float (&npx)[IPT],
                                         const float (&npy)[IPT], const float (&npz)[IPT], unsigned tis4,
                                         unsigned s_tab_addr, unsigned cst_addr, float sx, float sy,
                                         float sz, float cutu, float c2u, float pau, float pbu,
                                         T4AccScalar (&acc)[IPT], int (&cnt)[IPT], unsigned& n_tested,
                                         unsigned& n_live) {
    const u64 sx2 = tk_pack(sx, sx), sy2 = tk_pack(sy, sy), sz2 = tk_pack(sz, sz);
    constexpr unsigned STRIDE = T4_WARPS * 2 * T4_JC * 4; // bytes between the x, y, z, t arrays
    // live[k] bit q = (quad q, layer k) passed the box prefilter; quads with no live layer cost nothing
    unsigned any = 0;
#pragma unroll
    for (int k = 0; k < IPT; k++) any |= live[k];
#pragma unroll 1
    while (any) {
        const unsigned bit = any & (0u - any);
        any ^= bit;
        const unsigned qa = sbase + 16u * (unsigned)(__ffs(bit) - 1);
        float4 X, Y, Z;
        int4 Tq = make_int4(0, 0, 0, 0);
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(X.x), "=f"(X.y), "=f"(X.z), "=f"(X.w) : "r"(qa));
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(Y.x), "=f"(Y.y), "=f"(Y.z), "=f"(Y.w) : "r"(qa + STRIDE));
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(Z.x), "=f"(Z.y), "=f"(Z.z), "=f"(Z.w) : "r"(qa + 2 * STRIDE));
        if (MODE == 0)
            asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(Tq.x), "=r"(Tq.y), "=r"(Tq.z), "=r"(Tq.w) : "r"(qa + 3 * STRIDE));
        const u64 xa = tk_pack(X.x, X.y), xb = tk_pack(X.z, X.w);
        const u64 ya = tk_pack(Y.x, Y.y), yb = tk_pack(Y.z, Y.w);
        const u64 za = tk_pack(Z.x, Z.y), zb = tk_pack(Z.z, Z.w);
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            // MODE 0: layers whose box is not near skip the exact test (warp-uniform branch).  MODE 1 tests
            // every layer of a live quad: there the extra branch costs more than the tests it saves
            // (measured: eater 4.89 -> 4.76 ms without it, pulser 1.96 -> 2.30 ms without it)
            if (MODE == 0 && !(live[k] & bit)) continue;
            float4 cs = make_float4(c2u, pau, pbu, cutu);
            if (MODE == 1) // (c2, A, B, cut2) of (layer k, this lane): 32-bit shared address, no generic pointer
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(cs.x), "=f"(cs.y), "=f"(cs.z), "=f"(cs.w) : "r"(cst_addr + 512u * k));
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
            const float mn = fminf(fminf(a0, a1), fminf(b0, b1));
            if (COUNT) n_tested++; // instrumented build (option "count_blocks"): (layer, quad) blocks exact-tested
            if (__any_sync(0xffffffffu, mn < cs.w)) {
                if (COUNT) n_live++; // ... and blocks whose 128 pair-lanes are evaluated
                const unsigned fva = s_tab_addr + ((tis4 >> (8 * k)) & 255u);
                t4_live<MODE>(dxa, dya, dza, d2a, dxb, dyb, dzb, d2b, a0, a1, b0, b1, cs.w, cs.x, cs.y, cs.z, fva, Tq,
                              acc[k], cnt[k]);
            }
        }
    }
}

// The same blocks, for a chunk whose LIVE quads were stored COMPACTED (STAGE 4): quad r of the list sits at
// sbase + 16 r, r = 0 .. nq-1, in lane order (= the order t4_chunk walks them, so the summation order and every
// result bit are the same).  The loop is a running shared address: no bit scan (BREV + FLO are XU-pipe
// instructions, 16 cycles of the pipe the 8 MUFU of a live block need), no address arithmetic.
// clive[k] bit r = layer k passed the box prefilter for quad r (only read by MODE 0 with more than one layer).
template <int MODE, bool WRAP, bool COUNT, int IPT>
__device__ __forceinline__ void t4_chunk_compact(unsigned sbase, unsigned nq, const unsigned (&clive)[IPT],
                                                 const float (&npx)[IPT], const float (&npy)[IPT],
                                                 const float (&npz)[IPT], unsigned tis4, unsigned s_tab_addr,
                                                 unsigned cst_addr, float sx, float sy, float sz, float cutu, float c2u,
                                                 float pau, float pbu, T4AccScalar (&acc)[IPT], int (&cnt)[IPT],
                                                 unsigned& n_tested, unsigned& n_live) {
    const u64 sx2 = tk_pack(sx, sx), sy2 = tk_pack(sy, sy), sz2 = tk_pack(sz, sz);
    constexpr unsigned STRIDE = T4_WARPS * 2 * T4_JC * 4; // bytes between the x, y, z, t arrays
    unsigned qa = sbase;
    const unsigned qe = sbase + 16u * nq;
    unsigned bit = 1u;
    constexpr int UNROLL = T4_UNROLL;
#pragma unroll UNROLL
    do {
        float4 X, Y, Z;
        int4 Tq = make_int4(0, 0, 0, 0);
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(X.x), "=f"(X.y), "=f"(X.z), "=f"(X.w) : "r"(qa));
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(Y.x), "=f"(Y.y), "=f"(Y.z), "=f"(Y.w) : "r"(qa + STRIDE));
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(Z.x), "=f"(Z.y), "=f"(Z.z), "=f"(Z.w) : "r"(qa + 2 * STRIDE));
        if (MODE == 0)
            asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(Tq.x), "=r"(Tq.y), "=r"(Tq.z), "=r"(Tq.w) : "r"(qa + 3 * STRIDE));
        const u64 xa = tk_pack(X.x, X.y), xb = tk_pack(X.z, X.w);
        const u64 ya = tk_pack(Y.x, Y.y), yb = tk_pack(Y.z, Y.w);
        const u64 za = tk_pack(Z.x, Z.y), zb = tk_pack(Z.z, Z.w);
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            if (MODE == 0 && IPT > 1 && !(clive[k] & bit)) continue; // as t4_chunk: MODE 1 tests every layer of a live quad
            float4 cs = make_float4(c2u, pau, pbu, cutu);
            if (MODE == 1)
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(cs.x), "=f"(cs.y), "=f"(cs.z), "=f"(cs.w) : "r"(cst_addr + 512u * k));
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
            const float mn = fminf(fminf(a0, a1), fminf(b0, b1));
            if (COUNT) n_tested++;
            if (__any_sync(0xffffffffu, mn < cs.w)) {
                if (COUNT) n_live++;
                const unsigned fva = s_tab_addr + ((tis4 >> (8 * k)) & 255u);
                t4_live<MODE>(dxa, dya, dza, d2a, dxb, dyb, dzb, d2b, a0, a1, b0, b1, cs.w, cs.x, cs.y, cs.z, fva, Tq,
                              acc[k], cnt[k]);
            }
        }
        qa += 16u;
        if (MODE == 0 && IPT > 1) bit <<= 1;
    } while (qa != qe);
}

// STAGE 0: the j chunk goes global -> registers (4 x LDG.128 per lane) -> shared (3-4 x STS.128, transposed to
//          SoA), next chunk hinted into L1.
// STAGE 4: as STAGE 0, but the box prefilter runs on the registers BEFORE anything is stored, and only the quads
//          it lets through are stored, compacted to the front of the staging buffer (rank = popc of the live mask
//          below the lane): a dead chunk stores nothing, the block loop of a live one is a running address
//          (t4_chunk_compact).  Loads use clamped indices (no per-element predicates, no default moves; the quad
//          box of a partly valid quad is unaffected by the repeated last element), the invalid elements of the one
//          partly valid quad are pushed out of range just before the store.  Bit-identical to STAGE 0.
// STAGE 1 (MODE 1 only): the type-sorted copy also exists as SoA planes (jx, jy, jz); one elected lane brings the
//          next 128-j chunk in with three 512-byte cp.async.bulk copies that complete on the warp's mbarrier
//          (UBLKCP), no register round trip, no transpose; chunk starts are aligned down to 4 elements (16 bytes)
//          and the elements in front of the sub-run are masked.  A/B in profiles/r02_staging_ab.md.
// STAGE 3 (MODE 1 only): the SoA planes again, but through registers: 3 x LDG.128 (4 consecutive x, y, z of the
//          lane's quad, already in staging layout) -> quad box -> 3 x STS.128.  No transposition (the default's
//          4 x LDG.128 of (x,y,z,type) records needs 12 register moves), and a chunk that lies entirely inside its
//          sub-run takes a mask-free fast path: ~35 instead of 118 instructions per chunk.
// STAGE 2: the bounding boxes of all aligned quads of the j array are computed ONCE per step (quad_box_kernel,
//          32 B per 4 j); a chunk then starts with one 32-byte load per lane and the box prefilter, and only the
//          lanes whose quad survived load and stage their 4 positions — a dead chunk (about half of them) costs
//          ~35 instructions instead of ~100, and its positions are never read.
template <int MODE, bool COUNT = false, int STAGE = 0, int IPT = t4_ipt(MODE)>
__global__ void __launch_bounds__(T4_WARPS * 32, t4_minb(IPT))
force_tile4_kernel(const float4* __restrict__ pos4, const int* __restrict__ cell_start,
                   const float4* __restrict__ posj, const int* __restrict__ startj,
                   const int2* __restrict__ tiles, int* __restrict__ ctrl, float4* __restrict__ frc4, StepConst c,
                   const DeviceTables* __restrict__ tables, unsigned long long* __restrict__ block_counts = nullptr,
                   const float* __restrict__ jx = nullptr, const float* __restrict__ jy = nullptr,
                   const float* __restrict__ jz = nullptr, const float4* __restrict__ qbox = nullptr) {
    __shared__ __align__(16) T4Shared<MODE, IPT> sm;
    constexpr int T4_TI = 32 * IPT;
    // MODE 0: s_tab[tj*T + ti] = fv.  MODE 1: s_tab4[tj*T + ti] = (c2, fv*rep, -fv*att/Reff, cut2)
    __shared__ __align__(16) float s_tab[CF_TT_MAX * (MODE ? 4 : 1)];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = c.T;
    for (int i = tid; i < T * T; i += blockDim.x) {
        const int ti = i / T, tj = i % T; // tables are [ti][tj]
        if (MODE == 0) {
            s_tab[tj * T + ti] = tables->force[i];
        } else {
            const float inv = tables->inv_reff[i], fv = tables->force[i];
            float* e = &s_tab[(tj * T + ti) * 4];
            e[0] = c.nk_log2e * inv * inv;
            e[1] = fv * c.repulsion;
            e[2] = -(fv * (c.attraction * inv));
            e[3] = tables->cut2[i];
        }
    }
    unsigned bar_addr = 0, bar_phase = 0; // STAGE 1: this warp's two mbarriers (+ 8 * buffer), their parity bits
    if (STAGE == 1) {
        bar_addr = (unsigned)__cvta_generic_to_shared(&sm.bar[warp][0]);
        if (lane == 0) {
            t4_mbar_init(bar_addr, 1);
            t4_mbar_init(bar_addr + 8, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
    }
    __syncthreads(); // the only block-level barrier: tables are read-only afterwards
    const int ntiles = ctrl[0];
    const int ny = c.dims[1], nz = c.dims[2];
    const float cutu = c.cut2_uniform;
    const float c2u = c.nk_log2e * c.inv_reff_uniform * c.inv_reff_uniform;
    const float pau = c.repulsion, pbu = -(c.attraction * c.inv_reff_uniform);
    int2* const wsub = &sm.sub[warp][0];
    const unsigned cst_addr = (unsigned)__cvta_generic_to_shared(&sm.cst[MODE ? warp : 0][0][lane]); // + 512 * layer
    const unsigned s_tab_addr = (unsigned)__cvta_generic_to_shared(s_tab);
    const unsigned stage_addr = (unsigned)__cvta_generic_to_shared(&sm.x[warp][0][0]); // + buf*256 + 4*slot
    constexpr unsigned STRIDE = T4_WARPS * 2 * T4_JC * 4;
    const unsigned lane_lt = (1u << lane) - 1u; // STAGE 4: rank of a live quad = popc(live mask & lane_lt)

    // Tail: the tiles of the last, partly filled round (ntiles mod warps-of-the-grid) are handed out
    // as 2 or 4 sub-tiles of 2 or 1 layers when that shortens the round: a sub-tile costs its share
    // plus ~8 % per halving (the j stream is staged for fewer layers).
    const int nwarps = (int)gridDim.x * T4_WARPS;
    const int tail = ntiles % nwarps, full = ntiles - tail;
    int split = 1;
    if (tail > 0) {
        const float c1 = 1.0f;
        const float c2 = IPT >= 2 ? 0.5f * 1.08f * (float)((2 * tail + nwarps - 1) / nwarps) : 1e9f;
        const float c4 = IPT >= 4 ? 0.25f * 1.22f * (float)((4 * tail + nwarps - 1) / nwarps) : 1e9f;
        split = (c2 < c1 && c2 <= c4) ? 2 : ((c4 < c1 && c4 < c2) ? 4 : 1);
    }
    const int nvirtual = full + tail * split;
    unsigned n_tested = 0, n_live = 0;

    for (;;) {
        int tile = 0;
        if (lane == 0) tile = atomicAdd(&ctrl[1], 1);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= nvirtual) break;
        int first_layer = 0, max_i = T4_TI;
        if (tile >= full) { // a sub-tile of a tail tile
            const int v = tile - full;
            tile = full + v / split;
            max_i = T4_TI / split;
            first_layer = (v % split) * (IPT / split);
        }
        const int2 tl = tiles[tile];
        const int cell = tl.x;
        const int cz = cell % nz, cy = (cell / nz) % ny, cx = cell / (nz * ny);
        const int i_begin = cell_start[cell] + tl.y * T4_TI + 32 * first_layer;
        const int ni = min(cell_start[cell + 1] - i_begin, max_i);
        if (ni <= 0) continue; // this part of the tile holds no particle

        // ---- neighbour runs: lane r < 18 holds run r = (row r/2 of the 3x3 (x,y) rows, z segment r%2)
        int r_code = 0; // minimum-image shift of the run, 2 bits per axis: 1 = -W, 2 = +W
        int r_row = -1, r_z0 = 0, r_z1 = 0;
        if (lane < TK_MAX_RUNS) {
            int rho = lane >> 1, seg = lane & 1;
            int x = cx + rho / 3 - 1, y = cy + rho % 3 - 1;
            bool valid = true;
            if (c.periodic_x) {
                if (x < 0) { x = c.dims[0] - 1; r_code |= 1; } else if (x >= c.dims[0]) { x = 0; r_code |= 2; }
            } else { // slab mode: i-cells are layers 1..dims-2, so x stays inside [0, dims-1]
                if (x < 0 || x >= c.dims[0]) valid = false;
                else if (x == 0) r_code |= (c.gshift_lo < 0.f ? 1 : (c.gshift_lo > 0.f ? 2 : 0));
                else if (x == c.dims[0] - 1) r_code |= (c.gshift_hi < 0.f ? 1 : (c.gshift_hi > 0.f ? 2 : 0));
            }
            if (y < 0) { y = ny - 1; r_code |= 4; } else if (y >= ny) { y = 0; r_code |= 8; }
            if (seg == 0) {
                r_z0 = max(cz - 1, 0);
                r_z1 = min(cz + 1, nz - 1);
            } else if (cz == 0) {
                r_z0 = r_z1 = nz - 1;
                r_code |= 16;
            } else if (cz == nz - 1) {
                r_z0 = r_z1 = 0;
                r_code |= 32;
            } else {
                valid = false;
            }
            if (valid) r_row = x * ny + y;
        }
        // (sub-)run list in shared memory: MODE 0 entry r = run r; MODE 1 entry r*T + t = type t of run r
        __syncwarp();
        const int nsub = MODE ? TK_MAX_RUNS * T : TK_MAX_RUNS;
        for (int e0 = 0; e0 < nsub; e0 += 32) {
            const int e = e0 + lane;
            const int r = MODE ? e / T : e, t = MODE ? e % T : 0;
            const int row = __shfl_sync(0xffffffffu, r_row, r & 31);
            const int z0 = __shfl_sync(0xffffffffu, r_z0, r & 31), z1 = __shfl_sync(0xffffffffu, r_z1, r & 31);
            if (e < nsub) {
                int2 jj = make_int2(0, 0);
                if (row >= 0) {
                    const int b = MODE ? (row * T + t) * nz : row * nz;
                    jj.x = startj[b + z0];
                    jj.y = startj[b + z1 + 1];
                }
                wsub[e] = jj;
            }
        }

        // ---- my i particles: layer k holds slots i_begin + 32k + lane ----
        float npx[IPT], npy[IPT], npz[IPT];
        unsigned tis4 = 0; // byte k = 4 * type of this lane's particle in layer k
        T4AccScalar acc[IPT];
        int cnt[IPT];
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            const int il = k * 32 + lane;
            const bool v = il < ni;
            const float4 p = v ? pos4[i_begin + il] : make_float4(-TK_FAR, -TK_FAR, -TK_FAR, 0.f);
            const int ti = v ? (int)__float_as_uint(p.w) : 0;
            npx[k] = -p.x, npy[k] = -p.y, npz[k] = -p.z;
            tis4 |= (unsigned)(ti * 4) << (8 * k);
            acc[k].zero();
            cnt[k] = 0;
            const float lx = t4_warp_min(v ? p.x : T4_INF), hx = t4_warp_max(v ? p.x : -T4_INF);
            const float ly = t4_warp_min(v ? p.y : T4_INF), hy = t4_warp_max(v ? p.y : -T4_INF);
            const float lz = t4_warp_min(v ? p.z : T4_INF), hz = t4_warp_max(v ? p.z : -T4_INF);
            if (lane == 0) { // bounding box of the layer + prefilter threshold (squared box gap)
                sm.box[warp][k][0] = make_float4(lx, ly, lz, cutu * 1.0001f + 0.01f);
                sm.box[warp][k][1] = make_float4(hx, hy, hz, 0.f);
            }
        }
        __syncwarp();

        // ---- stream the chunks of all (sub-)runs; the next chunk is prefetched across run boundaries
        //      into L1 by a prefetch hint, which holds no registers ----
        int si = -1, off = 0, end = 0, lo = 0; // cursor: chunk [off, off + 128) of sub-run si = [lo, end)
        bool have;
#define T4_ADVANCE()                                \
    do {                                            \
        off += T4_JC;                               \
        have = true;                                \
        while (off >= end) {                        \
            if (++si >= nsub) { have = false; break; } \
            const int2 e_ = wsub[si];               \
            lo = e_.x, end = e_.y;                  \
            off = (STAGE >= 1 && STAGE <= 3) ? (lo & ~3) : lo; /* bulk copies start on 16-byte boundaries */ \
        }                                           \
    } while (0)
        T4_ADVANCE();
        int cur_si = -1, buf = 0;
        if (STAGE == 1 && have && lane == 0) { // first chunk of the tile: three plane copies onto buffer 0's barrier
            t4_mbar_expect(bar_addr, 3 * T4_JC * 4);
            t4_bulk_load(stage_addr, jx + off, T4_JC * 4, bar_addr);
            t4_bulk_load(stage_addr + STRIDE, jy + off, T4_JC * 4, bar_addr);
            t4_bulk_load(stage_addr + 2 * STRIDE, jz + off, T4_JC * 4, bar_addr);
        }
        float sx = 0.f, sy = 0.f, sz = 0.f;
        bool wrap = false;
        while (have) {
            if (STAGE == 4) {
                const int csi = si, coff = off, cend = end;
                if (csi != cur_si) { // a new (sub-)run: its minimum-image shift, MODE 1: its pair constants
                    cur_si = csi;
                    const int r = MODE ? csi / T : csi;
                    const int code = __shfl_sync(0xffffffffu, r_code, r);
                    sx = (code & 1) ? -c.W[0] : ((code & 2) ? c.W[0] : 0.f);
                    sy = (code & 4) ? -c.W[1] : ((code & 8) ? c.W[1] : 0.f);
                    sz = (code & 16) ? -c.W[2] : ((code & 32) ? c.W[2] : 0.f);
                    wrap = code != 0;
                    if (MODE == 1) {
                        const int tj = csi - r * T;
                        const char* row = reinterpret_cast<const char*>(s_tab) + tj * (T * 16);
#pragma unroll
                        for (int k = 0; k < IPT; k++) {
                            float4 e = *reinterpret_cast<const float4*>(row + ((tis4 >> (8 * k)) & 255u) * 4u);
                            if (!(k * 32 + lane < ni)) e.w = 0.f; // no particle: never accepts
                            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(cst_addr + 512u * k), "f"(e.x), "f"(e.y), "f"(e.z), "f"(e.w) : "memory");
                            const float mx = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(e.w))); // cut2 >= 0
                            if (lane == 0) sm.box[warp][k][0].w = mx * 1.0001f + 0.01f;
                        }
                        __syncwarp();
                    }
                }
                // ---- the chunk into registers: lane l holds the quad j = coff + 4l .. 4l+3 (clamped to the sub-run) ----
                const int j0 = coff + 4 * lane, jl = cend - 1;
                float4 q[4];
#pragma unroll
                for (int u = 0; u < 4; u++) q[u] = posj[min(j0 + u, jl)];
                T4_ADVANCE();
                if (have && lane < 17) { // the next chunk (possibly of the next run): prefetch hint, 2 KiB <= 17 lines
                    const float4* pf = posj + min(off + 8 * lane, end - 1);
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));
                }
                const float lox = fminf(fminf(q[0].x, q[1].x), fminf(q[2].x, q[3].x)), hix = fmaxf(fmaxf(q[0].x, q[1].x), fmaxf(q[2].x, q[3].x));
                const float loy = fminf(fminf(q[0].y, q[1].y), fminf(q[2].y, q[3].y)), hiy = fmaxf(fmaxf(q[0].y, q[1].y), fmaxf(q[2].y, q[3].y));
                const float loz = fminf(fminf(q[0].z, q[1].z), fminf(q[2].z, q[3].z)), hiz = fmaxf(fmaxf(q[0].z, q[1].z), fmaxf(q[2].z, q[3].z));
                // ---- box prefilter on the registers; quads past the end of the sub-run are masked out ----
                const int nvq = (cend - coff + 3) >> 2;
                const unsigned vmask = nvq >= 32 ? 0xffffffffu : ((1u << nvq) - 1u);
                unsigned live[IPT], any_live = 0;
#pragma unroll
                for (int k = 0; k < IPT; k++) {
                    const float4 b0 = sm.box[warp][k][0], b1 = sm.box[warp][k][1]; // broadcast reads
                    const float gx = fmaxf(fmaxf((lox - b1.x) + sx, (b0.x - hix) - sx), 0.f);
                    const float gy = fmaxf(fmaxf((loy - b1.y) + sy, (b0.y - hiy) - sy), 0.f);
                    const float gz = fmaxf(fmaxf((loz - b1.z) + sz, (b0.z - hiz) - sz), 0.f);
                    live[k] = __ballot_sync(0xffffffffu, fmaf(gz, gz, fmaf(gx, gx, gy * gy)) < b0.w) & vmask;
                    any_live |= live[k];
                }
                if (any_live) {
                    const unsigned sbase = stage_addr + (unsigned)buf * (T4_JC * 4);
                    // the repeated elements of the sub-run's last quad leave the range (selects that take the place of the
                    // register moves of the transposition)
                    const bool v1 = j0 + 1 < cend, v2 = j0 + 2 < cend, v3 = j0 + 3 < cend;
                    const unsigned rank = (unsigned)__popc(any_live & lane_lt);
                    if ((any_live >> lane) & 1u) {
                        const unsigned a = sbase + 16u * rank;
                        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(q[0].x), "f"(v1 ? q[1].x : TK_FAR), "f"(v2 ? q[2].x : TK_FAR), "f"(v3 ? q[3].x : TK_FAR) : "memory");
                        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a + STRIDE), "f"(q[0].y), "f"(v1 ? q[1].y : TK_FAR), "f"(v2 ? q[2].y : TK_FAR), "f"(v3 ? q[3].y : TK_FAR) : "memory");
                        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a + 2 * STRIDE), "f"(q[0].z), "f"(v1 ? q[1].z : TK_FAR), "f"(v2 ? q[2].z : TK_FAR), "f"(v3 ? q[3].z : TK_FAR) : "memory");
                        if (MODE == 0) {
                            const int rb = T * 4; // bytes per tj row of s_tab
                            asm volatile("st.shared.v4.s32 [%0], {%1,%2,%3,%4};" ::"r"(a + 3 * STRIDE),
                                         "r"((int)__float_as_uint(q[0].w) * rb), "r"((int)__float_as_uint(q[1].w) * rb),
                                         "r"((int)__float_as_uint(q[2].w) * rb), "r"((int)__float_as_uint(q[3].w) * rb) : "memory");
                        }
                    }
                    unsigned clive[IPT];
#pragma unroll
                    for (int k = 0; k < IPT; k++) // MODE 0, > 1 layer: the per-layer masks in list order
                        clive[k] = (MODE == 0 && IPT > 1) ? __reduce_or_sync(0xffffffffu, ((live[k] >> lane) & 1u) << rank) : 0u;
                    __syncwarp();
                    const unsigned nq = (unsigned)__popc(any_live);
                    if (wrap)
                        t4_chunk_compact<MODE, true, COUNT, IPT>(sbase, nq, clive, npx, npy, npz, tis4, s_tab_addr, cst_addr, sx, sy, sz,
                                                                 cutu, c2u, pau, pbu, acc, cnt, n_tested, n_live);
                    else
                        t4_chunk_compact<MODE, false, COUNT, IPT>(sbase, nq, clive, npx, npy, npz, tis4, s_tab_addr, cst_addr, sx, sy, sz,
                                                                  cutu, c2u, pau, pbu, acc, cnt, n_tested, n_live);
                    buf ^= 1; // the other buffer was last read one chunk ago by this same warp
                }
                continue;
            }
            const int csi = si, coff = off, cend = end, clo = lo;
            // ---- load and publish the chunk: lane l holds the quad j = coff + 4l .. 4l+3 ----
            const unsigned sbase = stage_addr + (unsigned)buf * (T4_JC * 4);
            float lox = T4_INF, loy = T4_INF, loz = T4_INF, hix = -T4_INF, hiy = -T4_INF, hiz = -T4_INF;
            if (STAGE == 3) {
                const int j0 = coff + 4 * lane; // aligned to 4 elements = 16 bytes in every plane
                float4 X = make_float4(TK_FAR, TK_FAR, TK_FAR, TK_FAR), Y = X, Z = X;
                if (j0 < cend) { // (the planes are padded: a quad may reach past the sub-run, never past the array)
                    X = *reinterpret_cast<const float4*>(jx + j0);
                    Y = *reinterpret_cast<const float4*>(jy + j0);
                    Z = *reinterpret_cast<const float4*>(jz + j0);
                }
                if (coff >= clo && coff + T4_JC <= cend) { // the whole chunk lies inside the sub-run: no masks
                    lox = fminf(fminf(X.x, X.y), fminf(X.z, X.w)), hix = fmaxf(fmaxf(X.x, X.y), fmaxf(X.z, X.w));
                    loy = fminf(fminf(Y.x, Y.y), fminf(Y.z, Y.w)), hiy = fmaxf(fmaxf(Y.x, Y.y), fmaxf(Y.z, Y.w));
                    loz = fminf(fminf(Z.x, Z.y), fminf(Z.z, Z.w)), hiz = fmaxf(fmaxf(Z.x, Z.y), fmaxf(Z.z, Z.w));
                } else {
                    const bool v0 = j0 >= clo && j0 < cend, v1 = j0 + 1 >= clo && j0 + 1 < cend;
                    const bool v2 = j0 + 2 >= clo && j0 + 2 < cend, v3 = j0 + 3 >= clo && j0 + 3 < cend;
                    if (!v0) X.x = Y.x = Z.x = TK_FAR; // elements of another sub-run (or past the end): never in range
                    if (!v1) X.y = Y.y = Z.y = TK_FAR;
                    if (!v2) X.z = Y.z = Z.z = TK_FAR;
                    if (!v3) X.w = Y.w = Z.w = TK_FAR;
                    lox = fminf(fminf(v0 ? X.x : T4_INF, v1 ? X.y : T4_INF), fminf(v2 ? X.z : T4_INF, v3 ? X.w : T4_INF));
                    hix = fmaxf(fmaxf(v0 ? X.x : -T4_INF, v1 ? X.y : -T4_INF), fmaxf(v2 ? X.z : -T4_INF, v3 ? X.w : -T4_INF));
                    loy = fminf(fminf(v0 ? Y.x : T4_INF, v1 ? Y.y : T4_INF), fminf(v2 ? Y.z : T4_INF, v3 ? Y.w : T4_INF));
                    hiy = fmaxf(fmaxf(v0 ? Y.x : -T4_INF, v1 ? Y.y : -T4_INF), fmaxf(v2 ? Y.z : -T4_INF, v3 ? Y.w : -T4_INF));
                    loz = fminf(fminf(v0 ? Z.x : T4_INF, v1 ? Z.y : T4_INF), fminf(v2 ? Z.z : T4_INF, v3 ? Z.w : T4_INF));
                    hiz = fmaxf(fmaxf(v0 ? Z.x : -T4_INF, v1 ? Z.y : -T4_INF), fmaxf(v2 ? Z.z : -T4_INF, v3 ? Z.w : -T4_INF));
                }
                const unsigned a = sbase + 16u * (unsigned)lane;
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(X.x), "f"(X.y), "f"(X.z), "f"(X.w) : "memory");
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a + STRIDE), "f"(Y.x), "f"(Y.y), "f"(Y.z), "f"(Y.w) : "memory");
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a + 2 * STRIDE), "f"(Z.x), "f"(Z.y), "f"(Z.z), "f"(Z.w) : "memory");
            } else if (STAGE == 2) {
                // precomputed box of my aligned quad (elements of a neighbouring sub-run only make it larger)
                const int j0 = coff + 4 * lane;
                if (j0 + 3 >= clo && j0 < cend) {
                    const float4 bl = qbox[2 * (j0 >> 2)], bh = qbox[2 * (j0 >> 2) + 1];
                    lox = bl.x, loy = bl.y, loz = bl.z, hix = bh.x, hiy = bh.y, hiz = bh.z;
                }
            } else if (STAGE == 1) {
                // the chunk was requested one iteration ago: wait for its bytes, then every lane reads its own quad
                while (!t4_mbar_try(bar_addr + 8u * (unsigned)buf, (bar_phase >> buf) & 1u)) {}
                bar_phase ^= 1u << buf;
                const unsigned a = sbase + 16u * (unsigned)lane;
                float4 X, Y, Z;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(X.x), "=f"(X.y), "=f"(X.z), "=f"(X.w) : "r"(a));
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(Y.x), "=f"(Y.y), "=f"(Y.z), "=f"(Y.w) : "r"(a + STRIDE));
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(Z.x), "=f"(Z.y), "=f"(Z.z), "=f"(Z.w) : "r"(a + 2 * STRIDE));
                const int j0 = coff + 4 * lane;
                const bool v0 = j0 >= clo && j0 < cend, v1 = j0 + 1 >= clo && j0 + 1 < cend;
                const bool v2 = j0 + 2 >= clo && j0 + 2 < cend, v3 = j0 + 3 >= clo && j0 + 3 < cend;
                if (!(v0 && v1 && v2 && v3)) { // elements of another sub-run (or past the end): never in range
                    if (!v0) X.x = Y.x = Z.x = TK_FAR;
                    if (!v1) X.y = Y.y = Z.y = TK_FAR;
                    if (!v2) X.z = Y.z = Z.z = TK_FAR;
                    if (!v3) X.w = Y.w = Z.w = TK_FAR;
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(X.x), "f"(X.y), "f"(X.z), "f"(X.w) : "memory");
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a + STRIDE), "f"(Y.x), "f"(Y.y), "f"(Y.z), "f"(Y.w) : "memory");
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a + 2 * STRIDE), "f"(Z.x), "f"(Z.y), "f"(Z.z), "f"(Z.w) : "memory");
                }
                lox = fminf(fminf(v0 ? X.x : T4_INF, v1 ? X.y : T4_INF), fminf(v2 ? X.z : T4_INF, v3 ? X.w : T4_INF));
                hix = fmaxf(fmaxf(v0 ? X.x : -T4_INF, v1 ? X.y : -T4_INF), fmaxf(v2 ? X.z : -T4_INF, v3 ? X.w : -T4_INF));
                loy = fminf(fminf(v0 ? Y.x : T4_INF, v1 ? Y.y : T4_INF), fminf(v2 ? Y.z : T4_INF, v3 ? Y.w : T4_INF));
                hiy = fmaxf(fmaxf(v0 ? Y.x : -T4_INF, v1 ? Y.y : -T4_INF), fmaxf(v2 ? Y.z : -T4_INF, v3 ? Y.w : -T4_INF));
                loz = fminf(fminf(v0 ? Z.x : T4_INF, v1 ? Z.y : T4_INF), fminf(v2 ? Z.z : T4_INF, v3 ? Z.w : T4_INF));
                hiz = fmaxf(fmaxf(v0 ? Z.x : -T4_INF, v1 ? Z.y : -T4_INF), fmaxf(v2 ? Z.z : -T4_INF, v3 ? Z.w : -T4_INF));
            } else {
                float4 q[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int j = coff + 4 * lane + u;
                    q[u] = j < cend ? posj[j] : make_float4(TK_FAR, TK_FAR, TK_FAR, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 4; u++) { // bounding box of the quad (valid j only)
                    const bool v = coff + 4 * lane + u < cend;
                    lox = fminf(lox, v ? q[u].x : T4_INF), hix = fmaxf(hix, v ? q[u].x : -T4_INF);
                    loy = fminf(loy, v ? q[u].y : T4_INF), hiy = fmaxf(hiy, v ? q[u].y : -T4_INF);
                    loz = fminf(loz, v ? q[u].z : T4_INF), hiz = fmaxf(hiz, v ? q[u].z : -T4_INF);
                }
                const unsigned a = sbase + 16u * (unsigned)lane;
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(q[0].x), "f"(q[1].x), "f"(q[2].x), "f"(q[3].x));
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a + STRIDE), "f"(q[0].y), "f"(q[1].y), "f"(q[2].y), "f"(q[3].y));
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a + 2 * STRIDE), "f"(q[0].z), "f"(q[1].z), "f"(q[2].z), "f"(q[3].z));
                if (MODE == 0) {
                    const int rb = T * 4; // bytes per tj row of s_tab
                    asm volatile("st.shared.v4.s32 [%0], {%1,%2,%3,%4};" ::"r"(a + 3 * STRIDE),
                                 "r"((int)__float_as_uint(q[0].w) * rb), "r"((int)__float_as_uint(q[1].w) * rb),
                                 "r"((int)__float_as_uint(q[2].w) * rb), "r"((int)__float_as_uint(q[3].w) * rb));
                }
            }
            __syncwarp();
            T4_ADVANCE();
            if (STAGE == 3) {
                if (have && lane < 12) { // the next chunk: 512 bytes = 4 lines in each of the three planes
                    const float* pl = lane < 4 ? jx : (lane < 8 ? jy : jz);
                    const float* pf = pl + min(off + 32 * (lane & 3), end - 1);
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));
                }
            } else if (STAGE == 2) {
                if (have && lane < 8) { // the next chunk's 32 quad boxes: 1 KiB = 8 lines
                    const float4* pf = qbox + 2 * (off >> 2) + min(8 * lane, 2 * ((end - 1 - off) >> 2)); // inside the array
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));
                }
            } else if (STAGE == 1) {
                // the next chunk (possibly of the next sub-run) -> the other buffer, which this warp finished
                // reading one chunk ago (the __syncwarp above orders those reads before the request)
                if (have && lane == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic accesses of that buffer first
                    const unsigned nb = bar_addr + 8u * (unsigned)(buf ^ 1);
                    const unsigned dst = stage_addr + (unsigned)(buf ^ 1) * (T4_JC * 4);
                    t4_mbar_expect(nb, 3 * T4_JC * 4);
                    t4_bulk_load(dst, jx + off, T4_JC * 4, nb);
                    t4_bulk_load(dst + STRIDE, jy + off, T4_JC * 4, nb);
                    t4_bulk_load(dst + 2 * STRIDE, jz + off, T4_JC * 4, nb);
                }
            } else if (have && lane < 17) {
                // the next chunk (possibly of the next run): prefetch hint, 128 float4 = 2 KiB <= 17 lines
                const float4* pf = posj + min(off + 8 * lane, end - 1);
                asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));
            }

            if (csi != cur_si) { // a new (sub-)run: its minimum-image shift, MODE 1: its pair constants
                cur_si = csi;
                const int r = MODE ? csi / T : csi;
                const int code = __shfl_sync(0xffffffffu, r_code, r);
                sx = (code & 1) ? -c.W[0] : ((code & 2) ? c.W[0] : 0.f);
                sy = (code & 4) ? -c.W[1] : ((code & 8) ? c.W[1] : 0.f);
                sz = (code & 16) ? -c.W[2] : ((code & 32) ? c.W[2] : 0.f);
                wrap = code != 0;
                if (MODE == 1) {
                    const int tj = csi - r * T;
                    const char* row = reinterpret_cast<const char*>(s_tab) + tj * (T * 16);
#pragma unroll
                    for (int k = 0; k < IPT; k++) {
                        float4 e = *reinterpret_cast<const float4*>(row + ((tis4 >> (8 * k)) & 255u) * 4u);
                        if (!(k * 32 + lane < ni)) e.w = 0.f; // no particle: never accepts
                        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(cst_addr + 512u * k), "f"(e.x), "f"(e.y), "f"(e.z), "f"(e.w) : "memory");
                        const float mx = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(e.w))); // cut2 >= 0
                        if (lane == 0) sm.box[warp][k][0].w = mx * 1.0001f + 0.01f;
                    }
                    __syncwarp();
                }
            }
            // ---- box prefilter: every lane tests its own quad against the box of each layer ----
            unsigned live[IPT];
#pragma unroll
            for (int k = 0; k < IPT; k++) {
                const float4 b0 = sm.box[warp][k][0], b1 = sm.box[warp][k][1]; // broadcast reads
                const float gx = fmaxf(fmaxf((lox - b1.x) + sx, (b0.x - hix) - sx), 0.f);
                const float gy = fmaxf(fmaxf((loy - b1.y) + sy, (b0.y - hiy) - sy), 0.f);
                const float gz = fmaxf(fmaxf((loz - b1.z) + sz, (b0.z - hiz) - sz), 0.f);
                live[k] = __ballot_sync(0xffffffffu, fmaf(gz, gz, fmaf(gx, gx, gy * gy)) < b0.w);
            }
            unsigned any_live = 0;
#pragma unroll
            for (int k = 0; k < IPT; k++) any_live |= live[k];
            if (STAGE == 2 && any_live) {
                // stage the surviving quads only (t4_chunk never reads the others)
                if ((any_live >> lane) & 1u) {
                    float4 q[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int j = coff + 4 * lane + u;
                        q[u] = (j >= clo && j < cend) ? posj[j] : make_float4(TK_FAR, TK_FAR, TK_FAR, 0.f);
                    }
                    const unsigned a = sbase + 16u * (unsigned)lane;
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(q[0].x), "f"(q[1].x), "f"(q[2].x), "f"(q[3].x) : "memory");
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a + STRIDE), "f"(q[0].y), "f"(q[1].y), "f"(q[2].y), "f"(q[3].y) : "memory");
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a + 2 * STRIDE), "f"(q[0].z), "f"(q[1].z), "f"(q[2].z), "f"(q[3].z) : "memory");
                    if (MODE == 0) {
                        const int rb = T * 4; // bytes per tj row of s_tab
                        asm volatile("st.shared.v4.s32 [%0], {%1,%2,%3,%4};" ::"r"(a + 3 * STRIDE),
                                     "r"((int)__float_as_uint(q[0].w) * rb), "r"((int)__float_as_uint(q[1].w) * rb),
                                     "r"((int)__float_as_uint(q[2].w) * rb), "r"((int)__float_as_uint(q[3].w) * rb) : "memory");
                    }
                }
                __syncwarp();
            }
            if (any_live) {
                if (wrap)
                    t4_chunk<MODE, true, COUNT, IPT>(sbase, live, npx, npy, npz, tis4, s_tab_addr, cst_addr, sx, sy, sz, cutu, c2u,
                                                pau, pbu, acc, cnt, n_tested, n_live);
                else
                    t4_chunk<MODE, false, COUNT, IPT>(sbase, live, npx, npy, npz, tis4, s_tab_addr, cst_addr, sx, sy, sz, cutu, c2u,
                                                 pau, pbu, acc, cnt, n_tested, n_live);
            }
            buf ^= 1; // the other buffer was last read one chunk ago by this same warp
        }
#undef T4_ADVANCE

        // ---- write back: the particle itself was tested too (d = 0, force term exactly 0) ----
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            const int il = k * 32 + lane;
            if (il < ni) {
                const float3 f = acc[k].sum();
                const int ti = (int)((tis4 >> (8 * k)) & 255u) >> 2;
                const int self_ok = (MODE ? s_tab[(ti * T + ti) * 4 + 3] : cutu) > 0.f ? 1 : 0;
                frc4[i_begin + il] = make_float4(f.x, f.y, f.z, __int_as_float(cnt[k] - self_ok));
            }
        }
    }
    if (COUNT && lane == 0 && block_counts) {
        atomicAdd(&block_counts[0], (unsigned long long)n_tested);
        atomicAdd(&block_counts[1], (unsigned long long)n_live);
    }
}

// ---------------------------------------------------------------------------------------------
// The type-homogeneous j copy (MODE 1): positions sorted by (xy row, type, z cell, Morton).
// Built from the cell-sorted array by ONE stable sort on key = row * T + type (the cell-sorted
// order already is (row, z cell, Morton)), a gather, and a lower-bound pass.
// ---------------------------------------------------------------------------------------------
__global__ void homog_key_kernel(const float4* __restrict__ pos4, const int* __restrict__ cell_start, int ncell,
                                 int nz, int T, int nslots_upper, const int* __restrict__ d_nslots,
                                 const int* __restrict__ d_first, uint32_t* __restrict__ keys,
                                 uint32_t* __restrict__ vals, int* __restrict__ cell_of) {
    // element e of the copy <-> slot first + e (slab mode: the used slots start at the left ghost layer)
    const int nslots = d_nslots ? min(*d_nslots, nslots_upper) : nslots_upper;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nslots) return;
    const int k = e + (d_first ? *d_first : 0);
    uint32_t key = (uint32_t)((ncell / nz) * T); // sentinel: slot holds no particle
    int cell = -1;
    if (k >= cell_start[0] && k < cell_start[ncell]) {
        int lo = 0, hi = ncell; // last cell with cell_start[cell] <= k
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (cell_start[mid] <= k) lo = mid; else hi = mid;
        }
        cell = lo;
        int t = (int)__float_as_uint(pos4[k].w);
        t = t < 0 ? 0 : (t > T - 1 ? T - 1 : t);
        key = (uint32_t)((cell / nz) * T + t);
    }
    keys[e] = key;
    vals[e] = (uint32_t)k;
    cell_of[k] = cell;
}

__global__ void homog_gather_kernel(const uint32_t* __restrict__ skeys, const uint32_t* __restrict__ svals,
                                    const float4* __restrict__ pos4, const int* __restrict__ cell_of, int nz,
                                    int nslots_upper, const int* __restrict__ d_nslots, float4* __restrict__ posj,
                                    uint32_t* __restrict__ comp, float* __restrict__ jx, float* __restrict__ jy,
                                    float* __restrict__ jz) {
    const int nslots = d_nslots ? min(*d_nslots, nslots_upper) : nslots_upper;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nslots) return;
    const uint32_t src = svals[k];
    const int cell = cell_of[src];
    const float4 p = pos4[src];
    posj[k] = p;
    if (jx) jx[k] = p.x, jy[k] = p.y, jz[k] = p.z; // SoA planes for the bulk-copy staging (STAGE 1)
    comp[k] = cell >= 0 ? skeys[k] * (uint32_t)nz + (uint32_t)(cell % nz) : 0xffffffffu;
}

// startj[e] = first index of the sorted copy whose composite key (row*T + type)*nz + cz is >= e
__global__ void homog_bounds_kernel(const uint32_t* __restrict__ comp, int nslots_upper, const int* __restrict__ d_nslots,
                                    int* __restrict__ startj, int nkeys) {
    const int nslots = d_nslots ? min(*d_nslots, nslots_upper) : nslots_upper;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e > nkeys) return;
    int lo = 0, hi = nslots;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (comp[mid] < (uint32_t)e) lo = mid + 1; else hi = mid;
    }
    startj[e] = lo;
}

// Bounding boxes of the aligned quads of a j array (STAGE 2): qbox[2q] = min xyz, qbox[2q+1] = max xyz of elements
// 4q .. 4q+3 inside [first, first + count) (empty box when none is).  One thread per quad.
__global__ void quad_box_kernel(const float4* __restrict__ posj, int first_host, const int* __restrict__ d_first,
                                int count_upper, const int* __restrict__ d_count, float4* __restrict__ qbox) {
    const int first = d_first ? *d_first : first_host;
    const int count = d_count ? min(*d_count, count_upper) : count_upper;
    const int q = (first >> 2) + blockIdx.x * blockDim.x + threadIdx.x;
    if (4 * q >= first + count) return;
    float lx = T4_INF, ly = T4_INF, lz = T4_INF, hx = -T4_INF, hy = -T4_INF, hz = -T4_INF;
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int j = 4 * q + u;
        if (j >= first && j < first + count) {
            const float4 p = posj[j];
            lx = fminf(lx, p.x), hx = fmaxf(hx, p.x);
            ly = fminf(ly, p.y), hy = fmaxf(hy, p.y);
            lz = fminf(lz, p.z), hz = fmaxf(hz, p.z);
        }
    }
    qbox[2 * q] = make_float4(lx, ly, lz, 0.f);
    qbox[2 * q + 1] = make_float4(hx, hy, hz, 0.f);
}
