// cf_device.cuh — constants and small device helpers shared by every kernel of the engine.
//
// Arithmetic contract: everything that decides a neighbour — displacement,
// minimum-image wrap, squared distance — uses exactly the rounding points of the reference as
// nvcc compiles it for sm_100a (ParticleSimulation.cu:92-112): d = o - p, exact wrap by +-W,
// d2 = fma(dz,dz, fma(dx,dx, dy*dy)).  Those are written with the __f*_rn intrinsics, which the
// compiler never contracts or reorders.  The comparison "sqrtf(d2 + 1e-4f) < Reff" is replaced
// by the equivalent "d2 < cut2[type pair]" with cut2 computed exactly on the host (tables.cpp).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define CF_T_MAX 10
#define CF_TT_MAX (CF_T_MAX * CF_T_MAX)

// Per-step constants, passed to kernels by value.
struct StepConst {
    float W[3];      // canvasWidth, canvasHeight, canvasDepth
    float halfW[3];  // W * 0.5f   (reference: ParticleSimulation.cu:97-102)
    float nhalfW[3]; // W * -0.5f
    float inv[3];    // (float)dims / W      — cell index = min((int)(pos * inv), dims - 1)
    int dims[3];     // cells per axis; linear cell = (cx * dims[1] + cy) * dims[2] + cz
    int periodic_x;  // 0 when x is slab-decomposed (ghost layers replace the wrap)
    // x cell coordinate = clamp((int)((x - x_org) * inv[0]), 0, x_cells - 1) + x_off.  Single GPU:
    // x_org = 0, x_off = 0, x_cells = dims[0].  Slab mode: x_org = slab lower bound, x_off = 1
    // (layer 0 and layer dims[0]-1 are ghost layers), x_cells = owned layers.
    float x_org;
    int x_off, x_cells;
    float gshift_lo, gshift_hi; // minimum-image shift of the two ghost layers (+-W at the seam)
    int gx_lo, gx_hi;           // x layers the (non-periodic) proximity graph may look at
    int T;
    float repulsion, attraction;
    float nk_log2e;  // -k * log2(e): exp(-k r^2) = exp2(nk_log2e * r^2)
    float dt, friction;
    float one_minus_balance, force_multiplier, max_expected; // .cu:136-143
    int uniform_radius; // 1 when every type pair has the same cut2 (all shipped presets)
    float cut2_uniform, inv_reff_uniform;
};

// Device tables (global memory, staged into shared memory by the kernels that index them):
//   [0      , TT)   cut2    : accept iff d2 < cut2[ti*T+tj]
//   [TT     , 2TT)  invReff : 1 / Reff[ti*T+tj]
//   [2TT    , 3TT)  force   : forceTable[ti*T+tj]  (row = self, column = other; .cu:116)
struct DeviceTables {
    float cut2[CF_TT_MAX];
    float inv_reff[CF_TT_MAX];
    float force[CF_TT_MAX];
};

__device__ __forceinline__ int cf_cell_coord(float x, float inv, int n) {
    int c = (int)__fmul_rn(x, inv);
    c = c > n - 1 ? n - 1 : c;
    return c < 0 ? 0 : c;
}

__device__ __forceinline__ int cf_cell_coord_x(float x, const StepConst& c) {
    int v = (int)__fmul_rn(__fsub_rn(x, c.x_org), c.inv[0]);
    v = v > c.x_cells - 1 ? c.x_cells - 1 : v;
    return (v < 0 ? 0 : v) + c.x_off;
}

__device__ __forceinline__ uint32_t cf_cell_key(float4 p, const StepConst& c) {
    int cx = cf_cell_coord_x(p.x, c);
    int cy = cf_cell_coord(p.y, c.inv[1], c.dims[1]);
    int cz = cf_cell_coord(p.z, c.inv[2], c.dims[2]);
    return (uint32_t)((cx * c.dims[1] + cy) * c.dims[2] + cz);
}

// Sort key = cell * 64 + Hilbert index of the particle's 4x4x4 sub-cell.  Inside a cell, particles
// that are adjacent in memory are adjacent in space, so the 32 i-particles of a warp layer and every
// quad of the j stream are compact blobs: whole (layer, quad) blocks are in range or out of range
// together.  A Hilbert curve has no jumps, so ANY run of consecutive particles is a connected blob;
// the Morton order of round 1 tears runs that straddle an octant boundary apart (model:
// tools/model/order_model.py, eater workload: 11% fewer exact block tests, 7% fewer live blocks).
// The cell part is exactly cf_cell_key (4*y truncates to 4*trunc(y) + sub for y >= 0).
#define CF_KEY_SUB 64u
__device__ __forceinline__ int cf_sub_coord(float y, int cell, int n) {
    int f = (int)__fmul_rn(y, 4.0f);
    f = f > 4 * n - 1 ? 4 * n - 1 : f;
    int sub = f - 4 * cell;
    return sub < 0 ? 0 : (sub > 3 ? 3 : sub);
}
// Hilbert index (Skilling's transform, 2 bits per axis) of sub-cell (sx, sy, sz), byte (sx*16 + sy*4 + sz)
// of this table; tests/test_parity_gpu.py::test_cell_assignment_bit_exact re-derives it.
__device__ __forceinline__ uint32_t cf_hilbert64(uint32_t sx, uint32_t sy, uint32_t sz) {
    const uint32_t i = (sx << 4) | (sy << 2) | sz;
    const uint32_t w = i >> 2;
    // 16 words selected with a chain of selects on the word index (no memory access, no local array)
    uint32_t v;
    switch (w) {
        case 0: v = 0x09080700u; break;
        case 1: v = 0x0e0f0601u; break;
        case 2: v = 0x1110191eu; break;
        case 3: v = 0x12131a1du; break;
        case 4: v = 0x0a0b0403u; break;
        case 5: v = 0x0d0c0502u; break;
        case 6: v = 0x1617181fu; break;
        case 7: v = 0x15141b1cu; break;
        case 8: v = 0x35343b3cu; break;
        case 9: v = 0x32333a3du; break;
        case 10: v = 0x29282720u; break;
        case 11: v = 0x2a2b2423u; break;
        case 12: v = 0x3637383fu; break;
        case 13: v = 0x3130393eu; break;
        case 14: v = 0x2e2f2621u; break;
        default: v = 0x2d2c2522u; break;
    }
    return (v >> (8u * (i & 3u))) & 255u;
}
__device__ __forceinline__ uint32_t cf_sort_key(float4 p, const StepConst& c) {
    float yx = __fmul_rn(__fsub_rn(p.x, c.x_org), c.inv[0]);
    float yy = __fmul_rn(p.y, c.inv[1]), yz = __fmul_rn(p.z, c.inv[2]);
    int cx = cf_cell_coord_x(p.x, c), cy = cf_cell_coord(p.y, c.inv[1], c.dims[1]),
        cz = cf_cell_coord(p.z, c.inv[2], c.dims[2]);
    uint32_t sx = (uint32_t)cf_sub_coord(yx, cx - c.x_off, c.x_cells);
    uint32_t sy = (uint32_t)cf_sub_coord(yy, cy, c.dims[1]);
    uint32_t sz = (uint32_t)cf_sub_coord(yz, cz, c.dims[2]);
    uint32_t cell = (uint32_t)((cx * c.dims[1] + cy) * c.dims[2] + cz);
    return cell * CF_KEY_SUB + cf_hilbert64(sx, sy, sz);
}

// Minimum-image wrap exactly as the reference: two dependent tests (.cu:97-98).  Both
// additions are exact in fp32 (Sterbenz), so any formulation of the same decision is bit-equal.
__device__ __forceinline__ float cf_wrap(float d, float W, float half, float nhalf) {
    d = d > half ? __fsub_rn(d, W) : d;
    d = d < nhalf ? __fadd_rn(d, W) : d;
    return d;
}

__device__ __forceinline__ float cf_dist2(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__device__ __forceinline__ float cf_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float cf_rsqrt(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Accepted-pair force term (ParticleSimulation.cu:113-131), fast-intrinsic form:
//   dist = sqrt(d2 + 1e-4), r = dist / Reff, net = (rep * exp(-k r^2) - att * r) * fv,
//   F += (d / dist) * net.
// MUFU.RSQ / MUFU.EX2 replace the IEEE sqrt, the four IEEE divides and expf of the reference;
// each is within 2 ulp, far inside the 1e-5 force tolerance (tests/test_parity_gpu.py).
__device__ __forceinline__ void cf_pair_force(float dx, float dy, float dz, float d2, float inv_reff,
                                              float fv, const StepConst& c, float& fx, float& fy,
                                              float& fz) {
    float x = __fadd_rn(d2, 0.0001f);
    float rinv = cf_rsqrt(x);
    float dist = x * rinv;
    float r = dist * inv_reff;
    float e = cf_ex2(r * r * c.nk_log2e);
    float net = fmaf(e, c.repulsion, -(r * c.attraction));
    float s = fv * net * rinv;
    fx = fmaf(s, dx, fx);
    fy = fmaf(s, dy, fy);
    fz = fmaf(s, dz, fz);
}

// Counter-based generator of cf_init_particles (restated in oracle/cellflow_oracle.c).
__host__ __device__ __forceinline__ uint64_t cf_mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ float cf_u01(uint64_t h) {
    return (float)(h >> 40) * 5.9604644775390625e-08f;
}
