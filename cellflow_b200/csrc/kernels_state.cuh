// kernels_state.cuh — particle state conversion, initialisation, key generation, reorder,
// fused integrate.  All HBM-bound, one thread per particle, float4 SoA.
//
// State layout in HBM (per particle slot, sorted by cell key of the last sort):
//   pos4 : x, y, z, bits(ptype)          16 B
//   vel4 : vx, vy, vz, bits(prevCount)   16 B   (prevCount = neighborCounts ping-pong, .cu:544-545)
//   frc4 : fx, fy, fz, bits(count)       16 B   (pair-force kernel output; after integrate it
//                                               holds the scaled force = reference p.acc, .cu:146)
//   id   : original particle index        4 B
#pragma once
#include "cf_device.cuh"

struct AosParticle { // reference Particle, SimulationParams.h:6-12 (44 B, 4-byte aligned)
    float pos[3], vel[3], acc[3];
    uint32_t ptype;
    float pad;
};

// Replaces initCurandKernel + initializeParticlesKernel (ParticleSimulation.cu:21-66): same spawn
// shape, counter-based generator keyed by (seed, id) — no RNG state array, reproducible.
__global__ void init_particles_kernel(float4* __restrict__ pos4, float4* __restrict__ vel4,
                                      float4* __restrict__ frc4, int* __restrict__ id, int n,
                                      int id0, int T, uint64_t seed, int mode, float W0, float W1,
                                      float W2) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint64_t pid = (uint64_t)(id0 + k);
    uint64_t s = cf_mix64(seed ^ cf_mix64(pid));
    float W[3] = {W0, W1, W2};
    float x[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float span = mode == 0 ? fminf(2000.0f, W[c]) : W[c];
        float off = __fmul_rn(__fsub_rn(W[c], span), 0.5f);
        float v = __fmaf_rn(cf_u01(cf_mix64(s + (uint64_t)c)), span, off);
        if (v >= W[c]) v = __uint_as_float(__float_as_uint(W[c]) - 1u);
        x[c] = v;
    }
    uint32_t t = (uint32_t)__fmul_rn(cf_u01(cf_mix64(s + 3u)), (float)T);
    t = t < (uint32_t)T ? t : (uint32_t)T - 1u;
    pos4[k] = make_float4(x[0], x[1], x[2], __uint_as_float(t));
    vel4[k] = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
    frc4[k] = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
    id[k] = id0 + k;
}

// Host AoS (reference layout) -> device SoA.  ids == nullptr: id = slot index.
__global__ void aos_to_soa_kernel(const AosParticle* __restrict__ aos, const int* __restrict__ counts,
                                  const int* __restrict__ ids, float4* __restrict__ pos4,
                                  float4* __restrict__ vel4, float4* __restrict__ frc4,
                                  int* __restrict__ id, int n, int T) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    AosParticle p = aos[k];
    int c = counts ? counts[k] : 0;
    // the reference's own spawn can produce ptype == numTypes (curand_uniform may return 1.0, .cu:62), which
    // indexes its tables out of bounds; here every kernel's tables are sized T, so the type is clamped on entry
    p.ptype = p.ptype < (uint32_t)T ? p.ptype : (uint32_t)(T - 1);
    pos4[k] = make_float4(p.pos[0], p.pos[1], p.pos[2], __uint_as_float(p.ptype));
    vel4[k] = make_float4(p.vel[0], p.vel[1], p.vel[2], __int_as_float(c));
    frc4[k] = make_float4(p.acc[0], p.acc[1], p.acc[2], __int_as_float(c));
    id[k] = ids ? ids[k] : k;
}

// Device SoA -> AoS in ORIGINAL particle order (scatter by id), or slot order when
// by_id == 0 (multi-GPU download, ids returned separately).
__global__ void soa_to_aos_kernel(const float4* __restrict__ pos4, const float4* __restrict__ vel4,
                                  const float4* __restrict__ frc4, const int* __restrict__ id,
                                  AosParticle* __restrict__ aos, int* __restrict__ counts, int n,
                                  int by_id) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float4 p = pos4[k], v = vel4[k], f = frc4[k];
    AosParticle o;
    o.pos[0] = p.x, o.pos[1] = p.y, o.pos[2] = p.z;
    o.vel[0] = v.x, o.vel[1] = v.y, o.vel[2] = v.z;
    o.acc[0] = f.x, o.acc[1] = f.y, o.acc[2] = f.z;
    o.ptype = __float_as_uint(p.w);
    o.pad = 0.f;
    int dst = by_id ? id[k] : k;
    aos[dst] = o;
    if (counts) counts[dst] = __float_as_int(v.w);
}

__global__ void scatter_counts_kernel(const int* __restrict__ counts, const int* __restrict__ id,
                                      float4* __restrict__ vel4, int n) {
    // counts are given in ORIGINAL order; slot k holds particle id[k]
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    vel4[k].w = __int_as_float(counts[id[k]]);
}

// moveParticlesKernel (.cu:169-185): pos = fmodf((pos + d) + W, W).
__global__ void move_universe_kernel(float4* __restrict__ pos4, int n, float dx, float dy, float dz,
                                     float W0, float W1, float W2) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float4 p = pos4[k];
    p.x = fmodf(__fadd_rn(__fadd_rn(p.x, dx), W0), W0);
    p.y = fmodf(__fadd_rn(__fadd_rn(p.y, dy), W1), W1);
    p.z = fmodf(__fadd_rn(__fadd_rn(p.z, dz), W2), W2);
    pos4[k] = p;
}

// Key source of the fused first sort pass (kernels_sort.cuh, rs_hist_kernel<.., true>): key = cf_sort_key.
struct CellKeyFn {
    const float4* pos4;
    StepConst c;
    __device__ __forceinline__ uint32_t operator()(int i) const { return cf_sort_key(pos4[i], c); }
};

// Reorder gather + cell bounds in one launch.
//   thread s < n      : slot s of the new order takes old slot perm[s];
//   thread c <= ncell : cellStart[c] = first slot whose key is >= c*64 (lower bound over the sorted keys);
//                       cell c occupies slots [cellStart[c], cellStart[c+1]).  One thread per cell, log2(n)
//                       probes of an L2-resident array; no worst case for clustered states (unlike a
//                       per-particle gap fill).
// n may live on the device (dn; slab mode).  Only the first n_keys sorted keys take part in the bounds
// (slab mode: the leavers' keys, class >= 1, sit behind them); *n_out = cellStart[ncell] - base when given.
__global__ void reorder_bounds_kernel(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ skeys,
                                      const float4* __restrict__ pos_in, const float4* __restrict__ vel_in,
                                      const int* __restrict__ id_in, float4* __restrict__ pos_out,
                                      float4* __restrict__ vel_out, int* __restrict__ id_out, int n_upper,
                                      const int* __restrict__ dn, int* __restrict__ cell_start, int ncell, int base,
                                      int* __restrict__ n_out) {
    const int n = dn ? min(max(*dn, 0), n_upper) : n_upper;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s <= ncell) {
        const uint32_t want = (uint32_t)s * CF_KEY_SUB; // first key of cell s
        int lo = 0, hi = n;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (skeys[mid] < want) lo = mid + 1; else hi = mid;
        }
        cell_start[s] = base + lo;
        if (s == ncell && n_out) *n_out = lo;
    }
    if (s < n) {
        const uint32_t src = perm[s];
        pos_out[s] = pos_in[src];
        vel_out[s] = vel_in[src];
        id_out[s] = id_in[src];
    }
}

// Reorder gather alone (slab initialisation): slot s of the new order takes old slot perm[s].
__global__ void reorder_kernel(const uint32_t* __restrict__ perm, const float4* __restrict__ pos_in,
                               const float4* __restrict__ vel_in, const int* __restrict__ id_in,
                               float4* __restrict__ pos_out, float4* __restrict__ vel_out,
                               int* __restrict__ id_out, int n) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    uint32_t src = perm[s];
    pos_out[s] = pos_in[src];
    vel_out[s] = vel_in[src];
    id_out[s] = id_in[src];
}

__global__ void fill_int_kernel(int* __restrict__ a, int n, int v) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) a[k] = v;
}

// Fused density-adaptive scale + friction + integrate + wrap (ParticleSimulation.cu:136-161),
// rounding points as the reference compiles them:
//   avg = (float)(count + prev) * 0.5;  dens = min(avg / maxExpected, 1)       (IEEE divide)
//   a = fma(dens, -(1 - balance), 1);   m = a * forceMultiplier;   acc = m * F
//   vel = fma(vel, friction, acc * dt);  pos = fma(vel, dt, pos);  pos = fmodf(pos + W, W)
// HBM traffic: reads pos 16 + vel 16 + frc 16, writes pos 16 + vel 16 + frc 16 = 96 B/particle.
__device__ __forceinline__ void cf_integrate_particle(float4& p, float4& v, float4& f, const StepConst& c) {
    int count = __float_as_int(f.w);
    int prev = __float_as_int(v.w);
    float avg = __fmul_rn((float)(count + prev), 0.5f);
    float dens = fminf(__fdiv_rn(avg, c.max_expected), 1.0f);
    float a = __fmaf_rn(dens, -c.one_minus_balance, 1.0f);
    float m = __fmul_rn(a, c.force_multiplier);
    float ax = __fmul_rn(m, f.x), ay = __fmul_rn(m, f.y), az = __fmul_rn(m, f.z);
    v.x = __fmaf_rn(v.x, c.friction, __fmul_rn(ax, c.dt));
    v.y = __fmaf_rn(v.y, c.friction, __fmul_rn(ay, c.dt));
    v.z = __fmaf_rn(v.z, c.friction, __fmul_rn(az, c.dt));
    p.x = fmodf(__fadd_rn(__fmaf_rn(v.x, c.dt, p.x), c.W[0]), c.W[0]);
    p.y = fmodf(__fadd_rn(__fmaf_rn(v.y, c.dt, p.y), c.W[1]), c.W[1]);
    p.z = fmodf(__fadd_rn(__fmaf_rn(v.z, c.dt, p.z), c.W[2]), c.W[2]);
    v.w = __int_as_float(count);
    f = make_float4(ax, ay, az, f.w);
}

__global__ void integrate_kernel(float4* __restrict__ pos4, float4* __restrict__ vel4,
                                 float4* __restrict__ frc4, int n, StepConst c) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float4 p = pos4[k], v = vel4[k], f = frc4[k];
    cf_integrate_particle(p, v, f, c);
    pos4[k] = p;
    vel4[k] = v;
    frc4[k] = f;
}

// Render feed (CellFlowWidget::updateParticleBuffer, CellFlowWidget.cpp:742-761, and
// getParticleTypeCounts, :875-886): (x, y, z, (float)type) per particle in ORIGINAL order plus
// per-type counts, produced on the device — 16 B per particle instead of the 44-byte AoS round
// trip and CPU repack the reference does every frame.
__global__ void render_feed_kernel(const float4* __restrict__ pos4, const int* __restrict__ id, int n, int by_id,
                                   float4* __restrict__ out, int* __restrict__ type_counts, int T) {
    __shared__ int hist[CF_T_MAX];
    if (threadIdx.x < CF_T_MAX) hist[threadIdx.x] = 0;
    __syncthreads();
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        float4 p = pos4[k];
        uint32_t t = __float_as_uint(p.w);
        out[by_id ? id[k] : k] = make_float4(p.x, p.y, p.z, (float)t);
        if (t < (uint32_t)T) atomicAdd(&hist[t], 1); // the widget ignores ptype >= numTypes too
    }
    __syncthreads();
    if (threadIdx.x < T && hist[threadIdx.x]) atomicAdd(&type_counts[threadIdx.x], hist[threadIdx.x]);
}

// Sum of neighbour counts (accepted ordered pairs) for the roofline figure.
__global__ void sum_counts_kernel(const float4* __restrict__ frc4, int n, unsigned long long* out) {
    unsigned long long local = 0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
        local += (unsigned long long)__float_as_int(frc4[k].w);
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, local);
}
