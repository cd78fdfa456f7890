// kernels_force.cuh — pairwise force + local-density count over the cell list.
//
// Law (ParticleSimulation.cu:86-133): for every other particle o within the type-pair radius
// Reff of p (minimum-image distance, dist = sqrt(d2 + 1e-4) < Reff): count++,
// F += dir * (repulsion * exp(-k r^2) - attraction * r) * forceTable[p.type][o.type], r = dist/Reff.
// Output per slot: frc4 = (Fx, Fy, Fz, bits(count)); the density scaling, friction and
// integration of .cu:136-161 are the integrate kernel's job.
#pragma once
#include "cf_device.cuh"

// Distinct neighbour coordinates of cell c along one axis of n cells (periodic or clamped).
// Returns how many (1..3); out[] holds them; wrap handling of the displacement is per pair.
__device__ __forceinline__ int cf_axis_cells(int c, int n, bool periodic, int out[3]) {
    if (n >= 3) {
        int k = 0;
        if (c > 0) out[k++] = c - 1; else if (periodic) out[k++] = n - 1;
        out[k++] = c;
        if (c < n - 1) out[k++] = c + 1; else if (periodic) out[k++] = 0;
        return k;
    }
    if (n == 2) {
        out[0] = c;
        if (periodic) { out[1] = 1 - c; return 2; }
        out[1] = 1 - c; // clamped grid of 2 cells: both are neighbours anyway
        return 2;
    }
    out[0] = c;
    return 1;
}

// ---- v1: one thread per particle, candidates streamed from the sorted array through L1/L2 ----
// Robust at any density (also the path for sparse grids, where cells hold a handful of
// particles and a CTA-per-cell tile would idle).  Threads of a warp are consecutive slots, i.e.
// the same or adjacent cells, so candidate loads are mostly warp-uniform broadcasts.
template <bool UNIFORM>
__global__ void __launch_bounds__(128)
force_pp_kernel(const float4* __restrict__ pos4, const int* __restrict__ cell_start,
                float4* __restrict__ frc4, int first, int n_upper, const int* __restrict__ dn, StepConst c,
                const DeviceTables* __restrict__ tables) {
    __shared__ float s_cut2[CF_TT_MAX], s_inv[CF_TT_MAX], s_fv[CF_TT_MAX];
    for (int i = threadIdx.x; i < c.T * c.T; i += blockDim.x) {
        s_cut2[i] = tables->cut2[i];
        s_inv[i] = tables->inv_reff[i];
        s_fv[i] = tables->force[i];
    }
    __syncthreads();
    const int n = dn ? min(*dn, n_upper) : n_upper; // slab mode: the owned count lives on the device
    int s = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= first + n) return;
    float4 p = pos4[s];
    int ti = (int)__float_as_uint(p.w);
    int cx = cf_cell_coord_x(p.x, c);
    int cy = cf_cell_coord(p.y, c.inv[1], c.dims[1]);
    int cz = cf_cell_coord(p.z, c.inv[2], c.dims[2]);
    int xs[3], ys[3], zs[3];
    int kx = cf_axis_cells(cx, c.dims[0], c.periodic_x != 0, xs);
    int ky = cf_axis_cells(cy, c.dims[1], true, ys);
    int kz = cf_axis_cells(cz, c.dims[2], true, zs);
    float fx = 0.f, fy = 0.f, fz = 0.f;
    int count = 0;
    const float* cut_row = s_cut2 + ti * c.T;
    const float* inv_row = s_inv + ti * c.T;
    const float* fv_row = s_fv + ti * c.T;
    // z-adjacent cells are adjacent in the sorted array: the <= 3 z cells of an (x, y) row are one or two slot
    // ranges (two when the periodic wrap splits them), walked in the same order as cell by cell — the summation
    // order, hence every result bit, is unchanged.  The kernel is bound by the latency of dependent loads (bounds
    // -> positions) at a handful of particles per cell, so one range per row instead of three cells, and the
    // bounds of the next range are fetched before the current one is walked (27 -> <= 9-18 rounds, half of them
    // hidden): pair force 0.092 -> 0.048 ms at BASELINE config 2 (100 k particles, ~4 per cell).
    int seg_lo[3], seg_hi[3], nseg = 0;
    for (int q = 0; q < kz; q++) {
        if (nseg > 0 && zs[q] == seg_hi[nseg - 1] + 1) seg_hi[nseg - 1] = zs[q];
        else seg_lo[nseg] = seg_hi[nseg] = zs[q], nseg++;
    }
    int a = 0, b = 0, sg = 0; // cursor of the NEXT range
    int row = (xs[0] * c.dims[1] + ys[0]) * c.dims[2];
    int nj0 = cell_start[row + seg_lo[0]], nj1 = cell_start[row + seg_hi[0] + 1];
    const int total = kx * ky * nseg;
    for (int it = 0; it < total; it++) {
        const int j0 = nj0, j1 = nj1;
        if (it + 1 < total) {
            if (++sg == nseg) {
                sg = 0;
                if (++b == ky) b = 0, a++;
                row = (xs[a] * c.dims[1] + ys[b]) * c.dims[2];
            }
            nj0 = cell_start[row + seg_lo[sg]], nj1 = cell_start[row + seg_hi[sg] + 1];
        }
        for (int j = j0; j < j1; j++) {
            float4 o = pos4[j];
            float dx = cf_wrap(__fsub_rn(o.x, p.x), c.W[0], c.halfW[0], c.nhalfW[0]);
            float dy = cf_wrap(__fsub_rn(o.y, p.y), c.W[1], c.halfW[1], c.nhalfW[1]);
            float dz = cf_wrap(__fsub_rn(o.z, p.z), c.W[2], c.halfW[2], c.nhalfW[2]);
            float d2 = cf_dist2(dx, dy, dz);
            int tj = (int)__float_as_uint(o.w);
            float cut = UNIFORM ? c.cut2_uniform : cut_row[tj];
            if (d2 < cut && j != s) {
                count++;
                float inv = UNIFORM ? c.inv_reff_uniform : inv_row[tj];
                cf_pair_force(dx, dy, dz, d2, inv, fv_row[tj], c, fx, fy, fz);
            }
        }
    }
    frc4[s] = make_float4(fx, fy, fz, __int_as_float(count));
}
