// kernels_tile.cuh — CTA-per-cell tiled pair-force kernel (dense regime).  See DESIGN.md 4.3.
#pragma once
#include "cf_device.cuh"
#include "kernels_force.cuh"

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

static inline bool tile_kernel_applicable(const StepConst&, int, int) { return false; }
static inline int launch_tile_force(cudaStream_t, const float4*, const int*, float4*, int, int,
                                    const StepConst&, const DeviceTables*, long long*) {
    return (int)cudaErrorNotSupported;
}
