// kernels_graph.cuh — same-type proximity graph over the cell list, warp-aggregated compaction.
//
// Rule (generateProximityGraphKernel, ParticleSimulation.cu:212-246), per particle i:
//   candidates = the first 2*maxConn particles j > i IN ORIGINAL INDEX ORDER with the same type
//   and plain (non-wrapped) d2 = fma(dz,dz,fma(dx,dx,dy*dy)) < dist^2; stable insertion sort by
//   d2; keep the first maxConn.
// The reference scans all j > i; here candidates come from the (2m+1)^3 cells around i (clamped,
// not periodic, because the rule does not wrap), the 2*maxConn smallest original ids are kept in
// a sorted per-thread list, and the same stable sort by d2 selects the edges.  The edge SET is
// identical; the reference's own edge ORDER is nondeterministic (one atomicAdd per thread).
#pragma once
#include "cf_device.cuh"

#define CF_GRAPH_K 32 // 2 * CF_MAX_GRAPH_CONN candidates, the reference's nearby[32] (.cu:210)

__global__ void __launch_bounds__(128)
graph_kernel(const float4* __restrict__ pos4, const int* __restrict__ id, const int* __restrict__ cell_start,
             int first, int n, StepConst c, float dist2, int max_conn, int m, int2* __restrict__ edges,
             int2* __restrict__ edge_slots, int capacity, int* __restrict__ edge_count) {
    int s = first + blockIdx.x * blockDim.x + threadIdx.x;
    bool active = s < first + n;
    int cand_id[CF_GRAPH_K];
    int cand_slot[CF_GRAPH_K];
    float cand_d2[CF_GRAPH_K];
    int ncand = 0;
    int my_id = 0;
    const int K = 2 * max_conn;
    if (active) {
        float4 p = pos4[s];
        my_id = id[s];
        uint32_t ti = __float_as_uint(p.w);
        int cx = cf_cell_coord_x(p.x, c);
        int cy = cf_cell_coord(p.y, c.inv[1], c.dims[1]);
        int cz = cf_cell_coord(p.z, c.inv[2], c.dims[2]);
        int x0 = max(cx - m, c.gx_lo), x1 = min(cx + m, c.gx_hi);
        int y0 = max(cy - m, 0), y1 = min(cy + m, c.dims[1] - 1);
        int z0 = max(cz - m, 0), z1 = min(cz + m, c.dims[2] - 1);
        for (int x = x0; x <= x1; x++)
            for (int y = y0; y <= y1; y++) {
                int row = (x * c.dims[1] + y) * c.dims[2];
                // z-adjacent cells are contiguous in the sorted array: one range per (x, y)
                int j0 = cell_start[row + z0], j1 = cell_start[row + z1 + 1];
                for (int j = j0; j < j1; j++) {
                    float4 o = pos4[j];
                    if (__float_as_uint(o.w) != ti) continue;
                    int jid = id[j];
                    if (jid <= my_id) continue;
                    float dx = __fsub_rn(o.x, p.x), dy = __fsub_rn(o.y, p.y), dz = __fsub_rn(o.z, p.z);
                    float d2 = cf_dist2(dx, dy, dz);
                    if (!(d2 < dist2)) continue;
                    if (ncand == K && jid > cand_id[K - 1]) continue;
                    // insert into the id-sorted candidate list (drop the largest id when full)
                    int pos = ncand < K ? ncand : K - 1;
                    while (pos > 0 && cand_id[pos - 1] > jid) {
                        cand_id[pos] = cand_id[pos - 1];
                        cand_slot[pos] = cand_slot[pos - 1];
                        cand_d2[pos] = cand_d2[pos - 1];
                        pos--;
                    }
                    cand_id[pos] = jid;
                    cand_slot[pos] = j;
                    cand_d2[pos] = d2;
                    if (ncand < K) ncand++;
                }
            }
        // stable insertion sort by d2 (.cu:235-243); the list is in index order, as the
        // reference's scan would have produced it
        for (int a = 1; a < ncand; a++) {
            float kd = cand_d2[a];
            int ki = cand_id[a], ks = cand_slot[a];
            int b = a - 1;
            while (b >= 0 && cand_d2[b] > kd) {
                cand_d2[b + 1] = cand_d2[b];
                cand_id[b + 1] = cand_id[b];
                cand_slot[b + 1] = cand_slot[b];
                b--;
            }
            cand_d2[b + 1] = kd;
            cand_id[b + 1] = ki;
            cand_slot[b + 1] = ks;
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
            edges[e] = make_int2(my_id, cand_id[a]);
            edge_slots[e] = make_int2(s, cand_slot[a]);
        }
    }
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
