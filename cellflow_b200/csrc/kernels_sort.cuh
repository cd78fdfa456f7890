// kernels_sort.cuh — stable LSD radix sort of (cell key, slot) pairs: the cell-list build.
//
// Keys are cell indices (< ncell), so only ceil(log2(ncell)) bits are sorted, in passes of at
// most 8 bits.  Each pass is histogram -> exclusive scan -> stable scatter; the sort is stable
// and therefore deterministic: the slot order inside a cell (hence the force summation order)
// is a pure function of the input, run to run and rank to rank.
//
// HBM traffic per pass: histogram reads 4 B/key; scatter reads 8 B and writes 8 B per pair.
#pragma once
#include "cf_device.cuh"

#define RS_THREADS 256
#define RS_WARPS (RS_THREADS / 32)
#define RS_BINS 256

// hist[d * nblocks + b] = number of keys of block b whose digit is d.
__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const uint32_t* __restrict__ keys, int n, int shift, uint32_t mask,
               uint32_t* __restrict__ hist, int nblocks, int items_per_block) {
    __shared__ uint32_t sh[RS_BINS];
    sh[threadIdx.x] = 0;
    __syncthreads();
    int base = blockIdx.x * items_per_block;
    int end = min(base + items_per_block, n);
    for (int i = base + threadIdx.x; i < end; i += RS_THREADS) {
        uint32_t d = (keys[i] >> shift) & mask;
        // warp-aggregate: nearly-sorted input puts whole warps into one bin
        uint32_t peers = __match_any_sync(__activemask(), d);
        if ((peers & ((1u << (threadIdx.x & 31)) - 1u)) == 0) atomicAdd(&sh[d], __popc(peers));
    }
    __syncthreads();
    hist[threadIdx.x * nblocks + blockIdx.x] = sh[threadIdx.x];
}

// In-place exclusive scan of `total` counters by one block of 1024 threads, in coalesced tiles of
// 16384 (four uint4 per thread, each uint4 load/store warp-contiguous): thread-local prefix,
// warp shuffle scan, 32 warp totals scanned by warp 0, running carry between tiles.  (A
// chunk-per-thread version read with a stride of total/1024 words and took 52 us for 62 k
// counters; a 4096-wide tiled one was bound by its three barriers per tile.)
#define RS_SCAN_VEC 4
__global__ void __launch_bounds__(1024) rs_scan_kernel(uint32_t* __restrict__ hist, int total) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    const bool vec_ok = (total & 3) == 0;
    for (int base = 0; base < total; base += 4096 * RS_SCAN_VEC) {
        // thread owns RS_SCAN_VEC groups of 4 consecutive counters: group g starts at
        // base + g*4096 + 4*tid, so the scan order inside a tile is (g, tid, element)
        uint32_t v[RS_SCAN_VEC][4];
        uint32_t gsum[RS_SCAN_VEC];
#pragma unroll
        for (int g = 0; g < RS_SCAN_VEC; g++) {
            const int i = base + g * 4096 + 4 * tid;
            v[g][0] = v[g][1] = v[g][2] = v[g][3] = 0;
            if (vec_ok && i + 3 < total) {
                uint4 q = *reinterpret_cast<const uint4*>(hist + i);
                v[g][0] = q.x, v[g][1] = q.y, v[g][2] = q.z, v[g][3] = q.w;
            } else {
#pragma unroll
                for (int e = 0; e < 4; e++)
                    if (i + e < total) v[g][e] = hist[i + e];
            }
            gsum[g] = v[g][0] + v[g][1] + v[g][2] + v[g][3];
        }
        const uint32_t carry = carry_s; // final since the barrier that ended the previous tile
        // inclusive scan of every group across the block (4 independent scans share the barriers)
        uint32_t incl[RS_SCAN_VEC];
#pragma unroll
        for (int g = 0; g < RS_SCAN_VEC; g++) {
            uint32_t x = gsum[g];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += t;
            }
            incl[g] = x;
        }
        __shared__ uint32_t ws[RS_SCAN_VEC][32];
        if (lane == 31) {
#pragma unroll
            for (int g = 0; g < RS_SCAN_VEC; g++) ws[g][warp] = incl[g];
        }
        __syncthreads();
        if (warp < RS_SCAN_VEC) { // warp g scans the 32 warp totals of group g
            uint32_t w = ws[warp][lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            ws[warp][lane] = wi - w;
            if (lane == 31) warp_sums[warp] = wi; // total of group `warp`
        }
        __syncthreads();
        uint32_t gbase = carry;
#pragma unroll
        for (int g = 0; g < RS_SCAN_VEC; g++) {
            const int i = base + g * 4096 + 4 * tid;
            uint32_t run = gbase + ws[g][warp] + incl[g] - gsum[g];
            if (vec_ok && i + 3 < total) {
                uint4 o4 = make_uint4(run, run + v[g][0], run + v[g][0] + v[g][1], run + v[g][0] + v[g][1] + v[g][2]);
                *reinterpret_cast<uint4*>(hist + i) = o4;
            } else {
                uint32_t r = run;
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    if (i + e < total) hist[i + e] = r;
                    r += v[g][e];
                }
            }
            gbase += warp_sums[g];
        }
        __syncthreads(); // everyone has read ws / warp_sums / carry_s
        if (tid == 0) carry_s = gbase;
        __syncthreads();
    }
}

// Stable scatter.  A block walks its items in rounds of RS_THREADS, in index order.  Per round:
// every warp ranks its lanes per digit with match_any, the first lane of each digit group
// publishes the group size, thread d (owner of digit d) turns the 8 per-warp sizes into bases
// on top of the scanned histogram plus what earlier rounds of this block already placed.
__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                  uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int n, int shift,
                  uint32_t mask, const uint32_t* __restrict__ hist, int nblocks, int items_per_block) {
    __shared__ uint32_t cnt[2][RS_WARPS][RS_BINS];
    int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int w = 0; w < RS_WARPS; w++) {
        cnt[0][w][tid] = 0;
        cnt[1][w][tid] = 0;
    }
    uint32_t running = hist[tid * nblocks + blockIdx.x]; // digit `tid`: next free output slot
    __syncthreads();
    int base = blockIdx.x * items_per_block;
    int rounds = items_per_block / RS_THREADS;
    for (int r = 0; r < rounds; r++) {
        int buf = r & 1;
        int i = base + r * RS_THREADS + tid;
        bool valid = i < n;
        uint32_t key = valid ? keys_in[i] : 0xffffffffu;
        uint32_t val = valid ? vals_in[i] : 0u;
        uint32_t d = valid ? ((key >> shift) & mask) : RS_BINS; // invalid lanes group apart
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) cnt[buf][warp][d] = __popc(peers);
        __syncthreads();
        {
            uint32_t run = running;
#pragma unroll
            for (int w = 0; w < RS_WARPS; w++) {
                uint32_t c = cnt[buf][w][tid];
                cnt[buf][w][tid] = run;
                run += c;
                cnt[buf ^ 1][w][tid] = 0;
            }
            running = run;
        }
        __syncthreads();
        if (valid) {
            uint32_t dst = cnt[buf][warp][d] + rank;
            keys_out[dst] = key;
            vals_out[dst] = val;
        }
        // the next round writes cnt[buf^1] (zeroed above, before the barrier) and only reads
        // cnt[buf] again two rounds later, after two more barriers
    }
}
