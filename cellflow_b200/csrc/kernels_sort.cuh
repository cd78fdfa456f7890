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
// 4096 (one uint4 per thread): thread-local prefix, warp shuffle scan, 32 warp totals scanned by
// warp 0, running carry between tiles.  (A chunk-per-thread version read with a stride of
// total/1024 words and took 52 us for 62 k counters; this one is bandwidth-limited.)
__global__ void __launch_bounds__(1024) rs_scan_kernel(uint32_t* __restrict__ hist, int total) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < total; base += 4096) {
        const int i = base + 4 * tid;
        uint32_t v0 = 0, v1 = 0, v2 = 0, v3 = 0;
        if (i + 3 < total && (total & 3) == 0) {
            uint4 v = *reinterpret_cast<const uint4*>(hist + i);
            v0 = v.x, v1 = v.y, v2 = v.z, v3 = v.w;
        } else {
            if (i < total) v0 = hist[i];
            if (i + 1 < total) v1 = hist[i + 1];
            if (i + 2 < total) v2 = hist[i + 2];
            if (i + 3 < total) v3 = hist[i + 3];
        }
        const uint32_t sum = v0 + v1 + v2 + v3;
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const uint32_t carry = carry_s; // final since the barrier that ended the previous tile
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_sums[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_sums[lane] = wi - w; // exclusive prefix of the warp totals
            if (lane == 31) carry_s = carry + wi;
        }
        __syncthreads();
        uint32_t run = carry + warp_sums[warp] + incl - sum;
        if (i + 3 < total && (total & 3) == 0) {
            uint4 o4 = make_uint4(run, run + v0, run + v0 + v1, run + v0 + v1 + v2);
            *reinterpret_cast<uint4*>(hist + i) = o4;
        } else {
            if (i < total) hist[i] = run;
            if (i + 1 < total) hist[i + 1] = run + v0;
            if (i + 2 < total) hist[i + 2] = run + v0 + v1;
            if (i + 3 < total) hist[i + 3] = run + v0 + v1 + v2;
        }
        __syncthreads(); // warp_sums / carry_s are rewritten by the next tile
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
