// kernels_slab.cuh — device side of the multi-GPU slab decomposition (no reference counterpart;
// SURVEY.md section 8e).  Rank r of G owns x in [bound(r), bound(r+1)); its cell grid has the
// owned x layers 1..nxl plus one ghost layer on each side (layer 0, layer nxl+1).
//
// Per step, before the force pass:
//   1. every owned particle gets a class (stay / to-left / to-right) from its x; the class is
//      the most significant part of the sort key, so ONE stable radix sort both orders the
//      stayers by (cell, type) and leaves the leavers as two contiguous tails;
//   2. leavers (pos, vel+prevCount, id = 36 B) go to the ring neighbours, arrivals are sorted
//      and merged into the stayers (stable, deterministic);
//   3. the first and last owned x layer — contiguous slot ranges, because x is the slowest cell
//      index — are copied to the neighbours as ghosts (pos+type, id = 20 B) and land directly
//      before / after the owned slots, already in cell order.
#pragma once
#include "cf_device.cuh"

#define SLAB_STAY 0u
#define SLAB_LEFT 1u
#define SLAB_RIGHT 2u

struct SlabGeom {
    float x_lo, x_hi;   // owned interval (global coordinates), x_hi == neighbour's x_lo bit for bit
    float slab_w;       // x_hi - x_lo in real terms (W / G)
    float W;            // global width
    uint32_t class_stride; // keys per class = ncell * 64
};

// Fixed-capacity messages with the element count in-band (no size exchange, no host sync):
//   migrants: pos4[cap] | vel4[cap] | id[cap] | count
//   halo    : pos4[cap] | id[cap]  | count
__host__ __device__ inline size_t slab_mig_bytes(int cap) { return (size_t)cap * 36 + 16; }
__host__ __device__ inline size_t slab_halo_bytes(int cap) { return (size_t)cap * 20 + 16; }
__host__ __device__ inline float4* mig_pos(char* m, int) { return (float4*)m; }
__host__ __device__ inline float4* mig_vel(char* m, int cap) { return (float4*)(m + (size_t)cap * 16); }
__host__ __device__ inline int* mig_id(char* m, int cap) { return (int*)(m + (size_t)cap * 32); }
__host__ __device__ inline int* mig_count(char* m, int cap) { return (int*)(m + (size_t)cap * 36); }
__host__ __device__ inline float4* halo_pos(char* m, int) { return (float4*)m; }
__host__ __device__ inline int* halo_id(char* m, int cap) { return (int*)(m + (size_t)cap * 16); }
__host__ __device__ inline int* halo_count(char* m, int cap) { return (int*)(m + (size_t)cap * 20); }

__device__ __forceinline__ uint32_t slab_class(float x, const SlabGeom& g, int* err) {
    if (x >= g.x_lo && x < g.x_hi) return SLAB_STAY;
    float d = x - g.x_lo;
    if (d < 0.f) d += g.W;
    if (d < 2.0f * g.slab_w) return SLAB_RIGHT; // one slab to the right (also covers G == 2)
    if (d >= g.W - 1.5f * g.slab_w) return SLAB_LEFT;
    if (err) atomicExch(err, 1); // moved further than one slab in a step
    return SLAB_STAY;
}

// key = class * (ncell*64) + cf_sort_key for owned particle i (slots base+i), val = i.
__global__ void slab_key_kernel(const float4* __restrict__ pos4, uint32_t* __restrict__ keys,
                                uint32_t* __restrict__ vals, int n, StepConst c, SlabGeom g,
                                int* __restrict__ err) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float4 p = pos4[k];
    uint32_t cls = slab_class(p.x, g, err);
    uint32_t key = cf_sort_key(p, c); // leavers: clamped cell, irrelevant
    keys[k] = cls * g.class_stride + key;
    vals[k] = (uint32_t)k;
}

// counts[0..2] = number of stay / left / right keys in the sorted key array.
__global__ void slab_class_counts_kernel(const uint32_t* __restrict__ skeys, int n, uint32_t class_stride,
                                         int* __restrict__ counts) {
    int t = threadIdx.x;
    if (t >= 2) return;
    uint32_t want = (uint32_t)(t + 1) * class_stride;
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (skeys[mid] < want) lo = mid + 1; else hi = mid;
    }
    __shared__ int b[2];
    b[t] = lo;
    __syncthreads();
    if (t == 0) {
        counts[0] = b[0];
        counts[1] = b[1] - b[0];
        counts[2] = n - b[1];
    }
}

// Gather the two leaver tails (sorted order) into the send messages.  The class counts are read
// from device memory (counts[0..2] = stay, left, right), so the host does not have to wait for
// them before the exchange is enqueued; a count above the message capacity is clamped here and
// reported by the host after its (single) synchronisation.
__global__ void slab_pack_migrants_kernel(const uint32_t* __restrict__ perm, const float4* __restrict__ pos4,
                                          const float4* __restrict__ vel4, const int* __restrict__ id,
                                          const int* __restrict__ counts, char* __restrict__ msg_left,
                                          char* __restrict__ msg_right, int cap) {
    const int n_stay = counts[0], n_left = min(counts[1], cap), n_right = min(counts[2], cap);
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) {
        *mig_count(msg_left, cap) = n_left;
        *mig_count(msg_right, cap) = n_right;
    }
    if (k >= n_left + n_right) return;
    char* msg = k < n_left ? msg_left : msg_right;
    int m = k < n_left ? k : k - n_left;
    uint32_t src = perm[n_stay + (k < n_left ? k : counts[1] + (k - n_left))];
    mig_pos(msg, cap)[m] = pos4[src];
    mig_vel(msg, cap)[m] = vel4[src];
    mig_id(msg, cap)[m] = id[src];
}

// Append arrivals behind the current owned particles (slots n .. n+nA) and emit their sort
// pairs (class = stay).
__global__ void slab_unpack_arrivals_kernel(char* __restrict__ msg_from_left, char* __restrict__ msg_from_right,
                                            int n_al, int n_ar, int cap, float4* __restrict__ pos4,
                                            float4* __restrict__ vel4, int* __restrict__ id, int n,
                                            uint32_t* __restrict__ akeys, uint32_t* __restrict__ avals,
                                            StepConst c) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_al + n_ar) return;
    char* msg = k < n_al ? msg_from_left : msg_from_right;
    int m = k < n_al ? k : k - n_al;
    float4 p = mig_pos(msg, cap)[m];
    pos4[n + k] = p;
    vel4[n + k] = mig_vel(msg, cap)[m];
    id[n + k] = mig_id(msg, cap)[m];
    akeys[k] = cf_sort_key(p, c);
    avals[k] = (uint32_t)(n + k);
}

// Stable merge of the sorted stayers S (first) with the sorted arrivals A: rank by binary search.
__global__ void slab_merge_kernel(const uint32_t* __restrict__ skeys, const uint32_t* __restrict__ svals, int ns,
                                  const uint32_t* __restrict__ akeys, const uint32_t* __restrict__ avals, int na,
                                  uint32_t* __restrict__ okeys, uint32_t* __restrict__ ovals) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < ns) {
        uint32_t key = skeys[k];
        int lo = 0, hi = na; // arrivals with key < mine go first
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (akeys[mid] < key) lo = mid + 1; else hi = mid;
        }
        okeys[k + lo] = key;
        ovals[k + lo] = svals[k];
    } else if (k < ns + na) {
        int a = k - ns;
        uint32_t key = akeys[a];
        int lo = 0, hi = ns; // stayers with key <= mine go first
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (skeys[mid] <= key) lo = mid + 1; else hi = mid;
        }
        okeys[a + lo] = key;
        ovals[a + lo] = avals[a];
    }
}

// Copy the first owned x layer (-> left neighbour) and the last one (-> right neighbour).
// Layer ranges come from cell_start on the device; counts travel in-band.
__global__ void slab_pack_halo_kernel(const float4* __restrict__ pos4, const int* __restrict__ id,
                                      const int* __restrict__ cell_start, int layer_cells, int nxl,
                                      char* __restrict__ msg_left, char* __restrict__ msg_right, int cap,
                                      int* __restrict__ err) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    int l0 = cell_start[layer_cells], l1 = cell_start[2 * layer_cells];
    int r0 = cell_start[nxl * layer_cells], r1 = cell_start[(nxl + 1) * layer_cells];
    int nl = l1 - l0, nr = r1 - r0;
    if (k == 0) {
        if (nl > cap || nr > cap) atomicExch(err, 2);
        *halo_count(msg_left, cap) = min(nl, cap);
        *halo_count(msg_right, cap) = min(nr, cap);
    }
    if (k < min(nl, cap)) {
        halo_pos(msg_left, cap)[k] = pos4[l0 + k];
        halo_id(msg_left, cap)[k] = id[l0 + k];
    }
    if (k < min(nr, cap)) {
        halo_pos(msg_right, cap)[k] = pos4[r0 + k];
        halo_id(msg_right, cap)[k] = id[r0 + k];
    }
}

// Ghosts from the left neighbour end right before the owned slots, ghosts from the right
// neighbour start right after them; their keys (forced ghost x layer) feed the bounds search.
__global__ void slab_unpack_ghosts_kernel(char* __restrict__ msg_from_left, char* __restrict__ msg_from_right,
                                          int cap, float4* __restrict__ pos4, int* __restrict__ id, int own_first,
                                          int n_own, uint32_t* __restrict__ gkeys_left,
                                          uint32_t* __restrict__ gkeys_right, StepConst c) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    int nl = *halo_count(msg_from_left, cap), nr = *halo_count(msg_from_right, cap);
    if (k < nl) {
        float4 p = halo_pos(msg_from_left, cap)[k];
        int slot = own_first - nl + k;
        pos4[slot] = p;
        id[slot] = halo_id(msg_from_left, cap)[k];
        int cy = cf_cell_coord(p.y, c.inv[1], c.dims[1]), cz = cf_cell_coord(p.z, c.inv[2], c.dims[2]);
        gkeys_left[k] = (uint32_t)((0 * c.dims[1] + cy) * c.dims[2] + cz) * CF_KEY_SUB; // cell part only
    }
    if (k < nr) {
        float4 p = halo_pos(msg_from_right, cap)[k];
        int slot = own_first + n_own + k;
        pos4[slot] = p;
        id[slot] = halo_id(msg_from_right, cap)[k];
        int cy = cf_cell_coord(p.y, c.inv[1], c.dims[1]), cz = cf_cell_coord(p.z, c.inv[2], c.dims[2]);
        gkeys_right[k] = (uint32_t)(((c.dims[0] - 1) * c.dims[1] + cy) * c.dims[2] + cz) * CF_KEY_SUB;
    }
}

// cell_start of the two ghost layers (lower bounds over the ghost key arrays).
__global__ void slab_ghost_bounds_kernel(const uint32_t* __restrict__ gkeys_left,
                                         const uint32_t* __restrict__ gkeys_right, char* __restrict__ msg_from_left,
                                         char* __restrict__ msg_from_right, int cap, int* __restrict__ cell_start,
                                         int layer_cells, int ncell, int own_first, int n_own,
                                         int* __restrict__ ghost_counts) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    int nl = *halo_count(msg_from_left, cap), nr = *halo_count(msg_from_right, cap);
    if (k == 0) {
        ghost_counts[0] = nl;
        ghost_counts[1] = nr;
    }
    if (k < layer_cells) { // cells of layer 0: c = k
        uint32_t want = (uint32_t)k * CF_KEY_SUB;
        int lo = 0, hi = nl;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (gkeys_left[mid] < want) lo = mid + 1; else hi = mid;
        }
        cell_start[k] = own_first - nl + lo;
    } else if (k < 2 * layer_cells + 1) { // cells of the last layer plus the end sentinel
        int c = ncell - layer_cells + (k - layer_cells);
        uint32_t want = (uint32_t)c * CF_KEY_SUB;
        int lo = 0, hi = nr;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (gkeys_right[mid] < want) lo = mid + 1; else hi = mid;
        }
        cell_start[c] = own_first + n_own + lo;
    }
}

// Initial conditions in slab mode: generate global particle k, keep it when it falls in the slab.
// key 0 = mine, 1 = not mine; a stable 1-bit radix pass then compacts "mine" in id order.
__global__ void slab_init_class_kernel(const float4* __restrict__ pos4, uint32_t* __restrict__ keys,
                                       uint32_t* __restrict__ vals, int n, SlabGeom g) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float x = pos4[k].x;
    keys[k] = (x >= g.x_lo && x < g.x_hi) ? 0u : 1u;
    vals[k] = (uint32_t)k;
}
