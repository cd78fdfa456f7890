// kernels_slab.cuh — device side of the multi-GPU slab decomposition (no reference counterpart;
// SURVEY.md section 8e).  Rank r of G owns x in [bound(r), bound(r+1)); its cell grid has the owned x
// layers 1..nxl plus one ghost layer on each side (layer 0, layer nxl+1).
//
// Round 2 design: NO host in the loop and NO library collective on the step path.  Every rank owns a
// MAILBOX in its own HBM that its two ring neighbours have mapped through CUDA IPC; a message is written
// straight into the receiver's mailbox with ordinary global stores over NVLink 5 by the kernel that
// produces it (the fused integrate kernel emits the migrants, the halo pack kernel copies the boundary
// layer), exact size, followed by a system-scope release of a sequence flag.  The receiver's stream holds a
// one-warp wait kernel that acquires the flag.  All counts (owned particles, migrants, arrivals, ghosts)
// stay on the device; kernels are launched over capacity-sized grids and read them there.
//
// Per step, between integrate(t) and force(t+1):
//   1. integrate classifies every particle by its new x (stay / to-left / to-right) and appends the leavers
//      (pos, vel+prevCount, id = 36 B) to the neighbour's mailbox;
//   2. arrivals are appended behind the owned slots (in id order: deterministic), ONE stable radix sort on
//      key = class * (ncell*64) + cell key orders the stayers + arrivals by cell and moves the leavers behind
//      them; reorder + cell bounds yield the new owned count on the device;
//   3. the first and last owned x layer — contiguous slot ranges, because x is the slowest cell index — are
//      copied into the neighbours' mailboxes (pos+type, id = 20 B) and land directly before / after the
//      owned slots of the receiver, already in cell order.
// Messages are double-buffered by sequence parity: a neighbour can only write message k+2 after it has
// consumed my message k+1, which I sent after consuming its message k (ring handshake, no credits needed).
#pragma once
#include "cf_device.cuh"
#include "kernels_state.cuh"

#define SLAB_STAY 0u
#define SLAB_GONE 1u

// error bits of the device status word (d_slab[SLAB_ERR]); the host reads it where the API synchronises
#define SLAB_ERR_FAR 1       // a particle moved further than the neighbouring slab in one step
#define SLAB_ERR_HALO 2      // a boundary layer exceeded the halo capacity
#define SLAB_ERR_MIG 4       // more leavers than the migrant capacity
#define SLAB_ERR_OWN 8       // a rank would own more particles than its capacity
#define SLAB_ERR_TIMEOUT 16  // a neighbour's message did not arrive in time

// device status / count words
enum { SLAB_NCUR = 0, SLAB_NTMP = 1, SLAB_ERR = 2, SLAB_GHOST_L = 4, SLAB_GHOST_R = 5, SLAB_SEND_L = 6, SLAB_SEND_R = 7,
       SLAB_TICKET_MIG = 8, SLAB_TICKET_HALO = 9, SLAB_ARR_L = 10, SLAB_ARR_R = 11,
       SLAB_FIRST = 12,  // first used slot (start of the left ghost layer)
       SLAB_NSLOTS = 13, // used slots: left ghosts + owned + right ghosts
       SLAB_WORDS = 32 };

struct SlabGeom {
    float x_lo, x_hi;       // owned interval (global coordinates), x_hi == right neighbour's x_lo bit for bit
    float w_own, w_left, w_right; // widths of this slab and of its two ring neighbours
    float W;                // global width
    uint32_t class_stride;  // keys per class = ncell * 64
};

// Mailbox layout, identical on every rank.  side 0 = "from my left neighbour", side 1 = "from my right
// neighbour"; parity = sequence number & 1.
//   ctl   : int flag_mig[2], flag_halo[2], cnt_mig[2][2], cnt_halo[2][2]   (64 B reserved)
//   mig   : [side][parity]  pos4[cap_mig] | vel4[cap_mig] | id[cap_mig]
//   halo  : [side][parity]  pos4[cap_halo] | id[cap_halo]
struct SlabMail {
    int cap_mig, cap_halo;
    __host__ __device__ size_t mig_bytes() const { return ((size_t)cap_mig * 36 + 255) & ~(size_t)255; }
    __host__ __device__ size_t halo_bytes() const { return ((size_t)cap_halo * 20 + 255) & ~(size_t)255; }
    __host__ __device__ size_t bytes() const { return 256 + 4 * mig_bytes() + 4 * halo_bytes(); }
    __host__ __device__ int* flag_mig(char* b, int side) const { return (int*)b + side; }
    __host__ __device__ int* flag_halo(char* b, int side) const { return (int*)b + 2 + side; }
    __host__ __device__ int* cnt_mig(char* b, int side, int par) const { return (int*)b + 4 + side * 2 + par; }
    __host__ __device__ int* cnt_halo(char* b, int side, int par) const { return (int*)b + 8 + side * 2 + par; }
    __host__ __device__ char* mig(char* b, int side, int par) const { return b + 256 + (size_t)(side * 2 + par) * mig_bytes(); }
    __host__ __device__ char* halo(char* b, int side, int par) const {
        return b + 256 + 4 * mig_bytes() + (size_t)(side * 2 + par) * halo_bytes();
    }
    __host__ __device__ float4* mig_pos(char* m) const { return (float4*)m; }
    __host__ __device__ float4* mig_vel(char* m) const { return (float4*)(m + (size_t)cap_mig * 16); }
    __host__ __device__ int* mig_id(char* m) const { return (int*)(m + (size_t)cap_mig * 32); }
    __host__ __device__ float4* halo_pos(char* m) const { return (float4*)m; }
    __host__ __device__ int* halo_id(char* m) const { return (int*)(m + (size_t)cap_halo * 16); }
};

// Everything a sending kernel needs: the two neighbours' mailboxes (mapped peer memory, or this rank's own
// mailbox when world == 1) and the local status words.
struct SlabPeers {
    char* left;   // mailbox of the left neighbour: I write its side 1 ("from my right neighbour")
    char* right;  // mailbox of the right neighbour: I write its side 0
    SlabMail mail;
    int* status;  // d_slab
};

__device__ __forceinline__ void slab_store_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int slab_load_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long slab_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// 0 = stay, 1 = to the left neighbour, 2 = to the right neighbour
__device__ __forceinline__ int slab_direction(float x, const SlabGeom& g, int* status) {
    if (x >= g.x_lo && x < g.x_hi) return 0;
    float d = x - g.x_lo;
    if (d < 0.f) d += g.W;
    if (d < g.w_own + g.w_right) return 2; // (also covers G == 2: both neighbours are the same rank)
    if (d >= g.W - g.w_left) return 1;
    atomicOr(&status[SLAB_ERR], SLAB_ERR_FAR);
    return 0;
}

// One leaver -> the neighbour's mailbox (parity = seq & 1).  Slot order is the order of the atomics; the
// receiver sorts its arrivals by id, so the result does not depend on it.
__device__ __forceinline__ void slab_emit_migrant(const SlabPeers& P, int dir, int par, float4 pos, float4 vel, int id) {
    const int slot = atomicAdd(&P.status[dir == 1 ? SLAB_SEND_L : SLAB_SEND_R], 1);
    if (slot >= P.mail.cap_mig) return; // overflow: reported by the closing block
    char* m = dir == 1 ? P.mail.mig(P.left, 1, par) : P.mail.mig(P.right, 0, par);
    P.mail.mig_pos(m)[slot] = pos;
    P.mail.mig_vel(m)[slot] = vel;
    P.mail.mig_id(m)[slot] = id;
}

// Called by every block of a migrant-emitting kernel after its last store: the last block to arrive
// publishes the two counts and releases the sequence flags.
__device__ __forceinline__ void slab_close_migrants(const SlabPeers& P, int seq) {
    // barrier, then ONE system-scope fence by thread 0: it orders every peer store of the block (observed
    // through the barrier) before the ticket — the pattern of a cooperative grid barrier.  A fence per thread
    // cost 40 us on a 1.5 M-slot grid.
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const int t = atomicAdd(&P.status[SLAB_TICKET_MIG], 1);
        if (t == (int)gridDim.x - 1) {
            __threadfence();
            int nl = atomicExch(&P.status[SLAB_SEND_L], 0), nr = atomicExch(&P.status[SLAB_SEND_R], 0);
            if (nl > P.mail.cap_mig || nr > P.mail.cap_mig) atomicOr(&P.status[SLAB_ERR], SLAB_ERR_MIG);
            nl = min(nl, P.mail.cap_mig), nr = min(nr, P.mail.cap_mig);
            P.status[SLAB_TICKET_MIG] = 0;
            const int par = seq & 1;
            *P.mail.cnt_mig(P.left, 1, par) = nl;
            *P.mail.cnt_mig(P.right, 0, par) = nr;
            __threadfence_system();
            slab_store_release_sys(P.mail.flag_mig(P.left, 1), seq);
            slab_store_release_sys(P.mail.flag_mig(P.right, 0), seq);
        }
    }
}

// Stand-alone classify + emit (the first build after an upload / spawn / universe move; every later step
// emits from the integrate kernel's epilogue, kernels_state.cuh).
__global__ void slab_emit_migrants_kernel(const float4* __restrict__ pos4, const float4* __restrict__ vel4,
                                          const int* __restrict__ id, int n_upper, SlabGeom g, SlabPeers P, int seq) {
    const int n = min(P.status[SLAB_NCUR], n_upper);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const float4 p = pos4[k];
        const int dir = slab_direction(p.x, g, P.status);
        if (dir) slab_emit_migrant(P, dir, seq & 1, p, vel4[k], id[k]);
    }
    slab_close_migrants(P, seq);
}

// Fused integrate + migrant emission (slab mode): the integrate kernel of kernels_state.cuh, whose epilogue
// classifies the particle by its NEW x and stores a leaver straight into the neighbour's mailbox — the
// compute step and its exchange are one kernel; the closing block releases the neighbours' flags.  The
// leaver also stays in the local arrays until the next cell-list build sorts it out (class GONE).
__global__ void integrate_slab_kernel(float4* __restrict__ pos4, float4* __restrict__ vel4, float4* __restrict__ frc4,
                                      const int* __restrict__ id, int n_upper, StepConst c, SlabGeom g, SlabPeers P,
                                      int seq) {
    const int n = min(P.status[SLAB_NCUR], n_upper);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) { // grid sized for the expected count
        float4 p = pos4[k], v = vel4[k], f = frc4[k];
        cf_integrate_particle(p, v, f, c);
        pos4[k] = p;
        vel4[k] = v;
        frc4[k] = f;
        const int dir = slab_direction(p.x, g, P.status);
        if (dir) slab_emit_migrant(P, dir, seq & 1, p, v, id[k]);
    }
    slab_close_migrants(P, seq);
}

// One warp: lanes 0 and 1 wait for the two flags of my mailbox to reach `seq`.
__global__ void slab_wait_kernel(const int* flag0, const int* flag1, int seq, int* status, unsigned long long timeout_ns) {
    if (threadIdx.x < 2) {
        const int* f = threadIdx.x == 0 ? flag0 : flag1;
        const unsigned long long t0 = slab_globaltimer();
        while (slab_load_acquire_sys(f) - seq < 0) {
            if (slab_globaltimer() - t0 > timeout_ns) {
                atomicOr(&status[SLAB_ERR], SLAB_ERR_TIMEOUT);
                break;
            }
            __nanosleep(200);
        }
    }
}

// Sort key of owned slot k (fused first sort pass): class * (ncell*64) + cf_sort_key.
struct SlabKeyFn {
    const float4* pos4; // owned slots (base applied)
    StepConst c;
    SlabGeom g;
    int* status;
    __device__ __forceinline__ uint32_t operator()(int k) const {
        const float4 p = pos4[k];
        const uint32_t gone = (p.x >= g.x_lo && p.x < g.x_hi) ? SLAB_STAY : SLAB_GONE;
        return gone * g.class_stride + cf_sort_key(p, c); // leavers: clamped cell, irrelevant
    }
};

// Arrivals -> slots [n, n + n_a) behind the owned particles, in id order (deterministic whatever order the
// senders' atomics produced).  One block; up to SLAB_ARR_SORT arrivals are sorted in shared memory (a step
// moves a few hundred particles across a face), more are appended unsorted (still correct; ties inside one
// sub-cell then depend on the arrival order).
#define SLAB_ARR_SORT 4096
__global__ void __launch_bounds__(1024)
slab_unpack_arrivals_kernel(char* __restrict__ mybox, SlabMail mail, int par, float4* __restrict__ pos4,
                            float4* __restrict__ vel4, int* __restrict__ id, int cap_own, int* __restrict__ status) {
    __shared__ unsigned long long skey[SLAB_ARR_SORT];
    const int n = min(status[SLAB_NCUR], cap_own);
    int n_al = min(*mail.cnt_mig(mybox, 0, par), mail.cap_mig), n_ar = min(*mail.cnt_mig(mybox, 1, par), mail.cap_mig);
    n_al = max(n_al, 0), n_ar = max(n_ar, 0);
    int n_a = n_al + n_ar;
    if (n + n_a > cap_own) { // drop what does not fit, report
        if (threadIdx.x == 0) atomicOr(&status[SLAB_ERR], SLAB_ERR_OWN);
        n_a = cap_own - n;
        n_al = min(n_al, n_a);
        n_ar = n_a - n_al;
    }
    char* ml = mail.mig(mybox, 0, par);
    char* mr = mail.mig(mybox, 1, par);
    const bool sorted = n_a <= SLAB_ARR_SORT;
    if (sorted) {
        int m = 1;
        while (m < n_a) m <<= 1;
        for (int k = threadIdx.x; k < m; k += blockDim.x) {
            unsigned long long v = ~0ull;
            if (k < n_a) {
                const int aid = k < n_al ? __ldcg(&mail.mig_id(ml)[k]) : __ldcg(&mail.mig_id(mr)[k - n_al]);
                v = ((unsigned long long)(unsigned)aid << 32) | (unsigned)k;
            }
            skey[k] = v;
        }
        __syncthreads();
        for (int size = 2; size <= m; size <<= 1)
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int k = threadIdx.x; k < m; k += blockDim.x) {
                    const int partner = k ^ stride;
                    if (partner > k) {
                        const bool up = (k & size) == 0;
                        const unsigned long long a = skey[k], b = skey[partner];
                        if ((a > b) == up) skey[k] = b, skey[partner] = a;
                    }
                }
                __syncthreads();
            }
    }
    for (int k = threadIdx.x; k < n_a; k += blockDim.x) {
        const int src = sorted ? (int)(skey[k] & 0xffffffffu) : k;
        char* m = src < n_al ? ml : mr;
        const int e = src < n_al ? src : src - n_al;
        pos4[n + k] = __ldcg(&mail.mig_pos(m)[e]);
        vel4[n + k] = __ldcg(&mail.mig_vel(m)[e]);
        id[n + k] = __ldcg(&mail.mig_id(m)[e]);
    }
    if (threadIdx.x == 0) {
        status[SLAB_NTMP] = n + n_a;
        status[SLAB_ARR_L] = n_al;
        status[SLAB_ARR_R] = n_ar;
    }
}

// Copy the first owned x layer (-> left neighbour's side 1) and the last one (-> right neighbour's side 0)
// into the neighbours' mailboxes; the last block publishes the counts and releases the flags.
// Layer ranges come from cell_start on the device.
__global__ void slab_pack_halo_kernel(const float4* __restrict__ pos4, const int* __restrict__ id,
                                      const int* __restrict__ cell_start, int layer_cells, int nxl, SlabPeers P, int seq) {
    const int l0 = cell_start[layer_cells], l1 = cell_start[2 * layer_cells];
    const int r0 = cell_start[nxl * layer_cells], r1 = cell_start[(nxl + 1) * layer_cells];
    const int cap = P.mail.cap_halo, par = seq & 1;
    const int nl = l1 - l0, nr = r1 - r0;
    char* const ml = P.mail.halo(P.left, 1, par);
    char* const mr = P.mail.halo(P.right, 0, par);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < min(max(nl, nr), cap); k += gridDim.x * blockDim.x) {
        if (k < nl) {
            P.mail.halo_pos(ml)[k] = pos4[l0 + k];
            P.mail.halo_id(ml)[k] = id[l0 + k];
        }
        if (k < nr) {
            P.mail.halo_pos(mr)[k] = pos4[r0 + k];
            P.mail.halo_id(mr)[k] = id[r0 + k];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system(); // (see slab_close_migrants)
        const int t = atomicAdd(&P.status[SLAB_TICKET_HALO], 1);
        if (t == (int)gridDim.x - 1) {
            __threadfence();
            P.status[SLAB_TICKET_HALO] = 0;
            if (nl > cap || nr > cap) atomicOr(&P.status[SLAB_ERR], SLAB_ERR_HALO);
            *P.mail.cnt_halo(P.left, 1, par) = min(nl, cap);
            *P.mail.cnt_halo(P.right, 0, par) = min(nr, cap);
            __threadfence_system();
            slab_store_release_sys(P.mail.flag_halo(P.left, 1), seq);
            slab_store_release_sys(P.mail.flag_halo(P.right, 0), seq);
        }
    }
}

// Ghosts from the left neighbour end right before the owned slots, ghosts from the right neighbour start
// right after them; their keys (forced ghost x layer) feed the bounds search.
__global__ void slab_unpack_ghosts_kernel(char* __restrict__ mybox, SlabMail mail, int par, float4* __restrict__ pos4,
                                          int* __restrict__ id, int own_first, const int* __restrict__ status,
                                          uint32_t* __restrict__ gkeys_left, uint32_t* __restrict__ gkeys_right,
                                          StepConst c) {
    const int n_own = status[SLAB_NCUR];
    const int nl = min(max(*mail.cnt_halo(mybox, 0, par), 0), mail.cap_halo);
    const int nr = min(max(*mail.cnt_halo(mybox, 1, par), 0), mail.cap_halo);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < max(nl, nr); k += gridDim.x * blockDim.x) {
    if (k < nl) {
        char* m = mail.halo(mybox, 0, par);
        const float4 p = __ldcg(&mail.halo_pos(m)[k]);
        const int slot = own_first - nl + k;
        pos4[slot] = p;
        id[slot] = __ldcg(&mail.halo_id(m)[k]);
        const int cy = cf_cell_coord(p.y, c.inv[1], c.dims[1]), cz = cf_cell_coord(p.z, c.inv[2], c.dims[2]);
        gkeys_left[k] = (uint32_t)((0 * c.dims[1] + cy) * c.dims[2] + cz) * CF_KEY_SUB; // cell part only
    }
    if (k < nr) {
        char* m = mail.halo(mybox, 1, par);
        const float4 p = __ldcg(&mail.halo_pos(m)[k]);
        const int slot = own_first + n_own + k;
        pos4[slot] = p;
        id[slot] = __ldcg(&mail.halo_id(m)[k]);
        const int cy = cf_cell_coord(p.y, c.inv[1], c.dims[1]), cz = cf_cell_coord(p.z, c.inv[2], c.dims[2]);
        gkeys_right[k] = (uint32_t)(((c.dims[0] - 1) * c.dims[1] + cy) * c.dims[2] + cz) * CF_KEY_SUB;
    }
    }
}

// cell_start of the two ghost layers (lower bounds over the ghost key arrays).
__global__ void slab_ghost_bounds_kernel(const uint32_t* __restrict__ gkeys_left, const uint32_t* __restrict__ gkeys_right,
                                         char* __restrict__ mybox, SlabMail mail, int par, int* __restrict__ cell_start,
                                         int layer_cells, int ncell, int own_first, int* __restrict__ status) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_own = status[SLAB_NCUR];
    const int nl = min(max(*mail.cnt_halo(mybox, 0, par), 0), mail.cap_halo);
    const int nr = min(max(*mail.cnt_halo(mybox, 1, par), 0), mail.cap_halo);
    if (k == 0) {
        status[SLAB_GHOST_L] = nl;
        status[SLAB_GHOST_R] = nr;
        status[SLAB_FIRST] = own_first - nl;
        status[SLAB_NSLOTS] = nl + n_own + nr;
    }
    if (k < layer_cells) { // cells of layer 0: c = k
        const uint32_t want = (uint32_t)k * CF_KEY_SUB;
        int lo = 0, hi = nl;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (gkeys_left[mid] < want) lo = mid + 1; else hi = mid;
        }
        cell_start[k] = own_first - nl + lo;
    } else if (k < 2 * layer_cells + 1) { // cells of the last layer plus the end sentinel
        const int c = ncell - layer_cells + (k - layer_cells);
        const uint32_t want = (uint32_t)c * CF_KEY_SUB;
        int lo = 0, hi = nr;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
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

// counts[0] = number of keys < class_stride in a sorted key array (= particles this rank keeps).
__global__ void slab_count_mine_kernel(const uint32_t* __restrict__ skeys, int n, int* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (skeys[mid] < 1u) lo = mid + 1; else hi = mid;
    }
    *out = lo;
}
