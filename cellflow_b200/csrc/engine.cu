// engine.cu — the cf_sim handle and the C ABI of include/cellflow_b200.h.
//
// Replaces the reference's host class ParticleSimulation (cuda-native/src/ParticleSimulation.cu:
// 426-687).  One handle = one device, one non-default stream, device-resident float4 SoA state
// kept in cell-sorted order between steps.  A step is
//     cell key -> stable radix sort -> reorder -> cell bounds -> pair force -> fused integrate
// with no host synchronisation inside (the reference synchronises after every launch, .cu:553).
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h> // header-only NVTX 3: ranges around the phases of a step

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "../../include/cellflow_b200.h"
#include "cf_device.cuh"
#include "kernels_force.cuh"
#include "kernels_graph.cuh"
#include "kernels_slab.cuh"
#include "kernels_sort.cuh"
#include "kernels_state.cuh"
#include "kernels_tile.cuh"
#include "kernels_tile4.cuh"

static_assert(sizeof(AosParticle) == 44 && sizeof(cf_particle) == 44, "reference Particle is 44 B");

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(CF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),  \
                        __FILE__, __LINE__);                                                  \
    } while (0)
#define ARG(cond)                                                                             \
    do {                                                                                      \
        if (!(cond)) return fail(CF_ERR_ARG, "bad argument: %s (%s:%d)", #cond, __FILE__, __LINE__); \
    } while (0)

extern "C" const char* cf_last_error(void) { return g_err; }
extern "C" const char* cf_version(void) { return "cellflow_b200 0.1 (sm_100a)"; }

// ---------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------
#define CF_STEP_EVENTS 8
struct StepEvents {
    // begin, after cell list, after force, after integrate; slab mode: migrants emitted, arrivals appended,
    // owned cells sorted, ghosts in place (e[4]..e[5] and e[6]..e[7] bracket the two mailbox exchanges)
    cudaEvent_t e[CF_STEP_EVENTS];
    bool has_exchange = false; // e[4..7] were recorded for this step
};

struct cf_sim {
    int device = 0;
    cudaStream_t stream = nullptr;
    int n = 0;   // owned particles
    int cap = 0; // slot capacity
    int T = 6;
    cf_params params;
    float raw[CF_TT_MAX];
    float radio[CF_T_MAX];
    float force[CF_TT_MAX];
    float half_host[CF_T_MAX];
    bool force_overridden = false;
    struct GlibcRand* rng = nullptr; // the handle's own copy of libc's never-seeded rand() stream (.cu:513-539)

    float4* pos[2] = {nullptr, nullptr};
    float4* vel[2] = {nullptr, nullptr};
    int* id[2] = {nullptr, nullptr};
    float4* frc = nullptr;
    int cur = 0;
    uint32_t* keys[2] = {nullptr, nullptr};
    uint32_t* vals[2] = {nullptr, nullptr};
    uint32_t* hist = nullptr;
    size_t hist_cap = 0;
    int* cell_start = nullptr;
    size_t cell_cap = 0;
    DeviceTables* d_tables = nullptr;
    StepConst sc;
    int ncell = 0;
    bool sorted_valid = false;
    unsigned long long state_gen = 1;   // bumped by every reorder and every change of positions
    unsigned long long graph_gen = 0;   // state_gen at the end of the last graph build: its edge slots are valid while equal

    float* d_vertices = nullptr;      // persistent vertex stream of the last graph (12 floats per edge)
    size_t vertex_cap = 0;            // ... in edges
    float* d_colors = nullptr;
    bool graph_count_pending = false; // the last build's edge count / occupancy have not been read back yet
    int graph_plan_count = 0;
    int2* edges = nullptr;
    int2* edge_slots = nullptr;
    int edge_cap = 0;
    int* d_edge_count = nullptr;
    int n_edges = 0;

    AosParticle* d_aos = nullptr;
    int* d_counts = nullptr;
    unsigned long long* d_accum = nullptr;

    // proximity graph: its own (type, cell) list
    uint32_t* gk[2] = {nullptr, nullptr};
    uint32_t* gv[2] = {nullptr, nullptr};
    float4* gpos = nullptr;
    size_t graph_cap = 0;
    int* gstart = nullptr;
    size_t gstart_cap = 0;
    unsigned long long* d_graph_occ = nullptr; // sum over (type, cell) keys of occupancy^2, last build
    unsigned long long h_graph_occ = 0;
    unsigned long long* h_graph_occ_pin = nullptr; // pinned mirror every build refreshes without synchronising
    double graph_mean_occ = 0.0;               // mean same-key companions per particle, last build
    int opt_graph_kernel = 0;                  // 0 auto, 1 thread per particle, 2 warp per particle
    int last_graph_kernel = 0;

    // tile kernel
    int2* d_tiles = nullptr;
    size_t tiles_cap = 0;
    int* d_tile_ctrl = nullptr;
    // particle-weighted cell occupancy (sum n_c^2) of the last completed step: device scalar and a
    // pinned host copy that every step refreshes asynchronously (read without synchronising)
    unsigned long long* d_cell_occ = nullptr;
    unsigned long long* h_cell_occ = nullptr;
    int planned_force_kernel = 1; // choice for the step being enqueued (part of the graph signature)
    // the occupancy the choice uses: taken over from h_cell_occ only where the API synchronises anyway
    // (cf_sync, cf_build_graph, cf_step_host), so the choice is a function of the call sequence, not of timing
    double policy_occ = 0.0;
    int occ_stride = 1; // the statistic samples every occ_stride-th cell
    float* d_half = nullptr;      // per-type conservative half radius (non-uniform radii)
    // generation-4 tile kernel, per-type radii: j copy sorted by (xy row, type, z cell, Morton)
    uint32_t* hk[2] = {nullptr, nullptr};
    uint32_t* hv[2] = {nullptr, nullptr};
    int* h_cell_of = nullptr;
    float4* h_pos = nullptr;
    float* h_xyz[3] = {nullptr, nullptr, nullptr}; // the same copy as SoA planes (bulk-copy staging)
    float4* d_qbox = nullptr;                      // STAGE 2: bounding boxes of the aligned quads of the j array
    size_t qbox_cap = 0;
    uint32_t* h_comp = nullptr;
    size_t homog_cap = 0;
    int* h_start = nullptr;
    size_t h_start_cap = 0;
    bool half_bound_ok = true;
    int sm_count = 148;

    // slab decomposition (multi-GPU); single GPU: base = 0, slab = false
    int base = 0;            // first owned slot
    bool slab = false;
    int rank = 0, world = 1;
    int cap_own = 0, cap_halo = 0, cap_mig = 0;
    int nxl = 0;             // owned x layers
    SlabGeom geom;
    std::vector<float> bounds;     // slab bounds set by cf_slab_set_bounds (empty: uniform split)
    SlabMail mail;                 // mailbox layout (identical on every rank)
    char* mailbox = nullptr;       // this rank's mailbox (device memory, exported through CUDA IPC)
    char* peer_box[2] = {nullptr, nullptr}; // the left / right neighbour's mailbox, mapped into this process
    bool connected = false;
    uint32_t* gkeys[2] = {nullptr, nullptr};
    int* d_slab = nullptr;         // device status words (kernels_slab.cuh: SLAB_NCUR ...): counts never leave the device
    int* h_slab = nullptr;         // pinned mirror, refreshed where the API synchronises
    int seq_mig = 0, seq_halo = 0; // sequence numbers of the two message kinds (same on every rank)
    bool mig_sent = false;         // the last integrate already emitted this step's migrants
    long long n_total = 0;         // global particle count (slab mode)
    double wait_timeout_ms = 20000.0;
    double opt_min_layer_width = 0.0; // slab mode: lower bound of the x layer width (same on every rank)
    double cell_edge = 1.0;        // cell edge and largest interaction radius of the current grid
    float rmax = 0.f;
    double ms_exchange = 0, ms_exchange_mig = 0, ms_exchange_halo = 0, ms_exchange_max = 0, ms_step_max = 0;

    // CUDA graphs of the (static) single-GPU step sequence: small problems are launch-bound
    struct StepGraph {
        cudaGraphExec_t exec = nullptr;
        std::vector<char> sig;      // everything the captured launches depend on
        std::vector<char> seen;     // signature met once (buffers sized): capture next time
        long long nodes = 0;
        bool swap_keys = false;
    };
    StepGraph step_graphs[4];
    StepGraph graph_graphs[4]; // cf_build_graph's device sequence (single GPU)
    int opt_graphs = 1;

    // options
    int opt_force_kernel = 0; // 0 auto, 1 per-particle, 2 tile (generation 3), 3 tile (generation 4)
    int opt_timing = 0;
    int opt_t4_stage = 4;     // staging of the j chunks in the tile kernel (kernels_tile4.cuh): 4 (default) box prefilter on the
                              // registers, live quads stored compacted; 0 every quad stored (round 1 .. mid round 2); 1
                              // cp.async.bulk + mbarrier from SoA planes; 2 precomputed quad boxes; 3 SoA planes through registers
                              // (1 and 3: per-type radii only).  All bit-identical.
    int opt_count_blocks = 0; // instrumented tile kernel: counts exact-tested / evaluated blocks (cf_stats)
    unsigned long long* d_block_counts = nullptr;
    bool block_counts_valid = false;
    int opt_t4_ctas = 0;      // experiment: resident CTAs per SM of the generation-4 tile kernel (0 = default 5)
    double opt_max_cells_per_particle = 16.0; // fine grids pay off for clustered states (cells are cheap)

    // stats
    std::vector<StepEvents> ev_pool;
    size_t ev_used = 0;
    std::vector<cudaEvent_t> gev_pool; // event pairs of the graph builds since the last fold (builds may be asynchronous)
    size_t gev_used = 0;
    double ms_sort = 0, ms_force = 0, ms_integrate = 0, ms_total = 0, ms_graph = 0, ms_graph_total = 0;
    long long graph_builds = 0;
    long long stat_steps = 0;
    long long launches = 0;
    long long tested_pairs = 0;
    int last_force_kernel = 0;
};

// NVTX range for the lifetime of the object (a few ns when no profiler is attached).
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

#define LAUNCH(sim, kernel, grid, block, smem, ...)                         \
    do {                                                                    \
        kernel<<<(grid), (block), (smem), (sim)->stream>>>(__VA_ARGS__);    \
        (sim)->launches++;                                                  \
    } while (0)

static inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// Owned particles occupy slots [base, base + n) of the state arrays (base = 0 on a single GPU;
// in slab mode the slots before/after hold the ghost layers).
static inline float4* opos(cf_sim* s) { return s->pos[s->cur] + s->base; }
static inline float4* ovel(cf_sim* s) { return s->vel[s->cur] + s->base; }
static inline int* oid(cf_sim* s) { return s->id[s->cur] + s->base; }
static inline float4* ofrc(cf_sim* s) { return s->frc + s->base; }

static int set_device(const cf_sim* sim) {
    CU(cudaSetDevice(sim->device));
    return 0;
}

static void free_particle_buffers(cf_sim* s) {
    for (int b = 0; b < 2; b++) {
        cudaFree(s->pos[b]);
        cudaFree(s->vel[b]);
        cudaFree(s->id[b]);
        cudaFree(s->keys[b]);
        cudaFree(s->vals[b]);
        s->pos[b] = s->vel[b] = nullptr;
        s->id[b] = nullptr;
        s->keys[b] = s->vals[b] = nullptr;
    }
    cudaFree(s->frc);
    cudaFree(s->d_aos);
    cudaFree(s->d_counts);
    cudaFree(s->edges);
    cudaFree(s->edge_slots);
    s->frc = nullptr;
    s->d_aos = nullptr;
    s->d_counts = nullptr;
    s->edges = s->edge_slots = nullptr;
    s->edge_cap = 0;
    s->cap = 0;
}

static int alloc_particle_buffers(cf_sim* s, int cap) {
    free_particle_buffers(s);
    size_t c = (size_t)std::max(cap, 1);
    for (int b = 0; b < 2; b++) {
        CU(cudaMalloc(&s->pos[b], c * sizeof(float4)));
        CU(cudaMalloc(&s->vel[b], c * sizeof(float4)));
        CU(cudaMalloc(&s->id[b], c * sizeof(int)));
        CU(cudaMalloc(&s->keys[b], c * sizeof(uint32_t)));
        CU(cudaMalloc(&s->vals[b], c * sizeof(uint32_t)));
    }
    CU(cudaMalloc(&s->frc, c * sizeof(float4)));
    CU(cudaMalloc(&s->d_aos, c * sizeof(AosParticle)));
    CU(cudaMalloc(&s->d_counts, std::max<size_t>(c, 16) * sizeof(int))); // also the render feed's type histogram
    for (int b = 0; b < 2; b++) {
        CU(cudaMemsetAsync(s->pos[b], 0, c * sizeof(float4), s->stream));
        CU(cudaMemsetAsync(s->vel[b], 0, c * sizeof(float4), s->stream));
        CU(cudaMemsetAsync(s->id[b], 0, c * sizeof(int), s->stream));
    }
    CU(cudaMemsetAsync(s->frc, 0, c * sizeof(float4), s->stream));
    s->cap = (int)c;
    s->cur = 0;
    s->sorted_valid = false, s->state_gen++;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// tables: force-table squash, per-pair radius, exact cut-off thresholds
// ---------------------------------------------------------------------------------------------

// updateForceTable, ParticleSimulation.cu:521-527: float tanh, separate product and sum.
static void squash_force_table(cf_sim* s, float range, float bias, float offset) {
    for (int i = 0; i < s->T * s->T; i++) {
        volatile float t = tanhf(s->raw[i] * offset);
        volatile float v = t * range;
        v = v + bias;
        s->force[i] = fmaxf(-1.0f, fminf(1.0f, v));
    }
    s->force_overridden = false;
}

// libc rand() tables of a fresh reference object (.cu:513-519, 533-539).  The reference never
// calls srand, so its first construction sees the default sequence; this library must not
// disturb the host's generator, so it replays glibc's TYPE_3 additive feedback generator
// (seed 1) privately.
struct GlibcRand {
    int32_t r[344 + 4096];
    int k = 0;
    GlibcRand() {
        r[0] = 1;
        for (int i = 1; i < 31; i++) {
            long long v = (16807LL * r[i - 1]) % 2147483647LL;
            if (v < 0) v += 2147483647LL;
            r[i] = (int32_t)v;
        }
        for (int i = 31; i < 34; i++) r[i] = r[i - 31];
        for (int i = 34; i < 344; i++) r[i] = (int32_t)((uint32_t)r[i - 31] + (uint32_t)r[i - 3]);
        k = 344;
    }
    int next() {
        if (k >= 344 + 4096) { // keep the last 34 values and continue (a handle may regenerate tables for ever)
            for (int i = 0; i < 34; i++) r[310 + i] = r[k - 34 + i];
            k = 344;
        }
        r[k] = (int32_t)((uint32_t)r[k - 31] + (uint32_t)r[k - 3]);
        int out = (int)(((uint32_t)r[k]) >> 1);
        k++;
        return out;
    }
};

static void default_tables(cf_sim* s, GlibcRand& g) {
    for (int i = 0; i < s->T * s->T; i++) s->raw[i] = (float)g.next() / 2147483647 * 2.0f - 1.0f;
    squash_force_table(s, 0.28f, -0.20f, 1.0f); // .cu:518
    for (int i = 0; i < s->T; i++) s->radio[i] = (float)g.next() / 2147483647 * 2.0f - 1.0f;
}

extern "C" int cf_reference_default_tables(int T, float* raw, float* radio, float* effective) {
    ARG(T >= 1 && T <= CF_MAX_PARTICLE_TYPES && raw && radio && effective);
    cf_sim tmp;
    tmp.T = T;
    GlibcRand g;
    default_tables(&tmp, g);
    memcpy(raw, tmp.raw, sizeof(float) * T * T);
    memcpy(radio, tmp.radio, sizeof(float) * T);
    memcpy(effective, tmp.force, sizeof(float) * T * T);
    return CF_OK;
}

// Reff of a type pair as the reference kernel evaluates it (.cu:108-110, SASS rounding points):
//   a_x = fma(radio[x], ratio, 1);  Reff = fma(a_p, radius, a_o * radius) * 0.5
static float reff_pair(const cf_sim* s, int tp, int to) {
    float ap = fmaf(s->radio[tp], s->params.ratioWithLFO, 1.0f);
    float ao = fmaf(s->radio[to], s->params.ratioWithLFO, 1.0f);
    volatile float prod = ao * s->params.radius;
    return fmaf(ap, s->params.radius, prod) * 0.5f;
}

// Largest-accepting threshold: the reference accepts iff sqrtf(d2 + 1e-4f) < Reff.  sqrtf and
// the addition are monotone, so there is a float cut2 with  accept <=> d2 < cut2 ; it is found
// here exactly (host sqrtf is IEEE, like the device's sqrt.rn).
static float exact_cut2(float reff) {
    if (!(reff > 0.0f)) return 0.0f;         // nothing is closer than sqrt(1e-4) ... or NaN
    if (std::isinf(reff)) return INFINITY;
    // c1 = smallest x with sqrtf(x) >= reff
    float c1 = reff * reff;
    while (sqrtf(c1) >= reff && c1 > 0.0f) c1 = nextafterf(c1, 0.0f);
    while (!(sqrtf(c1) >= reff)) c1 = nextafterf(c1, INFINITY);
    if (std::isinf(c1)) return INFINITY;
    // c2 = smallest d >= 0 with (d + 1e-4f) >= c1
    float c2 = c1 - 0.0001f;
    if (!(c2 > 0.0f)) c2 = 0.0f;
    auto ge = [&](float d) { volatile float x = d + 0.0001f; return x >= c1; };
    while (c2 > 0.0f && ge(c2)) c2 = nextafterf(c2, 0.0f);
    while (!ge(c2)) c2 = nextafterf(c2, INFINITY);
    return c2;
}

static float compute_tables(cf_sim* s, DeviceTables& t, bool& uniform) {
    float rmax = 0.f;
    int T = s->T;
    for (int a = 0; a < T; a++)
        for (int b = 0; b < T; b++) {
            float reff = reff_pair(s, a, b);
            t.cut2[a * T + b] = exact_cut2(reff);
            t.inv_reff[a * T + b] = reff > 0.f ? 1.0f / reff : 0.f;
            t.force[a * T + b] = s->force[a * T + b];
            if (reff > rmax) rmax = reff;
        }
    uniform = true;
    for (int i = 1; i < T * T; i++)
        if (t.cut2[i] != t.cut2[0] || t.inv_reff[i] != t.inv_reff[0]) uniform = false;
    // conservative per-type half radii for the tile kernel's prefilter:
    // (h_a + h_b)^2 evaluated in fp32 must not fall below cut2[a][b]
    s->half_bound_ok = true;
    for (int a = 0; a < T; a++) {
        double at = (double)fmaf(s->radio[a], s->params.ratioWithLFO, 1.0f);
        s->half_host[a] = (float)(0.5 * (double)s->params.radius * at * (1.0 + 4e-6));
    }
    for (int a = 0; a < T; a++)
        for (int b = 0; b < T; b++) {
            volatile float h = s->half_host[a] + s->half_host[b];
            volatile float h2 = h * h;
            if (t.cut2[a * T + b] > 0.f && !(h2 >= t.cut2[a * T + b] && h > 0.f)) s->half_bound_ok = false;
        }
    return rmax;
}

// x layers of a slab of width w (same rule on every rank, for every rank's slab)
static int slab_layers(const cf_sim* s, double w) {
    // option "slab_min_layer_width": wider x layers = wider ghost layers, for a proximity-graph distance larger than
    // the interaction radius (the graph may not reach further than one ghost layer)
    int nxl = std::max(1, std::min((int)floor(w / std::max(s->cell_edge, s->opt_min_layer_width)), 1022));
    while (nxl > 1 && w / nxl < (double)s->rmax * (1.0 + 4.0 * nxl * 1.1920929e-7)) nxl--;
    return nxl;
}
static float slab_bound(const cf_sim* s, int r);
static float slab_width(const cf_sim* s, int r);
static void slab_update_geom(cf_sim* s);

// Chooses the cell grid for the current parameters and uploads the tables.
static int prepare_step_const(cf_sim* s) {
    const cf_params& p = s->params;
    ARG(p.canvasWidth > 0.f && p.canvasHeight > 0.f && p.canvasDepth > 0.f);
    ARG(p.numParticleTypes == s->T);
    DeviceTables t;
    memset(&t, 0, sizeof(t));
    bool uniform = true;
    float rmax = compute_tables(s, t, uniform);
    StepConst c;
    memset(&c, 0, sizeof(c));
    float W[3] = {p.canvasWidth, p.canvasHeight, p.canvasDepth};
    // cell edge >= R_max (1e-5 margin covers the rounding of pos * inv), and coarse enough that
    // the grid has at most opt_max_cells_per_particle * n cells
    // Slab mode: the grid must be the same on every rank (ghost layers arrive sorted by the sender's cells, and
    // both ends of a link size their messages from it), so it is derived from rank-invariant numbers only —
    // the global count (or the capacity, equal on all ranks by contract), never this rank's own count.
    double vol = (double)W[0] * W[1] * W[2] / (s->slab ? s->world : 1);
    double n_ref = s->slab ? (s->n_total > 0 ? (double)s->n_total / s->world : (double)s->cap_own) : (double)s->n;
    double max_cells = std::max(64.0, s->opt_max_cells_per_particle * std::max(n_ref, 1.0));
    max_cells = std::min(max_cells, 16777216.0); // class * ncell * 64 must fit the 32-bit sort key
    double edge = std::max((double)rmax * (1.0 + 1e-5), cbrt(vol / max_cells));
    if (!(edge > 0.0)) edge = cbrt(vol / max_cells);
    long long ncell = 1;
    for (int a = 0; a < 3; a++) {
        int d = (int)floor((double)W[a] / edge);
        d = std::max(1, std::min(d, 1024));
        // the cell coordinate is fl(x * inv): its rounding error grows with the cell index (~d * 2^-24 cell
        // widths), so the slack over R_max must grow with the grid too, or a pair just inside the cut-off
        // could land two cells apart
        while (d > 1 && (double)W[a] / d < (double)rmax * (1.0 + 4.0 * d * 1.1920929e-7)) d--;
        c.dims[a] = d;
        c.W[a] = W[a];
        c.halfW[a] = W[a] * 0.5f;
        c.nhalfW[a] = W[a] * -0.5f;
        c.inv[a] = (float)d / W[a];
        ncell *= d;
    }
    c.periodic_x = 1;
    c.x_org = 0.f;
    c.x_off = 0;
    c.x_cells = c.dims[0];
    c.gshift_lo = c.gshift_hi = 0.f;
    c.gx_lo = 0;
    c.gx_hi = c.dims[0] - 1;
    if (s->slab) {
        // owned x layers over the slab, one ghost layer on each side
        slab_update_geom(s);
        double slab_w = (double)s->geom.x_hi - (double)s->geom.x_lo;
        ncell /= c.dims[0];
        s->cell_edge = edge;
        s->rmax = rmax;
        int nxl = slab_layers(s, slab_w);
        s->nxl = nxl;
        c.dims[0] = nxl + 2;
        c.inv[0] = (float)nxl / (s->geom.x_hi - s->geom.x_lo);
        c.periodic_x = 0;
        c.x_org = s->geom.x_lo;
        c.x_off = 1;
        c.x_cells = nxl;
        c.gshift_lo = s->rank == 0 ? -W[0] : 0.f;
        c.gshift_hi = s->rank == s->world - 1 ? W[0] : 0.f;
        c.gx_lo = s->rank == 0 ? 1 : 0;
        c.gx_hi = s->rank == s->world - 1 ? nxl : nxl + 1;
        ncell *= c.dims[0];
        // every slab (not only mine) must be at least one interaction radius wide: a ghost layer is one layer
        for (int r = 0; r < s->world; r++)
            if ((double)rmax * (1.0 + 1e-5) > (double)slab_width(s, r))
                return fail(CF_ERR_ARG, "interaction radius %.1f exceeds the width %.1f of slab %d: use fewer GPUs", rmax,
                            slab_width(s, r), r);
    }
    c.T = s->T;
    c.repulsion = p.repulsion;
    c.attraction = p.attraction;
    c.nk_log2e = -p.k * 1.4426950408889634f;
    c.dt = p.delta_t;
    c.friction = p.friction;
    c.one_minus_balance = 1.0f - p.balance;
    c.force_multiplier = p.forceMultiplier;
    c.max_expected = (float)p.maxExpectedNeighbors;
    c.uniform_radius = uniform ? 1 : 0;
    c.cut2_uniform = t.cut2[0];
    c.inv_reff_uniform = t.inv_reff[0];
    bool grid_changed = memcmp(c.dims, s->sc.dims, sizeof(c.dims)) != 0 ||
                        memcmp(c.W, s->sc.W, sizeof(c.W)) != 0;
    s->sc = c;
    s->ncell = (int)ncell;
    if (grid_changed) s->sorted_valid = false, s->state_gen++;
    if ((size_t)ncell + 1 > s->cell_cap) {
        cudaFree(s->cell_start);
        s->cell_start = nullptr;
        s->cell_cap = (size_t)ncell + 1 + (size_t)ncell / 4;
        CU(cudaMalloc(&s->cell_start, s->cell_cap * sizeof(int)));
    }
    CU(cudaMemcpyAsync(s->d_tables, &t, sizeof(t), cudaMemcpyHostToDevice, s->stream));
    CU(cudaMemcpyAsync(s->d_half, s->half_host, sizeof(float) * CF_T_MAX, cudaMemcpyHostToDevice, s->stream));
    // pageable source: the copy is staged before the call returns, `t` may go out of scope
    return 0;
}

// ---------------------------------------------------------------------------------------------
// cell-list build
// ---------------------------------------------------------------------------------------------
// Stable radix sort of (key, val) pairs (kernels_sort.cuh).  The pairs are in k[0]/v[0] — or, with GEN,
// generated by the first pass from `fn` (keys) and the identity permutation (vals) into k[0]/v[0].  The
// element count is n_upper, or *dn (<= n_upper) when dn is a device pointer.  Returns in *out_src the
// index (0/1) of the buffer pair holding the result.
template <class KeyFn, bool GEN>
static int radix_sort_run(cf_sim* s, KeyFn fn, uint32_t* k[2], uint32_t* v[2], int n_upper, const int* dn,
                          long long key_range, int* out_src) {
    // block shape from the EXPECTED count (slab mode launches over capacities; the real count is a device word)
    long long expect = n_upper;
    if (dn && s->slab) expect = std::min<long long>(n_upper, (s->n_total > 0 ? s->n_total / s->world : s->cap_own) * 5 / 4 + 4096);
    const RsPlan P = rs_make_plan(n_upper, key_range, (int)expect);
    const size_t hist_need = (size_t)RS_MAX_BINS * P.nblocks + RS_MAX_BINS;
    if (hist_need > s->hist_cap) {
        CU(cudaStreamSynchronize(s->stream));
        cudaFree(s->hist);
        s->hist = nullptr;
        s->hist_cap = hist_need * 2;
        CU(cudaMalloc(&s->hist, s->hist_cap * sizeof(uint32_t)));
    }
    uint32_t* totals = s->hist + (size_t)RS_MAX_BINS * P.nblocks;
    int src = 0;
    for (int p = 0; p < P.passes; p++) {
        const int shift = p * P.bits_per_pass;
        const uint32_t mask = (1u << P.bits_per_pass) - 1u;
        if (p == 0)
            LAUNCH(s, (rs_hist_kernel<KeyFn, GEN>), P.nblocks, RS_THREADS, 0, fn, k[0], v[0], n_upper, dn, shift, mask,
                   s->hist, P.nblocks, P.items);
        else
            LAUNCH(s, (rs_hist_kernel<RsKeysFromArray, false>), P.nblocks, RS_THREADS, 0, RsKeysFromArray{k[src]},
                   nullptr, nullptr, n_upper, dn, shift, mask, s->hist, P.nblocks, P.items);
        LAUNCH(s, rs_scan_rows_kernel, (int)mask + 1, RS_THREADS, 0, s->hist, P.nblocks, totals);
#define RS_SCATTER(R_)                                                                                          \
    LAUNCH(s, rs_scatter_kernel<R_>, P.nblocks, RS_THREADS, 0, k[src], v[src], k[src ^ 1], v[src ^ 1], n_upper, \
           dn, shift, mask, s->hist, totals, P.nblocks, P.groups)
        switch (P.R) {
            case 1: RS_SCATTER(1); break;
            case 2: RS_SCATTER(2); break;
            case 4: RS_SCATTER(4); break;
            default: RS_SCATTER(8); break;
        }
#undef RS_SCATTER
        src ^= 1;
    }
    *out_src = src;
    return 0;
}

static int radix_sort_pairs(cf_sim* s, uint32_t* k[2], uint32_t* v[2], int n, long long key_range, int* out_src,
                            const int* dn = nullptr) {
    return radix_sort_run<RsKeysFromArray, false>(s, RsKeysFromArray{k[0]}, k, v, n, dn, key_range, out_src);
}

static int ensure_sorted(cf_sim* s) {
    if (s->sorted_valid) return 0;
    int n = s->n;
    if (n <= 0) return 0;
    int cur = s->cur, nxt = cur ^ 1;
    int src = 0;
    // key generation is fused into the first histogram pass
    if (int rc = radix_sort_run<CellKeyFn, true>(s, CellKeyFn{s->pos[cur], s->sc}, s->keys, s->vals, n, nullptr,
                                                 (long long)s->ncell * CF_KEY_SUB, &src))
        return rc;
    LAUNCH(s, reorder_bounds_kernel, div_up(std::max(n, s->ncell + 1), 256), 256, 0, s->vals[src], s->keys[src],
           s->pos[cur], s->vel[cur], s->id[cur], s->pos[nxt], s->vel[nxt], s->id[nxt], n, nullptr, s->cell_start,
           s->ncell, 0, nullptr);
    if (src != 0) std::swap(s->keys[0], s->keys[1]), std::swap(s->vals[0], s->vals[1]);
    // keys[0] now holds the sorted keys of the current order
    s->cur = nxt, s->state_gen++;
    s->sorted_valid = true;
    CU(cudaGetLastError());
    return 0;
}

#include "slab_host.inl"

static int build_cell_list(cf_sim* s, cudaEvent_t* ev_x = nullptr) {
    return s->slab ? ensure_sorted_slab(s, ev_x) : ensure_sorted(s);
}

// ---------------------------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------------------------
extern "C" void cf_default_params(cf_params* p) {
    if (!p) return;
    p->radius = 42.07f;
    p->delta_t = 0.18f;
    p->friction = 0.51f;
    p->repulsion = 64.83f;
    p->attraction = 3.06f;
    p->k = 29.45f;
    p->balance = 0.79f;
    p->canvasWidth = 8000.0f;
    p->canvasHeight = 8000.0f;
    p->canvasDepth = 8000.0f;
    p->spawnRegionSize = 2000.0f;
    p->numParticleTypes = 6;
    p->ratioWithLFO = 0.0f;
    p->forceMultiplier = 2.33f;
    p->maxExpectedNeighbors = 400;
    p->forceRange = 0.28f;
    p->forceBias = -0.20f;
    p->ratio = 0.0f;
    p->lfoA = 0.0f;
    p->lfoS = 0.1f;
    p->forceOffset = 1.0f;
}

extern "C" int cf_create(int particle_count, int num_types, int device, cf_sim** out) {
    ARG(out != nullptr);
    *out = nullptr;
    ARG(particle_count >= 0);
    ARG(num_types >= 1 && num_types <= CF_MAX_PARTICLE_TYPES);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(CF_ERR_CUDA, "no CUDA device (%s): cellflow_b200 has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    ARG(device >= 0 && device < ndev);
    CU(cudaSetDevice(device));
    cf_sim* s = new cf_sim();
    s->device = device;
    s->T = num_types;
    s->n = particle_count;
    cf_default_params(&s->params);
    s->params.numParticleTypes = num_types;
    memset(&s->sc, 0, sizeof(s->sc));
    memset(s->raw, 0, sizeof(s->raw));
    memset(s->radio, 0, sizeof(s->radio));
    memset(s->force, 0, sizeof(s->force));
    cudaError_t ce = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (ce != cudaSuccess) {
        delete s;
        return fail(CF_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(ce));
    }
    int rc = alloc_particle_buffers(s, particle_count);
    if (rc == 0 && cudaMalloc(&s->d_tables, sizeof(DeviceTables)) != cudaSuccess) rc = fail(CF_ERR_CUDA, "cudaMalloc tables");
    if (rc == 0 && cudaMalloc(&s->d_edge_count, sizeof(int)) != cudaSuccess) rc = fail(CF_ERR_CUDA, "cudaMalloc");
    if (rc == 0 && cudaMalloc(&s->d_graph_occ, sizeof(unsigned long long)) != cudaSuccess) rc = fail(CF_ERR_CUDA, "cudaMalloc");
    if (rc == 0 && cudaMalloc(&s->d_tile_ctrl, 2 * sizeof(int)) != cudaSuccess) rc = fail(CF_ERR_CUDA, "cudaMalloc");
    if (rc == 0 && cudaMalloc(&s->d_cell_occ, 2 * sizeof(unsigned long long)) != cudaSuccess) rc = fail(CF_ERR_CUDA, "cudaMalloc");
    if (rc == 0 && cudaMemset(s->d_cell_occ, 0, 2 * sizeof(unsigned long long)) != cudaSuccess) rc = fail(CF_ERR_CUDA, "cudaMemset");
    if (rc == 0 && cudaMallocHost(&s->h_cell_occ, sizeof(unsigned long long)) != cudaSuccess) rc = fail(CF_ERR_CUDA, "cudaMallocHost");
    if (rc == 0) *s->h_cell_occ = 0;
    if (rc == 0 && cudaMallocHost(&s->h_graph_occ_pin, sizeof(unsigned long long)) != cudaSuccess) rc = fail(CF_ERR_CUDA, "cudaMallocHost");
    if (rc == 0) *s->h_graph_occ_pin = 0;
    if (rc == 0 && cudaMalloc(&s->d_half, CF_T_MAX * sizeof(float)) != cudaSuccess) rc = fail(CF_ERR_CUDA, "cudaMalloc");
    if (rc == 0) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) s->sm_count = prop.multiProcessorCount;
    }
    if (rc == 0 && cudaMalloc(&s->d_accum, sizeof(unsigned long long)) != cudaSuccess) rc = fail(CF_ERR_CUDA, "cudaMalloc");
    if (rc == 0 && cudaMalloc(&s->d_block_counts, 2 * sizeof(unsigned long long)) != cudaSuccess) rc = fail(CF_ERR_CUDA, "cudaMalloc");
    if (rc != 0) {
        cf_destroy(s);
        return rc;
    }
    s->rng = new GlibcRand();
    default_tables(s, *s->rng);
    *out = s;
    return CF_OK;
}

extern "C" int cf_destroy(cf_sim* s) {
    if (!s) return CF_OK;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (auto& g : s->step_graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    for (auto& g : s->graph_graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    slab_free(s);
    free_particle_buffers(s);
    cudaFree(s->hist);
    cudaFree(s->cell_start);
    cudaFree(s->d_tables);
    cudaFree(s->d_edge_count);
    cudaFree(s->d_graph_occ);
    cudaFree(s->d_vertices);
    cudaFree(s->d_colors);
    cudaFree(s->d_accum);
    cudaFree(s->d_block_counts);
    cudaFree(s->d_tiles);
    for (int b = 0; b < 2; b++) cudaFree(s->gk[b]), cudaFree(s->gv[b]);
    cudaFree(s->gpos);
    cudaFree(s->gstart);
    cudaFree(s->d_tile_ctrl);
    cudaFree(s->d_cell_occ);
    if (s->h_cell_occ) cudaFreeHost(s->h_cell_occ);
    if (s->h_graph_occ_pin) cudaFreeHost(s->h_graph_occ_pin);
    cudaFree(s->d_half);
    for (int b = 0; b < 2; b++) {
        cudaFree(s->hk[b]);
        cudaFree(s->hv[b]);
    }
    cudaFree(s->h_cell_of);
    cudaFree(s->h_pos);
    cudaFree(s->d_qbox);
    for (int a = 0; a < 3; a++) cudaFree(s->h_xyz[a]);
    cudaFree(s->h_comp);
    cudaFree(s->h_start);
    for (auto& ev : s->ev_pool)
        for (int i = 0; i < CF_STEP_EVENTS; i++) cudaEventDestroy(ev.e[i]);
    for (cudaEvent_t e : s->gev_pool) cudaEventDestroy(e);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s->rng;
    delete s;
    return CF_OK;
}

static int slab_refresh(cf_sim* s);
extern "C" int cf_get_particle_count(const cf_sim* s) {
    if (!s) return CF_ERR_ARG;
    if (s->slab) { // the owned count lives on the device
        cf_sim* m = const_cast<cf_sim*>(s);
        if (cudaSetDevice(m->device) != cudaSuccess) return CF_ERR_CUDA;
        if (int rc = slab_refresh(m)) return rc;
    }
    return s->n;
}
extern "C" int cf_get_num_particle_types(const cf_sim* s) { return s ? s->T : CF_ERR_ARG; }

// setParticleCount (.cu:564-572): reallocate and re-spawn when the count changes.
extern "C" int cf_set_particle_count(cf_sim* s, int count) {
    ARG(s && count >= 0);
    if (count == s->n) return CF_OK;
    if (int rc = set_device(s)) return rc;
    CU(cudaStreamSynchronize(s->stream));
    s->n = count;
    if (int rc = alloc_particle_buffers(s, count)) return rc;
    return cf_init_particles(s, 0x5EED0000ull, CF_INIT_SPAWN_CUBE);
}

// setNumParticleTypes (.cu:574-579): new random tables, particles re-spawned.
extern "C" int cf_set_num_particle_types(cf_sim* s, int types) {
    ARG(s && types >= 1 && types <= CF_MAX_PARTICLE_TYPES);
    s->T = types;
    s->params.numParticleTypes = types;
    // successive table draws continue ONE sequence, like the reference process's rand() (.cu:574-579)
    default_tables(s, *s->rng);
    return cf_init_particles(s, 0x5EED0000ull, CF_INIT_SPAWN_CUBE);
}

// ---------------------------------------------------------------------------------------------
// tables
// ---------------------------------------------------------------------------------------------
extern "C" int cf_regenerate_force_table(cf_sim* s) {
    ARG(s);
    GlibcRand& g = *s->rng; // successive calls continue the construction draws, like rand() would
    for (int i = 0; i < s->T * s->T; i++) s->raw[i] = (float)g.next() / 2147483647 * 2.0f - 1.0f;
    squash_force_table(s, 0.28f, -0.20f, 1.0f);
    return CF_OK;
}
extern "C" int cf_set_raw_force_table(cf_sim* s, const float* raw, int count) {
    ARG(s && raw && count >= 0);
    int m = std::min(count, s->T * s->T); // loadPreset writes min(array, T*T) entries
    memcpy(s->raw, raw, sizeof(float) * m);
    return CF_OK;
}
extern "C" int cf_get_raw_force_table(const cf_sim* s, float* raw, int count) {
    ARG(s && raw && count >= s->T * s->T);
    memcpy(raw, s->raw, sizeof(float) * s->T * s->T);
    return CF_OK;
}
extern "C" int cf_update_force_table(cf_sim* s, float range, float bias, float offset) {
    ARG(s);
    s->params.forceRange = range;
    s->params.forceBias = bias;
    s->params.forceOffset = offset;
    squash_force_table(s, range, bias, offset);
    return CF_OK;
}
extern "C" int cf_get_force_table(const cf_sim* s, float* eff, int count) {
    ARG(s && eff && count >= s->T * s->T);
    memcpy(eff, s->force, sizeof(float) * s->T * s->T);
    return CF_OK;
}
extern "C" int cf_set_force_table(cf_sim* s, const float* eff, int count) {
    ARG(s && eff && count == s->T * s->T);
    memcpy(s->force, eff, sizeof(float) * count);
    s->force_overridden = true;
    return CF_OK;
}
extern "C" int cf_set_radio_by_type(cf_sim* s, const float* radio, int count) {
    ARG(s && radio && count >= 0);
    memcpy(s->radio, radio, sizeof(float) * std::min(count, s->T));
    return CF_OK;
}
extern "C" int cf_set_radio_by_type_value(cf_sim* s, int index, float value) {
    ARG(s);
    if (index >= 0 && index < s->T) s->radio[index] = value; // out of range: ignored, .cu:616
    return CF_OK;
}
extern "C" int cf_get_radio_by_type(const cf_sim* s, float* radio, int count) {
    ARG(s && radio && count >= s->T);
    memcpy(radio, s->radio, sizeof(float) * s->T);
    return CF_OK;
}
extern "C" int cf_rotate_radio_by_type(cf_sim* s) { // .cu:594-600
    ARG(s);
    float tmp = s->radio[s->T - 1];
    for (int i = s->T - 1; i > 0; i--) s->radio[i] = s->radio[i - 1];
    s->radio[0] = tmp;
    return CF_OK;
}

// ---------------------------------------------------------------------------------------------
// particle state
// ---------------------------------------------------------------------------------------------
extern "C" int cf_init_particles(cf_sim* s, uint64_t seed, int mode) {
    ARG(s && (mode == CF_INIT_SPAWN_CUBE || mode == CF_INIT_UNIFORM));
    if (int rc = set_device(s)) return rc;
    if (s->slab) return slab_init_particles(s, s->n_total, seed, mode);
    if (s->n > 0)
        LAUNCH(s, init_particles_kernel, div_up(s->n, 256), 256, 0, opos(s), ovel(s), ofrc(s),
               oid(s), s->n, 0, s->T, seed, mode, s->params.canvasWidth, s->params.canvasHeight,
               s->params.canvasDepth);
    s->sorted_valid = false, s->state_gen++;
    CU(cudaGetLastError());
    return CF_OK;
}

extern "C" int cf_init_particles_global(cf_sim* s, int64_t n_total, uint64_t seed, int mode) {
    ARG(s && (mode == CF_INIT_SPAWN_CUBE || mode == CF_INIT_UNIFORM));
    if (!s->slab) return fail(CF_ERR_STATE, "cf_init_particles_global needs cf_comm_init first");
    if (int rc = set_device(s)) return rc;
    return slab_init_particles(s, n_total, seed, mode);
}

extern "C" int cf_slab_bounds(cf_sim* s, float* lo, float* hi) {
    ARG(s && lo && hi);
    if (!s->slab) {
        *lo = 0.f;
        *hi = s->params.canvasWidth;
        return CF_OK;
    }
    *lo = slab_bound(s, s->rank);
    *hi = slab_bound(s, s->rank + 1);
    return CF_OK;
}

static int upload_impl(cf_sim* s, const cf_particle* aos, const int32_t* counts, const int32_t* ids,
                       int count) {
    if (int rc = set_device(s)) return rc;
    if (s->slab) {
        if (count > s->cap_own) return fail(CF_ERR_CAPACITY, "%d particles exceed the rank capacity %d", count, s->cap_own);
        if (int rc = slab_drain(s)) return rc; // collective, like the upload itself
        if (int rc = slab_set_owned_count(s, count)) return rc;
    } else if (count > s->cap) {
        CU(cudaStreamSynchronize(s->stream));
        if (int rc = alloc_particle_buffers(s, count)) return rc;
    }
    s->n = count;
    if (count == 0) return CF_OK;
    CU(cudaMemcpyAsync(s->d_aos, aos, sizeof(AosParticle) * (size_t)count, cudaMemcpyHostToDevice, s->stream));
    int* d_ids = nullptr;
    if (counts)
        CU(cudaMemcpyAsync(s->d_counts, counts, sizeof(int) * (size_t)count, cudaMemcpyHostToDevice, s->stream));
    if (ids) {
        d_ids = (int*)s->keys[1]; // scratch: the sort is invalidated anyway
        CU(cudaMemcpyAsync(d_ids, ids, sizeof(int) * (size_t)count, cudaMemcpyHostToDevice, s->stream));
    }
    LAUNCH(s, aos_to_soa_kernel, div_up(count, 256), 256, 0, s->d_aos, counts ? s->d_counts : nullptr, d_ids,
           opos(s), ovel(s), ofrc(s), oid(s), count, s->T);
    s->sorted_valid = false, s->state_gen++;
    CU(cudaGetLastError());
    return CF_OK;
}

extern "C" int cf_upload_particles(cf_sim* s, const cf_particle* aos, int count) {
    ARG(s && aos && count >= 0);
    ARG(count == s->n);
    return upload_impl(s, aos, nullptr, nullptr, count);
}

extern "C" int cf_upload_particles_ids(cf_sim* s, const cf_particle* aos, const int32_t* counts,
                                       const int32_t* ids, int count) {
    ARG(s && (aos || count == 0) && count >= 0);
    return upload_impl(s, aos, counts, ids, count);
}

extern "C" int cf_download_particles(cf_sim* s, cf_particle* aos, int count) {
    ARG(s && aos);
    if (int rc = set_device(s)) return rc;
    if (int rc = slab_refresh(s)) return rc;
    ARG(count == s->n);
    if (count == 0) return CF_OK;
    LAUNCH(s, soa_to_aos_kernel, div_up(count, 256), 256, 0, opos(s), ovel(s), ofrc(s),
           oid(s), s->d_aos, s->d_counts, count, 1);
    CU(cudaMemcpyAsync(aos, s->d_aos, sizeof(AosParticle) * (size_t)count, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return CF_OK;
}

extern "C" int cf_download_particles_ids(cf_sim* s, cf_particle* aos, int32_t* counts, int32_t* ids,
                                         int capacity, int* count) {
    ARG(s && count);
    if (int rc = set_device(s)) return rc;
    if (int rc = slab_refresh(s)) return rc; // slab mode: the owned count lives on the device
    *count = s->n;
    ARG(capacity >= s->n);
    if (int rc = set_device(s)) return rc;
    if (s->n == 0) return CF_OK;
    LAUNCH(s, soa_to_aos_kernel, div_up(s->n, 256), 256, 0, opos(s), ovel(s), ofrc(s),
           oid(s), s->d_aos, s->d_counts, s->n, 0);
    if (aos) CU(cudaMemcpyAsync(aos, s->d_aos, sizeof(AosParticle) * (size_t)s->n, cudaMemcpyDeviceToHost, s->stream));
    if (counts) CU(cudaMemcpyAsync(counts, s->d_counts, sizeof(int) * (size_t)s->n, cudaMemcpyDeviceToHost, s->stream));
    if (ids) CU(cudaMemcpyAsync(ids, oid(s), sizeof(int) * (size_t)s->n, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return CF_OK;
}

extern "C" int cf_upload_neighbor_counts(cf_sim* s, const int32_t* counts, int count) {
    ARG(s && counts);
    if (int rc = set_device(s)) return rc;
    if (int rc = slab_refresh(s)) return rc;
    ARG(count == s->n);
    if (count == 0) return CF_OK;
    CU(cudaMemcpyAsync(s->d_counts, counts, sizeof(int) * (size_t)count, cudaMemcpyHostToDevice, s->stream));
    LAUNCH(s, scatter_counts_kernel, div_up(count, 256), 256, 0, s->d_counts, oid(s), ovel(s), count);
    CU(cudaGetLastError());
    return CF_OK;
}

extern "C" int cf_download_neighbor_counts(cf_sim* s, int32_t* counts, int count) {
    ARG(s && counts);
    if (int rc = set_device(s)) return rc;
    if (int rc = slab_refresh(s)) return rc;
    ARG(count == s->n);
    if (count == 0) return CF_OK;
    LAUNCH(s, soa_to_aos_kernel, div_up(count, 256), 256, 0, opos(s), ovel(s), ofrc(s),
           oid(s), s->d_aos, s->d_counts, count, 1);
    CU(cudaMemcpyAsync(counts, s->d_counts, sizeof(int) * (size_t)count, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return CF_OK;
}

extern "C" int cf_render_feed(cf_sim* s, float* xyzt, int capacity, int32_t* type_counts, const void** device_ptr) {
    ARG(s);
    if (int rc = set_device(s)) return rc;
    if (int rc = slab_refresh(s)) return rc;
    ARG(xyzt == nullptr || capacity >= s->n);
    if (device_ptr) *device_ptr = nullptr;
    if (s->n == 0) {
        if (type_counts) memset(type_counts, 0, sizeof(int32_t) * s->T);
        return CF_OK;
    }
    // d_aos (44 B per slot) doubles as the float4 output; d_counts holds the T counters
    float4* out = reinterpret_cast<float4*>(s->d_aos);
    CU(cudaMemsetAsync(s->d_counts, 0, sizeof(int) * CF_T_MAX, s->stream));
    LAUNCH(s, render_feed_kernel, div_up(s->n, 256), 256, 0, opos(s), oid(s), s->n, s->slab ? 0 : 1, out, s->d_counts, s->T);
    if (xyzt) CU(cudaMemcpyAsync(xyzt, out, sizeof(float4) * (size_t)s->n, cudaMemcpyDeviceToHost, s->stream));
    if (type_counts) CU(cudaMemcpyAsync(type_counts, s->d_counts, sizeof(int) * s->T, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if (device_ptr) *device_ptr = out; // valid until the next upload / download / render feed call
    return CF_OK;
}

extern "C" int cf_move_universe(cf_sim* s, float dx, float dy, float dz) {
    ARG(s);
    if (int rc = set_device(s)) return rc;
    if (s->slab) { // consume a pending migrant exchange first; the shift may move particles across a face
        if (int rc = slab_drain(s)) return rc;
        if (int rc = slab_refresh(s)) return rc;
    }
    if (s->n == 0) return CF_OK;
    // the reference passes the DEFAULT canvas here (.cu:586-589); the engine uses the live one
    LAUNCH(s, move_universe_kernel, div_up(s->n, 256), 256, 0, opos(s), s->n, dx, dy, dz,
           s->params.canvasWidth, s->params.canvasHeight, s->params.canvasDepth);
    s->sorted_valid = false, s->state_gen++;
    CU(cudaGetLastError());
    return CF_OK;
}

// ---------------------------------------------------------------------------------------------
// stepping
// ---------------------------------------------------------------------------------------------
extern "C" int cf_set_params(cf_sim* s, const cf_params* p) {
    ARG(s && p);
    ARG(p->numParticleTypes == s->T);
    s->params = *p;
    return CF_OK;
}
extern "C" int cf_get_params(const cf_sim* s, cf_params* p) {
    ARG(s && p);
    *p = s->params;
    return CF_OK;
}

extern "C" float cf_ratio_with_lfo(const cf_params* p, float t) { // CellFlowWidget.cpp:415-421
    if (!p) return 0.f;
    if (p->lfoA != 0.0f) {
        double lfo = (double)p->lfoA * sin(2.0f * M_PI * (double)p->lfoS * (double)t);
        return (float)((double)p->ratio + lfo);
    }
    return p->ratio;
}

static StepEvents* next_events(cf_sim* s) {
    if (!s->opt_timing) return nullptr;
    if (s->ev_used == s->ev_pool.size()) {
        if (s->ev_pool.size() >= 4096) return nullptr;
        StepEvents ev;
        for (int i = 0; i < CF_STEP_EVENTS; i++)
            if (cudaEventCreate(&ev.e[i]) != cudaSuccess) return nullptr;
        s->ev_pool.push_back(ev);
    }
    return &s->ev_pool[s->ev_used++];
}

// Builds the type-homogeneous j copy for the generation-4 tile kernel (per-type radii): one
// stable sort of the cell-sorted slots on key = row * T + type, a gather and a bounds pass.
static int build_homog_copy(cf_sim* s) {
    const int nz = s->sc.dims[2];
    const int nrow = s->ncell / nz;
    // slab mode: every slot up to the end of the right ghost layer, a count only the device knows
    // (cell_start[ncell]); slots before the left ghost layer hold nothing and get the sentinel key
    const int nslots = s->slab ? s->cap : s->n;
    const int* d_nslots = s->slab ? s->d_slab + SLAB_NSLOTS : nullptr;
    const int* d_first = s->slab ? s->d_slab + SLAB_FIRST : nullptr;
    const int nkeys = nrow * s->T * nz; // composite keys (row * T + type) * nz + cz
    if ((size_t)nslots > s->homog_cap) {
        CU(cudaStreamSynchronize(s->stream));
        for (int b = 0; b < 2; b++) {
            cudaFree(s->hk[b]);
            cudaFree(s->hv[b]);
            s->hk[b] = s->hv[b] = nullptr;
        }
        cudaFree(s->h_cell_of);
        cudaFree(s->h_pos);
        cudaFree(s->h_comp);
        s->h_cell_of = nullptr, s->h_pos = nullptr, s->h_comp = nullptr;
        for (int a = 0; a < 3; a++) cudaFree(s->h_xyz[a]), s->h_xyz[a] = nullptr;
        s->homog_cap = (size_t)nslots + (size_t)nslots / 8 + 1024;
        for (int b = 0; b < 2; b++) {
            CU(cudaMalloc(&s->hk[b], s->homog_cap * sizeof(uint32_t)));
            CU(cudaMalloc(&s->hv[b], s->homog_cap * sizeof(uint32_t)));
        }
        CU(cudaMalloc(&s->h_cell_of, s->homog_cap * sizeof(int)));
        CU(cudaMalloc(&s->h_pos, s->homog_cap * sizeof(float4)));
        for (int a = 0; a < 3; a++) { // + one chunk of slack: a bulk copy always moves 128 elements
            CU(cudaMalloc(&s->h_xyz[a], (s->homog_cap + 256) * sizeof(float)));
            CU(cudaMemsetAsync(s->h_xyz[a], 0, (s->homog_cap + 256) * sizeof(float), s->stream));
        }
        CU(cudaMalloc(&s->h_comp, s->homog_cap * sizeof(uint32_t)));
    }
    if ((size_t)nkeys + 2 > s->h_start_cap) {
        CU(cudaStreamSynchronize(s->stream));
        cudaFree(s->h_start);
        s->h_start = nullptr;
        s->h_start_cap = (size_t)nkeys + 2 + (size_t)nkeys / 4;
        CU(cudaMalloc(&s->h_start, s->h_start_cap * sizeof(int)));
    }
    const float4* pos = s->pos[s->cur];
    LAUNCH(s, homog_key_kernel, div_up(nslots, 256), 256, 0, pos, s->cell_start, s->ncell, nz, s->T, nslots, d_nslots,
           d_first, s->hk[0], s->hv[0], s->h_cell_of);
    int src = 0;
    if (int rc = radix_sort_pairs(s, s->hk, s->hv, nslots, (long long)nrow * s->T + 1, &src, d_nslots)) return rc;
    LAUNCH(s, homog_gather_kernel, div_up(nslots, 256), 256, 0, s->hk[src], s->hv[src], pos, s->h_cell_of, nz, nslots,
           d_nslots, s->h_pos, s->h_comp, (s->opt_t4_stage == 1 || s->opt_t4_stage == 3) ? s->h_xyz[0] : nullptr, s->h_xyz[1], s->h_xyz[2]);
    LAUNCH(s, homog_bounds_kernel, div_up(nkeys + 1, 256), 256, 0, s->h_comp, nslots, d_nslots, s->h_start, nkeys);
    return 0;
}

// Which pair-force kernel the next step uses (1 per-particle, 2 tile generation 3, 3 tile generation 4).
static int choose_force_kernel(const cf_sim* s) {
    int kernel = s->opt_force_kernel;
    const int global_nx = s->slab ? s->nxl * s->world : s->sc.dims[0];
    // the tile kernels decide the minimum-image wrap per neighbour cell: >= 4 cells per periodic axis
    const bool wrap_ok = s->sc.dims[1] >= 4 && s->sc.dims[2] >= 4 && global_nx >= 4;
    const bool v3_ok = wrap_ok && (s->sc.uniform_radius || s->half_bound_ok);
    // particles per cell as a particle sees it: the uniform estimate, or the measured sum n_c^2 / n of
    // a previous step when that is larger (clustered states)
    const double n = (double)std::max(s->n, 1);
    double occ = n / (double)std::max(s->ncell, 1);
    occ = std::max(occ, s->policy_occ);
    const bool tile_ok = wrap_ok && tile_kernel_applicable(occ);
    if (kernel == 0) {
        // generation 4 streams one sub-run per (neighbour run, type) when the radii differ per type:
        // it needs sub-runs long enough to fill its chunks
        const double subrun = 3.0 * occ / (s->sc.uniform_radius ? 1.0 : (double)s->T);
        kernel = !tile_ok ? 1 : (subrun >= 40.0 ? 3 : (v3_ok ? 2 : 1));
    }
    if (kernel == 3 && !wrap_ok) kernel = 1;
    if (kernel == 2 && !v3_ok) kernel = 1;
    return kernel;
}

// Call after a stream synchronisation: the statistic of the last completed step becomes the policy input.
static void refresh_policy(cf_sim* s) {
    if (s->h_cell_occ && s->n > 0) s->policy_occ = (double)s->occ_stride * (double)*s->h_cell_occ / (double)s->n;
}

static int launch_force(cf_sim* s) {
    int n = s->n;
    const float4* pos = s->pos[s->cur];
    const int kernel = s->planned_force_kernel;
    // refresh the occupancy statistic for the choice of a later step (one small launch, no synchronisation)
    const int occ_stride = s->ncell > (1 << 16) ? 8 : 1;
    s->occ_stride = occ_stride;
    LAUNCH(s, cell_occupancy_kernel, div_up(div_up(s->ncell, occ_stride), 256), 256, 0, s->cell_start, s->ncell, occ_stride,
           s->d_cell_occ, s->h_cell_occ);
    s->last_force_kernel = kernel;
    if (kernel == 2 || kernel == 3) {
        const bool homog = kernel == 3 && !s->sc.uniform_radius;
        // particles per tile: 128 (generation 3), 32 x layers of the generation-4 instantiation in use
        const int tile_i = kernel == 3 ? 32 * t4_ipt(homog ? 1 : 0) : TK_TI;
        size_t need = (size_t)s->ncell + (size_t)(s->slab ? s->cap_own : n) / tile_i + 2;
        if (need > s->tiles_cap) {
            CU(cudaStreamSynchronize(s->stream));
            cudaFree(s->d_tiles);
            s->d_tiles = nullptr;
            s->tiles_cap = need + need / 4;
            CU(cudaMalloc(&s->d_tiles, s->tiles_cap * sizeof(int2)));
        }
        if (homog)
            if (int rc = build_homog_copy(s)) return rc;
        CU(cudaMemsetAsync(s->d_tile_ctrl, 0, 2 * sizeof(int), s->stream));
        for (int part = 0; part < 2; part++) // full tiles first, partly filled ones last
            LAUNCH(s, build_tiles_kernel, div_up(s->ncell, 256), 256, 0, s->cell_start, s->ncell, s->sc.x_off,
                   s->sc.x_off + s->sc.x_cells - 1, s->sc.dims[1] * s->sc.dims[2], s->d_tiles, s->d_tile_ctrl, part,
                   tile_i);
        if (kernel == 3) { // persistent grid: 8 (per-type radii) or 6 (uniform radius) CTAs of 4 independent warps per SM
            // experiment knob (cf_set_option "t4_ctas_per_sm"): fewer resident CTAs, enforced with dynamic shared memory
            const int minb = t4_minb(t4_ipt(homog ? 1 : 0));
            const int ctas = s->opt_t4_ctas > 0 ? std::min(s->opt_t4_ctas, minb) : minb;
            const int grid = s->sm_count * ctas;
            const size_t pad = ctas < minb ? (size_t)(220 * 1024 / ctas) - 36 * 1024 : 0;
            if (pad) {
                cudaFuncSetAttribute(force_tile4_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
                cudaFuncSetAttribute(force_tile4_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
                cudaFuncSetAttribute(force_tile4_kernel<1, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
                cudaFuncSetAttribute(force_tile4_kernel<0, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
            }
            if (s->opt_count_blocks) { // instrumented build: exact-tested and evaluated (layer, quad) blocks
                CU(cudaMemsetAsync(s->d_block_counts, 0, 2 * sizeof(unsigned long long), s->stream));
                if (homog && s->opt_t4_stage == 4)
                    LAUNCH(s, (force_tile4_kernel<1, true, 4>), grid, T4_WARPS * 32, pad, pos, s->cell_start, s->h_pos, s->h_start,
                           s->d_tiles, s->d_tile_ctrl, s->frc, s->sc, s->d_tables, s->d_block_counts);
                else if (s->opt_t4_stage == 4)
                    LAUNCH(s, (force_tile4_kernel<0, true, 4>), grid, T4_WARPS * 32, pad, pos, s->cell_start, pos, s->cell_start,
                           s->d_tiles, s->d_tile_ctrl, s->frc, s->sc, s->d_tables, s->d_block_counts);
                else if (homog)
                    LAUNCH(s, (force_tile4_kernel<1, true>), grid, T4_WARPS * 32, pad, pos, s->cell_start, s->h_pos, s->h_start,
                           s->d_tiles, s->d_tile_ctrl, s->frc, s->sc, s->d_tables, s->d_block_counts);
                else
                    LAUNCH(s, (force_tile4_kernel<0, true>), grid, T4_WARPS * 32, pad, pos, s->cell_start, pos, s->cell_start,
                           s->d_tiles, s->d_tile_ctrl, s->frc, s->sc, s->d_tables, s->d_block_counts);
                s->block_counts_valid = true;
                return 0;
            }
            if (s->opt_t4_stage == 2) {
                // quad boxes of the j array the kernel streams: the type-sorted copy (element e), or the sorted slots
                const float4* jarr = homog ? s->h_pos : pos;
                const int upper = s->slab ? s->cap : s->n;
                const size_t nq = (size_t)upper / 4 + 2;
                if (nq > s->qbox_cap) {
                    CU(cudaStreamSynchronize(s->stream));
                    cudaFree(s->d_qbox);
                    s->d_qbox = nullptr;
                    s->qbox_cap = nq + nq / 8;
                    CU(cudaMalloc(&s->d_qbox, s->qbox_cap * 2 * sizeof(float4)));
                }
                const int* d_cnt = s->slab ? s->d_slab + SLAB_NSLOTS : nullptr;
                const int* d_first = (s->slab && !homog) ? s->d_slab + SLAB_FIRST : nullptr; // the copy starts at element 0
                LAUNCH(s, quad_box_kernel, div_up(upper / 4 + 1, 256), 256, 0, jarr, 0, d_first, upper, d_cnt, s->d_qbox);
                if (homog)
                    LAUNCH(s, (force_tile4_kernel<1, false, 2>), grid, T4_WARPS * 32, pad, pos, s->cell_start, s->h_pos, s->h_start,
                           s->d_tiles, s->d_tile_ctrl, s->frc, s->sc, s->d_tables, nullptr, nullptr, nullptr, nullptr, s->d_qbox);
                else
                    LAUNCH(s, (force_tile4_kernel<0, false, 2>), grid, T4_WARPS * 32, pad, pos, s->cell_start, pos, s->cell_start,
                           s->d_tiles, s->d_tile_ctrl, s->frc, s->sc, s->d_tables, nullptr, nullptr, nullptr, nullptr, s->d_qbox);
                return 0;
            }
            if (s->opt_t4_stage == 4) { // live quads stored compacted (kernels_tile4.cuh, STAGE 4)
                if (homog)
                    LAUNCH(s, (force_tile4_kernel<1, false, 4>), grid, T4_WARPS * 32, pad, pos, s->cell_start, s->h_pos, s->h_start,
                           s->d_tiles, s->d_tile_ctrl, s->frc, s->sc, s->d_tables, nullptr);
                else
                    LAUNCH(s, (force_tile4_kernel<0, false, 4>), grid, T4_WARPS * 32, pad, pos, s->cell_start, pos, s->cell_start,
                           s->d_tiles, s->d_tile_ctrl, s->frc, s->sc, s->d_tables, nullptr);
                return 0;
            }
            if (homog && s->opt_t4_stage == 3)
                LAUNCH(s, (force_tile4_kernel<1, false, 3>), grid, T4_WARPS * 32, pad, pos, s->cell_start, s->h_pos, s->h_start,
                       s->d_tiles, s->d_tile_ctrl, s->frc, s->sc, s->d_tables, nullptr, s->h_xyz[0], s->h_xyz[1], s->h_xyz[2]);
            else if (homog && s->opt_t4_stage == 1)
                LAUNCH(s, (force_tile4_kernel<1, false, 1>), grid, T4_WARPS * 32, pad, pos, s->cell_start, s->h_pos, s->h_start,
                       s->d_tiles, s->d_tile_ctrl, s->frc, s->sc, s->d_tables, nullptr, s->h_xyz[0], s->h_xyz[1], s->h_xyz[2]);
            else if (homog)
                LAUNCH(s, (force_tile4_kernel<1, false>), grid, T4_WARPS * 32, pad, pos, s->cell_start, s->h_pos, s->h_start,
                       s->d_tiles, s->d_tile_ctrl, s->frc, s->sc, s->d_tables, nullptr);
            else
                LAUNCH(s, (force_tile4_kernel<0, false>), grid, T4_WARPS * 32, pad, pos, s->cell_start, pos, s->cell_start,
                       s->d_tiles, s->d_tile_ctrl, s->frc, s->sc, s->d_tables, nullptr);
            return 0;
        }
        const int grid = s->sm_count * 4;
        if (s->sc.uniform_radius)
            LAUNCH(s, force_tile_kernel<true>, grid, TK_THREADS, 0, pos, s->cell_start, s->d_tiles, s->d_tile_ctrl,
                   s->frc, s->sc, s->d_tables, 0.f, s->d_half);
        else
            LAUNCH(s, force_tile_kernel<false>, grid, TK_THREADS, 0, pos, s->cell_start, s->d_tiles, s->d_tile_ctrl,
                   s->frc, s->sc, s->d_tables, 0.f, s->d_half);
        return 0;
    }
    const int nu = s->slab ? s->cap_own : n; // slab mode: the owned count is read on the device
    const int* dn = s->slab ? s->d_slab + SLAB_NCUR : nullptr;
    if (s->sc.uniform_radius)
        LAUNCH(s, force_pp_kernel<true>, div_up(nu, 128), 128, 0, pos, s->cell_start, s->frc, s->base, nu, dn, s->sc, s->d_tables);
    else
        LAUNCH(s, force_pp_kernel<false>, div_up(nu, 128), 128, 0, pos, s->cell_start, s->frc, s->base, nu, dn, s->sc, s->d_tables);
    return 0;
}

// One step, launched kernel by kernel.  Host-side state it changes: cur, keys/vals order,
// sorted_valid (mirrored by replay_step_host_state for graph replays).
static int step_direct(cf_sim* s, StepEvents* ev) {
    if (ev) CU(cudaEventRecord(ev->e[0], s->stream));
    if (ev) ev->has_exchange = s->slab && !s->sorted_valid;
    {
        NvtxRange r("cellflow:cell_list_build");
        if (int rc = build_cell_list(s, ev ? &ev->e[4] : nullptr)) return rc;
    }
    if (ev) CU(cudaEventRecord(ev->e[1], s->stream));
    if (s->n > 0 || s->slab) {
        NvtxRange r("cellflow:pair_force");
        if (int rc = launch_force(s)) return rc;
    }
    if (ev) CU(cudaEventRecord(ev->e[2], s->stream));
    NvtxRange r_int("cellflow:integrate");
    if (s->slab) {
        // fused integrate + migrant emission: the leavers go straight into the neighbours' mailboxes
        s->seq_mig++;
        LAUNCH(s, integrate_slab_kernel, slab_grid(s), 256, 0, opos(s), ovel(s), ofrc(s), oid(s), s->cap_own, s->sc,
               s->geom, slab_peers(s), s->seq_mig);
        s->mig_sent = true;
    } else if (s->n > 0) {
        LAUNCH(s, integrate_kernel, div_up(s->n, 256), 256, 0, opos(s), ovel(s), ofrc(s), s->n, s->sc);
    }
    if (ev) CU(cudaEventRecord(ev->e[3], s->stream));
    s->sorted_valid = false, s->state_gen++;
    return 0;
}

static std::vector<char> step_signature(const cf_sim* s) {
    std::vector<char> sig;
    auto put = [&](const void* p, size_t n) { sig.insert(sig.end(), (const char*)p, (const char*)p + n); };
    put(&s->sc, sizeof(s->sc));
    const void* ptrs[] = {s->pos[0], s->pos[1], s->vel[0], s->vel[1], s->id[0], s->id[1], s->frc, s->keys[0],
                          s->keys[1], s->vals[0], s->vals[1], s->hist, s->cell_start, s->d_tiles, s->d_tile_ctrl,
                          s->d_tables, s->d_half, s->hk[0], s->hk[1], s->hv[0], s->hv[1], s->h_cell_of, s->h_pos,
                          s->h_comp, s->h_start, s->h_xyz[0], s->h_xyz[1], s->h_xyz[2], s->d_qbox};
    put(ptrs, sizeof(ptrs));
    int ints[] = {s->n, s->ncell, s->cur, s->opt_force_kernel, s->sorted_valid ? 1 : 0, s->half_bound_ok ? 1 : 0,
                  s->planned_force_kernel, s->opt_t4_ctas, s->opt_count_blocks, s->opt_t4_stage};
    put(ints, sizeof(ints));
    return sig;
}

// Steps of the single-GPU engine are a fixed kernel sequence for fixed parameters, so after one
// direct run (which sizes every buffer) the sequence is captured into a CUDA graph and replayed:
// ~15 launches become one, which is what bounds ms/step at 100 k particles and below.
static int step_graphed(cf_sim* s, StepEvents* ev) {
    std::vector<char> sig = step_signature(s);
    cf_sim::StepGraph* hit = nullptr;
    cf_sim::StepGraph* seen = nullptr;
    cf_sim::StepGraph* spare = nullptr;
    for (auto& g : s->step_graphs) {
        if (g.exec && g.sig == sig) hit = &g;
        else if (!g.exec && g.seen == sig) seen = &g;
        else if (!g.exec && g.seen.empty() && !spare) spare = &g;
    }
    if (hit) {
        if (ev) {
            CU(cudaEventRecord(ev->e[0], s->stream));
            CU(cudaEventRecord(ev->e[1], s->stream));
            CU(cudaEventRecord(ev->e[2], s->stream));
            ev->has_exchange = false;
        }
        CU(cudaGraphLaunch(hit->exec, s->stream));
        if (ev) CU(cudaEventRecord(ev->e[3], s->stream));
        if (!s->sorted_valid) { // what ensure_sorted does on the host
            if (hit->swap_keys) std::swap(s->keys[0], s->keys[1]), std::swap(s->vals[0], s->vals[1]);
            s->cur ^= 1, s->state_gen++;
        }
        s->sorted_valid = false, s->state_gen++;
        s->launches += hit->nodes;
        return 0;
    }
    if (seen) {
        uint32_t* k0 = s->keys[0];
        long long before = s->launches;
        if (ev) {
            CU(cudaEventRecord(ev->e[0], s->stream));
            CU(cudaEventRecord(ev->e[1], s->stream));
            CU(cudaEventRecord(ev->e[2], s->stream));
            ev->has_exchange = false;
        }
        CU(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
        int rc = step_direct(s, nullptr);
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
        if (rc) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        if (e != cudaSuccess) return fail(CF_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&seen->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) return fail(CF_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
        seen->sig = sig;
        seen->seen.clear();
        seen->nodes = s->launches - before;
        seen->swap_keys = s->keys[0] != k0;
        CU(cudaGraphLaunch(seen->exec, s->stream)); // the capture recorded, it did not run
        if (ev) CU(cudaEventRecord(ev->e[3], s->stream));
        return 0;
    }
    // first time with these parameters: run directly (allocations happen here), remember it
    if (!spare) { // all slots taken by other signatures: recycle them
        for (auto& g : s->step_graphs) {
            if (g.exec) cudaGraphExecDestroy(g.exec);
            g = cf_sim::StepGraph();
        }
        spare = &s->step_graphs[0];
    }
    int rc = step_direct(s, ev);
    spare->seen = sig;
    return rc;
}

extern "C" int cf_step(cf_sim* s, const cf_params* p, int n_steps) {
    ARG(s && n_steps >= 0);
    if (p) {
        if (int rc = cf_set_params(s, p)) return rc;
    }
    if (int rc = set_device(s)) return rc;
    if (int rc = prepare_step_const(s)) return rc;
    if (s->n == 0 && !s->slab) return CF_OK;
    // per-phase timing (timing == 1) needs the individual launches; timing == 2 times whole steps
    const bool graphs = s->opt_graphs && !s->slab && s->opt_timing != 1 && s->n > 0;
    for (int it = 0; it < n_steps; it++) {
        StepEvents* ev = next_events(s);
        s->planned_force_kernel = choose_force_kernel(s);
        if (int rc = graphs ? step_graphed(s, ev) : step_direct(s, ev)) return rc;
    }
    CU(cudaGetLastError());
    return CF_OK;
}

extern "C" int cf_sync(cf_sim* s) {
    ARG(s);
    if (int rc = set_device(s)) return rc;
    CU(cudaStreamSynchronize(s->stream));
    refresh_policy(s);
    if (int rc = slab_refresh(s)) return rc; // slab mode: owned count + the device-side error word
    return CF_OK;
}

extern "C" int cf_step_host(cf_sim* s, const cf_params* p, const cf_particle* in, const int32_t* counts_in,
                            cf_particle* out, int32_t* counts_out, int count) {
    ARG(s && in && out && count >= 0);
    if (int rc = upload_impl(s, in, counts_in, nullptr, count)) return rc;
    if (int rc = cf_step(s, p, 1)) return rc;
    if (count == 0) return CF_OK;
    LAUNCH(s, soa_to_aos_kernel, div_up(count, 256), 256, 0, opos(s), ovel(s), ofrc(s),
           oid(s), s->d_aos, s->d_counts, count, 1);
    CU(cudaMemcpyAsync(out, s->d_aos, sizeof(AosParticle) * (size_t)count, cudaMemcpyDeviceToHost, s->stream));
    if (counts_out)
        CU(cudaMemcpyAsync(counts_out, s->d_counts, sizeof(int) * (size_t)count, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    refresh_policy(s);
    return CF_OK;
}

// ---------------------------------------------------------------------------------------------
// proximity graph
// ---------------------------------------------------------------------------------------------
// The graph events of the previous build have completed (every build ends with a stream
// synchronisation): fold them into the totals before the pair of events is reused.
static void fold_graph_timing(cf_sim* s) {
    size_t kept = 0;
    for (size_t i = 0; i + 1 < s->gev_used; i += 2) {
        float g = 0;
        cudaError_t e = cudaEventElapsedTime(&g, s->gev_pool[i], s->gev_pool[i + 1]);
        if (e == cudaSuccess) {
            s->ms_graph = g;
            s->ms_graph_total += g;
            s->graph_builds++;
        } else if (e == cudaErrorNotReady) { // an asynchronous build still in flight: keep its pair for later
            std::swap(s->gev_pool[kept], s->gev_pool[i]);
            std::swap(s->gev_pool[kept + 1], s->gev_pool[i + 1]);
            kept += 2;
        }
    }
    (void)cudaGetLastError();
    s->gev_used = kept;
}
// A pair of events for the build being enqueued (nullptr when the pool cannot grow).
static cudaEvent_t* next_graph_events(cf_sim* s) {
    if (s->gev_used + 2 > s->gev_pool.size()) {
        if (s->gev_pool.size() >= 8192) return nullptr;
        cudaEvent_t a, b;
        if (cudaEventCreate(&a) != cudaSuccess) return nullptr;
        if (cudaEventCreate(&b) != cudaSuccess) {
            cudaEventDestroy(a);
            return nullptr;
        }
        s->gev_pool.push_back(a);
        s->gev_pool.push_back(b);
    }
    cudaEvent_t* p = &s->gev_pool[s->gev_used];
    s->gev_used += 2;
    return p;
}

// Everything cf_build_graph launches depends on this plan (plus the buffers and the step constants).
struct GraphPlan {
    int first, count;            // slots [first, first + count) are keyed
    int use_lo_ghost, use_hi_ghost;
    int nkeys, mc, gkernel;
    float dist2;
    GraphGrid g;
};

static int graph_make_plan(cf_sim* s, float dist, int mc, GraphPlan& P) {
    memset(&P, 0, sizeof(P));
    // ---- slots that take part: owned, plus ghost layers that are real neighbours (not the seam) ----
    P.first = s->base, P.count = s->n;
    float ext_lo[3] = {0.f, 0.f, 0.f};
    float ext_hi[3] = {s->sc.W[0], s->sc.W[1], s->sc.W[2]};
    if (s->slab) {
        // a ghost layer is one x layer of the neighbour: the graph may not reach further than the thinnest of them
        float ex_left = 0.f, ex_right = 0.f, ex_min = INFINITY;
        for (int r = 0; r < s->world; r++) {
            const double w = slab_width(s, r);
            const float exr = (float)(w / slab_layers(s, w));
            ex_min = std::min(ex_min, exr);
            if (r == (s->rank + s->world - 1) % s->world) ex_left = exr;
            if (r == (s->rank + 1) % s->world) ex_right = exr;
        }
        if ((double)dist * (1.0 + 1e-5) > (double)ex_min)
            return fail(CF_ERR_ARG, "proximity distance %.1f exceeds the one-cell ghost layer (%.1f) of slab mode", dist, ex_min);
        // every slot that can hold a particle; the key kernel reads the real extents on the device
        P.use_lo_ghost = s->rank > 0, P.use_hi_ghost = s->rank < s->world - 1;
        P.first = 0;
        P.count = s->cap;
        ext_lo[0] = s->geom.x_lo - ex_left;
        ext_hi[0] = s->geom.x_hi + ex_right;
    }
    // ---- the graph's own (type, cell) list, cell edge >= dist ----
    GraphGrid& g = P.g;
    double edge = (double)dist * (1.0 + 1e-5), vol = 1.0;
    for (int a = 0; a < 3; a++) vol *= (double)(ext_hi[a] - ext_lo[a]);
    double max_keys = std::min(4.0 * std::max(s->slab ? s->cap_own : P.count, 0) + 4096.0, 1.6e7);
    edge = std::max(edge, cbrt(vol * s->T / max_keys));
    long long nc = 1;
    for (int a = 0; a < 3; a++) {
        int d = (int)floor((double)(ext_hi[a] - ext_lo[a]) / edge);
        d = std::max(1, std::min(d, 1024));
        g.dims[a] = d;
        g.org[a] = ext_lo[a];
        g.inv[a] = (float)d / (ext_hi[a] - ext_lo[a]);
        nc *= d;
    }
    g.ncell = (int)nc;
    g.T = s->T;
    g.x_min = -INFINITY;
    g.x_max = INFINITY;
    P.nkeys = s->T * g.ncell; // key nkeys = "not gridded"
    P.mc = mc;
    P.dist2 = dist * dist;
    // kernel choice from the occupancy of the PREVIOUS build (read back with its edge count, so no
    // extra synchronisation): ~27 * mean occupancy candidates per particle.  The warp-per-particle kernel wins
    // when a particle has hundreds of candidates AND there are too few particles for one thread each to fill
    // the GPU (tools/graph_crossover.py, the reference's spawn cube: 0.129 vs 0.163 ms at 25 k particles, 0.177 vs
    // 0.194 at 50 k, but 0.299 vs 0.278 at 100 k and 3.70 vs 3.25 at 800 k since the thread-per-particle kernel
    // keeps unordered lists); otherwise thread-per-particle
    P.gkernel = s->opt_graph_kernel;
    // (asynchronous builds never read the count back: the pinned mirror of an earlier build's occupancy stands in;
    //  both kernels produce the same edge set, so the choice may depend on timing)
    double occ = s->graph_mean_occ;
    if (s->h_graph_occ_pin && s->graph_plan_count > 0 && s->n > 0)
        occ = std::max(occ, (double)*s->h_graph_occ_pin / (double)s->graph_plan_count);
    const int n_graph = s->slab ? s->n : P.count;
    if (P.gkernel == 0) P.gkernel = (27.0 * occ >= 250.0 && n_graph < 75000) ? 2 : 1;
    return 0;
}

static int graph_ensure_buffers(cf_sim* s, const GraphPlan& P) {
    const int count = P.count, nkeys = P.nkeys;
    if ((size_t)count > s->graph_cap) {
        CU(cudaStreamSynchronize(s->stream));
        for (int b = 0; b < 2; b++) {
            cudaFree(s->gk[b]);
            cudaFree(s->gv[b]);
            s->gk[b] = s->gv[b] = nullptr;
        }
        cudaFree(s->gpos);
        s->gpos = nullptr;
        s->graph_cap = (size_t)count + (size_t)count / 8 + 1024;
        for (int b = 0; b < 2; b++) {
            CU(cudaMalloc(&s->gk[b], s->graph_cap * sizeof(uint32_t)));
            CU(cudaMalloc(&s->gv[b], s->graph_cap * sizeof(uint32_t)));
        }
        CU(cudaMalloc(&s->gpos, s->graph_cap * sizeof(float4)));
    }
    if ((size_t)nkeys + 2 > s->gstart_cap) {
        CU(cudaStreamSynchronize(s->stream));
        cudaFree(s->gstart);
        s->gstart = nullptr;
        s->gstart_cap = (size_t)nkeys + 2 + (size_t)nkeys / 4;
        CU(cudaMalloc(&s->gstart, s->gstart_cap * sizeof(int)));
    }
    return 0;
}

// The device work of one graph build: (cell list of the current positions,) key, sort, gather,
// bounds, occupancy, graph kernel.  No host synchronisation, no allocation: capturable.
static int graph_device_sequence(cf_sim* s, const GraphPlan& P, bool with_cell_list) {
    if (with_cell_list)
        if (int rc = build_cell_list(s)) return rc;
    CU(cudaMemsetAsync(s->d_edge_count, 0, sizeof(int), s->stream));
    CU(cudaMemsetAsync(s->d_graph_occ, 0, sizeof(unsigned long long), s->stream));
    if (!(P.count > 0 && (s->n > 0 || s->slab))) return 0;
    const int count = P.count, nkeys = P.nkeys;
    // slab mode: slots [0, cell_start[ncell]) take part and the owned count is a device word
    const int* d_n = s->slab ? s->d_slab + SLAB_NSLOTS : nullptr;
    const int* d_first = s->slab ? s->d_slab + SLAB_FIRST : nullptr;
    const int* d_own = s->slab ? s->d_slab + SLAB_NCUR : nullptr;
    LAUNCH(s, graph_key_kernel, div_up(count, 256), 256, 0, s->pos[s->cur], P.first, d_first, count, d_n, P.g, s->cell_start, s->ncell,
           s->base, s->n, d_own, P.use_lo_ghost, P.use_hi_ghost, s->gk[0], s->gv[0]);
    int src = 0;
    if (int rc = radix_sort_pairs(s, s->gk, s->gv, count, (long long)nkeys + 1, &src, d_n)) return rc;
    LAUNCH(s, graph_gather_kernel, div_up(count, 256), 256, 0, s->gv[src], s->pos[s->cur], s->id[s->cur], count, d_n, s->gpos);
    LAUNCH(s, graph_bounds_kernel, div_up(nkeys + 1, 256), 256, 0, s->gk[src], count, d_n, s->gstart, nkeys);
    LAUNCH(s, graph_occupancy_kernel, div_up(nkeys, 256), 256, 0, s->gstart, nkeys, s->d_graph_occ);
    if (P.gkernel == 2)
        LAUNCH(s, graph_warp_kernel, div_up(count, CF_GRAPHW_WARPS), CF_GRAPHW_WARPS * 32, 0, s->gpos, s->gv[src],
               s->gk[src], s->gstart, count, d_n, s->base, s->n, d_own, P.g, P.dist2, P.mc, s->edges, s->edge_slots,
               s->edge_cap, s->d_edge_count);
    else
        LAUNCH(s, graph_kernel, div_up(count, CF_GRAPH_THREADS), CF_GRAPH_THREADS,
               graph_list_bytes(P.mc), s->gpos, s->gv[src], s->gk[src], s->gstart, count, d_n,
               s->base, s->n, d_own, P.g, P.dist2, P.mc, s->edges, s->edge_slots, s->edge_cap, s->d_edge_count);
    return 0;
}

// Single GPU: the sequence above is ~25 small launches, which is what a graph build costs at
// 100 k particles; like the step it is run directly once per parameter set (buffers are sized),
// captured into a CUDA graph the second time and replayed from the third.
static int graph_sequence_cached(cf_sim* s, const GraphPlan& P) {
    std::vector<char> sig = step_signature(s);
    auto put = [&](const void* p, size_t n) { sig.insert(sig.end(), (const char*)p, (const char*)p + n); };
    put(&P, sizeof(P));
    const void* ptrs[] = {s->gk[0], s->gk[1], s->gv[0], s->gv[1], s->gpos, s->gstart, s->edges, s->edge_slots,
                          s->d_edge_count, s->d_graph_occ};
    put(ptrs, sizeof(ptrs));
    put(&s->edge_cap, sizeof(s->edge_cap));
    cf_sim::StepGraph* hit = nullptr;
    cf_sim::StepGraph* seen = nullptr;
    cf_sim::StepGraph* spare = nullptr;
    for (auto& g : s->graph_graphs) {
        if (g.exec && g.sig == sig) hit = &g;
        else if (!g.exec && g.seen == sig) seen = &g;
        else if (!g.exec && g.seen.empty() && !spare) spare = &g;
    }
    const bool was_sorted = s->sorted_valid;
    if (hit) {
        CU(cudaGraphLaunch(hit->exec, s->stream));
        if (!was_sorted) { // what ensure_sorted does on the host
            if (hit->swap_keys) std::swap(s->keys[0], s->keys[1]), std::swap(s->vals[0], s->vals[1]);
            s->cur ^= 1, s->state_gen++;
        }
        s->sorted_valid = true;
        s->launches += hit->nodes;
        return 0;
    }
    if (seen) {
        uint32_t* k0 = s->keys[0];
        long long before = s->launches;
        CU(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
        int rc = graph_device_sequence(s, P, true);
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
        if (rc) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        if (e != cudaSuccess) return fail(CF_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&seen->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) return fail(CF_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
        seen->sig = sig;
        seen->seen.clear();
        seen->nodes = s->launches - before;
        seen->swap_keys = s->keys[0] != k0;
        CU(cudaGraphLaunch(seen->exec, s->stream)); // the capture recorded, it did not run
        return 0;
    }
    if (!spare) {
        for (auto& g : s->graph_graphs) {
            if (g.exec) cudaGraphExecDestroy(g.exec);
            g = cf_sim::StepGraph();
        }
        spare = &s->graph_graphs[0];
    }
    int rc = graph_device_sequence(s, P, true);
    spare->seen = sig;
    return rc;
}

extern "C" int cf_build_graph(cf_sim* s, float dist, int max_conn, int* n_edges) {
    ARG(s);
    if (n_edges) *n_edges = 0;
    ARG(max_conn >= 0);
    if (int rc = set_device(s)) return rc;
    int mc = std::min(max_conn, CF_MAX_GRAPH_CONN);
    if ((s->n == 0 && !s->slab) || mc == 0 || !(dist > 0.f)) {
        s->n_edges = 0;
        return CF_OK;
    }
    if (int rc = prepare_step_const(s)) return rc;
    long long need = (long long)std::max(s->slab ? s->cap_own : s->n, 1) * mc;
    if (need > s->edge_cap) {
        CU(cudaStreamSynchronize(s->stream));
        cudaFree(s->edges);
        cudaFree(s->edge_slots);
        s->edges = s->edge_slots = nullptr;
        CU(cudaMalloc(&s->edges, sizeof(int2) * (size_t)need));
        CU(cudaMalloc(&s->edge_slots, sizeof(int2) * (size_t)need));
        s->edge_cap = (int)need;
    }
    NvtxRange r_graph("cellflow:proximity_graph");
    fold_graph_timing(s);
    cudaEvent_t* gev = s->opt_timing ? next_graph_events(s) : nullptr;
    if (gev) CU(cudaEventRecord(gev[0], s->stream));
    GraphPlan P;
    if (s->slab) {
        // the cell-list build migrates particles (and synchronises with the host): plan afterwards
        if (int rc = build_cell_list(s)) return rc;
        if (int rc = graph_make_plan(s, dist, mc, P)) return rc;
        if (int rc = graph_ensure_buffers(s, P)) return rc;
        if (int rc = graph_device_sequence(s, P, false)) return rc;
    } else {
        if (int rc = graph_make_plan(s, dist, mc, P)) return rc;
        if (int rc = graph_ensure_buffers(s, P)) return rc;
        if (int rc = s->opt_graphs ? graph_sequence_cached(s, P) : graph_device_sequence(s, P, true)) return rc;
    }
    s->last_graph_kernel = P.gkernel;
    if (gev) CU(cudaEventRecord(gev[1], s->stream));
    CU(cudaMemcpyAsync(s->h_graph_occ_pin, s->d_graph_occ, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
    s->graph_gen = s->state_gen;
    s->graph_count_pending = true;
    s->graph_plan_count = s->slab ? std::max(s->n, 1) : P.count;
    if (!n_edges) { // asynchronous: nothing is read back now
        CU(cudaGetLastError());
        return CF_OK;
    }
    return cf_get_graph_edge_count(s, n_edges);
}

extern "C" int cf_get_graph_edge_count(cf_sim* s, int* n_edges) {
    ARG(s && n_edges);
    if (int rc = set_device(s)) return rc;
    if (s->graph_count_pending) {
        CU(cudaMemcpyAsync(&s->h_graph_occ, s->d_graph_occ, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
        CU(cudaMemcpyAsync(&s->n_edges, s->d_edge_count, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        CU(cudaGetLastError());
        s->graph_count_pending = false;
        if (int rc = slab_refresh(s)) return rc;
        if (s->graph_plan_count > 0 && s->n > 0) s->graph_mean_occ = (double)s->h_graph_occ / (double)s->graph_plan_count;
        refresh_policy(s);
    }
    *n_edges = s->n_edges;
    return CF_OK;
}

extern "C" int cf_download_graph_edges(cf_sim* s, cf_edge* edges, int capacity) {
    ARG(s);
    if (s->graph_count_pending) {
        int ne = 0;
        if (int rc = cf_get_graph_edge_count(s, &ne)) return rc;
    }
    ARG(edges || s->n_edges == 0);
    ARG(capacity >= s->n_edges);
    if (int rc = set_device(s)) return rc;
    if (s->n_edges == 0) return CF_OK;
    CU(cudaMemcpyAsync(edges, s->edges, sizeof(int2) * (size_t)s->n_edges, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return CF_OK;
}

// Vertex stream of the last graph (reference VBO layout, .cu:248-276) written on the device: into memory the
// caller owns (device_dst: e.g. the mapped pointer of the widget's VBO, CellFlowWidget.cpp:606-617 — zero copy),
// or into the library's persistent buffer.  Asynchronous on the handle's stream.
extern "C" int cf_graph_vertices_device(cf_sim* s, const cf_color* colors, int num_colors, float* device_dst,
                                        int capacity_edges, const float** device_ptr) {
    ARG(s);
    if (device_ptr) *device_ptr = nullptr;
    if (int rc = set_device(s)) return rc;
    if (s->graph_count_pending) {
        int ne = 0;
        if (int rc = cf_get_graph_edge_count(s, &ne)) return rc;
    }
    ARG(colors && num_colors >= s->T);
    if (s->n_edges == 0) return CF_OK;
    // edge slots refer to the particle order the graph was built on: any later reorder or move invalidates them
    if (s->graph_gen != s->state_gen)
        return fail(CF_ERR_STATE, "graph vertices requested after the particles moved or were reordered: build the graph again");
    float* out = device_dst;
    if (out) {
        ARG(capacity_edges >= s->n_edges);
    } else {
        if ((size_t)s->n_edges > s->vertex_cap) {
            CU(cudaStreamSynchronize(s->stream));
            cudaFree(s->d_vertices);
            s->d_vertices = nullptr;
            s->vertex_cap = (size_t)s->n_edges + (size_t)s->n_edges / 4 + 1024;
            CU(cudaMalloc(&s->d_vertices, sizeof(float) * 12 * s->vertex_cap));
        }
        out = s->d_vertices;
    }
    if (!s->d_colors) CU(cudaMalloc(&s->d_colors, sizeof(float) * 3 * CF_T_MAX));
    CU(cudaMemcpyAsync(s->d_colors, colors, sizeof(float) * 3 * (size_t)std::min(num_colors, CF_T_MAX), cudaMemcpyHostToDevice,
                       s->stream));
    LAUNCH(s, graph_vertices_kernel, div_up(s->n_edges, 256), 256, 0, s->edge_slots, s->n_edges, s->pos[s->cur], s->d_colors,
           s->T, out);
    CU(cudaGetLastError());
    if (device_ptr) *device_ptr = out;
    return CF_OK;
}

extern "C" int cf_download_graph_vertices(cf_sim* s, const cf_color* colors, int num_colors, float* vertices,
                                          int capacity_edges) {
    ARG(s);
    const float* dev = nullptr;
    if (int rc = cf_graph_vertices_device(s, colors, num_colors, nullptr, 0, &dev)) return rc;
    if (s->n_edges == 0) return CF_OK;
    ARG(vertices && capacity_edges >= s->n_edges);
    CU(cudaMemcpyAsync(vertices, dev, sizeof(float) * 12 * (size_t)s->n_edges, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return CF_OK;
}

// ---------------------------------------------------------------------------------------------
// particle snapshot (no reference counterpart: savePreset persists parameters only,
// CellFlowWidget.cpp:1182-1269; the snapshot is what makes a run — and a parity failure — reproducible)
// ---------------------------------------------------------------------------------------------
struct SnapshotHeader {
    char magic[8]; // "CFSNAP02"
    uint32_t header_bytes;
    int32_t n, T;
    cf_params params;
    float raw[CF_TT_MAX], radio[CF_T_MAX], force[CF_TT_MAX];
    int32_t force_overridden;
    int32_t rank, world;
    int32_t reserved[5];
};

extern "C" int cf_save_snapshot(cf_sim* s, const char* path) {
    ARG(s && path);
    if (int rc = set_device(s)) return rc;
    if (int rc = slab_refresh(s)) return rc;
    const int n = s->n;
    std::vector<cf_particle> aos((size_t)std::max(n, 1));
    std::vector<int32_t> counts((size_t)std::max(n, 1)), ids((size_t)std::max(n, 1));
    int got = 0;
    // slot order + ids: a restored run sorts the same input order, so it continues bit for bit
    if (int rc = cf_download_particles_ids(s, aos.data(), counts.data(), ids.data(), std::max(n, 1), &got)) return rc;
    SnapshotHeader h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, "CFSNAP02", 8);
    h.header_bytes = (uint32_t)sizeof(h);
    h.n = got;
    h.T = s->T;
    h.params = s->params;
    memcpy(h.raw, s->raw, sizeof(h.raw));
    memcpy(h.radio, s->radio, sizeof(h.radio));
    memcpy(h.force, s->force, sizeof(h.force));
    h.force_overridden = s->force_overridden ? 1 : 0;
    h.rank = s->rank, h.world = s->world;
    FILE* f = fopen(path, "wb");
    if (!f) return fail(CF_ERR_IO, "cannot write snapshot %s", path);
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
    if (got > 0) {
        ok = ok && fwrite(aos.data(), sizeof(cf_particle), (size_t)got, f) == (size_t)got;
        ok = ok && fwrite(counts.data(), sizeof(int32_t), (size_t)got, f) == (size_t)got;
        ok = ok && fwrite(ids.data(), sizeof(int32_t), (size_t)got, f) == (size_t)got;
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) return fail(CF_ERR_IO, "short write to snapshot %s", path);
    return CF_OK;
}

extern "C" int cf_load_snapshot(cf_sim* s, const char* path) {
    ARG(s && path);
    if (int rc = set_device(s)) return rc;
    FILE* f = fopen(path, "rb");
    if (!f) return fail(CF_ERR_IO, "cannot read snapshot %s", path);
    SnapshotHeader h;
    bool ok = fread(&h, sizeof(h), 1, f) == 1 && memcmp(h.magic, "CFSNAP02", 8) == 0 && h.header_bytes == sizeof(h) &&
              h.n >= 0 && h.T >= 1 && h.T <= CF_MAX_PARTICLE_TYPES;
    std::vector<cf_particle> aos;
    std::vector<int32_t> counts, ids;
    if (ok && h.n > 0) {
        aos.resize((size_t)h.n), counts.resize((size_t)h.n), ids.resize((size_t)h.n);
        ok = fread(aos.data(), sizeof(cf_particle), (size_t)h.n, f) == (size_t)h.n &&
             fread(counts.data(), sizeof(int32_t), (size_t)h.n, f) == (size_t)h.n &&
             fread(ids.data(), sizeof(int32_t), (size_t)h.n, f) == (size_t)h.n;
    }
    fclose(f);
    if (!ok) return fail(CF_ERR_IO, "%s is not a cellflow_b200 snapshot (or is truncated)", path);
    if (s->slab && (h.rank != s->rank || h.world != s->world))
        return fail(CF_ERR_ARG, "snapshot of rank %d/%d loaded into rank %d/%d", h.rank, h.world, s->rank, s->world);
    s->T = h.T;
    s->params = h.params;
    s->params.numParticleTypes = h.T;
    memcpy(s->raw, h.raw, sizeof(h.raw));
    memcpy(s->radio, h.radio, sizeof(h.radio));
    memcpy(s->force, h.force, sizeof(h.force));
    s->force_overridden = h.force_overridden != 0;
    if (int rc = upload_impl(s, aos.data(), counts.data(), ids.data(), h.n)) return rc;
    CU(cudaStreamSynchronize(s->stream)); // the staging vectors go out of scope
    return CF_OK;
}

// ---------------------------------------------------------------------------------------------
// presets applied to a handle (CellFlowWidget::loadPreset order, CellFlowWidget.cpp:1079-1177)
// ---------------------------------------------------------------------------------------------
extern "C" int cf_apply_preset(cf_sim* s, const cf_preset* pr) {
    ARG(s && pr);
    if (pr->particleCount > 0 && pr->particleCount != s->n)
        if (int rc = cf_set_particle_count(s, pr->particleCount)) return rc;
    int T = pr->params.numParticleTypes;
    ARG(T >= 1 && T <= CF_MAX_PARTICLE_TYPES);
    if (T != s->T)
        if (int rc = cf_set_num_particle_types(s, T)) return rc;
    s->params = pr->params;
    s->params.numParticleTypes = T;
    if (pr->numRadio > 0) cf_set_radio_by_type(s, pr->radioByType, pr->numRadio);
    if (pr->numRawForce > 0) cf_set_raw_force_table(s, pr->rawForceTable, pr->numRawForce);
    return cf_update_force_table(s, pr->params.forceRange, pr->params.forceBias, pr->params.forceOffset);
}

// ---------------------------------------------------------------------------------------------
// introspection
// ---------------------------------------------------------------------------------------------
// bench_util.cu enqueues its L2 flush on the handle's stream (not part of the public header)
extern "C" int cf_internal_stream(cf_sim* s, cudaStream_t* stream, int* device) {
    ARG(s && stream && device);
    *stream = s->stream;
    *device = s->device;
    return CF_OK;
}

extern "C" int cf_set_option(cf_sim* s, const char* name, double value) {
    ARG(s && name);
    std::string k(name);
    if (k == "force_kernel") s->opt_force_kernel = (int)value;
    else if (k == "graph_kernel") s->opt_graph_kernel = (int)value; // 0 auto, 1 thread per particle, 2 warp per particle
    else if (k == "timing") s->opt_timing = (int)value;
    else if (k == "t4_ctas_per_sm") s->opt_t4_ctas = (int)value;
    else if (k == "t4_stage") s->opt_t4_stage = (int)value;
    else if (k == "count_blocks") s->opt_count_blocks = (int)value, s->block_counts_valid = false;
    else if (k == "max_cells_per_particle") s->opt_max_cells_per_particle = value;
    else if (k == "cuda_graphs") s->opt_graphs = (int)value;
    else if (k == "global_particle_count") s->n_total = (long long)value; // same value on every rank
    else if (k == "halo_capacity") s->cap_halo = (int)value;       // before cf_comm_init, same on every rank
    else if (k == "migrant_capacity") s->cap_mig = (int)value;     // before cf_comm_init, same on every rank
    else if (k == "wait_timeout_ms") s->wait_timeout_ms = value;
    else if (k == "slab_min_layer_width") s->opt_min_layer_width = value;
    else return fail(CF_ERR_ARG, "unknown option '%s'", name);
    s->sorted_valid = false, s->state_gen++;
    return CF_OK;
}

extern "C" int cf_stats_reset(cf_sim* s) {
    ARG(s);
    if (int rc = set_device(s)) return rc;
    CU(cudaStreamSynchronize(s->stream));
    s->ev_used = 0;
    s->ms_sort = s->ms_force = s->ms_integrate = s->ms_total = s->ms_graph = s->ms_exchange = 0;
    s->ms_exchange_mig = s->ms_exchange_halo = s->ms_exchange_max = s->ms_step_max = 0;
    s->ms_graph_total = 0;
    s->graph_builds = 0;
    s->stat_steps = 0;
    s->launches = 0;
    s->gev_used = 0;
    return CF_OK;
}

extern "C" int cf_get_stats(cf_sim* s, cf_stats* st) {
    ARG(s && st);
    if (int rc = set_device(s)) return rc;
    CU(cudaStreamSynchronize(s->stream));
    const int slab_rc = slab_refresh(s); // slab mode: owned count + error word (reported after the numbers are filled in)
    for (size_t i = 0; i < s->ev_used; i++) {
        float a = 0, b = 0, c = 0;
        StepEvents& ev = s->ev_pool[i];
        cudaEventElapsedTime(&a, ev.e[0], ev.e[1]);
        cudaEventElapsedTime(&b, ev.e[1], ev.e[2]);
        cudaEventElapsedTime(&c, ev.e[2], ev.e[3]);
        s->ms_sort += a;
        s->ms_force += b;
        s->ms_integrate += c;
        if (ev.has_exchange) { // the two mailbox exchanges: wait + append arrivals, halo pack + wait + ghosts
            float x = 0, y = 0;
            if (cudaEventElapsedTime(&x, ev.e[4], ev.e[5]) == cudaSuccess &&
                cudaEventElapsedTime(&y, ev.e[6], ev.e[7]) == cudaSuccess) {
                s->ms_exchange += x + y;
                s->ms_exchange_mig += x;
                s->ms_exchange_halo += y;
                s->ms_exchange_max = std::max(s->ms_exchange_max, (double)(x + y));
            }
        }
        s->ms_total += a + b + c;
        s->ms_step_max = std::max(s->ms_step_max, (double)(a + b + c));
        s->stat_steps++;
    }
    s->ev_used = 0;
    fold_graph_timing(s);
    memset(st, 0, sizeof(*st));
    st->ms_total = s->ms_total;
    st->ms_sort = s->ms_sort;
    st->ms_force = s->ms_force;
    st->ms_integrate = s->ms_integrate;
    st->ms_graph = s->ms_graph;
    st->ms_exchange = s->ms_exchange;
    st->ms_exchange_migrants = s->ms_exchange_mig;
    st->ms_exchange_halo = s->ms_exchange_halo;
    st->ms_exchange_max = s->ms_exchange_max;
    st->ms_step_max = s->ms_step_max;
    st->steps = s->stat_steps;
    st->launches = s->launches;
    for (int a = 0; a < 3; a++) st->grid[a] = s->sc.dims[a];
    st->stencil = 1;
    st->n_owned = s->n;
    st->n_ghost = 0;
    st->force_kernel = s->last_force_kernel;
    st->graph_kernel = s->last_graph_kernel;
    st->ms_graph_total = s->ms_graph_total;
    st->graph_builds = s->graph_builds;
    if (s->slab) st->n_ghost = s->h_slab[SLAB_GHOST_L] + s->h_slab[SLAB_GHOST_R];
    if (s->block_counts_valid) { // (layer of 32 i, quad of 4 j) blocks of the last instrumented force pass
        unsigned long long bc[2] = {0, 0};
        CU(cudaMemcpyAsync(bc, s->d_block_counts, sizeof(bc), cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        st->exact_tested_pairs = (long long)bc[0] * 128;
        st->evaluated_pair_lanes = (long long)bc[1] * 128;
    }
    if (s->n > 0) {
        unsigned long long acc = 0;
        CU(cudaMemsetAsync(s->d_accum, 0, sizeof(unsigned long long), s->stream));
        sum_counts_kernel<<<s->sm_count * 4, 256, 0, s->stream>>>(ofrc(s), s->n, s->d_accum);
        CU(cudaMemcpyAsync(&acc, s->d_accum, sizeof(acc), cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        st->accepted_pairs = (long long)acc;
        if (s->sorted_valid || s->cell_start) {
            // tests executed by the stencil of the last force pass: sum over cells of
            // n_cell * (particles in its neighbour cells); valid while cell_start is current
            CU(cudaMemsetAsync(s->d_accum, 0, sizeof(unsigned long long), s->stream));
            count_tests_kernel<<<div_up(s->ncell, 128), 128, 0, s->stream>>>(s->cell_start, s->ncell, s->sc, s->d_accum);
            CU(cudaMemcpyAsync(&acc, s->d_accum, sizeof(acc), cudaMemcpyDeviceToHost, s->stream));
            CU(cudaStreamSynchronize(s->stream));
            st->tested_pairs = (long long)acc;
        }
    }
    if (slab_rc) return slab_rc;
    return CF_OK;
}

extern "C" int cf_download_cell_keys(cf_sim* s, uint32_t* keys, int32_t* ids, int capacity, int* count) {
    ARG(s && count);
    *count = s->n;
    ARG(capacity >= s->n);
    if (int rc = set_device(s)) return rc;
    if (int rc = prepare_step_const(s)) return rc;
    if (s->n == 0 && !s->slab) return CF_OK;
    if (int rc = build_cell_list(s)) return rc;
    *count = s->n;
    ARG(capacity >= s->n);
    if (keys) CU(cudaMemcpyAsync(keys, s->keys[0], sizeof(uint32_t) * (size_t)s->n, cudaMemcpyDeviceToHost, s->stream));
    if (ids) CU(cudaMemcpyAsync(ids, oid(s), sizeof(int) * (size_t)s->n, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return CF_OK;
}

// ---------------------------------------------------------------------------------------------
// multi-GPU slabs: implemented in comm.cu
// ---------------------------------------------------------------------------------------------
