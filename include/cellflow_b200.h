/*
 * cellflow_b200.h — C ABI of the B200-native particle-life engine.
 *
 * This is the drop-in boundary for the reference's `class ParticleSimulation`
 * (reference: cuda-native/include/ParticleSimulation.cuh:10-75, implemented in
 * cuda-native/src/ParticleSimulation.cu:426-687).  The reference links that class statically
 * into its Qt executable; the replacement is a shared library (`libcellflow_b200.so`) with
 * plain-C entry points, plain pointers and sizes, and no CUDA, torch or C++ types in any
 * signature.  `include/ParticleSimulationB200.hpp` wraps these entry points in a C++ class with
 * the reference's method names so code written against ParticleSimulation.cuh keeps compiling.
 *
 * Conventions
 *   - every entry returns 0 on success and a negative cf_status on failure; nothing aborts the
 *     host process (the reference's CUDA_CHECK prints and exit(1)s, ParticleSimulation.cu:10-18);
 *     `cf_last_error()` returns a thread-local message for the last failure.
 *   - the library owns all device memory; the caller owns every host buffer it passes.
 *   - one handle is driven by one host thread at a time.  `cf_step` is asynchronous on the
 *     handle's stream; `cf_sync` and every download are synchronisation points.
 *   - there is no CPU fallback: without a CUDA device `cf_create` fails with CF_ERR_CUDA.
 */
#ifndef CELLFLOW_B200_H
#define CELLFLOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CF_MAX_PARTICLE_TYPES 10 /* reference: SimulationParams.h:64 (MAX_PARTICLE_TYPES) */
#define CF_MAX_GRAPH_CONN 16     /* reference kernel's nearby[32] holds 2*maxConn entries,
                                    ParticleSimulation.cu:210-215; values >16 overflow it there */

typedef enum cf_status {
    CF_OK = 0,
    CF_ERR_ARG = -1,      /* bad argument */
    CF_ERR_CUDA = -2,     /* CUDA runtime error (message in cf_last_error) */
    CF_ERR_STATE = -3,    /* call not valid in the handle's current state */
    CF_ERR_IO = -4,       /* preset file could not be read / parsed */
    CF_ERR_NCCL = -5,     /* (reserved; round 1 used NCCL for the slab exchange) */
    CF_ERR_CAPACITY = -6  /* a fixed-capacity buffer (halo, migrants, edges) overflowed */
} cf_status;

/* Reference `struct Particle`, SimulationParams.h:6-12: 44 bytes, 4-byte aligned.
 * `acc` holds the force applied in the last step (written at ParticleSimulation.cu:146). */
typedef struct cf_particle {
    float pos[3];
    float vel[3];
    float acc[3];
    uint32_t ptype;
    float pad;
} cf_particle;

/* Physics subset of the reference `struct SimulationParams`, SimulationParams.h:15-37.
 * Field names follow the reference.  `ratioWithLFO` is what the step kernel reads
 * (ParticleSimulation.cu:108); `ratio`, `lfoA`, `lfoS` feed it on the host
 * (CellFlowWidget.cpp:415-421).  `forceRange/Bias/Offset` drive cf_update_force_table. */
typedef struct cf_params {
    float radius;
    float delta_t;
    float friction;
    float repulsion;
    float attraction;
    float k;
    float balance;
    float canvasWidth;
    float canvasHeight;
    float canvasDepth;
    float spawnRegionSize;
    int32_t numParticleTypes;
    float ratioWithLFO;
    float forceMultiplier;
    int32_t maxExpectedNeighbors;
    float forceRange;
    float forceBias;
    float ratio;
    float lfoA;
    float lfoS;
    float forceOffset;
} cf_params;

/* Reference `struct ParticleColor`, SimulationParams.h:60-62. */
typedef struct cf_color {
    float r, g, b;
} cf_color;

/* Everything CellFlowWidget::loadPreset reads (CellFlowWidget.cpp:1070-1180).  Render-only
 * keys are carried so a preset round-trips; the engine ignores them. */
typedef struct cf_preset {
    int32_t particleCount; /* "PARTICLE_COUNT" */
    cf_params params;
    float pointSize;
    float depthFadeStart, depthFadeEnd, sizeAttenuationFactor, brightnessMin;
    float focusDistance, apertureSize;
    int32_t enableDepthFade, enableSizeAttenuation, enableBrightnessAttenuation, enableDOF;
    int32_t invertPan, invertForwardBack, invertRotation;
    int32_t effectType;
    int32_t numColors;
    cf_color particleColors[CF_MAX_PARTICLE_TYPES];
    int32_t numRadio;
    float radioByType[CF_MAX_PARTICLE_TYPES];
    int32_t numRawForce;
    float rawForceTable[CF_MAX_PARTICLE_TYPES * CF_MAX_PARTICLE_TYPES];
} cf_preset;

/* One proximity-graph edge: original particle indices, i < j (ParticleSimulation.cu:215). */
typedef struct cf_edge {
    int32_t i, j;
} cf_edge;

/* Device-side timings (CUDA events on the handle's stream) and counters of the last cf_step /
 * cf_build_graph call; used by bench.py for the roofline figures. */
typedef struct cf_stats {
    double ms_total;       /* whole cf_step call, all sub-steps */
    double ms_sort;        /* key + radix sort + reorder + cell bounds */
    double ms_force;       /* pair-force kernel(s) */
    double ms_integrate;   /* fused density/friction/integrate(+next key) kernel */
    double ms_exchange;    /* halo + migration exchange (multi-GPU) */
    double ms_graph;       /* last cf_build_graph */
    int64_t steps;         /* sub-steps covered by the timings above */
    int64_t launches;      /* kernels of this library launched since cf_stats_reset */
    int64_t accepted_pairs;/* sum over owned particles of neighbour counts of the last step */
    int64_t tested_pairs;  /* pair tests executed by the force kernel in the last step */
    int32_t grid[3];       /* cells per axis of the current cell grid */
    int32_t stencil;       /* half-width m of the (2m+1)^3 neighbour stencil */
    int32_t n_owned, n_ghost;
    int32_t force_kernel;  /* pair-force kernel of the last step: 1 per-particle, 2 tile gen. 3, 3 tile gen. 4 */
    int32_t graph_kernel;  /* kernel of the last cf_build_graph: 1 thread per particle, 2 warp per particle */
    double ms_graph_total; /* all cf_build_graph calls since cf_stats_reset (timing != 0) */
    int64_t graph_builds;  /* ... and how many they were */
    double ms_exchange_migrants; /* slab mode, part of ms_exchange: wait for + append the arrivals */
    double ms_exchange_halo;     /* slab mode, part of ms_exchange: halo pack + wait + ghost unpack + ghost bounds */
    int64_t exact_tested_pairs;  /* tile kernel, option "count_blocks": pairs that reached the exact per-pair test */
    int64_t evaluated_pair_lanes;/* tile kernel, option "count_blocks": pair-lanes whose force terms were evaluated */
    double ms_step_max;          /* slowest single step since cf_stats_reset (timing != 0) */
    double ms_exchange_max;      /* slab mode: largest ms_exchange of a single step (a spike = one late neighbour) */
} cf_stats;

typedef struct cf_sim cf_sim;

/* ---- lifetime ------------------------------------------------------------------------- */

/* Replaces ParticleSimulation::ParticleSimulation(int) + allocateMemory
 * (ParticleSimulation.cu:427-435, 451-467).  Particles start zeroed; call cf_init_particles or
 * cf_upload_particles.  Force table and radioByType start at the reference's defaults
 * (glibc rand() sequence, ParticleSimulation.cu:513-519, 533-539). */
int cf_create(int particle_count, int num_types, int device, cf_sim** out);
int cf_destroy(cf_sim* sim); /* ~ParticleSimulation + freeMemory, .cu:447-449, 469-476 */

/* setParticleCount / setNumParticleTypes, .cu:564-579 (re-inits like the reference does). */
int cf_set_particle_count(cf_sim* sim, int count);
int cf_get_particle_count(const cf_sim* sim);
int cf_set_num_particle_types(cf_sim* sim, int types);
int cf_get_num_particle_types(const cf_sim* sim);

/* ---- tables (ParticleSimulation.cu:513-539, 581-622) ------------------------------------ */

int cf_regenerate_force_table(cf_sim* sim); /* regenerateForceTable, .cu:581-583 */
/* Host-only (no device needed): the tables a freshly constructed reference object holds —
 * raw[T*T] and radio[T] from libc rand() in its never-seeded state, effective[T*T] =
 * updateForceTable(0.28, -0.20, 1.0) (.cu:513-519, 533-539).  "Default force matrix" of the
 * README headline configuration. */
int cf_reference_default_tables(int num_types, float* raw, float* radio, float* effective);
int cf_set_raw_force_table(cf_sim* sim, const float* raw, int count);  /* replaces the mutable
                                   pointer returned by getRawForceTableValues(), .cuh:47 */
int cf_get_raw_force_table(const cf_sim* sim, float* raw, int count);
int cf_update_force_table(cf_sim* sim, float forceRange, float forceBias,
                          float forceOffset);                           /* .cu:521-531 */
int cf_get_force_table(const cf_sim* sim, float* effective, int count); /* T*T, [self][other] */
int cf_set_force_table(cf_sim* sim, const float* effective, int count); /* bypass the transform */
int cf_set_radio_by_type(cf_sim* sim, const float* radio, int count);
int cf_set_radio_by_type_value(cf_sim* sim, int index, float value);    /* .cu:615-622 */
int cf_get_radio_by_type(const cf_sim* sim, float* radio, int count);   /* .cu:607-613 */
int cf_rotate_radio_by_type(cf_sim* sim);                               /* .cu:594-605 */

/* ---- particle state ----------------------------------------------------------------------- */

#define CF_INIT_SPAWN_CUBE 0 /* reference rule: centred min(2000,W)^3 cube, .cu:45-62, 489 */
#define CF_INIT_UNIFORM 1    /* uniform over the whole canvas */
/* Replaces initCurandKernel + initializeParticlesKernel (.cu:21-66).  Counter-based generator
 * keyed by (seed, particle id): any rank can regenerate any particle; no RNG state array.
 * Canvas size comes from the params last passed to cf_set_params / cf_step. */
int cf_init_particles(cf_sim* sim, uint64_t seed, int mode);

/* Host <-> device in the reference's 44-byte AoS layout and ORIGINAL particle order
 * (getParticleData, .cu:558-562).  `count` must equal cf_get_particle_count. */
int cf_upload_particles(cf_sim* sim, const cf_particle* aos, int count);
int cf_download_particles(cf_sim* sim, cf_particle* aos, int count);
/* neighborCounts ping-pong buffer of the last step (.cu:544-545, 165), original order. */
int cf_upload_neighbor_counts(cf_sim* sim, const int32_t* counts, int count);
int cf_download_neighbor_counts(cf_sim* sim, int32_t* counts, int count);
/* Render feed: what CellFlowWidget::updateParticleBuffer (CellFlowWidget.cpp:742-761) builds on the
 * CPU from getParticleData every frame — (x, y, z, (float)type) per particle, original order
 * (slot order in slab mode) — plus the per-type counts of getParticleTypeCounts (:875-886).
 * xyzt: capacity*4 floats on the host or NULL; type_counts: numParticleTypes ints or NULL;
 * device_ptr: receives a device pointer to the same float4 stream (zero-copy consumers) or NULL. */
int cf_render_feed(cf_sim* sim, float* xyzt, int capacity, int32_t* type_counts, const void** device_ptr);
/* updateCanvasDimensions / moveUniverse, .cu:507-511, 585-592. */
int cf_move_universe(cf_sim* sim, float dx, float dy, float dz);

/* ---- stepping (simulate, ParticleSimulation.cu:541-556) --------------------------------- */

int cf_set_params(cf_sim* sim, const cf_params* params);
int cf_get_params(const cf_sim* sim, cf_params* params);
/* Runs `n_steps` steps with `params` (NULL: keep the last ones).  Asynchronous. */
int cf_step(cf_sim* sim, const cf_params* params, int n_steps);
int cf_sync(cf_sim* sim);
/* Stateless form of one step for callers that keep particles on the host, as the reference's
 * widget does every frame (simulate + getParticleData, CellFlowWidget.cpp:424-427):
 * H2D(particles, counts) -> step -> D2H(particles, counts).  Blocking. */
int cf_step_host(cf_sim* sim, const cf_params* params, const cf_particle* in,
                 const int32_t* counts_in, cf_particle* out, int32_t* counts_out, int count);
/* LFO of the step driver, CellFlowWidget.cpp:415-421. */
float cf_ratio_with_lfo(const cf_params* params, float t_seconds);

/* ---- proximity graph (generateProximityGraph, ParticleSimulation.cu:188-277, 625-687) ---- */

/* Builds the edge set on the device; *n_edges receives the edge count (vertexCount/2 of the
 * reference).  max_conn is clamped to CF_MAX_GRAPH_CONN. */
int cf_build_graph(cf_sim* sim, float proximity_distance, int max_conn, int* n_edges);
/* n_edges == NULL: asynchronous build (nothing is read back; the step loop of a multi-GPU run never stops for
 * the host).  cf_get_graph_edge_count synchronises and returns the count of the last build. */
int cf_get_graph_edge_count(cf_sim* sim, int* n_edges);
int cf_download_graph_edges(cf_sim* sim, cf_edge* edges, int capacity);
/* Reference VBO layout: per edge 2 vertices x (pos xyz + colour rgb of i's type) = 12 floats
 * (ParticleSimulation.cu:255-275). */
int cf_download_graph_vertices(cf_sim* sim, const cf_color* colors, int num_colors,
                               float* vertices, int capacity_edges);
/* Zero-copy hand-off of the same stream (the reference writes it into a mapped GL VBO, .cu:633-676, consumer
 * CellFlowWidget.cpp:606-617): written on the device into `device_dst` — memory the caller owns, e.g. the
 * pointer cudaGraphicsResourceGetMappedPointer returns for the widget's VBO — or, with device_dst == NULL, into
 * a persistent buffer of the library whose address is returned in *device_ptr (valid until the next call).
 * Asynchronous on the handle's stream; cf_sync() orders it.  CF_ERR_STATE when the particles were moved or
 * reordered since cf_build_graph. */
int cf_graph_vertices_device(cf_sim* sim, const cf_color* colors, int num_colors, float* device_dst,
                             int capacity_edges, const float** device_ptr);

/* ---- presets (CellFlowWidget::loadPreset / savePreset, CellFlowWidget.cpp:1070-1269) ----- */

void cf_default_params(cf_params* params);  /* SimulationParams.h:15-37 defaults */
void cf_default_preset(cf_preset* preset);
int cf_load_preset(const char* path, cf_preset* preset);
int cf_save_preset(const char* path, const cf_preset* preset);
/* Applies a preset the way loadPreset does: count, types, radioByType, rawForceTable, then
 * updateForceTable(forceRange, forceBias, forceOffset) (CellFlowWidget.cpp:1079-1177). */
int cf_apply_preset(cf_sim* sim, const cf_preset* preset);

/* ---- particle snapshot (new: the reference's savePreset persists parameters only, CellFlowWidget.cpp:1182-1269) ---- */

/* Binary snapshot of everything a run continues from: parameters, tables, and per particle pos / vel / acc / type /
 * previous neighbour count / id in the engine's current slot order.  A handle restored with cf_load_snapshot
 * continues bit for bit like the one that wrote the file (same device, same options).  Slab mode: one file per
 * rank. */
int cf_save_snapshot(cf_sim* sim, const char* path);
int cf_load_snapshot(cf_sim* sim, const char* path);

/* ---- multi-GPU slabs (new; no reference counterpart) ---------------- */

/* Rank `rank` of `world` owns x in [bound(rank), bound(rank+1)) — the uniform split rank*W/world unless
 * cf_slab_set_bounds says otherwise.  `capacity` bounds the owned particle count of this rank and must be the
 * same on every rank (the cell grid and the mailbox sizes derive from rank-invariant numbers only: the
 * capacity, and the options "global_particle_count", "halo_capacity", "migrant_capacity").
 * Neighbour exchange is peer-to-peer: every rank owns a mailbox in device memory that its two ring neighbours
 * map through CUDA IPC and write with plain stores over NVLink from inside the producing kernels.  Setup:
 *   cf_comm_init on every rank -> cf_comm_mailbox_handle -> the host exchanges the 64-byte handles (any transport)
 *   -> cf_comm_connect(left neighbour's handle, right neighbour's handle).
 * world == 1 needs no connect (the rank's own mailbox stands in for both neighbours). */
#define CF_IPC_HANDLE_BYTES 64
int cf_comm_init(cf_sim* sim, int rank, int world, int capacity);
int cf_comm_mailbox_handle(cf_sim* sim, void* handle64);
int cf_comm_connect(cf_sim* sim, const void* left_handle64, const void* right_handle64);
/* Slab bounds for clustered states: world + 1 increasing x values, bounds[0] = 0, bounds[world] = canvas width,
 * the same array on every rank; every slab must stay at least one interaction radius wide.  Call it between
 * steps, after the particles have been redistributed to match (cellflow_b200/dist.py: rebalance). */
int cf_slab_set_bounds(cf_sim* sim, const float* bounds, int count);
/* Slab-mode initial condition: every rank generates the same n_total particles (counter-based
 * generator) and keeps those whose x lies in its slab.  Canvas = params last set. */
int cf_init_particles_global(cf_sim* sim, int64_t n_total, uint64_t seed, int mode);
/* Owned x interval [lo, hi) of this rank (valid after cf_comm_init + cf_set_params). */
int cf_slab_bounds(cf_sim* sim, float* lo, float* hi);
/* Upload a subset with explicit global ids (multi-GPU: each rank uploads what it owns). */
int cf_upload_particles_ids(cf_sim* sim, const cf_particle* aos, const int32_t* counts,
                            const int32_t* ids, int count);
int cf_download_particles_ids(cf_sim* sim, cf_particle* aos, int32_t* counts, int32_t* ids,
                              int capacity, int* count);

/* ---- introspection ------------------------------------------------------------------------- */

int cf_get_stats(cf_sim* sim, cf_stats* stats);
int cf_stats_reset(cf_sim* sim);
/* Sorted-order views for the cell-assignment parity tests: sort key (= cell * 64 + Hilbert index of the 4x4x4 sub-cell) and
 * original id per slot. */
int cf_download_cell_keys(cf_sim* sim, uint32_t* keys, int32_t* ids, int capacity, int* count);
/* Tuning knobs; every one defaults to the automatic choice and none changes a result beyond the
 * summation order of the force terms:
 *   "force_kernel"   0 auto | 1 one thread per particle | 2 tile kernel, generation 3 |
 *                    3 tile kernel, generation 4 (box prefilter, type-sorted j stream)
 *   "graph_kernel"   0 auto | 1 one thread per particle | 2 one warp per particle (dense states)
 *   "t4_stage"       staging of the j chunks in the generation-4 tile kernel, all bit-identical: 4 (default) box prefilter
 *                    on the registers, live quads stored compacted | 0 every quad stored | 1 cp.async.bulk + mbarrier from
 *                    SoA planes | 2 precomputed quad boxes | 3 SoA planes through registers (1, 3: per-type radii only)
 *   "t4_ctas_per_sm", "count_blocks"   experiment knobs of the same kernel: fewer resident CTAs; instrumented build that
 *                    fills cf_stats.exact_tested_pairs / evaluated_pair_lanes
 *   "cuda_graphs"    1 (default): single-GPU step and graph build replay captured CUDA graphs
 *   "timing"         0 off | 1 events between the phases of a step | 2 events around whole steps
 *   "max_cells_per_particle"   upper bound of grid cells per particle (default 16)
 *   "global_particle_count", "halo_capacity", "migrant_capacity"
 *                    slab mode, the same on every rank: grid sizing / mailbox capacities (before cf_comm_init)
 *   "wait_timeout_ms"  slab mode: how long a rank waits for a neighbour's message before it reports an error
 *   "slab_min_layer_width"  slab mode, the same on every rank: x layers (and with them the ghost layers) at least
 *                    this wide — set it to the proximity-graph distance when that exceeds the interaction radius
 * Unknown names return CF_ERR_ARG. */
int cf_set_option(cf_sim* sim, const char* name, double value);
/* Measurement helpers for bench.py (not on the simulation path): live FP32 FMA peak of the
 * device in TFLOP/s (roofline denominator of the pair-force kernel) and an L2 flush that
 * overwrites `bytes` of scratch (> L2 size) between timed iterations. */
int cf_bench_fp32_peak(int device, double* tflops, double* sm_mhz_effective);
int cf_bench_flush_l2(int device, size_t bytes);
/* The same flush enqueued on the handle's stream without synchronising (free-running multi-GPU step loops). */
int cf_bench_flush_l2_async(cf_sim* sim, size_t bytes);
const char* cf_last_error(void);
const char* cf_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CELLFLOW_B200_H */
