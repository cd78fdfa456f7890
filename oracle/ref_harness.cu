/*
 * ref_harness.cu — runs the reference's OWN kernels and host methods on caller-supplied data.
 * TEST / BASELINE INFRASTRUCTURE ONLY (built into oracle/_ref/, never shipped, never imported by
 * cellflow_b200).
 *
 * The reference translation unit is compiled from where it lies:
 *     #include "ParticleSimulation.cu"   ->  /root/reference/cuda-native/src/ParticleSimulation.cu
 * (the Makefile passes -I$(REF)/src -I$(REF)/include and a 2-line GL/gl.h typedef stub, because
 * the image has no OpenGL headers; the GL-interop methods are never called).  Nothing of the
 * reference is copied into this repository.
 *
 * Why a harness: the reference class seeds cuRAND with time(nullptr) and keeps d_particles
 * private (ParticleSimulation.cuh:85), so its public API can neither take an input nor repeat
 * one.  Its kernels are ordinary __global__ functions, so they are launched here directly on
 * buffers holding the caller's particles.
 *
 * Race-free evaluation: simulateParticlesKernel updates particles[] in place while other
 * threads still read it (ParticleSimulation.cu:89 vs :164).  ref_simulate_exact() launches it
 * as ONE warp (<<<1,32>>>) over a rotated copy of the input, 32 particles at a time: a single
 * warp reconverges every loop iteration (BSSY/BSYNC in the SASS), so all its loads precede
 * its stores and every particle sees exactly the state of step t.  That is the Jacobi reading
 * the oracle restates, evaluated by the reference's own machine code.
 */
#define private public /* reach d_forceTable / h_rawForceTable of the reference class in tests */
#include "ParticleSimulation.cu"
#undef private

#include <cstdint>
#include <vector>

#include "../include/cellflow_b200.h"

static_assert(sizeof(Particle) == sizeof(cf_particle), "Particle layout");

#define REF_CHECK(call)                                   \
    do {                                                  \
        cudaError_t e_ = (call);                          \
        if (e_ != cudaSuccess) {                          \
            fprintf(stderr, "ref_harness: %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return -2;                                    \
        }                                                 \
    } while (0)

static SimulationParams to_ref(const cf_params* p) {
    SimulationParams s;
    s.radius = p->radius;
    s.delta_t = p->delta_t;
    s.friction = p->friction;
    s.repulsion = p->repulsion;
    s.attraction = p->attraction;
    s.k = p->k;
    s.balance = p->balance;
    s.canvasWidth = p->canvasWidth;
    s.canvasHeight = p->canvasHeight;
    s.canvasDepth = p->canvasDepth;
    s.spawnRegionSize = p->spawnRegionSize;
    s.numParticleTypes = p->numParticleTypes;
    s.ratioWithLFO = p->ratioWithLFO;
    s.forceMultiplier = p->forceMultiplier;
    s.maxExpectedNeighbors = p->maxExpectedNeighbors;
    s.forceRange = p->forceRange;
    s.forceBias = p->forceBias;
    s.ratio = p->ratio;
    s.lfoA = p->lfoA;
    s.lfoS = p->lfoS;
    s.forceOffset = p->forceOffset;
    return s;
}

struct DevBufs {
    Particle* particles = nullptr;
    Particle* scratch = nullptr;
    float* table = nullptr;
    float* radio = nullptr;
    int* cntIn = nullptr;
    int* cntOut = nullptr;
    ~DevBufs() {
        cudaFree(particles);
        cudaFree(scratch);
        cudaFree(table);
        cudaFree(radio);
        cudaFree(cntIn);
        cudaFree(cntOut);
    }
};

static int upload(DevBufs& d, const cf_particle* in, const int32_t* cnt, int n, const cf_params* p,
                  const float* table, const float* radio) {
    int T = p->numParticleTypes;
    REF_CHECK(cudaMalloc(&d.particles, sizeof(Particle) * (size_t)n));
    REF_CHECK(cudaMalloc(&d.scratch, sizeof(Particle) * (size_t)n));
    REF_CHECK(cudaMalloc(&d.table, sizeof(float) * MAX_PARTICLE_TYPES * MAX_PARTICLE_TYPES));
    REF_CHECK(cudaMalloc(&d.radio, sizeof(float) * MAX_PARTICLE_TYPES));
    REF_CHECK(cudaMalloc(&d.cntIn, sizeof(int) * (size_t)n));
    REF_CHECK(cudaMalloc(&d.cntOut, sizeof(int) * ((size_t)n + 32))); /* +32: tail warp */
    REF_CHECK(cudaMemcpy(d.particles, in, sizeof(Particle) * (size_t)n, cudaMemcpyHostToDevice));
    REF_CHECK(cudaMemcpy(d.table, table, sizeof(float) * T * T, cudaMemcpyHostToDevice));
    REF_CHECK(cudaMemcpy(d.radio, radio, sizeof(float) * T, cudaMemcpyHostToDevice));
    if (cnt)
        REF_CHECK(cudaMemcpy(d.cntIn, cnt, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice));
    else
        REF_CHECK(cudaMemset(d.cntIn, 0, sizeof(int) * (size_t)n));
    REF_CHECK(cudaMemset(d.cntOut, 0, sizeof(int) * (size_t)n));
    return 0;
}

extern "C" {

/* One race-free step through the reference kernel (see header comment). */
int ref_simulate_exact(const cf_particle* in, const int32_t* cnt_in, int n, const cf_params* p,
                       const float* table, const float* radio, cf_particle* out,
                       int32_t* cnt_out) {
    DevBufs d;
    if (int rc = upload(d, in, cnt_in, n, p, table, radio)) return rc;
    SimulationParams sp = to_ref(p);
    std::vector<int32_t> zero(n, 0);
    int* cntRot = nullptr;
    REF_CHECK(cudaMalloc(&cntRot, sizeof(int) * (size_t)n));
    for (int b = 0; b < n; b += 32) {
        int head = n - b; /* scratch[k] = particles[(k + b) % n] */
        REF_CHECK(cudaMemcpy(d.scratch, d.particles + b, sizeof(Particle) * (size_t)head,
                             cudaMemcpyDeviceToDevice));
        if (b)
            REF_CHECK(cudaMemcpy(d.scratch + head, d.particles, sizeof(Particle) * (size_t)b,
                                 cudaMemcpyDeviceToDevice));
        REF_CHECK(cudaMemcpy(cntRot, d.cntIn + b, sizeof(int) * (size_t)head,
                             cudaMemcpyDeviceToDevice));
        simulateParticlesKernel<<<1, 32>>>(d.scratch, d.table, d.radio, cntRot, d.cntOut + b, sp, n);
        REF_CHECK(cudaGetLastError());
        int m = n - b < 32 ? n - b : 32;
        /* only threads 0..31 wrote; threads with idx >= n never exist because m <= head */
        REF_CHECK(cudaMemcpy(out + b, d.scratch, sizeof(Particle) * (size_t)m,
                             cudaMemcpyDeviceToHost));
    }
    REF_CHECK(cudaDeviceSynchronize());
    /* cntOut + b was indexed by thread idx 0..31 -> already in original order */
    REF_CHECK(cudaMemcpy(cnt_out, d.cntOut, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost));
    cudaFree(cntRot);
    return 0;
}

/* `steps` steps exactly as ParticleSimulation::simulate launches them (.cu:541-556): full grid,
 * in place, ping-pong counts.  Timed with CUDA events around the kernels only.  The result is
 * subject to the reference's own race. */
int ref_simulate_racy(const cf_particle* in, const int32_t* cnt_in, int n, const cf_params* p,
                      const float* table, const float* radio, int warmup, int steps,
                      cf_particle* out, int32_t* cnt_out, float* ms_per_step) {
    DevBufs d;
    if (int rc = upload(d, in, cnt_in, n, p, table, radio)) return rc;
    SimulationParams sp = to_ref(p);
    int blocks = (n + BLOCK_SIZE - 1) / BLOCK_SIZE;
    int* a = d.cntIn;
    int* b = d.cntOut;
    cudaEvent_t e0, e1;
    REF_CHECK(cudaEventCreate(&e0));
    REF_CHECK(cudaEventCreate(&e1));
    for (int s = 0; s < warmup + steps; s++) {
        if (s == warmup) REF_CHECK(cudaEventRecord(e0));
        simulateParticlesKernel<<<blocks, BLOCK_SIZE>>>(d.particles, d.table, d.radio, a, b, sp, n);
        int* t = a;
        a = b;
        b = t;
    }
    REF_CHECK(cudaEventRecord(e1));
    REF_CHECK(cudaDeviceSynchronize());
    REF_CHECK(cudaGetLastError());
    float ms = 0.f;
    REF_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms_per_step) *ms_per_step = steps > 0 ? ms / steps : 0.f;
    if (out)
        REF_CHECK(cudaMemcpy(out, d.particles, sizeof(Particle) * (size_t)n, cudaMemcpyDeviceToHost));
    if (cnt_out) REF_CHECK(cudaMemcpy(cnt_out, a, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}

/* generateProximityGraphKernel (.cu:188-277) into a plain device buffer instead of a mapped GL
 * VBO.  vertices: capacity_vertices * 6 floats.  Read-only on particles -> deterministic SET
 * of edges, nondeterministic order. */
int ref_graph(const cf_particle* in, int n, int num_types, float proximity_distance, int max_conn,
              const cf_color* colors, int num_colors, float* vertices, int capacity_vertices,
              int* vertex_count, float* ms) {
    Particle* dp = nullptr;
    ParticleColor* dc = nullptr;
    float* dv = nullptr;
    int* dn = nullptr;
    REF_CHECK(cudaMalloc(&dp, sizeof(Particle) * (size_t)n));
    REF_CHECK(cudaMalloc(&dc, sizeof(ParticleColor) * (size_t)num_colors));
    REF_CHECK(cudaMalloc(&dv, sizeof(float) * 6 * (size_t)capacity_vertices));
    REF_CHECK(cudaMalloc(&dn, sizeof(int)));
    REF_CHECK(cudaMemcpy(dp, in, sizeof(Particle) * (size_t)n, cudaMemcpyHostToDevice));
    REF_CHECK(cudaMemcpy(dc, colors, sizeof(ParticleColor) * (size_t)num_colors, cudaMemcpyHostToDevice));
    REF_CHECK(cudaMemset(dn, 0, sizeof(int)));
    cudaEvent_t e0, e1;
    REF_CHECK(cudaEventCreate(&e0));
    REF_CHECK(cudaEventCreate(&e1));
    int grid = (n + 255) / 256;
    REF_CHECK(cudaEventRecord(e0));
    generateProximityGraphKernel<<<grid, 256>>>(dp, n, proximity_distance * proximity_distance,
                                                max_conn, dc, num_types, dv, dn);
    REF_CHECK(cudaEventRecord(e1));
    REF_CHECK(cudaDeviceSynchronize());
    REF_CHECK(cudaGetLastError());
    float t = 0.f;
    REF_CHECK(cudaEventElapsedTime(&t, e0, e1));
    if (ms) *ms = t;
    REF_CHECK(cudaMemcpy(vertex_count, dn, sizeof(int), cudaMemcpyDeviceToHost));
    int nv = *vertex_count < capacity_vertices ? *vertex_count : capacity_vertices;
    REF_CHECK(cudaMemcpy(vertices, dv, sizeof(float) * 6 * (size_t)nv, cudaMemcpyDeviceToHost));
    cudaFree(dp);
    cudaFree(dc);
    cudaFree(dv);
    cudaFree(dn);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}

/* The reference class's own table code: construct -> (optionally) overwrite raw table ->
 * updateForceTable -> read the device table back (.cu:513-539). */
int ref_tables(int num_types, const float* raw_or_null, float range, float bias, float offset,
               float* raw_out, float* radio_out, float* effective_out) {
    srand(1); /* the state an unseeded process starts in (the reference never calls srand) */
    ParticleSimulation sim(32);
    if (num_types != 6) sim.setNumParticleTypes(num_types);
    int T = num_types;
    if (raw_or_null) {
        float* raw = sim.getRawForceTableValues();
        for (int i = 0; i < T * T; i++) raw[i] = raw_or_null[i];
        sim.updateForceTable(range, bias, offset);
    }
    for (int i = 0; i < T * T; i++) raw_out[i] = sim.getRawForceTableValues()[i];
    std::vector<float> r = sim.getRadioByType();
    for (int i = 0; i < T; i++) radio_out[i] = r[i];
    REF_CHECK(cudaMemcpy(effective_out, sim.d_forceTable, sizeof(float) * T * T, cudaMemcpyDeviceToHost));
    return 0;
}

/* moveParticlesKernel (.cu:169-185). */
int ref_move(cf_particle* inout, int n, float dx, float dy, float dz, float W, float H, float D) {
    Particle* dp = nullptr;
    REF_CHECK(cudaMalloc(&dp, sizeof(Particle) * (size_t)n));
    REF_CHECK(cudaMemcpy(dp, inout, sizeof(Particle) * (size_t)n, cudaMemcpyHostToDevice));
    moveParticlesKernel<<<(n + 255) / 256, 256>>>(dp, n, dx, dy, dz, W, H, D);
    REF_CHECK(cudaDeviceSynchronize());
    REF_CHECK(cudaMemcpy(inout, dp, sizeof(Particle) * (size_t)n, cudaMemcpyDeviceToHost));
    cudaFree(dp);
    return 0;
}

int ref_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

} /* extern "C" */
