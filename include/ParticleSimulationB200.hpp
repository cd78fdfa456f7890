// ParticleSimulationB200.hpp — header-only C++ shim over the C ABI (cellflow_b200.h) with the
// public interface of the reference's host class, so that code written against
// cuda-native/include/ParticleSimulation.cuh:10-75 (its only caller is CellFlowWidget.cpp) keeps
// compiling when it includes this header instead and links libcellflow_b200.so.
//
// Differences a maintainer should know (also in INTEGRATION.md):
//   * errors throw std::runtime_error instead of exit(1) (reference: CUDA_CHECK, .cu:10-18);
//   * getRawForceTableValues() returns a pointer into a host mirror owned by this object; the
//     mirror is pushed to the engine by updateForceTable(), which is the only way the reference's
//     caller ever commits its writes (CellFlowWidget.cpp:1170-1177);
//   * generateProximityGraph(vbo, ...) needs an uploader callback (CUDA-GL interop stays in the Qt
//     adapter, outside the timed path); generateProximityGraph(std::vector<float>&, ...) returns
//     the same vertex stream on the host.  generateTriangleMesh is dead code in the reference
//     (no caller) and is not provided.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "cellflow_b200.h"

namespace cellflow_b200 {

struct Float3 {
    float x, y, z;
};

// Layout-compatible with the reference Particle (SimulationParams.h:6-12) and with cf_particle.
struct Particle {
    Float3 pos, vel, acc;
    unsigned int ptype;
    float pad;
};
static_assert(sizeof(Particle) == sizeof(cf_particle), "Particle must stay 44 bytes");

// Same fields, order and defaults as the reference SimulationParams (SimulationParams.h:14-57).
struct SimulationParams {
    float radius = 42.07f;
    float delta_t = 0.18f;
    float friction = 0.51f;
    float repulsion = 64.83f;
    float attraction = 3.06f;
    float k = 29.45f;
    float balance = 0.79f;
    float canvasWidth = 8000.0f;
    float canvasHeight = 8000.0f;
    float canvasDepth = 8000.0f;
    float spawnRegionSize = 2000.0f;
    int numParticleTypes = 6;
    float ratioWithLFO = 0.0f;
    float forceMultiplier = 2.33f;
    int maxExpectedNeighbors = 400;
    float forceRange = 0.28f;
    float forceBias = -0.20f;
    float ratio = 0.0f;
    float lfoA = 0.0f;
    float lfoS = 0.1f;
    float forceOffset = 1.0f;
    float pointSize = 10.0f;
    float depthFadeStart = 10000.0f;
    float depthFadeEnd = 15000.0f;
    float sizeAttenuationFactor = 1000.0f;
    float brightnessMin = 0.4f;
    float focusDistance = 3000.0f;
    float apertureSize = 0.0f;
    bool enableDepthFade = false;
    bool enableSizeAttenuation = true;
    bool enableBrightnessAttenuation = true;
    bool enableDOF = false;

    cf_params physics() const {
        cf_params p;
        p.radius = radius, p.delta_t = delta_t, p.friction = friction, p.repulsion = repulsion;
        p.attraction = attraction, p.k = k, p.balance = balance;
        p.canvasWidth = canvasWidth, p.canvasHeight = canvasHeight, p.canvasDepth = canvasDepth;
        p.spawnRegionSize = spawnRegionSize, p.numParticleTypes = numParticleTypes;
        p.ratioWithLFO = ratioWithLFO, p.forceMultiplier = forceMultiplier;
        p.maxExpectedNeighbors = maxExpectedNeighbors, p.forceRange = forceRange, p.forceBias = forceBias;
        p.ratio = ratio, p.lfoA = lfoA, p.lfoS = lfoS, p.forceOffset = forceOffset;
        return p;
    }
};

using ParticleColor = cf_color;
constexpr int MAX_PARTICLE_TYPES = CF_MAX_PARTICLE_TYPES;

class ParticleSimulation {
public:
    explicit ParticleSimulation(int particleCount) : ParticleSimulation(particleCount, 8000.0f, 8000.0f) {}
    ParticleSimulation(int particleCount, float canvasWidth, float canvasHeight) {
        check(cf_create(particleCount, 6, 0, &sim_));
        canvasW_ = canvasWidth, canvasH_ = canvasHeight, canvasD_ = canvasHeight;
        pullTables();
        initializeParticles();
    }
    ~ParticleSimulation() { cf_destroy(sim_); }
    ParticleSimulation(const ParticleSimulation&) = delete;
    ParticleSimulation& operator=(const ParticleSimulation&) = delete;

    void initializeParticles() {
        cf_params p;
        check(cf_get_params(sim_, &p));
        p.canvasWidth = canvasW_, p.canvasHeight = canvasH_, p.canvasDepth = canvasD_;
        check(cf_set_params(sim_, &p));
        check(cf_init_particles(sim_, seed_++, CF_INIT_SPAWN_CUBE));
    }
    void initializeParticles(float canvasWidth, float canvasHeight) {
        updateCanvasDimensions(canvasWidth, canvasHeight);
        initializeParticles();
    }
    void updateCanvasDimensions(float canvasWidth, float canvasHeight) {
        canvasW_ = canvasWidth, canvasH_ = canvasHeight, canvasD_ = canvasHeight; // depth = height, .cu:510
    }
    void initializeForceTable() {
        check(cf_regenerate_force_table(sim_));
        pullTables();
    }
    void updateForceTable(float forceRange, float forceBias, float forceOffset) {
        check(cf_set_raw_force_table(sim_, raw_.data(), numTypes() * numTypes()));
        check(cf_update_force_table(sim_, forceRange, forceBias, forceOffset));
    }
    void initializeRadioByType() { pullTables(); }

    void simulate(const SimulationParams& params) {
        cf_params p = params.physics();
        check(cf_step(sim_, &p, 1));
        check(cf_sync(sim_)); // the reference synchronises inside simulate(), .cu:553
    }
    void getParticleData(std::vector<Particle>& particles) {
        particles.resize((size_t)getParticleCount());
        check(cf_download_particles(sim_, reinterpret_cast<cf_particle*>(particles.data()), (int)particles.size()));
    }

    void setParticleCount(int count) { check(cf_set_particle_count(sim_, count)); }
    int getParticleCount() const { return cf_get_particle_count(sim_); }
    void setNumParticleTypes(int types) {
        check(cf_set_num_particle_types(sim_, types));
        pullTables();
    }
    int getNumParticleTypes() const { return numTypes(); }

    void regenerateForceTable() { initializeForceTable(); }
    float* getRawForceTableValues() { return raw_.data(); }

    void moveUniverse(float dx, float dy) { check(cf_move_universe(sim_, dx, dy, 0.0f)); }
    void rotateRadioByType() { check(cf_rotate_radio_by_type(sim_)); }
    std::vector<float> getRadioByType() const {
        std::vector<float> r((size_t)numTypes());
        check(cf_get_radio_by_type(sim_, r.data(), (int)r.size()));
        return r;
    }
    void setRadioByTypeValue(int index, float value) { check(cf_set_radio_by_type_value(sim_, index, value)); }

    // Host-side vertex stream: 12 floats per edge, the reference's VBO layout (.cu:255-275).
    void generateProximityGraph(std::vector<float>& lineVertices, int& outVertexCount, float proximityDistance,
                                int maxConnectionsPerParticle, const std::vector<ParticleColor>& particleColors) {
        int ne = 0;
        check(cf_build_graph(sim_, proximityDistance, maxConnectionsPerParticle, &ne));
        lineVertices.resize((size_t)ne * 12);
        std::vector<ParticleColor> colors(particleColors);
        if ((int)colors.size() < numTypes()) colors.resize((size_t)numTypes(), ParticleColor{1.f, 1.f, 1.f});
        check(cf_download_graph_vertices(sim_, colors.data(), (int)colors.size(), lineVertices.data(), ne));
        outVertexCount = 2 * ne;
    }
    // Zero-copy form: the stream is written on the device straight into `mappedVbo`, the pointer the GL side
    // gets from cudaGraphicsResourceGetMappedPointer for its VBO (what the reference does inside
    // generateProximityGraph, .cu:633-676) — no host round trip.  capacityEdges = VBO size / 48 bytes.
    void generateProximityGraphDevice(float* mappedVbo, int capacityEdges, int& outVertexCount, float proximityDistance,
                                      int maxConnectionsPerParticle, const std::vector<ParticleColor>& particleColors) {
        int ne = 0;
        check(cf_build_graph(sim_, proximityDistance, maxConnectionsPerParticle, &ne));
        std::vector<ParticleColor> colors(particleColors);
        if ((int)colors.size() < numTypes()) colors.resize((size_t)numTypes(), ParticleColor{1.f, 1.f, 1.f});
        check(cf_graph_vertices_device(sim_, colors.data(), (int)colors.size(), mappedVbo, capacityEdges, nullptr));
        check(cf_sync(sim_));
        outVertexCount = 2 * ne;
    }
    // Particle snapshot (beyond the reference, whose savePreset keeps parameters only).
    void saveSnapshot(const std::string& path) { check(cf_save_snapshot(sim_, path.c_str())); }
    void loadSnapshot(const std::string& path) {
        check(cf_load_snapshot(sim_, path.c_str()));
        pullTables();
    }
    // Reference signature (.cuh:60-66).  `vboUploader(vbo, data, bytes)` is supplied by the GL side.
    std::function<void(unsigned int, const float*, size_t)> vboUploader;
    void generateProximityGraph(unsigned int openglVBO, int& outVertexCount, float proximityDistance,
                                int maxConnectionsPerParticle, const std::vector<ParticleColor>& particleColors) {
        if (!vboUploader) throw std::runtime_error("generateProximityGraph(vbo): set vboUploader first");
        std::vector<float> v;
        generateProximityGraph(v, outVertexCount, proximityDistance, maxConnectionsPerParticle, particleColors);
        vboUploader(openglVBO, v.data(), v.size() * sizeof(float));
    }

    cf_sim* handle() const { return sim_; }

private:
    static void check(int rc) {
        if (rc != 0) throw std::runtime_error(std::string("cellflow_b200: ") + cf_last_error());
    }
    int numTypes() const { return cf_get_num_particle_types(sim_); }
    void pullTables() {
        raw_.assign((size_t)MAX_PARTICLE_TYPES * MAX_PARTICLE_TYPES, 0.0f);
        check(cf_get_raw_force_table(sim_, raw_.data(), numTypes() * numTypes()));
    }
    cf_sim* sim_ = nullptr;
    std::vector<float> raw_;
    float canvasW_ = 8000.0f, canvasH_ = 8000.0f, canvasD_ = 8000.0f;
    uint64_t seed_ = 0x5EED0000ull;
};

}  // namespace cellflow_b200

#ifdef CELLFLOW_B200_REFERENCE_NAMES // drop-in: the reference's unqualified names
using cellflow_b200::MAX_PARTICLE_TYPES;
using cellflow_b200::Particle;
using cellflow_b200::ParticleColor;
using cellflow_b200::ParticleSimulation;
using cellflow_b200::SimulationParams;
#endif
