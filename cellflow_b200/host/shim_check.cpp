// Compile-and-run check of the C++ shim: code written against the reference's class interface
// (ParticleSimulation.cuh:10-75), using the reference's unqualified names.
#define CELLFLOW_B200_REFERENCE_NAMES
#include <cstdio>
#include "ParticleSimulationB200.hpp"

int main() {
    ParticleSimulation sim(4000);                    // CellFlowWidget.cpp:13
    SimulationParams params;
    sim.setNumParticleTypes(6);
    float* raw = sim.getRawForceTableValues();       // CellFlowWidget.cpp:1170-1177
    for (int i = 0; i < 36; i++) raw[i] = (i % 7) * 0.1f - 0.3f;
    sim.updateForceTable(params.forceRange, params.forceBias, params.forceOffset);
    sim.setRadioByTypeValue(2, 0.5f);
    params.radius = 300.f;
    for (int s = 0; s < 5; s++) sim.simulate(params);
    std::vector<Particle> particles;
    sim.getParticleData(particles);
    std::vector<ParticleColor> colors(6, ParticleColor{1.f, 0.5f, 0.25f});
    std::vector<float> verts;
    int nv = 0;
    sim.generateProximityGraph(verts, nv, 200.f, 5, colors);
    std::printf("SHIM_OK particles=%zu first=(%.2f %.2f %.2f) vertices=%d radio2=%.2f\n", particles.size(),
                particles[0].pos.x, particles[0].pos.y, particles[0].pos.z, nv, sim.getRadioByType()[2]);
    return particles.size() == 4000 ? 0 : 1;
}
