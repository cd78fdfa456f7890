// cellflow_headless — the reference's step loop without Qt: load a preset the way
// CellFlowWidget::loadPreset does (CellFlowWidget.cpp:1070-1180), then run the paintGL loop
// body (LFO -> simulate -> [proximity graph], CellFlowWidget.cpp:410-427, 552-575) headless.
// Everything on the timed path goes through the C ABI (include/cellflow_b200.h).
//
//   cellflow_headless --preset presets/eater.json [--n 1000000] [--steps 100] [--init uniform|spawn]
//                     [--seed S] [--graph DIST MAXCONN] [--fps-dt SECONDS] [--save out.json]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "cellflow_b200.h"

static void die(const char* what) {
    std::fprintf(stderr, "cellflow_headless: %s: %s\n", what, cf_last_error());
    std::exit(1);
}
#define OK(call) do { if ((call) != 0) die(#call); } while (0)

int main(int argc, char** argv) {
    std::string preset_path, save_path;
    int n_override = -1, steps = 100, mode = CF_INIT_SPAWN_CUBE, max_conn = 0;
    unsigned long long seed = 0x5EED0000ull;
    float graph_dist = 0.f, frame_dt = 1.0f / 60.0f;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&](const char* name) -> const char* {
            if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", name); std::exit(2); }
            return argv[++i];
        };
        if (a == "--preset") preset_path = next("--preset");
        else if (a == "--n") n_override = std::atoi(next("--n"));
        else if (a == "--steps") steps = std::atoi(next("--steps"));
        else if (a == "--seed") seed = std::strtoull(next("--seed"), nullptr, 0);
        else if (a == "--init") mode = std::strcmp(next("--init"), "uniform") == 0 ? CF_INIT_UNIFORM : CF_INIT_SPAWN_CUBE;
        else if (a == "--graph") { graph_dist = (float)std::atof(next("--graph")); max_conn = std::atoi(next("--graph")); }
        else if (a == "--fps-dt") frame_dt = (float)std::atof(next("--fps-dt"));
        else if (a == "--save") save_path = next("--save");
        else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    cf_preset preset;
    cf_default_preset(&preset);
    if (!preset_path.empty() && cf_load_preset(preset_path.c_str(), &preset) != 0) {
        std::fprintf(stderr, "cannot load preset %s\n", preset_path.c_str());
        return 1;
    }
    if (n_override >= 0) preset.particleCount = n_override;

    cf_sim* sim = nullptr;
    OK(cf_create(preset.particleCount, preset.params.numParticleTypes, 0, &sim));
    OK(cf_apply_preset(sim, &preset));         // tables + params, loadPreset order
    cf_params params;
    OK(cf_get_params(sim, &params));
    OK(cf_init_particles(sim, seed, mode));
    OK(cf_set_option(sim, "timing", 1));

    long long edges_total = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (int s = 0; s < steps; s++) {
        // paintGL: the LFO runs on wall-clock time there; headless it runs on a fixed frame time
        params.ratioWithLFO = cf_ratio_with_lfo(&params, (float)s * frame_dt);
        OK(cf_step(sim, &params, 1));
        if (max_conn > 0) {
            int ne = 0;
            OK(cf_build_graph(sim, graph_dist, max_conn, &ne));
            edges_total += ne;
        }
    }
    OK(cf_sync(sim));
    double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    cf_stats st;
    OK(cf_get_stats(sim, &st));
    std::printf("{\"particles\": %d, \"steps\": %d, \"wall_s\": %.6f, \"particle_steps_per_s\": %.1f, "
                "\"ms_cell_list\": %.4f, \"ms_force\": %.4f, \"ms_integrate\": %.4f, \"mean_neighbours\": %.2f, "
                "\"edges_per_step\": %.1f, \"grid\": [%d, %d, %d], \"launches\": %lld}\n",
                preset.particleCount, steps, wall, (double)preset.particleCount * steps / wall,
                st.ms_sort / (st.steps ? st.steps : 1), st.ms_force / (st.steps ? st.steps : 1),
                st.ms_integrate / (st.steps ? st.steps : 1),
                preset.particleCount ? (double)st.accepted_pairs / preset.particleCount : 0.0,
                steps ? (double)edges_total / steps : 0.0, st.grid[0], st.grid[1], st.grid[2], (long long)st.launches);
    if (!save_path.empty()) {
        preset.params = params;
        OK(cf_get_radio_by_type(sim, preset.radioByType, CF_MAX_PARTICLE_TYPES));
        preset.numRadio = params.numParticleTypes;
        OK(cf_get_raw_force_table(sim, preset.rawForceTable, CF_MAX_PARTICLE_TYPES * CF_MAX_PARTICLE_TYPES));
        preset.numRawForce = params.numParticleTypes * params.numParticleTypes;
        if (cf_save_preset(save_path.c_str(), &preset) != 0) std::fprintf(stderr, "cannot save %s\n", save_path.c_str());
    }
    cf_destroy(sim);
    return 0;
}
