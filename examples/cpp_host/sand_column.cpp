// sand_column.cpp — a C++ host driving libsparkl_b200 through include/sparkl_b200.hpp, written like examples3d/sand3.rs: a
// Drucker-Prager sand lattice on a corotated block, cuboid floor, `CudaMpmPipeline::step` per frame, positions read back.
// Build:  g++ -std=c++17 -Iinclude examples/cpp_host/sand_column.cpp -Lsparkl_b200 -lsparkl_b200 -o sand_column
// Usage:  sand_column [ni nj nk frames [positions.bin [heightfield]]]   (prints one summary line; optionally dumps the final positions;
//         `heightfield` replaces the cuboid floor by a heightfield ground)
#include <cstdio>
#include <cstdlib>
#include <string>

#include "sparkl_b200.hpp"

int main(int argc, char** argv) {
    using namespace sparkl;
    const int ni = argc > 1 ? atoi(argv[1]) : 20, nj = argc > 2 ? atoi(argv[2]) : 10, nk = argc > 3 ? atoi(argv[3]) : 10, frames = argc > 4 ? atoi(argv[4]) : 2;
    const float E = 1.0e7f, NU = 0.2f, h = 0.2f, r = h / 4.0f;
    try {
        SingleGpuMpmContext ctx(h);
        std::vector<ParticleModel> models = {ParticleModel::corotated_linear_elasticity(E, NU).with_drucker_prager_plasticity(E, NU),
                                             ParticleModel::corotated_linear_elasticity(E, NU)};
        ctx.models().from_models(models);
        const float y0 = h * 3.0f + 2.0f, ext = 2.0f * r;
        ParticleComponents pc;
        for (int layer = 0; layer < 2; ++layer) // upper lattice: sand (model 0), lower: elastic (model 1); examples3d/helper.rs order
            for (int i = 0; i < ni; ++i)
                for (int j = 0; j < nj; ++j)
                    for (int k = 0; k < nk; ++k) {
                        const float p[3] = {(float)i * ext, (layer == 0 ? y0 + ext * (float)nj : y0) + (float)j * ext, (float)k * ext};
                        pc.push(layer, p, r, 2700.0f);
                    }
        ctx.particles().from_particles(pc);
        const float he[3] = {(float)ni * ext / 2 + 6.0f, 1.0f, (float)nk * ext / 2 + 6.0f}, tr[3] = {(float)ni * ext / 2, 1.0f, (float)nk * ext / 2};
        if (argc > 6 && std::string(argv[6]) == "heightfield") {
            // a 41 x 41 heightfield like the reference scene's ground (examples3d/sand3.rs:23-38), with a bumpy surface made of exactly
            // representable samples (no transcendental: hosts in other languages reproduce it bit for bit); column-major, rows <-> z
            const int n = 41;
            std::vector<float> heights((size_t)n * n);
            for (int col = 0; col < n; ++col)
                for (int row = 0; row < n; ++row) heights[(size_t)col * n + row] = (float)((col * 3 + row * 5) % 7) * 0.001f;
            const float scale[3] = {40.0f, 10.0f, 40.0f}, at[3] = {(float)ni * ext / 2, y0 - 0.3f, (float)nk * ext / 2};
            ctx.colliders().from_collider_set({CudaColliderSet::heightfield(heights.data(), n, n, scale, at)});
        } else {
            ctx.colliders().from_collider_set({CudaColliderSet::cuboid(he, tr)});
        }
        CudaMpmPipeline pipeline;
        SpkSolverParameters params = default_solver_parameters();
        const float gravity[3] = {0.0f, -9.81f, 0.0f};
        std::vector<SingleGpuMpmContext*> contexts = {&ctx};
        SpkStepTimings t;
        float simulated = 0.0f;
        for (int f = 0; f < frames; ++f) simulated += pipeline.step(&t, params, gravity, contexts);
        const auto pos = ctx.particles().read_positions();
        double ymean = 0.0;
        for (const auto& p : pos) ymean += p.point[1];
        std::printf("particles %zu frames %d substeps %llu simulated %.6f last_dt %.3e active_blocks %u mean_y %.6f\n", pos.size(), frames,
                    (unsigned long long)pipeline.step_id, simulated, t.last_dt, t.num_active_blocks, ymean / (double)pos.size());
        if (argc > 5) {
            FILE* f = std::fopen(argv[5], "wb");
            std::fwrite(pos.data(), sizeof(pos[0]), pos.size(), f);
            std::fclose(f);
        }
        // error behaviour: the reference's assert on mismatched attribute lengths surfaces as an exception, not a crash
        ParticleComponents bad = pc;
        bad.velocity.pop_back();
        try { ctx.particles().from_particles(bad); std::printf("ERROR: length mismatch accepted\n"); return 2; }
        catch (const CudaError& e) { if (e.code != SPK_EINVAL) return 3; }
    } catch (const sparkl::CudaError& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return e.code > 0 ? 10 : 11; // 10: no usable CUDA device (there is no CPU fallback)
    }
    return 0;
}
