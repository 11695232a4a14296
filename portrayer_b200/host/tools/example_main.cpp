// The `cargo run --example X` of this repo: build a scene program, call
// Image::render (which goes through the C ABI to the GPU), save the image.
//   portrayer_example <name> [out.png]      env: SAMPLES, KD_DEPTH, KD_MESH_DEPTH, PORTRAYER_RNG, PORTRAYER_SEED
#include <cstdio>
#include <string>

#include "assets.hpp"
#include "examples/examples.hpp"
#include "render.hpp"

using namespace portrayer;

int main(int argc, char** argv) {
    if (argc < 2) {
        std::printf("usage: %s <example> [out.png]\nexamples:\n", argv[0]);
        for (const auto& kv : example_registry()) std::printf("  %s\n", kv.first.c_str());
        return 2;
    }
    if (const char* dir = std::getenv("PORTRAYER_ASSETS")) set_assets_dir(dir);
    try {
        auto it = example_registry().find(argv[1]);
        if (it == example_registry().end()) throw std::runtime_error("unknown example");
        ExampleScene ex = it->second();
        if (ex.prebuilt) throw std::runtime_error("known-answer scenes are driven from the tests");
        if (pt_init(-1) != PT_OK) throw std::runtime_error(pt_last_error());
        std::string out = argc > 2 ? argv[2] : ex.name + ".png";
        Image image(out, ex.width, ex.height);
        PtStats stats{};
        RenderOptions opts = RenderOptions::from_env();
        opts.stats = &stats;
        image.render<RenderProgress>(ex.scene, ex.cam, ex.background, opts);
        image.save();
        const double rays = double(stats.rays_primary + stats.rays_shadow + stats.rays_reflect + stats.rays_refract);
        std::printf("%s: %.3f ms on device, %.1f Mrays/s, %u launches\n", out.c_str(), stats.device_ms,
                    rays / (stats.device_ms * 1e3), stats.kernel_launches);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
