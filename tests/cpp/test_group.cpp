// Single-process multi-GPU render behind the reference's ONE entry point: Image::render (src/render.rs:216-223) of the
// host mirror, unchanged, first on one device (pt_init) and then on a device group (pt_init_devices): the pictures must
// be identical byte for byte, the group must have rendered every ray exactly once, and the call must get faster.
//   test_group [n_devices (default: all)] [example (default graphics-castle)] [samples (default 4)]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "assets.hpp"
#include "examples/examples.hpp"
#include "render.hpp"

using namespace portrayer;

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Shot {
    std::vector<uint8_t> rgb;
    PtStats stats{};
    double wall_ms = 0.0;
};

static Shot render_once(const ExampleScene& ex, uint32_t samples, int repeats) {
    Shot shot;
    for (int r = 0; r < repeats; ++r) {  // the last repeat is the warm one (frames, graphs and textures cached)
        Image image("/nonexistent/test_group.png", ex.width, ex.height);
        RenderOptions opts = RenderOptions::from_env();
        opts.samples = samples;
        opts.stats = &shot.stats;
        const double t0 = now_ms();
        image.render<NullProgress>(ex.scene, ex.cam, ex.background, opts);
        shot.wall_ms = now_ms() - t0;
        shot.rgb = image.buffer();
    }
    return shot;
}

int main(int argc, char** argv) {
    if (const char* dir = std::getenv("PORTRAYER_ASSETS")) set_assets_dir(dir);
    const int available = pt_device_count();
    int n = argc > 1 ? std::atoi(argv[1]) : available;
    const std::string name = argc > 2 ? argv[2] : "graphics-castle";
    const uint32_t samples = argc > 3 ? (uint32_t)std::atoi(argv[3]) : 4;
    if (available < 1) { std::printf("no CUDA device: %s\n", pt_last_error()); return 2; }
    if (n < 1 || n > available) n = available;
    int failures = 0;
    try {
        auto it = example_registry().find(name);
        if (it == example_registry().end()) throw std::runtime_error("unknown example " + name);
        ExampleScene ex = it->second();

        if (pt_init(0) != PT_OK) throw std::runtime_error(pt_last_error());
        if (pt_device_group_size() != 1) { std::printf("FAILED: group size %d after pt_init\n", pt_device_group_size()); ++failures; }
        const Shot one = render_once(ex, samples, 2);

        std::vector<int> ids(n);
        for (int i = 0; i < n; ++i) ids[i] = i;
        if (pt_init_devices(ids.data(), n) != PT_OK) throw std::runtime_error(pt_last_error());
        if (pt_device_group_size() != n) { std::printf("FAILED: group size %d, expected %d\n", pt_device_group_size(), n); ++failures; }
        const Shot all = render_once(ex, samples, 2);

        const bool same = one.rgb == all.rgb;
        size_t differing = 0;
        for (size_t i = 0; i < one.rgb.size() && i < all.rgb.size(); ++i) differing += one.rgb[i] != all.rgb[i];
        auto rays = [](const PtStats& s) { return s.rays_primary + s.rays_shadow + s.rays_reflect + s.rays_refract; };
        std::printf("%s %zux%zu x%u: 1 device %.2f ms wall (%.2f ms device), %d devices %.2f ms wall (%.2f ms device): x%.2f\n", name.c_str(),
                    ex.width, ex.height, samples, one.wall_ms, one.stats.device_ms, n, all.wall_ms, all.stats.device_ms, one.wall_ms / all.wall_ms);
        if (!same) { std::printf("FAILED: %zu bytes differ between the one-device and the group picture\n", differing); ++failures; }
        if (rays(one.stats) != rays(all.stats) || one.stats.rays_primary != all.stats.rays_primary) {
            std::printf("FAILED: ray counts differ (%llu vs %llu)\n", (unsigned long long)rays(one.stats), (unsigned long long)rays(all.stats));
            ++failures;
        }
        // a slice through the group: only the slice is written (render.rs:136-138)
        {
            Image image("/nonexistent/test_group.png", ex.width, ex.height);
            std::memset(image.buffer().data(), 7, image.buffer().size());
            RenderOptions opts = RenderOptions::from_env();
            opts.samples = samples;
            const size_t x1 = ex.width / 4, y1 = ex.height / 4, x2 = ex.width / 2, y2 = ex.height / 2 + 3;
            image.slice_mut({x1, y1}, {x2, y2}).render<NullProgress>(ex.scene, ex.cam, ex.background, opts);
            size_t bad = 0;
            for (size_t y = 0; y < ex.height; ++y)
                for (size_t x = 0; x < ex.width; ++x) {
                    const bool inside = x >= x1 && x <= x2 && y >= y1 && y <= y2;
                    for (int c = 0; c < 3; ++c) {
                        const uint8_t v = image.buffer()[(y * ex.width + x) * 3 + c];
                        bad += inside ? v != one.rgb[(y * ex.width + x) * 3 + c] : v != 7;
                    }
                }
            if (bad) { std::printf("FAILED: %zu bytes wrong in the slice render through the group\n", bad); ++failures; }
        }
        pt_shutdown();
    } catch (const std::exception& e) {
        std::printf("error: %s\n", e.what());
        return 1;
    }
    if (failures) { std::printf("%d group check(s) FAILED\n", failures); return 1; }
    std::printf("all group checks passed (%d device%s)\n", n, n == 1 ? "" : "s");
    return 0;
}
