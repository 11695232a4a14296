// pth_image_render: `Image::render` of the host mirror (render.cpp) behind the C ABI.  Lives in its own library
// (libportrayer_render.so) because it is the one entry point of the host mirror that needs libportrayer_gpu.so;
// everything else (libportrayer_host.so) is GPU-free, so the CPU reference arm of bench.py never maps the CUDA library.
#include <cstring>
#include <stdexcept>

#include "capi.h"
#include "capi_internal.hpp"
#include "render.hpp"

using namespace portrayer;

extern "C" {

int pth_image_render(const PthScene* s, uint32_t width, uint32_t height, uint32_t samples, uint32_t rng_mode,
                     uint64_t seed, uint8_t* rgb_inout, PtStats* stats) {
    try {
        if (s->example.prebuilt) throw std::runtime_error("prebuilt known-answer scenes have no HierScene to render");
        Image image("", width, height);
        std::memcpy(image.buffer().data(), rgb_inout, image.buffer().size());
        RenderOptions opts;
        opts.samples = samples;
        opts.rng_mode = rng_mode;
        opts.seed = seed;
        opts.stats = stats;
        image.render<NullProgress>(s->example.scene, s->example.cam, s->example.background, opts);
        std::memcpy(rgb_inout, image.buffer().data(), image.buffer().size());
        return 0;
    } catch (const std::exception& e) {
        capi_error() = e.what();
        return -1;
    }
}

}  // extern "C"
