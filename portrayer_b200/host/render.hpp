// Host mirror of the render entry point, src/render.rs:154-224 (Image) and
// :53-152 (ImageSliceMut), with the rayon pixel loop (:127-150) replaced by
// one call into the C ABI of include/portrayer_gpu.h.
// Everything above the loop is kept: Camera::new (:101), R::new (:103), the
// SAMPLES parse (:107-113), FlatScene::from + KDTreeScene::from (:124-126).
#pragma once
#include <atomic>
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

#include "portrayer_gpu.h"
#include "scene.hpp"

namespace portrayer {

// src/reporter.rs:10-13
struct NullProgress {  // reporter.rs:89-97
    explicit NullProgress(uint64_t) {}
    void report_finished_pixels(uint64_t) {}
};
class RenderProgress {  // reporter.rs:16-84 (CI=true style output: a percentage line per report)
  public:
    explicit RenderProgress(uint64_t pixels) : pixels_(pixels) {}
    ~RenderProgress();
    void report_finished_pixels(uint64_t finished);

  private:
    uint64_t pixels_;
    std::atomic<uint64_t> completed_{0};
    int last_percent_ = -1;
};

// How the deterministic jitter replaces thread_rng() for this render; read from
// the environment so example programs stay unchanged:
//   PORTRAYER_RNG=fixed|hash (default hash), PORTRAYER_SEED=<u64> (default 1)
struct RenderOptions {
    uint32_t samples = 0;  // 0 -> SAMPLES env, default 100 (render.rs:107-113)
    uint32_t rng_mode = PT_RNG_HASH;
    uint64_t seed = 1;
    bool linear_tlas = false;  // the reference's default build (no kdtree feature) scans linearly
    PtStats* stats = nullptr;
    static RenderOptions from_env();
};

class Image;

class ImageSliceMut {
  public:
    ImageSliceMut(Image& image, std::pair<size_t, size_t> top_left, std::pair<size_t, size_t> bottom_right);

    template <class R = RenderProgress>
    void render(const HierScene& scene, const CameraSettings& camera, const std::function<Rgb(Uv)>& background,
                RenderOptions opts = RenderOptions::from_env()) {
        R reporter(static_cast<uint64_t>(width()) * height());  // total = full image even for slices (render.rs:103)
        auto cb = [](void* user, uint64_t n) { static_cast<R*>(user)->report_finished_pixels(n); };
        render_impl(scene, camera, background, opts, cb, &reporter);
    }

  private:
    size_t width() const;
    size_t height() const;
    void render_impl(const HierScene& scene, const CameraSettings& camera, const std::function<Rgb(Uv)>& background,
                     const RenderOptions& opts, PtProgressFn cb, void* user);
    Image& image_;
    std::pair<size_t, size_t> top_left_, bottom_right_;
};

class Image {
  public:
    // Opens `path` if it exists with the same dimensions (keeps its pixels), else a black image. render.rs:165-188
    Image(const std::string& path, size_t width, size_t height);
    size_t width() const { return width_; }
    size_t height() const { return height_; }
    void save() const { save_as(path_); }
    void save_as(const std::string& path) const;  // .png (stored deflate) or .ppm
    ImageSliceMut slice_mut(std::pair<size_t, size_t> top_left, std::pair<size_t, size_t> bottom_right) {
        return ImageSliceMut(*this, top_left, bottom_right);
    }
    template <class R = RenderProgress>
    void render(const HierScene& scene, const CameraSettings& camera, const std::function<Rgb(Uv)>& background,
                RenderOptions opts = RenderOptions::from_env()) {
        ImageSliceMut(*this, {0, 0}, {width_ - 1, height_ - 1}).render<R>(scene, camera, background, opts);
    }
    std::vector<uint8_t>& buffer() { return buffer_; }
    const std::vector<uint8_t>& buffer() const { return buffer_; }

  private:
    std::string path_;
    size_t width_, height_;
    std::vector<uint8_t> buffer_;  // RGB8 row-major
};

// background.at(Uv{x / width, y / height}) for every integer pixel (render.rs:31-34).
// Returns PT_BG_PER_ROW data (H*3) when every row is uniform, else PT_BG_PER_PIXEL (W*H*3).
std::vector<double> evaluate_background(const std::function<Rgb(Uv)>& background, size_t width, size_t height,
                                        uint32_t* bg_mode_out);

}  // namespace portrayer
