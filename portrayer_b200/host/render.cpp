#include "render.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>

#include "kdtree.hpp"
#include "pack.hpp"

namespace portrayer {

RenderProgress::~RenderProgress() { std::printf("Done!\n"); }
void RenderProgress::report_finished_pixels(uint64_t finished) {
    uint64_t pos = completed_.fetch_add(finished) + finished;
    int percent = static_cast<int>(static_cast<double>(pos) / static_cast<double>(pixels_) * 100.0);
    if (percent != last_percent_) {
        last_percent_ = percent;
        std::printf("%d%%\n", percent);
        std::fflush(stdout);
    }
}

RenderOptions RenderOptions::from_env() {
    RenderOptions o;
    if (const char* m = std::getenv("PORTRAYER_RNG")) o.rng_mode = std::strcmp(m, "fixed") == 0 ? PT_RNG_FIXED : PT_RNG_HASH;
    if (const char* s = std::getenv("PORTRAYER_SEED")) o.seed = std::strtoull(s, nullptr, 10);
    if (const char* l = std::getenv("PORTRAYER_LINEAR")) o.linear_tlas = std::strcmp(l, "1") == 0;
    return o;
}

static uint32_t samples_from_env() {
    // Must be a valid number, must be positive, else the default. render.rs:107-113
    const char* v = std::getenv("SAMPLES");
    if (v && *v) {
        char* end = nullptr;
        unsigned long long n = std::strtoull(v, &end, 10);
        if (end != v && *end == '\0' && n > 0 && v[0] != '-') return static_cast<uint32_t>(n);
    }
    return PT_DEFAULT_SAMPLES;
}

std::vector<double> evaluate_background(const std::function<Rgb(Uv)>& background, size_t width, size_t height,
                                        uint32_t* bg_mode_out) {
    std::vector<double> px(width * height * 3);
    bool rows_uniform = true;
    for (size_t y = 0; y < height; ++y)
        for (size_t x = 0; x < width; ++x) {
            Rgb c = background(Uv{(double)x / (double)width, (double)y / (double)height});
            double* p = &px[(y * width + x) * 3];
            p[0] = c.r; p[1] = c.g; p[2] = c.b;
            if (x > 0 && std::memcmp(p, &px[(y * width) * 3], 3 * sizeof(double)) != 0) rows_uniform = false;
        }
    if (!rows_uniform) {
        *bg_mode_out = PT_BG_PER_PIXEL;
        return px;
    }
    std::vector<double> rows(height * 3);
    for (size_t y = 0; y < height; ++y) std::memcpy(&rows[y * 3], &px[(y * width) * 3], 3 * sizeof(double));
    *bg_mode_out = PT_BG_PER_ROW;
    return rows;
}

ImageSliceMut::ImageSliceMut(Image& image, std::pair<size_t, size_t> top_left, std::pair<size_t, size_t> bottom_right)
    : image_(image), top_left_(top_left), bottom_right_(bottom_right) {
    const size_t w = image.width(), h = image.height();
    if (top_left.first >= w || top_left.second >= h || bottom_right.first >= w || bottom_right.second >= h) {
        char msg[256];
        std::snprintf(msg, sizeof msg,
                      "The positions {x: %zu, y: %zu} and/or {x: %zu, y: %zu} are not within an image with width = %zu and height = %zu",
                      top_left.first, top_left.second, bottom_right.first, bottom_right.second, w, h);
        throw std::out_of_range(msg);  // render.rs:83-86 panics
    }
}
size_t ImageSliceMut::width() const { return image_.width(); }
size_t ImageSliceMut::height() const { return image_.height(); }

void ImageSliceMut::render_impl(const HierScene& scene, const CameraSettings& camera,
                                const std::function<Rgb(Uv)>& background, const RenderOptions& opts, PtProgressFn cb,
                                void* user) {
    const double width = static_cast<double>(image_.width());
    const double height = static_cast<double>(image_.height());
    PtCamera cam = make_camera(camera, width, height);

    PtRenderParams params{};
    params.width = static_cast<uint32_t>(image_.width());
    params.height = static_cast<uint32_t>(image_.height());
    params.x1 = static_cast<uint32_t>(top_left_.first);
    params.y1 = static_cast<uint32_t>(top_left_.second);
    params.x2 = static_cast<uint32_t>(bottom_right_.first);
    params.y2 = static_cast<uint32_t>(bottom_right_.second);
    params.samples = opts.samples ? opts.samples : samples_from_env();
    params.rng_mode = opts.rng_mode;
    params.seed = opts.seed;
    if (opts.linear_tlas) params.flags |= PT_RENDER_LINEAR_TLAS;

    // #[cfg(feature = "kdtree")] FlatScene::from + KDTreeScene::from. render.rs:123-126
    FlatScene flat = FlatScene::from(scene);
    KDTreeScene kd = KDTreeScene::from(std::move(flat));
    std::vector<uint8_t> blob = pack_scene(kd);

    std::vector<double> bg = evaluate_background(background, image_.width(), image_.height(), &params.bg_mode);

    PtScene* dev_scene = nullptr;
    int rc = pt_scene_upload(blob.data(), blob.size(), &dev_scene);
    if (rc != PT_OK) throw std::runtime_error(std::string("pt_scene_upload: ") + pt_last_error());
    rc = pt_render(dev_scene, &cam, &params, bg.data(), image_.buffer().data(), nullptr, nullptr, cb, user, opts.stats);
    pt_scene_free(dev_scene);
    // the reference panics (fused across rayon, render.rs:36,130); re-raise with the same text
    if (rc != PT_OK) throw std::runtime_error(pt_last_error());
}

// ------------------------------------------------------------------ Image I/O (host, outside the hot path)
namespace {
uint32_t crc32_of(const uint8_t* data, size_t n, uint32_t crc = 0) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ data[i]) & 0xFF] ^ (crc >> 8);
    return ~crc;
}
void put_be32(std::vector<uint8_t>& v, uint32_t x) {
    v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x);
}
void write_chunk(std::ofstream& out, const char* tag, const std::vector<uint8_t>& data) {
    std::vector<uint8_t> buf;
    put_be32(buf, static_cast<uint32_t>(data.size()));
    buf.insert(buf.end(), tag, tag + 4);
    buf.insert(buf.end(), data.begin(), data.end());
    put_be32(buf, crc32_of(buf.data() + 4, buf.size() - 4));
    out.write(reinterpret_cast<const char*>(buf.data()), buf.size());
}
bool read_ppm(const std::string& path, size_t w, size_t h, std::vector<uint8_t>& out) {
    std::ifstream in(path, std::ios::binary);
    if (!in) return false;
    std::string magic;
    size_t fw = 0, fh = 0, maxv = 0;
    in >> magic >> fw >> fh >> maxv;
    if (magic != "P6" || fw != w || fh != h || maxv != 255) return false;
    in.get();
    std::vector<uint8_t> buf(w * h * 3);
    in.read(reinterpret_cast<char*>(buf.data()), buf.size());
    if (!in) return false;
    out.swap(buf);
    return true;
}
}  // namespace

Image::Image(const std::string& path, size_t width, size_t height)
    : path_(path), width_(width), height_(height), buffer_(width * height * 3, 0) {
    // Existing image of equal size keeps its pixels (poor-man's resume, render.rs:165-176).
    // Only the PPM form is re-read here; PNG decode is host I/O left to the caller.
    if (path.size() > 4 && path.substr(path.size() - 4) == ".ppm") read_ppm(path, width, height, buffer_);
}

void Image::save_as(const std::string& path) const {
    std::ofstream out(path, std::ios::binary);
    if (!out) throw std::runtime_error("cannot write " + path);
    if (path.size() > 4 && path.substr(path.size() - 4) == ".ppm") {
        out << "P6\n" << width_ << " " << height_ << "\n255\n";
        out.write(reinterpret_cast<const char*>(buffer_.data()), buffer_.size());
        return;
    }
    // PNG with stored (uncompressed) deflate blocks
    const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    out.write(reinterpret_cast<const char*>(sig), 8);
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, static_cast<uint32_t>(width_));
    put_be32(ihdr, static_cast<uint32_t>(height_));
    ihdr.insert(ihdr.end(), {8, 2, 0, 0, 0});
    write_chunk(out, "IHDR", ihdr);
    std::vector<uint8_t> raw;
    raw.reserve((width_ * 3 + 1) * height_);
    for (size_t y = 0; y < height_; ++y) {
        raw.push_back(0);
        raw.insert(raw.end(), buffer_.begin() + y * width_ * 3, buffer_.begin() + (y + 1) * width_ * 3);
    }
    std::vector<uint8_t> z = {0x78, 0x01};
    uint32_t a = 1, b = 0;
    for (uint8_t v : raw) { a = (a + v) % 65521; b = (b + a) % 65521; }
    for (size_t pos = 0; pos < raw.size() || pos == 0;) {
        size_t n = std::min<size_t>(65535, raw.size() - pos);
        bool last = pos + n >= raw.size();
        z.push_back(last ? 1 : 0);
        z.push_back(n & 0xFF); z.push_back(n >> 8);
        z.push_back(~n & 0xFF); z.push_back((~n >> 8) & 0xFF);
        z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
        pos += n;
        if (last) break;
    }
    put_be32(z, (b << 16) | a);
    write_chunk(out, "IDAT", z);
    write_chunk(out, "IEND", {});
}

}  // namespace portrayer
