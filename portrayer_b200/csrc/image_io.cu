// The two byte-moving steps either side of the render loop, on the device.
//
//   texture ingest   src/texture.rs:96-102  RgbImageBuffer::open = image::open(path)?.to_rgb()
//       The entropy decode of the JPEG / PNG file stays on the host (the `image` crate there, Pillow in this repo's
//       mirror: neither reproduces jpeg-decoder 0.1.15 bit for bit off-line).  What follows the
//       decode is data-parallel: `to_rgb()` — whatever pixel layout the decoder produced (Luma8, LumaA8, Rgb8, Rgba8,
//       Bgr8, Bgra8: the DynamicImage variants of image 0.21) becomes RGB8 — and the texel lands in the pool the shade
//       kernel samples, under its residency key.  One kernel, HBM-bound: c bytes in, 3 bytes out per pixel.
//
//   PNG encode       src/render.rs:200-208  Image::save -> image::ImageBuffer::save -> PNG (8-bit RGB)
//       The frame is already in HBM when the render ends.  The file is assembled there: filter-0 scanlines inside
//       stored deflate blocks (zlib stream, Adler-32 trailer), one IDAT chunk with its CRC-32, IHDR and IEND — so what
//       crosses PCIe once is the finished file.  Both checksums are computed in parallel: Adler-32 over the monoid
//       (A, B, n) (+) (A', B', n') = (A + A', B + n' A + B', n + n') mod 65521, CRC-32 per 512-byte segment with a byte
//       table in shared memory and zlib's crc32_combine (multiplication by x^(8 n) in GF(2)[x] / P) to stitch the
//       segments.  Decoders see the reference's pixels; the BYTES of the file are not the `image` crate's (it deflates
//       with compression), which no test of the reference looks at.
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.h"

namespace ptd {
namespace {

constexpr int kIoBlock = 256;

// ------------------------------------------------------------------ to_rgb
// 4 pixels per thread; layout: channels per pixel and where R, G, B sit (gray: all from channel 0)
__global__ void __launch_bounds__(kIoBlock) to_rgb_kernel(const uint8_t* __restrict__ src, uint64_t n_pixels, uint32_t channels, uint32_t r_at,
                                                          uint32_t g_at, uint32_t b_at, uint8_t* __restrict__ dst) {
    const uint64_t first = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (first >= n_pixels) return;
    const uint32_t count = (uint32_t)min((uint64_t)4, n_pixels - first);
    uint8_t out[12];
    for (uint32_t k = 0; k < count; ++k) {
        const uint8_t* p = src + (first + k) * channels;
        out[3 * k] = p[r_at];
        out[3 * k + 1] = p[g_at];
        out[3 * k + 2] = p[b_at];
    }
    uint8_t* d = dst + first * 3;
    if (count == 4 && ((uintptr_t)d & 3u) == 0) {  // 12 bytes as three aligned words
        uint32_t* d32 = reinterpret_cast<uint32_t*>(d);
#pragma unroll
        for (int wd = 0; wd < 3; ++wd)
            d32[wd] = (uint32_t)out[4 * wd] | (uint32_t)out[4 * wd + 1] << 8 | (uint32_t)out[4 * wd + 2] << 16 | (uint32_t)out[4 * wd + 3] << 24;
    } else {
        for (uint32_t k = 0; k < count * 3; ++k) d[k] = out[k];
    }
}

// ------------------------------------------------------------------ PNG
struct PngLayout {
    uint32_t width, height;
    uint64_t row_bytes;   // 1 + 3 * width: filter byte + pixels
    uint64_t raw_len;     // height * row_bytes
    uint64_t n_blocks;    // stored deflate blocks of <= 65535 bytes
    uint64_t z_len;       // 2 + 5 * n_blocks + raw_len + 4
    uint64_t idat_data;   // offset of the zlib stream in the file = 8 + 25 + 8
    uint64_t total;
};
__host__ __device__ inline PngLayout png_layout(uint32_t width, uint32_t height) {
    PngLayout L;
    L.width = width;
    L.height = height;
    L.row_bytes = 1 + 3 * (uint64_t)width;
    L.raw_len = (uint64_t)height * L.row_bytes;
    L.n_blocks = L.raw_len ? (L.raw_len + 65534) / 65535 : 1;
    L.z_len = 2 + 5 * L.n_blocks + L.raw_len + 4;
    L.idat_data = 8 + 25 + 8;
    L.total = L.idat_data + L.z_len + 4 + 12;
    return L;
}

// byte r of the filtered scanline stream: a zero filter byte, then the row's pixels
__device__ __forceinline__ uint8_t raw_byte(const uint8_t* __restrict__ rgb, const PngLayout& L, uint64_t r) {
    const uint64_t y = r / L.row_bytes, c = r - y * L.row_bytes;
    return c == 0 ? (uint8_t)0 : rgb[y * (L.row_bytes - 1) + (c - 1)];
}

// the deflate part of the zlib stream (block headers + data), 4 output bytes per thread
__global__ void __launch_bounds__(kIoBlock) png_pack_kernel(const uint8_t* __restrict__ rgb, PngLayout L, uint8_t* __restrict__ png) {
    const uint64_t body = 5 * L.n_blocks + L.raw_len;
    const uint64_t first = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (first >= body) return;
    uint8_t* out = png + L.idat_data + 2;
    for (uint64_t p = first; p < min(first + 4, body); ++p) {
        const uint64_t b = p / 65540, off = p - b * 65540;
        uint8_t v;
        if (off < 5) {
            const uint64_t len = min((uint64_t)65535, L.raw_len - b * 65535);
            const uint32_t n = (uint32_t)len, nn = ~n & 0xFFFFu;
            v = off == 0 ? (uint8_t)(b + 1 == L.n_blocks ? 1 : 0)  // BFINAL, BTYPE = 00 (stored)
                : off == 1 ? (uint8_t)(n & 0xFF) : off == 2 ? (uint8_t)(n >> 8) : off == 3 ? (uint8_t)(nn & 0xFF) : (uint8_t)(nn >> 8);
        } else {
            v = raw_byte(rgb, L, b * 65535 + (off - 5));
        }
        out[p] = v;
    }
}

// Adler-32 partials: one (A, B, n) triple per block over kIoBlock * 64 raw bytes, reduced in order
struct Adler {
    uint64_t a, b, n;
};
__device__ __forceinline__ Adler adler_join(const Adler& x, const Adler& y) {  // x first, then y
    return Adler{(x.a + y.a) % 65521u, (x.b + (y.n % 65521u) * x.a + y.b) % 65521u, x.n + y.n};
}
constexpr uint32_t kAdlerPerThread = 64;
__global__ void __launch_bounds__(kIoBlock) png_adler_partial_kernel(const uint8_t* __restrict__ rgb, PngLayout L, Adler* __restrict__ partial) {
    const uint64_t first = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * kAdlerPerThread;
    Adler m{0, 0, 0};
    if (first < L.raw_len) {
        const uint32_t n = (uint32_t)min((uint64_t)kAdlerPerThread, L.raw_len - first);
        uint64_t a = 0, b = 0;
        for (uint32_t j = 0; j < n; ++j) {
            const uint64_t d = raw_byte(rgb, L, first + j);
            a += d;
            b += (uint64_t)(n - j) * d;
        }
        m = Adler{a, b, n};
    }
    // ordered tree reduction inside the warp, then across the block's warps
    for (int off = 1; off < 32; off <<= 1) {
        Adler o;
        o.a = __shfl_down_sync(0xFFFFFFFFu, m.a, off);
        o.b = __shfl_down_sync(0xFFFFFFFFu, m.b, off);
        o.n = __shfl_down_sync(0xFFFFFFFFu, m.n, off);
        if (((threadIdx.x & 31) & (2 * off - 1)) == 0) m = adler_join(m, o);
    }
    __shared__ Adler s[kIoBlock / 32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        Adler t = s[0];
        for (int w = 1; w < kIoBlock / 32; ++w) t = adler_join(t, s[w]);
        partial[blockIdx.x] = t;
    }
}
// fold the partials from the initial state (a = 1, b = 0) and write the big-endian trailer
__global__ void png_adler_final_kernel(const Adler* __restrict__ partial, uint32_t n_partials, PngLayout L, uint8_t* __restrict__ png) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint64_t a = 1, b = 0;
    for (uint32_t i = 0; i < n_partials; ++i) {
        const Adler p = partial[i];
        b = (b + (p.n % 65521u) * a + p.b) % 65521u;
        a = (a + p.a) % 65521u;
    }
    const uint32_t adler = (uint32_t)(b << 16 | a);
    uint8_t* t = png + L.idat_data + L.z_len - 4;
    t[0] = (uint8_t)(adler >> 24); t[1] = (uint8_t)(adler >> 16); t[2] = (uint8_t)(adler >> 8); t[3] = (uint8_t)adler;
}

// ---- CRC-32 (reflected, polynomial 0xEDB88320), zlib's formulation of GF(2) polynomial arithmetic
constexpr uint32_t kCrcPoly = 0xEDB88320u;
__host__ __device__ inline uint32_t crc_multmodp(uint32_t a, uint32_t b) {  // a(x) * b(x) mod P
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ kCrcPoly : b >> 1;
    }
    return p;
}
__host__ __device__ inline uint32_t crc_x8n(uint64_t n_bytes) {  // x^(8 n) mod P
    uint32_t x2n = 0x40000000u;  // x^1
    // x^(2^3) first: the exponent is 8 n = n * 2^3
    x2n = crc_multmodp(x2n, x2n);
    x2n = crc_multmodp(x2n, x2n);
    x2n = crc_multmodp(x2n, x2n);
    uint32_t p = 1u << 31;  // x^0
    while (n_bytes) {
        if (n_bytes & 1u) p = crc_multmodp(x2n, p);
        n_bytes >>= 1;
        x2n = crc_multmodp(x2n, x2n);
    }
    return p;
}
__device__ __forceinline__ uint32_t crc_join(uint32_t crc_first, uint32_t crc_second, uint32_t x8n_second) {  // zlib crc32_combine
    return crc_multmodp(x8n_second, crc_first) ^ crc_second;
}

constexpr uint32_t kCrcSeg = 512;  // bytes per thread
// standard CRC-32 (init ~0, final ~) of every kCrcSeg-byte segment of png[begin, begin + len)
__global__ void __launch_bounds__(kIoBlock) png_crc_segments_kernel(const uint8_t* __restrict__ data, uint64_t len, uint32_t* __restrict__ seg_crc) {
    __shared__ uint32_t table[256];
    {
        uint32_t c = threadIdx.x;
        for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ kCrcPoly : c >> 1;
        table[threadIdx.x] = c;  // kIoBlock == 256
    }
    __syncthreads();
    const uint64_t seg = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t first = seg * kCrcSeg;
    if (first >= len) return;
    const uint32_t n = (uint32_t)min((uint64_t)kCrcSeg, len - first);
    uint32_t crc = 0xFFFFFFFFu;
    const uint8_t* p = data + first;
    for (uint32_t j = 0; j < n; ++j) crc = table[(crc ^ p[j]) & 0xFFu] ^ (crc >> 8);
    seg_crc[seg] = ~crc;
}
// one block: thread t folds a contiguous run of segments, thread 0 folds the runs; then the framing is written:
// signature, IHDR, the IDAT length / type / zlib header, the IDAT CRC, IEND
__global__ void __launch_bounds__(kIoBlock) png_crc_final_kernel(const uint32_t* __restrict__ seg_crc, uint64_t n_segs, uint64_t len, PngLayout L,
                                                                 uint8_t* __restrict__ png) {
    __shared__ uint32_t run_crc[kIoBlock];
    __shared__ uint32_t run_x8n[kIoBlock];
    const uint64_t per = (n_segs + kIoBlock - 1) / kIoBlock;
    const uint64_t s0 = min((uint64_t)threadIdx.x * per, n_segs), s1 = min(s0 + per, n_segs);
    const uint32_t x_full = crc_x8n(kCrcSeg);
    const uint64_t last_len = len - (n_segs - 1) * kCrcSeg;
    const uint32_t x_last = crc_x8n(last_len);
    uint32_t crc = 0;
    uint64_t bytes = 0;
    for (uint64_t sgm = s0; sgm < s1; ++sgm) {
        const bool last = sgm + 1 == n_segs;
        crc = sgm == s0 ? seg_crc[sgm] : crc_join(crc, seg_crc[sgm], last ? x_last : x_full);
        bytes += last ? last_len : kCrcSeg;
    }
    run_crc[threadIdx.x] = crc;
    run_x8n[threadIdx.x] = crc_x8n(bytes);
    __syncthreads();
    if (threadIdx.x != 0) return;
    uint32_t total = run_crc[0];
    for (int t = 1; t < kIoBlock; ++t)
        if (min((uint64_t)t * per, n_segs) < n_segs) total = crc_join(total, run_crc[t], run_x8n[t]);
    auto be32 = [](uint8_t* d, uint32_t v) { d[0] = (uint8_t)(v >> 24); d[1] = (uint8_t)(v >> 16); d[2] = (uint8_t)(v >> 8); d[3] = (uint8_t)v; };
    // IDAT CRC (covers type + data) and IEND
    uint8_t* end = png + L.idat_data + L.z_len;
    be32(end, total);
    const uint8_t iend[12] = {0, 0, 0, 0, 'I', 'E', 'N', 'D', 0xAE, 0x42, 0x60, 0x82};
    for (int k = 0; k < 12; ++k) end[4 + k] = iend[k];
}
// signature + IHDR + the IDAT chunk's length, type and zlib header (everything in front of the deflate blocks)
__global__ void png_header_kernel(PngLayout L, uint8_t* __restrict__ png) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    auto be32 = [](uint8_t* d, uint32_t v) { d[0] = (uint8_t)(v >> 24); d[1] = (uint8_t)(v >> 16); d[2] = (uint8_t)(v >> 8); d[3] = (uint8_t)v; };
    const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    for (int k = 0; k < 8; ++k) png[k] = sig[k];
    uint8_t* ihdr = png + 8;
    be32(ihdr, 13);
    ihdr[4] = 'I'; ihdr[5] = 'H'; ihdr[6] = 'D'; ihdr[7] = 'R';
    be32(ihdr + 8, L.width);
    be32(ihdr + 12, L.height);
    ihdr[16] = 8; ihdr[17] = 2; ihdr[18] = 0; ihdr[19] = 0; ihdr[20] = 0;  // 8-bit, colour type 2 (RGB), deflate, adaptive, no interlace
    uint32_t crc = 0xFFFFFFFFu;
    for (int k = 4; k < 21; ++k) {
        crc ^= ihdr[k];
        for (int b = 0; b < 8; ++b) crc = (crc & 1u) ? (crc >> 1) ^ kCrcPoly : crc >> 1;
    }
    be32(ihdr + 21, ~crc);
    uint8_t* idat = png + 8 + 25;
    be32(idat, (uint32_t)L.z_len);
    idat[4] = 'I'; idat[5] = 'D'; idat[6] = 'A'; idat[7] = 'T';
    idat[8] = 0x78; idat[9] = 0x01;  // zlib: deflate, 32K window, no preset dictionary, fastest
}

}  // namespace

cudaError_t launch_to_rgb(const uint8_t* src, uint64_t n_pixels, uint32_t channels, uint32_t r_at, uint32_t g_at, uint32_t b_at, uint8_t* dst,
                          cudaStream_t st) {
    if (!n_pixels) return cudaSuccess;
    const uint64_t threads = (n_pixels + 3) / 4;
    to_rgb_kernel<<<(unsigned)((threads + kIoBlock - 1) / kIoBlock), kIoBlock, 0, st>>>(src, n_pixels, channels, r_at, g_at, b_at, dst);
    return cudaGetLastError();
}

uint64_t png_file_bytes(uint32_t width, uint32_t height) { return png_layout(width, height).total; }
uint64_t png_scratch_bytes(uint32_t width, uint32_t height) {
    const PngLayout L = png_layout(width, height);
    const uint64_t adler_blocks = (L.raw_len + (uint64_t)kIoBlock * kAdlerPerThread - 1) / ((uint64_t)kIoBlock * kAdlerPerThread) + 1;
    const uint64_t crc_len = 4 + L.z_len;
    const uint64_t n_segs = (crc_len + kCrcSeg - 1) / kCrcSeg;
    return adler_blocks * sizeof(Adler) + n_segs * sizeof(uint32_t) + 64;
}

// d_png: png_file_bytes() bytes; d_scratch: png_scratch_bytes() bytes.  Everything is enqueued on `st`.
cudaError_t launch_png_encode(const uint8_t* d_rgb, uint32_t width, uint32_t height, uint8_t* d_png, void* d_scratch, cudaStream_t st) {
    if (width == 0 || height == 0 || width > (1u << 24) || height > (1u << 24)) return cudaErrorInvalidValue;
    const PngLayout L = png_layout(width, height);
    if (L.z_len > 0x7FFFFFFFull) return cudaErrorInvalidValue;  // one IDAT chunk
    const uint64_t adler_blocks = (L.raw_len + (uint64_t)kIoBlock * kAdlerPerThread - 1) / ((uint64_t)kIoBlock * kAdlerPerThread);
    Adler* partial = static_cast<Adler*>(d_scratch);
    uint32_t* seg_crc = reinterpret_cast<uint32_t*>(partial + adler_blocks + 1);
    png_header_kernel<<<1, 32, 0, st>>>(L, d_png);
    const uint64_t body = 5 * L.n_blocks + L.raw_len;
    png_pack_kernel<<<(unsigned)(((body + 3) / 4 + kIoBlock - 1) / kIoBlock), kIoBlock, 0, st>>>(d_rgb, L, d_png);
    png_adler_partial_kernel<<<(unsigned)adler_blocks, kIoBlock, 0, st>>>(d_rgb, L, partial);
    png_adler_final_kernel<<<1, 32, 0, st>>>(partial, (uint32_t)adler_blocks, L, d_png);
    // the chunk CRC covers the type field and the data: png[idat_data - 4, idat_data + z_len)
    const uint64_t crc_len = 4 + L.z_len;
    const uint64_t n_segs = (crc_len + kCrcSeg - 1) / kCrcSeg;
    png_crc_segments_kernel<<<(unsigned)((n_segs + kIoBlock - 1) / kIoBlock), kIoBlock, 0, st>>>(d_png + L.idat_data - 4, crc_len, seg_crc);
    png_crc_final_kernel<<<1, kIoBlock, 0, st>>>(seg_crc, n_segs, crc_len, L, d_png);
    return cudaGetLastError();
}

}  // namespace ptd
