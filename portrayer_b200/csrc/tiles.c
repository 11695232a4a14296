/*
 * tiles.c — which pixels a rank renders, and in which order (pure host C, no CUDA).
 *
 * The image is cut into tile_w x tile_h tiles; tile k (row-major over the tile
 * grid) belongs to rank k % world, which interleaves sky and geometry across
 * GPUs.  Inside a tile pixels are listed in 8x4 micro-tiles so that 32
 * consecutive paths (one warp at SAMPLES=1) cover a compact block of the image.
 * Only pixels inside the inclusive slice are listed (src/render.rs:116-119,136-138).
 */
#include "portrayer_gpu.h"

uint64_t pt_owned_pixels(const PtRenderParams* p, uint32_t* index_out, uint64_t capacity) {
    if (!p || p->width == 0 || p->height == 0) return 0;
    const uint32_t tw = p->tile_w ? p->tile_w : 32, th = p->tile_h ? p->tile_h : 32;
    const uint32_t tiles_x = (p->width + tw - 1) / tw, tiles_y = (p->height + th - 1) / th;
    const uint32_t world = p->world > 1 ? p->world : 1;
    const uint32_t rank = world > 1 ? p->rank : 0;
    uint64_t n = 0;
    for (uint32_t ty = 0; ty < tiles_y; ++ty)
        for (uint32_t tx = 0; tx < tiles_x; ++tx) {
            const uint64_t tile = (uint64_t)ty * tiles_x + tx;
            if (tile % world != rank) continue;
            const uint32_t x0 = tx * tw, y0 = ty * th;
            for (uint32_t my = 0; my < th; my += 4)
                for (uint32_t mx = 0; mx < tw; mx += 8)
                    for (uint32_t dy = 0; dy < 4 && my + dy < th; ++dy)
                        for (uint32_t dx = 0; dx < 8 && mx + dx < tw; ++dx) {
                            const uint32_t x = x0 + mx + dx, y = y0 + my + dy;
                            if (x >= p->width || y >= p->height) continue;
                            if (x < p->x1 || x > p->x2 || y < p->y1 || y > p->y2) continue;
                            if (index_out && n < capacity) index_out[n] = y * p->width + x;
                            ++n;
                        }
        }
    return n;
}
