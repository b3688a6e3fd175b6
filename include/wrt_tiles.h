/* wrt_tiles.h — image sharding shared by the CUDA core (raygen / resolve / scatter
 * kernels), the host library (wrt_tile_pixel_map) and the multi-GPU driver.
 *
 * The reference renders pixels in one serial loop (Renderer.hpp:104-131); every
 * pixel is independent, so the frame shards as interleaved tiles.  The image is
 * cut into tile_w x tile_h tiles (multiples of the 8 x 4 pixel block a warp
 * traces; the default tile IS that block).  Rank r owns the tile sequence k = r, r + world, r + 2*world, ...;
 * sequence number k maps to image tile (k * perm_mul) % n_tiles with perm_mul
 * coprime to n_tiles, which scatters each rank's tiles over the whole image
 * (the bunny covers ~7 % of the pixels but spawns about half of the rays, so
 * contiguous or column-aligned ownership would not balance).
 * A rank's pixels are addressed by "slots": slot = local_tile * tile_pixels +
 * 32 * block + lane, blocks row-major inside the tile, lanes row-major inside
 * the 8 x 4 block.  Slots of clipped border tiles that fall outside the image
 * are padding.
 */
#ifndef WRT_TILES_H
#define WRT_TILES_H

#include <stdint.h>

#ifdef __CUDACC__
#define WRT_TILE_HD __host__ __device__ __forceinline__
#else
#define WRT_TILE_HD static inline
#endif

typedef struct WrtTileMap {
    int32_t width, height;
    int32_t tile_w, tile_h, tiles_x, tiles_y;
    int32_t rank, world;
    int64_t n_tiles;
    int64_t perm_mul;
} WrtTileMap;

WRT_TILE_HD int64_t wrt_tile_gcd(int64_t a, int64_t b) {
    while (b) { int64_t t = a % b; a = b; b = t; }
    return a;
}

WRT_TILE_HD WrtTileMap wrt_tilemap_make(int width, int height, int tile_w, int tile_h, int rank, int world) {
    WrtTileMap tm;
    tm.width = width; tm.height = height;
    tm.tile_w = tile_w; tm.tile_h = tile_h;
    tm.tiles_x = (width + tile_w - 1) / tile_w;
    tm.tiles_y = (height + tile_h - 1) / tile_h;
    tm.rank = rank; tm.world = world;
    tm.n_tiles = (int64_t)tm.tiles_x * tm.tiles_y;
    /* multiplier near the golden-ratio fraction of n_tiles, bumped until coprime */
    int64_t m = ((int64_t)((double)tm.n_tiles * 0.6180339887498949)) | 1;
    if (tm.n_tiles <= 2) m = 1;
    while (m > 1 && wrt_tile_gcd(m, tm.n_tiles) != 1) m += 2;
    tm.perm_mul = m;
    return tm;
}

/* Slots (including padding) owned by rank r of w. */
WRT_TILE_HD int64_t wrt_tilemap_slots(const WrtTileMap* tm, int r, int w) {
    int64_t mine = (tm->n_tiles - r + w - 1) / w;
    if (mine < 0) mine = 0;
    return mine * tm->tile_w * tm->tile_h;
}

/* Local slot of rank `r` -> pixel; returns 0 when the slot is padding. */
WRT_TILE_HD int wrt_tilemap_slot_to_pixel(const WrtTileMap* tm, int64_t slot, int r, int* px, int* py) {
    int tp = tm->tile_w * tm->tile_h;
    int64_t tl = slot / tp;
    int within = (int)(slot - tl * tp);
    int64_t k = tl * tm->world + r;
    if (k >= tm->n_tiles) return 0;
    int64_t t = (k * tm->perm_mul) % tm->n_tiles;
    int tx = (int)(t % tm->tiles_x), ty = (int)(t / tm->tiles_x);
    int bw = tm->tile_w / 8;
    int b = within >> 5, lane = within & 31;
    int bx = b % bw, by = b / bw;
    *px = tx * tm->tile_w + bx * 8 + (lane & 7);
    *py = ty * tm->tile_h + by * 4 + (lane >> 3);
    return *px < tm->width && *py < tm->height;
}

#endif /* WRT_TILES_H */
