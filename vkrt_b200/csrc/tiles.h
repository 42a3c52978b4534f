/* tiles.h — interleaved-tile partition of the image across GPUs (no reference equivalent: vkrt is single-GPU).
 *
 * The image is cut into tileW x tileH tiles (row-major tile ids). Tile (tx, ty) belongs to rank (ty*tilesX + tx + ty) % G:
 * the extra "+ ty" staggers rows so that tile columns do not alias onto one rank when G divides tilesX. Each rank stores
 * its tiles back to back ("tile-compact"): local pixel = localTile * tileW*tileH + (y_in_tile * tileW + x_in_tile).
 * Because every pixel-sample is seeded by (pixel, frame, sample) only (sampling/random.slang:21-27) and the film update is
 * per pixel (writeback.slang:87-114), any partition yields bit-identical pixels to a single-GPU render.
 * Plain C so that the C host, the CUDA library and the CPU tests share one definition. */
#ifndef VKRT_TILES_H
#define VKRT_TILES_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vkrt_tile_layout {
    uint32_t width, height;
    uint32_t tileW, tileH, tilesX, tilesY;
    uint32_t rank, worldSize;
    uint32_t localTileCount;
} vkrt_tile_layout;

static inline uint32_t vkrt_tile_owner(const vkrt_tile_layout* l, uint32_t tx, uint32_t ty) {
    return (uint32_t)(((uint64_t)ty * l->tilesX + tx + ty) % l->worldSize);
}

static inline void vkrt_tile_layout_init(vkrt_tile_layout* l, uint32_t width, uint32_t height, uint32_t tileW, uint32_t tileH, uint32_t rank,
                                         uint32_t worldSize) {
    l->width = width;
    l->height = height;
    l->tileW = tileW ? tileW : 32u;
    l->tileH = tileH ? tileH : 32u;
    l->tilesX = (width + l->tileW - 1u) / l->tileW;
    l->tilesY = (height + l->tileH - 1u) / l->tileH;
    l->rank = rank;
    l->worldSize = worldSize ? worldSize : 1u;
    l->localTileCount = 0;
    for (uint32_t ty = 0; ty < l->tilesY; ty++)
        for (uint32_t tx = 0; tx < l->tilesX; tx++)
            if (vkrt_tile_owner(l, tx, ty) == rank) l->localTileCount++;
}

/* out[localTile] = global tile id (ty * tilesX + tx), ascending */
static inline void vkrt_tile_layout_local_tiles(const vkrt_tile_layout* l, uint32_t* out) {
    uint32_t k = 0;
    for (uint32_t ty = 0; ty < l->tilesY; ty++)
        for (uint32_t tx = 0; tx < l->tilesX; tx++)
            if (vkrt_tile_owner(l, tx, ty) == l->rank) out[k++] = ty * l->tilesX + tx;
}

#ifdef __cplusplus
}
#endif
#endif
