/* wrt_host.h — C ABI of the host front-end (libwrt_host.so, no CUDA needed).
 *
 * Replaces, for the render hot path only, what the reference's driver does
 * before and after Renderer::render() (paths relative to /root/reference):
 *   wrt_scene_load      <- PPMGenerator::PPMGenerator + main()'s bunny.obj load   include/PPMGenerator.hpp:31-44, src/main.cpp:20-56
 *                          followed by Scene::initializeBVH (include/Scene.hpp:32-39) and flattening
 *   wrt_scene_desc      <- the Scene&/PPMGenerator* the Renderer reads             include/Renderer.hpp:38-49
 *   wrt_scene_camera    <- camera block of Renderer::render                        include/Renderer.hpp:65-100
 *   wrt_scene_output_name <- output file naming                                    include/PPMGenerator.hpp:62-74
 *   wrt_write_ppm_p3    <- PPMGenerator::writeHeader/writePixel                    include/PPMGenerator.hpp:631-646
 *
 * All functions return 0 on success, non-zero on failure; wrt_host_last_error()
 * then holds the text the reference would have printed after "ERROR:".
 */
#ifndef WRT_HOST_H
#define WRT_HOST_H

#include "wrt_scene.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct WrtScene WrtScene;

#define WRT_MATERIAL_MAIN_CPP 0   /* src/main.cpp:24-32, the active ("water") bunny material */
#define WRT_MATERIAL_GLASS    1   /* src/main.cpp:35-43, the commented-out variant */

/* config_path: scene config file (reference grammar).
 * obj_path:    mesh loaded the way main() loads "bunny.obj" (x20, y-3, z-3, hard-coded
 *              material); NULL or a missing file = no mesh, like the reference.
 * asset_dir:   directory that relative texture paths are resolved against;
 *              NULL = current working directory (reference behaviour). */
int wrt_scene_load(const char* config_path, const char* obj_path, const char* asset_dir,
                   int obj_material_variant, WrtScene** out);
/* Same, from an in-memory config text. */
int wrt_scene_load_text(const char* config_text, const char* obj_path, const char* asset_dir,
                        int obj_material_variant, WrtScene** out);
void wrt_scene_free(WrtScene* s);

const WrtSceneDesc* wrt_scene_desc(const WrtScene* s);
const WrtCamera*    wrt_scene_camera(const WrtScene* s);
/* Recomputes the camera for another image size (the derived 4K / 8K configs). */
int wrt_scene_set_imsize(WrtScene* s, int width, int height);
int wrt_scene_set_shadow_type(WrtScene* s, int soft);
int wrt_scene_bvh_depth(const WrtScene* s);
int64_t wrt_scene_upload_bytes(const WrtScene* s);
const char* wrt_scene_output_name(const WrtScene* s);

/* ASCII P3 writer, byte-identical to the reference's output for 8-bit data. */
int wrt_write_ppm_p3(const char* path, int width, int height, const uint8_t* rgb);

/* Image sharding (include/wrt_tiles.h) for the multi-GPU driver and its CPU tests.
 * wrt_tile_slot_count: slots (tile area incl. padding) rank `rank` of `world` owns.
 * wrt_tile_pixel_map:  out[slot] = y*width + x of that slot's pixel, or -1 for padding.
 * wrt_scatter_tiles_host: CPU twin of the rank-0 scatter kernel — `gathered` holds `world`
 *   tile-order RGB buffers of stride_bytes each; writes the row-major image. */
int64_t wrt_tile_slot_count(int width, int height, int tile_w, int tile_h, int rank, int world);
int wrt_tile_pixel_map(int width, int height, int tile_w, int tile_h, int rank, int world, int64_t* out, int64_t capacity);
int wrt_scatter_tiles_host(int width, int height, int tile_w, int tile_h, int world, const uint8_t* gathered,
                           int64_t stride_bytes, uint8_t* rgb_image);

const char* wrt_host_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* WRT_HOST_H */
