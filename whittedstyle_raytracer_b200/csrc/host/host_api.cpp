// host_api.cpp — C ABI over the host front-end (see include/wrt_host.h).
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <memory>
#include <thread>
#include <vector>

#include "../../../include/wrt_host.h"
#include "../../../include/wrt_tiles.h"
#include "host_scene.hpp"

struct WrtScene {
    wrt::HostScene hs;
    std::string out_name;
};

namespace {
thread_local std::string g_err;

int finish_load(std::unique_ptr<WrtScene>& s, const char* obj_path, int variant, WrtScene** out) {
    if (obj_path && *obj_path)
        s->hs.loadObjLikeMain(obj_path, wrt::main_cpp_material(variant == WRT_MATERIAL_GLASS));
    s->hs.buildAndFlatten();
    s->out_name = s->hs.outputName();
    *out = s.release();
    return 0;
}
} // namespace

extern "C" {

const char* wrt_host_last_error(void) { return g_err.c_str(); }

int wrt_scene_load(const char* config_path, const char* obj_path, const char* asset_dir, int variant, WrtScene** out) {
    try {
        std::unique_ptr<WrtScene> s(new WrtScene());
        if (asset_dir) s->hs.assetDir = asset_dir;
        s->hs.parseConfigFile(config_path ? config_path : "");
        return finish_load(s, obj_path, variant, out);
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

int wrt_scene_load_text(const char* text, const char* obj_path, const char* asset_dir, int variant, WrtScene** out) {
    try {
        std::unique_ptr<WrtScene> s(new WrtScene());
        if (asset_dir) s->hs.assetDir = asset_dir;
        s->hs.inputName = "scene.txt";
        s->hs.parseConfigText(text ? text : "");
        return finish_load(s, obj_path, variant, out);
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

void wrt_scene_free(WrtScene* s) { delete s; }

const WrtSceneDesc* wrt_scene_desc(const WrtScene* s) { return &s->hs.desc; }
const WrtCamera* wrt_scene_camera(const WrtScene* s) { return &s->hs.cam; }

int wrt_scene_set_imsize(WrtScene* s, int width, int height) {
    if (width < 0 || height < 0) { g_err = "imsize must be non-negative"; return 1; }
    s->hs.width = width;
    s->hs.height = height;
    s->hs.cam = s->hs.camera();
    return 0;
}

int wrt_scene_set_shadow_type(WrtScene* s, int soft) {
    s->hs.shadowType = soft ? 1 : 0;
    s->hs.desc.shadow_type = s->hs.shadowType;
    return 0;
}

int wrt_scene_bvh_depth(const WrtScene* s) { return s->hs.bvh_depth; }

int64_t wrt_scene_upload_bytes(const WrtScene* s) {
    const wrt::HostScene& h = s->hs;
    return (int64_t)(h.nodes.size() * sizeof(WrtNode) + h.prim_geom.size() * 4 + h.prim_normals.size() * 4 +
                     h.prim_uv.size() * 4 + h.prim_flags.size() * 4 * 6 + h.materials.size() * sizeof(WrtMaterial) +
                     h.lights.size() * sizeof(WrtLight) + h.texels.size() * 4 +
                     h.prim_flags.size() * 16);     // + the primitives' path codes wrt_upload_scene derives and stages (shadow_assoc.h)
}

const char* wrt_scene_output_name(const WrtScene* s) { return s->out_name.c_str(); }

int64_t wrt_tile_slot_count(int width, int height, int tile_w, int tile_h, int rank, int world) {
    if (width < 0 || height < 0 || tile_w <= 0 || tile_h <= 0 || tile_w % 8 || tile_h % 4 || world < 1 || rank < 0 || rank >= world)
        return -1;
    WrtTileMap tm = wrt_tilemap_make(width, height, tile_w, tile_h, rank, world);
    return wrt_tilemap_slots(&tm, rank, world);
}

int wrt_tile_pixel_map(int width, int height, int tile_w, int tile_h, int rank, int world, int64_t* out, int64_t capacity) {
    int64_t n = wrt_tile_slot_count(width, height, tile_w, tile_h, rank, world);
    if (n < 0) { g_err = "wrt_tile_pixel_map: bad tile geometry"; return 1; }
    if (capacity < n) { g_err = "wrt_tile_pixel_map: output too small"; return 1; }
    WrtTileMap tm = wrt_tilemap_make(width, height, tile_w, tile_h, rank, world);
    for (int64_t s = 0; s < n; s++) {
        int x, y;
        out[s] = wrt_tilemap_slot_to_pixel(&tm, s, rank, &x, &y) ? (int64_t)y * width + x : -1;
    }
    return 0;
}

int wrt_scatter_tiles_host(int width, int height, int tile_w, int tile_h, int world, const uint8_t* gathered,
                           int64_t stride_bytes, uint8_t* rgb_image) {
    if (wrt_tile_slot_count(width, height, tile_w, tile_h, 0, world) < 0) { g_err = "wrt_scatter_tiles_host: bad tile geometry"; return 1; }
    for (int r = 0; r < world; r++) {
        WrtTileMap tm = wrt_tilemap_make(width, height, tile_w, tile_h, r, world);
        int64_t n = wrt_tilemap_slots(&tm, r, world);
        if (n * 3 > stride_bytes) { g_err = "wrt_scatter_tiles_host: stride smaller than a rank's tile buffer"; return 1; }
        for (int64_t s = 0; s < n; s++) {
            int x, y;
            if (!wrt_tilemap_slot_to_pixel(&tm, s, r, &x, &y)) continue;
            const uint8_t* p = gathered + (size_t)r * stride_bytes + 3 * (size_t)s;
            uint8_t* q = rgb_image + 3 * ((size_t)y * width + x);
            q[0] = p[0]; q[1] = p[1]; q[2] = p[2];
        }
    }
    return 0;
}

// PPMGenerator::writeHeader / writePixel (include/PPMGenerator.hpp:631-646):
// "P3\nW\nH\n255\n" then "r g b\n" per pixel, row-major — byte-identical to `fout << int`.
// The reference's timer includes this write (src/main.cpp:59-64); at 4K it is 89 MB of text, so the
// formatter uses a 256-entry digit table and formats row bands on several threads.
int wrt_write_ppm_p3(const char* path, int width, int height, const uint8_t* rgb) {
    FILE* f = fopen(path, "wb");
    if (!f) { g_err = std::string("cannot open ") + path; return 1; }
    struct Lut { char txt[256][4]; unsigned char len[256]; };
    static const Lut lut = [] {
        Lut l;
        for (int v = 0; v < 256; v++) l.len[v] = (unsigned char)snprintf(l.txt[v], 4, "%d", v);
        return l;
    }();
    char hdr[64];
    int hl = snprintf(hdr, sizeof hdr, "P3\n%d\n%d\n255\n", width, height);
    bool ok = fwrite(hdr, 1, (size_t)hl, f) == (size_t)hl;
    const size_t npx = (size_t)width * height;
    const size_t band = (size_t)1 << 20;                        // pixels per task (<= 12 MB of text)
    const unsigned nthreads = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    std::vector<std::vector<char>> bufs(nthreads);
    for (size_t base = 0; base < npx && ok; base += band * nthreads) {
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nthreads; t++) {
            size_t p0 = base + (size_t)t * band, p1 = std::min(npx, p0 + band);
            bufs[t].clear();
            if (p0 >= p1) continue;
            pool.emplace_back([&, t, p0, p1] {
                std::vector<char>& out = bufs[t];
                out.resize((p1 - p0) * 12);
                char* w = out.data();
                for (size_t i = p0; i < p1; i++) {
                    for (int k = 0; k < 3; k++) {
                        unsigned v = rgb[i * 3 + k];
                        memcpy(w, lut.txt[v], 4);               // copies up to 3 digits (+1 byte overwritten next)
                        w += lut.len[v];
                        *w++ = k == 2 ? '\n' : ' ';
                    }
                }
                out.resize((size_t)(w - out.data()));
            });
        }
        for (auto& th : pool) th.join();
        for (unsigned t = 0; t < nthreads && ok; t++)
            if (!bufs[t].empty()) ok = fwrite(bufs[t].data(), 1, bufs[t].size(), f) == bufs[t].size();
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) { g_err = "short write"; return 1; }
    return 0;
}

} // extern "C"
