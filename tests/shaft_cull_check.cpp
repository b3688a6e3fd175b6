// TEST INFRASTRUCTURE (built and run by tests/test_shaft_cull.py).
// Runs the product's shaft test (csrc/cuda/shaft_cull.h — the same source k_surface_spawn compiles)
// on the CPU and checks every "empty" verdict by brute force: for many (u, v) samples, corner
// extremes included, the sample ray — built with the kernel's own float operations — must miss
// the own box of EVERY primitive under the exact BoundBox::IntersectRay arithmetic.  Also checks
// that each sample's 1/d lies inside the shaft's bounds, and that the shaft's candidate list
// (wrt_shaft_candidates, phase 1 of k_shadow_soft_list) contains every primitive whose own box a
// sample ray hits; and that the triangle-level pruning of that list (wrt_pyramid_triangle_may_block, k_soft_filter)
// only removes triangles that no sample ray hits with t < dis under Triangle::intersect's own arithmetic.
// Prints one JSON line.
//
// usage: shaft_cull_check <config.txt> <bunny.obj|-> <asset_dir> <n_requests> <seed>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#include "../include/wrt_host.h"
#include "../include/wrt_rng.h"
#include "../include/wrt_scene.h"
#include "../whittedstyle_raytracer_b200/csrc/cuda/fast_bvh.hpp"
#include "../whittedstyle_raytracer_b200/csrc/cuda/shaft_cull.h"

struct V { float x, y, z; };

static uint64_t g_state;
static uint32_t rnd() { g_state = g_state * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(g_state >> 32); }
static float unif() { return (float)(rnd() >> 8) * (1.0f / 16777216.0f); }

// BoundBox::IntersectRay (BoundBox.hpp:53-85) on a natural-order box
static bool slab(const float* mn, const float* mx, V o, V d, V inv) {
    float ax = (mn[0] - o.x) * inv.x, bx = (mx[0] - o.x) * inv.x;
    float ay = (mn[1] - o.y) * inv.y, by = (mx[1] - o.y) * inv.y;
    float az = (mn[2] - o.z) * inv.z, bz = (mx[2] - o.z) * inv.z;
    float tminx = d.x < 0 ? bx : ax, tmaxx = d.x < 0 ? ax : bx;
    float tminy = d.y < 0 ? by : ay, tmaxy = d.y < 0 ? ay : by;
    float tminz = d.z < 0 ? bz : az, tmaxz = d.z < 0 ? az : bz;
    float te = fmaxf(tminx, fmaxf(tminy, tminz)), tx = fminf(tmaxx, fminf(tmaxy, tmaxz));
    return te <= tx && tx >= 0;
}

int main(int argc, char** argv) {
    if (argc < 6) { fprintf(stderr, "usage\n"); return 2; }
    WrtScene* sc = nullptr;
    if (wrt_scene_load(argv[1], strcmp(argv[2], "-") ? argv[2] : nullptr, argv[3], 0, &sc) != 0) {
        fprintf(stderr, "load failed: %s\n", wrt_host_last_error());
        return 2;
    }
    const WrtSceneDesc* S = wrt_scene_desc(sc);
    const int n_req = atoi(argv[4]);
    g_state = strtoull(argv[5], nullptr, 10) * 2654435761ull + 12345;
    wrt::FastBvhBuilder fb;
    fb.build(S);
    std::vector<WrtNode> oct = fb.octant_copies();
    std::vector<float4> onodes(2 * oct.size());
    static_assert(sizeof(WrtNode) == 2 * sizeof(float4), "record = two float4");
    memcpy(onodes.data(), oct.data(), oct.size() * sizeof(WrtNode));
    const int np = S->n_prims, nn = (int)fb.nodes.size();
    // the 4-wide view of every octant copy, built by the product's own routine (wide_bvh.h, what k_wide4_copies runs)
    std::vector<float4> wnodes((size_t)WRT_WIDE_FLOAT4_PER_RECORD * 8 * (size_t)nn);
    for (int oct = 0; oct < 8; oct++)
        for (int c = 2; c + 1 < nn; c += 2)
            wrt_wide4_node(onodes.data() + 2 * (size_t)oct * nn, c, oct, wnodes.data() + WRT_WIDE_FLOAT4_PER_RECORD * ((size_t)oct * nn + c));
    long long wide_mismatch = 0;
    std::vector<float> box(6 * (size_t)np);
    float smin[3] = {INFINITY, INFINITY, INFINITY}, smax[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = 0; i < S->n_nodes; i++) {
        const WrtNode& nd = S->nodes[i];
        if (nd.link >= 0 || i == 1) continue;
        int p = ~nd.link;
        for (int k = 0; k < 3; k++) {
            box[6 * (size_t)p + k] = nd.pmin[k]; box[6 * (size_t)p + 3 + k] = nd.pmax[k];
            smin[k] = fminf(smin[k], nd.pmin[k]); smax[k] = fmaxf(smax[k], nd.pmax[k]);
        }
    }
    long long empty = 0, nonempty = 0, gave_up = 0, violations = 0, bound_violations = 0, rays_checked = 0, degenerate_rays = 0,
              lists = 0, list_items = 0, list_overflow = 0, list_violations = 0, list_rays = 0,
              filter_removed = 0, filter_kept = 0, filter_pairs = 0, filter_violations = 0;
    const float top = 1.0f - 1.0f / 16777216.0f;
    for (int r = 0; r < n_req; r++) {
        // origin: a point on a random primitive pushed off along +-normal like BVHStrategy.hpp:15, or a free point
        float o[3];
        int mode = rnd() % 4;
        if (mode < 3 && np > 0) {
            int p = rnd() % np;
            const float* g = S->prim_geom + 12 * (size_t)p;
            if ((S->prim_flags[p] & WRT_PRIM_KIND_MASK) == WRT_PRIM_TRIANGLE) {
                float a = unif(), b = unif();
                if (a + b > 1) { a = 1 - a; b = 1 - b; }
                V e1{g[4], g[5], g[6]}, e2{g[8], g[9], g[10]};
                V n{e1.y * e2.z - e1.z * e2.y, e1.z * e2.x - e1.x * e2.z, e1.x * e2.y - e1.y * e2.x};
                float len = sqrtf(n.x * n.x + n.y * n.y + n.z * n.z);
                float sgn = (rnd() & 1) ? 0.0005f : -0.0005f;
                if (len > 0) { n.x /= len; n.y /= len; n.z /= len; }
                o[0] = g[0] + a * e1.x + b * e2.x + sgn * n.x;
                o[1] = g[1] + a * e1.y + b * e2.y + sgn * n.y;
                o[2] = g[2] + a * e1.z + b * e2.z + sgn * n.z;
            } else {
                V n{unif() - 0.5f, unif() - 0.5f, unif() - 0.5f};
                float len = sqrtf(n.x * n.x + n.y * n.y + n.z * n.z) + 1e-20f;
                float rr = g[3] + 0.0005f;
                o[0] = g[0] + rr * n.x / len; o[1] = g[1] + rr * n.y / len; o[2] = g[2] + rr * n.z / len;
            }
        } else {
            for (int k = 0; k < 3; k++) o[k] = smin[k] + (smax[k] - smin[k]) * (unif() * 1.4f - 0.2f);
        }
        // a quarter of the origins share a coordinate with a light corner: that axis is dropped by the shaft
        // test, and when the coordinate is 0 every sample is axis-degenerate there (d_k == 0: the inf / NaN
        // paths of BoundBox::IntersectRay, which slab() below takes exactly like the kernels do)
        if (S->n_lights > 0 && rnd() % 4 == 0) {
            int k = rnd() % 3;
            o[k] = S->lights[rnd() % S->n_lights].tri[k + 3 * (rnd() % 3)];
        }
        for (int li = 0; li < S->n_lights; li++) {
            const WrtLight& L = S->lights[li];
            if (fabsf(L.pos[3] - 1.f) >= 0.00001f) continue;
            WrtShaft sh;
            const bool made = wrt_shaft_make(o, L.tri, &sh);
            const bool is_empty = wrt_shaft_is_empty(onodes.data(), nn, o, L.tri);
            if (!made) { gave_up++; if (is_empty) violations++; continue; }
            if (is_empty) empty++; else nonempty++;
            // candidate list of the same shaft (k_shadow_soft_list, phase 1)
            int stack[64], list[64];
            const int cnt = wrt_shaft_candidates(onodes.data(), nn, &sh, stack, 1, 64, list, 64);
            if (cnt < 0) list_overflow++; else { lists++; list_items += cnt; if (is_empty != (cnt == 0)) violations++; }
            {   // the wide walks must reach exactly the same leaves (their order may differ)
                if (wrt_shaft_is_empty4(onodes.data(), wnodes.data(), nn, o, L.tri) != is_empty) wide_mismatch++;
                // (the straight-line step checks the capacities once per step and gives up when fewer than 4 list / 3 stack
                // entries are free: 4 spare entries make its verdict on lists of up to 64 the binary walk's)
                int stack4[96], list4[68];
                int cnt4 = wrt_shaft_candidates4(onodes.data(), wnodes.data(), nn, &sh, stack4, 1, 96, list4, 68);
                if (cnt4 > 64) cnt4 = -1;
                if (cnt4 != cnt) wide_mismatch++;
                else for (int a = 0; a < cnt; a++) {
                    bool found = false;
                    for (int b = 0; b < cnt4; b++) found = found || list4[b] == list[a];
                    if (!found) wide_mismatch++;
                }
            }
            // triangle-level pruning of the list (k_soft_filter): the removed candidates must block no sample ray
            std::vector<int> removed;
            if (cnt > 0) {
                WrtShaftPyramid py;
                wrt_pyramid_make(o, L.tri, &py);
                for (int c = 0; c < cnt; c++) {
                    const int p = list[c];
                    const float* g = S->prim_geom + 12 * (size_t)p;
                    float aux[4];
                    wrt_triangle_aux(g + 4, g + 8, aux);                              // (k_pack_prims computes this at upload)
                    if ((S->prim_flags[p] & WRT_PRIM_KIND_MASK) != WRT_PRIM_TRIANGLE ||
                        (wrt_pyramid_triangle_may_block(&py, g, g + 4, g + 8, aux) && wrt_pyramid_triangle_may_block_edges(&py, g, g + 4, g + 8, aux))) filter_kept++;   // both stages of k_soft_filter
                    else { removed.push_back(p); filter_removed++; }
                }
            }
            V v0{L.tri[0], L.tri[1], L.tri[2]}, v1{L.tri[3], L.tri[4], L.tri[5]}, v2{L.tri[6], L.tri[7], L.tri[8]};
            const int ns = is_empty ? 40 : (removed.empty() ? 12 : 24);
            for (int s = 0; s < ns; s++) {
                float u, v;
                if (s < 4) { u = (s & 1) ? top : 0.f; v = (s & 2) ? top : 0.f; }
                else if (s < 8) { u = (s & 1) ? top : 0.f; v = unif(); if (s & 2) { float t = u; u = v; v = t; } }
                else wrt_light_sample_uv(WRT_DEFAULT_SEED, rnd(), 1u + rnd() % 511u, (uint32_t)li, (uint32_t)s, &u, &v);
                // SoftShadowQuery::begin (kernels.cuh) / area_light_shadow (oracle): same operations, same order
                float a = 1 - u - v;
                V lp{a * v0.x + u * v1.x + v * v2.x, a * v0.y + u * v1.y + v * v2.y, a * v0.z + u * v1.z + v * v2.z};
                V dl{lp.x - o[0], lp.y - o[1], lp.z - o[2]};
                float mag = sqrtf(dl.x * dl.x + dl.y * dl.y + dl.z * dl.z);
                V d = dl;
                if (mag > 0) { float mi = 1 / mag; d = V{dl.x * mi, dl.y * mi, dl.z * mi}; }
                V inv{1 / d.x, 1 / d.y, 1 / d.z};
                const float iv[3] = {inv.x, inv.y, inv.z}, dv[3] = {d.x, d.y, d.z};
                for (int k = 0; k < 3; k++) {
                    if (!((sh.use >> k) & 1)) continue;                   // dropped axis: nothing is claimed about it
                    bool neg = dv[k] < 0;
                    if (!(iv[k] >= sh.ilo[k] && iv[k] <= sh.ihi[k]) || neg != (((sh.octant >> k) & 1) != 0) || dv[k] == 0) bound_violations++;
                }
                const bool degenerate = dv[0] == 0 || dv[1] == 0 || dv[2] == 0;
                if (degenerate) degenerate_rays++;
                for (int p : removed) {                                  // Triangle::intersect, Triangle.hpp:22-41, and t < dis
                    const float* g = S->prim_geom + 12 * (size_t)p;
                    V tv0{g[0], g[1], g[2]}, E1{g[4], g[5], g[6]}, E2{g[8], g[9], g[10]};
                    V Sv{o[0] - tv0.x, o[1] - tv0.y, o[2] - tv0.z};
                    V S1{d.y * E2.z - d.z * E2.y, d.z * E2.x - d.x * E2.z, d.x * E2.y - d.y * E2.x};
                    V S2{Sv.y * E1.z - Sv.z * E1.y, Sv.z * E1.x - Sv.x * E1.z, Sv.x * E1.y - Sv.y * E1.x};
                    float rx = S2.x * E2.x + S2.y * E2.y + S2.z * E2.z, ry = S1.x * Sv.x + S1.y * Sv.y + S1.z * Sv.z, rz = S2.x * d.x + S2.y * d.y + S2.z * d.z;
                    float left = 1.0f / (S1.x * E1.x + S1.y * E1.y + S1.z * E1.z);
                    float t = rx * left, bu = ry * left, bv = rz * left;
                    const float EPS = 0.00001f;
                    filter_pairs++;
                    if (t + EPS > 0 && 1 - bu - bv + EPS > 0 && bu + EPS > 0 && bv + EPS > 0 && t < mag) filter_violations++;
                }
                if (cnt >= 0 && !degenerate) {
                    // every primitive whose own box this ray hits must be on the list
                    V ov{o[0], o[1], o[2]};
                    for (int p = 0; p < np; p++) {
                        if (!slab(&box[6 * (size_t)p], &box[6 * (size_t)p + 3], ov, d, inv)) continue;
                        bool found = false;
                        for (int c = 0; c < cnt; c++) found = found || list[c] == p;
                        if (!found) list_violations++;
                    }
                    list_rays++;
                }
                if (!is_empty) continue;
                rays_checked++;
                V ov{o[0], o[1], o[2]};
                for (int p = 0; p < np; p++)
                    if (slab(&box[6 * (size_t)p], &box[6 * (size_t)p + 3], ov, d, inv)) { violations++; break; }
            }
        }
    }
    printf("{\"prims\": %d, \"empty\": %lld, \"nonempty\": %lld, \"gave_up\": %lld, \"rays_checked\": %lld, "
           "\"degenerate_rays\": %lld, \"violations\": %lld, \"bound_violations\": %lld, \"lists\": %lld, \"list_items\": %lld, "
           "\"list_overflow\": %lld, \"list_rays\": %lld, \"list_violations\": %lld, \"filter_removed\": %lld, \"filter_kept\": %lld, "
           "\"filter_pairs\": %lld, \"filter_violations\": %lld, \"wide_mismatch\": %lld}\n", np, empty, nonempty, gave_up, rays_checked,
           degenerate_rays, violations, bound_violations, lists, list_items, list_overflow, list_rays, list_violations, filter_removed,
           filter_kept, filter_pairs, filter_violations, wide_mismatch);
    wrt_scene_free(sc);
    return (violations || bound_violations || list_violations || filter_violations || wide_mismatch) ? 1 : 0;
}
