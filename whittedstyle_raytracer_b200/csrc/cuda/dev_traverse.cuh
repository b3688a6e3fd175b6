// dev_traverse.cuh — device scene layout, ray/box/primitive tests and the three
// BVH traversals of the reference, restated as iterative stack traversals.
//
// HBM layout (all read-only during a frame, L2/L1 resident: ~0.6 MB for the bunny):
//   nodes  float4[2*n_nodes]   record i = {pMin.xyz, link} {pMax.xyz, pad}; sibling
//                              pairs adjacent and 64-byte aligned (one 2x LDG.128 per box)
//   geom   float4[3*n_prims]   triangle: {v0.xyz, radius=0} {E1.xyz, 1-alpha} {E2.xyz, flags}
//                              sphere:   {c.xyz,  radius  } {0,      1-alpha} {0,      flags}
//                              (E1 = v1-v0, E2 = v2-v0: the same subtractions Triangle.hpp:22-23 does)
// Primitive index == depth-first leaf rank, so BVH.hpp:157's "left subtree wins
// ties" is "smaller primitive index wins".
#pragma once
#include "dev_math.cuh"
#include "../../../include/wrt_scene.h"

namespace wrt {

struct DevScene {
    const float4* nodes;
    const float4* geom;
    const float4* attr;       // per prim 4x float4: {n0,uv0.x} {n1,uv0.y} {n2,uv1.x} {uv1.y,uv2.x,uv2.y,0}
    const int4*   ids;        // per prim {material, texture, normalmap, object}
    const int*    object_prim;
    const float4* materials;  // per material 3x float4: {Od.rgb, ka} {Os.rgb, kd} {ks, n, alpha, eta}
    const WrtLight* lights;
    const WrtTexture* textures;
    const WrtTexture* normalmaps;
    const float* texels;
    int n_nodes, n_prims, n_lights, n_point_lights, n_dir_lights;
    int has_light_prims;      // any WRT_PRIM_LIGHT primitive in the scene
    int shadow_type, depth_cueing;
    float bkg[3], eta;
    float dc[3], amin, amax, distmin, distmax;
    float eye[3];
};

struct Ray {
    f3 o, d;
    f3 inv;                   // 1/d, BoundBox.hpp:55 (hoisted: same value at every node)
};

__device__ __forceinline__ Ray make_ray(f3 o, f3 d) {
    Ray r;
    r.o = o; r.d = d;
    r.inv = mk3(1 / d.x, 1 / d.y, 1 / d.z);
    return r;
}

// BoundBox::IntersectRay, BoundBox.hpp:53-85.  The swap-on-negative-direction is
// done by selecting which plane feeds tmin/tmax: identical values, no extra ops.
__device__ __forceinline__ bool slab(const float4 lo, const float4 hi, const Ray& r, float& t_enter) {
    float ax = (lo.x - r.o.x) * r.inv.x, bx = (hi.x - r.o.x) * r.inv.x;
    float ay = (lo.y - r.o.y) * r.inv.y, by = (hi.y - r.o.y) * r.inv.y;
    float az = (lo.z - r.o.z) * r.inv.z, bz = (hi.z - r.o.z) * r.inv.z;
    bool sx = r.d.x < 0, sy = r.d.y < 0, sz = r.d.z < 0;
    float tmin_x = sx ? bx : ax, tmax_x = sx ? ax : bx;
    float tmin_y = sy ? by : ay, tmax_y = sy ? ay : by;
    float tmin_z = sz ? bz : az, tmax_z = sz ? az : bz;
    t_enter = fmaxf(tmin_x, fmaxf(tmin_y, tmin_z));
    float t_exit = fminf(tmax_x, fminf(tmax_y, tmax_z));
    return (t_enter <= t_exit) && (t_exit >= 0);
}

struct PrimHit {
    float t, u, v;            // u,v = barycentric b1,b2 (triangles only)
};

// Triangle::intersect acceptance test, Triangle.hpp:22-41
__device__ __forceinline__ bool tri_test(const float4 A, const float4 B, const float4 C, const Ray& r, PrimHit& h) {
    f3 v0 = mk3(A), E1 = mk3(B), E2 = mk3(C);
    f3 S = r.o - v0;
    f3 S1 = cross(r.d, E2);
    f3 S2 = cross(S, E1);
    float rx = dot(S2, E2), ry = dot(S1, S), rz = dot(S2, r.d);
    float left = 1.0f / dot(S1, E1);
    float t = rx * left, u = ry * left, v = rz * left;
    const float EPS = 0.00001f;
    if (t + EPS > 0 && 1 - u - v + EPS > 0 && u + EPS > 0 && v + EPS > 0) {
        h.t = t; h.u = u; h.v = v;
        return true;
    }
    return false;
}

// Sphere::intersect root selection, Sphere.hpp:25-47,80-90 + solveQuadratic global.hpp:105-125
__device__ __noinline__ bool sphere_test(const float4 A, const Ray& r, PrimHit& h) {
    float cx = A.x, cy = A.y, cz = A.z, radius = A.w;
    float Aq = 1.f;
    float Bq = 2 * (r.d.x * (r.o.x - cx) + r.d.y * (r.o.y - cy) + r.d.z * (r.o.z - cz));
    double dx = (double)(r.o.x - cx), dy = (double)(r.o.y - cy), dz = (double)(r.o.z - cz);
    float Cq = (float)(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)),
                                 -(double)(radius * radius)));
    float disc = Bq * Bq - 4 * Aq * Cq;
    float t1, t2;
    if (disc < 0) { t1 = FLT_MAX; t2 = FLT_MAX; }
    else if (disc == 0) { t1 = (-Bq + sqrtf(disc)) / 2 * Aq; t2 = t1; }
    else { t1 = (-Bq + sqrtf(disc)) / 2 * Aq; t2 = (-Bq - sqrtf(disc)) / 2 * Aq; }
    if (t1 > t2) { float s = t1; t1 = t2; t2 = s; }
    float t;
    if (float_equal(t1, FLT_MAX) && float_equal(t2, FLT_MAX)) return false;
    else if (float_equal(t1, t2)) {
        if (t1 < 0) return false;
        t = t1;
    } else {
        if (t1 > 0 && t2 > 0) t = t1;
        else if (t1 > 0 && t2 < 0) t = t1;
        else if (t1 < 0 && t2 > 0) t = t2;
        else return false;
    }
    h.t = t; h.u = 0.f; h.v = 0.f;
    return true;
}

__device__ __forceinline__ bool prim_test(const DevScene& s, int p, const Ray& r, PrimHit& h, float& one_minus_alpha,
                                          unsigned& flags) {
    const float4* g = s.geom + 3 * (size_t)p;
    float4 A = ldg4(g), B = ldg4(g + 1), C = ldg4(g + 2);
    one_minus_alpha = B.w;
    flags = __float_as_uint(C.w);
    if ((flags & WRT_PRIM_KIND_MASK) == WRT_PRIM_SPHERE) return sphere_test(A, r, h);
    return tri_test(A, B, C, r, h);
}

#define WRT_STACK_DEPTH 40   // median-split tree over N prims is ceil(log2 N) deep; 40 covers any 32-bit N

// Per-thread traversal stack in shared memory, column `tid` of a
// [WRT_STACK_DEPTH][blockDim.x] array: consecutive lanes hit consecutive banks.
struct Stack {
    int* base;
    int stride;
    int sp;
    __device__ __forceinline__ void init(int* smem, int tid, int nthreads) { base = smem + tid; stride = nthreads; sp = 0; }
    __device__ __forceinline__ void push(int v) { base[sp * stride] = v; ++sp; }
    __device__ __forceinline__ int pop() { --sp; return base[sp * stride]; }
    __device__ __forceinline__ bool empty() const { return sp == 0; }
};

struct Closest {
    float t; int prim; float u, v;
};

// getIntersection, BVH.hpp:137-159, over the subtree rooted at record `root`.
// The reference visits every node whose box the ray hits and keeps the minimum
// t, left subtree on ties.  Here: explicit stack, near child first; with
// `prune_rel >= 0` a box whose entry distance exceeds best_t*(1+prune_rel)+prune_rel
// is skipped (its leaves cannot hold a closer hit; the margin absorbs the ulp-level
// disagreement between slab and Moller-Trumbore distances).  prune_rel < 0 keeps
// the reference's exhaustive visit.  Ties still resolve to the smaller DFS rank.
__device__ __forceinline__ Closest closest_hit(const DevScene& s, int root, const Ray& r, Stack& st, float prune_rel) {
    Closest best;
    best.t = FLT_MAX; best.prim = -1; best.u = 0.f; best.v = 0.f;
    float limit = FLT_MAX;    // prune threshold derived from best.t
    auto leaf = [&](int p) {
        PrimHit h; float oma; unsigned fl;
        if (prim_test(s, p, r, h, oma, fl)) {
            // `linter.t <= rinter.t` keeps the left (smaller rank) candidate on ties; a candidate
            // with t == FLT_MAX can never displace the default miss on the left
            if (h.t < best.t || (h.t == best.t && p < best.prim)) {
                best.t = h.t; best.prim = p; best.u = h.u; best.v = h.v;
                if (prune_rel >= 0.f) limit = fabsf(h.t) * prune_rel + prune_rel + h.t;
            }
        }
    };
    float te;
    {
        float4 lo = ldg4(s.nodes + 2 * root), hi = ldg4(s.nodes + 2 * root + 1);
        if (!slab(lo, hi, r, te)) return best;
        int link = __float_as_int(lo.w);
        if (link < 0) { leaf(~link); return best; }
        root = link;
    }
    st.sp = 0;
    int cur = root;           // index of the left record of a sibling pair
    while (true) {
        const float4* n = s.nodes + 2 * cur;
        float4 l0 = ldg4(n), l1 = ldg4(n + 1), r0 = ldg4(n + 2), r1 = ldg4(n + 3);
        float tl, tr;
        bool hl = slab(l0, l1, r, tl), hr = slab(r0, r1, r, tr);
        hl = hl && !(tl > limit);
        hr = hr && !(tr > limit);
        int linkL = __float_as_int(l0.w), linkR = __float_as_int(r0.w);
        if (hl && linkL < 0) { leaf(~linkL); hl = false; }
        if (hr && linkR < 0) { hr = hr && !(tr > limit); if (hr) leaf(~linkR); hr = false; }
        if (hl && hr) {
            bool right_first = tr < tl;
            st.push(right_first ? linkL : linkR);
            cur = right_first ? linkR : linkL;
        } else if (hl) cur = linkL;
        else if (hr) cur = linkR;
        else {
            if (st.empty()) break;
            cur = st.pop();
        }
    }
    return best;
}

// BVHStrategy::ShadowHelper, BVHStrategy.hpp:24-48: product of (1-alpha) over every
// leaf reached through hit boxes whose primitive is hit with t < dis and is not a
// light avatar.  Leaves are visited left to right; the walk stops once the
// product is exactly 0 (0 * x == 0 for the finite factors that follow).
__device__ __forceinline__ float shadow_product(const DevScene& s, const Ray& r, float dis, Stack& st) {
    float res = 1.f;
    if (s.n_nodes == 0) return res;
    auto leaf = [&](int p) {
        PrimHit h; float oma; unsigned fl;
        if (prim_test(s, p, r, h, oma, fl) && h.t < dis && !(fl & WRT_PRIM_LIGHT)) res = res * oma;
    };
    float te;
    int cur;
    {
        float4 lo = ldg4(s.nodes), hi = ldg4(s.nodes + 1);
        if (!slab(lo, hi, r, te)) return res;
        int link = __float_as_int(lo.w);
        if (link < 0) { leaf(~link); return res; }
        cur = link;
    }
    st.sp = 0;
    while (true) {
        const float4* n = s.nodes + 2 * cur;
        float4 l0 = ldg4(n), l1 = ldg4(n + 1), r0 = ldg4(n + 2), r1 = ldg4(n + 3);
        float tl, tr;
        bool hl = slab(l0, l1, r, tl), hr = slab(r0, r1, r, tr);
        int linkL = __float_as_int(l0.w), linkR = __float_as_int(r0.w);
        if (hl && linkL < 0) { leaf(~linkL); hl = false; }
        if (hr && linkR < 0) { leaf(~linkR); hr = false; }
        if (res == 0.f) break;
        if (hl && hr) { st.push(linkR); cur = linkL; }
        else if (hl) cur = linkL;
        else if (hr) cur = linkR;
        else {
            if (st.empty()) break;
            cur = st.pop();
        }
    }
    return res;
}

// hasIntersection, BVH.hpp:162-186.  No root-box test.  The reference takes the
// CLOSEST hit of each root child and calls the ray occluded when that hit has
// t < dis and is not a light avatar.  Without light-avatar primitives this is
// "any primitive hit with t < dis" (the closest hit is < dis iff some hit is),
// which allows an early-out any-hit walk; with light avatars the two per-child
// closest-hit queries are done literally.
__device__ __forceinline__ bool occluded(const DevScene& s, const Ray& r, float dis, Stack& st, float prune_rel) {
    if (s.n_nodes == 0) return false;
    int root_link = __float_as_int(ldg4(s.nodes).w);
    if (root_link < 0) {
        PrimHit h; float oma; unsigned fl;
        // (the reference dereferences a null obj here when the lone primitive is missed)
        if (!prim_test(s, ~root_link, r, h, oma, fl)) return false;
        if (fl & WRT_PRIM_LIGHT) return false;
        return h.t < dis;
    }
    if (s.has_light_prims) {
        Closest a = closest_hit(s, root_link, r, st, prune_rel);
        if (a.prim >= 0 && a.t < dis && !(__float_as_uint(ldg4(s.geom + 3 * (size_t)a.prim + 2).w) & WRT_PRIM_LIGHT))
            return true;
        Closest b = closest_hit(s, root_link + 1, r, st, prune_rel);
        if (b.prim >= 0 && b.t < dis && !(__float_as_uint(ldg4(s.geom + 3 * (size_t)b.prim + 2).w) & WRT_PRIM_LIGHT))
            return true;
        return false;
    }
    bool occ = false;
    auto leaf = [&](int p) {
        PrimHit h; float oma; unsigned fl;
        if (prim_test(s, p, r, h, oma, fl) && h.t < dis) occ = true;
    };
    st.sp = 0;
    int cur = root_link;
    while (true) {
        const float4* n = s.nodes + 2 * cur;
        float4 l0 = ldg4(n), l1 = ldg4(n + 1), r0 = ldg4(n + 2), r1 = ldg4(n + 3);
        float tl, tr;
        bool hl = slab(l0, l1, r, tl), hr = slab(r0, r1, r, tr);
        int linkL = __float_as_int(l0.w), linkR = __float_as_int(r0.w);
        if (hl && linkL < 0) { leaf(~linkL); hl = false; }
        if (hr && linkR < 0) { leaf(~linkR); hr = false; }
        if (occ) break;
        if (hl && hr) {
            bool right_first = tr < tl;
            st.push(right_first ? linkL : linkR);
            cur = right_first ? linkR : linkL;
        } else if (hl) cur = linkL;
        else if (hr) cur = linkR;
        else {
            if (st.empty()) break;
            cur = st.pop();
        }
    }
    return occ;
}

// Renderer::getShadowCoeffi(Intersection&, Vector4f&), Renderer.hpp:381-400: every
// object in objList order, skipping self and light avatars; no distance bound.
__device__ __forceinline__ float directional_product(const DevScene& s, const Ray& r, int self_prim) {
    float res = 1.f;
    for (int k = 0; k < s.n_prims; k++) {
        int p = __ldg(s.object_prim + k);
        if (p == self_prim) continue;
        PrimHit h; float oma; unsigned fl;
        const float4* g = s.geom + 3 * (size_t)p;
        if (__float_as_uint(ldg4(g + 2).w) & WRT_PRIM_LIGHT) continue;
        if (prim_test(s, p, r, h, oma, fl)) res = res * oma;
        if (res == 0.f) break;
    }
    return res;
}

} // namespace wrt
