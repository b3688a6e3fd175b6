// ploc_bvh.h — device-side build of the traversal tree the kernels walk (replaces the host
// build of BVH.hpp:49-125 for everything but the reference-topology tree that axis-degenerate rays need).
//
// Any binary tree over the reference's per-primitive boxes with exact-union inner boxes gives bit-identical results
// (fast_bvh.hpp / DESIGN.md section 4), so the build is free to be whatever maps to the GPU.  This is PLOC —
// parallel locally-ordered clustering (Meister & Bittner 2018): sort the primitives along a Morton curve, then repeat
//     1. every cluster i looks at its `radius` neighbours on either side in the sorted order and picks the one whose
//        union box with it has the smallest surface area (ties: smaller index),
//     2. mutual nearest neighbours merge into a new inner node,
//     3. the surviving clusters are compacted, order preserved,
// until one cluster is left.  Every step is a data-parallel pass over an array; ~30 passes for the bunny.  The tree's
// surface-area cost is within 1 % (bunny) to 11 % (sphere field) of the host's binned-SAH build, measured by
// tests/ploc_check.cpp, and one primitive per leaf keeps the tie-break ranks (primitive index) defined.
// Progress: neighbours are ranked by a strict total order on pairs (ploc_nearest), so the globally smallest in-window
// pair is always mutual and every pass merges at least one pair.
//
// The per-item steps below are __host__ __device__: tests/ploc_check.cpp runs the SAME source on the CPU (passes
// emulated by loops) and checks the tree — every primitive exactly once, exact-union boxes, pair-adjacent layout,
// depth, cost.  The pass orchestration (grid barriers, ordered compaction) lives in bvh_build.cuh.
#ifndef WRT_PLOC_BVH_H
#define WRT_PLOC_BVH_H

#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define WRT_PLOC_HD __host__ __device__ __forceinline__
#else
#include <vector_types.h>
#include <vector_functions.h>
#define WRT_PLOC_HD static inline
#endif

#define WRT_PLOC_RADIUS 8

// Build tree: nodes 0 .. n-1 are the leaves (node id == primitive index), inner nodes are appended from n on.
struct PlocTree {
    float4* lo;        // {box min, left child  | -1}
    float4* hi;        // {box max, right child | -1}
    float4* dlo;       // dilated box (directional-shadow culling tree): {min, -}
    float4* dhi;       //                                               {max, -}
    int* parent;       // -1 for the root
    int* cnt;          // leaves below
};

WRT_PLOC_HD float ploc_fmin(float a, float b) { return a < b ? a : b; }
WRT_PLOC_HD float ploc_fmax(float a, float b) { return a > b ? a : b; }
WRT_PLOC_HD int ploc_f2i(float f) { union { float f; int i; } u; u.f = f; return u.i; }
WRT_PLOC_HD float ploc_i2f(int i) { union { float f; int i; } u; u.i = i; return u.f; }

// 63-bit Morton code of a point inside [bmin, bmax] (21 bits per axis).
WRT_PLOC_HD uint64_t ploc_expand21(uint64_t v) {
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}
WRT_PLOC_HD uint64_t ploc_morton(const float c[3], const float bmin[3], const float bmax[3]) {
    uint64_t code = 0;
    for (int k = 0; k < 3; k++) {
        float ext = bmax[k] - bmin[k];
        float t = ext > 0.f ? (c[k] - bmin[k]) / ext : 0.f;
        if (!(t > 0.f)) t = 0.f;                       // NaN and negatives
        if (t > 1.f) t = 1.f;
        uint64_t q = (uint64_t)(t * 2097151.0f);
        code |= ploc_expand21(q) << k;
    }
    return code;
}

// Leaf `p`: its own box (the reference's, from the flattened reference tree), the dilated copy
// (fast_bvh.hpp `dilated`: rel * max(extent, 1e-2 * |coordinate|) + abs on every side).
WRT_PLOC_HD void ploc_init_leaf(const PlocTree& t, int p, const float mn[3], const float mx[3], float dil_rel, float dil_abs) {
    t.lo[p] = make_float4(mn[0], mn[1], mn[2], ploc_i2f(-1));
    t.hi[p] = make_float4(mx[0], mx[1], mx[2], ploc_i2f(-1));
    float ext = 0.f;
    for (int k = 0; k < 3; k++) {
        ext = ploc_fmax(ext, mx[k] - mn[k]);
        ext = ploc_fmax(ext, ploc_fmax(fabsf(mn[k]), fabsf(mx[k])) * 1e-2f);
    }
    const float pad = dil_rel * ext + dil_abs;
    t.dlo[p] = make_float4(mn[0] - pad, mn[1] - pad, mn[2] - pad, 0.f);
    t.dhi[p] = make_float4(mx[0] + pad, mx[1] + pad, mx[2] + pad, 0.f);
    t.parent[p] = -1;
    t.cnt[p] = 1;
}

// half the surface area of the union of two boxes (symmetric in its arguments, bit for bit)
WRT_PLOC_HD float ploc_union_area(const float4 alo, const float4 ahi, const float4 blo, const float4 bhi) {
    float dx = ploc_fmax(ahi.x, bhi.x) - ploc_fmin(alo.x, blo.x);
    float dy = ploc_fmax(ahi.y, bhi.y) - ploc_fmin(alo.y, blo.y);
    float dz = ploc_fmax(ahi.z, bhi.z) - ploc_fmin(alo.z, blo.z);
    return dx * dy + dy * dz + dz * dx;
}

// Step 1: nearest neighbour of cluster i among cl[i-radius .. i+radius] (index into cl, -1 when alone).
// "Nearest" minimises a key that is a strict total order on unordered PAIRS and symmetric in the pair:
//     (union area, not buddies, |i - j|, min(i, j))        buddies: j == i ^ 1
// so the globally smallest pair is always mutual (progress), and equal areas — coincident or duplicated geometry, where
// every neighbour is equally near — pair up as (0,1)(2,3)... and halve the list per pass instead of merging one pair
// per pass.  A non-finite area counts as +inf.
WRT_PLOC_HD bool ploc_tie_before(int i, int j, int k) {       // pair (i, j) before pair (i, k) at equal area?
    const bool jb = j == (i ^ 1), kb = k == (i ^ 1);
    if (jb != kb) return jb;
    const int dj = j > i ? j - i : i - j, dk = k > i ? k - i : i - k;
    if (dj != dk) return dj < dk;
    return (j < i ? j : i) < (k < i ? k : i);
}
WRT_PLOC_HD int ploc_nearest(const PlocTree& t, const int* cl, int m, int i, int radius) {
    const int a = cl[i];
    const float4 alo = t.lo[a], ahi = t.hi[a];
    float best = INFINITY;
    int bj = -1;
    const int j0 = i - radius > 0 ? i - radius : 0, j1 = i + radius < m - 1 ? i + radius : m - 1;
    for (int j = j0; j <= j1; j++) {
        if (j == i) continue;
        const int b = cl[j];
        float d = ploc_union_area(alo, ahi, t.lo[b], t.hi[b]);
        if (!(d == d)) d = INFINITY;
        if (bj < 0 || d < best || (d == best && ploc_tie_before(i, j, bj))) { best = d; bj = j; }
    }
    return bj;
}

// Step 2: what becomes of cluster i.  0 = absorbed by its partner, 1 = stays, 2 = merges with nn[i] (i is the smaller
// index and creates the node).
WRT_PLOC_HD int ploc_fate(const int* nn, int i) {
    const int j = nn[i];
    if (j < 0 || nn[j] != i) return 1;
    return i < j ? 2 : 0;
}

// The new inner node `id` over children a (left) and b (right): exact union boxes (fminf/fmaxf of finite floats are
// exact, so the union is the same whatever the merge order), leaf count, parent links.
WRT_PLOC_HD void ploc_make_node(const PlocTree& t, int id, int a, int b) {
    const float4 alo = t.lo[a], ahi = t.hi[a], blo = t.lo[b], bhi = t.hi[b];
    t.lo[id] = make_float4(ploc_fmin(alo.x, blo.x), ploc_fmin(alo.y, blo.y), ploc_fmin(alo.z, blo.z), ploc_i2f(a));
    t.hi[id] = make_float4(ploc_fmax(ahi.x, bhi.x), ploc_fmax(ahi.y, bhi.y), ploc_fmax(ahi.z, bhi.z), ploc_i2f(b));
    const float4 dal = t.dlo[a], dah = t.dhi[a], dbl = t.dlo[b], dbh = t.dhi[b];
    t.dlo[id] = make_float4(ploc_fmin(dal.x, dbl.x), ploc_fmin(dal.y, dbl.y), ploc_fmin(dal.z, dbl.z), 0.f);
    t.dhi[id] = make_float4(ploc_fmax(dah.x, dbh.x), ploc_fmax(dah.y, dbh.y), ploc_fmax(dah.z, dbh.z), 0.f);
    t.cnt[id] = t.cnt[a] + t.cnt[b];
    t.parent[id] = -1;
    t.parent[a] = id;
    t.parent[b] = id;
}

// Layout: build node v -> its 32-byte record in the walked tree (include/wrt_scene.h WrtNode: sibling pairs adjacent,
// depth-first, record 0 = root, record 1 = padding).  A subtree over k leaves occupies 2k-2 records below its root, so
// with F(v) = first record below v:  F(root) = 2,  F(left child) = F(p) + 2,  F(right child) = F(p) + 2 + 2*cnt(left) - 2,
// and v's own record is F(parent) + (v is the right child).  Every node finds its F by walking up to the root —
// O(depth) each, no ordering between nodes.  Returns the record index; *link = children's pair (inner) or ~primitive
// (leaf); *depth = distance from the root.
WRT_PLOC_HD int ploc_record_of(const PlocTree& t, int n_leaves, int v, int* link, int* depth) {
    int acc = 0, d = 0, own_extra = 0, own_right = 0;
    int u = v;
    while (t.parent[u] >= 0) {
        const int p = t.parent[u];
        const int left = ploc_f2i(t.lo[p].w);
        const int right = u != left;
        const int extra = 2 + (right ? 2 * t.cnt[left] - 2 : 0);
        if (u == v) { own_extra = extra; own_right = right; }
        acc += extra;
        u = p;
        ++d;
    }
    *depth = d;
    const int F = 2 + acc;
    *link = v < n_leaves ? ~v : F;
    return d == 0 ? 0 : F - own_extra + own_right;
}

#endif /* WRT_PLOC_BVH_H */
