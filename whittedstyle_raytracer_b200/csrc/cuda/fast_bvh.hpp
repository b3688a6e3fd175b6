// fast_bvh.hpp — host-side build of the traversal tree the kernels actually walk.
//
// Why a second tree is legal ("Equivalence of acceleration structures").
// In the reference a primitive is tested iff the ray hits the box of every node on
// the path root -> leaf (BVH.hpp:141, BVHStrategy.hpp:27).  Every ancestor box is the
// exact fmin/fmax union of its children (BVH.hpp:67,121), so it CONTAINS the leaf's
// own box, and BoundBox::IntersectRay is monotonic under box inclusion for any ray
// whose direction has no zero component (round-to-nearest subtraction and
// multiplication are monotonic; no NaN can arise without a 0 * inf).  Hence
//     "all ancestor boxes hit"  <=>  "the primitive's own box is hit",
// and the set of primitives the reference tests — therefore the closest hit (ties:
// smaller reference DFS rank), the any-hit result and the hard-shadow product set —
// does not depend on the tree's topology.  Any binary tree over the same per-primitive
// boxes with exact-union inner boxes yields bit-identical results.  Rays with a zero
// direction component (axis-parallel; 1/0 = inf, 0*inf = NaN paths of BoundBox.hpp:55-84)
// are routed to the reference-topology tree instead.
//
// The reference's tree (median split of a centroid sort, BVH.hpp:83-114) mixes the four
// 40-unit wall/floor triangles with 4968 centimetre-sized bunny triangles, so most
// inner boxes near the root span the whole scene.  This tree is built with the
// surface-area heuristic (exact sweep for small ranges, 64 bins above), one primitive
// per leaf, same 32-byte record layout.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../../include/wrt_scene.h"

namespace wrt {

struct FastBvhBuilder {
    struct P { float mn[3], mx[3], c[3]; int prim; };
    std::vector<P> prims;
    std::vector<WrtNode> nodes;
    int max_depth = 0;

    static float half_area(const float* mn, const float* mx) {
        float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
        return dx * dy + dy * dz + dz * dx;
    }
    static void grow(float* mn, float* mx, const P& p) {
        for (int k = 0; k < 3; k++) { mn[k] = fminf(mn[k], p.mn[k]); mx[k] = fmaxf(mx[k], p.mx[k]); }
    }
    static void reset(float* mn, float* mx) {
        for (int k = 0; k < 3; k++) { mn[k] = INFINITY; mx[k] = -INFINITY; }
    }

    // Leaf boxes are taken from the reference tree's leaf records: prim_box[p] for every prim.
    void build(const WrtSceneDesc* s) {
        nodes.clear();
        max_depth = 0;
        const int n = s->n_prims;
        if (n == 0 || s->n_nodes == 0) return;
        prims.resize(n);
        for (int i = 0; i < s->n_nodes; i++) {
            const WrtNode& nd = s->nodes[i];
            if (nd.link >= 0 || i == 1) continue;            // record 1 is padding
            if (i == 1) continue;
            int p = ~nd.link;
            if (p < 0 || p >= n) continue;
            P& q = prims[p];
            q.prim = p;
            for (int k = 0; k < 3; k++) { q.mn[k] = nd.pmin[k]; q.mx[k] = nd.pmax[k]; q.c[k] = 0.5f * nd.pmin[k] + 0.5f * nd.pmax[k]; }
        }
        nodes.resize(2);
        memset(nodes.data(), 0, 2 * sizeof(WrtNode));
        nodes[1].link = ~0;
        rec_build(0, 0, n, 0);
    }

    void set_leaf(int rec, const P& p) {
        WrtNode& nd = nodes[rec];
        for (int k = 0; k < 3; k++) { nd.pmin[k] = p.mn[k]; nd.pmax[k] = p.mx[k]; }
        nd.link = ~p.prim;
    }

    // chooses the split of prims[b,e); returns mid in (b,e) after partitioning
    int split(int b, int e) {
        const int n = e - b;
        if (n == 2) return b + 1;
        float cmn[3], cmx[3];
        reset(cmn, cmx);
        for (int i = b; i < e; i++)
            for (int k = 0; k < 3; k++) { cmn[k] = fminf(cmn[k], prims[i].c[k]); cmx[k] = fmaxf(cmx[k], prims[i].c[k]); }
        float best_cost = INFINITY;
        int best_axis = -1, best_pos = -1;
        if (n <= 256) {                                       // exact sweep
            std::vector<float> right_area(n);
            for (int axis = 0; axis < 3; axis++) {
                if (!(cmx[axis] > cmn[axis])) continue;
                std::sort(prims.begin() + b, prims.begin() + e, [axis](const P& x, const P& y) {
                    return x.c[axis] < y.c[axis] || (x.c[axis] == y.c[axis] && x.prim < y.prim);
                });
                float mn[3], mx[3];
                reset(mn, mx);
                for (int i = n - 1; i > 0; i--) { grow(mn, mx, prims[b + i]); right_area[i] = half_area(mn, mx); }
                reset(mn, mx);
                for (int i = 1; i < n; i++) {
                    grow(mn, mx, prims[b + i - 1]);
                    float cost = half_area(mn, mx) * i + right_area[i] * (n - i);
                    if (cost < best_cost) { best_cost = cost; best_axis = axis; best_pos = i; }
                }
            }
            if (best_axis < 0) return b + n / 2;
            const int axis = best_axis;
            std::sort(prims.begin() + b, prims.begin() + e, [axis](const P& x, const P& y) {
                return x.c[axis] < y.c[axis] || (x.c[axis] == y.c[axis] && x.prim < y.prim);
            });
            return b + best_pos;
        }
        const int K = 64;
        for (int axis = 0; axis < 3; axis++) {
            if (!(cmx[axis] > cmn[axis])) continue;
            int cnt[K] = {0};
            float bmn[K][3], bmx[K][3];
            for (int j = 0; j < K; j++) reset(bmn[j], bmx[j]);
            const float scale = K / (cmx[axis] - cmn[axis]);
            for (int i = b; i < e; i++) {
                int j = std::min(K - 1, std::max(0, (int)((prims[i].c[axis] - cmn[axis]) * scale)));
                cnt[j]++;
                grow(bmn[j], bmx[j], prims[i]);
            }
            float ra[K];
            int rc[K];
            float mn[3], mx[3];
            reset(mn, mx);
            int c = 0;
            for (int j = K - 1; j > 0; j--) {
                if (cnt[j]) { for (int k = 0; k < 3; k++) { mn[k] = fminf(mn[k], bmn[j][k]); mx[k] = fmaxf(mx[k], bmx[j][k]); } }
                c += cnt[j];
                ra[j] = c ? half_area(mn, mx) : 0.f;
                rc[j] = c;
            }
            reset(mn, mx);
            c = 0;
            for (int j = 1; j < K; j++) {
                if (cnt[j - 1]) { for (int k = 0; k < 3; k++) { mn[k] = fminf(mn[k], bmn[j - 1][k]); mx[k] = fmaxf(mx[k], bmx[j - 1][k]); } }
                c += cnt[j - 1];
                if (c == 0 || rc[j] == 0) continue;
                float cost = half_area(mn, mx) * c + ra[j] * rc[j];
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_pos = j; }
            }
        }
        if (best_axis < 0) return b + n / 2;
        const int axis = best_axis;
        const float scale = K / (cmx[axis] - cmn[axis]);
        const float lo = cmn[axis];
        const int pos = best_pos;
        auto mid = std::partition(prims.begin() + b, prims.begin() + e, [=](const P& x) {
            int j = std::min(K - 1, std::max(0, (int)((x.c[axis] - lo) * scale)));
            return j < pos;
        });
        int m = (int)(mid - prims.begin());
        if (m == b || m == e) return b + n / 2;
        return m;
    }

    void rec_build(int rec, int b, int e, int depth) {
        max_depth = std::max(max_depth, depth);
        if (e - b == 1) { set_leaf(rec, prims[b]); return; }
        int mid = split(b, e);
        int pair = (int)nodes.size();
        nodes.resize(pair + 2);
        memset(&nodes[pair], 0, 2 * sizeof(WrtNode));
        nodes[rec].link = pair;
        rec_build(pair, b, mid, depth + 1);
        rec_build(pair + 1, mid, e, depth + 1);
        const WrtNode L = nodes[pair], R = nodes[pair + 1];
        WrtNode& nd = nodes[rec];
        for (int k = 0; k < 3; k++) { nd.pmin[k] = fminf(L.pmin[k], R.pmin[k]); nd.pmax[k] = fmaxf(L.pmax[k], R.pmax[k]); }
    }
};

} // namespace wrt
