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
#include <thread>
#include <vector>

#include "../../../include/wrt_scene.h"

namespace wrt {

// min/max without libm calls (finite inputs: same values as fminf/fmaxf)
static inline float fb_min(float a, float b) { return a < b ? a : b; }
static inline float fb_max(float a, float b) { return a > b ? a : b; }

struct FastBvhBuilder {
    struct P { float mn[3], mx[3], c[3]; };
    std::vector<P> prims;          // indexed by primitive
    std::vector<int> idx;          // permutation of primitive indices, partitioned in place
    std::vector<WrtNode> nodes;
    int max_depth = 0;

    static float half_area(const float* mn, const float* mx) {
        float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
        return dx * dy + dy * dz + dz * dx;
    }
    static void grow(float* mn, float* mx, const P& p) {
        for (int k = 0; k < 3; k++) { mn[k] = fb_min(mn[k], p.mn[k]); mx[k] = fb_max(mx[k], p.mx[k]); }
    }
    static void reset(float* mn, float* mx) {
        for (int k = 0; k < 3; k++) { mn[k] = INFINITY; mx[k] = -INFINITY; }
    }

    // Leaf boxes are taken from the reference tree's leaf records.
    void build(const WrtSceneDesc* s) {
        nodes.clear();
        max_depth = 0;
        const int n = s->n_prims;
        if (n == 0 || s->n_nodes == 0) return;
        prims.resize(n);
        idx.resize(n);
        for (int i = 0; i < s->n_nodes; i++) {
            const WrtNode& nd = s->nodes[i];
            if (nd.link >= 0 || i == 1) continue;            // inner node / padding record
            int p = ~nd.link;
            if (p < 0 || p >= n) continue;
            P& q = prims[p];
            for (int k = 0; k < 3; k++) { q.mn[k] = nd.pmin[k]; q.mx[k] = nd.pmax[k]; q.c[k] = 0.5f * nd.pmin[k] + 0.5f * nd.pmax[k]; }
        }
        for (int i = 0; i < n; i++) idx[i] = i;
        // A subtree over k primitives occupies exactly 2k-2 records below its root, so every subtree's
        // record range is known before it is built: sibling subtrees are built by independent threads
        // into disjoint ranges of one preallocated array (layout = depth-first, pairs adjacent).
        nodes.assign(2 * (size_t)n, WrtNode{});
        nodes[1].link = ~0;
        max_depth = rec_build(0, 0, n, 2, 0);
    }

    void set_leaf(int rec, int prim) {
        WrtNode& nd = nodes[rec];
        const P& p = prims[prim];
        for (int k = 0; k < 3; k++) { nd.pmin[k] = p.mn[k]; nd.pmax[k] = p.mx[k]; }
        nd.link = ~prim;
    }

    void sort_axis(int b, int e, int axis) {
        const std::vector<P>& pr = prims;
        std::sort(idx.begin() + b, idx.begin() + e, [&pr, axis](int x, int y) {
            float cx = pr[x].c[axis], cy = pr[y].c[axis];
            return cx < cy || (cx == cy && x < y);
        });
    }

    // chooses the split of idx[b,e); returns mid in (b,e) after partitioning
    int split(int b, int e) {
        const int n = e - b;
        if (n == 2) return b + 1;
        float cmn[3], cmx[3];
        reset(cmn, cmx);
        for (int i = b; i < e; i++)
            for (int k = 0; k < 3; k++) { cmn[k] = fb_min(cmn[k], prims[idx[i]].c[k]); cmx[k] = fb_max(cmx[k], prims[idx[i]].c[k]); }
        float best_cost = INFINITY;
        int best_axis = -1, best_pos = -1;
        if (n <= 16) {                                        // exact sweep over the three axes
            float right_area[16];
            int sorted_axis = -1;
            for (int axis = 0; axis < 3; axis++) {
                if (!(cmx[axis] > cmn[axis])) continue;
                sort_axis(b, e, axis);
                sorted_axis = axis;
                float mn[3], mx[3];
                reset(mn, mx);
                for (int i = n - 1; i > 0; i--) { grow(mn, mx, prims[idx[b + i]]); right_area[i] = half_area(mn, mx); }
                reset(mn, mx);
                for (int i = 1; i < n; i++) {
                    grow(mn, mx, prims[idx[b + i - 1]]);
                    float cost = half_area(mn, mx) * i + right_area[i] * (n - i);
                    if (cost < best_cost) { best_cost = cost; best_axis = axis; best_pos = i; }
                }
            }
            if (best_axis < 0) return b + n / 2;
            if (sorted_axis != best_axis) sort_axis(b, e, best_axis);
            return b + best_pos;
        }
        const int KMAX = 64, K = 64;                          // bins
        for (int axis = 0; axis < 3; axis++) {
            if (!(cmx[axis] > cmn[axis])) continue;
            int cnt[KMAX] = {0};
            float bmn[KMAX][3], bmx[KMAX][3];
            for (int j = 0; j < K; j++) reset(bmn[j], bmx[j]);
            const float scale = K / (cmx[axis] - cmn[axis]);
            for (int i = b; i < e; i++) {
                const P& p = prims[idx[i]];
                int j = std::min(K - 1, std::max(0, (int)((p.c[axis] - cmn[axis]) * scale)));
                cnt[j]++;
                grow(bmn[j], bmx[j], p);
            }
            float ra[KMAX];
            int rc[KMAX];
            float mn[3], mx[3];
            reset(mn, mx);
            int c = 0;
            for (int j = K - 1; j > 0; j--) {
                if (cnt[j]) { for (int k = 0; k < 3; k++) { mn[k] = fb_min(mn[k], bmn[j][k]); mx[k] = fb_max(mx[k], bmx[j][k]); } }
                c += cnt[j];
                ra[j] = c ? half_area(mn, mx) : 0.f;
                rc[j] = c;
            }
            reset(mn, mx);
            c = 0;
            for (int j = 1; j < K; j++) {
                if (cnt[j - 1]) { for (int k = 0; k < 3; k++) { mn[k] = fb_min(mn[k], bmn[j - 1][k]); mx[k] = fb_max(mx[k], bmx[j - 1][k]); } }
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
        const std::vector<P>& pr = prims;
        auto mid = std::partition(idx.begin() + b, idx.begin() + e, [&pr, axis, scale, lo, pos](int x) {
            int j = std::min(K - 1, std::max(0, (int)((pr[x].c[axis] - lo) * scale)));
            return j < pos;
        });
        int m = (int)(mid - idx.begin());
        if (m == b || m == e) return b + n / 2;
        return m;
    }

    // Eight copies of the tree, one per ray-direction octant (bit 0: d.x < 0, bit 1: d.y < 0, bit 2: d.z < 0),
    // with each record's planes pre-swapped into {entry planes}{exit planes} for that octant.
    std::vector<WrtNode> octant_copies() const { return octant_copies_of(nodes.data(), nodes.size()); }
    static std::vector<WrtNode> octant_copies_of(const WrtNode* nodes, size_t n_nodes) {
        std::vector<WrtNode> o;
        octant_copies_into(nodes, n_nodes, o);
        return o;
    }
    // Into a caller-owned vector (wrt_upload_scene keeps it between uploads: no fresh pages to fault in; this runs
    // inside the end-to-end time of a frame).  Copy, then swap the planes of the axes the octant looks down.
    static void octant_copies_into(const WrtNode* nodes, size_t n, std::vector<WrtNode>& o) {
        o.resize(8 * n);
        if (n == 0) return;
        for (int oct = 0; oct < 8; oct++) {
            WrtNode* dst = o.data() + (size_t)oct * n;
            memcpy(dst, nodes, n * sizeof(WrtNode));
            for (int k = 0; k < 3; k++) {
                if (!(oct & (1 << k))) continue;
                for (size_t i = 0; i < n; i++) { float t = dst[i].pmin[k]; dst[i].pmin[k] = dst[i].pmax[k]; dst[i].pmax[k] = t; }
            }
        }
    }

    // Copy of the tree whose leaf boxes are grown by rel * (largest extent, largest |coordinate|) + abs
    // on every side and whose inner boxes are the unions of those: a conservative culling structure
    // for queries the reference answers without boxes.
    std::vector<WrtNode> dilated(float rel, float abs_pad) const {
        std::vector<WrtNode> d = nodes;
        for (size_t i = 0; i < d.size(); i++) {
            if (d[i].link >= 0 || i == 1) continue;
            float ext = 0.f;
            for (int k = 0; k < 3; k++) {
                ext = fb_max(ext, d[i].pmax[k] - d[i].pmin[k]);
                ext = fb_max(ext, fb_max(fabsf(d[i].pmin[k]), fabsf(d[i].pmax[k])) * 1e-2f);
            }
            float pad = rel * ext + abs_pad;
            for (int k = 0; k < 3; k++) { d[i].pmin[k] -= pad; d[i].pmax[k] += pad; }
        }
        // inner boxes bottom-up: children always have larger indices than their parent
        for (size_t i = d.size(); i-- > 0;) {
            if (d[i].link < 0) continue;
            const WrtNode& L = d[d[i].link];
            const WrtNode& R = d[d[i].link + 1];
            for (int k = 0; k < 3; k++) { d[i].pmin[k] = fb_min(L.pmin[k], R.pmin[k]); d[i].pmax[k] = fb_max(L.pmax[k], R.pmax[k]); }
        }
        return d;
    }

    // Builds the subtree over idx[b,e) into record `rec`; its descendants use records [free, free + 2(e-b) - 2).
    // Returns the depth of the deepest leaf below (root = `depth`).
    int rec_build(int rec, int b, int e, int free, int depth) {
        if (e - b == 1) { set_leaf(rec, idx[b]); return depth; }
        int mid = split(b, e);
        const int pair = free;
        nodes[rec].link = pair;
        const int nl = mid - b;
        const int free_l = pair + 2, free_r = free_l + (2 * nl - 2);
        int dl = 0, dr = 0;
        if (e - b >= 2048 && depth < 4) {                      // big subtrees: build the halves concurrently
            std::thread t([&] { dl = rec_build(pair, b, mid, free_l, depth + 1); });
            dr = rec_build(pair + 1, mid, e, free_r, depth + 1);
            t.join();
        } else {
            dl = rec_build(pair, b, mid, free_l, depth + 1);
            dr = rec_build(pair + 1, mid, e, free_r, depth + 1);
        }
        const WrtNode L = nodes[pair], R = nodes[pair + 1];
        WrtNode& nd = nodes[rec];
        for (int k = 0; k < 3; k++) { nd.pmin[k] = fb_min(L.pmin[k], R.pmin[k]); nd.pmax[k] = fb_max(L.pmax[k], R.pmax[k]); }
        return dl > dr ? dl : dr;
    }
};

} // namespace wrt
