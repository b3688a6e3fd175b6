// prune_rule.h — how far beyond the best hit the pruned closest-hit walk still has to look.
//
// The reference's getIntersection (BVH.hpp:137-159) visits every node whose box the ray hits and keeps the minimum t
// (ties: left subtree = smaller primitive index).  The pruned walk skips a box entered beyond a LIMIT derived from the
// best hit so far; that is only sound if no primitive inside a skipped box can be accepted with a smaller t.
// "A primitive lies inside its box" is not enough:
//   * Triangle::intersect (Triangle.hpp:41) accepts barycentrics down to -1e-5, i.e. hit points up to
//     s_k = 1e-5 * (|E1_k| + |E2_k|) outside the triangle's box along axis k.  The hit point lies in the box dilated by
//     s, whose entry distance is the box's own minus at most max_k(s_k * |1/d_k|): a ray that sees a box face edge-on
//     (|1/d_k| large) can cross the triangle's plane long before it enters the box.
//   * t is a float result: the slab distances carry an absolute error of about eps * (|plane| + |o|) * |1/d_k|, and so
//     does Moller-Trumbore's t for the large axis-aligned triangles (walls, floors) where sin(angle to the plane) = |d_k|.
// So the limit is
//     t_best + 2e-3 * |t_best| + 1e-3                         (relative part + absolute floor)
//            + max_k |1/d_k| * (4 * slack + 64 * eps * M)      (this header)
// with slack = max over the scene's triangles of 1e-5 * max_k(|E1_k| + |E2_k|) and M = the largest |coordinate| of the ray
// origin.  The first term of the bracket is the geometric bound above with a factor 4 of safety; the second covers
// the float errors.  tests/prune_rule_check.cpp evaluates the rule by brute force on adversarial rays (targets on and
// just outside edges and vertices, elevations down to 1e-4 rad over the triangle's plane, box faces seen edge-on, the
// 40-unit wall triangles): the true closest primitive is never behind the limit of any other accepted primitive,
// whatever the visiting order (worst observed gap = 0.3 of the margin; round 1's t*(1+1e-3)+1e-3 fails it).
// What no finite margin covers: Moller-Trumbore's t carries an error of about eps * |o - v0| / sin(elevation); at
// 1e-5 rad that is 1 % of t and the reference's own answer is rounding noise (the checker reports that regime
// separately: 40 of 47 000 such rays would be decided differently).  WRT_TRAVERSAL_EXHAUSTIVE (no pruning) is the
// literal mode for callers who need the reference's noise reproduced as well.
#ifndef WRT_PRUNE_RULE_H
#define WRT_PRUNE_RULE_H

#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define WRT_PRUNE_HD __host__ __device__ __forceinline__
#else
#define WRT_PRUNE_HD static inline
#endif

#define WRT_PRUNE_ACCEPT_EPS 0.00001f      /* EPSILON of Triangle.hpp:41 */

/* slack of one triangle: how far outside its own box an accepted hit point can lie (largest axis) */
WRT_PRUNE_HD float wrt_prune_triangle_slack(const float E1[3], const float E2[3]) {
    float m = fmaxf(fabsf(E1[0]) + fabsf(E2[0]), fmaxf(fabsf(E1[1]) + fabsf(E2[1]), fabsf(E1[2]) + fabsf(E2[2])));
    return m < 1e30f ? WRT_PRUNE_ACCEPT_EPS * m : 0.f;        /* non-finite geometry: no claim, nothing finite to prune by */
}

/* host twin of the reduction k_pack_prims does at upload (prim_geom: 12 floats per primitive, wrt_scene.h) */
WRT_PRUNE_HD float wrt_prune_scene_slack(const float* prim_geom, const uint32_t* prim_flags, int n_prims) {
    float s = 0.f;
    for (int p = 0; p < n_prims; p++) {
        if ((prim_flags[p] & 1u) != 0u) continue;              /* spheres: the hit point lies on the sphere */
        s = fmaxf(s, wrt_prune_triangle_slack(prim_geom + 12 * (size_t)p + 4, prim_geom + 12 * (size_t)p + 8));
    }
    return s;
}

/* per ray, once: the direction- and position-dependent part of the margin */
WRT_PRUNE_HD float wrt_prune_ray_scale(const float o[3], const float inv[3], float slack) {
    const float maxinv = fmaxf(fabsf(inv[0]), fmaxf(fabsf(inv[1]), fabsf(inv[2])));
    const float M = fmaxf(fabsf(o[0]), fmaxf(fabsf(o[1]), fabsf(o[2])));
    return maxinv * (4.f * slack + 64.f * 5.9604645e-8f * M);
}

/* boxes entered beyond this cannot hold a closer accepted primitive */
WRT_PRUNE_HD float wrt_prune_limit(float t_best, float ray_scale) {
    return fabsf(t_best) * 2e-3f + 1e-3f + ray_scale + t_best;
}

#endif /* WRT_PRUNE_RULE_H */
