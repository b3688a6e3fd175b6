// wide_bvh.h — 4-wide view of the binary tree the kernels walk (host/device: the CPU checkers build the same nodes).
//
// The binary trees keep sibling pairs adjacent: one traversal step fetches the pair `c` (records c, c+1; 64 bytes) and
// tests two boxes, and the next fetch depends on the outcome.  Deep ray-tree levels and the soft-shadow shaft walks are
// chains of such dependent fetches (20-120 per walk).  A wide node holds, for pair c, the FOUR GRANDCHILD records (the
// children of record c and of record c+1) in one 128-byte block:
//     slot 0, 1 = children of record c     (or record c itself + an empty slot, when c is a leaf)
//     slot 2, 3 = children of record c+1   (likewise)
// Each slot is an unchanged 32-byte record {entry planes, link}{exit planes, pad} of the octant copy it was taken from,
// so the box tests, the leaf encoding (link < 0: ~primitive) and the child links (link >= 0: pair index, i.e. the next
// wide node) are the binary tree's own; a walk over wide nodes visits half as many nodes and never tests the two child
// boxes whose grandchildren it tests directly (measured on the metric frame's shaft walks: 119.6 -> 60.6 node visits,
// 239 -> 205 box tests per request in shadow).
// Exactness: the set of primitives a ray (or shaft) tests is "own box hit" (DESIGN.md section 4), whatever inner boxes
// are consulted on the way: every inner box is the exact union of its children, so skipping the test of a child box
// and testing its two children instead visits a superset of the leaves the binary walk visits, and each leaf is still
// tested against its own exact box.
// An empty slot holds planes no ray of the copy's octant can pass: entry planes at +-FLT_MAX beyond the exit planes.
#ifndef WRT_WIDE_BVH_H
#define WRT_WIDE_BVH_H

#include <float.h>

#ifdef __CUDACC__
#define WRT_WIDE_HD __host__ __device__ __forceinline__
#else
#include <vector_types.h>
#define WRT_WIDE_HD static inline
#endif

/* records per wide node: 4 slots x 2 float4; the wide node of pair c starts at float4 index 4 * c of its octant copy
 * (c is even: the array is twice the size of the binary copy, every node 128-byte aligned) */
#define WRT_WIDE_FLOAT4_PER_RECORD 4

WRT_WIDE_HD void wrt_wide4_node(const float4* src, int c, int oct, float4* out) {
    union { float f; int i; unsigned u; } w;
    for (int k = 0; k < 2; k++) {
        const float4 lo = src[2 * (size_t)(c + k)], hi = src[2 * (size_t)(c + k) + 1];
        w.f = lo.w;
        if (w.i >= 0) {
            const float4* ch = src + 2 * (size_t)w.i;
            out[4 * k + 0] = ch[0]; out[4 * k + 1] = ch[1];
            out[4 * k + 2] = ch[2]; out[4 * k + 3] = ch[3];
        } else {
            out[4 * k + 0] = lo; out[4 * k + 1] = hi;
            float4 e0, e1;                                   /* entry planes beyond the exit planes, for this octant's rays */
            e0.x = (oct & 1) ? -FLT_MAX : FLT_MAX; e1.x = -e0.x;
            e0.y = (oct & 2) ? -FLT_MAX : FLT_MAX; e1.y = -e0.y;
            e0.z = (oct & 4) ? -FLT_MAX : FLT_MAX; e1.z = -e0.z;
            w.u = 0x80000000u;                               /* link: a leaf nobody reaches */
            e0.w = w.f; e1.w = 0.f;
            out[4 * k + 2] = e0; out[4 * k + 3] = e1;
        }
    }
}

#endif /* WRT_WIDE_BVH_H */
