// shadow_assoc.h — the hard-shadow product in the association of the REFERENCE's tree, without walking that tree
// (host/device: tests/shadow_assoc_check.cpp runs the same source on the CPU).
//
// BVHStrategy::ShadowHelper (BVHStrategy.hpp:24-48) returns 1 for a missed box or leaf, the leaf's (1 - alpha) for a
// blocking leaf and `l * r` for an inner node: the product of the blocking leaves' factors, ASSOCIATED LIKE THE TREE.  The
// kernels walk another tree and meet the same leaves in another order.  IEEE multiplication is
// commutative but not associative: up to two factors != 1, or three equal ones, give the same bits in any association;
// from there on they need not (0.8^4: ((a*a)*a)*a and (a*a)*(a*a) differ by one ulp — and every shipped scene's bunny
// has alpha = 0.2).  Since x * 1 == x exactly, the reference's value is the product over the tree CONTRACTED to its
// blocking leaves, and that contraction follows from the leaves' root-to-leaf paths alone:
//   * primitives are stored in the reference tree's depth-first leaf order, so sorting the blocking leaves by primitive
//     index puts them in tree order;
//   * two neighbours in that order join at their lowest common ancestor, whose depth is the length of the common prefix
//     of their paths (WrtPathCode: one bit per step, 0 = left, 1 = right, root first);
//   * a stack reduction over the neighbours' join depths (the Cartesian tree of that sequence: deeper joins first)
//     multiplies exactly the pairs the recursion multiplies.  In a binary tree two neighbouring join depths never tie.
#ifndef WRT_SHADOW_ASSOC_H
#define WRT_SHADOW_ASSOC_H

#ifdef __CUDACC__
#define WRT_ASSOC_HD __host__ __device__ __forceinline__
#else
#define WRT_ASSOC_HD static inline
#endif

#define WRT_SHADOW_HITS 12             /* blocking leaves with a factor != 1 a ray may collect; beyond: visit-order product */
#define WRT_PATH_BITS 64               /* deeper reference trees: no path codes, visit-order product */

typedef struct WrtPathCode {
    unsigned hi, lo;                   /* bit (63 - k) of hi:lo = step k from the root (0 left, 1 right) */
    int depth;                         /* steps from the root to the leaf */
    int pad;
} WrtPathCode;

/* depth of the lowest common ancestor of two DIFFERENT leaves */
WRT_ASSOC_HD int wrt_lca_depth(WrtPathCode a, WrtPathCode b) {
    const unsigned long long x = ((unsigned long long)(a.hi ^ b.hi) << 32) | (unsigned long long)(a.lo ^ b.lo);
#ifdef __CUDA_ARCH__
    const int common = x ? __clzll((long long)x) : 64;
#else
    const int common = x ? __builtin_clzll(x) : 64;
#endif
    const int dmin = a.depth < b.depth ? a.depth : b.depth;
    return common < dmin ? common : dmin;
}

/* prim[0..k) ascending (tree order), f[i] the factor of prim[i]; k <= WRT_SHADOW_HITS */
WRT_ASSOC_HD float wrt_tree_product(int k, const int* prim, const float* f, const WrtPathCode* codes) {
    float v[WRT_SHADOW_HITS];
    int d[WRT_SHADOW_HITS];            /* d[j]: depth at which stack entry j joins entry j - 1 */
    int sp = 0;
    for (int i = 0; i < k; i++) {
        const int di = i == 0 ? -1 : wrt_lca_depth(codes[prim[i - 1]], codes[prim[i]]);
        while (sp >= 2 && d[sp - 1] > di) {            /* the two on top join below the new leaf's join: their node is complete */
            v[sp - 2] = v[sp - 2] * v[sp - 1];
            --sp;
        }
        v[sp] = f[i]; d[sp] = di; ++sp;
    }
    while (sp >= 2) {                                  /* join depths increase towards the top: deepest first */
        v[sp - 2] = v[sp - 2] * v[sp - 1];
        --sp;
    }
    return sp ? v[0] : 1.f;
}

#ifdef __cplusplus
#include <vector>
/* Host: the path codes of a flattened tree's primitives (include/wrt_scene.h — record 0 is the root; `link` >= 0: the
 * children are records `link` (left) and `link + 1` (right); `link` < 0: the leaf of primitive ~link).  What
 * wrt_upload_scene stages for the kernels; tests/shadow_assoc_check.cpp runs it on the scenes' own trees.  False (and no
 * codes) for a tree with an inner node at depth WRT_PATH_BITS or more than n_nodes reachable records (not a tree). */
template <class Node>
inline bool wrt_make_path_codes(const Node* nodes, int n_nodes, int n_prims, std::vector<WrtPathCode>& out) {
    out.clear();
    if (n_prims <= 0 || n_nodes <= 0) return false;
    out.assign((size_t)n_prims, WrtPathCode{0u, 0u, 0, 0});
    struct Item { int rec; int depth; unsigned long long path; };
    std::vector<Item> todo;
    todo.push_back(Item{0, 0, 0ull});
    long long visited = 0;
    while (!todo.empty()) {
        const Item it = todo.back();
        todo.pop_back();
        if (++visited > (long long)n_nodes || it.rec < 0 || it.rec >= n_nodes) { out.clear(); return false; }
        const int link = nodes[it.rec].link;
        if (link < 0) {
            const int prim = ~link;
            if (prim >= 0 && prim < n_prims) out[(size_t)prim] = WrtPathCode{(unsigned)(it.path >> 32), (unsigned)it.path, it.depth, 0};
        } else {
            if (it.depth >= WRT_PATH_BITS) { out.clear(); return false; }
            todo.push_back(Item{link, it.depth + 1, it.path});
            todo.push_back(Item{link + 1, it.depth + 1, it.path | (1ull << (63 - it.depth))});
        }
    }
    return true;
}
#endif

#endif /* WRT_SHADOW_ASSOC_H */
