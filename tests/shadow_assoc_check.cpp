// CPU check of csrc/cuda/shadow_assoc.h: on random binary trees (random shapes, depth up to 60) with random subsets of
// "blocking" leaves and random float factors, the stack reduction over path codes must give the bits of the recursion
// value(node) = value(left) * value(right), value(non-blocking leaf) = 1 — BVHStrategy::ShadowHelper's association.
// Prints {"cases": n, "mismatch": m, "order_matters": k} (order_matters: cases in which the left-to-right product differs,
// i.e. the check is not vacuous).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../whittedstyle_raytracer_b200/csrc/cuda/shadow_assoc.h"

struct Rng {
    uint64_t s;
    uint32_t next() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); }
    float unif() { return (next() >> 8) * (1.0f / 16777216.0f); }
};

struct Node { int left, right, leaf; };            // leaf >= 0: leaf rank

static std::vector<Node> nodes;
static std::vector<WrtPathCode> codes;

static int build(Rng& r, int n_leaves, int& next_leaf, unsigned long long path, int depth) {
    const int id = (int)nodes.size();
    nodes.push_back(Node{-1, -1, -1});
    if (n_leaves == 1 || depth >= 60) {
        // (depth cap: hang the remaining leaves as a right spine would exceed 64 bits; make this one leaf instead)
        nodes[id].leaf = next_leaf;
        WrtPathCode c;
        c.hi = (unsigned)(path >> 32); c.lo = (unsigned)path; c.depth = depth; c.pad = 0;
        codes.push_back(c);
        ++next_leaf;
        return id;
    }
    // random split, skewed now and then (the reference's median split is balanced; the rule must not depend on it)
    int nl = 1 + (int)(r.next() % (unsigned)(n_leaves - 1));
    if (r.next() % 4 == 0) nl = 1;
    if (r.next() % 4 == 0) nl = n_leaves - 1;
    const int l = build(r, nl, next_leaf, path, depth + 1);
    const int rr = build(r, n_leaves - nl, next_leaf, path | (1ull << (63 - depth)), depth + 1);
    nodes[id].left = l; nodes[id].right = rr;
    return id;
}

static float recurse(int id, const std::vector<float>& factor) {
    const Node& n = nodes[id];
    if (n.leaf >= 0) return factor[n.leaf];
    const float l = recurse(n.left, factor);
    const float r = recurse(n.right, factor);
    return l * r;
}

int main(int argc, char** argv) {
    const int cases = argc > 1 ? atoi(argv[1]) : 20000;
    Rng r{argc > 2 ? (uint64_t)atoll(argv[2]) : 7ull};
    long long mismatch = 0, order_matters = 0, done = 0;
    for (int c = 0; c < cases; c++) {
        nodes.clear(); codes.clear();
        const int n_leaves = 2 + (int)(r.next() % 300);
        int next_leaf = 0;
        const int root = build(r, n_leaves, next_leaf, 0ull, 0);
        const int nl = next_leaf;
        std::vector<float> factor(nl, 1.f);
        const int want = 1 + (int)(r.next() % WRT_SHADOW_HITS);
        const bool equal = r.next() % 3 == 0;                   // one glass material: equal factors
        const float a = 0.05f + 0.9f * r.unif();
        std::vector<int> prim;
        std::vector<float> f;
        for (int p = 0; p < nl && (int)prim.size() < want; p++)
            if ((int)(r.next() % (unsigned)nl) < want * 2) { factor[p] = equal ? a : 0.05f + 0.9f * r.unif(); prim.push_back(p); f.push_back(factor[p]); }
        if (prim.empty()) continue;
        const float ref = recurse(root, factor);
        const float got = wrt_tree_product((int)prim.size(), prim.data(), f.data(), codes.data());
        float seq = 1.f;
        for (float x : f) seq = seq * x;
        if (memcmp(&ref, &got, 4) != 0) ++mismatch;
        if (memcmp(&ref, &seq, 4) != 0) ++order_matters;
        ++done;
    }
    printf("{\"cases\": %lld, \"mismatch\": %lld, \"order_matters\": %lld}\n", done, mismatch, order_matters);
    return mismatch ? 1 : 0;
}
