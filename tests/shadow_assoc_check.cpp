// CPU check of csrc/cuda/shadow_assoc.h: on random binary trees (random shapes, depth up to 60) with random subsets of
// "blocking" leaves and random float factors, the stack reduction over path codes must give the bits of the recursion
// value(node) = value(left) * value(right), value(non-blocking leaf) = 1 — BVHStrategy::ShadowHelper's association.
// Prints {"cases": n, "mismatch": m, "order_matters": k} (order_matters: cases in which the left-to-right product differs,
// i.e. the check is not vacuous).
// `shadow_assoc_check tree <file> <cases> <seed>`: the same on a scene's own flattened tree (a file of WrtNode records,
// include/wrt_scene.h), with the path codes made by wrt_make_path_codes — the function wrt_upload_scene stages them with.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>
#include "../whittedstyle_raytracer_b200/csrc/cuda/shadow_assoc.h"
#include "../include/wrt_scene.h"

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

// ---- a scene's flattened tree ----
static float recurse_flat(const std::vector<WrtNode>& t, int rec, const std::vector<float>& factor) {
    const int link = t[rec].link;
    if (link < 0) return factor[~link];
    const float l = recurse_flat(t, link, factor);
    const float r = recurse_flat(t, link + 1, factor);
    return l * r;
}

static int check_flat_tree(const char* file, int cases, uint64_t seed) {
    FILE* f = fopen(file, "rb");
    if (!f) { printf("cannot open %s\n", file); return 2; }
    fseek(f, 0, SEEK_END);
    const long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<WrtNode> t((size_t)bytes / sizeof(WrtNode));
    if (fread(t.data(), sizeof(WrtNode), t.size(), f) != t.size()) { fclose(f); return 2; }
    fclose(f);
    const int nn = (int)t.size(), np = nn / 2;
    std::vector<WrtPathCode> pc;
    if (!wrt_make_path_codes(t.data(), nn, np, pc)) { printf("{\"codes\": false}\n"); return 3; }
    // primitives are numbered in the tree's depth-first leaf order (wrt_scene.h): the sort by primitive index relies on it
    {
        int next = 0, bad = 0, maxd = 0;
        std::vector<int> todo{0};
        while (!todo.empty()) {
            const int rec = todo.back(); todo.pop_back();
            const int link = t[rec].link;
            if (link < 0) { if (~link != next++) ++bad; if (pc[~link].depth > maxd) maxd = pc[~link].depth; }
            else { todo.push_back(link + 1); todo.push_back(link); }
        }
        if (bad || next != np) { printf("{\"leaf_order_violations\": %d}\n", bad + (next != np)); return 4; }
        printf("{\"prims\": %d, \"depth\": %d}\n", np, maxd);
    }
    Rng r{seed};
    long long mismatch = 0, order_matters = 0;
    std::vector<float> factor((size_t)np, 1.f);
    for (int c = 0; c < cases; c++) {
        const int cap = np < WRT_SHADOW_HITS ? np : WRT_SHADOW_HITS;          // (a small scene has fewer primitives than that)
        const int want = cap <= 3 ? cap : 3 + (int)(r.next() % (unsigned)(cap - 2));
        const bool equal = r.next() % 2 == 0;                   // the bunny: one glass material
        const float a = equal ? 0.8f : 0.05f + 0.9f * r.unif();
        // blockers the way a ray meets them: a few clusters of neighbouring primitives, in arbitrary (visit) order
        std::vector<int> prim;
        while ((int)prim.size() < want) {
            const int base = (int)(r.next() % (unsigned)np);
            const int run = 1 + (int)(r.next() % 3);
            for (int k = 0; k < run && (int)prim.size() < want; k++) {
                const int p = (base + (int)(r.next() % 8)) % np;
                bool dup = false;
                for (int q : prim) dup = dup || q == p;
                if (!dup) prim.push_back(p);
            }
        }
        std::vector<float> fvisit;
        for (int p : prim) { factor[p] = equal ? a : 0.05f + 0.9f * r.unif(); fvisit.push_back(factor[p]); }
        const float ref = recurse_flat(t, 0, factor);
        // what the kernel does: insertion sort by primitive index, then the stack reduction
        std::vector<int> sorted = prim;
        for (size_t i = 1; i < sorted.size(); i++) for (size_t k = i; k > 0 && sorted[k - 1] > sorted[k]; k--) std::swap(sorted[k - 1], sorted[k]);
        std::vector<float> fs;
        for (int p : sorted) fs.push_back(factor[p]);
        const float got = wrt_tree_product((int)sorted.size(), sorted.data(), fs.data(), pc.data());
        float seq = 1.f;
        for (float x : fvisit) seq = seq * x;
        if (memcmp(&ref, &got, 4) != 0) ++mismatch;
        if (memcmp(&ref, &seq, 4) != 0) ++order_matters;
        for (int p : prim) factor[p] = 1.f;
    }
    printf("{\"cases\": %d, \"mismatch\": %lld, \"order_matters\": %lld}\n", cases, mismatch, order_matters);
    return mismatch ? 1 : 0;
}

int main(int argc, char** argv) {
    if (argc > 2 && !strcmp(argv[1], "tree"))
        return check_flat_tree(argv[2], argc > 3 ? atoi(argv[3]) : 20000, argc > 4 ? (uint64_t)atoll(argv[4]) : 7ull);
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
