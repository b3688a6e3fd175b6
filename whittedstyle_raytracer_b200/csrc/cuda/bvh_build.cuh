// bvh_build.cuh — device-side build of the walked trees during wrt_upload_scene.
//
//   k_bvh_leaves     reference leaf records -> per-primitive boxes (prim_box), build-tree leaves, Morton keys
//   (radix sort of the keys: cub::DeviceRadixSort, the one library call of the build)
//   k_bvh_ploc       PLOC passes (ploc_bvh.h) in ONE cooperative launch: nearest neighbour | fate + per-block counts |
//                    ordered compaction + node creation, a grid barrier after each, until one cluster is left
//   k_bvh_layout     build tree -> 32-byte pair-adjacent records of the SAH-quality tree (fnodes) and of its dilated
//                    copy (dnodes); tree depth
//   k_octant_copies  8 copies with pre-swapped entry/exit planes (what BoundBox.hpp:68-70 does per box), for the built tree
//                    and for the reference-topology tree
// Nothing of this runs on the host any more: the host uploads the reference tree (it is part of the scene description, the
// contract with the oracle) and reads back one integer, the depth, to size the traversal stacks.
#pragma once
#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>

#include <cfloat>

#include "ploc_bvh.h"
#include "shaft_cull.h"
#include "wide_bvh.h"
#include "../../../include/wrt_scene.h"

namespace wrt {

#define WRT_PLOC_THREADS 256

enum BuildState { BS_ALLOC = 0, BS_ROOT = 1, BS_DEPTH = 2, BS_PASSES = 3, BS_ERROR = 4, BS_LEAVES = 5, BS_TOTAL = 8 };

struct BuildBuffers {
    PlocTree tree;                 // 2n entries each
    unsigned long long* keys[2];   // n
    int* vals[2];                  // n
    int* cl[2];                    // n: cluster lists (ping-pong)
    int* nn;                       // n
    int* blk;                      // per-block keep counts
    int* state;                    // BS_TOTAL
    void* cub_temp;
    size_t cub_temp_bytes;
};

// float atomic min / max through the integer units (works for any mix of signs; the cell starts at +-FLT_MAX)
__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
    if (v >= 0.f) atomicMin((int*)addr, __float_as_int(v));
    else atomicMax((unsigned*)addr, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.f) atomicMax((int*)addr, __float_as_int(v));
    else atomicMin((unsigned*)addr, __float_as_uint(v));
}

// bounds[0..2] = min, bounds[3..5] = max of the leaf-box centroids (the Morton grid spans the centroids, not the
// scene: measured on real rays, trees sorted on the centroid grid cost 4-5 % more node steps than the host's
// binned-SAH tree, on the scene-box grid 9 %)
__global__ void __launch_bounds__(256) k_bvh_leaves(const float4* __restrict__ ref_nodes, int n_nodes, int n_prims, BuildBuffers bb,
                                                    float4* prim_box, float* bounds, float dil_rel, float dil_abs) {
    float cmn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, cmx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_nodes; i += gridDim.x * blockDim.x) {
        if (i == 1) continue;                                  // padding record
        const float4 lo = ref_nodes[2 * (size_t)i], hi = ref_nodes[2 * (size_t)i + 1];
        const int link = __float_as_int(lo.w);
        if (link >= 0) continue;
        const int p = ~link;
        if (p < 0 || p >= n_prims) { atomicExch(bb.state + BS_ERROR, 1); continue; }
        if (atomicAdd(bb.tree.cnt + p, 1) != 0) { atomicExch(bb.state + BS_ERROR, 1); continue; }   // (cnt was zeroed) a primitive in two leaves
        atomicAdd(bb.state + BS_LEAVES, 1);
        const float mn[3] = {lo.x, lo.y, lo.z}, mx[3] = {hi.x, hi.y, hi.z};
        prim_box[2 * (size_t)p] = make_float4(mn[0], mn[1], mn[2], 0.f);
        prim_box[2 * (size_t)p + 1] = make_float4(mx[0], mx[1], mx[2], 0.f);
        ploc_init_leaf(bb.tree, p, mn, mx, dil_rel, dil_abs);
        for (int k = 0; k < 3; k++) {
            const float c = 0.5f * mn[k] + 0.5f * mx[k];
            if (c < cmn[k]) cmn[k] = c;                        // (NaN centroids never enter the bounds)
            if (c > cmx[k]) cmx[k] = c;
        }
        bb.vals[0][p] = p;
    }
    for (int k = 0; k < 3; k++) {
        float a = cmn[k], b = cmx[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            a = fminf(a, __shfl_xor_sync(0xffffffffu, a, off));
            b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, off));
        }
        if ((threadIdx.x & 31) == 0) {
            if (a < FLT_MAX) atomic_min_float(bounds + k, a);
            if (b > -FLT_MAX) atomic_max_float(bounds + 3 + k, b);
        }
    }
}

__global__ void __launch_bounds__(256) k_bvh_keys(int n_prims, BuildBuffers bb, const float* __restrict__ bounds) {
    const float bmin[3] = {bounds[0], bounds[1], bounds[2]}, bmax[3] = {bounds[3], bounds[4], bounds[5]};
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_prims; p += gridDim.x * blockDim.x) {
        const float4 lo = bb.tree.lo[p], hi = bb.tree.hi[p];
        const float c[3] = {0.5f * lo.x + 0.5f * hi.x, 0.5f * lo.y + 0.5f * hi.y, 0.5f * lo.z + 0.5f * hi.z};
        bb.keys[0][p] = ploc_morton(c, bmin, bmax);
    }
}

// exclusive scan of one int per thread over the CTA; returns the thread's offset, *total = CTA sum
__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
    __shared__ int warp_sums[WRT_PLOC_THREADS / 32];
    __shared__ int cta_total;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= (unsigned)off) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < WRT_PLOC_THREADS / 32 ? warp_sums[lane] : 0;
        int wi = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, off);
            if (lane >= (unsigned)off) wi += t;
        }
        if (lane < WRT_PLOC_THREADS / 32) warp_sums[lane] = wi - w;      // exclusive warp offsets
        if (lane == 31) cta_total = wi;
    }
    __syncthreads();
    const int res = warp_sums[warp] + incl - v;
    *total = cta_total;
    __syncthreads();                                       // the shared arrays are reused by the next call
    return res;
}

__global__ void __launch_bounds__(WRT_PLOC_THREADS) k_bvh_ploc(BuildBuffers bb, const int* __restrict__ sorted, int n, int radius) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const PlocTree t = bb.tree;
    int* cur = bb.cl[0];
    int* nxt = bb.cl[1];
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
    // a malformed description (a primitive without a leaf record, or in two) must never index with garbage: every thread
    // reads the same verdict of k_bvh_leaves and leaves before the first barrier
    if (bb.state[BS_ERROR] != 0 || bb.state[BS_LEAVES] != n) {
        if (gtid == 0 && bb.state[BS_ERROR] == 0) bb.state[BS_ERROR] = 1;
        return;
    }
    for (int i = gtid; i < n; i += gsize) cur[i] = sorted[i];
    if (gtid == 0) { bb.state[BS_ALLOC] = n; bb.state[BS_PASSES] = 0; }
    grid.sync();
    int m = n, passes = 0;
    while (m > 1) {
        // 1. nearest neighbour inside the window
        for (int i = gtid; i < m; i += gsize) bb.nn[i] = ploc_nearest(t, cur, m, i, radius);
        grid.sync();
        // 2. fates; every CTA owns one contiguous chunk of the cluster list and counts its survivors
        const int chunk = (m + gridDim.x - 1) / gridDim.x;
        const int b0 = min(m, (int)blockIdx.x * chunk), b1 = min(m, b0 + chunk);
        int keep = 0;
        for (int i = b0 + threadIdx.x; i < b1; i += blockDim.x) keep += ploc_fate(bb.nn, i) != 0;
        int cta_keep;
        block_exclusive_scan(keep, &cta_keep);
        if (threadIdx.x == 0) bb.blk[blockIdx.x] = cta_keep;
        grid.sync();
        // 3. ordered compaction: the CTA's base = survivors of the CTAs before it
        int base = 0, total = 0;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
            const int c = bb.blk[b];
            total += c;
            if (b < (int)blockIdx.x) base += c;
        }
        int dummy;
        // (two CTA-wide sums through the scan helper: offsets are not needed, only the totals)
        block_exclusive_scan(base, &base);
        block_exclusive_scan(total, &total);
        for (int i0 = b0; i0 < b1; i0 += blockDim.x) {
            const int i = i0 + threadIdx.x;
            const int fate = i < b1 ? ploc_fate(bb.nn, i) : 0;
            int tile_total;
            const int off = block_exclusive_scan(fate != 0 ? 1 : 0, &tile_total);
            if (fate == 2) {
                const int id = atomicAdd(bb.state + BS_ALLOC, 1);     // node ids are arbitrary; the topology is not
                ploc_make_node(t, id, cur[i], cur[bb.nn[i]]);
                nxt[base + off] = id;
            } else if (fate == 1) {
                nxt[base + off] = cur[i];
            }
            base += tile_total;
        }
        (void)dummy;
        ++passes;
        if (total >= m) {                                      // cannot happen (ploc_bvh.h: progress); never spin forever
            if (gtid == 0) atomicExch(bb.state + BS_ERROR, 2);
            break;
        }
        m = total;
        int* sw = cur; cur = nxt; nxt = sw;
        grid.sync();
    }
    if (gtid == 0) { bb.state[BS_ROOT] = n > 0 ? cur[0] : -1; bb.state[BS_PASSES] = passes; }
}

// Build tree -> records.  frec / drec: 2 float4 per record ({min, link}{max, 0}), 2n records (record 1 = padding).
__global__ void __launch_bounds__(256) k_bvh_layout(BuildBuffers bb, int n, float4* frec, float4* drec) {
    const PlocTree t = bb.tree;
    if (bb.state[BS_ERROR] != 0) return;
    const int total = n > 0 ? 2 * n - 1 : 0;
    int max_depth = 0;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < total; v += gridDim.x * blockDim.x) {
        int link, depth;
        const int r = ploc_record_of(t, n, v, &link, &depth);
        const float4 lo = t.lo[v], hi = t.hi[v], dlo = t.dlo[v], dhi = t.dhi[v];
        frec[2 * (size_t)r] = make_float4(lo.x, lo.y, lo.z, __int_as_float(link));
        frec[2 * (size_t)r + 1] = make_float4(hi.x, hi.y, hi.z, 0.f);
        drec[2 * (size_t)r] = make_float4(dlo.x, dlo.y, dlo.z, __int_as_float(link));
        drec[2 * (size_t)r + 1] = make_float4(dhi.x, dhi.y, dhi.z, 0.f);
        if (v < n && depth > max_depth) max_depth = depth;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && n > 0) {       // padding record, as the host build leaves it
        frec[2] = make_float4(0.f, 0.f, 0.f, __int_as_float(~0)); frec[3] = make_float4(0.f, 0.f, 0.f, 0.f);
        drec[2] = frec[2]; drec[3] = frec[3];
    }
    max_depth = __reduce_max_sync(0xffffffffu, max_depth);
    if ((threadIdx.x & 31) == 0 && max_depth > 0) atomicMax(bb.state + BS_DEPTH, max_depth);
}

// dst[oct][record]: the record with the planes of every axis the octant looks down swapped into {entry}{exit} order
__global__ void __launch_bounds__(256) k_octant_copies(const float4* __restrict__ src, int n_nodes, float4* dst) {
    const long long total = 8ll * n_nodes;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const int oct = (int)(g / n_nodes);
        const int i = (int)(g - (long long)oct * n_nodes);
        float4 lo = src[2 * (size_t)i], hi = src[2 * (size_t)i + 1];
        float t;
        if (oct & 1) { t = lo.x; lo.x = hi.x; hi.x = t; }
        if (oct & 2) { t = lo.y; lo.y = hi.y; hi.y = t; }
        if (oct & 4) { t = lo.z; lo.z = hi.z; hi.z = t; }
        dst[2 * ((size_t)oct * n_nodes + i)] = lo;
        dst[2 * ((size_t)oct * n_nodes + i) + 1] = hi;
    }
}

// dst[oct]: the 4-wide view (wide_bvh.h) of octant copy oct of a tree: one 128-byte node per sibling pair, holding the
// pair's four grandchild records.  One thread per (octant, pair).
__global__ void __launch_bounds__(256) k_wide4_copies(const float4* __restrict__ onodes, int n_nodes, float4* dst) {
    const int pairs = n_nodes / 2;
    const long long total = 8ll * pairs;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const int oct = (int)(g / pairs);
        const int c = 2 * (int)(g - (long long)oct * pairs);
        if (c == 0) continue;                               // records 0, 1: the root box and its padding, not a sibling pair
        float4 out[8];
        wrt_wide4_node(onodes + 2 * (size_t)oct * n_nodes, c, oct, out);
        float4* d = dst + WRT_WIDE_FLOAT4_PER_RECORD * ((size_t)oct * n_nodes + c);
#pragma unroll
        for (int i = 0; i < 8; i++) d[i] = out[i];
    }
}

// Scene arrays repacked on the device from the description's own layouts (one pass each; the host only stages bytes).
__global__ void __launch_bounds__(256) k_pack_prims(int n_prims, const float* __restrict__ prim_geom, const unsigned* __restrict__ prim_flags,
                                                    const int* __restrict__ prim_material, const int* __restrict__ prim_texture,
                                                    const int* __restrict__ prim_normalmap, const int* __restrict__ prim_object,
                                                    const float* __restrict__ prim_normals, const float* __restrict__ prim_uv,
                                                    const float* __restrict__ materials, int n_materials, int n_textures, int n_normalmaps,
                                                    float4* geom, float4* attr, int4* ids, float4* tri_aux, int* flags_out) {
    int bad = 0;
    float slack = 0.f;                                      // prune_rule.h: scene maximum of the triangles' acceptance slack
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_prims; p += gridDim.x * blockDim.x) {
        const float* g = prim_geom + 12 * (size_t)p;
        const int mat = prim_material[p];
        const unsigned flags = prim_flags[p];
        float alpha = 1.f;
        if (mat >= 0 && mat < n_materials) alpha = materials[12 * (size_t)mat + 10];        // WrtMaterial::alpha
        else bad |= 4;
        const float oma = 1 - alpha;                        // (1 - inter.mtlcolor.alpha), BVHStrategy.hpp:38
        geom[3 * (size_t)p + 0] = make_float4(g[0], g[1], g[2], g[3]);
        geom[3 * (size_t)p + 1] = make_float4(g[4], g[5], g[6], oma);
        geom[3 * (size_t)p + 2] = make_float4(g[8], g[9], g[10], __uint_as_float(flags));
        float aux[4] = {0.f, 0.f, 0.f, -1.f};
        if ((flags & WRT_PRIM_KIND_MASK) == WRT_PRIM_TRIANGLE) {
            wrt_triangle_aux(g + 4, g + 8, aux);
            slack = fmaxf(slack, wrt_prune_triangle_slack(g + 4, g + 8));
        }
        tri_aux[p] = make_float4(aux[0], aux[1], aux[2], aux[3]);
        const float* nn = prim_normals + 9 * (size_t)p;
        const float* uv = prim_uv + 6 * (size_t)p;
        attr[4 * (size_t)p + 0] = make_float4(nn[0], nn[1], nn[2], uv[0]);
        attr[4 * (size_t)p + 1] = make_float4(nn[3], nn[4], nn[5], uv[1]);
        attr[4 * (size_t)p + 2] = make_float4(nn[6], nn[7], nn[8], uv[2]);
        attr[4 * (size_t)p + 3] = make_float4(uv[3], uv[4], uv[5], 0.f);
        ids[p] = make_int4(mat, prim_texture[p], prim_normalmap[p], prim_object[p]);
        if (flags & WRT_PRIM_LIGHT) bad |= 1;
        if (prim_texture[p] >= n_textures || prim_normalmap[p] >= n_normalmaps) bad |= 2;
    }
    bad = __reduce_or_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicOr(flags_out, bad);       // bit 0: a light avatar exists; 1: a map is missing; 2: bad material
    const unsigned sb = __reduce_max_sync(0xffffffffu, __float_as_uint(slack));   // (non-negative floats order like their bits)
    if ((threadIdx.x & 31) == 0 && sb) atomicMax(reinterpret_cast<unsigned*>(flags_out) + 1, sb);
}

// WrtMaterial (12 floats) -> {Od.rgb, ka} {Os.rgb, kd} {ks, n, alpha, eta}
__global__ void __launch_bounds__(256) k_pack_materials(const float* __restrict__ materials, int n, float4* out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float* m = materials + 12 * (size_t)i;
        out[3 * (size_t)i + 0] = make_float4(m[0], m[1], m[2], m[6]);
        out[3 * (size_t)i + 1] = make_float4(m[3], m[4], m[5], m[7]);
        out[3 * (size_t)i + 2] = make_float4(m[8], m[9], m[10], m[11]);
    }
}

} // namespace wrt
