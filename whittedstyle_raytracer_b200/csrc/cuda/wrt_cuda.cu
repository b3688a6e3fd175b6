// wrt_cuda.cu — host side of libwrt_cuda.so: context, scene upload, frame
// scheduling and the C ABI of include/wrt_cuda.h.  Compiled for sm_100a only,
// with -fmad=false (see dev_math.cuh).  No CPU fallback exists in this library.
#include <algorithm>
#include <cstdio>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../../include/wrt_cuda.h"
#include "fast_bvh.hpp"
#include "kernels.cuh"

namespace {

thread_local std::string g_err;

int fail(const std::string& msg) {
    g_err = msg;
    return 1;
}

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            return fail(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
                        std::to_string(__LINE__) + ")");                                              \
        }                                                                                             \
    } while (0)

enum Family { F_RAYGEN, F_TRACE, F_SURFACE, F_SHADOW_HARD, F_SHADOW_SOFT, F_SHADOW_DIR, F_SHADE, F_COMBINE, F_RESOLVE, F_SOFT_LISTS };

struct TimedLaunch {
    int family;
    cudaEvent_t e0, e1;
};

} // namespace

struct WrtContext {
    int device = 0;
    int num_sms = 0;
    // All frame work runs on two internal streams: `chain` (high priority: the dependent closest-hit chain, the deep
    // levels' shadow kernels, combine + resolve) and `side` (level 0's shadow + shade kernels).  A caller-provided stream
    // (wrt_render_device) is joined to them by two events, nothing else runs on it.
    cudaStream_t chain = nullptr, side = nullptr;
    cudaEvent_t ev_in = nullptr, ev_lvl0 = nullptr, ev_side = nullptr, ev_out = nullptr;
    bool overlap = true;

    // scene
    wrt::DevScene ds{};
    std::vector<void*> scene_allocs;
    std::vector<size_t> scene_alloc_bytes;
    size_t scene_slot = 0;
    bool has_scene = false;
    bool textures_complete = true;
    int bvh_depth = 0;
    int stack_rows = 2;

    // camera / tiling / options
    WrtCamera cam{};
    bool has_cam = false;
    int tile_w = 8, tile_h = 4, rank = 0, world = 1;   // one warp-sized 8x4 block per tile: finest interleave
    int traversal = WRT_TRAVERSAL_PRUNED;
    uint32_t seed = WRT_DEFAULT_SEED;
    float queue_factor = 0.f;          // > 0: caller-fixed deep-level capacity (multiple of the batch); 0: automatic
    float deep_factor = 0.25f;         // automatic mode: current deep-level capacity factor (grows on overflow)
    float prune_rel = 1e-3f;
    bool kernel_timing = false;
    int refill = 16;                   // idle lanes that trigger a refill on deep ray-tree levels
    int refill_soft = 24;
    bool shaft_cull = true;            // soft shadows: answer requests whose light shaft is empty without tracing (shaft_cull.h)
    bool soft_lists = true;            // soft shadows: per-request candidate lists (k_soft_lists + k_soft_list_rays) instead of per-ray walks
    long long list_pool_cap_override = 0;
    int shaft_cull_max_level = 0;      // deepest ray-tree level whose surface stage runs the shaft test (when lists are on)
    wrt::SoftListBuffers list_bufs[2] = {};
    wrt::FastBvhBuilder fbvh;          // host scratch of wrt_upload_scene, kept between uploads
    std::vector<WrtNode> h_oct;
    bool unlit_cull = true;            // drop shadow requests of lights whose shading terms are exactly 0 at the point
    int cache_from_level = 0;          // per-ray soft-shadow kernel: occluder cache (99 = off)
    int chunk_div = 16;                // work claiming: 0 = one atomic per refill, k = chunks of n/(warps*k) items
    int refill0 = 32;                  // level 0 (coherent primary rays and their shadow rays)
    int trace_blocks_per_sm = 10;      // persistent shadow / unfused closest-hit kernels
    int fused_blocks_per_sm = WRT_FUSED_MIN_BLOCKS;
    int fuse_from = 99;                // levels >= this run the fused closest-hit + surface kernel (99: never; experimental, see kernels.cuh)
    bool shade0_separate = true;       // level 0 is shaded by its own launch on the side stream (else inside combine)
    bool small_batch_full_levels = true; // automatic sizing: batches under 1 M slots get full-size deep levels
    long long max_batch = 1ll << 25;   // primary slots per batch (8K = 33.2 M slots fits)

    // frame buffers
    wrt::FrameBuffers fb{};
    std::vector<void*> frame_allocs;
    unsigned batch_slots = 0;          // primary slots the buffers were sized for
    unsigned deep_slots = 0;           // slots of each deeper level
    int fb_lights = -1, fb_point = -1, fb_dir = -1;
    unsigned* h_counters = nullptr;    // pinned
    uint8_t* d_image = nullptr;        // full image, row-major (wrt_render)
    size_t d_image_bytes = 0;
    uint8_t* h_image = nullptr;        // pinned staging
    size_t h_image_bytes = 0;

    // batch-query scratch
    void* d_scratch[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t scratch_bytes[5] = {0, 0, 0, 0, 0};

    // bookkeeping
    int64_t launches = 0;
    int work_seq = 0;
    int coop_grid = 0;                 // co-resident CTAs of k_combine_resolve (cooperative launch)
    WrtStats stats{};
    std::vector<TimedLaunch> timed;
    std::vector<cudaEvent_t> event_pool;
    size_t event_next = 0;
    float family_ms[WRT_KERNEL_FAMILIES] = {0};
    int family_launches[WRT_KERNEL_FAMILIES] = {0};
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    cudaStream_t user_stream = nullptr;
    bool user_stream_joined = false;
    uint8_t* last_image = nullptr;     // targets of the frame in flight (needed again if an overflowed batch is redone)
    uint8_t* last_packed = nullptr;
    bool frame_pending = false;
    bool has_frame = false;

    wrt::TileMap tilemap() const {
        wrt::TileMap tm;
        static_cast<WrtTileMap&>(tm) = wrt_tilemap_make(cam.width, cam.height, tile_w, tile_h, rank, world);
        return tm;
    }
    long long local_slots(int r, int w) const {
        WrtTileMap tm = wrt_tilemap_make(cam.width, cam.height, tile_w, tile_h, r, w);
        return wrt_tilemap_slots(&tm, r, w);
    }
};

namespace {

using wrt::FrameBuffers;

// Scene arrays keep their device allocation across uploads when the new array fits
// (re-uploading a scene of the same size costs copies only, no cudaMalloc/cudaFree).
template <class T>
int dev_upload(WrtContext* c, const T* src, size_t count, const T** dst) {
    *dst = nullptr;
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    size_t slot = c->scene_slot++;
    if (slot >= c->scene_allocs.size()) { c->scene_allocs.push_back(nullptr); c->scene_alloc_bytes.push_back(0); }
    if (c->scene_alloc_bytes[slot] < bytes) {
        if (c->scene_allocs[slot]) cudaFree(c->scene_allocs[slot]);
        c->scene_allocs[slot] = nullptr; c->scene_alloc_bytes[slot] = 0;
        CK(cudaMalloc(&c->scene_allocs[slot], bytes));
        c->scene_alloc_bytes[slot] = bytes;
    }
    void* p = c->scene_allocs[slot];
    if (count) CK(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, c->chain));
    *dst = (const T*)p;
    return 0;
}

void free_scene(WrtContext* c) {
    for (void* p : c->scene_allocs) if (p) cudaFree(p);
    c->scene_allocs.clear();
    c->scene_alloc_bytes.clear();
    c->has_scene = false;
}

void free_frame(WrtContext* c) {
    for (void* p : c->frame_allocs) cudaFree(p);
    c->frame_allocs.clear();
    c->batch_slots = 0;
    c->deep_slots = 0;
    memset(&c->fb, 0, sizeof c->fb);
}

int tree_depth(const WrtSceneDesc* s) {
    if (s->n_nodes == 0) return 0;
    int maxd = 0;
    std::vector<std::pair<int, int>> st;
    st.push_back({0, 0});
    while (!st.empty()) {
        auto [n, d] = st.back();
        st.pop_back();
        maxd = std::max(maxd, d);
        int link = s->nodes[n].link;
        if (link >= 0) { st.push_back({link, d + 1}); st.push_back({link + 1, d + 1}); }
    }
    return maxd;
}

template <class T>
int frame_alloc(WrtContext* c, T** p, size_t count) {
    void* q = nullptr;
    CK(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
    c->frame_allocs.push_back(q);
    *p = (T*)q;
    return 0;
}

// Slots of each deeper level for a batch of `slots` primary slots.  Caller-fixed factor: factor x slots.  Automatic:
// small batches get a full-size level (memory is no object there), large ones `deep_factor` of the batch — in the bunny
// frames <= 16 % of the pixels spawn secondary rays; the factor doubles when a level overflows.
unsigned deep_slots_for(const WrtContext* c, unsigned slots) {
    double want;
    if (c->queue_factor > 0.f) want = (double)slots * c->queue_factor;
    else want = std::max((double)slots * c->deep_factor, c->small_batch_full_levels ? std::min((double)slots, 1048576.0) : 0.0);
    want = std::min(want, 2.0e9);
    return ((unsigned)std::max(want, 64.0) + 63u) & ~63u;
}

int ensure_frame_buffers(WrtContext* c, unsigned slots) {
    const wrt::DevScene& ds = c->ds;
    const unsigned capd = deep_slots_for(c, slots);
    if (c->batch_slots >= slots && c->deep_slots >= capd && c->fb_lights == ds.n_lights && c->fb_point == ds.n_point_lights &&
        c->fb_dir == ds.n_dir_lights)
        return 0;
    CK(cudaDeviceSynchronize());
    free_frame(c);
    FrameBuffers& fb = c->fb;
    const unsigned cap0 = slots + 64;
    const unsigned long long nodes64 = (unsigned long long)cap0 + (unsigned long long)(WRT_MAX_DEPTH - 1) * capd;
    if (nodes64 > 0x7fffff00ull) return fail("frame batch too large for 32-bit node indices");
    if (nodes64 * std::max(1, ds.n_lights) > 0xffffffffffull) return fail("too many (node, light) pairs");
    fb.cap0 = cap0; fb.capd = capd; fb.n_node_cap = (unsigned)nodes64;
    const unsigned long long deep_nodes = (unsigned long long)(WRT_MAX_DEPTH - 1) * capd;
    auto req_cap = [](unsigned long long nodes, int lights) {
        return (unsigned)std::min<unsigned long long>(nodes * (unsigned long long)std::max(1, lights), 0xffffff00ull);
    };
    fb.preq_cap[0] = req_cap(cap0, ds.n_point_lights); fb.preq_cap[1] = req_cap(deep_nodes, ds.n_point_lights);
    fb.dreq_cap[0] = req_cap(cap0, ds.n_dir_lights);   fb.dreq_cap[1] = req_cap(deep_nodes, ds.n_dir_lights);
    for (int k = 0; k < 2; k++) {
        if (frame_alloc(c, &fb.ray_o[k], capd)) return 1;
        if (frame_alloc(c, &fb.ray_d[k], capd)) return 1;
    }
    if (frame_alloc(c, &fb.hit, std::max(cap0, capd))) return 1;
    if (frame_alloc(c, &fb.surf, 4 * (size_t)nodes64)) return 1;
    if (frame_alloc(c, &fb.node_a, (size_t)nodes64)) return 1;
    if (frame_alloc(c, &fb.node_b, (size_t)nodes64)) return 1;
    if (frame_alloc(c, &fb.coeff, (size_t)nodes64 * std::max(1, ds.n_lights))) return 1;
    for (int q = 0; q < 2; q++) {
        if (frame_alloc(c, &fb.preq_o[q], ds.n_point_lights ? fb.preq_cap[q] : 1)) return 1;
        if (frame_alloc(c, &fb.preq_k[q], ds.n_point_lights ? fb.preq_cap[q] : 1)) return 1;
        if (frame_alloc(c, &fb.dreq_o[q], ds.n_dir_lights ? fb.dreq_cap[q] : 1)) return 1;
        if (frame_alloc(c, &fb.dreq_k[q], ds.n_dir_lights ? fb.dreq_cap[q] : 1)) return 1;
    }
    if (frame_alloc(c, &fb.counters, wrt::C_TOTAL)) return 1;
    // candidate lists of the soft-shadow path (kernels.cuh, K4b'), one set per request queue (their launches overlap):
    // walk scratch per thread, list pool, per-request {offset, count}.  A full pool only means per-ray walks for the
    // remaining requests.  Level-0 requests that survive the shaft test at spawn time have 1-2 candidates; deep-level
    // ones (origins on the bunny) ~30.
    for (int q = 0; q < 2; q++) {
        wrt::SoftListBuffers& lb = c->list_bufs[q];
        const unsigned long long per_req = q == 0 ? 2ull : 12ull;
        lb.pool_cap = (unsigned)std::min<unsigned long long>(std::max<unsigned long long>(8ull << 20, per_req * fb.preq_cap[q]), 1ull << 30);
        if (!ds.n_point_lights || ds.shadow_type == 0) lb.pool_cap = 64;
        if (c->list_pool_cap_override > 0) lb.pool_cap = (unsigned)c->list_pool_cap_override;     // tests: force the pool-full path
        const bool lists_possible = ds.n_point_lights && ds.shadow_type != 0;
        if (frame_alloc(c, &lb.scratch, lists_possible ? (size_t)c->num_sms * c->trace_blocks_per_sm * 128 * WRT_LIST_CAP : 1)) return 1;
        if (frame_alloc(c, &lb.pool, lb.pool_cap)) return 1;
        if (frame_alloc(c, &lb.ref, lists_possible ? fb.preq_cap[q] : 1)) return 1;
    }
    c->batch_slots = slots;
    c->deep_slots = capd;
    c->fb_lights = ds.n_lights; c->fb_point = ds.n_point_lights; c->fb_dir = ds.n_dir_lights;
    return 0;
}

int ensure_scratch(WrtContext* c, int k, size_t bytes) {
    if (c->scratch_bytes[k] >= bytes) return 0;
    if (c->d_scratch[k]) cudaFree(c->d_scratch[k]);
    c->d_scratch[k] = nullptr;
    c->scratch_bytes[k] = 0;
    CK(cudaMalloc(&c->d_scratch[k], bytes));
    c->scratch_bytes[k] = bytes;
    return 0;
}

cudaEvent_t next_event(WrtContext* c) {
    if (c->event_next == c->event_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        c->event_pool.push_back(e);
    }
    return c->event_pool[c->event_next++];
}

struct LaunchScope {                      // counts the launch; optionally brackets it with events
    WrtContext* c;
    cudaStream_t st;
    TimedLaunch tl{};
    bool timed;
    LaunchScope(WrtContext* ctx, cudaStream_t s, int family) : c(ctx), st(s), timed(ctx->kernel_timing) {
        ++c->launches;
        if (timed) {
            tl.family = family;
            tl.e0 = next_event(c);
            tl.e1 = next_event(c);
            cudaEventRecord(tl.e0, st);
        }
    }
    ~LaunchScope() {
        if (timed) {
            cudaEventRecord(tl.e1, st);
            c->timed.push_back(tl);
        }
    }
};

int grid_for(WrtContext* c, int blocks_per_sm) { return c->num_sms * blocks_per_sm; }

size_t stack_bytes(WrtContext* c, int threads) { return (size_t)c->stack_rows * threads * sizeof(int); }

float prune_value(const WrtContext* c) { return c->traversal == WRT_TRAVERSAL_PRUNED ? c->prune_rel : -1.f; }

// The shadow kernels of request queue q (0: level 0, 1: levels 1..8 together) on stream `st`.
int enqueue_shadows(WrtContext* c, cudaStream_t st, int q, int& work_seq) {
    using namespace wrt;
    FrameBuffers& fb = c->fb;
    const DevScene& ds = c->ds;
    const int TB = 128;
    const int trace_grid = grid_for(c, c->trace_blocks_per_sm), wide_grid = grid_for(c, 8);
    const size_t sb = stack_bytes(c, TB);
    auto work_slot = [&]() { int s = C_WORK + 2 * work_seq; ++work_seq; return s; };
    const int refill = (q == 0 ? c->refill0 : c->refill) | (c->chunk_div << 8);
    if (ds.n_point_lights > 0) {
        if (ds.shadow_type == 0) {
            LaunchScope ls(c, st, F_SHADOW_HARD);
            k_shadow_hard<<<trace_grid, TB, sb, st>>>(ds, fb, q, work_slot(), refill, c->traversal == WRT_TRAVERSAL_EXHAUSTIVE ? 1 : 0);
        } else {
            // per-request candidate lists (kernels.cuh, K4b'); scenes with light avatars keep the per-ray kernel,
            // whose literal hasIntersection path they need, and so does WRT_TRAVERSAL_EXHAUSTIVE
            const bool lists = c->soft_lists && !ds.has_light_prims && ds.n_nodes > 0 && c->traversal == WRT_TRAVERSAL_PRUNED &&
                               (unsigned long long)fb.preq_cap[q] * WRT_SOFT_SAMPLES < (1ull << 32);
            if (lists) {
                const SoftListBuffers& lb = c->list_bufs[q];
                {
                    LaunchScope ls(c, st, F_SOFT_LISTS);
                    k_soft_lists<<<trace_grid, TB, sb, st>>>(ds, fb, q, work_slot(), c->stack_rows, lb);
                }
                LaunchScope ls(c, st, F_SHADOW_SOFT);
                k_soft_list_rays<<<trace_grid, TB, sb, st>>>(ds, fb, q, work_slot(), c->seed, lb);
            } else {
                LaunchScope ls(c, st, F_SHADOW_SOFT);
                k_shadow_soft<<<trace_grid, TB, sb, st>>>(ds, fb, q, work_slot(), c->seed, c->refill_soft | (c->chunk_div << 8),
                                                          c->cache_from_level <= (q == 0 ? 0 : 1) ? 1 : 0);
            }
        }
    }
    if (ds.n_dir_lights > 0) {
        LaunchScope ls(c, st, F_SHADOW_DIR);
        k_shadow_directional<<<wide_grid, TB, sb, st>>>(ds, fb, q, c->traversal == WRT_TRAVERSAL_EXHAUSTIVE ? 1 : 0);
    }
    return 0;
}

// Enqueues one batch of primary slots [slot0, slot0+n): closest-hit chain + deep shadows on `chain`, level 0's shadow
// work on `side`.
int enqueue_batch(WrtContext* c, long long slot0, unsigned n, uint8_t* d_image, uint8_t* d_packed) {
    using namespace wrt;
    FrameBuffers& fb = c->fb;
    const DevScene& ds = c->ds;
    cudaStream_t st = c->chain;
    const bool overlap = c->overlap && !c->kernel_timing;
    cudaStream_t ss = overlap ? c->side : st;
    PrimaryGen pg;
    pg.cam = c->cam; pg.tm = c->tilemap(); pg.slot0 = slot0;
    CK(cudaMemsetAsync(fb.counters, 0, C_TOTAL * sizeof(unsigned), st));
    int work_seq = 0;
    auto work_slot = [&]() { int s = C_WORK + 2 * work_seq; ++work_seq; return s; };
    const int TB = 128;
    const int trace_grid = grid_for(c, c->trace_blocks_per_sm), fused_grid = grid_for(c, c->fused_blocks_per_sm), wide_grid = grid_for(c, 8);
    const size_t sb = stack_bytes(c, TB);
    const float prune = prune_value(c);
    auto cull_for = [&](int d) {
        // request culling belongs to the pruned mode; WRT_TRAVERSAL_EXHAUSTIVE traces every ray the reference traces
        int cull = 0;
        if (c->traversal == WRT_TRAVERSAL_PRUNED) {
            if (c->unlit_cull) cull |= WRT_CULL_UNLIT;
            // The shaft test at spawn time pays where most shafts are empty (primary hits: 79 % in the metric frame); on
            // deeper levels (~8 %) k_soft_lists finds the empty ones anyway (an empty list), off the critical
            // closest-hit chain.  Without the list kernels the test runs on every level.
            const bool lists_on = c->soft_lists && !ds.has_light_prims;
            if (c->shaft_cull && ds.shadow_type != 0 && (d <= c->shaft_cull_max_level || !lists_on)) cull |= WRT_CULL_SHAFT;
        }
        return cull;
    };
    auto level_kernels = [&](int d) {
        const int refill = (d == 0 ? c->refill0 : c->refill) | (c->chunk_div << 8);
        if (d >= c->fuse_from) {
            LaunchScope ls(c, st, F_TRACE);
            if (d == 0) k_trace_surface<true><<<fused_grid, TB, sb, st>>>(ds, fb, pg, d, n, work_slot(), prune, refill, cull_for(d));
            else k_trace_surface<false><<<fused_grid, TB, sb, st>>>(ds, fb, pg, d, n, work_slot(), prune, refill, cull_for(d));
        } else {
            {
                LaunchScope ls(c, st, F_TRACE);
                if (d == 0) k_trace_closest<true><<<trace_grid, TB, sb, st>>>(ds, fb, pg, d, n, work_slot(), prune, refill);
                else k_trace_closest<false><<<trace_grid, TB, sb, st>>>(ds, fb, pg, d, n, work_slot(), prune, refill);
            }
            LaunchScope ls(c, st, F_SURFACE);
            k_surface_spawn<<<wide_grid, 256, 0, st>>>(ds, fb, pg, d, n, cull_for(d));
        }
    };
    level_kernels(0);
    if (overlap) {
        CK(cudaEventRecord(c->ev_lvl0, st));
        CK(cudaStreamWaitEvent(ss, c->ev_lvl0, 0));
    }
    // level 0's shadow + shade kernels run beside the deep chain: a deep level holds few, long, incoherent rays and
    // leaves most of the SMs idle; a persistent traversal kernel ends with its longest ray
    if (enqueue_shadows(c, ss, 0, work_seq)) return 1;
    if (c->shade0_separate) {
        LaunchScope ls(c, ss, F_SHADE);
        k_shade<<<wide_grid, 256, 0, ss>>>(ds, fb, n, 0, 0);
    }
    if (overlap) CK(cudaEventRecord(c->ev_side, ss));
    for (int d = 1; d < WRT_MAX_DEPTH; d++) level_kernels(d);
    if (enqueue_shadows(c, st, 1, work_seq)) return 1;
    if (overlap) CK(cudaStreamWaitEvent(st, c->ev_side, 0));
    {
        LaunchScope ls(c, st, F_COMBINE);
        int shade_from = c->shade0_separate ? 1 : 0;
        TileMap tm = pg.tm;
        void* args[] = {(void*)&ds, (void*)&fb, (void*)&tm, (void*)&slot0, (void*)&n, (void*)&shade_from, (void*)&d_image, (void*)&d_packed};
        CK(cudaLaunchCooperativeKernel((const void*)k_combine_resolve, dim3(c->coop_grid), dim3(256), args, 0, st));
    }
    CK(cudaGetLastError());
    if (C_WORK + 2 * work_seq > C_TOTAL) return fail("internal: work counters exceed the counter block");
    return 0;
}

void add_batch_stats(WrtContext* c, const unsigned* cnt) {
    WrtStats& s = c->stats;
    const wrt::DevScene& ds = c->ds;
    s.rays_per_depth[0] += cnt[wrt::C_VALID0];
    s.closest_rays += cnt[wrt::C_VALID0];
    for (int d = 1; d < WRT_MAX_DEPTH; d++) {
        s.rays_per_depth[d] += cnt[wrt::C_NRAYS + d] + cnt[wrt::C_NTRAYS + d];
        s.closest_rays += cnt[wrt::C_NRAYS + d] + cnt[wrt::C_NTRAYS + d];
    }
    // requests answered without tracing count like the reference counts them (it traces them)
    int64_t culled = 0, skip_p = 0, skip_d = 0;
    for (int d = 0; d < WRT_MAX_DEPTH; d++) { culled += cnt[wrt::C_NCULL + d]; skip_p += cnt[wrt::C_NSKIP + d]; skip_d += cnt[wrt::C_NDSKIP + d]; }
    const int64_t queued_p = (int64_t)cnt[wrt::C_NPREQ] + cnt[wrt::C_NPREQ + 1], queued_d = (int64_t)cnt[wrt::C_NDREQ] + cnt[wrt::C_NDREQ + 1];
    const int64_t empty = cnt[wrt::C_NEMPTY];          // queued, but their candidate list came out empty: no rays built
    const int64_t p = queued_p + culled + skip_p, q = queued_d + skip_d;
    const int64_t per = ds.shadow_type ? WRT_SOFT_SAMPLES : 1;
    s.shadow_requests += p + q;
    s.shadow_rays += p * per + q;
    s.shaft_culled_requests += culled + empty;
    s.unlit_skipped_requests += skip_p + skip_d;
    s.shadow_rays_traced += (queued_p - empty) * per + queued_d;
}

#ifdef WRT_DEBUG_BOUNDS
int check_debug_bounds() {
    unsigned line = 0, stack = 0;
    CK(cudaMemcpyFromSymbol(&line, wrt::g_wrt_bounds_line, sizeof line));
    CK(cudaMemcpyFromSymbol(&stack, wrt::g_wrt_stack_overflow, sizeof stack));
    if (line || stack) {                                // report once, then re-arm
        const unsigned zero = 0;
        cudaMemcpyToSymbol(wrt::g_wrt_bounds_line, &zero, sizeof zero);
        cudaMemcpyToSymbol(wrt::g_wrt_stack_overflow, &zero, sizeof zero);
    }
    if (line) return fail("WRT_DEBUG_BOUNDS: index out of bounds at kernels.cuh:" + std::to_string(line));
    if (stack) return fail("WRT_DEBUG_BOUNDS: traversal stack overflow (sp " + std::to_string(stack - 1) + ")");
    return 0;
}
#else
int check_debug_bounds() { return 0; }
#endif

// A batch overflowed a deep-level queue.  Automatic sizing: double the deep-level capacity (kept for the following
// frames) while it is below 2x the batch; otherwise (or with a caller-fixed factor) the batch is split in halves.
// Returns true when the buffers grew and the same batch should simply be rendered again.
bool grow_after_overflow(WrtContext* c) {
    const float effective = (float)c->deep_slots / (float)std::max(1u, c->batch_slots);
    if (c->queue_factor > 0.f || effective >= 2.f) return false;
    const float old = c->deep_factor;
    c->deep_factor = std::min(2.f, std::max(old, effective) * 2.f);
    if (ensure_frame_buffers(c, c->batch_slots) == 0) return true;
    c->deep_factor = old;                              // out of memory: fall back to halving
    cudaGetLastError();
    return false;
}

// Renders the spans on `todo` (popped from the back) synchronously, re-rendering what overflows.
int render_spans_sync(WrtContext* c, std::vector<std::pair<long long, long long>>& todo, uint8_t* d_image, uint8_t* d_packed) {
    cudaStream_t st = c->chain;
    while (!todo.empty()) {
        auto [s0, n] = todo.back();
        todo.pop_back();
        if (n <= 0) continue;
        if (enqueue_batch(c, s0, (unsigned)n, d_image, d_packed)) return 1;
        CK(cudaMemcpyAsync(c->h_counters, c->fb.counters, wrt::C_TOTAL * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (c->h_counters[wrt::C_OVERFLOW]) {
            ++c->stats.overflow_retries;
            if (grow_after_overflow(c)) { todo.push_back({s0, n}); continue; }
            if (n <= 64) return fail("ray queue overflow on a 64-slot batch: raise queue_factor");
            long long half = ((n / 2 + 31) / 32) * 32;
            todo.push_back({s0 + half, n - half});
            todo.push_back({s0, half});
            continue;
        }
        add_batch_stats(c, c->h_counters);
    }
    return 0;
}

// Renders all local slots.  A single-batch frame (the common case) is enqueued without any host synchronisation;
// overflow is checked — and the frame redone — in finish_frame.
int render_all(WrtContext* c, cudaStream_t user, bool join_user, uint8_t* d_image, uint8_t* d_packed) {
    if (!c->has_scene) return fail("wrt_render: no scene uploaded");
    if (!c->has_cam) return fail("wrt_render: no camera set");
    if (c->cam.width <= 0 || c->cam.height <= 0) return fail("wrt_render: empty image");
    if (!c->textures_complete) return fail("wrt_render: a primitive references a texture / normal map that was not uploaded");
    const long long total = c->local_slots(c->rank, c->world);
    const long long max_batch = c->max_batch;
    memset(&c->stats, 0, sizeof c->stats);
    c->timed.clear();
    c->event_next = 0;
    c->user_stream = user;
    c->user_stream_joined = join_user;
    c->last_image = d_image; c->last_packed = d_packed;
    c->has_frame = true;
    cudaStream_t st = c->chain;
    if (total > 0) {
        unsigned want = (unsigned)std::min(total, max_batch);
        if (ensure_frame_buffers(c, want)) return 1;
    }
    if (join_user) {                                    // everything the caller enqueued before is visible to the frame
        CK(cudaEventRecord(c->ev_in, user));
        CK(cudaStreamWaitEvent(st, c->ev_in, 0));
    }
    CK(cudaEventRecord(c->ev_begin, st));
    std::vector<std::pair<long long, long long>> todo;
    for (long long s0 = total; s0 > 0;) {              // push in reverse so spans pop in order
        long long b = (s0 - 1) / max_batch * max_batch;
        todo.push_back({b, s0 - b});
        s0 = b;
    }
    if (todo.size() == 1) {                            // stay asynchronous, check in finish
        if (enqueue_batch(c, todo[0].first, (unsigned)todo[0].second, d_image, d_packed)) return 1;
        CK(cudaMemcpyAsync(c->h_counters, c->fb.counters, wrt::C_TOTAL * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        c->frame_pending = true;
    } else if (render_spans_sync(c, todo, d_image, d_packed)) return 1;
    CK(cudaEventRecord(c->ev_end, st));
    if (join_user) {
        CK(cudaEventRecord(c->ev_out, st));
        CK(cudaStreamWaitEvent(user, c->ev_out, 0));
    }
    return 0;
}

int finish_frame(WrtContext* c) {
    cudaStream_t st = c->chain;
    CK(cudaStreamSynchronize(st));
    if (c->frame_pending) {
        c->frame_pending = false;
        if (c->h_counters[wrt::C_OVERFLOW]) {
            // the single-batch frame overflowed a queue: redo it synchronously (larger deep levels, or in halves)
            const long long total = c->local_slots(c->rank, c->world);
            memset(&c->stats, 0, sizeof c->stats);
            c->stats.overflow_retries = 1;
            std::vector<std::pair<long long, long long>> todo;
            if (grow_after_overflow(c)) todo.push_back({0, total});
            else {
                if (total <= 64) return fail("ray queue overflow on a 64-slot batch: raise queue_factor");
                long long half = ((total / 2 + 31) / 32) * 32;
                todo.push_back({half, total - half});
                todo.push_back({0, half});
            }
            if (render_spans_sync(c, todo, c->last_image, c->last_packed)) return 1;
            CK(cudaEventRecord(c->ev_end, st));
            CK(cudaStreamSynchronize(st));
        } else {
            add_batch_stats(c, c->h_counters);
        }
    }
    if (check_debug_bounds()) return 1;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->ev_begin, c->ev_end) == cudaSuccess) c->stats.gpu_ms = ms;
    for (float& f : c->family_ms) f = 0.f;
    for (int& n : c->family_launches) n = 0;
    for (const TimedLaunch& tl : c->timed) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, tl.e0, tl.e1) == cudaSuccess) { c->family_ms[tl.family] += t; ++c->family_launches[tl.family]; }
    }
    return 0;
}

} // namespace

extern "C" {

const char* wrt_last_error(void) { return g_err.c_str(); }

int wrt_create(int device, WrtContext** out) {
    if (!out) return fail("wrt_create: null out pointer");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(std::string("wrt_create: no CUDA device (") + cudaGetErrorString(e) +
                    "); this library has no CPU fallback");
    if (device < 0 || device >= n) return fail("wrt_create: device index out of range");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(std::string("wrt_create: ") + prop.name + " is sm_" + std::to_string(prop.major) +
                    std::to_string(prop.minor) + "; this library is built for sm_100a only");
    WrtContext* c = new WrtContext();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    bool ev_ok = cudaStreamCreateWithPriority(&c->chain, cudaStreamNonBlocking, prio_greatest) == cudaSuccess &&
                 cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, prio_least) == cudaSuccess;
    for (cudaEvent_t* e : {&c->ev_in, &c->ev_lvl0, &c->ev_side, &c->ev_out})
        ev_ok = ev_ok && cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess;
    if (!ev_ok ||
        cudaEventCreate(&c->ev_begin) != cudaSuccess || cudaEventCreate(&c->ev_end) != cudaSuccess ||
        cudaMallocHost((void**)&c->h_counters, wrt::C_TOTAL * sizeof(unsigned)) != cudaSuccess) {
        delete c;
        return fail("wrt_create: stream/event/pinned allocation failed");
    }
    memset(c->h_counters, 0, wrt::C_TOTAL * sizeof(unsigned));
    {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wrt::k_combine_resolve, 256, 0) != cudaSuccess || per_sm < 1) {
            wrt_destroy(c);
            return fail("wrt_create: cooperative launch of k_combine_resolve is not possible on this device");
        }
        c->coop_grid = c->num_sms * std::min(per_sm, 4);
    }
    // tuning overrides (development only; defaults are what bench.py measures)
    if (const char* e = getenv("WRT_REFILL")) c->refill = std::max(1, std::min(32, atoi(e)));
    if (const char* e = getenv("WRT_REFILL_SOFT")) c->refill_soft = std::max(1, std::min(32, atoi(e)));
    if (const char* e = getenv("WRT_TRACE_BLOCKS")) c->trace_blocks_per_sm = std::max(1, std::min(32, atoi(e)));
    if (const char* e = getenv("WRT_OVERLAP")) c->overlap = atoi(e) != 0;
    if (const char* e = getenv("WRT_CACHE_FROM")) c->cache_from_level = atoi(e);
    if (const char* e = getenv("WRT_SHAFT_CULL")) c->shaft_cull = atoi(e) != 0;
    if (const char* e = getenv("WRT_UNLIT_CULL")) c->unlit_cull = atoi(e) != 0;
    if (const char* e = getenv("WRT_SOFT_LISTS")) c->soft_lists = atoi(e) != 0;
    if (const char* e = getenv("WRT_LIST_POOL_CAP")) c->list_pool_cap_override = atoll(e);
    if (const char* e = getenv("WRT_SHAFT_LEVELS")) c->shaft_cull_max_level = atoi(e);
    if (const char* e = getenv("WRT_CHUNK_DIV")) c->chunk_div = std::max(0, std::min(255, atoi(e)));
    if (const char* e = getenv("WRT_REFILL0")) c->refill0 = std::max(1, std::min(32, atoi(e)));
    if (const char* e = getenv("WRT_FUSED_BLOCKS")) c->fused_blocks_per_sm = std::max(1, std::min(32, atoi(e)));
    if (const char* e = getenv("WRT_FUSE_FROM")) c->fuse_from = atoi(e);
    if (const char* e = getenv("WRT_SHADE0_SEPARATE")) c->shade0_separate = atoi(e) != 0;
    if (const char* e = getenv("WRT_DEEP_FACTOR")) { c->deep_factor = std::max(0.001f, std::min(2.f, (float)atof(e))); c->small_batch_full_levels = false; }
    if (const char* e = getenv("WRT_MAX_BATCH")) c->max_batch = std::max(64ll, (atoll(e) + 31) / 32 * 32);   // tests: force multi-batch frames
    *out = c;
    return 0;
}

void wrt_destroy(WrtContext* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    free_scene(c);
    free_frame(c);
    for (int k = 0; k < 5; k++) if (c->d_scratch[k]) cudaFree(c->d_scratch[k]);
    if (c->d_image) cudaFree(c->d_image);
    if (c->h_image) cudaFreeHost(c->h_image);
    if (c->h_counters) cudaFreeHost(c->h_counters);
    for (cudaEvent_t e : c->event_pool) cudaEventDestroy(e);
    if (c->ev_begin) cudaEventDestroy(c->ev_begin);
    if (c->ev_end) cudaEventDestroy(c->ev_end);
    for (cudaEvent_t e : {c->ev_in, c->ev_lvl0, c->ev_side, c->ev_out}) if (e) cudaEventDestroy(e);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->chain) cudaStreamDestroy(c->chain);
    delete c;
}

int wrt_upload_scene(WrtContext* c, const WrtSceneDesc* s) {
    if (!c || !s) return fail("wrt_upload_scene: null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaDeviceSynchronize());                       // no frame may still read the old scene
    c->has_scene = false;
    c->scene_slot = 0;
    wrt::DevScene& ds = c->ds;
    memset(&ds, 0, sizeof ds);
    c->textures_complete = true;
    const int np = s->n_prims;
    // nodes: identical 32-byte records, viewed as float4 pairs on the device
    static_assert(sizeof(WrtNode) == 32, "WrtNode must be 32 bytes");
    if (dev_upload(c, (const float4*)s->nodes, 2 * (size_t)s->n_nodes, &ds.nodes)) return 1;
    // SAH tree over the same leaf boxes (fast_bvh.hpp explains why results are identical)
    wrt::FastBvhBuilder& fbvh = c->fbvh;               // host-side scratch lives in the context: uploads of same-sized
    auto t_sah0 = std::chrono::steady_clock::now();    // scenes (every frame of the end-to-end path) reuse its pages
    fbvh.build(s);
    auto t_sah1 = std::chrono::steady_clock::now();
    if (fbvh.nodes.size() != (size_t)s->n_nodes) return fail("wrt_upload_scene: fast BVH build failed");
    if (dev_upload(c, (const float4*)fbvh.nodes.data(), 2 * fbvh.nodes.size(), &ds.fnodes)) return 1;
    {
        std::vector<WrtNode>& oct = c->h_oct;
        wrt::FastBvhBuilder::octant_copies_into(fbvh.nodes.data(), fbvh.nodes.size(), oct);
        if (dev_upload(c, (const float4*)oct.data(), 2 * oct.size(), &ds.onodes)) return 1;
        wrt::FastBvhBuilder::octant_copies_into(s->nodes, (size_t)s->n_nodes, oct);      // (the previous copy was staged)
        if (dev_upload(c, (const float4*)oct.data(), 2 * oct.size(), &ds.ronodes)) return 1;
    }
    // dilated copy for the directional-shadow loop, which the reference runs without any box test
    std::vector<WrtNode> dil = fbvh.dilated(1e-3f, 1e-4f);
    if (dev_upload(c, (const float4*)dil.data(), 2 * dil.size(), &ds.dnodes)) return 1;
    std::vector<float4> geom(3 * (size_t)np), attr(4 * (size_t)np);
    std::vector<int4> ids(np);
    int has_light = 0;
    for (int p = 0; p < np; p++) {
        const float* g = s->prim_geom + 12 * (size_t)p;
        const WrtMaterial& m = s->materials[s->prim_material[p]];
        unsigned flags = s->prim_flags[p];
        if (flags & WRT_PRIM_LIGHT) has_light = 1;
        float oma = 1 - m.alpha;                       // (1 - inter.mtlcolor.alpha), BVHStrategy.hpp:38
        float fl;
        memcpy(&fl, &flags, 4);
        geom[3 * (size_t)p + 0] = make_float4(g[0], g[1], g[2], g[3]);
        geom[3 * (size_t)p + 1] = make_float4(g[4], g[5], g[6], oma);
        geom[3 * (size_t)p + 2] = make_float4(g[8], g[9], g[10], fl);
        const float* nn = s->prim_normals + 9 * (size_t)p;
        const float* uv = s->prim_uv + 6 * (size_t)p;
        attr[4 * (size_t)p + 0] = make_float4(nn[0], nn[1], nn[2], uv[0]);
        attr[4 * (size_t)p + 1] = make_float4(nn[3], nn[4], nn[5], uv[1]);
        attr[4 * (size_t)p + 2] = make_float4(nn[6], nn[7], nn[8], uv[2]);
        attr[4 * (size_t)p + 3] = make_float4(uv[3], uv[4], uv[5], 0.f);
        ids[p] = make_int4(s->prim_material[p], s->prim_texture[p], s->prim_normalmap[p], s->prim_object[p]);
        // the strategy queries never read texels; a frame render needs every referenced map present
        if (s->prim_texture[p] >= s->n_textures || s->prim_normalmap[p] >= s->n_normalmaps) c->textures_complete = false;
    }
    std::vector<float4> mats(3 * (size_t)s->n_materials);
    for (int i = 0; i < s->n_materials; i++) {
        const WrtMaterial& m = s->materials[i];
        mats[3 * (size_t)i + 0] = make_float4(m.diffuse[0], m.diffuse[1], m.diffuse[2], m.ka);
        mats[3 * (size_t)i + 1] = make_float4(m.specular[0], m.specular[1], m.specular[2], m.kd);
        mats[3 * (size_t)i + 2] = make_float4(m.ks, m.n, m.alpha, m.eta);
    }
    std::vector<float4> pbox(2 * (size_t)np, make_float4(0.f, 0.f, 0.f, 0.f));
    for (int i = 0; i < s->n_nodes; i++) {
        const WrtNode& nd = s->nodes[i];
        if (nd.link >= 0 || i == 1) continue;
        int p = ~nd.link;
        if (p < 0 || p >= np) continue;
        pbox[2 * (size_t)p] = make_float4(nd.pmin[0], nd.pmin[1], nd.pmin[2], 0.f);
        pbox[2 * (size_t)p + 1] = make_float4(nd.pmax[0], nd.pmax[1], nd.pmax[2], 0.f);
    }
    if (dev_upload(c, pbox.data(), pbox.size(), &ds.prim_box)) return 1;
    if (dev_upload(c, geom.data(), geom.size(), &ds.geom)) return 1;
    if (dev_upload(c, attr.data(), attr.size(), &ds.attr)) return 1;
    if (dev_upload(c, ids.data(), ids.size(), &ds.ids)) return 1;
    if (dev_upload(c, s->object_prim, (size_t)np, &ds.object_prim)) return 1;
    if (dev_upload(c, mats.data(), mats.size(), &ds.materials)) return 1;
    if (dev_upload(c, s->lights, (size_t)s->n_lights, &ds.lights)) return 1;
    if (dev_upload(c, s->textures, (size_t)s->n_textures, &ds.textures)) return 1;
    if (dev_upload(c, s->normalmaps, (size_t)s->n_normalmaps, &ds.normalmaps)) return 1;
    if (dev_upload(c, s->texels, 3 * (size_t)s->n_texels, &ds.texels)) return 1;
    for (int i = 0; i < s->n_lights && i < WRT_INLINE_LIGHTS; i++) ds.lights_c[i] = s->lights[i];
    ds.n_nodes = s->n_nodes; ds.n_prims = np; ds.n_lights = s->n_lights;
    ds.n_point_lights = ds.n_dir_lights = 0;
    for (int i = 0; i < s->n_lights; i++) {
        if (fabsf(s->lights[i].pos[3] - 1.f) < 0.00001f) ++ds.n_point_lights;      // FLOAT_EQUAL(w, 1), Renderer.hpp:275
        else ++ds.n_dir_lights;
    }
    ds.has_light_prims = has_light;
    ds.shadow_type = s->shadow_type; ds.depth_cueing = s->depth_cueing;
    for (int k = 0; k < 3; k++) { ds.bkg[k] = s->bkgcolor[k]; ds.dc[k] = s->dc[k]; ds.eye[k] = s->eye[k]; }
    ds.eta = s->eta;
    ds.amin = s->amin; ds.amax = s->amax; ds.distmin = s->distmin; ds.distmax = s->distmax;
    c->bvh_depth = tree_depth(s);
    c->stack_rows = std::max(c->bvh_depth, fbvh.max_depth) + 2;
#ifdef WRT_DEBUG_BOUNDS
    if (const char* e = getenv("WRT_DEBUG_STACK_ROWS")) c->stack_rows = std::max(1, atoi(e));   // provoke the stack check (tests)
#endif
    if (c->stack_rows > 90)     // 90 rows x 128 threads x 4 B = 45 KB of dynamic shared memory per CTA
        return fail("wrt_upload_scene: acceleration tree deeper than 88 levels (degenerate geometry?)");
    CK(cudaStreamSynchronize(c->chain));          // host staging vectors go out of scope here
    c->has_scene = true;
    if (getenv("WRT_VERBOSE"))
        fprintf(stderr, "[wrt] upload: %d prims, reference tree depth %d, SAH tree depth %d (built in %.2f ms)\n", np,
                c->bvh_depth, fbvh.max_depth, std::chrono::duration<double, std::milli>(t_sah1 - t_sah0).count());
    return 0;
}

int wrt_set_camera(WrtContext* c, const WrtCamera* cam) {
    if (!c || !cam) return fail("wrt_set_camera: null argument");
    if (cam->width < 0 || cam->height < 0) return fail("wrt_set_camera: negative image size");
    c->cam = *cam;
    c->has_cam = true;
    return 0;
}

int wrt_set_tiles(WrtContext* c, int tile_w, int tile_h, int rank, int world) {
    if (!c) return fail("wrt_set_tiles: null context");
    if (tile_w <= 0 || tile_h <= 0 || tile_w % 8 || tile_h % 4) return fail("wrt_set_tiles: tile must be a multiple of 8 x 4");
    if (world < 1 || rank < 0 || rank >= world) return fail("wrt_set_tiles: bad rank/world");
    c->tile_w = tile_w; c->tile_h = tile_h; c->rank = rank; c->world = world;
    return 0;
}

int wrt_set_options(WrtContext* c, int traversal, uint32_t seed, float queue_factor) {
    if (!c) return fail("wrt_set_options: null context");
    if (traversal != WRT_TRAVERSAL_EXHAUSTIVE && traversal != WRT_TRAVERSAL_PRUNED) return fail("wrt_set_options: bad traversal mode");
    c->traversal = traversal;
    c->seed = seed;
    if (queue_factor > 0 && queue_factor != c->queue_factor) {
        c->queue_factor = queue_factor;                // caller-fixed deep-level capacity: no automatic growth
        cudaSetDevice(c->device);
        cudaDeviceSynchronize();
        free_frame(c);
    }
    return 0;
}

int wrt_enable_kernel_timing(WrtContext* c, int on) {
    if (!c) return fail("wrt_enable_kernel_timing: null context");
    c->kernel_timing = on != 0;
    return 0;
}

// ---------------- batch queries ----------------

static int batch_common(WrtContext* c, int64_t n) {
    if (!c) return fail("null context");
    if (!c->has_scene) return fail("no scene uploaded");
    if (n < 0) return fail("negative count");
    CK(cudaSetDevice(c->device));
    return 0;
}

int wrt_trace_closest(WrtContext* c, const float* orig, const float* dir, int64_t n, WrtHit* hits) {
    if (batch_common(c, n)) return 1;
    if (n == 0) return 0;
    cudaStream_t st = c->chain;
    size_t vb = (size_t)n * 3 * sizeof(float);
    if (ensure_scratch(c, 0, vb) || ensure_scratch(c, 1, vb) || ensure_scratch(c, 2, (size_t)n * sizeof(WrtHit))) return 1;
    CK(cudaMemcpyAsync(c->d_scratch[0], orig, vb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->d_scratch[1], dir, vb, cudaMemcpyHostToDevice, st));
    int grid = (int)std::min<int64_t>((n + 127) / 128, grid_for(c, 8));
    ++c->launches;
    wrt::k_batch_closest<<<grid, 128, stack_bytes(c, 128), st>>>(c->ds, (const float*)c->d_scratch[0],
                                                                  (const float*)c->d_scratch[1], n,
                                                                  (WrtHit*)c->d_scratch[2], prune_value(c));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(hits, c->d_scratch[2], (size_t)n * sizeof(WrtHit), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

static int batch_shadow(WrtContext* c, const float* pos, const float* ndir, const float* lightpos, int64_t n,
                        float* coeff, int mode) {
    if (batch_common(c, n)) return 1;
    if (n == 0) return 0;
    cudaStream_t st = c->chain;
    size_t vb = (size_t)n * 3 * sizeof(float);
    for (int k = 0; k < 3; k++) if (ensure_scratch(c, k, std::max(vb, c->scratch_bytes[k]))) return 1;
    if (ensure_scratch(c, 3, (size_t)n * sizeof(float))) return 1;
    CK(cudaMemcpyAsync(c->d_scratch[0], pos, vb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->d_scratch[1], ndir, vb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->d_scratch[2], lightpos, vb, cudaMemcpyHostToDevice, st));
    int grid = (int)std::min<int64_t>((n + 127) / 128, grid_for(c, 8));
    ++c->launches;
    wrt::k_batch_shadow<<<grid, 128, stack_bytes(c, 128), st>>>(c->ds, (const float*)c->d_scratch[0],
                                                                 (const float*)c->d_scratch[1],
                                                                 (const float*)c->d_scratch[2], n,
                                                                 (float*)c->d_scratch[3], mode, prune_value(c));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(coeff, c->d_scratch[3], (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

int wrt_shadow_hard(WrtContext* c, const float* pos, const float* ndir, const float* lightpos, int64_t n, float* coeff) {
    return batch_shadow(c, pos, ndir, lightpos, n, coeff, 0);
}

int wrt_shadow_soft(WrtContext* c, const float* pos, const float* ndir, const float* lightpos, int64_t n, float* coeff) {
    return batch_shadow(c, pos, ndir, lightpos, n, coeff, 1);
}

int wrt_shadow_directional(WrtContext* c, const float* pos, const int32_t* self_object, const float* lightdir4,
                           int64_t n, float* coeff) {
    if (batch_common(c, n)) return 1;
    if (n == 0) return 0;
    cudaStream_t st = c->chain;
    if (ensure_scratch(c, 0, (size_t)n * 12) || ensure_scratch(c, 1, (size_t)n * 4) ||
        ensure_scratch(c, 2, (size_t)n * 16) || ensure_scratch(c, 3, (size_t)n * 4))
        return 1;
    CK(cudaMemcpyAsync(c->d_scratch[0], pos, (size_t)n * 12, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->d_scratch[1], self_object, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->d_scratch[2], lightdir4, (size_t)n * 16, cudaMemcpyHostToDevice, st));
    int grid = (int)std::min<int64_t>((n + 127) / 128, grid_for(c, 8));
    ++c->launches;
    wrt::k_batch_shadow_directional<<<grid, 128, stack_bytes(c, 128), st>>>(
        c->ds, (const float*)c->d_scratch[0], (const int*)c->d_scratch[1], (const float*)c->d_scratch[2], n,
        (float*)c->d_scratch[3], c->traversal == WRT_TRAVERSAL_EXHAUSTIVE ? 1 : 0);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(coeff, c->d_scratch[3], (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

// ---------------- frame ----------------

int wrt_render(WrtContext* c, uint8_t* rgb_host, WrtStats* stats) {
    if (!c || !rgb_host) return fail("wrt_render: null argument");
    CK(cudaSetDevice(c->device));
    if (!c->has_cam) return fail("wrt_render: no camera set");
    size_t bytes = (size_t)c->cam.width * c->cam.height * 3;
    if (bytes == 0) return fail("wrt_render: empty image");
    if (c->d_image_bytes < bytes) {
        if (c->d_image) cudaFree(c->d_image);
        c->d_image = nullptr; c->d_image_bytes = 0;
        CK(cudaMalloc((void**)&c->d_image, bytes));
        c->d_image_bytes = bytes;
    }
    if (c->h_image_bytes < bytes) {
        if (c->h_image) cudaFreeHost(c->h_image);
        c->h_image = nullptr; c->h_image_bytes = 0;
        CK(cudaMallocHost((void**)&c->h_image, bytes));
        c->h_image_bytes = bytes;
    }
    cudaStream_t st = c->chain;
    if (c->world > 1) CK(cudaMemsetAsync(c->d_image, 0, bytes, st));
    if (render_all(c, nullptr, false, c->d_image, nullptr)) return 1;
    if (finish_frame(c)) return 1;
    cudaPointerAttributes attr;
    bool user_pinned = c->world == 1 && cudaPointerGetAttributes(&attr, rgb_host) == cudaSuccess &&
                       attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (user_pinned) {                                 // caller's buffer is page-locked: no staging copy
        CK(cudaMemcpyAsync(rgb_host, c->d_image, bytes, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (stats) *stats = c->stats;
        return 0;
    }
    CK(cudaMemcpyAsync(c->h_image, c->d_image, bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (c->world == 1) memcpy(rgb_host, c->h_image, bytes);
    else {
        // only this rank's pixels are defined; copy them, leave the rest of the caller's buffer alone
        wrt::TileMap tm = c->tilemap();
        long long slots = c->local_slots(c->rank, c->world);
        for (long long s = 0; s < slots; s++) {
            int x, y;
            if (tm.slot_to_pixel(s, c->rank, x, y)) {
                size_t o = 3 * ((size_t)y * tm.width + x);
                rgb_host[o] = c->h_image[o]; rgb_host[o + 1] = c->h_image[o + 1]; rgb_host[o + 2] = c->h_image[o + 2];
            }
        }
    }
    if (stats) *stats = c->stats;
    return 0;
}

int wrt_render_device(WrtContext* c, void* d_rgb_tiles, void* cuda_stream) {
    if (!c || !d_rgb_tiles) return fail("wrt_render_device: null argument");
    CK(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;        // NULL = the legacy default stream, as in CUDA
    return render_all(c, st, true, nullptr, (uint8_t*)d_rgb_tiles);
}

int wrt_finish_device(WrtContext* c, WrtStats* stats) {
    if (!c) return fail("wrt_finish_device: null context");
    CK(cudaSetDevice(c->device));
    if (!c->has_frame) return fail("wrt_finish_device: no frame in flight");
    // A frame that overflowed a queue is re-rendered here (synchronously, into the same device buffer): when
    // stats->overflow_retries != 0 the caller must repeat whatever it had already enqueued behind the frame on
    // its own stream (parallel.py: the gather).
    if (finish_frame(c)) return 1;
    if (stats) *stats = c->stats;
    return 0;
}

int wrt_get_stats(WrtContext* c, WrtStats* stats) {
    if (!c || !stats) return fail("wrt_get_stats: null argument");
    *stats = c->stats;
    return 0;
}

int64_t wrt_tile_pixel_count(WrtContext* c, int rank, int world) {
    if (!c || !c->has_cam || world < 1 || rank < 0 || rank >= world) return -1;
    return c->local_slots(rank, world);
}

int wrt_scatter_tiles(WrtContext* c, const void* d_gathered, int world, int64_t stride_bytes, void* d_rgb_image,
                      void* cuda_stream) {
    if (!c || !d_gathered || !d_rgb_image) return fail("wrt_scatter_tiles: null argument");
    if (!c->has_cam) return fail("wrt_scatter_tiles: no camera set");
    CK(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;        // NULL = the legacy default stream, as in CUDA
    wrt::TileMap tm = c->tilemap();
    tm.world = world;
    long long slots_per_rank = c->local_slots(0, world);   // rank 0 owns the most tiles
    if (stride_bytes < slots_per_rank * 3) return fail("wrt_scatter_tiles: stride smaller than rank 0's tile buffer");
    ++c->launches;
    wrt::k_scatter_tiles<<<grid_for(c, 8), 256, 0, st>>>((const unsigned char*)d_gathered, stride_bytes, tm, world,
                                                         slots_per_rank, (unsigned char*)d_rgb_image);
    CK(cudaGetLastError());
    return 0;
}

int64_t wrt_kernel_launch_count(WrtContext* c) { return c ? c->launches : 0; }

int wrt_measure_fp32_peak(WrtContext* c, float* tflops_fma, float* tflops_mul_add) {
    if (!c || !tflops_fma || !tflops_mul_add) return fail("wrt_measure_fp32_peak: null argument");
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->chain;
    if (ensure_scratch(c, 4, 256)) return 1;
    const int iters = 1 << 15, blocks = c->num_sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best[2] = {0.f, 0.f};
    for (int mode = 0; mode < 2; mode++) {
        for (int rep = 0; rep < 4; rep++) {
            CK(cudaEventRecord(e0, st));
            wrt::k_fp32_peak<<<blocks, threads, 0, st>>>((float*)c->d_scratch[4], iters, mode);
            CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            double flops = (double)blocks * threads * (double)iters * 8.0 * 2.0;
            float tf = (float)(flops / (ms * 1e-3) / 1e12);
            if (rep > 0 && tf > best[mode]) best[mode] = tf;
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops_fma = best[0];
    *tflops_mul_add = best[1];
    return 0;
}

int wrt_get_kernel_times(WrtContext* c, float* ms, int capacity) {
    if (!c || !ms) return 0;
    int n = std::min(capacity, (int)WRT_KERNEL_FAMILIES);
    for (int i = 0; i < n; i++) ms[i] = c->family_ms[i];
    return n;
}

int wrt_get_kernel_launches(WrtContext* c, int32_t* launches, int capacity) {
    if (!c || !launches) return 0;
    int n = std::min(capacity, (int)WRT_KERNEL_FAMILIES);
    for (int i = 0; i < n; i++) launches[i] = c->family_launches[i];
    return n;
}

} // extern "C"
