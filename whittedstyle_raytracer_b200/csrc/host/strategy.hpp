// strategy.hpp — the reference's intersection-strategy plugin point
// (include/IIntersectStrategy.h:7-15) on the host types of this build, and the
// CUDA strategy that sits beside the reference's BaseInterStrategy / BVHStrategy.
//
// The two virtuals keep the reference's names, argument meaning and out-param
// convention (`inter` is overwritten wholesale, like BVHStrategy::UpdateInter).
// The single-ray forms forward to the batch C ABI with n = 1, so a CudaStrategy
// is a legal drop-in wherever the reference's Renderer holds an
// IIntersectStrategy*; the batch forms are what a wavefront caller should use.
#pragma once
#include <stdexcept>
#include <vector>

#include "../../../include/wrt_cuda.h"
#include "host_scene.hpp"

namespace wrt {

struct Intersection {                    // Intersection.hpp:13-28
    bool intersected = false;
    float t = FLT_MAX;
    V3 pos, nDir;
    V2 textPos;
    int textureIndex = -1, normalMapIndex = -1;
    Material mtlcolor;
    const Object* obj = nullptr;
};

class IIntersectStrategy {
public:
    virtual ~IIntersectStrategy() = default;
    virtual void UpdateInter(Intersection& inter, HostScene& sce, const V3& rayOrig, const V3& rayDir) = 0;
    virtual float getShadowCoeffi(HostScene& sce, Intersection& p, V3& lightpos) = 0;
};

class CudaStrategy : public IIntersectStrategy {
public:
    // The scene must have been flattened (HostScene::buildAndFlatten) before construction.
    CudaStrategy(HostScene& sce, int device = 0) {
        if (wrt_create(device, &ctx_) != 0) throw std::runtime_error(wrt_last_error());
        if (wrt_upload_scene(ctx_, &sce.desc) != 0 || wrt_set_camera(ctx_, &sce.cam) != 0) {
            std::string e = wrt_last_error();
            wrt_destroy(ctx_);
            throw std::runtime_error(e);
        }
    }
    ~CudaStrategy() override { wrt_destroy(ctx_); }
    CudaStrategy(const CudaStrategy&) = delete;
    CudaStrategy& operator=(const CudaStrategy&) = delete;

    WrtContext* context() const { return ctx_; }

    void UpdateInter(Intersection& inter, HostScene& sce, const V3& rayOrig, const V3& rayDir) override {
        WrtHit h;
        if (wrt_trace_closest(ctx_, &rayOrig.x, &rayDir.x, 1, &h) != 0) throw std::runtime_error(wrt_last_error());
        inter = fromHit(sce, h);
    }

    float getShadowCoeffi(HostScene&, Intersection& p, V3& lightpos) override {
        float c = 1.f;
        if (wrt_shadow_hard(ctx_, &p.pos.x, &p.nDir.x, &lightpos.x, 1, &c) != 0) throw std::runtime_error(wrt_last_error());
        return c;
    }

    // Batch forms (V3 is three packed floats).
    void UpdateInter(std::vector<Intersection>& inter, HostScene& sce, const std::vector<V3>& orig, const std::vector<V3>& dir) {
        std::vector<WrtHit> h(orig.size());
        if (wrt_trace_closest(ctx_, &orig[0].x, &dir[0].x, (int64_t)orig.size(), h.data()) != 0)
            throw std::runtime_error(wrt_last_error());
        inter.resize(orig.size());
        for (size_t i = 0; i < h.size(); i++) inter[i] = fromHit(sce, h[i]);
    }

    static Intersection fromHit(const HostScene& sce, const WrtHit& h) {
        Intersection in;
        if (!h.hit) return in;
        in.intersected = true;
        in.t = h.t;
        in.pos = V3(h.pos[0], h.pos[1], h.pos[2]);
        in.nDir = V3(h.ndir[0], h.ndir[1], h.ndir[2]);
        in.textPos = V2(h.uv[0], h.uv[1]);
        in.textureIndex = h.texture;
        in.normalMapIndex = h.normalmap;
        in.obj = &sce.objList[h.object];
        in.mtlcolor = in.obj->mtl;
        return in;
    }

private:
    WrtContext* ctx_ = nullptr;
};

static_assert(sizeof(V3) == 3 * sizeof(float), "V3 must be three packed floats");

} // namespace wrt
