// CudaStrategy.hpp — the file a maintainer of bobhansky/WhittedStyle_Raytracer adds to
// include/ to run intersection queries on a B200 through libwrt_cuda.so.
//
// It implements the reference's own plugin interface (include/IIntersectStrategy.h:7-15)
// on the reference's own types, beside BaseInterStrategy and BVHStrategy:
//     virtual void  UpdateInter(Intersection&, Scene&, const Vector3f&, const Vector3f&);
//     virtual float getShadowCoeffi(Scene&, Intersection&, Vector3f& lightpos);
// The scene is flattened once, lazily, from the Scene / BVHAccel the Renderer has already
// built (include/Renderer.hpp:38-49): the pointer tree becomes the 32-byte record array of
// include/wrt_scene.h, primitives are numbered in depth-first leaf order.
// Single-ray calls forward to the batch C ABI with n = 1 (correct, but latency-bound: the
// per-frame entry point wrt_render is what a renderer should call).
//
// Build: add -I<this repo>/include -L<this repo>/whittedstyle_raytracer_b200 -lwrt_cuda.
#pragma once

#include <cstring>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "IIntersectStrategy.h"
#include "BVH.hpp"
#include "Sphere.hpp"
#include "Triangle.hpp"
#include "Vector.hpp"

#include "wrt_cuda.h"

class CudaStrategy : public IIntersectStrategy {
public:
    explicit CudaStrategy(int device = 0) : device_(device) {}
    ~CudaStrategy() { if (ctx_) wrt_destroy(ctx_); }

    void UpdateInter(Intersection& inter, Scene& sce, const Vector3f& rayOrig, const Vector3f& rayDir) override {
        ensure(sce);
        float o[3] = { rayOrig.x, rayOrig.y, rayOrig.z }, d[3] = { rayDir.x, rayDir.y, rayDir.z };
        WrtHit h;
        if (wrt_trace_closest(ctx_, o, d, 1, &h) != 0) throw std::runtime_error(wrt_last_error());
        inter = Intersection();
        if (!h.hit) return;
        inter.intersected = true;
        inter.t = h.t;
        inter.pos = Vector3f(h.pos[0], h.pos[1], h.pos[2]);
        inter.nDir = Vector3f(h.ndir[0], h.ndir[1], h.ndir[2]);
        inter.textPos = Vector2f(h.uv[0], h.uv[1]);
        inter.textureIndex = h.texture;
        inter.normalMapIndex = h.normalmap;
        inter.obj = sce.objList[h.object].get();
        inter.mtlcolor = inter.obj->mtlcolor;
    }

    float getShadowCoeffi(Scene& sce, Intersection& p, Vector3f& lightpos) override {
        ensure(sce);
        float pos[3] = { p.pos.x, p.pos.y, p.pos.z }, nd[3] = { p.nDir.x, p.nDir.y, p.nDir.z };
        float lp[3] = { lightpos.x, lightpos.y, lightpos.z };
        float c = 1.f;
        if (wrt_shadow_hard(ctx_, pos, nd, lp, 1, &c) != 0) throw std::runtime_error(wrt_last_error());
        return c;
    }

private:
    int device_;
    WrtContext* ctx_ = nullptr;

    std::vector<WrtNode> nodes_;
    std::vector<Object*> leaf_objs_;

    void flattenNode(BVHNode* n, int rec) {
        nodes_[rec].pmin[0] = n->bound.pMin.x; nodes_[rec].pmin[1] = n->bound.pMin.y; nodes_[rec].pmin[2] = n->bound.pMin.z;
        nodes_[rec].pmax[0] = n->bound.pMax.x; nodes_[rec].pmax[1] = n->bound.pMax.y; nodes_[rec].pmax[2] = n->bound.pMax.z;
        if (!n->left && !n->right) {
            nodes_[rec].link = ~(int)leaf_objs_.size();
            leaf_objs_.push_back(n->obj);
            return;
        }
        int pair = (int)nodes_.size();
        nodes_.resize(pair + 2);
        memset(&nodes_[pair], 0, 2 * sizeof(WrtNode));
        nodes_[rec].link = pair;
        flattenNode(n->left, pair);
        flattenNode(n->right, pair + 1);
    }

    void ensure(Scene& sce) {
        if (ctx_) return;
        if (wrt_create(device_, &ctx_) != 0) throw std::runtime_error(wrt_last_error());
        const int n = (int)sce.objList.size();
        nodes_.clear(); leaf_objs_.clear();
        if (n > 0) {
            nodes_.resize(2);
            memset(nodes_.data(), 0, 2 * sizeof(WrtNode));
            nodes_[1].link = ~0;
            flattenNode(sce.BVHaccelerator->getNode(), 0);
        }
        std::unordered_map<const Object*, int> objIndex;
        for (int i = 0; i < n; i++) objIndex[sce.objList[i].get()] = i;

        std::vector<float> geom((size_t)n * 12, 0.f), normals((size_t)n * 9, 0.f), uv((size_t)n * 6, 0.f);
        std::vector<uint32_t> flags(n, 0);
        std::vector<int32_t> mat(n, 0), tex(n, -1), nmap(n, -1), prim_object(n, -1), object_prim(n, -1);
        std::vector<WrtMaterial> materials;
        for (int p = 0; p < n; p++) {
            Object* o = leaf_objs_[p];
            prim_object[p] = objIndex.at(o);
            object_prim[prim_object[p]] = p;
            WrtMaterial m;
            m.diffuse[0] = o->mtlcolor.diffuse.x; m.diffuse[1] = o->mtlcolor.diffuse.y; m.diffuse[2] = o->mtlcolor.diffuse.z;
            m.specular[0] = o->mtlcolor.specular.x; m.specular[1] = o->mtlcolor.specular.y; m.specular[2] = o->mtlcolor.specular.z;
            m.ka = o->mtlcolor.ka; m.kd = o->mtlcolor.kd; m.ks = o->mtlcolor.ks; m.n = o->mtlcolor.n;
            m.alpha = o->mtlcolor.alpha; m.eta = o->mtlcolor.eta;
            mat[p] = (int)materials.size();
            materials.push_back(m);
            tex[p] = o->textureIndex; nmap[p] = o->normalMapIndex;
            flags[p] = (o->objectType == SPEHRE ? WRT_PRIM_SPHERE : WRT_PRIM_TRIANGLE) | (o->isLight ? WRT_PRIM_LIGHT : 0u) |
                       (o->isTextureActivated ? WRT_PRIM_TEXTURED : 0u);
            float* g = &geom[(size_t)p * 12];
            if (o->objectType == TRIANGLE) {
                Triangle* t = static_cast<Triangle*>(o);
                Vector3f e1 = t->v1 - t->v0, e2 = t->v2 - t->v0;
                g[0] = t->v0.x; g[1] = t->v0.y; g[2] = t->v0.z;
                g[4] = e1.x; g[5] = e1.y; g[6] = e1.z;
                g[8] = e2.x; g[9] = e2.y; g[10] = e2.z;
                const Vector3f* ns[3] = { &t->n0, &t->n1, &t->n2 };
                const Vector2f* ts[3] = { &t->uv0, &t->uv1, &t->uv2 };
                for (int k = 0; k < 3; k++) {
                    normals[(size_t)p * 9 + 3 * k] = ns[k]->x; normals[(size_t)p * 9 + 3 * k + 1] = ns[k]->y;
                    normals[(size_t)p * 9 + 3 * k + 2] = ns[k]->z;
                    uv[(size_t)p * 6 + 2 * k] = ts[k]->x; uv[(size_t)p * 6 + 2 * k + 1] = ts[k]->y;
                }
            } else {
                Sphere* s = static_cast<Sphere*>(o);
                g[0] = s->centerPos.x; g[1] = s->centerPos.y; g[2] = s->centerPos.z; g[3] = s->radius;
            }
        }
        std::vector<WrtLight> lights;
        for (auto& l : sce.lightList) {
            WrtLight w;
            memset(&w, 0, sizeof w);
            w.pos[0] = l->pos.x; w.pos[1] = l->pos.y; w.pos[2] = l->pos.z; w.pos[3] = l->pos.w;
            w.color[0] = l->color.x; w.color[1] = l->color.y; w.color[2] = l->color.z;
            w.c1 = l->c1; w.c2 = l->c2; w.c3 = l->c3;
            const Vector3f* tv[3] = { &l->triangle.v0, &l->triangle.v1, &l->triangle.v2 };
            for (int k = 0; k < 3; k++) { w.tri[3 * k] = tv[k]->x; w.tri[3 * k + 1] = tv[k]->y; w.tri[3 * k + 2] = tv[k]->z; }
            lights.push_back(w);
        }
        WrtSceneDesc d;
        memset(&d, 0, sizeof d);
        d.n_nodes = (int32_t)nodes_.size(); d.n_prims = n;
        d.n_materials = (int32_t)materials.size(); d.n_lights = (int32_t)lights.size();
        d.nodes = nodes_.data(); d.prim_geom = geom.data(); d.prim_flags = flags.data();
        d.prim_material = mat.data(); d.prim_texture = tex.data(); d.prim_normalmap = nmap.data();
        d.prim_object = prim_object.data(); d.object_prim = object_prim.data();
        d.prim_normals = normals.data(); d.prim_uv = uv.data();
        d.materials = materials.data(); d.lights = lights.data();
        // textures and frame globals are only needed by wrt_render; the two strategy queries do not read them
        if (wrt_upload_scene(ctx_, &d) != 0) throw std::runtime_error(wrt_last_error());
    }
};
