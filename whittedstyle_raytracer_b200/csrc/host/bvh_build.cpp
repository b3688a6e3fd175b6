// bvh_build.cpp — reference-identical BVH build, flattening and camera setup.
//
// The tree topology must equal the reference's pointer tree
// (include/BVH.hpp:49-125) because closest-hit ties resolve to the left
// subtree (BVH.hpp:157) and the hard-shadow product visits every leaf whose
// ancestor boxes are hit (BVHStrategy.hpp:24-48).  So the build keeps:
//   * one object per leaf, the size-2 special case without sorting (BVH.hpp:55-69);
//   * union of ALL boxes -> maxExtent axis (ties: y, then z; BoundBox.hpp:41-50);
//   * libstdc++ std::sort on the box centroid `0.5*pMin + 0.5*pMax` with the
//     same strict-weak comparator (BVH.hpp:87-106) — the reference sorts a copy
//     of the sub-list, we sort the same sequence in place, which gives the same
//     permutation because introsort only sees comparator outcomes;
//   * split at size/2 (BVH.hpp:109-114).
// What is new is the layout: 32-byte records, sibling pairs adjacent and
// 64-byte aligned, primitives renumbered in depth-first leaf order.
#include <algorithm>
#include <cstring>
#include <thread>

#include "host_scene.hpp"

namespace wrt {

namespace {

struct Builder {
    HostScene& s;
    std::vector<int> order;      // object indices, permuted in place
    std::vector<float> cen;      // box centroids, 3 per object (same floats the reference recomputes in its comparator)
    std::vector<float> box;      // pMin, pMax per object, 6 floats: keeps the union loops out of the 200-byte Object records
    explicit Builder(HostScene& hs) : s(hs) {}

    static float centroid(const Object& o, int axis) {            // BoundBox.hpp:33
        const float* mn = &o.bmin.x;
        const float* mx = &o.bmax.x;
        return mn[axis] * 0.5f + mx[axis] * 0.5f;
    }

    // Builds the subtree over order[b,e) into record `rec`; its descendants take records [free, free + 2(e-b) - 2):
    // a subtree over k objects has exactly 2k-2 records below its root, so the layout the sequential depth-first
    // build produces (pair of a node = next free record when the node is visited) is known in advance, the leaf
    // of order[p] has depth-first rank p, and sibling subtrees can be built by different threads into disjoint
    // ranges of the preallocated array — same records, same order[] as one thread would produce.
    // Returns the depth of the deepest leaf.
    int build(int rec, int b, int e, int free, int depth) {
        int n = e - b;
        if (n == 1) {
            const Object& o = s.objList[order[b]];
            WrtNode& nd = s.nodes[rec];
            set_box(nd, o.bmin, o.bmax);
            nd.link = ~b;
            return depth;
        }
        if (n > 2) {
            const float* b0 = &box[6 * (size_t)order[b]];
            V3 mn(b0[0], b0[1], b0[2]), mx(b0[3], b0[4], b0[5]);
            for (int i = b + 1; i < e; i++) unite(mn, mx, &box[6 * (size_t)order[i]]);
            V3 d = mx - mn;                                        // BoundBox.hpp:41-50
            int axis = (d.x > d.y && d.x > d.z) ? 0 : (d.y > d.z ? 1 : 2);
            const float* cp = cen.data() + axis;
            std::sort(order.begin() + b, order.begin() + e, [cp](int a, int c) { return cp[3 * a] < cp[3 * c]; });
        }
        const int mid = (n == 2) ? b + 1 : b + n / 2;
        const int pair = free;
        s.nodes[rec].link = pair;
        const int free_l = free + 2, free_r = free + 2 * (mid - b);
        int dl, dr;
        if (n >= 8192 && depth < 4) {                              // up to 16 concurrent subtrees
            std::thread left([&] { dl = build(pair, b, mid, free_l, depth + 1); });
            dr = build(pair + 1, mid, e, free_r, depth + 1);
            left.join();
        } else {
            dl = build(pair, b, mid, free_l, depth + 1);
            dr = build(pair + 1, mid, e, free_r, depth + 1);
        }
        const WrtNode& L = s.nodes[pair];
        const WrtNode& R = s.nodes[pair + 1];
        WrtNode& nd = s.nodes[rec];
        for (int k = 0; k < 3; k++) {                              // Union(), BoundBox.hpp:90-102
            float lo = fminf(L.pmin[k], R.pmin[k]), hi = fmaxf(L.pmax[k], R.pmax[k]);
            nd.pmin[k] = fminf(lo, hi);
            nd.pmax[k] = fmaxf(lo, hi);
        }
        return std::max(dl, dr);
    }

    static void set_box(WrtNode& nd, const V3& mn, const V3& mx) {
        nd.pmin[0] = mn.x; nd.pmin[1] = mn.y; nd.pmin[2] = mn.z;
        nd.pmax[0] = mx.x; nd.pmax[1] = mx.y; nd.pmax[2] = mx.z;
    }
    static void unite(V3& mn, V3& mx, const float* o) {
        V3 lo(fminf(mn.x, o[0]), fminf(mn.y, o[1]), fminf(mn.z, o[2]));
        V3 hi(fmaxf(mx.x, o[3]), fmaxf(mx.y, o[4]), fmaxf(mx.z, o[5]));
        mn = V3(fminf(lo.x, hi.x), fminf(lo.y, hi.y), fminf(lo.z, hi.z));
        mx = V3(fmaxf(lo.x, hi.x), fmaxf(lo.y, hi.y), fmaxf(lo.z, hi.z));
    }
};

int material_index(std::vector<WrtMaterial>& table, const Material& m) {
    WrtMaterial w;
    w.diffuse[0] = m.diffuse.x; w.diffuse[1] = m.diffuse.y; w.diffuse[2] = m.diffuse.z;
    w.specular[0] = m.specular.x; w.specular[1] = m.specular.y; w.specular[2] = m.specular.z;
    w.ka = m.ka; w.kd = m.kd; w.ks = m.ks; w.n = m.n; w.alpha = m.alpha; w.eta = m.eta;
    for (size_t i = 0; i < table.size(); i++)
        if (memcmp(&table[i], &w, sizeof w) == 0) return (int)i;
    table.push_back(w);
    return (int)table.size() - 1;
}

void put3(std::vector<float>& v, size_t at, const V3& a) { v[at] = a.x; v[at + 1] = a.y; v[at + 2] = a.z; }

} // namespace

void HostScene::buildAndFlatten() {
    const int n = (int)objList.size();
    nodes.clear();
    materials.clear();
    Builder b(*this);
    if (n > 0) {
        b.order.resize(n);
        b.cen.resize(3 * (size_t)n);
        b.box.resize(6 * (size_t)n);
        for (int i = 0; i < n; i++) {
            b.order[i] = i;
            for (int k = 0; k < 3; k++) b.cen[3 * (size_t)i + k] = Builder::centroid(objList[i], k);
            const V3 &mn = objList[i].bmin, &mx = objList[i].bmax;
            float* bx = &b.box[6 * (size_t)i];
            bx[0] = mn.x; bx[1] = mn.y; bx[2] = mn.z; bx[3] = mx.x; bx[4] = mx.y; bx[5] = mx.z;
        }
        nodes.assign(2 * (size_t)n, WrtNode{});                   // root, padding record, 2n-2 descendants
        nodes[1].link = ~0;       // padding record, never referenced
        bvh_depth = b.build(0, 0, n, 2, 0);
    } else {
        bvh_depth = 0;
    }

    prim_geom.assign((size_t)n * 12, 0.f);
    prim_normals.assign((size_t)n * 9, 0.f);
    prim_uv.assign((size_t)n * 6, 0.f);
    prim_flags.assign(n, 0);
    prim_material.assign(n, 0);
    prim_texture.assign(n, -1);
    prim_normalmap.assign(n, -1);
    prim_object.assign(n, -1);
    object_prim.assign(n, -1);
    for (int p = 0; p < n; p++) {
        int oi = b.order[p];               // the leaf of order[p] has depth-first rank p
        const Object& o = objList[oi];
        prim_object[p] = oi;
        object_prim[oi] = p;
        uint32_t f = o.type == SPHERE ? WRT_PRIM_SPHERE : WRT_PRIM_TRIANGLE;
        if (o.isLight) f |= WRT_PRIM_LIGHT;
        if (o.isTextureActivated) f |= WRT_PRIM_TEXTURED;
        prim_flags[p] = f;
        prim_material[p] = material_index(materials, o.mtl);
        prim_texture[p] = o.textureIndex;
        prim_normalmap[p] = o.normalMapIndex;
        size_t g = (size_t)p * 12;
        if (o.type == TRIANGLE) {
            put3(prim_geom, g, o.v0);
            put3(prim_geom, g + 4, o.v1 - o.v0);     // E1, Triangle.hpp:22
            put3(prim_geom, g + 8, o.v2 - o.v0);     // E2, Triangle.hpp:23
            put3(prim_normals, (size_t)p * 9, o.n0);
            put3(prim_normals, (size_t)p * 9 + 3, o.n1);
            put3(prim_normals, (size_t)p * 9 + 6, o.n2);
            float* uv = &prim_uv[(size_t)p * 6];
            uv[0] = o.uv0.x; uv[1] = o.uv0.y; uv[2] = o.uv1.x; uv[3] = o.uv1.y; uv[4] = o.uv2.x; uv[5] = o.uv2.y;
        } else {
            put3(prim_geom, g, o.center);
            prim_geom[g + 3] = o.radius;
        }
    }

    lights.clear();
    for (const Light& l : lightList) {
        WrtLight w;
        memset(&w, 0, sizeof w);
        for (int k = 0; k < 4; k++) w.pos[k] = l.pos[k];
        w.color[0] = l.color.x; w.color[1] = l.color.y; w.color[2] = l.color.z;
        w.c1 = l.c1; w.c2 = l.c2; w.c3 = l.c3;
        const V3* tv[3] = {&l.tv0, &l.tv1, &l.tv2};
        for (int k = 0; k < 3; k++) { w.tri[3 * k] = tv[k]->x; w.tri[3 * k + 1] = tv[k]->y; w.tri[3 * k + 2] = tv[k]->z; }
        lights.push_back(w);
    }

    texels.clear();
    tex_desc.clear();
    nmap_desc.clear();
    auto pack = [&](const std::vector<Texture>& src, std::vector<WrtTexture>& dst) {
        for (const Texture& t : src) {
            WrtTexture d;
            d.width = t.width; d.height = t.height;
            d.offset = (int64_t)(texels.size() / 3);
            d.count = (int64_t)t.rgb.size();
            for (const V3& c : t.rgb) { texels.push_back(c.x); texels.push_back(c.y); texels.push_back(c.z); }
            dst.push_back(d);
        }
    };
    pack(textures, tex_desc);
    pack(normalMaps, nmap_desc);

    memset(&desc, 0, sizeof desc);
    desc.n_nodes = (int32_t)nodes.size();
    desc.n_prims = n;
    desc.n_materials = (int32_t)materials.size();
    desc.n_lights = (int32_t)lights.size();
    desc.n_textures = (int32_t)tex_desc.size();
    desc.n_normalmaps = (int32_t)nmap_desc.size();
    desc.n_texels = (int64_t)(texels.size() / 3);
    desc.nodes = nodes.data();
    desc.prim_geom = prim_geom.data();
    desc.prim_flags = prim_flags.data();
    desc.prim_material = prim_material.data();
    desc.prim_texture = prim_texture.data();
    desc.prim_normalmap = prim_normalmap.data();
    desc.prim_object = prim_object.data();
    desc.object_prim = object_prim.data();
    desc.prim_normals = prim_normals.data();
    desc.prim_uv = prim_uv.data();
    desc.materials = materials.data();
    desc.lights = lights.data();
    desc.textures = tex_desc.data();
    desc.normalmaps = nmap_desc.data();
    desc.texels = texels.data();
    desc.bkgcolor[0] = bkgcolor.x; desc.bkgcolor[1] = bkgcolor.y; desc.bkgcolor[2] = bkgcolor.z;
    desc.eta = eta;
    desc.shadow_type = shadowType;
    desc.depth_cueing = depthCueing ? 1 : 0;
    desc.dc[0] = dc.x; desc.dc[1] = dc.y; desc.dc[2] = dc.z;
    desc.amin = amin; desc.amax = amax; desc.distmin = distmin; desc.distmax = distmax;
    desc.eye[0] = eyePos.x; desc.eye[1] = eyePos.y; desc.eye[2] = eyePos.z;

    cam = camera();
}

// Renderer::render(), include/Renderer.hpp:65-100, expression for expression.
// M_PI is the reference's own re-definition (global.hpp:14).
WrtCamera HostScene::camera() const {
    const double REF_M_PI = 3.1415926535897;
    V3 u = cross(viewdir, updir);
    u = normalized(u);
    V3 v = cross(u, viewdir);
    v = normalized(v);
    float d = 1.f;
    if (parallel_projection) d = 4.f;
    // degree2Radians(const float&) returns float: d * M_PI / 180.f in double, narrowed (global.hpp:87-89);
    // tan() then resolves to the float overload.
    float half_deg = hfov / 2.f;
    float rad = (float)(half_deg * REF_M_PI / 180.f);
    float width_half = std::fabs(std::tan(rad) * d);
    float aspect_ratio = width / (float)height;
    float height_half = width_half / aspect_ratio;
    V3 n = normalized(viewdir);
    V3 ul = eyePos + d * n - width_half * u + height_half * v;
    V3 ur = eyePos + d * n + width_half * u + height_half * v;
    V3 ll = eyePos + d * n - width_half * u - height_half * v;
    V3 delta_h, delta_v;
    if (width != 1) delta_h = (ur - ul) / (float)(width - 1);
    if (height != 1) delta_v = (ll - ul) / (float)(height - 1);
    V3 c_off_h = (ur - ul) / (float)(width * 2);
    V3 c_off_v = (ll - ul) / (float)(height * 2);
    WrtCamera c;
    memset(&c, 0, sizeof c);
    auto put = [](float* dst, const V3& a) { dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; };
    put(c.eye, eyePos); put(c.ul, ul); put(c.delta_h, delta_h); put(c.delta_v, delta_v);
    put(c.c_off_h, c_off_h); put(c.c_off_v, c_off_v); put(c.n, n);
    c.d = d;
    c.parallel = parallel_projection;
    c.width = width;
    c.height = height;
    return c;
}

} // namespace wrt
