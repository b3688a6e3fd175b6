// dev_shade.cuh — hit-record completion, texturing, normal mapping, optics and
// Blinn-Phong, each in the reference's operation order.
#pragma once
#include "dev_traverse.cuh"

namespace wrt {

#define REF_M_PI 3.1415926535897   // the reference's own M_PI, global.hpp:14 (double)

struct Surface {              // the fields of Intersection the shader reads, Intersection.hpp:13-28
    f3 pos, nDir;
    float u, v;               // textPos
    int textureIndex, normalMapIndex;
    int prim, material;
    unsigned flags;
};

struct Mtl {                  // Material.hpp:6-16
    f3 diffuse, specular;
    float ka, kd, ks, n, alpha, eta;
};

__device__ __forceinline__ Mtl load_material(const DevScene& s, int m) {
    const float4* p = s.materials + 3 * (size_t)m;
    float4 a = ldg4(p), b = ldg4(p + 1), c = ldg4(p + 2);
    Mtl r;
    r.diffuse = mk3(a); r.ka = a.w;
    r.specular = mk3(b); r.kd = b.w;
    r.ks = c.x; r.n = c.y; r.alpha = c.z; r.eta = c.w;
    return r;
}

// The hit half of Triangle::intersect (Triangle.hpp:43-59) and Sphere::intersect
// (Sphere.hpp:49-74 / :93-117), from (t, prim, b1, b2) found by the traversal.
__device__ __forceinline__ Surface complete_hit(const DevScene& s, f3 orig, f3 dir, float t, int prim, float b1, float b2) {
    Surface sf;
    sf.prim = prim;
    sf.u = -1.f; sf.v = -1.f; sf.textureIndex = -1; sf.normalMapIndex = -1;
    int4 id = __ldg(s.ids + prim);
    sf.material = id.x;
    sf.flags = __float_as_uint(ldg4(s.geom + 3 * (size_t)prim + 2).w);
    sf.pos = orig + t * dir;
    if ((sf.flags & WRT_PRIM_KIND_MASK) == WRT_PRIM_TRIANGLE) {
        const float4* a = s.attr + 4 * (size_t)prim;
        float4 a0 = ldg4(a), a1 = ldg4(a + 1), a2 = ldg4(a + 2);
        float w0 = 1 - b1 - b2;
        sf.nDir = normalized((mk3(a0) * w0) + mk3(a1) * b1 + mk3(a2) * b2);
        if (sf.flags & WRT_PRIM_TEXTURED) {
            float4 a3 = ldg4(a + 3);
            // uv0 * w0 + uv1 * b1 + uv2 * b2  (Vector2f ops, Vector.hpp:48-54)
            sf.u = a0.w * w0 + a2.w * b1 + a3.y * b2;
            sf.v = a1.w * w0 + a3.x * b1 + a3.z * b2;
            sf.textureIndex = id.y;
            sf.normalMapIndex = id.z;
        }
    } else {
        float4 A = ldg4(s.geom + 3 * (size_t)prim);
        sf.nDir = normalized(sf.pos - mk3(A));
        if (sf.flags & WRT_PRIM_TEXTURED) {
            // acos/atan2 resolve to the float overloads in the reference; glibc's are
            // correctly rounded in practice, so round the double results once.
            float phi = (float)acos((double)sf.nDir.z);
            float v = (float)((double)phi / REF_M_PI);
            float theta = (float)atan2((double)sf.nDir.y, (double)sf.nDir.x);
            if (theta < 0) theta = (float)((double)theta + 2 * REF_M_PI);
            float u = (float)((double)theta / (2.0 * REF_M_PI));
            sf.u = u; sf.v = v;
            sf.textureIndex = id.y;
            sf.normalMapIndex = id.z;
        }
    }
    return sf;
}

// Texture::getRGBat, Texture.hpp:16-29: nearest texel by truncation, flat-index clamp.
__device__ __forceinline__ f3 texture_at(const DevScene& s, const WrtTexture* tex, float u, float v) {
    int w = tex->width, h = tex->height;
    if (w == 0 && h == 0) return mk3(0.f, 0.f, 0.f);
    int x = (int)(u * w);
    int y = (int)(v * h);
    int index = y * w + x;
    if (index < 0) index = 0;
    if ((long long)index >= tex->count) index = (int)(tex->count - 1);
    const float* p = s.texels + 3 * (size_t)(tex->offset + index);
    return mk3(__ldg(p), __ldg(p + 1), __ldg(p + 2));
}

// Renderer::changeNormalDir, Renderer.hpp:417-474 (including deltaV1 = uv1.y - uv1.y).
__device__ __noinline__ f3 change_normal_dir(const DevScene& s, const Surface& sf) {
    f3 color = texture_at(s, s.normalmaps + sf.normalMapIndex, sf.u, sf.v);
    f3 T, B, nDir;
    if ((sf.flags & WRT_PRIM_KIND_MASK) == WRT_PRIM_TRIANGLE) {
        const float4* g = s.geom + 3 * (size_t)sf.prim;
        f3 e1 = mk3(ldg4(g + 1)), e2 = mk3(ldg4(g + 2));
        nDir = normalized(cross(e1, e2));
        const float4* a = s.attr + 4 * (size_t)sf.prim;
        float4 a0 = ldg4(a), a1 = ldg4(a + 1), a2 = ldg4(a + 2), a3 = ldg4(a + 3);
        float uv0x = a0.w, uv0y = a1.w, uv1x = a2.w, uv1y = a3.x, uv2x = a3.y, uv2y = a3.z;
        float deltaU1 = uv1x - uv0x;
        float deltaV1 = uv1y - uv1y;
        float deltaU2 = uv2x - uv0x;
        float deltaV2 = uv2y - uv0y;
        float coef = 1 / (-deltaU1 * deltaV2 + deltaV1 * deltaU2);
        T = coef * (-deltaV2 * e1 + deltaV1 * e2);
        B = coef * (-deltaU2 * e1 + deltaU1 * e2);
        T = normalized(T);
        B = normalized(B);
    } else {
        nDir = sf.nDir;
        T = mk3(-nDir.y / sqrtf(nDir.x * nDir.x + nDir.y * nDir.y), nDir.x / sqrtf(nDir.x * nDir.x + nDir.y * nDir.y), 0.f);
        B = cross(nDir, T);
    }
    f3 res;
    res.x = T.x * color.x + B.x * color.y + nDir.x * color.z;
    res.y = T.y * color.x + B.y * color.y + nDir.y * color.z;
    res.z = T.z * color.x + B.z * color.y + nDir.z * color.z;
    return normalized(res);
}

// fresnel, global.hpp:185-205
__device__ __forceinline__ float fresnel(f3 Incident, f3 normal, float eta_i, float eta_t) {
    f3 I = normalized(-Incident);
    f3 N = normalized(normal);
    float cosI_N = dot(I, N);
    if (cosI_N < 0) N = -N;
    float F0 = ref_pow2((eta_t - eta_i) / (eta_t + eta_i));
    float Fr = F0 + (1 - F0) * (ref_powf(1 - (dot(I, N)), 5.f));
    return Fr;
}

// getReflectionDir, global.hpp:208-213
__device__ __forceinline__ f3 reflection_dir(f3 incident, f3 normal) {
    f3 I = -normalized(incident);
    f3 N = normalized(normal);
    return (2 * (dot(N, I))) * N - I;
}

// getRefractionDir, global.hpp:219-248 (zero vector on total internal reflection)
__device__ __forceinline__ f3 refraction_dir(f3 incident, f3 normal, float eta_i, float eta_t) {
    f3 I = normalized(-incident);
    f3 N = normalized(normal);
    float cos_theta_i = dot(N, I);
    {   // clamp(-1, 1, v) = std::max(lo, std::min(hi, v)), global.hpp:23-26
        float m = (cos_theta_i < 1.f) ? cos_theta_i : 1.f;
        cos_theta_i = (-1.f < m) ? m : -1.f;
    }
    if (cos_theta_i < 0) { N = -N; cos_theta_i = -cos_theta_i; }
    float sin_theta_i = sqrtf(1 - ref_pow2(cos_theta_i));
    float sin_theta_t = (eta_i / eta_t) * sin_theta_i;
    if (sin_theta_i > (eta_t / eta_i)) return mk3(0.f, 0.f, 0.f);
    float cos_theta_t = sqrtf(1 - ref_pow2(sin_theta_t));
    return cos_theta_t * (-N) + (eta_i / eta_t) * (cos_theta_i * N - I);
}

// The two geometric factors of one light's Blinn-Phong terms (Renderer.hpp:290-296, :320-326):
// mx = max(L.N, 0) and pw = pow(max(H.N, 0), n).  One function for k_shade and for the request
// culling in k_surface_spawn, so both see bit-identical values.  LAZY: pw is only evaluated when
// mx == 0 (the only case the culling asks about); otherwise it is returned as 1.
template <bool LAZY>
__device__ __forceinline__ void light_factors(const WrtLight* L, f3 p_eye_dir, f3 pos, f3 nDir, float n_exp, float& mx, float& pw) {
    f3 p_light_dir;
    if (float_equal(L->pos[3], 1.f)) p_light_dir = normalized(mk3(L->pos[0], L->pos[1], L->pos[2]) - pos);
    else p_light_dir = normalized(mk3(-L->pos[0], -L->pos[1], -L->pos[2]));
    float ndl = dot(p_light_dir, normalized(nDir));
    mx = (ndl < 0.f) ? 0.f : ndl;
    pw = 1.f;
    if (LAZY && mx != 0.f) return;
    f3 h = normalized(p_light_dir + p_eye_dir);
    float hn = dot(h, nDir);
    float mh = (hn < 0.f) ? 0.f : hn;
    // pow(+0, n) = +0 for n > 0 (C99 F.9.4.4): the common back-facing case needs no double-precision pow
    pw = (mh == 0.f && n_exp > 0.f) ? 0.f : ref_powf(mh, n_exp);
}

// True when light L adds exactly the same value to the pixel whatever its shadow coefficient is: both
// factors are 0, so the diffuse and the specular term are (finite, sign fixed by the colours) * 0 for
// every coefficient in [0, 1] — the light faces the back of the surface.  The reference still traces
// these shadow rays (Renderer.hpp:283-287 runs before the max()); their result is multiplied by 0.
__device__ __forceinline__ bool light_terms_vanish(const WrtLight* L, f3 p_eye_dir, f3 pos, f3 nDir, float n_exp) {
    float mx, pw;
    light_factors<true>(L, p_eye_dir, pos, nDir, n_exp, mx, pw);
    return mx == 0.f && pw == 0.f;
}

// Renderer::blinnPhongShader, Renderer.hpp:265-341.  `shadow[l]` holds the
// coefficient the shadow kernels produced for light l (hard: product; soft:
// number of unoccluded samples, divided by 50 here; directional: product).
__device__ __forceinline__ f3 blinn_phong(const DevScene& s, f3 rayOrig, f3 pos, f3 nDir, const Mtl& m,
                                          const float* shadow) {
    f3 p_eye_dir = normalized(rayOrig - pos);
    f3 ambient = m.ka * m.diffuse;
    f3 diffuse = mk3(0.f, 0.f, 0.f), specular = mk3(0.f, 0.f, 0.f);
    for (int li = 0; li < s.n_lights; li++) {
        const WrtLight* L = s.lights + li;
        f3 lcolor = mk3(L->color[0], L->color[1], L->color[2]);
        float sh = shadow[li];
        float mx, pw;
        light_factors<false>(L, p_eye_dir, pos, nDir, m.n, mx, pw);
        if (float_equal(L->pos[3], 1.f)) {
            f3 lightPos = mk3(L->pos[0], L->pos[1], L->pos[2]);
            float d_p_light = norm(lightPos - pos);
            float attenuation = 1.f;
            if (L->c1 >= 0.f) attenuation = 1.f / (L->c1 + L->c2 * d_p_light + L->c3 * d_p_light * d_p_light);
            if (s.shadow_type != 0) sh = sh / (float)WRT_SOFT_SAMPLES;     // sum / sampleNum, Renderer.hpp:413
            diffuse = diffuse + (((((sh * lcolor) * m.kd) * m.diffuse) * attenuation) * mx);
            specular = specular + (((((sh * lcolor) * m.ks) * m.specular) * attenuation) * pw);
        } else {
            diffuse = diffuse + ((((sh * lcolor) * m.kd) * m.diffuse) * mx);
            specular = specular + ((((sh * lcolor) * m.ks) * m.specular) * pw);
        }
    }
    f3 res = ambient + diffuse + specular;
    if (s.depth_cueing) {
        float p_eye_dist = norm(pos - mk3(s.eye[0], s.eye[1], s.eye[2]));
        float alpha = 0;
        if (p_eye_dist <= s.distmin) alpha = s.amax;
        else if (p_eye_dist >= s.distmax) alpha = s.amin;
        else alpha = s.amin + (s.amax - s.amin) * (s.distmax - p_eye_dist) / (s.distmax - s.distmin);
        res = alpha * res + (1 - alpha) * mk3(s.dc[0], s.dc[1], s.dc[2]);
    }
    return res;
}

} // namespace wrt
