/* wrt_scene.h — plain-C description of a flattened Whitted scene.
 *
 * This is the data contract between the host front-end (config parser, OBJ
 * reader, reference-identical BVH build), the CUDA render core (libwrt_cuda.so)
 * and the CPU oracle (oracle/whitted_oracle.c).  Everything is a pointer+count
 * into caller-owned host memory; the CUDA library copies during upload.
 *
 * What each piece replaces in the reference (paths relative to /root/reference):
 *   WrtNode      <- BVHNode / BoundBox          include/BVH.hpp:15-25, include/BoundBox.hpp:8-11
 *   prim_geom    <- Triangle::v0..v2, Sphere    include/Triangle.hpp:11-13, include/Sphere.hpp:8-9
 *   prim_normals <- Triangle::n0..n2            include/Triangle.hpp:16
 *   prim_uv      <- Triangle::uv0..uv2          include/Triangle.hpp:17
 *   WrtMaterial  <- Material                    include/Material.hpp:6-16
 *   WrtLight     <- Light (+ area triangle)     include/Light.hpp:8-43
 *   WrtTexture   <- Texture                     include/Texture.hpp:7-14
 *   globals      <- PPMGenerator private state  include/PPMGenerator.hpp:139-164
 *   WrtCamera    <- locals of Renderer::render  include/Renderer.hpp:65-100
 *   WrtHit       <- Intersection                include/Intersection.hpp:13-28
 */
#ifndef WRT_SCENE_H
#define WRT_SCENE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One BVH record, 32 bytes, 32-byte aligned.  Record 0 is the root (record 1
 * is padding so that every sibling pair starts on a 64-byte boundary).
 *   link >= 0 : internal node; its children are records link (left) and link+1 (right)
 *   link <  0 : leaf holding exactly one primitive, index ~link
 * Primitives are stored in depth-first leaf order, so a primitive's index is
 * also its left-to-right rank in the reference's pointer tree (the closest-hit
 * tie-break of BVH.hpp:157 becomes "smaller primitive index wins"). */
typedef struct WrtNode {
    float   pmin[3];
    int32_t link;
    float   pmax[3];
    int32_t pad;
} WrtNode;

#define WRT_PRIM_TRIANGLE   0u
#define WRT_PRIM_SPHERE     1u
#define WRT_PRIM_KIND_MASK  1u
#define WRT_PRIM_LIGHT      2u   /* Object::isLight (light avatar)      */
#define WRT_PRIM_TEXTURED   4u   /* Object::isTextureActivated          */

typedef struct WrtMaterial {      /* 48 bytes */
    float diffuse[3];
    float specular[3];
    float ka, kd, ks, n, alpha, eta;
} WrtMaterial;

typedef struct WrtLight {         /* 80 bytes */
    float pos[4];                 /* w == 1 point light, else directional */
    float color[3];
    float c1, c2, c3;             /* c1 < 0: no attenuation */
    float tri[9];                 /* area-light triangle v0,v1,v2 after Light::intialize() */
    float pad;
} WrtLight;

typedef struct WrtTexture {
    int32_t width, height;
    int64_t offset;               /* first texel, in texels, into WrtSceneDesc::texels */
    int64_t count;                /* rgb.size() */
} WrtTexture;

typedef struct WrtSceneDesc {
    int32_t n_nodes;              /* 0 when the scene holds no object */
    int32_t n_prims;
    int32_t n_materials;
    int32_t n_lights;
    int32_t n_textures;
    int32_t n_normalmaps;
    int64_t n_texels;             /* rgb triples in texels[] */

    const WrtNode*     nodes;         /* n_nodes */
    const float*       prim_geom;     /* n_prims x 12: tri = v0.xyz,0, E1.xyz,0, E2.xyz,0 (E1=v1-v0, E2=v2-v0);
                                                         sphere = c.xyz,r, 0... */
    const uint32_t*    prim_flags;    /* n_prims: WRT_PRIM_* */
    const int32_t*     prim_material; /* n_prims */
    const int32_t*     prim_texture;  /* n_prims: Object::textureIndex   (-1 none) */
    const int32_t*     prim_normalmap;/* n_prims: Object::normalMapIndex (-1 none) */
    const int32_t*     prim_object;   /* n_prims: index of the object in Scene::objList order */
    const int32_t*     object_prim;   /* n_prims: inverse of prim_object */
    const float*       prim_normals;  /* n_prims x 9: n0,n1,n2 (triangles; zeros for spheres) */
    const float*       prim_uv;       /* n_prims x 6: uv0,uv1,uv2 */
    const WrtMaterial* materials;     /* n_materials */
    const WrtLight*    lights;        /* n_lights */
    const WrtTexture*  textures;      /* n_textures */
    const WrtTexture*  normalmaps;    /* n_normalmaps (texels already remapped 2c-1) */
    const float*       texels;        /* n_texels x 3 */

    float   bkgcolor[3];
    float   eta;                  /* scene index of refraction (4th bkgcolor value) */
    int32_t shadow_type;          /* 0 hard, 1 soft */
    int32_t depth_cueing;
    float   dc[3];
    float   amin, amax, distmin, distmax;
    float   eye[3];               /* PPMGenerator::eyePos, used by depth cueing */
} WrtSceneDesc;

/* Camera vectors, computed on the host with the reference's own expressions
 * (Renderer.hpp:65-100) so that tan()/double promotion never runs on the GPU. */
typedef struct WrtCamera {
    float   eye[3];
    float   ul[3];
    float   delta_h[3];
    float   delta_v[3];
    float   c_off_h[3];
    float   c_off_v[3];
    float   n[3];                 /* normalized viewdir */
    float   d;                    /* 1 perspective, 4 parallel */
    int32_t parallel;
    int32_t width, height;
} WrtCamera;

/* Batch form of the reference's Intersection out-parameter. */
typedef struct WrtHit {           /* 60 bytes */
    int32_t hit;                  /* Intersection::intersected */
    int32_t object;               /* index into Scene::objList, -1 on miss */
    float   t;                    /* FLT_MAX on miss */
    float   pos[3];
    float   ndir[3];
    float   uv[2];                /* (-1,-1) when the object carries no texture */
    int32_t texture;
    int32_t normalmap;
    int32_t material;
    int32_t prim;                 /* flattened primitive index (DFS rank), -1 on miss */
} WrtHit;

/* Ray counters, same definition as SURVEY.md section 3.3: one "ray" is one
 * closest-hit query (UpdateInter) or one shadow query. */
typedef struct WrtStats {
    int64_t closest_rays;
    int64_t shadow_rays;
    int64_t rays_per_depth[9];
    int64_t shadow_requests;      /* (hit, light) pairs; x50 samples when soft */
    int64_t box_tests;            /* filled by the oracle only */
    int64_t prim_tests;           /* filled by the oracle only */
    int32_t overflow_retries;
    int32_t pad;
    float   gpu_ms;               /* device time of the last wrt_render* call */
    float   pad2;
    /* CUDA path only.  Shadow requests answered without tracing; they stay counted in shadow_requests /
     * shadow_rays, which follow the reference's definition (it traces them): */
    int64_t shaft_culled_requests;  /* soft shadows: the whole shaft to the area light misses every leaf box => 50 lit */
    int64_t unlit_skipped_requests; /* the light's diffuse and specular factors are exactly 0 => coefficient unused */
    int64_t shadow_rays_traced;     /* shadow rays the kernels really traced */
} WrtStats;

#define WRT_MAX_DEPTH 9           /* Renderer.hpp:25 */
#define WRT_SOFT_SAMPLES 50       /* Renderer.hpp:407 */
#ifndef WRT_DEFAULT_SEED
#define WRT_DEFAULT_SEED 0x5EEDu   /* soft-shadow RNG seed, see wrt_rng.h */
#endif

#ifdef __cplusplus
}
#endif
#endif /* WRT_SCENE_H */
