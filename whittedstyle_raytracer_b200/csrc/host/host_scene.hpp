// host_scene.hpp — host-side source of truth for a parsed scene.
//
// Mirrors the state the reference keeps in PPMGenerator / Scene
// (include/PPMGenerator.hpp:139-164, include/Scene.hpp:12-45) as plain data,
// plus the flattened arrays (WrtSceneDesc) handed to the CUDA core.
#pragma once
#include <cfloat>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/wrt_scene.h"
#include "vecmath.hpp"

namespace wrt {

struct Material {                       // Material.hpp:6-16
    V3 diffuse, specular;
    float ka = 0, kd = 0, ks = 0, n = 0;
    float alpha = 0, eta = 0;           // uninitialised in the reference; 0 here
};

enum ObjType { TRIANGLE = 0, SPHERE = 1 };   // Object.hpp:9-13

struct Object {                         // Object.hpp:15-40 + Triangle.hpp:11-17 + Sphere.hpp:8-9
    ObjType type = TRIANGLE;
    Material mtl;
    bool isLight = false;
    bool isTextureActivated = false;
    int textureIndex = -1;
    int normalMapIndex = -1;
    V3 v0, v1, v2;
    V3 n0, n1, n2;
    V2 uv0, uv1, uv2;
    V3 center;
    float radius = 1.f;
    V3 bmin, bmax;                      // BoundBox
    void initializeBound();
};

struct Light {                          // Light.hpp:8-43
    float pos[4] = {0, 0, 0, 0};
    V3 color;
    float c1 = -1, c2 = -1, c3 = -1;
    V3 tv0, tv1, tv2;
    void initialize();
};

struct Texture {                        // Texture.hpp:7-14
    std::string name;
    int width = 0, height = 0;
    std::vector<V3> rgb;
};

struct ParseError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// The hard-coded material of src/main.cpp:22-32 (the active, "water" variant)
// and the commented-out "glass" variant of :35-43.
Material main_cpp_material(bool glass_variant);

struct HostScene {
    // ---- PPMGenerator state ----
    int width = -1, height = -1;
    V3 eyePos{FLT_MAX, 0, 0}, viewdir{FLT_MAX, 0, 0}, updir{FLT_MAX, 0, 0}, bkgcolor{FLT_MAX, 0, 0};
    int hfov = -1;
    float eta = 0;
    int parallel_projection = 0;
    int shadowType = 0;
    bool depthCueing = false;
    V3 dc;
    float amin = 0, amax = 0, distmin = 0, distmax = 0;
    std::vector<Object> objList;
    std::vector<Light> lightList;
    std::vector<Texture> textures, normalMaps;
    std::string inputName;
    std::string assetDir;               // "" = cwd (reference behaviour)

    // ---- parsing (config_parser.cpp / obj_reader.cpp) ----
    void parseConfigFile(const std::string& path);
    void parseConfigText(const std::string& text);
    // src/main.cpp:46-56: returns false (scene unchanged) when the file is absent
    bool loadObjLikeMain(const std::string& path, const Material& mtl,
                         float scale = 20.f, float dy = -3.f, float dz = -3.f);
    std::string outputName() const;     // PPMGenerator.hpp:62-74

    // ---- flattening (bvh_build.cpp) ----
    void buildAndFlatten();             // reference-identical BVH + SoA arrays
    WrtCamera camera() const;           // Renderer.hpp:65-100

    // flattened storage
    std::vector<WrtNode> nodes;
    std::vector<float> prim_geom, prim_normals, prim_uv, texels;
    std::vector<uint32_t> prim_flags;
    std::vector<int32_t> prim_material, prim_texture, prim_normalmap, prim_object, object_prim;
    std::vector<WrtMaterial> materials;
    std::vector<WrtLight> lights;
    std::vector<WrtTexture> tex_desc, nmap_desc;
    int bvh_depth = 0;
    WrtSceneDesc desc{};
    WrtCamera cam{};
};

} // namespace wrt
