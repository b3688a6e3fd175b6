// obj_reader.cpp — Wavefront OBJ ingest for the hard-coded `bunny.obj` load of
// the reference's driver (src/main.cpp:46-56) and PPMGenerator::loadObj
// (include/PPMGenerator.hpp:86-125).
//
// Scope: `v`, `vt`, `vn` and triangular `f` records in the
// four index forms the vendored loader understands (include/OBJ_Loader.h:755-836).
// n-gon triangulation and .mtl parsing are out of scope; a face with more than
// three corners is rejected loudly instead of being triangulated differently.
//
// Behaviour that matters for pixel parity:
//   * coordinates go through std::stof (OBJ_Loader.h:545-547);
//   * a face without normals gets the UN-normalised cross(p1-p0, p2-p1) of the
//     UNSCALED positions on all three corners (OBJ_Loader.h:821-835);
//   * missing texture coordinates are (0,0) (OBJ_Loader.h:782), but loadObj is
//     called with textureIndex = -1 so the triangles are untextured;
//   * main.cpp scales positions by 20 and subtracts 3 from y and z afterwards.
#include <fstream>
#include <sstream>

#include "host_scene.hpp"

namespace wrt {

namespace {

std::vector<std::string> split_ws(const std::string& s) {
    std::vector<std::string> out;
    size_t i = 0;
    while (i < s.size()) {
        while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\r')) ++i;
        size_t b = i;
        while (i < s.size() && !(s[i] == ' ' || s[i] == '\t' || s[i] == '\r')) ++i;
        if (i > b) out.emplace_back(s, b, i - b);
    }
    return out;
}

template <class T>
const T& element(const std::vector<T>& v, const std::string& index) {   // OBJ_Loader.h:397-405
    int idx = std::stoi(index);
    if (idx < 0) idx = (int)v.size() + idx;
    else idx--;
    if (idx < 0 || idx >= (int)v.size()) throw ParseError("obj: index " + index + " out of range");
    return v[idx];
}

} // namespace

bool HostScene::loadObjLikeMain(const std::string& path, const Material& mtl, float scale, float dy, float dz) {
    if (path.size() < 4 || path.substr(path.size() - 4) != ".obj") return false;
    std::ifstream file(path);
    if (!file.is_open()) return false;

    std::vector<V3> P, N;
    std::vector<V2> T;
    std::string line;
    size_t added = 0;
    while (std::getline(file, line)) {
        std::vector<std::string> tok = split_ws(line);
        if (tok.empty()) continue;
        try {
            if (tok[0] == "v" && tok.size() >= 4) {
                P.emplace_back(std::stof(tok[1]), std::stof(tok[2]), std::stof(tok[3]));
            } else if (tok[0] == "vt" && tok.size() >= 3) {
                T.emplace_back(std::stof(tok[1]), std::stof(tok[2]));
            } else if (tok[0] == "vn" && tok.size() >= 4) {
                N.emplace_back(std::stof(tok[1]), std::stof(tok[2]), std::stof(tok[3]));
            } else if (tok[0] == "f") {
                if (tok.size() != 4)
                    throw ParseError("obj: only triangular faces are supported (n-gon triangulation is out of scope)");
                Object t;
                t.type = TRIANGLE;
                V3 pos[3], nor[3];
                V2 uv[3];
                bool noNormal = false;
                for (int i = 0; i < 3; i++) {
                    const std::string& c = tok[i + 1];
                    size_t s1 = c.find('/');
                    size_t s2 = s1 == std::string::npos ? std::string::npos : c.find('/', s1 + 1);
                    std::string a = c.substr(0, s1);
                    std::string b = s1 == std::string::npos ? "" : c.substr(s1 + 1, s2 == std::string::npos ? std::string::npos : s2 - s1 - 1);
                    std::string d = s2 == std::string::npos ? "" : c.substr(s2 + 1);
                    pos[i] = element(P, a);
                    uv[i] = V2(0.f, 0.f);
                    if (s1 == std::string::npos) {                 // P
                        noNormal = true;
                    } else if (s2 == std::string::npos) {          // P/T
                        uv[i] = element(T, b);
                        noNormal = true;
                    } else if (b.empty()) {                        // P//N
                        nor[i] = element(N, d);
                    } else {                                       // P/T/N
                        uv[i] = element(T, b);
                        nor[i] = element(N, d);
                    }
                }
                if (noNormal) {
                    V3 A = pos[1] - pos[0];
                    V3 B = pos[2] - pos[1];
                    V3 n = cross(A, B);
                    nor[0] = nor[1] = nor[2] = n;
                }
                for (int i = 0; i < 3; i++) {                      // main.cpp:48-54
                    pos[i] = pos[i] * scale;
                    pos[i].y += dy;
                    pos[i].z += dz;
                }
                t.v0 = pos[0]; t.v1 = pos[1]; t.v2 = pos[2];
                t.n0 = nor[0]; t.n1 = nor[1]; t.n2 = nor[2];
                t.uv0 = uv[0]; t.uv1 = uv[1]; t.uv2 = uv[2];
                t.mtl = mtl;
                t.textureIndex = -1;
                t.normalMapIndex = -1;
                t.isTextureActivated = false;
                t.initializeBound();
                objList.emplace_back(std::move(t));
                ++added;
            }
        } catch (const std::invalid_argument&) {
            throw ParseError("obj: malformed number in line: " + line);
        } catch (const std::out_of_range&) {
            throw ParseError("obj: number out of range in line: " + line);
        }
    }
    return added > 0 || !P.empty();
}

} // namespace wrt
