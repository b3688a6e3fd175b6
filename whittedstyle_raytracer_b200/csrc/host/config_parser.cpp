// config_parser.cpp — the reference's scene-config grammar, rewritten.
//
// Behavioural contract (all line numbers: /root/reference/include/PPMGenerator.hpp):
//   * whitespace-separated tokens read with `fin >> str` semantics, loop shape
//     `fin >> kw; while (!fin.eof()) { process(kw); fin >> kw; }`  (:179-183)
//     — so a keyword that is the very last token with no trailing whitespace is
//     silently ignored, exactly as in the reference;
//   * every argument read is preceded by checkFin() (:671-675);
//   * keyword table :371-618, objects :209-365, faces :679-823, textures :827-881;
//   * errors do not exit here: they throw ParseError carrying the text the
//     reference prints after "ERROR: " (or the whole line when the reference
//     prints it directly and the text starts with "ERROR:"); the CLI turns that
//     into the same console line + exit(-1).
#include <fstream>
#include <sstream>

#include "host_scene.hpp"

namespace wrt {

namespace {

// Emulates std::ifstream >> std::string, including eofbit/failbit behaviour.
struct TokenStream {
    std::string data;
    size_t pos = 0;
    bool eofbit = false, failbit = false;

    static bool is_space(char c) {
        return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f';
    }
    bool eof() const { return eofbit; }
    void read(std::string& s) {
        if (failbit || eofbit) {          // sentry fails: string untouched, failbit set
            failbit = true;
            return;
        }
        while (pos < data.size() && is_space(data[pos])) ++pos;
        if (pos >= data.size()) {         // nothing extracted
            eofbit = failbit = true;
            s.clear();
            return;
        }
        size_t b = pos;
        while (pos < data.size() && !is_space(data[pos])) ++pos;
        s.assign(data, b, pos - b);
        if (pos >= data.size()) eofbit = true;
    }
};

void checkPosInt(const std::string& str) {            // global.hpp:32-38
    for (char i : str)
        if (i < 48 || i > 57) throw ParseError(str + ": expect a positive number");
}

void checkFloat(const std::string& str) {             // global.hpp:41-84
    auto bad = [&]() { return ParseError(str + ": not a valid float number"); };
    if (str.empty()) throw bad();
    bool dotAppeared = false;
    size_t first = 0;
    if (str[0] == '-') {
        if (str.size() == 1) throw bad();
        first = 1;
    }
    for (size_t i = first; i < str.size(); i++) {
        // the reference's `i == first && str[first] < 48 || str[i] > 57`
        if ((i == first && str[first] < 48) || str[i] > 57) throw bad();
        if (str[i] == '.' && !dotAppeared && i != str.size() - 1) dotAppeared = true;
        else if (str[i] < 48 || str[i] > 57) throw bad();
    }
}

int to_int(const std::string& s) {
    try { return std::stoi(s); }
    catch (const std::exception&) { throw ParseError(s + ": stoi failed (the reference aborts here)"); }
}
float to_float(const std::string& s) {
    try { return std::stof(s); }
    catch (const std::exception&) { throw ParseError(s + ": stof failed (the reference aborts here)"); }
}

bool is_light_avatar(const Material& m) {             // PPMGenerator.hpp:268-273, :314-319
    return float_equal(1.f, m.diffuse.x) && float_equal(1.f, m.diffuse.y) && float_equal(1.f, m.diffuse.z) &&
           float_equal(1.f, m.specular.x) && float_equal(1.f, m.specular.y) && float_equal(1.f, m.specular.z) &&
           float_equal(1.f, m.ka) && float_equal(1.f, m.kd) && float_equal(1.f, m.ks) && float_equal(0.f, m.n);
}

struct Parser {
    HostScene& g;
    TokenStream fin;
    std::vector<V3> vertices, normals;
    std::vector<V2> textCoords;
    bool isTextureOn = false;
    int textIndex = -1, bumpIndex = -1;
    Material mtlcolor;

    explicit Parser(HostScene& s) : g(s) {}

    void checkFin() {
        if (fin.eof()) throw ParseError("Insufficient or invalid data as input, check your config file\n");
    }
    std::string next() { std::string s; checkFin(); fin.read(s); return s; }

    V3 vertexAt(int index) {                          // getEleIn, global.hpp:169-174
        if (index >= (int)vertices.size() || index < 0) throw ParseError("vertex index is out of bound");
        return vertices[index];
    }
    V3 normalAt(int index) {
        if (index >= (int)normals.size() || index < 0) throw ParseError("normal index is out of bound");
        return normals[index];
    }
    V2 uvAt(int index) {
        if (index >= (int)textCoords.size() || index < 0) throw ParseError("texture coordinate index is out of bound");
        return textCoords[index];
    }

    // ---- ASCII P3 texture reader, PPMGenerator.hpp:827-881 ----
    void loadTexture(const std::string& name, std::vector<Texture>& list) {
        for (auto& t : list) if (t.name == name) return;
        std::string path = name;
        if (!g.assetDir.empty() && !(name.size() && name[0] == '/')) path = g.assetDir + "/" + name;
        std::ifstream input(path, std::ios::in | std::ios::binary);
        if (!input.is_open()) throw ParseError("ERROR:: texture file does not exits, program terminates.\n");
        TokenStream in;
        in.data.assign(std::istreambuf_iterator<char>(input), std::istreambuf_iterator<char>());
        std::string b0, b1, b2;
        in.read(b0); in.read(b1); in.read(b2);
        if (b0 != "P3") throw ParseError("ERROR:: Need P3 keyword, program terminates.\n");
        Texture t;
        t.name = name;
        checkPosInt(b1); checkPosInt(b2);
        t.width = to_int(b1);
        t.height = to_int(b2);
        in.read(b0);                                  // maxval, ignored by the reference
        t.rgb.reserve((size_t)t.width * t.height);
        // fast path for the common case; semantics identical to three `>>` + stoi
        const char* p = in.data.data() + in.pos;
        const char* e = in.data.data() + in.data.size();
        for (int j = 0; j < t.height; j++)
            for (int i = 0; i < t.width; i++) {
                int c[3];
                for (int k = 0; k < 3; k++) {
                    while (p < e && TokenStream::is_space(*p)) ++p;
                    if (p >= e) throw ParseError(name + ": texture data is truncated (the reference aborts here)");
                    long v = 0;
                    const char* b = p;
                    while (p < e && !TokenStream::is_space(*p)) {
                        if (*p < 48 || *p > 57) throw ParseError(std::string(b, p + 1) + ": expect a positive number");
                        v = v * 10 + (*p - 48);
                        if (v > 2147483647L) throw ParseError(name + ": texel out of int range");
                        ++p;
                    }
                    c[k] = (int)v;
                }
                t.rgb.emplace_back(c[0] / 255.f, c[1] / 255.f, c[2] / 255.f);
            }
        list.emplace_back(std::move(t));
    }

    // ---- faces, PPMGenerator.hpp:679-823 ----
    void flatNormal(Object& t) {
        V3 e1 = t.v1 - t.v0, e2 = t.v2 - t.v0;
        V3 normal = normalized(cross(e1, e2));
        t.n0 = t.n1 = t.n2 = normal;
    }
    // Which of the reference's four corner forms a token is (std::regex_match against "[0-9]+", "[0-9]+//[0-9]+",
    // "[0-9]+/[0-9]+", "[0-9]+/[0-9]+/[0-9]+", PPMGenerator.hpp:289-292), without std::regex: 0 = none.
    enum FaceForm { FORM_NONE = 0, FORM_FLAT, FORM_SMOOTH, FORM_FLAT_TEXT, FORM_SMOOTH_TEXT };
    static FaceForm classify(const std::string& s, size_t& p, size_t& q) {
        size_t i = 0, n = s.size();
        auto digits = [&]() { size_t b = i; while (i < n && s[i] >= '0' && s[i] <= '9') ++i; return i > b; };
        if (!digits()) return FORM_NONE;
        if (i == n) return FORM_FLAT;
        if (s[i] != '/') return FORM_NONE;
        p = i++;
        if (i < n && s[i] == '/') {                       // v//n
            q = i++;
            if (!digits() || i != n) return FORM_NONE;
            return FORM_SMOOTH;
        }
        if (!digits()) return FORM_NONE;
        if (i == n) return FORM_FLAT_TEXT;                // v/t
        if (s[i] != '/') return FORM_NONE;
        q = i++;
        if (!digits() || i != n) return FORM_NONE;
        return FORM_SMOOTH_TEXT;                          // v/t/n
    }
    void processFace(const std::string tok[3], Object& t) {
        size_t p[3] = {0, 0, 0}, q[3] = {0, 0, 0};
        FaceForm f0 = classify(tok[0], p[0], q[0]), f1 = classify(tok[1], p[1], q[1]), f2 = classify(tok[2], p[2], q[2]);
        // all three corners must have the same form (the reference tests the forms in the order
        // flat, smooth, smooth_text, flat_text; the forms are mutually exclusive, so order is irrelevant)
        if (f0 == FORM_NONE || f0 != f1 || f0 != f2) throw ParseError("f face information is not valid");
        V3* vs[3] = {&t.v0, &t.v1, &t.v2};
        V3* ns[3] = {&t.n0, &t.n1, &t.n2};
        V2* ts[3] = {&t.uv0, &t.uv1, &t.uv2};
        for (int i = 0; i < 3; i++) {
            const std::string& s = tok[i];
            switch (f0) {
            case FORM_FLAT:
                *vs[i] = vertexAt(to_int(s) - 1);
                break;
            case FORM_SMOOTH:
                *vs[i] = vertexAt(to_int(s.substr(0, p[i])) - 1);
                *ns[i] = normalAt(to_int(s.substr(q[i] + 1)) - 1);
                break;
            case FORM_FLAT_TEXT:
                *vs[i] = vertexAt(to_int(s.substr(0, p[i])) - 1);
                *ts[i] = uvAt(to_int(s.substr(p[i] + 1)) - 1);
                break;
            default:
                *vs[i] = vertexAt(to_int(s.substr(0, p[i])) - 1);
                *ts[i] = uvAt(to_int(s.substr(p[i] + 1, q[i] - p[i] - 1)) - 1);
                *ns[i] = normalAt(to_int(s.substr(q[i] + 1)) - 1);
                break;
            }
        }
        if (f0 == FORM_FLAT || f0 == FORM_FLAT_TEXT) flatNormal(t);
    }

    void applyTextureState(Object& s) {               // :256-266, :322-331
        if (isTextureOn) {
            s.isTextureActivated = true;
            s.textureIndex = textIndex;
            if (bumpIndex != -1) {
                s.normalMapIndex = bumpIndex;
                bumpIndex = -1;
            }
        }
    }

    void readObject(const std::string& key) {         // :209-365
        if (key == "v") {
            std::string t0 = next(), t1 = next(), t2 = next();
            vertices.emplace_back(to_float(t0), to_float(t1), to_float(t2));
        } else if (key == "sphere") {
            Object s;
            s.type = SPHERE;
            s.mtl = mtlcolor;
            std::string t0 = next(), t1 = next(), t2 = next(), t3 = next();
            checkFloat(t0); checkFloat(t1); checkFloat(t2); checkFloat(t3);
            s.center = V3(to_float(t0), to_float(t1), to_float(t2));
            s.radius = to_float(t3);
            applyTextureState(s);
            if (is_light_avatar(s.mtl)) s.isLight = true;
            s.initializeBound();
            g.objList.emplace_back(std::move(s));
        } else if (key == "f") {
            std::string tok[3];
            tok[0] = next(); tok[1] = next(); tok[2] = next();
            Object t;
            t.type = TRIANGLE;
            t.mtl = mtlcolor;
            processFace(tok, t);
            if (is_light_avatar(t.mtl)) t.isLight = true;
            applyTextureState(t);
            t.initializeBound();
            g.objList.emplace_back(std::move(t));
        } else if (key == "vn") {
            std::string t0 = next(), t1 = next(), t2 = next();
            checkFloat(t0); checkFloat(t1); checkFloat(t2);
            normals.emplace_back(normalized(V3(to_float(t0), to_float(t1), to_float(t2))));
        } else if (key == "vt") {
            std::string t0 = next(), t1 = next();
            checkFloat(t0); checkFloat(t1);
            textCoords.emplace_back(to_float(t0), to_float(t1));
        }
    }

    V3 readVec3Checked() {
        std::string a = next(), b = next(), c = next();
        checkFloat(a); checkFloat(b); checkFloat(c);
        return V3(to_float(a), to_float(b), to_float(c));
    }

    void readLight(bool attenuated) {                 // :441-491
        std::string t[10];
        int n = attenuated ? 10 : 7;
        for (int i = 0; i < n; i++) t[i] = next();
        for (int i = 0; i < n; i++) checkFloat(t[i]);
        Light l;
        for (int i = 0; i < 4; i++) l.pos[i] = to_float(t[i]);
        l.color = V3(to_float(t[4]), to_float(t[5]), to_float(t[6]));
        if (attenuated) { l.c1 = to_float(t[7]); l.c2 = to_float(t[8]); l.c3 = to_float(t[9]); }
        l.initialize();
        g.lightList.emplace_back(l);
    }

    void processKeyword(const std::string& key) {     // :371-618
        if (key == "imsize") {
            std::string a = next(), b = next();
            checkPosInt(a); g.width = to_int(a);
            checkPosInt(b); g.height = to_int(b);
        } else if (key == "eye") {
            g.eyePos = readVec3Checked();
        } else if (key == "viewdir") {
            g.viewdir = readVec3Checked();
        } else if (key == "hfov") {
            std::string a = next();
            checkPosInt(a);
            g.hfov = to_int(a);
        } else if (key == "updir") {
            g.updir = readVec3Checked();
        } else if (key == "bkgcolor") {
            g.bkgcolor = readVec3Checked();
            std::string a = next();
            checkFloat(a);
            g.eta = to_float(a);
        } else if (key == "projection") {
            if (next() == "parallel") g.parallel_projection = 1;
        } else if (key == "light") {
            readLight(false);
        } else if (key == "attlight") {
            readLight(true);
        } else if (key == "mtlcolor") {
            std::string t[12];
            for (auto& s : t) s = next();
            for (auto& s : t) checkFloat(s);
            mtlcolor.diffuse = V3(to_float(t[0]), to_float(t[1]), to_float(t[2]));
            mtlcolor.specular = V3(to_float(t[3]), to_float(t[4]), to_float(t[5]));
            mtlcolor.ka = to_float(t[6]); mtlcolor.kd = to_float(t[7]); mtlcolor.ks = to_float(t[8]);
            mtlcolor.n = to_float(t[9]); mtlcolor.alpha = to_float(t[10]); mtlcolor.eta = to_float(t[11]);
            isTextureOn = false;
        } else if (key == "shadow") {
            if (next() == "soft") g.shadowType = 1;
        } else if (key == "depthcueing") {
            g.depthCueing = true;
            std::string t[7];
            for (auto& s : t) s = next();
            for (auto& s : t) checkFloat(s);
            g.dc = V3(to_float(t[0]), to_float(t[1]), to_float(t[2]));
            g.amax = to_float(t[3]); g.amin = to_float(t[4]);
            g.distmax = to_float(t[5]); g.distmin = to_float(t[6]);
        } else if (key == "texture" || key == "textrue") {   // "textrue": README typo, accepted as an alias
            size_t size0 = g.textures.size();
            std::string a = next();
            loadTexture(a, g.textures);
            isTextureOn = true;
            if (size0 == g.textures.size()) {
                for (size_t i = 0; i < g.textures.size(); i++)
                    if (g.textures[i].name == a) { textIndex = (int)i; break; }
            } else textIndex = (int)g.textures.size() - 1;
        } else if (key == "bump") {
            size_t size0 = g.normalMaps.size();
            std::string a = next();
            loadTexture(a, g.normalMaps);
            isTextureOn = true;
            if (size0 == g.normalMaps.size()) {
                for (size_t i = 0; i < g.normalMaps.size(); i++)
                    if (g.normalMaps[i].name == a) { bumpIndex = (int)i; break; }
            } else {
                bumpIndex = (int)g.normalMaps.size() - 1;
                for (V3& c : g.normalMaps[bumpIndex].rgb) {   // :600-606
                    c = c * 2.f;
                    c.x = c.x - 1.f; c.y = c.y - 1.f; c.z = c.z - 1.f;
                }
            }
        } else if (key == "sphere" || key == "v" || key == "f" || key == "vn" || key == "vt") {
            readObject(key);
        } else {
            throw ParseError("extraneous string in the input file\n");
        }
    }

    void run() {                                      // :174-197
        std::string keyWord;
        checkFin();
        fin.read(keyWord);
        while (!fin.eof()) {
            processKeyword(keyWord);
            fin.read(keyWord);
        }
        bool inited = g.width != -1 && g.height != -1 && !float_equal(g.eyePos.x, FLT_MAX) &&
                      !float_equal(g.viewdir.x, FLT_MAX) && g.hfov != -1 && !float_equal(g.updir.x, FLT_MAX) &&
                      !float_equal(g.bkgcolor.x, FLT_MAX);
        if (!inited) throw ParseError("insufficient input data: unable to initialize the program\n");
        if (float_equal(g.viewdir.x, g.updir.x) && float_equal(g.viewdir.y, g.updir.y) &&
            float_equal(g.viewdir.z, g.updir.z))
            throw ParseError("invalid viewPlane infomation: updir and view dir can't be the same");
    }
};

} // namespace

void Object::initializeBound() {
    if (type == TRIANGLE) {                           // Triangle.hpp:131-134, BoundBox.hpp:13-25,105-117
        bmin = V3(fminf(v0.x, v1.x), fminf(v0.y, v1.y), fminf(v0.z, v1.z));
        bmax = V3(fmaxf(v0.x, v1.x), fmaxf(v0.y, v1.y), fmaxf(v0.z, v1.z));
        V3 mn(fminf(bmin.x, v2.x), fminf(bmin.y, v2.y), fminf(bmin.z, v2.z));
        V3 mx(fmaxf(bmax.x, v2.x), fmaxf(bmax.y, v2.y), fmaxf(bmax.z, v2.z));
        // BoundBox(min,max) re-applies fmin/fmax pairwise
        bmin = V3(fminf(mn.x, mx.x), fminf(mn.y, mx.y), fminf(mn.z, mx.z));
        bmax = V3(fmaxf(mn.x, mx.x), fmaxf(mn.y, mx.y), fmaxf(mn.z, mx.z));
    } else {                                          // Sphere.hpp:123-127
        V3 mn(center.x - radius, center.y - radius, center.z - radius);
        V3 mx(center.x + radius, center.y + radius, center.z + radius);
        bmin = V3(fminf(mn.x, mx.x), fminf(mn.y, mx.y), fminf(mn.z, mx.z));
        bmax = V3(fmaxf(mn.x, mx.x), fmaxf(mn.y, mx.y), fmaxf(mn.z, mx.z));
    }
}

void Light::initialize() {                            // Light.hpp:28-42
    if (float_equal(pos[3], 0.f)) return;
    tv0 = V3(pos[0], pos[1], pos[2]);
    tv1 = V3(pos[0] + 7, pos[1], pos[2] - 7);
    tv2 = V3(pos[0], pos[1], pos[2] - 7);
    // `0.33333 * v` converts the double literal to float (friend operator*(float, Vector3f))
    V3 center = 0.33333f * tv0 + 0.33333f * tv1 + 0.33333f * tv2;
    V3 offset = tv0 - center;
    tv0 = tv0 + offset;
    tv1 = tv1 + offset;
    tv2 = tv2 + offset;
}

Material main_cpp_material(bool glass_variant) {      // src/main.cpp:22-43
    Material m;
    m.diffuse = V3(0.529, 0.807, 0.921);
    m.specular = V3(0.33, 0.66, 0.99);
    m.ka = 0.05;
    m.kd = 0.1;
    m.ks = glass_variant ? 0.2 : 0.1;
    m.n = 64;
    m.alpha = 0.2;
    m.eta = glass_variant ? 1.33 : 1.52;
    return m;
}

void HostScene::parseConfigText(const std::string& text) {
    Parser p(*this);
    p.fin.data = text;
    p.run();
}

void HostScene::parseConfigFile(const std::string& path) {
    std::ifstream f(path, std::ios::in | std::ios::binary);
    if (!f.is_open()) throw ParseError("ERROR:: inputfile does not exits, program terminates.\n");
    inputName = path;
    std::string text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    parseConfigText(text);
}

std::string HostScene::outputName() const {           // PPMGenerator.hpp:62-74
    size_t pos = inputName.find(".txt");
    if (pos == std::string::npos) return inputName + ".ppm";
    if (pos == 0) return ".ppm";
    return inputName.substr(0, pos) + ".ppm";
}

} // namespace wrt
