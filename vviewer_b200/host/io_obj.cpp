/*
 * Wavefront OBJ import with the conventions the reference gets from assimp 5.2.5
 * (src/lib/vengine/core/io/AssimpLoadModel.cpp:60-135, 543-563: aiProcess_Triangulate | aiProcess_FlipUVs |
 * aiProcess_CalcTangentSpace, then uv.y = 1 - uv.y):
 *   - one Mesh per `o` object ("defaultobject" when the file has none), fan triangulation,
 *   - vertices de-duplicated per (v, vt, vn) triplet,
 *   - mesh uv = the file's (u, v) (FlipUVs and the engine's 1 - v cancel),
 *   - normals normalised, colour = 1,
 *   - tangent / bitangent = per-face UV-derivative frame computed on the FLIPPED uv (what CalcTangentSpace
 *     sees), Gram-Schmidt against the vertex normal, averaged over the faces sharing the vertex.
 * assimp is not available in this image; only the tangent matters downstream because the path tracer
 * rebuilds the bitangent as cross(n, t) (shaders/include/frame.glsl:23-33).
 */
#include "vengine.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <tuple>
#include <algorithm>

namespace vengine {

void computeTangents(Mesh &mesh) {
    size_t nv = mesh.vertices.size();
    std::vector<vec3> tan(nv, vec3(0, 0, 0)), bit(nv, vec3(0, 0, 0));
    for (size_t f = 0; f + 2 < mesh.indices.size(); f += 3) {
        uint32_t i0 = mesh.indices[f], i1 = mesh.indices[f + 1], i2 = mesh.indices[f + 2];
        const Vertex &a = mesh.vertices[i0], &b = mesh.vertices[i1], &c = mesh.vertices[i2];
        vec3 v(b.position[0] - a.position[0], b.position[1] - a.position[1], b.position[2] - a.position[2]);
        vec3 w(c.position[0] - a.position[0], c.position[1] - a.position[1], c.position[2] - a.position[2]);
        /* uv as assimp sees it after FlipUVs: (u, 1 - v) */
        float sx = b.uv[0] - a.uv[0], sy = (1.0f - b.uv[1]) - (1.0f - a.uv[1]);
        float tx = c.uv[0] - a.uv[0], ty = (1.0f - c.uv[1]) - (1.0f - a.uv[1]);
        float dirCorrection = (tx * sy - ty * sx) < 0.0f ? -1.0f : 1.0f;
        if (sx * ty == sy * tx) {
            sx = 0.0f; sy = 1.0f; tx = 1.0f; ty = 0.0f;
        }
        vec3 tangent = (w * sy - v * ty) * dirCorrection;
        vec3 bitangent = (v * tx - w * sx) * dirCorrection;
        for (uint32_t idx : {i0, i1, i2}) {
            const Vertex &p = mesh.vertices[idx];
            vec3 n(p.normal[0], p.normal[1], p.normal[2]);
            vec3 lt = tangent - n * vm::dot(tangent, n);
            vec3 lb = bitangent - n * vm::dot(bitangent, n);
            float ltl = vm::length(lt), lbl = vm::length(lb);
            if (ltl > 0 && std::isfinite(ltl)) tan[idx] = tan[idx] + lt / ltl;
            if (lbl > 0 && std::isfinite(lbl)) bit[idx] = bit[idx] + lb / lbl;
        }
    }
    for (size_t i = 0; i < nv; i++) {
        Vertex &p = mesh.vertices[i];
        vec3 n(p.normal[0], p.normal[1], p.normal[2]);
        vec3 t = tan[i], b = bit[i];
        float tl = vm::length(t), bl = vm::length(b);
        if (!(tl > 1e-12f) || !std::isfinite(tl)) {
            /* reconstruction used by the reference's sanitisation path (AssimpLoadModel.cpp:112-124) */
            vec3 t1 = vm::cross(n, vec3(0, 0, 1)), t2 = vm::cross(n, vec3(1, 0, 0));
            t = vm::length(t1) > vm::length(t2) ? t1 : t2;
            tl = vm::length(t);
            if (!(tl > 0)) { t = vec3(1, 0, 0); tl = 1; }
        }
        t = t / tl;
        if (!(bl > 1e-12f) || !std::isfinite(bl)) {
            b = vm::cross(n, t);
            bl = vm::length(b);
            if (!(bl > 0)) { b = vec3(0, 0, 1); bl = 1; }
        }
        b = b / bl;
        p.tangent[0] = t.x; p.tangent[1] = t.y; p.tangent[2] = t.z;
        p.bitangent[0] = b.x; p.bitangent[1] = b.y; p.bitangent[2] = b.z;
    }
}

namespace {

std::string folderOf(const std::string &p) {
    size_t k = p.find_last_of('/');
    return k == std::string::npos ? std::string("") : p.substr(0, k + 1);
}
std::string fileOf(const std::string &p) {
    size_t k = p.find_last_of('/');
    return k == std::string::npos ? p : p.substr(k + 1);
}
std::string stemOf(const std::string &p) {
    std::string f = fileOf(p);
    size_t dot = f.find_last_of('.');
    return dot == std::string::npos ? f : f.substr(0, dot);
}
std::string trimmed(const std::string &s) {
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string("") : s.substr(a, b - a + 1);
}

/* the .mtl statements the reference consumes through assimp (AssimpLoadModel.cpp:252-360): Kd, Ks, Ns, map_Kd, map_Bump / bump */
struct MtlRecord {
    std::string name;
    vec3 kd{0.6f, 0.6f, 0.6f}, ks{0, 0, 0}; /* assimp's ObjFile::Material defaults */
    float ns = 0.0f;
    std::string mapKd, mapBump;
};

/* texture statements may carry options ("-bm 1.0 file.png"): the file name is the last token */
std::string textureFileOf(const std::string &args) {
    std::istringstream ss(args);
    std::string tok, last;
    while (ss >> tok) last = tok;
    for (char &c : last)
        if (c == '\\') c = '/';
    return last;
}

void parseMTL(const std::string &path, std::vector<MtlRecord> &out) {
    std::ifstream in(path);
    if (!in) {
        std::fprintf(stderr, "loadOBJ(): material library %s not found\n", path.c_str());
        return;
    }
    std::string line;
    MtlRecord *cur = nullptr;
    while (std::getline(in, line)) {
        line = trimmed(line);
        if (line.empty() || line[0] == '#') continue;
        size_t sp = line.find_first_of(" \t");
        std::string key = line.substr(0, sp), args = sp == std::string::npos ? "" : trimmed(line.substr(sp));
        if (key == "newmtl") {
            out.push_back(MtlRecord());
            cur = &out.back();
            cur->name = args;
        } else if (!cur) {
            continue;
        } else if (key == "Kd") {
            std::sscanf(args.c_str(), "%f %f %f", &cur->kd.x, &cur->kd.y, &cur->kd.z);
        } else if (key == "Ks") {
            std::sscanf(args.c_str(), "%f %f %f", &cur->ks.x, &cur->ks.y, &cur->ks.z);
        } else if (key == "Ns") {
            std::sscanf(args.c_str(), "%f", &cur->ns);
        } else if (key == "map_Kd") {
            cur->mapKd = textureFileOf(args);
        } else if (key == "map_Bump" || key == "map_bump" || key == "bump") {
            cur->mapBump = textureFileOf(args);
        }
    }
}

bool black(vec3 c) { return std::fabs(c.x) <= 1e-6f && std::fabs(c.y) <= 1e-6f && std::fabs(c.z) <= 1e-6f; }

/* the pseudo-PBR mapping of assimpLoadMaterialsOBJ (AssimpLoadModel.cpp:252-360) */
ImportedMaterial materialFromMTL(const MtlRecord &m, const std::string &prefix, const std::string &objPath, const std::string &folder) {
    ImportedMaterial im;
    im.info = AssetInfo(prefix + ":" + m.name, objPath);
    im.info.embedded = true;
    im.type = ImportedMaterialType::PBR_STANDARD;
    im.albedo = vec4(m.kd, 1.0f);
    if (!m.mapKd.empty()) {
        auto img = std::make_shared<ImageU8>();
        int srcChannels = 0;
        if (loadImageU8(folder + m.mapKd, *img, true, &srcChannels)) {
            if (img->channels == 1) { /* STBI_rgb_alpha */
                auto wide = std::make_shared<ImageU8>();
                wide->width = img->width; wide->height = img->height; wide->channels = 4;
                wide->data.resize(img->data.size() * 4);
                for (size_t p = 0; p < img->data.size(); p++) {
                    wide->data[4 * p] = wide->data[4 * p + 1] = wide->data[4 * p + 2] = img->data[p];
                    wide->data[4 * p + 3] = 255;
                }
                img = wide;
            }
            ImportedTexture t;
            t.name = im.info.name + ":albedo";
            t.image = img;
            t.colorSpace = ColorSpace::sRGB;
            im.albedoTexture = t;
            if (black(vec3(im.albedo.x, im.albedo.y, im.albedo.z))) im.albedo = vec4(1, 1, 1, im.albedo.w);
            if (srcChannels == 4) {
                ImportedTexture a;
                a.name = im.info.name + ":alpha";
                a.colorSpace = ColorSpace::LINEAR;
                auto one = std::make_shared<ImageU8>();
                one->width = img->width; one->height = img->height; one->channels = 1;
                one->data.resize((size_t)img->width * img->height);
                for (size_t p = 0; p < one->data.size(); p++) one->data[p] = img->data[4 * p + 3];
                a.image = one;
                im.alphaTexture = a;
                if (im.albedo.w == 0) im.albedo.w = 1;
                im.transparent = true;
            }
        } else {
            std::fprintf(stderr, "loadOBJ(): failed to load texture %s\n", (folder + m.mapKd).c_str());
        }
    }
    float diffuseAmount = std::max(std::max(im.albedo.x, im.albedo.y), im.albedo.z);
    float specularAmount = std::max(std::max(m.ks.x, m.ks.y), m.ks.z);
    im.metallic = 1.0f - diffuseAmount / (diffuseAmount + specularAmount);
    im.roughness = 1.0f - m.ns / 100.0f;
    if (black(vec3(im.albedo.x, im.albedo.y, im.albedo.z)) && specularAmount != 0) im.albedo = vec4(m.ks, im.albedo.w);
    if (!m.mapBump.empty()) {
        auto img = std::make_shared<ImageU8>();
        int srcChannels = 0;
        /* some .mtl files use map_Bump for displacement: a 1-channel file is not taken as a normal map */
        if (loadImageU8(folder + m.mapBump, *img, true, &srcChannels) && srcChannels != 1) {
            ImportedTexture t;
            t.name = im.info.name + ":normal";
            t.image = img;
            t.colorSpace = ColorSpace::LINEAR;
            im.normalTexture = t;
        }
    }
    return im;
}

}  // namespace

/* Node tree as assimp's OBJ importer builds it: a root named after the file with one child per object ("o" and "g"
 * statements both open one), each holding one mesh per material used inside it (all of them named after the object). */
bool loadOBJ(const std::string &path, ImportedModelNode &root, std::vector<ImportedMaterial> *materials, std::string *err) {
    std::ifstream in(path);
    if (!in) {
        if (err) *err = "cannot open " + path;
        return false;
    }
    const std::string folder = folderOf(path);
    std::vector<vec3> P, N;
    std::vector<vec2> T;
    std::vector<MtlRecord> mtl;
    struct Builder {
        std::string object;
        int material; /* index into mtl, -1 = none */
        std::map<std::tuple<int, int, int>, uint32_t> lut;
        std::unique_ptr<Mesh> mesh;
    };
    std::vector<Builder> builders;
    std::vector<std::string> objectOrder;
    auto current = [&](const std::string &object, int material) -> Builder & {
        for (auto &b : builders)
            if (b.object == object && b.material == material) return b;
        bool known = false;
        for (auto &o : objectOrder) known = known || o == object;
        if (!known) objectOrder.push_back(object);
        builders.push_back(Builder());
        builders.back().object = object;
        builders.back().material = material;
        builders.back().mesh = std::make_unique<Mesh>();
        builders.back().mesh->name = object;
        return builders.back();
    };
    std::string curName = "defaultobject";
    int curMaterial = -1;
    bool usedWithoutMaterial = false;
    std::string line;
    while (std::getline(in, line)) {
        if (line.empty() || line[0] == '#') continue;
        const char *s = line.c_str();
        if (s[0] == 'v' && s[1] == ' ') {
            vec3 p;
            std::sscanf(s + 2, "%f %f %f", &p.x, &p.y, &p.z);
            P.push_back(p);
        } else if (s[0] == 'v' && s[1] == 't') {
            vec2 t;
            std::sscanf(s + 3, "%f %f", &t.x, &t.y);
            T.push_back(t);
        } else if (s[0] == 'v' && s[1] == 'n') {
            vec3 n;
            std::sscanf(s + 3, "%f %f %f", &n.x, &n.y, &n.z);
            N.push_back(n);
        } else if ((s[0] == 'o' || s[0] == 'g') && (s[1] == ' ' || s[1] == '\t')) {
            std::istringstream ss(line.substr(2));
            std::string name;
            if (ss >> name) curName = name;
        } else if (line.compare(0, 7, "mtllib ") == 0) {
            if (materials) parseMTL(folder + trimmed(line.substr(7)), mtl);
        } else if (line.compare(0, 7, "usemtl ") == 0) {
            std::string name = trimmed(line.substr(7));
            curMaterial = -1;
            for (size_t i = 0; i < mtl.size(); i++)
                if (mtl[i].name == name) curMaterial = (int)i;
        } else if (s[0] == 'f' && s[1] == ' ') {
            Builder &b = current(curName, materials ? curMaterial : -1);
            if (curMaterial < 0) usedWithoutMaterial = true;
            std::istringstream ss(line.substr(2));
            std::string tok;
            std::vector<uint32_t> poly;
            while (ss >> tok) {
                int vi = 0, ti = 0, ni = 0;
                const char *c = tok.c_str();
                vi = std::atoi(c);
                const char *s1 = std::strchr(c, '/');
                if (s1) {
                    if (s1[1] != '/') ti = std::atoi(s1 + 1);
                    const char *s2 = std::strchr(s1 + 1, '/');
                    if (s2) ni = std::atoi(s2 + 1);
                }
                if (vi < 0) vi = (int)P.size() + vi + 1;
                if (ti < 0) ti = (int)T.size() + ti + 1;
                if (ni < 0) ni = (int)N.size() + ni + 1;
                auto key = std::make_tuple(vi, ti, ni);
                auto it = b.lut.find(key);
                uint32_t idx;
                if (it == b.lut.end()) {
                    Vertex v{};
                    if (vi >= 1 && vi <= (int)P.size()) {
                        v.position[0] = P[vi - 1].x; v.position[1] = P[vi - 1].y; v.position[2] = P[vi - 1].z;
                    }
                    if (ti >= 1 && ti <= (int)T.size()) {
                        v.uv[0] = T[ti - 1].x;
                        v.uv[1] = T[ti - 1].y;
                    }
                    if (ni >= 1 && ni <= (int)N.size()) {
                        vec3 n = vm::normalize(N[ni - 1]);
                        v.normal[0] = n.x; v.normal[1] = n.y; v.normal[2] = n.z;
                    }
                    v.color[0] = v.color[1] = v.color[2] = 1.0f;
                    idx = (uint32_t)b.mesh->vertices.size();
                    b.mesh->vertices.push_back(v);
                    b.lut[key] = idx;
                } else {
                    idx = it->second;
                }
                poly.push_back(idx);
            }
            for (size_t k = 1; k + 1 < poly.size(); k++) {
                b.mesh->indices.push_back(poly[0]);
                b.mesh->indices.push_back(poly[k]);
                b.mesh->indices.push_back(poly[k + 1]);
            }
        }
    }
    if (builders.empty()) {
        if (err) *err = "no faces in " + path;
        return false;
    }
    /* materials: one record per newmtl in file order; faces without usemtl get assimp's "DefaultMaterial" */
    int defaultMaterial = -1;
    if (materials) {
        const std::string prefix = stemOf(path);
        materials->clear();
        for (const MtlRecord &m : mtl) materials->push_back(materialFromMTL(m, prefix, path, folder));
        if (usedWithoutMaterial) {
            MtlRecord def;
            def.name = "DefaultMaterial";
            defaultMaterial = (int)materials->size();
            materials->push_back(materialFromMTL(def, prefix, path, folder));
        }
    }
    root = ImportedModelNode();
    root.name = fileOf(path);
    for (const std::string &object : objectOrder) {
        ImportedModelNode child;
        child.name = object;
        for (auto &b : builders) {
            if (b.object != object || !b.mesh) continue;
            bool anyNormal = false;
            for (const Vertex &v : b.mesh->vertices) anyNormal = anyNormal || v.normal[0] != 0 || v.normal[1] != 0 || v.normal[2] != 0;
            if (!anyNormal) computeNormals(*b.mesh);
            computeTangents(*b.mesh);
            child.materialIndices.push_back(materials ? (b.material >= 0 ? b.material : defaultMaterial) : -1);
            child.meshes.push_back(std::move(b.mesh));
        }
        root.children.push_back(std::move(child));
    }
    return true;
}

}  // namespace vengine
