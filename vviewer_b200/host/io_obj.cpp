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

bool loadOBJ(const std::string &path, Model3D &out, std::string *err) {
    std::ifstream in(path);
    if (!in) {
        if (err) *err = "cannot open " + path;
        return false;
    }
    std::vector<vec3> P, N;
    std::vector<vec2> T;
    struct Builder {
        std::string name;
        std::map<std::tuple<int, int, int>, uint32_t> lut;
        std::unique_ptr<Mesh> mesh;
    };
    std::vector<Builder> builders;
    auto current = [&](const std::string &name) -> Builder & {
        for (auto &b : builders)
            if (b.name == name) return b;
        builders.push_back(Builder());
        builders.back().name = name;
        builders.back().mesh = std::make_unique<Mesh>();
        builders.back().mesh->name = name;
        return builders.back();
    };
    std::string curName = "defaultobject";
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
        } else if (s[0] == 'o' && s[1] == ' ') {
            std::istringstream ss(line.substr(2));
            ss >> curName;
        } else if (s[0] == 'f' && s[1] == ' ') {
            Builder &b = current(curName);
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
    for (auto &b : builders) {
        computeTangents(*b.mesh);
        out.meshes.push_back(std::move(b.mesh));
    }
    return true;
}

}  // namespace vengine
