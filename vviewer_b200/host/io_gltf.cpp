/*
 * glTF 2.0 import (.gltf with embedded base64 / external buffers and images, and binary .glb) with the conventions
 * the reference obtains from assimp 5.2.5's glTF2 importer + its own post-processing
 * (src/lib/vengine/core/io/AssimpLoadModel.cpp:52-160 meshes and nodes, :363-530 materials, :533-613 entry points):
 *   - one Mesh per primitive, named after the glTF mesh ("-<i>" appended when a mesh has several primitives);
 *   - the node tree of the default scene; a scene with several roots hangs them under a node called "ROOT";
 *     node matrices are decomposed into translation / rotation / scale (glm::decompose, :154-160);
 *   - mesh uv = (u, 1 - v_gltf): assimp converts to v-up, aiProcess_FlipUVs and the engine's `1 - v` cancel (:86);
 *   - normals normalised, colour = 1; tangents from the file when present (bitangent = cross(n, t) * w), otherwise the
 *     UV-derivative frame of computeTangents();
 *   - materials become PBR_STANDARD records named "<file stem>:<material name>" with textures "<material>:albedo",
 *     ":roughness" (G of metallicRoughness), ":metallic" (B), ":ao" (G of the occlusion map, :463), ":emissive",
 *     ":normal", ":alpha" (only when the base-colour image has 4 channels, which also flags the material transparent);
 *     an emissive texture forces a non-black emissive colour and a non-zero strength (:505-511);
 *   - every image is decoded bottom-row-first, as stbi does once the engine has switched on its global vertical flip
 *     (core/Image.cpp:39, SURVEY trap T11).
 * Not handled (as in the reference path): skins, animations, morph targets, cameras, KHR_texture_transform, Draco.
 */
#include "json.hpp"
#include "vengine.hpp"

#include <cstdio>
#include <cstring>
#include <fstream>

namespace vengine {

Transform decomposeTransform(const mat4 &m) {
    vec3 t(m[3][0], m[3][1], m[3][2]);
    vec3 c0(m[0][0], m[0][1], m[0][2]), c1(m[1][0], m[1][1], m[1][2]), c2(m[2][0], m[2][1], m[2][2]);
    /* Gram-Schmidt over the columns, scale = the lengths that remain (shear is dropped) */
    vec3 s;
    s.x = vm::length(c0);
    if (s.x > 0) c0 = c0 / s.x;
    c1 = c1 - c0 * vm::dot(c0, c1);
    s.y = vm::length(c1);
    if (s.y > 0) c1 = c1 / s.y;
    c2 = c2 - c0 * vm::dot(c0, c2);
    c2 = c2 - c1 * vm::dot(c1, c2);
    s.z = vm::length(c2);
    if (s.z > 0) c2 = c2 / s.z;
    /* a mirrored basis becomes a rotation with all three scales negated */
    if (vm::dot(c0, vm::cross(c1, c2)) < 0) {
        s = s * -1.0f;
        c0 = c0 * -1.0f;
        c1 = c1 * -1.0f;
        c2 = c2 * -1.0f;
    }
    mat4 r(1.0f);
    r[0][0] = c0.x; r[0][1] = c0.y; r[0][2] = c0.z;
    r[1][0] = c1.x; r[1][1] = c1.y; r[1][2] = c1.z;
    r[2][0] = c2.x; r[2][1] = c2.y; r[2][2] = c2.z;
    return Transform(t, s, vm::quat_cast(r));
}

void computeNormals(Mesh &mesh) {
    std::vector<vec3> acc(mesh.vertices.size(), vec3(0, 0, 0));
    for (size_t f = 0; f + 2 < mesh.indices.size(); f += 3) {
        uint32_t i[3] = {mesh.indices[f], mesh.indices[f + 1], mesh.indices[f + 2]};
        vec3 p[3];
        for (int k = 0; k < 3; k++) p[k] = vec3(mesh.vertices[i[k]].position[0], mesh.vertices[i[k]].position[1], mesh.vertices[i[k]].position[2]);
        vec3 n = vm::cross(p[1] - p[0], p[2] - p[0]); /* length = 2 x area */
        for (int k = 0; k < 3; k++) acc[i[k]] = acc[i[k]] + n;
    }
    for (size_t v = 0; v < mesh.vertices.size(); v++) {
        float l = vm::length(acc[v]);
        vec3 n = l > 0 ? acc[v] / l : vec3(0, 1, 0);
        mesh.vertices[v].normal[0] = n.x;
        mesh.vertices[v].normal[1] = n.y;
        mesh.vertices[v].normal[2] = n.z;
    }
}

namespace {

using json::Value;

std::string dirOfPath(const std::string &p) {
    size_t k = p.find_last_of('/');
    return k == std::string::npos ? std::string("") : p.substr(0, k + 1);
}
std::string stemOfPath(const std::string &p) {
    size_t slash = p.find_last_of('/');
    size_t start = slash == std::string::npos ? 0 : slash + 1;
    size_t dot = p.find_last_of('.');
    if (dot == std::string::npos || dot < start) dot = p.size();
    return p.substr(start, dot - start);
}

bool readWholeFile(const std::string &path, std::vector<uint8_t> &out) {
    std::ifstream in(path, std::ios::binary);
    if (!in) return false;
    in.seekg(0, std::ios::end);
    std::streamoff n = in.tellg();
    in.seekg(0, std::ios::beg);
    out.resize((size_t)std::max<std::streamoff>(n, 0));
    if (n > 0) in.read(reinterpret_cast<char *>(out.data()), n);
    return (bool)in || in.eof();
}

bool decodeBase64(const char *s, size_t n, std::vector<uint8_t> &out) {
    static int8_t lut[256];
    static bool init = false;
    if (!init) {
        std::memset(lut, -1, sizeof(lut));
        const char *abc = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
        for (int i = 0; i < 64; i++) lut[(uint8_t)abc[i]] = (int8_t)i;
        lut[(uint8_t)'-'] = 62; /* URL-safe alphabet */
        lut[(uint8_t)'_'] = 63;
        init = true;
    }
    out.clear();
    out.reserve(n / 4 * 3 + 3);
    uint32_t acc = 0;
    int bits = 0;
    for (size_t i = 0; i < n; i++) {
        uint8_t c = (uint8_t)s[i];
        if (c == '=') break;
        if (c == '\n' || c == '\r' || c == ' ') continue;
        int v = lut[c];
        if (v < 0) return false;
        acc = (acc << 6) | (uint32_t)v;
        bits += 6;
        if (bits >= 8) {
            bits -= 8;
            out.push_back((uint8_t)(acc >> bits));
        }
    }
    return true;
}

std::string percentDecode(const std::string &s) {
    std::string out;
    for (size_t i = 0; i < s.size(); i++) {
        if (s[i] == '%' && i + 2 < s.size() + 0 && std::isxdigit((unsigned char)s[i + 1]) && std::isxdigit((unsigned char)s[i + 2])) {
            out.push_back((char)std::strtol(s.substr(i + 1, 2).c_str(), nullptr, 16));
            i += 2;
        } else {
            out.push_back(s[i]);
        }
    }
    return out;
}

struct Loader {
    std::string path, folder, stem;
    Value doc;
    std::vector<uint8_t> glbBin;
    bool hasGlbBin = false;
    std::vector<std::vector<uint8_t>> buffers;
    std::vector<bool> bufferLoaded;
    struct DecodedImage {
        bool tried = false, ok = false;
        std::shared_ptr<ImageU8> rgba;
        int srcChannels = 0;
    };
    std::vector<DecodedImage> images;
    std::string error;

    bool fail(const std::string &msg) {
        error = msg;
        return false;
    }

    const Value *arrayOf(const char *key) const {
        const Value *v = doc.find(key);
        return v && v->isArray() ? v : nullptr;
    }
    size_t count(const char *key) const {
        const Value *v = arrayOf(key);
        return v ? v->size() : 0;
    }

    /* data: URI or a path relative to the .gltf */
    bool loadUri(const std::string &uri, std::vector<uint8_t> &out) {
        if (uri.compare(0, 5, "data:") == 0) {
            size_t comma = uri.find(',');
            if (comma == std::string::npos) return false;
            if (uri.substr(0, comma).find(";base64") == std::string::npos) return false;
            return decodeBase64(uri.data() + comma + 1, uri.size() - comma - 1, out);
        }
        return readWholeFile(folder + percentDecode(uri), out);
    }

    bool open() {
        std::vector<uint8_t> file;
        if (!readWholeFile(path, file)) return fail("cannot open " + path);
        std::string text;
        if (file.size() >= 12 && std::memcmp(file.data(), "glTF", 4) == 0) {
            /* GLB container: 12-byte header, then chunks (length, type, data); JSON first, BIN optional */
            auto le32 = [&](size_t o) { return (uint32_t)file[o] | ((uint32_t)file[o + 1] << 8) | ((uint32_t)file[o + 2] << 16) | ((uint32_t)file[o + 3] << 24); };
            if (le32(4) != 2) return fail("unsupported GLB version");
            size_t pos = 12;
            while (pos + 8 <= file.size()) {
                uint32_t len = le32(pos), type = le32(pos + 4);
                if (pos + 8 + (size_t)len > file.size()) return fail("truncated GLB chunk");
                if (type == 0x4E4F534Au) { /* "JSON" */
                    text.assign(reinterpret_cast<const char *>(&file[pos + 8]), len);
                } else if (type == 0x004E4942u && !hasGlbBin) { /* "BIN\0" */
                    glbBin.assign(file.begin() + (long)pos + 8, file.begin() + (long)pos + 8 + len);
                    hasGlbBin = true;
                }
                pos += 8 + (size_t)((len + 3u) & ~3u);
            }
            if (text.empty()) return fail("GLB without a JSON chunk");
        } else {
            text.assign(reinterpret_cast<const char *>(file.data()), file.size());
        }
        try {
            doc = json::parse(text);
        } catch (std::exception &e) {
            return fail(std::string("malformed glTF: ") + e.what());
        }
        if (!doc.isObject()) return fail("malformed glTF: root is not an object");
        const Value *asset = doc.find("asset");
        if (!asset || asset->string("version", "")[0] != '2') return fail("only glTF 2.x is supported");
        buffers.resize(count("buffers"));
        bufferLoaded.assign(buffers.size(), false);
        images.resize(count("images"));
        return true;
    }

    const std::vector<uint8_t> *buffer(size_t i) {
        if (i >= buffers.size()) return nullptr;
        if (!bufferLoaded[i]) {
            const Value &b = (*arrayOf("buffers"))[i];
            const Value *uri = b.find("uri");
            if (uri && uri->isString()) {
                if (!loadUri(uri->getString(), buffers[i])) return nullptr;
            } else if (i == 0 && hasGlbBin) {
                buffers[i] = glbBin;
            } else {
                return nullptr;
            }
            bufferLoaded[i] = true;
        }
        return &buffers[i];
    }

    /* ---- accessors */
    static int componentSize(int ct) {
        switch (ct) {
            case 5120: case 5121: return 1;
            case 5122: case 5123: return 2;
            case 5125: case 5126: return 4;
        }
        return 0;
    }
    static int typeWidth(const std::string &t) {
        if (t == "SCALAR") return 1;
        if (t == "VEC2") return 2;
        if (t == "VEC3") return 3;
        if (t == "VEC4") return 4;
        if (t == "MAT2") return 4;
        if (t == "MAT3") return 9;
        if (t == "MAT4") return 16;
        return 0;
    }
    static double readComponent(const uint8_t *p, int ct, bool normalized) {
        switch (ct) {
            case 5120: { int8_t v; std::memcpy(&v, p, 1); return normalized ? std::max(v / 127.0, -1.0) : v; }
            case 5121: { uint8_t v = *p; return normalized ? v / 255.0 : v; }
            case 5122: { int16_t v; std::memcpy(&v, p, 2); return normalized ? std::max(v / 32767.0, -1.0) : v; }
            case 5123: { uint16_t v; std::memcpy(&v, p, 2); return normalized ? v / 65535.0 : v; }
            case 5125: { uint32_t v; std::memcpy(&v, p, 4); return v; }
            case 5126: { float v; std::memcpy(&v, p, 4); return v; }
        }
        return 0.0;
    }
    /* raw view of count x width components inside a bufferView */
    bool viewSpan(const Value &view, size_t byteOffset, size_t elemBytes, size_t n, const uint8_t *&base, size_t &stride) {
        const std::vector<uint8_t> *buf = buffer((size_t)view.number("buffer", -1));
        if (!buf) return fail("glTF buffer missing or unreadable");
        size_t vOff = (size_t)view.number("byteOffset", 0), vLen = (size_t)view.number("byteLength", 0);
        stride = (size_t)view.number("byteStride", 0);
        if (stride == 0) stride = elemBytes;
        if (n == 0) {
            base = buf->data();
            return true;
        }
        size_t last = byteOffset + (n - 1) * stride + elemBytes;
        if (last > vLen || vOff + last > buf->size()) return fail("glTF accessor out of range");
        base = buf->data() + vOff + byteOffset;
        return true;
    }
    /* accessor -> count x width doubles (sparse accessors applied) */
    bool readAccessor(int index, std::vector<double> &out, int &width, size_t &n) {
        const Value *accs = arrayOf("accessors");
        if (!accs || index < 0 || (size_t)index >= accs->size()) return fail("glTF accessor index out of range");
        const Value &a = (*accs)[(size_t)index];
        int ct = (int)a.number("componentType", 0);
        width = typeWidth(a.string("type", ""));
        n = (size_t)a.number("count", 0);
        bool normalized = a.boolean("normalized", false);
        int cs = componentSize(ct);
        if (!cs || !width) return fail("glTF accessor with unknown type");
        out.assign(n * (size_t)width, 0.0);
        const Value *views = arrayOf("bufferViews");
        if (a.has("bufferView")) {
            size_t vi = (size_t)a.number("bufferView", 0);
            if (!views || vi >= views->size()) return fail("glTF bufferView index out of range");
            const uint8_t *base;
            size_t stride;
            if (!viewSpan((*views)[vi], (size_t)a.number("byteOffset", 0), (size_t)cs * width, n, base, stride)) return false;
            for (size_t i = 0; i < n; i++)
                for (int c = 0; c < width; c++) out[i * width + c] = readComponent(base + i * stride + (size_t)c * cs, ct, normalized);
        }
        const Value *sparse = a.find("sparse");
        if (sparse && sparse->isObject()) {
            size_t sn = (size_t)sparse->number("count", 0);
            const Value &si = (*sparse)["indices"], &sv = (*sparse)["values"];
            int ict = (int)si.number("componentType", 0);
            const uint8_t *ib, *vb;
            size_t istride, vstride;
            if (!views) return fail("glTF sparse accessor without bufferViews");
            if (!viewSpan((*views)[(size_t)si.number("bufferView", 0)], (size_t)si.number("byteOffset", 0), (size_t)componentSize(ict), sn, ib, istride)) return false;
            if (!viewSpan((*views)[(size_t)sv.number("bufferView", 0)], (size_t)sv.number("byteOffset", 0), (size_t)cs * width, sn, vb, vstride)) return false;
            for (size_t k = 0; k < sn; k++) {
                size_t target = (size_t)readComponent(ib + k * istride, ict, false);
                if (target >= n) return fail("glTF sparse index out of range");
                for (int c = 0; c < width; c++) out[target * width + c] = readComponent(vb + k * vstride + (size_t)c * cs, ct, normalized);
            }
        }
        return true;
    }

    /* ---- images */
    DecodedImage &image(size_t i) {
        static DecodedImage none;
        if (i >= images.size()) return none;
        DecodedImage &im = images[i];
        if (im.tried) return im;
        im.tried = true;
        const Value &rec = (*arrayOf("images"))[i];
        std::vector<uint8_t> bytes;
        const Value *uri = rec.find("uri");
        if (uri && uri->isString()) {
            if (!loadUri(uri->getString(), bytes)) {
                std::fprintf(stderr, "loadGLTF(): failed to load image %zu of %s\n", i, path.c_str());
                return im;
            }
        } else if (rec.has("bufferView")) {
            const Value *views = arrayOf("bufferViews");
            size_t vi = (size_t)rec.number("bufferView", 0);
            if (!views || vi >= views->size()) return im;
            const Value &view = (*views)[vi];
            const std::vector<uint8_t> *buf = buffer((size_t)view.number("buffer", -1));
            size_t off = (size_t)view.number("byteOffset", 0), len = (size_t)view.number("byteLength", 0);
            if (!buf || off + len > buf->size()) return im;
            bytes.assign(buf->begin() + (long)off, buf->begin() + (long)(off + len));
        } else {
            return im;
        }
        ImageU8 decoded;
        /* flipped: trap T11 */
        if (!decodeImageU8(bytes.data(), bytes.size(), decoded, &im.srcChannels, true)) {
            std::fprintf(stderr, "loadGLTF(): failed to decode image %zu of %s\n", i, path.c_str());
            return im;
        }
        /* stbi is asked for STBI_rgb_alpha: grey files are widened too */
        if (decoded.channels == 1) {
            ImageU8 wide;
            wide.width = decoded.width;
            wide.height = decoded.height;
            wide.channels = 4;
            wide.data.resize(decoded.data.size() * 4);
            for (size_t p = 0; p < decoded.data.size(); p++) {
                wide.data[4 * p] = wide.data[4 * p + 1] = wide.data[4 * p + 2] = decoded.data[p];
                wide.data[4 * p + 3] = 255;
            }
            decoded = std::move(wide);
        }
        im.rgba = std::make_shared<ImageU8>(std::move(decoded));
        im.ok = true;
        return im;
    }
    /* glTF textureInfo {index} -> decoded source image, or nullptr */
    DecodedImage *textureImage(const Value *info) {
        if (!info || !info->isObject() || !info->has("index")) return nullptr;
        const Value *texs = arrayOf("textures");
        size_t ti = (size_t)info->number("index", -1);
        if (!texs || ti >= texs->size()) return nullptr;
        const Value &t = (*texs)[ti];
        if (!t.has("source")) return nullptr;
        DecodedImage &im = image((size_t)t.number("source", -1));
        return im.ok ? &im : nullptr;
    }
    static ImportedTexture wholeImage(const std::string &name, const DecodedImage &im, ColorSpace cs) {
        ImportedTexture t;
        t.name = name;
        t.embedded = true;
        t.image = im.rgba;
        t.colorSpace = cs;
        return t;
    }
    /* assimpCreateImage(..., channel): one channel of the RGBA image as a 1-channel texture (AssimpLoadModel.cpp:236-249) */
    static ImportedTexture oneChannel(const std::string &name, const DecodedImage &im, int channel) {
        ImportedTexture t;
        t.name = name;
        t.embedded = true;
        t.colorSpace = ColorSpace::LINEAR;
        auto img = std::make_shared<ImageU8>();
        img->width = im.rgba->width;
        img->height = im.rgba->height;
        img->channels = 1;
        size_t n = (size_t)img->width * img->height;
        img->data.resize(n);
        for (size_t p = 0; p < n; p++) img->data[p] = im.rgba->data[4 * p + (size_t)channel];
        t.image = img;
        return t;
    }

    static vec4 vec4Of(const Value *v, vec4 def) {
        if (!v || !v->isArray() || v->size() < 3) return def;
        def.x = (*v)[0].getFloat();
        def.y = (*v)[1].getFloat();
        def.z = (*v)[2].getFloat();
        if (v->size() > 3) def.w = (*v)[3].getFloat();
        return def;
    }

    void loadMaterials(std::vector<ImportedMaterial> &out) {
        const Value *mats = arrayOf("materials");
        size_t n = mats ? mats->size() : 0;
        out.resize(n);
        for (size_t i = 0; i < n; i++) {
            const Value &m = (*mats)[i];
            ImportedMaterial &im = out[i];
            std::string name = m.string("name", "");
            /* assimp leaves unnamed materials without a name, which would alias all of them to "<stem>:" in the
             * material map; they are kept apart here */
            if (name.empty()) name = "material_" + std::to_string(i);
            im.info = AssetInfo(stem + ":" + name, path);
            im.info.embedded = true;
            im.type = ImportedMaterialType::PBR_STANDARD;
            const Value *pbr = m.find("pbrMetallicRoughness");
            static const Value emptyObject = Value::object();
            if (!pbr || !pbr->isObject()) pbr = &emptyObject;

            im.albedo = vec4Of(pbr->find("baseColorFactor"), vec4(1, 1, 1, 1));
            if (DecodedImage *base = textureImage(pbr->find("baseColorTexture"))) {
                im.albedoTexture = wholeImage(im.info.name + ":albedo", *base, ColorSpace::sRGB);
                if (base->srcChannels == 4) {
                    im.alphaTexture = oneChannel(im.info.name + ":alpha", *base, 3);
                    im.transparent = true;
                }
            }
            im.roughness = (float)pbr->number("roughnessFactor", 1.0);
            im.metallic = (float)pbr->number("metallicFactor", 1.0);
            if (DecodedImage *mr = textureImage(pbr->find("metallicRoughnessTexture"))) {
                im.roughnessTexture = oneChannel(im.info.name + ":roughness", *mr, 1);
                im.metallicTexture = oneChannel(im.info.name + ":metallic", *mr, 2);
            }
            if (DecodedImage *occ = textureImage(m.find("occlusionTexture"))) im.aoTexture = oneChannel(im.info.name + ":ao", *occ, 1);

            vec4 e = vec4Of(m.find("emissiveFactor"), vec4(0, 0, 0, 0));
            im.emissiveColor = vec3(e.x, e.y, e.z);
            im.emissiveStrength = 1.0f; /* AI_MATKEY_EMISSIVE_INTENSITY is absent unless the extension below is */
            if (const Value *ext = m.find("extensions"))
                if (const Value *es = ext->find("KHR_materials_emissive_strength")) im.emissiveStrength = (float)es->number("emissiveStrength", 1.0);
            if (DecodedImage *em = textureImage(m.find("emissiveTexture"))) {
                im.emissiveTexture = wholeImage(im.info.name + ":emissive", *em, ColorSpace::sRGB);
                const float eps = 1e-6f; /* isBlack(): math/MathUtils */
                if (std::fabs(im.emissiveColor.x) <= eps && std::fabs(im.emissiveColor.y) <= eps && std::fabs(im.emissiveColor.z) <= eps)
                    im.emissiveColor = vec3(1.0f);
                if (im.emissiveStrength == 0.0f) im.emissiveStrength = 1.0f;
            }
            if (DecodedImage *nm = textureImage(m.find("normalTexture"))) im.normalTexture = wholeImage(im.info.name + ":normal", *nm, ColorSpace::LINEAR);
        }
    }

    /* ---- meshes */
    bool loadPrimitive(const Value &prim, const std::string &name, ImportedModelNode &node) {
        int mode = (int)prim.number("mode", 4);
        if (mode != 4 && mode != 5 && mode != 6) return true; /* points and lines are not renderable here */
        const Value *attrs = prim.find("attributes");
        if (!attrs || !attrs->isObject() || !attrs->has("POSITION")) return true;
        std::vector<double> pos, nrm, uv, tan;
        int w;
        size_t nv, n2;
        if (!readAccessor((int)attrs->number("POSITION", -1), pos, w, nv)) return false;
        if (w != 3) return fail("glTF POSITION accessor is not VEC3");
        bool hasNormals = false, hasUVs = false, hasTangents = false;
        if (attrs->has("NORMAL")) {
            if (!readAccessor((int)attrs->number("NORMAL", -1), nrm, w, n2)) return false;
            hasNormals = w == 3 && n2 == nv;
        }
        if (attrs->has("TEXCOORD_0")) {
            if (!readAccessor((int)attrs->number("TEXCOORD_0", -1), uv, w, n2)) return false;
            hasUVs = w == 2 && n2 == nv;
        }
        if (attrs->has("TANGENT") && hasNormals) {
            if (!readAccessor((int)attrs->number("TANGENT", -1), tan, w, n2)) return false;
            hasTangents = w == 4 && n2 == nv;
        }
        auto mesh = std::make_unique<Mesh>();
        mesh->name = name;
        mesh->vertices.resize(nv);
        for (size_t i = 0; i < nv; i++) {
            Vertex &v = mesh->vertices[i];
            std::memset(&v, 0, sizeof(v));
            for (int c = 0; c < 3; c++) v.position[c] = (float)pos[i * 3 + (size_t)c];
            if (hasNormals) {
                vec3 n = vm::normalize(vec3((float)nrm[i * 3], (float)nrm[i * 3 + 1], (float)nrm[i * 3 + 2]));
                v.normal[0] = n.x; v.normal[1] = n.y; v.normal[2] = n.z;
            }
            if (hasUVs) {
                v.uv[0] = (float)uv[i * 2];
                v.uv[1] = 1.0f - (float)uv[i * 2 + 1];
            }
            v.color[0] = v.color[1] = v.color[2] = 1.0f;
        }
        /* indices -> triangle list */
        std::vector<uint32_t> idx;
        if (prim.has("indices")) {
            std::vector<double> raw;
            size_t ni;
            if (!readAccessor((int)prim.number("indices", -1), raw, w, ni)) return false;
            idx.resize(ni);
            for (size_t i = 0; i < ni; i++) {
                if (raw[i] < 0 || raw[i] >= (double)nv) return fail("glTF index out of range");
                idx[i] = (uint32_t)raw[i];
            }
        } else {
            idx.resize(nv);
            for (size_t i = 0; i < nv; i++) idx[i] = (uint32_t)i;
        }
        if (mode == 4) {
            idx.resize(idx.size() / 3 * 3);
            mesh->indices = std::move(idx);
        } else if (mode == 5) { /* strip: winding alternates */
            for (size_t i = 0; i + 2 < idx.size(); i++) {
                if (i & 1) mesh->indices.insert(mesh->indices.end(), {idx[i + 1], idx[i], idx[i + 2]});
                else mesh->indices.insert(mesh->indices.end(), {idx[i], idx[i + 1], idx[i + 2]});
            }
        } else { /* fan */
            for (size_t i = 1; i + 1 < idx.size(); i++) mesh->indices.insert(mesh->indices.end(), {idx[0], idx[i], idx[i + 1]});
        }
        if (mesh->indices.empty()) return true;
        if (!hasNormals) computeNormals(*mesh);
        if (hasTangents) {
            for (size_t i = 0; i < nv; i++) {
                Vertex &v = mesh->vertices[i];
                vec3 n(v.normal[0], v.normal[1], v.normal[2]);
                vec3 t = vm::normalize(vec3((float)tan[i * 4], (float)tan[i * 4 + 1], (float)tan[i * 4 + 2]));
                vec3 b = vm::normalize(vm::cross(n, t) * (float)tan[i * 4 + 3]);
                bool bad = !(std::isfinite(t.x) && std::isfinite(t.y) && std::isfinite(t.z) && std::isfinite(b.x) && std::isfinite(b.y) && std::isfinite(b.z));
                if (bad) {
                    /* sanitisation path of the reference (AssimpLoadModel.cpp:105-124) */
                    vec3 t1 = vm::cross(n, vec3(0, 0, 1)), t2 = vm::cross(n, vec3(1, 0, 0));
                    t = vm::normalize(vm::length(t1) > vm::length(t2) ? t1 : t2);
                    b = vm::normalize(vm::cross(n, t));
                }
                v.tangent[0] = t.x; v.tangent[1] = t.y; v.tangent[2] = t.z;
                v.bitangent[0] = b.x; v.bitangent[1] = b.y; v.bitangent[2] = b.z;
            }
        } else {
            computeTangents(*mesh);
        }
        node.meshes.push_back(std::move(mesh));
        node.materialIndices.push_back(prim.has("material") ? (int32_t)prim.number("material", -1) : -1);
        return true;
    }

    bool loadNode(size_t index, ImportedModelNode &out, int depth) {
        const Value *nodes = arrayOf("nodes");
        if (!nodes || index >= nodes->size()) return fail("glTF node index out of range");
        if (depth > 256) return fail("glTF node hierarchy too deep (cycle?)");
        const Value &n = (*nodes)[index];
        out.name = n.string("name", "");
        if (out.name.empty()) out.name = "nodes[" + std::to_string(index) + "]"; /* assimp names unnamed nodes by id */
        mat4 m(1.0f);
        const Value *mat = n.find("matrix");
        if (mat && mat->isArray() && mat->size() == 16) {
            for (int c = 0; c < 4; c++)
                for (int r = 0; r < 4; r++) m[c][r] = (*mat)[(size_t)(c * 4 + r)].getFloat();
        } else {
            vec3 t(0, 0, 0), s(1, 1, 1);
            quat q;
            if (const Value *v = n.find("translation")) t = vec3((*v)[0].getFloat(), (*v)[1].getFloat(), (*v)[2].getFloat());
            if (const Value *v = n.find("scale")) s = vec3((*v)[0].getFloat(), (*v)[1].getFloat(), (*v)[2].getFloat());
            if (const Value *v = n.find("rotation")) q = quat((*v)[3].getFloat(), (*v)[0].getFloat(), (*v)[1].getFloat(), (*v)[2].getFloat());
            m = vm::translate(t) * vm::toMat4(q) * vm::scale(s);
        }
        out.transform = decomposeTransform(m);
        if (n.has("mesh")) {
            const Value *meshes = arrayOf("meshes");
            size_t mi = (size_t)n.number("mesh", -1);
            if (!meshes || mi >= meshes->size()) return fail("glTF mesh index out of range");
            const Value &mesh = (*meshes)[mi];
            const Value *prims = mesh.find("primitives");
            std::string base = mesh.string("name", "");
            if (base.empty()) base = "meshes[" + std::to_string(mi) + "]";
            size_t np = prims && prims->isArray() ? prims->size() : 0;
            for (size_t p = 0; p < np; p++)
                if (!loadPrimitive((*prims)[p], np > 1 ? base + "-" + std::to_string(p) : base, out)) return false;
        }
        const Value *children = n.find("children");
        if (children && children->isArray()) {
            out.children.resize(children->size());
            for (size_t c = 0; c < children->size(); c++)
                if (!loadNode((size_t)(*children)[c].getDouble(), out.children[c], depth + 1)) return false;
        }
        return true;
    }

    bool loadScene(ImportedModelNode &root) {
        const Value *scenes = arrayOf("scenes");
        std::vector<size_t> roots;
        if (scenes && scenes->size() > 0) {
            size_t si = (size_t)doc.number("scene", 0);
            if (si >= scenes->size()) si = 0;
            const Value *ns = (*scenes)[si].find("nodes");
            if (ns && ns->isArray())
                for (const Value &v : ns->items()) roots.push_back((size_t)v.getDouble());
        } else if (count("nodes") > 0) {
            roots.push_back(0);
        }
        if (roots.empty()) return fail("glTF file has no scene nodes");
        if (roots.size() == 1) return loadNode(roots[0], root, 0);
        root.name = "ROOT";
        root.children.resize(roots.size());
        for (size_t i = 0; i < roots.size(); i++)
            if (!loadNode(roots[i], root.children[i], 1)) return false;
        return true;
    }
};

}  // namespace

bool loadGLTF(const std::string &path, ImportedModelNode &root, std::vector<ImportedMaterial> *materials, std::string *err) {
    Loader L;
    L.path = path;
    L.folder = dirOfPath(path);
    L.stem = stemOfPath(path);
    bool ok = false;
    try {
        ok = L.open() && L.loadScene(root);
        if (ok && materials) L.loadMaterials(*materials);
    } catch (std::exception &e) {
        L.error = std::string("malformed glTF: ") + e.what();
        ok = false;
    }
    if (!ok && err) *err = L.error.empty() ? "failed to load " + path : L.error + " (" + path + ")";
    return ok;
}

}  // namespace vengine
