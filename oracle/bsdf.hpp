/*
 * ORACLE — test infrastructure only (see omath.hpp).
 *
 * Pure functions of the reference estimator, restated from the GLSL includes. All paths are under
 * /root/reference/src/lib/vengine/shaders/.
 */
#pragma once
#include "omath.hpp"

namespace orc {

/* include/constants.glsl:1-9 */
static const float EPSILON = 0.000001f;
static const float PI = 3.14159265358979323846f;
static const float INV_PI = 0.31830988618379067154f;
static const float INV_4PI = 0.079577472f;
static const float PI_OVER_TWO = 1.57079632679489661923f;
static const float PI_OVER_FOUR = 0.78539816339744830961f;

/* include/utils.glsl:22-30 */
inline bool isBlack(vec3 c) { return c.x == 0 && c.y == 0 && c.z == 0; }
inline bool isBlack(vec3 c, float eps) { return std::fabs(c.x) <= eps && std::fabs(c.y) <= eps && std::fabs(c.z) <= eps; }

/* include/rng/rng_def.glsl:4-36 */
inline float uintToFloat(uint32_t x) {
    uint32_t bits = 0x3f800000u | (x >> 9);
    float f;
    std::memcpy(&f, &bits, 4);
    return f - 1.0f;
}
inline uint32_t xorshift(uint32_t &s) {
    s ^= s << 13;
    s ^= s >> 17;
    s ^= s << 5;
    return s;
}
inline uint32_t jenkinsHash(uint32_t x) {
    x += x << 10;
    x ^= x >> 6;
    x += x << 3;
    x ^= x >> 11;
    x += x << 15;
    return x;
}
/* rng_def.glsl:32-36 with `frame` = global sample index (batch * batchSize + s): one stream per
 * (pixel, sample) instead of one per (pixel, batch) so samples of a batch can run concurrently.
 * pixelIdx keeps the reference's x * width + y form (trap T9). */
inline uint32_t initRNG(uint32_t px, uint32_t py, uint32_t resx, uint32_t frame) {
    uint32_t pixelIdx = px * resx + py;
    uint32_t s = pixelIdx ^ jenkinsHash(frame);
    return jenkinsHash(s);
}
/* Low-discrepancy mode (PTC_FLAG_SAMPLER_SOBOL): shuffled + Owen-scrambled Sobol points (Burley 2020), the role the
 * reference gives to its optional PMJ02BN sampler (include/rng/rng_pmj.glsl:66-107: rand1D / rand2D keyed by pixel,
 * sample index and a running dimension).  Restated integer for integer like vviewer_b200/csrc/bsdf.cuh. */
inline uint32_t hashCombine(uint32_t seed, uint32_t v) { return seed ^ (v + (seed << 6) + (seed >> 2)); }
inline uint32_t reverseBits32(uint32_t x) {
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8);
    return (x >> 16) | (x << 16);
}
inline uint32_t laineKarras(uint32_t x, uint32_t seed) {
    x += seed;
    x ^= x * 0x6c50b47cu;
    x ^= x * 0xb82f1e52u;
    x ^= x * 0xc7afe638u;
    x ^= x * 0x8d22f6e6u;
    return x;
}
inline uint32_t owenScramble(uint32_t x, uint32_t seed) { return reverseBits32(laineKarras(reverseBits32(x), seed)); }
inline uint32_t sobolDim1(uint32_t i) {
    uint32_t v = 0x80000000u, r = 0;
    for (; i; i >>= 1) {
        if (i & 1u) r ^= v;
        v ^= v >> 1;
    }
    return r;
}
/* ---- the reference's optional PMJ02BN sampler, include/rng/rng_pmj.glsl:20-107 (pbrt-v4 style) */
struct PmjTables { /* the two storage buffers of vulkan/resources/VulkanRandom.cpp:40-72 */
    const float *pmj = nullptr;  /* [PMJ_N_SEQUENCES][PMJ_N_SAMPLES][2] */
    const float *blue = nullptr; /* [BLUE_NOISE_TEXTURES][BLUE_NOISE_RESOLUTION][BLUE_NOISE_RESOLUTION] */
};
static constexpr uint32_t PMJ_N_SEQUENCES = 16, PMJ_N_SAMPLES = 16384, BLUE_NOISE_TEXTURES = 48, BLUE_NOISE_RESOLUTION = 128; /* rng_pmj_defines.glsl, bluenoise_defines.glsl */
static constexpr uint32_t PMJ_SEED = 2873468793u;
static constexpr float ONEMINUSEPSILON = 0.999999f; /* include/constants.glsl:2 */
inline uint64_t mixBits(uint64_t v) { /* rng_pmj.glsl:30-37 */
    v ^= (v >> 31);
    v *= 9202493588570546565ull;
    v ^= (v >> 27);
    v *= 9357036318526133325ull;
    v ^= (v >> 33);
    return v;
}
inline uint32_t permutationElement(uint32_t i, uint32_t l, uint32_t p) { /* rng_pmj.glsl:39-69 */
    uint32_t w = l - 1;
    w |= w >> 1;
    w |= w >> 2;
    w |= w >> 4;
    w |= w >> 8;
    w |= w >> 16;
    do {
        i ^= p;
        i *= 0xe170893du;
        i ^= p >> 16;
        i ^= (i & w) >> 4;
        i ^= p >> 8;
        i *= 0x0929eb3fu;
        i ^= p >> 23;
        i ^= (i & w) >> 1;
        i *= 1u | p >> 27;
        i *= 0x6935fa69u;
        i ^= (i & w) >> 11;
        i *= 0x74dcb303u;
        i ^= (i & w) >> 2;
        i *= 0x9e501cc3u;
        i ^= (i & w) >> 2;
        i *= 0xc860a3dfu;
        i &= w;
        i ^= i >> 5;
    } while (i >= l);
    return (i + p) % l;
}
inline uint32_t pmjHash(uint32_t px, uint32_t py, uint32_t dimension) { /* rng_pmj.glsl:73-74, 92-94 */
    return (uint32_t)mixBits(((uint64_t)px << 48) ^ ((uint64_t)py << 32) ^ ((uint64_t)dimension << 16) ^ (uint64_t)PMJ_SEED);
}
inline float blueNoise(const PmjTables &t, uint32_t texture, uint32_t px, uint32_t py) { /* include/rng/bluenoise.glsl:1-8: [texture][x][y] */
    return t.blue[((size_t)(texture % BLUE_NOISE_TEXTURES) * BLUE_NOISE_RESOLUTION + px % BLUE_NOISE_RESOLUTION) * BLUE_NOISE_RESOLUTION + py % BLUE_NOISE_RESOLUTION];
}

struct Rng {
    uint32_t state;         /* xorshift state, or the next dimension */
    uint32_t pixelSeed = 0; /* low discrepancy only */
    uint32_t index = 0;     /* global sample index */
    bool ld = false;
    /* PMJ02BN (PTC_FLAG_SAMPLER_PMJ): state = dimension, index = sampleIndex */
    bool pmj = false;
    uint32_t px = 0, py = 0, spp = 1;
    PmjTables tables;
    void init(uint32_t px_, uint32_t py_, uint32_t resx, uint32_t sampleIndex, bool lowDiscrepancy) {
        ld = lowDiscrepancy;
        index = sampleIndex;
        pixelSeed = jenkinsHash(px_ * resx + py_);
        state = ld ? 0u : initRNG(px_, py_, resx, sampleIndex);
    }
    /* raygen.rgen.glsl:30-33, 41-43, 57-61: per sample the dimension restarts at pixel.y * width + pixel.y (sic) */
    void initPmj(uint32_t px_, uint32_t py_, uint32_t resx, uint32_t sampleIndex, uint32_t samplesPerPixel, const PmjTables &t) {
        pmj = true;
        ld = false;
        px = px_;
        py = py_;
        spp = samplesPerPixel;
        tables = t;
        index = sampleIndex;
        state = py_ * resx + py_;
    }
    float rand1D() {
        if (pmj) { /* rng_pmj.glsl:71-83 */
            const uint32_t idx = permutationElement(index, spp, pmjHash(px, py, state));
            const float delta = blueNoise(tables, state, px, py);
            state++;
            return std::min(((float)idx + delta) / (float)spp, ONEMINUSEPSILON);
        }
        if (ld) {
            uint32_t seed = jenkinsHash(hashCombine(pixelSeed, state));
            state += 1;
            uint32_t idx = owenScramble(index, seed);
            return uintToFloat(owenScramble(reverseBits32(idx), hashCombine(seed, 1u)));
        }
        return uintToFloat(xorshift(state));
    }
    vec2 rand2D() {
        if (pmj) { /* rng_pmj.glsl:85-107 (BLUE_NOISE_2D is not defined) */
            uint32_t idx = index;
            const uint32_t inst = state / 2u;
            if (inst >= PMJ_N_SEQUENCES) idx = permutationElement(index, spp, pmjHash(px, py, state));
            const float *u = tables.pmj + ((size_t)(inst % PMJ_N_SEQUENCES) * PMJ_N_SAMPLES + idx % PMJ_N_SAMPLES) * 2;
            state += 2u;
            return {std::min(u[0], ONEMINUSEPSILON), std::min(u[1], ONEMINUSEPSILON)};
        }
        if (ld) {
            uint32_t seed = jenkinsHash(hashCombine(pixelSeed, state));
            state += 2;
            uint32_t idx = owenScramble(index, seed);
            uint32_t x = owenScramble(reverseBits32(idx), hashCombine(seed, 1u));
            uint32_t y = owenScramble(sobolDim1(idx), hashCombine(seed, 2u));
            return {uintToFloat(x), uintToFloat(y)};
        }
        float a = rand1D();
        float b = rand1D();
        return {a, b};
    }
};

/* include/sampling.glsl:3-35 */
inline vec3 cosineSampleHemisphere(vec2 r, float &pdf) {
    float rx = 2.0f * r.x - 1.0f;
    float ry = 2.0f * r.y - 1.0f;
    vec3 dir(0, 0, 0);
    if (rx == 0 && ry == 0) {
        dir.x = 0;
        dir.z = 0;
    } else if (std::fabs(rx) > std::fabs(ry)) {
        float rr = rx;
        float phi = PI_OVER_FOUR * (ry / rx);
        dir.x = rr * std::cos(phi);
        dir.z = rr * std::sin(phi);
    } else {
        float rr = ry;
        float phi = PI_OVER_TWO - PI_OVER_FOUR * (rx / ry);
        dir.x = rr * std::cos(phi);
        dir.z = rr * std::sin(phi);
    }
    dir.y = std::sqrt(std::max(0.0f, 1.0f - dir.x * dir.x - dir.z * dir.z));
    pdf = INV_PI * dir.y;
    pdf = std::max(pdf, EPSILON);
    return dir;
}
/* sampling.glsl:37-40 */
inline float cosineSampleHemispherePdf(float cosTheta) { return cosTheta * INV_PI; }
/* sampling.glsl:42-46 */
inline vec2 uniformSampleTriangle(vec2 r) {
    float a = std::sqrt(1.0f - r.x);
    return {1.0f - a, a * r.y};
}
/* sampling.glsl:48-67 */
inline vec2 concentricSampleDisk(vec2 r) {
    float ox = 2.0f * r.x - 1.0f, oy = 2.0f * r.y - 1.0f;
    if (ox == 0 && oy == 0) return {0, 0};
    float theta, rr;
    if (std::fabs(ox) > std::fabs(oy)) {
        rr = ox;
        theta = PI_OVER_FOUR * (oy / ox);
    } else {
        rr = oy;
        theta = PI_OVER_TWO - PI_OVER_FOUR * (ox / oy);
    }
    return {rr * std::cos(theta), rr * std::sin(theta)};
}

/* pt/MIS.glsl:5-10.  One guard the shader lacks: a density above sqrt(FLT_MAX) = 1.8e19 (a mesh light seen exactly edge-on:
 * d^2 / cos -> inf, lightSampling.glsl:92) squares to inf and the shader's inf / inf is NaN, which poisons the pixel for good
 * (seen once in 134 M samples of the MeshLight recipe).  The limit of the expression is returned instead; finite cases are
 * the shader's arithmetic unchanged.  The device code carries the same guard. */
inline float PowerHeuristic(int nf, float fPdf, int ng, float gPdf) {
    float f = nf * fPdf;
    float g = ng * gPdf;
    float f2 = f * f, g2 = g * g;
    if (std::isinf(f2)) return std::isinf(g2) ? 0.5f : 1.0f;
    return f2 / (f2 + g2);
}

/* include/brdfs/common.glsl:3-7 */
inline float schlickWeight(float cosTheta) {
    float m = clampf(1.0f - cosTheta, 0.0f, 1.0f);
    return (m * m) * (m * m) * m;
}
/* common.glsl:19-24 */
inline float GTR2(float NdotH, float a) {
    float a2 = a * a;
    float t = 1.0f + (a2 - 1.0f) * NdotH * NdotH;
    return a2 / (PI * t * t);
}
/* common.glsl:45-49 */
inline float smithG_GGX(float NdotV, float alphaG) {
    float a = alphaG * alphaG;
    float b = NdotV * NdotV;
    return 1.0f / (std::fabs(NdotV) + std::max(std::sqrt(a + b - a * b), EPSILON));
}

struct PBRStandard {
    vec3 albedo;
    float metallic;
    float roughness;
};

/* include/brdfs/pbrStandard.glsl:10-19 */
inline vec3 evalDisneyDiffuse(float NdotL, float NdotV, float LdotH, const PBRStandard &pbr) {
    float FL = schlickWeight(NdotL);
    float FV = schlickWeight(NdotV);
    float Fd90 = 0.5f + 2.0f * LdotH * LdotH * pbr.roughness;
    float Fd = mixf(1.0f, Fd90, FL) * mixf(1.0f, Fd90, FV);
    return pbr.albedo * ((1.0f / PI) * Fd);
}
/* pbrStandard.glsl:25-44 */
inline vec3 evalDisneyMicrofacetIsotropic(float NdotL, float NdotV, float NdotH, float LdotH, const PBRStandard &pbr) {
    float Cdlum = 0.3f * pbr.albedo.x + 0.6f * pbr.albedo.y + 0.1f * pbr.albedo.z;
    vec3 Ctint = Cdlum > 0.0f ? pbr.albedo / Cdlum : vec3(1.0f);
    float specular = 0.5f;
    float specularTint = 0.0f;
    vec3 Cspec0 = vmix(vmix(vec3(1.0f), Ctint, specularTint) * (specular * 0.08f), pbr.albedo, pbr.metallic);
    float a = std::max(0.001f, pbr.roughness * pbr.roughness);
    float Ds = GTR2(NdotH, a);
    float FH = schlickWeight(LdotH);
    vec3 Fs = vmix(Cspec0, vec3(1.0f), FH);
    float Gs = smithG_GGX(NdotL, a);
    Gs *= smithG_GGX(NdotV, a);
    return Fs * (Gs * Ds);
}
/* pbrStandard.glsl:46-63 */
inline float pdfDisneyMicrofacetIsotropic(vec3 wi, vec3 wo, const PBRStandard &pbr) {
    if (!(wo.y > 0)) return 0.0f;
    if (!(wi.y > 0)) return 0.0f;
    vec3 wh = normalize(wo + wi);
    float NdotH = std::max(wh.y, EPSILON);
    float alpha2 = pbr.roughness * pbr.roughness;
    alpha2 *= alpha2;
    float cos2Theta = NdotH * NdotH;
    float denom = cos2Theta * (alpha2 - 1.0f) + 1.0f;
    if (denom == 0.0f) return 0.0f;
    float pdfDistribution = alpha2 * NdotH / (PI * denom * denom);
    return pdfDistribution / (4.0f * dot(wo, wh));
}
/* pbrStandard.glsl:65-81 */
inline void sampleDisneyMicrofacetIsotropic(vec3 &wi, vec3 wo, float &pdf, vec2 u, const PBRStandard &pbr) {
    float phi = (2.0f * PI) * u.y;
    float alpha = pbr.roughness * pbr.roughness;
    float tanTheta2 = alpha * alpha * u.x / (1.0f - u.x);
    float cosTheta = 1.0f / std::sqrt(1.0f + tanTheta2);
    float sinTheta = std::sqrt(std::max(EPSILON, 1.0f - cosTheta * cosTheta));
    vec3 wh(sinTheta * std::cos(phi), cosTheta, sinTheta * std::sin(phi));
    if (!(wh.y > 0)) wh = wh * -1.0f;
    wi = reflect(-wo, wh);
    pdf = pdfDisneyMicrofacetIsotropic(wi, wo, pbr);
}
/* pbrStandard.glsl:83-90 */
inline float getDiffuseSamplingRatio(const PBRStandard &pbr) {
    float d = std::max(1.0f - pbr.metallic, 0.1f);
    float g = std::max(1.0f - pbr.roughness, 0.1f);
    return d / (d + g);
}
/* pbrStandard.glsl:92-105 */
inline vec3 evalPBRStandard(const PBRStandard &pbr, vec3 wi, vec3 wo, vec3 H) {
    float NdotL = wi.y;
    float NdotV = wo.y;
    if (NdotL < 0.0f || NdotV < 0.0f) return vec3(0.0f);
    float NdotH = H.y;
    float LdotH = dot(wi, H);
    vec3 diffuse = evalDisneyDiffuse(NdotL, NdotV, LdotH, pbr);
    vec3 glossy = evalDisneyMicrofacetIsotropic(NdotL, NdotV, NdotH, LdotH, pbr);
    return (diffuse * (1.0f - pbr.metallic) + glossy) * NdotL;
}
/* pbrStandard.glsl:123-137 */
inline float pdfPBRStandard(vec3 wi, vec3 wo, const PBRStandard &pbr) {
    float cosTheta = wi.y;
    if (cosTheta < 0) return 0.0f;
    float pdfDiffuse = cosineSampleHemispherePdf(cosTheta);
    float pdfMicrofacet = pdfDisneyMicrofacetIsotropic(wi, wo, pbr);
    float r = getDiffuseSamplingRatio(pbr);
    return pdfDiffuse * r + pdfMicrofacet * (1.0f - r);
}
/* pbrStandard.glsl:139-165 */
inline vec3 samplePBRStandard(vec3 &wi, vec3 wo, float &pdf, const PBRStandard &pbr, vec2 u, float rnd) {
    pdf = 0.0f;
    wi = vec3(0.0f);
    float r = getDiffuseSamplingRatio(pbr);
    if (rnd <= r) {
        wi = cosineSampleHemisphere(u, pdf);
    } else {
        sampleDisneyMicrofacetIsotropic(wi, wo, pdf, u, pbr);
    }
    vec3 F = evalPBRStandard(pbr, wi, wo, normalize(wi + wo));
    pdf = pdfPBRStandard(wi, wo, pbr);
    if (pdf < EPSILON) return vec3(0.0f);
    return F;
}

/* include/phaseFunctions.glsl:1-8 */
inline float HenyeyGreenstein(float cosTheta, float g) {
    float denom = 1.0f + g * g + 2.0f * g * cosTheta;
    return INV_4PI * (1.0f - g * g) / (denom * std::sqrt(denom));
}
inline float HG_p(vec3 wo, vec3 wi, float g) { return HenyeyGreenstein(dot(wo, wi), g); }

/* include/frame.glsl:63-74 */
inline void createCoordinateSystem(vec3 v1, vec3 &v2, vec3 &v3) {
    if (std::fabs(v1.x) > std::fabs(v1.y)) {
        v2 = vec3(-v1.z, 0, v1.x) / std::sqrt(v1.x * v1.x + v1.z * v1.z);
    } else {
        v2 = vec3(0, v1.z, -v1.y) / std::sqrt(v1.y * v1.y + v1.z * v1.z);
    }
    v3 = cross(v1, v2);
}
/* phaseFunctions.glsl:10-32 */
inline float HG_Sample(vec3 wo, vec3 &wi, vec2 r, float g) {
    float cosTheta;
    if (std::fabs(g) < 1e-3f) {
        cosTheta = 1.0f - 2.0f * r.x;
    } else {
        float sqrTerm = (1.0f - g * g) / (1.0f + g - 2.0f * g * r.x);
        cosTheta = -(1.0f + g * g - sqrTerm * sqrTerm) / (2.0f * g);
    }
    float sinTheta = std::sqrt(std::max(0.0f, 1.0f - cosTheta * cosTheta));
    float phi = 2.0f * PI * r.y;
    vec3 v1, v2;
    createCoordinateSystem(wo, v1, v2);
    wi = v1 * (sinTheta * std::cos(phi)) + v2 * (sinTheta * std::sin(phi)) + wo * cosTheta;
    return HenyeyGreenstein(cosTheta, g);
}

/* include/frame.glsl:1-58 */
struct Frame {
    vec3 normal, tangent, bitangent;
};
inline bool fixFrame(vec3 &normal, vec3 &tangent, vec3 &bitangent, vec3 ray) {
    bool flipped = false;
    if (dot(normal, ray) > 0) {
        normal = -normal;
        tangent = -tangent;
        flipped = true;
    }
    normal = normalize(normal);
    tangent = normalize(tangent);
    bitangent = normalize(bitangent);
    if (std::fabs(dot(normal, tangent)) > 0.999f) {
        bitangent = std::fabs(normal.z) < 0.999f ? vec3(0, 0, 1) : vec3(1, 0, 0);
        tangent = cross(bitangent, normal);
        bitangent = cross(normal, tangent);
    } else {
        bitangent = cross(normal, tangent);
        tangent = cross(bitangent, normal);
    }
    return flipped;
}
inline vec3 localToWorld(const Frame &f, vec3 v) { return f.tangent * v.x + f.normal * v.y + f.bitangent * v.z; }
inline vec3 worldToLocal(const Frame &f, vec3 v) { return {dot(v, f.tangent), dot(v, f.normal), dot(v, f.bitangent)}; }
inline void applyNormalToFrame(Frame &f, vec3 newNormal) {
    vec3 nw = localToWorld(f, newNormal);
    f.normal = nw;
    f.tangent = cross(f.normal, f.bitangent);
    f.bitangent = cross(f.tangent, f.normal);
}
/* include/utils.glsl:10-14 */
inline vec3 processNormalFromNormalMap(vec3 n) {
    vec3 N = n * 2.0f - vec3(1.0f);
    return normalize(vec3(N.x, N.z, -N.y));
}

}  // namespace orc
