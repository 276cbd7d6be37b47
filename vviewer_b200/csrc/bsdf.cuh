/*
 * Device functions of the estimator: RNG, samplers, MIS, Disney/GGX "PBR standard" BRDF, Lambert,
 * Henyey-Greenstein phase function, shading frame.  Behaviour follows the reference's GLSL includes
 * (paths under src/lib/vengine/shaders/): include/rng/rng_def.glsl, include/sampling.glsl, pt/MIS.glsl,
 * include/brdfs/common.glsl, include/brdfs/pbrStandard.glsl, include/phaseFunctions.glsl,
 * include/frame.glsl, include/utils.glsl.  Local shading frame: y = normal.
 */
#pragma once
#include "common.cuh"

#define PT_EPSILON 0.000001f
#define PT_PI 3.14159265358979323846f
#define PT_INV_PI 0.31830988618379067154f
#define PT_INV_4PI 0.079577472f
#define PT_PI_2 1.57079632679489661923f
#define PT_PI_4 0.78539816339744830961f

/* ------------------------------------------------------------------ RNG (rng_def.glsl:4-36) */
PTC_HD uint32_t jenkins(uint32_t x) {
    x += x << 10;
    x ^= x >> 6;
    x += x << 3;
    x ^= x >> 11;
    x += x << 15;
    return x;
}
/* one stream per (pixel, global sample index); pixel index = x * width + y as in rng_def.glsl:33 */
PTC_HD uint32_t rngSeed(uint32_t px, uint32_t py, uint32_t width, uint32_t sampleIndex) {
    return jenkins((px * width + py) ^ jenkins(sampleIndex));
}
/* Sampler state of one path.  Two modes:
 *   default            the reference's default generator (rng_def.glsl): xorshift32 on `s`;
 *   low discrepancy    (PTC_FLAG_SAMPLER_SOBOL; the reference's optional PMJ02BN sampler, rng_pmj.glsl:66-107, plays this
 *                      role, its tables are not reproduced) shuffled + Owen-scrambled Sobol points (Burley 2020, "Practical
 *                      hash-based Owen scrambling"): `s` counts dimensions, the global sample index is shuffled per
 *                      (pixel, dimension) and the first two Sobol dimensions are scrambled per (pixel, dimension), so that
 *                      the samples of one pixel are a scrambled (0, m, 2)-net in every dimension pair.
 * rnd() = rand1D, rnd2() = rand2D of the reference; both modes draw in the same program order. */
struct Rng {
    uint32_t s;         /* xorshift state, or the next dimension */
    uint32_t pixelSeed; /* low discrepancy only; PMJ02BN: pixel.x | pixel.y << 16 */
    uint32_t index;     /* global sample index of the path (batch * batchSize + s) */
    uint32_t ld;        /* 0 xorshift stream, 1 Owen-scrambled Sobol, 2 PMJ02BN (rng_pmj.glsl) */
};

/* ---- the reference's optional PMJ02BN sampler (include/rng/rng_pmj.glsl:20-107).  The two tables are the reference's storage
 * buffers (vulkan/resources/VulkanRandom.cpp:40-72), handed over through ptc_set_sampler_tables; one render uses one
 * samplesPerPixel (raygen.rgen.glsl:30-33), so tables + count sit in constant memory for the kernels of that render. */
struct PmjConst {
    const float *pmj;  /* [16][16384][2] */
    const float *blue; /* [48][128][128] */
    uint32_t spp;
};
__constant__ PmjConst g_pmj;
#define PMJ_N_SEQUENCES 16u
#define PMJ_N_SAMPLES 16384u
#define BLUE_NOISE_TEXTURES 48u
#define BLUE_NOISE_RESOLUTION 128u
#define PMJ_SEED 2873468793u
#define ONEMINUSEPSILON 0.999999f
PTC_D uint64_t mixBits(uint64_t v) { /* rng_pmj.glsl:30-37 */
    v ^= (v >> 31);
    v *= 9202493588570546565ull;
    v ^= (v >> 27);
    v *= 9357036318526133325ull;
    v ^= (v >> 33);
    return v;
}
PTC_D uint32_t permutationElement(uint32_t i, uint32_t l, uint32_t p) { /* rng_pmj.glsl:39-69 */
    uint32_t w = l - 1u;
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
PTC_D uint32_t pmjHash(uint32_t px, uint32_t py, uint32_t dimension) { /* rng_pmj.glsl:73-74, 92-94 */
    return (uint32_t)mixBits(((uint64_t)px << 48) ^ ((uint64_t)py << 32) ^ ((uint64_t)dimension << 16) ^ (uint64_t)PMJ_SEED);
}
PTC_D float pmjRand1D(Rng &r) { /* rng_pmj.glsl:71-83 */
    const uint32_t px = r.pixelSeed & 0xffffu, py = r.pixelSeed >> 16;
    const uint32_t idx = permutationElement(r.index, g_pmj.spp, pmjHash(px, py, r.s));
    /* include/rng/bluenoise.glsl:1-8: data[texture][pixel.x][pixel.y] */
    const float delta = __ldg(g_pmj.blue + ((size_t)(r.s % BLUE_NOISE_TEXTURES) * BLUE_NOISE_RESOLUTION + px % BLUE_NOISE_RESOLUTION) * BLUE_NOISE_RESOLUTION + py % BLUE_NOISE_RESOLUTION);
    r.s += 1u;
    return fminf(__fdiv_rn(__fadd_rn((float)idx, delta), (float)g_pmj.spp), ONEMINUSEPSILON);
}
PTC_D float2 pmjRand2D(Rng &r) { /* rng_pmj.glsl:85-107 (BLUE_NOISE_2D is not defined) */
    uint32_t idx = r.index;
    const uint32_t inst = r.s / 2u;
    if (inst >= PMJ_N_SEQUENCES) idx = permutationElement(r.index, g_pmj.spp, pmjHash(r.pixelSeed & 0xffffu, r.pixelSeed >> 16, r.s));
    const float2 u = __ldg((const float2 *)g_pmj.pmj + (size_t)(inst % PMJ_N_SEQUENCES) * PMJ_N_SAMPLES + idx % PMJ_N_SAMPLES);
    r.s += 2u;
    return make_float2(fminf(u.x, ONEMINUSEPSILON), fminf(u.y, ONEMINUSEPSILON));
}
PTC_HD uint32_t hashCombine(uint32_t seed, uint32_t v) { return seed ^ (v + (seed << 6) + (seed >> 2)); }
PTC_HD uint32_t reverseBits32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8);
    return (x >> 16) | (x << 16);
#endif
}
PTC_HD uint32_t laineKarras(uint32_t x, uint32_t seed) { /* a hash whose bit k only depends on bits <= k */
    x += seed;
    x ^= x * 0x6c50b47cu;
    x ^= x * 0xb82f1e52u;
    x ^= x * 0xc7afe638u;
    x ^= x * 0x8d22f6e6u;
    return x;
}
PTC_HD uint32_t owenScramble(uint32_t x, uint32_t seed) { return reverseBits32(laineKarras(reverseBits32(x), seed)); }
PTC_HD uint32_t sobolDim1(uint32_t i) { /* second Sobol dimension: direction numbers v_k = v_(k-1) ^ (v_(k-1) >> 1) */
    uint32_t v = 0x80000000u, r = 0;
    for (; i; i >>= 1) {
        if (i & 1u) r ^= v;
        v ^= v >> 1;
    }
    return r;
}
PTC_HD float bitsToUnitFloat(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(0x3f800000u | (x >> 9)) - 1.0f;
#else
    uint32_t b = 0x3f800000u | (x >> 9);
    float f;
    memcpy(&f, &b, 4);
    return f - 1.0f;
#endif
}
/* sampler: 0 default xorshift stream, 1 Sobol, 2 PMJ02BN */
PTC_D uint32_t samplerOfFlags(uint32_t flags) { return (flags & PTC_FLAG_SAMPLER_PMJ) ? 2u : ((flags & PTC_FLAG_SAMPLER_SOBOL) ? 1u : 0u); }
PTC_D Rng rngInit(uint32_t px, uint32_t py, uint32_t width, uint32_t sampleIndex, uint32_t sampler) {
    Rng r;
    r.ld = sampler;
    r.index = sampleIndex;
    if (sampler == 2u) { /* raygen.rgen.glsl:57-61: every sample starts at dimension pixel.y * width + pixel.y (sic) */
        r.pixelSeed = (px & 0xffffu) | (py << 16);
        r.s = py * width + py;
        return r;
    }
    r.pixelSeed = jenkins(px * width + py);
    r.s = sampler ? 0u : rngSeed(px, py, width, sampleIndex);
    return r;
}
/* the per-path part of the state that is not stored in the slot (everything but `s`) is a function of (pixel, sample index) */
PTC_D void rngRestore(Rng &r, uint32_t sampler, uint32_t px, uint32_t py, uint32_t width, uint32_t sampleIndex) {
    r.ld = sampler;
    r.index = sampleIndex;
    r.pixelSeed = sampler == 2u ? ((px & 0xffffu) | (py << 16)) : (sampler == 1u ? jenkins(px * width + py) : 0u);
}
PTC_D float rnd(Rng &r) {
    if (r.ld == 2u) return pmjRand1D(r);
    if (r.ld) {
        const uint32_t seed = jenkins(hashCombine(r.pixelSeed, r.s));
        r.s += 1u;
        const uint32_t idx = owenScramble(r.index, seed);
        return bitsToUnitFloat(owenScramble(reverseBits32(idx), hashCombine(seed, 1u)));
    }
    r.s ^= r.s << 13;
    r.s ^= r.s >> 17;
    r.s ^= r.s << 5;
    return __uint_as_float(0x3f800000u | (r.s >> 9)) - 1.0f;
}
PTC_D float2 rnd2(Rng &r) {
    if (r.ld == 2u) return pmjRand2D(r);
    if (r.ld) {
        const uint32_t seed = jenkins(hashCombine(r.pixelSeed, r.s));
        r.s += 2u;
        const uint32_t idx = owenScramble(r.index, seed);
        const uint32_t x = owenScramble(reverseBits32(idx), hashCombine(seed, 1u));
        const uint32_t y = owenScramble(sobolDim1(idx), hashCombine(seed, 2u));
        return make_float2(bitsToUnitFloat(x), bitsToUnitFloat(y));
    }
    const float a = rnd(r);
    const float b = rnd(r);
    return make_float2(a, b);
}

/* ------------------------------------------------------------------ samplers (sampling.glsl) */
PTC_D float2 concentricDisk(float u0, float u1) { /* :48-67 */
    float ox = 2.0f * u0 - 1.0f, oy = 2.0f * u1 - 1.0f;
    if (ox == 0.0f && oy == 0.0f) return make_float2(0.0f, 0.0f);
    float r, th;
    if (fabsf(ox) > fabsf(oy)) {
        r = ox;
        th = PT_PI_4 * (oy / ox);
    } else {
        r = oy;
        th = PT_PI_2 - PT_PI_4 * (ox / oy);
    }
    float s, c;
    sincosf(th, &s, &c);
    return make_float2(r * c, r * s);
}
PTC_D float3 cosineHemisphere(float u0, float u1, float &pdf) { /* :3-35 */
    float2 d = concentricDisk(u0, u1);
    float y = sqrtf(fmaxf(0.0f, 1.0f - d.x * d.x - d.y * d.y));
    pdf = fmaxf(PT_INV_PI * y, PT_EPSILON);
    return f3(d.x, y, d.y);
}
PTC_D float2 sampleTriangle(float u0, float u1) { /* :42-46 */
    float a = sqrtf(1.0f - u0);
    return make_float2(1.0f - a, a * u1);
}
/* MIS.glsl:5-10 with nf = ng = 1.  One guard the shader lacks (same in oracle/bsdf.hpp): a density above sqrt(FLT_MAX) - a mesh light
 * seen exactly edge-on - squares to inf and inf / inf is NaN; the limit of the expression is returned instead. */
PTC_D float powerHeuristic(float fPdf, float gPdf) {
    float f2 = fPdf * fPdf, g2 = gPdf * gPdf;
    if (isinf(f2)) return isinf(g2) ? 0.5f : 1.0f;
    return f2 / (f2 + g2);
}

/* ------------------------------------------------------------------ PBR standard (pbrStandard.glsl, common.glsl) */
/* evalPBRStandard / pdfPBRStandard are written with explicit round-to-nearest intrinsics (never contracted into
 * FMAs) in the exact operation order of the oracle: GTR2's 1 + (a^2 - 1) NdotH^2 cancels catastrophically for
 * glossy lobes, so a single differently-rounded product upstream would already break the 1e-5 BSDF parity bar. */
struct Pbr {
    float3 albedo;
    float metallic, roughness;
};
PTC_D float rmul(float a, float b) { return __fmul_rn(a, b); }
PTC_D float radd(float a, float b) { return __fadd_rn(a, b); }
PTC_D float rsub(float a, float b) { return __fsub_rn(a, b); }
PTC_D float rdiv(float a, float b) { return __fdiv_rn(a, b); }
PTC_D float rdot(float3 a, float3 b) { return radd(radd(rmul(a.x, b.x), rmul(a.y, b.y)), rmul(a.z, b.z)); }
PTC_D float3 rnormalize(float3 a) {
    const float l = __fsqrt_rn(rdot(a, a));
    return f3(rdiv(a.x, l), rdiv(a.y, l), rdiv(a.z, l));
}
PTC_D float rmix(float a, float b, float t) { return radd(rmul(a, rsub(1.0f, t)), rmul(b, t)); }
PTC_D float schlickW(float c) {
    float m = clampf(rsub(1.0f, c), 0.0f, 1.0f);
    return rmul(rmul(rmul(m, m), rmul(m, m)), m);
}
PTC_D float gtr2(float NdotH, float a) {
    float a2 = rmul(a, a);
    float t = radd(1.0f, rmul(rmul(rsub(a2, 1.0f), NdotH), NdotH));
    return rdiv(a2, rmul(rmul(PT_PI, t), t));
}
PTC_D float smithGGX(float NdotV, float alphaG) {
    float a = rmul(alphaG, alphaG), b = rmul(NdotV, NdotV);
    return rdiv(1.0f, radd(fabsf(NdotV), fmaxf(__fsqrt_rn(rsub(radd(a, b), rmul(a, b))), PT_EPSILON)));
}
PTC_D float diffuseRatio(const Pbr &p) { /* :83-90 */
    float d = fmaxf(rsub(1.0f, p.metallic), 0.1f), g = fmaxf(rsub(1.0f, p.roughness), 0.1f);
    return rdiv(d, radd(d, g));
}
PTC_D float3 pbrEval(const Pbr &p, float3 wi, float3 wo) { /* evalPBRStandard :92-105 with H = normalize(wo + wi) */
    float NdotL = wi.y, NdotV = wo.y;
    if (NdotL < 0.0f || NdotV < 0.0f) return f3(0.0f);
    float3 H = rnormalize(f3(radd(wo.x, wi.x), radd(wo.y, wi.y), radd(wo.z, wi.z)));
    float NdotH = H.y, LdotH = rdot(wi, H);
    /* Disney diffuse :10-19 */
    float FL = schlickW(NdotL), FV = schlickW(NdotV);
    float Fd90 = radd(0.5f, rmul(rmul(rmul(2.0f, LdotH), LdotH), p.roughness));
    float Fd = rmul(rmix(1.0f, Fd90, FL), rmix(1.0f, Fd90, FV));
    float kd = rmul(rdiv(1.0f, PT_PI), Fd);
    float3 diffuse = f3(rmul(p.albedo.x, kd), rmul(p.albedo.y, kd), rmul(p.albedo.z, kd));
    /* microfacet :25-44 (specular 0.5, specularTint 0 -> Cspec0 = mix(0.04, albedo, metallic)) */
    const float s0 = rmul(0.5f, 0.08f);
    float3 Cspec0 = f3(rmix(s0, p.albedo.x, p.metallic), rmix(s0, p.albedo.y, p.metallic), rmix(s0, p.albedo.z, p.metallic));
    float a = fmaxf(0.001f, rmul(p.roughness, p.roughness));
    float Ds = gtr2(NdotH, a);
    float FH = schlickW(LdotH);
    float3 Fs = f3(rmix(Cspec0.x, 1.0f, FH), rmix(Cspec0.y, 1.0f, FH), rmix(Cspec0.z, 1.0f, FH));
    float Gs = rmul(smithGGX(NdotL, a), smithGGX(NdotV, a));
    float gd = rmul(Gs, Ds);
    float km = rsub(1.0f, p.metallic);
    return f3(rmul(radd(rmul(diffuse.x, km), rmul(Fs.x, gd)), NdotL), rmul(radd(rmul(diffuse.y, km), rmul(Fs.y, gd)), NdotL),
              rmul(radd(rmul(diffuse.z, km), rmul(Fs.z, gd)), NdotL));
}
PTC_D float pbrPdfMicrofacet(float3 wi, float3 wo, const Pbr &p) { /* :46-63 */
    if (!(wo.y > 0.0f) || !(wi.y > 0.0f)) return 0.0f;
    float3 wh = rnormalize(f3(radd(wo.x, wi.x), radd(wo.y, wi.y), radd(wo.z, wi.z)));
    float NdotH = fmaxf(wh.y, PT_EPSILON);
    float a2 = rmul(p.roughness, p.roughness);
    a2 = rmul(a2, a2);
    float denom = radd(rmul(rmul(NdotH, NdotH), rsub(a2, 1.0f)), 1.0f);
    if (denom == 0.0f) return 0.0f;
    float pd = rdiv(rmul(a2, NdotH), rmul(rmul(PT_PI, denom), denom));
    return rdiv(pd, rmul(4.0f, rdot(wo, wh)));
}
PTC_D float pbrPdf(float3 wi, float3 wo, const Pbr &p) { /* :123-137 */
    if (wi.y < 0.0f) return 0.0f;
    float r = diffuseRatio(p);
    return radd(rmul(rmul(wi.y, PT_INV_PI), r), rmul(pbrPdfMicrofacet(wi, wo, p), rsub(1.0f, r)));
}
PTC_D float3 pbrSample(float3 &wi, float3 wo, float &pdf, const Pbr &p, float u0, float u1, float lobe) { /* :139-165 */
    if (lobe <= diffuseRatio(p)) {
        float unused;
        wi = cosineHemisphere(u0, u1, unused);
    } else { /* :65-81 */
        float phi = (2.0f * PT_PI) * u1;
        float alpha = p.roughness * p.roughness;
        float tan2 = alpha * alpha * u0 / (1.0f - u0);
        float cosT = 1.0f / sqrtf(1.0f + tan2);
        float sinT = sqrtf(fmaxf(PT_EPSILON, 1.0f - cosT * cosT));
        float sp, cp;
        sincosf(phi, &sp, &cp);
        float3 wh = f3(sinT * cp, cosT, sinT * sp);
        if (!(wh.y > 0.0f)) wh = -wh;
        /* reflect(-wo, wh) */
        wi = -wo + wh * (2.0f * dot(wh, wo));
    }
    float3 F = pbrEval(p, wi, wo);
    pdf = pbrPdf(wi, wo, p);
    if (pdf < PT_EPSILON) return f3(0.0f);
    return F;
}

/* ------------------------------------------------------------------ Henyey-Greenstein (phaseFunctions.glsl) */
PTC_D float hg(float cosT, float g) {
    float d = 1.0f + g * g + 2.0f * g * cosT;
    return PT_INV_4PI * (1.0f - g * g) / (d * sqrtf(d));
}
PTC_D void coordinateSystem(float3 v1, float3 &v2, float3 &v3) { /* frame.glsl:63-74 */
    if (fabsf(v1.x) > fabsf(v1.y))
        v2 = f3(-v1.z, 0.0f, v1.x) / sqrtf(v1.x * v1.x + v1.z * v1.z);
    else
        v2 = f3(0.0f, v1.z, -v1.y) / sqrtf(v1.y * v1.y + v1.z * v1.z);
    v3 = cross(v1, v2);
}
PTC_D float hgSample(float3 wo, float3 &wi, float u0, float u1, float g) { /* :10-32 */
    float cosT;
    if (fabsf(g) < 1e-3f) {
        cosT = 1.0f - 2.0f * u0;
    } else {
        float sq = (1.0f - g * g) / (1.0f + g - 2.0f * g * u0);
        cosT = -(1.0f + g * g - sq * sq) / (2.0f * g);
    }
    float sinT = sqrtf(fmaxf(0.0f, 1.0f - cosT * cosT));
    float sp, cp;
    sincosf(2.0f * PT_PI * u1, &sp, &cp);
    float3 v1, v2;
    coordinateSystem(wo, v1, v2);
    wi = v1 * (sinT * cp) + v2 * (sinT * sp) + wo * cosT;
    return hg(cosT, g);
}

/* ------------------------------------------------------------------ shading frame (frame.glsl) */
struct Frame {
    float3 n, t, b;
};
PTC_D bool fixFrame(Frame &f, float3 ray) { /* :7-36 */
    bool flipped = false;
    if (dot(f.n, ray) > 0.0f) {
        f.n = -f.n;
        f.t = -f.t;
        flipped = true;
    }
    f.n = normalize(f.n);
    f.t = normalize(f.t);
    if (fabsf(dot(f.n, f.t)) > 0.999f) {
        f.b = fabsf(f.n.z) < 0.999f ? f3(0.0f, 0.0f, 1.0f) : f3(1.0f, 0.0f, 0.0f);
        f.t = cross(f.b, f.n);
        f.b = cross(f.n, f.t);
    } else {
        f.b = cross(f.n, f.t);
        f.t = cross(f.b, f.n);
    }
    return flipped;
}
PTC_D float3 toWorld(const Frame &f, float3 v) { return f.t * v.x + f.n * v.y + f.b * v.z; }
PTC_D float3 toLocal(const Frame &f, float3 v) { return f3(dot(v, f.t), dot(v, f.n), dot(v, f.b)); }
PTC_D void applyNormal(Frame &f, float3 nLocal) { /* :51-58 */
    f.n = toWorld(f, nLocal);
    f.t = cross(f.n, f.b);
    f.b = cross(f.t, f.n);
}
PTC_D float3 normalFromMap(float3 c) { /* utils.glsl:10-14 */
    float3 N = c * 2.0f - f3(1.0f);
    return normalize(f3(N.x, N.z, -N.y));
}
