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
PTC_D float rnd(uint32_t &s) {
    s ^= s << 13;
    s ^= s >> 17;
    s ^= s << 5;
    return __uint_as_float(0x3f800000u | (s >> 9)) - 1.0f;
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
PTC_D float powerHeuristic(float fPdf, float gPdf) { /* MIS.glsl:5-10 with nf = ng = 1 */
    float f2 = fPdf * fPdf, g2 = gPdf * gPdf;
    return f2 / (f2 + g2);
}

/* ------------------------------------------------------------------ PBR standard (pbrStandard.glsl, common.glsl) */
struct Pbr {
    float3 albedo;
    float metallic, roughness;
};
PTC_D float schlickW(float c) {
    float m = clampf(1.0f - c, 0.0f, 1.0f);
    return (m * m) * (m * m) * m;
}
PTC_D float gtr2(float NdotH, float a) {
    float a2 = a * a;
    float t = 1.0f + (a2 - 1.0f) * NdotH * NdotH;
    return a2 / (PT_PI * t * t);
}
PTC_D float smithGGX(float NdotV, float alphaG) {
    float a = alphaG * alphaG, b = NdotV * NdotV;
    return 1.0f / (fabsf(NdotV) + fmaxf(sqrtf(a + b - a * b), PT_EPSILON));
}
PTC_D float diffuseRatio(const Pbr &p) { /* :83-90 */
    float d = fmaxf(1.0f - p.metallic, 0.1f), g = fmaxf(1.0f - p.roughness, 0.1f);
    return d / (d + g);
}
PTC_D float3 pbrEval(const Pbr &p, float3 wi, float3 wo) { /* evalPBRStandard :92-105 with H = normalize(wi + wo) */
    float NdotL = wi.y, NdotV = wo.y;
    if (NdotL < 0.0f || NdotV < 0.0f) return f3(0.0f);
    float3 H = normalize(wo + wi);
    float NdotH = H.y, LdotH = dot(wi, H);
    /* Disney diffuse :10-19 */
    float FL = schlickW(NdotL), FV = schlickW(NdotV);
    float Fd90 = 0.5f + 2.0f * LdotH * LdotH * p.roughness;
    float Fd = mixf(1.0f, Fd90, FL) * mixf(1.0f, Fd90, FV);
    float3 diffuse = p.albedo * ((1.0f / PT_PI) * Fd);
    /* microfacet :25-44 (specular 0.5, specularTint 0 -> Cspec0 = mix(0.04, albedo, metallic)) */
    float3 Cspec0 = mix3(f3(0.5f * 0.08f), p.albedo, p.metallic);
    float a = fmaxf(0.001f, p.roughness * p.roughness);
    float Ds = gtr2(NdotH, a);
    float3 Fs = mix3(Cspec0, f3(1.0f), schlickW(LdotH));
    float Gs = smithGGX(NdotL, a) * smithGGX(NdotV, a);
    float3 glossy = Fs * (Gs * Ds);
    return (diffuse * (1.0f - p.metallic) + glossy) * NdotL;
}
PTC_D float pbrPdfMicrofacet(float3 wi, float3 wo, const Pbr &p) { /* :46-63 */
    if (!(wo.y > 0.0f) || !(wi.y > 0.0f)) return 0.0f;
    float3 wh = normalize(wo + wi);
    float NdotH = fmaxf(wh.y, PT_EPSILON);
    float a2 = p.roughness * p.roughness;
    a2 *= a2;
    float denom = NdotH * NdotH * (a2 - 1.0f) + 1.0f;
    if (denom == 0.0f) return 0.0f;
    return (a2 * NdotH / (PT_PI * denom * denom)) / (4.0f * dot(wo, wh));
}
PTC_D float pbrPdf(float3 wi, float3 wo, const Pbr &p) { /* :123-137 */
    if (wi.y < 0.0f) return 0.0f;
    float r = diffuseRatio(p);
    return (wi.y * PT_INV_PI) * r + pbrPdfMicrofacet(wi, wo, p) * (1.0f - r);
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
