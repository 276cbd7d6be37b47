/*
 * Environment importance sampling for PTC_FLAG_ENV_IMPORTANCE (include/ptc.h): an extension without a reference
 * counterpart - the reference never light-samples the environment (lightSampling.glsl:101-106 TODO, SURVEY trap T3).
 *
 * 512 x 256 luminance x cos(latitude) table over the equirectangular domain of include/environmentMap.glsl:1-10
 * (u = atan(z, x) * 0.1591 + 0.5 + 0.25 mod 1, v = asin(y) * 0.3183 + 0.5): a marginal CDF over rows and one conditional
 * CDF per row (513 KB, L2 resident).  The tables are built on the host at scene upload (one pass over the equirect input,
 * double accumulators, index order) so that they are bit-identical to the CPU definition the parity tests compare with.
 */
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "common.cuh"

namespace envd {

constexpr int EW = 512, EH = 256;

struct DeviceTables {
    const float *cdfV; /* EH + 1 */
    const float *cdfU; /* EH rows of EW + 1 */
};

/* host: fills cdfV (EH + 1) and cdfU (EH * (EW + 1)); false without an environment */
inline bool buildTables(const float *rgba, uint32_t W, uint32_t H, std::vector<float> &cdfV, std::vector<float> &cdfU) {
    cdfV.assign(EH + 1, 0.0f);
    cdfU.assign((size_t)EH * (EW + 1), 0.0f);
    if (!rgba || W == 0 || H == 0) return false;
    const size_t nb = (size_t)EW * EH;
    std::vector<double> lumSum(nb, 0.0), weight(nb, 0.0);
    std::vector<uint32_t> count(nb, 0u);
    auto lum = [](const float *p) { return 0.2126 * (double)p[0] + 0.7152 * (double)p[1] + 0.0722 * (double)p[2]; };
    for (uint32_t y = 0; y < H; y++) {
        const size_t rowBin = (size_t)(((uint64_t)y * EH) / H) * EW;
        const float *row = rgba + (size_t)y * W * 4;
        for (uint32_t x = 0; x < W; x++) {
            const size_t k = rowBin + (size_t)(((uint64_t)x * EW) / W);
            lumSum[k] += lum(row + (size_t)x * 4);
            count[k]++;
        }
    }
    double total = 0.0;
    for (int j = 0; j < EH; j++) {
        const double latitude = (((double)j + 0.5) / EH - 0.5) / 0.3183;
        const double cosLat = std::max(std::cos(latitude), 0.0);
        for (int i = 0; i < EW; i++) {
            const size_t k = (size_t)j * EW + i;
            double l;
            if (count[k]) {
                l = lumSum[k] / (double)count[k];
            } else { /* grid finer than the image: nearest texel */
                const uint32_t x = std::min<uint32_t>((uint32_t)(((uint64_t)i * W) / EW), W - 1);
                const uint32_t y = std::min<uint32_t>((uint32_t)(((uint64_t)j * H) / EH), H - 1);
                l = lum(rgba + ((size_t)y * W + x) * 4);
            }
            if (!(l >= 0.0) || !std::isfinite(l)) l = 0.0;
            weight[k] = l * cosLat;
            total += weight[k];
        }
    }
    /* a floor keeps the density positive wherever the environment can hold radiance */
    const double floorW = total > 0.0 ? 1e-3 * total / ((double)EW * EH) : 1.0;
    std::vector<double> rowSum(EH, 0.0);
    double all = 0.0;
    for (int j = 0; j < EH; j++) {
        float *row = &cdfU[(size_t)j * (EW + 1)];
        double acc = 0.0;
        for (int i = 0; i < EW; i++) {
            row[i] = (float)acc;
            acc += weight[(size_t)j * EW + i] + floorW;
        }
        for (int i = 0; i < EW; i++) row[i] = (float)((double)row[i] / acc);
        row[EW] = 1.0f;
        rowSum[j] = acc;
        all += acc;
    }
    double acc = 0.0;
    for (int j = 0; j < EH; j++) {
        cdfV[j] = (float)(acc / all);
        acc += rowSum[j];
    }
    cdfV[EH] = 1.0f;
    return true;
}

#ifdef __CUDACC__
/* largest k in [0, n - 1] with cdf[k] <= u */
PTC_D int findInterval(const float *__restrict__ cdf, int n, float u) {
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(cdf + mid) <= u) lo = mid;
        else hi = mid;
    }
    return lo;
}

PTC_D float cosLatitude(float v) { return cosf((v - 0.5f) / 0.3183f); }
PTC_D float solidAnglePdf(float pdfuv, float v) { return pdfuv * (0.3183f * 0.1591f) / fmaxf(cosLatitude(v), 1e-6f); }

/* density over (u, v) of the bin that holds (u, v) */
PTC_D float pdfUV(const DeviceTables &t, float u, float v) {
    const int j = min(max((int)(v * (float)EH), 0), EH - 1), i = min(max((int)(u * (float)EW), 0), EW - 1);
    const float *row = t.cdfU + (size_t)j * (EW + 1);
    return ((__ldg(t.cdfV + j + 1) - __ldg(t.cdfV + j)) * (float)EH) * ((__ldg(row + i + 1) - __ldg(row + i)) * (float)EW);
}

PTC_D void sampleUV(const DeviceTables &t, float u1, float u2, float &u, float &v, float &pdf) {
    const int j = findInterval(t.cdfV, EH, u1);
    const float c0 = __ldg(t.cdfV + j), dv = __ldg(t.cdfV + j + 1) - c0;
    const float fv = dv > 0.0f ? (u1 - c0) / dv : 0.5f;
    const float *row = t.cdfU + (size_t)j * (EW + 1);
    const int i = findInterval(row, EW, u2);
    const float r0 = __ldg(row + i), du = __ldg(row + i + 1) - r0;
    const float fu = du > 0.0f ? (u2 - r0) / du : 0.5f;
    u = fminf(((float)i + fu) / (float)EW, 0.99999994f);
    v = fminf(((float)j + fv) / (float)EH, 0.99999994f);
    pdf = (dv * (float)EH) * (du * (float)EW);
}

/* inverse of include/environmentMap.glsl:1-10 */
PTC_D float3 direction(float u, float v) {
    const float latitude = (v - 0.5f) / 0.3183f;
    float uu = u - 0.25f;
    uu = uu - floorf(uu);
    const float phi = (uu - 0.5f) / 0.1591f;
    const float r = fmaxf(cosf(latitude), 0.0f);
    return make_float3(r * cosf(phi), sinf(latitude), r * sinf(phi));
}

/* include/environmentMap.glsl:1-10 */
PTC_D float2 equirectUV(float3 d) {
    float ux = atan2f(d.z, d.x) * 0.1591f + 0.5f;
    const float uy = asinf(fminf(fmaxf(d.y, -1.0f), 1.0f)) * 0.3183f + 0.5f;
    ux = ux + 0.25f;
    ux = ux - floorf(ux);
    return make_float2(ux, uy);
}
#endif

}  // namespace envd
