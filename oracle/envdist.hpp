/*
 * TEST INFRASTRUCTURE (oracle): luminance importance distribution over the environment, for PTC_FLAG_ENV_IMPORTANCE.
 *
 * NOT reference behaviour: the reference never light-samples the environment (lightSampling.glsl:101-106 is a TODO,
 * rayNEE.rmiss.glsl:12-19 adds nothing - SURVEY trap T3), so there is no reference file to restate.  The option delivers
 * the "HDR environment + MIS/NEE" of BASELINE config 2; this header is the CPU definition the device implementation
 * (vviewer_b200/csrc/envdist.cuh) has to agree with, table for table.
 *
 * Definition.  Domain: the equirectangular texture coordinates (u, v) in [0,1)^2 that include/environmentMap.glsl:1-10 maps a
 * direction to (u = atan(z, x) * 0.1591 + 0.5 + 0.25 mod 1, v = asin(y) * 0.3183 + 0.5).  Grid: 512 x 256 bins.
 * Weight of bin (i, j) = mean Rec.709 luminance of the equirect texels (x, y) with x * 512 / W == i and y * 256 / H == j,
 * times cos(latitude of the bin centre), plus 1e-3 of the mean weight (the pdf is positive wherever radiance can be).
 * Tables: marginal CDF over rows and one conditional CDF per row, floats, accumulated in double in index order.
 * Sampling: piecewise-constant inversion (row by u1, column by u2), direction by the inverse of the mapping above.
 * Solid-angle density: p(u, v) * 0.3183 * 0.1591 / cos(latitude).
 */
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace envdist {

constexpr int EW = 512, EH = 256;

struct Tables {
    std::vector<float> cdfV; /* EH + 1 */
    std::vector<float> cdfU; /* EH * (EW + 1) */
    bool valid = false;
};

inline void build(const float *rgba, uint32_t W, uint32_t H, Tables &t) {
    t.valid = false;
    t.cdfV.assign(EH + 1, 0.0f);
    t.cdfU.assign((size_t)EH * (EW + 1), 0.0f);
    if (!rgba || W == 0 || H == 0) return;
    std::vector<double> sum((size_t)EW * EH, 0.0);
    std::vector<uint32_t> cnt((size_t)EW * EH, 0u);
    for (uint32_t y = 0; y < H; y++) {
        const uint32_t j = (uint32_t)(((uint64_t)y * EH) / H);
        for (uint32_t x = 0; x < W; x++) {
            const uint32_t i = (uint32_t)(((uint64_t)x * EW) / W);
            const float *p = rgba + ((size_t)y * W + x) * 4;
            sum[(size_t)j * EW + i] += 0.2126 * (double)p[0] + 0.7152 * (double)p[1] + 0.0722 * (double)p[2];
            cnt[(size_t)j * EW + i]++;
        }
    }
    std::vector<double> w((size_t)EW * EH, 0.0);
    double total = 0.0;
    for (int j = 0; j < EH; j++) {
        const double lat = (((double)j + 0.5) / EH - 0.5) / 0.3183;
        const double c = std::max(std::cos(lat), 0.0);
        for (int i = 0; i < EW; i++) {
            const size_t k = (size_t)j * EW + i;
            /* bins without a texel (image smaller than the grid) take the nearest texel */
            double lum;
            if (cnt[k]) {
                lum = sum[k] / (double)cnt[k];
            } else {
                const uint32_t x = std::min<uint32_t>((uint32_t)(((uint64_t)i * W) / EW), W - 1), y = std::min<uint32_t>((uint32_t)(((uint64_t)j * H) / EH), H - 1);
                const float *p = rgba + ((size_t)y * W + x) * 4;
                lum = 0.2126 * (double)p[0] + 0.7152 * (double)p[1] + 0.0722 * (double)p[2];
            }
            if (!(lum >= 0.0) || !std::isfinite(lum)) lum = 0.0;
            w[k] = lum * c;
            total += w[k];
        }
    }
    const double floorW = total > 0.0 ? 1e-3 * total / ((double)EW * EH) : 1.0;
    std::vector<double> rowSum(EH, 0.0);
    double all = 0.0;
    for (int j = 0; j < EH; j++) {
        double acc = 0.0;
        float *row = &t.cdfU[(size_t)j * (EW + 1)];
        for (int i = 0; i < EW; i++) {
            row[i] = (float)acc; /* unnormalised for now */
            acc += w[(size_t)j * EW + i] + floorW;
        }
        rowSum[j] = acc;
        for (int i = 0; i < EW; i++) row[i] = (float)((double)row[i] / acc);
        row[EW] = 1.0f;
        all += acc;
    }
    double acc = 0.0;
    for (int j = 0; j < EH; j++) {
        t.cdfV[j] = (float)(acc / all);
        acc += rowSum[j];
    }
    t.cdfV[EH] = 1.0f;
    t.valid = true;
}

/* largest k in [0, n - 1] with cdf[k] <= u */
inline int findInterval(const float *cdf, int n, float u) {
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (cdf[mid] <= u) lo = mid;
        else hi = mid;
    }
    return lo;
}

/* density over (u, v) of the bin that holds (u, v) */
inline float pdfUV(const Tables &t, float u, float v) {
    const int j = std::min(std::max((int)(v * EH), 0), EH - 1), i = std::min(std::max((int)(u * EW), 0), EW - 1);
    const float *row = &t.cdfU[(size_t)j * (EW + 1)];
    return ((t.cdfV[j + 1] - t.cdfV[j]) * (float)EH) * ((row[i + 1] - row[i]) * (float)EW);
}

inline float cosLatitude(float v) { return std::cos((v - 0.5f) / 0.3183f); }

/* (u1, u2) -> (u, v) and the density over (u, v) */
inline void sampleUV(const Tables &t, float u1, float u2, float &u, float &v, float &pdf) {
    const int j = findInterval(t.cdfV.data(), EH, u1);
    const float dv = t.cdfV[j + 1] - t.cdfV[j];
    const float fv = dv > 0.0f ? (u1 - t.cdfV[j]) / dv : 0.5f;
    const float *row = &t.cdfU[(size_t)j * (EW + 1)];
    const int i = findInterval(row, EW, u2);
    const float du = row[i + 1] - row[i];
    const float fu = du > 0.0f ? (u2 - row[i]) / du : 0.5f;
    u = std::min(((float)i + fu) / (float)EW, 0.99999994f);
    v = std::min(((float)j + fv) / (float)EH, 0.99999994f);
    pdf = (dv * (float)EH) * (du * (float)EW);
}

/* inverse of include/environmentMap.glsl:1-10 */
inline void direction(float u, float v, float &x, float &y, float &z) {
    const float lat = (v - 0.5f) / 0.3183f;
    float uu = u - 0.25f;
    uu = uu - std::floor(uu);
    const float phi = (uu - 0.5f) / 0.1591f;
    const float r = std::max(std::cos(lat), 0.0f);
    y = std::sin(lat);
    x = r * std::cos(phi);
    z = r * std::sin(phi);
}

inline float solidAnglePdf(float pdfuv, float v) { return pdfuv * (0.3183f * 0.1591f) / std::max(cosLatitude(v), 1e-6f); }

}  // namespace envdist
