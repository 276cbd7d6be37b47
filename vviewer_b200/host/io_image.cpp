/*
 * Image I/O for the host side: PNG (8/16-bit, non-interlaced) in/out through zlib, Radiance RGBE in/out.
 * Replaces the stb_image / stb_image_write calls of the reference (src/lib/vengine/core/Image.cpp:9-43,
 * src/lib/vengine/core/ImageUtils.cpp:34-76) with the same numeric conventions:
 *   - RGBE decode: mantissa * 2^(e - 136), e == 0 -> 0           (stbi__hdr_convert)
 *   - RGBE encode: frexp(max) * 256 / max, truncation            (stbiw__linear_to_rgbe)
 *   - PNG out: clamp, linear -> sRGB, uchar(255 * x) truncation  (ImageUtils.cpp:58-66)
 */
#include "vengine.hpp"

#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace vengine {

static bool readFile(const std::string &path, std::vector<uint8_t> &out) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out.resize((size_t)n);
    size_t got = n > 0 ? std::fread(out.data(), 1, (size_t)n, f) : 0;
    std::fclose(f);
    return got == (size_t)n;
}

static void flipRows(uint8_t *data, int rowBytes, int h) {
    std::vector<uint8_t> tmp(rowBytes);
    for (int y = 0; y < h / 2; y++) {
        uint8_t *a = data + (size_t)y * rowBytes, *b = data + (size_t)(h - 1 - y) * rowBytes;
        std::memcpy(tmp.data(), a, rowBytes);
        std::memcpy(a, b, rowBytes);
        std::memcpy(b, tmp.data(), rowBytes);
    }
}

/* ---------------------------------------------------------------- PNG in */
static uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

static const uint8_t kPngSignature[8] = {137, 80, 78, 71, 13, 10, 26, 10};

/* PNG from memory; rows top first.  srcChannels = the channel count stbi reports for the file */
static bool decodePNG(const uint8_t *bytes, size_t nBytes, ImageU8 &out, int *srcChannels) {
    if (nBytes < 33 || std::memcmp(bytes, kPngSignature, 8) != 0) return false;
    struct View {
        const uint8_t *d;
        size_t n;
        size_t size() const { return n; }
        const uint8_t &operator[](size_t i) const { return d[i]; }
    } file{bytes, nBytes};
    uint32_t w = 0, h = 0;
    int bitDepth = 0, colorType = 0, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    size_t pos = 8;
    while (pos + 12 <= file.size()) {
        uint32_t len = be32(&file[pos]);
        const uint8_t *type = &file[pos + 4];
        const uint8_t *data = &file[pos + 8];
        if (pos + 12 + len > file.size()) return false;
        if (!std::memcmp(type, "IHDR", 4)) {
            w = be32(data);
            h = be32(data + 4);
            bitDepth = data[8];
            colorType = data[9];
            interlace = data[12];
        } else if (!std::memcmp(type, "PLTE", 4)) {
            plte.assign(data, data + len);
        } else if (!std::memcmp(type, "tRNS", 4)) {
            trns.assign(data, data + len);
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + len;
    }
    if (w == 0 || h == 0 || interlace != 0) return false;
    if (bitDepth != 8 && bitDepth != 16) return false;
    int srcCh = colorType == 0 ? 1 : colorType == 2 ? 3 : colorType == 3 ? 1 : colorType == 4 ? 2 : colorType == 6 ? 4 : 0;
    if (!srcCh) return false;
    int bps = bitDepth / 8;
    size_t stride = (size_t)w * srcCh * bps;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf rawLen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rawLen, idat.data(), (uLong)idat.size()) != Z_OK || rawLen != raw.size()) return false;
    /* unfilter */
    int bpp = srcCh * bps;
    std::vector<uint8_t> img(stride * h);
    for (uint32_t y = 0; y < h; y++) {
        int ft = raw[y * (stride + 1)];
        const uint8_t *src = &raw[y * (stride + 1) + 1];
        uint8_t *dst = &img[y * stride];
        const uint8_t *up = y ? &img[(y - 1) * stride] : nullptr;
        for (size_t i = 0; i < stride; i++) {
            int a = i >= (size_t)bpp ? dst[i - bpp] : 0;
            int b = up ? up[i] : 0;
            int c = (up && i >= (size_t)bpp) ? up[i - bpp] : 0;
            int v = src[i];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: {
                    int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                    v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    break;
                }
                default: return false;
            }
            dst[i] = (uint8_t)v;
        }
    }
    /* channel count as stbi_info reports it; 1 stays 1, everything else is forced to RGBA
     * (Image<stbi_uc>::loadDiskImage) */
    int infoCh = colorType == 3 ? (trns.empty() ? 3 : 4) : srcCh;
    out.width = (int)w;
    out.height = (int)h;
    out.channels = infoCh == 1 ? 1 : 4;
    out.data.resize((size_t)w * h * out.channels);
    for (size_t p = 0; p < (size_t)w * h; p++) {
        const uint8_t *s = &img[p * bpp];
        uint8_t px[4] = {0, 0, 0, 255};
        switch (colorType) {
            case 0: px[0] = px[1] = px[2] = s[0]; break;
            case 2: px[0] = s[0]; px[1] = s[bps]; px[2] = s[2 * bps]; break;
            case 3: {
                size_t k = s[0];
                if (k * 3 + 2 < plte.size()) { px[0] = plte[k * 3]; px[1] = plte[k * 3 + 1]; px[2] = plte[k * 3 + 2]; }
                if (k < trns.size()) px[3] = trns[k];
                break;
            }
            case 4: px[0] = px[1] = px[2] = s[0]; px[3] = s[bps]; break;
            case 6: px[0] = s[0]; px[1] = s[bps]; px[2] = s[2 * bps]; px[3] = s[3 * bps]; break;
        }
        if (out.channels == 1)
            out.data[p] = px[0];
        else
            std::memcpy(&out.data[p * 4], px, 4);
    }
    if (srcChannels) *srcChannels = infoCh;
    return true;
}

bool decodeImageU8(const uint8_t *bytes, size_t nBytes, ImageU8 &out, int *srcChannels, bool flipVertically) {
    bool ok = false;
    if (nBytes >= 8 && std::memcmp(bytes, kPngSignature, 8) == 0)
        ok = decodePNG(bytes, nBytes, out, srcChannels);
    else if (nBytes >= 3 && bytes[0] == 0xFF && bytes[1] == 0xD8 && bytes[2] == 0xFF)
        ok = decodeJPEG(bytes, nBytes, out, srcChannels);
    if (ok && flipVertically) flipRows(out.data.data(), out.width * out.channels, out.height);
    return ok;
}

bool loadImageU8(const std::string &path, ImageU8 &out, bool flipVertically, int *srcChannels) {
    std::vector<uint8_t> file;
    if (!readFile(path, file)) return false;
    return decodeImageU8(file.data(), file.size(), out, srcChannels, flipVertically);
}

/* ---------------------------------------------------------------- PNG out */
bool writeImagePNG(const std::string &path, int w, int h, int channels, const uint8_t *data) {
    if (channels < 1 || channels > 4) return false;
    size_t stride = (size_t)w * channels;
    std::vector<uint8_t> raw((stride + 1) * h);
    for (int y = 0; y < h; y++) {
        raw[y * (stride + 1)] = 0;
        std::memcpy(&raw[y * (stride + 1) + 1], data + (size_t)y * stride, stride);
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    std::fwrite(kPngSignature, 1, 8, f);
    auto chunk = [&](const char *type, const uint8_t *d, uint32_t len) {
        uint8_t hdr[8] = {(uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len, (uint8_t)type[0], (uint8_t)type[1],
                          (uint8_t)type[2], (uint8_t)type[3]};
        std::fwrite(hdr, 1, 8, f);
        if (len) std::fwrite(d, 1, len, f);
        uLong crc = crc32(0L, hdr + 4, 4);
        if (len) crc = crc32(crc, d, len);
        uint8_t c[4] = {(uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc};
        std::fwrite(c, 1, 4, f);
    };
    static const int ctype[5] = {0, 0, 4, 2, 6};
    uint8_t ihdr[13] = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w, (uint8_t)(h >> 24), (uint8_t)(h >> 16),
                        (uint8_t)(h >> 8), (uint8_t)h, 8, (uint8_t)ctype[channels], 0, 0, 0};
    chunk("IHDR", ihdr, 13);
    chunk("IDAT", comp.data(), (uint32_t)clen);
    chunk("IEND", nullptr, 0);
    std::fclose(f);
    return true;
}

/* ---------------------------------------------------------------- Radiance HDR in */
bool loadImageHDR(const std::string &path, ImageF32 &out, bool flipVertically) {
    std::vector<uint8_t> file;
    if (!readFile(path, file)) return false;
    size_t pos = 0;
    auto readLine = [&](std::string &line) {
        line.clear();
        while (pos < file.size() && file[pos] != '\n') line.push_back((char)file[pos++]);
        if (pos < file.size()) pos++;
        return true;
    };
    std::string line;
    readLine(line);
    if (line.rfind("#?RADIANCE", 0) != 0 && line.rfind("#?RGBE", 0) != 0) return false;
    while (pos < file.size()) {
        readLine(line);
        if (line.empty()) break;
    }
    readLine(line);
    int w = 0, h = 0;
    if (std::sscanf(line.c_str(), "-Y %d +X %d", &h, &w) != 2 || w <= 0 || h <= 0) return false;
    std::vector<uint8_t> rgbe((size_t)w * h * 4);
    for (int y = 0; y < h; y++) {
        uint8_t *row = &rgbe[(size_t)y * w * 4];
        if (pos + 4 > file.size()) return false;
        bool rle = w >= 8 && w < 32768 && file[pos] == 2 && file[pos + 1] == 2 && !(file[pos + 2] & 0x80) &&
                   (((int)file[pos + 2] << 8) | file[pos + 3]) == w;
        if (!rle) {
            /* flat scanline(s): the rest of the file is uncompressed from here on */
            size_t need = (size_t)w * 4;
            if (pos + need > file.size()) return false;
            std::memcpy(row, &file[pos], need);
            pos += need;
            continue;
        }
        pos += 4;
        for (int c = 0; c < 4; c++) {
            int x = 0;
            while (x < w) {
                if (pos >= file.size()) return false;
                int count = file[pos++];
                if (count > 128) {
                    count -= 128;
                    if (pos >= file.size() || x + count > w) return false;
                    uint8_t v = file[pos++];
                    for (int k = 0; k < count; k++) row[(x++) * 4 + c] = v;
                } else {
                    if (count == 0 || pos + count > file.size() || x + count > w) return false;
                    for (int k = 0; k < count; k++) row[(x++) * 4 + c] = file[pos++];
                }
            }
        }
    }
    out.width = w;
    out.height = h;
    out.channels = 4;
    out.data.resize((size_t)w * h * 4);
    for (int y = 0; y < h; y++) {
        int sy = flipVertically ? (h - 1 - y) : y;
        for (int x = 0; x < w; x++) {
            const uint8_t *p = &rgbe[((size_t)sy * w + x) * 4];
            float *o = &out.data[((size_t)y * w + x) * 4];
            if (p[3] != 0) {
                float f = std::ldexp(1.0f, (int)p[3] - (128 + 8));
                o[0] = p[0] * f;
                o[1] = p[1] * f;
                o[2] = p[2] * f;
            } else {
                o[0] = o[1] = o[2] = 0.0f;
            }
            o[3] = 1.0f;
        }
    }
    return true;
}

/* ---------------------------------------------------------------- Radiance HDR out */
static void toRGBE(const float *rgb, uint8_t *o) {
    float m = std::max(rgb[0], std::max(rgb[1], rgb[2]));
    if (m < 1e-32f) {
        o[0] = o[1] = o[2] = o[3] = 0;
    } else {
        int e;
        float n = std::frexp(m, &e) * 256.0f / m;
        o[0] = (uint8_t)(rgb[0] * n);
        o[1] = (uint8_t)(rgb[1] * n);
        o[2] = (uint8_t)(rgb[2] * n);
        o[3] = (uint8_t)(e + 128);
    }
}

bool writeImageHDR(const std::string &path, int w, int h, int channels, const float *data) {
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    std::fprintf(f, "#?RADIANCE\n# Written by vviewer_b200\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=          1.0000000000000\n\n-Y %d +X %d\n", h, w);
    std::vector<uint8_t> line((size_t)w * 4), enc;
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            float rgb[3];
            const float *p = data + ((size_t)y * w + x) * channels;
            if (channels >= 3) {
                rgb[0] = p[0]; rgb[1] = p[1]; rgb[2] = p[2];
            } else {
                rgb[0] = rgb[1] = rgb[2] = p[0];
            }
            toRGBE(rgb, &line[(size_t)x * 4]);
        }
        if (w < 8 || w >= 32768) {
            std::fwrite(line.data(), 1, line.size(), f);
            continue;
        }
        uint8_t hdr[4] = {2, 2, (uint8_t)(w >> 8), (uint8_t)(w & 255)};
        std::fwrite(hdr, 1, 4, f);
        for (int c = 0; c < 4; c++) {
            enc.clear();
            int x = 0;
            while (x < w) {
                /* find the next run of >= 3 equal values */
                int r = x;
                while (r + 2 < w) {
                    if (line[r * 4 + c] == line[(r + 1) * 4 + c] && line[r * 4 + c] == line[(r + 2) * 4 + c]) break;
                    r++;
                }
                if (r + 2 >= w) r = w;
                while (x < r) { /* literals */
                    int len = std::min(r - x, 128);
                    enc.push_back((uint8_t)len);
                    for (int k = 0; k < len; k++) enc.push_back(line[(x + k) * 4 + c]);
                    x += len;
                }
                if (r + 2 < w) { /* run */
                    while (r < w && line[r * 4 + c] == line[x * 4 + c]) r++;
                    while (x < r) {
                        int len = std::min(r - x, 127);
                        enc.push_back((uint8_t)(len + 128));
                        enc.push_back(line[x * 4 + c]);
                        x += len;
                    }
                }
            }
            std::fwrite(enc.data(), 1, enc.size(), f);
        }
    }
    std::fclose(f);
    return true;
}

/* ---------------------------------------------------------------- ImageUtils.cpp */
float linearToSRGB(float v) {
    if (v <= 0.0031308f) return 12.92f * v;
    return 1.055f * std::pow(v, 1.0f / 2.4f) - 0.055f;
}

void applyExposure(std::vector<float> &in, float exposure, uint32_t channels) {
    uint32_t step = std::min(channels, 3u);
    float s = std::pow(2.0f, exposure);
    for (size_t i = 0; i + channels <= in.size(); i += channels)
        for (uint32_t c = 0; c < step; c++) in[i + c] = in[i + c] * s;
}

void writeToDisk(const std::vector<float> &in, const std::string &filename, FileType type, uint32_t w, uint32_t h, uint32_t channels) {
    switch (type) {
        case FileType::PNG: {
            std::vector<uint8_t> im((size_t)w * h * channels, 255);
            for (size_t i = 0; i < im.size(); i++)
                im[i] = static_cast<uint8_t>(255.0f * linearToSRGB(std::min(std::max(in[i], 0.0f), 1.0f)));
            writeImagePNG(filename + ".png", (int)w, (int)h, (int)channels, im.data());
            break;
        }
        case FileType::HDR: writeImageHDR(filename + ".hdr", (int)w, (int)h, (int)channels, in.data()); break;
    }
}

}  // namespace vengine
