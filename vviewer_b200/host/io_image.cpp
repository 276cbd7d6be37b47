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
    if (n < 0) { /* not seekable / error */
        std::fclose(f);
        return false;
    }
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

/* PNG from memory; rows top first.  srcChannels = the channel count stbi reports for the file.  Accepts what the reference's decoder
 * (stb_image) accepts: colour types 0 / 2 / 3 / 4 / 6, bit depths 1 / 2 / 4 / 8 / 16 (16 keeps the high byte, sub-byte grey is scaled
 * to 0..255), Adam7 interlacing; like stb it does not verify chunk CRCs.  The file is not trusted: chunk lengths, the IHDR size, the
 * image dimensions (w * h <= 2^28 pixels) and the inflated size are checked before anything is allocated or indexed. */
static bool decodePNG(const uint8_t *bytes, size_t nBytes, ImageU8 &out, int *srcChannels) {
    if (nBytes < 33 || std::memcmp(bytes, kPngSignature, 8) != 0) return false;
    uint32_t w = 0, h = 0;
    int bitDepth = 0, colorType = 0, interlace = 0;
    bool haveHeader = false;
    std::vector<uint8_t> idat, plte, trns;
    size_t pos = 8;
    while (pos + 12 <= nBytes) {
        const uint32_t len = be32(bytes + pos);
        const uint8_t *type = bytes + pos + 4;
        const uint8_t *data = bytes + pos + 8;
        if ((size_t)len > nBytes - pos - 12) return false;
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len != 13 || haveHeader) return false;
            w = be32(data);
            h = be32(data + 4);
            bitDepth = data[8];
            colorType = data[9];
            if (data[10] != 0 || data[11] != 0) return false; /* compression / filter method */
            interlace = data[12];
            haveHeader = true;
        } else if (!haveHeader) {
            return false; /* IHDR must come first */
        } else if (!std::memcmp(type, "PLTE", 4)) {
            plte.assign(data, data + len);
        } else if (!std::memcmp(type, "tRNS", 4)) {
            trns.assign(data, data + len);
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (!haveHeader || w == 0 || h == 0 || interlace > 1) return false;
    if ((uint64_t)w * h > (1ull << 28)) return false;
    const int srcCh = colorType == 0 ? 1 : colorType == 2 ? 3 : colorType == 3 ? 1 : colorType == 4 ? 2 : colorType == 6 ? 4 : 0;
    if (!srcCh) return false;
    const bool depthOk = colorType == 0 ? (bitDepth == 1 || bitDepth == 2 || bitDepth == 4 || bitDepth == 8 || bitDepth == 16)
                       : colorType == 3 ? (bitDepth == 1 || bitDepth == 2 || bitDepth == 4 || bitDepth == 8)
                                        : (bitDepth == 8 || bitDepth == 16);
    if (!depthOk) return false;
    const int bitsPerPixel = srcCh * bitDepth;
    const int bpp = std::max(1, bitsPerPixel / 8); /* filter distance in bytes */
    /* the (sub-)images the stream holds: the whole image, or the seven Adam7 passes */
    struct Pass {
        uint32_t x0, y0, dx, dy, pw, ph;
        size_t stride;
    };
    std::vector<Pass> passes;
    if (interlace == 0) {
        passes.push_back({0, 0, 1, 1, w, h, 0});
    } else {
        const uint32_t xo[7] = {0, 4, 0, 2, 0, 1, 0}, yo[7] = {0, 0, 4, 0, 2, 0, 1}, xs[7] = {8, 8, 4, 4, 2, 2, 1}, ys[7] = {8, 8, 8, 4, 4, 2, 2};
        for (int k = 0; k < 7; k++) {
            const uint32_t pw = w > xo[k] ? (w - xo[k] + xs[k] - 1) / xs[k] : 0, ph = h > yo[k] ? (h - yo[k] + ys[k] - 1) / ys[k] : 0;
            if (pw && ph) passes.push_back({xo[k], yo[k], xs[k], ys[k], pw, ph, 0});
        }
    }
    size_t rawSize = 0;
    for (Pass &ps : passes) {
        ps.stride = ((size_t)ps.pw * bitsPerPixel + 7) / 8;
        rawSize += (ps.stride + 1) * ps.ph;
    }
    std::vector<uint8_t> raw(rawSize);
    uLongf rawLen = (uLongf)raw.size();
    if (idat.empty() || uncompress(raw.data(), &rawLen, idat.data(), (uLong)idat.size()) != Z_OK || rawLen != raw.size()) return false;
    /* unfilter every pass in place (row by row, previous row of the SAME pass), then scatter its samples: one byte per sample
     * (the high byte of 16-bit samples, sub-byte samples unpacked) into img[h][w][srcCh] */
    std::vector<uint8_t> img((size_t)w * h * srcCh);
    std::vector<uint8_t> cur, prev;
    size_t rp = 0;
    const int greyScale = (colorType == 0 && bitDepth < 8) ? 255 / ((1 << bitDepth) - 1) : 1;
    for (const Pass &ps : passes) {
        cur.assign(ps.stride, 0);
        prev.assign(ps.stride, 0);
        for (uint32_t y = 0; y < ps.ph; y++) {
            const int ft = raw[rp];
            const uint8_t *src = &raw[rp + 1];
            rp += ps.stride + 1;
            for (size_t i = 0; i < ps.stride; i++) {
                const int a = i >= (size_t)bpp ? cur[i - bpp] : 0;
                const int b = y ? prev[i] : 0;
                const int c = (y && i >= (size_t)bpp) ? prev[i - bpp] : 0;
                int v = src[i];
                switch (ft) {
                    case 0: break;
                    case 1: v += a; break;
                    case 2: v += b; break;
                    case 3: v += (a + b) >> 1; break;
                    case 4: {
                        const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                        v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                        break;
                    }
                    default: return false;
                }
                cur[i] = (uint8_t)v;
            }
            const uint32_t oy = ps.y0 + y * ps.dy;
            for (uint32_t x = 0; x < ps.pw; x++) {
                uint8_t *dst = &img[((size_t)oy * w + (ps.x0 + x * ps.dx)) * srcCh];
                for (int ch = 0; ch < srcCh; ch++) {
                    const size_t sample = (size_t)x * srcCh + ch;
                    if (bitDepth == 16)
                        dst[ch] = cur[sample * 2];
                    else if (bitDepth == 8)
                        dst[ch] = cur[sample];
                    else {
                        const size_t bit = sample * bitDepth;
                        const int v = (cur[bit >> 3] >> (8 - bitDepth - (int)(bit & 7))) & ((1 << bitDepth) - 1);
                        dst[ch] = (uint8_t)(v * greyScale);
                    }
                }
            }
            cur.swap(prev);
        }
    }
    /* channel count as stbi_info reports it; 1 stays 1, everything else is forced to RGBA
     * (Image<stbi_uc>::loadDiskImage) */
    const int infoCh = colorType == 3 ? (trns.empty() ? 3 : 4) : srcCh;
    out.width = (int)w;
    out.height = (int)h;
    out.channels = infoCh == 1 ? 1 : 4;
    out.data.resize((size_t)w * h * out.channels);
    for (size_t p = 0; p < (size_t)w * h; p++) {
        const uint8_t *sp = &img[p * srcCh];
        uint8_t px[4] = {0, 0, 0, 255};
        switch (colorType) {
            case 0: px[0] = px[1] = px[2] = sp[0]; break;
            case 2: px[0] = sp[0]; px[1] = sp[1]; px[2] = sp[2]; break;
            case 3: {
                const size_t k = sp[0];
                if (k * 3 + 2 < plte.size()) { px[0] = plte[k * 3]; px[1] = plte[k * 3 + 1]; px[2] = plte[k * 3 + 2]; }
                if (k < trns.size()) px[3] = trns[k];
                break;
            }
            case 4: px[0] = px[1] = px[2] = sp[0]; px[3] = sp[1]; break;
            case 6: px[0] = sp[0]; px[1] = sp[1]; px[2] = sp[2]; px[3] = sp[3]; break;
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
