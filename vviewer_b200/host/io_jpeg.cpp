/*
 * JPEG decoder for the host side (baseline and progressive DCT, 8-bit, Huffman; 1 or 3 components; restart
 * intervals; any h/v sampling factors).  It replaces stb_image's `stbi_load_from_memory`, which the reference
 * uses for glTF-embedded and on-disk material textures (src/lib/vengine/core/io/AssimpLoadModel.cpp:172-225;
 * bundled case: the five 2048x2048 JPEGs of assets/models/DamagedHelmet.gltf, one of them progressive).
 *
 * Written from the JPEG specification (ITU-T T.81: marker syntax A/B, Huffman decoding F.2.2, progressive
 * decoding G.1.2) with the numeric choices that stb_image makes, because the texels feed the renderer:
 *   - inverse DCT: the Loeffler-Ligtenberg-Moschytz integer scheme with 12-bit constants, 2 extra bits kept
 *     between the passes, +128 level shift and clamp folded into the row pass;
 *   - chroma upsampling: "triangle" filter (3/4 near + 1/4 far, separable) for 2x horizontal / vertical /
 *     both, pixel replication for other factors; the near/far row pairing starts half a step in;
 *   - YCbCr -> RGB in 20-bit fixed point with the 1.402 / 0.71414 / 0.34414 / 1.772 constants.
 * tests/test_import.py checks the output against stb_image itself (oracle/_ref/stb_decode, compiled from the
 * reference's vendored header where it lies) and against committed checksums.
 */
#include "vengine.hpp"

#include <cstring>

namespace vengine {
namespace {

const uint8_t kZigzag[64 + 15] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,
                                  6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
                                  39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
                                  /* corrupt streams can run past 63: land on the last coefficient */
                                  63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};

struct HuffTable {
    bool present = false;
    uint8_t fast[512];  /* 9-bit prefix -> symbol slot, 255 = longer code */
    uint8_t size[257];  /* code length of slot k */
    uint16_t code[256]; /* canonical code of slot k */
    uint8_t value[256]; /* symbol of slot k */
    uint32_t maxcode[18];
    int delta[17];

    bool build(const uint8_t counts[16], const uint8_t *symbols, int nSymbols) {
        int k = 0;
        for (int len = 1; len <= 16; len++)
            for (int j = 0; j < counts[len - 1]; j++) {
                if (k >= 256) return false;
                size[k++] = (uint8_t)len;
            }
        if (k != nSymbols) return false;
        size[k] = 0;
        std::memcpy(value, symbols, (size_t)nSymbols);
        /* canonical codes: consecutive within a length, doubled when the length grows (T.81 C.2) */
        uint32_t next = 0;
        int slot = 0;
        for (int len = 1; len <= 16; len++) {
            delta[len] = slot - (int)next;
            while (size[slot] == len) code[slot++] = (uint16_t)next++;
            if (next > (1u << len)) return false;
            maxcode[len] = next << (16 - len);
            next <<= 1;
        }
        maxcode[17] = 0xffffffffu;
        std::memset(fast, 255, sizeof(fast));
        for (int i = 0; i < k; i++) {
            int s = size[i];
            if (s <= 9) {
                int first = code[i] << (9 - s), n = 1 << (9 - s);
                for (int j = 0; j < n; j++) fast[first + j] = (uint8_t)i;
            }
        }
        present = true;
        return true;
    }
};

/* MSB-first bit reader over entropy-coded data: removes FF00 stuffing, stops at any marker and feeds zero bits after it */
struct BitReader {
    const uint8_t *p = nullptr, *end = nullptr;
    uint64_t acc = 0;
    int bits = 0;
    int marker = -1; /* marker byte seen in the stream (p stays on its FF) */
    int starved = 0; /* zero bytes fed after the END of the data (not after a marker): a truncated file */

    void fill() {
        while (bits <= 56) {
            uint32_t byte = 0;
            if (marker < 0 && p < end) {
                if (*p != 0xFF) {
                    byte = *p++;
                } else if (p + 1 >= end) {
                    p = end;
                    ++starved;
                } else if (p[1] == 0x00) { /* stuffed data byte FF */
                    byte = 0xFF;
                    p += 2;
                } else if (p[1] == 0xFF) { /* fill byte before a marker */
                    ++p;
                    continue;
                } else {
                    marker = p[1];
                }
            } else if (marker < 0) {
                ++starved;
            }
            acc |= (uint64_t)byte << (56 - bits);
            bits += 8;
        }
    }
    uint32_t peek(int n) {
        if (bits < n) fill();
        return (uint32_t)(acc >> (64 - n));
    }
    void skip(int n) {
        acc <<= n;
        bits -= n;
    }
    uint32_t get(int n) {
        if (n == 0) return 0;
        uint32_t v = peek(n);
        skip(n);
        return v;
    }
    int getBit() { return (int)get(1); }
    void reset() {
        acc = 0;
        bits = 0;
    }
};

inline int extendSign(uint32_t v, int s) { return (v < (1u << (s - 1))) ? (int)v - (1 << s) + 1 : (int)v; }

struct Component {
    int id = 0, h = 1, v = 1, tq = 0;
    int td = 0, ta = 0;     /* Huffman table ids of the current scan */
    int x = 0, y = 0;       /* true size in samples */
    int w2 = 0, h2 = 0;     /* size padded to whole MCUs */
    int blocksW = 0, blocksH = 0;
    int dcPred = 0;
    std::vector<uint8_t> samples;  /* w2 x h2 */
    std::vector<int16_t> coeff;    /* progressive: blocksW x blocksH x 64 */
};

struct Decoder {
    const uint8_t *data, *end;
    uint16_t quant[4][64]; /* natural (de-zigzagged) order */
    bool quantPresent[4] = {false, false, false, false};
    HuffTable dc[4], ac[4];
    Component comp[4];
    int nComp = 0, width = 0, height = 0;
    int hmax = 1, vmax = 1, mcuX = 0, mcuY = 0;
    bool progressive = false, sawSOF = false;
    int restartInterval = 0;
    int adobeTransform = -1;
    bool jfif = false;
    /* scan parameters */
    int scanN = 0, order[4];
    int ss = 0, se = 63, ah = 0, al = 0;
    int eobrun = 0;
    BitReader br;

    Decoder(const uint8_t *d, size_t n) : data(d), end(d + n) {}

    static int be16(const uint8_t *p) { return (p[0] << 8) | p[1]; }

    int decodeSymbol(const HuffTable &h) {
        uint32_t top = br.peek(16);
        int slot = h.fast[top >> 7];
        if (slot < 255) {
            br.skip(h.size[slot]);
            return h.value[slot];
        }
        int len = 10;
        while (len <= 16 && top >= h.maxcode[len]) ++len;
        if (len > 16) return -1;
        int idx = (int)(top >> (16 - len)) + h.delta[len];
        if (idx < 0 || idx >= 256 || h.size[idx] != len) return -1;
        br.skip(len);
        return h.value[idx];
    }

    /* ---- sequential block (F.2.2) */
    bool blockSequential(int16_t *blk, Component &c) {
        const HuffTable &hd = dc[c.td], &ha = ac[c.ta];
        const uint16_t *q = quant[c.tq];
        int t = decodeSymbol(hd);
        if (t < 0 || t > 15) return false;
        std::memset(blk, 0, 64 * sizeof(int16_t));
        int diff = t ? extendSign(br.get(t), t) : 0;
        c.dcPred += diff;
        blk[0] = (int16_t)(c.dcPred * q[0]);
        for (int k = 1; k < 64;) {
            int rs = decodeSymbol(ha);
            if (rs < 0) return false;
            int r = rs >> 4, s = rs & 15;
            if (s == 0) {
                if (rs != 0xF0) break; /* end of block */
                k += 16;
            } else {
                k += r;
                int z = kZigzag[k++];
                blk[z] = (int16_t)(extendSign(br.get(s), s) * q[z]);
            }
        }
        return true;
    }

    /* ---- progressive DC (G.1.2.1) */
    bool blockProgressiveDC(int16_t *blk, Component &c) {
        if (se != 0) return false;
        if (ah == 0) {
            std::memset(blk, 0, 64 * sizeof(int16_t));
            int t = decodeSymbol(dc[c.td]);
            if (t < 0 || t > 15) return false;
            int diff = t ? extendSign(br.get(t), t) : 0;
            c.dcPred += diff;
            blk[0] = (int16_t)(c.dcPred * (1 << al));
        } else if (br.getBit()) {
            blk[0] = (int16_t)(blk[0] + (1 << al));
        }
        return true;
    }

    /* ---- progressive AC: first pass G.1.2.2, refinement G.1.2.3 */
    bool blockProgressiveAC(int16_t *blk, const HuffTable &h) {
        if (ss == 0) return false;
        if (ah == 0) {
            if (eobrun) {
                --eobrun;
                return true;
            }
            int k = ss;
            do {
                int rs = decodeSymbol(h);
                if (rs < 0) return false;
                int r = rs >> 4, s = rs & 15;
                if (s == 0) {
                    if (r < 15) {
                        eobrun = (1 << r);
                        if (r) eobrun += (int)br.get(r);
                        --eobrun;
                        break;
                    }
                    k += 16;
                } else {
                    k += r;
                    int z = kZigzag[k++];
                    blk[z] = (int16_t)(extendSign(br.get(s), s) * (1 << al));
                }
            } while (k <= se);
            return true;
        }
        /* refinement: one more bit for coefficients that are already non-zero, new +-1 coefficients in between */
        const int16_t bit = (int16_t)(1 << al);
        auto refine = [&](int16_t &c) {
            if (br.getBit() && (c & bit) == 0) c = (int16_t)(c > 0 ? c + bit : c - bit);
        };
        if (eobrun) {
            --eobrun;
            for (int k = ss; k <= se; k++) {
                int16_t &c = blk[kZigzag[k]];
                if (c != 0) refine(c);
            }
            return true;
        }
        int k = ss;
        do {
            int rs = decodeSymbol(h);
            if (rs < 0) return false;
            int r = rs >> 4, s = rs & 15;
            int newValue = 0;
            if (s == 0) {
                if (r < 15) {
                    eobrun = (1 << r) - 1;
                    if (r) eobrun += (int)br.get(r);
                    r = 64; /* run to the end of the band */
                }
                /* r == 15: skip 16 zero coefficients, nothing new */
            } else {
                if (s != 1) return false;
                newValue = br.getBit() ? bit : -bit;
            }
            while (k <= se) {
                int16_t &c = blk[kZigzag[k++]];
                if (c != 0) {
                    refine(c);
                } else {
                    if (r == 0) {
                        c = (int16_t)newValue;
                        break;
                    }
                    --r;
                }
            }
        } while (k <= se);
        return true;
    }

    /* ---- inverse DCT, one 8x8 block -> 8-bit samples */
    static inline uint8_t clamp8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
#define JF(x) ((int)(((x) * 4096 + 0.5)))
    static inline void idct1d(const int s[8], int x[4], int t[4]) {
        int p1, p2, p3, p4, p5;
        /* even part */
        p1 = (s[2] + s[6]) * JF(0.5411961f);
        int e2 = p1 + s[6] * JF(-1.847759065f);
        int e3 = p1 + s[2] * JF(0.765366865f);
        int e0 = (s[0] + s[4]) * 4096;
        int e1 = (s[0] - s[4]) * 4096;
        x[0] = e0 + e3;
        x[3] = e0 - e3;
        x[1] = e1 + e2;
        x[2] = e1 - e2;
        /* odd part */
        int t0 = s[7], t1 = s[5], t2 = s[3], t3 = s[1];
        p3 = t0 + t2;
        p4 = t1 + t3;
        p1 = t0 + t3;
        p2 = t1 + t2;
        p5 = (p3 + p4) * JF(1.175875602f);
        t0 = t0 * JF(0.298631336f);
        t1 = t1 * JF(2.053119869f);
        t2 = t2 * JF(3.072711026f);
        t3 = t3 * JF(1.501321110f);
        p1 = p5 + p1 * JF(-0.899976223f);
        p2 = p5 + p2 * JF(-2.562915447f);
        p3 = p3 * JF(-1.961570560f);
        p4 = p4 * JF(-0.390180644f);
        t[3] = t3 + p1 + p4;
        t[2] = t2 + p2 + p3;
        t[1] = t1 + p2 + p4;
        t[0] = t0 + p1 + p3;
    }
#undef JF
    static void idctBlock(uint8_t *out, int stride, const int16_t *d) {
        int tmp[64];
        for (int c = 0; c < 8; c++) {
            int s[8], x[4], t[4];
            for (int r = 0; r < 8; r++) s[r] = d[r * 8 + c];
            idct1d(s, x, t);
            for (int i = 0; i < 4; i++) x[i] += 512; /* 12-bit constants, keep 2 bits: round at bit 10 */
            tmp[0 * 8 + c] = (x[0] + t[3]) >> 10;
            tmp[7 * 8 + c] = (x[0] - t[3]) >> 10;
            tmp[1 * 8 + c] = (x[1] + t[2]) >> 10;
            tmp[6 * 8 + c] = (x[1] - t[2]) >> 10;
            tmp[2 * 8 + c] = (x[2] + t[1]) >> 10;
            tmp[5 * 8 + c] = (x[2] - t[1]) >> 10;
            tmp[3 * 8 + c] = (x[3] + t[0]) >> 10;
            tmp[4 * 8 + c] = (x[3] - t[0]) >> 10;
        }
        for (int r = 0; r < 8; r++) {
            int x[4], t[4];
            idct1d(tmp + r * 8, x, t);
            /* remove 12 + 2 + 3 bits, round, and shift the level by +128 before the shift */
            for (int i = 0; i < 4; i++) x[i] += 65536 + (128 << 17);
            uint8_t *o = out + (size_t)r * stride;
            o[0] = clamp8((x[0] + t[3]) >> 17);
            o[7] = clamp8((x[0] - t[3]) >> 17);
            o[1] = clamp8((x[1] + t[2]) >> 17);
            o[6] = clamp8((x[1] - t[2]) >> 17);
            o[2] = clamp8((x[2] + t[1]) >> 17);
            o[5] = clamp8((x[2] - t[1]) >> 17);
            o[3] = clamp8((x[3] + t[0]) >> 17);
            o[4] = clamp8((x[3] - t[0]) >> 17);
        }
    }

    /* ---- marker segments */
    bool readDQT(const uint8_t *p, int len) {
        while (len > 0) {
            int pq = p[0] >> 4, tq = p[0] & 15;
            if (pq > 1 || tq > 3) return false;
            int need = 1 + 64 * (pq ? 2 : 1);
            if (len < need) return false;
            for (int i = 0; i < 64; i++) quant[tq][kZigzag[i]] = (uint16_t)(pq ? be16(p + 1 + 2 * i) : p[1 + i]);
            quantPresent[tq] = true;
            p += need;
            len -= need;
        }
        return len == 0;
    }
    bool readDHT(const uint8_t *p, int len) {
        while (len > 0) {
            if (len < 17) return false;
            int tc = p[0] >> 4, th = p[0] & 15;
            if (tc > 1 || th > 3) return false;
            int n = 0;
            for (int i = 0; i < 16; i++) n += p[1 + i];
            if (n > 256 || len < 17 + n) return false;
            if (!(tc == 0 ? dc[th] : ac[th]).build(p + 1, p + 17, n)) return false;
            p += 17 + n;
            len -= 17 + n;
        }
        return len == 0;
    }
    bool readSOF(const uint8_t *p, int len, bool prog) {
        if (sawSOF || len < 6) return false;
        if (p[0] != 8) return false; /* 8-bit samples only */
        height = be16(p + 1);
        width = be16(p + 3);
        nComp = p[5];
        if (width <= 0 || height <= 0) return false;
        if (nComp != 1 && nComp != 3) return false;
        if (len != 6 + 3 * nComp) return false;
        progressive = prog;
        for (int i = 0; i < nComp; i++) {
            Component &c = comp[i];
            c.id = p[6 + 3 * i];
            c.h = p[7 + 3 * i] >> 4;
            c.v = p[7 + 3 * i] & 15;
            c.tq = p[8 + 3 * i];
            if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) return false;
            hmax = std::max(hmax, c.h);
            vmax = std::max(vmax, c.v);
        }
        /* all sampling factors must divide the maxima (the upsampler works in whole ratios) */
        for (int i = 0; i < nComp; i++)
            if (hmax % comp[i].h != 0 || vmax % comp[i].v != 0) return false;
        mcuX = (width + 8 * hmax - 1) / (8 * hmax);
        mcuY = (height + 8 * vmax - 1) / (8 * vmax);
        for (int i = 0; i < nComp; i++) {
            Component &c = comp[i];
            c.x = (width * c.h + hmax - 1) / hmax;
            c.y = (height * c.v + vmax - 1) / vmax;
            c.w2 = mcuX * c.h * 8;
            c.h2 = mcuY * c.v * 8;
            c.blocksW = c.w2 / 8;
            c.blocksH = c.h2 / 8;
            c.samples.assign((size_t)c.w2 * c.h2, 0);
            if (progressive) c.coeff.assign((size_t)c.blocksW * c.blocksH * 64, 0);
        }
        sawSOF = true;
        return true;
    }
    bool readSOS(const uint8_t *p, int len) {
        if (!sawSOF || len < 1) return false;
        scanN = p[0];
        if (scanN < 1 || scanN > nComp || len != 4 + 2 * scanN) return false;
        for (int i = 0; i < scanN; i++) {
            int id = p[1 + 2 * i], which = -1;
            for (int k = 0; k < nComp; k++)
                if (comp[k].id == id) which = k;
            if (which < 0) return false;
            comp[which].td = p[2 + 2 * i] >> 4;
            comp[which].ta = p[2 + 2 * i] & 15;
            if (comp[which].td > 3 || comp[which].ta > 3) return false;
            order[i] = which;
        }
        ss = p[1 + 2 * scanN];
        se = p[2 + 2 * scanN];
        ah = p[3 + 2 * scanN] >> 4;
        al = p[3 + 2 * scanN] & 15;
        if (progressive) {
            if (ss > 63 || se > 63 || ss > se || ah > 13 || al > 13) return false;
        } else {
            if (ss != 0 || ah != 0 || al != 0) return false;
            se = 63;
        }
        return true;
    }

    void restartState() {
        br.reset();
        for (int i = 0; i < nComp; i++) comp[i].dcPred = 0;
        eobrun = 0;
    }
    /* at a restart boundary: drop the padding bits and step over the RSTn marker */
    bool consumeRestart() {
        br.reset();
        if (br.marker < 0) br.fill(), br.reset();
        if (br.marker >= 0xD0 && br.marker <= 0xD7) {
            br.p += 2;
            br.marker = -1;
            restartState();
            return true;
        }
        return false; /* anything else ends the scan */
    }

    bool decodeUnit(Component &c, int bx, int by) {
        if (progressive) {
            int16_t *blk = &c.coeff[((size_t)by * c.blocksW + bx) * 64];
            if (ss == 0) return blockProgressiveDC(blk, c);
            return blockProgressiveAC(blk, ac[c.ta]);
        }
        int16_t blk[64];
        if (!quantPresent[c.tq]) return false;
        if (!blockSequential(blk, c)) return false;
        idctBlock(&c.samples[(size_t)by * 8 * c.w2 + (size_t)bx * 8], c.w2, blk);
        return true;
    }

    bool decodeScan() {
        const bool needDC = !progressive || (ss == 0 && ah == 0), needAC = !progressive || ss != 0;
        for (int i = 0; i < scanN; i++) {
            const Component &c = comp[order[i]];
            if (needDC && !dc[c.td].present) return false;
            if (needAC && !ac[c.ta].present) return false;
        }
        restartState();
        br.marker = -1;
        int todo = restartInterval ? restartInterval : 0x7fffffff;
        if (scanN == 1) {
            /* non-interleaved: the component's own block grid, not the MCU-padded one */
            Component &c = comp[order[0]];
            int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
            for (int by = 0; by < h; by++)
                for (int bx = 0; bx < w; bx++) {
                    if (!decodeUnit(c, bx, by) || br.starved > 16) return false;
                    if (--todo <= 0) {
                        if (!consumeRestart()) return true;
                        todo = restartInterval;
                    }
                }
            return true;
        }
        if (progressive && ss != 0) return false; /* AC scans are never interleaved */
        for (int my = 0; my < mcuY; my++)
            for (int mx = 0; mx < mcuX; mx++) {
                for (int i = 0; i < scanN; i++) {
                    Component &c = comp[order[i]];
                    for (int v = 0; v < c.v; v++)
                        for (int h = 0; h < c.h; h++)
                            if (!decodeUnit(c, mx * c.h + h, my * c.v + v)) return false;
                }
                if (br.starved > 16) return false; /* ran off the end of the file inside the entropy-coded data */
                if (--todo <= 0) {
                    if (!consumeRestart()) return true;
                    todo = restartInterval;
                }
            }
        return true;
    }

    void finishProgressive() {
        for (int i = 0; i < nComp; i++) {
            Component &c = comp[i];
            int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
            const uint16_t *q = quant[c.tq];
            for (int by = 0; by < h; by++)
                for (int bx = 0; bx < w; bx++) {
                    int16_t *blk = &c.coeff[((size_t)by * c.blocksW + bx) * 64];
                    for (int k = 0; k < 64; k++) blk[k] = (int16_t)(blk[k] * q[k]);
                    idctBlock(&c.samples[(size_t)by * 8 * c.w2 + (size_t)bx * 8], c.w2, blk);
                }
        }
    }

    bool parse() {
        const uint8_t *p = data;
        if (end - p < 4 || p[0] != 0xFF || p[1] != 0xD8) return false;
        p += 2;
        bool sawScan = false;
        while (true) {
            /* next marker */
            while (p < end && *p != 0xFF) ++p;
            while (p < end && *p == 0xFF) ++p;
            if (p >= end) break;
            int m = *p++;
            if (m == 0xD9) break; /* EOI */
            if (m == 0x00 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
            if (end - p < 2) return false;
            int len = be16(p) - 2;
            if (len < 0 || end - (p + 2) < len) return false;
            const uint8_t *seg = p + 2;
            p = seg + len;
            switch (m) {
                case 0xDB:
                    if (!readDQT(seg, len)) return false;
                    break;
                case 0xC4:
                    if (!readDHT(seg, len)) return false;
                    break;
                case 0xC0:
                case 0xC1:
                    if (!readSOF(seg, len, false)) return false;
                    break;
                case 0xC2:
                    if (!readSOF(seg, len, true)) return false;
                    break;
                case 0xDD:
                    if (len != 2) return false;
                    restartInterval = be16(seg);
                    break;
                case 0xE0:
                    if (len >= 5 && std::memcmp(seg, "JFIF", 5) == 0) jfif = true;
                    break;
                case 0xEE:
                    if (len >= 12 && std::memcmp(seg, "Adobe", 5) == 0) adobeTransform = seg[11];
                    break;
                case 0xDA: {
                    if (!readSOS(seg, len)) return false;
                    br = BitReader();
                    br.p = p;
                    br.end = end;
                    if (!decodeScan()) return false;
                    sawScan = true;
                    p = br.p; /* entropy data never contains an unstuffed marker: the scan loop finds the next one */
                    break;
                }
                default:
                    /* arithmetic coding, lossless, hierarchical, 12-bit: not supported */
                    if ((m >= 0xC3 && m <= 0xCF) && m != 0xC4 && m != 0xC8 && m != 0xCC) return false;
                    break; /* APPn, COM and others are skipped */
            }
        }
        if (!sawSOF || !sawScan) return false;
        if (progressive) {
            for (int i = 0; i < nComp; i++)
                if (!quantPresent[comp[i].tq]) return false;
            finishProgressive();
        }
        return true;
    }

    /* ---- upsampling rows */
    static const uint8_t *rowCopy(uint8_t *, const uint8_t *nearRow, const uint8_t *, int, int) { return nearRow; }
    static const uint8_t *rowV2(uint8_t *out, const uint8_t *n, const uint8_t *f, int w, int) {
        for (int i = 0; i < w; i++) out[i] = (uint8_t)((3 * n[i] + f[i] + 2) >> 2);
        return out;
    }
    static const uint8_t *rowH2(uint8_t *out, const uint8_t *in, const uint8_t *, int w, int) {
        if (w == 1) {
            out[0] = out[1] = in[0];
            return out;
        }
        out[0] = in[0];
        out[1] = (uint8_t)((in[0] * 3 + in[1] + 2) >> 2);
        int i;
        for (i = 1; i < w - 1; i++) {
            int n = 3 * in[i] + 2;
            out[i * 2 + 0] = (uint8_t)((n + in[i - 1]) >> 2);
            out[i * 2 + 1] = (uint8_t)((n + in[i + 1]) >> 2);
        }
        out[i * 2 + 0] = (uint8_t)((in[w - 2] * 3 + in[w - 1] + 2) >> 2);
        out[i * 2 + 1] = in[w - 1];
        return out;
    }
    static const uint8_t *rowHV2(uint8_t *out, const uint8_t *n, const uint8_t *f, int w, int) {
        if (w == 1) {
            out[0] = out[1] = (uint8_t)((3 * n[0] + f[0] + 2) >> 2);
            return out;
        }
        int t1 = 3 * n[0] + f[0];
        out[0] = (uint8_t)((t1 + 2) >> 2);
        for (int i = 1; i < w; i++) {
            int t0 = t1;
            t1 = 3 * n[i] + f[i];
            out[i * 2 - 1] = (uint8_t)((3 * t0 + t1 + 8) >> 4);
            out[i * 2] = (uint8_t)((3 * t1 + t0 + 8) >> 4);
        }
        out[w * 2 - 1] = (uint8_t)((t1 + 2) >> 2);
        return out;
    }
    static const uint8_t *rowGeneric(uint8_t *out, const uint8_t *n, const uint8_t *, int w, int hs) {
        for (int i = 0; i < w; i++)
            for (int j = 0; j < hs; j++) out[i * hs + j] = n[i];
        return out;
    }

    /* produces RGBA8 rows, top row first */
    bool output(ImageU8 &img, int *srcChannels) {
        typedef const uint8_t *(*RowFn)(uint8_t *, const uint8_t *, const uint8_t *, int, int);
        struct Resampler {
            RowFn fn;
            const uint8_t *line0, *line1;
            int hs, vs, wLores, ystep, ypos;
            std::vector<uint8_t> buf;
        } rs[3];
        for (int k = 0; k < nComp; k++) {
            Resampler &r = rs[k];
            r.hs = hmax / comp[k].h;
            r.vs = vmax / comp[k].v;
            r.ystep = r.vs >> 1;
            r.wLores = (width + r.hs - 1) / r.hs;
            r.ypos = 0;
            r.line0 = r.line1 = comp[k].samples.data();
            r.buf.resize((size_t)width + 2 * r.hs + 16);
            if (r.hs == 1 && r.vs == 1) r.fn = rowCopy;
            else if (r.hs == 1 && r.vs == 2) r.fn = rowV2;
            else if (r.hs == 2 && r.vs == 1) r.fn = rowH2;
            else if (r.hs == 2 && r.vs == 2) r.fn = rowHV2;
            else r.fn = rowGeneric;
        }
        /* three components are RGB only when their ids spell it or an Adobe segment says "no transform" outside JFIF */
        const bool idsRGB = nComp == 3 && comp[0].id == 'R' && comp[1].id == 'G' && comp[2].id == 'B';
        const bool isRGB = nComp == 3 && (idsRGB || (adobeTransform == 0 && !jfif));
        img.width = width;
        img.height = height;
        img.channels = 4;
        img.data.resize((size_t)width * height * 4);
        if (srcChannels) *srcChannels = nComp >= 3 ? 3 : 1;
        const int crR = ((int)(1.40200f * 4096.0f + 0.5f)) << 8, crG = -(((int)(0.71414f * 4096.0f + 0.5f)) << 8);
        const int cbG = -(((int)(0.34414f * 4096.0f + 0.5f)) << 8), cbB = ((int)(1.77200f * 4096.0f + 0.5f)) << 8;
        for (int j = 0; j < height; j++) {
            const uint8_t *row[3] = {nullptr, nullptr, nullptr};
            for (int k = 0; k < nComp; k++) {
                Resampler &r = rs[k];
                bool bottom = r.ystep >= (r.vs >> 1);
                row[k] = r.fn(r.buf.data(), bottom ? r.line1 : r.line0, bottom ? r.line0 : r.line1, r.wLores, r.hs);
                if (++r.ystep >= r.vs) {
                    r.ystep = 0;
                    r.line0 = r.line1;
                    if (++r.ypos < comp[k].y) r.line1 += comp[k].w2;
                }
            }
            uint8_t *out = &img.data[(size_t)j * width * 4];
            if (nComp == 1) {
                for (int i = 0; i < width; i++) {
                    out[4 * i] = out[4 * i + 1] = out[4 * i + 2] = row[0][i];
                    out[4 * i + 3] = 255;
                }
            } else if (isRGB) {
                for (int i = 0; i < width; i++) {
                    out[4 * i] = row[0][i];
                    out[4 * i + 1] = row[1][i];
                    out[4 * i + 2] = row[2][i];
                    out[4 * i + 3] = 255;
                }
            } else {
                for (int i = 0; i < width; i++) {
                    int yf = (row[0][i] << 20) + (1 << 19);
                    int cb = row[1][i] - 128, cr = row[2][i] - 128;
                    int r = yf + cr * crR;
                    int g = yf + cr * crG + (int)(((unsigned)(cb * cbG)) & 0xffff0000u);
                    int b = yf + cb * cbB;
                    out[4 * i] = clamp8(r >> 20);
                    out[4 * i + 1] = clamp8(g >> 20);
                    out[4 * i + 2] = clamp8(b >> 20);
                    out[4 * i + 3] = 255;
                }
            }
        }
        return true;
    }
};

}  // namespace

bool decodeJPEG(const uint8_t *bytes, size_t n, ImageU8 &out, int *srcChannels) {
    if (n < 4) return false;
    Decoder d(bytes, n);
    if (!d.parse()) return false;
    return d.output(out, srcChannels);
}

}  // namespace vengine
