/*
 * ORACLE — test infrastructure only (see omath.hpp).
 *
 * (1) World-space triangle soup + ray/triangle test + brute-force and median-split-BVH queries used by
 *     the CPU estimator.  In the reference this work is done by the Vulkan driver / RT cores
 *     (VulkanAccelerationStructure.cpp:137,258; traceRayEXT at raygen.rgen.glsl:110), so the hit rule is
 *     ours to define: smallest (t, world triangle id) with tmin < t < tmax, no face culling
 *     (VulkanScene.cpp:364-377 disables culling).
 * (2) The CPU reference LBVH build (Morton codes -> stable sort -> Karras hierarchy -> AABB fit) that the
 *     device build must match bit for bit.
 */
#pragma once
#include <vector>
#include <cstdint>
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <functional>
#include "omath.hpp"

namespace orc {

struct WorldTri {
    vec3 v0, e1, e2; /* e1 = v1 - v0, e2 = v2 - v0 */
    uint32_t inst, prim;
};

struct Hit {
    float t, u, v;
    int32_t tri; /* world triangle id, -1 = miss */
};

/* Moeller-Trumbore; identical operation order to the device kernel (vviewer_b200/csrc/traverse.cuh) */
inline bool intersectTri(const WorldTri &T, vec3 o, vec3 d, float tmin, float tmax, float &t, float &u, float &v) {
    vec3 p = cross(d, T.e2);
    float det = dot(T.e1, p);
    if (det == 0.0f) return false;
    float inv = 1.0f / det;
    vec3 s = o - T.v0;
    u = dot(s, p) * inv;
    if (u < 0.0f || u > 1.0f) return false;
    vec3 q = cross(s, T.e1);
    v = dot(d, q) * inv;
    if (v < 0.0f || u + v > 1.0f) return false;
    t = dot(T.e2, q) * inv;
    return t > tmin && t < tmax;
}

struct AABB {
    vec3 lo, hi;
    AABB() : lo(FLT_MAX), hi(-FLT_MAX) {}
    void grow(vec3 p) {
        lo = vmin(lo, p);
        hi = vmax(hi, p);
    }
    void grow(const AABB &b) {
        lo = vmin(lo, b.lo);
        hi = vmax(hi, b.hi);
    }
};

inline AABB triBounds(const WorldTri &T) {
    AABB b;
    b.grow(T.v0);
    b.grow(T.v0 + T.e1);
    b.grow(T.v0 + T.e2);
    return b;
}

/* ------------------------------------------------------------------ oracle's own query BVH */
struct QNode {
    AABB box;
    int32_t left, right; /* children, or (first, -count) for leaves */
};

struct QueryBVH {
    const std::vector<WorldTri> *tris = nullptr;
    std::vector<QNode> nodes;
    std::vector<uint32_t> ids;
    std::vector<vec3> cent;

    void build(const std::vector<WorldTri> &t) {
        tris = &t;
        nodes.clear();
        ids.resize(t.size());
        cent.resize(t.size());
        for (size_t i = 0; i < t.size(); i++) {
            ids[i] = (uint32_t)i;
            AABB b = triBounds(t[i]);
            cent[i] = (b.lo + b.hi) * 0.5f;
        }
        if (!t.empty()) {
            nodes.reserve(t.size() * 2);
            rec(0, (int)t.size());
        }
    }
    int rec(int a, int b) {
        int me = (int)nodes.size();
        nodes.push_back(QNode());
        AABB box, cb;
        for (int i = a; i < b; i++) {
            box.grow(triBounds((*tris)[ids[i]]));
            cb.grow(cent[ids[i]]);
        }
        nodes[me].box = box;
        if (b - a <= 4) {
            nodes[me].left = a;
            nodes[me].right = -(b - a);
            return me;
        }
        vec3 ext = cb.hi - cb.lo;
        int ax = ext.x > ext.y ? (ext.x > ext.z ? 0 : 2) : (ext.y > ext.z ? 1 : 2);
        float mid = 0.5f * (cb.lo[ax] + cb.hi[ax]);
        auto it = std::partition(ids.begin() + a, ids.begin() + b, [&](uint32_t id) { return cent[id][ax] < mid; });
        int m = (int)(it - ids.begin());
        if (m == a || m == b) {
            m = (a + b) / 2;
            std::nth_element(ids.begin() + a, ids.begin() + m, ids.begin() + b,
                             [&](uint32_t x, uint32_t y) { return cent[x][ax] < cent[y][ax] || (cent[x][ax] == cent[y][ax] && x < y); });
        }
        int l = rec(a, m);
        int r = rec(m, b);
        nodes[me].left = l;
        nodes[me].right = r;
        return me;
    }
    static bool hitBox(const AABB &b, vec3 o, vec3 inv, float tmin, float tmax) {
        /* conservative slab test (slightly widened) so the BVH never loses a hit brute force would find */
        float t0 = tmin, t1 = tmax;
        for (int a = 0; a < 3; a++) {
            float ta = (b.lo[a] - o[a]) * inv[a];
            float tb = (b.hi[a] - o[a]) * inv[a];
            if (ta > tb) std::swap(ta, tb);
            if (ta != ta || tb != tb) continue; /* 0 * inf */
            t0 = std::max(t0, ta * (1.0f - 4e-7f) - 1e-7f);
            t1 = std::min(t1, tb * (1.0f + 4e-7f) + 1e-7f);
        }
        return t0 <= t1;
    }
    /* closest hit: smallest (t, tri id) */
    Hit closest(vec3 o, vec3 d, float tmin, float tmax) const {
        Hit best{tmax, 0, 0, -1};
        if (nodes.empty()) return best;
        vec3 inv(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        int stack[128];
        int sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const QNode &n = nodes[stack[--sp]];
            /* <= so that equal-t candidates are still visited for the id tie-break */
            if (!hitBox(n.box, o, inv, tmin, best.t)) continue;
            if (n.right < 0) {
                for (int i = 0; i < -n.right; i++) {
                    uint32_t id = ids[n.left + i];
                    float t, u, v;
                    if (intersectTri((*tris)[id], o, d, tmin, tmax, t, u, v)) {
                        if (t < best.t || (t == best.t && (best.tri < 0 || (int32_t)id < best.tri))) best = Hit{t, u, v, (int32_t)id};
                    }
                }
            } else {
                stack[sp++] = n.left;
                stack[sp++] = n.right;
            }
        }
        return best;
    }
    /* all hits in (tmin, tmax), sorted by (t, tri id) */
    void all(vec3 o, vec3 d, float tmin, float tmax, std::vector<Hit> &out) const {
        out.clear();
        if (nodes.empty()) return;
        vec3 inv(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        int stack[128];
        int sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const QNode &n = nodes[stack[--sp]];
            if (!hitBox(n.box, o, inv, tmin, tmax)) continue;
            if (n.right < 0) {
                for (int i = 0; i < -n.right; i++) {
                    uint32_t id = ids[n.left + i];
                    float t, u, v;
                    if (intersectTri((*tris)[id], o, d, tmin, tmax, t, u, v)) out.push_back(Hit{t, u, v, (int32_t)id});
                }
            } else {
                stack[sp++] = n.left;
                stack[sp++] = n.right;
            }
        }
        std::sort(out.begin(), out.end(), [](const Hit &a, const Hit &b) { return a.t < b.t || (a.t == b.t && a.tri < b.tri); });
    }
};

inline Hit bruteClosest(const std::vector<WorldTri> &tris, vec3 o, vec3 d, float tmin, float tmax) {
    Hit best{tmax, 0, 0, -1};
    for (size_t i = 0; i < tris.size(); i++) {
        float t, u, v;
        if (intersectTri(tris[i], o, d, tmin, tmax, t, u, v)) {
            if (t < best.t || (t == best.t && (best.tri < 0 || (int32_t)i < best.tri))) best = Hit{t, u, v, (int32_t)i};
        }
    }
    return best;
}

/* ------------------------------------------------------------------ reference LBVH build */
inline uint64_t expandBits21(uint64_t v) {
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}
inline uint64_t expandBits10(uint64_t v) {
    v &= 0x3ffull;
    v = (v * 0x00010001ull) & 0xFF0000FFull;
    v = (v * 0x00000101ull) & 0x0F00F00Full;
    v = (v * 0x00000011ull) & 0xC30C30C3ull;
    v = (v * 0x00000005ull) & 0x49249249ull;
    return v;
}

/* number of Morton bits per axis as a function of the triangle count (both sides must agree) */
/* 30-bit codes up to 65 536 primitives, 48-bit (16 per axis: cells of 1 / 65 536 of the scene, six 8-bit sort passes) up to 2^26, the
 * full 63 bits beyond */
inline int mortonBitsPerAxis(uint64_t nTris) {
    if (const char *e = getenv("PTC_MORTON_BITS")) { /* tests: force a tier (10, 16 or 21) on both sides */
        const int b = atoi(e);
        if (b == 10 || b == 16 || b == 21) return b;
    }
    return nTris <= 65536ull ? 10 : (nTris <= (1ull << 26) ? 16 : 21);
}

struct LBVH {
    uint64_t n = 0;
    AABB scene;
    std::vector<uint64_t> morton; /* sorted */
    std::vector<uint32_t> order;  /* sorted position -> world triangle id */
    std::vector<int32_t> parent, left, right; /* 2n-1 nodes; internal 0..n-2, leaf k -> n-1+k */
    std::vector<uint32_t> subCount;           /* triangles below each internal node */
    int32_t root = 0;                         /* Karras: 0; PLOC: n - 2 (the last node created); single triangle: leaf 0 */
    uint32_t hierarchy = 1, plocRadius = 16;  /* PTC_HIERARCHY_PLOC by default, like the device build */
    std::vector<AABB> box;

    static inline int clz64(uint64_t x) { return x ? __builtin_clzll(x) : 64; }
    static inline int clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }
    int delta(int64_t i, int64_t j) const {
        if (j < 0 || j >= (int64_t)n) return -1;
        uint64_t a = morton[i], b = morton[j];
        if (a == b) return 64 + clz32((uint32_t)i ^ (uint32_t)j);
        return clz64(a ^ b);
    }

    void build(const std::vector<WorldTri> &tris) {
        std::vector<AABB> tb(tris.size());
        for (size_t i = 0; i < tris.size(); i++) tb[i] = triBounds(tris[i]);
        buildBounds(tb);
    }
    /* the build only ever looks at the primitives' boxes: triangles (single level, bottom level) or instances (top level) */
    void buildBounds(const std::vector<AABB> &tb) {
        n = tb.size();
        morton.assign(n, 0);
        order.resize(n);
        if (n == 0) return;
        scene = AABB();
        for (uint64_t i = 0; i < n; i++) scene.grow(tb[i]);
        vec3 ext = scene.hi - scene.lo;
        vec3 inv(ext.x > 0 ? 1.0f / ext.x : 0.0f, ext.y > 0 ? 1.0f / ext.y : 0.0f, ext.z > 0 ? 1.0f / ext.z : 0.0f);
        int bits = mortonBitsPerAxis(n);
        float scale = (float)(1u << bits);
        float qmax = scale - 1.0f;
        std::vector<uint64_t> code(n);
        for (uint64_t i = 0; i < n; i++) {
            vec3 c = (tb[i].lo + tb[i].hi) * 0.5f;
            vec3 q = (c - scene.lo) * inv;
            uint64_t qx = (uint64_t)(uint32_t)std::min(std::max(q.x * scale, 0.0f), qmax);
            uint64_t qy = (uint64_t)(uint32_t)std::min(std::max(q.y * scale, 0.0f), qmax);
            uint64_t qz = (uint64_t)(uint32_t)std::min(std::max(q.z * scale, 0.0f), qmax);
            if (bits == 10)
                code[i] = (expandBits10(qx) << 2) | (expandBits10(qy) << 1) | expandBits10(qz);
            else
                code[i] = (expandBits21(qx) << 2) | (expandBits21(qy) << 1) | expandBits21(qz);
            order[i] = (uint32_t)i;
        }
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return code[a] < code[b]; });
        for (uint64_t i = 0; i < n; i++) morton[i] = code[order[i]];

        uint64_t nn = 2 * n - 1;
        parent.assign(nn, -1);
        left.assign(nn, -1);
        right.assign(nn, -1);
        box.assign(nn, AABB());
        subCount.assign(n, 0);
        root = 0;
        for (uint64_t k = 0; k < n; k++) box[n - 1 + k] = tb[order[k]];
        if (hierarchy == 1 && n > 1) {
            ploc();
            return;
        }
        for (int64_t i = 0; i + 1 < (int64_t)n; i++) {
            /* Karras 2012, "Maximizing parallelism in the construction of BVHs, octrees and k-d trees" */
            int d = (delta(i, i + 1) - delta(i, i - 1)) >= 0 ? 1 : -1;
            int dmin = delta(i, i - d);
            int64_t lmax = 2;
            while (delta(i, i + lmax * d) > dmin) lmax *= 2;
            int64_t l = 0;
            for (int64_t t = lmax / 2; t >= 1; t /= 2)
                if (delta(i, i + (l + t) * d) > dmin) l += t;
            int64_t j = i + l * d;
            int dnode = delta(i, j);
            int64_t s = 0;
            int64_t t = l;
            do {
                t = (t + 1) / 2;
                if (delta(i, i + (s + t) * d) > dnode) s += t;
            } while (t > 1);
            int64_t gamma = i + s * d + std::min(d, 0);
            int64_t lo = std::min(i, j), hi = std::max(i, j);
            int32_t L = (lo == gamma) ? (int32_t)(n - 1 + gamma) : (int32_t)gamma;
            int32_t R = (hi == gamma + 1) ? (int32_t)(n - 1 + gamma + 1) : (int32_t)(gamma + 1);
            left[i] = L;
            right[i] = R;
            parent[L] = (int32_t)i;
            parent[R] = (int32_t)i;
            subCount[i] = (uint32_t)(hi - lo + 1);
        }
        if (n > 1) fit(0);
    }
    static float mergedHalfArea(const AABB &a, const AABB &b) {
        AABB m = a;
        m.grow(b);
        float ex = m.hi.x - m.lo.x, ey = m.hi.y - m.lo.y, ez = m.hi.z - m.lo.z;
        float p = ex * ey;
        float q = ey * ez;
        float r = ez * ex;
        return (p + q) + r;
    }
    /* Parallel locally-ordered clustering (Meister & Bittner 2018), restated sequentially round by round exactly like
     * vviewer_b200/csrc/lbvh.cuh: nearest neighbour by merged half-area within +-radius positions of the Morton-ordered
     * cluster list (ties: smaller position), mutual pairs merge into a node numbered in position order, list compacted. */
    void ploc() {
        const int64_t r = (int64_t)std::min<uint32_t>(std::max<uint32_t>(plocRadius, 1u), 32u);
        std::vector<int32_t> cid(n), cidNext;
        for (uint64_t k = 0; k < n; k++) cid[k] = (int32_t)(n - 1 + k);
        uint32_t nextNode = 0;
        auto tris = [&](int32_t node) { return node >= (int32_t)(n - 1) ? 1u : subCount[node]; };
        uint32_t rounds = 0;
        while (cid.size() > 1) {
            const int64_t c = (int64_t)cid.size();
            rounds++;
            std::vector<int64_t> nn(c);
            for (int64_t i = 0; i < c; i++) {
                float bestA = 0.0f;
                int64_t best = -1;
                for (int64_t j = std::max<int64_t>(0, i - r); j <= std::min<int64_t>(c - 1, i + r); j++) {
                    if (j == i) continue;
                    float a = mergedHalfArea(box[cid[i]], box[cid[j]]);
                    if (best < 0 || a < bestA) {
                        bestA = a;
                        best = j;
                    }
                }
                nn[i] = best;
            }
            cidNext.clear();
            for (int64_t i = 0; i < c; i++) {
                const int64_t j = nn[i];
                const bool mutual = nn[j] == i;
                if (mutual && i > j) continue; /* merged into its partner */
                if (mutual) {
                    const int32_t id = (int32_t)nextNode++;
                    const int32_t L = cid[i], R = cid[j];
                    left[id] = L;
                    right[id] = R;
                    parent[L] = id;
                    parent[R] = id;
                    subCount[id] = tris(L) + tris(R);
                    AABB b = box[L];
                    b.grow(box[R]);
                    box[id] = b;
                    cidNext.push_back(id);
                } else {
                    cidNext.push_back(cid[i]);
                }
            }
            cid.swap(cidNext);
        }
        root = (int32_t)(n - 2);
        if (getenv("PTC_VERBOSE")) fprintf(stderr, "[oracle] PLOC: %llu triangles, %u rounds\n", (unsigned long long)n, rounds);
    }
    void fit(int32_t root) {
        /* iterative post-order */
        std::vector<int32_t> st;
        std::vector<int32_t> ord;
        st.push_back(root);
        while (!st.empty()) {
            int32_t x = st.back();
            st.pop_back();
            ord.push_back(x);
            if (x < (int32_t)(n - 1)) {
                st.push_back(left[x]);
                st.push_back(right[x]);
            }
        }
        for (size_t k = ord.size(); k-- > 0;) {
            int32_t x = ord[k];
            if (x < (int32_t)(n - 1)) {
                AABB b = box[left[x]];
                b.grow(box[right[x]]);
                box[x] = b;
            }
        }
    }
};


/* ------------------------------------------------------------------ reference collapse to the 8-wide compressed BVH
 * CPU restatement of the device collapse (vviewer_b200/csrc/lbvh.cuh, which documents the 80-byte node layout and the
 * child selection / slot assignment rules); sequential, breadth first through a FIFO, which yields the same numbering as
 * the device's level-by-level passes.  Float arithmetic is plain IEEE single (this file is compiled with
 * -ffp-contract=off), matching the device's explicit round-to-nearest intrinsics. */
struct WideBVH {
    std::vector<uint32_t> words;    /* 20 per node */
    std::vector<uint32_t> triOrder; /* traversal position -> world triangle id */
    uint64_t nNodes = 0;

    static float bitsToFloat(uint32_t b) {
        float f;
        memcpy(&f, &b, 4);
        return f;
    }
    static uint32_t floatToBits(float f) {
        uint32_t b;
        memcpy(&b, &f, 4);
        return b;
    }
    static float halfArea(const AABB &b) {
        float ex = b.hi.x - b.lo.x, ey = b.hi.y - b.lo.y, ez = b.hi.z - b.lo.z;
        float a = ex * ey;
        float c = ey * ez;
        float e = ez * ex;
        return (a + c) + e;
    }

    void build(const LBVH &L) {
        words.clear();
        triOrder.clear();
        nNodes = 0;
        const int64_t n = (int64_t)L.n;
        if (n == 0) return;
        triOrder.resize(n);
        auto subTris = [&](int32_t node) -> uint32_t { return node >= n - 1 ? 1u : L.subCount[node]; };
        /* sorted positions of the triangles below a small subtree, left to right */
        std::function<void(int32_t, std::vector<uint32_t> &)> subLeaves = [&](int32_t node, std::vector<uint32_t> &out) {
            if (node >= n - 1) {
                out.push_back((uint32_t)(node - (n - 1)));
                return;
            }
            subLeaves(L.left[node], out);
            subLeaves(L.right[node], out);
        };
        const uint32_t LEAF = 3;
        std::vector<int32_t> fifo; /* binary root of wide node k */
        fifo.push_back(n == 1 ? 0 : L.root);
        uint32_t triBase = 0;
        for (size_t id = 0; id < fifo.size(); id++) {
            const int32_t root = fifo[id];
            /* ---- children */
            std::vector<int32_t> list{root};
            for (int phase = 0; phase < 2; phase++) {
                const uint32_t thr = phase == 0 ? LEAF : 1u;
                while (list.size() < 8) {
                    int bi = -1;
                    float ba = -1.0f;
                    for (size_t i = 0; i < list.size(); i++) {
                        float a = halfArea(L.box[list[i]]);
                        if (subTris(list[i]) > thr && a > ba) {
                            ba = a;
                            bi = (int)i;
                        }
                    }
                    if (bi < 0) break;
                    int32_t c = list[bi];
                    list[bi] = L.left[c];
                    list.push_back(L.right[c]);
                }
            }
            /* ---- slots */
            const AABB &rb = L.box[root];
            const float ncx = (rb.lo.x + rb.hi.x) * 0.5f, ncy = (rb.lo.y + rb.hi.y) * 0.5f, ncz = (rb.lo.z + rb.hi.z) * 0.5f;
            const int len = (int)list.size();
            float vx[8], vy[8], vz[8];
            for (int i = 0; i < len; i++) {
                const AABB &b = L.box[list[i]];
                vx[i] = (b.lo.x + b.hi.x) * 0.5f - ncx;
                vy[i] = (b.lo.y + b.hi.y) * 0.5f - ncy;
                vz[i] = (b.lo.z + b.hi.z) * 0.5f - ncz;
            }
            int32_t slotChild[8];
            for (int s = 0; s < 8; s++) slotChild[s] = -1;
            uint32_t childDone = 0, slotDone = 0;
            for (int it = 0; it < len; it++) {
                int bc = -1, bs = -1;
                float bcost = 0.0f;
                for (int c = 0; c < len; c++) {
                    if (childDone & (1u << c)) continue;
                    for (int sl = 0; sl < 8; sl++) {
                        if (slotDone & (1u << sl)) continue;
                        float cost = (((sl & 4) ? vx[c] : -vx[c]) + ((sl & 2) ? vy[c] : -vy[c])) + ((sl & 1) ? vz[c] : -vz[c]);
                        if (bc < 0 || cost > bcost) {
                            bcost = cost;
                            bc = c;
                            bs = sl;
                        }
                    }
                }
                childDone |= 1u << bc;
                slotDone |= 1u << bs;
                slotChild[bs] = list[bc];
            }
            /* ---- quantisation frame */
            const float plo[3] = {rb.lo.x, rb.lo.y, rb.lo.z}, phi[3] = {rb.hi.x, rb.hi.y, rb.hi.z};
            uint32_t e[3];
            float scale[3], inv[3];
            for (int a = 0; a < 3; a++) {
                const float extent = phi[a] - plo[a];
                const float sdiv = extent / 255.0f;
                const uint32_t b = floatToBits(sdiv);
                uint32_t ee = (b >> 23) & 0xffu;
                if (b & 0x7fffffu) ee++;
                ee = std::min(std::max(ee, 1u), 253u);
                while (ee < 253u && extent * bitsToFloat((254u - ee) << 23) > 255.0f) ee++;
                e[a] = ee;
                scale[a] = bitsToFloat(ee << 23);
                inv[a] = bitsToFloat((254u - ee) << 23);
            }
            /* ---- record */
            uint32_t meta[8], qlo[3][8], qhi[3][8];
            uint32_t imask = 0, off = 0;
            const uint32_t childBase = (uint32_t)fifo.size();
            for (int sl = 0; sl < 8; sl++) {
                const int32_t c = slotChild[sl];
                if (c < 0) {
                    meta[sl] = 0;
                    for (int a = 0; a < 3; a++) {
                        qlo[a][sl] = 255u;
                        qhi[a][sl] = 0u;
                    }
                    continue;
                }
                const AABB &cb = L.box[c];
                const float clo[3] = {cb.lo.x, cb.lo.y, cb.lo.z}, chi[3] = {cb.hi.x, cb.hi.y, cb.hi.z};
                for (int a = 0; a < 3; a++) {
                    float ql = std::floor((clo[a] - plo[a]) * inv[a]);
                    ql = std::min(std::max(ql, 0.0f), 255.0f);
                    while (ql > 0.0f && plo[a] + ql * scale[a] > clo[a]) ql -= 1.0f;
                    float qh = std::ceil((chi[a] - plo[a]) * inv[a]);
                    qh = std::min(std::max(qh, 0.0f), 255.0f);
                    while (qh < 255.0f && plo[a] + qh * scale[a] < chi[a]) qh += 1.0f;
                    qlo[a][sl] = (uint32_t)ql;
                    qhi[a][sl] = (uint32_t)qh;
                }
                const uint32_t ct = subTris(c);
                if (ct > LEAF) {
                    meta[sl] = 0x20u | (24u + (uint32_t)sl);
                    imask |= 1u << sl;
                    fifo.push_back(c);
                } else {
                    meta[sl] = (((1u << ct) - 1u) << 5) | off;
                    std::vector<uint32_t> leaves;
                    subLeaves(c, leaves);
                    for (uint32_t j = 0; j < ct; j++) triOrder[triBase + off + j] = L.order[leaves[j]];
                    off += ct;
                }
            }
            auto pack4 = [](const uint32_t *b) { return b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24); };
            uint32_t w[20];
            w[0] = floatToBits(rb.lo.x);
            w[1] = floatToBits(rb.lo.y);
            w[2] = floatToBits(rb.lo.z);
            w[3] = e[0] | (e[1] << 8) | (e[2] << 16) | (imask << 24);
            w[4] = childBase;
            w[5] = triBase;
            w[6] = pack4(meta);
            w[7] = pack4(meta + 4);
            for (int a = 0; a < 3; a++) {
                w[8 + 2 * a] = pack4(qlo[a]);
                w[9 + 2 * a] = pack4(qlo[a] + 4);
                w[14 + 2 * a] = pack4(qhi[a]);
                w[15 + 2 * a] = pack4(qhi[a] + 4);
            }
            words.insert(words.end(), w, w + 20);
            triBase += off;
        }
        nNodes = fifo.size();
    }
};

}  // namespace orc
