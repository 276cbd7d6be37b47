/*
 * libptc_cuda.so — the product: include/ptc.h implemented with hand-written sm_100a CUDA.
 * No CPU fallback exists; every entry point fails loudly when CUDA is unavailable.
 *
 * Host side of the hot path = what VulkanRendererPathTracing::render() does below the scene model
 * (src/lib/vengine/vulkan/renderers/VulkanRendererPathTracing.cpp:121-226, 791-956): upload scene records,
 * (re)build the acceleration structure, loop over batches, read the three RGBA32F targets back.
 */
#include "common.cuh"
#include "bsdf.cuh"
#include "lbvh.cuh"
#include "traverse.cuh"
#include "wavefront.cuh"

#include <array>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <system_error>
#include <thread>
#include <vector>

#include <dlfcn.h>
#include <nccl.h> /* types and prototypes only: the library is opened at run time (NcclApi) */

namespace {

/* all textures of one (width, height, sRGB) class live in one layered array behind one texture object */
struct TextureClass {
    uint32_t width = 0, height = 0, srgb = 0, layers = 0;
    cudaArray_t array = nullptr;
    cudaTextureObject_t tex = 0;
};

}  // namespace

/* everything one GPU holds: scene, acceleration structure, wavefront state, streams */
struct Dev {
    int device = 0;
    cudaStream_t stream = nullptr;
    int smCount = 148;
    wf::ExtendTune tune{EXTEND_MIN_ACTIVE, EXTEND_TRI_ENTER, EXTEND_TRI_LEAVE, EXTEND_BLOCKED};
    size_t persistMax = 0, windowMax = 0; /* L2 persisting carve-out and access-policy window limits of the device */

    /* scene */
    DBuf<ptc_vertex> vertices;
    DBuf<uint32_t> indices;
    DBuf<DInstance> instances;
    DBuf<ptc_material> materials;
    DBuf<ptc_light_data> lightData;
    DBuf<ptc_light_instance> lightInstances;
    DBuf<cudaTextureObject_t> texClassTable;
    DBuf<uint32_t> texRef;
    std::vector<TextureClass> texClasses;
    cudaArray_t cubeArray = nullptr;
    cudaTextureObject_t cubeTex = 0;
    uint32_t cubeN = 0;
    DBuf<float> envCdfV, envCdfU; /* PTC_FLAG_ENV_IMPORTANCE tables, valid while cubeTex is */
    DBuf<float> pmjTable, blueTable; /* ptc_set_sampler_tables (PTC_FLAG_SAMPLER_PMJ) */
    DBuf<float4> emissiveBoxes;   /* DScene::emissiveBoxes */
    uint32_t nEmissiveBoxes = 0;
    uint32_t nInstances = 0, nMaterials = 0, nLightInstances = 0, nTextures = 0, nWorldTris = 0;
    bool anyEmissive = false, anyTransparent = false, anyVolumeChange = false;
    bool sceneUploaded = false, accelBuilt = false;
    lbvh::Build accel;
    /* two-level alternative (one tree per mesh + one over the instances), for heavily instanced scenes */
    lbvh::TwoLevel accel2;
    std::vector<lbvh::MeshRange> meshes; /* of the uploaded scene */
    uint64_t uniqueTris = 0;             /* triangles of the meshes that are instanced at least once */
    uint32_t accelMode = PTC_ACCEL_AUTO;
    uint64_t twoLevelMinBytes = ~0ull;   /* AUTO: flattened bytes above which two levels are taken (a quarter of the device memory) */
    bool twoLevel = false;               /* what the last build produced */
    size_t travBytes() const { return twoLevel ? accel2.traversalBytes() : (accel.n ? accel.traversalBytes() : 0); }
    const void *travBase() const { return twoLevel ? (const void *)accel2.trav.p : (const void *)accel.trav.p; }

    /* render state */
    DBuf<float4> acc; /* the three accumulation targets (radiance, albedo, normal), contiguous: ONE ncclReduce sums them */
    /* two wavefronts in flight (see renderImpl): each has its own path state, queues, counters and stream */
    struct WaveBufs {
        DBuf<float4> orgRng, dirFlags, beta, radiance, hit, aovA, aovN, shOrg, shDir, shContrib, prBeta;
        DBuf<uint32_t> queue0, queue1, qShadow, qProbe, counters, hitInst;
        DBuf<unsigned long long> stats;
        size_t capacity = 0;
    } wave[2];
    DBuf<uint32_t> pixmap, tileOffsets, tileIds; /* tile split: local pixel -> global pixel, built on the device, kept across calls */
    uint32_t pixmapKey[5] = {0, 0, 0, 0, 0}, pixmapCount = 0;
    cudaEvent_t evStart = nullptr, evStop = nullptr; /* device time of one render */
    cudaStream_t stream2 = nullptr; /* second wavefront; c->stream carries the first and everything else */
    cudaEvent_t evA = nullptr, evB = nullptr, evAcc[2] = {nullptr, nullptr}, evFork = nullptr;
    static constexpr int RING = 8;
    cudaEvent_t evItem[RING] = {};
    /* Two wavefronts (half batches) in flight on two streams; the numbers cap the resident blocks per SM of the traversal / shading
     * kernels, 0 = one wavefront at a time.  Default: uncapped - every kernel asks for the whole GPU, so the other wavefront's kernel
     * simply fills the SMs that the running kernel's last, longest rays leave idle (atrium +0.2 %, fog +1.2 %, Cornell +3.4 %).
     * SHARING the SMs between the issue-bound k_extend and the DRAM-bound k_shade loses: 4+3 blocks 2030, 5+2 1888, 6+1 1604 against
     * 2069 Mseg/s alone (profiles/r1_v4_kernel_experiments.log) - k_shade needs >= 3 blocks per SM to keep DRAM busy and k_extend
     * loses as much with 4 as the overlap wins.  PTC_OVERLAP=t,s / PTC_OVERLAP=0 for experiments. */
    int overlapTrace = 64, overlapShade = 64;
    /* readback of the images: two pinned staging buffers (allocated at the first host readback) and the event of each one's copy */
    static constexpr size_t STAGE_BYTES = 16u << 20;
    char *stage[2] = {nullptr, nullptr};
    cudaEvent_t evStage[2] = {nullptr, nullptr};
    uint32_t tileOrder = 0; /* PTC_TILE_ORDER: tile edge of the single-rank pixel walk, 0 = row major */
    float sceneLo[3] = {0, 0, 0}, sceneHi[3] = {0, 0, 0}; /* world box of the scene (from the build) */

    std::atomic<float> progress{0.0f};
    ptc_stats stats{};
    std::string err; /* what went wrong on this device (collected by the context) */

    /* identity of what the texture arrays / the cubemap were built from (ptc_texture.uid, ptc_env.uid): an upload whose
     * textures are the same immutable objects keeps the device copies, like the reference, which uploads at import */
    std::vector<uint64_t> texSignature;
    /* (normal map, roughness map) pairs that createTextures packed into one RGBA8 layer (normal.rgb, roughness.r): pair -> texture index */
    std::map<std::pair<uint32_t, uint32_t>, uint32_t> packedPairs;
    uint32_t nSceneTextures = 0; /* textures of the scene description (nTextures also counts the packed copies) */
    uint64_t envSignature[3] = {0, 0, 0};

    void freeTextureClasses() {
        for (auto &t : texClasses) {
            if (t.tex) cudaDestroyTextureObject(t.tex);
            if (t.array) cudaFreeArray(t.array);
        }
        texClasses.clear();
        texSignature.clear();
        packedPairs.clear();
    }
    void freeCubemap() {
        if (cubeTex) cudaDestroyTextureObject(cubeTex);
        if (cubeArray) cudaFreeArray(cubeArray);
        cubeTex = 0;
        cubeArray = nullptr;
        cubeN = 0;
        envSignature[0] = envSignature[1] = envSignature[2] = 0;
    }
    void freeTextures() {
        freeTextureClasses();
        freeCubemap();
    }
    ~Dev() {
        cudaSetDevice(device);
        freeTextures();
        if (evA) cudaEventDestroy(evA);
        if (evB) cudaEventDestroy(evB);
        if (evFork) cudaEventDestroy(evFork);
        if (evStart) cudaEventDestroy(evStart);
        if (evStop) cudaEventDestroy(evStop);
        for (cudaEvent_t e : evAcc)
            if (e) cudaEventDestroy(e);
        for (cudaEvent_t e : evItem)
            if (e) cudaEventDestroy(e);
        for (cudaEvent_t e : evStage)
            if (e) cudaEventDestroy(e);
        for (char *p : stage)
            if (p) cudaFreeHost(p);
        if (stream2) cudaStreamDestroy(stream2);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace {

/* 3x3 inverse of the upper-left block of a column-major 4x4; row-major output */
void inverse3(const float *M, float *out) {
    float a = M[0], b = M[4], c = M[8];
    float d = M[1], e = M[5], f = M[9];
    float g = M[2], h = M[6], i = M[10];
    float co00 = e * i - f * h, co01 = -(d * i - f * g), co02 = d * h - e * g;
    float det = a * co00 + b * co01 + c * co02;
    float id = 1.0f / det;
    out[0] = co00 * id;
    out[1] = -(b * i - c * h) * id;
    out[2] = (b * f - c * e) * id;
    out[3] = co01 * id;
    out[4] = (a * i - c * g) * id;
    out[5] = -(a * f - c * d) * id;
    out[6] = co02 * id;
    out[7] = -(a * h - b * g) * id;
    out[8] = (a * e - b * d) * id;
}

DScene makeDScene(Dev *c) {
    DScene s{};
    s.vertices = c->vertices.p;
    s.indices = c->indices.p;
    s.instances = c->instances.p;
    s.materials = c->materials.p;
    s.lightData = c->lightData.p;
    s.lightInstances = c->lightInstances.p;
    s.texClasses = c->texClassTable.p;
    s.texRef = c->texRef.p;
    s.shading = c->twoLevel ? c->accel2.shading.p : c->accel.shading.p;
    s.cubemap = c->cubeTex;
    s.envCdfV = c->cubeTex ? c->envCdfV.p : nullptr;
    s.envCdfU = c->cubeTex ? c->envCdfU.p : nullptr;
    s.nInstances = c->nInstances;
    s.nMaterials = c->nMaterials;
    s.nLightInstances = c->nLightInstances;
    s.nTextures = c->nTextures;
    s.hasCubemap = c->cubeTex ? 1u : 0u;
    s.twoLevel = c->twoLevel ? 1u : 0u;
    if (c->twoLevel) {
        s.bvhNodes = c->accel2.nodes();
        s.tris = c->accel2.tris();
        s.nTris = (c->accelBuilt && c->accel2.tlas.nNodes) ? c->accel2.nTris : 0u; /* 0 = nothing to hit */
        s.nWideNodes = c->accel2.nNodes;
        s.tlasInst = c->accel2.tlasInst.p;
        s.meshRoot = c->accel2.meshRoot.p;
    } else {
        s.bvhNodes = c->accel.wideNodes();
        s.tris = c->accel.sortedTris();
        s.nTris = c->accelBuilt ? c->accel.n : 0u;
        s.nWideNodes = c->accel.nWide;
        s.tlasInst = nullptr;
        s.meshRoot = nullptr;
    }
    s.prmtMagic = 0x47000000u;
    s.anyEmissive = c->anyEmissive ? 1u : 0u;
    s.emissiveBoxes = c->emissiveBoxes.p;
    s.nEmissiveBoxes = c->nEmissiveBoxes;
    s.anyTransparent = c->anyTransparent ? 1u : 0u;
    s.anyVolume = c->anyVolumeChange ? 1u : 0u;
    return s;
}

/* Uploads the scene's 8-bit textures: every texture becomes RGBA8 (an R8 source reads back as (r, 0, 0, 1) like
 * VK_FORMAT_R8_UNORM), grouped into layered arrays by (width, height, sRGB).  Sampler = the reference's
 * (VulkanTexture.cpp:219-232): linear, REPEAT, normalised coordinates, sRGB decode before filtering. */
/* Which (normal map, roughness map) pairs can share one tap: both linear (the alpha channel of an sRGB texture is not decoded either,
 * but the normal map is linear anyway), the same size, neither all white.  Metallic-roughness assets sample exactly these two maps
 * at the same coordinates on every surface event (rayPrimaryPBRStandard.rchit.glsl:100-110). */
bool packablePair(const ptc_scene_desc *sd, const std::vector<uint32_t> &ref, uint32_t normalTex, uint32_t roughTex) {
    if (normalTex >= sd->n_textures || roughTex >= sd->n_textures || normalTex == roughTex) return false;
    if (ref[normalTex] == TEX_WHITE || ref[roughTex] == TEX_WHITE) return false;
    const ptc_texture &a = sd->textures[normalTex], &b = sd->textures[roughTex];
    return !a.srgb && !b.srgb && a.channels == 4 && a.width == b.width && a.height == b.height;
}

void createTextures(Dev *c, const ptc_scene_desc *sd) {
    const uint32_t n = sd->n_textures;
    struct Source { /* a layer: a texture of the scene, or a packed (normal, roughness) pair */
        uint32_t tex, rough;
    };
    std::vector<uint32_t> ref(n, 0);
    std::vector<std::vector<Source>> members;
    auto classOf = [&](uint32_t width, uint32_t height, uint32_t srgb) {
        uint32_t cls = 0;
        for (; cls < c->texClasses.size(); cls++) {
            const TextureClass &k = c->texClasses[cls];
            if (k.width == width && k.height == height && k.srgb == srgb && members[cls].size() < 2048) break;
        }
        if (cls == c->texClasses.size()) {
            TextureClass k;
            k.width = width;
            k.height = height;
            k.srgb = srgb;
            c->texClasses.push_back(k);
            members.emplace_back();
        }
        if (cls > 0xfffeu) throw CudaError{"too many texture classes"};
        return cls;
    };
    for (uint32_t t = 0; t < n; t++) {
        const ptc_texture &in = sd->textures[t];
        const size_t texels = (size_t)in.width * in.height;
        bool white = in.channels == 4;
        for (size_t p = 0; white && p < texels * 4; p++) white = in.data[p] == 255;
        if (white) {
            ref[t] = TEX_WHITE;
            continue;
        }
        const uint32_t cls = classOf(in.width, in.height, in.srgb ? 1u : 0u);
        ref[t] = (cls << 16) | (uint32_t)members[cls].size();
        members[cls].push_back(Source{t, 0xffffffffu});
    }
    /* packed copies: one extra layer per distinct pair; the materials of the device table are pointed at it (patchMaterials) */
    c->packedPairs.clear();
    for (uint32_t m = 0; m < sd->n_materials; m++) {
        const uint32_t nt = sd->materials[m].tex2[1], rt = sd->materials[m].tex1[2];
        if (!packablePair(sd, ref, nt, rt) || c->packedPairs.count({nt, rt})) continue;
        const ptc_texture &in = sd->textures[nt];
        const uint32_t cls = classOf(in.width, in.height, 0u);
        c->packedPairs[{nt, rt}] = (uint32_t)ref.size();
        ref.push_back((cls << 16) | (uint32_t)members[cls].size());
        members[cls].push_back(Source{nt, rt});
    }
    std::vector<cudaTextureObject_t> table(c->texClasses.size());
    for (size_t cls = 0; cls < c->texClasses.size(); cls++) {
        TextureClass &k = c->texClasses[cls];
        k.layers = (uint32_t)members[cls].size();
        const size_t texels = (size_t)k.width * k.height;
        cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uchar4>();
        CUDA_TRY(cudaMalloc3DArray(&k.array, &fmt, make_cudaExtent(k.width, k.height, k.layers), cudaArrayLayered));
        std::vector<uint8_t> rgba; /* staging for R8 sources and packed pairs */
        for (uint32_t l = 0; l < k.layers; l++) {
            const Source &srcId = members[cls][l];
            const ptc_texture &in = sd->textures[srcId.tex];
            const uint8_t *src = in.data;
            if (in.channels != 4) {
                rgba.resize(texels * 4);
                for (size_t p = 0; p < texels; p++) {
                    uint8_t px[4] = {0, 0, 0, 255};
                    for (uint32_t ch = 0; ch < in.channels && ch < 4; ch++) px[ch] = in.data[p * in.channels + ch];
                    std::memcpy(rgba.data() + p * 4, px, 4);
                }
                src = rgba.data();
            }
            if (srcId.rough != 0xffffffffu) { /* (normal.r, normal.g, normal.b, roughness.r) */
                const ptc_texture &ro = sd->textures[srcId.rough];
                rgba.resize(texels * 4);
                for (size_t p = 0; p < texels; p++) {
                    std::memcpy(rgba.data() + p * 4, in.data + p * 4, 3);
                    rgba[p * 4 + 3] = ro.data[p * ro.channels];
                }
                src = rgba.data();
            }
            cudaMemcpy3DParms cp{};
            cp.srcPtr = make_cudaPitchedPtr((void *)src, (size_t)k.width * 4, k.width, k.height);
            cp.dstArray = k.array;
            cp.dstPos = make_cudaPos(0, 0, l);
            cp.extent = make_cudaExtent(k.width, k.height, 1);
            cp.kind = cudaMemcpyHostToDevice;
            CUDA_TRY(cudaMemcpy3D(&cp)); /* synchronous: the staging vector is reused */
        }
        cudaResourceDesc rd{};
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = k.array;
        cudaTextureDesc td{};
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeWrap; /* REPEAT */
        td.filterMode = cudaFilterModeLinear;
        td.readMode = cudaReadModeNormalizedFloat;
        td.normalizedCoords = 1;
        td.sRGB = k.srgb ? 1 : 0; /* VK_FORMAT_R8G8B8A8_SRGB: decode before filtering */
        CUDA_TRY(cudaCreateTextureObject(&k.tex, &rd, &td, nullptr));
        table[cls] = k.tex;
    }
    c->nSceneTextures = n;
    c->nTextures = (uint32_t)ref.size();
    c->texRef.upload(ref.data(), ref.size(), c->stream);
    c->texClassTable.upload(table.data(), table.size(), c->stream);
    CUDA_TRY(cudaStreamSynchronize(c->stream)); /* the host vectors die here */
}

/* the device copy of the material table: materials whose (normal, roughness) maps were packed fetch the packed layer once */
void uploadMaterials(Dev *c, const ptc_scene_desc *sd) {
    std::vector<ptc_material> mats(sd->materials, sd->materials + sd->n_materials);
    if (getenv("PTC_NO_TEXTURE_PACKING") == nullptr)
        for (ptc_material &m : mats) {
            auto it = c->packedPairs.find({m.tex2[1], m.tex1[2]});
            if (it == c->packedPairs.end()) continue;
            m.tex2[1] = it->second;
            m.tex1[2] = TEX_IN_NORMAL_ALPHA;
        }
    c->materials.upload(mats.data(), mats.size(), c->stream);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
}

void createCubemap(Dev *c, const ptc_env &env) {
    if (!env.equirect_rgba || !env.width || !env.height) return;
    const uint32_t N = std::max(1u, std::min(env.width / 4u, 1080u)); /* VulkanRendererSkybox.cpp:100 */
    /* equirect as a float4 texture: linear, REPEAT, like the reference's sampler for the HDR image */
    cudaArray_t eqArray = nullptr;
    cudaTextureObject_t eqTex = 0;
    cudaChannelFormatDesc f4 = cudaCreateChannelDesc<float4>();
    CUDA_TRY(cudaMallocArray(&eqArray, &f4, env.width, env.height));
    CUDA_TRY(cudaMemcpy2DToArray(eqArray, 0, 0, env.equirect_rgba, (size_t)env.width * 16, (size_t)env.width * 16, env.height, cudaMemcpyHostToDevice));
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = eqArray;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeWrap;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 1;
    CUDA_TRY(cudaCreateTextureObject(&eqTex, &rd, &td, nullptr));

    DBuf<float4> faces;
    faces.alloc((size_t)6 * N * N);
    dim3 block(16, 16, 1), grid((N + 15) / 16, (N + 15) / 16, 6);
    wf::k_equirect_to_cube<<<grid, block, 0, c->stream>>>(eqTex, N, faces.p);
    CUDA_TRY(cudaGetLastError());

    CUDA_TRY(cudaMalloc3DArray(&c->cubeArray, &f4, make_cudaExtent(N, N, 6), cudaArrayCubemap));
    cudaMemcpy3DParms cp{};
    cp.srcPtr = make_cudaPitchedPtr(faces.p, (size_t)N * 16, N, N);
    cp.dstArray = c->cubeArray;
    cp.extent = make_cudaExtent(N, N, 6);
    cp.kind = cudaMemcpyDeviceToDevice;
    CUDA_TRY(cudaMemcpy3DAsync(&cp, c->stream));
    cudaResourceDesc crd{};
    crd.resType = cudaResourceTypeArray;
    crd.res.array.array = c->cubeArray;
    cudaTextureDesc ctd{};
    ctd.addressMode[0] = ctd.addressMode[1] = ctd.addressMode[2] = cudaAddressModeClamp;
    ctd.filterMode = cudaFilterModeLinear;
    ctd.readMode = cudaReadModeElementType;
    ctd.normalizedCoords = 1;
    ctd.seamlessCubemap = 1;
    CUDA_TRY(cudaCreateTextureObject(&c->cubeTex, &crd, &ctd, nullptr));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    cudaDestroyTextureObject(eqTex);
    cudaFreeArray(eqArray);
    c->cubeN = N;
    /* importance tables of the same input (host pass, ~10 ms for 3072 x 1536; kept with the cubemap by ptc_env.uid) */
    std::vector<float> cdfV, cdfU;
    envd::buildTables(env.equirect_rgba, env.width, env.height, cdfV, cdfU);
    c->envCdfV.upload(cdfV.data(), cdfV.size(), c->stream);
    c->envCdfU.upload(cdfU.data(), cdfU.size(), c->stream);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
}

/* L2 persistence for what every ray touches: the wide nodes (breadth first, so the top levels come first) and as much of
 * the triangle array as the persisting carve-out holds.  The path state streams through L2 (6 GB per batch at 1080p x 16)
 * and would otherwise keep evicting the BVH. */
void setTraversalWindow(Dev *c) {
    cudaStreamAttrValue attr{};
    cudaCtxResetPersistingL2Cache(); /* lines of a previous scene */
    const size_t bytes = c->travBytes();
    if (bytes == 0 || c->persistMax == 0 || c->windowMax == 0) {
        attr.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
        cudaStreamSetAttribute(c->stream2, cudaStreamAttributeAccessPolicyWindow, &attr);
        return;
    }
    const size_t carve = std::min(c->persistMax, bytes);
    CUDA_TRY(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
    const size_t window = std::min(bytes, c->windowMax);
    attr.accessPolicyWindow.base_ptr = const_cast<void *>(c->travBase());
    attr.accessPolicyWindow.num_bytes = window;
    attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)window);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    CUDA_TRY(cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    CUDA_TRY(cudaStreamSetAttribute(c->stream2, cudaStreamAttributeAccessPolicyWindow, &attr));
}

void ensureWave(Dev *c, int k, size_t slots, uint32_t depth) {
    Dev::WaveBufs &w = c->wave[k];
    if (c->twoLevel && w.hitInst.n < std::max(slots, w.capacity)) w.hitInst.alloc(std::max(slots, w.capacity));
    if (slots > w.capacity) {
        w.orgRng.alloc(slots);
        w.dirFlags.alloc(slots);
        w.beta.alloc(slots);
        w.radiance.alloc(slots);
        w.hit.alloc(slots);
        w.aovA.alloc(slots);
        w.aovN.alloc(slots);
        w.shOrg.alloc(slots);
        w.shDir.alloc(slots);
        w.shContrib.alloc(slots);
        w.prBeta.alloc(slots);
        w.queue0.alloc(slots);
        w.queue1.alloc(slots);
        w.qShadow.alloc(slots);
        w.qProbe.alloc(slots);
        if (c->twoLevel) w.hitInst.alloc(slots);
        w.capacity = slots;
    }
    w.counters.alloc((size_t)(depth + 2) * wf::CNT_STRIDE);
    w.stats.alloc(wf::ST_COUNT);
}

wf::Wave makeWave(Dev *c, int k) {
    Dev::WaveBufs &b = c->wave[k];
    wf::Wave w{};
    w.orgRng = b.orgRng.p;
    w.dirFlags = b.dirFlags.p;
    w.beta = b.beta.p;
    w.radiance = b.radiance.p;
    w.hit = b.hit.p;
    w.aovAlbedo = b.aovA.p;
    w.aovNormal = b.aovN.p;
    w.shOrgTmax = b.shOrg.p;
    w.shDirVol = b.shDir.p;
    w.shContrib = b.shContrib.p;
    w.prBetaPdf = b.prBeta.p;
    w.queue[0] = b.queue0.p;
    w.queue[1] = b.queue1.p;
    w.qShadow = b.qShadow.p;
    w.qProbe = b.qProbe.p;
    w.hitInst = b.hitInst.p;
    w.counters = b.counters.p;
    w.stats = b.stats.p;
    return w;
}

int renderImpl(Dev *c, const ptc_render_params *rp, float4 *dR, float4 *dA, float4 *dN) {
    const uint32_t W = rp->width, H = rp->height;
    const uint32_t batches = rp->samples / rp->batch_size; /* VulkanRendererPathTracing.cpp:798-799 (T7) */
    const uint32_t totalSamples = batches * rp->batch_size;
    const uint32_t world = rp->world ? rp->world : 1u;
    const size_t nPix = (size_t)W * H;
    cudaStream_t s = c->stream;
    c->progress = 0.0f; /* renderProgress() restarts with every render (VulkanRendererPathTracing.cpp:228-231) */

    /* pixel set of this rank: tiles dealt round-robin, enumerated tile by tile (a warp's 32 camera rays then cover a compact
     * screen area).  The host only walks the rank's TILES for their offsets; k_tile_pixmap fills the pixels; the result is kept
     * until the partition changes. */
    uint32_t nPixLocal = (uint32_t)nPix;
    const uint32_t *pixmapPtr = nullptr;
    /* a single rank walks the image in small tiles too (PTC_TILE_ORDER = tile edge, 0 = row major): the 32 camera rays of a warp and the
     * entries of a queue window then cover a compact screen area */
    const bool tileSplit = rp->split_mode == PTC_SPLIT_TILE && world > 1;
    const uint32_t orderTile = tileSplit ? 0u : c->tileOrder;
    if (tileSplit || orderTile > 0u) {
        const uint32_t tile = tileSplit ? (rp->tile_size ? rp->tile_size : 32u) : orderTile;
        const uint32_t world = tileSplit ? (rp->world ? rp->world : 1u) : 1u;
        const uint32_t key[5] = {W, H, tile, tileSplit ? rp->rank : 0u, world};
        if (memcmp(key, c->pixmapKey, sizeof(key)) != 0 || !c->pixmap.p) {
            const uint32_t tilesX = (W + tile - 1) / tile, tilesY = (H + tile - 1) / tile, nTiles = tilesX * tilesY;
            /* tile (tx, ty) belongs to rank (tx + ty) mod world: diagonal stripes, so that neither the columns of a colonnade nor the
             * horizon of a landscape line up with one rank's share (plain t mod world with a tile count per row that is a multiple of
             * world gave vertical stripes) */
            std::vector<uint32_t> offsets, tileIds;
            uint32_t total = 0;
            for (uint32_t t = 0; t < nTiles; t++) {
                const uint32_t tx = t % tilesX, ty = t / tilesX;
                if ((tx + ty) % world != key[3]) continue;
                offsets.push_back(total);
                tileIds.push_back(t);
                total += std::min(tile, W - tx * tile) * std::min(tile, H - ty * tile);
            }
            offsets.push_back(total);
            c->tileOffsets.upload(offsets.data(), offsets.size(), s);
            c->tileIds.alloc(std::max<size_t>(tileIds.size(), 1));
            if (!tileIds.empty()) c->tileIds.upload(tileIds.data(), tileIds.size(), s);
            c->pixmap.alloc(std::max<size_t>(total, 1));
            const uint32_t nLocalTiles = (uint32_t)tileIds.size();
            if (nLocalTiles) wf::k_tile_pixmap<<<nLocalTiles, 256, 0, s>>>(W, H, tile, c->tileIds.p, c->tileOffsets.p, c->pixmap.p);
            CUDA_TRY(cudaStreamSynchronize(s)); /* the host vector dies here */
            memcpy(c->pixmapKey, key, sizeof(key));
            c->pixmapCount = total;
        }
        nPixLocal = c->pixmapCount;
        pixmapPtr = c->pixmap.p;
    }

    /* explicit clear (the reference relies on fresh device memory being zero, trap T8) */
    CUDA_TRY(cudaMemsetAsync(dR, 0, nPix * sizeof(float4), s));
    CUDA_TRY(cudaMemsetAsync(dA, 0, nPix * sizeof(float4), s));
    CUDA_TRY(cudaMemsetAsync(dN, 0, nPix * sizeof(float4), s));

    /* Samples of one batch run concurrently as a wavefront; very large batches are cut into chunks that fit the wave.
     *
     * Two wavefronts are in flight, each on its own stream (PTC_OVERLAP, see ptc_ctx::overlapTrace): the kernels of one fill the SMs
     * that the tail of the other's running kernel leaves idle.  All of them are persistent kernels that pull work from device
     * queues, so any number of resident blocks makes progress.  A batch is split into two half-batch wavefronts, so the
     * path-state memory is what one full batch uses.  Work items are accumulated in a fixed order (event chain), which
     * keeps the image deterministic. */
    const bool timeKernels = (rp->flags & PTC_FLAG_TIME_KERNELS) != 0;
    const bool overlap = c->overlapTrace > 0 && c->overlapShade > 0 && !timeKernels && nPixLocal > 0;
    const int nWaves = overlap ? 2 : 1;
    size_t maxSlots = (size_t)1 << (overlap ? 25 : 26);
    if (const char *ms = getenv("PTC_MAX_SLOTS")) maxSlots = std::max<size_t>(1, (size_t)atoll(ms)); /* tests: force chunking */
    /* Small wavefronts waste the GPU (every bounce is a handful of persistent-kernel launches whose tails and late, nearly empty
     * bounces cost the same whatever the ray count): a rank that renders an eighth of a 4K image in batches of 8 samples ran 21 %
     * below the full render's rate (profiles/r2_partition_probe.log).  So when one batch fills less than half a wavefront, consecutive
     * batches are rendered as ONE work item (the sample index is global, so nothing but the grouping of the float additions into the
     * accumulators changes).  Sample split keeps its batches apart: rank r owns every world-th batch. */
    uint32_t mergeBatches = 1;
    if (nPixLocal > 0 && batches > 1 && !(rp->split_mode == PTC_SPLIT_SAMPLE && world > 1) && getenv("PTC_NO_BATCH_MERGE") == nullptr) {
        const size_t perBatch = (size_t)nPixLocal * rp->batch_size;
        /* (at least 8 work items stay, so that renderProgress() keeps moving like the reference's batch counter) */
        if (perBatch * 2 <= maxSlots) mergeBatches = (uint32_t)std::min<size_t>(std::max<uint32_t>(1u, batches / 8u), std::max<size_t>(1, (maxSlots * (size_t)nWaves) / perBatch));
    }
    const uint32_t itemSamplesMax = rp->batch_size * mergeBatches; /* samples of one work item */
    uint32_t chunkSamples = itemSamplesMax;
    if (nPixLocal > 0) chunkSamples = (uint32_t)std::max<size_t>(1, std::min<size_t>(itemSamplesMax, maxSlots / nPixLocal));
    if (overlap && chunkSamples == itemSamplesMax && itemSamplesMax >= 2) chunkSamples = (itemSamplesMax + 1) / 2;
    wf::Wave waves[2];
    cudaStream_t streams[2] = {s, c->stream2};
    for (int k = 0; k < nWaves; k++) {
        ensureWave(c, k, (size_t)chunkSamples * std::max(1u, nPixLocal), rp->depth);
        waves[k] = makeWave(c, k);
        CUDA_TRY(cudaMemsetAsync(c->wave[k].stats.p, 0, wf::ST_COUNT * sizeof(unsigned long long), s));
    }

    if (rp->flags & PTC_FLAG_SAMPLER_PMJ) {
        if (!c->pmjTable.p || !c->blueTable.p) throw CudaError{"PTC_FLAG_SAMPLER_PMJ needs ptc_set_sampler_tables"};
        if (W > 65535u || H > 65535u) throw CudaError{"the PMJ02BN sampler packs pixel coordinates into 16 bits"};
        const PmjConst pc{c->pmjTable.p, c->blueTable.p, totalSamples};
        CUDA_TRY(cudaMemcpyToSymbolAsync(g_pmj, &pc, sizeof(pc), 0, cudaMemcpyHostToDevice, s));
    }
    wf::RenderConst rc{};
    rc.sd = rp->scene;
    rc.width = W;
    rc.height = H;
    rc.depth = rp->depth;
    rc.totalSamples = totalSamples;
    rc.cameraType = rp->camera_type;
    rc.orthoW = rp->ortho_width;
    rc.orthoH = rp->ortho_height;
    rc.totalLights = c->nLightInstances;
    rc.envLight = ((rp->flags & PTC_FLAG_ENV_IMPORTANCE) && c->cubeTex && (rp->scene.background[3] == 1.0f || rp->scene.background[3] == 2.0f)) ? 1u : 0u;
    rc.totalLights += rc.envLight;
    rc.flags = rp->flags;
    rc.nPixLocal = nPixLocal;
    rc.pixmap = pixmapPtr;

    DScene sc = makeDScene(c);

    /* persistent kernels: exactly as many blocks as are resident at once (multiples of the SM count) */
    auto residentGrid = [&](const void *kernel, int block, int cap) {
        int perSm = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, block, 0) != cudaSuccess || perSm < 1) perSm = 1;
        if (cap > 0) perSm = std::min(perSm, cap);
        return c->smCount * perSm;
    };
    const int capTrace = overlap ? c->overlapTrace : 0, capShade = overlap ? c->overlapShade : 0;
    using ExtendFn = void (*)(wf::Wave, const DScene, uint32_t, wf::ExtendTune);
    const bool mediaInScene = c->anyVolumeChange || rp->scene.volumes[0] != -1.0f; /* = hasVolumes below: k_shade<*, true> writes the next tmax */
    const bool TL = c->twoLevel;
    const ExtendFn extendFn = mediaInScene ? (TL ? (ExtendFn)wf::k_extend<true, true> : (ExtendFn)wf::k_extend<true, false>)
                                           : (TL ? (ExtendFn)wf::k_extend<false, true> : (ExtendFn)wf::k_extend<false, false>);
    const int gridExtend = residentGrid((const void *)extendFn, TRV_BLOCK, capTrace);
    /* k_shade specialisation (wavefront.cuh): lights in the pick / media reachable */
    const bool hasLights = rc.totalLights > 0, hasVolumes = c->anyVolumeChange || rp->scene.volumes[0] != -1.0f;
    rc.fuseProbe = (c->anyEmissive && hasLights && !c->anyTransparent && !hasVolumes && getenv("PTC_NO_PROBE_FUSION") == nullptr) ? 1u : 0u;
    using ShadeFn = void (*)(wf::Wave, const DScene, const wf::RenderConst, uint32_t, uint32_t);
    /* k_shade<lights, media, sampler, environment light-sampled> (the environment as a light implies a non-empty light pick) */
#define SHADE_ROW(L, V, E) {wf::k_shade<L, V, 0, E>, wf::k_shade<L, V, 1, E>, wf::k_shade<L, V, 2, E>}
    static const ShadeFn shadeTable[2][2][2][3] = {{{SHADE_ROW(false, false, false), SHADE_ROW(false, false, false)}, {SHADE_ROW(false, true, false), SHADE_ROW(false, true, false)}},
                                                   {{SHADE_ROW(true, false, false), SHADE_ROW(true, false, true)}, {SHADE_ROW(true, true, false), SHADE_ROW(true, true, true)}}};
#undef SHADE_ROW
    const int samplerKind = (rp->flags & PTC_FLAG_SAMPLER_PMJ) ? 2 : ((rp->flags & PTC_FLAG_SAMPLER_SOBOL) ? 1 : 0);
    const ShadeFn shadeFn = shadeTable[hasLights ? 1 : 0][hasVolumes ? 1 : 0][rc.envLight ? 1 : 0][samplerKind];
    const int gridShade = residentGrid((const void *)shadeFn, 128, capShade);
    using ChainFn = void (*)(wf::Wave, const DScene, const wf::RenderConst, uint32_t, wf::ExtendTune);
    const ChainFn shadowFn = c->anyTransparent ? (TL ? (ChainFn)wf::k_shadow<false, true> : (ChainFn)wf::k_shadow<false, false>)
                                               : (TL ? (ChainFn)wf::k_shadow<true, true> : (ChainFn)wf::k_shadow<true, false>);
    const ChainFn probeFn = TL ? (ChainFn)wf::k_probe<true> : (ChainFn)wf::k_probe<false>;
    const int gridShadow = residentGrid((const void *)shadowFn, TRV_BLOCK, capTrace);
    const int gridProbe = residentGrid((const void *)probeFn, TRV_BLOCK, capTrace);
    uint64_t launches = 0, traceLaunches = 0;
    double traceMs = 0, shadeMs = 0, shadowMs = 0;
    auto timed = [&](double &acc, cudaStream_t st, auto &&launch) {
        if (timeKernels) CUDA_TRY(cudaEventRecord(c->evA, st));
        launch();
        if (timeKernels) {
            CUDA_TRY(cudaEventRecord(c->evB, st));
            CUDA_TRY(cudaEventSynchronize(c->evB));
            float ms = 0;
            CUDA_TRY(cudaEventElapsedTime(&ms, c->evA, c->evB));
            acc += ms;
        }
    };

    cudaEvent_t evStart = c->evStart, evStop = c->evStop; /* owned by the device state: nothing leaks when a call below throws */
    CUDA_TRY(cudaEventRecord(evStart, s));
    /* the second stream starts after the clears above */
    if (overlap) {
        CUDA_TRY(cudaEventRecord(c->evFork, s));
        CUDA_TRY(cudaStreamWaitEvent(c->stream2, c->evFork, 0));
    }
    uint32_t myBatches = 0;
    for (uint32_t b = 0; b < batches; b++)
        if (!(rp->split_mode == PTC_SPLIT_SAMPLE && world > 1 && (b % world) != rp->rank)) myBatches++;
    uint64_t totalItems = 0;
    for (uint32_t b = 0; b < batches; b += mergeBatches) {
        if (rp->split_mode == PTC_SPLIT_SAMPLE && world > 1 && (b % world) != rp->rank) continue;
        const uint32_t itemSamples = std::min(mergeBatches, batches - b) * rp->batch_size;
        totalItems += (itemSamples + chunkSamples - 1) / chunkSamples;
    }
    (void)myBatches;

    uint64_t item = 0;
    if (nPixLocal > 0) {
        for (uint32_t b = 0; b < batches; b += mergeBatches) {
            if (rp->split_mode == PTC_SPLIT_SAMPLE && world > 1 && (b % world) != rp->rank) continue;
            const uint32_t itemSamples = std::min(mergeBatches, batches - b) * rp->batch_size;
            for (uint32_t s0 = 0; s0 < itemSamples; s0 += chunkSamples, item++) {
                const int k = (int)(item % (uint64_t)nWaves);
                cudaStream_t st = streams[k];
                const wf::Wave &w = waves[k];
                /* at most RING work items in flight; renderProgress() follows the completed ones without draining the pipeline */
                if (item >= (uint64_t)Dev::RING) {
                    CUDA_TRY(cudaEventSynchronize(c->evItem[item % Dev::RING]));
                    c->progress = (float)(item - Dev::RING + 1) / (float)totalItems;
                }
                const uint32_t ns = std::min(chunkSamples, itemSamples - s0);
                const uint32_t nSlots = ns * nPixLocal;
                CUDA_TRY(cudaMemsetAsync(w.counters, 0, (size_t)(rp->depth + 2) * wf::CNT_STRIDE * sizeof(uint32_t), st));
                wf::k_raygen<<<(nSlots + 255) / 256, 256, 0, st>>>(w, rc, nSlots, b * rp->batch_size + s0);
                launches++;
                for (uint32_t d = 0; d < rp->depth; d++) {
                    timed(traceMs, st, [&] { extendFn<<<gridExtend, TRV_BLOCK, 0, st>>>(w, sc, d, c->tune); });
                    timed(shadeMs, st, [&] { shadeFn<<<gridShade, 128, 0, st>>>(w, sc, rc, d, b * rp->batch_size + s0); });
                    launches += 2;
                    traceLaunches++;
                    if (rc.totalLights > 0) {
                        timed(shadowMs, st, [&] { shadowFn<<<gridShadow, TRV_BLOCK, 0, st>>>(w, sc, rc, d, c->tune); });
                        launches++;
                    }
                    if (c->anyEmissive) {
                        timed(shadowMs, st, [&] { probeFn<<<gridProbe, TRV_BLOCK, 0, st>>>(w, sc, rc, d, c->tune); });
                        launches++;
                    }
                }
                wf::k_collect_stats<<<1, 32, 0, st>>>(w, rp->depth);
                /* accumulate in item order: item i adds after item i - 1 (which ran on the other stream) */
                if (overlap && item > 0) CUDA_TRY(cudaStreamWaitEvent(st, c->evAcc[(item - 1) & 1u], 0));
                wf::k_accumulate<<<(nPixLocal + 255) / 256, 256, 0, st>>>(w, rc, ns, dR, dA, dN);
                if (overlap) CUDA_TRY(cudaEventRecord(c->evAcc[item & 1u], st));
                CUDA_TRY(cudaEventRecord(c->evItem[item % Dev::RING], st));
                launches += 2;
                CUDA_TRY(cudaGetLastError());
            }
        }
    }
    /* join: the first stream waits for the last work item of the second */
    if (overlap && item > 0) CUDA_TRY(cudaStreamWaitEvent(s, c->evAcc[(item - 1) & 1u], 0));
    if (overlap && item > 1) CUDA_TRY(cudaStreamWaitEvent(s, c->evAcc[(item - 2) & 1u], 0));
    CUDA_TRY(cudaEventRecord(evStop, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, evStart, evStop));

    unsigned long long hs[wf::ST_COUNT] = {};
    for (int k = 0; k < nWaves; k++) {
        unsigned long long one[wf::ST_COUNT];
        CUDA_TRY(cudaMemcpy(one, c->wave[k].stats.p, sizeof(one), cudaMemcpyDeviceToHost));
        for (int i = 0; i < wf::ST_COUNT; i++) hs[i] += one[i];
    }
    c->stats.segments = hs[wf::ST_SEGMENTS];
    c->stats.path_rays = hs[wf::ST_SEGMENTS];
    c->stats.shadow_rays = hs[wf::ST_SHADOW_RAYS];
    c->stats.shadow_hops = hs[wf::ST_SHADOW_HOPS];
    c->stats.probe_rays = hs[wf::ST_PROBE_RAYS];
    c->stats.probe_hops = hs[wf::ST_PROBE_HOPS];
    c->stats.reserved[0] = hs[wf::ST_NODE_VISITS]; /* filled only by -DPTC_TRAV_STATS builds (tools/) */
    c->stats.reserved[1] = hs[wf::ST_TRI_TESTS];
    c->stats.reserved[2] = hs[wf::ST_NODE_ITERS];
    c->stats.reserved[3] = hs[wf::ST_TRI_ITERS];
    c->stats.render_ms = ms;
    c->stats.trace_ms = traceMs;
    c->stats.shade_ms = shadeMs;
    c->stats.shadow_ms = shadowMs;
    c->stats.trace_launches = traceLaunches;
    c->stats.kernel_launches = launches;
    c->progress = 1.0f;
    return 0;
}

/* ------------------------------------------------------------------ NCCL, resolved at run time
 * The only exchange of the path is the sum (sample split) / gather (tile split: disjoint pixels, zeros elsewhere) of the
 * accumulation buffers onto one GPU.  libnccl is opened on first use, so a single-GPU process needs no NCCL at
 * all and a process that already holds one (torch) shares it. */
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
    bool load() {
        if (handle) return true;
        const char *names[] = {getenv("PTC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n || !*n) continue;
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) {
            error = std::string("cannot open libnccl.so.2: ") + dlerror();
            return false;
        }
#define NCCL_SYM(field, sym)                                                        \
    field = reinterpret_cast<decltype(field)>(dlsym(handle, #sym));                 \
    if (!field) {                                                                   \
        error = "libnccl lacks " #sym;                                              \
        return false;                                                               \
    }
        NCCL_SYM(GetUniqueId, ncclGetUniqueId)
        NCCL_SYM(CommInitRank, ncclCommInitRank)
        NCCL_SYM(CommInitAll, ncclCommInitAll)
        NCCL_SYM(CommDestroy, ncclCommDestroy)
        NCCL_SYM(Reduce, ncclReduce)
        NCCL_SYM(Broadcast, ncclBroadcast)
        NCCL_SYM(GroupStart, ncclGroupStart)
        NCCL_SYM(GroupEnd, ncclGroupEnd)
        NCCL_SYM(GetErrorString, ncclGetErrorString)
#undef NCCL_SYM
        return true;
    }
};
NcclApi g_nccl;
std::mutex g_ncclMutex;
#define NCCL_TRY(expr)                                                                                                                    \
    do {                                                                                                                                  \
        ncclResult_t _r = (expr);                                                                                                         \
        if (_r != ncclSuccess) throw CudaError{std::string(#expr) + " -> " + g_nccl.GetErrorString(_r) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"}; \
    } while (0)

}  // namespace

/* The context: one Dev per GPU it drives.
 *   one device                      the plain case
 *   several devices, one process    ptc_create(.., ids, n > 1): one worker thread + streams per GPU, ncclCommInitAll over them;
 *                                   upload / build run on all of them, a render is partitioned (tiles or sample batches) and the
 *                                   buffers are reduced onto the first device
 *   one device, several processes   ptc_comm_init_rank: this context is rank r of `world` single-device contexts (one process per
 *                                   GPU under torchrun / MPI); ptc_render* reduce onto rank 0 */
struct ptc_ctx {
    std::vector<std::unique_ptr<Dev>> devs;
    std::string err;
    std::vector<ncclComm_t> comms; /* one per Dev (in-process group) or one (multi-process) */
    int commRank = 0, commWorld = 1; /* multi-process mode */
    ptc_stats stats{};
    double reduceMs = 0.0;
    ~ptc_ctx() {
        for (ncclComm_t cm : comms)
            if (cm && g_nccl.CommDestroy) g_nccl.CommDestroy(cm);
    }
};

namespace {

int fail(ptc_ctx *c, const std::string &msg) {
    if (c) c->err = msg;
    return 1;
}
Dev *dev0(ptc_ctx *ctx) { return (ctx && !ctx->devs.empty()) ? ctx->devs[0].get() : nullptr; }

/* runs fn(dev, index) for every device of the context, on one thread per device when there are several; the first error wins */
template <class Fn>
void forEachDev(ptc_ctx *ctx, Fn &&fn) {
    const size_t n = ctx->devs.size();
    std::vector<std::string> errs(n);
    auto body = [&](size_t i) {
        try {
            CUDA_TRY(cudaSetDevice(ctx->devs[i]->device));
            fn(ctx->devs[i].get(), (uint32_t)i);
        } catch (const CudaError &e) {
            errs[i] = e.msg;
        } catch (const std::exception &e) {
            errs[i] = e.what();
        }
    };
    if (n == 1) {
        body(0);
    } else {
        std::vector<std::thread> th;
        th.reserve(n);
        for (size_t i = 0; i < n; i++) {
            try {
                th.emplace_back(body, i);
            } catch (const std::system_error &e) { /* joinable threads must not be destroyed: report and stop starting more */
                errs[i] = std::string("cannot start the device thread: ") + e.what();
                break;
            }
        }
        for (auto &t : th) t.join();
    }
    for (size_t i = 0; i < n; i++)
        if (!errs[i].empty()) throw CudaError{n > 1 ? "device " + std::to_string(ctx->devs[i]->device) + ": " + errs[i] : errs[i]};
}

/* every index a kernel will follow unchecked is checked here, once, on the host (a malformed description must fail with a message,
 * not read out of bounds on the device) */
void validateScene(const ptc_scene_desc *sd) {
    if ((sd->n_vertices && !sd->vertices) || (sd->n_indices && !sd->indices) || (sd->n_meshes && !sd->meshes) || (sd->n_instances && !sd->instances) ||
        (sd->n_materials && !sd->materials) || (sd->n_light_data && !sd->light_data) || (sd->n_light_instances && !sd->light_instances) ||
        (sd->n_textures && !sd->textures))
        throw CudaError{"scene array is null but its count is not"};
    std::vector<uint8_t> meshChecked(sd->n_meshes, 0);
    auto volumeIndexOk = [&](float v) { return v == -1.0f || (v >= 0.0f && v < (float)sd->n_materials && v <= 65535.0f && v == floorf(v)); };
    for (uint32_t i = 0; i < sd->n_instances; i++) {
        const ptc_instance &in = sd->instances[i];
        if (in.mesh_index >= sd->n_meshes) throw CudaError{"instance mesh index out of range"};
        if (in.material_index >= sd->n_materials) throw CudaError{"instance material index out of range"};
        const ptc_mesh &m = sd->meshes[in.mesh_index];
        if (in.num_triangles != m.tri_count) throw CudaError{"instance num_triangles differs from its mesh's tri_count"};
        if (!volumeIndexOk(in.id[1]) || !volumeIndexOk(in.id[2])) throw CudaError{"instance volume material index out of range (must be -1 or a material index below 65536)"};
        if (meshChecked[in.mesh_index]) continue;
        meshChecked[in.mesh_index] = 1;
        if ((uint64_t)m.first_index + 3ull * m.tri_count > sd->n_indices) throw CudaError{"mesh index range out of bounds"};
        if ((uint64_t)m.first_vertex + m.vertex_count > sd->n_vertices) throw CudaError{"mesh vertex range out of bounds"};
        const uint32_t *ind = sd->indices + m.first_index;
        uint32_t mx = 0;
        for (uint64_t k = 0; k < 3ull * m.tri_count; k++) mx = std::max(mx, ind[k]);
        if (m.tri_count && mx >= m.vertex_count) throw CudaError{"mesh index exceeds its vertex count"};
    }
    for (uint32_t i = 0; i < sd->n_light_instances; i++) {
        const ptc_light_instance &L = sd->light_instances[i];
        if (L.info[3] > 2u) throw CudaError{"light instance type must be 0, 1 or 2"};
        if (L.info[3] == 2u) {
            if (L.info[1] >= sd->n_instances) throw CudaError{"mesh light instance index out of range"};
            if (sd->instances[L.info[1]].num_triangles == 0u) throw CudaError{"mesh light without triangles"};
        } else if (L.info[0] >= sd->n_light_data) {
            throw CudaError{"light data index out of range"};
        }
    }
}

/* ptc_upload_scene on one device.  geometryFollows: the vertex / index pools are only allocated here and arrive from the first device
 * of the context (broadcastGeometry) */
void uploadSceneDev(Dev *c, const ptc_scene_desc *sd, bool geometryFollows) {
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    c->sceneUploaded = false;
    c->accelBuilt = false;
    if (geometryFollows) {
        c->vertices.alloc(sd->n_vertices);
        c->indices.alloc(sd->n_indices);
    } else {
        c->vertices.upload(sd->vertices, sd->n_vertices, s);
        c->indices.upload(sd->indices, sd->n_indices, s);
    }
    c->lightData.upload(sd->light_data, sd->n_light_data, s);
    c->lightInstances.upload(sd->light_instances, sd->n_light_instances, s);
    c->nMaterials = sd->n_materials;
    c->nLightInstances = sd->n_light_instances;
    c->nInstances = sd->n_instances;

    std::vector<DInstance> inst(sd->n_instances);
    std::vector<float4> emBoxes;
    std::map<uint32_t, std::array<float, 6>> meshBox; /* object-space boxes of the meshes of emissive instances */
    float unionLo[3] = {1e30f, 1e30f, 1e30f}, unionHi[3] = {-1e30f, -1e30f, -1e30f};
    uint32_t nEmissiveInst = 0;
    uint64_t tri = 0;
    c->anyEmissive = c->anyTransparent = c->anyVolumeChange = false;
    for (uint32_t i = 0; i < sd->n_instances; i++) {
        const ptc_instance &in = sd->instances[i];
        if (in.mesh_index >= sd->n_meshes) throw CudaError{"instance mesh index out of range"};
        if (in.material_index >= sd->n_materials) throw CudaError{"instance material index out of range"};
        const ptc_mesh &m = sd->meshes[in.mesh_index];
        DInstance &d = inst[i];
        const float *M = in.model;
        for (int r = 0; r < 3; r++)
            for (int col = 0; col < 4; col++) d.m[r * 4 + col] = M[col * 4 + r];
        inverse3(M, d.nrm);
        /* world->object = [A^-1 | -A^-1 t] */
        for (int r = 0; r < 3; r++) {
            for (int col = 0; col < 3; col++) d.w2o[r * 4 + col] = d.nrm[r * 3 + col];
            d.w2o[r * 4 + 3] = -(d.nrm[r * 3 + 0] * M[12] + d.nrm[r * 3 + 1] * M[13] + d.nrm[r * 3 + 2] * M[14]);
        }
        d.volFront = in.id[1];
        d.volBack = in.id[2];
        d.material = in.material_index;
        d.firstIndex = m.first_index;
        d.firstVertex = m.first_vertex;
        d.numTriangles = in.num_triangles;
        d.firstWorldTri = (uint32_t)tri;
        d.mesh = in.mesh_index;
        tri += m.tri_count;
        const ptc_material &mat = sd->materials[in.material_index];
        const float ei = mat.emissive[3];
        if (std::fabs(ei * mat.emissive[0]) > 0.05f || std::fabs(ei * mat.emissive[1]) > 0.05f || std::fabs(ei * mat.emissive[2]) > 0.05f) {
            c->anyEmissive = true;
            /* world box of this emitter (DScene::emissiveBoxes): the mesh's object-space box (computed once per mesh) through the
             * instance transform, padded against the rounding of the device's own transform */
            nEmissiveInst++;
            if (m.vertex_count > 0) {
                if ((uint64_t)m.first_vertex + m.vertex_count > sd->n_vertices) throw CudaError{"mesh vertex range out of bounds"};
                auto it = meshBox.find(in.mesh_index);
                if (it == meshBox.end()) {
                    std::array<float, 6> ob = {1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f};
                    for (uint32_t v = 0; v < m.vertex_count; v++) {
                        const float *p = sd->vertices[m.first_vertex + v].position;
                        for (int r = 0; r < 3; r++) {
                            ob[r] = std::min(ob[r], p[r]);
                            ob[3 + r] = std::max(ob[3 + r], p[r]);
                        }
                    }
                    it = meshBox.emplace(in.mesh_index, ob).first;
                }
                const std::array<float, 6> &ob = it->second;
                float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
                for (int corner = 0; corner < 8; corner++) {
                    const float p[3] = {ob[(corner & 1) ? 3 : 0], ob[(corner & 2) ? 4 : 1], ob[(corner & 4) ? 5 : 2]};
                    for (int r = 0; r < 3; r++) {
                        const float wv = M[r] * p[0] + M[4 + r] * p[1] + M[8 + r] * p[2] + M[12 + r];
                        lo[r] = std::min(lo[r], wv);
                        hi[r] = std::max(hi[r], wv);
                    }
                }
                for (int r = 0; r < 3; r++) {
                    const float pad = 1e-4f * (std::fabs(lo[r]) + std::fabs(hi[r]) + (hi[r] - lo[r])) + 1e-5f;
                    lo[r] -= pad;
                    hi[r] += pad;
                    unionLo[r] = std::min(unionLo[r], lo[r]);
                    unionHi[r] = std::max(unionHi[r], hi[r]);
                }
                if (nEmissiveInst <= PTC_MAX_EMISSIVE_BOXES) {
                    emBoxes.push_back(make_float4(lo[0], lo[1], lo[2], 0.0f));
                    emBoxes.push_back(make_float4(hi[0], hi[1], hi[2], 0.0f));
                }
            }
        }
        if (mat.metallic_roughness_ao[3] > 0.0f) c->anyTransparent = true;
        if (in.id[1] != in.id[2]) c->anyVolumeChange = true;
    }
    if (tri >= 0x7fffffffull) throw CudaError{"more than 2^31 world triangles"};
    c->nWorldTris = (uint32_t)tri;
    c->instances.upload(inst.data(), inst.size(), s);
    c->meshes.resize(sd->n_meshes);
    {
        std::vector<uint8_t> used(sd->n_meshes, 0);
        for (uint32_t i = 0; i < sd->n_instances; i++) used[sd->instances[i].mesh_index] = 1;
        c->uniqueTris = 0;
        for (uint32_t m = 0; m < sd->n_meshes; m++) {
            /* a mesh no instance refers to gets an empty tree */
            c->meshes[m] = lbvh::MeshRange{sd->meshes[m].first_index, used[m] ? sd->meshes[m].tri_count : 0u, sd->meshes[m].first_vertex};
            if (used[m]) c->uniqueTris += sd->meshes[m].tri_count;
        }
    }
    /* few emitters: one box each; many: the box around all of them */
    if (nEmissiveInst > PTC_MAX_EMISSIVE_BOXES) {
        emBoxes.clear();
        emBoxes.push_back(make_float4(unionLo[0], unionLo[1], unionLo[2], 0.0f));
        emBoxes.push_back(make_float4(unionHi[0], unionHi[1], unionHi[2], 0.0f));
    }
    c->nEmissiveBoxes = nEmissiveInst >= 1 ? (uint32_t)(emBoxes.size() / 2) : 0u;
    if (c->nEmissiveBoxes) c->emissiveBoxes.upload(emBoxes.data(), emBoxes.size(), s);

    std::vector<uint64_t> sig;
    bool allIdentified = sd->n_textures > 0;
    for (uint32_t t = 0; t < sd->n_textures; t++) {
        const ptc_texture &in = sd->textures[t];
        if (in.channels != 1 && in.channels != 4) throw CudaError{"texture channels must be 1 or 4"};
        if (!in.data || !in.width || !in.height) throw CudaError{"texture without data"};
        allIdentified = allIdentified && in.uid != 0;
        sig.push_back(in.uid);
        sig.push_back(((uint64_t)in.width << 32) | in.height);
        sig.push_back(((uint64_t)in.channels << 32) | in.srgb);
    }
    uint64_t uploadBytes = sd->n_vertices * sizeof(ptc_vertex) + sd->n_indices * 4ull + (uint64_t)sd->n_materials * sizeof(ptc_material) +
                           (uint64_t)sd->n_light_data * sizeof(ptc_light_data) + (uint64_t)sd->n_light_instances * sizeof(ptc_light_instance) +
                           (uint64_t)sd->n_instances * sizeof(DInstance);
    /* (which maps are packed together depends on the materials: part of the identity of the device copies) */
    sig.push_back(0x7061636b65647321ull);
    for (uint32_t m = 0; m < sd->n_materials; m++) sig.push_back(((uint64_t)sd->materials[m].tex2[1] << 32) | sd->materials[m].tex1[2]);
    if (!(allIdentified && sig == c->texSignature && c->nSceneTextures == sd->n_textures)) {
        c->freeTextureClasses();
        createTextures(c, sd);
        if (allIdentified) c->texSignature = sig;
        for (uint32_t t = 0; t < sd->n_textures; t++) uploadBytes += (uint64_t)sd->textures[t].width * sd->textures[t].height * sd->textures[t].channels;
    }
    uploadMaterials(c, sd);
    const ptc_env &env = sd->env;
    const bool hasEnv = env.equirect_rgba && env.width && env.height;
    if (!(hasEnv && env.uid != 0 && c->cubeTex && c->envSignature[0] == env.uid && c->envSignature[1] == env.width && c->envSignature[2] == env.height)) {
        c->freeCubemap();
        createCubemap(c, env);
        if (hasEnv) uploadBytes += (uint64_t)env.width * env.height * 16ull;
        if (hasEnv && env.uid != 0) {
            c->envSignature[0] = env.uid;
            c->envSignature[1] = env.width;
            c->envSignature[2] = env.height;
        }
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    c->stats.upload_bytes = uploadBytes;
    c->sceneUploaded = true;
}


/* One process, several GPUs: the geometry pools cross PCIe ONCE, to the first device, and reach the others through the context's
 * communicator (NVLink / NVSwitch) instead of N pageable host copies competing for the host's memory system (a 180 MB pool to 8 GPUs:
 * 74 ms, profiles/r2_offlinerender_c4_8gpu.log).  The small per-render arrays (instances, materials, lights) stay plain copies. */
void broadcastGeometry(ptc_ctx *ctx, const ptc_scene_desc *sd) {
    const size_t nDev = ctx->devs.size();
    NCCL_TRY(g_nccl.GroupStart());
    for (size_t i = 0; i < nDev; i++) {
        Dev *c = ctx->devs[i].get();
        CUDA_TRY(cudaSetDevice(c->device));
        if (sd->n_vertices)
            NCCL_TRY(g_nccl.Broadcast(c->vertices.p, c->vertices.p, sd->n_vertices * sizeof(ptc_vertex), ncclUint8, 0, ctx->comms[i], c->stream));
        if (sd->n_indices) NCCL_TRY(g_nccl.Broadcast(c->indices.p, c->indices.p, sd->n_indices * 4ull, ncclUint8, 0, ctx->comms[i], c->stream));
    }
    NCCL_TRY(g_nccl.GroupEnd());
    for (size_t i = 0; i < nDev; i++) {
        CUDA_TRY(cudaSetDevice(ctx->devs[i]->device));
        CUDA_TRY(cudaStreamSynchronize(ctx->devs[i]->stream));
    }
}

__global__ void k_fill_alpha(float4 *a, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) a[i].w = 1.0f;
}

/* render settings a backend cannot run */
void checkRenderParams(ptc_ctx *ctx, const ptc_render_params *rp) {
    for (auto &d : ctx->devs)
        if (!d->accelBuilt) throw CudaError{"ptc_build_accel has not been called"};
    if (rp->batch_size == 0 || rp->width == 0 || rp->height == 0 || rp->depth == 0) throw CudaError{"bad render params"};
    if (rp->depth > 255) throw CudaError{"depth must be <= 255"};
    if ((uint64_t)rp->width * rp->height >= 0x7fffffffull) throw CudaError{"image too large"};
    const float cv = rp->scene.volumes[0];
    if (!(cv == -1.0f || (cv >= 0.0f && cv < (float)ctx->devs[0]->nMaterials && cv <= 65535.0f && cv == floorf(cv))))
        throw CudaError{"camera volume material index out of range"};
    if (rp->split_mode > PTC_SPLIT_SAMPLE) throw CudaError{"unknown split mode"};
    if (rp->split_mode != PTC_SPLIT_NONE && rp->world > 1 && rp->rank >= rp->world) throw CudaError{"rank must be below world"};
}

/* Device -> pageable host memory through the device's two pinned staging buffers: the DMA of chunk i runs while the host moves chunk
 * i - 1 out of the other buffer (on up to four threads).  A plain cudaMemcpy into pageable memory does the same staging on one thread
 * without the overlap: 400 MB of 4K images took 79 ms (5 GB/s, profiles/r2_offlinerender_c4_8gpu.log). */
void hostCopyParallel(char *dst, const char *src, size_t n) {
    const size_t ways = n >= (8u << 20) ? 4 : (n >= (2u << 20) ? 2 : 1);
    if (ways == 1) {
        std::memcpy(dst, src, n);
        return;
    }
    const size_t part = (n / ways + 63) & ~(size_t)63;
    std::vector<std::thread> th;
    th.reserve(ways);
    for (size_t w = 1; w < ways; w++) {
        const size_t b = std::min(n, w * part), e = w + 1 == ways ? n : std::min(n, (w + 1) * part);
        if (e <= b) continue;
        try {
            th.emplace_back([=] { std::memcpy(dst + b, src + b, e - b); });
        } catch (const std::system_error &) { /* no thread to be had: this part is copied here */
            std::memcpy(dst + b, src + b, e - b);
        }
    }
    std::memcpy(dst, src, std::min(n, part));
    for (auto &t : th) t.join();
}

void readbackStaged(Dev *c, float *const hOut[3], float4 *const img[3], size_t bytesPerImage) {
    struct Chunk {
        char *dst;
        const char *src;
        size_t n;
    };
    std::vector<Chunk> chunks;
    for (int k = 0; k < 3; k++)
        if (hOut[k])
            for (size_t off = 0; off < bytesPerImage; off += Dev::STAGE_BYTES)
                chunks.push_back({(char *)hOut[k] + off, (const char *)img[k] + off, std::min(Dev::STAGE_BYTES, bytesPerImage - off)});
    if (chunks.empty()) return;
    for (int b = 0; b < 2; b++) {
        if (!c->stage[b]) CUDA_TRY(cudaHostAlloc((void **)&c->stage[b], Dev::STAGE_BYTES, cudaHostAllocDefault));
        if (!c->evStage[b]) CUDA_TRY(cudaEventCreateWithFlags(&c->evStage[b], cudaEventDisableTiming));
    }
    auto issue = [&](size_t i) {
        CUDA_TRY(cudaMemcpyAsync(c->stage[i & 1u], chunks[i].src, chunks[i].n, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaEventRecord(c->evStage[i & 1u], c->stream));
    };
    issue(0);
    for (size_t i = 0; i < chunks.size(); i++) {
        if (i + 1 < chunks.size()) issue(i + 1); /* its buffer was emptied in the previous iteration */
        CUDA_TRY(cudaEventSynchronize(c->evStage[i & 1u]));
        hostCopyParallel(chunks[i].dst, c->stage[i & 1u], chunks[i].n);
    }
}

/* The partitioned render of a context: every device renders its share into its own three targets (one allocation), the targets are
 * summed onto the root with ONE ncclReduce per device (tiles are disjoint and zero elsewhere, so the sum serves both split modes),
 * and the root writes alpha = 1 behind the reduce.  dOut (optional, device pointers on the root's device) / hOut (optional, host)
 * receive the three images.  Returns false on a non-root rank of a multi-process group (no image there). */
bool renderAll(ptc_ctx *ctx, const ptc_render_params *rp, void *const dOut[3], float *const hOut[3]) {
    checkRenderParams(ctx, rp);
    const bool verbose = getenv("PTC_VERBOSE") != nullptr;
    auto tick = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) { /* host wall clock of the phases of one call */
        const auto now = std::chrono::steady_clock::now();
        if (verbose) fprintf(stderr, "[ptc]    %-12s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - tick).count());
        tick = now;
    };
    const size_t nPix = (size_t)rp->width * rp->height;
    const uint32_t nDev = (uint32_t)ctx->devs.size();
    const bool group = nDev > 1, ranks = ctx->commWorld > 1;
    ptc_render_params base = *rp;
    if (group || ranks) {
        /* the context knows the partition; the caller only chooses the mode (none = sample batches: any batch count balances) */
        if (base.split_mode == PTC_SPLIT_NONE) base.split_mode = PTC_SPLIT_SAMPLE;
        base.world = group ? nDev : (uint32_t)ctx->commWorld;
    }
    const bool isRoot = !ranks || ctx->commRank == 0;
    /* device targets: the caller's (single device, device pointers given) or the device's own contiguous triple */
    std::vector<float4 *> target(nDev, nullptr);
    const bool callerTargets = dOut && dOut[0] && dOut[1] && dOut[2] && !group && !ranks;
    forEachDev(ctx, [&](Dev *c, uint32_t i) {
        ptc_render_params mine = base;
        if (group) mine.rank = i;
        if (ranks) mine.rank = (uint32_t)ctx->commRank;
        float4 *r, *a, *n;
        if (callerTargets) {
            r = (float4 *)dOut[0], a = (float4 *)dOut[1], n = (float4 *)dOut[2];
        } else {
            c->acc.alloc(3 * nPix);
            r = c->acc.p, a = r + nPix, n = a + nPix;
            target[i] = r;
        }
        renderImpl(c, &mine, r, a, n);
    });
    lap("render");
    /* the exchange */
    ctx->reduceMs = 0.0;
    if (group || ranks) {
        Dev *root = ctx->devs[0].get();
        CUDA_TRY(cudaSetDevice(root->device));
        CUDA_TRY(cudaEventRecord(root->evA, root->stream));
        NCCL_TRY(g_nccl.GroupStart());
        for (uint32_t i = 0; i < nDev; i++) {
            Dev *c = ctx->devs[i].get();
            CUDA_TRY(cudaSetDevice(c->device));
            NCCL_TRY(g_nccl.Reduce(target[i], target[i], 3 * nPix * 4, ncclFloat, ncclSum, 0, ctx->comms[i], c->stream));
        }
        NCCL_TRY(g_nccl.GroupEnd());
        CUDA_TRY(cudaSetDevice(root->device));
        CUDA_TRY(cudaEventRecord(root->evB, root->stream));
        for (uint32_t i = 0; i < nDev; i++) {
            CUDA_TRY(cudaSetDevice(ctx->devs[i]->device));
            CUDA_TRY(cudaStreamSynchronize(ctx->devs[i]->stream));
        }
        CUDA_TRY(cudaSetDevice(root->device));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, root->evA, root->evB));
        ctx->reduceMs = ms;
    }
    lap("reduce");
    /* statistics of the whole context */
    ptc_stats st{};
    const ptc_stats &b0 = ctx->devs[0]->stats;
    st.build_ms = b0.build_ms, st.n_triangles = b0.n_triangles, st.n_bvh_nodes = b0.n_bvh_nodes, st.scene_bytes = b0.scene_bytes;
    st.upload_bytes = b0.upload_bytes;
    st.accel_levels = b0.accel_levels, st.traversal_bytes = b0.traversal_bytes;
    st.reduce_ms = ctx->reduceMs;
    for (auto &d : ctx->devs) {
        const ptc_stats &x = d->stats;
        st.segments += x.segments, st.path_rays += x.path_rays, st.shadow_rays += x.shadow_rays, st.shadow_hops += x.shadow_hops;
        st.probe_rays += x.probe_rays, st.probe_hops += x.probe_hops, st.trace_launches += x.trace_launches, st.kernel_launches += x.kernel_launches;
        st.render_ms = std::max(st.render_ms, x.render_ms), st.trace_ms = std::max(st.trace_ms, x.trace_ms);
        st.shade_ms = std::max(st.shade_ms, x.shade_ms), st.shadow_ms = std::max(st.shadow_ms, x.shadow_ms);
        st.build_ms = std::max(st.build_ms, x.build_ms);
        for (int k = 0; k < 4; k++) st.reserved[k] += x.reserved[k];
    }
    st.render_ms += ctx->reduceMs;
    ctx->stats = st;
    if (!isRoot) return false;
    /* alpha = 1 is written in ONE place: here on the image's owner, or - for a caller who partitions by hand and sums the parts
     * himself (rank / world given, no communicator) - on rank 0 only */
    Dev *root = ctx->devs[0].get();
    CUDA_TRY(cudaSetDevice(root->device));
    const bool manualPart = !group && !ranks && rp->split_mode != PTC_SPLIT_NONE && rp->world > 1;
    float4 *img[3];
    for (int k = 0; k < 3; k++) img[k] = callerTargets ? (float4 *)dOut[k] : target[0] + (size_t)k * nPix;
    if (!manualPart || rp->rank == 0)
        for (int k = 0; k < 3; k++) k_fill_alpha<<<(unsigned)((nPix + 255) / 256), 256, 0, root->stream>>>(img[k], nPix);
    if (!callerTargets && dOut)
        for (int k = 0; k < 3; k++)
            if (dOut[k]) CUDA_TRY(cudaMemcpyAsync(dOut[k], img[k], nPix * 16, cudaMemcpyDeviceToDevice, root->stream));
    /* readback like getRenderTargetData x3 (…PathTracing.cpp:890-893) */
    if (hOut) readbackStaged(root, hOut, img, nPix * 16);
    CUDA_TRY(cudaStreamSynchronize(root->stream));
    lap("readback");
    return true;
}

}  // namespace

/* ====================================================================== C-ABI */
#define PTC_GUARD_BEGIN try {
#define PTC_GUARD_END(ctx)                          \
    }                                               \
    catch (const CudaError &e) {                    \
        return fail(ctx, e.msg);                    \
    }                                               \
    catch (const std::exception &e) {               \
        return fail(ctx, e.what());                 \
    }

extern "C" {

PTC_API const char *ptc_backend_name(void) { return "cuda-sm_100a"; }

static void createDev(Dev *c, int device) {
    c->device = device;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, c->device));
    if (prop.major != 10) throw CudaError{std::string("device '") + prop.name + "' is not sm_100; this build targets B200 only"};
    c->smCount = prop.multiProcessorCount;
    c->persistMax = (size_t)std::max(0, prop.persistingL2CacheMaxSize);
    c->windowMax = (size_t)std::max(0, prop.accessPolicyMaxWindowSize);
    if (const char *h = getenv("PTC_HIERARCHY")) {
        if (!strcmp(h, "lbvh")) c->accel.hierarchy = PTC_HIERARCHY_LBVH;
        if (!strcmp(h, "ploc")) c->accel.hierarchy = PTC_HIERARCHY_PLOC;
    }
    if (const char *r = getenv("PTC_PLOC_RADIUS")) {
        const int v = atoi(r);
        if (v >= 1 && v <= PLOC_MAX_RADIUS) c->accel.plocRadius = (uint32_t)v;
    }
    if (const char *t = getenv("PTC_EXTEND_TUNE")) { /* "minActive,triEnter,triLeave,blocked" */
        unsigned a, b, d, e;
        if (sscanf(t, "%u,%u,%u,%u", &a, &b, &d, &e) == 4) c->tune = wf::ExtendTune{a, b, d, e};
    }
    c->twoLevelMinBytes = (uint64_t)prop.totalGlobalMem / 4;
    if (const char *b = getenv("PTC_TWO_LEVEL_MIN_BYTES")) c->twoLevelMinBytes = (uint64_t)atoll(b);
    if (const char *a = getenv("PTC_ACCEL")) { /* flat | two | auto: default mode of ptc_set_accel_mode, for tuning runs */
        if (!strcmp(a, "flat")) c->accelMode = PTC_ACCEL_FLAT;
        if (!strcmp(a, "two")) c->accelMode = PTC_ACCEL_TWO_LEVEL;
    }
    if (const char *t = getenv("PTC_TILE_ORDER")) c->tileOrder = (uint32_t)std::min(64, std::max(0, atoi(t)));
    if (const char *o = getenv("PTC_OVERLAP")) { /* "traceBlocksPerSM,shadeBlocksPerSM"; "0" = one wavefront at a time */
        int a = 0, b = 0;
        const int got = sscanf(o, "%d,%d", &a, &b);
        if (got == 2 && a > 0 && b > 0) c->overlapTrace = a, c->overlapShade = b;
        else if (got >= 1 && a == 0) c->overlapTrace = c->overlapShade = 0;
    }
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&c->evA));
    CUDA_TRY(cudaEventCreate(&c->evB));
    CUDA_TRY(cudaEventCreate(&c->evStart));
    CUDA_TRY(cudaEventCreate(&c->evStop));
    CUDA_TRY(cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming));
    for (cudaEvent_t &e : c->evAcc) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (cudaEvent_t &e : c->evItem) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
}

PTC_API int ptc_create(ptc_ctx **out, const int *device_ids, int n_devices) {
    if (!out) return 1;
    *out = nullptr;
    ptc_ctx *ctx = new ptc_ctx();
    *out = ctx; /* returned even on failure so that ptc_last_error can explain */
    PTC_GUARD_BEGIN
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(ctx, std::string("no CUDA device available (") + cudaGetErrorString(e) + "); this library has no CPU fallback");
    std::vector<int> ids;
    if (device_ids && n_devices > 0) {
        ids.assign(device_ids, device_ids + n_devices);
    } else {
        int cur = 0;
        if (cudaGetDevice(&cur) != cudaSuccess) cur = 0;
        ids.push_back(cur);
    }
    for (size_t i = 0; i < ids.size(); i++) {
        if (ids[i] < 0 || ids[i] >= count) return fail(ctx, "device id " + std::to_string(ids[i]) + " out of range (" + std::to_string(count) + " devices)");
        for (size_t j = 0; j < i; j++)
            if (ids[j] == ids[i]) return fail(ctx, "device id " + std::to_string(ids[i]) + " listed twice");
    }
    for (int id : ids) {
        ctx->devs.emplace_back(new Dev());
        createDev(ctx->devs.back().get(), id);
    }
    if (ids.size() > 1) { /* one process, several GPUs: ncclCommInitAll */
        std::lock_guard<std::mutex> lock(g_ncclMutex);
        if (!g_nccl.load()) return fail(ctx, g_nccl.error);
        ctx->comms.assign(ids.size(), nullptr);
        NCCL_TRY(g_nccl.CommInitAll(ctx->comms.data(), (int)ids.size(), ids.data()));
    }
    CUDA_TRY(cudaSetDevice(ids[0]));
    return 0;
    PTC_GUARD_END(ctx)
}

PTC_API void ptc_destroy(ptc_ctx *ctx) { delete ctx; }
PTC_API const char *ptc_last_error(const ptc_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
PTC_API int ptc_device_count(const ptc_ctx *ctx) { return ctx ? (int)ctx->devs.size() : 0; }

PTC_API int ptc_comm_unique_id(uint8_t *out128) {
    if (!out128) return 1;
    std::lock_guard<std::mutex> lock(g_ncclMutex);
    if (!g_nccl.load()) return 1;
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ptc_comm_unique_id hands out 128 bytes");
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return 1;
    memcpy(out128, &id, 128);
    return 0;
}

PTC_API int ptc_comm_init_rank(ptc_ctx *ctx, const uint8_t *id128, int rank, int world) {
    if (!ctx || !id128) return fail(ctx, "null argument");
    if (ctx->devs.size() != 1) return fail(ctx, "ptc_comm_init_rank needs a single-device context (one process per GPU)");
    if (world < 1 || rank < 0 || rank >= world) return fail(ctx, "bad rank / world");
    if (!ctx->comms.empty()) return fail(ctx, "context already has a communicator");
    PTC_GUARD_BEGIN
    {
        std::lock_guard<std::mutex> lock(g_ncclMutex);
        if (!g_nccl.load()) return fail(ctx, g_nccl.error);
    }
    if (world == 1) return 0;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    CUDA_TRY(cudaSetDevice(ctx->devs[0]->device));
    ctx->comms.assign(1, nullptr);
    NCCL_TRY(g_nccl.CommInitRank(&ctx->comms[0], world, id, rank));
    ctx->commRank = rank;
    ctx->commWorld = world;
    return 0;
    PTC_GUARD_END(ctx)
}

PTC_API int ptc_upload_scene(ptc_ctx *ctx, const ptc_scene_desc *sd) {
    if (!ctx || !sd) return fail(ctx, "null argument");
    if (ctx->devs.empty()) return fail(ctx, "context has no CUDA device");
    PTC_GUARD_BEGIN
    for (auto &d : ctx->devs) d->sceneUploaded = d->accelBuilt = false;
    validateScene(sd);
    /* the scene is replicated: every device gets its own copy */
    const uint64_t geometryBytes = sd->n_vertices * sizeof(ptc_vertex) + sd->n_indices * 4ull;
    const char *minEnv = getenv("PTC_SCENE_BROADCAST_MIN_BYTES"); /* negative = never */
    const long long minBytes = minEnv ? atoll(minEnv) : (1ll << 20);
    const bool viaLink = ctx->devs.size() > 1 && ctx->comms.size() == ctx->devs.size() && minBytes >= 0 && geometryBytes >= (uint64_t)minBytes;
    forEachDev(ctx, [&](Dev *c, uint32_t i) { uploadSceneDev(c, sd, viaLink && i > 0); });
    if (viaLink) broadcastGeometry(ctx, sd);
    ctx->stats.upload_bytes = ctx->devs[0]->stats.upload_bytes;
    return 0;
    PTC_GUARD_END(ctx)
}

/* replaces VulkanRandom::createBuffers (vulkan/resources/VulkanRandom.cpp:40-72): the two sampler tables as device arrays */
PTC_API int ptc_set_sampler_tables(ptc_ctx *ctx, const float *pmj, uint32_t n_sequences, uint32_t n_samples, const float *blue, uint32_t n_textures,
                                   uint32_t resolution) {
    if (!ctx || !pmj || !blue) return fail(ctx, "null argument");
    if (ctx->devs.empty()) return fail(ctx, "context has no CUDA device");
    if (n_sequences != PMJ_N_SEQUENCES || n_samples != PMJ_N_SAMPLES || n_textures != BLUE_NOISE_TEXTURES || resolution != BLUE_NOISE_RESOLUTION)
        return fail(ctx, "sampler tables must be 16 x 16384 x 2 and 48 x 128 x 128 (rng_pmj_defines.glsl, bluenoise_defines.glsl)");
    PTC_GUARD_BEGIN
    forEachDev(ctx, [&](Dev *c, uint32_t) {
        c->pmjTable.upload(pmj, (size_t)n_sequences * n_samples * 2, c->stream);
        c->blueTable.upload(blue, (size_t)n_textures * resolution * resolution, c->stream);
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    });
    return 0;
    PTC_GUARD_END(ctx)
}

PTC_API int ptc_set_build_options(ptc_ctx *ctx, uint32_t hierarchy, uint32_t ploc_radius) {
    if (!ctx) return 1;
    if (hierarchy != PTC_HIERARCHY_LBVH && hierarchy != PTC_HIERARCHY_PLOC) return fail(ctx, "unknown hierarchy");
    if (ploc_radius > PLOC_MAX_RADIUS) return fail(ctx, "ploc_radius must be <= 32");
    for (auto &c : ctx->devs) {
        c->accel.hierarchy = hierarchy;
        c->accel.plocRadius = ploc_radius ? ploc_radius : 16u;
        c->accelBuilt = false;
    }
    return 0;
}

PTC_API int ptc_set_accel_mode(ptc_ctx *ctx, uint32_t mode) {
    if (!ctx) return 1;
    if (mode > PTC_ACCEL_TWO_LEVEL) return fail(ctx, "unknown acceleration-structure mode");
    for (auto &c : ctx->devs) {
        c->accelMode = mode;
        c->accelBuilt = false;
    }
    return 0;
}

PTC_API int ptc_build_accel(ptc_ctx *ctx) {
    if (!ctx) return 1;
    if (ctx->devs.empty()) return fail(ctx, "context has no CUDA device");
    for (auto &d : ctx->devs)
        if (!d->sceneUploaded) return fail(ctx, "ptc_upload_scene has not been called");
    PTC_GUARD_BEGIN
    /* every device builds its own copy from the same input: the build is deterministic, so the copies are identical */
    forEachDev(ctx, [&](Dev *c, uint32_t) {
        CUDA_TRY(cudaEventRecord(c->evA, c->stream));
        /* One tree over world-space triangles, or one per mesh below one over the instances (VulkanScene.cpp:306-381)?  Measured on C4
         * (42.5 M world / 0.6 M unique triangles, profiles/r2_c4_flat_vs_two_level.log): flat 1025 Mseg/s, two levels 648 - the boxes of
         * a forest's instances overlap, a ray enters many bottom-level trees, and 180 GB of HBM hold the flattened 8 GB easily.  Two
         * levels buy MEMORY (traversal + shading data 8.6 GB -> 0.13 GB), so AUTO takes them only when the flattened data would exceed a
         * quarter of the device memory (PTC_TWO_LEVEL_MIN_BYTES overrides the threshold). */
        bool two = c->accelMode == PTC_ACCEL_TWO_LEVEL;
        if (c->accelMode == PTC_ACCEL_AUTO) {
            const uint64_t flatBytes = (uint64_t)c->nWorldTris * (48 + 144 + 14); /* triangles + shading records + ~0.17 nodes of 80 B each */
            two = flatBytes > c->twoLevelMinBytes && (uint64_t)c->nWorldTris >= 2ull * std::max<uint64_t>(c->uniqueTris, 1);
        }
        c->twoLevel = two;
        if (two)
            c->accel2.run(c->vertices.p, c->indices.p, c->instances.p, c->nInstances, c->meshes, c->accel.hierarchy, c->accel.plocRadius, c->stream);
        else
            c->accel.run(c->vertices.p, c->indices.p, c->instances.p, c->nInstances, c->nWorldTris, c->stream);
        CUDA_TRY(cudaEventRecord(c->evB, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, c->evA, c->evB));
        c->accelBuilt = true;
        if (two) {
            for (int a = 0; a < 3; a++) c->sceneLo[a] = c->accel2.tlas.lo[a], c->sceneHi[a] = c->accel2.tlas.hi[a];
        } else if (c->accel.n > 0) {
            uint32_t sb[6];
            CUDA_TRY(cudaMemcpy(sb, c->accel.sceneBounds.p, sizeof(sb), cudaMemcpyDeviceToHost));
            for (int a = 0; a < 3; a++) c->sceneLo[a] = lbvh::floatUnflip(sb[a]), c->sceneHi[a] = lbvh::floatUnflip(sb[3 + a]);
        }
        setTraversalWindow(c);
        c->stats.build_ms = ms;
        c->stats.n_triangles = c->nWorldTris;
        c->stats.n_bvh_nodes = two ? c->accel2.nNodes : c->accel.nWide;
        c->stats.accel_levels = two ? 2u : 1u;
        c->stats.traversal_bytes = c->travBytes();
        c->stats.scene_bytes = (two ? c->accel2.bytes() : c->accel.bytes()) + c->vertices.bytes() + c->indices.bytes() + c->instances.bytes() + c->materials.bytes();
    });
    const ptc_stats &b0 = ctx->devs[0]->stats;
    ctx->stats.build_ms = b0.build_ms, ctx->stats.n_triangles = b0.n_triangles, ctx->stats.n_bvh_nodes = b0.n_bvh_nodes, ctx->stats.scene_bytes = b0.scene_bytes;
    ctx->stats.accel_levels = b0.accel_levels, ctx->stats.traversal_bytes = b0.traversal_bytes;
    for (auto &d : ctx->devs) ctx->stats.build_ms = std::max(ctx->stats.build_ms, d->stats.build_ms);
    return 0;
    PTC_GUARD_END(ctx)
}

PTC_API int ptc_render_device(ptc_ctx *ctx, const ptc_render_params *rp, void *dR, void *dA, void *dN) {
    if (!ctx || !rp) return fail(ctx, "null argument");
    if (ctx->devs.empty()) return fail(ctx, "context has no CUDA device");
    PTC_GUARD_BEGIN
    void *const d[3] = {dR, dA, dN};
    renderAll(ctx, rp, d, nullptr);
    return 0;
    PTC_GUARD_END(ctx)
}

PTC_API int ptc_render(ptc_ctx *ctx, const ptc_render_params *rp, float *radiance, float *albedo, float *normal) {
    if (!ctx || !rp) return fail(ctx, "null argument");
    if (ctx->devs.empty()) return fail(ctx, "context has no CUDA device");
    PTC_GUARD_BEGIN
    float *const h[3] = {radiance, albedo, normal};
    renderAll(ctx, rp, nullptr, h);
    return 0;
    PTC_GUARD_END(ctx)
}

PTC_API float ptc_progress(const ptc_ctx *ctx) {
    if (!ctx || ctx->devs.empty()) return 0.0f;
    float p = 0.0f;
    for (auto &d : ctx->devs) p += d->progress.load();
    return p / (float)ctx->devs.size();
}
PTC_API int ptc_get_stats(ptc_ctx *ctx, ptc_stats *out) {
    if (!ctx || !out) return 1;
    *out = ctx->stats;
    return 0;
}

PTC_API int ptc_trace_closest(ptc_ctx *ctx, const float *rays, int n, int *inst, int *prim, float *t, float *u, float *v) {
    Dev *c = dev0(ctx);
    if (!c || !rays) return fail(ctx, "null argument");
    if (!c->accelBuilt) return fail(ctx, "ptc_build_accel has not been called");
    if (n <= 0) return 0;
    PTC_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    DBuf<float> dRays, dT, dU, dV;
    DBuf<int> dInst, dPrim;
    DBuf<uint32_t> dCounter;
    dRays.upload(rays, (size_t)n * 8, s);
    dT.alloc(n); dU.alloc(n); dV.alloc(n); dInst.alloc(n); dPrim.alloc(n);
    dCounter.alloc(1);
    CUDA_TRY(cudaMemsetAsync(dCounter.p, 0, 4, s));
    DScene sc = makeDScene(c);
    const int grid = std::max(1, std::min((n + TRV_BLOCK - 1) / TRV_BLOCK, c->smCount * 4));
    if (c->twoLevel)
        wf::k_trace_closest<true><<<grid, TRV_BLOCK, 0, s>>>(sc, dRays.p, (uint32_t)n, dCounter.p, dInst.p, dPrim.p, dT.p, dU.p, dV.p, c->tune);
    else
        wf::k_trace_closest<false><<<grid, TRV_BLOCK, 0, s>>>(sc, dRays.p, (uint32_t)n, dCounter.p, dInst.p, dPrim.p, dT.p, dU.p, dV.p, c->tune);
    CUDA_TRY(cudaGetLastError());
    if (inst) CUDA_TRY(cudaMemcpyAsync(inst, dInst.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    if (prim) CUDA_TRY(cudaMemcpyAsync(prim, dPrim.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    if (t) CUDA_TRY(cudaMemcpyAsync(t, dT.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    if (u) CUDA_TRY(cudaMemcpyAsync(u, dU.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    if (v) CUDA_TRY(cudaMemcpyAsync(v, dV.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return 0;
    PTC_GUARD_END(ctx)
}

/* two-level structure dump: level -1 = the top-level tree (primitives = instances), level m >= 0 = the tree of mesh m (primitives =
 * its triangles); node words with indices RELATIVE to the tree, like ptc_get_wide_bvh */
PTC_API int ptc_get_accel_level(ptc_ctx *ctx, int32_t level, uint64_t *n_nodes_out, uint64_t *n_prims_out, uint32_t *node_words, uint32_t *prim_order, float *box6) {
    Dev *c = dev0(ctx);
    if (!c) return 1;
    if (!c->accelBuilt) return fail(ctx, "ptc_build_accel has not been called");
    if (!c->twoLevel) return fail(ctx, "the acceleration structure is single level (ptc_set_accel_mode)");
    if (level < -1 || level >= (int32_t)c->accel2.blas.size()) return fail(ctx, "no such level");
    PTC_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    const lbvh::TwoLevel::Tree &t = level < 0 ? c->accel2.tlas : *c->accel2.blas[level];
    if (n_nodes_out) *n_nodes_out = t.nNodes;
    if (n_prims_out) *n_prims_out = t.nPrims;
    if (box6)
        for (int a = 0; a < 3; a++) box6[a] = t.lo[a], box6[3 + a] = t.hi[a];
    if (node_words && t.nNodes) CUDA_TRY(cudaMemcpy(node_words, t.nodes.p, (size_t)t.nNodes * 80, cudaMemcpyDeviceToHost));
    if (prim_order && t.nPrims) CUDA_TRY(cudaMemcpy(prim_order, t.order.p, (size_t)t.nPrims * 4, cudaMemcpyDeviceToHost));
    return 0;
    PTC_GUARD_END(ctx)
}

PTC_API int ptc_get_lbvh(ptc_ctx *ctx, uint64_t *n_out, uint64_t *morton, uint32_t *order, int32_t *parent, int32_t *left, int32_t *right,
                         float *aabb) {
    Dev *c = dev0(ctx);
    if (!c) return 1;
    if (!c->accelBuilt) return fail(ctx, "ptc_build_accel has not been called");
    PTC_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    if (c->twoLevel) return fail(ctx, "the acceleration structure has two levels: use ptc_get_accel_level");
    const lbvh::Build &B = c->accel;
    const size_t n = B.n;
    if (n_out) *n_out = n;
    if (n == 0) return 0;
    const size_t nn = 2 * n - 1;
    if (morton) CUDA_TRY(cudaMemcpy(morton, B.keysSorted.p, n * 8, cudaMemcpyDeviceToHost));
    if (order) CUDA_TRY(cudaMemcpy(order, B.order.p, n * 4, cudaMemcpyDeviceToHost));
    if (parent) CUDA_TRY(cudaMemcpy(parent, B.parent.p, nn * 4, cudaMemcpyDeviceToHost));
    if (left) CUDA_TRY(cudaMemcpy(left, B.left.p, nn * 4, cudaMemcpyDeviceToHost));
    if (right) CUDA_TRY(cudaMemcpy(right, B.right.p, nn * 4, cudaMemcpyDeviceToHost));
    if (aabb) {
        std::vector<float4> lo(nn), hi(nn);
        CUDA_TRY(cudaMemcpy(lo.data(), B.nodeLo.p, nn * 16, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(hi.data(), B.nodeHi.p, nn * 16, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < nn; i++) {
            aabb[i * 6 + 0] = lo[i].x; aabb[i * 6 + 1] = lo[i].y; aabb[i * 6 + 2] = lo[i].z;
            aabb[i * 6 + 3] = hi[i].x; aabb[i * 6 + 4] = hi[i].y; aabb[i * 6 + 5] = hi[i].z;
        }
    }
    return 0;
    PTC_GUARD_END(ctx)
}

PTC_API int ptc_get_wide_bvh(ptc_ctx *ctx, uint64_t *n_nodes_out, uint64_t *n_tris_out, uint32_t *node_words, uint32_t *tri_order) {
    Dev *c = dev0(ctx);
    if (!c) return 1;
    if (!c->accelBuilt) return fail(ctx, "ptc_build_accel has not been called");
    PTC_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    if (c->twoLevel) return fail(ctx, "the acceleration structure has two levels: use ptc_get_accel_level");
    const lbvh::Build &B = c->accel;
    if (n_nodes_out) *n_nodes_out = B.n ? B.nWide : 0;
    if (n_tris_out) *n_tris_out = B.n;
    if (B.n == 0) return 0;
    if (node_words) CUDA_TRY(cudaMemcpy(node_words, B.wideNodes(), (size_t)B.nWide * 80, cudaMemcpyDeviceToHost));
    if (tri_order) CUDA_TRY(cudaMemcpy(tri_order, B.wideOrder.p, (size_t)B.n * 4, cudaMemcpyDeviceToHost));
    return 0;
    PTC_GUARD_END(ctx)
}

PTC_API int ptc_bsdf_eval(ptc_ctx *ctx, int n, const float *params, const float *wi, const float *wo, float *out_f, float *out_pdf) {
    Dev *c = dev0(ctx);
    if (!c || !c->stream) return fail(ctx, "context has no CUDA device");
    if (n <= 0) return 0;
    PTC_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    DBuf<float> dP, dWi, dWo, dF, dPdf;
    dP.upload(params, (size_t)n * 5, s);
    dWi.upload(wi, (size_t)n * 3, s);
    dWo.upload(wo, (size_t)n * 3, s);
    dF.alloc((size_t)n * 3);
    dPdf.alloc(n);
    wf::k_bsdf_eval<<<(n + 255) / 256, 256, 0, s>>>(n, dP.p, dWi.p, dWo.p, dF.p, dPdf.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out_f, dF.p, (size_t)n * 12, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(out_pdf, dPdf.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return 0;
    PTC_GUARD_END(ctx)
}

PTC_API int ptc_bsdf_sample(ptc_ctx *ctx, int n, const float *params, const float *wo, const float *u, float *out_wi, float *out_f,
                            float *out_pdf) {
    Dev *c = dev0(ctx);
    if (!c || !c->stream) return fail(ctx, "context has no CUDA device");
    if (n <= 0) return 0;
    PTC_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    DBuf<float> dP, dWo, dU, dWi, dF, dPdf;
    dP.upload(params, (size_t)n * 5, s);
    dWo.upload(wo, (size_t)n * 3, s);
    dU.upload(u, (size_t)n * 3, s);
    dWi.alloc((size_t)n * 3);
    dF.alloc((size_t)n * 3);
    dPdf.alloc(n);
    wf::k_bsdf_sample<<<(n + 255) / 256, 256, 0, s>>>(n, dP.p, dWo.p, dU.p, dWi.p, dF.p, dPdf.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out_wi, dWi.p, (size_t)n * 12, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(out_f, dF.p, (size_t)n * 12, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(out_pdf, dPdf.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return 0;
    PTC_GUARD_END(ctx)
}

PTC_API int ptc_sampler_points(ptc_ctx *ctx, uint32_t px, uint32_t py, uint32_t width, uint32_t first_index, uint32_t count, uint32_t dimension,
                               uint32_t flags, float *out_xy) {
    Dev *c = dev0(ctx);
    if (!c || !c->stream) return fail(ctx, "context has no CUDA device");
    if (count == 0) return 0;
    if (!out_xy) return fail(ctx, "null argument");
    PTC_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    if (flags & PTC_FLAG_SAMPLER_PMJ) {
        if (!c->pmjTable.p || !c->blueTable.p) return fail(ctx, "ptc_set_sampler_tables has not been called");
        const PmjConst pc{c->pmjTable.p, c->blueTable.p, first_index + count};
        CUDA_TRY(cudaMemcpyToSymbolAsync(g_pmj, &pc, sizeof(pc), 0, cudaMemcpyHostToDevice, c->stream));
    }
    DBuf<float> dO;
    dO.alloc((size_t)count * 2);
    wf::k_sampler_points<<<(count + 255) / 256, 256, 0, c->stream>>>(px, py, width, first_index, count, dimension, flags, dO.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out_xy, dO.p, (size_t)count * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
    PTC_GUARD_END(ctx)
}

PTC_API int ptc_srgb_table(ptc_ctx *ctx, float *out256) {
    Dev *c = dev0(ctx);
    if (!c || !c->stream || !out256) return fail(ctx, "context has no CUDA device");
    PTC_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    /* a 256 x 1 sRGB texture whose texel i holds the code i, created exactly like the scene's textures (createTextures) */
    cudaArray_t arr = nullptr;
    cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uchar4>();
    CUDA_TRY(cudaMalloc3DArray(&arr, &fmt, make_cudaExtent(256, 1, 1), cudaArrayLayered));
    std::vector<uint8_t> px(256 * 4);
    for (int i = 0; i < 256; i++) px[4 * i] = px[4 * i + 1] = px[4 * i + 2] = (uint8_t)i, px[4 * i + 3] = 255;
    cudaMemcpy3DParms cp{};
    cp.srcPtr = make_cudaPitchedPtr(px.data(), 256 * 4, 256, 1);
    cp.dstArray = arr;
    cp.extent = make_cudaExtent(256, 1, 1);
    cp.kind = cudaMemcpyHostToDevice;
    cudaError_t e = cudaMemcpy3D(&cp);
    cudaTextureObject_t tex = 0;
    if (e == cudaSuccess) {
        cudaResourceDesc rd{};
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = arr;
        cudaTextureDesc td{};
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeWrap;
        td.filterMode = cudaFilterModeLinear;
        td.readMode = cudaReadModeNormalizedFloat;
        td.normalizedCoords = 1;
        td.sRGB = 1;
        e = cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    }
    DBuf<float> dO;
    if (e == cudaSuccess) {
        dO.alloc(256);
        wf::k_srgb_table<<<1, 256, 0, c->stream>>>(tex, dO.p);
        e = cudaMemcpyAsync(out256, dO.p, 256 * 4, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    }
    if (tex) cudaDestroyTextureObject(tex);
    cudaFreeArray(arr);
    CUDA_TRY(e);
    return 0;
    PTC_GUARD_END(ctx)
}

PTC_API int ptc_env_lookup(ptc_ctx *ctx, int n, const float *dirs, float *out_rgb) {
    Dev *c = dev0(ctx);
    if (!c || !c->stream) return fail(ctx, "context has no CUDA device");
    if (n <= 0) return 0;
    PTC_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    DBuf<float> dD, dO;
    dD.upload(dirs, (size_t)n * 3, s);
    dO.alloc((size_t)n * 3);
    DScene sc = makeDScene(c);
    wf::k_env_lookup<<<(n + 255) / 256, 256, 0, s>>>(sc, n, dD.p, dO.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out_rgb, dO.p, (size_t)n * 12, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return 0;
    PTC_GUARD_END(ctx)
}

PTC_API int ptc_env_sample(ptc_ctx *ctx, int n, const float *u01, float *out_dirs, float *out_pdf) {
    Dev *c = dev0(ctx);
    if (!c || !c->stream) return fail(ctx, "context has no CUDA device");
    if (!c->cubeTex || !c->envCdfV.p) return fail(ctx, "no environment");
    if (n <= 0) return 0;
    PTC_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    DBuf<float> dU, dD, dP;
    dU.upload(u01, (size_t)n * 2, s);
    dD.alloc((size_t)n * 3);
    dP.alloc((size_t)n);
    wf::k_env_sample<<<(n + 255) / 256, 256, 0, s>>>(makeDScene(c), n, dU.p, dD.p, dP.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out_dirs, dD.p, (size_t)n * 12, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(out_pdf, dP.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return 0;
    PTC_GUARD_END(ctx)
}

PTC_API int ptc_env_pdf(ptc_ctx *ctx, int n, const float *dirs, float *out_pdf) {
    Dev *c = dev0(ctx);
    if (!c || !c->stream) return fail(ctx, "context has no CUDA device");
    if (!c->cubeTex || !c->envCdfV.p) return fail(ctx, "no environment");
    if (n <= 0) return 0;
    PTC_GUARD_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    DBuf<float> dD, dP;
    dD.upload(dirs, (size_t)n * 3, s);
    dP.alloc((size_t)n);
    wf::k_env_pdf<<<(n + 255) / 256, 256, 0, s>>>(makeDScene(c), n, dD.p, dP.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out_pdf, dP.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return 0;
    PTC_GUARD_END(ctx)
}

} /* extern "C" */
