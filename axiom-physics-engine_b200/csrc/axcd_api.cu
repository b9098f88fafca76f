// C ABI (include/axcd.h) over the CUDA kernels: context, device buffers, stage orchestration.
// No CPU fallback exists here: every stage is a kernel launch on the context's stream.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include <dlfcn.h>

#include "axcd.h"
#include "axcd_common.cuh"
#include "axcd_lbvh.cuh"
#include "axcd_epa_warp.cuh"
#include "axcd_narrow.cuh"
#include "axcd_manifold.cuh"
#include "axcd_ccd.cuh"
#include "axcd_query.cuh"
#include "axcd_refit.cuh"
#include "axcd_sort.cuh"

using namespace axcd;

// launch shapes of the latency-bound helper kernels (swept on the B200, see profiles/r01_experiments.md)
#ifndef AXCD_SCATTER_BLOCKS
#define AXCD_SCATTER_BLOCKS 32    // blocks per SM of scatterPairsKernel (sweep 4/8/16/32)
#endif
#ifndef AXCD_TOPO_THREADS
#define AXCD_TOPO_THREADS 64
#endif

namespace {

enum Stage { ST_NONE = 0, ST_SHAPES = 1, ST_POSES = 2, ST_REFIT = 3, ST_BROAD = 4, ST_NARROW = 5 };
enum Ev { EV_START, EV_REFIT, EV_SORT, EV_BUILD, EV_PAIR, EV_PAIRSORT, EV_BROAD_END, EV_N0, EV_GJK, EV_END, EV_COUNT };

}  // namespace

struct AxcdContext {
    AxcdConfig cfg;
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    int stage = ST_NONE;
    int numSMs = kNumSMs;    // multiprocessor count of the device (grid sizing of the persistent kernels)
    uint32_t n = 0;          // bodies
    uint32_t nHull = 0;
    bool hasWorlds = false;
    int mortonBits = 10;     // per axis
    int worldBits = 0;
    uint32_t numPairs = 0;   // pairs held (<= maxPairs)
    uint32_t foundPairs = 0; // pairs found (may exceed capacity)
    uint32_t numContacts = 0, foundContacts = 0;
    uint32_t launches[3] = {0, 0, 0};   // kernels launched by refit / broadphase / narrowphase
    Counters hostCtr;
    char lastErr[256] = {0};

    // device buffers
    float* dXf = nullptr;            // n * 10 floats (axiom::math::Transform AoS)
    float* dPoseStage = nullptr;     // n * 7 floats: landing area of axcd_set_poses
    AxcdContact* sinkDev = nullptr;  // axcd_set_contact_sink: device-side address of the caller's page-locked buffer
    AxcdContact* sinkHost = nullptr;
    // progress words of the fused narrowphase (page-locked, mapped): [0] pair count + 1, [1 + t] contacts through tile t + 1
    volatile uint32_t* hProgress = nullptr;
    uint32_t* dProgress = nullptr;       // the same words as the kernel addresses them
    uint32_t progressWords = 0;
    cudaStream_t copyStream = nullptr;   // the sink's DMA chunks travel here while the narrowphase still runs
    bool sinkPending = false;            // a narrowphase that reports progress was launched and its sink not yet drained
    bool narrowReportsProgress = false;  // what the last axcd_narrowphase call launched (saved with a captured graph)
    uint4* dShapes = nullptr;
    uint8_t* dType8 = nullptr;       // shape type per body, rewritten by every refit
    float4* dHull = nullptr;
    uint32_t* dWorld = nullptr;
    uint32_t* dBodyKeys = nullptr;   // slab mode: global id of every local body
    uint4* dFilters = nullptr;       // (category, mask, group, pad) per body
    bool filtersOn = false;
    bool slabOn = false;
    float slabLo = 0.0f, slabHi = 0.0f;
    uint32_t nOwned = 0;             // slab mode: bodies [0, nOwned) are owned, the rest are ghosts
    float4* dGhostSend = nullptr;    // slab mode: numRanks send buffers of ghost records
    uint32_t* dGhostCount = nullptr;
    uint32_t ghostRanks = 0, ghostCap = 0;
    // x-slab mode with the exchange inside the library (axcd_slab_init / axcd_slab_step): NCCL communicator,
    // slab edges on the device, the size matrix of the handshake and the receive buffer
    void* slabComm = nullptr;        // ncclComm_t
    bool slabOwnsComm = false;
    uint32_t slabRank = 0, slabRanks = 0;
    float* dEdges = nullptr;         // slabRanks + 1 slab edges
    uint32_t* dCountMatrix = nullptr;   // slabRanks x slabRanks: row r = records rank r sends to each rank
    uint32_t* hCountMatrix = nullptr;   // pinned host copy
    float4* dGhostRecv = nullptr;       // ghostCap records
    float4* dGhostVertSend = nullptr;   // slabRanks x ghostVertCap hull vertices of the ghosts sent
    uint32_t ghostVertCap = 0;          // maxHullVerts - nHull: room behind the owned vertices in the hull pool
    uint32_t lastGhosts = 0;
    float lastExchangeMs = 0.0f;
    cudaEvent_t evX0 = nullptr, evX1 = nullptr;
    float* dAabb = nullptr;          // n * 6 floats (axiom::math::AABB AoS)
    uint32_t* dKeys[2] = {nullptr, nullptr};
    uint32_t* dVals[2] = {nullptr, nullptr};
    float4* dSegLo = nullptr;        // segment tree over sorted leaves, heap layout, 2*P entries
    float4* dSegHi = nullptr;
    uint32_t segP = 0;               // leaf level size (power of two >= n)
    Node32* dNodes32 = nullptr;      // compact nodes of the pair traversal (32 B per internal node)
    BvhNode* dNodes = nullptr;       // float nodes for the scene queries: allocated and built by the first query after a broadphase
    bool queryNodesValid = false;
    bool hasHulls = false;           // any convex-hull shape in the current scene (ghosts are never hulls)
    bool hasGenericShapes = false;   // any hull, capsule or cylinder among the owned bodies: pairs that need GJK can exist
    bool hasCylinders = false;       // any cylinder among the owned bodies: the cylinder-capable kernel instantiations run
    bool hasSweptRayShapes = false;  // any hull or cylinder: ray casts against them run the conservative advancement
    uint32_t* dWorldEnd = nullptr;
    uint2* dPairsTmp = nullptr;      // candidate pairs as found (unordered)
    uint2* dPairs = nullptr;         // candidate pairs, canonical order
    uint32_t* dBodyCount = nullptr;  // pairs per body a; [n, 2n) = fill cursors of the scatter
    uint32_t* dBodyStart = nullptr;  // exclusive scan of dBodyCount
    uint32_t* dSegB = nullptr;       // per-body segments of partner indices
    uint32_t* dScanStatus = nullptr;
    EpaWork* dEpaWork = nullptr;     // GJK -> EPA queue (maxContacts)
    uint32_t* dEpaOverflow = nullptr;
    float* dEpaSpill = nullptr;      // polytopes of pairs that outgrew the fast EPA caps
    uint32_t spillCap = 0;
    uint32_t* dSlotStatus = nullptr; // look-back status, one word per slot-scan tile
    uint32_t* dChunks = nullptr;     // class-homogeneous chunks of 32 pair indices for the GJK kernel
    uint8_t* dFlags = nullptr;       // per pair: 0 none, 1 shallow contact, 2 EPA contact
    uint32_t* dSlots = nullptr;      // per pair: contact slot
    AxcdContact* dTmpContacts = nullptr;   // per pair: shallow contact records before compaction
    AxcdContact* dContacts = nullptr;
    AxcdManifold* dManifolds = nullptr;   // allocated by the first axcd_build_manifolds
    bool manifoldsValid = false;          // built for the current narrowphase result
    uint8_t* dAwake = nullptr;            // per body: 0 = sleeping (NULL semantics when !awakeOn)
    bool awakeOn = false;
    bool fatValid = false;                // coherent mode: dAabb holds fat boxes of the current shapes
    bool pairsCached = false;             // coherent mode: dPairs / cachedPairs describe the current fat boxes
    uint32_t cachedFound = 0;             // pair count of the cached broadphase (as found, may exceed capacity)
    bool broadSkipped = false;            // the last axcd_broadphase reused the cached pairs
    uint32_t movedLast = 0;
    int sortedBuf = 0;                    // which of dKeys/dVals holds the sorted order of the last broadphase
    // scene-query scratch (grow-only)
    void* dQIn = nullptr;      size_t qInBytes = 0;      // query boxes / rays (+ worlds)
    void* dQCount = nullptr;   size_t qCountBytes = 0;   // counts, starts, cursors, scan status, ticket/total
    void* dQSeg = nullptr;     size_t qSegBytes = 0;     // hit bodies per query segment
    void* dQOut = nullptr;     size_t qOutBytes = 0;     // sorted (query, body) hits / ray hits
    float* dPairDist = nullptr;
    uint32_t* dSortHist = nullptr;
    uint32_t* dStepHist = nullptr;   // the step's digit histograms: filled by the Morton kernel, cleared again behind the sort
    uint32_t* dSortStatus = nullptr;
    // bucket sort of the Morton keys (axcd_sort.cuh): counts / starts / cursors per bucket, (key, index) staging
    uint32_t* dBucketCounts = nullptr;
    uint32_t* dBucketStarts = nullptr;
    uint32_t* dBucketCursors = nullptr;
    uint2* dBucketTmp = nullptr;
    bool bucketSortOff = true;       // the bucket sort is opt-in (AXCD_BUCKET_SORT=1): measured slower than the LSD sort
                                     // at 1 M keys so far (profiles/r02_experiments.md)
    bool travInline = false;         // AXCD_TRAV_INLINE=1: the traversal kernel with inline leaf tests
    bool splitNarrow = false;        // AXCD_SPLIT_NARROW=1: classify + closed forms + slots as separate kernels even
                                     // when the fused closed-form narrowphase applies (A/B measurements)
    Counters* dCtr = nullptr;
    Counters* dCtrBase = nullptr;    // two counter blocks: step k uses one, its refit kernel resets the other for step k+1
    int ctrParity = 0;
    Counters* dCtrInit = nullptr;    // per-step initial value of the counters (device copy: async reset)
    cudaEvent_t ev[EV_COUNT];
    bool evValid[EV_COUNT];
    // ---- CUDA graph of the fused step (axcd_step / axcd_step_async) --------------------------------
    // One executable graph per counter-block parity.  `gen` counts every call that changes what the
    // stage functions would launch (body count, filters, slab rule, ...); a graph is replayed only while
    // its generation is current, and captured only once the configuration has been stable for a step.
    struct StepGraph {
        cudaGraphExec_t exec = nullptr;
        uint64_t gen = 0;
        uint32_t launches[3] = {0, 0, 0};
        int sortedBuf = 0;
        bool reportsProgress = false;   // its narrowphase reports tile progress to the host (contact sink attached)
    } graph[2];
    uint64_t gen = 1, lastStepGen = 0;
    bool capturing = false;      // stage functions are being recorded into a graph: no event records
    bool graphsOff = false;      // AXCD_NO_GRAPH=1, or a capture failed once
    bool lastStepGraph = false;  // the last fused step was a graph launch (no per-stage times)
};

namespace {

int fail(AxcdContext* c, cudaError_t e, const char* what) {
    if (c) snprintf(c->lastErr, sizeof(c->lastErr), "%s: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
    if (e == cudaErrorMemoryAllocation) return AXCD_ERR_GPU_ALLOC;
    return AXCD_ERR_GPU_FAILED;
}

#define CU(call)                                                   \
    do {                                                           \
        cudaError_t _e = (call);                                   \
        if (_e != cudaSuccess) return fail(ctx, _e, #call);        \
    } while (0)

void slabRelease(AxcdContext* ctx);   // NCCL communicator and slab buffers (defined with the slab entry points)

template <typename T>
cudaError_t dalloc(T** p, size_t count) {
    return cudaMalloc(reinterpret_cast<void**>(p), sizeof(T) * (count ? count : 1));
}

int bitsFor(uint32_t n) {   // bits needed to represent values in [0, n)
    int b = 1;
    while (b < 32 && (1ull << b) < n) ++b;
    return b;
}

void recordEv(AxcdContext* c, int e) {
    if (c->capturing) return;   // timing events are not part of the step graph
    cudaEventRecord(c->ev[e], c->stream);
    c->evValid[e] = true;
}

float evMs(AxcdContext* c, int a, int b) {
    if (!c->evValid[a] || !c->evValid[b]) return 0.0f;
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]) != cudaSuccess) {
        cudaGetLastError();
        return 0.0f;
    }
    return ms;
}

// Upper levels of the range tree over the sorted leaf boxes (leaves at [P, 2P) are in place); returns the launch count.
// With sortedIdx the bottom pass also gathers the leaf records (segGatherBottomKernel).
uint32_t launchRangeTree(AxcdContext* ctx, cudaStream_t st, const uint32_t* sortedIdx = nullptr,
                         const uint32_t* sortedKeys = nullptr, int worldShift = 31) {
    const uint32_t P = ctx->segP;
    uint32_t launches = 1;
    const uint32_t bottomBlocks = P > (uint32_t)kSegLeaves ? P / kSegLeaves : 1;
    if (sortedIdx)
        segGatherBottomKernel<<<bottomBlocks, kSegThreads, 0, st>>>(ctx->dAabb, sortedIdx, sortedKeys, ctx->n, worldShift,
                                                                    ctx->dSegLo, ctx->dSegHi, P, ctx->dStepHist,
                                                                    (uint32_t)(kMaxPasses * kRadix));
    else
        segBuildBottomKernel<<<bottomBlocks, kSegThreads, 0, st>>>(ctx->dSegLo, ctx->dSegHi, P);
    uint32_t count = bottomBlocks;   // nodes in the level the bottom pass ended on
    while (count > 1) {
        ++launches;
        if (count <= (uint32_t)kSegLeaves) {
            segBuildTopKernel<<<1, kSegThreads, 0, st>>>(ctx->dSegLo, ctx->dSegHi, count);
            count = 1;
        } else {
            // a middle pass: treat the level as leaves of a smaller tree
            segBuildBottomKernel<<<count / kSegLeaves, kSegThreads, 0, st>>>(ctx->dSegLo, ctx->dSegHi, count);
            count /= kSegLeaves;
        }
    }
    return launches;
}

uint32_t sortTilesFor(uint64_t n) { return (uint32_t)((n + kSortTile - 1) / kSortTile); }

// Waits for the stream and brings the device counters (pair / contact counts) to the host.


}  // namespace

// Everything that depends on the body count: range-tree size (empty leaf slots beyond n) and the
// Morton resolution.
int resizeBodies(AxcdContext* ctx, uint32_t n) {
    ctx->n = n;
    ctx->gen++;
    ctx->fatValid = false;      // a different body set: no fat box or cached pair list carries over
    ctx->pairsCached = false;
    uint32_t P = 1;
    while (P < n) P <<= 1;
    ctx->segP = P;
    // leaf slots beyond n stay empty boxes; upper levels are rebuilt every step
    fillEmptyBoxesKernel<<<(2 * P + 255) / 256, 256, 0, ctx->stream>>>(ctx->dSegLo, ctx->dSegHi, 2 * P);
    CU(cudaGetLastError());
    ctx->worldBits = ctx->hasWorlds ? bitsFor(ctx->cfg.numWorlds) : 0;
    // Morton resolution: ~1 bit/axis finer than one body per cell, within the 32-bit key
    const uint32_t perWorld = ctx->hasWorlds ? (n / ctx->cfg.numWorlds + 1) : n;
    int mb = (bitsFor(perWorld > 1 ? perWorld : 2) + 2) / 3 + 1;
    const int room = (32 - ctx->worldBits) / 3;
    if (mb > room) mb = room;
    if (mb > 10) mb = 10;
    if (mb < 1) mb = 1;
    ctx->mortonBits = mb;
    return AXCD_OK;
}

// ---- contact sink of the fused narrowphase: host side ---------------------------------------------------------------
// The kernel reports, tile by tile and in pair order, how far the device contact array is complete (progress words in
// page-locked memory, see narrowClosedFusedKernel).  drainContactSink runs inside the next blocking call: it follows
// the words while the kernel is still running and hands every finished stretch (256 KB growing to 8 MB) to the copy engine on a second
// stream, so the contacts cross PCIe as DMA bursts behind the narrowphase instead of after it.
int drainContactSink(AxcdContext* ctx) {
    if (!ctx->sinkPending) return AXCD_OK;
    ctx->sinkPending = false;
    if (!ctx->sinkHost || !ctx->hProgress) return AXCD_OK;
    volatile uint32_t* pr = ctx->hProgress;
    const uint32_t maxTiles = ctx->progressWords - 1u;
    // chunk sizes double from 256 KB to 8 MB: the copy engine starts as soon as the first tiles are in, and the later,
    // larger bursts keep the per-copy set-up off the wire (profiles/r02_experiments.md: fixed 1 / 2 / 4 / 8 / 16 MB
    // chunks 2.28 / 2.27 / 2.23 / 2.22 / 2.23 ms per e2e step on a box where stores from the SMs gave 2.31)
    const uint32_t maxChunk = (8u << 20) / (uint32_t)sizeof(AxcdContact);
    uint32_t chunk = (256u << 10) / (uint32_t)sizeof(AxcdContact);
    uint32_t tiles = 0xffffffffu;   // unknown until the kernel has stored the pair count
    uint32_t next = 0, end = 0, copied = 0, spins = 0;
    bool computeDone = false;
    while (true) {
        if (tiles == 0xffffffffu) {
            const uint32_t np1 = pr[0];
            if (np1) tiles = (uint32_t)(((uint64_t)(np1 - 1u) + kFusedTile - 1) / kFusedTile);
        }
        const uint32_t limit = tiles == 0xffffffffu ? maxTiles : (tiles < maxTiles ? tiles : maxTiles);
        while (next < limit) {
            const uint32_t f = pr[1u + next];
            if (!f) break;
            end = f - 1u;
            ++next;
        }
        const bool all = tiles != 0xffffffffu && next >= limit;
        if (end > copied && (end - copied >= chunk || all)) {
            CU(cudaMemcpyAsync(ctx->sinkHost + copied, ctx->dContacts + copied, sizeof(AxcdContact) * (size_t)(end - copied),
                               cudaMemcpyDeviceToHost, ctx->copyStream));
            copied = end;
            chunk = chunk * 2u < maxChunk ? chunk * 2u : maxChunk;
        }
        if (all) break;
        if (computeDone) return AXCD_ERR_GPU_FAILED;   // the step has ended and the words are still incomplete
        if ((++spins & 1023u) == 0u) {
            const cudaError_t q = cudaStreamQuery(ctx->stream);
            if (q == cudaSuccess) computeDone = true;   // one more pass over the words, then give up
            else if (q != cudaErrorNotReady) CU(q);
        }
    }
    CU(cudaStreamSynchronize(ctx->copyStream));
    return AXCD_OK;
}

// Before a narrowphase that reports progress is launched: the previous step's sink is complete and the words are clear.
int armContactSink(AxcdContext* ctx) {
    const int rc = drainContactSink(ctx);
    if (rc) return rc;
    memset(const_cast<uint32_t*>(ctx->hProgress), 0, sizeof(uint32_t) * ctx->progressWords);
    ctx->sinkPending = true;
    return AXCD_OK;
}

int refreshCountersImpl(AxcdContext* ctx) {
    {
        const int rc = drainContactSink(ctx);
        if (rc) return rc;
    }
    CU(cudaMemcpyAsync(&ctx->hostCtr, ctx->dCtr, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->stage >= ST_BROAD) {
        ctx->foundPairs = ctx->hostCtr.pairCount;
        ctx->numPairs = ctx->foundPairs < ctx->cfg.maxPairs ? ctx->foundPairs : ctx->cfg.maxPairs;
    }
    if (ctx->stage >= ST_NARROW) {
        ctx->foundContacts = ctx->hostCtr.contactCount;
        ctx->numContacts = ctx->foundContacts < ctx->cfg.maxContacts ? ctx->foundContacts : ctx->cfg.maxContacts;
    }
    return AXCD_OK;
}

extern "C" {

void axcd_default_config(AxcdConfig* cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->maxBodies = 1024;
    cfg->maxPairs = 8192;
    cfg->maxContacts = 8192;
    cfg->maxHullVerts = 0;
    cfg->numWorlds = 1;
    cfg->aabbMargin = 0.0f;
    cfg->gjkMaxIters = 32;
    cfg->epaMaxIters = 32;
    cfg->epaMaxFaces = 64;
    cfg->gjkTol = 1e-6f;
    cfg->epaTol = 1e-4f;
    cfg->flags = 0;
    cfg->deviceOrdinal = 0;
    cfg->stream = nullptr;
}

const char* axcd_error_string(int32_t code) {
    // text of axiom::core::errorCodeToString (reference: src/core/error_code.cpp:5-62)
    switch (code) {
        case 0: return "Success";
        case 100: return "Division by zero";
        case 101: return "Cannot normalize zero vector";
        case 102: return "Matrix is singular and cannot be inverted";
        case 103: return "Quaternion is invalid (not normalized or contains NaN)";
        case 200: return "Out of memory";
        case 201: return "Invalid allocation parameters";
        case 202: return "Unexpected null pointer";
        case 300: return "Invalid collision shape";
        case 301: return "GJK algorithm failed to converge";
        case 302: return "EPA algorithm failed to converge";
        case 400: return "Invalid rigid body properties";
        case 401: return "Constraint solver failed to converge";
        case 500: return "Vulkan initialization failed";
        case 501: return "Shader compilation failed";
        case 502: return "GPU buffer allocation failed";
        case 503: return "Invalid GPU operation or state";
        case 504: return "GPU operation timed out";
        case 505: return "GPU operation failed";
        case 600: return "Invalid parameter";
        case 601: return "Value out of range";
        default: return "Unknown error";
    }
}

int32_t axcd_pin_host_buffer(void* hostPtr, uint64_t bytes) {
    if (!hostPtr) return AXCD_ERR_NULL_POINTER;
    if (bytes == 0) return AXCD_ERR_INVALID_PARAM;
    const cudaError_t e = cudaHostRegister(hostPtr, (size_t)bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? AXCD_ERR_GPU_ALLOC : AXCD_ERR_GPU_FAILED;
    }
    return AXCD_OK;
}

int32_t axcd_unpin_host_buffer(void* hostPtr) {
    if (!hostPtr) return AXCD_ERR_NULL_POINTER;
    if (cudaHostUnregister(hostPtr) != cudaSuccess) {
        cudaGetLastError();
        return AXCD_ERR_GPU_FAILED;
    }
    return AXCD_OK;
}

const char* axcd_last_device_error(AxcdContext* ctx) { return ctx ? ctx->lastErr : ""; }

int32_t axcd_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void axcd_destroy(AxcdContext* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    void* bufs[] = {ctx->dPoseStage, ctx->dXf, ctx->dShapes, ctx->dType8, ctx->dHull, ctx->dWorld, ctx->dBodyKeys, ctx->dFilters, ctx->dAwake, ctx->dGhostSend, ctx->dGhostCount, ctx->dAabb, ctx->dKeys[0], ctx->dKeys[1],
                    ctx->dVals[0], ctx->dVals[1], ctx->dSegLo, ctx->dSegHi, ctx->dNodes, ctx->dNodes32,
                    ctx->dWorldEnd, ctx->dPairsTmp, ctx->dPairs, ctx->dBodyCount, ctx->dBodyStart, ctx->dSegB, ctx->dScanStatus, ctx->dEpaWork,
                    ctx->dEpaOverflow, ctx->dEpaSpill, ctx->dSlotStatus, ctx->dChunks, ctx->dFlags, ctx->dSlots, ctx->dTmpContacts, ctx->dContacts, ctx->dManifolds, ctx->dQIn, ctx->dQCount, ctx->dQSeg, ctx->dQOut, ctx->dPairDist, ctx->dSortHist, ctx->dStepHist,
                    ctx->dSortStatus, ctx->dCtrBase, ctx->dCtrInit, ctx->dBucketCounts, ctx->dBucketStarts, ctx->dBucketCursors,
                    ctx->dBucketTmp};
    for (void* b : bufs)
        if (b) cudaFree(b);
    for (int i = 0; i < EV_COUNT; ++i)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (auto& g : ctx->graph)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    slabRelease(ctx);
    if (ctx->copyStream) cudaStreamDestroy(ctx->copyStream);
    if (ctx->hProgress) cudaFreeHost(const_cast<uint32_t*>(ctx->hProgress));
    if (ctx->ownStream && ctx->stream) cudaStreamDestroy(ctx->stream);
    cudaGetLastError();
    delete ctx;
}

int32_t axcd_create(const AxcdConfig* cfg, AxcdContext** out) {
    if (!cfg || !out) return AXCD_ERR_NULL_POINTER;
    *out = nullptr;
    if (cfg->maxBodies == 0 || cfg->maxBodies > (1u << 28) || cfg->maxPairs == 0 || cfg->maxPairs > (1u << 30) ||
        cfg->maxContacts == 0 || cfg->numWorlds == 0 || cfg->numWorlds > (1u << 20) || cfg->epaMaxFaces < 4 ||
        !(cfg->gjkTol >= 0.0f) || !(cfg->epaTol >= 0.0f) || !(cfg->aabbMargin == cfg->aabbMargin))
        return AXCD_ERR_INVALID_PARAM;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return AXCD_ERR_GPU_INIT;   // no device: there is no CPU path to fall back to
    }
    if (cfg->deviceOrdinal < 0 || cfg->deviceOrdinal >= ndev) return AXCD_ERR_INVALID_PARAM;
    if (cudaSetDevice(cfg->deviceOrdinal) != cudaSuccess) {
        cudaGetLastError();
        return AXCD_ERR_GPU_INIT;
    }
    AxcdContext* ctx = new (std::nothrow) AxcdContext();
    if (!ctx) return AXCD_ERR_OUT_OF_MEMORY;
    ctx->cfg = *cfg;
    {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->deviceOrdinal) == cudaSuccess && sms > 0)
            ctx->numSMs = sms;
        cudaGetLastError();
    }
    if (ctx->cfg.epaMaxFaces > (uint32_t)kEpaHardFaces) ctx->cfg.epaMaxFaces = kEpaHardFaces;
    {
        const char* ng = getenv("AXCD_NO_GRAPH");
        if (ng && ng[0] == '1') ctx->graphsOff = true;
        const char* ti = getenv("AXCD_TRAV_INLINE");
        if (ti && ti[0] == '1') ctx->travInline = true;
        const char* sn = getenv("AXCD_SPLIT_NARROW");
        if (sn && sn[0] == '1') ctx->splitNarrow = true;
        const char* nbs = getenv("AXCD_BUCKET_SORT");
        if (nbs && nbs[0] == '1') ctx->bucketSortOff = false;
    }
    for (int i = 0; i < EV_COUNT; ++i) {
        ctx->ev[i] = nullptr;
        ctx->evValid[i] = false;
    }
    int rc = AXCD_OK;
    auto body = [&]() -> int {
        if (cfg->stream) {
            ctx->stream = static_cast<cudaStream_t>(cfg->stream);
        } else {
            CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
            ctx->ownStream = true;
        }
        for (int i = 0; i < EV_COUNT; ++i) CU(cudaEventCreate(&ctx->ev[i]));
        const size_t nb = cfg->maxBodies, np = cfg->maxPairs;
        // +64 floats of slack: the staged 128-bit loads may touch the tail of the last block
        CU(dalloc(&ctx->dXf, nb * 10 + 64));
        CU(dalloc(&ctx->dPoseStage, nb * 7));
        CU(dalloc(&ctx->dShapes, nb));
        CU(dalloc(&ctx->dType8, nb));
        CU(dalloc(&ctx->dHull, (size_t)cfg->maxHullVerts));
        if (cfg->numWorlds > 1) {
            CU(dalloc(&ctx->dWorld, nb));
            CU(dalloc(&ctx->dWorldEnd, (size_t)cfg->numWorlds));
        }
        CU(dalloc(&ctx->dAabb, nb * 6 + 64));
        CU(dalloc(&ctx->dBodyKeys, nb));
        for (int k = 0; k < 2; ++k) {
            CU(dalloc(&ctx->dKeys[k], nb));
            CU(dalloc(&ctx->dVals[k], nb));
        }
        CU(dalloc(&ctx->dPairsTmp, np));
        CU(dalloc(&ctx->dPairs, np));
        CU(dalloc(&ctx->dBodyCount, 2 * nb));
        CU(dalloc(&ctx->dBodyStart, nb + 1));
        CU(dalloc(&ctx->dSegB, np));
        CU(dalloc(&ctx->dScanStatus, nb / kScanTile + 2 + 128));
        {
            size_t P = 1;
            while (P < nb) P <<= 1;
            CU(dalloc(&ctx->dSegLo, 2 * P));
            CU(dalloc(&ctx->dSegHi, 2 * P));
        }
        CU(dalloc(&ctx->dNodes32, nb));
        // scene-query buffers up front (float nodes, scratch for 64 k queries / 1 M hits): no first-use allocation
        // on the query path; the scratch still grows on demand for larger batches
        CU(dalloc(&ctx->dNodes, nb));
        {
            const size_t nq0 = 65536, hits0 = 1u << 20;
            CU(cudaMalloc(&ctx->dQIn, nq0 * 32));        ctx->qInBytes = nq0 * 32;
            CU(cudaMalloc(&ctx->dQCount, nq0 * 16));     ctx->qCountBytes = nq0 * 16;
            CU(cudaMalloc(&ctx->dQSeg, hits0 * 4));      ctx->qSegBytes = hits0 * 4;
            CU(cudaMalloc(&ctx->dQOut, hits0 * 8));      ctx->qOutBytes = hits0 * 8;
        }
        {
            // CUDA loads a kernel's code at its first launch; ask for the scene-query and manifold kernels now, so that
            // the first query after a step does not pay for it (measured: 20-300 ms on a freshly started box)
            cudaFuncAttributes fa;
            const void* fns[] = {(const void*)buildTopologyKernel<false>, (const void*)queryAabbKernel<false>,
                                 (const void*)queryAabbKernel<true>,      (const void*)raycastKernel<false>,
                                 (const void*)raycastKernel<true>,        (const void*)manifoldKernel};
            for (const void* f : fns) cudaFuncGetAttributes(&fa, f);
            cudaGetLastError();
        }
        CU(dalloc(&ctx->dEpaWork, (size_t)cfg->maxContacts));
        CU(dalloc(&ctx->dEpaOverflow, (size_t)cfg->maxContacts));
        ctx->spillCap = cfg->maxContacts < 65536u ? cfg->maxContacts : 65536u;
        CU(dalloc(&ctx->dEpaSpill, (size_t)ctx->spillCap * (Poly<kEpaFastVerts, kEpaFastFaces, kEpaFastEdges, 1>::kWords + kSpillStateWords)));
        // one status word per tile of the slot scan (2048 pairs) or of the fused narrowphase (1024 pairs), plus the
        // slack the 128-wide look-back may read past the last tile
        CU(dalloc(&ctx->dSlotStatus, np / kFusedTile + 2 + 128));
        CU(dalloc(&ctx->dChunks, (size_t)chunkCapFor(cfg->maxPairs) * 32));
        CU(dalloc(&ctx->dFlags, np + kSlotTile));
        CU(dalloc(&ctx->dSlots, np));
        CU(dalloc(&ctx->dTmpContacts, np));
        CU(cudaFuncSetAttribute(epaKernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEpaSmemBytes));
        CU(cudaFuncSetAttribute(epaKernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEpaSmemBytes));
        CU(dalloc(&ctx->dContacts, (size_t)cfg->maxContacts));
        if (cfg->flags & AXCD_FLAG_PAIR_DISTANCES) CU(dalloc(&ctx->dPairDist, np));
        CU(dalloc(&ctx->dSortHist, (size_t)kMaxPasses * kRadix));
        CU(dalloc(&ctx->dStepHist, (size_t)kMaxPasses * kRadix));
        CU(cudaMemsetAsync(ctx->dStepHist, 0, sizeof(uint32_t) * kMaxPasses * kRadix, ctx->stream));
        const size_t maxTiles = sortTilesFor(nb > np ? nb : np);
        CU(dalloc(&ctx->dSortStatus, (size_t)kMaxPasses * maxTiles * kRadix));
        CU(dalloc(&ctx->dBucketCounts, (size_t)(1u << kMaxBucketBits)));
        CU(dalloc(&ctx->dBucketStarts, (size_t)(1u << kMaxBucketBits) + 1));
        CU(dalloc(&ctx->dBucketCursors, (size_t)(1u << kMaxBucketBits)));
        CU(dalloc(&ctx->dBucketTmp, nb));
        CU(cudaMemsetAsync(ctx->dBucketCounts, 0, sizeof(uint32_t) << kMaxBucketBits, ctx->stream));
        CU(dalloc(&ctx->dCtrBase, 2));
        ctx->dCtr = ctx->dCtrBase;
        CU(dalloc(&ctx->dCtrInit, 1));
        {
            Counters init;
            memset(&init, 0, sizeof(init));
            for (int k = 0; k < 3; ++k) {
                init.boundsMin[k] = 0xffffffffu;   // ordered-uint encodings of +inf / -inf
                init.boundsMax[k] = 0u;
            }
            CU(cudaMemcpy(ctx->dCtrInit, &init, sizeof(Counters), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(ctx->dCtrBase + 1, &init, sizeof(Counters), cudaMemcpyHostToDevice));
        }
        CU(cudaMemsetAsync(ctx->dCtr, 0, sizeof(Counters), ctx->stream));
        // slotKernel reads the per-pair flags 8 bytes at a time and masks the tail: keep the tail defined
        CU(cudaMemsetAsync(ctx->dFlags, 0, np + kSlotTile, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        return AXCD_OK;
    };
    rc = body();
    if (rc != AXCD_OK) {
        axcd_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return AXCD_OK;
}

int32_t axcd_set_shapes(AxcdContext* ctx, const AxcdShape* shapes, uint32_t n, const float* hullXYZ,
                        uint32_t nHullVerts, const uint32_t* worldId) {
    if (!ctx) return AXCD_ERR_NULL_POINTER;
    if (n && !shapes) return AXCD_ERR_NULL_POINTER;
    if (nHullVerts && !hullXYZ) return AXCD_ERR_NULL_POINTER;
    if (n > ctx->cfg.maxBodies || nHullVerts > ctx->cfg.maxHullVerts) return AXCD_ERR_OUT_OF_RANGE;
    if (ctx->cfg.numWorlds > 1 && n && !worldId) return AXCD_ERR_INVALID_PARAM;
    bool anyHull = false, anyCapsule = false, anyCylinder = false;
    for (uint32_t i = 0; i < n; ++i) {
        const AxcdShape& s = shapes[i];
        if (s.type == AXCD_SHAPE_CONVEX) {
            anyHull = true;
            uint32_t first, cnt;
            memcpy(&first, &s.p0, 4);
            memcpy(&cnt, &s.p1, 4);
            if (cnt == 0 || cnt > 65535u || (uint64_t)first + cnt > nHullVerts) return AXCD_ERR_INVALID_SHAPE;
        } else if (s.type != AXCD_SHAPE_SPHERE && s.type != AXCD_SHAPE_BOX && s.type != AXCD_SHAPE_CAPSULE &&
                   s.type != AXCD_SHAPE_CYLINDER) {
            return AXCD_ERR_INVALID_SHAPE;   // Plane / Mesh are not in scope
        }
        if (s.type == AXCD_SHAPE_CAPSULE) anyCapsule = true;
        if (s.type == AXCD_SHAPE_CYLINDER) {
            anyCylinder = true;
            if (!(s.p0 >= 0.0f && s.p0 < 3.0e38f && s.p1 >= 0.0f && s.p1 < 3.0e38f)) return AXCD_ERR_INVALID_SHAPE;
        }
        // sizes must be finite and non-negative (a flat box or a zero radius is a valid degenerate shape)
        if (s.type == AXCD_SHAPE_SPHERE && !(s.p0 >= 0.0f && s.p0 < 3.0e38f)) return AXCD_ERR_INVALID_SHAPE;
        if (s.type == AXCD_SHAPE_BOX && !(s.p0 >= 0.0f && s.p0 < 3.0e38f && s.p1 >= 0.0f && s.p1 < 3.0e38f && s.p2 >= 0.0f && s.p2 < 3.0e38f))
            return AXCD_ERR_INVALID_SHAPE;
        if (s.type == AXCD_SHAPE_CAPSULE && !(s.p0 >= 0.0f && s.p0 < 3.0e38f && s.p1 >= 0.0f && s.p1 < 3.0e38f)) return AXCD_ERR_INVALID_SHAPE;
        if (ctx->cfg.numWorlds > 1 && worldId[i] >= ctx->cfg.numWorlds) return AXCD_ERR_OUT_OF_RANGE;
    }
    ctx->hasHulls = anyHull;
    ctx->hasSweptRayShapes = anyHull || anyCylinder;
    ctx->hasCylinders = anyCylinder;
    ctx->hasGenericShapes = anyHull || anyCapsule || anyCylinder;
    ctx->gen++;
    ctx->filtersOn = false;   // per-body filter words describe the previous body set: set them again
    if (ctx->awakeOn && ctx->dAwake) cudaMemsetAsync(ctx->dAwake, 1, ctx->cfg.maxBodies, ctx->stream);
    ctx->awakeOn = false;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    CU(cudaMemcpyAsync(ctx->dShapes, shapes, sizeof(AxcdShape) * n, cudaMemcpyHostToDevice, ctx->stream));
    if (nHullVerts) {
        float4* tmp = static_cast<float4*>(malloc(sizeof(float4) * nHullVerts));
        if (!tmp) return AXCD_ERR_OUT_OF_MEMORY;
        for (uint32_t i = 0; i < nHullVerts; ++i)
            tmp[i] = make_float4(hullXYZ[3 * i], hullXYZ[3 * i + 1], hullXYZ[3 * i + 2], 0.0f);
        cudaError_t e = cudaMemcpyAsync(ctx->dHull, tmp, sizeof(float4) * nHullVerts, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        free(tmp);
        if (e != cudaSuccess) return fail(ctx, e, "hull upload");
    }
    ctx->hasWorlds = ctx->cfg.numWorlds > 1;
    if (ctx->hasWorlds)
        CU(cudaMemcpyAsync(ctx->dWorld, worldId, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->nHull = nHullVerts;
    ctx->nOwned = n;
    {
        const int rc = resizeBodies(ctx, n);
        if (rc) return rc;
    }
    ctx->fatValid = false;      // new bodies: every fat box is rebuilt by the next refit
    ctx->pairsCached = false;
    ctx->stage = ST_SHAPES;
    return AXCD_OK;
}

int32_t axcd_set_transforms(AxcdContext* ctx, const void* transforms, uint32_t n, uint32_t strideBytes) {
    if (!ctx) return AXCD_ERR_NULL_POINTER;
    if (ctx->stage < ST_SHAPES) return AXCD_ERR_GPU_INVALID_OP;
    if ((n != ctx->n && n != ctx->nOwned) || strideBytes < 40 || (strideBytes & 3u)) return AXCD_ERR_INVALID_PARAM;
    if (n && !transforms) return AXCD_ERR_NULL_POINTER;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    if (n) {
        if (strideBytes == 40) {
            CU(cudaMemcpyAsync(ctx->dXf, transforms, (size_t)n * 40, cudaMemcpyHostToDevice, ctx->stream));
        } else {
            CU(cudaMemcpy2DAsync(ctx->dXf, 40, transforms, strideBytes, 40, n, cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    ctx->stage = ST_POSES;
    return AXCD_OK;
}

int32_t axcd_set_poses(AxcdContext* ctx, const void* poses, uint32_t n, uint32_t strideBytes) {
    if (!ctx) return AXCD_ERR_NULL_POINTER;
    if (ctx->stage < ST_POSES) return AXCD_ERR_GPU_INVALID_OP;   // the scales come from an earlier axcd_set_transforms
    if ((n != ctx->n && n != ctx->nOwned) || strideBytes < 28 || (strideBytes & 3u)) return AXCD_ERR_INVALID_PARAM;
    if (n && !poses) return AXCD_ERR_NULL_POINTER;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    if (n) {
        if (strideBytes == 28) {
            CU(cudaMemcpyAsync(ctx->dPoseStage, poses, (size_t)n * 28, cudaMemcpyHostToDevice, ctx->stream));
        } else {
            CU(cudaMemcpy2DAsync(ctx->dPoseStage, 28, poses, strideBytes, 28, n, cudaMemcpyHostToDevice, ctx->stream));
        }
        const uint32_t words = n * 7u;
        mergePosesKernel<<<(words + 255u) / 256u, 256, 0, ctx->stream>>>(ctx->dPoseStage, ctx->dXf, words);
        CU(cudaGetLastError());
    }
    ctx->stage = ST_POSES;
    return AXCD_OK;
}

int32_t axcd_refit(AxcdContext* ctx) {
    if (!ctx) return AXCD_ERR_NULL_POINTER;
    if (ctx->stage < ST_POSES) return AXCD_ERR_GPU_INVALID_OP;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    recordEv(ctx, EV_START);
    if (!ctx->capturing) ctx->lastStepGraph = false;   // a staged (or direct) step: per-stage events are recorded
    ctx->manifoldsValid = false;
    // temporal coherence: the skip decision reads the moved-body count of the LAST refit only.  If the
    // previous refit was never consumed by a broadphase, the bodies it moved are not in the cached pair
    // list and this refit will not count them again: the cache cannot be trusted.
    if (ctx->stage == ST_REFIT) ctx->pairsCached = false;
    // reset per-step counters: bounds (min = +inf encoding, max = -inf encoding) and counts
    // per-step counters (scene bounds at +-inf encodings, counts at zero): this step takes the block the
    // previous refit kernel reset, and its own refit kernel resets the other one for the next step — no
    // memcpy node in front of the first kernel
    ctx->ctrParity ^= 1;
    ctx->dCtr = ctx->dCtrBase + ctx->ctrParity;
    Counters* ctrNext = ctx->dCtrBase + (ctx->ctrParity ^ 1);
    if (!ctx->n) CU(cudaMemcpyAsync(ctrNext, ctx->dCtrInit, sizeof(Counters), cudaMemcpyDeviceToDevice, ctx->stream));
    if (ctx->n) {
        const uint32_t blocks = (ctx->n + kRefitThreads - 1) / kRefitThreads;
        if (ctx->cfg.flags & AXCD_FLAG_TEMPORAL_COHERENCE) {
            refitKernel<true><<<blocks, kRefitThreads, 0, ctx->stream>>>(
                reinterpret_cast<const float4*>(ctx->dXf), ctx->dShapes, ctx->dHull,
                reinterpret_cast<float4*>(ctx->dAabb), ctx->dType8, ctx->n, ctx->cfg.aabbMargin,
                ctx->fatValid ? 0u : 1u, ctx->dCtr, ctrNext, ctx->dCtrInit);
            ctx->fatValid = true;
        } else {
            if (ctx->cfg.flags & AXCD_FLAG_REFIT_MAT4_ROUTE) {
                // boxes through AABB::transform(Transform::toMatrix()) (row a15): the block-staged kernel
                refitKernel<false, true><<<blocks, kRefitThreads, 0, ctx->stream>>>(
                    reinterpret_cast<const float4*>(ctx->dXf), ctx->dShapes, ctx->dHull,
                    reinterpret_cast<float4*>(ctx->dAabb), ctx->dType8, ctx->n, ctx->cfg.aabbMargin, 0u, ctx->dCtr,
                    ctrNext, ctx->dCtrInit);
            } else {
#if defined(AXCD_REFIT_NO_TMA)
            // diagnostic build: the block-staged kernel (compute-sanitizer's initcheck does not see the
            // bulk-store writes of the TMA kernel and reports every later read of the AABBs)
            refitKernel<false><<<blocks, kRefitThreads, 0, ctx->stream>>>(
                reinterpret_cast<const float4*>(ctx->dXf), ctx->dShapes, ctx->dHull,
                reinterpret_cast<float4*>(ctx->dAabb), ctx->dType8, ctx->n, ctx->cfg.aabbMargin, 0u, ctx->dCtr,
                ctrNext, ctx->dCtrInit);
#else
            (void)blocks;
            uint32_t tiles = (ctx->n + kRefitThreads - 1) / kRefitThreads;
            const uint32_t grid = tiles < (uint32_t)ctx->numSMs * kRefitTmaBlocksPerSM ? tiles : ctx->numSMs * kRefitTmaBlocksPerSM;
            refitTmaKernel<<<grid, kRefitThreads, 0, ctx->stream>>>(
                ctx->dXf, ctx->dShapes, ctx->dHull, ctx->dAabb, ctx->dType8, ctx->n, ctx->cfg.aabbMargin, ctx->dCtr,
                ctrNext, ctx->dCtrInit);
#endif
            }
        }
        CU(cudaGetLastError());
    }
    ctx->launches[0] = ctx->n ? 1 : 0;
    recordEv(ctx, EV_REFIT);
    ctx->stage = ST_REFIT;
    return AXCD_OK;
}

int32_t axcd_broadphase(AxcdContext* ctx) {
    if (!ctx) return AXCD_ERR_NULL_POINTER;
    // the per-step device counters are reset by axcd_refit: each stage runs once per refit
    if (ctx->stage != ST_REFIT) return AXCD_ERR_GPU_INVALID_OP;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    const uint32_t n = ctx->n;
    cudaStream_t st = ctx->stream;
    ctx->numPairs = ctx->foundPairs = 0;
    ctx->launches[1] = 0;
    ctx->broadSkipped = false;
    uint32_t bucketLaunches = 0;
    if ((ctx->cfg.flags & AXCD_FLAG_TEMPORAL_COHERENCE) && n >= 2) {
        // Temporal coherence: the candidate set is a function of the fat boxes only.  If no body left
        // its fat box since the cached broadphase, the cached canonical pair list is still exact.
        uint32_t moved = 0;
        CU(cudaMemcpyAsync(&moved, &ctx->dCtr->movedBodies, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        ctx->movedLast = moved;
        if (moved == 0 && ctx->pairsCached) {
            // the per-step counter reset cleared the device-side pair count: restore it
            CU(cudaMemcpyAsync(&ctx->dCtr->pairCount, &ctx->cachedFound, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            for (int e : {EV_SORT, EV_BUILD, EV_PAIR, EV_PAIRSORT, EV_BROAD_END}) recordEv(ctx, e);
            ctx->broadSkipped = true;
            ctx->stage = ST_BROAD;
            return AXCD_OK;
        }
    }
    if (n >= 2) {
        // ---- Morton keys + radix sort ----------------------------------------------------------
        const uint32_t blocks = (n + kRefitThreads - 1) / kRefitThreads;
        const int keyBits = 3 * ctx->mortonBits + ctx->worldBits;
        const int passes = (keyBits + 7) / 8;
        // scratch the later kernels expect zeroed: cleared by the Morton kernel, not by memset nodes
        const uint32_t scanTilesZ = (n + kScanTile - 1) / kScanTile;
        const uint32_t slotTilesZ = (ctx->cfg.maxPairs + kFusedTile - 1) / kFusedTile + 128;
        ZeroList zl;
        zl.ptr[0] = ctx->dBodyCount;   zl.words[0] = n;
        zl.ptr[1] = ctx->dScanStatus;  zl.words[1] = scanTilesZ + 1 + 128;
        zl.ptr[2] = ctx->dSlotStatus;  zl.words[2] = slotTilesZ + 1;
        zl.ptr[3] = ctx->dSortHist;    zl.words[3] = 0;   // (the step's histograms live in dStepHist: see mortonKernel)
        zl.ptr[4] = ctx->dSortStatus;  zl.words[4] = (uint32_t)passes * sortTilesFor(n) * kRadix;
        // Morton keys, then the bucket sort (one MSD pass + per-bucket shared-memory sorts); the LSD radix kernels
        // are launched behind it and run only if a bucket overflowed (device-side flag, no host round trip)
        const BucketPlan bp = ctx->bucketSortOff ? BucketPlan{0, 0} : bucketPlanFor(n, keyBits);
        const uint32_t mortonBlocks = blocks < (uint32_t)ctx->numSMs * 4u ? blocks : (uint32_t)ctx->numSMs * 4u;
        mortonKernel<<<mortonBlocks, kRefitThreads, 0, st>>>(reinterpret_cast<const float4*>(ctx->dAabb),
                                                             ctx->hasWorlds ? ctx->dWorld : nullptr, ctx->dKeys[0],
                                                             ctx->dVals[0], n, ctx->mortonBits, ctx->dCtr, zl,
                                                             bp.bucketBits ? ctx->dBucketCounts : nullptr, bp.shift,
                                                             ctx->dStepHist, passes);
        CU(cudaGetLastError());
        const uint32_t* lsdEnable = nullptr;
        bucketLaunches = 0;
        if (bp.bucketBits) {
            const uint32_t nbk = 1u << bp.bucketBits;
            const int outBuf = passes & 1;   // where the LSD sort would leave its result: everything downstream reads that
            bucketScanKernel<<<1, 1024, 0, st>>>(ctx->dBucketCounts, ctx->dBucketStarts, ctx->dBucketCursors, nbk,
                                                 &ctx->dCtr->sortFallback, &ctx->dCtr->sortMaxBucket);
            bucketScatterKernel<<<ctx->numSMs * 8, 256, 0, st>>>(ctx->dKeys[0], n, bp.shift, ctx->dBucketCursors, ctx->dBucketTmp,
                                                                 &ctx->dCtr->sortFallback);
            const uint32_t sortBlocks = nbk < (uint32_t)ctx->numSMs * 16 ? nbk : (uint32_t)ctx->numSMs * 16;
            bucketSortKernel<<<sortBlocks, kBucketThreads, 0, st>>>(ctx->dBucketTmp, ctx->dBucketStarts, nbk, ctx->dKeys[outBuf],
                                                                    ctx->dVals[outBuf], &ctx->dCtr->sortFallback);
            CU(cudaGetLastError());
            lsdEnable = &ctx->dCtr->sortFallback;
            bucketLaunches = 3;
        }
        // the radix tickets live in the per-step counter block, which the refit kernel reset
        const int sb = radixSort<uint32_t, true>(ctx->dKeys[0], ctx->dKeys[1], ctx->dVals[0], ctx->dVals[1], n, 0,
                                                 passes, ctx->dStepHist, ctx->dSortStatus, ctx->dCtr->sortTicket, st,
                                                 ctx->numSMs, true, lsdEnable, true);
        CU(cudaGetLastError());
        recordEv(ctx, EV_SORT);
        ctx->sortedBuf = sb;
        const uint32_t* sKeys = ctx->dKeys[sb];
        const uint32_t* sVals = ctx->dVals[sb];
        // ---- LBVH ------------------------------------------------------------------------------
        const int worldShift = 3 * ctx->mortonBits;
        const uint32_t b256 = (n + 255) / 256;
        const uint32_t P = ctx->segP;
        float4* leafLo = ctx->dSegLo + P;
        float4* leafHi = ctx->dSegHi + P;
        // leaf gather + bottom levels of the range tree in one kernel, then its upper levels
        const uint32_t segLaunches = launchRangeTree(ctx, st, sVals, sKeys, ctx->hasWorlds ? worldShift : 31);
        if (ctx->hasWorlds) markWorldEndsKernel<<<b256, 256, 0, st>>>(sKeys, n, worldShift, ctx->dWorldEnd);
        buildTopologyKernel<true><<<(n + AXCD_TOPO_THREADS - 1) / AXCD_TOPO_THREADS, AXCD_TOPO_THREADS, 0, st>>>(sKeys, n, ctx->dSegLo, ctx->dSegHi, P, ctx->dNodes32);
        ctx->queryNodesValid = false;
        CU(cudaGetLastError());
        recordEv(ctx, EV_BUILD);
        // ---- traversal ---------------------------------------------------------------------------
        const uint32_t tb = (n + kTravThreads - 1) / kTravThreads;
        const SlabRule slabRule{ctx->slabOn ? 1 : 0, ctx->slabLo, ctx->slabHi, ctx->dBodyKeys};
        if (ctx->travInline)   // AXCD_TRAV_INLINE=1: leaf tests inline in the walk (the round-1 kernel; A/B measurements)
            findPairsKernel<<<tb, kTravThreads, 0, st>>>(leafLo, leafHi, ctx->dNodes32, ctx->dSegLo, ctx->dSegHi,
                                                         ctx->hasWorlds ? ctx->dWorldEnd : nullptr, n, ctx->dPairsTmp,
                                                         ctx->cfg.maxPairs, ctx->dBodyCount, slabRule,
                                                         ctx->filtersOn ? ctx->dFilters : nullptr,
                                                         ctx->awakeOn ? ctx->dAwake : nullptr, ctx->dCtr);
        else if (ctx->hasWorlds)
            findPairsDenseKernel<true><<<tb, kTravThreads, 0, st>>>(leafLo, leafHi, ctx->dNodes32, ctx->dSegLo, ctx->dSegHi,
                                                                    ctx->dWorldEnd, n, ctx->dPairsTmp,
                                                                    ctx->cfg.maxPairs, ctx->dBodyCount, slabRule,
                                                                    ctx->filtersOn ? ctx->dFilters : nullptr,
                                                                    ctx->awakeOn ? ctx->dAwake : nullptr, ctx->dCtr);
        else
            findPairsDenseKernel<false><<<tb, kTravThreads, 0, st>>>(leafLo, leafHi, ctx->dNodes32, ctx->dSegLo, ctx->dSegHi,
                                                                     nullptr, n, ctx->dPairsTmp,
                                                                     ctx->cfg.maxPairs, ctx->dBodyCount, slabRule,
                                                                     ctx->filtersOn ? ctx->dFilters : nullptr,
                                                                     ctx->awakeOn ? ctx->dAwake : nullptr, ctx->dCtr);
        CU(cudaGetLastError());
        recordEv(ctx, EV_PAIR);
        // ---- canonical order: counting sort by body a, then tiny per-body sorts by b ------------------
        const uint32_t scanTiles = (n + kScanTile - 1) / kScanTile;
        // the scan writes the segment starts twice: dBodyStart stays, the copy is the scatter's fill cursor
        exclusiveScanKernel<<<scanTiles, kScanThreads, 0, st>>>(ctx->dBodyCount, ctx->dBodyStart, ctx->dBodyCount + n, n,
                                                                ctx->dScanStatus, &ctx->dCtr->scanTicket,
                                                                &ctx->dCtr->storedPairs);
        scatterPairsKernel<<<ctx->numSMs * AXCD_SCATTER_BLOCKS, 256, 0, st>>>(ctx->dPairsTmp, &ctx->dCtr->pairCount, ctx->cfg.maxPairs,
                                                        ctx->dBodyCount + n, ctx->dSegB);
        sortSegmentsCoopKernel<<<(n + kCoopBodies - 1) / kCoopBodies, kCoopBodies, 0, st>>>(ctx->dBodyStart, ctx->dBodyCount, n, ctx->dSegB, ctx->dPairs);
        CU(cudaGetLastError());
        recordEv(ctx, EV_PAIRSORT);
        // no host round trip here: the narrowphase kernels read the pair count on the device
        // morton, (hist, scan, passes), gather, [worldEnds], range tree, topology+fit, traversal, scan, scatter,
        // segment sort
        ctx->launches[1] = 1 + bucketLaunches + passes + 1 + segLaunches + (ctx->hasWorlds ? 1 : 0) + 1 + 3;
    } else {
        recordEv(ctx, EV_SORT);
        recordEv(ctx, EV_BUILD);
        recordEv(ctx, EV_PAIR);
        recordEv(ctx, EV_PAIRSORT);
    }
    recordEv(ctx, EV_BROAD_END);
    ctx->stage = ST_BROAD;
    if (ctx->cfg.flags & AXCD_FLAG_TEMPORAL_COHERENCE) {
        // remember the pair count of this broadphase for the steps that can reuse it
        CU(cudaMemcpyAsync(&ctx->cachedFound, &ctx->dCtr->pairCount, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        ctx->pairsCached = true;
    }
    return AXCD_OK;
}

int32_t axcd_narrowphase(AxcdContext* ctx) {
    if (!ctx) return AXCD_ERR_NULL_POINTER;
    if (ctx->stage != ST_BROAD) return AXCD_ERR_GPU_INVALID_OP;   // once per broadphase (see above)
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    cudaStream_t st = ctx->stream;
    recordEv(ctx, EV_N0);
    ctx->narrowReportsProgress = false;
    ctx->numContacts = ctx->foundContacts = 0;
    ctx->manifoldsValid = false;
    ctx->launches[2] = 0;
    if (ctx->n >= 2) {
        NarrowParams p;
        p.gjkMaxIters = ctx->cfg.gjkMaxIters;
        p.epaMaxIters = ctx->cfg.epaMaxIters;
        p.epaMaxFaces = ctx->cfg.epaMaxFaces;
        p.gjkTol = ctx->cfg.gjkTol;
        p.epaTol = ctx->cfg.epaTol;
        p.wantDistances = (ctx->cfg.flags & AXCD_FLAG_PAIR_DISTANCES) ? 1u : 0u;
        p.boxBoxGeneric = (ctx->cfg.flags & AXCD_FLAG_BOXBOX_GJK_EPA) ? 1u : 0u;
        // The pair count lives on the device: persistent grids sized for the capacity, capped at a
        // few resident waves.
        const uint32_t mp = ctx->cfg.maxPairs;
        uint32_t tiles = (mp + kGjkThreads - 1) / kGjkThreads;
        uint32_t closedTiles = tiles;
        if (tiles > (uint32_t)ctx->numSMs * AXCD_GJK_MIN_BLOCKS) tiles = ctx->numSMs * AXCD_GJK_MIN_BLOCKS;   // one resident wave
        if (closedTiles > (uint32_t)ctx->numSMs * AXCD_CLOSED_MIN_BLOCKS) closedTiles = ctx->numSMs * AXCD_CLOSED_MIN_BLOCKS;
        const uint32_t slotTilesMax = (mp + kSlotTile - 1) / kSlotTile;
        const uint32_t slotBlocks = slotTilesMax < (uint32_t)ctx->numSMs * 4 ? slotTilesMax : ctx->numSMs * 4;
        // the slot scan's status words were cleared by this step's Morton kernel; a step that reused the
        // cached broadphase (temporal coherence) did not run it
        if (ctx->broadSkipped)
            CU(cudaMemsetAsync(ctx->dSlotStatus, 0, sizeof(uint32_t) * ((mp + kFusedTile - 1) / kFusedTile + 129), st));
        NarrowQueues q{ctx->dEpaWork, ctx->dEpaOverflow, ctx->dEpaSpill, ctx->spillCap};
        const uint2* pairs = ctx->dPairs;
        const uint32_t* pairCount = &ctx->dCtr->pairCount;
        const uint32_t chunkCap = chunkCapFor(mp);
        // Pair classes decided in closed form (sphere-sphere, sphere-box, box-box by SAT) and the classes that
        // need GJK (hulls, capsules; box-box when forced, or when its separation distance is wanted) go to
        // separate chunk lists and kernels.  A scene that cannot produce a generic pair skips those launches.
        const bool boxGeneric = p.boxBoxGeneric || p.wantDistances;
        const uint32_t genericMask = kGenericClassMask | (boxGeneric ? (1u << 4) : 0u);
        const bool anyGeneric = ctx->hasGenericShapes || boxGeneric || ctx->slabOn || ctx->n != ctx->nOwned;
        if (!anyGeneric && !ctx->dPairDist && !ctx->splitNarrow) {
            // every pair is decided in closed form: classification, closed forms and in-order compaction in ONE kernel
            const uint32_t fusedTilesMax = (mp + kFusedTile - 1) / kFusedTile;
            const uint32_t fusedBlocks = fusedTilesMax < (uint32_t)ctx->numSMs * AXCD_FUSED_MINBLOCKS ? fusedTilesMax : (uint32_t)ctx->numSMs * AXCD_FUSED_MINBLOCKS;
            if (ctx->sinkDev && ctx->dProgress) {
                if (!ctx->capturing) {   // (a captured step is armed by axcd_step_async when the graph is launched)
                    const int rc = armContactSink(ctx);
                    if (rc) return rc;
                }
                narrowClosedFusedKernel<true><<<fusedBlocks, kFusedThreads, 0, st>>>(pairs, pairCount, mp, ctx->dType8, ctx->dXf, ctx->dShapes,
                                                                                     ctx->dContacts, ctx->cfg.maxContacts, ctx->dProgress,
                                                                                     ctx->dSlotStatus, ctx->dCtr);
                ctx->narrowReportsProgress = true;
            } else
                narrowClosedFusedKernel<false><<<fusedBlocks, kFusedThreads, 0, st>>>(pairs, pairCount, mp, ctx->dType8, ctx->dXf, ctx->dShapes,
                                                                                      ctx->dContacts, ctx->cfg.maxContacts, nullptr,
                                                                                      ctx->dSlotStatus, ctx->dCtr);
            CU(cudaGetLastError());
            recordEv(ctx, EV_GJK);
            ctx->launches[2] = 1;
        } else {
        classifyPairsKernel<<<classifyBlocksFor(mp), kClsThreads, 0, st>>>(pairs, pairCount, mp, ctx->dType8, ctx->dChunks,
                                                                            chunkCap, genericMask, ctx->dCtr);
        closedFormKernel<<<closedTiles, kGjkThreads, 0, st>>>(pairs, ctx->dChunks, chunkCap, ctx->dXf, ctx->dShapes, ctx->dFlags,
                                                        ctx->dTmpContacts, ctx->dPairDist, ctx->dCtr);
        // cylinder-capable instantiations only where a cylinder can occur (ghost bodies of a slab may be cylinders)
        const bool cyl = ctx->hasCylinders || ctx->slabOn || ctx->n != ctx->nOwned;
        if (anyGeneric) {
            if (cyl)
                gjkKernel<true><<<tiles, kGjkThreads, 0, st>>>(pairs, ctx->dChunks, chunkCap, ctx->dXf, ctx->dShapes, ctx->dHull, p,
                                                               ctx->dFlags, ctx->dTmpContacts, q, ctx->cfg.maxContacts,
                                                               ctx->dPairDist, ctx->dCtr);
            else
                gjkKernel<false><<<tiles, kGjkThreads, 0, st>>>(pairs, ctx->dChunks, chunkCap, ctx->dXf, ctx->dShapes, ctx->dHull, p,
                                                                ctx->dFlags, ctx->dTmpContacts, q, ctx->cfg.maxContacts,
                                                                ctx->dPairDist, ctx->dCtr);
        }
        slotKernel<<<slotBlocks, kSlotThreads, 0, st>>>(ctx->dFlags, pairCount, mp, ctx->dTmpContacts, ctx->dContacts,
                                                        ctx->cfg.maxContacts, ctx->dSlots, ctx->dSlotStatus, ctx->dCtr);
        CU(cudaGetLastError());
        recordEv(ctx, EV_GJK);
        // EPA: persistent grids, queue lengths are read on the device
        if (anyGeneric) {
            if (cyl) {
                epaKernel<true><<<ctx->numSMs * kEpaBlocksPerSM, kEpaThreads, kEpaSmemBytes, st>>>(
                    q, ctx->cfg.maxContacts, pairs, ctx->dXf, ctx->dShapes, ctx->dHull, p, ctx->dContacts, ctx->cfg.maxContacts,
                    ctx->dSlots, ctx->dPairDist, ctx->dCtr);
                epaWarpFallbackKernel<true><<<ctx->numSMs * kWarpFbBlocksPerSM, kWarpFbThreads, 0, st>>>(
                    q, pairs, ctx->dXf, ctx->dShapes, ctx->dHull, p, ctx->dContacts, ctx->cfg.maxContacts, ctx->dSlots,
                    ctx->dPairDist, ctx->dCtr);
            } else {
                epaKernel<false><<<ctx->numSMs * kEpaBlocksPerSM, kEpaThreads, kEpaSmemBytes, st>>>(
                    q, ctx->cfg.maxContacts, pairs, ctx->dXf, ctx->dShapes, ctx->dHull, p, ctx->dContacts, ctx->cfg.maxContacts,
                    ctx->dSlots, ctx->dPairDist, ctx->dCtr);
                epaWarpFallbackKernel<false><<<ctx->numSMs * kWarpFbBlocksPerSM, kWarpFbThreads, 0, st>>>(
                    q, pairs, ctx->dXf, ctx->dShapes, ctx->dHull, p, ctx->dContacts, ctx->cfg.maxContacts, ctx->dSlots,
                    ctx->dPairDist, ctx->dCtr);
            }
            CU(cudaGetLastError());
        }
        ctx->launches[2] = anyGeneric ? 6 : 3;   // classify, closed forms, [GJK], slots, [EPA, EPA fallback]
        if (ctx->sinkDev) {   // the fused kernel writes the sink itself; this path copies the finished array
            copyContactsKernel<<<ctx->numSMs * 4, 256, 0, st>>>(reinterpret_cast<const float2*>(ctx->dContacts),
                                                               reinterpret_cast<float2*>(ctx->sinkDev), &ctx->dCtr->contactCount,
                                                               ctx->cfg.maxContacts);
            CU(cudaGetLastError());
            ctx->launches[2] += 1;
        }
        }
    } else {
        recordEv(ctx, EV_GJK);
    }
    recordEv(ctx, EV_END);
    ctx->stage = ST_NARROW;
    return AXCD_OK;
}

int32_t axcd_get_stats(AxcdContext* ctx, AxcdStats* out) {
    if (!ctx || !out) return AXCD_ERR_NULL_POINTER;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    memset(out, 0, sizeof(*out));
    {
        const int rc = refreshCountersImpl(ctx);
        if (rc) return rc;
    }
    out->numBodies = ctx->n;
    out->movedBodies = (ctx->cfg.flags & AXCD_FLAG_TEMPORAL_COHERENCE) ? ctx->hostCtr.movedBodies : ctx->n;
    out->broadphaseSkipped = (ctx->stage >= ST_BROAD && ctx->broadSkipped) ? 1u : 0u;
    if (ctx->stage >= ST_BROAD) {
        out->numPairs = ctx->numPairs;
        out->requiredPairs = ctx->foundPairs;
    }
    if (ctx->stage >= ST_NARROW) {
        out->numContacts = ctx->numContacts;
        out->requiredContacts = ctx->foundContacts;
        out->numPenetrating = ctx->hostCtr.epaCount;
        out->gjkFailures = ctx->hostCtr.gjkFailures;
        out->epaFailures = ctx->hostCtr.epaFailures;
        out->contactPointCount = ctx->manifoldsValid ? ctx->hostCtr.manifoldPoints : 0;
    }
    out->refitMs = evMs(ctx, EV_START, EV_REFIT);
    if (ctx->stage >= ST_BROAD) {
        out->sortMs = evMs(ctx, EV_REFIT, EV_SORT);
        out->buildMs = evMs(ctx, EV_SORT, EV_BUILD);
        out->pairMs = evMs(ctx, EV_BUILD, EV_PAIR);
        out->pairSortMs = evMs(ctx, EV_PAIR, EV_PAIRSORT);
        out->broadphaseTime = evMs(ctx, EV_START, EV_BROAD_END);
    }
    if (ctx->stage >= ST_NARROW) {
        out->gjkMs = evMs(ctx, EV_N0, EV_GJK);
        out->epaMs = evMs(ctx, EV_GJK, EV_END);
        out->narrowphaseTime = evMs(ctx, EV_N0, EV_END);
        out->totalMs = evMs(ctx, EV_START, EV_END);   // a graph-launched step has this one only
    }
    out->graphLaunched = ctx->lastStepGraph ? 1u : 0u;
    out->sortFallback = (ctx->stage >= ST_BROAD) ? ctx->hostCtr.sortFallback : 0u;
    out->sortMaxBucket = (ctx->stage >= ST_BROAD) ? ctx->hostCtr.sortMaxBucket : 0u;
    // algorithmic bytes: refit 80 B/body, Morton 32 B/body, sort (16 B * passes + 4) per
    // element, pair sort (16 B * passes + 8) per pair, narrowphase gather 96 B/pair + 40 B/contact
    {
        const uint64_t n = ctx->n, np = out->numPairs, nc = out->numContacts;
        const int passes = (3 * ctx->mortonBits + ctx->worldBits + 7) / 8;
        out->bytesMoved = n * 80 + n * 32 + n * (16ull * passes + 4) + np * 40 + np * 96 + nc * 40;
    }
    out->kernelLaunches = ctx->launches[0] + (ctx->stage >= ST_BROAD ? ctx->launches[1] : 0) +
                          (ctx->stage >= ST_NARROW ? ctx->launches[2] : 0);
    if (ctx->stage >= ST_BROAD && ctx->hostCtr.travOverflow) {
        snprintf(ctx->lastErr, sizeof(ctx->lastErr), "LBVH traversal stack overflow: candidate pairs would be missing");
        return AXCD_ERR_GPU_FAILED;
    }
    if (ctx->foundPairs > ctx->cfg.maxPairs || ctx->foundContacts > ctx->cfg.maxContacts) return AXCD_ERR_OUT_OF_RANGE;
    return AXCD_OK;
}

// The fused step, asynchronous.  Replays the CUDA graph of refit + broadphase + narrowphase when one is
// current for this counter parity; captures one when the launch configuration has been stable since the
// previous step; otherwise (first step, configuration just changed, temporal coherence with its host-side
// decision, graphs switched off) calls the stage functions directly.
int32_t axcd_step_async(AxcdContext* ctx) {
    if (!ctx) return AXCD_ERR_NULL_POINTER;
    if (ctx->stage < ST_POSES) return AXCD_ERR_GPU_INVALID_OP;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    const bool graphable = !ctx->graphsOff && !(ctx->cfg.flags & (AXCD_FLAG_TEMPORAL_COHERENCE | AXCD_FLAG_NO_GRAPH)) &&
                           ctx->n >= 2;
    const bool stable = ctx->lastStepGen == ctx->gen;
    ctx->lastStepGen = ctx->gen;
    ctx->lastStepGraph = false;
    if (graphable) {
        const int parity = ctx->ctrParity ^ 1;   // the parity axcd_refit is about to switch to
        AxcdContext::StepGraph& g = ctx->graph[parity];
        if (g.exec && g.gen != ctx->gen) {
            cudaGraphExecDestroy(g.exec);
            g.exec = nullptr;
        }
        if (!g.exec && stable) {
            // record the stage functions into a graph (they skip their timing events while capturing)
            cudaGraph_t graph = nullptr;
            const int parityBefore = ctx->ctrParity;
            const int stageBefore = ctx->stage;
            if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
                ctx->capturing = true;
                int32_t rc = axcd_refit(ctx);
                if (!rc) rc = axcd_broadphase(ctx);
                if (!rc) rc = axcd_narrowphase(ctx);
                ctx->capturing = false;
                const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
                if (!rc && e == cudaSuccess && graph && cudaGraphInstantiate(&g.exec, graph, 0) == cudaSuccess) {
                    g.gen = ctx->gen;
                    g.launches[0] = ctx->launches[0];
                    g.launches[1] = ctx->launches[1];
                    g.launches[2] = ctx->launches[2];
                    g.sortedBuf = ctx->sortedBuf;
                    g.reportsProgress = ctx->narrowReportsProgress;
                } else {
                    g.exec = nullptr;
                    ctx->graphsOff = true;   // something on the path is not capturable here: stay on direct launches
                }
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
            } else {
                cudaGetLastError();
                ctx->graphsOff = true;
            }
            // nothing has run yet: rewind the host-side state the recording advanced
            ctx->ctrParity = parityBefore;
            ctx->dCtr = ctx->dCtrBase + ctx->ctrParity;
            ctx->stage = stageBefore;
        }
        if (g.exec) {
            if (g.reportsProgress) {
                const int rc = armContactSink(ctx);
                if (rc) return rc;
            }
            ctx->ctrParity = parity;
            ctx->dCtr = ctx->dCtrBase + parity;
            for (int e = 0; e < EV_COUNT; ++e) ctx->evValid[e] = false;
            recordEv(ctx, EV_START);
            CU(cudaGraphLaunch(g.exec, ctx->stream));
            recordEv(ctx, EV_END);
            ctx->launches[0] = g.launches[0];
            ctx->launches[1] = g.launches[1];
            ctx->launches[2] = g.launches[2];
            ctx->sortedBuf = g.sortedBuf;
            ctx->numPairs = ctx->foundPairs = 0;
            ctx->numContacts = ctx->foundContacts = 0;
            ctx->manifoldsValid = false;
            ctx->queryNodesValid = false;
            ctx->broadSkipped = false;
            ctx->stage = ST_NARROW;
            ctx->lastStepGraph = true;
            return AXCD_OK;
        }
    }
    int32_t rc = axcd_refit(ctx);
    if (rc) return rc;
    rc = axcd_broadphase(ctx);
    if (rc) return rc;
    return axcd_narrowphase(ctx);
}

int32_t axcd_step(AxcdContext* ctx, AxcdStats* outStats) {
    if (!ctx) return AXCD_ERR_NULL_POINTER;
    int32_t rc = axcd_step_async(ctx);
    if (rc) return rc;
    AxcdStats tmp;
    rc = axcd_get_stats(ctx, outStats ? outStats : &tmp);
    return rc;
}

int32_t axcd_get_aabbs(AxcdContext* ctx, void* outAabb24, uint32_t cap) {
    if (!ctx || (!outAabb24 && ctx->n)) return AXCD_ERR_NULL_POINTER;
    if (ctx->stage < ST_REFIT) return AXCD_ERR_GPU_INVALID_OP;
    if (cap < ctx->n) return AXCD_ERR_OUT_OF_RANGE;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    if (ctx->n) CU(cudaMemcpyAsync(outAabb24, ctx->dAabb, (size_t)ctx->n * 24, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return AXCD_OK;
}

int32_t axcd_get_pairs(AxcdContext* ctx, uint32_t* outPairs2, uint32_t cap, uint32_t* outCount) {
    if (!ctx || !outCount) return AXCD_ERR_NULL_POINTER;
    if (ctx->stage < ST_BROAD) return AXCD_ERR_GPU_INVALID_OP;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    {
        const int rc = refreshCountersImpl(ctx);
        if (rc) return rc;
    }
    *outCount = ctx->numPairs;
    if (ctx->numPairs == 0) return AXCD_OK;
    if (!outPairs2) return AXCD_ERR_NULL_POINTER;
    if (cap < ctx->numPairs) return AXCD_ERR_OUT_OF_RANGE;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    CU(cudaMemcpyAsync(outPairs2, ctx->dPairs, sizeof(uint2) * ctx->numPairs, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return AXCD_OK;
}

int32_t axcd_get_pair_distances(AxcdContext* ctx, float* outDist, uint32_t cap, uint32_t* outCount) {
    if (!ctx || !outCount) return AXCD_ERR_NULL_POINTER;
    if (ctx->stage < ST_NARROW || !ctx->dPairDist) return AXCD_ERR_GPU_INVALID_OP;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    {
        const int rc = refreshCountersImpl(ctx);
        if (rc) return rc;
    }
    *outCount = ctx->numPairs;
    if (ctx->numPairs == 0) return AXCD_OK;
    if (!outDist) return AXCD_ERR_NULL_POINTER;
    if (cap < ctx->numPairs) return AXCD_ERR_OUT_OF_RANGE;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    CU(cudaMemcpyAsync(outDist, ctx->dPairDist, sizeof(float) * ctx->numPairs, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return AXCD_OK;
}

int32_t axcd_get_contacts(AxcdContext* ctx, AxcdContact* out, uint32_t cap, uint32_t* outCount) {
    if (!ctx || !outCount) return AXCD_ERR_NULL_POINTER;
    if (ctx->stage < ST_NARROW) return AXCD_ERR_GPU_INVALID_OP;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    {
        const int rc = refreshCountersImpl(ctx);
        if (rc) return rc;
    }
    *outCount = ctx->numContacts;
    if (ctx->numContacts == 0) return AXCD_OK;
    if (!out) return AXCD_ERR_NULL_POINTER;
    if (cap < ctx->numContacts) return AXCD_ERR_OUT_OF_RANGE;
    CU(cudaMemcpyAsync(out, ctx->dContacts, sizeof(AxcdContact) * ctx->numContacts, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return AXCD_OK;
}

int32_t axcd_set_contact_sink(AxcdContext* ctx, AxcdContact* hostBuffer, uint32_t capacity) {
    if (!ctx) return AXCD_ERR_NULL_POINTER;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    {   // a step in flight still delivers into the buffer it was launched with
        const int rc = drainContactSink(ctx);
        if (rc) return rc;
    }
    if (!hostBuffer) {
        if (ctx->sinkDev) ctx->gen++;   // the launch configuration changed: graphs are re-captured
        ctx->sinkDev = ctx->sinkHost = nullptr;
        return AXCD_OK;
    }
    if (capacity < ctx->cfg.maxContacts) return AXCD_ERR_INVALID_PARAM;
    void* dev = nullptr;
    if (cudaHostGetDevicePointer(&dev, hostBuffer, 0) != cudaSuccess || !dev) {   // not page-locked / not mapped
        cudaGetLastError();
        return AXCD_ERR_INVALID_PARAM;
    }
    if (!ctx->hProgress) {   // progress words of the fused narrowphase + the stream its DMA chunks use
        const uint32_t words = (ctx->cfg.maxPairs + kFusedTile - 1) / kFusedTile + 2u;
        void* h = nullptr;
        void* d = nullptr;
        if (cudaHostAlloc(&h, sizeof(uint32_t) * words, cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess ||
            cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking) != cudaSuccess) {
            cudaGetLastError();
            if (h) cudaFreeHost(h);
            return AXCD_ERR_OUT_OF_MEMORY;
        }
        memset(h, 0, sizeof(uint32_t) * words);
        ctx->hProgress = static_cast<volatile uint32_t*>(h);
        ctx->dProgress = static_cast<uint32_t*>(d);
        ctx->progressWords = words;
    }
    if (dev != ctx->sinkDev) ctx->gen++;
    ctx->sinkDev = static_cast<AxcdContact*>(dev);
    ctx->sinkHost = hostBuffer;
    return AXCD_OK;
}

int32_t axcd_build_manifolds(AxcdContext* ctx) {
    if (!ctx) return AXCD_ERR_NULL_POINTER;
    if (ctx->stage < ST_NARROW) return AXCD_ERR_GPU_INVALID_OP;
    if (ctx->manifoldsValid) return AXCD_ERR_GPU_INVALID_OP;   // once per narrowphase, like the other stages
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    if (!ctx->dManifolds) CU(dalloc(&ctx->dManifolds, (size_t)ctx->cfg.maxContacts));
    if (ctx->n >= 2) {
        uint32_t blocks = (ctx->cfg.maxContacts + kManThreads - 1) / kManThreads;
        if (blocks > (uint32_t)ctx->numSMs * 8) blocks = ctx->numSMs * 8;   // the contact count lives on the device
        manifoldKernel<<<blocks, kManThreads, 0, ctx->stream>>>(ctx->dContacts, &ctx->dCtr->contactCount, ctx->cfg.maxContacts,
                                                                ctx->dXf, ctx->dShapes,
                                                                reinterpret_cast<float4*>(ctx->dManifolds),
                                                                &ctx->dCtr->manifoldPoints);
        CU(cudaGetLastError());
    }
    ctx->manifoldsValid = true;
    return AXCD_OK;
}

int32_t axcd_get_manifolds(AxcdContext* ctx, AxcdManifold* out, uint32_t cap, uint32_t* outCount, uint32_t* outPointCount) {
    if (!ctx || !outCount) return AXCD_ERR_NULL_POINTER;
    if (ctx->stage < ST_NARROW || !ctx->manifoldsValid) return AXCD_ERR_GPU_INVALID_OP;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    {
        const int rc = refreshCountersImpl(ctx);
        if (rc) return rc;
    }
    *outCount = ctx->numContacts;
    if (outPointCount) *outPointCount = ctx->hostCtr.manifoldPoints;
    if (ctx->numContacts == 0) return AXCD_OK;
    if (!out) return AXCD_ERR_NULL_POINTER;
    if (cap < ctx->numContacts) return AXCD_ERR_OUT_OF_RANGE;
    CU(cudaMemcpyAsync(out, ctx->dManifolds, sizeof(AxcdManifold) * ctx->numContacts, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return AXCD_OK;
}

namespace {
cudaError_t growScratch(void** p, size_t* cur, size_t need) {
    if (need <= *cur) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cur = 0;
    const cudaError_t e = cudaMalloc(p, need);
    if (e == cudaSuccess) *cur = need;
    return e;
}
// The scene queries walk 64-byte float nodes; the step only builds the compact ones, so the first query
// after a broadphase fits the float nodes from the same sorted keys and range tree (still resident).
int ensureQueryNodes(AxcdContext* ctx) {
    if (ctx->queryNodesValid || ctx->n < 2) return AXCD_OK;
    if (!ctx->dNodes) CU(dalloc(&ctx->dNodes, (size_t)ctx->cfg.maxBodies));
    buildTopologyKernel<false><<<(ctx->n + 255) / 256, 256, 0, ctx->stream>>>(ctx->dKeys[ctx->sortedBuf], ctx->n, ctx->dSegLo,
                                                                             ctx->dSegHi, ctx->segP, ctx->dNodes);
    CU(cudaGetLastError());
    ctx->queryNodesValid = true;
    return AXCD_OK;
}
QueryTree queryTreeOf(const AxcdContext* ctx) {
    QueryTree T;
    T.leafLo = ctx->dSegLo + ctx->segP;
    T.nodes = ctx->dNodes;
    T.sortedKeys = ctx->dKeys[ctx->sortedBuf];
    T.aabb = ctx->dAabb;
    T.worldId = ctx->hasWorlds ? ctx->dWorld : nullptr;
    T.n = ctx->n;
    T.worldShift = 3 * ctx->mortonBits;
    T.hasWorlds = ctx->hasWorlds ? 1 : 0;
    return T;
}
}  // namespace

int32_t axcd_query_aabbs(AxcdContext* ctx, const float* boxes6, const uint32_t* queryWorld, uint32_t nq,
                         uint32_t* outHits2, uint32_t cap, uint32_t* outCount) {
    if (!ctx || !outCount || (nq && !boxes6)) return AXCD_ERR_NULL_POINTER;
    if (ctx->stage < ST_BROAD) return AXCD_ERR_GPU_INVALID_OP;
    *outCount = 0;
    if (nq == 0 || ctx->n == 0) return AXCD_OK;
    if (nq > (1u << 28)) return AXCD_ERR_OUT_OF_RANGE;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    cudaStream_t st = ctx->stream;
    const uint32_t scanTiles = (nq + kScanTile - 1) / kScanTile;
    // layout of dQCount: counts[nqPad] | starts[nqPad] | cursor copy[nqPad] | status[scanTiles + 1] | ticket | total
    const uint32_t nqPad = (nq + 3u) & ~3u;   // the scan stores 16-byte vectors: keep every sub-buffer aligned
    const uint32_t statusWords = ((scanTiles + 1 + 128 + 3u) & ~3u);   // + the look-back's read slack, 16-byte multiple
    const size_t words = 3 * (size_t)nqPad + statusWords + 2;
    CU(growScratch(&ctx->dQIn, &ctx->qInBytes, (size_t)nq * 28));
    CU(growScratch(&ctx->dQCount, &ctx->qCountBytes, words * 4));
    float* dBoxes = static_cast<float*>(ctx->dQIn);
    uint32_t* dQW = reinterpret_cast<uint32_t*>(dBoxes + (size_t)nq * 6);
    uint32_t* dCounts = static_cast<uint32_t*>(ctx->dQCount);
    uint32_t* dStarts = dCounts + nqPad;
    uint32_t* dCursor = dStarts + nqPad;
    uint32_t* dStatus = dCursor + nqPad;
    uint32_t* dTicket = dStatus + statusWords;
    uint32_t* dTotal = dTicket + 1;
    CU(cudaMemcpyAsync(dBoxes, boxes6, (size_t)nq * 24, cudaMemcpyHostToDevice, st));
    const bool useWorld = ctx->hasWorlds && queryWorld;
    if (useWorld) CU(cudaMemcpyAsync(dQW, queryWorld, (size_t)nq * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(dStatus, 0, sizeof(uint32_t) * (statusWords + 2), st));
    {
        const int rc = ensureQueryNodes(ctx);
        if (rc) return rc;
    }
    const QueryTree T = queryTreeOf(ctx);
    const uint32_t blocks = (nq + kQueryThreads - 1) / kQueryThreads;
    queryAabbKernel<false><<<blocks, kQueryThreads, 0, st>>>(T, dBoxes, useWorld ? dQW : nullptr, nq, dCounts, nullptr, nullptr);
    exclusiveScanKernel<<<scanTiles, kScanThreads, 0, st>>>(dCounts, dStarts, dCursor, nq, dStatus, dTicket, dTotal);
    CU(cudaGetLastError());
    uint32_t total = 0;
    CU(cudaMemcpyAsync(&total, dTotal, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *outCount = total;
    if (total == 0) return AXCD_OK;
    if (total > cap) return AXCD_ERR_OUT_OF_RANGE;   // *outCount says how many slots are needed
    if (!outHits2) return AXCD_ERR_NULL_POINTER;
    CU(growScratch(&ctx->dQSeg, &ctx->qSegBytes, (size_t)total * 4));
    CU(growScratch(&ctx->dQOut, &ctx->qOutBytes, (size_t)total * 8));
    queryAabbKernel<true><<<blocks, kQueryThreads, 0, st>>>(T, dBoxes, useWorld ? dQW : nullptr, nq, dCounts, dStarts,
                                                            static_cast<uint32_t*>(ctx->dQSeg));
    sortSegmentsCoopKernel<<<(nq + kCoopBodies - 1) / kCoopBodies, kCoopBodies, 0, st>>>(dStarts, dCounts, nq, static_cast<uint32_t*>(ctx->dQSeg),
                                                                                         static_cast<uint2*>(ctx->dQOut));
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(outHits2, ctx->dQOut, (size_t)total * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return AXCD_OK;
}

int32_t axcd_raycast(AxcdContext* ctx, const AxcdRay* rays, uint32_t nq, AxcdRayHit* outHits) {
    if (!ctx || (nq && (!rays || !outHits))) return AXCD_ERR_NULL_POINTER;
    if (ctx->stage < ST_BROAD) return AXCD_ERR_GPU_INVALID_OP;
    if (nq == 0) return AXCD_OK;
    if (nq > (1u << 26)) return AXCD_ERR_OUT_OF_RANGE;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    cudaStream_t st = ctx->stream;
    CU(growScratch(&ctx->dQIn, &ctx->qInBytes, (size_t)nq * sizeof(AxcdRay)));
    CU(growScratch(&ctx->dQOut, &ctx->qOutBytes, (size_t)nq * sizeof(AxcdRayHit)));
    CU(cudaMemcpyAsync(ctx->dQIn, rays, (size_t)nq * sizeof(AxcdRay), cudaMemcpyHostToDevice, st));
    {
        const int rc = ensureQueryNodes(ctx);
        if (rc) return rc;
    }
    const QueryTree T = queryTreeOf(ctx);
    NarrowParams p;
    p.gjkMaxIters = ctx->cfg.gjkMaxIters;
    p.epaMaxIters = ctx->cfg.epaMaxIters;
    p.epaMaxFaces = ctx->cfg.epaMaxFaces;
    p.gjkTol = ctx->cfg.gjkTol;
    p.epaTol = ctx->cfg.epaTol;
    p.wantDistances = 1u;
    p.boxBoxGeneric = 0u;
    const uint32_t rb = (nq + kQueryThreads - 1) / kQueryThreads;
    if (ctx->hasSweptRayShapes)
        raycastKernel<true><<<rb, kQueryThreads, 0, st>>>(T, static_cast<const float4*>(ctx->dQIn), nq, ctx->dXf, ctx->dShapes,
                                                          ctx->dHull, p, static_cast<uint32_t*>(ctx->dQOut));
    else
        raycastKernel<false><<<rb, kQueryThreads, 0, st>>>(T, static_cast<const float4*>(ctx->dQIn), nq, ctx->dXf, ctx->dShapes,
                                                           ctx->dHull, p, static_cast<uint32_t*>(ctx->dQOut));
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(outHits, ctx->dQOut, (size_t)nq * sizeof(AxcdRayHit), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return AXCD_OK;
}

static int32_t ccdPairsImpl(AxcdContext* ctx, const uint32_t* pairs2, uint32_t npairs, const float* displacement3,
                            const float* rotation3, AxcdSweep* out) {
    if (!ctx || (npairs && (!pairs2 || !displacement3 || !out))) return AXCD_ERR_NULL_POINTER;
    if (ctx->stage < ST_POSES) return AXCD_ERR_GPU_INVALID_OP;
    if (npairs == 0) return AXCD_OK;
    if (npairs > (1u << 28)) return AXCD_ERR_OUT_OF_RANGE;
    for (uint32_t k = 0; k < 2 * npairs; ++k)
        if (pairs2[k] >= ctx->n) return AXCD_ERR_OUT_OF_RANGE;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    cudaStream_t st = ctx->stream;
    // scratch: pairs | displacements | rotation vectors in dQIn, results in dQOut
    const size_t pairBytes = ((size_t)npairs * 8 + 15) & ~(size_t)15;
    const size_t vecBytes = ((size_t)ctx->n * 12 + 15) & ~(size_t)15;
    CU(growScratch(&ctx->dQIn, &ctx->qInBytes, pairBytes + 2 * vecBytes));
    CU(growScratch(&ctx->dQOut, &ctx->qOutBytes, (size_t)npairs * sizeof(AxcdSweep)));
    uint2* dPairs = static_cast<uint2*>(ctx->dQIn);
    float* dDisp = reinterpret_cast<float*>(static_cast<char*>(ctx->dQIn) + pairBytes);
    float* dRot = reinterpret_cast<float*>(static_cast<char*>(ctx->dQIn) + pairBytes + vecBytes);
    CU(cudaMemcpyAsync(dPairs, pairs2, (size_t)npairs * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(dDisp, displacement3, (size_t)ctx->n * 12, cudaMemcpyHostToDevice, st));
    if (rotation3) CU(cudaMemcpyAsync(dRot, rotation3, (size_t)ctx->n * 12, cudaMemcpyHostToDevice, st));
    NarrowParams p;
    p.gjkMaxIters = ctx->cfg.gjkMaxIters;
    p.epaMaxIters = ctx->cfg.epaMaxIters;
    p.epaMaxFaces = ctx->cfg.epaMaxFaces;
    p.gjkTol = ctx->cfg.gjkTol;
    p.epaTol = ctx->cfg.epaTol;
    p.wantDistances = 1u;
    p.boxBoxGeneric = 0u;
    const bool cyl = ctx->hasCylinders || ctx->n != ctx->nOwned;
    const uint32_t blocks = (npairs + kCcdThreads - 1) / kCcdThreads;
    uint32_t* dOut = static_cast<uint32_t*>(ctx->dQOut);
    if (rotation3) {
        if (cyl) ccdAngularKernel<true><<<blocks, kCcdThreads, 0, st>>>(dPairs, npairs, ctx->dXf, ctx->dShapes, ctx->dHull, dDisp, dRot, p, dOut);
        else ccdAngularKernel<false><<<blocks, kCcdThreads, 0, st>>>(dPairs, npairs, ctx->dXf, ctx->dShapes, ctx->dHull, dDisp, dRot, p, dOut);
    } else {
        if (cyl) ccdKernel<true><<<blocks, kCcdThreads, 0, st>>>(dPairs, npairs, ctx->dXf, ctx->dShapes, ctx->dHull, dDisp, p, dOut);
        else ccdKernel<false><<<blocks, kCcdThreads, 0, st>>>(dPairs, npairs, ctx->dXf, ctx->dShapes, ctx->dHull, dDisp, p, dOut);
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, ctx->dQOut, (size_t)npairs * sizeof(AxcdSweep), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return AXCD_OK;
}

int32_t axcd_ccd_pairs(AxcdContext* ctx, const uint32_t* pairs2, uint32_t npairs, const float* displacement3,
                       AxcdSweep* out) {
    return ccdPairsImpl(ctx, pairs2, npairs, displacement3, nullptr, out);
}

int32_t axcd_ccd_pairs_angular(AxcdContext* ctx, const uint32_t* pairs2, uint32_t npairs, const float* displacement3,
                               const float* rotation3, AxcdSweep* out) {
    if (npairs && !rotation3) return AXCD_ERR_NULL_POINTER;
    return ccdPairsImpl(ctx, pairs2, npairs, displacement3, rotation3, out);
}

int32_t axcd_set_awake(AxcdContext* ctx, const uint8_t* awake, uint32_t n) {
    if (!ctx) return AXCD_ERR_NULL_POINTER;
    ctx->pairsCached = false;   // the candidate rule changed
    ctx->gen++;
    if (!awake) {
        ctx->awakeOn = false;
        return AXCD_OK;
    }
    if (n > ctx->cfg.maxBodies) return AXCD_ERR_OUT_OF_RANGE;
    if (n != ctx->nOwned) return AXCD_ERR_INVALID_PARAM;   // one flag per owned body (ghosts count as awake)
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    if (!ctx->dAwake) {
        CU(dalloc(&ctx->dAwake, (size_t)ctx->cfg.maxBodies));
        CU(cudaMemsetAsync(ctx->dAwake, 1, ctx->cfg.maxBodies, ctx->stream));   // bodies beyond n (ghosts) are awake
    }
    if (n) CU(cudaMemcpyAsync(ctx->dAwake, awake, n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->awakeOn = true;
    return AXCD_OK;
}

int32_t axcd_set_filters(AxcdContext* ctx, const AxcdFilter* filters, uint32_t n) {
    if (!ctx) return AXCD_ERR_NULL_POINTER;
    ctx->gen++;
    if (!filters) {
        ctx->filtersOn = false;
        ctx->pairsCached = false;
        return AXCD_OK;
    }
    // one record per body held (the traversal reads the filter of every body it pairs); ghost bodies of the
    // slab mode carry no filter words, so filtering and slabs do not combine in this version
    if (ctx->slabOn || ctx->n != ctx->nOwned) return AXCD_ERR_INVALID_PARAM;
    if (ctx->stage < ST_SHAPES || n != ctx->n) return AXCD_ERR_INVALID_PARAM;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    if (!ctx->dFilters) CU(dalloc(&ctx->dFilters, (size_t)ctx->cfg.maxBodies));
    uint4* tmp = static_cast<uint4*>(malloc(sizeof(uint4) * (n ? n : 1)));
    if (!tmp) return AXCD_ERR_OUT_OF_MEMORY;
    for (uint32_t i = 0; i < n; ++i)
        tmp[i] = make_uint4(filters[i].categoryBits, filters[i].maskBits, (uint32_t)(int32_t)filters[i].groupIndex, 0u);
    cudaError_t e = cudaMemcpyAsync(ctx->dFilters, tmp, sizeof(uint4) * n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    free(tmp);
    if (e != cudaSuccess) return fail(ctx, e, "filter upload");
    ctx->filtersOn = true;
    ctx->pairsCached = false;
    return AXCD_OK;
}

int32_t axcd_set_slab(AxcdContext* ctx, float xLo, float xHi, uint32_t enable) {
    if (!ctx) return AXCD_ERR_NULL_POINTER;
    if (enable && !(xLo <= xHi)) return AXCD_ERR_INVALID_PARAM;
    if (enable && ctx->cfg.numWorlds > 1) return AXCD_ERR_INVALID_PARAM;   // slabs split ONE scene
    // the ghost set changes every step, so there is no persistent fat-box state to be coherent with
    if (enable && (ctx->cfg.flags & AXCD_FLAG_TEMPORAL_COHERENCE)) return AXCD_ERR_INVALID_PARAM;
    if (enable && ctx->filtersOn) return AXCD_ERR_INVALID_PARAM;   // ghost records carry no filter words
    if (ctx->slabOn != (enable != 0) || ctx->slabLo != xLo || ctx->slabHi != xHi) ctx->gen++;
    ctx->slabOn = enable != 0;
    ctx->slabLo = xLo;
    ctx->slabHi = xHi;
    return AXCD_OK;
}

int32_t axcd_set_body_keys(AxcdContext* ctx, const uint32_t* keys, uint32_t first, uint32_t count) {
    if (!ctx || (count && !keys)) return AXCD_ERR_NULL_POINTER;
    if ((uint64_t)first + count > ctx->cfg.maxBodies) return AXCD_ERR_OUT_OF_RANGE;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    if (count) CU(cudaMemcpyAsync(ctx->dBodyKeys + first, keys, sizeof(uint32_t) * count, cudaMemcpyHostToDevice, ctx->stream));
    return AXCD_OK;
}

int32_t axcd_get_body_keys(AxcdContext* ctx, uint32_t* outKeys, uint32_t cap) {
    if (!ctx || (!outKeys && ctx->n)) return AXCD_ERR_NULL_POINTER;
    if (cap < ctx->n) return AXCD_ERR_OUT_OF_RANGE;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    if (ctx->n) CU(cudaMemcpyAsync(outKeys, ctx->dBodyKeys, sizeof(uint32_t) * ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return AXCD_OK;
}

int32_t axcd_set_ghosts(AxcdContext* ctx, uint32_t nOwned, uint32_t nGhosts, const void* transforms40,
                        const AxcdShape* shapes, const uint32_t* keys) {
    if (!ctx) return AXCD_ERR_NULL_POINTER;
    if (ctx->stage < ST_POSES) return AXCD_ERR_GPU_INVALID_OP;   // owned bodies first
    if (nGhosts && (!transforms40 || !shapes || !keys)) return AXCD_ERR_NULL_POINTER;
    if (nOwned != ctx->nOwned) return AXCD_ERR_INVALID_PARAM;
    if ((uint64_t)nOwned + nGhosts > ctx->cfg.maxBodies) return AXCD_ERR_OUT_OF_RANGE;
    ctx->fatValid = false;
    ctx->pairsCached = false;
    for (uint32_t i = 0; i < nGhosts; ++i)
        if (shapes[i].type != AXCD_SHAPE_SPHERE && shapes[i].type != AXCD_SHAPE_BOX &&
            shapes[i].type != AXCD_SHAPE_CAPSULE && shapes[i].type != AXCD_SHAPE_CYLINDER)
            return AXCD_ERR_INVALID_SHAPE;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    cudaStream_t st = ctx->stream;
    if (nGhosts) {
        CU(cudaMemcpyAsync(ctx->dXf + (size_t)nOwned * 10, transforms40, (size_t)nGhosts * 40, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->dShapes + nOwned, shapes, sizeof(AxcdShape) * nGhosts, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->dBodyKeys + nOwned, keys, sizeof(uint32_t) * nGhosts, cudaMemcpyHostToDevice, st));
    }
    if (ctx->n != nOwned + nGhosts) {
        const int rc = resizeBodies(ctx, nOwned + nGhosts);
        if (rc) return rc;
    }
    ctx->stage = ST_POSES;
    return AXCD_OK;
}

int32_t axcd_pack_ghosts(AxcdContext* ctx, const float* edges, uint32_t numRanks, uint32_t myRank, void** outDevPtrs,
                         uint32_t* outCounts) {
    if (!ctx || !edges || !outDevPtrs || !outCounts) return AXCD_ERR_NULL_POINTER;
    if (ctx->stage < ST_REFIT) return AXCD_ERR_GPU_INVALID_OP;
    if (numRanks == 0 || numRanks > 64 || myRank >= numRanks) return AXCD_ERR_INVALID_PARAM;
    if (ctx->nHull) return AXCD_ERR_INVALID_SHAPE;   // hull ghosts would need their vertices shipped too
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    cudaStream_t st = ctx->stream;
    const uint32_t cap = ctx->cfg.maxBodies - ctx->nOwned;
    if (ctx->ghostRanks != numRanks || ctx->ghostCap != cap || !ctx->dGhostSend) {
        if (ctx->dGhostSend) cudaFree(ctx->dGhostSend);
        if (ctx->dGhostCount) cudaFree(ctx->dGhostCount);
        ctx->dGhostSend = nullptr;
        ctx->dGhostCount = nullptr;
        CU(dalloc(&ctx->dGhostSend, (size_t)numRanks * cap * (kGhostWords / 4)));
        CU(dalloc(&ctx->dGhostCount, (size_t)numRanks));
        ctx->ghostRanks = numRanks;
        ctx->ghostCap = cap;
    }
    CU(cudaMemsetAsync(ctx->dGhostCount, 0, sizeof(uint32_t) * numRanks, st));
    const uint32_t blocks = (ctx->nOwned + 255) / 256;
    for (uint32_t r = 0; r < numRanks; ++r) {
        outDevPtrs[r] = nullptr;
        outCounts[r] = 0;
        if (r == myRank || ctx->nOwned == 0) continue;
        float4* buf = ctx->dGhostSend + (size_t)r * cap * (kGhostWords / 4);
        packGhostsKernel<<<blocks, 256, 0, st>>>(ctx->dAabb, ctx->dXf, ctx->dShapes, ctx->dBodyKeys, ctx->nOwned, edges[r],
                                                 edges[r + 1], buf, cap, ctx->dGhostCount + r);
        outDevPtrs[r] = buf;
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(outCounts, ctx->dGhostCount, sizeof(uint32_t) * numRanks, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    bool over = false;
    for (uint32_t r = 0; r < numRanks; ++r)
        if (outCounts[r] > cap) over = true;
    return over ? AXCD_ERR_OUT_OF_RANGE : AXCD_OK;
}

int32_t axcd_set_ghosts_device(AxcdContext* ctx, uint32_t nOwned, uint32_t nGhosts, const void* devRecords) {
    if (!ctx) return AXCD_ERR_NULL_POINTER;
    if (ctx->stage < ST_POSES) return AXCD_ERR_GPU_INVALID_OP;
    if (nGhosts && !devRecords) return AXCD_ERR_NULL_POINTER;
    if (nOwned != ctx->nOwned) return AXCD_ERR_INVALID_PARAM;
    if ((uint64_t)nOwned + nGhosts > ctx->cfg.maxBodies) return AXCD_ERR_OUT_OF_RANGE;
    ctx->fatValid = false;
    ctx->pairsCached = false;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    if (nGhosts) {
        GhostOffsets none;
        none.numRanks = 0;
        unpackGhostsKernel<<<(nGhosts + 255) / 256, 256, 0, ctx->stream>>>(static_cast<const float4*>(devRecords), nGhosts, nOwned,
                                                                            ctx->dXf, ctx->dShapes, ctx->dBodyKeys, none, 0u);
        CU(cudaGetLastError());
    }
    if (ctx->n != nOwned + nGhosts) {
        const int rc = resizeBodies(ctx, nOwned + nGhosts);
        if (rc) return rc;
    }
    ctx->stage = ST_POSES;
    return AXCD_OK;
}

// ---- x-slab mode with the ghost exchange inside the library --------------------------------
// NCCL is bound at run time (dlopen of libnccl.so.2), so libaxcd.so loads and every single-GPU entry point
// works on a machine without NCCL; a process that already loaded NCCL (torch) shares that copy.
}  // extern "C"
namespace {
typedef struct ncclComm* NcclComm;
struct NcclUniqueId { char internal[128]; };
enum { kNcclUint32 = 3, kNcclFloat32 = 7 };   // ncclDataType_t values (nccl.h)
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi* ncclApi() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return &api;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) return &api;
    auto sym = [&](const char* s) { return dlsym(api.lib, s); };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.Send && api.Recv &&
             api.GroupStart && api.GroupEnd;
    return &api;
}
int ncclFail(AxcdContext* c, int rc, const char* what) {
    NcclApi* N = ncclApi();
    if (c) snprintf(c->lastErr, sizeof(c->lastErr), "%s: NCCL error %d (%s)", what, rc,
                    (N->GetErrorString ? N->GetErrorString(rc) : "?"));
    return AXCD_ERR_GPU_FAILED;
}
#define NC(call)                                              \
    do {                                                      \
        const int _r = (call);                                \
        if (_r != 0) return ncclFail(ctx, _r, #call);         \
    } while (0)

void slabRelease(AxcdContext* ctx) {
    if (ctx->slabComm && ctx->slabOwnsComm && ncclApi()->ok) ncclApi()->CommDestroy(static_cast<NcclComm>(ctx->slabComm));
    ctx->slabComm = nullptr;
    if (ctx->dEdges) cudaFree(ctx->dEdges);
    if (ctx->dCountMatrix) cudaFree(ctx->dCountMatrix);
    if (ctx->dGhostRecv) cudaFree(ctx->dGhostRecv);
    if (ctx->dGhostVertSend) cudaFree(ctx->dGhostVertSend);
    ctx->dGhostVertSend = nullptr;
    if (ctx->hCountMatrix) cudaFreeHost(ctx->hCountMatrix);
    if (ctx->evX0) cudaEventDestroy(ctx->evX0);
    if (ctx->evX1) cudaEventDestroy(ctx->evX1);
    ctx->dEdges = nullptr;
    ctx->dCountMatrix = nullptr;
    ctx->dGhostRecv = nullptr;
    ctx->hCountMatrix = nullptr;
    ctx->evX0 = ctx->evX1 = nullptr;
    cudaGetLastError();
}

int slabSetup(AxcdContext* ctx, void* comm, bool owns, uint32_t rank, uint32_t numRanks, const float* edges) {
    if (ctx->stage < ST_POSES) return AXCD_ERR_GPU_INVALID_OP;   // owned shapes and poses first
    if (ctx->cfg.numWorlds > 1 || (ctx->cfg.flags & AXCD_FLAG_TEMPORAL_COHERENCE)) return AXCD_ERR_INVALID_PARAM;
    for (uint32_t r = 0; r < numRanks; ++r)
        if (!(edges[r] <= edges[r + 1])) return AXCD_ERR_INVALID_PARAM;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    slabRelease(ctx);
    ctx->slabComm = comm;
    ctx->slabOwnsComm = owns;
    ctx->slabRank = rank;
    ctx->slabRanks = numRanks;
    const uint32_t cap = ctx->cfg.maxBodies - ctx->nOwned;
    if (ctx->dGhostSend) cudaFree(ctx->dGhostSend);
    if (ctx->dGhostCount) cudaFree(ctx->dGhostCount);
    ctx->dGhostSend = nullptr;
    ctx->dGhostCount = nullptr;
    CU(dalloc(&ctx->dGhostSend, (size_t)numRanks * cap * (kGhostWords / 4)));
    CU(dalloc(&ctx->dGhostCount, (size_t)2 * numRanks));   // records, then hull vertices, per destination
    ctx->ghostRanks = numRanks;
    ctx->ghostCap = cap;
    // hull ghosts bring their vertices: they land behind the owned ones in the hull pool
    ctx->ghostVertCap = ctx->cfg.maxHullVerts > ctx->nHull ? ctx->cfg.maxHullVerts - ctx->nHull : 0u;
    if (ctx->ghostVertCap) CU(dalloc(&ctx->dGhostVertSend, (size_t)numRanks * ctx->ghostVertCap));
    CU(dalloc(&ctx->dGhostRecv, (size_t)cap * (kGhostWords / 4)));
    CU(dalloc(&ctx->dEdges, (size_t)numRanks + 1));
    CU(dalloc(&ctx->dCountMatrix, (size_t)2 * numRanks * numRanks));
    CU(cudaMallocHost(reinterpret_cast<void**>(&ctx->hCountMatrix), sizeof(uint32_t) * 2 * numRanks * numRanks));
    CU(cudaEventCreate(&ctx->evX0));
    CU(cudaEventCreate(&ctx->evX1));
    CU(cudaMemcpyAsync(ctx->dEdges, edges, sizeof(float) * (numRanks + 1), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->hasSweptRayShapes = true;   // ghost bodies may be hulls or cylinders
    return axcd_set_slab(ctx, edges[rank], edges[rank + 1], 1u);
}
}  // namespace
extern "C" {

int32_t axcd_nccl_unique_id(void* out128) {
    if (!out128) return AXCD_ERR_NULL_POINTER;
    NcclApi* N = ncclApi();
    if (!N->ok) return AXCD_ERR_GPU_INIT;
    NcclUniqueId id;
    if (N->GetUniqueId(&id) != 0) return AXCD_ERR_GPU_FAILED;
    memcpy(out128, &id, sizeof(id));
    return AXCD_OK;
}

int32_t axcd_slab_init(AxcdContext* ctx, const void* uniqueId128, uint32_t rank, uint32_t numRanks, const float* edges) {
    if (!ctx || !uniqueId128 || !edges) return AXCD_ERR_NULL_POINTER;
    if (numRanks == 0 || numRanks > 64 || rank >= numRanks) return AXCD_ERR_INVALID_PARAM;
    NcclApi* N = ncclApi();
    if (!N->ok) {
        snprintf(ctx->lastErr, sizeof(ctx->lastErr), "libnccl.so.2 could not be loaded");
        return AXCD_ERR_GPU_INIT;
    }
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    NcclUniqueId id;
    memcpy(&id, uniqueId128, sizeof(id));
    NcclComm comm = nullptr;
    NC(N->CommInitRank(&comm, (int)numRanks, id, (int)rank));
    const int rc = slabSetup(ctx, comm, true, rank, numRanks, edges);
    if (rc != AXCD_OK && ctx->slabComm != comm) N->CommDestroy(comm);
    return rc;
}

int32_t axcd_slab_init_comm(AxcdContext* ctx, void* ncclComm, uint32_t rank, uint32_t numRanks, const float* edges) {
    if (!ctx || !ncclComm || !edges) return AXCD_ERR_NULL_POINTER;
    if (numRanks == 0 || numRanks > 64 || rank >= numRanks) return AXCD_ERR_INVALID_PARAM;
    if (!ncclApi()->ok) return AXCD_ERR_GPU_INIT;
    return slabSetup(ctx, ncclComm, false, rank, numRanks, edges);
}

// One slab step, asynchronous after the size handshake: refit of the owned bodies -> device-side ghost
// selection for every other rank (one kernel) -> ncclAllGather of the per-destination counts -> one host
// read of the size matrix (the only host synchronisation of the step; NCCL receive counts are host
// arguments) -> grouped ncclSend / ncclRecv of the 64-byte ghost records, GPU to GPU over NVLink -> ghosts
// appended behind the owned bodies -> the fused step (CUDA graph once the ghost count is stable).
int32_t axcd_slab_step_async(AxcdContext* ctx) {
    if (!ctx) return AXCD_ERR_NULL_POINTER;
    if (!ctx->slabComm || ctx->stage < ST_POSES) return AXCD_ERR_GPU_INVALID_OP;
    NcclApi* N = ncclApi();
    NcclComm comm = static_cast<NcclComm>(ctx->slabComm);
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    cudaStream_t st = ctx->stream;
    const uint32_t R = ctx->slabRanks, me = ctx->slabRank, cap = ctx->ghostCap, nOwned = ctx->nOwned;
    CU(cudaEventRecord(ctx->evX0, st));
    {
        const int rc = axcd_refit(ctx);   // boxes of the owned bodies (stale ghosts behind them are refit too; harmless)
        if (rc) return rc;
    }
    const uint32_t vcap = ctx->ghostVertCap;
    CU(cudaMemsetAsync(ctx->dGhostCount, 0, sizeof(uint32_t) * 2 * R, st));
    if (nOwned && R > 1) {
        packGhostsAllKernel<<<(nOwned + 255) / 256, 256, 0, st>>>(ctx->dAabb, ctx->dXf, ctx->dShapes, ctx->dBodyKeys, ctx->dHull, nOwned,
                                                                 ctx->dEdges, R, me, ctx->dGhostSend, cap, ctx->dGhostVertSend, vcap,
                                                                 ctx->dGhostCount);
        CU(cudaGetLastError());
    }
    // size handshake: everybody learns the whole count matrix (row r = what rank r sends: R record counts, R vertex counts)
    NC(N->AllGather(ctx->dGhostCount, ctx->dCountMatrix, 2 * R, kNcclUint32, comm, st));
    CU(cudaMemcpyAsync(ctx->hCountMatrix, ctx->dCountMatrix, sizeof(uint32_t) * 2 * R * R, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const uint32_t* M = ctx->hCountMatrix;
    auto recs = [&](uint32_t src, uint32_t dst) { return M[src * 2 * R + dst]; };
    auto verts = [&](uint32_t src, uint32_t dst) { return M[src * 2 * R + R + dst]; };
    uint64_t total = 0, totalVerts = 0;
    for (uint32_t r = 0; r < R; ++r) {
        if (r != me) {
            total += recs(r, me);
            totalVerts += verts(r, me);
        }
    }
    // a send buffer that overflowed anywhere fails the step on every rank (the matrix is the same everywhere)
    for (uint32_t r = 0; r < R; ++r)
        for (uint32_t d = 0; d < R; ++d)
            if (recs(r, d) > ctx->cfg.maxBodies || verts(r, d) > (1u << 30)) return AXCD_ERR_OUT_OF_RANGE;
    for (uint32_t d = 0; d < R; ++d)
        if (d != me && (recs(me, d) > cap || verts(me, d) > vcap)) return AXCD_ERR_OUT_OF_RANGE;
    if (total > cap || totalVerts > vcap) return AXCD_ERR_OUT_OF_RANGE;
    // the records (and the hull vertices, straight into the hull pool behind the owned ones), rank to rank
    GhostOffsets go;
    go.numRanks = R;
    NC(N->GroupStart());
    uint64_t off = 0, voff = 0;
    for (uint32_t r = 0; r < R; ++r) {
        go.vertBase[r] = ctx->nHull + (uint32_t)voff;
        if (r != me) {
            const uint32_t nSend = recs(me, r), nRecv = recs(r, me), vSend = verts(me, r), vRecv = verts(r, me);
            if (nSend)
                NC(N->Send(ctx->dGhostSend + (size_t)r * cap * (kGhostWords / 4), (size_t)nSend * kGhostWords, kNcclFloat32, (int)r, comm, st));
            if (nRecv)
                NC(N->Recv(ctx->dGhostRecv + off * (kGhostWords / 4), (size_t)nRecv * kGhostWords, kNcclFloat32, (int)r, comm, st));
            if (vSend) NC(N->Send(ctx->dGhostVertSend + (size_t)r * vcap, (size_t)vSend * 4, kNcclFloat32, (int)r, comm, st));
            if (vRecv) NC(N->Recv(ctx->dHull + ctx->nHull + voff, (size_t)vRecv * 4, kNcclFloat32, (int)r, comm, st));
            off += nRecv;
            voff += vRecv;
        }
        go.recEnd[r] = (uint32_t)off;
    }
    NC(N->GroupEnd());
    if ((uint64_t)nOwned + total > ctx->cfg.maxBodies) return AXCD_ERR_OUT_OF_RANGE;
    ctx->fatValid = false;
    ctx->pairsCached = false;
    if (total) {
        unpackGhostsKernel<<<((uint32_t)total + 255) / 256, 256, 0, st>>>(ctx->dGhostRecv, (uint32_t)total, nOwned, ctx->dXf, ctx->dShapes,
                                                                         ctx->dBodyKeys, go, 1u);
        CU(cudaGetLastError());
    }
    if (ctx->n != nOwned + (uint32_t)total) {
        const int rc = resizeBodies(ctx, nOwned + (uint32_t)total);
        if (rc) return rc;
    }
    ctx->stage = ST_POSES;
    CU(cudaEventRecord(ctx->evX1, st));
    ctx->lastGhosts = (uint32_t)total;
    return axcd_step_async(ctx);
}

int32_t axcd_slab_step(AxcdContext* ctx, AxcdStats* outStats) {
    int32_t rc = axcd_slab_step_async(ctx);
    if (rc) return rc;
    AxcdStats tmp;
    AxcdStats* out = outStats ? outStats : &tmp;
    rc = axcd_get_stats(ctx, out);
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, ctx->evX0, ctx->evX1) == cudaSuccess) out->exchangeMs = ms;
    cudaGetLastError();
    out->ghostBodies = ctx->lastGhosts;
    return rc;
}

// ---- test hooks ----------------------------------------------------------------------------------
int32_t axcd_test_sort_pairs32(AxcdContext* ctx, uint32_t* keys, uint32_t* vals, uint32_t n, uint32_t keyBits) {
    if (!ctx || (n && (!keys || !vals))) return AXCD_ERR_NULL_POINTER;
    if (n > ctx->cfg.maxBodies || keyBits == 0 || keyBits > 32) return AXCD_ERR_OUT_OF_RANGE;
    if (n == 0) return AXCD_OK;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    cudaStream_t st = ctx->stream;
    CU(cudaMemcpyAsync(ctx->dKeys[0], keys, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->dVals[0], vals, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
    const int sb = radixSort<uint32_t, true>(ctx->dKeys[0], ctx->dKeys[1], ctx->dVals[0], ctx->dVals[1], n, 0,
                                             (int)(keyBits + 7) / 8, ctx->dSortHist, ctx->dSortStatus,
                                             ctx->dCtr->sortTicket, st);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(keys, ctx->dKeys[sb], sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(vals, ctx->dVals[sb], sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return AXCD_OK;
}

int32_t axcd_test_sort_keys64(AxcdContext* ctx, uint64_t* keys, uint32_t n, uint32_t keyBits) {
    if (!ctx || (n && !keys)) return AXCD_ERR_NULL_POINTER;
    if (n > ctx->cfg.maxPairs || keyBits == 0 || keyBits > 64) return AXCD_ERR_OUT_OF_RANGE;
    if (n == 0) return AXCD_OK;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    cudaStream_t st = ctx->stream;
    uint64_t* kb[2] = {reinterpret_cast<uint64_t*>(ctx->dPairsTmp), reinterpret_cast<uint64_t*>(ctx->dPairs)};
    CU(cudaMemcpyAsync(kb[0], keys, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, st));
    const int sb = radixSort<uint64_t, false>(kb[0], kb[1], nullptr, nullptr, n, 0,
                                              (int)(keyBits + 7) / 8, ctx->dSortHist, ctx->dSortStatus,
                                              ctx->dCtr->sortTicket, st);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(keys, kb[sb], sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return AXCD_OK;
}

int32_t axcd_test_sort_bench(AxcdContext* ctx, uint32_t n, uint32_t keyBits, uint32_t iters, float* outMsPerSort) {
    if (!ctx || !outMsPerSort) return AXCD_ERR_NULL_POINTER;
    if (n == 0 || n > ctx->cfg.maxBodies || keyBits == 0 || keyBits > 32 || iters == 0) return AXCD_ERR_OUT_OF_RANGE;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    cudaStream_t st = ctx->stream;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    float total = 0.0f;
    const int passes = (int)(keyBits + 7) / 8;
    for (uint32_t it = 0; it <= iters; ++it) {   // iteration 0 is a warm-up
        fillRandomKeysKernel<<<ctx->numSMs * 8, 256, 0, st>>>(ctx->dKeys[0], ctx->dVals[0], n, keyBits, 0x9E3779B9u * (it + 1));
        CU(cudaEventRecord(e0, st));
        radixSort<uint32_t, true>(ctx->dKeys[0], ctx->dKeys[1], ctx->dVals[0], ctx->dVals[1], n, 0, passes,
                                  ctx->dSortHist, ctx->dSortStatus, ctx->dCtr->sortTicket, st);
        CU(cudaEventRecord(e1, st));
        CU(cudaStreamSynchronize(st));
        float ms = 0.0f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (it) total += ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *outMsPerSort = total / (float)iters;
    return AXCD_OK;
}

// Sorts n keys with an identity payload the way a step sorts its Morton keys (mode 1: bucket sort with the LSD
// fallback armed, mode 0: LSD only); outputs sorted keys and the permutation, *outFallback = 1 if the LSD
// kernels did the work.
int32_t axcd_test_sort_morton(AxcdContext* ctx, const uint32_t* keys, uint32_t n, uint32_t keyBits, uint32_t mode,
                              uint32_t* outKeys, uint32_t* outVals, uint32_t* outFallback) {
    if (!ctx || (n && (!keys || !outKeys || !outVals))) return AXCD_ERR_NULL_POINTER;
    if (n > ctx->cfg.maxBodies || keyBits == 0 || keyBits > 32) return AXCD_ERR_OUT_OF_RANGE;
    if (n == 0) return AXCD_OK;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    cudaStream_t st = ctx->stream;
    const int passes = (int)(keyBits + 7) / 8;
    const BucketPlan bp = mode ? bucketPlanFor(n, (int)keyBits) : BucketPlan{0, 0};
    CU(cudaMemcpyAsync(ctx->dKeys[0], keys, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
    uint32_t* flag = &ctx->dCtr->sortFallback;
    CU(cudaMemsetAsync(flag, 0, 4, st));
    iotaCountKernel<<<ctx->numSMs * 8, 256, 0, st>>>(ctx->dKeys[0], ctx->dVals[0], n, bp.bucketBits ? ctx->dBucketCounts : nullptr, bp.shift);
    const uint32_t* enable = nullptr;
    if (bp.bucketBits) {
        const uint32_t nbk = 1u << bp.bucketBits;
        const int outBuf = passes & 1;
        bucketScanKernel<<<1, 1024, 0, st>>>(ctx->dBucketCounts, ctx->dBucketStarts, ctx->dBucketCursors, nbk, flag, &ctx->dCtr->sortMaxBucket);
        bucketScatterKernel<<<ctx->numSMs * 8, 256, 0, st>>>(ctx->dKeys[0], n, bp.shift, ctx->dBucketCursors, ctx->dBucketTmp, flag);
        const uint32_t sortBlocks = nbk < (uint32_t)ctx->numSMs * 16 ? nbk : (uint32_t)ctx->numSMs * 16;
        bucketSortKernel<<<sortBlocks, kBucketThreads, 0, st>>>(ctx->dBucketTmp, ctx->dBucketStarts, nbk, ctx->dKeys[outBuf], ctx->dVals[outBuf], flag);
        enable = flag;
    }
    const int sb = radixSort<uint32_t, true>(ctx->dKeys[0], ctx->dKeys[1], ctx->dVals[0], ctx->dVals[1], n, 0, passes,
                                             ctx->dSortHist, ctx->dSortStatus, ctx->dCtr->sortTicket, st, ctx->numSMs, false, enable);
    CU(cudaGetLastError());
    uint32_t fb = 0;
    CU(cudaMemcpyAsync(&fb, flag, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(outKeys, ctx->dKeys[sb], sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(outVals, ctx->dVals[sb], sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (outFallback) *outFallback = bp.bucketBits ? fb : 1u;
    return AXCD_OK;
}

// Device-resident timing of that sort on n pseudo-random keys (identity payload): average ms of `iters` sorts,
// key generation + bucket counting excluded for mode 0 and included (it is part of the method) for mode 1.
int32_t axcd_test_sort_bench_morton(AxcdContext* ctx, uint32_t n, uint32_t keyBits, uint32_t iters, uint32_t mode,
                                    float* outMsPerSort) {
    if (!ctx || !outMsPerSort) return AXCD_ERR_NULL_POINTER;
    if (n == 0 || n > ctx->cfg.maxBodies || keyBits == 0 || keyBits > 32 || iters == 0) return AXCD_ERR_OUT_OF_RANGE;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    cudaStream_t st = ctx->stream;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    float total = 0.0f;
    const int passes = (int)(keyBits + 7) / 8;
    const BucketPlan bp = mode ? bucketPlanFor(n, (int)keyBits) : BucketPlan{0, 0};
    uint32_t* flag = &ctx->dCtr->sortFallback;
    for (uint32_t it = 0; it <= iters; ++it) {   // iteration 0 is a warm-up
        fillRandomKeysKernel<<<ctx->numSMs * 8, 256, 0, st>>>(ctx->dKeys[0], ctx->dVals[0], n, keyBits, 0x9E3779B9u * (it + 1));
        CU(cudaMemsetAsync(flag, 0, 4, st));
        CU(cudaEventRecord(e0, st));
        const uint32_t* enable = nullptr;
        if (bp.bucketBits) {
            const uint32_t nbk = 1u << bp.bucketBits;
            const int outBuf = passes & 1;
            iotaCountKernel<<<ctx->numSMs * 8, 256, 0, st>>>(ctx->dKeys[0], ctx->dVals[0], n, ctx->dBucketCounts, bp.shift);
            bucketScanKernel<<<1, 1024, 0, st>>>(ctx->dBucketCounts, ctx->dBucketStarts, ctx->dBucketCursors, nbk, flag, &ctx->dCtr->sortMaxBucket);
            bucketScatterKernel<<<ctx->numSMs * 8, 256, 0, st>>>(ctx->dKeys[0], n, bp.shift, ctx->dBucketCursors, ctx->dBucketTmp, flag);
            const uint32_t sortBlocks = nbk < (uint32_t)ctx->numSMs * 16 ? nbk : (uint32_t)ctx->numSMs * 16;
            bucketSortKernel<<<sortBlocks, kBucketThreads, 0, st>>>(ctx->dBucketTmp, ctx->dBucketStarts, nbk, ctx->dKeys[outBuf], ctx->dVals[outBuf], flag);
            enable = flag;
        }
        radixSort<uint32_t, true>(ctx->dKeys[0], ctx->dKeys[1], ctx->dVals[0], ctx->dVals[1], n, 0, passes, ctx->dSortHist,
                                  ctx->dSortStatus, ctx->dCtr->sortTicket, st, ctx->numSMs, false, enable);
        CU(cudaEventRecord(e1, st));
        CU(cudaStreamSynchronize(st));
        float ms = 0.0f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (it) total += ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *outMsPerSort = total / (float)iters;
    return AXCD_OK;
}

// FP32 FMA-chain peak of the device (the roof the GJK / EPA kernels are measured against): every thread
// runs 8 independent chains of explicit fmaf (the library is built -fmad=false, explicit calls still fuse),
// 2 flops per FMA, timed with CUDA events on the context stream.
namespace {
__global__ void __launch_bounds__(256) fmaChainKernel(float* __restrict__ out, uint32_t iters, float a, float b) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
          x7 = x0 + 7.f;
    for (uint32_t i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456f) out[0] = s;   // keeps the chains alive; never true in practice
}
}  // namespace

int32_t axcd_test_fp32_peak(AxcdContext* ctx, uint32_t iters, float* outTflops) {
    if (!ctx || !outTflops) return AXCD_ERR_NULL_POINTER;
    if (iters == 0 || iters > (1u << 20)) return AXCD_ERR_OUT_OF_RANGE;
    cudaSetDevice(ctx->cfg.deviceOrdinal);
    cudaStream_t st = ctx->stream;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    const int blocks = ctx->numSMs * 8, threads = 256;   // 2048 resident threads per SM
    float best = 0.0f;
    for (int rep = 0; rep < 4; ++rep) {   // rep 0 warms up
        CU(cudaEventRecord(e0, st));
        fmaChainKernel<<<blocks, threads, 0, st>>>(reinterpret_cast<float*>(ctx->dSortHist), iters, 0.999f, 1e-3f);
        CU(cudaEventRecord(e1, st));
        CU(cudaStreamSynchronize(st));
        float ms = 0.0f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 8.0 * 16.0 * (double)iters * (double)blocks * threads;
        const float tf = (float)(flops / (ms * 1e-3) / 1e12);
        if (rep && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *outTflops = best;
    return AXCD_OK;
}

}  // extern "C"
