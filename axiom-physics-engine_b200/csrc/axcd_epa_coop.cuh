// Cooperative EPA: 8 lanes expand one polytope together (4 pairs per warp).  EXPERIMENTAL, off by
// default (AXCD_FLAG_EPA_COOPERATIVE): bit-identical to the per-thread kernel but 3x slower on the
// headline scene (profiles/r01_experiments.md) — the unrolled per-slot loops issue for every lane
// of the warp whether or not its slot takes part.
//
// The one-thread-per-pair EPA keeps a 0.9 KB polytope per thread in shared memory, which caps an SM
// at 8 warps and leaves it latency-bound.  Here the faces live in REGISTERS, striped over the 8
// lanes of a group (slot f belongs to lane f & 7, register f >> 3; 32 face slots), and only the
// vertices, the two cores and a few scratch words sit in shared memory (592 B per pair).  Every
// per-face loop of the algorithm (visibility, closest-face search, face construction) becomes one
// step per lane, the group never diverges internally, and an SM holds 24 warps.
//
// The arithmetic per face / per vertex is exactly that of epaIterate (axcd_narrow.cuh) — same
// expression trees, same tie rules (closest face: lowest slot; horizon: visible faces by ascending
// slot, edges in winding order; new faces into the lowest free slots) — so results are bit-identical
// to the oracle.  Pairs that need more than 18 vertices / 32 faces / 24 horizon edges, or whose GJK
// end simplex is not a tetrahedron, go to the full-cap fallback kernel.
#pragma once

#include "axcd_narrow.cuh"

namespace axcd {

constexpr int kCoopG = 8;
constexpr int kCoopGroups = 32 / kCoopG;   // pairs per warp
constexpr int kCoopVerts = 18;
constexpr int kCoopFaces = 32;
constexpr int kCoopRegs = kCoopFaces / kCoopG;   // face slots per lane
constexpr int kCoopEdges = 24;
constexpr int kCoopWarps = 8;
constexpr int kCoopThreads = kCoopWarps * 32;

// shared-memory words of one pair
constexpr int CW_CORE_A = 0;     // 19 words: c, e0, e1, e2, r, s, kind, nv, vertsOffset
constexpr int CW_CORE_B = 19;
constexpr int CW_ORIGIN = 38;    // 3
constexpr int CW_PAIR = 41;      // pairIdx, ia, ib, gjkStatus, queueIdx
constexpr int CW_Y = 46;         // 3 * kCoopVerts
constexpr int CW_ID = CW_Y + 3 * kCoopVerts;
constexpr int CW_ROW = CW_ID + kCoopVerts;       // one 32-bit "visible edge a->b" row per vertex
constexpr int CW_EDGE = CW_ROW + kCoopVerts;     // kCoopEdges packed 16-bit edges
constexpr int kCoopWords = CW_EDGE + kCoopEdges / 2;

__device__ __forceinline__ void storeCore(float* P, const Core& k, const float4* hullBase) {
    P[0] = k.c.x; P[1] = k.c.y; P[2] = k.c.z;
    P[3] = k.e0.x; P[4] = k.e0.y; P[5] = k.e0.z;
    P[6] = k.e1.x; P[7] = k.e1.y; P[8] = k.e1.z;
    P[9] = k.e2.x; P[10] = k.e2.y; P[11] = k.e2.z;
    P[12] = k.r;
    P[13] = k.s.x; P[14] = k.s.y; P[15] = k.s.z;
    P[16] = __int_as_float(k.kind);
    P[17] = __uint_as_float(k.nv);
    P[18] = __uint_as_float(k.verts ? (uint32_t)(k.verts - hullBase) : 0u);
}
__device__ __forceinline__ Core loadCore(const float* P, const float4* hullBase) {
    Core k;
    k.c = mk3(P[0], P[1], P[2]);
    k.e0 = mk3(P[3], P[4], P[5]);
    k.e1 = mk3(P[6], P[7], P[8]);
    k.e2 = mk3(P[9], P[10], P[11]);
    k.r = P[12];
    k.s = mk3(P[13], P[14], P[15]);
    k.kind = __float_as_int(P[16]);
    k.nv = __float_as_uint(P[17]);
    k.verts = hullBase + __float_as_uint(P[18]);
    return k;
}

struct CoopFaces {   // this lane's face slots: slot = j * 8 + (lane & 7)
    float nx[kCoopRegs], ny[kCoopRegs], nz[kCoopRegs], d[kCoopRegs];
    uint32_t fi[kCoopRegs];
};

__device__ __forceinline__ V3 coopY(const float* P, int i) { return mk3(P[CW_Y + 3 * i], P[CW_Y + 3 * i + 1], P[CW_Y + 3 * i + 2]); }

// plane of face (i0,i1,i2): same arithmetic as epaSetFace
__device__ __forceinline__ void coopPlane(const float* P, int i0, int i1, int i2, float& nx, float& ny, float& nz,
                                          float& d, uint32_t& fi) {
    const V3 p0 = coopY(P, i0);
    V3 n = cross3(coopY(P, i1) - p0, coopY(P, i2) - p0);
    const float len2 = dot3(n, n);
    fi = (uint32_t)i0 | ((uint32_t)i1 << 8) | ((uint32_t)i2 << 16);
    if (len2 <= 1e-30f) {
        nx = ny = nz = 0.0f;
        d = FLT_MAX;
        return;
    }
    const float inv = 1.0f / sqrtf(len2);
    n = n * inv;
    nx = n.x; ny = n.y; nz = n.z;
    d = dot3(n, p0);
}

// (d, slot) minimum over the 8 lanes of a group; slot < 0 means "none"
__device__ __forceinline__ void groupMinFace(uint32_t gmask, float& d, int& slot) {
#pragma unroll
    for (int off = 1; off < kCoopG; off <<= 1) {
        const float od = __shfl_xor_sync(gmask, d, off);
        const int os = __shfl_xor_sync(gmask, slot, off);
        if (os >= 0 && (slot < 0 || od < d || (od == d && os < slot))) {
            d = od;
            slot = os;
        }
    }
}

__global__ void __launch_bounds__(kCoopThreads, 3)
epaCoopKernel(NarrowQueues q, uint32_t queueCap, const uint2* __restrict__ pairs, const float* __restrict__ xf,
              const uint4* __restrict__ shapes, const float4* __restrict__ hull, NarrowParams cfg,
              AxcdContact* __restrict__ contacts, uint32_t maxContacts, const uint32_t* __restrict__ slots,
              float* __restrict__ pairDist, Counters* __restrict__ ctr) {
    __shared__ float sPairs[kCoopWarps * kCoopGroups * kCoopWords];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane & (kCoopG - 1), grp = lane / kCoopG, gbase = grp * kCoopG;
    const uint32_t gmask = 0xffu << gbase;
    float* P = sPairs + (warp * kCoopGroups + grp) * kCoopWords;
    uint32_t* PU = reinterpret_cast<uint32_t*>(P);
    const uint32_t count = min(ctr->epaCount, queueCap);
    const int maxFaces = (int)min(cfg.epaMaxFaces, (uint32_t)kEpaHardFaces);

    enum { EMPTY = 0, RUNNING = 1, DONE = 2 };
    int state = EMPTY;            // uniform within a group
    CoopFaces F;
    uint32_t alive = 0;           // face slots in use (group-uniform)
    int nv = 0, nf = 0, best = -1;
    float bd = FLT_MAX;
    uint32_t it = 0, status = 0;
    bool overflow = false, degenerate = false;
    uint32_t cur = 0, end = 0;    // warp's claimed queue range
    bool drained = false;

    while (true) {
        const uint32_t running = __ballot_sync(0xffffffffu, state == RUNNING);
        const uint32_t done = __ballot_sync(0xffffffffu, state == DONE);
        // ---- finalise, batched (>= 2 groups, or nothing is running) ---------------------------------
        if (done && (__popc(done) >= 2 * kCoopG || !running)) {
            if (state == DONE) {
                if (overflow) {
                    if (gl == 0) {
                        const uint32_t o = atomicAdd(&ctr->epaOverflow, 1u);
                        q.overflow[o] = PU[CW_PAIR + 4];   // (no spill: the full-cap kernel restarts the pair)
                    }
                } else {
                    // every lane computes the (uniform) result; lane 0 writes it
                    const Core A = loadCore(P + CW_CORE_A, hull);
                    EpaResult r;
                    if (degenerate) {
                        r = epaTouching(mk3(1.f, 0.f, 0.f), pointFromId(A, PU[CW_ID] & 0xffffu));
                    } else {
                        // fetch the closest face from its owner lane
                        const int owner = gbase + (best & (kCoopG - 1)), j = best >> 3;
                        float sx = F.nx[0], sy = F.ny[0], sz = F.nz[0], sd = F.d[0];
                        uint32_t sfi = F.fi[0];
#pragma unroll
                        for (int t = 1; t < kCoopRegs; ++t)
                            if (j == t) { sx = F.nx[t]; sy = F.ny[t]; sz = F.nz[t]; sd = F.d[t]; sfi = F.fi[t]; }
                        sx = __shfl_sync(gmask, sx, owner);
                        sy = __shfl_sync(gmask, sy, owner);
                        sz = __shfl_sync(gmask, sz, owner);
                        sd = __shfl_sync(gmask, sd, owner);
                        sfi = __shfl_sync(gmask, sfi, owner);
                        const int i0 = sfi & 0xffu, i1 = (sfi >> 8) & 0xffu, i2 = (sfi >> 16) & 0xffu;
                        r.n = mk3(sx, sy, sz);
                        r.depth = (sd > 0.0f) ? sd : 0.0f;
                        float la, lb, lc;
                        int m;
                        const V3 p = closestTriangle(coopY(P, i0), coopY(P, i1), coopY(P, i2), la, lb, lc, m);
                        r.pa = (pointFromId(A, PU[CW_ID + i0] & 0xffffu) * la + pointFromId(A, PU[CW_ID + i1] & 0xffffu) * lb) +
                               pointFromId(A, PU[CW_ID + i2] & 0xffffu) * lc;
                        r.pb = r.pa - p;
                        r.status = status;
                        r.overflow = false;
                    }
                    if (gl == 0) {
                        EpaLane L;
                        L.A = A;
                        L.B = loadCore(P + CW_CORE_B, hull);
                        L.origin = mk3(P[CW_ORIGIN], P[CW_ORIGIN + 1], P[CW_ORIGIN + 2]);
                        L.pairIdx = PU[CW_PAIR];
                        L.ia = PU[CW_PAIR + 1];
                        L.ib = PU[CW_PAIR + 2];
                        L.status = PU[CW_PAIR + 3];
                        L.queueIdx = PU[CW_PAIR + 4];
                        epaEmit(L, r, contacts, maxContacts, slots, pairDist, ctr);
                    }
                }
                state = EMPTY;
            }
        }
        // ---- refill, batched ---------------------------------------------------------------------------
        const uint32_t empty = __ballot_sync(0xffffffffu, state == EMPTY);
        if (!drained && empty && (__popc(empty) >= 2 * kCoopG || !running)) {
            if (cur >= end) {
                uint32_t c = 0;
                if (lane == 0) c = atomicAdd(&ctr->epaCursor, (uint32_t)kEpaChunk);
                cur = __shfl_sync(0xffffffffu, c, 0);
                end = min(cur + kEpaChunk, count);
                if (cur >= count) drained = true;
            }
            if (!drained) {
                // groups (not lanes) take items: rank among the empty groups
                uint32_t groupsEmptyBelow = 0, groupsEmpty = 0;
#pragma unroll
                for (int g = 0; g < kCoopGroups; ++g) {
                    const bool e = (empty >> (g * kCoopG)) & 1u;
                    groupsEmpty += e ? 1u : 0u;
                    groupsEmptyBelow += (e && g < grp) ? 1u : 0u;
                }
                const uint32_t mine = cur + groupsEmptyBelow;
                if (state == EMPTY && mine < end) {
                    // all lanes of the group load the item redundantly; lane 0 writes shared memory
                    const EpaWork* wk = q.work + mine;
                    const uint4 h = __ldg(reinterpret_cast<const uint4*>(wk));
                    const float4 f0 = __ldg(reinterpret_cast<const float4*>(wk) + 1),
                                 f1 = __ldg(reinterpret_cast<const float4*>(wk) + 2),
                                 f2 = __ldg(reinterpret_cast<const float4*>(wk) + 3);
                    const uint4 idv = __ldg(reinterpret_cast<const uint4*>(wk) + 4);
                    const int n0 = (int)(h.z & 0xffu);
                    const uint2 pk = __ldg(pairs + h.x);
                    const BodyPose ta = loadPose(xf, pk.x), tb = loadPose(xf, pk.y);
                    const uint4 sa = __ldg(shapes + pk.x), sb = __ldg(shapes + pk.y);
                    const V3 origin = ta.p;
                    if (gl == 0) {
                        storeCore(P + CW_CORE_A, makeCore(ta, sa, hull, origin), hull);
                        storeCore(P + CW_CORE_B, makeCore(tb, sb, hull, origin), hull);
                        P[CW_ORIGIN] = origin.x; P[CW_ORIGIN + 1] = origin.y; P[CW_ORIGIN + 2] = origin.z;
                        PU[CW_PAIR] = h.x;
                        PU[CW_PAIR + 1] = pk.x;
                        PU[CW_PAIR + 2] = pk.y;
                        PU[CW_PAIR + 3] = (h.z & 0x80000000u) ? (uint32_t)AXCD_ERR_GJK_NO_CONVERGE : 0u;
                        PU[CW_PAIR + 4] = mine;
                        P[CW_Y + 0] = f0.x; P[CW_Y + 1] = f0.y; P[CW_Y + 2] = f0.z;
                        P[CW_Y + 3] = f0.w; P[CW_Y + 4] = f1.x; P[CW_Y + 5] = f1.y;
                        P[CW_Y + 6] = f1.z; P[CW_Y + 7] = f1.w; P[CW_Y + 8] = f2.x;
                        P[CW_Y + 9] = f2.y; P[CW_Y + 10] = f2.z; P[CW_Y + 11] = f2.w;
                        PU[CW_ID] = idv.x; PU[CW_ID + 1] = idv.y; PU[CW_ID + 2] = idv.z; PU[CW_ID + 3] = idv.w;
                    }
                    __syncwarp(gmask);
                    overflow = false;
                    degenerate = false;
                    status = 0;
                    it = 0;
                    if (n0 != 4) {
                        overflow = true;   // needs the simplex blow-up: full-cap path
                        state = DONE;
                    } else {
                        // orientation: make (0,1,2) face away from vertex 3
                        const V3 y0 = coopY(P, 0), y1 = coopY(P, 1), y2 = coopY(P, 2), y3 = coopY(P, 3);
                        const bool swap = dot3(cross3(y1 - y0, y2 - y0), y3 - y0) > 0.0f;
                        __syncwarp(gmask);
                        if (swap && gl == 0) {
                            P[CW_Y + 0] = y1.x; P[CW_Y + 1] = y1.y; P[CW_Y + 2] = y1.z;
                            P[CW_Y + 3] = y0.x; P[CW_Y + 4] = y0.y; P[CW_Y + 5] = y0.z;
                            const uint32_t t0 = PU[CW_ID];
                            PU[CW_ID] = PU[CW_ID + 1];
                            PU[CW_ID + 1] = t0;
                        }
                        __syncwarp(gmask);
#pragma unroll
                        for (int t = 0; t < kCoopRegs; ++t) {
                            F.nx[t] = F.ny[t] = F.nz[t] = 0.0f;
                            F.d[t] = FLT_MAX;
                            F.fi[t] = 0;
                        }
                        if (gl < 4) {
                            const int a0 = (gl == 3) ? 1 : 0;
                            const int a1 = (gl == 0) ? 1 : ((gl == 2) ? 2 : 3);
                            const int a2 = (gl == 0) ? 2 : ((gl == 1) ? 1 : ((gl == 2) ? 3 : 2));
                            coopPlane(P, a0, a1, a2, F.nx[0], F.ny[0], F.nz[0], F.d[0], F.fi[0]);
                        }
                        alive = 0xfu;
                        nf = 4;
                        nv = 4;
                        float md = (gl < 4) ? F.d[0] : FLT_MAX;
                        int ms = (gl < 4 && F.d[0] < FLT_MAX) ? gl : -1;
                        groupMinFace(gmask, md, ms);
                        best = ms;
                        bd = md;
                        state = RUNNING;
                    }
                }
                cur = min(cur + groupsEmpty, end);
            }
        }
        // ---- one expansion step for every running group ----------------------------------------------------
        const uint32_t nowRunning = __ballot_sync(0xffffffffu, state == RUNNING);
        if (!nowRunning) {
            if (drained && !__ballot_sync(0xffffffffu, state == DONE)) break;
            continue;
        }
        if (state != RUNNING) continue;

        if (best < 0) {   // every face degenerate
            degenerate = true;
            state = DONE;
            continue;
        }
        // closest face normal from its owner lane
        float bnx, bny, bnz;
        {
            const int owner = gbase + (best & (kCoopG - 1)), j = best >> 3;
            float sx = F.nx[0], sy = F.ny[0], sz = F.nz[0];
#pragma unroll
            for (int t = 1; t < kCoopRegs; ++t)
                if (j == t) { sx = F.nx[t]; sy = F.ny[t]; sz = F.nz[t]; }
            bnx = __shfl_sync(gmask, sx, owner);
            bny = __shfl_sync(gmask, sy, owner);
            bnz = __shfl_sync(gmask, sz, owner);
        }
        const V3 bn = mk3(bnx, bny, bnz);
        uint32_t wid;
        V3 w;
        {
            const Core A = loadCore(P + CW_CORE_A, hull);
            const Core B = loadCore(P + CW_CORE_B, hull);
            w = supportDiff(A, B, bn, wid);
        }
        const float dw = dot3(w, bn);
        const float scale = (bd > 1.0f) ? bd : 1.0f;
        if (dw - bd <= cfg.epaTol * scale) {
            state = DONE;
            continue;
        }
        {
            bool dup = false;
            for (int v = gl; v < nv; v += kCoopG) dup = dup || same3(w, coopY(P, v));
            if (__ballot_sync(gmask, dup) & gmask) {
                state = DONE;
                continue;
            }
        }
        if (it >= cfg.epaMaxIters || nv >= kEpaHardVerts) {
            status = AXCD_ERR_EPA_NO_CONVERGE;
            state = DONE;
            continue;
        }
        // ---- visibility (one step per lane-owned face) + closest surviving face ----------------------
        const float wl = fabsf(w.x) + fabsf(w.y) + fabsf(w.z);
        const float visEps = 1e-6f * ((wl > 1.0f) ? wl : 1.0f);
        uint32_t vis = 0;
        float nbd = FLT_MAX;
        int nbest = -1;
#pragma unroll
        for (int j = 0; j < kCoopRegs; ++j) {
            const int slot = j * kCoopG + gl;
            const bool al = (alive >> slot) & 1u;
            const bool v = al && (dot3(mk3(F.nx[j], F.ny[j], F.nz[j]), w) - F.d[j] > visEps);
            vis |= ((__ballot_sync(gmask, v) >> gbase) & 0xffu) << (j * kCoopG);
            if (al && !v && F.d[j] < nbd) {   // ascending slot order within the lane: j-major
                nbd = F.d[j];
                nbest = slot;
            }
        }
        // ---- visible directed edges -------------------------------------------------------------------------
        for (int v = gl; v < nv; v += kCoopG) PU[CW_ROW + v] = 0u;
        __syncwarp(gmask);
#pragma unroll
        for (int j = 0; j < kCoopRegs; ++j) {
            if ((vis >> (j * kCoopG + gl)) & 1u) {
                const uint32_t fi = F.fi[j];
                const uint32_t v0 = fi & 0xffu, v1 = (fi >> 8) & 0xffu, v2 = (fi >> 16) & 0xffu;
                atomicOr(&PU[CW_ROW + v0], 1u << v1);
                atomicOr(&PU[CW_ROW + v1], 1u << v2);
                atomicOr(&PU[CW_ROW + v2], 1u << v0);
            }
        }
        __syncwarp(gmask);
        // ---- horizon in canonical order ---------------------------------------------------------------------
        int nh = 0;
        uint32_t starts = 0, ends = 0;
        bool dupEdge = false;
#pragma unroll
        for (int j = 0; j < kCoopRegs; ++j) {
            const bool isVis = (vis >> (j * kCoopG + gl)) & 1u;
            uint32_t keep = 0, e0 = 0, e1 = 0, e2 = 0;
            if (isVis) {
                const uint32_t fi = F.fi[j];
                const uint32_t v0 = fi & 0xffu, v1 = (fi >> 8) & 0xffu, v2 = (fi >> 16) & 0xffu;
                if (!((PU[CW_ROW + v1] >> v0) & 1u)) { keep |= 1u; e0 = v0 | (v1 << 8); }
                if (!((PU[CW_ROW + v2] >> v1) & 1u)) { keep |= 2u; e1 = v1 | (v2 << 8); }
                if (!((PU[CW_ROW + v0] >> v2) & 1u)) { keep |= 4u; e2 = v2 | (v0 << 8); }
            }
            const uint32_t c = __popc(keep);
            const uint32_t b0 = (__ballot_sync(gmask, c & 1u) >> gbase) & 0xffu;
            const uint32_t b1 = (__ballot_sync(gmask, c & 2u) >> gbase) & 0xffu;
            const uint32_t lt = (1u << gl) - 1u;
            int pos = nh + __popc(b0 & lt) + 2 * __popc(b1 & lt);
            nh += __popc(b0) + 2 * __popc(b1);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (!((keep >> k) & 1u)) continue;
                const uint32_t ed = (k == 0) ? e0 : ((k == 1) ? e1 : e2);
                const uint32_t ea = ed & 0xffu, eb = ed >> 8;
                if (((starts >> ea) & 1u) || ((ends >> eb) & 1u)) dupEdge = true;
                starts |= 1u << ea;
                ends |= 1u << eb;
                if (pos < kCoopEdges) {
                    // two 16-bit edges per word; halfword stores from different lanes do not interfere
                    reinterpret_cast<unsigned short*>(PU + CW_EDGE)[pos] = (unsigned short)ed;
                }
                ++pos;
            }
        }
        // loop check: every vertex starts / ends at most one horizon edge (across the whole group)
        const uint32_t allStarts = __reduce_or_sync(gmask, starts), allEnds = __reduce_or_sync(gmask, ends);
        const bool anyDup = (__ballot_sync(gmask, dupEdge) & gmask) != 0;
        const bool loopOk = nh >= 3 && !anyDup && __popc(allStarts) == nh && __popc(allEnds) == nh;
        const int nalive = __popc(alive), nvis = __popc(vis);
        if (!loopOk || nalive - nvis + nh > maxFaces) {
            status = AXCD_ERR_EPA_NO_CONVERGE;
            state = DONE;
            continue;
        }
        if (nh > kCoopEdges || nv >= kCoopVerts || nalive - nvis + nh > kCoopFaces) {
            overflow = true;
            state = DONE;
            continue;
        }
        // ---- add the vertex, build the new faces in the lowest free slots -------------------------------------
        const int wi = nv;
        if (gl == 0) {
            P[CW_Y + 3 * wi] = w.x; P[CW_Y + 3 * wi + 1] = w.y; P[CW_Y + 3 * wi + 2] = w.z;
            PU[CW_ID + wi] = wid;
        }
        nv++;
        alive &= ~vis;
        __syncwarp(gmask);
        // S = the nh lowest free slots
        uint32_t S = 0;
        {
            uint32_t freeSlots = ~alive;
            for (int h = 0; h < nh; ++h) {
                const uint32_t low = freeSlots & (0u - freeSlots);
                S |= low;
                freeSlots ^= low;
            }
        }
        const unsigned short* e16 = reinterpret_cast<const unsigned short*>(PU + CW_EDGE);
#pragma unroll
        for (int j = 0; j < kCoopRegs; ++j) {
            const int slot = j * kCoopG + gl;
            if ((S >> slot) & 1u) {
                const int h = __popc(S & ((1u << slot) - 1u));
                const uint32_t ed = e16[h];
                coopPlane(P, (int)(ed & 0xffu), (int)(ed >> 8), wi, F.nx[j], F.ny[j], F.nz[j], F.d[j], F.fi[j]);
                if (F.d[j] < nbd || (F.d[j] == nbd && slot < nbest)) {
                    nbd = F.d[j];
                    nbest = slot;
                }
            }
        }
        alive |= S;
        nf = max(nf, 32 - __clz(S));
        groupMinFace(gmask, nbd, nbest);
        best = nbest;
        bd = nbd;
        it++;
        __syncwarp(gmask);
    }
}

}  // namespace axcd
