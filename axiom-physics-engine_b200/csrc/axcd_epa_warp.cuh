// Full-cap EPA for the pairs whose polytope outgrew the fast path: ONE WARP PER PAIR.
//
// These are the longest expansions of the step (up to cfg.epaMaxIters steps on up to 64 faces), and
// there are few of them (~1 % of the EPA pairs), so what matters is the latency of one expansion
// step, not the instruction count: a thread that walks 64 faces serially leaves the step waiting
// ~100 us for a handful of pairs.  Here the 64 face slots are striped over the 32 lanes (slot s
// lives in lane s & 31, register s >> 5), so every per-face loop of epaIterate becomes one or two
// steps: visibility = two ballots, closest face = one shuffle reduction, horizon = prefix sums over
// the visible faces, new faces = one plane computation per lane.  Vertices, edge-bit rows and the
// horizon list sit in shared memory (2.6 KB per warp, laid out like the full-cap Poly so epaInit
// and the spill restore can be reused).
//
// Arithmetic and tie rules are exactly those of epaIterate / the oracle: closest face = lowest slot
// among the minima; horizon = visible faces by ascending slot, their edges in winding order, an
// edge kept iff its reverse is not an edge of a visible face; new faces fill the lowest free slots
// in horizon order.  Results are bit-identical to the per-thread path.
#pragma once

#include "axcd_narrow.cuh"

namespace axcd {

#ifndef AXCD_WARPFB_BLOCKS
#define AXCD_WARPFB_BLOCKS 4
#endif
constexpr int kWarpFbBlocksPerSM = AXCD_WARPFB_BLOCKS;
constexpr int kWarpFbWarps = 4;
constexpr int kWarpFbThreads = kWarpFbWarps * 32;
// MAXE = 2 x 192: the horizon list keeps one edge per word here (lanes write it concurrently)
using WarpPoly = Poly<kEpaHardVerts, kEpaHardFaces, kEpaHardFaces * 6, 1>;
constexpr int kWarpFbScratch = 4;   // start / end vertex masks of the horizon (2 x 64 bit)
constexpr int kWarpFbWords = WarpPoly::kWords + kWarpFbScratch;
static_assert(kEpaHardFaces == 64 && kEpaHardVerts <= 64, "two face slots per lane, 64-bit vertex masks");

struct WarpFaces {   // this lane's two face slots: slot = j * 32 + lane
    float nx[2], ny[2], nz[2], d[2];
    uint32_t fi[2];
};

// plane of face (i0,i1,i2): the arithmetic of epaSetFace, result into registers
__device__ __forceinline__ void warpPlane(const WarpPoly& e, int i0, int i1, int i2, WarpFaces& F, int j) {
    const V3 p0 = e.y(i0);
    V3 n = cross3(e.y(i1) - p0, e.y(i2) - p0);
    const float len2 = dot3(n, n);
    F.fi[j] = (uint32_t)i0 | ((uint32_t)i1 << 8) | ((uint32_t)i2 << 16);
    if (len2 <= 1e-30f) {
        F.nx[j] = 0.f; F.ny[j] = 0.f; F.nz[j] = 0.f; F.d[j] = FLT_MAX;
        return;
    }
    const float inv = 1.0f / sqrtf(len2);
    n = n * inv;
    F.nx[j] = n.x; F.ny[j] = n.y; F.nz[j] = n.z;
    F.d[j] = dot3(n, p0);
}

template <bool CYL>
__global__ void __launch_bounds__(kWarpFbThreads, AXCD_WARPFB_BLOCKS)
epaWarpFallbackKernel(NarrowQueues q, const uint2* __restrict__ pairs, const float* __restrict__ xf,
                      const uint4* __restrict__ shapes, const float4* __restrict__ hull, NarrowParams cfg,
                      AxcdContact* __restrict__ contacts, uint32_t maxContacts, const uint32_t* __restrict__ slots,
                      float* __restrict__ pairDist, Counters* __restrict__ ctr) {
    __shared__ float sMem[kWarpFbWarps * kWarpFbWords];
    constexpr uint32_t kFull = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    WarpPoly poly;
    poly.base = sMem + warp * kWarpFbWords;
    uint32_t* scratch = reinterpret_cast<uint32_t*>(poly.base + WarpPoly::kWords);
    const uint32_t count = ctr->epaOverflow;
    const int maxFaces = (int)min(cfg.epaMaxFaces, (uint32_t)kEpaHardFaces);

    while (true) {   // persistent warps claim overflow items one at a time (their lengths vary widely)
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(&ctr->fallbackCursor, 1u);
        item = __shfl_sync(kFull, item, 0);
        if (item >= count) break;
        EpaLaneT<CYL> L;
        EpaState<WarpPoly::Mask> st;
        EpaResult r;
        WarpFaces F;
        const uint32_t ov = q.overflow[item];
        const bool resume = (ov & 0x80000000u) != 0;
        L.queueIdx = ov & 0x7fffffffu;
        __syncwarp();   // the previous pair's shared-memory state is dead
        // every lane builds the two cores (registers); lane 0 alone runs the scalar epaInit on the
        // warp's polytope when there is no spilled state to resume from
        int finished = epaBegin(q.work + L.queueIdx, pairs, xf, shapes, hull, poly, L, st, r, resume || lane != 0);
        finished = __shfl_sync(kFull, finished, 0);
        __syncwarp();
        if (finished) {
            if (lane == 0) epaEmit(L, r, contacts, maxContacts, slots, pairDist, ctr);
            continue;
        }
        uint64_t alive;
        int nv;
        uint32_t it;
#pragma unroll
        for (int j = 0; j < 2; ++j) { F.nx[j] = F.ny[j] = F.nz[j] = 0.f; F.d[j] = FLT_MAX; F.fi[j] = 0u; }
        if (resume) {
            using FP = Poly<kEpaFastVerts, kEpaFastFaces, kEpaFastEdges, 1>;
            const float* src = q.spill + (size_t)item * (FP::kWords + kSpillStateWords);
            const uint32_t* su = reinterpret_cast<const uint32_t*>(src + FP::kWords);
            nv = (int)su[0];
            const int nf = (int)su[1];
            alive = (uint64_t)su[4];
            it = su[5];
            FP fp;
            fp.base = const_cast<float*>(src);
            if (lane < nv) {   // kEpaFastVerts <= 16
                poly.setY(lane, fp.y(lane));
                poly.setId(lane, fp.id(lane));
            }
            if (lane < nf) {   // kEpaFastFaces <= 32: all in register 0
                const V3 n = fp.fn(lane);
                F.nx[0] = n.x; F.ny[0] = n.y; F.nz[0] = n.z; F.d[0] = fp.fd(lane);
                F.fi[0] = fp.fi(lane);
            }
        } else {
            nv = __shfl_sync(kFull, st.nv, 0);
            alive = 0xfull;   // the four faces epaInit wrote to the shared polytope
            it = 0;
            if (lane < 4) {
                const V3 n = poly.fn(lane);
                F.nx[0] = n.x; F.ny[0] = n.y; F.nz[0] = n.z; F.d[0] = poly.fd(lane);
                F.fi[0] = poly.fi(lane);
            }
        }
        __syncwarp();

        uint32_t status = 0;
        bool degenerate = false;
        int best = -1;
        while (true) {
            // ---- closest face: lowest slot among the minima ----------------------------------------
            float bd = FLT_MAX;
            int bs = 64;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int s = j * 32 + lane;
                if (((alive >> s) & 1ull) && F.d[j] < bd) { bd = F.d[j]; bs = s; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float od = __shfl_xor_sync(kFull, bd, off);
                const int os = __shfl_xor_sync(kFull, bs, off);
                if (od < bd || (od == bd && os < bs)) { bd = od; bs = os; }
            }
            if (bs >= 64) { degenerate = true; break; }
            best = bs;
            const int bl = best & 31, bj = best >> 5;
            const V3 bn = mk3(__shfl_sync(kFull, bj ? F.nx[1] : F.nx[0], bl), __shfl_sync(kFull, bj ? F.ny[1] : F.ny[0], bl),
                              __shfl_sync(kFull, bj ? F.nz[1] : F.nz[0], bl));
            uint32_t wid;
            const V3 w = supportDiff(L.A, L.B, bn, wid);
            const float dw = dot3(w, bn);
            const float scale = (bd > 1.0f) ? bd : 1.0f;
            if (dw - bd <= cfg.epaTol * scale) break;
            bool dup = false;
            for (int i = lane; i < nv; i += 32) dup = dup || same3(w, poly.y(i));
            if (__any_sync(kFull, dup)) break;
            if (it >= cfg.epaMaxIters || nv >= kEpaHardVerts) { status = AXCD_ERR_EPA_NO_CONVERGE; break; }
            // ---- visible faces ---------------------------------------------------------------------
            const float wl = fabsf(w.x) + fabsf(w.y) + fabsf(w.z);
            const float visEps = 1e-6f * ((wl > 1.0f) ? wl : 1.0f);
            bool v[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int s = j * 32 + lane;
                v[j] = ((alive >> s) & 1ull) && (dot3(mk3(F.nx[j], F.ny[j], F.nz[j]), w) - F.d[j] > visEps);
            }
            const uint64_t vis = (uint64_t)__ballot_sync(kFull, v[0]) | ((uint64_t)__ballot_sync(kFull, v[1]) << 32);
            // ---- directed edges of the visible faces, one bit row per start vertex -------------------
            for (int i = lane; i < 2 * nv; i += 32) poly.u(WarpPoly::kRowBase + i) = 0u;
            if (lane < kWarpFbScratch) scratch[lane] = 0u;
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (!v[j]) continue;
                const uint32_t fi = F.fi[j];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const uint32_t ea = (fi >> (8 * k)) & 0xffu, eb = (fi >> (8 * ((k + 1) % 3))) & 0xffu;
                    atomicOr(&poly.u(WarpPoly::kRowBase + 2 * ea + (eb >> 5)), 1u << (eb & 31));
                }
            }
            __syncwarp();
            // ---- horizon in canonical order + simple-loop check --------------------------------------
            int nh = 0;
            bool bad = false;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint32_t keep = 0;
                const uint32_t fi = F.fi[j];
                if (v[j]) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const uint32_t ea = (fi >> (8 * k)) & 0xffu, eb = (fi >> (8 * ((k + 1) % 3))) & 0xffu;
                        if (!poly.edgeBit(eb, ea)) keep |= 1u << k;
                    }
                }
                const int cnt = __popc(keep);
                int incl = cnt;   // inclusive prefix sum over the lanes = ascending slot order
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const int t = __shfl_up_sync(kFull, incl, off);
                    if (lane >= off) incl += t;
                }
                int pos = nh + incl - cnt;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (!((keep >> k) & 1u)) continue;
                    const uint32_t ea = (fi >> (8 * k)) & 0xffu, eb = (fi >> (8 * ((k + 1) % 3))) & 0xffu;
                    poly.edgeWord(pos) = ea | (eb << 8);
                    ++pos;
                    // every vertex may start one horizon edge and end one
                    const uint32_t so = atomicOr(&scratch[ea >> 5], 1u << (ea & 31));
                    const uint32_t eo = atomicOr(&scratch[2 + (eb >> 5)], 1u << (eb & 31));
                    if (((so >> (ea & 31)) & 1u) || ((eo >> (eb & 31)) & 1u)) bad = true;
                }
                nh += __shfl_sync(kFull, incl, 31);
            }
            const bool loopOk = !__any_sync(kFull, bad) && nh >= 3;
            const int nalive = __popcll(alive), nvis = __popcll(vis);
            if (!loopOk || nalive - nvis + nh > maxFaces) { status = AXCD_ERR_EPA_NO_CONVERGE; break; }
            // ---- add the vertex, replace the visible faces by the fan over the horizon ------------------
            const int wi = nv;
            if (lane == 0) {
                poly.setY(wi, w);
                poly.setId(wi, wid);
            }
            nv++;
            alive &= ~vis;
            __syncwarp();   // vertex + horizon list visible to every lane
            const uint64_t freeSlots = ~alive;
            bool built[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int s = j * 32 + lane;
                built[j] = false;
                if ((freeSlots >> s) & 1ull) {
                    const int rank = __popcll(freeSlots & ((1ull << s) - 1ull));   // h-th edge -> h-th lowest free slot
                    if (rank < nh) {
                        const uint32_t ed = poly.edgeWord(rank);
                        warpPlane(poly, (int)(ed & 0xffu), (int)(ed >> 8), wi, F, j);
                        built[j] = true;
                    }
                }
            }
            alive |= (uint64_t)__ballot_sync(kFull, built[0]) | ((uint64_t)__ballot_sync(kFull, built[1]) << 32);
            it++;
            __syncwarp();   // horizon list consumed before the next step overwrites it
        }

        // ---- result: the closest face's plane, witness point from the support ids ------------------------
        if (degenerate) {
            r = epaTouching(mk3(1.f, 0.f, 0.f), pointFromId(L.A, poly.id(0) & 0xffffu));
        } else {
            const int bl = best & 31, bj = best >> 5;
            const uint32_t fi = __shfl_sync(kFull, bj ? F.fi[1] : F.fi[0], bl);
            const float fdist = __shfl_sync(kFull, bj ? F.d[1] : F.d[0], bl);
            r.n = mk3(__shfl_sync(kFull, bj ? F.nx[1] : F.nx[0], bl), __shfl_sync(kFull, bj ? F.ny[1] : F.ny[0], bl),
                      __shfl_sync(kFull, bj ? F.nz[1] : F.nz[0], bl));
            const int i0 = fi & 0xffu, i1 = (fi >> 8) & 0xffu, i2 = (fi >> 16) & 0xffu;
            r.depth = (fdist > 0.0f) ? fdist : 0.0f;
            r.status = status;
            r.overflow = false;
            float la, lb, lc;
            int m;
            const V3 p = closestTriangle(poly.y(i0), poly.y(i1), poly.y(i2), la, lb, lc, m);
            r.pa = (pointFromId(L.A, poly.id(i0) & 0xffffu) * la + pointFromId(L.A, poly.id(i1) & 0xffffu) * lb) +
                   pointFromId(L.A, poly.id(i2) & 0xffffu) * lc;
            r.pb = r.pa - p;
        }
        if (lane == 0) epaEmit(L, r, contacts, maxContacts, slots, pairDist, ctr);
    }
}

}  // namespace axcd
