// Shared device/host declarations for the collision path kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "axcd.h"

namespace axcd {

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs.  Sizing constant (chunk capacity, classifier grid); launch grids of the
                               // persistent kernels use the device's own count (AxcdContext::numSMs, queried at create)

// decoupled look-back status words: 2 flag bits + 30 value bits
constexpr uint32_t kFlagAggregate = 1u << 30;
constexpr uint32_t kFlagInclusive = 2u << 30;
constexpr uint32_t kFlagMask = 3u << 30;
constexpr uint32_t kValueMask = ~kFlagMask;

// ---- float3 helpers.  The library is built -fmad=false, so nothing is contracted implicitly: every
// expression rounds exactly as written and the only fused operations are the explicit fmaf calls
// below (reference operators: include/axiom/math/vec3.hpp:52-191). ----
struct V3 {
    float x, y, z;
};
__host__ __device__ __forceinline__ V3 mk3(float x, float y, float z) { return V3{x, y, z}; }
__host__ __device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ __forceinline__ V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__host__ __device__ __forceinline__ V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
// dot / cross use ONE explicit fused pattern (fmaf = a single rounding), the same one the CPU oracle uses.
// It is the pattern GCC 13 generates for the reference's own formulas (vec3.hpp:179-191) under the reference's
// own x86 flags (-mavx2 -mfma, default contraction): tests/test_oracle_vs_reference.py compares the two bit for
// bit on 10^5 random inputs.  dot: fma(z, z', fma(x, x', y*y')); cross: (y z' - z y', z x' - x z') with the
// FIRST product rounded and the second fused, (x y' - y x') with the first fused and the second rounded.
__host__ __device__ __forceinline__ float dot3(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.x, b.x, a.y * b.y)); }
__host__ __device__ __forceinline__ V3 cross3(V3 a, V3 b) {
    return mk3(fmaf(-a.z, b.y, a.y * b.z), fmaf(-a.x, b.z, a.z * b.x), fmaf(a.x, b.y, -(a.y * b.x)));
}
__host__ __device__ __forceinline__ bool same3(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

// glm::quat * glm::vec3 as reached from Quat::operator*(Vec3) (reference: src/math/quat.cpp:33-38;
// formula: SURVEY.md Appendix B).  q = (x,y,z,w).
__device__ __forceinline__ V3 quatRotate(float4 q, V3 v) {
    V3 u = mk3(q.x, q.y, q.z);
    V3 uv = cross3(u, v);
    V3 uuv = cross3(u, uv);
    V3 t = mk3(fmaf(uv.x, q.w, uuv.x), fmaf(uv.y, q.w, uuv.y), fmaf(uv.z, q.w, uuv.z));
    return mk3(fmaf(t.x, 2.0f, v.x), fmaf(t.y, 2.0f, v.y), fmaf(t.z, 2.0f, v.z));
}

// order-preserving float <-> uint mapping for atomicMin/atomicMax on floats
__device__ __forceinline__ uint32_t floatToOrdered(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float orderedToFloat(uint32_t u) {
    uint32_t v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(v);
#else
    float f;
    memcpy(&f, &v, 4);
    return f;
#endif
}

// Exclusive prefix of `tile` over per-tile status words (flag | value) of a single-pass scan: warp-wide look-back,
// 128 predecessor tiles per round (one 16-byte volatile load per lane).  All in-flight tiles sit in the "aggregate"
// state together, so the walk back to the last inclusive prefix is about one wave of tiles long: 4 words per lane
// make it 4x fewer dependent L2 round trips than a word per lane.  `status` must be readable up to the next
// multiple of 128 words past `tile`.  Call with all 32 lanes of a warp.
__device__ __forceinline__ uint32_t lookbackWide(const volatile uint32_t* status, uint32_t tile, int lane) {
    uint32_t excl = 0;
    for (int chunk = (int)((tile - 1u) >> 7); chunk >= 0; --chunk) {
        const uint32_t base = (uint32_t)chunk * 128u + 4u * (uint32_t)lane;
        uint32_t w[4];
        bool ready;
        do {
            asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3])
                         : "l"(status + base)
                         : "memory");
            ready = true;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (base + k < tile && (w[k] & kFlagMask) == 0u) ready = false;   // not published yet
        } while (!__all_sync(0xffffffffu, ready));
        int myTop = -1;   // highest of this lane's words that holds an inclusive prefix
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (base + k < tile && (w[k] & kFlagMask) == kFlagInclusive) myTop = k;
        const uint32_t hasInc = __ballot_sync(0xffffffffu, myTop >= 0);
        const int top = hasInc ? 31 - __clz(hasInc) : -1;   // the lane holding the nearest inclusive prefix
        uint32_t sum = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool take = base + k < tile && (lane > top || (lane == top && k >= myTop));
            sum += take ? (w[k] & kValueMask) : 0u;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        excl += sum;
        if (hasInc) break;
    }
    return excl;
}

constexpr int kNumClasses = 12;   // pair classes of the narrowphase (pairClass in axcd_narrow.cuh)

// One 64-byte LBVH internal node: both children's boxes live in the parent, so a traversal step is
// one aligned 64-byte read.  Leaf-ness and child indices follow from (first, split, last):
//   left  child covers sorted leaves [first, split]   -> leaf `split`   if first == split,
//                                                        else internal node `split`
//   right child covers sorted leaves [split+1, last]  -> leaf `split+1` if split+1 == last,
//                                                        else internal node `split+1`
struct __align__(64) BvhNode {
    float lminx, lminy, lminz, lmaxx;
    float lmaxy, lmaxz, rminx, rminy;
    float rminz, rmaxx, rmaxy, rmaxz;
    uint32_t first, split, last, pad;
};
static_assert(sizeof(BvhNode) == 64, "BvhNode must be one 64-byte record");

// device-side counters block (one per context)
struct Counters {
    uint32_t boundsMin[3];   // ordered-uint encoded minima of AABB centres
    uint32_t boundsMax[3];
    uint32_t pairCount;      // candidate pairs found (may exceed capacity)
    uint32_t contactCount;
    uint32_t epaCount;
    uint32_t gjkFailures;
    uint32_t epaFailures;
    uint32_t sortTicket[8];  // dynamic tile tickets, one per radix pass
    uint32_t gjkTicket;      // dynamic tile ticket of the GJK kernel
    uint32_t epaOverflow;    // EPA pairs that outgrew the shared-memory polytope caps
    uint32_t epaCursor;      // next unclaimed EPA queue item
    uint32_t scanTicket;     // tile ticket of the body-count scan
    uint32_t storedPairs;    // pairs actually stored (<= capacity) = total of the body-count scan
    uint32_t fallbackCursor; // next unclaimed overflow item (full-cap EPA)
    uint32_t gjkChunks;      // 32-pair class-homogeneous chunks queued for the GJK kernel
    uint32_t gjkChunkCursor; // next unclaimed chunk
    uint32_t gjkChunksGeneric;       // chunks of the classes that need GJK (stored from the back of the chunk array)
    uint32_t gjkChunkCursorGeneric;  // next unclaimed chunk of that list
    uint32_t travOverflow;   // set if a traversal stack ever filled up (the step then reports 505 instead of losing pairs)
    uint32_t movedBodies;    // temporal coherence: bodies whose tight box left their fat box this step
    uint32_t manifoldPoints; // contact points over all manifolds (PhysicsWorldStats::contactPointCount)
    uint32_t sortFallback;   // bucket sort gave up (a bucket over its capacity): the LSD radix kernels sort instead
    uint32_t sortMaxBucket;  // largest Morton bucket of the step
};

// ---- lanes of a warp that hold the same small value ---------------------------------------------------------------
// Peer mask from one ballot per value bit instead of __match_any_sync: the match operation is throughput-limited on
// B200 (profiles/r02_experiments.md: the radix ranking loop went from 0.081 to 0.069 ms per 1 M-key sort with this).
template <int BITS>
__device__ __forceinline__ uint32_t peersByBallot(uint32_t v) {
    uint32_t peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < BITS; ++b) {
        const bool bit = (v >> b) & 1u;
        const uint32_t bal = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? bal : ~bal;
    }
    return peers;
}

}  // namespace axcd
