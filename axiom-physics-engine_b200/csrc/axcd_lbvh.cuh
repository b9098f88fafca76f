// Stage 2 (after the Morton sort): LBVH build (Karras 2012 topology over the sorted keys), bottom-up
// box fit, and the candidate-pair traversal.
//
// The candidate set is defined by the reference's closed-interval predicate AABB::intersects
// (include/axiom/math/aabb.hpp:132-135) on the exact refit floats: leaf boxes are copied, never
// recomputed, and internal boxes are exact min/max unions, so the tree only prunes — it can never
// change the set.  Each unordered pair is reported once: leaf i (sorted position) only descends into
// subtrees that contain a sorted position > i.
#pragma once

#include "axcd_common.cuh"

namespace axcd {

// ---- leaves in Morton order ------------------------------------------------------------------------
// leafLo[k] = (min.xyz, bits(bodyIndex)), leafHi[k] = (max.xyz, bits(world id)) for the body at sorted position k;
// written by segGatherBottomKernel below, on its way up the range tree.

// leafHi.w currently holds the world id; replace it by the last sorted index of that world so the
// traversal can prune other worlds with one compare.  worldEnd[w] filled by markWorldEndsKernel.
__global__ void markWorldEndsKernel(const uint32_t* __restrict__ sortedKeys, uint32_t n, int worldShift,
                                    uint32_t* __restrict__ worldEnd) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t w = sortedKeys[k] >> worldShift;
    if (k + 1 == n || (sortedKeys[k + 1] >> worldShift) != w) worldEnd[w] = k;
}

// ---- Karras topology ----------------------------------------------------------------------------
// delta(i,j): length of the common prefix of the (key, position) pairs; -1 outside [0,n).
__device__ __forceinline__ int karrasDelta(const uint32_t* __restrict__ keys, int n, int i, uint32_t ki, int j) {
    if (j < 0 || j >= n) return -1;
    const uint32_t kj = __ldg(keys + j);
    const uint32_t x = ki ^ kj;
    return x ? __clz(x) : 32 + __clz((uint32_t)i ^ (uint32_t)j);
}

// ---- compact traversal nodes -----------------------------------------------------------------------
// The pair traversal is bound by L1TEX throughput (ncu: 80 % of peak, every visited node costs four
// divergent 16-byte loads per lane), so the nodes it walks are 32 bytes instead of 64: both child boxes
// quantised to 16 bits per coordinate on the grid of the scene box, plus (split | leaf flags, last).
// `first` is not stored: the left child inherits its parent's, the right child's equals its own index.
// The quantiser Q is ONE monotone function (subtract, multiply, floor, clamp — each monotone under
// round-to-nearest) applied to every coordinate, min and max alike, so a.min <= b.max implies
// Q(a.min) <= Q(b.max): the quantised test can only add candidates, never lose one, and every leaf hit
// is confirmed with AABB::intersects on the exact leaf boxes.
struct __align__(32) Node32 {
    uint32_t la, lb, lc;      // left child box, packed for the SWAR overlap test (see packChild)
    uint32_t ra, rb, rc;      // right child box
    uint32_t split;           // bits 0..29 split, bit 30: left child is a leaf, bit 31: right child is a leaf
    uint32_t last;
};
static_assert(sizeof(Node32) == 32, "Node32 must be one 32-byte record");
#ifndef AXCD_LDG256
#define AXCD_LDG256 1
#endif
// One node = one 256-bit load (sm_100: LDG.E.256) instead of two 128-bit ones: the traversal is co-limited by the LSU
// data pipe (ncu: 68 % of its wavefront peak with two loads per node) and by issue slots, and this halves its node
// fetch instructions.  Read-only path (.nc), 32-byte aligned by the record type.
__device__ __forceinline__ void loadNode32(const Node32* __restrict__ p, uint4& lo, uint4& hi) {
#if AXCD_LDG256
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                 : "l"(p));
#else
    const uint4* q = reinterpret_cast<const uint4*>(p);
    lo = __ldg(q);
    hi = __ldg(q + 1);
#endif
}
constexpr uint32_t kSplitMask = 0x3fffffffu, kLeftLeaf = 1u << 30, kRightLeaf = 1u << 31;
constexpr uint32_t kGuard = 0x80008000u;   // bit 15 of each half: the borrow guard of the SWAR compare
constexpr float kQuantCells = 32767.0f;    // 15 bits per coordinate

struct QuantFrame {
    float s0x, s0y, s0z, ivx, ivy, ivz;
};
// Grid of the scene box = root of the range tree (heap index 1): 32767 cells per axis.
__device__ __forceinline__ QuantFrame loadQuantFrame(const float4* __restrict__ segLo, const float4* __restrict__ segHi) {
    const float4 a = __ldg(segLo + 1), b = __ldg(segHi + 1);
    QuantFrame f;
    f.s0x = a.x; f.s0y = a.y; f.s0z = a.z;
    const float ex = b.x - a.x, ey = b.y - a.y, ez = b.z - a.z;
    f.ivx = (ex > 0.0f && ex < 3.0e38f) ? kQuantCells / ex : 0.0f;   // degenerate / infinite axis: everything in cell 0
    f.ivy = (ey > 0.0f && ey < 3.0e38f) ? kQuantCells / ey : 0.0f;
    f.ivz = (ez > 0.0f && ez < 3.0e38f) ? kQuantCells / ez : 0.0f;
    return f;
}
__device__ __forceinline__ uint32_t quantCoord(float x, float s0, float iv) {
    const float t = floorf((x - s0) * iv);
    return (uint32_t)fminf(fmaxf(t, 0.0f), kQuantCells);   // NaN -> 0
}
// A box is stored as three words of two 15-bit fields each, arranged so that the closed-interval
// overlap test against a query box is three guarded subtractions:
//   child:  a = min.x | min.y << 16      b = min.z | (C - max.z) << 16      c = (C - max.x) | (C - max.y) << 16
//   query:  a = max.x | max.y << 16      b = max.z | (C - min.z) << 16      c = (C - min.x) | (C - min.y) << 16   (| guard)
// with C = 32767.  Field by field, query >= child says: q.max >= c.min on x, y, z and c.max >= q.min on
// z, x, y — exactly AABB::intersects on the quantised boxes.  (query | guard) - child keeps the guard
// bit of a half iff that half of the query is >= the child's (15-bit fields cannot borrow past it).
__device__ __forceinline__ void packChild(const QuantFrame& f, const float* b /*6*/, uint32_t& a, uint32_t& bb, uint32_t& c) {
    const uint32_t C = (uint32_t)kQuantCells;
    a = quantCoord(b[0], f.s0x, f.ivx) | (quantCoord(b[1], f.s0y, f.ivy) << 16);
    bb = quantCoord(b[2], f.s0z, f.ivz) | ((C - quantCoord(b[5], f.s0z, f.ivz)) << 16);
    c = (C - quantCoord(b[3], f.s0x, f.ivx)) | ((C - quantCoord(b[4], f.s0y, f.ivy)) << 16);
}
__device__ __forceinline__ void packQuery(const QuantFrame& f, const float* b /*6*/, uint32_t& a, uint32_t& bb, uint32_t& c) {
    const uint32_t C = (uint32_t)kQuantCells;
    a = (quantCoord(b[3], f.s0x, f.ivx) | (quantCoord(b[4], f.s0y, f.ivy) << 16)) | kGuard;
    bb = (quantCoord(b[5], f.s0z, f.ivz) | ((C - quantCoord(b[2], f.s0z, f.ivz)) << 16)) | kGuard;
    c = ((C - quantCoord(b[0], f.s0x, f.ivx)) | ((C - quantCoord(b[1], f.s0y, f.ivy)) << 16)) | kGuard;
}
__device__ __forceinline__ bool quantIntersect(uint32_t qa, uint32_t qb, uint32_t qc, uint32_t ca, uint32_t cb, uint32_t cc) {
    return (((qa - ca) & (qb - cb) & (qc - cc)) & kGuard) == kGuard;
}

// ---- boxes of aligned leaf ranges (segment tree over the Morton-sorted leaves) --------------------
// Heap layout over P = 2^k >= n leaves: node h covers the leaves of its subtree, leaves live at
// [P, 2P) (slots >= n hold empty boxes), parents of h are h >> 1.  Any sorted range [l, r] is then the
// union of O(log(r-l)) tree nodes, which is how buildTopologyKernel fits the LBVH boxes without
// atomics or fences.  Unions use fminf/fmaxf so a NaN leaf box (which can never intersect anything)
// cannot poison its ancestors.
constexpr int kSegThreads = 256;
constexpr int kSegLeaves = 2 * kSegThreads;   // leaves per block in the bottom pass

__global__ void fillEmptyBoxesKernel(float4* __restrict__ lo, float4* __restrict__ hi, uint32_t count) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const float inf = __int_as_float(0x7f800000);
    lo[k] = make_float4(inf, inf, inf, 0.f);
    hi[k] = make_float4(-inf, -inf, -inf, 0.f);
}

// Reduces `width` consecutive nodes starting at heap index `base` (a whole level or a block's slice
// of it) upward while more than `stopAt` nodes remain, writing every produced level.
__device__ __forceinline__ void segReduceUp(float4* __restrict__ lo, float4* __restrict__ hi, uint32_t base,
                                            uint32_t width, float (*sLo)[3], float (*sHi)[3]) {
    // sLo/sHi hold the current level (width entries) on entry
    uint32_t w = width, b = base;
    while (w > 1) {
        const uint32_t half = w >> 1;
        __syncthreads();
        float l0, l1, l2, h0, h1, h2;
        const uint32_t t = threadIdx.x;
        const bool act = t < half;
        if (act) {
            l0 = fminf(sLo[2 * t][0], sLo[2 * t + 1][0]);
            l1 = fminf(sLo[2 * t][1], sLo[2 * t + 1][1]);
            l2 = fminf(sLo[2 * t][2], sLo[2 * t + 1][2]);
            h0 = fmaxf(sHi[2 * t][0], sHi[2 * t + 1][0]);
            h1 = fmaxf(sHi[2 * t][1], sHi[2 * t + 1][1]);
            h2 = fmaxf(sHi[2 * t][2], sHi[2 * t + 1][2]);
        }
        __syncthreads();
        b >>= 1;
        if (act) {
            sLo[t][0] = l0; sLo[t][1] = l1; sLo[t][2] = l2;
            sHi[t][0] = h0; sHi[t][1] = h1; sHi[t][2] = h2;
            lo[b + t] = make_float4(l0, l1, l2, 0.f);
            hi[b + t] = make_float4(h0, h1, h2, 0.f);
        }
        w = half;
    }
}

// Bottom pass: each block reduces its kSegLeaves leaves up to one node.
__global__ void __launch_bounds__(kSegThreads)
segBuildBottomKernel(float4* __restrict__ lo, float4* __restrict__ hi, uint32_t P) {
    __shared__ float sLo[kSegLeaves][3], sHi[kSegLeaves][3];
    const uint32_t width = min((uint32_t)kSegLeaves, P);
    const uint32_t base = P + blockIdx.x * kSegLeaves;
    for (uint32_t t = threadIdx.x; t < width; t += kSegThreads) {
        const float4 a = lo[base + t], b = hi[base + t];
        sLo[t][0] = a.x; sLo[t][1] = a.y; sLo[t][2] = a.z;
        sHi[t][0] = b.x; sHi[t][1] = b.y; sHi[t][2] = b.z;
    }
    segReduceUp(lo, hi, base, width, sLo, sHi);
}

// The bottom pass of a step, fused with the leaf gather: the block fetches the boxes of the
// bodies at its sorted positions, stores the leaf records the traversal reads, and reduces them upward from shared
// memory — the leaves are not written and read back in between.  Slots >= n keep the empty boxes resizeBodies left.
__global__ void __launch_bounds__(kSegThreads)
segGatherBottomKernel(const float* __restrict__ aabb, const uint32_t* __restrict__ sortedIdx,
                      const uint32_t* __restrict__ sortedKeys, uint32_t n, int worldShift, float4* __restrict__ lo,
                      float4* __restrict__ hi, uint32_t P, uint32_t* __restrict__ zeroPtr, uint32_t zeroWords) {
    __shared__ float sLo[kSegLeaves][3], sHi[kSegLeaves][3];
    // zero tail: the digit histograms the Morton kernel accumulated for the sort that has just finished; the next
    // step's Morton kernel expects them cleared
    for (uint32_t i = blockIdx.x * kSegThreads + threadIdx.x; i < zeroWords; i += gridDim.x * kSegThreads) zeroPtr[i] = 0u;
    const uint32_t width = min((uint32_t)kSegLeaves, P);
    const uint32_t k0 = blockIdx.x * kSegLeaves;
    const uint32_t base = P + k0;
    const float inf = __int_as_float(0x7f800000);
    for (uint32_t t = threadIdx.x; t < width; t += kSegThreads) {
        const uint32_t k = k0 + t;
        float4 a = make_float4(inf, inf, inf, 0.f), b = make_float4(-inf, -inf, -inf, 0.f);
        if (k < n) {
            const uint32_t i = __ldg(sortedIdx + k);
            const float2* r = reinterpret_cast<const float2*>(aabb + (size_t)i * 6);   // 24-byte record: three 8-byte loads
            const float2 b0 = __ldg(r), b1 = __ldg(r + 1), b2 = __ldg(r + 2);
            a = make_float4(b0.x, b0.y, b1.x, __uint_as_float(i));
            b = make_float4(b1.y, b2.x, b2.y, __uint_as_float(__ldg(sortedKeys + k) >> worldShift));
            lo[base + t] = a;
            hi[base + t] = b;
        }
        sLo[t][0] = a.x; sLo[t][1] = a.y; sLo[t][2] = a.z;
        sHi[t][0] = b.x; sHi[t][1] = b.y; sHi[t][2] = b.z;
    }
    segReduceUp(lo, hi, base, width, sLo, sHi);
}

// Top pass (one block): reduces the level holding `count` <= kSegLeaves nodes (heap base = count)
// to the root.
__global__ void __launch_bounds__(kSegThreads)
segBuildTopKernel(float4* __restrict__ lo, float4* __restrict__ hi, uint32_t count) {
    __shared__ float sLo[kSegLeaves][3], sHi[kSegLeaves][3];
    for (uint32_t t = threadIdx.x; t < count; t += kSegThreads) {
        const float4 a = lo[count + t], b = hi[count + t];
        sLo[t][0] = a.x; sLo[t][1] = a.y; sLo[t][2] = a.z;
        sHi[t][0] = b.x; sHi[t][1] = b.y; sHi[t][2] = b.z;
    }
    segReduceUp(lo, hi, count, count, sLo, sHi);
}

// Box of the sorted leaf range [l, r] (inclusive).
__device__ __forceinline__ void segQuery(const float4* __restrict__ lo, const float4* __restrict__ hi, uint32_t P,
                                         uint32_t l, uint32_t r, float* out /*6*/) {
    const float inf = __int_as_float(0x7f800000);
    float a0 = inf, a1 = inf, a2 = inf, b0 = -inf, b1 = -inf, b2 = -inf;
    uint32_t x = l + P, y = r + P + 1;
    while (x < y) {
        if (x & 1u) {
            const float4 p = __ldg(lo + x), q = __ldg(hi + x);
            a0 = fminf(a0, p.x); a1 = fminf(a1, p.y); a2 = fminf(a2, p.z);
            b0 = fmaxf(b0, q.x); b1 = fmaxf(b1, q.y); b2 = fmaxf(b2, q.z);
            ++x;
        }
        if (y & 1u) {
            --y;
            const float4 p = __ldg(lo + y), q = __ldg(hi + y);
            a0 = fminf(a0, p.x); a1 = fminf(a1, p.y); a2 = fminf(a2, p.z);
            b0 = fmaxf(b0, q.x); b1 = fmaxf(b1, q.y); b2 = fmaxf(b2, q.z);
        }
        x >>= 1;
        y >>= 1;
    }
    out[0] = a0; out[1] = a1; out[2] = a2; out[3] = b0; out[4] = b1; out[5] = b2;
}

// One thread per internal node i in [0, n-1): finds its leaf range and split (Karras 2012), then
// fits both children's boxes with two range queries and writes the finished 64-byte node.
// COMPACT = true writes the 32-byte quantised nodes the pair traversal walks; false writes the 64-byte
// float nodes the scene queries use (built on demand by the first query after a broadphase).
template <bool COMPACT>
__global__ void buildTopologyKernel(const uint32_t* __restrict__ keys, uint32_t n, const float4* __restrict__ segLo,
                                    const float4* __restrict__ segHi, uint32_t P, void* __restrict__ nodesOut) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int N = (int)n;
    if (i >= N - 1) return;
    const uint32_t ki = __ldg(keys + i);
    const int d = (karrasDelta(keys, N, i, ki, i + 1) - karrasDelta(keys, N, i, ki, i - 1)) >= 0 ? 1 : -1;
    const int dmin = karrasDelta(keys, N, i, ki, i - d);
    int lmax = 2;
    while (karrasDelta(keys, N, i, ki, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (karrasDelta(keys, N, i, ki, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = karrasDelta(keys, N, i, ki, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (karrasDelta(keys, N, i, ki, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    float L[6], R[6];
    segQuery(segLo, segHi, P, (uint32_t)first, (uint32_t)gamma, L);
    segQuery(segLo, segHi, P, (uint32_t)gamma + 1u, (uint32_t)last, R);
    if (COMPACT) {
        const QuantFrame f = loadQuantFrame(segLo, segHi);
        uint32_t a0, a1, a2, b0, b1, b2;
        packChild(f, L, a0, a1, a2);
        packChild(f, R, b0, b1, b2);
        const uint32_t sp = (uint32_t)gamma | (first == gamma ? kLeftLeaf : 0u) | (gamma + 1 == last ? kRightLeaf : 0u);
        uint4* o = reinterpret_cast<uint4*>(static_cast<Node32*>(nodesOut) + i);
        o[0] = make_uint4(a0, a1, a2, b0);
        o[1] = make_uint4(b1, b2, sp, (uint32_t)last);
    } else {
        float4* o = reinterpret_cast<float4*>(static_cast<BvhNode*>(nodesOut) + i);
        o[0] = make_float4(L[0], L[1], L[2], L[3]);
        o[1] = make_float4(L[4], L[5], R[0], R[1]);
        o[2] = make_float4(R[2], R[3], R[4], R[5]);
        reinterpret_cast<uint4*>(o)[3] = make_uint4((uint32_t)first, (uint32_t)gamma, (uint32_t)last, 0u);
    }
}

// ---- traversal ------------------------------------------------------------------------------------
#ifndef AXCD_TRAV_THREADS
#define AXCD_TRAV_THREADS 128
#endif
#ifndef AXCD_TRAV_POOL
#define AXCD_TRAV_POOL 1024
#endif
constexpr int kTravThreads = AXCD_TRAV_THREADS;
constexpr int kTravPool = AXCD_TRAV_POOL;   // pairs staged per block before the coalesced flush
constexpr int kTravStack = 64;
#ifndef AXCD_TRAV_WIDE
#define AXCD_TRAV_WIDE 1   // 1: two nodes in flight per thread (0: one)
#endif

// AABB::intersects (aabb.hpp:132-135): closed intervals, any NaN -> false
__device__ __forceinline__ bool boxesIntersect(float ax0, float ay0, float az0, float ax1, float ay1, float az1,
                                               float bx0, float by0, float bz0, float bx1, float by1, float bz1) {
    return ax0 <= bx1 && ax1 >= bx0 && ay0 <= by1 && ay1 >= by0 && az0 <= bz1 && az1 >= bz0;
}

// gui::FilterInfo rule (include/axiom/gui/body_inspector.hpp:38-42); records are (category, mask, group, pad)
__device__ __forceinline__ bool shouldCollide(const uint4* __restrict__ filt, uint32_t i, uint32_t j) {
    const uint4 a = __ldg(filt + i), b = __ldg(filt + j);
    const int ga = (int)a.z, gb = (int)b.z;
    if (ga == gb && ga != 0) return ga > 0;
    return (a.y & b.x) != 0u && (a.x & b.y) != 0u;
}

struct SlabRule {
    int enabled;
    float lo, hi;            // this rank's slab [lo, hi) on the x axis
    const uint32_t* keys;    // global id of every local body
};

// Emits each candidate pair once as (a, b) = (min, max) of the two body indices, unordered, into
// `pairs` (staged per block in shared memory, flushed with coalesced stores), and counts the
// pairs of every body a in bodyCount[a] for the counting sort that follows (orderPairs*).
__global__ void __launch_bounds__(kTravThreads)
findPairsKernel(const float4* __restrict__ leafLo, const float4* __restrict__ leafHi,
                const Node32* __restrict__ nodes, const float4* __restrict__ segLo, const float4* __restrict__ segHi,
                const uint32_t* __restrict__ worldEnd, uint32_t n,
                uint2* __restrict__ pairs, uint32_t maxPairs, uint32_t* __restrict__ bodyCount,
                SlabRule slab, const uint4* __restrict__ filters, const uint8_t* __restrict__ awake,
                Counters* __restrict__ ctr) {
    __shared__ uint2 sPool[kTravPool];
    __shared__ uint32_t sCount, sBase;
    if (threadIdx.x == 0) sCount = 0;
    __syncthreads();

    const uint32_t i = blockIdx.x * kTravThreads + threadIdx.x;
    if (i < n && n >= 2) {
        const float4 lo = leafLo[i], hi = leafHi[i];
        const uint32_t bodyI = __float_as_uint(lo.w);
        const uint32_t wEnd = worldEnd ? worldEnd[__float_as_uint(hi.w)] : n - 1;
        uint32_t qxy, qzX, qYZ;
        {
            const QuantFrame f = loadQuantFrame(segLo, segHi);
            const float b[6] = {lo.x, lo.y, lo.z, hi.x, hi.y, hi.z};
            packQuery(f, b, qxy, qzX, qYZ);
        }
#if AXCD_TRAV_WIDE
        // Two nodes in flight per thread: the node carried in registers plus one popped from the stack are
        // loaded together, so two of the dependent L2 round trips of a walk overlap (the order in which
        // nodes are visited does not change the pair set).  Stack entries carry (node, first leaf).
        constexpr uint32_t kNone = 0xffffffffu;
        // With two nodes expanded per trip the stack can hold two pending subtrees per level; the Karras
        // tree is at most 64 levels deep (32 key bits + 32 tie-break bits), so 128 entries cannot fill up.
        // Should that reasoning ever fail, the overflow is reported (505) rather than pairs being dropped.
        constexpr int kWideStack = 2 * kTravStack;
        uint32_t stackN[kWideStack], stackF[kWideStack];
        int sp = 0;
        uint32_t ni = 0, first = 0;
        while (true) {
            const bool haveB = sp > 0;
            uint32_t nb = ni, firstB = first;
            if (haveB) {
                --sp;
                nb = stackN[sp];
                firstB = stackF[sp];
            }
            const uint4* npA = reinterpret_cast<const uint4*>(nodes + ni);
            const uint4* npB = reinterpret_cast<const uint4*>(nodes + nb);
            const uint4 a0 = __ldg(npA), a1 = __ldg(npA + 1);
            const uint4 b0 = __ldg(npB), b1 = __ldg(npB + 1);
            uint32_t next = kNone, nextFirst = 0;
#pragma unroll
            for (int which = 0; which < 2; ++which) {
                if (which == 1 && !haveB) break;
                const uint4 q0 = which ? b0 : a0, q1 = which ? b1 : a1;
                const uint32_t fst = which ? firstB : first;
                const uint32_t split = q1.z & kSplitMask, last = q1.w;
                const bool hitL = split > i && fst <= wEnd && quantIntersect(qxy, qzX, qYZ, q0.x, q0.y, q0.z);
                const bool hitR = last > i && split + 1 <= wEnd && quantIntersect(qxy, qzX, qYZ, q0.w, q1.x, q1.y);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const bool hit = c ? hitR : hitL;
                    if (!hit) continue;
                    const bool leaf = (q1.z & (c ? kRightLeaf : kLeftLeaf)) != 0u;
                    const uint32_t child = c ? split + 1 : split;
                    if (!leaf) {
                        const uint32_t cf = c ? child : fst;
                        if (next == kNone) {
                            next = child;
                            nextFirst = cf;
                        } else if (sp < kWideStack) {
                            stackN[sp] = child;
                            stackF[sp] = cf;
                            ++sp;
                        } else {
                            atomicExch(&ctr->travOverflow, 1u);
                        }
                        continue;
                    }
                    const float4 jl = __ldg(leafLo + child), jh = __ldg(leafHi + child);
                    if (!boxesIntersect(lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, jl.x, jl.y, jl.z, jh.x, jh.y, jh.z)) continue;
                    const uint32_t bodyJ = __float_as_uint(jl.w);
                    uint2 pr = make_uint2(min(bodyI, bodyJ), max(bodyI, bodyJ));
                    if (filters && !shouldCollide(filters, bodyI, bodyJ)) continue;
                    if (awake && !(__ldg(awake + bodyI) | __ldg(awake + bodyJ))) continue;
                    if (slab.enabled) {
                        const float xs = (lo.x > jl.x) ? lo.x : jl.x;
                        if (!(xs >= slab.lo && xs < slab.hi)) continue;
                        if (slab.keys[pr.x] > slab.keys[pr.y]) pr = make_uint2(pr.y, pr.x);
                    }
                    const uint32_t slot = atomicAdd(&sCount, 1u);
                    if (slot < kTravPool) {
                        sPool[slot] = pr;
                    } else {
                        const uint32_t g = atomicAdd(&ctr->pairCount, 1u);
                        if (g < maxPairs) {
                            pairs[g] = pr;
                            atomicAdd(&bodyCount[pr.x], 1u);
                        }
                    }
                }
            }
            if (next != kNone) {
                ni = next;
                first = nextFirst;
            } else if (sp > 0) {
                --sp;
                ni = stackN[sp];
                first = stackF[sp];
            } else {
                break;
            }
        }
    }
#else
        // The node to visit next is carried in registers (index and the first leaf of its range); the
        // stack (local memory) only ever holds right children, whose first leaf equals their index, so a
        // plain descent never round-trips through it.
        constexpr uint32_t kNone = 0xffffffffu;
        uint32_t stack[kTravStack];
        int sp = 0;
        uint32_t ni = 0, first = 0;
        while (true) {
            const uint4* np = reinterpret_cast<const uint4*>(nodes + ni);
            const uint4 q0 = __ldg(np), q1 = __ldg(np + 1);
            const uint32_t split = q1.z & kSplitMask, last = q1.w;
            // left child: sorted leaves [first, split]
            const bool hitL = split > i && first <= wEnd && quantIntersect(qxy, qzX, qYZ, q0.x, q0.y, q0.z);
            // right child: sorted leaves [split+1, last]
            const bool hitR = last > i && split + 1 <= wEnd && quantIntersect(qxy, qzX, qYZ, q0.w, q1.x, q1.y);
            uint32_t next = kNone, nextFirst = 0;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const bool hit = c ? hitR : hitL;
                if (!hit) continue;
                const bool leaf = (q1.z & (c ? kRightLeaf : kLeftLeaf)) != 0u;
                const uint32_t child = c ? split + 1 : split;
                if (!leaf) {
                    if (next == kNone) {
                        next = child;
                        nextFirst = c ? child : first;
                    } else if (sp < kTravStack) {
                        stack[sp++] = child;   // c == 1 here: a right child
                    }
                    continue;
                }
                // exact test on the leaf's own box (AABB::intersects, closed intervals)
                const float4 jl = __ldg(leafLo + child), jh = __ldg(leafHi + child);
                if (!boxesIntersect(lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, jl.x, jl.y, jl.z, jh.x, jh.y, jh.z)) continue;
                const uint32_t bodyJ = __float_as_uint(jl.w);
                uint2 pr = make_uint2(min(bodyI, bodyJ), max(bodyI, bodyJ));
                if (filters && !shouldCollide(filters, bodyI, bodyJ)) continue;
                // sleeping bodies (debug::DebugRigidBody::isAwake, physics_debug_draw.hpp:123): a pair of
                // two sleeping bodies is not a candidate
                if (awake && !(__ldg(awake + bodyI) | __ldg(awake + bodyJ))) continue;
                if (slab.enabled) {
                    // one huge scene split into x-slabs: this rank reports the pair only if the left end
                    // of the pair's x-overlap, max(min_i.x, min_j.x), lies in its slab (exactly one rank
                    // does), and orients it by global id so both bodies play the same role as in a
                    // single-GPU run
                    const float xs = (lo.x > jl.x) ? lo.x : jl.x;
                    if (!(xs >= slab.lo && xs < slab.hi)) continue;
                    if (slab.keys[pr.x] > slab.keys[pr.y]) pr = make_uint2(pr.y, pr.x);
                }
                const uint32_t slot = atomicAdd(&sCount, 1u);
                if (slot < kTravPool) {
                    sPool[slot] = pr;
                } else {   // pool full: append directly
                    const uint32_t g = atomicAdd(&ctr->pairCount, 1u);
                    if (g < maxPairs) {
                        pairs[g] = pr;
                        atomicAdd(&bodyCount[pr.x], 1u);
                    }
                }
            }
            if (next != kNone) {
                ni = next;
                first = nextFirst;
            } else if (sp > 0) {
                ni = stack[--sp];
                first = ni;
            } else {
                break;
            }
        }
    }
#endif
    __syncthreads();
    const uint32_t cnt = min(sCount, (uint32_t)kTravPool);
    if (threadIdx.x == 0 && cnt) sBase = atomicAdd(&ctr->pairCount, cnt);
    __syncthreads();
    if (cnt) {
        const uint32_t base = sBase;
        for (uint32_t k = threadIdx.x; k < cnt; k += kTravThreads)
            if (base + k < maxPairs) {
                const uint2 pr = sPool[k];
                pairs[base + k] = pr;
                atomicAdd(&bodyCount[pr.x], 1u);
            }
    }
}

// ---- traversal with deferred, dense leaf tests (the default) ----------------------------------------------------
// Profile of the kernel above (ncu, headline scene): a third of its instructions belong to the leaf branch — load the
// other leaf's float box, exact AABB::intersects, filter / sleep / slab rules, pool slot — and execute with ~3 of 32
// lanes active, because in any given trip only a few lanes of a warp hit a leaf while the others wait.  Here a lane
// that reaches a leaf only RECORDS the candidate (its own sorted index, the other leaf's) in a per-warp shared-memory
// buffer (one warp-wide prefix sum per trip), and whenever 32 candidates have accumulated the warp tests them
// densely, one per lane, with a ballot-aggregated append to the block's pair pool.  The internal-node walk is the
// same two-nodes-in-flight loop; the candidate set and therefore the pair set are unchanged.
#ifndef AXCD_TRAV_PROLOGUE
#define AXCD_TRAV_PROLOGUE 1
#endif
#ifndef AXCD_TRAV_NODES
#define AXCD_TRAV_NODES 2   // nodes in flight per lane and trip in findPairsDenseKernel
#endif
constexpr int kCandCap = 32 + 32 * 2 * AXCD_TRAV_NODES;   // leftover (< 32) + at most two new candidates per node, lane and trip

__device__ __forceinline__ void testLeafCandidates(uint2 cand, bool valid, const float4* __restrict__ leafLo,
                                                   const float4* __restrict__ leafHi, const SlabRule& slab,
                                                   const uint4* __restrict__ filters, const uint8_t* __restrict__ awake,
                                                   uint2* __restrict__ sPool, uint32_t* __restrict__ sCount,
                                                   uint2* __restrict__ pairs, uint32_t maxPairs,
                                                   uint32_t* __restrict__ bodyCount, Counters* __restrict__ ctr) {
    const int lane = threadIdx.x & 31;
    bool keep = false;
    uint2 pr = make_uint2(0u, 0u);
    if (valid) {
        const float4 il = __ldg(leafLo + cand.x), ih = __ldg(leafHi + cand.x);
        const float4 jl = __ldg(leafLo + cand.y), jh = __ldg(leafHi + cand.y);
        // exact test on the leaves' own boxes (AABB::intersects, closed intervals)
        keep = boxesIntersect(il.x, il.y, il.z, ih.x, ih.y, ih.z, jl.x, jl.y, jl.z, jh.x, jh.y, jh.z);
        const uint32_t bodyI = __float_as_uint(il.w), bodyJ = __float_as_uint(jl.w);
        pr = make_uint2(min(bodyI, bodyJ), max(bodyI, bodyJ));
        if (keep && filters) keep = shouldCollide(filters, bodyI, bodyJ);
        // sleeping bodies (debug::DebugRigidBody::isAwake, physics_debug_draw.hpp:123): a pair of two sleeping
        // bodies is not a candidate
        if (keep && awake) keep = (__ldg(awake + bodyI) | __ldg(awake + bodyJ)) != 0;
        if (keep && slab.enabled) {
            // one huge scene split into x-slabs: this rank reports the pair only if the left end of the pair's
            // x-overlap, max(min_i.x, min_j.x), lies in its slab (exactly one rank does), and orients it by
            // global id so both bodies play the same role as in a single-GPU run
            const float xs = (il.x > jl.x) ? il.x : jl.x;
            keep = xs >= slab.lo && xs < slab.hi;
            if (keep && slab.keys[pr.x] > slab.keys[pr.y]) pr = make_uint2(pr.y, pr.x);
        }
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    if (!bal) return;
    uint32_t base = 0;
    if (lane == __ffs(bal) - 1) base = atomicAdd(sCount, (uint32_t)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
    if (!keep) return;
    const uint32_t slot = base + __popc(bal & ((1u << lane) - 1u));
    if (slot < (uint32_t)kTravPool) {
        sPool[slot] = pr;
    } else {   // pool full: append directly
        const uint32_t g = atomicAdd(&ctr->pairCount, 1u);
        if (g < maxPairs) {
            pairs[g] = pr;
            atomicAdd(&bodyCount[pr.x], 1u);
        }
    }
}

// WORLDS = false (one world): the first leaf of a node's range is only needed for the world cut, so the walk neither
// carries nor stacks it.
template <bool WORLDS>
__global__ void __launch_bounds__(kTravThreads)
findPairsDenseKernel(const float4* __restrict__ leafLo, const float4* __restrict__ leafHi,
                     const Node32* __restrict__ nodes, const float4* __restrict__ segLo, const float4* __restrict__ segHi,
                     const uint32_t* __restrict__ worldEnd, uint32_t n,
                     uint2* __restrict__ pairs, uint32_t maxPairs, uint32_t* __restrict__ bodyCount,
                     SlabRule slab, const uint4* __restrict__ filters, const uint8_t* __restrict__ awake,
                     Counters* __restrict__ ctr) {
    __shared__ uint2 sPool[kTravPool];
    __shared__ uint2 sCand[kTravThreads / 32][kCandCap];
    __shared__ uint32_t sCount, sBase;
    if (threadIdx.x == 0) sCount = 0;
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t i = blockIdx.x * kTravThreads + threadIdx.x;
    constexpr uint32_t kNone = 0xffffffffu;
    constexpr int kWideStack = 2 * kTravStack;
    uint32_t stackN[kWideStack], stackF[WORLDS ? kWideStack : 1];
    int sp = 0;
    uint32_t ni = 0, first = 0;
    bool active = i < n && n >= 2;
    uint32_t wEnd = 0, qxy = 0, qzX = 0, qYZ = 0;
    if (active) {
        const float4 lo = leafLo[i], hi = leafHi[i];
        if (WORLDS) wEnd = worldEnd[__float_as_uint(hi.w)];
        const QuantFrame f = loadQuantFrame(segLo, segHi);
        const float b[6] = {lo.x, lo.y, lo.z, hi.x, hi.y, hi.z};
        packQuery(f, b, qxy, qzX, qYZ);
    }
#if AXCD_TRAV_PROLOGUE
    // ---- warp-uniform descent to the lowest node that holds all of the warp's leaves ------------------------------------
    // The 32 lanes of a warp own consecutive sorted leaves [w0, w1], so the upper part of their walks is the same
    // path: down the spine towards their own leaves.  Walking it once per warp, one node per trip and without the
    // per-lane stack logic, replaces ~10 full trips of the loop below.  At a node whose left child holds the whole warp,
    // the right child lies entirely behind every lane's leaf: each lane tests it against its own box and stacks it on
    // a hit (exactly what the general loop would have done); a left child that lies entirely in front of the warp is
    // pruned for every lane by the "sorted position > i" rule.  The descent stops at the first node whose split
    // separates two of the warp's leaves, or that has a leaf child; the general loop starts there.
    {
        const uint32_t w0 = i - (uint32_t)lane;
        const uint32_t w1 = min(w0 + 31u, n - 1u);
        if (w0 < n && n >= 2u) {
            while (true) {
                uint4 a0, a1;
                loadNode32(nodes + ni, a0, a1);
                const uint32_t split = a1.z & kSplitMask;
                if (a1.z & (kLeftLeaf | kRightLeaf)) break;
                if (w1 <= split) {          // the whole warp lives in the left child
                    if (active && (!WORLDS || split + 1u <= wEnd) && quantIntersect(qxy, qzX, qYZ, a0.w, a1.x, a1.y)) {
                        if (sp < kWideStack) {
                            stackN[sp] = split + 1u;
                            if (WORLDS) stackF[sp] = split + 1u;
                            ++sp;
                        } else {
                            atomicExch(&ctr->travOverflow, 1u);
                        }
                    }
                    ni = split;             // (first is unchanged)
                } else if (w0 > split) {    // ... in the right child: the left one is in front of every lane
                    ni = split + 1u;
                    first = split + 1u;
                } else {
                    break;
                }
            }
        }
    }
#endif
    uint32_t cnt = 0;   // candidates waiting in this warp's buffer (warp-uniform)
    constexpr int K = AXCD_TRAV_NODES;   // nodes in flight per lane and trip: the carried one + up to K - 1 popped
    while (__any_sync(0xffffffffu, active)) {
        uint32_t cand[2 * K];
#pragma unroll
        for (int k = 0; k < 2 * K; ++k) cand[k] = kNone;
        if (active) {
            uint32_t nd[K], fs[K];
            bool on[K];
            nd[0] = ni;
            fs[0] = first;
            on[0] = true;
#pragma unroll
            for (int k = 1; k < K; ++k) {
                on[k] = sp > 0;
                nd[k] = ni;
                fs[k] = first;
                if (on[k]) {
                    --sp;
                    nd[k] = stackN[sp];
                    if (WORLDS) fs[k] = stackF[sp];
                }
            }
            uint4 r0[K], r1[K];
#pragma unroll
            for (int k = 0; k < K; ++k) loadNode32(nodes + nd[k], r0[k], r1[k]);
            uint32_t next = kNone, nextFirst = 0;
#pragma unroll
            for (int which = 0; which < K; ++which) {
                const uint4 q0 = r0[which], q1 = r1[which];
                const uint32_t fst = fs[which];
                const uint32_t split = q1.z & kSplitMask, last = q1.w;
                const bool hitL = on[which] && split > i && (!WORLDS || fst <= wEnd) && quantIntersect(qxy, qzX, qYZ, q0.x, q0.y, q0.z);
                const bool hitR = on[which] && last > i && (!WORLDS || split + 1 <= wEnd) && quantIntersect(qxy, qzX, qYZ, q0.w, q1.x, q1.y);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const bool hit = c ? hitR : hitL;
                    const bool leaf = (q1.z & (c ? kRightLeaf : kLeftLeaf)) != 0u;
                    const uint32_t child = c ? split + 1 : split;
                    if (hit && leaf) cand[which * 2 + c] = child;
                    if (hit && !leaf) {
                        const uint32_t cf = c ? child : fst;
                        if (next == kNone) {
                            next = child;
                            nextFirst = cf;
                        } else if (sp < kWideStack) {
                            stackN[sp] = child;
                            if (WORLDS) stackF[sp] = cf;
                            ++sp;
                        } else {
                            atomicExch(&ctr->travOverflow, 1u);
                        }
                    }
                }
            }
            if (next != kNone) {
                ni = next;
                first = nextFirst;
            } else if (sp > 0) {
                --sp;
                ni = stackN[sp];
                if (WORLDS) first = stackF[sp];
            } else {
                active = false;
            }
        }
        // ---- append this trip's leaf candidates to the warp's buffer: one prefix sum over the lanes' counts --------
        uint32_t mine = 0;
#pragma unroll
        for (int k = 0; k < 2 * K; ++k) mine += (cand[k] != kNone) ? 1u : 0u;
        if (__any_sync(0xffffffffu, mine != 0u)) {
            uint32_t inc = mine;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
                if (lane >= off) inc += t;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
            uint32_t at = cnt + inc - mine;
#pragma unroll
            for (int k = 0; k < 2 * K; ++k)
                if (cand[k] != kNone) sCand[warp][at++] = make_uint2(i, cand[k]);
            cnt += total;
            __syncwarp();
            while (cnt >= 32u) {   // a full warp's worth: test them densely, one candidate per lane
                cnt -= 32u;
                testLeafCandidates(sCand[warp][cnt + lane], true, leafLo, leafHi, slab, filters, awake, sPool, &sCount, pairs,
                                   maxPairs, bodyCount, ctr);
            }
            __syncwarp();
        }
    }
    if (cnt) testLeafCandidates(sCand[warp][min(cnt - 1u, (uint32_t)lane)], (uint32_t)lane < cnt, leafLo, leafHi, slab, filters, awake,
                                sPool, &sCount, pairs, maxPairs, bodyCount, ctr);
    __syncthreads();
    const uint32_t poolCnt = min(sCount, (uint32_t)kTravPool);
    if (threadIdx.x == 0 && poolCnt) sBase = atomicAdd(&ctr->pairCount, poolCnt);
    __syncthreads();
    if (poolCnt) {
        const uint32_t base = sBase;
        for (uint32_t k = threadIdx.x; k < poolCnt; k += kTravThreads)
            if (base + k < maxPairs) {
                const uint2 pr = sPool[k];
                pairs[base + k] = pr;
                atomicAdd(&bodyCount[pr.x], 1u);
            }
    }
}

// ---- canonical pair order by counting sort ---------------------------------------------------------
// bodyStart = exclusive scan of bodyCount (done with exclusiveScanKernel, which also writes the copy
// that serves as the fill cursor).  scatterPairsKernel drops
// every pair's b into its body-a segment (order inside a segment is arbitrary); sortSegmentsCoopKernel
// then sorts each segment and writes the final (a, b) list, which is thereby sorted by (a, b).
#ifndef AXCD_SCATTER_MATCH
#define AXCD_SCATTER_MATCH 1   // 1: one returning atomic per group of equal a's in a warp (pairSort 0.081 -> 0.075 ms)
#endif
__global__ void scatterPairsKernel(const uint2* __restrict__ pairs, const uint32_t* __restrict__ pairCount,
                                   uint32_t maxPairs, uint32_t* __restrict__ bodyCursor,   // starts as a copy of bodyStart
                                   uint32_t* __restrict__ segB) {
    const uint32_t np = min(*pairCount, maxPairs);
#if AXCD_SCATTER_MATCH
    // lanes of a warp that hold pairs of the same body a share one returning atomic (the traversal emits a query's
    // candidates next to each other, so equal a's sit in neighbouring lanes)
    const int lane = threadIdx.x & 31;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t trips = (np + stride - 1u) / stride;
    for (uint32_t t = 0, k = blockIdx.x * blockDim.x + threadIdx.x; t < trips; ++t, k += stride) {
        const bool valid = k < np;
        const uint2 pr = valid ? __ldg(pairs + k) : make_uint2(0xffffffffu, 0u);
        const uint32_t peers = __match_any_sync(0xffffffffu, pr.x);   // (runs of neighbouring lanes by shuffle + ballot: 2 us slower)
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (lane == leader && valid) base = atomicAdd(&bodyCursor[pr.x], (uint32_t)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (valid) segB[base + __popc(peers & ((1u << lane) - 1u))] = pr.y;
    }
#else
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < np; k += gridDim.x * blockDim.x) {
        const uint2 pr = pairs[k];
        const uint32_t pos = atomicAdd(&bodyCursor[pr.x], 1u);
        segB[pos] = pr.y;
    }
#endif
}

constexpr int kSegLocal = 32;

// One body's segment, sorted by one thread.
__device__ __forceinline__ void sortOneSegment(uint32_t a, uint32_t start, uint32_t len, const uint32_t* __restrict__ segB,
                                               uint2* __restrict__ out) {
    if (len == 0) return;
    if (len <= kSegLocal) {
        uint32_t v[kSegLocal];
        for (uint32_t k = 0; k < len; ++k) v[k] = segB[start + k];
        for (uint32_t k = 1; k < len; ++k) {   // insertion sort (segments average ~3 entries)
            const uint32_t x = v[k];
            int j = (int)k - 1;
            while (j >= 0 && v[j] > x) {
                v[j + 1] = v[j];
                --j;
            }
            v[j + 1] = x;
        }
        for (uint32_t k = 0; k < len; ++k) out[start + k] = make_uint2(a, v[k]);
    } else {
        // long segment (a body touching > 32 others): rank each element by counting smaller ones
        for (uint32_t k = 0; k < len; ++k) {
            const uint32_t x = segB[start + k];
            uint32_t r = 0;
            for (uint32_t m = 0; m < len; ++m) r += (segB[start + m] < x) ? 1u : 0u;
            out[start + r] = make_uint2(a, x);   // b values within a segment are distinct
        }
    }
}

// Block-cooperative form of the segment sort.  The segments of 256 consecutive bodies are one contiguous stretch of
// segB (bodyStart is a scan), so the block loads that stretch coalesced into shared memory, marks every element with
// its body, and then works per ELEMENT instead of per body: an element's place in its segment is the number of smaller
// partners in that segment (partners of one body are distinct).  Loads and stores are coalesced, no lane idles on a
// short segment while its neighbour sorts a long one, and nothing lives in local memory.  A stretch that does not fit
// the shared buffer (a very dense neighbourhood) takes the per-body path above.
constexpr int kCoopBodies = 256;
constexpr int kCoopCap = 6144;   // elements per block stretch held in shared memory (256 bodies x 24 partners)

__global__ void __launch_bounds__(kCoopBodies)
sortSegmentsCoopKernel(const uint32_t* __restrict__ bodyStart, const uint32_t* __restrict__ bodyCount, uint32_t n,
                       const uint32_t* __restrict__ segB, uint2* __restrict__ out) {
    __shared__ uint32_t sV[kCoopCap];
    __shared__ uint8_t sOwner[kCoopCap];
    __shared__ uint32_t sStart[kCoopBodies], sLen[kCoopBodies];
    __shared__ uint32_t sEnd;
    const int tid = threadIdx.x;
    const uint32_t a0 = blockIdx.x * kCoopBodies;
    const uint32_t a = a0 + tid;
    const uint32_t start = a < n ? __ldg(bodyStart + a) : 0u;
    const uint32_t len = a < n ? __ldg(bodyCount + a) : 0u;
    sStart[tid] = start;
    sLen[tid] = len;
    const uint32_t lastBody = min(n - a0, (uint32_t)kCoopBodies) - 1u;
    if ((uint32_t)tid == lastBody) sEnd = start + len;
    __syncthreads();
    const uint32_t base = sStart[0];
    const uint32_t total = sEnd - base;
    if (total == 0u) return;
    if (total > (uint32_t)kCoopCap) {
        sortOneSegment(a, start, len, segB, out);   // oversized stretch: one thread per body
        return;
    }
    for (uint32_t e = tid; e < total; e += kCoopBodies) sV[e] = __ldg(segB + base + e);
    for (uint32_t k = 0; k < len; ++k) sOwner[start - base + k] = (uint8_t)tid;
    __syncthreads();
    for (uint32_t e = tid; e < total; e += kCoopBodies) {
        const uint32_t t = sOwner[e];
        const uint32_t s0 = sStart[t] - base, l = sLen[t];
        const uint32_t x = sV[e];
        uint32_t r = 0;
        for (uint32_t m = 0; m < l; ++m) r += (sV[s0 + m] < x) ? 1u : 0u;
        out[base + s0 + r] = make_uint2(a0 + t, x);
    }
}

}  // namespace axcd
