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

// ---- gather leaves into Morton order ------------------------------------------------------------
// leafLo[k] = (min.xyz, bits(bodyIndex)), leafHi[k] = (max.xyz, bits(last sorted index of the
// body's world)) for the body at sorted position k.
__global__ void gatherLeavesKernel(const float* __restrict__ aabb, const uint32_t* __restrict__ sortedIdx,
                                   const uint32_t* __restrict__ sortedKeys, float4* __restrict__ leafLo,
                                   float4* __restrict__ leafHi, uint32_t n, int worldShift) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t i = sortedIdx[k];
    const float* b = aabb + (size_t)i * 6;
    // 24-byte record: three aligned 8-byte loads
    const float2 b0 = __ldg(reinterpret_cast<const float2*>(b));
    const float2 b1 = __ldg(reinterpret_cast<const float2*>(b) + 1);
    const float2 b2 = __ldg(reinterpret_cast<const float2*>(b) + 2);
    leafLo[k] = make_float4(b0.x, b0.y, b1.x, __uint_as_float(i));
    leafHi[k] = make_float4(b1.y, b2.x, b2.y, __uint_as_float(sortedKeys[k] >> worldShift));
}

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

// One thread per internal node i in [0, n-1): finds its leaf range and split, records parents.
// parent[] is indexed by leaf (0..n-1) then internal node (n + i); bit 31 marks "right child".
__global__ void buildTopologyKernel(const uint32_t* __restrict__ keys, uint32_t n,
                                    BvhNode* __restrict__ nodes, uint32_t* __restrict__ parent) {
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
    BvhNode* nd = nodes + i;
    nd->first = (uint32_t)first;
    nd->split = (uint32_t)gamma;
    nd->last = (uint32_t)last;
    nd->pad = 0;
    // left child: leaf gamma if first == gamma else internal gamma
    parent[(first == gamma) ? gamma : N + gamma] = (uint32_t)i;
    parent[(gamma + 1 == last) ? gamma + 1 : N + gamma + 1] = (uint32_t)i | 0x80000000u;
    if (i == 0) parent[N + 0] = 0xffffffffu;   // root
}

// One thread per leaf climbs; the second arrival at a node continues with the union.  Boxes are
// written into the parent's child slots.  Internal unions use fminf/fmaxf so a NaN leaf box (which
// can never intersect anything) cannot poison its ancestors.
__global__ void fitBoxesKernel(const float4* __restrict__ leafLo, const float4* __restrict__ leafHi,
                               uint32_t n, BvhNode* __restrict__ nodes, const uint32_t* __restrict__ parent,
                               uint32_t* __restrict__ visit) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n || n < 2) return;
    const float4 lo = leafLo[k], hi = leafHi[k];
    float bx0 = lo.x, by0 = lo.y, bz0 = lo.z, bx1 = hi.x, by1 = hi.y, bz1 = hi.z;
    uint32_t p = parent[k];
    while (true) {
        const bool right = (p & 0x80000000u) != 0;
        const uint32_t pi = p & 0x7fffffffu;
        volatile float* f = reinterpret_cast<volatile float*>(nodes + pi);
        if (!right) {
            f[0] = bx0; f[1] = by0; f[2] = bz0; f[3] = bx1; f[4] = by1; f[5] = bz1;
        } else {
            f[6] = bx0; f[7] = by0; f[8] = bz0; f[9] = bx1; f[10] = by1; f[11] = bz1;
        }
        __threadfence();
        if (atomicAdd(visit + pi, 1u) == 0u) return;   // first arrival: sibling not done yet
        __threadfence();
        const int o = right ? 0 : 6;                   // sibling's slot
        bx0 = fminf(bx0, f[o + 0]); by0 = fminf(by0, f[o + 1]); bz0 = fminf(bz0, f[o + 2]);
        bx1 = fmaxf(bx1, f[o + 3]); by1 = fmaxf(by1, f[o + 4]); bz1 = fmaxf(bz1, f[o + 5]);
        p = parent[n + pi];
        if (p == 0xffffffffu) return;   // root done
    }
}

// ---- traversal ------------------------------------------------------------------------------------
constexpr int kTravThreads = 128;
constexpr int kTravPool = 2048;    // pairs staged per block before the coalesced flush
constexpr int kTravStack = 64;

// AABB::intersects (aabb.hpp:132-135): closed intervals, any NaN -> false
__device__ __forceinline__ bool boxesIntersect(float ax0, float ay0, float az0, float ax1, float ay1, float az1,
                                               float bx0, float by0, float bz0, float bx1, float by1, float bz1) {
    return ax0 <= bx1 && ax1 >= bx0 && ay0 <= by1 && ay1 >= by0 && az0 <= bz1 && az1 >= bz0;
}

// Packed candidate pair: (min(bodyA,bodyB) << idxBits) | max(...).
__global__ void __launch_bounds__(kTravThreads)
findPairsKernel(const float4* __restrict__ leafLo, const float4* __restrict__ leafHi,
                const BvhNode* __restrict__ nodes, const uint32_t* __restrict__ worldEnd, uint32_t n,
                int idxBits, uint64_t* __restrict__ pairs, uint32_t maxPairs, Counters* __restrict__ ctr) {
    __shared__ uint64_t sPool[kTravPool];
    __shared__ uint32_t sCount, sBase;
    if (threadIdx.x == 0) sCount = 0;
    __syncthreads();

    const uint32_t i = blockIdx.x * kTravThreads + threadIdx.x;
    if (i < n && n >= 2) {
        const float4 lo = leafLo[i], hi = leafHi[i];
        const uint32_t bodyI = __float_as_uint(lo.w);
        const uint32_t wEnd = worldEnd ? worldEnd[__float_as_uint(hi.w)] : n - 1;
        uint32_t stack[kTravStack];
        int sp = 0;
        stack[sp++] = 0;
        while (sp > 0) {
            const uint32_t ni = stack[--sp];
            const float4* np = reinterpret_cast<const float4*>(nodes + ni);
            const float4 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2);
            const uint4 q3 = __ldg(reinterpret_cast<const uint4*>(np) + 3);
            const uint32_t first = q3.x, split = q3.y, last = q3.z;
            // left child: sorted leaves [first, split]
            const bool hitL = split > i && first <= wEnd &&
                              boxesIntersect(lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y);
            // right child: sorted leaves [split+1, last]
            const bool hitR = last > i && split + 1 <= wEnd &&
                              boxesIntersect(lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const bool hit = c ? hitR : hitL;
                if (!hit) continue;
                const bool leaf = c ? (split + 1 == last) : (first == split);
                const uint32_t child = c ? split + 1 : split;
                if (!leaf) {
                    if (sp < kTravStack) stack[sp++] = child;
                    continue;
                }
                const uint32_t bodyJ = __float_as_uint(__ldg(&leafLo[child].w));
                const uint32_t a = min(bodyI, bodyJ), b = max(bodyI, bodyJ);
                const uint64_t pk = ((uint64_t)a << idxBits) | b;
                const uint32_t slot = atomicAdd(&sCount, 1u);
                if (slot < kTravPool) {
                    sPool[slot] = pk;
                } else {   // pool full: append directly
                    const uint32_t g = atomicAdd(&ctr->pairCount, 1u);
                    if (g < maxPairs) pairs[g] = pk;
                }
            }
        }
    }
    __syncthreads();
    const uint32_t cnt = min(sCount, (uint32_t)kTravPool);
    if (threadIdx.x == 0 && cnt) sBase = atomicAdd(&ctr->pairCount, cnt);
    __syncthreads();
    if (cnt) {
        const uint32_t base = sBase;
        for (uint32_t k = threadIdx.x; k < cnt; k += kTravThreads)
            if (base + k < maxPairs) pairs[base + k] = sPool[k];
    }
}

}  // namespace axcd
