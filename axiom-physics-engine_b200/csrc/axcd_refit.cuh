// Stage 1: per-body world-space AABB refit + scene-bounds reduction, and Morton key generation.
//
// HBM layout: transforms stay in the reference's own 40-byte AoS `axiom::math::Transform`
// (position 0, rotation 12, scale 28; include/axiom/math/transform.hpp:18-22) so the upload is one
// cudaMemcpy of the caller's buffer; AABBs are written as the reference's 24-byte `AABB`
// (include/axiom/math/aabb.hpp:18-21).  Neither record is 16-byte aligned per element, so each
// 256-body block moves its 10,240 B of transforms / 6,144 B of AABBs with coalesced 128-bit
// accesses through a shared-memory stage, and threads read their record from shared memory.
// Algorithmic bytes: 40 (Transform) + 16 (shape) in, 24 (AABB) out = 80 B/body.
#pragma once

#include "axcd_common.cuh"

namespace axcd {

#ifndef AXCD_REFIT_THREADS
#define AXCD_REFIT_THREADS 256
#endif
constexpr int kRefitThreads = AXCD_REFIT_THREADS;

// Transform::transformPoint (reference: src/math/transform.cpp:86-93)
__device__ __forceinline__ V3 transformPoint(V3 pos, float4 q, V3 scale, V3 p) {
    V3 scaled = mk3(p.x * scale.x, p.y * scale.y, p.z * scale.z);
    V3 rotated = quatRotate(q, scaled);
    return rotated + pos;
}

// AABB::expand(Vec3) (reference: include/axiom/math/aabb.hpp:143-150), NaN-ignoring selects
__device__ __forceinline__ void expandPoint(V3& lo, V3& hi, V3 p) {
    lo.x = (p.x < lo.x) ? p.x : lo.x;
    lo.y = (p.y < lo.y) ? p.y : lo.y;
    lo.z = (p.z < lo.z) ? p.z : lo.z;
    hi.x = (p.x > hi.x) ? p.x : hi.x;
    hi.y = (p.y > hi.y) ? p.y : hi.y;
    hi.z = (p.z > hi.z) ? p.z : hi.z;
}

// The alternative box route: AABB(-h, h).transform(Transform::toMatrix()), reference:
// src/math/aabb.cpp:8-35 over src/math/transform.cpp:13-24.  M = T * R * S has mat3_cast(q) with column j scaled
// by scale_j as its upper-left 3x3 and the position as its last column (products with the identity's zeros and
// ones are exact); the 8 corners of the LOCAL box in aabb.cpp's order go through glm's mat4 * vec4(c, 1) =
// (m[0]*x + m[1]*y) + (m[2]*z + m[3]), each product-sum one fused operation; AABB(Vec3) then expand().
__device__ __forceinline__ void fitBoxMat4(V3 pos, float4 q, V3 scl, float hx, float hy, float hz, V3& lo, V3& hi) {
    const float qxx = q.x * q.x, qyy = q.y * q.y, qzz = q.z * q.z;
    const float qxz = q.x * q.z, qxy = q.x * q.y, qyz = q.y * q.z;
    const float qwx = q.w * q.x, qwy = q.w * q.y, qwz = q.w * q.z;
    const V3 c0 = mk3(1.0f - 2.0f * (qyy + qzz), 2.0f * (qxy + qwz), 2.0f * (qxz - qwy)) * scl.x;
    const V3 c1 = mk3(2.0f * (qxy - qwz), 1.0f - 2.0f * (qxx + qzz), 2.0f * (qyz + qwx)) * scl.y;
    const V3 c2 = mk3(2.0f * (qxz + qwy), 2.0f * (qyz - qwx), 1.0f - 2.0f * (qxx + qyy)) * scl.z;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float cx = (k & 1) ? hx : -hx, cy = (k & 2) ? hy : -hy, cz = (k & 4) ? hz : -hz;
        const V3 p = mk3(fmaf(c1.x, cy, c0.x * cx) + fmaf(c2.x, cz, pos.x), fmaf(c1.y, cy, c0.y * cx) + fmaf(c2.y, cz, pos.y),
                         fmaf(c1.z, cy, c0.z * cx) + fmaf(c2.z, cz, pos.z));
        if (k == 0) {
            lo = p;
            hi = p;
        } else {
            expandPoint(lo, hi, p);
        }
    }
}

// Tight world-space box of one body (no margin): the per-shape refit rules of SURVEY.md A.2.
template <bool MAT4 = false>
__device__ __forceinline__ void fitBody(V3 pos, float4 q, V3 scl, uint4 sh, const float4* __restrict__ hull,
                                        V3& lo, V3& hi) {
    const float p0 = __uint_as_float(sh.y), p1 = __uint_as_float(sh.z), p2 = __uint_as_float(sh.w);
    if (MAT4 && sh.x == AXCD_SHAPE_BOX) {
        fitBoxMat4(pos, q, scl, p0, p1, p2, lo, hi);
        return;
    }
    if (sh.x == AXCD_SHAPE_SPHERE) {
        // AABB::fromCenterExtents(position, Vec3(r)) (aabb.hpp:213-215); rotation and scale
        // ignored as in the reference's sphere placement (src/debug/physics_debug_draw.cpp:246-248)
        lo = pos - mk3(p0, p0, p0);
        hi = pos + mk3(p0, p0, p0);
    } else if (sh.x == AXCD_SHAPE_BOX) {
        // 8 corners in the order of src/debug/debug_draw.cpp:99-108, AABB(Vec3) then expand().
        // The corners come in antipodal pairs +-c, and every step of transformPoint before the
        // final "+ position" is odd under round-to-nearest: (-c)*scale == -(c*scale) and
        // quatRotate(q, -v) == -quatRotate(q, v) bit for bit (products and fmaf negate exactly).
        // So 4 rotations r_k give all 8 corners as pos +- r_k, and because fl(pos + r) is
        // monotone in r the select-based min/max over the 8 corners equals pos -+ max_k |r_k|.
        // Exceptions, sent down the literal 8-corner path: a zero position component (the sign
        // of an exact zero sum, -0 + -0, depends on the sign of a zero r) and non-finite data
        // (NaN corners are skipped by expand(), which is order-dependent).
        const V3 sc = mk3(p0 * scl.x, p1 * scl.y, p2 * scl.z);   // corner 6 = (+,+,+), scaled
        const V3 r0 = quatRotate(q, sc);                               // corner 6 (-> 0)
        const V3 r1 = quatRotate(q, mk3(sc.x, -sc.y, -sc.z));           // corner 1 (-> 7)
        const V3 r2 = quatRotate(q, mk3(sc.x, -sc.y, sc.z));            // corner 2 (-> 4)
        const V3 r3 = quatRotate(q, mk3(sc.x, sc.y, -sc.z));            // corner 5 (-> 3)
        const V3 m = mk3(fmaxf(fmaxf(fabsf(r0.x), fabsf(r1.x)), fmaxf(fabsf(r2.x), fabsf(r3.x))),
                         fmaxf(fmaxf(fabsf(r0.y), fabsf(r1.y)), fmaxf(fabsf(r2.y), fabsf(r3.y))),
                         fmaxf(fmaxf(fabsf(r0.z), fabsf(r1.z)), fmaxf(fabsf(r2.z), fabsf(r3.z))));
        // finite check on everything the result depends on; NaN anywhere makes the sum NaN
        // (fmaxf would drop a NaN operand, so the r_k are summed, not m)
        const float chk = (fabsf(r0.x) + fabsf(r0.y) + fabsf(r0.z)) + (fabsf(r1.x) + fabsf(r1.y) + fabsf(r1.z)) +
                          (fabsf(r2.x) + fabsf(r2.y) + fabsf(r2.z)) + (fabsf(r3.x) + fabsf(r3.y) + fabsf(r3.z)) +
                          (fabsf(pos.x) + fabsf(pos.y) + fabsf(pos.z));
        if (chk < 3.0e38f && pos.x != 0.0f && pos.y != 0.0f && pos.z != 0.0f) {
            lo = pos - m;
            hi = pos + m;
        } else {
            // corner k of debug_draw.cpp:99-108: x sign + for k in {1,2,5,6}, y sign + for k >= 4,
            // z sign + for k in {2,3,6,7}
            V3 c0 = transformPoint(pos, q, scl, mk3(-p0, -p1, -p2));
            lo = c0;
            hi = c0;
#pragma unroll 1
            for (int k = 1; k < 8; ++k) {
                const float cx = ((k ^ (k >> 1)) & 1) ? p0 : -p0;
                const float cy = (k & 4) ? p1 : -p1;
                const float cz = (k & 2) ? p2 : -p2;
                expandPoint(lo, hi, transformPoint(pos, q, scl, mk3(cx, cy, cz)));
            }
        }
    } else if (sh.x == AXCD_SHAPE_CAPSULE) {
        // p0 = radius, p1 = height: box of the two end spheres, endpoints placed as the reference
        // does (src/debug/physics_debug_draw.cpp:254-266: local (0, -+height/2, 0) through
        // transformPoint); the radius is not scaled
        const V3 half = mk3(0.0f, p1 * 0.5f, 0.0f);
        const V3 start = transformPoint(pos, q, scl, -half);
        const V3 end = transformPoint(pos, q, scl, half);
        lo = start;
        hi = start;
        expandPoint(lo, hi, end);
        lo = lo - mk3(p0, p0, p0);
        hi = hi + mk3(p0, p0, p0);
    } else if (sh.x == AXCD_SHAPE_CYLINDER) {
        // p0 = radius, p1 = height, local Y axis.  Exact box of the scaled cylinder: along world axis k the half
        // extent is the support distance h/2 * |l.y| + r * |(l.x, l.z)| of the local direction l = scale * (R^T e_k)
        const float qxx = q.x * q.x, qyy = q.y * q.y, qzz = q.z * q.z;
        const float qxz = q.x * q.z, qxy = q.x * q.y, qyz = q.y * q.z;
        const float qwx = q.w * q.x, qwy = q.w * q.y, qwz = q.w * q.z;
        const float cx[3] = {1.0f - 2.0f * (qyy + qzz), 2.0f * (qxy + qwz), 2.0f * (qxz - qwy)};
        const float cy[3] = {2.0f * (qxy - qwz), 1.0f - 2.0f * (qxx + qzz), 2.0f * (qyz + qwx)};
        const float cz[3] = {2.0f * (qxz + qwy), 2.0f * (qyz - qwx), 1.0f - 2.0f * (qxx + qyy)};
        float ext[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float lx = cx[k] * scl.x, ly = cy[k] * scl.y, lz = cz[k] * scl.z;
            ext[k] = (p1 * 0.5f) * fabsf(ly) + p0 * sqrtf(lx * lx + lz * lz);
        }
        lo = pos - mk3(ext[0], ext[1], ext[2]);
        hi = pos + mk3(ext[0], ext[1], ext[2]);
    } else {   // AXCD_SHAPE_CONVEX (validated on the host): min/max over transformPoint(v_i)
        const uint32_t first = sh.y, count = sh.z;
        float4 v = __ldg(hull + first);
        V3 c0 = transformPoint(pos, q, scl, mk3(v.x, v.y, v.z));
        lo = c0;
        hi = c0;
        for (uint32_t k = 1; k < count; ++k) {
            v = __ldg(hull + first + k);
            expandPoint(lo, hi, transformPoint(pos, q, scl, mk3(v.x, v.y, v.z)));
        }
    }
}

// COHERENT = false: every body gets its tight box expanded by `margin` (AABB::expand(float)).
// COHERENT = true (temporal coherence): aabb4 holds persistent FAT boxes.  A body
// whose tight box still lies inside its fat box keeps it; otherwise (or when `force` is set: first step
// after the shapes changed) the fat box becomes the tight box expanded by `margin` and the body is
// counted in ctr->movedBodies.  If nothing moved the candidate-pair set cannot have changed and the host
// skips the broadphase (axcd_broadphase).
template <bool COHERENT, bool MAT4 = false>
__global__ void __launch_bounds__(kRefitThreads)
refitKernel(const float4* __restrict__ xf4,      // n*40 bytes viewed as float4 (base 16B aligned)
            const uint4* __restrict__ shapes,    // AxcdShape as uint4
            const float4* __restrict__ hull,     // hull vertices padded to float4
            float4* __restrict__ aabb4,          // n*24 bytes viewed as float4
            uint8_t* __restrict__ type8,         // out: shape type per body, compact (pair classification gathers it)
            uint32_t n, float margin, uint32_t force, Counters* __restrict__ ctr,
            Counters* __restrict__ ctrNext,      // the next step's counter block: reset here (nobody reads it before)
            const Counters* __restrict__ ctrInit) {
    __shared__ __align__(16) float sIn[kRefitThreads * 10];
    __shared__ __align__(16) float sOut[kRefitThreads * 6];
    __shared__ uint32_t sRed[6][kRefitThreads / 32];

    const uint32_t base = blockIdx.x * kRefitThreads;
    const uint32_t cnt = min((uint32_t)kRefitThreads, n - base);
    const int tid = threadIdx.x;
    static_assert(sizeof(Counters) / 4 <= kRefitThreads, "one thread per counter word");
    if (blockIdx.x == 0 && tid < (int)(sizeof(Counters) / 4))
        reinterpret_cast<uint32_t*>(ctrNext)[tid] = reinterpret_cast<const uint32_t*>(ctrInit)[tid];

    if (COHERENT && !force) {   // stage the block's current fat boxes into sOut (same path as the stores below)
        const uint32_t nFloats = cnt * 6, nVec = nFloats / 4;
        const float4* src = aabb4 + (size_t)base * 6 / 4;
        float4* dst = reinterpret_cast<float4*>(sOut);
        for (uint32_t i = tid; i < nVec; i += kRefitThreads) dst[i] = src[i];
        const float* srcF = reinterpret_cast<const float*>(src);
        for (uint32_t i = nVec * 4 + tid; i < nFloats; i += kRefitThreads) sOut[i] = srcF[i];
    }

    // ---- stage the block's transforms: coalesced 128-bit loads --------------------------------
    {
        const uint32_t nFloats = cnt * 10;
        const uint32_t nVec = nFloats / 4;                 // whole float4s
        const float4* src = xf4 + (size_t)base * 10 / 4;   // base*40 B is a multiple of 16 B
        float4* dst = reinterpret_cast<float4*>(sIn);
        for (uint32_t i = tid; i < nVec; i += kRefitThreads) dst[i] = __ldg(src + i);
        const float* srcF = reinterpret_cast<const float*>(src);
        for (uint32_t i = nVec * 4 + tid; i < nFloats; i += kRefitThreads) sIn[i] = __ldg(srcF + i);
    }
    __syncthreads();

    V3 lo = mk3(0.f, 0.f, 0.f), hi = lo;
    bool valid = tid < (int)cnt;
    if (valid) {
        const float* t = sIn + tid * 10;
        const V3 pos = mk3(t[0], t[1], t[2]);
        const float4 q = make_float4(t[3], t[4], t[5], t[6]);
        const V3 scl = mk3(t[7], t[8], t[9]);
        const uint4 sh = __ldg(shapes + base + tid);
        type8[base + tid] = (uint8_t)sh.x;
        fitBody<MAT4>(pos, q, scl, sh, hull, lo, hi);
        float* o = sOut + tid * 6;
        bool keep = false;
        if (COHERENT && !force) {
            // contained (closed): keep the fat box.  Any NaN compares false -> refit.
            keep = o[0] <= lo.x && o[1] <= lo.y && o[2] <= lo.z && lo.x <= hi.x && lo.y <= hi.y && lo.z <= hi.z &&
                   hi.x <= o[3] && hi.y <= o[4] && hi.z <= o[5];
        }
        if (keep) {
            lo = mk3(o[0], o[1], o[2]);
            hi = mk3(o[3], o[4], o[5]);
        } else {
            if (margin != 0.0f) {   // AABB::expand(float) (aabb.hpp:156-160)
                lo = lo - mk3(margin, margin, margin);
                hi = hi + mk3(margin, margin, margin);
            }
            o[0] = lo.x; o[1] = lo.y; o[2] = lo.z;
            o[3] = hi.x; o[4] = hi.y; o[5] = hi.z;
        }
        if (COHERENT) {
            const uint32_t bal = __ballot_sync(__activemask(), !keep);
            if (!keep && (tid & 31) == __ffs(bal) - 1) atomicAdd(&ctr->movedBodies, (uint32_t)__popc(bal));
        }
    }

    // ---- scene bounds of the AABB centres (for Morton normalisation; quality only) ------------
    {
        // AABB::center() (aabb.hpp:62)
        V3 c = (lo + hi) * 0.5f;
        const float big = 3.0e38f;
        bool fin = valid && fabsf(c.x) < big && fabsf(c.y) < big && fabsf(c.z) < big;   // false for NaN
        // ordered-uint encodings reduce with one redux.sync each (integer min/max on the warp)
        uint32_t v[6] = {floatToOrdered(fin ? c.x : big), floatToOrdered(fin ? c.y : big), floatToOrdered(fin ? c.z : big),
                         floatToOrdered(fin ? c.x : -big), floatToOrdered(fin ? c.y : -big), floatToOrdered(fin ? c.z : -big)};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            v[k] = __reduce_min_sync(0xffffffffu, v[k]);
            v[k + 3] = __reduce_max_sync(0xffffffffu, v[k + 3]);
        }
        if ((tid & 31) == 0) {
#pragma unroll
            for (int k = 0; k < 6; ++k) sRed[k][tid >> 5] = v[k];
        }
    }
    __syncthreads();
    if (tid < 6) {
        uint32_t r = sRed[tid][0];
        for (int w = 1; w < kRefitThreads / 32; ++w)
            r = (tid < 3) ? min(r, sRed[tid][w]) : max(r, sRed[tid][w]);
        if (tid < 3) {
            if (r < floatToOrdered(3.0e38f)) atomicMin(&ctr->boundsMin[tid], r);
        } else {
            if (r > floatToOrdered(-3.0e38f)) atomicMax(&ctr->boundsMax[tid - 3], r);
        }
    }

    // ---- write the block's AABBs: coalesced 128-bit stores -----------------------------------
    {
        const uint32_t nFloats = cnt * 6;
        const uint32_t nVec = nFloats / 4;
        float4* dst = aabb4 + (size_t)base * 6 / 4;   // base*24 B is a multiple of 16 B
        const float4* src = reinterpret_cast<const float4*>(sOut);
        for (uint32_t i = tid; i < nVec; i += kRefitThreads) dst[i] = src[i];
        float* dstF = reinterpret_cast<float*>(dst);
        for (uint32_t i = nVec * 4 + tid; i < nFloats; i += kRefitThreads) dstF[i] = sOut[i];
    }
}

// ---- TMA-pipelined refit (the default, non-coherent path) ------------------------------------------------
// Persistent blocks walk the 256-body tiles of the scene.  One elected thread moves each tile's 10,240 B
// of transforms global -> shared with a bulk async copy (cp.async.bulk, completion on an mbarrier) one
// tile AHEAD of the compute, and sends each finished 6,144 B AABB tile shared -> global with a bulk
// store, so the LSU instruction stream of the old staging loops disappears and loads, math and stores of
// consecutive tiles overlap inside one block.  Scene bounds accumulate in registers across tiles (one
// block reduction at the end).  The last, partial tile (and any tile whose byte count is not a multiple
// of 16) goes through plain loads.
#ifndef AXCD_REFIT_TMA_BLOCKS
#define AXCD_REFIT_TMA_BLOCKS 5   // resident blocks per SM the grid is sized for
#endif
constexpr int kRefitTmaBlocksPerSM = AXCD_REFIT_TMA_BLOCKS;
#ifndef AXCD_REFIT_STAGES
#define AXCD_REFIT_STAGES 2       // input tiles in flight per block (prefetch distance = stages - 1)
#endif
constexpr int kRefitStages = AXCD_REFIT_STAGES;

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smemAddr(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulkLoad(void* dstShared, const void* srcGlobal, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smemAddr(dstShared)), "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar))
                 : "memory");
}
__device__ __forceinline__ void bulkStore(void* dstGlobal, const void* srcShared, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dstGlobal), "r"(smemAddr(srcShared)),
                 "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(kRefitThreads, kRefitTmaBlocksPerSM)
refitTmaKernel(const float* __restrict__ xf,        // n*10 floats (base 16B aligned)
               const uint4* __restrict__ shapes, const float4* __restrict__ hull,
               float* __restrict__ aabb,            // n*6 floats (base 16B aligned)
               uint8_t* __restrict__ type8, uint32_t n, float margin, Counters* __restrict__ ctr,
               Counters* __restrict__ ctrNext, const Counters* __restrict__ ctrInit) {
    __shared__ __align__(128) float sIn[kRefitStages][kRefitThreads * 10];
    __shared__ __align__(128) float sOut[2][kRefitThreads * 6];
    __shared__ __align__(8) uint64_t sBar[kRefitStages];
    __shared__ uint32_t sRed[6][kRefitThreads / 32];
    constexpr uint32_t kInBytes = kRefitThreads * 40, kOutBytes = kRefitThreads * 24;
    static_assert(kInBytes % 16 == 0 && kOutBytes % 16 == 0, "bulk copies move multiples of 16 bytes");

    const int tid = threadIdx.x;
    if (blockIdx.x == 0 && tid < (int)(sizeof(Counters) / 4))
        reinterpret_cast<uint32_t*>(ctrNext)[tid] = reinterpret_cast<const uint32_t*>(ctrInit)[tid];
    const uint32_t fullTiles = n / kRefitThreads;
    if (tid == 0) {
        for (int k = 0; k < kRefitStages; ++k) mbarInit(&sBar[k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const float big = 3.0e38f;
    float bmin[3] = {big, big, big}, bmax[3] = {-big, -big, -big};   // this thread's running bounds of box centres
    auto accumulate = [&](V3 lo, V3 hi) {
        const V3 c = (lo + hi) * 0.5f;   // AABB::center() (aabb.hpp:62)
        if (fabsf(c.x) < big && fabsf(c.y) < big && fabsf(c.z) < big) {   // false for NaN
            bmin[0] = fminf(bmin[0], c.x); bmin[1] = fminf(bmin[1], c.y); bmin[2] = fminf(bmin[2], c.z);
            bmax[0] = fmaxf(bmax[0], c.x); bmax[1] = fmaxf(bmax[1], c.y); bmax[2] = fmaxf(bmax[2], c.z);
        }
    };

    uint32_t tile = blockIdx.x;
    int stage = 0, ostage = 0;
    uint32_t phaseBits = 0;
    if (tid == 0) {   // prologue: the first stages-1 tiles of this block
        for (int k = 0; k < kRefitStages - 1; ++k) {
            const uint32_t t = tile + (uint32_t)k * gridDim.x;
            if (t < fullTiles) {
                mbarExpectTx(&sBar[k], kInBytes);
                bulkLoad(sIn[k], xf + (size_t)t * kRefitThreads * 10, kInBytes, &sBar[k]);
            }
        }
    }
    for (; tile < fullTiles; tile += gridDim.x) {
        // prefetch into the stage consumed in the previous iteration (last read before its second __syncthreads)
        const uint32_t next = tile + (uint32_t)(kRefitStages - 1) * gridDim.x;
        const int pstage = (stage + kRefitStages - 1) % kRefitStages;
        if (tid == 0 && next < fullTiles) {
            mbarExpectTx(&sBar[pstage], kInBytes);
            bulkLoad(sIn[pstage], xf + (size_t)next * kRefitThreads * 10, kInBytes, &sBar[pstage]);
        }
        const uint32_t body = tile * kRefitThreads + tid;
        const uint4 sh = __ldg(shapes + body);
        mbarWait(&sBar[stage], (phaseBits >> stage) & 1u);
        phaseBits ^= 1u << stage;
        const float2* t = reinterpret_cast<const float2*>(sIn[stage] + tid * 10);   // 40-byte record, 8-byte aligned
        const float2 t0 = t[0], t1 = t[1], t2 = t[2], t3 = t[3], t4 = t[4];
        V3 lo, hi;
        fitBody(mk3(t0.x, t0.y, t1.x), make_float4(t1.y, t2.x, t2.y, t3.x), mk3(t3.y, t4.x, t4.y), sh, hull, lo, hi);
        type8[body] = (uint8_t)sh.x;
        if (margin != 0.0f) {   // AABB::expand(float) (aabb.hpp:156-160)
            lo = lo - mk3(margin, margin, margin);
            hi = hi + mk3(margin, margin, margin);
        }
        accumulate(lo, hi);
        // sOut[ostage] was handed to a bulk store two iterations ago: its reads must be finished
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();
        float2* o = reinterpret_cast<float2*>(sOut[ostage] + tid * 6);              // 24-byte record, 8-byte aligned
        o[0] = make_float2(lo.x, lo.y);
        o[1] = make_float2(lo.z, hi.x);
        o[2] = make_float2(hi.y, hi.z);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the bulk store
        __syncthreads();
        if (tid == 0) bulkStore(aabb + (size_t)tile * kRefitThreads * 6, sOut[ostage], kOutBytes);
        stage = (stage + 1) % kRefitStages;
        ostage ^= 1;
    }
    // ---- the partial last tile: plain loads and stores, one block -----------------------------------------------
    if (blockIdx.x == fullTiles % gridDim.x) {
        const uint32_t body = fullTiles * kRefitThreads + tid;
        if (body < n) {
            const float2* t = reinterpret_cast<const float2*>(xf + (size_t)body * 10);
            const float2 t0 = __ldg(t), t1 = __ldg(t + 1), t2 = __ldg(t + 2), t3 = __ldg(t + 3), t4 = __ldg(t + 4);
            const uint4 sh = __ldg(shapes + body);
            V3 lo, hi;
            fitBody(mk3(t0.x, t0.y, t1.x), make_float4(t1.y, t2.x, t2.y, t3.x), mk3(t3.y, t4.x, t4.y), sh, hull, lo, hi);
            type8[body] = (uint8_t)sh.x;
            if (margin != 0.0f) {
                lo = lo - mk3(margin, margin, margin);
                hi = hi + mk3(margin, margin, margin);
            }
            accumulate(lo, hi);
            float2* o = reinterpret_cast<float2*>(aabb + (size_t)body * 6);
            o[0] = make_float2(lo.x, lo.y);
            o[1] = make_float2(lo.z, hi.x);
            o[2] = make_float2(hi.y, hi.z);
        }
    }
    // ---- scene bounds: one block reduction, one atomic per component --------------------------------------------
    {
        uint32_t v[6] = {floatToOrdered(bmin[0]), floatToOrdered(bmin[1]), floatToOrdered(bmin[2]),
                         floatToOrdered(bmax[0]), floatToOrdered(bmax[1]), floatToOrdered(bmax[2])};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            v[k] = __reduce_min_sync(0xffffffffu, v[k]);
            v[k + 3] = __reduce_max_sync(0xffffffffu, v[k + 3]);
        }
        if ((tid & 31) == 0) {
#pragma unroll
            for (int k = 0; k < 6; ++k) sRed[k][tid >> 5] = v[k];
        }
    }
    __syncthreads();
    if (tid < 6) {
        uint32_t r = sRed[tid][0];
        for (int w = 1; w < kRefitThreads / 32; ++w)
            r = (tid < 3) ? min(r, sRed[tid][w]) : max(r, sRed[tid][w]);
        if (tid < 3) {
            if (r < floatToOrdered(3.0e38f)) atomicMin(&ctr->boundsMin[tid], r);
        } else {
            if (r > floatToOrdered(-3.0e38f)) atomicMax(&ctr->boundsMax[tid - 3], r);
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // shared memory must outlive the stores
}

// ---- Morton keys -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t expandBits10(uint32_t v) {   // 10 bits -> every third bit
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// key = worldId << (3*bitsPerAxis) | morton(centre).  Reads the 24-byte AABBs through a shared
// stage (coalesced 128-bit loads), writes key (4 B) + body index (4 B): 32 B/body.
// Scratch words the later kernels of a step expect zeroed (per-body pair counts, look-back status words of
// the scans and of the radix passes, the radix histograms).  mortonKernel clears them on its way — one
// thread per body is more than enough — so no memset node precedes those kernels.
struct ZeroList {
    uint32_t* ptr[5];
    uint32_t words[5];
};

// The kernel also accumulates the radix sort's digit histograms (one 256-bin histogram per 8-bit digit of the key,
// `histPasses` of them) — per block in shared memory over all of the block's tiles, then one reduction per non-empty
// bin — so the sort needs no histogram pass of its own.  `hist` must be zero on entry (cleared by a later kernel of
// the previous broadphase, see segGatherBottomKernel's zero tail); the grid is a few blocks per SM, each walking
// several 256-body tiles, which keeps the number of global reductions at that of a dedicated histogram kernel.
__global__ void __launch_bounds__(kRefitThreads)
mortonKernel(const float4* __restrict__ aabb4, const uint32_t* __restrict__ worldId,
             uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t n,
             int bitsPerAxis, const Counters* __restrict__ ctr, ZeroList zl,
             uint32_t* __restrict__ bucketCounts /* or nullptr */, int bucketShift,
             uint32_t* __restrict__ hist /* or nullptr */, int histPasses) {
    __shared__ __align__(16) float sIn[kRefitThreads * 6];
    __shared__ uint32_t sHist[4 * 256];
    const int tid = threadIdx.x;
#pragma unroll
    for (int k = 0; k < 5; ++k)
        for (uint32_t i = blockIdx.x * kRefitThreads + tid; i < zl.words[k]; i += gridDim.x * kRefitThreads) zl.ptr[k][i] = 0u;
    if (hist)
        for (int i = tid; i < histPasses * 256; i += kRefitThreads) sHist[i] = 0u;
    float lo3[3], inv3[3];
    const float cells = (float)(1u << bitsPerAxis);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        lo3[k] = orderedToFloat(ctr->boundsMin[k]);
        inv3[k] = orderedToFloat(ctr->boundsMax[k]) - lo3[k];   // extent
    }
    const uint32_t tiles = (n + kRefitThreads - 1) / kRefitThreads;
    for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const uint32_t base = tile * kRefitThreads;
        const uint32_t cnt = min((uint32_t)kRefitThreads, n - base);
        __syncthreads();   // the previous tile's staged boxes have been read (and sHist is cleared, first trip)
        {
            const uint32_t nFloats = cnt * 6, nVec = nFloats / 4;
            const float4* src = aabb4 + (size_t)base * 6 / 4;
            float4* dst = reinterpret_cast<float4*>(sIn);
            for (uint32_t i = tid; i < nVec; i += kRefitThreads) dst[i] = __ldg(src + i);
            const float* srcF = reinterpret_cast<const float*>(src);
            for (uint32_t i = nVec * 4 + tid; i < nFloats; i += kRefitThreads) sIn[i] = __ldg(srcF + i);
        }
        __syncthreads();
        if (tid >= (int)cnt) continue;
        const float* b = sIn + tid * 6;
        uint32_t code = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float lo = lo3[k];
            const float c = (b[k] + b[k + 3]) * 0.5f;
            const float ext = inv3[k];
            float t = (ext > 0.0f) ? (c - lo) / ext : 0.0f;
            t = t * cells;
            // NaN / out-of-range centres clamp into the grid (quality only; never affects the pair set)
            int q = (t >= 0.0f) ? ((t < cells) ? (int)t : (int)cells - 1) : 0;
            q = min(max(q, 0), (int)cells - 1);
            code |= expandBits10((uint32_t)q) << (2 - k);
        }
        if (worldId) code |= __ldg(worldId + base + tid) << (3 * bitsPerAxis);
        keys[base + tid] = code;
        vals[base + tid] = base + tid;
        if (bucketCounts) atomicAdd(bucketCounts + (code >> bucketShift), 1u);   // bucket sort, step 1 (axcd_sort.cuh)
        if (hist)
            for (int p = 0; p < histPasses; ++p) atomicAdd(&sHist[p * 256 + ((code >> (8 * p)) & 0xffu)], 1u);
    }
    if (hist) {
        __syncthreads();
        for (int i = tid; i < histPasses * 256; i += kRefitThreads) {
            const uint32_t v = sHist[i];
            if (v) atomicAdd(&hist[i], v);
        }
    }
}

// ---- x-slab mode: ghost records --------------------------------------------------------------------
// A ghost record is 16 floats (64 B): Transform (10), shape (4 words), global id, pad.
constexpr int kGhostWords = 16;

// Packs every owned body whose closed x-interval [min.x, max.x] meets the half-open slab [lo, hi) of
// a destination rank into that rank's send buffer (warp-aggregated append; order is irrelevant, the
// pair orientation comes from the global ids).
__global__ void packGhostsKernel(const float* __restrict__ aabb, const float* __restrict__ xf,
                                 const uint4* __restrict__ shapes, const uint32_t* __restrict__ keys,
                                 uint32_t nOwned, float lo, float hi, float4* __restrict__ out, uint32_t cap,
                                 uint32_t* __restrict__ counter) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool take = false;
    if (i < nOwned) {
        const float mn = __ldg(aabb + (size_t)i * 6), mx = __ldg(aabb + (size_t)i * 6 + 3);
        take = mn < hi && mx >= lo;
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, take);
    if (!bal) return;
    const int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == __ffs(bal) - 1) base = atomicAdd(counter, (uint32_t)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
    if (!take) return;
    const uint32_t slot = base + __popc(bal & ((1u << lane) - 1u));
    if (slot >= cap) return;   // counted, not stored: the host sees counter > cap
    const float* t = xf + (size_t)i * 10;
    const uint4 sh = __ldg(shapes + i);
    float4* o = out + (size_t)slot * (kGhostWords / 4);
    o[0] = make_float4(t[0], t[1], t[2], t[3]);
    o[1] = make_float4(t[4], t[5], t[6], t[7]);
    o[2] = make_float4(t[8], t[9], __uint_as_float(sh.x), __uint_as_float(sh.y));
    o[3] = make_float4(__uint_as_float(sh.z), __uint_as_float(sh.w), __uint_as_float(__ldg(keys + i)), 0.0f);
}

// The same selection for ALL destination ranks in one pass over the owned bodies: slab r = [edges[r],
// edges[r+1]) for r != myRank; out holds numRanks send buffers of `cap` records each, counters one word per
// rank.  (Slabs are contiguous, so a body usually matches none or one neighbour.)  A convex hull's vertices travel
// too: they are appended to the destination's vertex send buffer (vertOut, vertCap float4 per rank; counters
// numRanks .. 2*numRanks-1) and the record's `first` becomes the offset inside that buffer — the receiver rebases it.
__global__ void packGhostsAllKernel(const float* __restrict__ aabb, const float* __restrict__ xf,
                                    const uint4* __restrict__ shapes, const uint32_t* __restrict__ keys,
                                    const float4* __restrict__ hull, uint32_t nOwned, const float* __restrict__ edges,
                                    uint32_t numRanks, uint32_t myRank, float4* __restrict__ out, uint32_t cap,
                                    float4* __restrict__ vertOut, uint32_t vertCap, uint32_t* __restrict__ counters) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    float mn = 0.0f, mx = 0.0f;
    const bool live = i < nOwned;
    if (live) {
        mn = __ldg(aabb + (size_t)i * 6);
        mx = __ldg(aabb + (size_t)i * 6 + 3);
    }
    for (uint32_t r = 0; r < numRanks; ++r) {
        if (r == myRank) continue;
        const bool take = live && mn < __ldg(edges + r + 1) && mx >= __ldg(edges + r);
        const uint32_t bal = __ballot_sync(0xffffffffu, take);
        if (!bal) continue;
        uint32_t base = 0;
        if (lane == __ffs(bal) - 1) base = atomicAdd(counters + r, (uint32_t)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
        if (!take) continue;
        const uint32_t slot = base + __popc(bal & ((1u << lane) - 1u));
        uint4 sh = __ldg(shapes + i);
        if (sh.x == AXCD_SHAPE_CONVEX) {
            const uint32_t vbase = atomicAdd(counters + numRanks + r, sh.z);   // counted even when it does not fit
            if (slot < cap && vbase + sh.z <= vertCap) {
                float4* vo = vertOut + (size_t)r * vertCap + vbase;
                for (uint32_t k = 0; k < sh.z; ++k) vo[k] = __ldg(hull + sh.y + k);
            }
            sh.y = vbase;
        }
        if (slot >= cap) continue;   // counted, not stored: the host sees counter > cap
        const float* t = xf + (size_t)i * 10;
        float4* o = out + ((size_t)r * cap + slot) * (kGhostWords / 4);
        o[0] = make_float4(t[0], t[1], t[2], t[3]);
        o[1] = make_float4(t[4], t[5], t[6], t[7]);
        o[2] = make_float4(t[8], t[9], __uint_as_float(sh.x), __uint_as_float(sh.y));
        o[3] = make_float4(__uint_as_float(sh.z), __uint_as_float(sh.w), __uint_as_float(__ldg(keys + i)), 0.0f);
    }
}

// Offsets of each source rank's block inside the receive buffers (records and hull vertices).
struct GhostOffsets {
    uint32_t recEnd[64];     // records of source ranks 0..r end here (prefix sums, own rank contributes 0)
    uint32_t vertBase[64];   // index in the hull pool where source rank r's vertices start
    uint32_t numRanks;
};

// Appends received ghost records after the owned bodies.  With `off` (records grouped by source rank) a hull's
// vertex range is rebased onto where that rank's vertices landed in the hull pool.
__global__ void unpackGhostsKernel(const float4* __restrict__ in, uint32_t count, uint32_t first,
                                   float* __restrict__ xf, uint4* __restrict__ shapes, uint32_t* __restrict__ keys,
                                   GhostOffsets off, uint32_t rebase) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= count) return;
    const float4* r = in + (size_t)g * (kGhostWords / 4);
    const float4 a = r[0], b = r[1], c = r[2], d = r[3];
    float* t = xf + (size_t)(first + g) * 10;
    t[0] = a.x; t[1] = a.y; t[2] = a.z; t[3] = a.w;
    t[4] = b.x; t[5] = b.y; t[6] = b.z; t[7] = b.w;
    t[8] = c.x; t[9] = c.y;
    uint4 sh = make_uint4(__float_as_uint(c.z), __float_as_uint(c.w), __float_as_uint(d.x), __float_as_uint(d.y));
    if (rebase && sh.x == AXCD_SHAPE_CONVEX) {
        uint32_t src = 0;
        while (src + 1 < off.numRanks && g >= off.recEnd[src]) ++src;
        sh.y += off.vertBase[src];
    }
    shapes[first + g] = sh;
    keys[first + g] = __float_as_uint(d.z);
}

// axcd_set_poses: position + rotation (7 floats per body, packed) into the 10-float Transform records; the
// scales stay as the last axcd_set_transforms left them.
__global__ void mergePosesKernel(const float* __restrict__ poses, float* __restrict__ xf, uint32_t words) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= words) return;
    const uint32_t body = i / 7u, comp = i - body * 7u;
    xf[(size_t)body * 10 + comp] = poses[i];
}

}  // namespace axcd
