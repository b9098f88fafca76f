// Hand-written LSD radix sort (8-bit digits), "onesweep" style: one histogram kernel for all
// passes, then one kernel per pass that ranks a 4096-key tile in shared memory, resolves its global
// digit offsets with a decoupled look-back over per-tile status words, and scatters with coalesced
// runs.  Keys are uint32 (+ uint32 payload) for the Morton sort and uint64 (no payload) for the
// candidate-pair sort.  Stable.  Algorithmic HBM bytes per element per pass: read key(+val) and
// write key(+val); plus one key read for the histogram.
#pragma once

#include "axcd_common.cuh"

namespace axcd {

constexpr int kSortThreads = 256;
#ifndef AXCD_SORT_ITEMS
#define AXCD_SORT_ITEMS 16
#endif
constexpr int kSortItems = AXCD_SORT_ITEMS;
#ifndef AXCD_HIST_ITEMS
#define AXCD_HIST_ITEMS 8   // keys per thread of the histogram kernel (sets its grid)
#endif
#ifndef AXCD_SORT_LOOKBACK
#define AXCD_SORT_LOOKBACK 4   // predecessor tiles read per look-back round (sweep 1..32: 4-6 best at 1 M keys)
#endif
#ifndef AXCD_SORT_BALLOT_RANK
#define AXCD_SORT_BALLOT_RANK 1   // 1: peer masks from eight ballots; 0: from one __match_any_sync (slower on B200:
#endif                            //    sort stage 0.081 vs 0.069 ms at 1 M keys, 1.91 vs 1.67 ms at 64 M)
constexpr int kSortTile = kSortThreads * kSortItems;   // 4096 keys per tile
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kRadix = 256;
constexpr int kMaxPasses = 8;


template <typename K>
__device__ __forceinline__ uint32_t digitOf(K key, int shift) {
    return (uint32_t)(key >> shift) & 0xffu;
}

// hist[p*256 + d] += number of keys whose p-th digit is d (all passes in one read of the keys)
template <typename K>
__global__ void __launch_bounds__(kSortThreads)
radixHistogramKernel(const K* __restrict__ keys, uint32_t n, int bitStart, int passes,
                     uint32_t* __restrict__ hist, const uint32_t* __restrict__ enable = nullptr) {
    __shared__ uint32_t sh[kMaxPasses * kRadix];
    if (enable && !*enable) return;   // fallback of the bucket sort: runs only when that one gave up
    for (int i = threadIdx.x; i < passes * kRadix; i += kSortThreads) sh[i] = 0;
    __syncthreads();
    const uint32_t stride = gridDim.x * kSortThreads;
    for (uint32_t i = blockIdx.x * kSortThreads + threadIdx.x; i < n; i += stride) {
        const K k = keys[i];
        for (int p = 0; p < passes; ++p) atomicAdd(&sh[p * kRadix + digitOf(k, bitStart + 8 * p)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * kRadix; i += kSortThreads) {
        const uint32_t v = sh[i];
        if (v) atomicAdd(&hist[i], v);
    }
}

// One radix pass.  `status` holds numTiles*256 words for this pass (zero-initialised); `ticket`
// hands out tile indices in launch order so that a tile's predecessors have always started.
// `digitHist` is this pass's RAW digit histogram over all keys (256 counts): every block turns it into the exclusive
// digit bases itself — the same shuffle tree that scans its own tile's digit counts carries the second value — so no
// scan kernel runs between the histogram and the passes.
template <typename K, bool HAS_VAL>
__global__ void __launch_bounds__(kSortThreads)
radixOnesweepKernel(const K* __restrict__ keysIn, K* __restrict__ keysOut,
                    const uint32_t* __restrict__ valsIn, uint32_t* __restrict__ valsOut,
                    uint32_t n, int shift, const uint32_t* __restrict__ digitHist,
                    volatile uint32_t* __restrict__ status, uint32_t* __restrict__ ticket,
                    const uint32_t* __restrict__ enable = nullptr) {
    __shared__ K sKeys[kSortTile];
    if (enable && !*enable) return;
    __shared__ uint32_t sVals[HAS_VAL ? kSortTile : 1];
    __shared__ uint32_t sWarpHist[kSortWarps][kRadix];
    __shared__ uint32_t sDigitStart[kRadix];
    __shared__ uint32_t sGlobalOff[kRadix];
    __shared__ uint32_t sWarpTot[kSortWarps], sWarpTotH[kSortWarps];
    __shared__ uint32_t sTile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sTile = atomicAdd(ticket, 1u);
    for (int i = tid; i < kSortWarps * kRadix; i += kSortThreads) (&sWarpHist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = sTile;
    const uint32_t tileBase = tile * kSortTile;
    const uint32_t tileCount = min((uint32_t)kSortTile, n - tileBase);
    const uint32_t histRaw = __ldg(digitHist + tid);   // kSortThreads == kRadix: thread d owns digit d

    // ---- load (warp-striped: warp w owns a contiguous 512-key span) and rank ------------------
    K key[kSortItems];
    uint32_t val[kSortItems];
    uint32_t rank[kSortItems];
    const uint32_t warpBase = tileBase + warp * (32 * kSortItems);
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const uint32_t idx = warpBase + k * 32 + lane;
        const bool ok = idx < n;
        key[k] = ok ? keysIn[idx] : (K)~(K)0;   // padding sorts last (digit 255, highest index)
        if (HAS_VAL) val[k] = ok ? valsIn[idx] : 0u;
    }
    const uint32_t lanemaskLt = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const uint32_t d = digitOf(key[k], shift);
#if AXCD_SORT_BALLOT_RANK
        // lanes with the same digit, from eight ballots (one per digit bit) instead of one match operation: ncu showed
        // 35 % of the kernel's stall samples waiting for match results, and issuing the matches ahead changed nothing
        // — the match unit's throughput was the limit, while the issue slots were 83 % idle
        const uint32_t peers = peersByBallot<8>(d);
#else
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
#endif
        const int leader = __ffs(peers) - 1;
        uint32_t prev = 0;
        if (lane == leader) {
            prev = sWarpHist[warp][d];
            sWarpHist[warp][d] = prev + __popc(peers);
        }
        prev = __shfl_sync(0xffffffffu, prev, leader);
        rank[k] = prev + __popc(peers & lanemaskLt);
        __syncwarp();
    }
    __syncthreads();

    // ---- per-digit: exclusive prefix over warps, tile total, look-back -------------------------
    uint32_t digitCount;
    {
        const int d = tid;   // kSortThreads == kRadix
        uint32_t sum = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const uint32_t t = sWarpHist[w][d];
            sWarpHist[w][d] = sum;
            sum += t;
        }
        digitCount = sum;
        uint32_t pub = sum;
        if (d == kRadix - 1) pub -= (kSortTile - tileCount);   // padding keys are not real
        volatile uint32_t* st = status + (size_t)tile * kRadix;
        uint32_t excl = 0;
        if (tile == 0) {
            st[d] = kFlagInclusive | pub;
        } else {
            st[d] = kFlagAggregate | pub;
            // Decoupled look-back, kLookback predecessors per round: the loads of a round are
            // independent (all in flight together), so a deep walk costs one L2 round trip per
            // kLookback tiles instead of one per tile.
            constexpr int kLookback = AXCD_SORT_LOOKBACK;
            int t = (int)tile - 1;
            bool found = false;
            while (!found && t >= 0) {
                uint32_t sv[kLookback];
#pragma unroll
                for (int w = 0; w < kLookback; ++w) {
                    const int idx = t - w;
                    sv[w] = (idx >= 0) ? status[(size_t)idx * kRadix + d] : uint32_t(kFlagInclusive);
                }
#pragma unroll
                for (int w = 0; w < kLookback; ++w) {
                    if (found) continue;
                    uint32_t sw = sv[w];
                    while ((sw & kFlagMask) == 0) sw = status[(size_t)(t - w) * kRadix + d];   // not published yet
                    excl += sw & kValueMask;
                    found = (sw & kFlagMask) == kFlagInclusive;
                }
                t -= kLookback;
            }
            st[d] = kFlagInclusive | (excl + pub);
        }
        // ---- exclusive scans over the digits: the tile's digit counts (position of each digit's run in the tile) and
        //      the global histogram (first output position of each digit) -------------------------------------------------
        uint32_t inc = digitCount, incH = histRaw;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
            const uint32_t tH = __shfl_up_sync(0xffffffffu, incH, off);
            if (lane >= off) {
                inc += t;
                incH += tH;
            }
        }
        if (lane == 31) {
            sWarpTot[warp] = inc;
            sWarpTotH[warp] = incH;
        }
        __syncthreads();
        uint32_t basev = 0, basevH = 0;
#pragma unroll
        for (int i = 0; i < kSortWarps; ++i) {
            basev += (i < warp) ? sWarpTot[i] : 0u;
            basevH += (i < warp) ? sWarpTotH[i] : 0u;
        }
        sDigitStart[tid] = basev + inc - digitCount;
        sGlobalOff[d] = (basevH + incH - histRaw) + excl;
    }
    __syncthreads();

    // ---- local scatter into sorted order --------------------------------------------------------
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const uint32_t d = digitOf(key[k], shift);
        const uint32_t pos = sDigitStart[d] + sWarpHist[warp][d] + rank[k];
        sKeys[pos] = key[k];
        if (HAS_VAL) sVals[pos] = val[k];
    }
    __syncthreads();

    // ---- coalesced global scatter ----------------------------------------------------------------
    for (uint32_t i = tid; i < tileCount; i += kSortThreads) {
        const K kk = sKeys[i];
        const uint32_t d = digitOf(kk, shift);
        const uint32_t dst = sGlobalOff[d] + (i - sDigitStart[d]);
        keysOut[dst] = kk;
        if (HAS_VAL) valsOut[dst] = sVals[i];
    }
}

// Host driver.  Sorts by key bits [bitStart, bitStart + 8*passes).  The sorted data ends in
// (keysA, valsA) if `passes` is even, else in (keysB, valsB); returns which (0 = A, 1 = B).
// scratchHist: kMaxPasses*256 words; scratchStatus: passes * numTiles * 256 words.
template <typename K, bool HAS_VAL>
inline int radixSort(K* keysA, K* keysB, uint32_t* valsA, uint32_t* valsB, uint32_t n, int bitStart,
                     int passes, uint32_t* scratchHist, uint32_t* scratchStatus,
                     uint32_t* tickets /* kMaxPasses words */, cudaStream_t stream, int numSMs = kNumSMs,
                     bool scratchZeroed = false /* the caller already cleared hist / status / tickets */,
                     const uint32_t* enable = nullptr /* device flag: kernels return at once while it is 0 */,
                     bool histReady = false /* scratchHist already holds the raw digit histograms (the Morton kernel's) */) {
    if (n == 0 || passes == 0) return 0;
    const uint32_t numTiles = (n + kSortTile - 1) / kSortTile;
    if (!scratchZeroed) {
        if (!histReady) cudaMemsetAsync(scratchHist, 0, sizeof(uint32_t) * kMaxPasses * kRadix, stream);
        cudaMemsetAsync(scratchStatus, 0, sizeof(uint32_t) * (size_t)passes * numTiles * kRadix, stream);
        cudaMemsetAsync(tickets, 0, sizeof(uint32_t) * kMaxPasses, stream);
    }
    if (!histReady) {
        uint32_t histBlocks = (n + kSortThreads * AXCD_HIST_ITEMS - 1) / (kSortThreads * AXCD_HIST_ITEMS);
        if (histBlocks > (uint32_t)numSMs * 8) histBlocks = numSMs * 8;
        radixHistogramKernel<K><<<histBlocks, kSortThreads, 0, stream>>>(keysA, n, bitStart, passes, scratchHist, enable);
    }
    K* kin = keysA;
    K* kout = keysB;
    uint32_t* vin = valsA;
    uint32_t* vout = valsB;
    for (int p = 0; p < passes; ++p) {
        radixOnesweepKernel<K, HAS_VAL><<<numTiles, kSortThreads, 0, stream>>>(
            kin, kout, vin, vout, n, bitStart + 8 * p, scratchHist + p * kRadix,
            scratchStatus + (size_t)p * numTiles * kRadix, tickets + p, enable);
        K* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    return passes & 1;
}

// ---- bucket sort of the Morton keys (the step's default; the LSD sort above is its fallback) --------------------
// The Morton sort of a step orders (key, body index) pairs whose payload starts out as the identity.  For keys
// that spread over their top bits — bodies spread over the scene — one MSD pass is enough:
//   1. mortonKernel counts the keys per bucket (top `bucketBits` bits; about 256 keys per bucket) while it
//      writes them,
//   2. bucketScanKernel turns the counts into bucket starts and finds the largest bucket,
//   3. bucketScatterKernel drops every (key, index) into its bucket (one returning atomic per key; the order
//      inside a bucket does not matter),
//   4. bucketSortKernel sorts each bucket by (key, index) in shared memory (bitonic, one block per bucket) and
//      writes keys and indices out — the same total order as the stable LSD sort.
// Against 3 LSD passes (16 B read+write per key and pass, plus per-tile look-back chains that serialise when
// all tiles are co-resident) this moves 4 + 8 + 16 B per key in three short kernels.  If any bucket exceeds
// kBucketCap (a clustered scene: most bodies in a few Morton cells) step 2 raises ctr->sortFallback, steps 3-4
// return at once and the LSD kernels — always launched behind them, returning at once otherwise — sort instead.
constexpr int kBucketCap = 1024;       // keys a bucket-sort block takes
constexpr int kBucketThreads = 256;
constexpr int kMaxBucketBits = 16;

struct BucketPlan {
    int bucketBits;   // 0: bucket sort off
    int shift;        // bucket = key >> shift
};
// about 256 keys per bucket, never more buckets than the key has leading bits for
inline BucketPlan bucketPlanFor(uint32_t n, int keyBits) {
    int nb = 1;
    while (nb < 32 && (1ull << nb) < n) ++nb;
    int b = nb - 8;
    if (b > kMaxBucketBits) b = kMaxBucketBits;
    if (b > keyBits) b = keyBits;
    if (b < 1 || n < 4096) return BucketPlan{0, 0};
    return BucketPlan{b, keyBits - b};
}

// One block: exclusive scan of the bucket counts -> starts[0..numBuckets] (and a copy as the scatter cursors),
// largest bucket -> fallback flag; clears the counts for the next step.
__global__ void __launch_bounds__(1024)
bucketScanKernel(uint32_t* __restrict__ counts, uint32_t* __restrict__ starts, uint32_t* __restrict__ cursors,
                 uint32_t numBuckets, uint32_t* __restrict__ fallbackFlag, uint32_t* __restrict__ maxBucketOut) {
    __shared__ uint32_t sWarp[32];
    __shared__ uint32_t sCarry, sMax;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        sCarry = 0;
        sMax = 0;
    }
    __syncthreads();
    uint32_t localMax = 0;
    for (uint32_t base = 0; base < numBuckets; base += 1024) {
        const uint32_t i = base + tid;
        const uint32_t v = (i < numBuckets) ? counts[i] : 0u;
        if (i < numBuckets) counts[i] = 0u;
        localMax = max(localMax, v);
        uint32_t inc = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += t;
        }
        if (lane == 31) sWarp[warp] = inc;
        __syncthreads();
        uint32_t wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 32; ++w) {
            wbase += (w < warp) ? sWarp[w] : 0u;
            total += sWarp[w];
        }
        const uint32_t excl = sCarry + wbase + inc - v;
        if (i < numBuckets) {
            starts[i] = excl;
            cursors[i] = excl;
        }
        __syncthreads();
        if (tid == 0) sCarry += total;
        __syncthreads();
    }
    localMax = __reduce_max_sync(0xffffffffu, localMax);
    if (lane == 0) atomicMax(&sMax, localMax);
    __syncthreads();
    if (tid == 0) {
        starts[numBuckets] = sCarry;
        *maxBucketOut = sMax;
        *fallbackFlag = (sMax > (uint32_t)kBucketCap) ? 1u : 0u;
    }
}

// (key, index) of every element into its bucket; tmp holds uint2 (key, index).
__global__ void __launch_bounds__(256)
bucketScatterKernel(const uint32_t* __restrict__ keys, uint32_t n, int shift, uint32_t* __restrict__ cursors,
                    uint2* __restrict__ tmp, const uint32_t* __restrict__ fallbackFlag) {
    if (*fallbackFlag) return;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const uint32_t k = keys[i];
        const uint32_t pos = atomicAdd(&cursors[k >> shift], 1u);
        tmp[pos] = make_uint2(k, i);
    }
}

// One block per bucket: bitonic sort of (key << 32 | index) in shared memory, padded to the next power of two.
__global__ void __launch_bounds__(kBucketThreads)
bucketSortKernel(const uint2* __restrict__ tmp, const uint32_t* __restrict__ starts, uint32_t numBuckets,
                 uint32_t* __restrict__ keysOut, uint32_t* __restrict__ valsOut, const uint32_t* __restrict__ fallbackFlag) {
    __shared__ unsigned long long sElem[kBucketCap];
    if (*fallbackFlag) return;
    const int tid = threadIdx.x;
    for (uint32_t b = blockIdx.x; b < numBuckets; b += gridDim.x) {
        const uint32_t start = starts[b], count = starts[b + 1] - start;
        if (count == 0) continue;
        uint32_t S = 1;
        while (S < count) S <<= 1;
        for (uint32_t i = tid; i < S; i += kBucketThreads) {
            unsigned long long e = ~0ull;
            if (i < count) {
                const uint2 kv = tmp[start + i];
                e = ((unsigned long long)kv.x << 32) | kv.y;
            }
            sElem[i] = e;
        }
        __syncthreads();
        for (uint32_t k = 2; k <= S; k <<= 1) {
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                for (uint32_t t = tid; t < (S >> 1); t += kBucketThreads) {
                    // t-th compare-exchange of this stage: partner indices (lo, lo | j)
                    const uint32_t lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const uint32_t hi = lo | j;
                    const bool up = (lo & k) == 0;
                    const unsigned long long a = sElem[lo], c = sElem[hi];
                    if ((a > c) == up) {
                        sElem[lo] = c;
                        sElem[hi] = a;
                    }
                }
                __syncthreads();
            }
        }
        for (uint32_t i = tid; i < count; i += kBucketThreads) {
            const unsigned long long e = sElem[i];
            keysOut[start + i] = (uint32_t)(e >> 32);
            valsOut[start + i] = (uint32_t)e;
        }
        __syncthreads();
    }
}

// identity payload + bucket counts for the sort test hooks (what mortonKernel does for a step's keys)
__global__ void iotaCountKernel(const uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t n,
                                uint32_t* __restrict__ bucketCounts, int shift) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        vals[i] = i;
        if (bucketCounts) atomicAdd(bucketCounts + (keys[i] >> shift), 1u);
    }
}

// pseudo-random keys for the sort timing hook (splitmix-style hash of the index)
__global__ void fillRandomKeysKernel(uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t n, uint32_t keyBits,
                                     uint32_t seed) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t x = i * 0x9E3779B9u + seed;
        x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
        keys[i] = keyBits >= 32 ? x : (x & ((1u << keyBits) - 1u));
        vals[i] = i;
    }
}

// ---- single-pass exclusive scan of uint32 (decoupled look-back), used for compaction ------------
#ifndef AXCD_SCAN_THREADS
#define AXCD_SCAN_THREADS 256
#endif
constexpr int kScanThreads = AXCD_SCAN_THREADS;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;
static_assert(kScanItems == 8, "the scan kernel moves 8 words per thread as two uint4");

// out[i] = out2[i] = sum of in[0..i); *total = sum of all.  status: numTiles + 128 words, zero-initialised.
__global__ void __launch_bounds__(kScanThreads)
exclusiveScanKernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t* __restrict__ out2, uint32_t n,
                    volatile uint32_t* __restrict__ status, uint32_t* __restrict__ ticket,
                    uint32_t* __restrict__ total) {
    __shared__ uint32_t sWarp[kScanThreads / 32];
    __shared__ uint32_t sTile, sExcl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sTile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = sTile;
    const uint32_t base = tile * kScanTile + tid * kScanItems;
    uint32_t v[kScanItems];
    uint32_t sum = 0;
    if (base + kScanItems <= n) {   // cudaMalloc'd input, base is a multiple of 8 words: two 16-byte loads
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(in + base)), b = __ldg(reinterpret_cast<const uint4*>(in + base) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) v[k] = (base + k < n) ? in[base + k] : 0u;
    }
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) sum += v[k];
    uint32_t inc = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += t;
    }
    if (lane == 31) sWarp[warp] = inc;
    __syncthreads();
    uint32_t wbase = 0, tileTotal = 0;
#pragma unroll
    for (int i = 0; i < kScanThreads / 32; ++i) {
        wbase += (i < warp) ? sWarp[i] : 0u;
        tileTotal += sWarp[i];
    }
    if (warp == 0) {   // warp-parallel decoupled look-back: 128 predecessor tiles per round (status: +128 words of slack)
        uint32_t excl = 0;
        if (tile == 0) {
            if (lane == 0) status[0] = kFlagInclusive | tileTotal;
        } else {
            if (lane == 0) status[tile] = kFlagAggregate | tileTotal;
            excl = lookbackWide(status, tile, lane);
            if (lane == 0) status[tile] = kFlagInclusive | (excl + tileTotal);
        }
        if (lane == 0) {
            sExcl = excl;
            if ((tile + 1) * (uint64_t)kScanTile >= n) *total = excl + tileTotal;
        }
    }
    __syncthreads();
    uint32_t run = sExcl + wbase + inc - sum;
    if (base + kScanItems <= n) {
        uint32_t o[kScanItems];
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            o[k] = run;
            run += v[k];
        }
        reinterpret_cast<uint4*>(out + base)[0] = make_uint4(o[0], o[1], o[2], o[3]);
        reinterpret_cast<uint4*>(out + base)[1] = make_uint4(o[4], o[5], o[6], o[7]);
        if ((reinterpret_cast<uintptr_t>(out2 + base) & 15u) == 0) {
            reinterpret_cast<uint4*>(out2 + base)[0] = make_uint4(o[0], o[1], o[2], o[3]);
            reinterpret_cast<uint4*>(out2 + base)[1] = make_uint4(o[4], o[5], o[6], o[7]);
        } else {
#pragma unroll
            for (int k = 0; k < kScanItems; ++k) out2[base + k] = o[k];
        }
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            if (base + k < n) {
                out[base + k] = run;
                out2[base + k] = run;
            }
            run += v[k];
        }
    }
}

}  // namespace axcd
