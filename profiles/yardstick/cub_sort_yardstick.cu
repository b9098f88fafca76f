// Same-box yardstick for the hand-written radix sort: CUB DeviceRadixSort::SortPairs on the same
// key distribution (consequence 4).  Measurement tooling only — not part of
// the product library, never loaded by it.   Build: see profiles/yardstick/Makefile
#include <cub/device/device_radix_sort.cuh>
#include <cstdio>
#include <cstdint>

__global__ void fillKeys(uint32_t* keys, uint32_t* vals, uint32_t n, uint32_t keyBits, uint32_t seed) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t x = i * 0x9E3779B9u + seed;
        x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
        keys[i] = keyBits >= 32 ? x : (x & ((1u << keyBits) - 1u));
        vals[i] = i;
    }
}

int main() {
    const uint32_t sizes[4] = {1u << 20, 1u << 22, 1u << 24, 1u << 26};
    for (uint32_t n : sizes) {
        for (int bits : {24, 32}) {
            uint32_t *k0, *k1, *v0, *v1;
            cudaMalloc(&k0, 4ull * n); cudaMalloc(&k1, 4ull * n); cudaMalloc(&v0, 4ull * n); cudaMalloc(&v1, 4ull * n);
            size_t tmpBytes = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, k0, k1, v0, v1, (int)n, 0, bits);
            void* tmp; cudaMalloc(&tmp, tmpBytes);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            float total = 0;
            const int iters = 10;
            for (int it = 0; it <= iters; ++it) {
                fillKeys<<<148 * 8, 256>>>(k0, v0, n, bits, 0x9E3779B9u * (it + 1));
                cudaEventRecord(e0);
                cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, k0, k1, v0, v1, (int)n, 0, bits);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (it) total += ms;
            }
            const int passes = (bits + 7) / 8;
            const double ms = total / iters, bytes = (double)n * (16.0 * passes + 4);
            printf("{\"impl\": \"cub\", \"n\": %u, \"key_bits\": %d, \"ms\": %.4f, \"GBps\": %.1f}\n", n, bits, ms, bytes / (ms * 1e-3) / 1e9);
            cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1); cudaFree(tmp);
        }
    }
    return 0;
}
