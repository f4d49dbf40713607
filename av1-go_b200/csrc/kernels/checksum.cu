// K9' -- device-side 64-bit plane digest, returned per frame instead of copying pixels back.
// digest = sum over visible pixels of (v + 1) * (((y << 16 | x) + 1) * GOLDEN | 1)  (mod 2^64).
// Order independent (a sum), position sensitive.  Streaming read of one plane: F bytes.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../../include/av1r_stages.h"
#include "dev_common.cuh"

namespace av1r {

__host__ __device__ __forceinline__ uint64_t cks_term(uint32_t v, uint32_t x, uint32_t y) {
    uint64_t pos = ((uint64_t)y << 16 | x) + 1;
    uint64_t wgt = (pos * 0x9E3779B97F4A7C15ull) | 1ull;
    return (uint64_t)(v + 1) * wgt;
}

template <typename T>
__global__ void __launch_bounds__(256) checksum_kernel(const uint8_t* __restrict__ src, size_t pitch, int w, int h,
                                                       unsigned long long* __restrict__ out) {
    constexpr int VEC = PixTraits<T>::VEC;
    const int ipr = (w + VEC - 1) / VEC;
    const long long total = (long long)ipr * h;
    uint64_t acc = 0;
    for (long long it = blockIdx.x * 256ll + threadIdx.x; it < total; it += (long long)gridDim.x * 256) {
        const int y = (int)(it / ipr), c = (int)(it - (long long)y * ipr);
        const int x = c * VEC;
        const T* row = (const T*)(src + (size_t)y * pitch);
        if (x + VEC <= w) {
            int px[VEC];
            unpack16(ld_stream128(row + x), px, T());
#pragma unroll
            for (int j = 0; j < VEC; j++) acc += cks_term((uint32_t)px[j], x + j, y);
        } else {
            for (int j = 0; x + j < w; j++) acc += cks_term(row[x + j], x + j, y);
        }
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ uint64_t part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t s = 0;
        for (int i = 0; i < 8; i++) s += part[i];
        atomicAdd(out, (unsigned long long)s);
    }
}

cudaError_t launch_plane_checksum(const void* src, size_t pitch, int w, int h, int bpc, uint64_t* out_dev, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(out_dev, 0, sizeof(uint64_t), s);
    if (e != cudaSuccess) return e;
    {
        static bool carve_done = false;
        if (!carve_done) {
            prefer_max_smem(checksum_kernel<uint8_t>);
            prefer_max_smem(checksum_kernel<uint16_t>);
            carve_done = true;
        }
    }
    const int vec = bpc == 8 ? 16 : 8;
    long long items = (long long)((w + vec - 1) / vec) * h;
    int blocks = (int)((items + 256 * 4 - 1) / (256 * 4));
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (bpc == 8) checksum_kernel<uint8_t><<<blocks, 256, 0, s>>>((const uint8_t*)src, pitch, w, h, (unsigned long long*)out_dev);
    else checksum_kernel<uint16_t><<<blocks, 256, 0, s>>>((const uint8_t*)src, pitch, w, h, (unsigned long long*)out_dev);
    return cudaGetLastError();
}

}  // namespace av1r

extern "C" int av1r_stage_plane_checksum(const void* src, size_t pitch, int w, int h, int bpc, uint64_t* out_dev, void* stream) {
    cudaError_t e = av1r::launch_plane_checksum(src, pitch, w, h, bpc, out_dev, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : -5;
}

extern "C" uint64_t av1r_plane_checksum_host(const void* src, size_t pitch, int w, int h, int bpc) {
    uint64_t acc = 0;
    for (int y = 0; y < h; y++) {
        const uint8_t* row = (const uint8_t*)src + (size_t)y * pitch;
        for (int x = 0; x < w; x++) {
            uint32_t v = bpc == 8 ? row[x] : ((const uint16_t*)row)[x];
            acc += av1r::cks_term(v, x, y);
        }
    }
    return acc;
}
