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

struct CksPlanes {
    const uint8_t* src[3];
    size_t pitch[3];
    int w[3], h[3];
};

template <typename T>
__global__ void __launch_bounds__(256) checksum_kernel(CksPlanes P, unsigned long long* __restrict__ out3) {
    // blockIdx.y = plane: the three planes of a frame in one launch
    const uint8_t* __restrict__ src = P.src[blockIdx.y];
    const size_t pitch = P.pitch[blockIdx.y];
    const int w = P.w[blockIdx.y], h = P.h[blockIdx.y];
    unsigned long long* out = out3 + blockIdx.y;
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

// all planes of a frame: one memset of the accumulators + one launch (grid.y = plane)
cudaError_t launch_frame_checksum(const void* const src[3], const size_t pitch[3], const int w[3], const int h[3], int nplanes, int bpc,
                                  uint64_t* out_dev, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(out_dev, 0, 3 * sizeof(uint64_t), s);
    if (e != cudaSuccess) return e;
    {
        static bool carve_done = false;
        if (!carve_done) {
            prefer_max_smem(checksum_kernel<uint8_t>);
            prefer_max_smem(checksum_kernel<uint16_t>);
            carve_done = true;
        }
    }
    CksPlanes P;
    for (int p = 0; p < 3; p++) {
        const int q = p < nplanes ? p : 0;
        P.src[p] = (const uint8_t*)src[q];
        P.pitch[p] = pitch[q];
        P.w[p] = p < nplanes ? w[q] : 0;
        P.h[p] = p < nplanes ? h[q] : 0;
    }
    const int vec = bpc == 8 ? 16 : 8;
    long long items = (long long)((w[0] + vec - 1) / vec) * h[0];
    int blocks = (int)((items + 256 * 4 - 1) / (256 * 4));
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 4) blocks = 148 * 4;
    dim3 grid(blocks, nplanes);
    if (bpc == 8) checksum_kernel<uint8_t><<<grid, 256, 0, s>>>(P, (unsigned long long*)out_dev);
    else checksum_kernel<uint16_t><<<grid, 256, 0, s>>>(P, (unsigned long long*)out_dev);
    return cudaGetLastError();
}

cudaError_t launch_plane_checksum(const void* src, size_t pitch, int w, int h, int bpc, uint64_t* out_dev, cudaStream_t s) {
    // single plane (stage API): same kernel with one plane; only out_dev[0] is touched
    cudaError_t e = cudaMemsetAsync(out_dev, 0, sizeof(uint64_t), s);
    if (e != cudaSuccess) return e;
    CksPlanes P;
    for (int p = 0; p < 3; p++) { P.src[p] = (const uint8_t*)src; P.pitch[p] = pitch; P.w[p] = w; P.h[p] = h; }
    const int vec = bpc == 8 ? 16 : 8;
    long long items = (long long)((w + vec - 1) / vec) * h;
    int blocks = (int)((items + 256 * 4 - 1) / (256 * 4));
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (bpc == 8) checksum_kernel<uint8_t><<<dim3(blocks, 1), 256, 0, s>>>(P, (unsigned long long*)out_dev);
    else checksum_kernel<uint16_t><<<dim3(blocks, 1), 256, 0, s>>>(P, (unsigned long long*)out_dev);
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
