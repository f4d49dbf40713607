// K6 -- super-resolution upscaling (AV1 spec 7.16) for sm_100a.
//
// Frames coded with use_superres are reconstructed, deblocked and CDEF-filtered at the coded (downscaled) width; this kernel
// then stretches every row to the upscaled width with the normative 8-tap filter (positions in 1/16384 sample, 64 filter
// phases), before loop restoration.  One thread per output sample, rows along threadIdx.x: the eight taps of neighbouring
// outputs overlap, so the source row segment of a warp stays in L1 and the output row is written with coalesced stores.
// A CTA handles a 256-sample strip of one row of one plane; blockIdx.z walks the planes.
// Algorithmic bytes: F_down read + F_up written = (d/8 + 1) * F_down for denominator d (9..16).
#include <cuda_runtime.h>
#include <stdint.h>

#include "dev_common.cuh"
#include "devframe.h"
#include "intra.h"
#include "../tables/tables_filter.inc"

namespace av1r {

// (global memory, read as one 128-bit row per sample: neighbouring samples have different phases, and a constant-memory table
// indexed per lane is read one address at a time)
__device__ __align__(16) int16_t d_upscale_filter[64][8];
static bool g_sr_const_loaded[64] = {false};

template <typename T>
__global__ void __launch_bounds__(256) superres_kernel(SuperresLaunch L) {
    const int plane = blockIdx.z;
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    const int up_w = L.up_w[plane];
    if (x >= up_w || y >= L.h[plane]) return;
    const int src_x = -(1 << 14) + L.initial_subpel_x[plane] + x * L.step_x[plane];
    const int src_px = src_x >> 14, sub = (src_x & ((1 << 14) - 1)) >> 8;
    const int max_x = L.src_cw[plane] - 1;
    const T* src = reinterpret_cast<const T*>(L.src.p[plane] + (size_t)y * L.src.pitch[plane]);
    const int4 fq = __ldg(reinterpret_cast<const int4*>(d_upscale_filter[sub]));
    const int f[8] = {(int)(short)(fq.x & 0xffff), fq.x >> 16, (int)(short)(fq.y & 0xffff), fq.y >> 16,
                      (int)(short)(fq.z & 0xffff), fq.z >> 16, (int)(short)(fq.w & 0xffff), fq.w >> 16};
    int sum = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) sum += (int)src[min(max(src_px + k - 3, 0), max_x)] * f[k];
    const int pixmax = (1 << L.bd) - 1;
    reinterpret_cast<T*>(L.dst.p[plane] + (size_t)y * L.dst.pitch[plane])[x] = (T)min(max((sum + 64) >> 7, 0), pixmax);
}

cudaError_t launch_superres(const SuperresLaunch& L, cudaStream_t s) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (!(dev < 64 && g_sr_const_loaded[dev])) {
        if ((e = cudaMemcpyToSymbol(d_upscale_filter, av1t_upscale_filter, sizeof(av1t_upscale_filter))) != cudaSuccess) return e;
        if (dev < 64) g_sr_const_loaded[dev] = true;
    }
    // chroma planes are at most as large as luma: the grid is sized for luma and chroma CTAs outside their plane exit at once
    {
        static bool carve_done = false;
        if (!carve_done) {
            prefer_max_smem(superres_kernel<uint8_t>);
            prefer_max_smem(superres_kernel<uint16_t>);
            carve_done = true;
        }
    }
    dim3 grid((L.up_w[0] + 255) / 256, L.h[0], L.planes);
    if (L.bd == 8) superres_kernel<uint8_t><<<grid, 256, 0, s>>>(L);
    else superres_kernel<uint16_t><<<grid, 256, 0, s>>>(L);
    return cudaGetLastError();
}

}  // namespace av1r
