// Small device helpers shared by the reconstruction kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace av1r {

// Every kernel of the engine asks for the same (largest) shared-memory carve-out: the SMs then never have to drain to be
// re-partitioned between L1 and shared memory when kernels of different frames (K3 needs 60-75 KB per CTA, the filters a few KB)
// land on the same SM, which is what lets the stages of the frames in flight overlap.
template <class F>
static inline void prefer_max_smem(F f) {
    cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

__device__ __forceinline__ int d_round2(int x, int n) { return n == 0 ? x : (x + (1 << (n - 1))) >> n; }
__device__ __forceinline__ int d_clip3(int lo, int hi, int x) { return min(max(x, lo), hi); }

// 128-bit streaming load/store (pixels are touched once per stage: keep them out of L1).
__device__ __forceinline__ uint4 ld_stream128(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream128(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

template <typename T> struct PixTraits;
template <> struct PixTraits<uint8_t> { static constexpr int VEC = 16; };
template <> struct PixTraits<uint16_t> { static constexpr int VEC = 8; };

// unpack / pack a 16-byte vector to ints
__device__ __forceinline__ void unpack16(const uint4& v, int* o, uint8_t) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        o[4 * i + 0] = w[i] & 0xff;
        o[4 * i + 1] = (w[i] >> 8) & 0xff;
        o[4 * i + 2] = (w[i] >> 16) & 0xff;
        o[4 * i + 3] = (w[i] >> 24);
    }
}
__device__ __forceinline__ void unpack16(const uint4& v, int* o, uint16_t) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        o[2 * i + 0] = w[i] & 0xffff;
        o[2 * i + 1] = w[i] >> 16;
    }
}
__device__ __forceinline__ uint4 pack16(const int* o, uint8_t) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
        w[i] = (uint32_t)o[4 * i] | ((uint32_t)o[4 * i + 1] << 8) | ((uint32_t)o[4 * i + 2] << 16) | ((uint32_t)o[4 * i + 3] << 24);
    return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ uint4 pack16(const int* o, uint16_t) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; i++) w[i] = (uint32_t)o[2 * i] | ((uint32_t)o[2 * i + 1] << 16);
    return make_uint4(w[0], w[1], w[2], w[3]);
}

}  // namespace av1r
