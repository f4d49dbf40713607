// K7 -- loop restoration (AV1 spec 7.17): Wiener 7-tap separable and self-guided projection, sm_100a.
//
// One CTA per (64-column tile, 64-luma-row stripe, plane).  The CDEF output tile plus a 3-sample halo
// is staged in shared memory once, with the normative stripe rule applied while loading (rows outside
// the stripe come from the deblocked, pre-CDEF frame and only the 2 nearest are used; columns/rows clamp
// at the plane edges) -- after that both filters run purely out of shared memory.  Out of place; units
// whose type is NONE are copied through.  Algorithmic bytes ~2.06 F (4 boundary rows per 64-row stripe).
#include <cuda_runtime.h>
#include <stdint.h>

#include "dev_common.cuh"
#include "devframe.h"
#include "intra.h"
#include "../tables/tables_filter.inc"

namespace av1r {

static constexpr int LR_TW = 64, LR_TH = 64, LR_H = 3;
static constexpr int LR_SW = LR_TW + 2 * LR_H;      // staged tile width
static constexpr int LR_SH = LR_TH + 2 * LR_H;

__constant__ int16_t c_sgr_params[16][4];
static bool g_lr_const_loaded[64] = {false};

struct LrSmem {
    uint16_t tile[LR_SH * LR_SW];
    union {
        int16_t inter[LR_SH * LR_TW];                          // Wiener horizontal pass
        struct {
            uint16_t A[(LR_TH + 2) * (LR_TW + 2)];
            int32_t B[(LR_TH + 2) * (LR_TW + 2)];
        } sg;
    } u;
};

template <typename T>
__global__ void __launch_bounds__(256) lr_kernel(LrLaunch L) {
    __shared__ LrSmem sm;
    const int plane = blockIdx.z;
    const DevFrameParams& fp = L.fp;
    const int sx = plane ? fp.subx : 0, sy = plane ? fp.suby : 0;
    const int pw = fp.w[plane], ph = fp.h[plane];
    const int tid = threadIdx.x;
    const int stripe = blockIdx.y;
    const int ls = -8 + stripe * 64;
    const int ys = ls >> sy, ye = ys + (64 >> sy) - 1;          // StripeStartY / StripeEndY in plane rows
    const int y0 = max(ys, 0), y1 = min(ye, ph - 1);
    const int x0 = blockIdx.x * LR_TW;
    if (y0 > ph - 1 || x0 >= pw) return;
    const int w = min(LR_TW, pw - x0), h = y1 - y0 + 1;
    const int bd = fp.bd, pixmax = (1 << bd) - 1;
    const T* cdef = (const T*)L.cdef.p[plane];
    const T* dbl = (const T*)L.deblocked.p[plane];
    const int cpe = L.cdef.pitch[plane] / sizeof(T), dpe = L.deblocked.pitch[plane] / sizeof(T);
    T* dst = (T*)L.dst.p[plane];
    const int ope = L.dst.pitch[plane] / sizeof(T);
    // unit of this tile
    int type = RESTORE_NONE_D;
    LrUnitDev u;
    if (L.lr_type[plane] != 0) {
        const int unit_size = L.unit_size[plane];
        const int unit_row = min(L.unit_rows[plane] - 1, ((max(ls, 0) + 8) >> sy) / unit_size);
        const int unit_col = min(L.unit_cols[plane] - 1, x0 / unit_size);
        u = L.units[plane][unit_row * L.unit_cols[plane] + unit_col];
        type = u.type;
    }
    if (type == RESTORE_NONE_D) {
        for (int i = tid; i < w * h; i += 256) {
            const int r = i / w, c = i - r * w;
            dst[(size_t)(y0 + r) * ope + x0 + c] = cdef[(size_t)(y0 + r) * cpe + x0 + c];
        }
        return;
    }
    // ---- stage the tile: sample(x0 - 3 + c, y0 - 3 + r) with the stripe rule
    for (int i = tid; i < (h + 6) * (w + 6); i += 256) {
        const int r = i / (w + 6), c = i - r * (w + 6);
        int x = min(max(x0 - 3 + c, 0), pw - 1);
        int y = min(max(y0 - 3 + r, 0), ph - 1);
        int v;
        if (y < ys) v = dbl[(size_t)max(ys - 2, y) * dpe + x];
        else if (y > ye) v = dbl[(size_t)min(ye + 2, y) * dpe + x];
        else v = cdef[(size_t)y * cpe + x];
        sm.tile[r * LR_SW + c] = (uint16_t)v;
    }
    __syncthreads();
    if (type == RESTORE_WIENER_D) {
        const int round0 = bd == 12 ? 5 : 3, round1 = bd == 12 ? 9 : 11;
        int vf[7], hf[7];
        vf[3] = hf[3] = 128;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            vf[i] = vf[6 - i] = u.wiener[0][i];
            vf[3] -= 2 * u.wiener[0][i];
            hf[i] = hf[6 - i] = u.wiener[1][i];
            hf[3] -= 2 * u.wiener[1][i];
        }
        const int offset = 1 << (bd + 7 - round0 - 1);
        const int limit = (1 << (bd + 1 + 7 - round0)) - 1;
        for (int i = tid; i < (h + 6) * w; i += 256) {
            const int r = i / w, c = i - r * w;
            int s = 0;
#pragma unroll
            for (int t = 0; t < 7; t++) s += hf[t] * sm.tile[r * LR_SW + c + t];
            const int v = (s + (1 << (round0 - 1))) >> round0;
            sm.u.inter[r * LR_TW + c] = (int16_t)min(max(v, -offset), limit - offset);
        }
        __syncthreads();
        for (int i = tid; i < h * w; i += 256) {
            const int r = i / w, c = i - r * w;
            int s = 0;
#pragma unroll
            for (int t = 0; t < 7; t++) s += vf[t] * sm.u.inter[(r + t) * LR_TW + c];
            const int v = (s + (1 << (round1 - 1))) >> round1;
            dst[(size_t)(y0 + r) * ope + x0 + c] = (T)min(max(v, 0), pixmax);
        }
        return;
    }
    // ---- self-guided
    const int r0 = c_sgr_params[u.sgr_set][0], r1 = c_sgr_params[u.sgr_set][1];
    const int s0 = c_sgr_params[u.sgr_set][2], s1 = c_sgr_params[u.sgr_set][3];
    constexpr int PPT = LR_TW * LR_TH / 256;      // pixels per thread
    int f0[PPT];
    const int gw = w + 2;
    for (int pass = 0; pass < 2; pass++) {
        const int r = pass ? r1 : r0, sp = pass ? s1 : s0;
        if (r) {
            const int n = (2 * r + 1) * (2 * r + 1);
            const uint32_t one_by_n = ((1u << 12) + n / 2) / n;
            for (int i = tid; i < (h + 2) * gw; i += 256) {
                const int gi = i / gw, gj = i - gi * gw;     // grid point (gi - 1, gj - 1) -> tile centre (gi + 2, gj + 2)
                uint32_t a = 0, b = 0;
                for (int dy = -r; dy <= r; dy++)
                    for (int dx = -r; dx <= r; dx++) {
                        const uint32_t v = sm.tile[(gi + 2 + dy) * LR_SW + gj + 2 + dx];
                        a += v * v;
                        b += v;
                    }
                const int sh = bd - 8;
                a = sh ? (a + (1u << (2 * sh - 1))) >> (2 * sh) : a;
                const uint32_t d = sh ? (b + (1u << (sh - 1))) >> sh : b;
                const uint32_t p = a * n < d * d ? 0 : a * n - d * d;
                const uint32_t z = (uint32_t)(((uint64_t)p * (uint32_t)sp + (1u << 19)) >> 20);
                uint32_t a2;
                if (z >= 255) a2 = 256;
                else if (z == 0) a2 = 1;
                else a2 = ((z << 8) + z / 2) / (z + 1);
                const uint32_t b2 = (256 - a2) * b * one_by_n;
                sm.u.sg.A[gi * (LR_TW + 2) + gj] = (uint16_t)a2;
                sm.u.sg.B[gi * (LR_TW + 2) + gj] = (int32_t)((b2 + (1u << 11)) >> 12);
            }
        }
        __syncthreads();
        int k = 0;
        for (int i = tid; i < h * w; i += 256, k++) {
            const int pr = i / w, pc = i - pr * w;
            const int uu = (int)sm.tile[(pr + 3) * LR_SW + pc + 3];
            int f = uu << 4;
            if (r) {
                int a = 0, b = 0;
#pragma unroll
                for (int dy = -1; dy <= 1; dy++)
#pragma unroll
                    for (int dx = -1; dx <= 1; dx++) {
                        int wgt;
                        if (pass == 0) wgt = ((pr + dy) & 1) ? (dx == 0 ? 6 : 5) : 0;
                        else wgt = (dx == 0 || dy == 0) ? 4 : 3;
                        a += wgt * (int)sm.u.sg.A[(pr + 1 + dy) * (LR_TW + 2) + pc + 1 + dx];
                        b += wgt * sm.u.sg.B[(pr + 1 + dy) * (LR_TW + 2) + pc + 1 + dx];
                    }
                int shift = 5;
                if (pass == 0 && (pr & 1)) shift = 4;
                const int v = a * uu + b;
                const int rs = 8 + shift - 4;
                f = (v + (1 << (rs - 1))) >> rs;
            }
            if (pass == 0) {
                f0[k] = f;
            } else {
                const int w0 = u.sgr_xqd[0], w1 = u.sgr_xqd[1], w2 = 128 - w0 - w1;
                const int v = w1 * (uu << 4) + w0 * f0[k] + w2 * f;
                dst[(size_t)(y0 + pr) * ope + x0 + pc] = (T)min(max((v + (1 << 10)) >> 11, 0), pixmax);
            }
        }
        __syncthreads();
    }
}

cudaError_t launch_lr(const LrLaunch& L, cudaStream_t s) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (!(dev < 64 && g_lr_const_loaded[dev])) {
        if ((e = cudaMemcpyToSymbol(c_sgr_params, av1t_sgr_params, sizeof(av1t_sgr_params))) != cudaSuccess) return e;
        if (dev < 64) g_lr_const_loaded[dev] = true;
    }
    {
        static bool carve_done = false;
        if (!carve_done) {
            prefer_max_smem(lr_kernel<uint8_t>);
            prefer_max_smem(lr_kernel<uint16_t>);
            carve_done = true;
        }
    }
    const int nstripes = (L.fp.h[0] + 8 + 63) / 64;
    dim3 grid((L.fp.w[0] + LR_TW - 1) / LR_TW, nstripes, L.fp.mono ? 1 : 3);
    if (L.fp.bd == 8) lr_kernel<uint8_t><<<grid, 256, 0, s>>>(L);
    else lr_kernel<uint16_t><<<grid, 256, 0, s>>>(L);
    return cudaGetLastError();
}

}  // namespace av1r
