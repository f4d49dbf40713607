// K7 -- loop restoration (AV1 spec 7.17): Wiener 7-tap separable and self-guided projection, sm_100a.
//
// One CTA per (64-column tile, 64-luma-row stripe, plane).  The CDEF output tile plus a 3-sample halo
// is staged in shared memory once, with the normative stripe rule applied while loading (rows outside
// the stripe come from the deblocked, pre-CDEF frame and only the 2 nearest are used; columns/rows clamp
// at the plane edges) -- after that both filters run purely out of shared memory.  Out of place; units
// whose type is NONE are copied through with 128-bit streaming loads / stores.  Algorithmic bytes ~2.06 F
// (4 boundary rows per 64-row stripe).
#include <cuda_runtime.h>
#include <stdint.h>

#include "dev_common.cuh"
#include "devframe.h"
#include "intra.h"
#include "../tables/tables_filter.inc"

namespace av1r {

static constexpr int LR_TW = 64, LR_TH = 64, LR_H = 3;
static constexpr int LR_SH = LR_TH + 2 * LR_H;      // staged rows
// Staged tile: row stride 80 samples (160 B, 16-byte aligned); tile column c (sample x0 - 3 + c) lives at row offset c + 1, so that
// offset 0 is sample x0 - 4 -- an 8-byte aligned address in the frame (x0 is a multiple of 64) and rows come in as 64-bit loads.
static constexpr int LR_SW = 80;
static constexpr int LR_IW = LR_TW + 8;             // row stride of the Wiener intermediate (144 B, 16-byte aligned)

__constant__ int16_t c_sgr_params[16][4];
static bool g_lr_const_loaded[64] = {false};

struct LrSmem {
    __align__(16) uint16_t tile[LR_SH * LR_SW];
    union {
        __align__(16) int16_t inter[LR_SH * LR_IW];            // Wiener horizontal pass
        struct {
            uint16_t A[(LR_TH + 2) * (LR_TW + 2)];
            int32_t B[(LR_TH + 2) * (LR_TW + 2)];
        } sg;
    } u;
};

// K7: one CTA per (64-column tile, 64-luma-row stripe, plane).
//   staging : a warp per row; rows of interior tiles arrive as 64-bit loads (four 16-bit samples), edge tiles sample by sample with
//             the column clamp.  The stripe rule picks the source row: CDEF output inside the stripe, the deblocked frame for the
//             (at most two) rows used above / below it.
//   Wiener  : horizontal pass 8 outputs per thread from two 128-bit shared loads (symmetric taps: 4 multiplies per output), packed
//             int16 intermediate; vertical pass two columns x four rows per thread, 32-bit stores.
//   SGR     : box sums from shared memory, projection, stores.
template <typename T>
__global__ void __launch_bounds__(256) lr_kernel(LrLaunch L) {
    __shared__ LrSmem sm;
    const int plane = blockIdx.z;
    const DevFrameParams& fp = L.fp;
    const int sy = plane ? fp.suby : 0;
    const int pw = fp.w[plane], ph = fp.h[plane];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
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
        if (sizeof(T) == 2 && (w & 7) == 0) {   // rows of 16-byte pieces (x0 is a multiple of 64 samples)
            const int lg = 31 - __clz(w >> 3);
            if ((w >> 3) == (1 << lg)) {
                for (int i = tid; i < (h << lg); i += 256) {
                    const int r = i >> lg, c = (i & ((1 << lg) - 1)) << 3;
                    st_stream128(dst + (size_t)(y0 + r) * ope + x0 + c, ld_stream128(cdef + (size_t)(y0 + r) * cpe + x0 + c));
                }
                return;
            }
        }
        for (int i = tid; i < w * h; i += 256) {
            const int r = i / w, c = i - r * w;
            dst[(size_t)(y0 + r) * ope + x0 + c] = cdef[(size_t)(y0 + r) * cpe + x0 + c];
        }
        return;
    }
    // ---- stage the tile: sample(x0 - 3 + c, y0 - 3 + r) with the stripe rule, at tile[r * LR_SW + c + 1]
    {
        const bool fast = sizeof(T) == 2 && x0 >= 4 && x0 + w + 3 <= pw - 1 && (size_t)(x0 + 72) * sizeof(T) <= L.cdef.pitch[plane] &&
                          (size_t)(x0 + 72) * sizeof(T) <= L.deblocked.pitch[plane];
        for (int r = warp; r < h + 6; r += 8) {
            const int y = min(max(y0 - 3 + r, 0), ph - 1);
            const T* row;
            if (y < ys) row = dbl + (size_t)max(ys - 2, y) * dpe;
            else if (y > ye) row = dbl + (size_t)min(ye + 2, y) * dpe;
            else row = cdef + (size_t)y * cpe;
            if (fast) {
                if (lane < 19) {
                    const uint2 v = __ldg(reinterpret_cast<const uint2*>(row + x0 - 4) + lane);
                    reinterpret_cast<uint2*>(sm.tile + r * LR_SW)[lane] = v;
                }
            } else {
                for (int c = lane; c < w + 6; c += 32) sm.tile[r * LR_SW + c + 1] = (uint16_t)row[min(max(x0 - 3 + c, 0), pw - 1)];
            }
        }
    }
    __syncthreads();
    if (type == RESTORE_WIENER_D) {
        const int round0 = bd == 12 ? 5 : 3, round1 = bd == 12 ? 9 : 11;
        int vf[4], hf[4];                                        // taps 0..2 and the centre tap (symmetric filters)
        vf[3] = hf[3] = 128;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            vf[i] = u.wiener[0][i];
            vf[3] -= 2 * u.wiener[0][i];
            hf[i] = u.wiener[1][i];
            hf[3] -= 2 * u.wiener[1][i];
        }
        const int offset = 1 << (bd + 7 - round0 - 1);
        const int limit = (1 << (bd + 1 + 7 - round0)) - 1;
        const int rnd0 = 1 << (round0 - 1), lo = -offset, hi = limit - offset;
        for (int idx = tid; idx < (h + 6) * 8; idx += 256) {   // horizontal: (row, group of 8 columns)
            const int r = idx >> 3, c0 = (idx & 7) << 3;
            if (c0 >= w) continue;
            const uint4* wp = reinterpret_cast<const uint4*>(sm.tile + r * LR_SW + c0);
            const uint4 a = wp[0], b = wp[1];
            const uint32_t wv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            int x[16];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                x[2 * k] = (int)(wv[k] & 0xffffu);
                x[2 * k + 1] = (int)(wv[k] >> 16);
            }
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 8; k++) {                       // output column c0 + k reads tile columns c0 + k .. c0 + k + 6 = x[k + 1 .. k + 7]
                int s = rnd0 + hf[3] * x[k + 4];
                s += hf[0] * (x[k + 1] + x[k + 7]);
                s += hf[1] * (x[k + 2] + x[k + 6]);
                s += hf[2] * (x[k + 3] + x[k + 5]);
                const int v = min(max(s >> round0, lo), hi);
                if (k & 1) o[k >> 1] |= (uint32_t)v << 16;
                else o[k >> 1] = (uint32_t)v & 0xffffu;
            }
            *reinterpret_cast<uint4*>(sm.u.inter + r * LR_IW + c0) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        __syncthreads();
        const int rnd1 = 1 << (round1 - 1);
        for (int idx = tid; idx < ((h + 3) >> 2) * 32; idx += 256) {   // vertical: (column pair, group of 4 rows)
            const int cp = idx & 31, r0 = (idx >> 5) << 2;
            if (2 * cp >= w) continue;
            const uint32_t* mp = reinterpret_cast<const uint32_t*>(sm.u.inter + r0 * LR_IW) + cp;
            int lo_[10], hi_[10];
#pragma unroll
            for (int t = 0; t < 10; t++) {
                const uint32_t v = (r0 + t < h + 6) ? mp[t * (LR_IW / 2)] : 0u;
                lo_[t] = (int)(short)(v & 0xffffu);
                hi_[t] = (int)v >> 16;
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (r0 + k >= h) break;
                int s0 = rnd1 + vf[3] * lo_[k + 3], s1 = rnd1 + vf[3] * hi_[k + 3];
                s0 += vf[0] * (lo_[k] + lo_[k + 6]) + vf[1] * (lo_[k + 1] + lo_[k + 5]) + vf[2] * (lo_[k + 2] + lo_[k + 4]);
                s1 += vf[0] * (hi_[k] + hi_[k + 6]) + vf[1] * (hi_[k + 1] + hi_[k + 5]) + vf[2] * (hi_[k + 2] + hi_[k + 4]);
                const int v0 = min(max(s0 >> round1, 0), pixmax), v1 = min(max(s1 >> round1, 0), pixmax);
                T* dp = dst + (size_t)(y0 + r0 + k) * ope + x0 + 2 * cp;
                if (2 * cp + 1 < w) {
                    if (sizeof(T) == 2) *reinterpret_cast<uint32_t*>(dp) = (uint32_t)v0 | ((uint32_t)v1 << 16);
                    else *reinterpret_cast<uint16_t*>(dp) = (uint16_t)(v0 | (v1 << 8));
                } else {
                    dp[0] = (T)v0;
                }
            }
        }
        return;
    }
    // ---- self-guided
    __shared__ uint16_t s_xbx[256];               // x / (x + 1) in 8-bit fixed point (spec 7.17.3), instead of a division per grid point
    s_xbx[tid] = (uint16_t)(tid == 255 ? 256 : (tid == 0 ? 1 : ((tid << 8) + tid / 2) / (tid + 1)));
    __syncthreads();
    const int r0 = c_sgr_params[u.sgr_set][0], r1 = c_sgr_params[u.sgr_set][1];
    const int s0 = c_sgr_params[u.sgr_set][2], s1 = c_sgr_params[u.sgr_set][3];
    constexpr int PPT = LR_TW * LR_TH / 256;      // pixels per thread
    int f0[PPT];
    const int gw = w + 2;
    const uint16_t* tile = sm.tile + 1;           // tile column c at tile[r * LR_SW + c]
    for (int pass = 0; pass < 2; pass++) {
        const int r = pass ? r1 : r0, sp = pass ? s1 : s0;
        if (r) {
            const int n = (2 * r + 1) * (2 * r + 1);
            const uint32_t one_by_n = ((1u << 12) + n / 2) / n;
            // pass 0 (r = 2) only ever reads the odd grid rows (weights of the even ones are zero): skip the others
            // (grid rows are walked 66 columns at a time whatever the tile width: a multiply-shift instead of a division by gw)
            for (int i = tid; i < (h + 2) * (LR_TW + 2); i += 256) {
                const int gi = (i * 993) >> 16, gj = i - gi * (LR_TW + 2);     // i / 66 for i < 66 * 66; grid point (gi - 1, gj - 1)
                if (gj >= gw) continue;
                if (pass == 0 && !((gi - 1) & 1)) continue;
                uint32_t a = 0, b = 0;
                const uint16_t* tc = tile + (gi + 2) * LR_SW + gj + 2;
                if (r == 2) {
#pragma unroll
                    for (int dy = -2; dy <= 2; dy++)
#pragma unroll
                        for (int dx = -2; dx <= 2; dx++) {
                            const uint32_t v = tc[dy * LR_SW + dx];
                            a += v * v;
                            b += v;
                        }
                } else {
#pragma unroll
                    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
                        for (int dx = -1; dx <= 1; dx++) {
                            const uint32_t v = tc[dy * LR_SW + dx];
                            a += v * v;
                            b += v;
                        }
                }
                const int sh = bd - 8;
                a = sh ? (a + (1u << (2 * sh - 1))) >> (2 * sh) : a;
                const uint32_t d = sh ? (b + (1u << (sh - 1))) >> sh : b;
                const uint32_t p = a * n < d * d ? 0 : a * n - d * d;
                const uint32_t z = (uint32_t)(((uint64_t)p * (uint32_t)sp + (1u << 19)) >> 20);
                const uint32_t a2 = s_xbx[min(z, 255u)];
                const uint32_t b2 = (256 - a2) * b * one_by_n;
                sm.u.sg.A[gi * (LR_TW + 2) + gj] = (uint16_t)a2;
                sm.u.sg.B[gi * (LR_TW + 2) + gj] = (int32_t)((b2 + (1u << 11)) >> 12);
            }
        }
        __syncthreads();
        int k = 0;
        for (int i = tid; i < h * LR_TW; i += 256, k++) {   // 64 columns per row whatever the tile width: shifts instead of a division
            const int pr = i >> 6, pc = i & (LR_TW - 1);
            if (pc >= w) continue;
            const int uu = (int)tile[(pr + 3) * LR_SW + pc + 3];
            int f = uu << 4;
            if (r) {
                int a = 0, b = 0;
                if (pass == 0) {
                    if (pr & 1) {     // odd row: the grid row itself, weights 5 6 5
#pragma unroll
                        for (int dx = -1; dx <= 1; dx++) {
                            const int wgt = dx == 0 ? 6 : 5;
                            a += wgt * (int)sm.u.sg.A[(pr + 1) * (LR_TW + 2) + pc + 1 + dx];
                            b += wgt * sm.u.sg.B[(pr + 1) * (LR_TW + 2) + pc + 1 + dx];
                        }
                    } else {          // even row: the odd grid rows above and below, weights 5 6 5 each
#pragma unroll
                        for (int dy = -1; dy <= 1; dy += 2)
#pragma unroll
                            for (int dx = -1; dx <= 1; dx++) {
                                const int wgt = dx == 0 ? 6 : 5;
                                a += wgt * (int)sm.u.sg.A[(pr + 1 + dy) * (LR_TW + 2) + pc + 1 + dx];
                                b += wgt * sm.u.sg.B[(pr + 1 + dy) * (LR_TW + 2) + pc + 1 + dx];
                            }
                    }
                } else {
#pragma unroll
                    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
                        for (int dx = -1; dx <= 1; dx++) {
                            const int wgt = (dx == 0 || dy == 0) ? 4 : 3;
                            a += wgt * (int)sm.u.sg.A[(pr + 1 + dy) * (LR_TW + 2) + pc + 1 + dx];
                            b += wgt * sm.u.sg.B[(pr + 1 + dy) * (LR_TW + 2) + pc + 1 + dx];
                        }
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
