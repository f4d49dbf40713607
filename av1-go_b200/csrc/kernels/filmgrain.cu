// K8 -- film grain synthesis (AV1 spec 7.18.3) for sm_100a.
//
// Two kernels per frame:
//   fg_prepare : one CTA. Grain templates (LFSR jump-ahead in parallel, AR filter as a row
//                wavefront), full-resolution scaling LUTs, per-32x32-block random offsets.
//   fg_apply   : streaming kernel, one CTA per 32-luma-row stripe x 512 luma columns (plus the
//                co-located chroma).  Templates + LUTs are staged in shared memory, pixels move
//                as 128-bit L1-bypassing loads/stores.  Out of place: the reference frame stays
//                grain-free, the display copy gets the noise.
// Algorithmic bytes: 2F per frame (read F, write F); the luma re-read for chroma averaging hits
// L1/L2 (same CTA touched those rows).  HBM-bound by design; no tensor-core work exists here.
//
// Replaces (in the reference's pipeline) the film-grain pass libdav1d performs inside the ffmpeg
// child the daemon would spawn (/root/reference/internal/ffmpeg/transcode.go:195).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../../include/av1r_stages.h"
#include "dev_common.cuh"
#include "../tables/tables_fg.inc"

namespace av1r {

static constexpr int FG_MAX_STRIPES = 512;   // frame height <= 16384
static constexpr int FG_MAX_BLOCKS = 512;    // frame width  <= 16384
static constexpr int FG_TW = 512;            // luma columns per CTA

struct FgDev {
    int16_t luma[73 * 82];
    int16_t cb[73 * 82];
    int16_t cr[73 * 82];
    uint8_t lut[3][4096];
    uint8_t offs[FG_MAX_STRIPES * FG_MAX_BLOCKS];
};

struct FgK {   // kernel parameter block (by value)
    av1r_film_grain_params p;
    int bd, w, h, subx, suby, mono, mc_identity;
    int nstripes, nblocks;
};

__constant__ int16_t c_gauss[2048];
static bool g_gauss_loaded[64] = {false};

__device__ __forceinline__ unsigned lfsr_step(unsigned r) {
    unsigned bit = (r ^ (r >> 1) ^ (r >> 3) ^ (r >> 12)) & 1u;
    return (r >> 1) | (bit << 15);
}
__device__ __forceinline__ unsigned mat_apply(const uint16_t* m, unsigned r) {
    unsigned o = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) o |= (unsigned)(__popc(m[i] & r) & 1) << i;
    return o;
}
__device__ __forceinline__ unsigned lfsr_jump(const uint16_t (*jm)[16], unsigned r, int k) {
    for (int j = 0; k; j++, k >>= 1)
        if (k & 1) r = mat_apply(jm[j], r);
    return r;
}

__device__ void build_lut256(int n, const int* val, const int* sc, int* lut, int tid, int nthr) {
    for (int i = tid; i < 256; i += nthr) {
        int v;
        if (n == 0) v = 0;
        else if (i < val[0]) v = sc[0];
        else if (i >= val[n - 1]) v = sc[n - 1];
        else {
            int s = 0;
            while (s < n - 2 && i >= val[s + 1]) s++;
            int dy = sc[s + 1] - sc[s], dx = val[s + 1] - val[s];
            int delta = dy * ((65536 + (dx >> 1)) / dx);
            int x = i - val[s];
            v = sc[s] + ((x * delta + 32768) >> 16);
        }
        lut[i] = v;
    }
}

// Auto-regressive grain filter (spec 7.18.3.3), row wavefront: row y may process column x once row y-1 has finished x + lag.
// The lag is a template parameter so that the 2 * lag * (lag + 1) taps unroll into independent shared-memory loads with the
// coefficients in registers: a step of the wavefront costs one load latency instead of a 24-iteration dependent loop.
template <int LAG>
__device__ __forceinline__ void fg_ar_filter(int16_t* g_luma, int16_t* g_cb, int16_t* g_cr, const int (*s_coef)[25], int ash, int gmin, int gmax,
                                             int cw, int ch, int subx, int suby, int mono, bool luma_term, int tid) {
    constexpr int NT = 2 * LAG * (LAG + 1);
    {
        constexpr int rows = 70, cols = 76, steps = cols + (LAG + 1) * (rows - 1);
        int coef[NT > 0 ? NT : 1];
#pragma unroll
        for (int i = 0; i < NT; i++) coef[i] = s_coef[0][i];
        const int y = tid + 3;
        for (int t = 0; t < steps; t++) {
            const int x = 3 + t - (LAG + 1) * tid;
            if (tid < rows && x >= 3 && x < 3 + cols) {
                int sum = 0;
#pragma unroll
                for (int dr = -LAG; dr <= 0; dr++)
#pragma unroll
                    for (int dc = -LAG; dc <= LAG; dc++)
                        if (dr < 0 || dc < 0) sum += coef[(dr + LAG) * (2 * LAG + 1) + dc + LAG] * g_luma[(y + dr) * 82 + x + dc];
                const int v = g_luma[y * 82 + x] + d_round2(sum, ash);
                g_luma[y * 82 + x] = (int16_t)d_clip3(gmin, gmax, v);
            }
            __syncthreads();
        }
    }
    if (!mono) {
        const int rows = ch - 3, cols = cw - 6;
        const int steps = cols + (LAG + 1) * (rows - 1);
        const int pl = tid >> 7;          // 0: cb (threads 0..127)  1: cr (threads 128..255)
        const int ry = tid & 127;
        int16_t* gp = pl ? g_cr : g_cb;
        int coef[NT + 1];
#pragma unroll
        for (int i = 0; i <= NT; i++) coef[i] = s_coef[1 + pl][i];
        const int y = ry + 3;
        for (int t = 0; t < steps; t++) {
            const int x = 3 + t - (LAG + 1) * ry;
            if (ry < rows && x >= 3 && x < 3 + cols) {
                int sum = 0;
#pragma unroll
                for (int dr = -LAG; dr <= 0; dr++)
#pragma unroll
                    for (int dc = -LAG; dc <= LAG; dc++)
                        if (dr < 0 || dc < 0) sum += coef[(dr + LAG) * (2 * LAG + 1) + dc + LAG] * gp[(y + dr) * cw + x + dc];
                if (luma_term) {
                    int luma = 0;
                    const int lx = ((x - 3) << subx) + 3, ly = ((y - 3) << suby) + 3;
                    for (int i = 0; i <= suby; i++)
                        for (int j = 0; j <= subx; j++) luma += g_luma[(ly + i) * 82 + lx + j];
                    sum += d_round2(luma, subx + suby) * coef[NT];
                }
                const int v = gp[y * cw + x] + d_round2(sum, ash);
                gp[y * cw + x] = (int16_t)d_clip3(gmin, gmax, v);
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(256) fg_prepare_kernel(FgK k, FgDev* __restrict__ st) {
    __shared__ uint16_t jm[14][16];
    __shared__ int lut256[3][256];
    __shared__ int s_coef[3][25];
    __shared__ int16_t g_luma[73 * 82], g_cb[73 * 82], g_cr[73 * 82];
    const int tid = threadIdx.x;
    const av1r_film_grain_params& p = k.p;
    const int bd = k.bd;
    // --- LFSR jump matrices: jm[j] = A^(2^j), row i = mask of input bits feeding output bit i
    if (tid < 16) jm[0][tid] = tid < 15 ? (uint16_t)(1u << (tid + 1)) : (uint16_t)0x100B;
    if (tid < 25) {
        s_coef[0][tid] = tid < 24 ? p.ar_coeffs_y[tid] : 0;
        s_coef[1][tid] = p.ar_coeffs_cb[tid];
        s_coef[2][tid] = p.ar_coeffs_cr[tid];
    }
    __syncthreads();
    for (int j = 1; j < 14; j++) {
        if (tid < 16) {
            unsigned row = jm[j - 1][tid], acc = 0;
            for (int b = 0; b < 16; b++)
                if ((row >> b) & 1) acc ^= jm[j - 1][b];
            jm[j][tid] = (uint16_t)acc;
        }
        __syncthreads();
    }
    // --- raw Gaussian grain, chunked per thread
    const int gshift = 12 - bd + p.grain_scale_shift;
    const int cw = k.subx ? 44 : 82, ch = k.suby ? 38 : 73;
    {
        const int n = 73 * 82;
        const int chunk = (n + 255) / 256;
        int e0 = tid * chunk, e1 = min(n, e0 + chunk);
        if (e0 < n) {
            unsigned r = lfsr_jump(jm, (unsigned)p.grain_seed, e0);
            for (int e = e0; e < e1; e++) {
                r = lfsr_step(r);
                int g = p.num_y_points > 0 ? c_gauss[(r >> 5) & 2047] : 0;
                g_luma[e] = (int16_t)d_round2(g, gshift);
            }
        }
    }
    if (!k.mono) {
        const int n = ch * cw;
        const int chunk = (n + 255) / 256;
        int e0 = tid * chunk, e1 = min(n, e0 + chunk);
        if (e0 < n) {
            unsigned r0 = lfsr_jump(jm, (unsigned)p.grain_seed ^ 0xb524u, e0);
            unsigned r1 = lfsr_jump(jm, (unsigned)p.grain_seed ^ 0x49d8u, e0);
            const bool on0 = p.num_cb_points || p.chroma_scaling_from_luma;
            const bool on1 = p.num_cr_points || p.chroma_scaling_from_luma;
            for (int e = e0; e < e1; e++) {
                r0 = lfsr_step(r0);
                r1 = lfsr_step(r1);
                g_cb[e] = (int16_t)d_round2(on0 ? (int)c_gauss[(r0 >> 5) & 2047] : 0, gshift);
                g_cr[e] = (int16_t)d_round2(on1 ? (int)c_gauss[(r1 >> 5) & 2047] : 0, gshift);
            }
        }
    }
    __syncthreads();
    // --- AR filter: row wavefront (row y may process column x once row y-1 has finished x+lag)
    const int lag = p.ar_coeff_lag;
    const int ash = p.ar_coeff_shift;
    const int gcenter = 128 << (bd - 8);
    const int gmin = -gcenter, gmax = (256 << (bd - 8)) - 1 - gcenter;
    switch (lag) {
        case 0: fg_ar_filter<0>(g_luma, g_cb, g_cr, s_coef, ash, gmin, gmax, cw, ch, k.subx, k.suby, k.mono, p.num_y_points > 0, tid); break;
        case 1: fg_ar_filter<1>(g_luma, g_cb, g_cr, s_coef, ash, gmin, gmax, cw, ch, k.subx, k.suby, k.mono, p.num_y_points > 0, tid); break;
        case 2: fg_ar_filter<2>(g_luma, g_cb, g_cr, s_coef, ash, gmin, gmax, cw, ch, k.subx, k.suby, k.mono, p.num_y_points > 0, tid); break;
        default: fg_ar_filter<3>(g_luma, g_cb, g_cr, s_coef, ash, gmin, gmax, cw, ch, k.subx, k.suby, k.mono, p.num_y_points > 0, tid); break;
    }
        for (int i = tid; i < 73 * 82; i += 256) {
        st->luma[i] = g_luma[i];
        st->cb[i] = g_cb[i];
        st->cr[i] = g_cr[i];
    }
    // --- scaling LUTs at full sample resolution
    build_lut256(p.num_y_points, p.point_y_value, p.point_y_scaling, lut256[0], tid, 256);
    if (p.chroma_scaling_from_luma) {
        build_lut256(p.num_y_points, p.point_y_value, p.point_y_scaling, lut256[1], tid, 256);
        build_lut256(p.num_y_points, p.point_y_value, p.point_y_scaling, lut256[2], tid, 256);
    } else {
        build_lut256(p.num_cb_points, p.point_cb_value, p.point_cb_scaling, lut256[1], tid, 256);
        build_lut256(p.num_cr_points, p.point_cr_value, p.point_cr_scaling, lut256[2], tid, 256);
    }
    __syncthreads();
    {
        const int shift = bd - 8;
        for (int pl = 0; pl < 3; pl++)
            for (int idx = tid; idx < (1 << bd); idx += 256) {
                int x = idx >> shift;
                int rem = idx - (x << shift);
                int v;
                if (bd == 8 || x == 255) v = lut256[pl][x];
                else {
                    int s = lut256[pl][x], e = lut256[pl][x + 1];
                    v = s + d_round2((e - s) * rem, shift);
                }
                st->lut[pl][idx] = (uint8_t)v;
            }
    }
    // --- per-block random offsets, one stripe per thread
    for (int s = tid; s < k.nstripes; s += 256) {
        unsigned r = (unsigned)p.grain_seed;
        r ^= (unsigned)((s * 37 + 178) & 255) << 8;
        r ^= (unsigned)((s * 173 + 105) & 255);
        for (int b = 0; b < k.nblocks; b++) {
            r = lfsr_step(r);
            st->offs[s * FG_MAX_BLOCKS + b] = (uint8_t)(r >> 8);
        }
    }
}

struct FgPlanes {
    const uint8_t* src[3];
    uint8_t* dst[3];
    size_t spitch[3], dpitch[3];
};

// Shared-memory view used by fg_apply.
struct FgSmem {
    const int16_t* tpl[3];
    const uint8_t* lut[3];
    const uint8_t* offs;     // [2][nb_local]: row 0 = stripe s-1, row 1 = stripe s
    int nb_local, b_first;   // offs column c <-> block b_first + c
    int tw[3];               // template row widths
};

template <int PSX, int PSY>
__device__ __forceinline__ int fg_stripe_val(const FgSmem& sm, int pl, int srow, int i, int x, int overlap, int gmin, int gmax) {
    // srow: 0 = previous stripe, 1 = current stripe; x = plane column
    constexpr int BS = 32 >> PSX;
    const int b = x / BS, j = x - b * BS;
    const int cw = sm.tw[pl];
    const int off = sm.offs[srow * sm.nb_local + (b - sm.b_first)];
    const int pox = PSX ? 6 + (off >> 4) : 9 + 2 * (off >> 4);
    const int poy = PSY ? 6 + (off & 15) : 9 + 2 * (off & 15);
    int g = sm.tpl[pl][(poy + i) * cw + pox + j];
    if (overlap && b > 0 && j < (2 >> PSX)) {
        const int offp = sm.offs[srow * sm.nb_local + (b - 1 - sm.b_first)];
        const int poxp = PSX ? 6 + (offp >> 4) : 9 + 2 * (offp >> 4);
        const int poyp = PSY ? 6 + (offp & 15) : 9 + 2 * (offp & 15);
        const int old = sm.tpl[pl][(poyp + i) * cw + poxp + j + BS];
        if (PSX == 0) g = (j == 0) ? old * 27 + g * 17 : old * 17 + g * 27;
        else g = old * 23 + g * 22;
        g = d_clip3(gmin, gmax, d_round2(g, 5));
    }
    return g;
}

template <int PSX, int PSY>
__device__ __forceinline__ int fg_noise(const FgSmem& sm, int pl, int x, int y, int overlap, int gmin, int gmax) {
    constexpr int RS = 32 >> PSY;
    const int s = y / RS, i = y - s * RS;
    int g = fg_stripe_val<PSX, PSY>(sm, pl, 1, i, x, overlap, gmin, gmax);
    if (overlap && s > 0 && i < (2 >> PSY)) {
        const int old = fg_stripe_val<PSX, PSY>(sm, pl, 0, i + RS, x, overlap, gmin, gmax);
        if (PSY == 0) g = (i == 0) ? old * 27 + g * 17 : old * 17 + g * 27;
        else g = old * 23 + g * 22;
        g = d_clip3(gmin, gmax, d_round2(g, 5));
    }
    return g;
}

template <typename T, int SSX, int SSY>
__global__ void __launch_bounds__(256) fg_apply_kernel(FgK k, const FgDev* __restrict__ st, FgPlanes pp) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    constexpr int VEC = PixTraits<T>::VEC;
    const av1r_film_grain_params& p = k.p;
    const int bd = k.bd, tid = threadIdx.x;
    const int s = blockIdx.y;                 // stripe
    const int x0 = blockIdx.x * FG_TW;        // first luma column
    const int cw = k.subx ? 44 : 82, ch = k.suby ? 38 : 73;
    const int nlut = 1 << bd;
    // shared layout: luma tpl | cb | cr | lut[3] | offs
    int16_t* s_luma = (int16_t*)smem_raw;
    int16_t* s_cb = s_luma + 73 * 82;
    int16_t* s_cr = s_cb + ch * cw;
    uint8_t* s_lut = (uint8_t*)(s_cr + ch * cw);
    uint8_t* s_offs = s_lut + 3 * nlut;
    const int b_first = max(0, x0 / 32 - 1);
    const int nb_local = FG_TW / 32 + 2;
    for (int i = tid; i < 73 * 82 / 2; i += 256) ((uint32_t*)s_luma)[i] = ((const uint32_t*)st->luma)[i];
    if (!k.mono) {
        for (int i = tid; i < ch * cw / 2; i += 256) {
            ((uint32_t*)s_cb)[i] = ((const uint32_t*)st->cb)[i];
            ((uint32_t*)s_cr)[i] = ((const uint32_t*)st->cr)[i];
        }
    }
    for (int pl = 0; pl < 3; pl++)
        for (int i = tid; i < nlut / 4; i += 256) ((uint32_t*)(s_lut + pl * nlut))[i] = ((const uint32_t*)st->lut[pl])[i];
    for (int i = tid; i < 2 * nb_local; i += 256) {
        int r = i / nb_local, c = i - r * nb_local;
        int ss = s - 1 + r, b = b_first + c;
        s_offs[i] = (ss >= 0 && b < k.nblocks) ? st->offs[ss * FG_MAX_BLOCKS + b] : 0;
    }
    __syncthreads();
    FgSmem sm;
    sm.tpl[0] = s_luma; sm.tpl[1] = s_cb; sm.tpl[2] = s_cr;
    sm.lut[0] = s_lut; sm.lut[1] = s_lut + nlut; sm.lut[2] = s_lut + 2 * nlut;
    sm.offs = s_offs; sm.nb_local = nb_local; sm.b_first = b_first;
    sm.tw[0] = 82; sm.tw[1] = cw; sm.tw[2] = cw;

    const int gcenter = 128 << (bd - 8);
    const int gmin = -gcenter, gmax = (256 << (bd - 8)) - 1 - gcenter;
    int min_value, max_luma, max_chroma;
    if (p.clip_to_restricted_range) {
        min_value = 16 << (bd - 8);
        max_luma = 235 << (bd - 8);
        max_chroma = k.mc_identity ? max_luma : (240 << (bd - 8));
    } else {
        min_value = 0;
        max_luma = max_chroma = (256 << (bd - 8)) - 1;
    }
    const int pixmax = (1 << bd) - 1;
    const int sshift = p.grain_scaling;
    const int overlap = p.overlap_flag;
    const int w = k.w, h = k.h;

    // ---- luma: 32 rows x FG_TW columns, VEC pixels per item
    {
        constexpr int IPR = FG_TW / VEC;   // items per row
        const bool on = p.num_y_points > 0;
        for (int it = tid; it < 32 * IPR; it += 256) {
            const int r = it / IPR, c = it - r * IPR;
            const int y = s * 32 + r, x = x0 + c * VEC;
            if (y >= h || x >= w) continue;
            const T* srow = (const T*)(pp.src[0] + (size_t)y * pp.spitch[0]);
            T* drow = (T*)(pp.dst[0] + (size_t)y * pp.dpitch[0]);
            int px[VEC];
            if (x + VEC <= w) {
                uint4 v = ld_stream128(srow + x);
                unpack16(v, px, T());
#pragma unroll
                for (int j = 0; j < VEC; j++) {
                    if (on) {
                        int g = fg_noise<0, 0>(sm, 0, x + j, y, overlap, gmin, gmax);
                        int noise = d_round2((int)sm.lut[0][px[j]] * g, sshift);
                        px[j] = d_clip3(min_value, max_luma, px[j] + noise);
                    }
                }
                st_stream128(drow + x, pack16(px, T()));
            } else {
                for (int j = 0; x + j < w; j++) {
                    int o = srow[x + j];
                    if (on) {
                        int g = fg_noise<0, 0>(sm, 0, x + j, y, overlap, gmin, gmax);
                        int noise = d_round2((int)sm.lut[0][o] * g, sshift);
                        o = d_clip3(min_value, max_luma, o + noise);
                    }
                    drow[x + j] = (T)o;
                }
            }
        }
    }
    if (k.mono) return;
    // ---- chroma planes co-located with this luma tile
    {
        const int pw = (w + SSX) >> SSX, ph = (h + SSY) >> SSY;
        constexpr int CROWS = 32 >> SSY;
        constexpr int CTW = FG_TW >> SSX;
        constexpr int IPR = CTW / VEC;
        const int cx0 = x0 >> SSX;
        for (int it = tid; it < 2 * CROWS * IPR; it += 256) {
            const int pl = 1 + it / (CROWS * IPR);
            const int rem = it - (pl - 1) * (CROWS * IPR);
            const int r = rem / IPR, c = rem - r * IPR;
            const int y = s * CROWS + r, x = cx0 + c * VEC;
            if (y >= ph || x >= pw) continue;
            const int npts = pl == 1 ? p.num_cb_points : p.num_cr_points;
            const bool on = npts > 0 || p.chroma_scaling_from_luma;
            const int lm = (pl == 1 ? p.cb_luma_mult : p.cr_luma_mult) - 128;
            const int mm = (pl == 1 ? p.cb_mult : p.cr_mult) - 128;
            const int off = ((pl == 1 ? p.cb_offset : p.cr_offset) - 256) << (bd - 8);
            const T* srow = (const T*)(pp.src[pl] + (size_t)y * pp.spitch[pl]);
            T* drow = (T*)(pp.dst[pl] + (size_t)y * pp.dpitch[pl]);
            const T* lrow = (const T*)(pp.src[0] + (size_t)(y << SSY) * pp.spitch[0]);
            const bool full = (x + VEC <= pw) && (((x + VEC) << SSX) <= w);
            if (full) {
                int px[VEC], la[VEC << SSX];
                unpack16(ld_stream128(srow + x), px, T());
                unpack16(ld_stream128(lrow + (x << SSX)), la, T());
                if (SSX) unpack16(ld_stream128(lrow + (x << SSX) + VEC), la + VEC, T());
#pragma unroll
                for (int j = 0; j < VEC; j++) {
                    if (on) {
                        int avg = SSX ? ((la[2 * j] + la[2 * j + 1] + 1) >> 1) : la[j];
                        int merged;
                        if (p.chroma_scaling_from_luma) merged = avg;
                        else merged = d_clip3(0, pixmax, ((avg * lm + px[j] * mm) >> 6) + off);
                        int g = fg_noise<SSX, SSY>(sm, pl, x + j, y, overlap, gmin, gmax);
                        int noise = d_round2((int)sm.lut[pl][merged] * g, sshift);
                        px[j] = d_clip3(min_value, max_chroma, px[j] + noise);
                    }
                }
                st_stream128(drow + x, pack16(px, T()));
            } else {
                for (int j = 0; x + j < pw; j++) {
                    int o = srow[x + j];
                    if (on) {
                        int lx = (x + j) << SSX;
                        int lnx = min(lx + 1, w - 1);
                        int avg = SSX ? ((lrow[lx] + lrow[lnx] + 1) >> 1) : lrow[lx];
                        int merged;
                        if (p.chroma_scaling_from_luma) merged = avg;
                        else merged = d_clip3(0, pixmax, ((avg * lm + o * mm) >> 6) + off);
                        int g = fg_noise<SSX, SSY>(sm, pl, x + j, y, overlap, gmin, gmax);
                        int noise = d_round2((int)sm.lut[pl][merged] * g, sshift);
                        o = d_clip3(min_value, max_chroma, o + noise);
                    }
                    drow[x + j] = (T)o;
                }
            }
        }
    }
}

static thread_local const char* g_stage_err = "";

template <typename T, int SSX, int SSY>
static cudaError_t launch_apply(const FgK& k, const FgDev* st, const FgPlanes& pp, cudaStream_t s) {
    const int cw = k.subx ? 44 : 82, ch = k.suby ? 38 : 73;
    size_t smem = 73 * 82 * 2 + 2 * (size_t)ch * cw * 2 + 3 * (size_t)(1 << k.bd) + 2 * (FG_TW / 32 + 2);
    smem = (smem + 15) & ~(size_t)15;
    auto kern = fg_apply_kernel<T, SSX, SSY>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid((k.w + FG_TW - 1) / FG_TW, k.nstripes);
    kern<<<grid, 256, smem, s>>>(k, st, pp);
    return cudaGetLastError();
}

}  // namespace av1r

using namespace av1r;

extern "C" size_t av1r_film_grain_scratch_bytes(void) { return sizeof(FgDev); }

extern "C" const char* av1r_stage_last_error(void) { return g_stage_err; }

// Two halves of the stage, so that the engine can start the (single-CTA, latency-bound) template preparation on a side stream as
// soon as the frame is issued -- it depends on the header's grain parameters only -- and apply it when the frame's pixels exist.
static int fg_setup(const av1r_film_grain_params* p, int bpc, int w, int h, int subx, int suby, int mono, int mc_identity, void* scratch, FgK& k) {
    if (!p || !scratch || w <= 0 || h <= 0 || (bpc != 8 && bpc != 10 && bpc != 12)) {
        g_stage_err = "film_grain: bad arguments";
        return -22;
    }
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        g_stage_err = "film_grain: no CUDA device";
        return -5;
    }
    if (dev < 64 && !g_gauss_loaded[dev]) {
        if (cudaMemcpyToSymbol(c_gauss, av1t_gaussian_sequence, sizeof(av1t_gaussian_sequence)) != cudaSuccess) {
            g_stage_err = "film_grain: constant upload failed";
            return -5;
        }
        g_gauss_loaded[dev] = true;
    }
    k.p = *p;
    k.bd = bpc; k.w = w; k.h = h; k.subx = subx; k.suby = suby; k.mono = mono; k.mc_identity = mc_identity;
    k.nstripes = (h + 31) / 32;
    k.nblocks = (((w + 1) / 2) + 15) / 16;
    if (k.nstripes > FG_MAX_STRIPES || k.nblocks > FG_MAX_BLOCKS) {
        g_stage_err = "film_grain: frame too large";
        return -38;
    }
    return 0;
}

namespace av1r {
int fg_launch_prepare(const av1r_film_grain_params* p, int bpc, int w, int h, int subx, int suby, int mono, int mc_identity, void* scratch,
                      cudaStream_t s) {
    FgK k;
    const int rc = fg_setup(p, bpc, w, h, subx, suby, mono, mc_identity, scratch, k);
    if (rc) return rc;
    fg_prepare_kernel<<<1, 256, 0, s>>>(k, (FgDev*)scratch);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_stage_err = cudaGetErrorString(e); return -5; }
    return 0;
}

int fg_launch_apply(const av1r_film_grain_params* p, int bpc, int w, int h, int subx, int suby, int mono, int mc_identity, const void* const src[3],
                    const size_t src_pitch[3], void* const dst[3], const size_t dst_pitch[3], void* scratch, cudaStream_t s) {
    FgK k;
    const int rc = fg_setup(p, bpc, w, h, subx, suby, mono, mc_identity, scratch, k);
    if (rc) return rc;
    FgDev* st = (FgDev*)scratch;
    FgPlanes pp;
    for (int i = 0; i < 3; i++) {
        pp.src[i] = (const uint8_t*)src[mono ? 0 : i];
        pp.dst[i] = (uint8_t*)dst[mono ? 0 : i];
        pp.spitch[i] = src_pitch[mono ? 0 : i];
        pp.dpitch[i] = dst_pitch[mono ? 0 : i];
    }
    cudaError_t e;
    if (bpc == 8) {
        if (subx && suby) e = launch_apply<uint8_t, 1, 1>(k, st, pp, s);
        else if (subx) e = launch_apply<uint8_t, 1, 0>(k, st, pp, s);
        else e = launch_apply<uint8_t, 0, 0>(k, st, pp, s);
    } else {
        if (subx && suby) e = launch_apply<uint16_t, 1, 1>(k, st, pp, s);
        else if (subx) e = launch_apply<uint16_t, 1, 0>(k, st, pp, s);
        else e = launch_apply<uint16_t, 0, 0>(k, st, pp, s);
    }
    if (e != cudaSuccess) { g_stage_err = cudaGetErrorString(e); return -5; }
    return 0;
}
}  // namespace av1r

extern "C" int av1r_stage_film_grain(const av1r_film_grain_params* p, int bpc, int w, int h, int subx, int suby,
                                     int mono, int mc_identity, const void* const src[3], const size_t src_pitch[3],
                                     void* const dst[3], const size_t dst_pitch[3], void* scratch, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int rc = av1r::fg_launch_prepare(p, bpc, w, h, subx, suby, mono, mc_identity, scratch, s);
    if (rc) return rc;
    return av1r::fg_launch_apply(p, bpc, w, h, subx, suby, mono, mc_identity, src, src_pitch, dst, dst_pitch, scratch, s);
}
