// K2 -- inter prediction (AV1 spec 7.11.3): one CTA per prediction rectangle of the host's InterBlk list.
//
//   translation : separable 8-tap sub-pel filter (regular / smooth / sharp / bilinear, 4-tap variants for
//                 dimensions <= 4), reference coordinates clamped to the visible reference frame, intermediate
//                 rounding InterRound0 = 3, InterRound1 = 11 (single) / 7 (compound)              (7.11.3.3/4)
//   warp        : local (least-squares) or global affine model, 8x8 sub-blocks, 193x8 filter table   (7.11.3.5)
//   compound    : average, distance weights, wedge mask, difference-weighted mask (luma mask reused by chroma)
//                                                                                        (7.11.3.11/12/14/15)
//   OBMC        : blend with predictions formed from the above / left neighbours' motion             (7.11.3.10)
// Inter-intra blocks get their (clipped) inter predictor here; the intra part is blended by the wavefront kernel
// (K3), which is the only stage that may read reconstructed neighbours.  Every block is independent of every
// other block of the frame (it reads reference frames only), so the whole frame is one launch.
// The block is processed in 32x32 tiles so that the working set (two int32 predictions + the 39x32 intermediate)
// stays in 13 KB of shared memory whatever the block size; reference samples come through the read-only path
// (L1/L2 hits: neighbouring blocks share most of their 8-tap support).
// Algorithmic bytes: Rbar * F_inter read + F_inter written.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "../av1_consts.h"
#include "dev_common.cuh"
#include "devframe.h"
#include "inter.h"
#include "../tables/tables_inter.inc"

namespace av1r {

__constant__ int16_t c_subpel[6][16][8];
__constant__ uint8_t c_wedge_codebook[3][16][3];
__constant__ uint8_t c_wedge_signflip[BLOCK_SIZES_ALL][16];
__constant__ uint8_t c_blk_w[BLOCK_SIZES_ALL];
__constant__ uint8_t c_blk_h[BLOCK_SIZES_ALL];
__device__ __align__(16) int16_t d_warped_filter[193][8];
__device__ uint8_t d_obmc_mask[7][64];
__device__ uint8_t d_wedge_master[6][64][64];
static bool g_inter_const_loaded[64] = {false};

static constexpr int IT = 32;                 // tile edge
static constexpr int INTER_THREADS = 128;       // CTA size of the general kernel
static constexpr int INTER_THREADS_SMALL = 64;  // blocks of at most 16x16 luma samples: two warps (half the idle lanes per barrier)

__device__ __forceinline__ int wedge_mask_d(int bsize, int flip, int wedge, int i, int j) {
    const int w = c_blk_w[bsize], h = c_blk_h[bsize];
    const uint8_t* cb = c_wedge_codebook[h > w ? 0 : (h < w ? 1 : 2)][wedge];
    const int dir = cb[0], xoff = 32 - ((cb[1] * w) >> 3), yoff = 32 - ((cb[2] * h) >> 3);
    const int m = d_wedge_master[dir][yoff + i][xoff + j];
    return (flip ^ c_wedge_signflip[bsize][wedge]) ? 64 - m : m;
}

__device__ __forceinline__ int filter_index_d(int type, int len) {
    if (len <= 4) {
        if (type == 0 || type == 2) return 4;
        if (type == 1) return 5;
    }
    return type;
}

static constexpr int RW = IT + 16;            // reference window row stride in samples (tw + 7 used; 96 B rows: 16-byte aligned)
static constexpr int MW = IT + 8;             // row stride of the packed 16-bit intermediate (80 B rows: 16-byte aligned)
struct InterSmem {
    int32_t pred[2][IT * IT];
    // intermediate of the separable filter: int32 [39][32] on the generic / warped paths, int16 [39][MW] on the fast path
    __align__(16) int32_t mid[(IT + 7) * IT];
    __align__(16) uint16_t refwin[(IT + 7) * RW];   // clamped reference samples of the tile's 8-tap support, loaded once
};

// idx / d for idx <= 39 * 39 and d <= 39 as a multiply and a shift (inv = ceil(65536 / d); exact in that range): the tile loops
// below split a linear thread index into (row, column) for every sample
__constant__ uint32_t c_recip16[41];   // ceil(65536 / d), d = 1 .. 40
__device__ __forceinline__ int recip16(int d) { return c_recip16[d]; }
__device__ __forceinline__ int div16(int idx, int inv) { return (idx * inv) >> 16; }

template <typename T>
__device__ __forceinline__ int ld_ref(const uint8_t* base, uint32_t pitch, int x, int y) {
    return (int)__ldg((const T*)(base + (size_t)y * pitch) + x);
}

// ---- fast path of the translational predictor: tiles 8 / 16 / 32 samples wide, a multiple of 4 high -------------------------
// Written for instruction count (the stage is issue-bound, not bandwidth-bound: profiles/r1d_c3_ncu_summary.md):
//   fetch : the (tw + 7) x (th + 7) support comes in with one 32-bit load per lane and row (a warp per row, ten rows in flight per
//           warp); a funnel shift with the neighbour lane's word removes the sub-word start offset, so that the window in shared
//           memory starts at sample ix - 3 as packed 16-bit pairs.  Windows that leave the reference frame take a clamped
//           per-sample path.
//   H pass: a thread turns 16 window samples (two 128-bit shared loads) into 8 outputs and stores them as one 128-bit row of
//           packed int16 (the intermediate fits 16 bits: spec 7.11.3.4).
//   V pass: a thread owns two columns x four rows: eleven 32-bit loads of packed pairs feed 64 multiply-adds.
template <typename T, int NT>
__device__ __forceinline__ void fetch_window_fast(const uint8_t* ref, uint32_t pitch, int lastx, int lasty, int ix, int iy, int ww, int wh,
                                                  uint16_t* win) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x0 = ix - 3, y0 = iy - 3;
    constexpr int SPW = 4 / (int)sizeof(T);                  // samples per 32-bit word: 2 (uint16) or 4 (uint8)
    // interior: every 32-bit word the rows are fetched with lies inside the visible reference frame
    const bool interior = x0 >= 0 && y0 >= 0 && ((x0 + ww - 1) | (SPW - 1)) <= lastx && y0 + wh - 1 <= lasty;
    if (interior) {
        const int xs = x0 & ~(SPW - 1), sh = (x0 - xs) * 8 * (int)sizeof(T);
        const int nwords = ((x0 + ww - 1 - xs) / SPW) + 1;   // <= 21 (uint16) / 11 (uint8)
        const uint8_t* base = ref + (size_t)y0 * pitch + (size_t)xs * sizeof(T);
        constexpr int NW = NT / 32, KMAX = NT == 128 ? 10 : 12;   // 4 warps x 10 rows >= 39; 2 warps x 12 rows >= 23 (small blocks)
        uint32_t w[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; k++) {
            const int r = warp + NW * k;
            w[k] = (r < wh && lane < nwords) ? __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)r * pitch) + lane) : 0u;
        }
#pragma unroll
        for (int k = 0; k < KMAX; k++) {
            const int r = warp + NW * k;
            const uint32_t nxt = __shfl_down_sync(0xffffffffu, w[k], 1);
            const uint32_t v = __funnelshift_r(w[k], nxt, sh);
            if (r < wh && lane < nwords) {
                if (sizeof(T) == 2) {
                    reinterpret_cast<uint32_t*>(win + r * RW)[lane] = v;
                } else {
                    uint2 e;
                    e.x = __byte_perm(v, 0, 0x4140);
                    e.y = __byte_perm(v, 0, 0x4342);
                    reinterpret_cast<uint2*>(win + r * RW)[lane] = e;
                }
            }
        }
    } else {
        for (int r = warp; r < wh; r += NT / 32) {
            const int y = min(max(y0 + r, 0), lasty);
            const T* row = (const T*)(ref + (size_t)y * pitch);
            for (int c = lane; c < ww; c += 32) win[r * RW + c] = (uint16_t)__ldg(row + min(max(x0 + c, 0), lastx));
        }
    }
}

template <typename T, int NT>
__device__ void predict_tile_fast(const uint8_t* ref, uint32_t pitch, int lastx, int lasty, int ix, int iy, const int16_t* fh, const int16_t* fv,
                                  int tw, int th, int round1, InterSmem& sm, int32_t* out) {
    const int ww = tw + 7, wh = th + 7;
    fetch_window_fast<T, NT>(ref, pitch, lastx, lasty, ix, iy, ww, wh, sm.refwin);
    __syncthreads();
    int16_t* mid = reinterpret_cast<int16_t*>(sm.mid);
    {   // horizontal pass: (row, group of 8 columns) per thread
        const int lg = tw == 32 ? 2 : (tw == 16 ? 1 : 0), ng = 1 << lg;
        int f[8];
#pragma unroll
        for (int t = 0; t < 8; t++) f[t] = fh[t];
        for (int idx = threadIdx.x; idx < (wh << lg); idx += NT) {
            const int r = idx >> lg, c0 = (idx & (ng - 1)) << 3;
            const uint4* wp = reinterpret_cast<const uint4*>(sm.refwin + r * RW + c0);
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
            for (int k = 0; k < 4; k++) {
                int s0 = 4, s1 = 4;
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    s0 += f[t] * x[2 * k + t];
                    s1 += f[t] * x[2 * k + 1 + t];
                }
                o[k] = ((uint32_t)(s0 >> 3) & 0xffffu) | ((uint32_t)(s1 >> 3) << 16);
            }
            *reinterpret_cast<uint4*>(mid + r * MW + c0) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
    __syncthreads();
    {   // vertical pass: (column pair, group of 4 rows) per thread
        const int lp = tw == 32 ? 4 : (tw == 16 ? 3 : 2), np = 1 << lp;
        const int rnd = 1 << (round1 - 1);
        int f[8];
#pragma unroll
        for (int t = 0; t < 8; t++) f[t] = fv[t];
        for (int idx = threadIdx.x; idx < ((th >> 2) << lp); idx += NT) {
            const int cp = idx & (np - 1), r0 = (idx >> lp) << 2;
            const uint32_t* mp = reinterpret_cast<const uint32_t*>(mid + r0 * MW) + cp;
            int lo[11], hi[11];
#pragma unroll
            for (int t = 0; t < 11; t++) {
                const uint32_t v = mp[t * (MW / 2)];
                lo[t] = (int)(short)(v & 0xffffu);
                hi[t] = (int)v >> 16;
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                int s0 = rnd, s1 = rnd;
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    s0 += f[t] * lo[k + t];
                    s1 += f[t] * hi[k + t];
                }
                *reinterpret_cast<int2*>(out + (r0 + k) * IT + 2 * cp) = make_int2(s0 >> round1, s1 >> round1);
            }
        }
    }
    __syncthreads();
}

// translational prediction of a tw x th tile; (fx, fy) = 1/16 phases, (ix, iy) = integer reference position of the tile's
// top-left sample.  The (tw + 7) x (th + 7) support is fetched once (coordinates clamped to the visible reference frame, spec
// 7.11.3.4) into shared memory; both filter passes then run out of shared memory.
template <typename T, int NT>
__device__ void predict_tile(const uint8_t* ref, uint32_t pitch, int lastx, int lasty, int ix, int iy, int fx, int fy, int fidx_h, int fidx_v,
                             int tw, int th, int round1, InterSmem& sm, int32_t* out) {
    const int16_t* fh = c_subpel[fidx_h][fx];
    const int16_t* fv = c_subpel[fidx_v][fy];
    const int ww = tw + 7, wh = th + 7;
    if ((tw == 8 || tw == 16 || tw == 32) && (th & 3) == 0) {
        predict_tile_fast<T, NT>(ref, pitch, lastx, lasty, ix, iy, fh, fv, tw, th, round1, sm, out);
        return;
    }
    const int inv_ww = recip16(ww), inv_tw = recip16(tw);
    // four independent loads in flight per thread before the first store (the window fetch is pure L2 / L1 latency)
    for (int base = threadIdx.x; base < ww * wh; base += 4 * NT) {
        int v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int idx = base + u * NT;
            const int r = div16(min(idx, ww * wh - 1), inv_ww), c = min(idx, ww * wh - 1) - r * ww;
            v[u] = ld_ref<T>(ref, pitch, min(max(ix + c - 3, 0), lastx), min(max(iy + r - 3, 0), lasty));
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int idx = base + u * NT;
            if (idx < ww * wh) {
                const int r = div16(idx, inv_ww), c = idx - r * ww;
                sm.refwin[r * RW + c] = (uint16_t)v[u];
            }
        }
    }
    __syncthreads();
    int32_t* mid = sm.mid;
    const int rnd = 1 << (round1 - 1);
    if (((tw | th) & 3) == 0) {
        // register sliding window: a thread produces 4 neighbouring outputs from 11 inputs (32 MACs per 11 shared-memory loads
        // instead of 8 loads per output), horizontally then vertically
        const int gw = tw >> 2, inv_gw = recip16(gw);
        for (int idx = threadIdx.x; idx < wh * gw; idx += NT) {
            const int r = div16(idx, inv_gw), c = (idx - r * gw) << 2;
            const uint16_t* w = sm.refwin + r * RW + c;
            int x[11];
#pragma unroll
            for (int t = 0; t < 11; t++) x[t] = (int)w[t];
#pragma unroll
            for (int o = 0; o < 4; o++) {
                int s = 0;
#pragma unroll
                for (int t = 0; t < 8; t++) s += fh[t] * x[o + t];
                mid[r * IT + c + o] = (s + 4) >> 3;
            }
        }
        __syncthreads();
        const int gh = th >> 2;
        for (int idx = threadIdx.x; idx < gh * tw; idx += NT) {
            const int g = div16(idx, inv_tw), c = idx - g * tw, r = g << 2;
            int x[11];
#pragma unroll
            for (int t = 0; t < 11; t++) x[t] = mid[(r + t) * IT + c];
#pragma unroll
            for (int o = 0; o < 4; o++) {
                int s = 0;
#pragma unroll
                for (int t = 0; t < 8; t++) s += fv[t] * x[o + t];
                out[(r + o) * IT + c] = (s + rnd) >> round1;
            }
        }
        __syncthreads();
        return;
    }
    const int n1 = wh * tw;
    for (int idx = threadIdx.x; idx < n1; idx += NT) {
        const int r = div16(idx, inv_tw), c = idx - r * tw;
        const uint16_t* w = sm.refwin + r * RW + c;
        int s = 0;
#pragma unroll
        for (int t = 0; t < 8; t++) s += fh[t] * (int)w[t];
        mid[r * IT + c] = (s + 4) >> 3;
    }
    __syncthreads();
    const int n2 = th * tw;
    for (int idx = threadIdx.x; idx < n2; idx += NT) {
        const int r = div16(idx, inv_tw), c = idx - r * tw;
        int s = 0;
#pragma unroll
        for (int t = 0; t < 8; t++) s += fv[t] * mid[(r + t) * IT + c];
        out[r * IT + c] = (s + rnd) >> round1;
    }
    __syncthreads();
}

// ---- warped prediction, 16-bit samples: packed dot products ---------------------------------------------------------------------
// The warp filter has a different 8-tap phase for every sample, so the work per sample is fixed: what can shrink is the number of
// instructions around the 8 + 8 multiply-adds.  All 193 x 8 taps fit a signed byte, samples and the horizontal intermediate fit a
// signed 16-bit lane, so a pair of taps times a pair of samples is one IDP.2A (same issue rate as IMAD on sm_100a:
// tools/micro/idp_bench.cu): 4 or 5 per output instead of 8 IMAD + 8 unpacks, with the odd start positions handled by shifting the
// 64-bit tap vector by one byte instead of re-packing samples.  The 15 x 15 support is fetched as 32-bit words (no clamps) when it
// lies inside the reference; the horizontal intermediate is stored transposed so that the vertical pass reads packed row pairs.
__device__ __align__(8) int8_t d_warped_filter8[193][8];

// dot products of 8 taps with 8 consecutive 16-bit samples that start at an even / odd sample of the packed words q[]
__device__ __forceinline__ int dot8_even(uint32_t q0, uint32_t q1, uint32_t q2, uint32_t q3, uint2 taps) {
    int s = __dp2a_lo((int)q0, (int)taps.x, 0);
    s = __dp2a_hi((int)q1, (int)taps.x, s);
    s = __dp2a_lo((int)q2, (int)taps.y, s);
    return __dp2a_hi((int)q3, (int)taps.y, s);
}
__device__ __forceinline__ int dot8_odd(uint32_t q0, uint32_t q1, uint32_t q2, uint32_t q3, uint32_t q4, uint2 taps) {
    const int t0 = (int)(taps.x << 8), t1 = (int)__funnelshift_l(taps.x, taps.y, 8), t2 = (int)(taps.y >> 24);
    int s = __dp2a_lo((int)q0, t0, 0);
    s = __dp2a_hi((int)q1, t0, s);
    s = __dp2a_lo((int)q2, t1, s);
    s = __dp2a_hi((int)q3, t1, s);
    return __dp2a_lo((int)q4, t2, s);
}

// A lane finishes TWO neighbouring outputs of a row (horizontal pass) or of a column (vertical pass): their supports overlap in
// seven of eight samples, so five shared-memory words serve both (ten before), and the parity of the start sample -- which
// decides between the four-word and the five-word dot product -- is the same for every lane of the warp.  The model's matrix and
// shears are read once per tile into registers (two 128-bit loads): through the reference the compiler re-read them from global
// memory in every 8x8 block, because the shared-memory stores in between may alias.
static constexpr int WS = 12;   // words per window row: 8 used; 12 keeps the eight rows a warp reads at once on different banks
template <int NT>
__device__ void warp_tile16(const uint8_t* ref, uint32_t pitch, int lastx, int lasty, int x0, int y0, int sx, int sy, const WarpRec& wr_g, int tw,
                            int th, int round1, InterSmem& sm, int32_t* out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* win = reinterpret_cast<uint32_t*>(sm.refwin) + warp * (15 * WS);   // 15 rows x 8 words
    uint32_t* wmt = reinterpret_cast<uint32_t*>(sm.mid) + warp * 80;             // transposed intermediate: 8 columns x 9 words (15 int16 + pad)
    const int4 ma = __ldg(reinterpret_cast<const int4*>(&wr_g)), mb = __ldg(reinterpret_cast<const int4*>(&wr_g) + 1);
    const int alpha = (int)(short)(mb.z & 0xffff), beta = mb.z >> 16, gamma = (int)(short)(mb.w & 0xffff), delta = mb.w >> 16;
    const int nbx = tw >> 3, lnbx = 31 - __clz(nbx), nb = nbx * (th >> 3);
    const int rnd = 1 << (round1 - 1);
    for (int b = warp; b < nb; b += NT / 32) {
        const int i8 = b >> lnbx, j8 = b & (nbx - 1);
        const int src_x = (x0 + j8 * 8 + 4) << sx, src_y = (y0 + i8 * 8 + 4) << sy;
        const long long dst_x = (long long)ma.z * src_x + (long long)ma.w * src_y + ma.x;
        const long long dst_y = (long long)mb.x * src_x + (long long)mb.y * src_y + ma.y;
        const long long x4 = dst_x >> sx, y4 = dst_y >> sy;
        const int ix4 = (int)(x4 >> 16), sx4 = (int)(x4 & 0xFFFF), iy4 = (int)(y4 >> 16), sy4 = (int)(y4 & 0xFFFF);
        const int xl = ix4 - 7, yt = iy4 - 7;
        int off;                                                     // window column 0 sits at sample `off` of a row's first word
        if (xl >= 0 && yt >= 0 && (xl & ~1) + 15 <= lastx && yt + 14 <= lasty) {   // all eight words of every row inside the frame
            const int xs = xl & ~1;
            off = xl - xs;
            const uint8_t* base = ref + (size_t)yt * pitch + (size_t)xs * 2;
            uint32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int idx = lane + 32 * u, r = idx >> 3, c = idx & 7;
                v[u] = r < 15 ? __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)r * pitch) + c) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int idx = lane + 32 * u, r = idx >> 3, c = idx & 7;
                if (r < 15) win[r * WS + c] = v[u];
            }
        } else {                                                     // leaves the reference frame: clamped samples, packed by pairs
            off = 0;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int idx = lane + 32 * u, r = idx >> 3, c = idx & 7;
                if (r < 15) {
                    const uint16_t* row = (const uint16_t*)(ref + (size_t)min(max(yt + r, 0), lasty) * pitch);
                    const uint32_t a = __ldg(row + min(max(xl + 2 * c, 0), lastx)), bb = __ldg(row + min(max(xl + 2 * c + 1, 0), lastx));
                    win[r * WS + c] = a | (bb << 16);
                }
            }
        }
        __syncwarp();
        int16_t* wmt16 = reinterpret_cast<int16_t*>(wmt);
#pragma unroll
        for (int u = 0; u < 2; u++) {                                // horizontal: 15 rows x 4 column pairs
            const int t = lane + 32 * u;
            if (t < 15 * 4) {
                const int r = t >> 2, c = (t & 3) << 1;              // i1 = r - 7, i2 = c - 4 (and c - 3)
                const int sxa = sx4 + alpha * (c - 4) + beta * (r - 7);
                const uint2 ta = __ldg(reinterpret_cast<const uint2*>(d_warped_filter8[((sxa + 512) >> 10) + 64]));
                const uint2 tb = __ldg(reinterpret_cast<const uint2*>(d_warped_filter8[((sxa + alpha + 512) >> 10) + 64]));
                const uint32_t* q = win + r * WS + (c >> 1);
                const uint32_t q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3], q4 = q[4];
                int sa, sb;
                if (off == 0) {                                      // warp-uniform
                    sa = dot8_even(q0, q1, q2, q3, ta);
                    sb = dot8_odd(q0, q1, q2, q3, q4, tb);
                } else {
                    sa = dot8_odd(q0, q1, q2, q3, q4, ta);
                    sb = dot8_even(q1, q2, q3, q4, tb);
                }
                wmt16[c * 18 + r] = (int16_t)((sa + 4) >> 3);
                wmt16[c * 18 + 18 + r] = (int16_t)((sb + 4) >> 3);
            }
        }
        __syncwarp();
        {                                                            // vertical: 8 columns x 4 row pairs, one pair per lane
            const int c = lane & 7, orow = (lane >> 3) << 1;         // i1 = orow - 4 (and orow - 3), i2 = c - 4
            const int sya = sy4 + gamma * (c - 4) + delta * (orow - 4);
            const uint2 ta = __ldg(reinterpret_cast<const uint2*>(d_warped_filter8[((sya + 512) >> 10) + 64]));
            const uint2 tb = __ldg(reinterpret_cast<const uint2*>(d_warped_filter8[((sya + delta + 512) >> 10) + 64]));
            const uint32_t* q = wmt + c * 9 + (orow >> 1);
            const uint32_t q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3], q4 = q[4];
            const int sa = dot8_even(q0, q1, q2, q3, ta), sb = dot8_odd(q0, q1, q2, q3, q4, tb);
            int32_t* o = out + (i8 * 8 + orow) * IT + j8 * 8 + c;
            o[0] = (sa + rnd) >> round1;
            o[IT] = (sb + rnd) >> round1;
        }
        __syncwarp();
    }
    __syncthreads();
}

// ---- scaled references (spec 7.11.3.3 motion vector scaling + 7.11.3.4 block inter prediction) --------------------------------
// The reference has another size than the frame (spatial resize, or super-resolution on an inter frame): sample positions advance
// in 1/1024 steps of xStep / yStep from a start position computed once per *block* (the rounding of the start position and of the
// steps is part of the normative result, so tiles of a block derive their positions from the block origin, not from their own).
// Rare in practice (the daemon's encoder never resizes): a plain per-sample implementation, reference samples straight from L2.
__device__ __forceinline__ long long round2s64_d(long long x, int n) {
    return x >= 0 ? (x + (1ll << (n - 1))) >> n : -((-x + (1ll << (n - 1))) >> n);
}
template <typename T, int NT>
__device__ void predict_tile_scaled(const uint8_t* ref, uint32_t pitch, int lastx, int lasty, int bx, int by, int mv_row, int mv_col, int sx, int sy,
                                    int xscale, int yscale, int fidx_h, int fidx_v, int tx, int ty, int tw, int th, int round1, InterSmem& sm,
                                    int32_t* out) {
    const long long origx = ((long long)bx << 4) + ((2 * mv_col) >> sx) + 8, origy = ((long long)by << 4) + ((2 * mv_row) >> sy) + 8;
    const int startx = (int)(round2s64_d(origx * xscale - (8ll << 14), 8) + 32), starty = (int)(round2s64_d(origy * yscale - (8ll << 14), 8) + 32);
    const int stepx = (int)round2s64_d(xscale, 4), stepy = (int)round2s64_d(yscale, 4);
    const int fy0 = starty & 1023;
    const int rnd = 1 << (round1 - 1);
    int32_t* mid = sm.mid;
    // 16 output rows at a time: at most ((15 * 2048 + 1023) >> 10) + 8 = 38 intermediate rows, which fit the 39-row scratch
    for (int sub = 0; sub < th; sub += 16) {
        const int sh = min(16, th - sub);
        const int row0 = (fy0 + stepy * (ty + sub)) >> 10;
        const int nrows = ((fy0 + stepy * (ty + sub + sh - 1)) >> 10) + 8 - row0;
        for (int idx = threadIdx.x; idx < nrows * tw; idx += NT) {
            const int r = idx / tw, c = idx - r * tw;
            const int p = startx + stepx * (tx + c);
            const int16_t* f = c_subpel[fidx_h][(p >> 6) & 15];
            const T* row = (const T*)(ref + (size_t)min(max((starty >> 10) + row0 + r - 3, 0), lasty) * pitch);
            const int x0 = (p >> 10) - 3;
            int s = 0;
#pragma unroll
            for (int t = 0; t < 8; t++) s += f[t] * (int)__ldg(row + min(max(x0 + t, 0), lastx));
            mid[r * IT + c] = (s + 4) >> 3;
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < sh * tw; idx += NT) {
            const int r = idx / tw, c = idx - r * tw;
            const int p = fy0 + stepy * (ty + sub + r);
            const int16_t* f = c_subpel[fidx_v][(p >> 6) & 15];
            const int base = (p >> 10) - row0;
            int s = 0;
#pragma unroll
            for (int t = 0; t < 8; t++) s += f[t] * mid[(base + t) * IT + c];
            out[(sub + r) * IT + c] = (s + rnd) >> round1;
        }
        __syncthreads();
    }
}

// warped prediction of a tile (multiples of 8): one warp per 8x8 sub-block
template <typename T, int NT>
__device__ void warp_tile(const uint8_t* ref, uint32_t pitch, int lastx, int lasty, int x0, int y0, int sx, int sy, const WarpRec& wr, int tw, int th,
                          int round1, InterSmem& sm, int32_t* out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int32_t* wm = sm.mid + warp * 128;
    uint16_t* win = sm.refwin + warp * 240;   // 15 x 15 support of one 8x8 block (16-sample rows)
    const int nbx = tw >> 3, nb = nbx * (th >> 3);
    const int rnd = 1 << (round1 - 1);
    for (int b = warp; b < nb; b += NT / 32) {
        const int i8 = b / nbx, j8 = b - i8 * nbx;
        const int src_x = (x0 + j8 * 8 + 4) << sx, src_y = (y0 + i8 * 8 + 4) << sy;
        const long long dst_x = (long long)wr.mat[2] * src_x + (long long)wr.mat[3] * src_y + wr.mat[0];
        const long long dst_y = (long long)wr.mat[4] * src_x + (long long)wr.mat[5] * src_y + wr.mat[1];
        const long long x4 = dst_x >> sx, y4 = dst_y >> sy;
        const int ix4 = (int)(x4 >> 16), sx4 = (int)(x4 & 0xFFFF), iy4 = (int)(y4 >> 16), sy4 = (int)(y4 & 0xFFFF);
        {   // 15 x 15 support: eight loads in flight per lane before the stores
            int v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int idx = min(lane + 32 * u, 15 * 15 - 1);
                const int wr_ = (idx * 4370) >> 16, wc = idx - wr_ * 15;   // idx / 15 for idx < 225
                v[u] = ld_ref<T>(ref, pitch, min(max(ix4 + wc - 7, 0), lastx), min(max(iy4 + wr_ - 7, 0), lasty));
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int idx = lane + 32 * u;
                if (idx < 15 * 15) {
                    const int wr_ = (idx * 4370) >> 16, wc = idx - wr_ * 15;
                    win[wr_ * 16 + wc] = (uint16_t)v[u];
                }
            }
        }
        __syncwarp();
        for (int idx = lane; idx < 15 * 8; idx += 32) {
            const int i1 = (idx >> 3) - 7, i2 = (idx & 7) - 4;
            const int sxx = sx4 + wr.alpha * i2 + wr.beta * i1;
            const int offs = ((sxx + 512) >> 10) + 64;
            const uint16_t* w = win + (i1 + 7) * 16 + (i2 - 3 + 7);
            // the eight taps of one filter phase are one 16-byte row of the table: a single vector load instead of eight
            const uint4 fq = __ldg(reinterpret_cast<const uint4*>(d_warped_filter[offs]));
            const int f[8] = {(int16_t)(fq.x & 0xffff), (int16_t)(fq.x >> 16), (int16_t)(fq.y & 0xffff), (int16_t)(fq.y >> 16),
                              (int16_t)(fq.z & 0xffff), (int16_t)(fq.z >> 16), (int16_t)(fq.w & 0xffff), (int16_t)(fq.w >> 16)};
            int s = 0;
#pragma unroll
            for (int i3 = 0; i3 < 8; i3++) s += f[i3] * (int)w[i3];
            wm[idx] = (s + 4) >> 3;
        }
        __syncwarp();
        for (int idx = lane; idx < 64; idx += 32) {
            const int i1 = (idx >> 3) - 4, i2 = (idx & 7) - 4;
            const int syy = sy4 + wr.gamma * i2 + wr.delta * i1;
            const int offs = ((syy + 512) >> 10) + 64;
            const uint4 fq = __ldg(reinterpret_cast<const uint4*>(d_warped_filter[offs]));
            const int f[8] = {(int16_t)(fq.x & 0xffff), (int16_t)(fq.x >> 16), (int16_t)(fq.y & 0xffff), (int16_t)(fq.y >> 16),
                              (int16_t)(fq.z & 0xffff), (int16_t)(fq.z >> 16), (int16_t)(fq.w & 0xffff), (int16_t)(fq.w >> 16)};
            int s = 0;
#pragma unroll
            for (int i3 = 0; i3 < 8; i3++) s += f[i3] * wm[(i1 + i3 + 4) * 8 + i2 + 4];
            out[(i8 * 8 + i1 + 4) * IT + j8 * 8 + i2 + 4] = (s + rnd) >> round1;
        }
        __syncwarp();
    }
    __syncthreads();
}

template <typename T, int NT>
// (8 / 12 CTAs per SM; 10 and 12 / 8 and 16 were measured and change nothing: the kernel is not occupancy-limited, profiles/r2_work_order.md)
__global__ void __launch_bounds__(NT, NT == 128 ? 8 : 12) inter_pred_kernel(InterLaunch L, int first_item) {
    __shared__ InterSmem sm;
    const uint32_t item = L.tiles[first_item + blockIdx.x];
    const InterBlk r = L.blks[item & 0x0fffffffu];
    const int qx0 = ((item >> 28) & 1) << 6, qy0 = ((item >> 29) & 1) << 6;   // luma origin of this CTA's 64x64 quadrant inside the block
    const DevFrameParams& fp = L.fp;
    const int pixmax = (1 << fp.bd) - 1;
    const int is_compound = r.ref[1] >= 0;
    const int round1 = is_compound ? 7 : 11, post = 11 - round1;
    // difference-weighted masks live in a frame-sized luma plane in global memory (written by this CTA's luma pass, read by its
    // chroma passes after the barrier + fence below): rare, so it does not deserve 16 KB of shared memory per CTA
    uint8_t* gmask = L.mask + (size_t)r.y * L.mask_pitch + r.x;
    for (int plane = 0; plane < 3; plane++) {
        if (plane == 0 && !(r.planes & 1)) continue;
        if (plane > 0 && !(r.planes & 2)) continue;
        const int sx = plane ? fp.subx : 0, sy = plane ? fp.suby : 0;
        const int px = r.x >> sx, py = r.y >> sy, pw = r.w >> sx, ph = r.h >> sy;
        T* cur = (T*)L.cur.p[plane];
        const int cpe = L.cur.pitch[plane] / sizeof(T);
        const int rx0 = qx0 >> sx, ry0 = qy0 >> sy, rx1 = min(pw, (qx0 + 64) >> sx), ry1 = min(ph, (qy0 + 64) >> sy);
        for (int ty = ry0; ty < ry1; ty += IT)
            for (int tx = rx0; tx < rx1; tx += IT) {
                const int tw = min(IT, pw - tx), th = min(IT, ph - ty);
                for (int l = 0; l < 1 + is_compound; l++) {
                    const DevPlanes& rf = L.refs[r.ref[l]];
                    const int lastx = L.ref_w[r.ref[l]][plane] - 1, lasty = L.ref_h[r.ref[l]][plane] - 1;
                    if (L.xscale[r.ref[l]] != (1 << 14) || L.yscale[r.ref[l]] != (1 << 14)) {
                        predict_tile_scaled<T, NT>(rf.p[plane], rf.pitch[plane], lastx, lasty, px, py, r.mv[l][0], r.mv[l][1], sx, sy, L.xscale[r.ref[l]],
                                                   L.yscale[r.ref[l]], filter_index_d(r.filt[1], pw), filter_index_d(r.filt[0], ph), tx, ty, tw, th, round1,
                                                   sm, sm.pred[l]);
                    } else if (r.warp[l] >= 0 && pw >= 8 && ph >= 8) {
                        if (sizeof(T) == 2)
                            warp_tile16<NT>(rf.p[plane], rf.pitch[plane], lastx, lasty, px + tx, py + ty, sx, sy, L.warps[r.warp[l]], tw, th, round1, sm,
                                        sm.pred[l]);
                        else
                            warp_tile<T, NT>(rf.p[plane], rf.pitch[plane], lastx, lasty, px + tx, py + ty, sx, sy, L.warps[r.warp[l]], tw, th, round1, sm,
                                         sm.pred[l]);
                    } else {
                        const int posx = ((px + tx) << 4) + ((2 * r.mv[l][1]) >> sx), posy = ((py + ty) << 4) + ((2 * r.mv[l][0]) >> sy);
                        predict_tile<T, NT>(rf.p[plane], rf.pitch[plane], lastx, lasty, posx >> 4, posy >> 4, posx & 15, posy & 15,
                                        filter_index_d(r.filt[1], pw), filter_index_d(r.filt[0], ph), tw, th, round1, sm, sm.pred[l]);
                    }
                }
                const bool masked = is_compound && (r.comp_type == COMPOUND_WEDGE || r.comp_type == COMPOUND_DIFFWTD);
                if (!masked && (tw == 8 || tw == 16 || tw == 32)) {
                    // unmasked predictions (single, average, distance weights): a thread finishes 8 neighbouring samples of a row
                    // and stores them with one 128-bit (64-bit at 8 bits per sample) store where the row position allows it
                    const int lg = tw == 32 ? 2 : (tw == 16 ? 1 : 0);
                    const int wa = !is_compound ? 1 : (r.comp_type == COMPOUND_DISTANCE ? r.fwd_w : 1);
                    const int wb = !is_compound ? 0 : (r.comp_type == COMPOUND_DISTANCE ? r.bck_w : 1);
                    const int sh = !is_compound ? 0 : (r.comp_type == COMPOUND_DISTANCE ? 4 + post : 1 + post);
                    const int rnd = sh ? 1 << (sh - 1) : 0;
                    const int cwp = fp.cw[plane], chp = fp.ch[plane];
                    for (int idx = threadIdx.x; idx < (th << lg); idx += NT) {
                        const int i = idx >> lg, j0 = (idx & ((1 << lg) - 1)) << 3;
                        const int gx = px + tx + j0, gy = py + ty + i;
                        if (gy >= chp || gx >= cwp) continue;
                        const int4* pa = reinterpret_cast<const int4*>(sm.pred[0] + i * IT + j0);
                        const int4 a0 = pa[0], a1 = pa[1];
                        int v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                        if (is_compound) {
                            const int4* pb = reinterpret_cast<const int4*>(sm.pred[1] + i * IT + j0);
                            const int4 b0 = pb[0], b1 = pb[1];
                            const int b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                            for (int k = 0; k < 8; k++) v[k] = (v[k] * wa + b[k] * wb + rnd) >> sh;
                        }
#pragma unroll
                        for (int k = 0; k < 8; k++) v[k] = min(max(v[k], 0), pixmax);
                        T* dp = cur + (size_t)gy * cpe + gx;
                        if (sizeof(T) == 2) {
                            const uint32_t w0 = (uint32_t)v[0] | ((uint32_t)v[1] << 16), w1 = (uint32_t)v[2] | ((uint32_t)v[3] << 16);
                            const uint32_t w2 = (uint32_t)v[4] | ((uint32_t)v[5] << 16), w3 = (uint32_t)v[6] | ((uint32_t)v[7] << 16);
                            if (gx + 8 <= cwp && (gx & 7) == 0) {
                                *reinterpret_cast<uint4*>(dp) = make_uint4(w0, w1, w2, w3);
                            } else {   // rows of chroma blocks start at even samples only: 32-bit stores, pair by pair
                                uint32_t* d32 = reinterpret_cast<uint32_t*>(dp);
                                d32[0] = w0;
                                if (gx + 2 < cwp) d32[1] = w1;
                                if (gx + 4 < cwp) d32[2] = w2;
                                if (gx + 6 < cwp) d32[3] = w3;
                            }
                        } else {
                            const uint32_t w0 = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
                            const uint32_t w1 = (uint32_t)v[4] | ((uint32_t)v[5] << 8) | ((uint32_t)v[6] << 16) | ((uint32_t)v[7] << 24);
                            if (gx + 8 <= cwp && (gx & 7) == 0) {
                                *reinterpret_cast<uint2*>(dp) = make_uint2(w0, w1);
                            } else {
                                uint16_t* d16 = reinterpret_cast<uint16_t*>(dp);
                                d16[0] = (uint16_t)w0;
                                if (gx + 2 < cwp) d16[1] = (uint16_t)(w0 >> 16);
                                if (gx + 4 < cwp) d16[2] = (uint16_t)w1;
                                if (gx + 6 < cwp) d16[3] = (uint16_t)(w1 >> 16);
                            }
                        }
                    }
                    __syncthreads();
                    continue;
                }
                const int inv_tw = recip16(tw);
                for (int idx = threadIdx.x; idx < tw * th; idx += NT) {
                    const int i = div16(idx, inv_tw), j = idx - i * tw;
                    const int gx = px + tx + j, gy = py + ty + i;
                    const int a = sm.pred[0][i * IT + j];
                    int v;
                    if (!is_compound) {
                        v = a;
                    } else {
                        const int b = sm.pred[1][i * IT + j];
                        if (r.comp_type == COMPOUND_WEDGE || r.comp_type == COMPOUND_DIFFWTD) {
                            const int bi = ty + i, bj = tx + j;
                            int m;
                            if (r.comp_type == COMPOUND_DIFFWTD) {
                                if (plane == 0) {
                                    int diff = abs(a - b);
                                    diff = d_round2(diff, (fp.bd - 8) + post);
                                    m = min(max(38 + diff / 16, 0), 64);
                                    if (r.mask_type) m = 64 - m;
                                    gmask[(size_t)bi * L.mask_pitch + bj] = (uint8_t)m;
                                } else if (sx && sy) {
                                    m = (__ldcg(gmask + (size_t)(2 * bi) * L.mask_pitch + 2 * bj) + __ldcg(gmask + (size_t)(2 * bi) * L.mask_pitch + 2 * bj + 1) +
                                         __ldcg(gmask + (size_t)(2 * bi + 1) * L.mask_pitch + 2 * bj) +
                                         __ldcg(gmask + (size_t)(2 * bi + 1) * L.mask_pitch + 2 * bj + 1) + 2) >> 2;
                                } else if (sx) {
                                    m = (__ldcg(gmask + (size_t)bi * L.mask_pitch + 2 * bj) + __ldcg(gmask + (size_t)bi * L.mask_pitch + 2 * bj + 1) + 1) >> 1;
                                } else {
                                    m = __ldcg(gmask + (size_t)bi * L.mask_pitch + bj);
                                }
                            } else {
                                if (!sx && !sy) m = wedge_mask_d(r.bsize, r.wedge_sign, r.wedge_index, bi, bj);
                                else if (sx && !sy)
                                    m = (wedge_mask_d(r.bsize, r.wedge_sign, r.wedge_index, bi, 2 * bj) +
                                         wedge_mask_d(r.bsize, r.wedge_sign, r.wedge_index, bi, 2 * bj + 1) + 1) >> 1;
                                else
                                    m = (wedge_mask_d(r.bsize, r.wedge_sign, r.wedge_index, 2 * bi, 2 * bj) +
                                         wedge_mask_d(r.bsize, r.wedge_sign, r.wedge_index, 2 * bi, 2 * bj + 1) +
                                         wedge_mask_d(r.bsize, r.wedge_sign, r.wedge_index, 2 * bi + 1, 2 * bj) +
                                         wedge_mask_d(r.bsize, r.wedge_sign, r.wedge_index, 2 * bi + 1, 2 * bj + 1) + 2) >> 2;
                            }
                            v = d_round2(m * a + (64 - m) * b, 6 + post);
                        } else if (r.comp_type == COMPOUND_DISTANCE) {
                            v = d_round2(a * r.fwd_w + b * r.bck_w, 4 + post);
                        } else {
                            v = d_round2(a + b, 1 + post);
                        }
                    }
                    if (gx < fp.cw[plane] && gy < fp.ch[plane]) cur[(size_t)gy * cpe + gx] = (T)min(max(v, 0), pixmax);
                }
                __syncthreads();
            }
        if (plane == 0 && is_compound && r.comp_type == COMPOUND_DIFFWTD) {
            __threadfence_block();
            __syncthreads();
        }
        // ---- overlapped motion compensation: blend with the neighbours' predictions, above pass then left pass
        const int n_nb = r.obmc_above + r.obmc_left;
        for (int k = 0; k < n_nb; k++) {
            const int above = k < r.obmc_above;
            if (above && plane > 0 && !r.obmc_chroma_above) continue;
            const ObmcNb nb = L.obmc[r.obmc_first + k];
            int ow, oh;
            if (above) {
                ow = min(pw, (nb.step4 * 4) >> sx);
                oh = min(ph >> 1, 32 >> sy);
            } else {
                ow = min(pw >> 1, 32 >> sx);
                oh = min(ph, (nb.step4 * 4) >> sy);
            }
            const int ox = (nb.x4 * 4) >> sx, oy = (nb.y4 * 4) >> sy;
            const DevPlanes& rf = L.refs[nb.ref];
            const int lastx = L.ref_w[nb.ref][plane] - 1, lasty = L.ref_h[nb.ref][plane] - 1;
            const bool nb_scaled = L.xscale[nb.ref] != (1 << 14) || L.yscale[nb.ref] != (1 << 14);
            const uint8_t* m = d_obmc_mask[31 - __clz(above ? oh : ow)];
            // absolute plane rectangle of this CTA's quadrant: only overlap samples inside it are blended here
            const int ax0 = px + rx0, ay0 = py + ry0, ax1 = px + rx1, ay1 = py + ry1;
            for (int ty = 0; ty < oh; ty += IT)
                for (int tx = 0; tx < ow; tx += IT) {
                    const int tw = min(IT, ow - tx), th = min(IT, oh - ty);
                    if (ox + tx >= ax1 || ox + tx + tw <= ax0 || oy + ty >= ay1 || oy + ty + th <= ay0) continue;   // CTA-uniform
                    const int posx = ((ox + tx) << 4) + ((2 * nb.mv[1]) >> sx), posy = ((oy + ty) << 4) + ((2 * nb.mv[0]) >> sy);
                    if (nb_scaled)
                        predict_tile_scaled<T, NT>(rf.p[plane], rf.pitch[plane], lastx, lasty, ox, oy, nb.mv[0], nb.mv[1], sx, sy, L.xscale[nb.ref],
                                                   L.yscale[nb.ref], filter_index_d(nb.filt[1], ow), filter_index_d(nb.filt[0], oh), tx, ty, tw, th, 11, sm,
                                                   sm.pred[0]);
                    else
                        predict_tile<T, NT>(rf.p[plane], rf.pitch[plane], lastx, lasty, posx >> 4, posy >> 4, posx & 15, posy & 15,
                                            filter_index_d(nb.filt[1], ow), filter_index_d(nb.filt[0], oh), tw, th, 11, sm, sm.pred[0]);
                    const int inv_tw = recip16(tw);
                    for (int idx = threadIdx.x; idx < tw * th; idx += NT) {
                        const int i = div16(idx, inv_tw), j = idx - i * tw;
                        const int gx = ox + tx + j, gy = oy + ty + i;
                        if (gx >= fp.cw[plane] || gy >= fp.ch[plane] || gx < ax0 || gx >= ax1 || gy < ay0 || gy >= ay1) continue;
                        const int o = min(max(sm.pred[0][i * IT + j], 0), pixmax);
                        const int mm = above ? m[ty + i] : m[tx + j];
                        T* p = cur + (size_t)gy * cpe + gx;
                        *p = (T)((mm * (int)*p + (64 - mm) * o + 32) >> 6);
                    }
                    __syncthreads();
                }
        }
    }
}

// residual of plain inter transform blocks: frame += residual (records with mode TXM_INTER that are not part of an
// inter-intra block; those wait for the blend in K3).  One warp per record.
template <typename T>
__global__ void __launch_bounds__(128) inter_residual_kernel(const TxRec* recs, const uint32_t* order, int n, DevPlanes cur, DevResidual res,
                                                             DevFrameParams fp) {
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= n) return;
    const TxRec r = recs[order[wid]];
    if (r.mode != TXM_INTER || (r.flags & TXF_II)) return;
    static const uint8_t kW[TX_SIZES_ALL] = {2, 3, 4, 5, 6, 2, 3, 3, 4, 4, 5, 5, 6, 2, 4, 3, 5, 4, 6};
    static const uint8_t kH[TX_SIZES_ALL] = {2, 3, 4, 5, 6, 3, 2, 4, 3, 5, 4, 6, 5, 4, 2, 5, 3, 6, 4};
    const int plane = r.plane, lw = kW[r.txsz], w = 1 << lw, h = 1 << kH[r.txsz];
    const int x = r.x4 * 4, y = r.y4 * 4;
    const int xe = min(w, fp.cw[plane] - x), ye = min(h, fp.ch[plane] - y);
    const int pixmax = (1 << fp.bd) - 1;
    T* out = (T*)(cur.p[plane] + (size_t)y * cur.pitch[plane]) + x;
    const int ope = cur.pitch[plane] / sizeof(T);
    const int16_t* rp = res_ptr(res, plane, x, y);
    const int rpe = 1 << res.tw_log2[plane];
    if (w >= 8 && xe == w) {
        // transform blocks are aligned to their width: rows of 8-sample groups move as 128-bit residual / pixel vectors
        const int lg = lw - 3;
        for (int idx = lane; idx < (ye << lg); idx += 32) {
            const int i = idx >> lg, j = (idx & ((1 << lg) - 1)) << 3;
            const uint4 rv = *reinterpret_cast<const uint4*>(rp + i * rpe + j);
            const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
            T* op = out + i * ope + j;
            int px[8];
            if (sizeof(T) == 2) {
                const uint4 pv = *reinterpret_cast<const uint4*>(op);
                const uint32_t pw_[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    px[2 * k] = (int)(pw_[k] & 0xffffu);
                    px[2 * k + 1] = (int)(pw_[k] >> 16);
                }
            } else {
                const uint2 pv = *reinterpret_cast<const uint2*>(op);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    px[k] = (int)((pv.x >> (8 * k)) & 0xffu);
                    px[4 + k] = (int)((pv.y >> (8 * k)) & 0xffu);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                px[2 * k] = min(max(px[2 * k] + (int)(short)(rw[k] & 0xffffu), 0), pixmax);
                px[2 * k + 1] = min(max(px[2 * k + 1] + ((int)rw[k] >> 16), 0), pixmax);
            }
            if (sizeof(T) == 2) {
                *reinterpret_cast<uint4*>(op) = make_uint4((uint32_t)px[0] | ((uint32_t)px[1] << 16), (uint32_t)px[2] | ((uint32_t)px[3] << 16),
                                                           (uint32_t)px[4] | ((uint32_t)px[5] << 16), (uint32_t)px[6] | ((uint32_t)px[7] << 16));
            } else {
                *reinterpret_cast<uint2*>(op) = make_uint2((uint32_t)px[0] | ((uint32_t)px[1] << 8) | ((uint32_t)px[2] << 16) | ((uint32_t)px[3] << 24),
                                                           (uint32_t)px[4] | ((uint32_t)px[5] << 8) | ((uint32_t)px[6] << 16) | ((uint32_t)px[7] << 24));
            }
        }
        return;
    }
    for (int idx = lane; idx < w * h; idx += 32) {
        const int i = idx >> lw, j = idx & (w - 1);
        if (i < ye && j < xe) {
            const int v = (int)out[i * ope + j] + (int)rp[i * rpe + j];
            out[i * ope + j] = (T)min(max(v, 0), pixmax);
        }
    }
}

static cudaError_t inter_upload_constants() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && g_inter_const_loaded[dev]) return cudaSuccess;
    if ((e = cudaMemcpyToSymbol(c_subpel, av1t_subpel_filters, sizeof(av1t_subpel_filters))) != cudaSuccess) return e;
    {
        uint32_t rc[41] = {0, 65536};
        for (int d = 2; d <= 40; d++) rc[d] = (uint32_t)((65536 + d - 1) / d);
        if ((e = cudaMemcpyToSymbol(c_recip16, rc, sizeof(rc))) != cudaSuccess) return e;
    }
    if ((e = cudaMemcpyToSymbol(c_wedge_codebook, av1t_wedge_codebook, sizeof(av1t_wedge_codebook))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_wedge_signflip, av1t_wedge_signflip, sizeof(av1t_wedge_signflip))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_blk_w, kBlockW, sizeof(kBlockW))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_blk_h, kBlockH, sizeof(kBlockH))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(d_warped_filter, av1t_warped_filter, sizeof(av1t_warped_filter))) != cudaSuccess) return e;
    {
        static int8_t f8[193][8];
        for (int i = 0; i < 193; i++)
            for (int t = 0; t < 8; t++) f8[i][t] = (int8_t)av1t_warped_filter[i][t];   // all taps lie in -22 .. 127
        if ((e = cudaMemcpyToSymbol(d_warped_filter8, f8, sizeof(f8))) != cudaSuccess) return e;
    }
    if ((e = cudaMemcpyToSymbol(d_obmc_mask, av1t_obmc_mask, sizeof(av1t_obmc_mask))) != cudaSuccess) return e;
    // wedge master masks (spec 7.11.3.11)
    static uint8_t master[6][64][64];
    enum { WH = 0, WV = 1, W27 = 2, W63 = 3, W117 = 4, W153 = 5 };
    for (int j = 0; j < 64; j++) {
        int shift = 16;
        for (int i = 0; i < 64; i += 2) {
            master[W63][i][j] = av1t_wedge_master_oblique_even[std::min(std::max(j - shift, 0), 63)];
            shift--;
            master[W63][i + 1][j] = av1t_wedge_master_oblique_odd[std::min(std::max(j - shift, 0), 63)];
            master[WV][i][j] = master[WV][i + 1][j] = av1t_wedge_master_vertical[j];
        }
    }
    for (int i = 0; i < 64; i++)
        for (int j = 0; j < 64; j++) {
            const int msk = master[W63][i][j];
            master[W27][j][i] = (uint8_t)msk;
            master[W117][i][63 - j] = (uint8_t)(64 - msk);
            master[W153][63 - j][i] = (uint8_t)(64 - msk);
            master[WH][j][i] = master[WV][i][j];
        }
    if ((e = cudaMemcpyToSymbol(d_wedge_master, master, sizeof(master))) != cudaSuccess) return e;
    if (dev < 64) g_inter_const_loaded[dev] = true;
    return cudaSuccess;
}

cudaError_t inter_copy_wedge_master(uint8_t* dst_dev, cudaStream_t s) {
    cudaError_t e = inter_upload_constants();
    if (e != cudaSuccess) return e;
    void* src = nullptr;
    if ((e = cudaGetSymbolAddress(&src, d_wedge_master)) != cudaSuccess) return e;
    return cudaMemcpyAsync(dst_dev, src, 6 * 64 * 64, cudaMemcpyDeviceToDevice, s);
}

cudaError_t launch_inter(const InterLaunch& L, cudaStream_t s) {
    if (L.n <= 0 || L.n_tiles <= 0) return cudaSuccess;
    cudaError_t e = inter_upload_constants();
    if (e != cudaSuccess) return e;
    {
        static bool carve_done = false;
        if (!carve_done) {
            prefer_max_smem(inter_pred_kernel<uint8_t, INTER_THREADS>);
            prefer_max_smem(inter_pred_kernel<uint16_t, INTER_THREADS>);
            prefer_max_smem(inter_pred_kernel<uint8_t, INTER_THREADS_SMALL>);
            prefer_max_smem(inter_pred_kernel<uint16_t, INTER_THREADS_SMALL>);
            prefer_max_smem(inter_residual_kernel<uint8_t>);
            prefer_max_smem(inter_residual_kernel<uint16_t>);
            carve_done = true;
        }
    }
    // the host lists the work items of small blocks (at most 16x16 luma samples) first: they run as two-warp CTAs
    const int n_small = L.n_tiles_small, n_large = L.n_tiles - L.n_tiles_small;
    // the large-block launch first: its heaviest CTAs (listed first) are what the stage ends on
    if (n_large > 0) {
        if (L.fp.bd == 8) inter_pred_kernel<uint8_t, INTER_THREADS><<<n_large, INTER_THREADS, 0, s>>>(L, n_small);
        else inter_pred_kernel<uint16_t, INTER_THREADS><<<n_large, INTER_THREADS, 0, s>>>(L, n_small);
    }
    if (n_small > 0) {
        if (L.fp.bd == 8) inter_pred_kernel<uint8_t, INTER_THREADS_SMALL><<<n_small, INTER_THREADS_SMALL, 0, s>>>(L, 0);
        else inter_pred_kernel<uint16_t, INTER_THREADS_SMALL><<<n_small, INTER_THREADS_SMALL, 0, s>>>(L, 0);
    }
    return cudaGetLastError();
}

cudaError_t launch_inter_residual(const TxRec* recs, const uint32_t* order, int n, const DevPlanes& cur, const DevResidual& res,
                                  const DevFrameParams& fp, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    const int blocks = (n + 3) / 4;
    if (fp.bd == 8) inter_residual_kernel<uint8_t><<<blocks, 128, 0, s>>>(recs, order, n, cur, res, fp);
    else inter_residual_kernel<uint16_t><<<blocks, 128, 0, s>>>(recs, order, n, cur, res, fp);
    return cudaGetLastError();
}

}  // namespace av1r
