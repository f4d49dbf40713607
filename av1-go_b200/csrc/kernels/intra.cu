// K3 -- intra prediction + residual add as a *record-level dataflow* kernel (AV1 spec 7.11.2, 7.11.4, 7.11.5, 7.12.3).
//
// Intra prediction reads reconstructed neighbours, so transform blocks form a dependency DAG: a block needs the blocks that
// own its above row (plus above-right when the mode looks there), its left column (plus below-left) and, for CfL, the luma
// blocks under it.  Walking the blocks of a superblock row in decode order (v1-v4 of this kernel) serialises ~5000 blocks per
// row although the DAG is only a few hundred blocks deep (a 2:1 wavefront at *block* granularity).  v5 therefore gives every
// record to its own warp:
//   1. `wmap_scatter_kernel` writes, for every 4x4 cell of every plane, the position (in K3 order) of the record that
//      reconstructs it (atomicMax: for inter-intra blocks the residual record wins over the blend record).
//   2. `intra_dataflow_kernel`: persistent warps claim records in decode order through an atomic ticket, look up the owners of
//      the edge cells they are about to read, spin (with back-off) on those owners' done-flags, predict, add the residual,
//      store, fence and raise their own flag.  A warp only ever waits for lower tickets, which are held by resident warps, so
//      the lowest unfinished record can always run: no deadlock, no host-built dependency lists.
// Edge samples come from L2 (ld.global.cg; the frame is being written by other SMs), the residual from K1's unit-major
// buffer.  The critical path is the DAG depth (W/bw + 2H/bh blocks) times one L2 round trip, instead of the record count.
// Algorithmic bytes: F_intra written + 2A residual read + 32 B/record (+ edge re-reads, all L2 hits).
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "../av1_consts.h"
#include "dev_common.cuh"
#include "devframe.h"
#include "intra.h"
#include "../tables/tables_pred.inc"
#include "../tables/tables_inter.inc"

namespace av1r {

__constant__ int16_t c_dr_deriv[90];
__constant__ uint8_t c_sm_weights[124];
__constant__ int8_t c_fi_taps[5][8][8];
__constant__ uint8_t c_itxw_log2[TX_SIZES_ALL];
__constant__ uint8_t c_itxh_log2[TX_SIZES_ALL];
__constant__ uint8_t c_ii_weights[128];
__constant__ uint8_t c_ii_codebook[3][16][3];
__constant__ uint8_t c_ii_signflip[BLOCK_SIZES_ALL][16];
__constant__ uint8_t c_ii_blk_w[BLOCK_SIZES_ALL];
__constant__ uint8_t c_ii_blk_h[BLOCK_SIZES_ALL];
static bool g_intra_const_loaded[64] = {false};

static constexpr int INTRA_WARPS = 2;
static constexpr int EDGE_PAD = 16;
static constexpr int EDGE_LEN = EDGE_PAD + 2 * 129 + 16;   // room for upsampled edges (index -2 .. 2*(w+h))

struct IntraSmem {
    int32_t above[2][EDGE_LEN];
    int32_t left[2][EDGE_LEN];
    int16_t tile[64 * 64 / 4];   // 32x32 int16: filter-intra predictions / CfL luma terms
    int16_t res[256];            // residual of blocks up to 256 samples, fetched before the dependency wait
};

// Samples of the frame under reconstruction, read through L2 (other SMs are writing it).
template <typename T>
struct FrameView {
    const uint8_t* p[3];
    uint32_t pitch[3];
    __device__ __forceinline__ int px(int plane, int x, int y) const { return (int)__ldcg((const T*)(p[plane] + (size_t)y * pitch[plane]) + x); }
};

template <typename T>
__device__ __forceinline__ int ldpx(const uint8_t* base, uint32_t pitch, int x, int y) {
    return (int)__ldcg((const T*)(base + (size_t)y * pitch) + x);
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ int edge_filter_strength_d(int w, int h, int filter_type, int delta) {
    const int d = abs(delta), blk = w + h;
    int s = 0;
    if (filter_type == 0) {
        if (blk <= 8) { if (d >= 56) s = 1; }
        else if (blk <= 16) { if (d >= 40) s = 1; }
        else if (blk <= 24) { if (d >= 8) s = 1; if (d >= 16) s = 2; if (d >= 32) s = 3; }
        else if (blk <= 32) { if (d >= 1) s = 1; if (d >= 4) s = 2; if (d >= 32) s = 3; }
        else { if (d >= 1) s = 3; }
    } else {
        if (blk <= 8) { if (d >= 40) s = 1; if (d >= 64) s = 2; }
        else if (blk <= 16) { if (d >= 20) s = 1; if (d >= 48) s = 2; }
        else if (blk <= 24) { if (d >= 4) s = 3; }
        else { if (d >= 1) s = 3; }
    }
    return s;
}
__device__ __forceinline__ int edge_upsample_d(int w, int h, int filter_type, int delta) {
    const int d = abs(delta), blk = w + h;
    if (d <= 0 || d >= 40) return 0;
    return filter_type == 0 ? blk <= 16 : blk <= 8;
}

// src/dst point at element 0 (index -1 is the corner).  sz counts the corner.
__device__ __forceinline__ void edge_filter_d(const int32_t* src, int32_t* dst, int sz, int strength, int total, int lane) {
    const int k0 = strength == 3 ? 2 : 0, k1 = strength == 1 ? 4 : (strength == 2 ? 5 : 4), k2 = strength == 1 ? 8 : (strength == 2 ? 6 : 4);
    for (int i = lane; i < total + 2; i += 32) {
        // element index e = i - 1 over [-1, total]; filtered for 1 <= i < sz, copied otherwise
        int v;
        if (i >= 1 && i < sz) {
            const int a = src[max(i - 2, 0) - 1], b = src[max(i - 1, 0) - 1], c = src[i - 1], d = src[min(i + 1, sz - 1) - 1],
                      e = src[min(i + 2, sz - 1) - 1];
            v = (k0 * a + k1 * b + k2 * c + k1 * d + k0 * e + 8) >> 4;
        } else {
            v = src[i - 1];
        }
        dst[i - 1] = v;
    }
}

// upsample numPx samples: dst gets indices -2 .. 2*numPx-2
__device__ __forceinline__ void edge_upsample_d(const int32_t* src, int32_t* dst, int num_px, int pixmax, int lane) {
    for (int i = lane; i < num_px; i += 32) {
        // dup[k] = src[k-2] for k = 1..numPx+1, dup[0] = src[-1], dup[numPx+2] = src[numPx-1]
        const int d0 = src[max(i - 2, -1)], d1 = src[i - 1], d2 = src[i], d3 = src[min(i + 1, num_px - 1)];
        int s = -d0 + 9 * d1 + 9 * d2 - d3;
        s = min(max((s + 8) >> 4, 0), pixmax);
        dst[2 * i - 1] = s;
        dst[2 * i] = d2;
    }
    if (lane == 0) dst[-2] = src[-1];
}

// inter-intra blend weight of sample (i, j) of a w x h plane block (spec 7.11.3.13 / wedge 7.11.3.11)
__device__ __forceinline__ int ii_mask(int pk, const uint8_t* master, int i, int j, int w, int h, int sx, int sy) {
    const int wedge = pk & 1, wedge_index = (pk >> 1) & 15, ii_mode = (pk >> 5) & 3, bsize = (pk >> 7) & 31;
    if (wedge) {
        const int bw = c_ii_blk_w[bsize], bh = c_ii_blk_h[bsize];
        const uint8_t* cb = c_ii_codebook[bh > bw ? 0 : (bh < bw ? 1 : 2)][wedge_index];
        const int xoff = 32 - ((cb[1] * bw) >> 3), yoff = 32 - ((cb[2] * bh) >> 3);
        const uint8_t* mm = master + cb[0] * 4096;
        const int flip = c_ii_signflip[bsize][wedge_index];
        int acc = 0;
        for (int dy = 0; dy <= sy; dy++)
            for (int dx = 0; dx <= sx; dx++) {
                const int m = __ldg(mm + (yoff + (i << sy) + dy) * 64 + xoff + (j << sx) + dx);
                acc += flip ? 64 - m : m;
            }
        const int sh = sx + sy;
        return sh ? (acc + (1 << (sh - 1))) >> sh : acc;
    }
    const int scale = 128 / max(w, h);
    if (ii_mode == II_V_PRED) return c_ii_weights[i * scale];
    if (ii_mode == II_H_PRED) return c_ii_weights[j * scale];
    if (ii_mode == II_SMOOTH_PRED) return c_ii_weights[min(i, j) * scale];
    return 32;
}

template <typename T>
__device__ void intra_block(const TxRec& r, const DevPlanes& fr, const DevResidual& res, const DevFrameParams& fp, IntraSmem& sm,
                            const FrameView<T>& uv, int lane, const uint8_t* wedge_master, const uint8_t* pal) {
    const int plane = r.plane;
    const int lw = c_itxw_log2[r.txsz], lh = c_itxh_log2[r.txsz];
    const int w = 1 << lw, h = 1 << lh;
    const int x = r.x4 * 4, y = r.y4 * 4;
    const int bd = fp.bd, pixmax = (1 << bd) - 1;
    const int max_x = fp.cw[plane] - 1, max_y = fp.ch[plane] - 1;
    const uint32_t pitch = fr.pitch[plane];
    const int xe = min(w, fp.cw[plane] - x), ye = min(h, fp.ch[plane] - y);
    T* out = (T*)(fr.p[plane] + (size_t)y * pitch) + x;
    const int opitch = pitch / sizeof(T);
    const bool has_res = r.eob > 0;
    const bool res_pre = (w * h) <= 256;     // prefetched into sm.res (row-major w x h) by the caller
    const int16_t* rp = res_pre ? sm.res : res_ptr(res, plane, x, y);
    const int rpitch = res_pre ? w : (1 << res.tw_log2[plane]);
    const bool ii = (r.flags & TXF_II) != 0;
    const int ii_pk = (uint16_t)r.cfl_alpha;
    const int psx = plane ? fp.subx : 0, psy = plane ? fp.suby : 0;
    auto emit = [&](int i, int j, int v) {
        if (i < ye && j < xe) {
            if (ii) {   // blend the intra predictor over the inter predictor K2 left in the frame
                const int m = ii_mask(ii_pk, wedge_master, i, j, w, h, psx, psy);
                v = (m * v + (64 - m) * (int)__ldcg(out + i * opitch + j) + 32) >> 6;
            }
            if (has_res) v = min(max(v + (int)rp[i * rpitch + j], 0), pixmax);
            out[i * opitch + j] = (T)v;
        }
    };
    if (r.mode == TXM_INTER) {
        // plain inter residuals were added by the K2 residual kernel; only inter-intra blocks wait for their blend
        if (has_res && ii)
            for (int idx = lane; idx < w * h; idx += 32) {
                const int i = idx >> lw, j = idx & (w - 1);
                if (i < ye && j < xe) {
                    int v = (int)__ldcg(out + i * opitch + j) + (int)rp[i * rpitch + j];
                    out[i * opitch + j] = (T)min(max(v, 0), pixmax);
                }
            }
        return;
    }
    if (r.mode == TXM_PALETTE) {   // spec 7.11.4; entry layout documented at TileDecoder::palette_tokens
        const uint8_t* e = pal + r.pal_off;
        const uint16_t* hdr = reinterpret_cast<const uint16_t*>(e);
        const uint8_t* map = e + 24;
        const int ox = hdr[8], oy = hdr[9], stride = hdr[10];
        for (int idx = lane; idx < w * h; idx += 32) {
            const int i = idx >> lw, j = idx & (w - 1);
            emit(i, j, (int)hdr[map[(size_t)(y - oy + i) * stride + (x - ox + j)]]);
        }
        return;
    }
    const int have_left = r.flags & TXF_HAVE_LEFT, have_above = r.flags & TXF_HAVE_ABOVE;
    const int have_ar = r.flags & TXF_HAVE_ABOVE_RIGHT, have_bl = r.flags & TXF_HAVE_BELOW_LEFT;
    int32_t* above = sm.above[0] + EDGE_PAD;
    int32_t* left = sm.left[0] + EDGE_PAD;
    const int n = w + h;
    // ---- edges
    for (int i = lane; i < n; i += 32) {
        int a, l;
        if (have_above) {
            const int idx = have_ar ? min(i, 2 * w - 1) : min(i, w - 1);
            a = uv.px(plane, min(max_x, x + idx), y - 1);
        } else if (have_left) {
            a = uv.px(plane, x - 1, y);
        } else {
            a = (1 << (bd - 1)) - 1;
        }
        if (have_left) {
            const int idx = have_bl ? min(i, 2 * h - 1) : min(i, h - 1);
            l = uv.px(plane, x - 1, min(max_y, y + idx));
        } else if (have_above) {
            l = uv.px(plane, x, y - 1);
        } else {
            l = (1 << (bd - 1)) + 1;
        }
        above[i] = a;
        left[i] = l;
    }
    if (lane == 0) {
        int c;
        if (have_above && have_left) c = uv.px(plane, x - 1, y - 1);
        else if (have_above) c = uv.px(plane, x, y - 1);
        else if (have_left) c = uv.px(plane, x - 1, y);
        else c = 1 << (bd - 1);
        above[-1] = c;
        left[-1] = c;
    }
    __syncwarp();
    int mode = r.mode;
    if (mode == TXM_CFL) mode = DC_PRED;

    if (mode == TXM_FILTER_INTRA) {
        const int w4 = w >> 2, h2 = h >> 1;
        int16_t* pt = sm.tile;   // w x h predictions
        const int fm = r.fi_mode;
        for (int d = 0; d < h2 + w4 - 1; d++) {
            const int i2_lo = max(0, d - (w4 - 1)), i2_hi = min(h2 - 1, d);
            const int nblk = i2_hi - i2_lo + 1;
            for (int t = lane; t < nblk * 8; t += 32) {
                const int i2 = i2_lo + (t >> 3), j4 = d - i2, o = t & 7;
                int p[7];
#pragma unroll
                for (int i = 0; i < 7; i++) {
                    int v;
                    if (i < 5) {
                        if (i2 == 0) v = above[(j4 << 2) + i - 1];
                        else if (j4 == 0 && i == 0) v = left[(i2 << 1) - 1];
                        else v = pt[((i2 << 1) - 1) * w + (j4 << 2) + i - 1];
                    } else {
                        if (j4 == 0) v = left[(i2 << 1) + i - 5];
                        else v = pt[((i2 << 1) + i - 5) * w + (j4 << 2) - 1];
                    }
                    p[i] = v;
                }
                int pr = 0;
#pragma unroll
                for (int i = 0; i < 7; i++) pr += c_fi_taps[fm][o][i] * p[i];
                const int v = pr >= 0 ? (pr + 8) >> 4 : -((-pr + 8) >> 4);
                pt[((i2 << 1) + (o >> 2)) * w + (j4 << 2) + (o & 3)] = (int16_t)min(max(v, 0), pixmax);
            }
            __syncwarp();
        }
        for (int idx = lane; idx < w * h; idx += 32) emit(idx >> lw, idx & (w - 1), pt[idx]);
        return;
    }
    if (mode >= V_PRED && mode <= D67_PRED) {
        const int kModeToAngle[9] = {0, 90, 180, 45, 135, 113, 157, 203, 67};
        const int p_angle = kModeToAngle[mode] + r.angle_delta * 3;
        int up_above = 0, up_left = 0;
        if (fp.enable_edge_filter) {
            const int filter_type = (r.flags & TXF_SMOOTH_EDGE) ? 1 : 0;
            if (p_angle != 90 && p_angle != 180) {
                if (p_angle > 90 && p_angle < 180 && (w + h) >= 24) {
                    if (lane == 0) {
                        const int v = (left[0] * 5 + above[-1] * 6 + above[0] * 5 + 8) >> 4;
                        above[-1] = v;
                        left[-1] = v;
                    }
                    __syncwarp();
                }
                if (have_above) {
                    const int strength = edge_filter_strength_d(w, h, filter_type, p_angle - 90);
                    if (strength) {
                        const int num_px = min(w, max_x - x + 1) + (p_angle < 90 ? h : 0) + 1;
                        int32_t* dst = (above == sm.above[0] + EDGE_PAD) ? sm.above[1] + EDGE_PAD : sm.above[0] + EDGE_PAD;
                        edge_filter_d(above, dst, num_px, strength, n - 1, lane);
                        above = dst;
                        __syncwarp();
                    }
                }
                if (have_left) {
                    const int strength = edge_filter_strength_d(w, h, filter_type, p_angle - 180);
                    if (strength) {
                        const int num_px = min(h, max_y - y + 1) + (p_angle > 180 ? w : 0) + 1;
                        int32_t* dst = (left == sm.left[0] + EDGE_PAD) ? sm.left[1] + EDGE_PAD : sm.left[0] + EDGE_PAD;
                        edge_filter_d(left, dst, num_px, strength, n - 1, lane);
                        left = dst;
                        __syncwarp();
                    }
                }
            }
            up_above = edge_upsample_d(w, h, filter_type, p_angle - 90);
            if (up_above) {
                int32_t* dst = (above == sm.above[0] + EDGE_PAD) ? sm.above[1] + EDGE_PAD : sm.above[0] + EDGE_PAD;
                edge_upsample_d(above, dst, w + (p_angle < 90 ? h : 0), pixmax, lane);
                above = dst;
                __syncwarp();
            }
            up_left = edge_upsample_d(w, h, filter_type, p_angle - 180);
            if (up_left) {
                int32_t* dst = (left == sm.left[0] + EDGE_PAD) ? sm.left[1] + EDGE_PAD : sm.left[0] + EDGE_PAD;
                edge_upsample_d(left, dst, h + (p_angle > 180 ? w : 0), pixmax, lane);
                left = dst;
                __syncwarp();
            }
        }
        int dx = 0, dy = 0;
        if (p_angle < 90) dx = c_dr_deriv[p_angle];
        else if (p_angle > 90 && p_angle < 180) dx = c_dr_deriv[180 - p_angle];
        if (p_angle > 90 && p_angle < 180) dy = c_dr_deriv[p_angle - 90];
        else if (p_angle > 180) dy = c_dr_deriv[270 - p_angle];
        for (int idx = lane; idx < w * h; idx += 32) {
            const int i = idx >> lw, j = idx & (w - 1);
            int v;
            if (p_angle < 90) {
                const int id = (i + 1) * dx;
                const int b = (id >> (6 - up_above)) + (j << up_above);
                const int sh = ((id << up_above) >> 1) & 0x1F;
                const int max_base = (w + h - 1) << up_above;
                v = b < max_base ? (above[b] * (32 - sh) + above[b + 1] * sh + 16) >> 5 : above[max_base];
            } else if (p_angle == 90) {
                v = above[j];
            } else if (p_angle < 180) {
                int id = (j << 6) - (i + 1) * dx;
                int b = id >> (6 - up_above);
                if (b >= -(1 << up_above)) {
                    const int sh = ((id << up_above) >> 1) & 0x1F;
                    v = (above[b] * (32 - sh) + above[b + 1] * sh + 16) >> 5;
                } else {
                    id = (i << 6) - (j + 1) * dy;
                    b = id >> (6 - up_left);
                    const int sh = ((id << up_left) >> 1) & 0x1F;
                    v = (left[b] * (32 - sh) + left[b + 1] * sh + 16) >> 5;
                }
            } else if (p_angle == 180) {
                v = left[i];
            } else {
                const int id = (j + 1) * dy;
                const int b = (id >> (6 - up_left)) + (i << up_left);
                const int sh = ((id << up_left) >> 1) & 0x1F;
                const int max_base = (w + h - 1) << up_left;
                v = b < max_base ? (left[b] * (32 - sh) + left[b + 1] * sh + 16) >> 5 : left[max_base];
            }
            emit(i, j, v);
        }
        return;
    }
    if (mode == SMOOTH_PRED || mode == SMOOTH_V_PRED || mode == SMOOTH_H_PRED) {
        const uint8_t* ww = c_sm_weights + (w - 4);
        const uint8_t* wh = c_sm_weights + (h - 4);
        const int bl = left[h - 1], tr = above[w - 1];
        for (int idx = lane; idx < w * h; idx += 32) {
            const int i = idx >> lw, j = idx & (w - 1);
            int v;
            if (mode == SMOOTH_PRED) v = (wh[i] * above[j] + (256 - wh[i]) * bl + ww[j] * left[i] + (256 - ww[j]) * tr + 256) >> 9;
            else if (mode == SMOOTH_V_PRED) v = (wh[i] * above[j] + (256 - wh[i]) * bl + 128) >> 8;
            else v = (ww[j] * left[i] + (256 - ww[j]) * tr + 128) >> 8;
            emit(i, j, v);
        }
        return;
    }
    if (mode == PAETH_PRED) {
        const int tl = above[-1];
        for (int idx = lane; idx < w * h; idx += 32) {
            const int i = idx >> lw, j = idx & (w - 1);
            const int b = above[j] + left[i] - tl;
            const int pl = abs(b - left[i]), pt = abs(b - above[j]), ptl = abs(b - tl);
            const int v = (pl <= pt && pl <= ptl) ? left[i] : (pt <= ptl ? above[j] : tl);
            emit(i, j, v);
        }
        return;
    }
    // ---- DC (and CfL on top of it)
    int dc;
    {
        int s = 0;
        if (have_left)
            for (int k = lane; k < h; k += 32) s += left[k];
        if (have_above)
            for (int k = lane; k < w; k += 32) s += above[k];
        s = warp_sum(s);
        if (have_left && have_above) dc = (s + ((w + h) >> 1)) / (w + h);
        else if (have_left) dc = (s + (h >> 1)) >> lh;
        else if (have_above) dc = (s + (w >> 1)) >> lw;
        else dc = 1 << (bd - 1);
    }
    if (r.mode != TXM_CFL) {
        for (int idx = lane; idx < w * h; idx += 32) emit(idx >> lw, idx & (w - 1), dc);
        return;
    }
    {
        const int sx = fp.subx, sy = fp.suby;
        const int max_lw = r.cfl_max_w4 * 4, max_lh = r.cfl_max_h4 * 4;
        int16_t* L = sm.tile;
        int s = 0;
        for (int idx = lane; idx < w * h; idx += 32) {
            const int i = idx >> lw, j = idx & (w - 1);
            const int ly = min((y + i) << sy, max_lh - (1 << sy)), lx = min((x + j) << sx, max_lw - (1 << sx));
            int t = 0;
            for (int dy2 = 0; dy2 <= sy; dy2++)
                for (int dx2 = 0; dx2 <= sx; dx2++) t += uv.px(0, lx + dx2, ly + dy2);
            const int v = t << (3 - sx - sy);
            L[idx] = (int16_t)v;
            s += v;
        }
        s = warp_sum(s);
        const int sh = lw + lh;
        const int avg = (s + (1 << (sh - 1))) >> sh;
        const int alpha = r.cfl_alpha;
        __syncwarp();
        for (int idx = lane; idx < w * h; idx += 32) {
            const int t = alpha * ((int)L[idx] - avg);
            const int scaled = t >= 0 ? (t + 32) >> 6 : -((-t + 32) >> 6);
            emit(idx >> lw, idx & (w - 1), min(max(dc + scaled, 0), pixmax));
        }
    }
}

// 1. owner map: position (K3 order) of the record that reconstructs each 4x4 cell
__global__ void __launch_bounds__(128) wmap_scatter_kernel(IntraLaunch L) {
    const int pos = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (pos >= L.n) return;
    const TxRec r = L.recs[L.order[pos]];
    const int plane = r.plane;
    const int w4 = 1 << (c_itxw_log2[r.txsz] - 2), h4 = 1 << (c_itxh_log2[r.txsz] - 2);
    const int pw4 = L.fp.pw4[plane], ph4 = L.fp.ph4[plane];
    int32_t* m = L.wmap[plane];
    for (int c = lane; c < w4 * h4; c += 32) {
        const int cy = r.y4 + c / w4, cx = r.x4 + c % w4;
        if (cx < pw4 && cy < ph4) atomicMax(m + (size_t)cy * pw4 + cx, pos);
    }
}

__device__ __forceinline__ void wait_flag(const int* flags, int id) {
    const volatile int* f = flags + id;
    int ns = 32;
    while (*f == 0) {
        __nanosleep(ns);
        if (ns < 256) ns <<= 1;
    }
}

// Warp-collective wait: every lane names one owner position (or -1).  All pending flags are sampled once per round with a
// single warp-wide load; only lane 0 then spins, on the highest pending position (the one most likely to finish last), so a
// waiting warp costs L2 one poll per back-off period instead of one per lane.
__device__ __forceinline__ void wait_deps_warp(const int* flags, int id, int lane) {
    while (true) {
        const bool pend = id >= 0 && *(const volatile int*)(flags + id) == 0;
        if (!__any_sync(0xffffffffu, pend)) return;
        const int mx = __reduce_max_sync(0xffffffffu, pend ? id : -1);
        if (lane == 0) wait_flag(flags, mx);
        __syncwarp();
    }
}

// 2. the dataflow kernel
template <typename T>
__global__ void __launch_bounds__(INTRA_WARPS * 32) intra_dataflow_kernel(IntraLaunch L) {
    __shared__ IntraSmem s_sm[INTRA_WARPS];
    const int warp_in = threadIdx.x >> 5, lane = threadIdx.x & 31;
    IntraSmem& sm = s_sm[warp_in];
    const DevFrameParams& fp = L.fp;
    FrameView<T> uv;
    for (int pl = 0; pl < 3; pl++) {
        uv.p[pl] = L.frame.p[pl];
        uv.pitch[pl] = L.frame.pitch[pl];
    }
    while (true) {
        // persistent warps claim records in K3 order; the grid (L.ctas) bounds the ticket window of one frame so that the frames
        // in flight on other streams share the machine instead of one frame's spinning warps monopolising it
        int pos = 0;
        if (lane == 0) pos = atomicAdd(L.ticket, 1);
        pos = __shfl_sync(0xffffffffu, pos, 0);
        if (pos >= L.n) return;
        __syncwarp();
        const TxRec r = L.recs[L.order[pos]];
        // residual of small blocks: fetch now (it has no dependency), so it is off the critical path after the wait
        if (r.eob > 0) {
            const int lw = c_itxw_log2[r.txsz], w = 1 << lw, n = w << c_itxh_log2[r.txsz];
            if (n <= 256) {
                const int16_t* rp = res_ptr(L.res, r.plane, r.x4 * 4, r.y4 * 4);
                const int rpitch = 1 << L.res.tw_log2[r.plane];
                for (int idx = lane; idx < n; idx += 32) sm.res[idx] = __ldg(rp + (idx >> lw) * rpitch + (idx & (w - 1)));
            }
        }
        // ---- wait for the owners of everything this record reads
        if (r.mode == TXM_INTER) {
            if (r.flags & TXF_II) {   // residual of an inter-intra block: after its blend record
                if (lane == 0) wait_flag(L.flags, (int)r.pal_off);
            }
        } else if (r.mode != TXM_PALETTE) {
            const int plane = r.plane;
            const int w4 = 1 << (c_itxw_log2[r.txsz] - 2), h4 = 1 << (c_itxh_log2[r.txsz] - 2);
            const int pw4 = fp.pw4[plane], ph4 = fp.ph4[plane];
            const int have_left = r.flags & TXF_HAVE_LEFT, have_above = r.flags & TXF_HAVE_ABOVE;
            int need_ar = 0, need_bl = 0;
            if (r.mode >= V_PRED && r.mode <= D67_PRED) {
                const int kModeToAngle[9] = {0, 90, 180, 45, 135, 113, 157, 203, 67};
                const int p_angle = kModeToAngle[r.mode] + r.angle_delta * 3;
                need_ar = p_angle < 90 && (r.flags & TXF_HAVE_ABOVE_RIGHT);
                need_bl = p_angle > 180 && (r.flags & TXF_HAVE_BELOW_LEFT);
            }
            // V_PRED reads only the row above, H_PRED only the left column (unless that edge is missing and the other one stands in)
            const int want_above = have_above && !(r.mode == H_PRED && r.angle_delta == 0 && have_left);
            const int want_left = have_left && !(r.mode == V_PRED && r.angle_delta == 0 && have_above);
            const int na = want_above ? w4 * (need_ar ? 2 : 1) + 1 : 0;      // cells x4-1 .. on row y4-1 (the first is the corner)
            const int nl = want_left ? h4 * (need_bl ? 2 : 1) : 0;            // cells y4 .. on column x4-1
            const int32_t* m = L.wmap[plane];
            for (int c0 = 0; c0 < na + nl; c0 += 32) {       // warp-uniform trip count: the wait is collective
                const int c = c0 + lane;
                int id = -1;
                if (c < na + nl) {
                    int cx, cy;
                    if (c < na) {
                        cx = r.x4 - 1 + c;
                        cy = r.y4 - 1;
                    } else {
                        cx = r.x4 - 1;
                        cy = r.y4 + (c - na);
                    }
                    cx = min(max(cx, 0), pw4 - 1);
                    cy = min(max(cy, 0), ph4 - 1);
                    id = __ldg(m + (size_t)cy * pw4 + cx);
                    if (id >= pos) id = -1;
                }
                wait_deps_warp(L.flags, id, lane);
            }
            if (r.mode == TXM_CFL) {   // luma samples under this chroma block
                const int sx = fp.subx, sy = fp.suby;
                const int lx0 = (r.x4 << sx), ly0 = (r.y4 << sy);
                const int lx1 = min((r.x4 + w4) << sx, (int)r.cfl_max_w4), ly1 = min((r.y4 + h4) << sy, (int)r.cfl_max_h4);
                const int lw4 = max(lx1 - lx0, 0), lh4 = max(ly1 - ly0, 0);
                const int lpw4 = fp.pw4[0];
                for (int c0 = 0; c0 < lw4 * lh4; c0 += 32) {
                    const int c = c0 + lane;
                    int id = -1;
                    if (c < lw4 * lh4) {
                        id = __ldg(L.wmap[0] + (size_t)(ly0 + c / lw4) * lpw4 + lx0 + c % lw4);
                        if (id >= pos) id = -1;
                    }
                    wait_deps_warp(L.flags, id, lane);
                }
            }
        }
        __syncwarp();
        __threadfence();   // acquire: the owners' samples are visible in L2
        intra_block<T>(r, L.frame, L.res, fp, sm, uv, lane, L.wedge_master, L.pal);
        __threadfence();   // release: this record's samples before its flag
        __syncwarp();
        if (lane == 0) *(volatile int*)(L.flags + pos) = 1;
    }
}

static cudaError_t intra_upload_constants() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && g_intra_const_loaded[dev]) return cudaSuccess;
    if ((e = cudaMemcpyToSymbol(c_dr_deriv, av1t_dr_intra_derivative, sizeof(av1t_dr_intra_derivative))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_sm_weights, av1t_smooth_weights, sizeof(av1t_smooth_weights))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_fi_taps, av1t_filter_intra_taps, sizeof(av1t_filter_intra_taps))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_itxw_log2, kTxWLog2, sizeof(kTxWLog2))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_itxh_log2, kTxHLog2, sizeof(kTxHLog2))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_ii_weights, av1t_ii_weights1d, sizeof(av1t_ii_weights1d))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_ii_codebook, av1t_wedge_codebook, sizeof(av1t_wedge_codebook))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_ii_signflip, av1t_wedge_signflip, sizeof(av1t_wedge_signflip))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_ii_blk_w, kBlockW, sizeof(kBlockW))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_ii_blk_h, kBlockH, sizeof(kBlockH))) != cudaSuccess) return e;
    if (dev < 64) g_intra_const_loaded[dev] = true;
    return cudaSuccess;
}

cudaError_t launch_intra(const IntraLaunch& L, cudaStream_t s) {
    if (L.n <= 0) return cudaSuccess;
    cudaError_t e = intra_upload_constants();
    if (e != cudaSuccess) return e;
    wmap_scatter_kernel<<<(L.n + 3) / 4, 128, 0, s>>>(L);
    const int blocks = std::min((L.n + INTRA_WARPS - 1) / INTRA_WARPS, std::max(1, L.ctas));
    static bool attr_done = false;
    if (!attr_done) {   // 14 KB of static shared memory per 2-warp CTA: ask for the large carve-out so 15 CTAs fit per SM
        cudaFuncSetAttribute(intra_dataflow_kernel<uint8_t>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(intra_dataflow_kernel<uint16_t>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr_done = true;
    }
    if (L.fp.bd == 8) intra_dataflow_kernel<uint8_t><<<blocks, INTRA_WARPS * 32, 0, s>>>(L);
    else intra_dataflow_kernel<uint16_t><<<blocks, INTRA_WARPS * 32, 0, s>>>(L);
    return cudaGetLastError();
}

}  // namespace av1r
