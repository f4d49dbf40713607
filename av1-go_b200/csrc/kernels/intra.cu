// K3 -- intra prediction + residual add, wavefront over superblock rows with the current 64x64 unit
// resident in shared memory (AV1 spec 7.11.2, 7.11.5, 7.12.3).
//
// Intra prediction reads *reconstructed* neighbours and the decode order inside a superblock is a
// Z-order whose bottom-left quadrant depends on the top-right one, so the blocks of a superblock form
// one long dependency chain; the only parallelism is across superblock rows (row r may process SB c once
// row r-1 has finished SB c+1), across tiles and across frames.  What can be optimised is the *latency
// of one link of the chain*: one warp owns one (tile, SB row) item and keeps the 64x64 luma unit (+ its
// chroma) it is working on in shared memory together with the row above / column left of the unit
// (loaded from HBM/L2 once per unit), so the neighbour fetches of every block are shared-memory reads
// instead of L2 round trips; reconstructed samples are written through to the frame in HBM.
// Inter-row hand-off: per-item progress counters (release: __threadfence + store, acquire: volatile
// load + __threadfence); halo reads use ld.cg so a stale L1 line can never be observed.  Items are
// claimed through an atomic ticket, so a warp only waits on items that are already running.
// v4: the chain link is kept off HBM/L2 latency entirely -- the unit's int16 residual (12 KB, unit-major layout written by K1),
// the unit's 32-byte records (chunks of 64) and the item's unit table are brought into shared memory by TMA bulk copies
// (cp.async.bulk + mbarrier complete_tx), double-buffered so that unit k+1 / chunk q+1 are in flight while unit k / chunk q
// are being reconstructed; a block then costs shared-memory reads, ALU work and fire-and-forget stores only.
// Algorithmic bytes: F_intra written + 2A residual read + 32 B/record; halos are L2 hits.
#include <cuda_runtime.h>
#include <stdint.h>

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

static constexpr int INTRA_WARPS = 4;
static constexpr int EDGE_PAD = 16;
static constexpr int EDGE_LEN = EDGE_PAD + 2 * 129 + 16;   // room for upsampled edges (index -2 .. 2*(w+h))

struct IntraSmem {
    int32_t above[2][EDGE_LEN];
    int32_t left[2][EDGE_LEN];
    int16_t tile[64 * 64 / 4];   // 32x32 int16: filter-intra predictions / CfL luma terms
};

// The 64x64 unit being reconstructed, resident in shared memory (per warp).
template <typename T>
struct UnitView {
    T* tile[3];        // [th][tw] samples of the unit
    T* above[3];       // index -1 .. 2*tw-1 : frame row just above the unit
    T* left[3];        // index 0 .. 2*th-1  : frame column just left of the unit
    int ux0[3], uy0[3], tw[3], th[3];
    __device__ __forceinline__ int px(int plane, int x, int y) const {
        const int dx = x - ux0[plane], dy = y - uy0[plane];
        if (dy < 0) return above[plane][dx];
        if (dx < 0) return left[plane][dy];
        return tile[plane][dy * tw[plane] + dx];
    }
};

// ---- TMA bulk copy + mbarrier helpers (sm_90+; SASS UBLKCP / SYNCS)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
    } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

static constexpr int REC_CHUNK = 64;     // records per bulk copy
static constexpr int SBS_CAP = 128;      // unit descriptors staged per item

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
__device__ void intra_block(const TxRec& r, const DevPlanes& fr, const int16_t* res_s, const DevResidual& res, const DevFrameParams& fp, IntraSmem& sm,
                            const UnitView<T>& uv, int lane, const uint8_t* wedge_master, const uint8_t* pal) {
    const int plane = r.plane;
    const int lw = c_itxw_log2[r.txsz], lh = c_itxh_log2[r.txsz];
    const int w = 1 << lw, h = 1 << lh;
    const int x = r.x4 * 4, y = r.y4 * 4;
    const int bd = fp.bd, pixmax = (1 << bd) - 1;
    const int max_x = fp.cw[plane] - 1, max_y = fp.ch[plane] - 1;
    const uint32_t pitch = fr.pitch[plane];
    const int xe = min(w, fp.cw[plane] - x), ye = min(h, fp.ch[plane] - y);
    T* tl = uv.tile[plane] + (y - uv.uy0[plane]) * uv.tw[plane] + (x - uv.ux0[plane]);
    const int tpitch = uv.tw[plane];
    T* out = (T*)(fr.p[plane] + (size_t)y * pitch) + x;
    const int opitch = pitch / sizeof(T);
    // residual of the unit, resident in shared memory (same tile geometry as the sample tile)
    const int16_t* rp = res_s + res.plane_off[plane] + (y - uv.uy0[plane]) * uv.tw[plane] + (x - uv.ux0[plane]);
    const int rpitch = uv.tw[plane];
    const bool has_res = r.eob > 0;
    const bool ii = (r.flags & TXF_II) != 0;
    const int ii_pk = (uint16_t)r.cfl_alpha;
    const int psx = plane ? fp.subx : 0, psy = plane ? fp.suby : 0;
    auto emit = [&](int i, int j, int v) {
        if (i < ye && j < xe) {
            if (ii) {   // blend the intra predictor over the inter predictor already in the unit
                const int m = ii_mask(ii_pk, wedge_master, i, j, w, h, psx, psy);
                v = (m * v + (64 - m) * (int)tl[i * tpitch + j] + 32) >> 6;
            }
            if (has_res) v = min(max(v + (int)rp[i * rpitch + j], 0), pixmax);
            out[i * opitch + j] = (T)v;
            tl[i * tpitch + j] = (T)v;
        }
    };
    if (r.mode == TXM_INTER) {
        // plain inter residuals were added by the K2 residual kernel; only inter-intra blocks wait for their blend
        if (has_res && ii)
            for (int idx = lane; idx < w * h; idx += 32) {
                const int i = idx >> lw, j = idx & (w - 1);
                if (i < ye && j < xe) {
                    int v = (int)tl[i * tpitch + j] + (int)rp[i * rpitch + j];
                    v = min(max(v, 0), pixmax);
                    out[i * opitch + j] = (T)v;
                    tl[i * tpitch + j] = (T)v;
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

template <typename T>
__global__ void __launch_bounds__(INTRA_WARPS * 32) intra_wavefront_kernel(IntraLaunch L) {
    extern __shared__ __align__(128) uint8_t s_raw[];
    const int warp_in = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* wbase = s_raw + (size_t)warp_in * L.smem_per_warp;
    // per-warp shared memory: [residual x2][records x2][unit table][barriers][IntraSmem][sample tiles + halos]
    const DevFrameParams& fp0 = L.frames[0].fp;
    const int unit_elems = L.frames[0].res.unit_elems;
    const uint32_t unit_bytes = (uint32_t)unit_elems * 2;
    int16_t* res_s[2] = {reinterpret_cast<int16_t*>(wbase), reinterpret_cast<int16_t*>(wbase + unit_bytes)};
    TxRec* rec_s[2];
    rec_s[0] = reinterpret_cast<TxRec*>(wbase + 2 * unit_bytes);
    rec_s[1] = rec_s[0] + REC_CHUNK;
    SbRange* sbs_s = reinterpret_cast<SbRange*>(rec_s[1] + REC_CHUNK);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sbs_s + SBS_CAP);   // [0,1] residual, [2,3] records, [4] unit table
    IntraSmem& sm = *reinterpret_cast<IntraSmem*>(bars + 8);
    UnitView<T> uv;
    {
        T* p = reinterpret_cast<T*>(reinterpret_cast<uint8_t*>(&sm) + sizeof(IntraSmem));
        for (int pl = 0; pl < 3; pl++) {
            const int sx = pl ? fp0.subx : 0, sy = pl ? fp0.suby : 0;
            uv.tw[pl] = 64 >> sx;
            uv.th[pl] = 64 >> sy;
            uv.tile[pl] = p;
            p += uv.tw[pl] * uv.th[pl];
            uv.above[pl] = p + 8;            // index -1 valid
            p += 2 * uv.tw[pl] + 16;
            uv.left[pl] = p;
            p += 2 * uv.th[pl] + 8;
        }
    }
    const uint32_t bar0 = smem_u32(bars);
    if (lane == 0) {
        for (int i = 0; i < 5; i++) mbar_init(bar0 + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t par = 0;   // bit i = parity to wait for on barrier i
    while (true) {
        int item = 0;
        if (lane == 0) item = atomicAdd(L.ticket, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= L.n_items) return;
        const SbRowItem it = L.items[item];
        const IntraFrame& F = L.frames[it.frame];
        const DevFrameParams& fp = F.fp;
        const int nplanes = fp.mono ? 1 : 3;
        const int n_units = (int)it.n_units;
        volatile int* dep = it.dep_item >= 0 ? (volatile int*)(L.progress + it.dep_item) : nullptr;
        // ---- stage the item's unit table, then start the first residual / record transfers
        const int n_tab = min(n_units, SBS_CAP);
        if (lane == 0) {
            fence_proxy_async();
            mbar_expect_tx(bar0 + 32, (uint32_t)(n_tab * sizeof(SbRange)));
            bulk_g2s(smem_u32(sbs_s), F.sbs + it.first_unit, (uint32_t)(n_tab * sizeof(SbRange)), bar0 + 32);
        }
        mbar_wait(bar0 + 32, (par >> 4) & 1);
        par ^= 16;
        auto unit_desc = [&](int k) -> SbRange { return k < SBS_CAP ? sbs_s[k] : F.sbs[it.first_unit + k]; };
        auto issue_res = [&](int k) {
            const SbRange u = unit_desc(k);
            if (lane == 0) {
                fence_proxy_async();
                mbar_expect_tx(bar0 + 8 * (k & 1), unit_bytes);
                bulk_g2s(smem_u32(res_s[k & 1]), F.res.base + ((size_t)u.uy * F.res.units_x + u.ux) * unit_elems, unit_bytes, bar0 + 8 * (k & 1));
            }
        };
        int ld_unit = 0, ld_chunk = 0, q_load = 0, q_use = 0;
        auto issue_rec = [&]() {
            if (ld_unit >= n_units) return;
            const SbRange u = unit_desc(ld_unit);
            const int first = (int)u.first + ld_chunk * REC_CHUNK;
            const int cnt = min(REC_CHUNK, (int)u.count - ld_chunk * REC_CHUNK);
            if (lane == 0) {
                fence_proxy_async();
                mbar_expect_tx(bar0 + 16 + 8 * (q_load & 1), (uint32_t)(cnt * sizeof(TxRec)));
                bulk_g2s(smem_u32(rec_s[q_load & 1]), F.recs + first, (uint32_t)(cnt * sizeof(TxRec)), bar0 + 16 + 8 * (q_load & 1));
            }
            q_load++;
            ld_chunk++;
            if (ld_chunk * REC_CHUNK >= (int)u.count) {
                ld_unit++;
                ld_chunk = 0;
            }
        };
        __syncwarp();
        issue_res(0);
        issue_rec();
        int sb_done = 0;
        for (int k = 0; k < n_units; k++) {
            const SbRange un = unit_desc(k);
            __syncwarp();
            if (k + 1 < n_units) issue_res(k + 1);   // buffer (k+1)&1 was last read by unit k-1
            const bool first_of_sb = (k == 0) || (unit_desc(k - 1).sb_col != un.sb_col);
            if (first_of_sb && dep) {
                const int c = un.sb_col - un.tile_sb_col0;
                const int need = min(c + 2, (int)it.n_sb);
                if (lane == 0)
                    while (*dep < need) __nanosleep(32);
                __syncwarp();
            }
            __threadfence();
            // ---- load the halo of the unit: row above (-1 .. 2tw-1) and column left (0 .. 2th-1)
            for (int pl = 0; pl < nplanes; pl++) {
                const int sx = pl ? fp.subx : 0, sy = pl ? fp.suby : 0;
                const int ux0 = (un.ux * 64) >> sx, uy0 = (un.uy * 64) >> sy;
                uv.ux0[pl] = ux0;
                uv.uy0[pl] = uy0;
                const T* base = (const T*)F.frame.p[pl];
                const int pe = F.frame.pitch[pl] / sizeof(T);
                const int max_x = fp.cw[pl] - 1, max_y = fp.ch[pl] - 1;
                if (uy0 > 0)
                    for (int i = lane - 1; i < 2 * uv.tw[pl]; i += 32) {
                        const int x = min(max(ux0 + i, 0), max_x);
                        uv.above[pl][i] = __ldcg(base + (size_t)(uy0 - 1) * pe + x);
                    }
                if (ux0 > 0)
                    for (int i = lane; i < 2 * uv.th[pl]; i += 32) {
                        const int y = min(uy0 + i, max_y);
                        uv.left[pl][i] = __ldcg(base + (size_t)y * pe + ux0 - 1);
                    }
                if (F.inter_frame) {
                    // inter-predicted (and residual-added) samples of this unit were produced by K2: bring them on chip
                    const int tw = uv.tw[pl], th = uv.th[pl];
                    for (int i = lane; i < tw * th; i += 32) {
                        const int yy = i / tw, xx = i - yy * tw;
                        uv.tile[pl][i] = __ldcg(base + (size_t)min(uy0 + yy, max_y) * pe + min(ux0 + xx, max_x));
                    }
                }
            }
            mbar_wait(bar0 + 8 * (k & 1), (par >> (k & 1)) & 1);
            par ^= 1u << (k & 1);
            __syncwarp();
            const int16_t* rs = res_s[k & 1];
            for (int c0 = 0; c0 < (int)un.count; c0 += REC_CHUNK) {
                const int qb = q_use & 1;
                mbar_wait(bar0 + 16 + 8 * qb, (par >> (2 + qb)) & 1);
                par ^= 4u << qb;
                __syncwarp();
                issue_rec();                              // next chunk -> the buffer chunk q_use-1 used
                const int cnt = min(REC_CHUNK, (int)un.count - c0);
                const TxRec* rb = rec_s[qb];
                for (int t = 0; t < cnt; t++) {
                    intra_block<T>(rb[t], F.frame, rs, F.res, fp, sm, uv, lane, F.wedge_master, F.pal);
                    __syncwarp();
                }
                q_use++;
            }
            const bool last_of_sb = (k + 1 == n_units) || (unit_desc(k + 1).sb_col != un.sb_col);
            if (last_of_sb) {
                __threadfence();
                __syncwarp();
                sb_done = un.sb_col - un.tile_sb_col0 + 1;
                if (lane == 0) *(volatile int*)(L.progress + item) = sb_done;
            }
        }
        // superblocks without any intra record never appear in the list: publish the full row at the end
        __threadfence();
        __syncwarp();
        if (lane == 0) *(volatile int*)(L.progress + item) = (int)it.n_sb;
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

size_t intra_smem_per_warp(int bd, int subx, int suby) {
    const size_t ts = bd == 8 ? 1 : 2;
    size_t unit_elems = 0, n = 0;
    for (int pl = 0; pl < 3; pl++) {
        const int sx = pl ? subx : 0, sy = pl ? suby : 0;
        const int tw = 64 >> sx, th = 64 >> sy;
        unit_elems += (size_t)tw * th;
        n += ts * (tw * th + 2 * tw + 16 + 2 * th + 8);
    }
    n += 2 * unit_elems * sizeof(int16_t) + 2 * REC_CHUNK * sizeof(TxRec) + SBS_CAP * sizeof(SbRange) + 8 * sizeof(uint64_t) + sizeof(IntraSmem);
    return (n + 127) & ~(size_t)127;
}

cudaError_t launch_intra(const IntraLaunch& L_, int bd, int subx, int suby, cudaStream_t s) {
    if (L_.n_items <= 0) return cudaSuccess;
    cudaError_t e = intra_upload_constants();
    if (e != cudaSuccess) return e;
    IntraLaunch L = L_;
    L.smem_per_warp = (int)intra_smem_per_warp(bd, subx, suby);
    const size_t smem = (size_t)L.smem_per_warp * INTRA_WARPS;
    const int blocks = (L.n_items + INTRA_WARPS - 1) / INTRA_WARPS;
    static bool attr_done[2] = {false, false};
    if (bd == 8) {
        auto k = intra_wavefront_kernel<uint8_t>;
        if (!attr_done[0]) {
            if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024)) != cudaSuccess) return e;
            cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            attr_done[0] = true;
        }
        k<<<blocks, INTRA_WARPS * 32, smem, s>>>(L);
    } else {
        auto k = intra_wavefront_kernel<uint16_t>;
        if (!attr_done[1]) {
            if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024)) != cudaSuccess) return e;
            cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            attr_done[1] = true;
        }
        k<<<blocks, INTRA_WARPS * 32, smem, s>>>(L);
    }
    return cudaGetLastError();
}

}  // namespace av1r
