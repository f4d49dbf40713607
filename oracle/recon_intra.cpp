// TEST INFRASTRUCTURE ONLY -- scalar CPU reconstruction of one frame from the parser's work-lists:
// dequantisation (spec 7.12.3), 2-D inverse transform (7.13.3, itx_oracle.cpp) and intra prediction
// (7.11.2: edge preparation, directional with edge filter / upsampling, smooth, Paeth, DC,
// recursive filter intra, chroma-from-luma 7.11.5).  The arithmetic the reference daemon would get
// from libdav1d inside its ffmpeg child (/root/reference/internal/ffmpeg/transcode.go:195); pinned
// against dav1d 1.5.3 with inloop_filters=0 by tests/test_decode_intra.py.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../av1-go_b200/csrc/frame_state.h"
#include "oracle_frame.h"

#include "../av1-go_b200/csrc/tables/tables_pred.inc"
#include "../av1-go_b200/csrc/tables/tables_quant.inc"
#include "../av1-go_b200/csrc/tables/tables_qm.inc"

extern "C" void orc_inverse_transform_2d(const int32_t* coef, int txsz, int txtp, int bd, int32_t* res);

namespace orc {
using namespace av1r;

static inline int round2(int x, int n) { return n == 0 ? x : (x + (1 << (n - 1))) >> n; }
static inline int round2s(int x, int n) { return x >= 0 ? round2(x, n) : -round2(-x, n); }
static inline int clip3(int lo, int hi, int x) { return x < lo ? lo : (x > hi ? hi : x); }

static const int kModeToAngle[13] = {0, 90, 180, 45, 135, 113, 157, 203, 67, 0, 0, 0, 0};

static int edge_filter_strength(int w, int h, int filter_type, int delta) {
    const int d = abs(delta), blk = w + h;
    int s = 0;
    if (filter_type == 0) {
        if (blk <= 8) { if (d >= 56) s = 1; }
        else if (blk <= 12) { if (d >= 40) s = 1; }
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
static int edge_upsample(int w, int h, int filter_type, int delta) {
    const int d = abs(delta), blk = w + h;
    if (d <= 0 || d >= 40) return 0;
    return filter_type == 0 ? blk <= 16 : blk <= 8;
}
// buf points at element index 0; valid range [-1, sz-2]  (sz counts the corner)
static void edge_filter(int* buf, int sz, int strength) {
    if (!strength) return;
    static const int kern[3][5] = {{0, 4, 8, 4, 0}, {0, 5, 6, 5, 0}, {2, 4, 4, 4, 2}};
    int edge[160];
    for (int i = 0; i < sz; i++) edge[i] = buf[i - 1];
    for (int i = 1; i < sz; i++) {
        int s = 0;
        for (int j = 0; j < 5; j++) {
            int k = clip3(0, sz - 1, i - 2 + j);
            s += kern[strength - 1][j] * edge[k];
        }
        buf[i - 1] = (s + 8) >> 4;
    }
}
static void edge_upsample_apply(int* buf, int num_px, int pixmax) {
    int dup[80];
    dup[0] = buf[-1];
    for (int i = -1; i < num_px; i++) dup[i + 2] = buf[i];
    dup[num_px + 2] = buf[num_px - 1];
    buf[-2] = dup[0];
    for (int i = 0; i < num_px; i++) {
        int s = -dup[i] + 9 * dup[i + 1] + 9 * dup[i + 2] - dup[i + 3];
        s = clip3(0, pixmax, round2(s, 4));
        buf[2 * i - 1] = s;
        buf[2 * i] = dup[i + 2];
    }
}

// Predict one transform block into pred[h][w] (stride w).
void predict_intra_block(const Plane& pl, const FrameGeom& g, const TxRec& r, int plane, int* pred) {
    const int w = kTxW[r.txsz], h = kTxH[r.txsz];
    const int log2w = kTxWLog2[r.txsz], log2h = kTxHLog2[r.txsz];
    const int x = r.x4 * 4, y = r.y4 * 4;
    const int bd = g.bd;
    const int pixmax = (1 << bd) - 1;
    const int max_x = g.cw[plane] - 1, max_y = g.ch[plane] - 1;
    const int have_left = !!(r.flags & TXF_HAVE_LEFT), have_above = !!(r.flags & TXF_HAVE_ABOVE);
    const int have_ar = !!(r.flags & TXF_HAVE_ABOVE_RIGHT), have_bl = !!(r.flags & TXF_HAVE_BELOW_LEFT);
    int above_buf[16 + 2 * 64 + 64 + 16], left_buf[16 + 2 * 64 + 64 + 16];
    int* above = above_buf + 16;
    int* left = left_buf + 16;
    auto px = [&](int yy, int xx) -> int { return pl.at(xx, yy); };
    const int n = w + h;
    // ---- edges (spec 7.11.2; top-right / bottom-left limited to one transform size, then replicated)
    if (!have_above && have_left) {
        for (int i = 0; i < n; i++) above[i] = px(y, x - 1);
    } else if (!have_above && !have_left) {
        for (int i = 0; i < n; i++) above[i] = (1 << (bd - 1)) - 1;
    } else {
        for (int i = 0; i < w; i++) above[i] = px(y - 1, std::min(max_x, x + i));
        for (int i = w; i < n; i++) {
            if (have_ar && i < 2 * w) above[i] = px(y - 1, std::min(max_x, x + i));
            else above[i] = above[i - 1];
        }
    }
    if (!have_left && have_above) {
        for (int i = 0; i < n; i++) left[i] = px(y - 1, x);
    } else if (!have_left && !have_above) {
        for (int i = 0; i < n; i++) left[i] = (1 << (bd - 1)) + 1;
    } else {
        for (int i = 0; i < h; i++) left[i] = px(std::min(max_y, y + i), x - 1);
        for (int i = h; i < n; i++) {
            if (have_bl && i < 2 * h) left[i] = px(std::min(max_y, y + i), x - 1);
            else left[i] = left[i - 1];
        }
    }
    if (have_above && have_left) above[-1] = px(y - 1, x - 1);
    else if (have_above) above[-1] = px(y - 1, x);
    else if (have_left) above[-1] = px(y, x - 1);
    else above[-1] = 1 << (bd - 1);
    left[-1] = above[-1];

    int mode = r.mode;
    if (mode == TXM_CFL) mode = DC_PRED;
    if (mode == TXM_FILTER_INTRA) {
        const int w4 = w >> 2, h2 = h >> 1;
        for (int i2 = 0; i2 < h2; i2++)
            for (int j4 = 0; j4 < w4; j4++) {
                int p[7];
                for (int i = 0; i < 7; i++) {
                    if (i < 5) {
                        if (i2 == 0) p[i] = above[(j4 << 2) + i - 1];
                        else if (j4 == 0 && i == 0) p[i] = left[(i2 << 1) - 1];
                        else p[i] = pred[((i2 << 1) - 1) * w + (j4 << 2) + i - 1];
                    } else {
                        if (j4 == 0) p[i] = left[(i2 << 1) + i - 5];
                        else p[i] = pred[((i2 << 1) + i - 5) * w + (j4 << 2) - 1];
                    }
                }
                for (int i1 = 0; i1 < 2; i1++)
                    for (int j1 = 0; j1 < 4; j1++) {
                        int pr = 0;
                        for (int i = 0; i < 7; i++) pr += av1t_filter_intra_taps[r.fi_mode][(i1 << 2) + j1][i] * p[i];
                        pred[((i2 << 1) + i1) * w + (j4 << 2) + j1] = clip3(0, pixmax, round2s(pr, 4));
                    }
            }
        return;
    }
    if (is_directional_mode(mode)) {
        const int p_angle = kModeToAngle[mode] + r.angle_delta * 3;
        int up_above = 0, up_left = 0;
        if (g.enable_edge_filter) {
            const int filter_type = !!(r.flags & TXF_SMOOTH_EDGE);
            if (p_angle != 90 && p_angle != 180) {
                if (p_angle > 90 && p_angle < 180 && (w + h) >= 24) {
                    int v = round2(left[0] * 5 + above[-1] * 6 + above[0] * 5, 4);
                    above[-1] = left[-1] = v;
                }
                if (have_above) {
                    int strength = edge_filter_strength(w, h, filter_type, p_angle - 90);
                    int num_px = std::min(w, max_x - x + 1) + (p_angle < 90 ? h : 0) + 1;
                    edge_filter(above, num_px, strength);
                }
                if (have_left) {
                    int strength = edge_filter_strength(w, h, filter_type, p_angle - 180);
                    int num_px = std::min(h, max_y - y + 1) + (p_angle > 180 ? w : 0) + 1;
                    edge_filter(left, num_px, strength);
                }
            }
            up_above = edge_upsample(w, h, filter_type, p_angle - 90);
            if (up_above) edge_upsample_apply(above, w + (p_angle < 90 ? h : 0), pixmax);
            up_left = edge_upsample(w, h, filter_type, p_angle - 180);
            if (up_left) edge_upsample_apply(left, h + (p_angle > 180 ? w : 0), pixmax);
        }
        int dx = 0, dy = 0;
        if (p_angle < 90) dx = av1t_dr_intra_derivative[p_angle];
        else if (p_angle > 90 && p_angle < 180) dx = av1t_dr_intra_derivative[180 - p_angle];
        if (p_angle > 90 && p_angle < 180) dy = av1t_dr_intra_derivative[p_angle - 90];
        else if (p_angle > 180) dy = av1t_dr_intra_derivative[270 - p_angle];
        for (int i = 0; i < h; i++)
            for (int j = 0; j < w; j++) {
                int v;
                if (p_angle < 90) {
                    int idx = (i + 1) * dx;
                    int base = (idx >> (6 - up_above)) + (j << up_above);
                    int shift = ((idx << up_above) >> 1) & 0x1F;
                    int max_base = (w + h - 1) << up_above;
                    if (base < max_base) v = round2(above[base] * (32 - shift) + above[base + 1] * shift, 5);
                    else v = above[max_base];
                } else if (p_angle == 90) {
                    v = above[j];
                } else if (p_angle < 180) {
                    int idx = (j << 6) - (i + 1) * dx;
                    int base = idx >> (6 - up_above);
                    if (base >= -(1 << up_above)) {
                        int shift = ((idx << up_above) >> 1) & 0x1F;
                        v = round2(above[base] * (32 - shift) + above[base + 1] * shift, 5);
                    } else {
                        idx = (i << 6) - (j + 1) * dy;
                        base = idx >> (6 - up_left);
                        int shift = ((idx << up_left) >> 1) & 0x1F;
                        v = round2(left[base] * (32 - shift) + left[base + 1] * shift, 5);
                    }
                } else if (p_angle == 180) {
                    v = left[i];
                } else {
                    int idx = (j + 1) * dy;
                    int base = (idx >> (6 - up_left)) + (i << up_left);
                    int shift = ((idx << up_left) >> 1) & 0x1F;
                    int max_base = (w + h - 1) << up_left;
                    if (base < max_base) v = round2(left[base] * (32 - shift) + left[base + 1] * shift, 5);
                    else v = left[max_base];
                }
                pred[i * w + j] = v;
            }
        return;
    }
    if (mode == SMOOTH_PRED || mode == SMOOTH_V_PRED || mode == SMOOTH_H_PRED) {
        const uint8_t* ww = av1t_smooth_weights + (w - 4);
        const uint8_t* wh = av1t_smooth_weights + (h - 4);
        for (int i = 0; i < h; i++)
            for (int j = 0; j < w; j++) {
                int v;
                if (mode == SMOOTH_PRED) {
                    int s = wh[i] * above[j] + (256 - wh[i]) * left[h - 1] + ww[j] * left[i] + (256 - ww[j]) * above[w - 1];
                    v = round2(s, 9);
                } else if (mode == SMOOTH_V_PRED) {
                    v = round2(wh[i] * above[j] + (256 - wh[i]) * left[h - 1], 8);
                } else {
                    v = round2(ww[j] * left[i] + (256 - ww[j]) * above[w - 1], 8);
                }
                pred[i * w + j] = v;
            }
        return;
    }
    if (mode == PAETH_PRED) {
        for (int i = 0; i < h; i++)
            for (int j = 0; j < w; j++) {
                int base = above[j] + left[i] - above[-1];
                int pl_ = abs(base - left[i]), pt = abs(base - above[j]), ptl = abs(base - above[-1]);
                int v;
                if (pl_ <= pt && pl_ <= ptl) v = left[i];
                else if (pt <= ptl) v = above[j];
                else v = above[-1];
                pred[i * w + j] = v;
            }
        return;
    }
    // DC_PRED
    int dc;
    if (have_left && have_above) {
        int sum = 0;
        for (int k = 0; k < h; k++) sum += left[k];
        for (int k = 0; k < w; k++) sum += above[k];
        sum += (w + h) >> 1;
        dc = sum / (w + h);
    } else if (have_left) {
        int sum = 0;
        for (int k = 0; k < h; k++) sum += left[k];
        dc = (sum + (h >> 1)) >> log2h;
    } else if (have_above) {
        int sum = 0;
        for (int k = 0; k < w; k++) sum += above[k];
        dc = (sum + (w >> 1)) >> log2w;
    } else {
        dc = 1 << (bd - 1);
    }
    for (int i = 0; i < w * h; i++) pred[i] = dc;
}

static void cfl_apply(const Plane& luma, const FrameGeom& g, const TxRec& r, int* pred) {
    const int w = kTxW[r.txsz], h = kTxH[r.txsz];
    const int sx = g.subx, sy = g.suby;
    const int start_x = r.x4 * 4, start_y = r.y4 * 4;
    const int max_lw = r.cfl_max_w4 * 4, max_lh = r.cfl_max_h4 * 4;
    const int pixmax = (1 << g.bd) - 1;
    static thread_local int L[64 * 64];
    int avg = 0;
    for (int i = 0; i < h; i++) {
        int ly = (start_y + i) << sy;
        ly = std::min(ly, max_lh - (1 << sy));
        for (int j = 0; j < w; j++) {
            int lx = (start_x + j) << sx;
            lx = std::min(lx, max_lw - (1 << sx));
            int t = 0;
            for (int dy = 0; dy <= sy; dy++)
                for (int dx = 0; dx <= sx; dx++) t += luma.at(lx + dx, ly + dy);
            int v = t << (3 - sx - sy);
            L[i * w + j] = v;
            avg += v;
        }
    }
    avg = round2(avg, kTxWLog2[r.txsz] + kTxHLog2[r.txsz]);
    for (int i = 0; i < w * h; i++) {
        int scaled = round2s(r.cfl_alpha * (L[i] - avg), 6);
        pred[i] = clip3(0, pixmax, pred[i] + scaled);
    }
}

static void dequant_block(const FrameWork& fw, const FrameGeom& g, const TxRec& r, int32_t* coef) {
    const int w = kTxW[r.txsz], h = kTxH[r.txsz];
    const int cw = std::min(w, 32), chh = std::min(h, 32);
    memset(coef, 0, sizeof(int32_t) * cw * chh);
    const int bdi = (g.bd - 8) >> 1;
    const int dcq = av1t_dc_qlookup[bdi][clip3(0, 255, r.qidx + g.dq_dc[r.plane])];
    const int acq = av1t_ac_qlookup[bdi][clip3(0, 255, r.qidx + g.dq_ac[r.plane])];
    const int pels = w * h;
    const int dq_denom = (pels > 256) + (pels > 1024);
    const int mx = (1 << (7 + g.bd)) - 1, mn = -(1 << (7 + g.bd));
    // quantiser matrix (spec 7.12.3): 2-D transform types only, level 15 = flat
    const uint8_t* qm = (r.qm_level < 15 && r.txtp < IDTX) ? av1t_iqmatrix[r.qm_level][r.plane > 0] + av1t_qm_offset[r.txsz] : nullptr;
    for (int k = 0; k < r.ntok; k++) {
        uint32_t t = fw.coefs[r.coef_off + k];
        int pos = coef_token_pos(t), level = coef_token_level(t);
        int q = pos == 0 ? dcq : acq;
        if (qm) q = (q * qm[pos] + 16) >> 5;
        int64_t dq = (int64_t)abs(level) * q;
        dq &= 0xFFFFFF;
        dq >>= dq_denom;
        int v = level < 0 ? -(int)dq : (int)dq;
        coef[pos] = clip3(mn, mx, v);
    }
}

void predict_inter_frame(const FrameWork& fw, Frame& f, const Frame* const refs[8]);
void interintra_blend(Frame& f, const TxRec& r, const int* intra);

void reconstruct_frame(const FrameWork& fw, Frame& f, const Frame* const refs[8]) {
    const FrameGeom& g = f.g;
    if (!fw.inter.empty()) predict_inter_frame(fw, f, refs);
    static thread_local int pred[64 * 64];
    static thread_local int32_t coef[32 * 32], res[64 * 64];
    const int pixmax = (1 << g.bd) - 1;
    for (size_t ti = 0; ti < fw.tx.size(); ti++) {
        const TxRec& r = fw.tx[ti];
        Plane& pl = f.p[r.plane];
        const int w = kTxW[r.txsz], h = kTxH[r.txsz];
        const int x = r.x4 * 4, y = r.y4 * 4;
        // the whole transform block is reconstructed, also the part beyond the coded frame edge (it lands in the planes' margin):
        // chroma-from-luma of a block that straddles the edge averages those luma samples
        const int xe = x + w, ye = y + h;
        if (r.mode != TXM_INTER && (r.flags & TXF_II)) {
            predict_intra_block(pl, g, r, r.plane, pred);
            interintra_blend(f, r, pred);
            continue;
        }
        if (r.mode == TXM_PALETTE) {
            // spec 7.11.4: pred = palette[ColorMap]; entry layout documented at TileDecoder::palette_tokens
            const uint8_t* e = fw.pal.data() + r.pal_off;
            const uint16_t* hdr = reinterpret_cast<const uint16_t*>(e);
            const uint8_t* map = e + 24;
            const int ox = hdr[8], oy = hdr[9], stride = hdr[10];
            for (int i = 0; i < h; i++)
                for (int j = 0; j < w; j++) pred[i * w + j] = hdr[map[(size_t)(y - oy + i) * stride + (x - ox + j)]];
            for (int yy = y; yy < ye; yy++)
                for (int xx = x; xx < xe; xx++) pl.at(xx, yy) = (uint16_t)pred[(yy - y) * w + (xx - x)];
        } else if (r.mode == TXM_INTRABC) {
            // spec 7.11.3.2 - 7.11.3.4 with use_intrabc: the reference is the frame being decoded (no filter has run), bilinear
            // taps at 1/16 sample (chroma of an odd luma vector is a half-sample position), rounding 3 then 11, positions clamped
            // to the coded plane.  Decode order guarantees the source samples are final.
            const int sx = r.plane ? g.subx : 0, sy = r.plane ? g.suby : 0;
            const int dvx = (int16_t)r.cfl_max_w4, dvy = (int16_t)r.cfl_max_h4;
            const int posx = (x << 4) + ((2 * dvx) >> sx), posy = (y << 4) + ((2 * dvy) >> sy);
            const int ix = posx >> 4, fx = posx & 15, iy = posy >> 4, fy = posy & 15;
            const int lastx = g.cw[r.plane] - 1, lasty = g.ch[r.plane] - 1;
            const int round0 = g.bd == 12 ? 5 : 3, round1 = g.bd == 12 ? 9 : 11;
            for (int i = 0; i < ye - y; i++)
                for (int j = 0; j < xe - x; j++) {
                    int t[2];
                    for (int k = 0; k < 2; k++) {
                        const int yy = clip3(0, lasty, iy + i + k);
                        const int a = pl.at(clip3(0, lastx, ix + j), yy), b = pl.at(clip3(0, lastx, ix + j + 1), yy);
                        t[k] = round2((128 - 8 * fx) * a + 8 * fx * b, round0);
                    }
                    pred[i * w + j] = clip3(0, pixmax, round2((128 - 8 * fy) * t[0] + 8 * fy * t[1], round1));
                }
            for (int yy = y; yy < ye; yy++)
                for (int xx = x; xx < xe; xx++) pl.at(xx, yy) = (uint16_t)pred[(yy - y) * w + (xx - x)];
        } else if (r.mode != TXM_INTER) {
            predict_intra_block(pl, g, r, r.plane, pred);
            if (r.mode == TXM_CFL) cfl_apply(f.p[0], g, r, pred);
            for (int yy = y; yy < ye; yy++)
                for (int xx = x; xx < xe; xx++) pl.at(xx, yy) = (uint16_t)pred[(yy - y) * w + (xx - x)];
        }
        if (r.eob > 0) {
            dequant_block(fw, g, r, coef);
            orc_inverse_transform_2d(coef, r.txsz, r.txtp, g.bd, res);
            for (int yy = y; yy < ye; yy++)
                for (int xx = x; xx < xe; xx++)
                    pl.at(xx, yy) = (uint16_t)clip3(0, pixmax, pl.at(xx, yy) + res[(yy - y) * w + (xx - x)]);
        }
    }
}

}  // namespace orc
