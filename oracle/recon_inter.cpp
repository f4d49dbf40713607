// TEST INFRASTRUCTURE ONLY -- scalar CPU restatement of AV1 inter prediction from the parser's K2 work-list
// (spec 7.11.3): sub-pel 8-tap block prediction with reference edge clamping (7.11.3.3/4), warped motion
// (7.11.3.5), compound averaging / distance weights (7.11.3.15), wedge and difference-weighted masks
// (7.11.3.11/12), mask blend (7.11.3.14), overlapped motion compensation (7.11.3.10) and the inter-intra blend
// (7.11.3.13).  This is the arithmetic the reference daemon would get from libdav1d inside its ffmpeg child
// (/root/reference/internal/ffmpeg/transcode.go:195); pinned against dav1d 1.5.3 by tests/test_decode_inter.py.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../av1-go_b200/csrc/frame_state.h"
#include "oracle_frame.h"

#include "../av1-go_b200/csrc/tables/tables_inter.inc"

namespace orc {
using namespace av1r;

static inline int round2(int x, int n) { return n == 0 ? x : (x + (1 << (n - 1))) >> n; }
static inline int clip3(int lo, int hi, int x) { return x < lo ? lo : (x > hi ? hi : x); }

static const uint8_t kWedgeBitsO[BLOCK_SIZES_ALL] = {0, 0, 0, 4, 4, 4, 4, 4, 4, 4, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 0, 0};

// ---- wedge masks (7.11.3.11)
static uint8_t g_master[6][64][64];
static bool g_master_done = false;
static void init_master() {
    if (g_master_done) return;
    enum { WH = 0, WV = 1, W27 = 2, W63 = 3, W117 = 4, W153 = 5 };
    for (int j = 0; j < 64; j++) {
        int shift = 16;
        for (int i = 0; i < 64; i += 2) {
            g_master[W63][i][j] = av1t_wedge_master_oblique_even[clip3(0, 63, j - shift)];
            shift--;
            g_master[W63][i + 1][j] = av1t_wedge_master_oblique_odd[clip3(0, 63, j - shift)];
            g_master[WV][i][j] = av1t_wedge_master_vertical[j];
            g_master[WV][i + 1][j] = av1t_wedge_master_vertical[j];
        }
    }
    for (int i = 0; i < 64; i++)
        for (int j = 0; j < 64; j++) {
            const int msk = g_master[W63][i][j];
            g_master[W27][j][i] = (uint8_t)msk;
            g_master[W117][i][63 - j] = (uint8_t)(64 - msk);
            g_master[W153][63 - j][i] = (uint8_t)(64 - msk);
            g_master[WH][j][i] = g_master[WV][i][j];
        }
    g_master_done = true;
}
// WedgeMasks[bsize][flip][wedge][i][j]
static int wedge_mask(int bsize, int flip, int wedge, int i, int j) {
    const int w = kBlockW[bsize], h = kBlockH[bsize];
    const uint8_t* cb = av1t_wedge_codebook[h > w ? 0 : (h < w ? 1 : 2)][wedge];
    const int dir = cb[0], xoff = 32 - ((cb[1] * w) >> 3), yoff = 32 - ((cb[2] * h) >> 3);
    const int m = g_master[dir][yoff + i][xoff + j];
    return (flip ^ av1t_wedge_signflip[bsize][wedge]) ? 64 - m : m;
}

static int filter_index(int type, int len) {
    if (len <= 4) {
        if (type == INTERP_EIGHTTAP || type == INTERP_SHARP) return 4;
        if (type == INTERP_SMOOTH) return 5;
    }
    return type;
}

static inline int64_t round2s64(int64_t x, int n) { return x >= 0 ? (x + ((int64_t)1 << (n - 1))) >> n : -((-x + ((int64_t)1 << (n - 1))) >> n); }

// 7.11.3.3 (motion vector scaling) + 7.11.3.4 (block inter prediction) for a reference whose size differs from the frame's
// (cur_w x cur_h = luma size of the frame being predicted): positions advance in 1/1024 sample steps of xStep / yStep.
static void block_pred_scaled(const Frame& ref, int cur_w, int cur_h, int plane, int px, int py, int w, int h, int mv_row, int mv_col,
                              const uint8_t filt[2], int round0, int round1, int* out) {
    const int sx = plane ? ref.g.subx : 0, sy = plane ? ref.g.suby : 0;
    const Plane& rp = ref.p[plane];
    const int lastx = ref.g.w[plane] - 1, lasty = ref.g.h[plane] - 1;
    const int xscale = (int)((((int64_t)ref.g.w[0] << 14) + cur_w / 2) / cur_w), yscale = (int)((((int64_t)ref.g.h[0] << 14) + cur_h / 2) / cur_h);
    const int half = 8;
    const int64_t origx = ((int64_t)px << 4) + ((2 * mv_col) >> sx) + half, origy = ((int64_t)py << 4) + ((2 * mv_row) >> sy) + half;
    const int64_t basex = origx * xscale - ((int64_t)half << 14), basey = origy * yscale - ((int64_t)half << 14);
    const int startx = (int)(round2s64(basex, 14 + 4 - 10) + 32), starty = (int)(round2s64(basey, 14 + 4 - 10) + 32);
    const int stepx = (int)round2s64(xscale, 4), stepy = (int)round2s64(yscale, 4);
    const int ih = (((h - 1) * stepy + (1 << 10) - 1) >> 10) + 8;
    const int fih = filter_index(filt[1], w), fiv = filter_index(filt[0], h);
    std::vector<int> inter((size_t)ih * w);
    for (int r = 0; r < ih; r++) {
        const int yy = clip3(0, lasty, (starty >> 10) + r - 3);
        for (int c = 0; c < w; c++) {
            const int p = startx + stepx * c;
            const int16_t* f = av1t_subpel_filters[fih][(p >> 6) & 15];
            int s = 0;
            for (int t = 0; t < 8; t++) s += f[t] * rp.at(clip3(0, lastx, (p >> 10) + t - 3), yy);
            inter[(size_t)r * w + c] = round2(s, round0);
        }
    }
    for (int r = 0; r < h; r++) {
        const int p = (starty & 1023) + stepy * r;
        const int16_t* f = av1t_subpel_filters[fiv][(p >> 6) & 15];
        for (int c = 0; c < w; c++) {
            int s = 0;
            for (int t = 0; t < 8; t++) s += f[t] * inter[(size_t)((p >> 10) + t) * w + c];
            out[r * w + c] = round2(s, round1);
        }
    }
}

// 7.11.3.3 + 7.11.3.4 (unscaled references): out[h][w] at intermediate precision
static void block_pred(const Frame& ref, int cur_w, int cur_h, int plane, int px, int py, int w, int h, int mv_row, int mv_col, const uint8_t filt[2],
                       int round0, int round1, int* out) {
    if (ref.g.w[0] != cur_w || ref.g.h[0] != cur_h) {
        block_pred_scaled(ref, cur_w, cur_h, plane, px, py, w, h, mv_row, mv_col, filt, round0, round1, out);
        return;
    }
    const int sx = plane ? ref.g.subx : 0, sy = plane ? ref.g.suby : 0;
    const Plane& rp = ref.p[plane];
    const int lastx = ref.g.w[plane] - 1, lasty = ref.g.h[plane] - 1;
    const int posx = (px << 4) + ((2 * mv_col) >> sx), posy = (py << 4) + ((2 * mv_row) >> sy);
    const int ix = posx >> 4, fx = posx & 15, iy = posy >> 4, fy = posy & 15;
    const int16_t* fh = av1t_subpel_filters[filter_index(filt[1], w)][fx];
    const int16_t* fv = av1t_subpel_filters[filter_index(filt[0], h)][fy];
    std::vector<int> inter((size_t)(h + 7) * w);
    for (int r = 0; r < h + 7; r++) {
        const int yy = clip3(0, lasty, iy + r - 3);
        for (int c = 0; c < w; c++) {
            int s = 0;
            for (int t = 0; t < 8; t++) s += fh[t] * rp.at(clip3(0, lastx, ix + c + t - 3), yy);
            inter[(size_t)r * w + c] = round2(s, round0);
        }
    }
    for (int r = 0; r < h; r++)
        for (int c = 0; c < w; c++) {
            int s = 0;
            for (int t = 0; t < 8; t++) s += fv[t] * inter[(size_t)(r + t) * w + c];
            out[r * w + c] = round2(s, round1);
        }
}

// 7.11.3.5 block warp, whole block in 8x8 units
static void warp_pred(const Frame& ref, int plane, int px, int py, int w, int h, const WarpRec& wr, int round0, int round1, int* out) {
    const int sx = plane ? ref.g.subx : 0, sy = plane ? ref.g.suby : 0;
    const Plane& rp = ref.p[plane];
    const int lastx = ref.g.w[plane] - 1, lasty = ref.g.h[plane] - 1;
    for (int i8 = 0; i8 < h / 8; i8++)
        for (int j8 = 0; j8 < w / 8; j8++) {
            const int src_x = (px + j8 * 8 + 4) << sx, src_y = (py + i8 * 8 + 4) << sy;
            const int64_t dst_x = (int64_t)wr.mat[2] * src_x + (int64_t)wr.mat[3] * src_y + wr.mat[0];
            const int64_t dst_y = (int64_t)wr.mat[4] * src_x + (int64_t)wr.mat[5] * src_y + wr.mat[1];
            const int64_t x4 = dst_x >> sx, y4 = dst_y >> sy;
            const int ix4 = (int)(x4 >> 16), sx4 = (int)(x4 & 0xFFFF), iy4 = (int)(y4 >> 16), sy4 = (int)(y4 & 0xFFFF);
            int inter[15][8];
            for (int i1 = -7; i1 < 8; i1++)
                for (int i2 = -4; i2 < 4; i2++) {
                    const int sxx = sx4 + wr.alpha * i2 + wr.beta * i1;
                    const int offs = round2(sxx, 10) + 64;
                    int s = 0;
                    for (int i3 = 0; i3 < 8; i3++)
                        s += av1t_warped_filter[offs][i3] * rp.at(clip3(0, lastx, ix4 + i2 - 3 + i3), clip3(0, lasty, iy4 + i1));
                    inter[i1 + 7][i2 + 4] = round2(s, round0);
                }
            for (int i1 = -4; i1 < 4; i1++)
                for (int i2 = -4; i2 < 4; i2++) {
                    const int syy = sy4 + wr.gamma * i2 + wr.delta * i1;
                    const int offs = round2(syy, 10) + 64;
                    int s = 0;
                    for (int i3 = 0; i3 < 8; i3++) s += av1t_warped_filter[offs][i3] * inter[i1 + i3 + 4][i2 + 4];
                    out[(i8 * 8 + i1 + 4) * w + j8 * 8 + i2 + 4] = round2(s, round1);
                }
        }
}

void predict_inter_frame(const FrameWork& fw, Frame& f, const Frame* const refs[8]) {
    init_master();
    const FrameGeom& g = f.g;
    const int pixmax = (1 << g.bd) - 1;
    std::vector<int> p0(128 * 128), p1(128 * 128), ob(128 * 128);
    std::vector<uint8_t> mask(128 * 128);
    for (const InterBlk& r : fw.inter) {
        const int is_compound = r.ref[1] >= 0;
        const int round0 = 3, round1 = is_compound ? 7 : 11, post = 14 - round0 - round1;
        for (int plane = 0; plane < 3; plane++) {
            if (plane == 0 && !(r.planes & 1)) continue;
            if (plane > 0 && !(r.planes & 2)) continue;
            const int sx = plane ? g.subx : 0, sy = plane ? g.suby : 0;
            const int px = r.x >> sx, py = r.y >> sy, pw = r.w >> sx, ph = r.h >> sy;
            Plane& cur = f.p[plane];
            int* preds[2] = {p0.data(), p1.data()};
            for (int l = 0; l < 1 + is_compound; l++) {
                const Frame& ref = *refs[r.ref[l]];
                if (r.warp[l] >= 0 && pw >= 8 && ph >= 8) warp_pred(ref, plane, px, py, pw, ph, fw.warps[r.warp[l]], round0, round1, preds[l]);
                else block_pred(ref, g.w[0], g.h[0], plane, px, py, pw, ph, r.mv[l][0], r.mv[l][1], r.filt, round0, round1, preds[l]);
            }
            const int xe = std::min(pw, g.cw[plane] - px), ye = std::min(ph, g.ch[plane] - py);
            if (!is_compound) {
                for (int i = 0; i < ye; i++)
                    for (int j = 0; j < xe; j++) cur.at(px + j, py + i) = (uint16_t)clip3(0, pixmax, p0[i * pw + j]);
            } else {
                if (r.comp_type == COMPOUND_DIFFWTD && plane == 0) {
                    for (int i = 0; i < ph; i++)
                        for (int j = 0; j < pw; j++) {
                            int diff = abs(p0[i * pw + j] - p1[i * pw + j]);
                            diff = round2(diff, (g.bd - 8) + post);
                            int m = clip3(0, 64, 38 + diff / 16);
                            if (r.mask_type) m = 64 - m;
                            mask[i * 128 + j] = (uint8_t)m;
                        }
                } else if (r.comp_type == COMPOUND_WEDGE && plane == 0) {
                    for (int i = 0; i < ph; i++)
                        for (int j = 0; j < pw; j++) mask[i * 128 + j] = (uint8_t)wedge_mask(r.bsize, r.wedge_sign, r.wedge_index, i, j);
                }
                for (int i = 0; i < ye; i++)
                    for (int j = 0; j < xe; j++) {
                        const int a = p0[i * pw + j], b = p1[i * pw + j];
                        int v;
                        if (r.comp_type == COMPOUND_WEDGE || r.comp_type == COMPOUND_DIFFWTD) {
                            int m;
                            if (!sx && !sy) m = mask[i * 128 + j];
                            else if (sx && !sy) m = round2(mask[i * 128 + 2 * j] + mask[i * 128 + 2 * j + 1], 1);
                            else m = round2(mask[2 * i * 128 + 2 * j] + mask[2 * i * 128 + 2 * j + 1] + mask[(2 * i + 1) * 128 + 2 * j] +
                                                mask[(2 * i + 1) * 128 + 2 * j + 1], 2);
                            v = round2(m * a + (64 - m) * b, 6 + post);
                        } else if (r.comp_type == COMPOUND_DISTANCE) {
                            v = round2(a * r.fwd_w + b * r.bck_w, 4 + post);
                        } else {
                            v = round2(a + b, 1 + post);
                        }
                        cur.at(px + j, py + i) = (uint16_t)clip3(0, pixmax, v);
                    }
            }
            // overlapped motion compensation
            if (r.obmc_above + r.obmc_left > 0) {
                for (int k = 0; k < r.obmc_above + r.obmc_left; k++) {
                    const int above = k < r.obmc_above;
                    if (above && plane > 0 && !r.obmc_chroma_above) continue;
                    const ObmcNb& nb = fw.obmc[r.obmc_first + k];
                    int ow, oh;
                    if (above) {
                        ow = std::min(pw, (nb.step4 * 4) >> sx);
                        oh = std::min(ph >> 1, 32 >> sy);
                    } else {
                        ow = std::min(pw >> 1, 32 >> sx);
                        oh = std::min(ph, (nb.step4 * 4) >> sy);
                    }
                    const int ox = (nb.x4 * 4) >> sx, oy = (nb.y4 * 4) >> sy;
                    block_pred(*refs[nb.ref], g.w[0], g.h[0], plane, ox, oy, ow, oh, nb.mv[0], nb.mv[1], nb.filt, 3, 11, ob.data());
                    int lg = 0;
                    while ((1 << lg) < (above ? oh : ow)) lg++;
                    const uint8_t* m = av1t_obmc_mask[lg];
                    for (int i = 0; i < oh && oy + i < g.ch[plane]; i++)
                        for (int j = 0; j < ow && ox + j < g.cw[plane]; j++) {
                            const int o = clip3(0, pixmax, ob[i * ow + j]);
                            const int mm = above ? m[i] : m[j];
                            cur.at(ox + j, oy + i) = (uint16_t)round2(mm * cur.at(ox + j, oy + i) + (64 - mm) * o, 6);
                        }
                }
            }
        }
    }
}

// inter-intra blend of one plane block: intra[h][w] over the inter predictor already in the frame (7.11.3.13/14)
void interintra_blend(Frame& f, const TxRec& r, const int* intra) {
    init_master();
    const FrameGeom& g = f.g;
    const int plane = r.plane;
    const int sx = plane ? g.subx : 0, sy = plane ? g.suby : 0;
    const int w = kTxW[r.txsz], h = kTxH[r.txsz];
    const int x = r.x4 * 4, y = r.y4 * 4;
    const int pk = (uint16_t)r.cfl_alpha;
    const int wedge = pk & 1, wedge_index = (pk >> 1) & 15, ii_mode = (pk >> 5) & 3, bsize = (pk >> 7) & 31;
    Plane& cur = f.p[plane];
    const int size_scale = 128 / std::max(h, w);
    for (int i = 0; i < h && y + i < g.ch[plane]; i++)
        for (int j = 0; j < w && x + j < g.cw[plane]; j++) {
            int m;
            if (wedge) {
                if (!sx && !sy) m = wedge_mask(bsize, 0, wedge_index, i, j);
                else if (sx && !sy) m = round2(wedge_mask(bsize, 0, wedge_index, i, 2 * j) + wedge_mask(bsize, 0, wedge_index, i, 2 * j + 1), 1);
                else m = round2(wedge_mask(bsize, 0, wedge_index, 2 * i, 2 * j) + wedge_mask(bsize, 0, wedge_index, 2 * i, 2 * j + 1) +
                                    wedge_mask(bsize, 0, wedge_index, 2 * i + 1, 2 * j) + wedge_mask(bsize, 0, wedge_index, 2 * i + 1, 2 * j + 1), 2);
            } else if (ii_mode == II_V_PRED) {
                m = av1t_ii_weights1d[i * size_scale];
            } else if (ii_mode == II_H_PRED) {
                m = av1t_ii_weights1d[j * size_scale];
            } else if (ii_mode == II_SMOOTH_PRED) {
                m = av1t_ii_weights1d[std::min(i, j) * size_scale];
            } else {
                m = 32;
            }
            const int inter = cur.at(x + j, y + i);
            cur.at(x + j, y + i) = (uint16_t)round2(m * intra[i * w + j] + (64 - m) * inter, 6);
        }
    (void)kWedgeBitsO;
}

}  // namespace orc
