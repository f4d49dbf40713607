// TEST INFRASTRUCTURE ONLY -- scalar CPU restatement of the in-loop filters:
//   deblocking (AV1 spec 7.14: edge loop filter, filter size / level / limits, narrow + wide filters)
//   CDEF       (7.15: direction search, primary/secondary taps with damping, min/max clamp)
// consuming the per-4x4 edge descriptors / cdef indices produced by the product's host parser.
// Pinned against dav1d 1.5.3 with inloop_filters = 1 and 3 (tests/test_decode_intra.py) and, at unit
// level, against libaom's aom_lpf_*_c / cdef_filter_*_c (tests/test_filters_unit.py).
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>

#include "oracle_frame.h"

namespace orc {
using namespace av1r;

static inline int clip3(int lo, int hi, int x) { return x < lo ? lo : (x > hi ? hi : x); }
static inline int round2(int x, int n) { return n == 0 ? x : (x + (1 << (n - 1))) >> n; }

// Filter the 1 sample-wide line across an edge.  px points at q0; step = distance between p/q samples.
void lf_sample(uint16_t* q0p, ptrdiff_t step, int plane, int fsz, int lvl, int sharpness, int bd) {
    const int shift = sharpness > 4 ? 2 : (sharpness > 0 ? 1 : 0);
    const int limit = sharpness > 0 ? clip3(1, 9 - sharpness, lvl >> shift) : std::max(1, lvl >> shift);
    const int blimit = 2 * (lvl + 2) + limit;
    const int thresh = lvl >> 4;
    const int s = bd - 8;
    const int limit_bd = limit << s, blimit_bd = blimit << s, thresh_bd = thresh << s, one = 1 << s;
    int q[7], p[7];
    const int flen = fsz == 4 ? 4 : (plane != 0 ? 6 : (fsz == 8 ? 8 : 16));
    const int nread = flen == 4 ? 2 : (flen == 6 ? 3 : (flen == 8 ? 4 : 7));
    for (int i = 0; i < nread; i++) {
        q[i] = q0p[i * step];
        p[i] = q0p[-(i + 1) * step];
    }
    const int hev = abs(p[1] - p[0]) > thresh_bd || abs(q[1] - q[0]) > thresh_bd;
    int mask = abs(p[1] - p[0]) > limit_bd || abs(q[1] - q[0]) > limit_bd || (abs(p[0] - q[0]) * 2 + abs(p[1] - q[1]) / 2) > blimit_bd;
    if (flen >= 6) mask |= abs(p[2] - p[1]) > limit_bd || abs(q[2] - q[1]) > limit_bd;
    if (flen >= 8) mask |= abs(p[3] - p[2]) > limit_bd || abs(q[3] - q[2]) > limit_bd;
    if (mask) return;
    int flat = 0, flat2 = 0;
    if (flen >= 6) {
        flat = abs(p[1] - p[0]) <= one && abs(q[1] - q[0]) <= one && abs(p[2] - p[0]) <= one && abs(q[2] - q[0]) <= one;
        if (flen >= 8) flat = flat && abs(p[3] - p[0]) <= one && abs(q[3] - q[0]) <= one;
    }
    if (flen >= 16) {
        flat2 = 1;
        for (int i = 4; i < 7; i++) flat2 = flat2 && abs(p[i] - p[0]) <= one && abs(q[i] - q[0]) <= one;
    }
    if (fsz == 4 || !flat) {
        // narrow filter
        const int lo = -(1 << (bd - 1)), hi = (1 << (bd - 1)) - 1, off = 0x80 << s;
        const int ps1 = p[1] - off, ps0 = p[0] - off, qs0 = q[0] - off, qs1 = q[1] - off;
        int f = hev ? clip3(lo, hi, ps1 - qs1) : 0;
        f = clip3(lo, hi, f + 3 * (qs0 - ps0));
        const int f1 = clip3(lo, hi, f + 4) >> 3, f2 = clip3(lo, hi, f + 3) >> 3;
        q0p[0] = (uint16_t)(clip3(lo, hi, qs0 - f1) + off);
        q0p[-step] = (uint16_t)(clip3(lo, hi, ps0 + f2) + off);
        if (!hev) {
            const int f3 = round2(f1, 1);
            q0p[step] = (uint16_t)(clip3(lo, hi, qs1 - f3) + off);
            q0p[-2 * step] = (uint16_t)(clip3(lo, hi, ps1 + f3) + off);
        }
        return;
    }
    // wide filters
    int log2size, n, n2;
    if (fsz == 8 || !flat2) {
        log2size = 3;
        n = plane == 0 ? 3 : 2;
        n2 = plane == 0 ? 0 : 1;
    } else {
        log2size = 4;
        n = 6;
        n2 = 1;
    }
    auto at = [&](int pos) -> int { return pos >= 0 ? q[pos] : p[-pos - 1]; };
    int F[12];
    for (int i = -n; i < n; i++) {
        int t = 0;
        for (int j = -n; j <= n; j++) {
            int pp = clip3(-(n + 1), n, i + j);
            int tap = abs(j) <= n2 ? 2 : 1;
            t += at(pp) * tap;
        }
        F[i + n] = round2(t, log2size);
    }
    for (int i = -n; i < n; i++) q0p[i * step] = (uint16_t)F[i + n];
}

void deblock_frame(const FrameWork& fw, Frame& f) {
    const FrameGeom& g = f.g;
    const int sharp = fw.fh.lf.sharpness;
    for (int plane = 0; plane < 3; plane++) {
        if (plane > 0 && !fw.fh.lf.level[1 + plane]) continue;
        Plane& pl = f.p[plane];
        const int pw4 = fw.plane_w4(plane), ph4 = fw.plane_h4(plane);
        for (int pass = 0; pass < 2; pass++)
            for (int r4 = 0; r4 < ph4; r4++)
                for (int c4 = 0; c4 < pw4; c4++) {
                    const LfEdge& e = fw.lf[plane][(size_t)r4 * pw4 + c4];
                    const int len = pass ? e.len_h : e.len_v, lvl = pass ? e.lvl_h : e.lvl_v;
                    if (!len) continue;
                    for (int i = 0; i < 4; i++) {
                        const int x = c4 * 4 + (pass ? i : 0), y = r4 * 4 + (pass ? 0 : i);
                        if (x >= g.cw[plane] || y >= g.ch[plane]) continue;
                        lf_sample(&pl.at(x, y), pass ? pl.stride : 1, plane, len, lvl, sharp, g.bd);
                    }
                }
    }
}

// ------------------------------------------------------------------------------------ CDEF
static const int8_t kCdefDir[8][2][2] = {{{-1, 1}, {-2, 2}}, {{0, 1}, {-1, 2}}, {{0, 1}, {0, 2}}, {{0, 1}, {1, 2}},
                                         {{1, 1}, {2, 2}},   {{1, 0}, {2, 1}},  {{1, 0}, {2, 0}}, {{1, 0}, {2, -1}}};
static const int kCdefUvDir[2][2][8] = {{{0, 1, 2, 3, 4, 5, 6, 7}, {1, 2, 2, 2, 3, 4, 6, 0}}, {{7, 0, 2, 4, 5, 6, 6, 6}, {0, 1, 2, 3, 4, 5, 6, 7}}};

static inline int floor_log2(unsigned x) { int s = 0; while (x > 1) { x >>= 1; s++; } return s; }

static int constrain(int diff, int threshold, int damping) {
    if (!threshold) return 0;
    const int adj = std::max(0, damping - floor_log2(threshold));
    const int mag = abs(diff);
    const int v = clip3(0, mag, threshold - (mag >> adj));
    return diff < 0 ? -v : v;
}

void cdef_direction(const Plane& pl, int x0, int y0, int bd, int* dir_out, int* var_out) {
    static const int div_table[9] = {0, 840, 420, 280, 210, 168, 140, 120, 105};
    int cost[8] = {0};
    int partial[8][15] = {{0}};
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) {
            const int x = (pl.at(x0 + j, y0 + i) >> (bd - 8)) - 128;
            partial[0][i + j] += x;
            partial[1][i + j / 2] += x;
            partial[2][i] += x;
            partial[3][3 + i - j / 2] += x;
            partial[4][7 + i - j] += x;
            partial[5][3 - i / 2 + j] += x;
            partial[6][j] += x;
            partial[7][i / 2 + j] += x;
        }
    for (int i = 0; i < 8; i++) {
        cost[2] += partial[2][i] * partial[2][i];
        cost[6] += partial[6][i] * partial[6][i];
    }
    cost[2] *= div_table[8];
    cost[6] *= div_table[8];
    for (int i = 0; i < 7; i++) {
        cost[0] += (partial[0][i] * partial[0][i] + partial[0][14 - i] * partial[0][14 - i]) * div_table[i + 1];
        cost[4] += (partial[4][i] * partial[4][i] + partial[4][14 - i] * partial[4][14 - i]) * div_table[i + 1];
    }
    cost[0] += partial[0][7] * partial[0][7] * div_table[8];
    cost[4] += partial[4][7] * partial[4][7] * div_table[8];
    for (int i = 1; i < 8; i += 2) {
        for (int j = 0; j < 5; j++) cost[i] += partial[i][3 + j] * partial[i][3 + j];
        cost[i] *= div_table[8];
        for (int j = 0; j < 3; j++)
            cost[i] += (partial[i][j] * partial[i][j] + partial[i][10 - j] * partial[i][10 - j]) * div_table[2 * j + 2];
    }
    int best = 0, dir = 0;
    for (int d = 0; d < 8; d++)
        if (cost[d] > best) { best = cost[d]; dir = d; }
    *dir_out = dir;
    *var_out = (best - cost[(dir + 4) & 7]) >> 10;
}

static void cdef_filter_block(const Plane& in, Plane& out, const FrameGeom& g, int plane, int x0, int y0, int w, int h,
                              int pri, int sec, int damping, int dir) {
    static const int pri_taps[2][2] = {{4, 2}, {3, 3}}, sec_taps[2][2] = {{2, 1}, {2, 1}};
    const int cs = g.bd - 8;
    const int cwid = g.cw[plane], chei = g.ch[plane];
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            const int x = in.at(x0 + j, y0 + i);
            int sum = 0, mx = x, mn = x;
            for (int k = 0; k < 2; k++)
                for (int sign = -1; sign <= 1; sign += 2) {
                    {
                        const int yy = y0 + i + sign * kCdefDir[dir][k][0], xx = x0 + j + sign * kCdefDir[dir][k][1];
                        if (yy >= 0 && yy < chei && xx >= 0 && xx < cwid) {
                            const int p = in.at(xx, yy);
                            sum += pri_taps[(pri >> cs) & 1][k] * constrain(p - x, pri, damping);
                            mx = std::max(mx, p);
                            mn = std::min(mn, p);
                        }
                    }
                    for (int off = -2; off <= 2; off += 4) {
                        const int d2 = (dir + off) & 7;
                        const int yy = y0 + i + sign * kCdefDir[d2][k][0], xx = x0 + j + sign * kCdefDir[d2][k][1];
                        if (yy >= 0 && yy < chei && xx >= 0 && xx < cwid) {
                            const int s = in.at(xx, yy);
                            sum += sec_taps[(pri >> cs) & 1][k] * constrain(s - x, sec, damping);
                            mx = std::max(mx, s);
                            mn = std::min(mn, s);
                        }
                    }
                }
            out.at(x0 + j, y0 + i) = (uint16_t)clip3(mn, mx, x + ((8 + sum - (sum < 0)) >> 4));
        }
}

void cdef_frame(const FrameWork& fw, const Frame& in, Frame& out) {
    const FrameGeom& g = in.g;
    const FrameHdr& fh = fw.fh;
    const int c64 = (fw.mi_cols + 15) >> 4;
    const int cs = g.bd - 8;
    for (int r = 0; r < fw.mi_rows; r += 2)
        for (int c = 0; c < fw.mi_cols; c += 2) {
            const int idx = fw.cdef_idx[(size_t)(r >> 4) * c64 + (c >> 4)];
            if (idx < 0) continue;
            const uint8_t* s0 = &fw.skip_mi[(size_t)r * fw.mi_cols + c];
            const uint8_t* s1 = &fw.skip_mi[(size_t)(r + 1) * fw.mi_cols + c];
            if (s0[0] && s0[1] && s1[0] && s1[1]) continue;
            int ydir, var;
            cdef_direction(in.p[0], c * 4, r * 4, g.bd, &ydir, &var);
            {
                int pri = fh.cdef_y_pri[idx] << cs, sec = fh.cdef_y_sec[idx] << cs;
                const int dir = pri == 0 ? 0 : ydir;
                const int var_str = (var >> 6) ? std::min(floor_log2(var >> 6), 12) : 0;
                pri = var ? (pri * (4 + var_str) + 8) >> 4 : 0;
                cdef_filter_block(in.p[0], out.p[0], g, 0, c * 4, r * 4, 8, 8, pri, sec, fh.cdef_damping + cs, dir);
            }
            if (g.mono) continue;
            const int pri = fh.cdef_uv_pri[idx] << cs, sec = fh.cdef_uv_sec[idx] << cs;
            const int dir = pri == 0 ? 0 : kCdefUvDir[g.subx][g.suby][ydir];
            for (int plane = 1; plane < 3; plane++)
                cdef_filter_block(in.p[plane], out.p[plane], g, plane, (c * 4) >> g.subx, (r * 4) >> g.suby, 8 >> g.subx, 8 >> g.suby, pri,
                                  sec, fh.cdef_damping + cs - 1, dir);
        }
}

// ------------------------------------------------------------------------------------ loop restoration
// AV1 spec 7.17: 64-row stripes offset by 8 luma rows; rows outside the stripe come from the deblocked
// (pre-CDEF) frame, at most 2 rows deep; Wiener 7-tap separable, self-guided with two box radii.
#include "../av1-go_b200/csrc/tables/tables_filter.inc"

struct LrCtx {
    const Plane* cdef;
    const Plane* deblocked;
    int plane_end_x, plane_end_y, stripe_start, stripe_end;
    int sample(int x, int y) const {
        x = std::max(0, std::min(plane_end_x, x));
        y = std::max(0, std::min(plane_end_y, y));
        if (y < stripe_start) {
            y = std::max(stripe_start - 2, y);
            return deblocked->at(x, y);
        }
        if (y > stripe_end) {
            y = std::min(stripe_end + 2, y);
            return deblocked->at(x, y);
        }
        return cdef->at(x, y);
    }
};

static void wiener_block(const LrCtx& c, const LrUnit& u, int bd, int x, int y, int w, int h, Plane& out) {
    const int round0 = bd == 12 ? 5 : 3, round1 = bd == 12 ? 9 : 11;
    int vf[7], hf[7];
    auto get_filter = [](const int8_t* co, int* f) {
        f[3] = 128;
        for (int i = 0; i < 3; i++) {
            f[i] = f[6 - i] = co[i];
            f[3] -= 2 * co[i];
        }
    };
    get_filter(u.wiener[0], vf);
    get_filter(u.wiener[1], hf);
    const int offset = 1 << (bd + 7 - round0 - 1);
    const int limit = (1 << (bd + 1 + 7 - round0)) - 1;
    std::vector<int> inter((size_t)(h + 6) * w);
    for (int r = 0; r < h + 6; r++)
        for (int cc = 0; cc < w; cc++) {
            int s = 0;
            for (int t = 0; t < 7; t++) s += hf[t] * c.sample(x + cc + t - 3, y + r - 3);
            int v = round2(s, round0);
            inter[(size_t)r * w + cc] = clip3(-offset, limit - offset, v);
        }
    const int pixmax = (1 << bd) - 1;
    for (int r = 0; r < h; r++)
        for (int cc = 0; cc < w; cc++) {
            int s = 0;
            for (int t = 0; t < 7; t++) s += vf[t] * inter[(size_t)(r + t) * w + cc];
            out.at(x + cc, y + r) = (uint16_t)clip3(0, pixmax, round2(s, round1));
        }
}

static void sgr_box(const LrCtx& c, int bd, int x, int y, int w, int h, int r, int s_param, int pass, std::vector<int>& F) {
    const int n = (2 * r + 1) * (2 * r + 1);
    const int one_by_n = ((1 << 12) + n / 2) / n;
    std::vector<int> A((size_t)(h + 2) * (w + 2)), B((size_t)(h + 2) * (w + 2));
    for (int i = -1; i <= h; i++)
        for (int j = -1; j <= w; j++) {
            uint32_t a = 0, b = 0;
            for (int dy = -r; dy <= r; dy++)
                for (int dx = -r; dx <= r; dx++) {
                    const uint32_t v = (uint32_t)c.sample(x + j + dx, y + i + dy);
                    a += v * v;
                    b += v;
                }
            a = (uint32_t)round2((int)a, 2 * (bd - 8));
            const uint32_t d = (uint32_t)round2((int)b, bd - 8);
            const uint32_t p = a * n < d * d ? 0 : a * n - d * d;
            const uint32_t z = (uint32_t)(((uint64_t)p * s_param + (1 << 19)) >> 20);
            uint32_t a2;
            if (z >= 255) a2 = 256;
            else if (z == 0) a2 = 1;
            else a2 = ((z << 8) + z / 2) / (z + 1);
            const uint32_t b2 = (256 - a2) * b * one_by_n;
            A[(size_t)(i + 1) * (w + 2) + j + 1] = (int)a2;
            B[(size_t)(i + 1) * (w + 2) + j + 1] = (int)((b2 + (1 << 11)) >> 12);
        }
    F.assign((size_t)w * h, 0);
    for (int i = 0; i < h; i++) {
        int shift = 5;
        if (pass == 0 && (i & 1)) shift = 4;
        for (int j = 0; j < w; j++) {
            int a = 0, b = 0;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    int weight;
                    if (pass == 0) {
                        if ((i + dy) & 1) weight = dx == 0 ? 6 : 5;
                        else weight = 0;
                    } else {
                        weight = (dx == 0 || dy == 0) ? 4 : 3;
                    }
                    a += weight * A[(size_t)(i + 1 + dy) * (w + 2) + j + 1 + dx];
                    b += weight * B[(size_t)(i + 1 + dy) * (w + 2) + j + 1 + dx];
                }
            const int v = a * c.cdef->at(x + j, y + i) + b;
            F[(size_t)i * w + j] = round2(v, 8 + shift - 4);
        }
    }
}

static void sgr_block(const LrCtx& c, const LrUnit& u, int bd, int x, int y, int w, int h, Plane& out) {
    const int r0 = av1t_sgr_params[u.sgr_set][0], r1 = av1t_sgr_params[u.sgr_set][1];
    const int s0 = av1t_sgr_params[u.sgr_set][2], s1 = av1t_sgr_params[u.sgr_set][3];
    std::vector<int> f0, f1;
    if (r0) sgr_box(c, bd, x, y, w, h, r0, s0, 0, f0);
    if (r1) sgr_box(c, bd, x, y, w, h, r1, s1, 1, f1);
    const int w0 = u.sgr_xqd[0], w1 = u.sgr_xqd[1], w2 = 128 - w0 - w1;
    const int pixmax = (1 << bd) - 1;
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            const int uu = c.cdef->at(x + j, y + i) << 4;
            int v = w1 * uu;
            v += w0 * (r0 ? f0[(size_t)i * w + j] : uu);
            v += w2 * (r1 ? f1[(size_t)i * w + j] : uu);
            out.at(x + j, y + i) = (uint16_t)clip3(0, pixmax, round2(v, 11));
        }
}

// Upscaling process, AV1 spec 7.16: horizontal 8-tap filter with 1/16384-sample positions, per plane.
void upscale_frame(const FrameHdr& fh, const Frame& in, Frame& out) {
    const FrameGeom& g = in.g;
    for (int plane = 0; plane < (g.mono ? 1 : 3); plane++) {
        const int sx = plane ? g.subx : 0;
        const int down_w = (fh.frame_width + sx) >> sx, up_w = (fh.upscaled_width + sx) >> sx;
        const int plane_h = out.g.h[plane];
        const int step_x = ((down_w << 14) + (up_w / 2)) / up_w;
        const int err = up_w * step_x - (down_w << 14);
        const int initial_subpel_x = ((-((up_w - down_w) << 13) + up_w / 2) / up_w + (1 << 7) - err / 2) & ((1 << 14) - 1);
        const int max_x = g.cw[plane] - 1;     // (MiCols >> subX) * MI_SIZE - 1
        const int pixmax = (1 << g.bd) - 1;
        for (int y = 0; y < plane_h; y++)
            for (int x = 0; x < up_w; x++) {
                const int src_x = -(1 << 14) + initial_subpel_x + x * step_x;
                const int src_px = src_x >> 14, sub = (src_x & ((1 << 14) - 1)) >> 8;
                int sum = 0;
                for (int k = 0; k < 8; k++) {
                    const int sxp = std::min(std::max(src_px + k - 3, 0), max_x);
                    sum += in.p[plane].at(sxp, y) * av1t_upscale_filter[sub][k];
                }
                out.p[plane].at(x, y) = (uint16_t)std::min(std::max((sum + 64) >> 7, 0), pixmax);
            }
    }
}

void lr_frame(const FrameWork& fw, const Frame& deblocked, const Frame& cdef, Frame& out) {
    const FrameGeom& g = cdef.g;
    out = cdef;
    for (int plane = 0; plane < (g.mono ? 1 : 3); plane++) {
        if (fw.fh.lr_type[plane] == RESTORE_NONE) continue;
        const int sx = plane ? g.subx : 0, sy = plane ? g.suby : 0;
        const int unit_size = fw.fh.lr_size[plane];
        const int pw = g.w[plane], ph = g.h[plane];
        const int unit_rows = fw.lr_rows[plane], unit_cols = fw.lr_cols[plane];
        if (fw.lr[plane].empty()) continue;
        LrCtx c;
        c.cdef = &cdef.p[plane];
        c.deblocked = &deblocked.p[plane];
        c.plane_end_x = pw - 1;
        c.plane_end_y = ph - 1;
        for (int stripe = 0;; stripe++) {
            const int ls = -8 + stripe * 64;                 // luma stripe start
            int ys = ls >> sy, ye = ys + (64 >> sy) - 1;      // plane rows [StripeStartY, StripeEndY]
            if (std::max(ys, 0) > ph - 1) break;
            c.stripe_start = ys;
            c.stripe_end = ye;
            const int y0 = std::max(ys, 0), y1 = std::min(ye, ph - 1);
            // all rows of a stripe belong to one unit row: ((lumaY + 8) >> sy) / unitSize
            const int unit_row = std::min(unit_rows - 1, ((std::max(ls, 0) + 8) >> sy) / unit_size);
            for (int uc = 0; uc < unit_cols; uc++) {
                const int x0 = uc * unit_size, x1 = uc == unit_cols - 1 ? pw - 1 : (uc + 1) * unit_size - 1;
                const LrUnit& u = fw.lr[plane][(size_t)unit_row * unit_cols + uc];
                if (u.type == RESTORE_WIENER) wiener_block(c, u, g.bd, x0, y0, x1 - x0 + 1, y1 - y0 + 1, out.p[plane]);
                else if (u.type == RESTORE_SGRPROJ) sgr_block(c, u, g.bd, x0, y0, x1 - x0 + 1, y1 - y0 + 1, out.p[plane]);
            }
        }
    }
}

}  // namespace orc
