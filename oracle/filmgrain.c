/* TEST INFRASTRUCTURE ONLY -- CPU restatement of AV1 film grain synthesis (K8).
 *
 * The reference repo has no pixel code (the daemon shells out to ffmpeg:
 * /root/reference/internal/ffmpeg/transcode.go:195); the arithmetic it relies on lives in the
 * third-party libdav1d inside that FFmpeg build (unpinned "latest" URL,
 * /root/reference/internal/config/config.go:33).  This file restates the published algorithm
 * (AV1 spec 7.18.3: random number process, generate grain, scaling LUT, add noise synthesis)
 * in scalar C.  It is pinned against dav1d 1.5.3 output (apply_grain 0 vs 1) by
 * tests/test_filmgrain.py and tests/golden/filmgrain_*.npz.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may link this.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../av1-go_b200/csrc/tables/tables_fg.inc"

typedef struct {
    int apply_grain, grain_seed, update_grain;
    int num_y_points, point_y_value[16], point_y_scaling[16];
    int chroma_scaling_from_luma;
    int num_cb_points, point_cb_value[16], point_cb_scaling[16];
    int num_cr_points, point_cr_value[16], point_cr_scaling[16];
    int grain_scaling;
    int ar_coeff_lag;
    int ar_coeffs_y[24], ar_coeffs_cb[25], ar_coeffs_cr[25];
    int ar_coeff_shift;
    int grain_scale_shift;
    int cb_mult, cb_luma_mult, cb_offset, cr_mult, cr_luma_mult, cr_offset;
    int overlap_flag, clip_to_restricted_range;
} orc_fg_params;

static int round2(int x, int n) { return n == 0 ? x : (x + (1 << (n - 1))) >> n; }
static int clip3(int lo, int hi, int x) { return x < lo ? lo : (x > hi ? hi : x); }

static unsigned rnd_reg;
static int get_random(int bits) {
    unsigned r = rnd_reg;
    unsigned bit = ((r >> 0) ^ (r >> 1) ^ (r >> 3) ^ (r >> 12)) & 1;
    r = (r >> 1) | (bit << 15);
    rnd_reg = r;
    return (r >> (16 - bits)) & ((1 << bits) - 1);
}

static void build_lut(int n, const int* val, const int* sc, int* lut) {
    if (n == 0) { memset(lut, 0, 256 * sizeof(int)); return; }
    for (int i = 0; i < val[0]; i++) lut[i] = sc[0];
    for (int i = 0; i < n - 1; i++) {
        int dy = sc[i + 1] - sc[i], dx = val[i + 1] - val[i];
        int delta = dy * ((65536 + (dx >> 1)) / dx);
        for (int x = 0; x < dx; x++) lut[val[i] + x] = sc[i] + ((x * delta + 32768) >> 16);
    }
    for (int i = val[n - 1]; i < 256; i++) lut[i] = sc[n - 1];
}

static int scale_lut(const int* lut, int index, int bd) {
    int shift = bd - 8;
    int x = index >> shift;
    int rem = index - (x << shift);
    if (bd == 8 || x == 255) return lut[x];
    int start = lut[x], end = lut[x + 1];
    return start + round2((end - start) * rem, shift);
}

static int px_get(const void* p, int stride, int x, int y, int bd) {
    if (bd == 8) return ((const uint8_t*)p)[(size_t)y * stride + x];
    return ((const uint16_t*)((const uint8_t*)p + (size_t)y * stride))[x];
}
static void px_put(void* p, int stride, int x, int y, int bd, int v) {
    if (bd == 8) ((uint8_t*)p)[(size_t)y * stride + x] = (uint8_t)v;
    else ((uint16_t*)((uint8_t*)p + (size_t)y * stride))[x] = (uint16_t)v;
}

/* strides in bytes */
void orc_film_grain(const orc_fg_params* g, int bd, int w, int h, int subx, int suby, int mono, int mc_identity,
                    const void* const in[3], const int in_stride[3], void* const out[3], const int out_stride[3]) {
    static int luma_grain[73][82], cb_grain[73][82], cr_grain[73][82];
    int lut[3][256];
    const int grain_center = 128 << (bd - 8);
    const int grain_min = -grain_center, grain_max = (256 << (bd - 8)) - 1 - grain_center;
    const int nplanes = mono ? 1 : 3;

    /* 7.18.3.3 generate grain */
    rnd_reg = g->grain_seed;
    int shift = 12 - bd + g->grain_scale_shift;
    for (int y = 0; y < 73; y++)
        for (int x = 0; x < 82; x++) {
            int v = g->num_y_points > 0 ? av1t_gaussian_sequence[get_random(11)] : 0;
            luma_grain[y][x] = round2(v, shift);
        }
    int lag = g->ar_coeff_lag;
    int ash = g->ar_coeff_shift;
    for (int y = 3; y < 73; y++)
        for (int x = 3; x < 82 - 3; x++) {
            int sum = 0, pos = 0;
            for (int dr = -lag; dr <= 0; dr++) {
                for (int dc = -lag; dc <= lag; dc++) {
                    if (dr == 0 && dc == 0) break;
                    sum += g->ar_coeffs_y[pos] * luma_grain[y + dr][x + dc];
                    pos++;
                }
            }
            luma_grain[y][x] = clip3(grain_min, grain_max, luma_grain[y][x] + round2(sum, ash));
        }
    int cw = subx ? 44 : 82, ch = suby ? 38 : 73;
    if (!mono) {
        rnd_reg = g->grain_seed ^ 0xb524;
        for (int y = 0; y < ch; y++)
            for (int x = 0; x < cw; x++) {
                int v = (g->num_cb_points || g->chroma_scaling_from_luma) ? av1t_gaussian_sequence[get_random(11)] : 0;
                cb_grain[y][x] = round2(v, shift);
            }
        rnd_reg = g->grain_seed ^ 0x49d8;
        for (int y = 0; y < ch; y++)
            for (int x = 0; x < cw; x++) {
                int v = (g->num_cr_points || g->chroma_scaling_from_luma) ? av1t_gaussian_sequence[get_random(11)] : 0;
                cr_grain[y][x] = round2(v, shift);
            }
        for (int y = 3; y < ch; y++)
            for (int x = 3; x < cw - 3; x++) {
                int s0 = 0, s1 = 0, pos = 0;
                for (int dr = -lag; dr <= 0; dr++) {
                    for (int dc = -lag; dc <= lag; dc++) {
                        int c0 = g->ar_coeffs_cb[pos], c1 = g->ar_coeffs_cr[pos];
                        if (dr == 0 && dc == 0) {
                            if (g->num_y_points > 0) {
                                int luma = 0;
                                int lx = ((x - 3) << subx) + 3, ly = ((y - 3) << suby) + 3;
                                for (int i = 0; i <= suby; i++)
                                    for (int j = 0; j <= subx; j++) luma += luma_grain[ly + i][lx + j];
                                luma = round2(luma, subx + suby);
                                s0 += luma * c0;
                                s1 += luma * c1;
                            }
                            break;
                        }
                        s0 += c0 * cb_grain[y + dr][x + dc];
                        s1 += c1 * cr_grain[y + dr][x + dc];
                        pos++;
                    }
                }
                cb_grain[y][x] = clip3(grain_min, grain_max, cb_grain[y][x] + round2(s0, ash));
                cr_grain[y][x] = clip3(grain_min, grain_max, cr_grain[y][x] + round2(s1, ash));
            }
    }
    /* 7.18.3.4 scaling LUTs */
    build_lut(g->num_y_points, g->point_y_value, g->point_y_scaling, lut[0]);
    if (g->chroma_scaling_from_luma) {
        memcpy(lut[1], lut[0], sizeof(lut[0]));
        memcpy(lut[2], lut[0], sizeof(lut[0]));
    } else {
        build_lut(g->num_cb_points, g->point_cb_value, g->point_cb_scaling, lut[1]);
        build_lut(g->num_cr_points, g->point_cr_value, g->point_cr_scaling, lut[2]);
    }
    /* 7.18.3.5 add noise: noise stripes */
    int nstripes = (h + 31) / 32;
    int sw = ((w + 1) / 2) * 2 + 64; /* stripe row width incl. slack */
    int* stripe[3];
    for (int p = 0; p < 3; p++) stripe[p] = (int*)calloc((size_t)nstripes * 34 * sw, sizeof(int));
#define STRIPE(p, n, i, x) stripe[p][((size_t)(n) * 34 + (i)) * sw + (x)]
    int luma_num = 0;
    for (int y = 0; y < (h + 1) / 2; y += 16) {
        rnd_reg = g->grain_seed;
        rnd_reg ^= ((luma_num * 37 + 178) & 255) << 8;
        rnd_reg ^= ((luma_num * 173 + 105) & 255);
        for (int x = 0; x < (w + 1) / 2; x += 16) {
            int rand = get_random(8);
            int offx = rand >> 4, offy = rand & 15;
            for (int p = 0; p < nplanes; p++) {
                int psx = p > 0 ? subx : 0, psy = p > 0 ? suby : 0;
                int pox = psx ? 6 + offx : 9 + offx * 2;
                int poy = psy ? 6 + offy : 9 + offy * 2;
                for (int i = 0; i < (34 >> psy); i++)
                    for (int j = 0; j < (34 >> psx); j++) {
                        int gr = p == 0 ? luma_grain[poy + i][pox + j] : (p == 1 ? cb_grain[poy + i][pox + j] : cr_grain[poy + i][pox + j]);
                        if (psx == 0) {
                            if (x * 2 + j >= sw) continue;
                            if (j < 2 && g->overlap_flag && x > 0) {
                                int old = STRIPE(p, luma_num, i, x * 2 + j);
                                if (j == 0) gr = old * 27 + gr * 17;
                                else gr = old * 17 + gr * 27;
                                gr = clip3(grain_min, grain_max, round2(gr, 5));
                            }
                            STRIPE(p, luma_num, i, x * 2 + j) = gr;
                        } else {
                            if (j == 0 && g->overlap_flag && x > 0) {
                                int old = STRIPE(p, luma_num, i, x + j);
                                gr = old * 23 + gr * 22;
                                gr = clip3(grain_min, grain_max, round2(gr, 5));
                            }
                            STRIPE(p, luma_num, i, x + j) = gr;
                        }
                    }
            }
        }
        luma_num++;
    }
    /* noise image + blend; chroma first (it reads un-noised luma) */
    int min_value, max_luma, max_chroma;
    if (g->clip_to_restricted_range) {
        min_value = 16 << (bd - 8);
        max_luma = 235 << (bd - 8);
        max_chroma = mc_identity ? max_luma : (240 << (bd - 8));
    } else {
        min_value = 0;
        max_luma = max_chroma = (256 << (bd - 8)) - 1;
    }
    int pixmax = (1 << bd) - 1;
    int sshift = g->grain_scaling;
    for (int p = nplanes - 1; p >= 0; p--) {
        int psx = p > 0 ? subx : 0, psy = p > 0 ? suby : 0;
        int pw = (w + psx) >> psx, ph = (h + psy) >> psy;
        for (int y = 0; y < ph; y++) {
            int ln = y >> (5 - psy);
            int i = y - (ln << (5 - psy));
            for (int x = 0; x < pw; x++) {
                int gr = STRIPE(p, ln, i, x);
                if (psy == 0) {
                    if (i < 2 && ln > 0 && g->overlap_flag) {
                        int old = STRIPE(p, ln - 1, i + 32, x);
                        if (i == 0) gr = old * 27 + gr * 17;
                        else gr = old * 17 + gr * 27;
                        gr = clip3(grain_min, grain_max, round2(gr, 5));
                    }
                } else {
                    if (i < 1 && ln > 0 && g->overlap_flag) {
                        int old = STRIPE(p, ln - 1, i + 16, x);
                        gr = old * 23 + gr * 22;
                        gr = clip3(grain_min, grain_max, round2(gr, 5));
                    }
                }
                int orig = px_get(in[p], in_stride[p], x, y, bd);
                int res = orig;
                if (p == 0) {
                    if (g->num_y_points > 0) {
                        int noise = round2(scale_lut(lut[0], orig, bd) * gr, sshift);
                        res = clip3(min_value, max_luma, orig + noise);
                    }
                } else {
                    int npts = p == 1 ? g->num_cb_points : g->num_cr_points;
                    if (npts > 0 || g->chroma_scaling_from_luma) {
                        int lx = x << subx, ly = y << suby;
                        int lnx = lx + 1 < w - 1 ? lx + 1 : w - 1;
                        int avg;
                        if (subx) avg = round2(px_get(in[0], in_stride[0], lx, ly, bd) + px_get(in[0], in_stride[0], lnx, ly, bd), 1);
                        else avg = px_get(in[0], in_stride[0], lx, ly, bd);
                        int merged;
                        if (g->chroma_scaling_from_luma) {
                            merged = avg;
                        } else {
                            int lm = p == 1 ? g->cb_luma_mult : g->cr_luma_mult;
                            int m = p == 1 ? g->cb_mult : g->cr_mult;
                            int off = p == 1 ? g->cb_offset : g->cr_offset;
                            int combined = avg * (lm - 128) + orig * (m - 128);
                            merged = clip3(0, pixmax, (combined >> 6) + ((off - 256) << (bd - 8)));
                        }
                        int noise = round2(scale_lut(lut[p], merged, bd) * gr, sshift);
                        res = clip3(min_value, max_chroma, orig + noise);
                    }
                }
                px_put(out[p], out_stride[p], x, y, bd, res);
            }
        }
    }
    for (int p = 0; p < 3; p++) free(stripe[p]);
#undef STRIPE
}
