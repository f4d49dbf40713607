// TEST INFRASTRUCTURE ONLY -- scalar 2-D inverse transform (K1) for the CPU oracle.
// Restates AV1 spec 7.13.3 (2-D inverse transform process): row pass with rectangular 1/sqrt2
// pre-scale and bd+8 clamp, Round2(rowShift), clamp to max(bd+6,16) bits, column pass, Round2(4).
// The 1-D butterflies come from av1-go_b200/csrc/kernels/itx1d.h (shared with the CUDA kernel) and
// are unit-pinned against libaom 3.13.1's av1_idct*/av1_iadst*/av1_inv_txfm2d_add_*_c.
#include <stdint.h>
#include <string.h>

#include "../av1-go_b200/csrc/av1_consts.h"
#include "../av1-go_b200/csrc/kernels/itx1d.h"

using namespace av1r;

static inline int32_t clampi(int32_t v, int bits) {
    const int32_t mx = (1 << (bits - 1)) - 1, mn = -(1 << (bits - 1));
    return v < mn ? mn : (v > mx ? mx : v);
}
static inline int32_t round2(int32_t x, int n) { return n == 0 ? x : (x + (1 << (n - 1))) >> n; }

static int row_shift(int txsz) {
    static const int8_t s[TX_SIZES_ALL] = {0, 1, 2, 2, 2, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2};
    return s[txsz];
}

// coef: dequantised coefficients, row-major, stride = min(w,32), rows = min(h,32).
// res: output residual, row-major w x h.
extern "C" void orc_inverse_transform_2d(const int32_t* coef, int txsz, int txtp, int bd, int32_t* res) {
    const int w = kTxW[txsz], h = kTxH[txsz];
    const int cw = w < 32 ? w : 32, chh = h < 32 ? h : 32;
    static thread_local int32_t buf[64 * 64];
    int32_t t[64];
    if (txtp == WHT_WHT) {
        for (int i = 0; i < 4; i++) {
            for (int j = 0; j < 4; j++) t[j] = coef[i * 4 + j];
            iwht4(t, 2);
            for (int j = 0; j < 4; j++) buf[i * 4 + j] = t[j];
        }
        for (int j = 0; j < 4; j++) {
            for (int i = 0; i < 4; i++) t[i] = buf[i * 4 + j];
            iwht4(t, 0);
            for (int i = 0; i < 4; i++) res[i * 4 + j] = t[i];
        }
        return;
    }
    int vk, hk, ud, lr;
    txtp_decompose(txtp, vk, hk, ud, lr);
    const int rect = (kTxWLog2[txsz] - kTxHLog2[txsz] == 1) || (kTxHLog2[txsz] - kTxWLog2[txsz] == 1);
    const int rs = row_shift(txsz);
    const int mid_bits = bd + 6 > 16 ? bd + 6 : 16;
    memset(buf, 0, sizeof(int32_t) * w * h);
    for (int i = 0; i < chh; i++) {
        for (int j = 0; j < w; j++) {
            int32_t v = j < cw ? coef[i * cw + j] : 0;
            if (rect) v = round2(v * 2896, 12);
            t[j] = clampi(v, bd + 8);
        }
        itx_1d(t, w, hk);
        for (int j = 0; j < w; j++) buf[i * w + j] = round2(t[j], rs);
    }
    for (int j = 0; j < w; j++) {
        const int sj = lr ? w - 1 - j : j;
        for (int i = 0; i < h; i++) t[i] = clampi(buf[i * w + sj], mid_bits);
        itx_1d(t, h, vk);
        for (int i = 0; i < h; i++) res[(ud ? h - 1 - i : i) * w + j] = round2(t[i], 4);
    }
}

extern "C" void orc_itx_1d(int32_t* t, int n, int kind) { itx_1d(t, n, kind); }
