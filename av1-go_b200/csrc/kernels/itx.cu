// K1 -- dequantisation + 2-D inverse transform (AV1 spec 7.12.3, 7.13.3) for sm_100a.
//
// One warp per transform block.  The sparse coefficient tokens of the block (the only HBM read
// stream: 4 bytes per non-zero coefficient) are dequantised and scattered into a padded
// shared-memory tile, lanes then run one row transform each (butterflies in registers, itx1d.h),
// and after a warp barrier one column transform each, writing the int16 residual with coalesced
// row-contiguous stores.  All transform blocks of a frame are independent -> one launch per frame.
// Algorithmic bytes: C (tokens) + 32 B/record + 2A (residual write).  No tensor-core work exists here:
// the butterflies need the normative Round2 after every rotation, which a GEMM cannot reproduce.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../av1_consts.h"
#include "dev_common.cuh"
#include "devframe.h"
#include "itx1d.h"
#include "../tables/tables_quant.inc"
#include "../tables/tables_qm.inc"

namespace av1r {

__constant__ int16_t c_dc_q[3][256];
__constant__ int16_t c_ac_q[3][256];
__constant__ uint8_t c_txw_log2[TX_SIZES_ALL];
__constant__ uint8_t c_txh_log2[TX_SIZES_ALL];
__constant__ uint16_t c_qm_offset[TX_SIZES_ALL];
__device__ uint8_t d_iqmatrix[15][2][3344];   // Quantizer_Matrix (spec 7.12.3), row-major per transform size; 100 KB, read through L1/L2
static bool g_itx_const_loaded[64] = {false};

static constexpr int ITX_WARPS = 4;
static constexpr int ITX_BUF_WORDS = 32 * 65;   // worst case: 32 coded rows x (64 + 1) columns

template <int N>
__device__ __forceinline__ void run_1d(int32_t* t, int kind) {
    if (kind == ITX_DCT) {
        if (N == 4) idct4(t);
        else if (N == 8) idct8(t);
        else if (N == 16) idct16(t);
        else if (N == 32) idct32(t);
        else idct64(t);
    } else if (kind == ITX_IDENTITY) {
        if (N == 4) iidentity4(t);
        else if (N == 8) iidentity8(t);
        else if (N == 16) iidentity16(t);
        else iidentity32(t);
    } else {
        if (N == 4) iadst4(t);
        else if (N == 8) iadst8(t);
        else iadst16(t);
    }
}

__device__ __forceinline__ int32_t clamp_bits(int32_t v, int bits) {
    const int32_t mx = (1 << (bits - 1)) - 1, mn = -(1 << (bits - 1));
    return min(max(v, mn), mx);
}

// row pass for one row of width W held in buf (stride W+1); only columns < cw are non-zero on input
template <int W>
__device__ __forceinline__ void row_pass(int32_t* row, int cw, int kind, int rect, int bd, int rs) {
    int32_t t[W];
#pragma unroll
    for (int j = 0; j < W; j++) {
        int32_t v = (j < cw) ? row[j] : 0;
        if (rect) v = (v * 2896 + 2048) >> 12;
        t[j] = clamp_bits(v, bd + 8);
    }
    run_1d<W>(t, kind);
#pragma unroll
    for (int j = 0; j < W; j++) row[j] = rs ? ((t[j] + (1 << (rs - 1))) >> rs) : t[j];
}

// column pass for one column: rows < ch come from buf (stride), the rest are zero
template <int H>
__device__ __forceinline__ void col_pass(const int32_t* col, int stride, int ch, int kind, int mid_bits, int ud_flip, int16_t* dst,
                                         int dst_pitch_elems, int rows_valid) {
    int32_t t[H];
#pragma unroll
    for (int i = 0; i < H; i++) t[i] = (i < ch) ? clamp_bits(col[i * stride], mid_bits) : 0;
    run_1d<H>(t, kind);
#pragma unroll
    for (int i = 0; i < H; i++) {
        const int oi = ud_flip ? H - 1 - i : i;
        if (oi < rows_valid) dst[oi * dst_pitch_elems] = (int16_t)((t[i] + 8) >> 4);
    }
}

__global__ void __launch_bounds__(ITX_WARPS * 32) itx_kernel(const TxRec* __restrict__ recs, const uint32_t* __restrict__ order, int n,
                                                            const uint32_t* __restrict__ coefs, DevResidual res, DevFrameParams fp) {
    __shared__ int32_t s_buf[ITX_WARPS][ITX_BUF_WORDS];
    const int warp_in = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bi = blockIdx.x * ITX_WARPS + warp_in;
    if (bi >= n) return;
    const TxRec r = recs[order[bi]];
    int32_t* buf = s_buf[warp_in];
    const int txsz = r.txsz, plane = r.plane;
    const int lw = c_txw_log2[txsz], lh = c_txh_log2[txsz];
    const int w = 1 << lw, h = 1 << lh;
    const int cw = min(w, 32), ch = min(h, 32);
    const int lcw = min(lw, 5);
    const int stride = w + 1;
    const int bd = fp.bd;
    // ---- zero + dequantise + scatter
    for (int i = lane; i < ch * stride; i += 32) buf[i] = 0;
    __syncwarp();
    {
        const int bdi = (bd - 8) >> 1;
        const int dcq = c_dc_q[bdi][min(max(r.qidx + fp.dq_dc[plane], 0), 255)];
        const int acq = c_ac_q[bdi][min(max(r.qidx + fp.dq_ac[plane], 0), 255)];
        const int pels = w * h;
        const int dq_denom = (pels > 256) + (pels > 1024);
        const int mx = (1 << (7 + bd)) - 1, mn = -(1 << (7 + bd));
        const uint32_t* tk = coefs + r.coef_off;
        // quantiser matrix: 2-D transform types only, level 15 = flat
        const uint8_t* qm = (r.qm_level < 15 && r.txtp < IDTX) ? d_iqmatrix[r.qm_level][plane > 0] + c_qm_offset[txsz] : nullptr;
        for (int k = lane; k < r.ntok; k += 32) {
            const uint32_t t = tk[k];
            const int pos = (int)(t & 1023), level = (int32_t)t >> 10;
            int q = pos == 0 ? dcq : acq;
            if (qm) q = (q * (int)__ldg(qm + pos) + 16) >> 5;
            uint32_t dq = ((uint32_t)abs(level) * (uint32_t)q) & 0xFFFFFFu;
            dq >>= dq_denom;
            int v = level < 0 ? -(int)dq : (int)dq;
            v = min(max(v, mn), mx);
            buf[(pos >> lcw) * stride + (pos & (cw - 1))] = v;
        }
    }
    __syncwarp();
    int16_t* dst = res_ptr(res, plane, r.x4 * 4, r.y4 * 4);   // a transform block never straddles a unit
    const int dpe = 1 << res.tw_log2[plane];
    // luma blocks are written in full also where they straddle the coded frame edge: chroma-from-luma averages the reconstructed
    // luma of the whole transform block (spec MaxLumaW / MaxLumaH), and the residual tile of the unit has room for it
    const int cols_valid = plane ? min(w, fp.cw[plane] - r.x4 * 4) : w, rows_valid = plane ? min(h, fp.ch[plane] - r.y4 * 4) : h;
    if (r.txtp == WHT_WHT) {
        if (lane < 4) {
            int32_t t[4];
            for (int j = 0; j < 4; j++) t[j] = buf[lane * stride + j];
            iwht4(t, 2);
            for (int j = 0; j < 4; j++) buf[lane * stride + j] = t[j];
        }
        __syncwarp();
        if (lane < 4) {
            int32_t t[4];
            for (int i = 0; i < 4; i++) t[i] = buf[i * stride + lane];
            iwht4(t, 0);
            if (lane < cols_valid)
                for (int i = 0; i < 4; i++)
                    if (i < rows_valid) dst[i * dpe + lane] = (int16_t)t[i];
        }
        return;
    }
    int vk, hk, ud, lr;
    txtp_decompose(r.txtp, vk, hk, ud, lr);
    const int rect = (lw - lh == 1) || (lh - lw == 1);
    static const int8_t kRowShift[TX_SIZES_ALL] = {0, 1, 2, 2, 2, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2};
    const int rs = kRowShift[txsz];
    // ---- row pass: lane i owns coded row i
    if (lane < ch) {
        int32_t* row = buf + lane * stride;
        switch (lw) {
            case 2: row_pass<4>(row, cw, hk, rect, bd, rs); break;
            case 3: row_pass<8>(row, cw, hk, rect, bd, rs); break;
            case 4: row_pass<16>(row, cw, hk, rect, bd, rs); break;
            case 5: row_pass<32>(row, cw, hk, rect, bd, rs); break;
            default: row_pass<64>(row, cw, hk, rect, bd, rs); break;
        }
    }
    __syncwarp();
    // ---- column pass: lane j owns output column j (two rounds when w == 64)
    const int mid_bits = max(bd + 6, 16);
    for (int j = lane; j < w; j += 32) {
        if (j >= cols_valid) continue;
        const int sj = lr ? w - 1 - j : j;
        const int32_t* col = buf + sj;
        switch (lh) {
            case 2: col_pass<4>(col, stride, ch, vk, mid_bits, ud, dst + j, dpe, rows_valid); break;
            case 3: col_pass<8>(col, stride, ch, vk, mid_bits, ud, dst + j, dpe, rows_valid); break;
            case 4: col_pass<16>(col, stride, ch, vk, mid_bits, ud, dst + j, dpe, rows_valid); break;
            case 5: col_pass<32>(col, stride, ch, vk, mid_bits, ud, dst + j, dpe, rows_valid); break;
            default: col_pass<64>(col, stride, ch, vk, mid_bits, ud, dst + j, dpe, rows_valid); break;
        }
    }
}

// Transform blocks of at most 16x16 (the bulk of every stream) -- half a warp per block, t[16] registers: ~1/4 of the registers
// of the general kernel, so that K1 no longer evicts the persistent K3 CTAs of the frames in flight on other streams.
static constexpr int ITXS_HALF_WARPS = 8;
__global__ void __launch_bounds__(ITXS_HALF_WARPS * 16) itx_small_kernel(const TxRec* __restrict__ recs, const uint32_t* __restrict__ order, int n,
                                                                        const uint32_t* __restrict__ coefs, DevResidual res, DevFrameParams fp) {
    __shared__ int32_t s_buf[ITXS_HALF_WARPS][16 * 17];
    const int hw = threadIdx.x >> 4, lane = threadIdx.x & 15;
    const unsigned mask = 0xFFFFu << (threadIdx.x & 16);   // the two halves of a warp work on different blocks
    const int bi = blockIdx.x * ITXS_HALF_WARPS + hw;
    if (bi >= n) return;
    const TxRec r = recs[order[bi]];
    int32_t* buf = s_buf[hw];
    const int txsz = r.txsz, plane = r.plane;
    const int lw = c_txw_log2[txsz], lh = c_txh_log2[txsz];
    const int w = 1 << lw, h = 1 << lh;
    const int stride = w + 1;
    const int bd = fp.bd;
    for (int i = lane; i < h * stride; i += 16) buf[i] = 0;
    __syncwarp(mask);
    {
        const int bdi = (bd - 8) >> 1;
        const int dcq = c_dc_q[bdi][min(max(r.qidx + fp.dq_dc[plane], 0), 255)];
        const int acq = c_ac_q[bdi][min(max(r.qidx + fp.dq_ac[plane], 0), 255)];
        const int mx = (1 << (7 + bd)) - 1, mn = -(1 << (7 + bd));
        const uint32_t* tk = coefs + r.coef_off;
        const uint8_t* qm = (r.qm_level < 15 && r.txtp < IDTX) ? d_iqmatrix[r.qm_level][plane > 0] + c_qm_offset[txsz] : nullptr;
        for (int k = lane; k < r.ntok; k += 16) {
            const uint32_t t = tk[k];
            const int pos = (int)(t & 1023), level = (int32_t)t >> 10;
            int q = pos == 0 ? dcq : acq;
            if (qm) q = (q * (int)__ldg(qm + pos) + 16) >> 5;
            const uint32_t dq = ((uint32_t)abs(level) * (uint32_t)q) & 0xFFFFFFu;   // dqDenom is 1 up to 256 samples
            int v = level < 0 ? -(int)dq : (int)dq;
            v = min(max(v, mn), mx);
            buf[(pos >> lw) * stride + (pos & (w - 1))] = v;
        }
    }
    __syncwarp(mask);
    int16_t* dst = res_ptr(res, plane, r.x4 * 4, r.y4 * 4);
    const int dpe = 1 << res.tw_log2[plane];
    // luma blocks are written in full also where they straddle the coded frame edge: chroma-from-luma averages the reconstructed
    // luma of the whole transform block (spec MaxLumaW / MaxLumaH), and the residual tile of the unit has room for it
    const int cols_valid = plane ? min(w, fp.cw[plane] - r.x4 * 4) : w, rows_valid = plane ? min(h, fp.ch[plane] - r.y4 * 4) : h;
    if (r.txtp == WHT_WHT) {
        if (lane < 4) {
            int32_t t[4];
            for (int j = 0; j < 4; j++) t[j] = buf[lane * stride + j];
            iwht4(t, 2);
            for (int j = 0; j < 4; j++) buf[lane * stride + j] = t[j];
        }
        __syncwarp(mask);
        if (lane < 4) {
            int32_t t[4];
            for (int i = 0; i < 4; i++) t[i] = buf[i * stride + lane];
            iwht4(t, 0);
            if (lane < cols_valid)
                for (int i = 0; i < 4; i++)
                    if (i < rows_valid) dst[i * dpe + lane] = (int16_t)t[i];
        }
        return;
    }
    int vk, hk, ud, lr;
    txtp_decompose(r.txtp, vk, hk, ud, lr);
    const int rect = (lw - lh == 1) || (lh - lw == 1);
    static const int8_t kRowShiftS[TX_SIZES_ALL] = {0, 1, 2, 2, 2, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2};
    const int rs = kRowShiftS[txsz];
    if (lane < h) {
        int32_t* row = buf + lane * stride;
        switch (lw) {
            case 2: row_pass<4>(row, w, hk, rect, bd, rs); break;
            case 3: row_pass<8>(row, w, hk, rect, bd, rs); break;
            default: row_pass<16>(row, w, hk, rect, bd, rs); break;
        }
    }
    __syncwarp(mask);
    const int mid_bits = max(bd + 6, 16);
    if (lane < cols_valid) {
        const int sj = lr ? w - 1 - lane : lane;
        const int32_t* col = buf + sj;
        switch (lh) {
            case 2: col_pass<4>(col, stride, h, vk, mid_bits, ud, dst + lane, dpe, rows_valid); break;
            case 3: col_pass<8>(col, stride, h, vk, mid_bits, ud, dst + lane, dpe, rows_valid); break;
            default: col_pass<16>(col, stride, h, vk, mid_bits, ud, dst + lane, dpe, rows_valid); break;
        }
    }
}

cudaError_t itx_upload_constants() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && g_itx_const_loaded[dev]) return cudaSuccess;
    e = cudaMemcpyToSymbol(c_dc_q, av1t_dc_qlookup, sizeof(av1t_dc_qlookup));
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_ac_q, av1t_ac_qlookup, sizeof(av1t_ac_qlookup));
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_txw_log2, kTxWLog2, sizeof(kTxWLog2));
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_txh_log2, kTxHLog2, sizeof(kTxHLog2));
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_qm_offset, av1t_qm_offset, sizeof(av1t_qm_offset));
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(d_iqmatrix, av1t_iqmatrix, sizeof(av1t_iqmatrix));
    if (e != cudaSuccess) return e;
    if (dev < 64) g_itx_const_loaded[dev] = true;
    return cudaSuccess;
}

// order: indices of the records with eob > 0 (device array of n entries); the first n_small of them are at most 16x16
cudaError_t launch_itx(const TxRec* recs, const uint32_t* order, int n, int n_small, const uint32_t* coefs, const DevResidual& res,
                       const DevFrameParams& fp, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    cudaError_t e = itx_upload_constants();
    if (e != cudaSuccess) return e;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(itx_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr_done = true;
    }
    {
        static bool carve_done = false;
        if (!carve_done) {
            prefer_max_smem(itx_small_kernel);
            carve_done = true;
        }
    }
    if (n_small > 0)
        itx_small_kernel<<<(n_small + ITXS_HALF_WARPS - 1) / ITXS_HALF_WARPS, ITXS_HALF_WARPS * 16, 0, s>>>(recs, order, n_small, coefs, res, fp);
    if (n > n_small) {
        const int nl = n - n_small;
        itx_kernel<<<(nl + ITX_WARPS - 1) / ITX_WARPS, ITX_WARPS * 32, 0, s>>>(recs, order + n_small, nl, coefs, res, fp);
    }
    return cudaGetLastError();
}

}  // namespace av1r
