// K4 -- deblocking loop filter (AV1 spec 7.14) for sm_100a.
//
// The host parse already classified every 4x4 unit edge (filter size 4/8/16 and level, LfEdge), so
// the device work is pure pixel filtering.  Within one pass the read/modify sets of different edges
// never overlap (a filter of size L touches < L/2 samples per side and L <= the transform size on
// both sides), so every edge sample line of a pass is an independent thread; only the vertical ->
// horizontal pass order is a dependency (two launches per frame, all planes in each).
// In place; algorithmic bytes 2F (the horizontal pass re-reads lines that are still L2 resident).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../av1_consts.h"
#include "dev_common.cuh"
#include "devframe.h"
#include "intra.h"

namespace av1r {

template <typename T>
__device__ __forceinline__ void lf_line(T* q0p, int step, int plane, int fsz, int lvl, int sharp, int bd) {
    const int shift = sharp > 4 ? 2 : (sharp > 0 ? 1 : 0);
    const int limit = sharp > 0 ? min(max(lvl >> shift, 1), 9 - sharp) : max(1, lvl >> shift);
    const int blimit = 2 * (lvl + 2) + limit;
    const int thresh = lvl >> 4;
    const int s = bd - 8;
    const int limit_bd = limit << s, blimit_bd = blimit << s, thresh_bd = thresh << s, one = 1 << s;
    const int flen = fsz == 4 ? 4 : (plane != 0 ? 6 : (fsz == 8 ? 8 : 16));
    const int nread = flen == 4 ? 2 : (flen == 6 ? 3 : (flen == 8 ? 4 : 7));
    int q[7], p[7];
#pragma unroll
    for (int i = 0; i < 7; i++) {
        if (i < nread) {
            q[i] = q0p[i * step];
            p[i] = q0p[-(i + 1) * step];
        } else {
            q[i] = p[i] = 0;
        }
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
#pragma unroll
        for (int i = 4; i < 7; i++) flat2 = flat2 && abs(p[i] - p[0]) <= one && abs(q[i] - q[0]) <= one;
    }
    if (fsz == 4 || !flat) {
        const int lo = -(1 << (bd - 1)), hi = (1 << (bd - 1)) - 1, off = 0x80 << s;
        const int ps1 = p[1] - off, ps0 = p[0] - off, qs0 = q[0] - off, qs1 = q[1] - off;
        int f = hev ? d_clip3(lo, hi, ps1 - qs1) : 0;
        f = d_clip3(lo, hi, f + 3 * (qs0 - ps0));
        const int f1 = d_clip3(lo, hi, f + 4) >> 3, f2 = d_clip3(lo, hi, f + 3) >> 3;
        q0p[0] = (T)(d_clip3(lo, hi, qs0 - f1) + off);
        q0p[-step] = (T)(d_clip3(lo, hi, ps0 + f2) + off);
        if (!hev) {
            const int f3 = (f1 + 1) >> 1;
            q0p[step] = (T)(d_clip3(lo, hi, qs1 - f3) + off);
            q0p[-2 * step] = (T)(d_clip3(lo, hi, ps1 + f3) + off);
        }
        return;
    }
    if (fsz == 8 || !flat2) {
        if (plane == 0) {
            // 8-tap: n = 3, centre tap doubled, >> 3
            const int v[8] = {p[3], p[2], p[1], p[0], q[0], q[1], q[2], q[3]};   // positions -4 .. 3
            int F[6];
#pragma unroll
            for (int i = -3; i < 3; i++) {
                int t = 0;
#pragma unroll
                for (int j = -3; j <= 3; j++) {
                    const int pp = min(max(i + j, -4), 3);
                    t += v[pp + 4] * (j == 0 ? 2 : 1);
                }
                F[i + 3] = (t + 4) >> 3;
            }
#pragma unroll
            for (int i = -3; i < 3; i++) q0p[i * step] = (T)F[i + 3];
        } else {
            // chroma 6-tap: n = 2, taps |j| <= 1 doubled, >> 3
            const int v[6] = {p[2], p[1], p[0], q[0], q[1], q[2]};               // positions -3 .. 2
            int F[4];
#pragma unroll
            for (int i = -2; i < 2; i++) {
                int t = 0;
#pragma unroll
                for (int j = -2; j <= 2; j++) {
                    const int pp = min(max(i + j, -3), 2);
                    t += v[pp + 3] * (abs(j) <= 1 ? 2 : 1);
                }
                F[i + 2] = (t + 4) >> 3;
            }
#pragma unroll
            for (int i = -2; i < 2; i++) q0p[i * step] = (T)F[i + 2];
        }
        return;
    }
    {
        // 14-tap: n = 6, taps |j| <= 1 doubled, >> 4
        const int v[14] = {p[6], p[5], p[4], p[3], p[2], p[1], p[0], q[0], q[1], q[2], q[3], q[4], q[5], q[6]};   // -7 .. 6
        int F[12];
#pragma unroll
        for (int i = -6; i < 6; i++) {
            int t = 0;
#pragma unroll
            for (int j = -6; j <= 6; j++) {
                const int pp = min(max(i + j, -7), 6);
                t += v[pp + 7] * (abs(j) <= 1 ? 2 : 1);
            }
            F[i + 6] = (t + 8) >> 4;
        }
#pragma unroll
        for (int i = -6; i < 6; i++) q0p[i * step] = (T)F[i + 6];
    }
}

// one thread per (4x4 unit, line 0..3); blockIdx.z = plane
template <typename T, int PASS>
__global__ void __launch_bounds__(256) deblock_kernel(LfLaunch L) {
    const int plane = blockIdx.z;
    if (!L.plane_on[plane]) return;
    const int pw4 = L.fp.pw4[plane], ph4 = L.fp.ph4[plane];
    int c4, r4, line;
    if (PASS == 0) {
        // vertical edges: consecutive threads walk down the rows of one unit column -> pixel row = thread
        const int col = blockIdx.x * 8 + (threadIdx.x & 7);        // unit column
        const int prow = blockIdx.y * 32 + (threadIdx.x >> 3);     // pixel row
        c4 = col;
        r4 = prow >> 2;
        line = prow & 3;
    } else {
        // horizontal edges: consecutive threads = consecutive pixel columns (coalesced rows)
        const int pcol = blockIdx.x * 64 + (threadIdx.x & 63);
        r4 = blockIdx.y * 4 + (threadIdx.x >> 6);
        c4 = pcol >> 2;
        line = pcol & 3;
    }
    if (c4 >= pw4 || r4 >= ph4) return;
    const LfEdge e = L.edges[plane][(size_t)r4 * pw4 + c4];
    const int len = PASS ? e.len_h : e.len_v, lvl = PASS ? e.lvl_h : e.lvl_v;
    if (!len) return;
    const int x = c4 * 4 + (PASS ? line : 0), y = r4 * 4 + (PASS ? 0 : line);
    if (x >= L.fp.cw[plane] || y >= L.fp.ch[plane]) return;
    const int pitch_e = L.frame.pitch[plane] / sizeof(T);
    T* q0 = (T*)L.frame.p[plane] + (size_t)y * pitch_e + x;
    lf_line<T>(q0, PASS ? pitch_e : 1, plane, len, lvl, L.fp.lf_sharpness, L.fp.bd);
}

// ---- edge classification on the device (what the host pre-pass build_loopfilter_edges computes, stream_parser.cpp) -----------
// Width / height of a transform size and of a block size as shifts of packed immediates (log2 - 2, three bits per entry): these
// are looked up per thread with a different index in every lane, which a constant-memory table serves one address at a time.
__device__ __forceinline__ int lf_txw(int t) { return 4 << (int)((0x1132846d2244688ull >> (3 * t)) & 7); }
__device__ __forceinline__ int lf_txh(int t) { return 4 << (int)((0xa161389940c688ull >> (3 * t)) & 7); }
__device__ __forceinline__ int lf_bw(int b) { return 4 << (b < 21 ? (int)((0x2650b648db491240ull >> (3 * b)) & 7) : 4); }   // BLOCK_64X16 is entry 21
__device__ __forceinline__ int lf_bh(int b) { return 4 << (b < 21 ? (int)((0x42c2b2c71a68a208ull >> (3 * b)) & 7) : 2); }

// one warp per block: its mi cells (clipped to the frame) get the block's index
__global__ void __launch_bounds__(256) lf_scatter_kernel(const LfBlk* __restrict__ blks, int n, uint32_t* __restrict__ mi_blk, int mi_cols, int mi_rows) {
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= n) return;
    const LfBlk b = blks[wid];
    const int w4 = lf_bw(b.bsize) >> 2, h4 = lf_bh(b.bsize) >> 2;
    const int cw = min(w4, mi_cols - b.mi_col), ch = min(h4, mi_rows - b.mi_row);
    const int lw = 31 - __clz(w4);
    for (int i = lane; i < (h4 << lw); i += 32) {
        const int y = i >> lw, x = i & (w4 - 1);
        if (x < cw && y < ch) mi_blk[(size_t)(b.mi_row + y) * mi_cols + b.mi_col + x] = (uint32_t)wid;
    }
}

__global__ void __launch_bounds__(256) lf_classify_kernel(LfClassify L) {
    const int plane = blockIdx.z;
    if (!L.plane_on[plane]) return;
    const DevFrameParams& fp = L.fp;
    const int sx = plane ? fp.subx : 0, sy = plane ? fp.suby : 0;
    const int pw4 = fp.pw4[plane], ph4 = fp.ph4[plane];
    const int c4 = blockIdx.x * 64 + (threadIdx.x & 63), r4 = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (c4 >= pw4 || r4 >= ph4) return;
    const int mi_cols = fp.mi_cols, mi_rows = fp.mi_rows;
    const size_t idx = (size_t)r4 * pw4 + c4;
    const int row = r4 << sy, col = c4 << sx;
    LfEdge e = {0, 0, 0, 0};
    // for sub-sampled planes the spec addresses mode info at the odd (bottom-right) luma mi
    const int mrow = min(mi_rows - 1, row | sy), mcol = min(mi_cols - 1, col | sx);
    if (row * 4 < fp.h[0] && col * 4 < fp.w[0]) {
        const uint8_t* lf_tx = L.lf_tx[plane];
        const LfBlk b = L.blks[L.mi_blk[(size_t)mrow * mi_cols + mcol]];
        const int max_len = plane ? 8 : 16;
        const int li0 = plane == 0 ? 0 : plane + 1, li1 = plane == 0 ? 1 : plane + 1;
        const int txsz = lf_tx[idx];
        const int txw = lf_txw(txsz), txh = lf_txh(txsz);
        const int bwp = max(4, lf_bw(b.bsize) >> sx), bhp = max(4, lf_bh(b.bsize) >> sy);
        const int xp = c4 * 4, yp = r4 * 4;
        if (c4 > 0 && (xp & (txw - 1)) == 0 && (b.filt_inside || (xp & (bwp - 1)) == 0)) {
            int lvl = b.lvl[li0];
            if (!lvl) {
                const int pcol = min(mi_cols - 1, ((c4 - 1) << sx) | sx);
                lvl = L.blks[L.mi_blk[(size_t)mrow * mi_cols + pcol]].lvl[li0];
            }
            if (lvl) {
                e.len_v = (uint8_t)min(max_len, min(lf_txw(lf_tx[idx - 1]), txw));
                e.lvl_v = (uint8_t)lvl;
            }
        }
        if (r4 > 0 && (yp & (txh - 1)) == 0 && (b.filt_inside || (yp & (bhp - 1)) == 0)) {
            int lvl = b.lvl[li1];
            if (!lvl) {
                const int prow = min(mi_rows - 1, ((r4 - 1) << sy) | sy);
                lvl = L.blks[L.mi_blk[(size_t)prow * mi_cols + mcol]].lvl[li1];
            }
            if (lvl) {
                e.len_h = (uint8_t)min(max_len, min(lf_txh(lf_tx[idx - pw4]), txh));
                e.lvl_h = (uint8_t)lvl;
            }
        }
    }
    L.edges[plane][idx] = e;
}

cudaError_t launch_lf_classify(const LfClassify& L, cudaStream_t s) {
    if (L.n_blks <= 0) return cudaErrorInvalidValue;
    lf_scatter_kernel<<<(L.n_blks * 32 + 255) / 256, 256, 0, s>>>(L.blks, L.n_blks, L.mi_blk, L.fp.mi_cols, L.fp.mi_rows);
    dim3 grid((L.fp.pw4[0] + 63) / 64, (L.fp.ph4[0] + 3) / 4, L.fp.mono ? 1 : 3);
    lf_classify_kernel<<<grid, 256, 0, s>>>(L);
    return cudaGetLastError();
}

cudaError_t launch_deblock(const LfLaunch& L, cudaStream_t s) {
    {
        static bool carve_done = false;
        if (!carve_done) {
            prefer_max_smem(deblock_kernel<uint8_t, 0>);
            prefer_max_smem(deblock_kernel<uint16_t, 0>);
            prefer_max_smem(deblock_kernel<uint8_t, 1>);
            prefer_max_smem(deblock_kernel<uint16_t, 1>);
            carve_done = true;
        }
    }
    int nplanes = L.fp.mono ? 1 : 3;
    {
        dim3 grid((L.fp.pw4[0] + 7) / 8, (L.fp.ch[0] + 31) / 32, nplanes);
        if (L.fp.bd == 8) deblock_kernel<uint8_t, 0><<<grid, 256, 0, s>>>(L);
        else deblock_kernel<uint16_t, 0><<<grid, 256, 0, s>>>(L);
    }
    {
        dim3 grid((L.fp.cw[0] + 63) / 64, (L.fp.ph4[0] + 3) / 4, nplanes);
        if (L.fp.bd == 8) deblock_kernel<uint8_t, 1><<<grid, 256, 0, s>>>(L);
        else deblock_kernel<uint16_t, 1><<<grid, 256, 0, s>>>(L);
    }
    return cudaGetLastError();
}

}  // namespace av1r
