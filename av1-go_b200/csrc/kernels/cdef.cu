// K5 -- constrained directional enhancement filter (AV1 spec 7.15) for sm_100a.
//
// One CTA per 64x64 luma filter block (plus its co-located chroma).  The deblocked input tile and a
// 2-sample halo are staged in shared memory once (samples outside the coded frame are marked
// unavailable), the 8x8 direction search runs from shared memory, then every thread filters
// 16 luma + 8 chroma samples.  Out of place (CDEF reads pre-CDEF neighbours); samples of skipped
// 8x8 blocks are copied through.  Algorithmic bytes 2F; the halo re-reads are L2 hits.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "dev_common.cuh"
#include "devframe.h"
#include "intra.h"

namespace av1r {

__constant__ int8_t c_cdef_dir[8][2][2] = {{{-1, 1}, {-2, 2}}, {{0, 1}, {-1, 2}}, {{0, 1}, {0, 2}}, {{0, 1}, {1, 2}},
                                           {{1, 1}, {2, 2}},   {{1, 0}, {2, 1}},  {{1, 0}, {2, 0}}, {{1, 0}, {2, -1}}};
__constant__ uint8_t c_cdef_uv_dir[2][2][8] = {{{0, 1, 2, 3, 4, 5, 6, 7}, {1, 2, 2, 2, 3, 4, 6, 0}}, {{7, 0, 2, 4, 5, 6, 6, 6}, {0, 1, 2, 3, 4, 5, 6, 7}}};

static constexpr int CDEF_LT = 68;   // luma tile edge (64 + 2*2)
// Row stride of the staged tiles.  Tile column tx (sample x0 - 2 + tx) lives at row offset tx + 2, so that offset 0 is sample
// x0 - 4: an 8-byte aligned address in the frame, and the rows of interior tiles arrive as 64-bit loads.
static constexpr int CDEF_LS = 72;
#define CDEF_TIDX(ty, tx) ((ty) * CDEF_LS + (tx) + 2)

// Direction search (spec 7.15.2) as 90 line sums per 8x8 block: direction d partitions the block into lines (15 diagonals for d = 0, 4;
// 8 rows / columns for d = 2, 6; 11 half-slope lines for the odd directions); cost[d] = sum over its lines of (line sum)^2 * 840 / (samples
// on the line).  One warp searches one block: every lane owns three lines.  The table lists, per line, its samples (i * 8 + j, 0xff =
// none), its direction and its weight; it is built once on the host.
static constexpr int CDEF_LINES = 96;   // 90 used, padded to 3 per lane
struct CdefLineTable {
    uint8_t pix[CDEF_LINES][8];
    uint16_t weight[CDEF_LINES];
    uint8_t dir[CDEF_LINES];
};
__device__ CdefLineTable g_cdef_lines;
static bool g_cdef_lines_loaded[64] = {false};

static void cdef_build_lines(CdefLineTable& t) {
    memset(&t, 0xff, sizeof(t));
    int n = 0;
    for (int d = 0; d < 8; d++) {
        const int nl = (d == 0 || d == 4) ? 15 : ((d == 2 || d == 6) ? 8 : 11);
        for (int k = 0; k < nl; k++) {
            int cnt = 0;
            for (int i = 0; i < 8; i++)
                for (int j = 0; j < 8; j++) {
                    int f;
                    switch (d) {
                        case 0: f = i + j; break;
                        case 1: f = i + j / 2; break;
                        case 2: f = i; break;
                        case 3: f = 3 + i - j / 2; break;
                        case 4: f = 7 + i - j; break;
                        case 5: f = 3 - i / 2 + j; break;
                        case 6: f = j; break;
                        default: f = i / 2 + j; break;
                    }
                    if (f == k) t.pix[n][cnt++] = (uint8_t)(i * 8 + j);
                }
            t.weight[n] = (uint16_t)(840 / cnt);
            t.dir[n] = (uint8_t)d;
            n++;
        }
    }
    for (; n < CDEF_LINES; n++) {
        t.weight[n] = 0;
        t.dir[n] = 0;
    }
}

// constrain() of spec 7.15.3 with the damping shift (adj = max(0, damping - floor(log2(threshold)))) hoisted out of the tap loop
__device__ __forceinline__ int cdef_constrain(int diff, int threshold, int adj) {
    const int mag = abs(diff);
    const int v = min(max(threshold - (mag >> adj), 0), mag);
    return diff < 0 ? -v : v;
}

// One filtered sample (spec 7.15.3), branch free: a tap that is unavailable (marked -1 in the tile) or whose strength class is off is
// replaced by the centre sample, which makes its difference 0 and leaves the min / max unchanged -- exactly the effect of skipping it.
// offsets: op[k] primary, oa[k] / ob[k] the two secondary directions (tile offsets dy * stride + dx), k = tap distance 1 / 2.
struct CdefBlk {
    int pri, sec, adj_p, adj_s, pt0, pt1;
    int op[2], oa[2], ob[2];
};
__device__ __forceinline__ int cdef_tap(int p, int x, bool on, int thr, int adj, int wgt, int& mx, int& mn) {
    p = (on && p >= 0) ? p : x;
    const int d = p - x;
    const int v = cdef_constrain(d, thr, adj);
    mx = max(mx, p);
    mn = min(mn, p);
    return wgt * v;
}
__device__ __forceinline__ int cdef_pixel(const int16_t* tile, int pos, const CdefBlk& B) {
    const int x = tile[pos];
    int sum = 0, mx = x, mn = x;
    const bool pon = B.pri != 0, son = B.sec != 0;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int ptap = k ? B.pt1 : B.pt0, stap = k ? 1 : 2;
#pragma unroll
        for (int sg = -1; sg <= 1; sg += 2) {
            sum += cdef_tap(tile[pos + sg * B.op[k]], x, pon, B.pri, B.adj_p, ptap, mx, mn);
            sum += cdef_tap(tile[pos + sg * B.oa[k]], x, son, B.sec, B.adj_s, stap, mx, mn);
            sum += cdef_tap(tile[pos + sg * B.ob[k]], x, son, B.sec, B.adj_s, stap, mx, mn);
        }
    }
    return min(max(x + ((8 + sum - (sum < 0)) >> 4), mn), mx);
}

template <typename T>
// (six CTAs per SM: 40 registers with a few bytes of spill beat 60 registers at four CTAs by 5 % -- the kernel waits on its tile
// loads, profiles/r2_work_order.md)
__global__ void __launch_bounds__(256, 6) cdef_kernel(CdefLaunch L) {
    __shared__ __align__(16) int16_t s_luma[CDEF_LT * CDEF_LS];
    __shared__ __align__(16) int16_t s_chroma[2][CDEF_LT * CDEF_LS];   // sized for 4:4:4
    __shared__ uint8_t s_dir[64], s_skip[64];
    __shared__ int s_var[64];
    __shared__ __align__(16) CdefLineTable s_lines;
    __shared__ int16_t s_dtab[8][2];
    const DevFrameParams& fp = L.fp;
    const int tid = threadIdx.x;
    const int fbx = blockIdx.x, fby = blockIdx.y;
    const int c64 = (fp.mi_cols + 15) >> 4;
    const int idx = L.cdef_idx[(size_t)fby * c64 + fbx];
    const int bd = fp.bd, cs = bd - 8;
    const int nplanes = fp.mono ? 1 : 3;
    if (tid < 16) s_dtab[tid >> 1][tid & 1] = (int16_t)(c_cdef_dir[tid >> 1][tid & 1][0] * CDEF_LS + c_cdef_dir[tid >> 1][tid & 1][1]);
    for (int i = tid; i < (int)(sizeof(CdefLineTable) / 4); i += 256) reinterpret_cast<uint32_t*>(&s_lines)[i] = reinterpret_cast<const uint32_t*>(&g_cdef_lines)[i];
    // ---- stage tiles (a warp per row; samples outside the coded frame are marked -1 = unavailable)
    {
        const int warp = tid >> 5, lane = tid & 31;
        for (int plane = 0; plane < nplanes; plane++) {
            const int sx = plane ? fp.subx : 0, sy = plane ? fp.suby : 0;
            const int tw = (64 >> sx) + 4, th = (64 >> sy) + 4;
            const int x0 = (fbx * 64 >> sx) - 2, y0 = (fby * 64 >> sy) - 2;
            int16_t* tile = plane == 0 ? s_luma : s_chroma[plane - 1];
            const T* src = (const T*)L.src.p[plane];
            const int pitch_e = L.src.pitch[plane] / sizeof(T);
            // interior in x: all of x0 .. x0 + tw - 1 inside the coded width and the 64-bit pieces starting at x0 - 2 inside the row
            const int nq = (tw + 2 + 3) >> 2;                     // 64-bit pieces from sample x0 - 2: 18 (luma) / 10 (4:2:0 chroma)
            const bool fast = sizeof(T) == 2 && x0 >= 2 && x0 + tw <= fp.cw[plane] && (size_t)(x0 - 2 + 4 * nq) * sizeof(T) <= L.src.pitch[plane];
            for (int ty = warp; ty < th; ty += 8) {
                const int y = y0 + ty;
                const bool row_ok = y >= 0 && y < fp.ch[plane];
                if (fast && row_ok) {
                    if (lane < nq) {
                        const uint2 v = __ldg(reinterpret_cast<const uint2*>(src + (size_t)y * pitch_e + x0 - 2) + lane);
                        reinterpret_cast<uint2*>(tile + ty * CDEF_LS)[lane] = v;
                    }
                } else {
                    for (int tx = lane; tx < tw; tx += 32) {
                        const int x = x0 + tx;
                        int v = -1;
                        if (row_ok && x >= 0 && x < fp.cw[plane]) v = src[(size_t)y * pitch_e + x];
                        tile[CDEF_TIDX(ty, tx)] = (int16_t)v;
                    }
                }
            }
        }
    }
    // ---- per 8x8: skip flag
    if (tid < 64) {
        const int by = tid >> 3, bx = tid & 7;
        const int r = fby * 16 + by * 2, c = fbx * 16 + bx * 2;
        int skip = 1;
        if (idx >= 0 && r < fp.mi_rows && c < fp.mi_cols) {
            const uint8_t* s0 = L.skip_mi + (size_t)r * fp.mi_cols + c;
            const uint8_t* s1 = s0 + fp.mi_cols;
            skip = s0[0] && s0[1] && s1[0] && s1[1];
        }
        s_skip[tid] = (uint8_t)skip;
    }
    __syncthreads();
    // ---- direction search: one warp per 8x8 block, three line sums per lane (skipped when no primary strength needs a direction)
    const bool need_dir = idx >= 0 && (fp.cdef_y_pri[max(idx, 0)] | fp.cdef_uv_pri[max(idx, 0)]) != 0;
    if (need_dir) {
        const int warp = tid >> 5, lane = tid & 31;
        for (int blk = warp; blk < 64; blk += 8) {
            if (s_skip[blk]) continue;   // warp-uniform
            const int by = blk >> 3, bx = blk & 7;
            const int16_t* base = s_luma + CDEF_TIDX(by * 8 + 2, bx * 8 + 2);
            int c3[3], d3[3];
#pragma unroll
            for (int q = 0; q < 3; q++) {
                const int line = lane + 32 * q;
                const uint2 pk = *reinterpret_cast<const uint2*>(s_lines.pix[line]);
                int sum = 0;
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    const uint32_t px = ((t < 4 ? pk.x : pk.y) >> (8 * (t & 3))) & 0xff;
                    if (px != 0xff) sum += (base[(px >> 3) * CDEF_LS + (px & 7)] >> cs) - 128;
                }
                c3[q] = sum * sum * (int)s_lines.weight[line];
                d3[q] = s_lines.dir[line];
            }
            int cost[8];
#pragma unroll
            for (int d = 0; d < 8; d++)
                cost[d] = __reduce_add_sync(0xffffffffu, (d3[0] == d ? c3[0] : 0) + (d3[1] == d ? c3[1] : 0) + (d3[2] == d ? c3[2] : 0));
            if (lane == 0) {
                int best = 0, dir = 0;
#pragma unroll
                for (int d = 0; d < 8; d++)
                    if (cost[d] > best) { best = cost[d]; dir = d; }
                int opp = cost[0];
#pragma unroll
                for (int d = 1; d < 8; d++)
                    if (((dir + 4) & 7) == d) opp = cost[d];
                s_dir[blk] = (uint8_t)dir;
                s_var[blk] = (best - opp) >> 10;
            }
        }
    } else if (tid < 64) {
        s_dir[tid] = 0;
        s_var[tid] = 0;
    }
    __syncthreads();
    // ---- filter: a thread finishes one row of a block -- 8 luma / 4 (4:2:0) chroma samples that share strengths, direction and
    // damping -- and stores it as one 128-bit (64-bit) vector
    for (int plane = 0; plane < nplanes; plane++) {
        const int sx = plane ? fp.subx : 0, sy = plane ? fp.suby : 0;
        const int x0 = fbx * 64 >> sx, y0 = fby * 64 >> sy;
        const int16_t* tile = plane == 0 ? s_luma : s_chroma[plane - 1];
        T* dst = (T*)L.dst.p[plane];
        const int pitch_e = L.dst.pitch[plane] / sizeof(T);
        const int G = 8 >> sx, bh = 64 >> sy;                // samples per block row, plane rows in this filter block
        for (int it = tid; it < bh * 8; it += 256) {
            const int bx = it & 7, py = it >> 3, px = bx * G;
            const int x = x0 + px, y = y0 + py;
            if (x >= fp.cw[plane] || y >= fp.ch[plane]) continue;
            const int blk = ((py << sy) >> 3) * 8 + bx;
            const int pos = CDEF_TIDX(py + 2, px + 2);
            int v[8];
            bool filt = false;
            CdefBlk B;
            if (!s_skip[blk]) {
                const int ydir = s_dir[blk];
                int pri, sec, dir, damping;
                if (plane == 0) {
                    pri = fp.cdef_y_pri[idx] << cs;
                    sec = fp.cdef_y_sec[idx] << cs;
                    dir = pri == 0 ? 0 : ydir;
                    const int var = s_var[blk];
                    const int var_str = (var >> 6) ? min(31 - __clz(var >> 6), 12) : 0;
                    pri = var ? (pri * (4 + var_str) + 8) >> 4 : 0;
                    damping = fp.cdef_damping + cs;
                } else {
                    pri = fp.cdef_uv_pri[idx] << cs;
                    sec = fp.cdef_uv_sec[idx] << cs;
                    dir = pri == 0 ? 0 : c_cdef_uv_dir[fp.subx][fp.suby][ydir];
                    damping = fp.cdef_damping + cs - 1;
                }
                if (pri | sec) {
                    filt = true;
                    B.pri = pri;
                    B.sec = sec;
                    B.pt0 = ((pri >> cs) & 1) ? 3 : 4;
                    B.pt1 = ((pri >> cs) & 1) ? 3 : 2;
                    B.adj_p = pri ? max(0, damping - (31 - __clz(pri))) : 0;
                    B.adj_s = sec ? max(0, damping - (31 - __clz(sec))) : 0;
                    const int d2a = (dir + 2) & 7, d2b = (dir - 2) & 7;
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        B.op[k] = s_dtab[dir][k];
                        B.oa[k] = s_dtab[d2a][k];
                        B.ob[k] = s_dtab[d2b][k];
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (k < G) v[k] = filt ? cdef_pixel(tile, pos + k, B) : (int)tile[pos + k];
            T* dp = dst + (size_t)y * pitch_e + x;
            if (x + G <= fp.cw[plane]) {
                if (sizeof(T) == 2) {
                    if (G == 8) *reinterpret_cast<uint4*>(dp) = make_uint4((uint32_t)v[0] | ((uint32_t)v[1] << 16), (uint32_t)v[2] | ((uint32_t)v[3] << 16),
                                                                             (uint32_t)v[4] | ((uint32_t)v[5] << 16), (uint32_t)v[6] | ((uint32_t)v[7] << 16));
                    else *reinterpret_cast<uint2*>(dp) = make_uint2((uint32_t)v[0] | ((uint32_t)v[1] << 16), (uint32_t)v[2] | ((uint32_t)v[3] << 16));
                } else {
                    if (G == 8) *reinterpret_cast<uint2*>(dp) = make_uint2((uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24),
                                                                             (uint32_t)v[4] | ((uint32_t)v[5] << 8) | ((uint32_t)v[6] << 16) | ((uint32_t)v[7] << 24));
                    else *reinterpret_cast<uint32_t*>(dp) = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (k < G && x + k < fp.cw[plane]) dp[k] = (T)v[k];
            }
        }
    }
}

cudaError_t launch_cdef(const CdefLaunch& L, cudaStream_t s) {
    dim3 grid((L.fp.mi_cols + 15) >> 4, (L.fp.mi_rows + 15) >> 4);
    {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (!(dev < 64 && g_cdef_lines_loaded[dev])) {
            static CdefLineTable host_tab;
            cdef_build_lines(host_tab);
            if ((e = cudaMemcpyToSymbol(g_cdef_lines, &host_tab, sizeof(host_tab))) != cudaSuccess) return e;
            if (dev < 64) g_cdef_lines_loaded[dev] = true;
        }
    }
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(cdef_kernel<uint8_t>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(cdef_kernel<uint16_t>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr_done = true;
    }
    if (L.fp.bd == 8) cdef_kernel<uint8_t><<<grid, 256, 0, s>>>(L);
    else cdef_kernel<uint16_t><<<grid, 256, 0, s>>>(L);
    return cudaGetLastError();
}

}  // namespace av1r
