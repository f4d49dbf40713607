// K3 -- intra prediction + residual add, *unit-resident dataflow* kernel (AV1 spec 7.11.2, 7.11.4, 7.11.5, 7.12.3).
//
// Intra prediction reads reconstructed neighbours, so transform blocks form a dependency DAG; the stage is bound by the latency
// of one link of that DAG times its depth, not by bytes.  v5 gave every record its own warp and handed samples over through
// L2 (one link = flag poll + fence + L2 edge gather + fence: ~1.5 us).  v6 keeps the hand-over on chip:
//   * one CTA owns one 64x64 luma unit (+ its two 32x32 chroma tiles) at a time.  The unit's samples live in a shared-memory
//     canvas together with the halo they can read: the row above (corner .. above-right) and the column to the left
//     (.. below-left; above-right samples beside the unit never exist: the unit to the right is decoded later).
//   * inside the CTA the unit's records are dealt round-robin to NW warps; a warp looks up the owners of the 4x4 cells its edges
//     come from in a shared-memory owner map and waits on those records' completion barriers (one mbarrier per record:
//     try_wait suspends the warp in hardware), predicts from the canvas into the canvas, adds the residual (brought in by one
//     TMA bulk copy per unit: the residual is stored unit-major) and arrives on its own barrier.  One link = one predictor
//     call (~3000 cycles of dependent instructions), no L2 round trip.
//   * units are listed by the host in wavefront order with the table indices of the (<= 5) neighbour units they read; a
//     persistent grid claims units through a global ticket.  Hand-over between units is either whole-unit (wait for the
//     neighbours' flags, load the halo, write the unit back, release the flag) or -- default -- cell-level: every unit has a
//     64-bit progress word with a bit per 4-sample cell of its bottom row and right column; the record that last writes such
//     cells stores them, fences and publishes the bits, and the border records of the neighbour units wait for exactly the bits
//     they need and fetch those samples into their halo on demand, so a unit starts while its neighbours are still busy.
//     A CTA only waits for lower tickets, which are held by CTAs that already run: no deadlock.
// Algorithmic bytes: F_intra written + 2A residual read + 32 B/record (+ halo re-reads, L2 hits).
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
// base angle of the directional modes V_PRED .. D67_PRED (index = PredMode 1..8), packed one byte each
__device__ __forceinline__ int mode_angle(int mode) {
    constexpr unsigned long long pk = (90ull << 8) | (180ull << 16) | (45ull << 24) | (135ull << 32) | (113ull << 40) | (157ull << 48) | (203ull << 56);
    return mode == 8 ? 67 : (int)((pk >> (8 * mode)) & 0xff);
}
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

typedef int16_t edge_t;                     // samples are <= 12 bits
static constexpr int K3_MAX_WARPS = 8;
static constexpr int EDGE_PAD = 16;
static constexpr int EDGE_LEN = EDGE_PAD + 128 + 16;   // w + h <= 128 samples per edge; upsampled edges (w + h <= 16) need 2x + 2
static constexpr int CV_LEFT = 16;          // canvas columns to the left of the unit (only column -1 is used; 16 keeps rows 16-byte aligned)
static constexpr int CV_W0 = 64, CV_W1 = 32;            // unit tile sizes (4:2:0)
static constexpr int CV_CS0 = CV_LEFT + 2 * CV_W0;      // canvas row stride, luma (covers the above-right samples)
static constexpr int CV_CS1 = CV_LEFT + 2 * CV_W1;
static constexpr int CV_ELEMS = (CV_W0 + 1) * CV_CS0 + 2 * (CV_W1 + 1) * CV_CS1;   // rows -1 .. H-1
static constexpr int RES_ELEMS = CV_W0 * CV_W0 + 2 * CV_W1 * CV_W1;
static constexpr int OWNER_CELLS = 16 * 16 + 2 * 8 * 8;

struct WarpScratch {
    edge_t above[2][EDGE_LEN];
    edge_t left[2][EDGE_LEN];
    int16_t tile[32 * 32];       // filter-intra predictions / CfL luma terms
};

// Where the unit's samples are: pure arithmetic on the plane index (the canvas geometry is fixed at 4:2:0), kept in registers.
template <typename T>
struct UnitCtx {
    T* canvas;           // plane p: rows -1 .. H-1 of stride cs(p), columns -CV_LEFT .. 2W-1; planes back to back
    T* lext;             // below-left columns: sample (x0 - 1, y0 + H + k) of plane p at lext[lext_off(p) + k]
    const int16_t* res;  // residual of the unit (plane tiles of W x H int16)
    int ux, uy;
    __device__ __forceinline__ static int W(int p) { return p ? CV_W1 : CV_W0; }
    __device__ __forceinline__ static int cs(int p) { return p ? CV_CS1 : CV_CS0; }
    __device__ __forceinline__ static int cv_off(int p) {   // element offset of sample (x0, y0) of plane p
        return (p == 0 ? 0 : (p == 1 ? (CV_W0 + 1) * CV_CS0 : (CV_W0 + 1) * CV_CS0 + (CV_W1 + 1) * CV_CS1)) + cs(p) + CV_LEFT;
    }
    __device__ __forceinline__ static int lext_off(int p) { return p == 0 ? 0 : (p == 1 ? CV_W0 : CV_W0 + CV_W1); }
    __device__ __forceinline__ static int res_off(int p) { return p == 0 ? 0 : (p == 1 ? CV_W0 * CV_W0 : CV_W0 * CV_W0 + CV_W1 * CV_W1); }
    __device__ __forceinline__ int x0(int p) const { return ux * W(p); }
    __device__ __forceinline__ int y0(int p) const { return uy * W(p); }
    __device__ __forceinline__ T* cv(int p) const { return canvas + cv_off(p); }
    __device__ __forceinline__ int px(int plane, int x, int y) const {
        const int dx = x - x0(plane), dy = y - y0(plane);
        const T* p = dy < W(plane) ? canvas + (cv_off(plane) + dy * cs(plane) + dx) : lext + (lext_off(plane) + dy - W(plane));   // one load, selected address
        return (int)*p;
    }
};

// log2 of the transform width / height minus 2, three bits per TxSize, as immediates (no table load on the record's critical path)
__device__ __forceinline__ int tx_lw(int txsz) {
    constexpr unsigned long long pk = 0ull | (0ull << 0) | (1ull << 3) | (2ull << 6) | (3ull << 9) | (4ull << 12) |   // 4x4 8x8 16x16 32x32 64x64
                                      (0ull << 15) | (1ull << 18) | (1ull << 21) | (2ull << 24) | (2ull << 27) | (3ull << 30) |   // 4x8 8x4 8x16 16x8 16x32 32x16
                                      (3ull << 33) | (4ull << 36) | (0ull << 39) | (2ull << 42) | (1ull << 45) | (3ull << 48) |   // 32x64 64x32 4x16 16x4 8x32 32x8
                                      (2ull << 51) | (4ull << 54);                                                               // 16x64 64x16
    return 2 + (int)((pk >> (3 * txsz)) & 7);
}
__device__ __forceinline__ int tx_lh(int txsz) {
    constexpr unsigned long long pk = 0ull | (0ull << 0) | (1ull << 3) | (2ull << 6) | (3ull << 9) | (4ull << 12) |
                                      (1ull << 15) | (0ull << 18) | (2ull << 21) | (1ull << 24) | (3ull << 27) | (2ull << 30) |
                                      (4ull << 33) | (3ull << 36) | (2ull << 39) | (0ull << 42) | (3ull << 45) | (1ull << 48) |
                                      (4ull << 51) | (2ull << 54);
    return 2 + (int)((pk >> (3 * txsz)) & 7);
}

struct UnitSync {
    unsigned long long bar;   // mbarrier of the residual bulk copy
    int unit, next;
};

// ---- TMA bulk copy + mbarrier helpers (SASS UBLKCP / SYNCS)
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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return ok != 0;
}
// Polling a neighbour unit's flag uses a relaxed load (served by L2, no side effects); the acquire fence -- which invalidates the
// SM's whole L1 (SASS CCTL.IVALL) -- is paid once, after the flag has been seen, not on every poll of every waiting CTA.
__device__ __forceinline__ int ld_relaxed(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Progressive hand-over: one 64-bit word per unit, a bit per 4-sample cell of the unit's bottom row and right column, per plane:
//   luma bottom 0..15, luma right 16..31, U bottom 32..39, U right 40..47, V bottom 48..55, V right 56..63.
// A record that writes such cells stores them to the frame, fences and ORs its bits in; records of neighbour units that read
// them wait for the bits, then fetch the samples into their canvas halo.
__device__ __forceinline__ int prog_boff(int p) { return p == 0 ? 0 : (p == 1 ? 32 : 48); }
__device__ __forceinline__ int prog_roff(int p) { return p == 0 ? 16 : (p == 1 ? 40 : 56); }
__device__ __forceinline__ unsigned long long ld_relaxed64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acquire_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// Release of border cells without touching L1: `fence.acq_rel.gpu` / __threadfence() compile to MEMBAR.ALL.GPU + CCTL.IVALL, and the
// CCTL invalidates the whole L1 of the SM -- the records, order lists and halo lines of *every* K3 CTA resident there (co-resident
// frames slowed each other by 1.6x at two CTAs per SM).  A release RED orders the warp's earlier stores (cumulative over the
// __syncwarp before it) with a MEMBAR only; the consumers read the cells with L1-bypassing loads after they have seen the bits,
// so nothing stale can be hit.
__device__ __forceinline__ void red_release_or64(unsigned long long* p, unsigned long long v) {
    asm volatile("red.release.gpu.global.or.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
template <typename T>
__device__ __forceinline__ T ld_cell(const T* p) {   // coherent at L2, never served by L1
    if (sizeof(T) == 1) {
        unsigned short v;   // (no 8-bit register class in inline PTX: load into a 16-bit one)
        asm volatile("ld.relaxed.gpu.global.u8 %0, [%1];" : "=h"(v) : "l"(p) : "memory");
        return (T)v;
    } else {
        unsigned short v;
        asm volatile("ld.relaxed.gpu.global.u16 %0, [%1];" : "=h"(v) : "l"(p) : "memory");
        return (T)v;
    }
}
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

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
__device__ __forceinline__ void edge_filter_d(const edge_t* src, edge_t* dst, int sz, int strength, int total, int lane) {
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
        dst[i - 1] = (edge_t)v;
    }
}

// upsample numPx samples: dst gets indices -2 .. 2*numPx-2
__device__ __forceinline__ void edge_upsample_d(const edge_t* src, edge_t* dst, int num_px, int pixmax, int lane) {
    for (int i = lane; i < num_px; i += 32) {
        // dup[k] = src[k-2] for k = 1..numPx+1, dup[0] = src[-1], dup[numPx+2] = src[numPx-1]
        const int d0 = src[max(i - 2, -1)], d1 = src[i - 1], d2 = src[i], d3 = src[min(i + 1, num_px - 1)];
        int s = -d0 + 9 * d1 + 9 * d2 - d3;
        s = min(max((s + 8) >> 4, 0), pixmax);
        dst[2 * i - 1] = (edge_t)s;
        dst[2 * i] = (edge_t)d2;
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

#ifndef K3_RARE_ATTR
#define K3_RARE_ATTR __forceinline__
#endif
// Geometry of a record inside its unit (shared by the hot predictor and the out-of-line rare paths).
template <typename T>
struct BlkCtx {
    int plane, lw, lh, w, h, x, y, pixmax, max_x, max_y, xe, ye, xw, yw, opitch, rpitch;
    T* out;
    const int16_t* rp;
    bool has_res;
    __device__ __forceinline__ BlkCtx(const TxRec& r, const UnitCtx<T>& uv, int bd, int pcw, int pch) {
        plane = r.plane;
        lw = tx_lw(r.txsz);
        lh = tx_lh(r.txsz);
        w = 1 << lw;
        h = 1 << lh;
        x = r.x4 * 4;
        y = r.y4 * 4;
        pixmax = (1 << bd) - 1;
        max_x = pcw - 1;
        max_y = pch - 1;
        xe = min(w, pcw - x);
        ye = min(h, pch - y);
        // luma is reconstructed in full also beyond the coded frame edge (the canvas holds the whole unit; the write-back clips):
        // chroma-from-luma of a block that straddles the edge reads those samples (spec MaxLumaW / MaxLumaH)
        xw = plane ? xe : w;
        yw = plane ? ye : h;
        opitch = uv.cs(plane);
        out = uv.cv(plane) + (y - uv.y0(plane)) * opitch + (x - uv.x0(plane));   // the unit's samples live in shared memory
        has_res = r.eob > 0;
        rpitch = uv.W(plane);            // residual unit: plane tiles of W x H int16, staged by the bulk copy
        rp = uv.res + uv.res_off(plane) + (y - uv.y0(plane)) * rpitch + (x - uv.x0(plane));
    }
    __device__ __forceinline__ void emit(int i, int j, int v) const {
        if (i < yw && j < xw) {
            if (has_res) v = min(max(v + (int)rp[i * rpitch + j], 0), pixmax);
            out[i * opitch + j] = (T)v;
        }
    }
};

// Record kinds that read no neighbour edge and are rare in camera content: residual of an inter-intra block, intra block copy,
// palette.  Out of line on purpose: the wavefront kernel is bound by the latency of its dependency chain, its instruction cache hit
// rate is ~91 % (ncu sm__icc_request_hit_rate), and every kilobyte of cold code inlined into the record loop showed up as lost
// throughput when several frames share an SM (profiles/r2_k3_sweep.md).  Scalars are passed by value: taking the address of the
// kernel's parameter struct would copy it to local memory.
template <typename T, int KIND>
__device__ K3_RARE_ATTR void rare_block(const TxRec r, const UnitCtx<T> uv, int bd, int pcw, int pch, int psx, int psy, int lane,
                                        const uint8_t* pal, const uint8_t* fb, uint32_t pitch) {
    const BlkCtx<T> c(r, uv, bd, pcw, pch);
    const int w = c.w, h = c.h, lw = c.lw, x = c.x, y = c.y;
    if (r.mode == TXM_INTER) {
        // plain inter residuals were added by the K2 residual kernel; only inter-intra blocks wait for their blend
        if (c.has_res && (r.flags & TXF_II))
            for (int idx = lane; idx < w * h; idx += 32) {
                const int i = idx >> lw, j = idx & (w - 1);
                if (i < c.ye && j < c.xe) {
                    int v = (int)c.out[i * c.opitch + j] + (int)c.rp[i * c.rpitch + j];
                    c.out[i * c.opitch + j] = (T)min(max(v, 0), c.pixmax);
                }
            }
        return;
    }
    if (KIND < 2) return;
    if (r.mode == TXM_INTRABC) {
        // spec 7.11.3.2 - 7.11.3.4 with use_intrabc: the reference is this frame (no filter runs on such frames), displaced by the
        // block vector; bilinear taps at 1/16 sample (the chroma of an odd luma vector sits on a half sample), rounding 3 then 11,
        // positions clamped to the coded plane.  The source units are complete and released (the caller waited for their flags);
        // the loads bypass L1.
        const int dvx = (int16_t)r.cfl_max_w4, dvy = (int16_t)r.cfl_max_h4;
        const int posx = (x << 4) + ((2 * dvx) >> psx), posy = (y << 4) + ((2 * dvy) >> psy);
        const int ix = posx >> 4, fx = posx & 15, iy = posy >> 4, fy = posy & 15;
        for (int idx = lane; idx < w * h; idx += 32) {
            const int i = idx >> lw, j = idx & (w - 1);
            const int xa = min(max(ix + j, 0), c.max_x), xb = min(max(ix + j + 1, 0), c.max_x);
            const int ya = min(max(iy + i, 0), c.max_y), yb = min(max(iy + i + 1, 0), c.max_y);
            const T* r0 = reinterpret_cast<const T*>(fb + (size_t)ya * pitch);
            const T* r1 = reinterpret_cast<const T*>(fb + (size_t)yb * pitch);
            const int t0 = ((128 - 8 * fx) * (int)ld_cell<T>(r0 + xa) + 8 * fx * (int)ld_cell<T>(r0 + xb) + 4) >> 3;
            const int t1 = ((128 - 8 * fx) * (int)ld_cell<T>(r1 + xa) + 8 * fx * (int)ld_cell<T>(r1 + xb) + 4) >> 3;
            c.emit(i, j, min(max(((128 - 8 * fy) * t0 + 8 * fy * t1 + 1024) >> 11, 0), c.pixmax));
        }
        return;
    }
    {   // TXM_PALETTE, spec 7.11.4; entry layout documented at TileDecoder::palette_tokens
        const uint8_t* e = pal + r.pal_off;
        const uint16_t* hdr = reinterpret_cast<const uint16_t*>(e);
        const uint8_t* map = e + 24;
        const int ox = hdr[8], oy = hdr[9], stride = hdr[10];
        for (int idx = lane; idx < w * h; idx += 32) {
            const int i = idx >> lw, j = idx & (w - 1);
            c.emit(i, j, (int)hdr[map[(size_t)(y - oy + i) * stride + (x - ox + j)]]);
        }
    }
}

// Inter-intra (inter frames only): the intra predictor of the whole block is blended over the inter predictor K2 left in the
// frame.  ii_save parks the inter predictor in the warp's scratch tile (blocks are at most 32x32), the ordinary predictor then
// writes the intra predictor into the canvas, ii_blend mixes the two (spec 7.11.3.13).  Both out of line: see rare_block.
template <typename T>
__device__ K3_RARE_ATTR void ii_save(const TxRec r, const UnitCtx<T> uv, int bd, int pcw, int pch, int16_t* tile, int lane) {
    const BlkCtx<T> c(r, uv, bd, pcw, pch);
    for (int idx = lane; idx < c.w * c.h; idx += 32) {
        const int i = idx >> c.lw, j = idx & (c.w - 1);
        if (i < c.yw && j < c.xw) tile[idx] = (int16_t)c.out[i * c.opitch + j];
    }
}
template <typename T>
__device__ K3_RARE_ATTR void ii_blend(const TxRec r, const UnitCtx<T> uv, int bd, int pcw, int pch, int psx, int psy, const int16_t* tile,
                                      const uint8_t* wedge_master, int lane) {
    const BlkCtx<T> c(r, uv, bd, pcw, pch);
    const int ii_pk = (uint16_t)r.cfl_alpha;
    for (int idx = lane; idx < c.w * c.h; idx += 32) {
        const int i = idx >> c.lw, j = idx & (c.w - 1);
        if (i < c.yw && j < c.xw) {
            const int m = ii_mask(ii_pk, wedge_master, i, j, c.w, c.h, psx, psy);
            c.out[i * c.opitch + j] = (T)((m * (int)c.out[i * c.opitch + j] + (64 - m) * (int)tile[idx] + 32) >> 6);
        }
    }
}

// Filter-intra (spec 7.11.2.3): 4x2 sub-blocks along anti-diagonals, each from seven neighbours.  Rare -> out of line (see rare_block).
template <typename T>
__device__ K3_RARE_ATTR void filter_intra_block(const TxRec r, const UnitCtx<T> uv, int bd, int pcw, int pch, const edge_t* above, const edge_t* left,
                                                int16_t* pt, int lane) {
    const BlkCtx<T> c(r, uv, bd, pcw, pch);
    const int w = c.w, h = c.h, lw = c.lw, pixmax = c.pixmax;
    {
        const int w4 = w >> 2, h2 = h >> 1;
        const int fm = r.fi_mode;
        // a lane always computes the same sample of a 4x2 sub-block (t strides by 32, o = t & 7): its seven taps are loaded once;
        // indexed by o inside the loop they were an 8-way divergent constant-memory access per tap and iteration
        int tap[7];
#pragma unroll
        for (int i = 0; i < 7; i++) tap[i] = c_fi_taps[fm][lane & 7][i];
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
                for (int i = 0; i < 7; i++) pr += tap[i] * p[i];
                const int v = pr >= 0 ? (pr + 8) >> 4 : -((-pr + 8) >> 4);
                pt[((i2 << 1) + (o >> 2)) * w + (j4 << 2) + (o & 3)] = (int16_t)min(max(v, 0), pixmax);
            }
            __syncwarp();
        }
        for (int idx = lane; idx < w * h; idx += 32) c.emit(idx >> lw, idx & (w - 1), pt[idx]);
    }
}

template <typename T>
__device__ __forceinline__ void intra_pred(const TxRec& r, const UnitCtx<T>& uv, const DevFrameParams& fp, WarpScratch& sm, int lane,
                                           const uint8_t* smw) {
    const int bd = fp.bd;
    // (two uniform parameter loads and a select instead of a register-indexed constant load: planes 1 and 2 have the same size)
    const BlkCtx<T> c(r, uv, bd, r.plane ? fp.cw[1] : fp.cw[0], r.plane ? fp.ch[1] : fp.ch[0]);
    const int plane = c.plane, lw = c.lw, lh = c.lh, w = c.w, h = c.h, x = c.x, y = c.y, pixmax = c.pixmax, max_x = c.max_x, max_y = c.max_y;
    auto emit = [&](int i, int j, int v) { c.emit(i, j, v); };
    const int have_left = r.flags & TXF_HAVE_LEFT, have_above = r.flags & TXF_HAVE_ABOVE;
    const int have_ar = r.flags & TXF_HAVE_ABOVE_RIGHT, have_bl = r.flags & TXF_HAVE_BELOW_LEFT;
    edge_t* above = sm.above[0] + EDGE_PAD;
    edge_t* left = sm.left[0] + EDGE_PAD;
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
        above[i] = (edge_t)a;
        left[i] = (edge_t)l;
    }
    if (lane == 0) {
        int c;
        if (have_above && have_left) c = uv.px(plane, x - 1, y - 1);
        else if (have_above) c = uv.px(plane, x, y - 1);
        else if (have_left) c = uv.px(plane, x - 1, y);
        else c = 1 << (bd - 1);
        above[-1] = (edge_t)c;
        left[-1] = (edge_t)c;
    }
    __syncwarp();
    int mode = r.mode;
    if (mode == TXM_CFL) mode = DC_PRED;

    if (mode == TXM_FILTER_INTRA) {
        filter_intra_block<T>(r, uv, bd, plane ? fp.cw[1] : fp.cw[0], plane ? fp.ch[1] : fp.ch[0], above, left, sm.tile, lane);
        return;
    }
    if (mode >= V_PRED && mode <= D67_PRED) {
        const int p_angle = mode_angle(mode) + r.angle_delta * 3;
        int up_above = 0, up_left = 0;
        if (fp.enable_edge_filter) {
            const int filter_type = (r.flags & TXF_SMOOTH_EDGE) ? 1 : 0;
            if (p_angle != 90 && p_angle != 180) {
                if (p_angle > 90 && p_angle < 180 && (w + h) >= 24) {
                    if (lane == 0) {
                        const int v = (left[0] * 5 + above[-1] * 6 + above[0] * 5 + 8) >> 4;
                        above[-1] = (edge_t)v;
                        left[-1] = (edge_t)v;
                    }
                    __syncwarp();
                }
                if (have_above) {
                    const int strength = edge_filter_strength_d(w, h, filter_type, p_angle - 90);
                    if (strength) {
                        const int num_px = min(w, max_x - x + 1) + (p_angle < 90 ? h : 0) + 1;
                        edge_t* dst = (above == sm.above[0] + EDGE_PAD) ? sm.above[1] + EDGE_PAD : sm.above[0] + EDGE_PAD;
                        edge_filter_d(above, dst, num_px, strength, n - 1, lane);
                        above = dst;
                        __syncwarp();
                    }
                }
                if (have_left) {
                    const int strength = edge_filter_strength_d(w, h, filter_type, p_angle - 180);
                    if (strength) {
                        const int num_px = min(h, max_y - y + 1) + (p_angle > 180 ? w : 0) + 1;
                        edge_t* dst = (left == sm.left[0] + EDGE_PAD) ? sm.left[1] + EDGE_PAD : sm.left[0] + EDGE_PAD;
                        edge_filter_d(left, dst, num_px, strength, n - 1, lane);
                        left = dst;
                        __syncwarp();
                    }
                }
            }
            up_above = edge_upsample_d(w, h, filter_type, p_angle - 90);
            if (up_above) {
                edge_t* dst = (above == sm.above[0] + EDGE_PAD) ? sm.above[1] + EDGE_PAD : sm.above[0] + EDGE_PAD;
                edge_upsample_d(above, dst, w + (p_angle < 90 ? h : 0), pixmax, lane);
                above = dst;
                __syncwarp();
            }
            up_left = edge_upsample_d(w, h, filter_type, p_angle - 180);
            if (up_left) {
                edge_t* dst = (left == sm.left[0] + EDGE_PAD) ? sm.left[1] + EDGE_PAD : sm.left[0] + EDGE_PAD;
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
        // (weights from the CTA's shared-memory copy: indexed per lane, the constant-memory table was read with up to 8 different
        // addresses per warp and load, i.e. serialised)
        const uint8_t* ww = smw + (w - 4);
        const uint8_t* wh = smw + (h - 4);
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
        s = __reduce_add_sync(0xffffffffu, s);
        if (have_left && have_above) {   // (s + (w + h) / 2) / (w + h); w + h = {1, 3, 5} * 2^k: shift, then exact division by 3 or 5
            const int lo = min(lw, lh), d = abs(lw - lh);
            const unsigned q = (unsigned)(s + ((w + h) >> 1)) >> lo;      // divisor is now 2, 3 or 5 (square, 2:1, 4:1)
            dc = d == 0 ? (int)(q >> 1) : (d == 1 ? (int)(__umulhi(q, 0xAAAAAAABu) >> 1) : (int)(__umulhi(q, 0xCCCCCCCDu) >> 2));
        }
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
        s = __reduce_add_sync(0xffffffffu, s);
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


// Warp-collective wait on the completion barriers of records of the same unit: every lane names one local record index (or
// -1).  A record's barrier (one mbarrier, arrival count 1) completes one phase per unit the CTA processes -- when the record's
// samples are in the canvas (barriers the unit does not use are stepped in the prologue, so all stay on the CTA's parity);
// try_wait suspends the warp in hardware, so waiting warps do not steal issue slots or shared-memory bandwidth from the warps
// that predict.
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return ok != 0;
}
// `stuck` (one int per frame) is raised instead of hanging if a wait makes no progress for seconds: a wrong work-list must
// end in a digest mismatch, not in a wedged GPU.
__device__ __forceinline__ void wait_local(uint32_t rbar, int id, uint32_t parity, int* stuck) {
    for (int rounds = 0; rounds < 64; rounds++) {
        const bool pend = id >= 0 && !mbar_test(rbar + 8u * (uint32_t)id, parity);
        if (!__any_sync(0xffffffffu, pend)) return;
        // sleep on the youngest pending record (the one most likely to finish last); only the try_wait itself is re-issued while
        // it times out, not the whole test / vote / reduce sequence
        const uint32_t b = rbar + 8u * (uint32_t)__reduce_max_sync(0xffffffffu, pend ? id : -1);
        int spins = 0;
        while (!mbar_try(b, parity))   // (sleeping between attempts was measured: no effect on throughput, profiles/r2_k3_sweep.md)
            if (++spins > (1 << 22)) { *stuck = 1; return; }
    }
    *stuck = 1;
}

__device__ __forceinline__ TxRec load_rec(const TxRec* p) {
    union { TxRec r; uint4 q[2]; } u;
    u.q[0] = __ldg(reinterpret_cast<const uint4*>(p));
    u.q[1] = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    return u.r;
}

// PROG (cell-by-cell hand-over between units or whole units) and PROF (phase timers) are compile-time: the record loop is bound by
// the latency of its dependency chain, and every runtime test or dead branch in it costs throughput (profiles/r2_k3_sweep.md).
// KIND selects which record kinds the build knows: 0 = camera-content intra frame (no inter-intra, inter-residual, block-copy or
// palette record can occur: those paths are compiled out of the record loop), 1 = inter frame without screen-content tools
// (adds inter-intra blends and their residuals), 2 = everything (adds palette and intra block copy).
template <typename T, int NW, bool PROG, bool PROF, int KIND>
__global__ void __launch_bounds__(NW * 32, NW == 8 ? 3 : 6) intra_unit_kernel(IntraLaunch L) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    // ---- carve (mirrored by k3_smem_bytes)
    int16_t* res_s = reinterpret_cast<int16_t*>(smem_raw);
    T* canvas = reinterpret_cast<T*>(smem_raw + RES_ELEMS * sizeof(int16_t));
    T* lext_s = canvas + CV_ELEMS;                                  // 64 + 32 + 32 samples
    int32_t* owner = reinterpret_cast<int32_t*>(lext_s + 128);
    unsigned long long* rbar_s = reinterpret_cast<unsigned long long*>(owner + OWNER_CELLS);   // one completion barrier per record
    UnitSync* us = reinterpret_cast<UnitSync*>(rbar_s + K3_UNIT_MAX_RECS);
    WarpScratch* wscr = reinterpret_cast<WarpScratch*>(reinterpret_cast<uint8_t*>(us) + 64);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nthr = NW * 32, nw = NW;
    const DevFrameParams& fp = L.fp;
    const uint32_t bar = smem_u32(&us->bar);
    constexpr uint32_t unit_bytes = RES_ELEMS * sizeof(int16_t);
    __shared__ uint8_t s_smw[128];   // smooth-predictor weights (see intra_pred)
    if (tid < (int)sizeof(c_sm_weights)) s_smw[tid] = c_sm_weights[tid];
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t parity = 0;
    int inited = 0;   // record barriers [0, inited) hold valid mbarrier objects
    uint32_t rpar = 0;   // phase parity the record barriers complete during the current unit
    const uint32_t rbar = smem_u32(rbar_s);
    // optional phase timers (AV1R_K3_PROF): thread 0 of the CTA for the unit phases, lane 0 of every warp for the record phases
    long long tp = 0;
    auto lap = [&](int slot, bool who) {
        if (PROF && who) {
            const long long t = clock64();
            atomicAdd(L.prof + slot, (unsigned long long)(t - tp));
            tp = t;
        }
    };
    if (PROF) tp = clock64();
    while (true) {
        if (tid == 0) us->unit = atomicAdd(L.ticket, 1);
        __syncthreads();
        const int u = us->unit;
        if (u >= L.n_units) return;
        lap(0, tid == 0);   // ticket
        const K3Unit* up = L.units + u;
        const int first = (int)__ldg(&up->first), count = (int)__ldg(&up->count);
        const int ux = __ldg(&up->ux), uy = __ldg(&up->uy);
        // ---- prologue (needs nothing from other units): residual bulk copy, context, owner map
        if (tid == 0) {
            us->next = 0;
            mbar_expect_tx(bar, unit_bytes);
            bulk_g2s(smem_u32(res_s), L.res.base + ((size_t)uy * L.res.units_x + ux) * L.res.unit_elems, unit_bytes, bar);
        }
        UnitCtx<T> uc;
        uc.canvas = canvas;
        uc.lext = lext_s;
        uc.res = res_s;
        uc.ux = ux;
        uc.uy = uy;
        for (int i = tid; i < OWNER_CELLS; i += nthr) owner[i] = -1;
        for (int i = tid; i < max(count, inited); i += nthr) {
            if (i >= inited) {   // first use: create the barrier and bring it to the CTA's current phase
                mbar_init(rbar + 8u * i, 1);
                if (rpar) mbar_arrive(rbar + 8u * i);
            } else if (i >= count) {
                mbar_arrive(rbar + 8u * i);   // not used by this unit: keep its phase in step
            }
        }
        inited = max(inited, count);
        __syncthreads();
        for (int k0 = warp * 32; k0 < count; k0 += nw * 32) {   // owner map: local index of the record that reconstructs each 4x4 cell
            uint2 hd = make_uint2(0, 0);
            if (k0 + lane < count) hd = __ldg(reinterpret_cast<const uint2*>(L.recs + __ldg(L.order + first + k0 + lane)));
            const int nb = min(32, count - k0);
            for (int j = 0; j < nb; j++) {
                const uint32_t hx = __shfl_sync(0xffffffffu, hd.x, j), hy = __shfl_sync(0xffffffffu, hd.y, j);
                const int x4 = hx & 0xffff, y4 = hx >> 16, plane = hy & 0xff, txsz = (hy >> 8) & 0xff;
                const int lw4 = tx_lw(txsz) - 2, w4 = 1 << lw4, h4 = 1 << (tx_lh(txsz) - 2);
                const int uw4 = plane ? 8 : 16;
                const int lx4 = x4 - ux * uw4, ly4 = y4 - uy * uw4;
                int32_t* m = owner + (plane == 0 ? 0 : (plane == 1 ? 256 : 320));
                for (int c = lane; c < w4 * h4; c += 32) {
                    const int cy = ly4 + (c >> lw4), cx = lx4 + (c & (w4 - 1));
                    if (cx >= 0 && cy >= 0 && cx < uw4 && cy < uw4) atomicMax(m + cy * uw4 + cx, k0 + j);
                }
            }
        }
        // inter frame: the unit's own inter-predicted samples (final since K2 ran: fetched before the neighbour wait, off the chain)
        if (KIND >= 1 && L.load_tile) {   // (a row per warp, a 32-bit word per lane: a row of the unit is at most 32 words)
#pragma unroll 1
            for (int p = 0; p < (fp.mono ? 1 : 3); p++) {
                const int W = p ? CV_W1 : CV_W0, cs = p ? CV_CS1 : CV_CS0;
                const int x0 = ux * W, y0 = uy * W;
                const uint8_t* fb = p == 0 ? L.frame.p[0] : (p == 1 ? L.frame.p[1] : L.frame.p[2]);
                const uint32_t pitch = p == 0 ? L.frame.pitch[0] : L.frame.pitch[1];
                T* cv = uc.cv(p);
                const int vw = min(W, (p ? fp.cw[1] : fp.cw[0]) - x0), vh = min(W, (p ? fp.ch[1] : fp.ch[0]) - y0);
                const int wpr = vw * (int)sizeof(T) / 4;
                if (lane < wpr)
                    for (int row = warp; row < vh; row += nw)
                        reinterpret_cast<uint32_t*>(cv + row * cs)[lane] =
                            __ldcg(reinterpret_cast<const uint32_t*>(fb + (size_t)(y0 + row) * pitch + (size_t)x0 * sizeof(T)) + lane);
            }
        }
        // the first record of every warp does not depend on the neighbours either
        int k = warp;
        TxRec r_next;
        if (k < count) r_next = load_rec(L.recs + __ldg(L.order + first + k));
        lap(1, tid == 0);   // prologue: context, owner map
        if (PROG) {
            // border cells that no record of this unit reconstructs (inter-predicted samples, cells outside the frame) are final
            __syncthreads();   // owner map complete
            if (warp == 0) {
                unsigned long long pre = 0;
#pragma unroll
                for (int p = 0; p < 3; p++) {
                    const int uw4 = p ? 8 : 16;
                    const int32_t* m = owner + (p == 0 ? 0 : (p == 1 ? 256 : 320));
                    const bool inb = lane < uw4;
                    const unsigned bb = __ballot_sync(0xffffffffu, inb && m[(uw4 - 1) * uw4 + lane] < 0);
                    const unsigned rb = __ballot_sync(0xffffffffu, inb && m[lane * uw4 + uw4 - 1] < 0);
                    pre |= (unsigned long long)bb << prog_boff(p);
                    pre |= (unsigned long long)rb << prog_roff(p);
                }
                if (lane == 0 && pre) atomicOr(L.uprog + u, pre);
            }
        }
        // ---- wait for the neighbour units whose samples this unit reads
        if (warp == 0 && !PROG) {
            if (lane < 5) {
                const int d = __ldg(&up->dep[lane]);
                if (d >= 0) {
                    int spins = 0;
                    while (ld_relaxed(L.uflags + d) == 0) {   // <= 5 polling lanes per CTA
                        __nanosleep(100);
                        if (++spins > (1 << 25)) { *L.stuck = 2; break; }
                    }
                    fence_acquire_gpu();
                }
            }
            __syncwarp();
        }
        __syncthreads();
        lap(2, tid == 0);   // neighbour units
        // ---- halo (and, in inter frames, the unit's own inter-predicted samples) from L2
        if (!PROG) {   // row above (corner .. above-right) and column to the left (.. below-left) of the three planes: one flat index space, all
            // loads of a thread issued before the first store, so the halo costs one L2 round trip
            constexpr int SEG0 = 2 * (2 * CV_W0 + 1), SEG1 = 2 * (2 * CV_W1 + 1);   // per plane: [row: 2W + 1][col: 2W] (+1 pad)
            constexpr int TOT = SEG0 + 2 * SEG1;
            constexpr int PER = (TOT + nthr - 1) / nthr;
            // decode element e -> global source (null: nothing to fetch) and canvas destination
            auto halo = [&](int e, const T*& src, T*& dst) {
                src = nullptr;
                dst = nullptr;
                if (e >= TOT) return;
                const int p = e < SEG0 ? 0 : (e < SEG0 + SEG1 ? 1 : 2);
                const int W = p ? CV_W1 : CV_W0, cs = p ? CV_CS1 : CV_CS0;
                const int i = e - (p == 0 ? 0 : (p == 1 ? SEG0 : SEG0 + SEG1));
                const int x0 = ux * W, y0 = uy * W;
                const uint8_t* fb = L.frame.p[p];
                const uint32_t pitch = L.frame.pitch[p];
                if (i < 2 * W + 1) {
                    const int dx = i - 1;
                    if (y0 > 0 && x0 + dx >= 0 && x0 + dx < (p ? fp.cw[1] : fp.cw[0])) {
                        src = reinterpret_cast<const T*>(fb + (size_t)(y0 - 1) * pitch) + x0 + dx;
                        dst = uc.cv(p) - cs + dx;
                    }
                } else {
                    const int dy = i - (2 * W + 1);
                    if (x0 > 0 && dy < 2 * W && y0 + dy < (p ? fp.ch[1] : fp.ch[0])) {
                        src = reinterpret_cast<const T*>(fb + (size_t)(y0 + dy) * pitch) + x0 - 1;
                        dst = dy < W ? uc.cv(p) + dy * cs - 1 : uc.lext + uc.lext_off(p) + (dy - W);
                    }
                }
            };
            T v[PER];
#pragma unroll
            for (int j = 0; j < PER; j++) {
                const T* src;
                T* dst;
                halo(tid + j * nthr, src, dst);
                v[j] = src ? __ldcg(src) : (T)0;
            }
#pragma unroll
            for (int j = 0; j < PER; j++) {
                const T* src;
                T* dst;
                halo(tid + j * nthr, src, dst);
                if (dst) *dst = v[j];
            }
        }
        lap(3, tid == 0);   // halo loads
        mbar_wait(bar, parity);
        parity ^= 1;
        __syncthreads();
        lap(4, tid == 0);   // residual bulk copy
        if (PROF && lane == 0 && tid != 0) tp = clock64();
        // ---- record-level dataflow inside the unit
        // Records are dealt round-robin to the warps (record k -> warp k mod NW; a warp walks its records in order, so the lowest
        // unfinished record can always run) and the next record is fetched while the current one waits and predicts.
        WarpScratch& sm = wscr[warp];
        for (; k < count; k += nw) {
            const TxRec r = r_next;
            if (k + nw < count) r_next = load_rec(L.recs + __ldg(L.order + first + k + nw));
            lap(8, lane == 0);   // record fetch
            if (KIND >= 1 && r.mode == TXM_INTER) {
                if (r.flags & TXF_II) wait_local(rbar, lane == 0 ? (int)r.pal_off - first : -1, rpar, L.stuck);   // residual of an inter-intra block: after its blend
            } else if (KIND == 2 && r.mode == TXM_INTRABC) {
                // block copy: wait until every unit the source rectangle touches is complete in the frame (whole-unit flags; the
                // plan guarantees they come earlier in the table, so this cannot wait on a unit that waits on us)
#ifndef K3_NO_IBC_WAIT
                if (L.upos) {
                    const int plane = r.plane;
                    const int psx = plane ? fp.subx : 0, psy = plane ? fp.suby : 0;
                    const int pcw = plane ? fp.cw[1] : fp.cw[0], pch = plane ? fp.ch[1] : fp.ch[0];
                    const int bx = r.x4 * 4, by = r.y4 * 4;
                    const int bw = max(1, min(1 << tx_lw(r.txsz), pcw - bx)), bh = max(1, min(1 << tx_lh(r.txsz), pch - by));
                    const int posx = (bx << 4) + ((2 * (int)(int16_t)r.cfl_max_w4) >> psx), posy = (by << 4) + ((2 * (int)(int16_t)r.cfl_max_h4) >> psy);
                    const int us = 6 - psx, vs = 6 - psy;
                    const int sx0 = (posx >> 4) >> us, sy0 = (posy >> 4) >> vs;
                    const int sx1 = ((posx >> 4) + bw - 1 + ((posx & 15) != 0)) >> us, sy1 = ((posy >> 4) + bh - 1 + ((posy & 15) != 0)) >> vs;
                    const int nx = sx1 - sx0 + 1, ny = sy1 - sy0 + 1;
                    const int UX = (fp.mi_cols + 15) >> 4, UY = (fp.mi_rows + 15) >> 4;
                    if (lane < nx * ny && lane < 32) {
                        const int xx = sx0 + lane % nx, yy = sy0 + lane / nx;
                        if (xx >= 0 && yy >= 0 && xx < UX && yy < UY) {
                            const int d = __ldg(L.upos + yy * UX + xx);
                            if (d >= 0 && d < u) {
                                int spins = 0;
                                while (ld_relaxed(L.uflags + d) == 0) {
                                    __nanosleep(200);
                                    if (++spins > (1 << 24)) { *L.stuck = 4; break; }
                                }
                            }
                        }
                    }
                    __syncwarp();
                    fence_acquire_gpu();
                }
#endif
            } else if (KIND < 2 || r.mode != TXM_PALETTE) {
                const int plane = r.plane;
                const int w4 = 1 << (tx_lw(r.txsz) - 2), h4 = 1 << (tx_lh(r.txsz) - 2);
                const int pw4 = plane ? fp.pw4[1] : fp.pw4[0], ph4 = plane ? fp.ph4[1] : fp.ph4[0];
                const int uw4 = plane ? 8 : 16;
                const int bx4 = ux * uw4, by4 = uy * uw4;
                const int have_left = r.flags & TXF_HAVE_LEFT, have_above = r.flags & TXF_HAVE_ABOVE;
                int need_ar = 0, need_bl = 0;
                if (r.mode >= V_PRED && r.mode <= D67_PRED) {
                    const int p_angle = mode_angle(r.mode) + r.angle_delta * 3;
                    need_ar = p_angle < 90 && (r.flags & TXF_HAVE_ABOVE_RIGHT);
                    need_bl = p_angle > 180 && (r.flags & TXF_HAVE_BELOW_LEFT);
                }
                // V_PRED reads only the row above, H_PRED only the left column (unless that edge is missing and the other one stands in)
                const int want_above = have_above && !(r.mode == H_PRED && r.angle_delta == 0 && have_left);
                const int want_left = have_left && !(r.mode == V_PRED && r.angle_delta == 0 && have_above);
                const int na = want_above ? w4 * (need_ar ? 2 : 1) + 1 : 0;      // cells x4-1 .. on row y4-1 (the first is the corner)
                const int nl = want_left ? h4 * (need_bl ? 2 : 1) : 0;            // cells y4 .. on column x4-1
                const int32_t* m = owner + (plane == 0 ? 0 : (plane == 1 ? 256 : 320));
                const int pcw = plane ? fp.cw[1] : fp.cw[0], pch = plane ? fp.ch[1] : fp.ch[0];
                for (int c0 = 0; c0 < na + nl; c0 += 32) {       // warp-uniform trip count: the wait is collective
                    const int c = c0 + lane;
                    int id = -1;
                    int nbr = -1, bit = 0, cxl = 0, cyl = 0;     // cell outside the unit: neighbour (index into K3Unit::dep) and its progress bit
                    if (c < na + nl) {
                        int cx, cy;
                        if (c < na) {
                            cx = r.x4 - 1 + c;
                            cy = r.y4 - 1;
                        } else {
                            cx = r.x4 - 1;
                            cy = r.y4 + (c - na);
                        }
                        cx = min(max(cx, 0), pw4 - 1) - bx4;
                        cy = min(max(cy, 0), ph4 - 1) - by4;
                        cxl = cx;
                        cyl = cy;
                        if (cx >= 0 && cy >= 0 && cx < uw4 && cy < uw4) {   // (whole-unit hand-over: cells of other units were final before the unit started)
                            id = m[cy * uw4 + cx];
                            if (id >= k) id = -1;
                        } else if (PROG) {
                            if (cy < 0) {            // row above the unit: above-left / above / above-right unit, bottom row of cells
                                nbr = cx < 0 ? 2 : (cx < uw4 ? 3 : 4);
                                bit = prog_boff(plane) + (cx < 0 ? uw4 - 1 : (cx < uw4 ? cx : cx - uw4));
                            } else if (cx < 0) {     // column left of the unit: left / below-left unit, right column of cells
                                nbr = cy < uw4 ? 0 : 1;
                                bit = prog_roff(plane) + (cy < uw4 ? cy : cy - uw4);
                            }
                        }
                    }
                    if (PROG && __any_sync(0xffffffffu, nbr >= 0)) {
                        // lane n < 5 collects the cells wanted from neighbour n and polls that unit's progress word
                        const unsigned long long mine = nbr >= 0 ? 1ull << bit : 0ull;
                        unsigned long long want = 0;
#pragma unroll
                        for (int n = 0; n < 5; n++) {
                            const unsigned lo = __reduce_or_sync(0xffffffffu, nbr == n ? (unsigned)mine : 0u);
                            const unsigned hi = __reduce_or_sync(0xffffffffu, nbr == n ? (unsigned)(mine >> 32) : 0u);
                            if (lane == n) want = ((unsigned long long)hi << 32) | lo;
                        }
                        if (lane < 5 && want) {
                            const int d = __ldg(&up->dep[lane]);
                            if (d >= 0) {
                                int spins = 0;
                                // back-off 100 ns .. 0.8 us: a frame's chain crosses ~100 unit borders, so even the longest sleep adds a few
                                // per cent to its latency, while the polls were a fifth of the kernel's issued instructions
                                unsigned ns = 100;
                                while ((ld_relaxed64(L.uprog + d) & want) != want) {
                                    __nanosleep(ns);
                                    if (ns < 800) ns <<= 1;
                                    if (++spins > (1 << 23)) { *L.stuck = 3; break; }
                                }
                            }
                        }
                        __syncwarp();   // every lane is past the polls that saw the bits: the cells are in L2 (released before the bits)
                        if (nbr >= 0) {   // bring the cell's four samples next to the unit into the canvas halo
                            T* cv = uc.cv(plane);
                            const int cs = uc.cs(plane), W = uc.W(plane);
                            const int x0 = ux * W, y0 = uy * W;
                            const uint8_t* fb = L.frame.p[plane];
                            const uint32_t pitch = L.frame.pitch[plane];
                            if (cyl < 0) {
                                const T* src = reinterpret_cast<const T*>(fb + (size_t)(y0 - 1) * pitch);
#pragma unroll
                                for (int j = 0; j < 4; j++) {
                                    const int dx = cxl * 4 + j;
                                    if (x0 + dx < pcw) cv[-cs + dx] = ld_cell<T>(src + x0 + dx);
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 4; j++) {
                                    const int dy = cyl * 4 + j;
                                    if (y0 + dy < pch) {
                                        const T v = ld_cell<T>(reinterpret_cast<const T*>(fb + (size_t)(y0 + dy) * pitch) + x0 - 1);
                                        if (dy < W) cv[dy * cs - 1] = v;
                                        else uc.lext[uc.lext_off(plane) + dy - W] = v;
                                    }
                                }
                            }
                        }
                        __syncwarp();
                    }
                    wait_local(rbar, id, rpar, L.stuck);
                }
                if (r.mode == TXM_CFL) {   // luma samples under this chroma block
                    const int sx = fp.subx, sy = fp.suby;
                    const int lx0 = (r.x4 << sx), ly0 = (r.y4 << sy);
                    const int lx1 = min((r.x4 + w4) << sx, (int)r.cfl_max_w4), ly1 = min((r.y4 + h4) << sy, (int)r.cfl_max_h4);
                    const int lw4 = max(lx1 - lx0, 0), lh4 = max(ly1 - ly0, 0);
                    for (int c0 = 0; c0 < lw4 * lh4; c0 += 32) {
                        const int c = c0 + lane;
                        int id = -1;
                        if (c < lw4 * lh4) {
                            const int cx = lx0 + c % lw4 - ux * 16, cy = ly0 + c / lw4 - uy * 16;
                            if (cx >= 0 && cy >= 0 && cx < 16 && cy < 16) {
                                id = owner[cy * 16 + cx];
                                if (id >= k) id = -1;
                            }
                        }
                        wait_local(rbar, id, rpar, L.stuck);
                    }
                }
            }
            // the waits acquire (mbarrier test/try_wait) what the owners released with their arrive; __syncwarp orders the lanes
            __syncwarp();
            lap(9, lane == 0);   // dependency wait
            {
                const int plane = r.plane;
                const int pcw = plane ? fp.cw[1] : fp.cw[0], pch = plane ? fp.ch[1] : fp.ch[0];
                const int psx = plane ? fp.subx : 0, psy = plane ? fp.suby : 0;
                if (KIND >= 1 && (r.mode == TXM_INTER || (KIND == 2 && (r.mode == TXM_INTRABC || r.mode == TXM_PALETTE)))) {
                    rare_block<T, KIND>(r, uc, fp.bd, pcw, pch, psx, psy, lane, L.pal, plane == 0 ? L.frame.p[0] : (plane == 1 ? L.frame.p[1] : L.frame.p[2]),
                                  plane == 0 ? L.frame.pitch[0] : L.frame.pitch[1]);
                } else {
                    const bool ii = KIND >= 1 && (r.flags & TXF_II) != 0;
                    if (ii) {
                        ii_save<T>(r, uc, fp.bd, pcw, pch, sm.tile, lane);
                        __syncwarp();
                    }
                    intra_pred<T>(r, uc, fp, sm, lane, s_smw);
                    if (ii) {
                        __syncwarp();
                        ii_blend<T>(r, uc, fp.bd, pcw, pch, psx, psy, sm.tile, L.wedge_master, lane);
                    }
                }
            }
            __syncwarp();        // all lanes' samples are in the canvas before lane 0 releases the record's barrier
            if (lane == 0) mbar_arrive(rbar + 8u * (uint32_t)k);
            if (PROG) {
                // cells of the unit's bottom row / right column this record is the last writer of: store them to the frame now and
                // publish them, so that the units below and to the right can start on them while this unit is still busy
                const int plane = r.plane;
                const int uw4 = plane ? 8 : 16, W = uc.W(plane), cs = uc.cs(plane);
                const int w4 = 1 << (tx_lw(r.txsz) - 2), h4 = 1 << (tx_lh(r.txsz) - 2);
                const int lx4 = r.x4 - ux * uw4, ly4 = r.y4 - uy * uw4;
                const int32_t* m = owner + (plane == 0 ? 0 : (plane == 1 ? 256 : 320));
                const int pcw = plane ? fp.cw[1] : fp.cw[0], pch = plane ? fp.ch[1] : fp.ch[0];
                const int x0 = ux * W, y0 = uy * W;
                const T* cv = uc.cv(plane);
                uint8_t* fb = L.frame.p[plane];
                const uint32_t pitch = L.frame.pitch[plane];
                unsigned long long bits = 0;
                if (ly4 + h4 == uw4) {
                    const unsigned ob = __ballot_sync(0xffffffffu, lane < w4 && m[(uw4 - 1) * uw4 + lx4 + lane] == k);
                    T* dst = reinterpret_cast<T*>(fb + (size_t)(y0 + W - 1) * pitch) + x0;
                    for (int i = lane; i < 4 * w4; i += 32) {
                        const int dx = lx4 * 4 + i;
                        if (((ob >> (i >> 2)) & 1) && x0 + dx < pcw && y0 + W - 1 < pch) dst[dx] = cv[(W - 1) * cs + dx];
                    }
                    bits |= (unsigned long long)ob << (prog_boff(plane) + lx4);
                }
                if (lx4 + w4 == uw4) {
                    const unsigned ob = __ballot_sync(0xffffffffu, lane < h4 && m[(ly4 + lane) * uw4 + uw4 - 1] == k);
                    for (int i = lane; i < 4 * h4; i += 32) {
                        const int dy = ly4 * 4 + i;
                        if (((ob >> (i >> 2)) & 1) && y0 + dy < pch && x0 + W - 1 < pcw)
                            reinterpret_cast<T*>(fb + (size_t)(y0 + dy) * pitch)[x0 + W - 1] = cv[dy * cs + W - 1];
                    }
                    bits |= (unsigned long long)ob << (prog_roff(plane) + ly4);
                }
                if (bits) {
                    __syncwarp();
                    if (lane == 0) red_release_or64(L.uprog + u, bits);
                }
            }
            lap(10, lane == 0);  // predict + reconstruct
            if (PROF && lane == 0) atomicAdd(L.prof + 12, 1ull);
        }
        lap(11, lane == 0 && tid != 0);   // warp idle at the end of the unit
        __syncthreads();
        lap(5, tid == 0);   // dataflow (CTA view)
        // ---- write the unit back (coalesced 32-bit words; coded plane widths are multiples of 4 samples)
#pragma unroll 1
        for (int p = 0; p < (fp.mono ? 1 : 3); p++) {   // (a row per warp, a 32-bit word per lane)
            const int W = p ? CV_W1 : CV_W0, cs = p ? CV_CS1 : CV_CS0;
            const int x0 = ux * W, y0 = uy * W;
            const int vw = min(W, (p ? fp.cw[1] : fp.cw[0]) - x0), vh = min(W, (p ? fp.ch[1] : fp.ch[0]) - y0);
            uint8_t* fb = p == 0 ? L.frame.p[0] : (p == 1 ? L.frame.p[1] : L.frame.p[2]);
            const uint32_t pitch = p == 0 ? L.frame.pitch[0] : L.frame.pitch[1];
            const T* cv = uc.cv(p);
            const int wpr = vw * (int)sizeof(T) / 4;
            if (lane < wpr)
                for (int row = warp; row < vh; row += nw)
                    reinterpret_cast<uint32_t*>(fb + (size_t)(y0 + row) * pitch + (size_t)x0 * sizeof(T))[lane] =
                        reinterpret_cast<const uint32_t*>(cv + row * cs)[lane];
        }
        __syncthreads();
        if (tid == 0) st_release(L.uflags + u, 1);   // release after the CTA barrier: cumulative over every thread's write-back stores
        rpar ^= 1;
        lap(6, tid == 0);   // write-back + release
        if (PROF && tid == 0) atomicAdd(L.prof + 13, 1ull);
    }
}

static size_t k3_smem_bytes(int bps, int warps) {
    return RES_ELEMS * sizeof(int16_t) + (size_t)(CV_ELEMS + 128) * bps + OWNER_CELLS * 4 + K3_UNIT_MAX_RECS * 8 + 64 + (size_t)warps * sizeof(WarpScratch);
}

template <typename F>
static void k3_for_each_kernel(F f) {   // T x PROG x {camera intra, inter, everything, everything + timers}
    f(intra_unit_kernel<uint8_t, 8, false, false, 0>); f(intra_unit_kernel<uint8_t, 8, false, false, 1>); f(intra_unit_kernel<uint8_t, 8, false, false, 2>); f(intra_unit_kernel<uint8_t, 8, false, true, 2>);
    f(intra_unit_kernel<uint8_t, 8, true, false, 0>); f(intra_unit_kernel<uint8_t, 8, true, false, 1>); f(intra_unit_kernel<uint8_t, 8, true, false, 2>); f(intra_unit_kernel<uint8_t, 8, true, true, 2>);
    f(intra_unit_kernel<uint16_t, 8, false, false, 0>); f(intra_unit_kernel<uint16_t, 8, false, false, 1>); f(intra_unit_kernel<uint16_t, 8, false, false, 2>); f(intra_unit_kernel<uint16_t, 8, false, true, 2>);
    f(intra_unit_kernel<uint16_t, 8, true, false, 0>); f(intra_unit_kernel<uint16_t, 8, true, false, 1>); f(intra_unit_kernel<uint16_t, 8, true, false, 2>); f(intra_unit_kernel<uint16_t, 8, true, true, 2>);
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
    const size_t mx = k3_smem_bytes(2, K3_MAX_WARPS);
    {
        cudaError_t ea = cudaSuccess;
        auto set = [&](auto k) { if (ea == cudaSuccess) ea = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mx); };
        k3_for_each_kernel(set);
        if (ea != cudaSuccess) return ea;
    }
    if (dev < 64) g_intra_const_loaded[dev] = true;
    return cudaSuccess;
}

cudaError_t launch_intra(const IntraLaunch& L, cudaStream_t s) {
    if (L.n_units <= 0) return cudaSuccess;
    if (!L.fp.mono && (L.fp.subx != 1 || L.fp.suby != 1)) return cudaErrorInvalidValue;   // the canvas geometry is 4:2:0 (or luma only)
    cudaError_t e = intra_upload_constants();
    if (e != cudaSuccess) return e;
    {
        static bool carve_done = false;
        if (!carve_done) {
            k3_for_each_kernel([](auto k) { prefer_max_smem(k); });
            carve_done = true;
        }
    }
    const int blocks = std::min(L.n_units, std::max(1, L.ctas));
    const bool prof = L.prof != nullptr, prog = L.progressive != 0;
    const int kind = prof ? 2 : L.kind;
    auto go = [&](auto tag_t, auto tag_prog) {
        using TT = decltype(tag_t);
        constexpr bool PG = decltype(tag_prog)::value;
        const size_t sm = k3_smem_bytes((int)sizeof(TT), 8);
        if (prof) intra_unit_kernel<TT, 8, PG, true, 2><<<blocks, 256, sm, s>>>(L);
        else if (kind == 0) intra_unit_kernel<TT, 8, PG, false, 0><<<blocks, 256, sm, s>>>(L);
        else if (kind == 1) intra_unit_kernel<TT, 8, PG, false, 1><<<blocks, 256, sm, s>>>(L);
        else intra_unit_kernel<TT, 8, PG, false, 2><<<blocks, 256, sm, s>>>(L);
    };
    if (L.fp.bd == 8) {
        if (prog) go(uint8_t(), std::true_type());
        else go(uint8_t(), std::false_type());
    } else {
        if (prog) go(uint16_t(), std::true_type());
        else go(uint16_t(), std::false_type());
    }
    return cudaGetLastError();
}

}  // namespace av1r
