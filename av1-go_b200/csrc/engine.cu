// Decode engine: host parse -> pinned staging -> one H2D per frame -> reconstruction kernels.
//
//   per frame:  [H2D work-list arena] -> K1 itx -> K3 intra wavefront -> K4 deblock (V,H) -> K5 CDEF
//               -> K8 film grain (display copy) -> plane digests -> [D2H 24 B | D2H planes in parity mode]
// Frames are issued round-robin over `streams` CUDA streams with `frames_in_flight` resource slots,
// so the low-occupancy wavefront kernel of one frame overlaps the streaming filters of others.
// No CPU fallback exists: without a CUDA device av1r_open fails.
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstring>
#include <deque>
#include <functional>
#include <future>
#include <memory>
#include <vector>

#include "../../include/av1r_stages.h"
#include "demux.h"
#include "engine.h"
#include "kernels/inter.h"
#include "kernels/intra.h"
#include "k3_plan.h"
#include "md5.h"
#include "parallel.h"
#include "stream_parser.h"

struct av1r_clip;

namespace av1r {

#define CK(call)                                                         \
    do {                                                                 \
        cudaError_t _e = (call);                                         \
        if (_e != cudaSuccess) {                                         \
            err = std::string(#call) + ": " + cudaGetErrorString(_e);    \
            return AV1R_EIO;                                             \
        }                                                                \
    } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Slot buffers of the same kind see similar sizes on every slot: `hw` (a high-water mark shared by that kind) lets a slot that has
// to grow jump straight to the largest size any slot has needed so far, so a new engine stops (re)allocating after a few frames.
struct DevBuf {
    uint8_t* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n, size_t* hw = nullptr) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = align_up(n + n / 2, 1 << 20);
        if (hw) {
            want = std::max(want, *hw);
            *hw = want;
        }
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    ~DevBuf() { if (p) cudaFree(p); }
};
struct PinBuf {
    uint8_t* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n, size_t* hw = nullptr) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = align_up(n + n / 2, 1 << 20);
        if (hw) {
            want = std::max(want, *hw);
            *hw = want;
        }
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    ~PinBuf() { if (p) cudaFreeHost(p); }
};

// A frame buffer on the device (3 planes in one allocation).
struct FrameSlab {                 // one cudaMalloc shared by several frame buffers
    uint8_t* p = nullptr;
    ~FrameSlab() { if (p) cudaFree(p); }
};

struct DevFrameBuf {
    uint8_t* base = nullptr;
    std::shared_ptr<FrameSlab> slab;   // owner of `base` when set
    DevPlanes pl;
    int cw[3], ch[3], bps;
    int w[3] = {0, 0, 0}, h[3] = {0, 0, 0};   // visible size of the frame currently held (the reference geometry scaled prediction reads)
    size_t bytes = 0;
    cudaEvent_t ready = nullptr;   // recorded when the frame's last kernel has been queued; readers on other streams wait on it
    ~DevFrameBuf() {
        if (base && !slab) cudaFree(base);
        if (ready) cudaEventDestroy(ready);
    }
};

// Allocates `count` frame buffers of one geometry out of a single device allocation (cudaMalloc is a synchronising, millisecond-scale
// call: a decoder that keeps 2-3 buffers per in-flight frame must not pay it per frame).
static bool alloc_frames(const DevFrameParams& fp, int count, std::vector<std::shared_ptr<DevFrameBuf>>& out, std::string& err) {
    const int bps = fp.bd == 8 ? 1 : 2;
    size_t off[3], total = 0;
    uint32_t pitch[3];
    for (int p = 0; p < 3; p++) {
        pitch[p] = (uint32_t)align_up((size_t)fp.cw[p] * bps, 256);
        off[p] = total;
        total += align_up((size_t)pitch[p] * fp.ch[p], 256);
    }
    auto slab = std::make_shared<FrameSlab>();
    if (cudaMalloc(&slab->p, total * count) != cudaSuccess) {
        err = "cudaMalloc(frame buffers) failed";
        return false;
    }
    for (int i = 0; i < count; i++) {
        auto f = std::make_shared<DevFrameBuf>();
        f->bps = bps;
        f->slab = slab;
        f->base = slab->p + (size_t)i * total;
        for (int p = 0; p < 3; p++) {
            f->cw[p] = fp.cw[p];
            f->ch[p] = fp.ch[p];
            f->pl.pitch[p] = pitch[p];
            f->pl.p[p] = f->base + off[p];
        }
        f->bytes = total;
        if (cudaEventCreateWithFlags(&f->ready, cudaEventDisableTiming) != cudaSuccess) {
            err = "cudaEventCreate(frame) failed";
            return false;
        }
        out.push_back(f);
    }
    return true;
}

// Layout of the per-frame work-list arena (same offsets on host staging and device).
struct WorkLayout {
    size_t recs, coefs, order, lf[3], lfblk, lftx[3], cdef_idx, skip_mi, lr[3], inter, itiles, obmc, warps, pal, k3order, k3units, k3upos, total;
    int n_lfblk, lf_device, k3upos_on;
    int n_recs, n_coefs, n_order, n_order_small, n_inter, n_itiles, n_itiles_small, n_obmc, n_warps, n_k3, n_k3units;
};

static_assert(sizeof(LrUnit) == sizeof(LrUnitDev), "LrUnit layouts must match");

struct DevWork {
    WorkLayout lay;
    DevFrameParams fp;       // geometry the frame is reconstructed at (coded width; downscaled when use_superres)
    DevFrameParams fp_up;    // geometry of the frame that is output / kept as a reference (== fp without superres)
    FrameHdr fh;
    int lf_on = 0, lf_plane_on[3] = {0, 0, 0}, cdef_on = 0, lr_on = 0;
    uint64_t coded_samples = 0, coef_tokens = 0;
    double parse_ms = 0;
    int grain_on = 0;
    int lr_rows[3] = {0, 0, 0}, lr_cols[3] = {0, 0, 0}, lr_has[3] = {0, 0, 0};
};

static void fill_params(const SeqHdr& seq, const FrameWork& fw, DevFrameParams& fp) {
    memset(&fp, 0, sizeof(fp));
    const FrameHdr& fh = fw.fh;
    fp.bd = seq.bit_depth;
    fp.subx = seq.subsampling_x;
    fp.suby = seq.subsampling_y;
    fp.mono = seq.mono_chrome;
    fp.mi_cols = fh.mi_cols;
    fp.mi_rows = fh.mi_rows;
    fp.sb128 = seq.use_128x128_superblock;
    for (int p = 0; p < 3; p++) {
        const int sx = p ? fp.subx : 0, sy = p ? fp.suby : 0;
        fp.w[p] = (fh.frame_width + sx) >> sx;
        fp.h[p] = (fh.frame_height + sy) >> sy;
        fp.cw[p] = (fh.mi_cols * 4) >> sx;
        fp.ch[p] = (fh.mi_rows * 4) >> sy;
        fp.pw4[p] = (fh.mi_cols + sx) >> sx;
        fp.ph4[p] = (fh.mi_rows + sy) >> sy;
    }
    fp.dq_dc[0] = fh.delta_q_y_dc; fp.dq_ac[0] = 0;
    fp.dq_dc[1] = fh.delta_q_u_dc; fp.dq_ac[1] = fh.delta_q_u_ac;
    fp.dq_dc[2] = fh.delta_q_v_dc; fp.dq_ac[2] = fh.delta_q_v_ac;
    fp.enable_edge_filter = seq.enable_intra_edge_filter;
    fp.lf_sharpness = fh.lf.sharpness;
    fp.cdef_damping = fh.cdef_damping;
    for (int i = 0; i < 8; i++) {
        fp.cdef_y_pri[i] = fh.cdef_y_pri[i];
        fp.cdef_y_sec[i] = fh.cdef_y_sec[i];
        fp.cdef_uv_pri[i] = fh.cdef_uv_pri[i];
        fp.cdef_uv_sec[i] = fh.cdef_uv_sec[i];
    }
}

// Computes the arena layout of a parsed frame.
static void plan_layout(const FrameWork& fw, DevWork& dw) {
    WorkLayout& L = dw.lay;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    L.n_recs = (int)fw.tx.size();
    L.n_coefs = (int)fw.coefs.size();
    int n_order = 0;
    for (const TxRec& r : fw.tx) n_order += r.eob > 0;
    L.n_order = n_order;
    L.recs = take(sizeof(TxRec) * std::max(1, L.n_recs));
    L.coefs = take(sizeof(uint32_t) * std::max(1, L.n_coefs));
    L.order = take(sizeof(uint32_t) * std::max(1, L.n_order));
    // deblocking: host-classified edges (4 B per 4x4 cell and plane), or -- the engine's own parsers -- the block list and the
    // per-4x4 transform-size maps from which the device classifies the edges itself (a quarter of the bytes, none of the host time)
    L.lf_device = !fw.host_lf;
    L.n_lfblk = (int)fw.lf_blocks.size();
    if (L.lf_device) {
        for (int p = 0; p < 3; p++) L.lf[p] = 0;
        L.lfblk = take(sizeof(LfBlk) * std::max(1, L.n_lfblk));
        for (int p = 0; p < 3; p++) L.lftx[p] = take(std::max<size_t>(1, (size_t)fw.plane_w4(p) * fw.plane_h4(p)));
    } else {
        for (int p = 0; p < 3; p++) L.lf[p] = take(sizeof(LfEdge) * std::max<size_t>(1, fw.lf[p].size()));
        L.lfblk = 0;
        for (int p = 0; p < 3; p++) L.lftx[p] = 0;
    }
    L.cdef_idx = take(std::max<size_t>(1, fw.cdef_idx.size()));
    L.skip_mi = take(std::max<size_t>(1, fw.skip_mi.size()));
    for (int p = 0; p < 3; p++) {
        L.lr[p] = take(sizeof(LrUnit) * std::max<size_t>(1, fw.lr[p].size()));
        dw.lr_rows[p] = fw.lr_rows[p];
        dw.lr_cols[p] = fw.lr_cols[p];
        dw.lr_has[p] = !fw.lr[p].empty();
    }
    L.n_inter = (int)fw.inter.size();
    L.n_obmc = (int)fw.obmc.size();
    L.n_warps = (int)fw.warps.size();
    L.inter = take(sizeof(InterBlk) * std::max(1, L.n_inter));
    // K2 work items: one per 64x64 luma quadrant of a prediction rectangle (a 128x128 block is four CTAs' worth of work, not one)
    L.n_itiles = 0;
    for (const InterBlk& b : fw.inter) L.n_itiles += ((b.w + 63) >> 6) * ((b.h + 63) >> 6);
    L.itiles = take(sizeof(uint32_t) * std::max(1, L.n_itiles));
    L.obmc = take(sizeof(ObmcNb) * std::max(1, L.n_obmc));
    L.warps = take(sizeof(WarpRec) * std::max(1, L.n_warps));
    L.pal = take(std::max<size_t>(4, fw.pal.size()));
    int n_k3 = 0;
    for (const TxRec& r : fw.tx) n_k3 += (r.mode != TXM_INTER) || (r.flags & TXF_II);
    L.n_k3 = n_k3;
    L.k3order = take(sizeof(uint32_t) * std::max(1, n_k3));
    // unit table of the intra kernel: at most one entry per 64x64 luma unit (the exact count is known after the ordering pass)
    const int units_cap = ((fw.fh.mi_cols + 15) >> 4) * ((fw.fh.mi_rows + 15) >> 4);
    L.n_k3units = n_k3 > 0 ? units_cap : 0;
    L.k3units = take(sizeof(K3Unit) * std::max(1, units_cap));
    // frames with intra block copy: unit position -> table index, so that the kernel finds the units a block vector points into
    L.k3upos_on = fw.fh.allow_intrabc && n_k3 > 0;
    L.k3upos = L.k3upos_on ? take(sizeof(int32_t) * std::max(1, units_cap)) : 0;
    L.total = o;
}

static void fill_arena(const FrameWork& fw, DevWork& dw, uint8_t* h) {
    const WorkLayout& L = dw.lay;
    if (L.n_recs) memcpy(h + L.recs, fw.tx.data(), sizeof(TxRec) * L.n_recs);
    if (L.n_coefs) memcpy(h + L.coefs, fw.coefs.data(), sizeof(uint32_t) * L.n_coefs);
    {   // K1 order: coded transform blocks, those of at most 16x16 (itx_small_kernel) first, the 32- and 64-point ones after them
        // Inside each group the blocks are ordered by (transform size, transform type), stable in decode order: every size / type pair
        // is its own unrolled butterfly network, and warps that run side by side on an SM should be fetching the same one (ncu on
        // the 32/64-point kernel in decode order: instruction cache hit rate 65 %, 6.4 warp-cycles of fetch stall per instruction).
        uint32_t* ord = (uint32_t*)(h + L.order);
        static const bool nosort = getenv("AV1R_K1_NOSORT") != nullptr;   // (A/B switch of the experiment)
        constexpr int NKEY = TX_SIZES_ALL * 17;
        uint32_t cnt[2][NKEY + 1];
        memset(cnt, 0, sizeof(cnt));
        auto grp = [&](const TxRec& r) { return (kTxW[r.txsz] <= 16 && kTxH[r.txsz] <= 16) ? 0 : 1; };
        // Largest transforms first: a kernel ends with its last CTAs, so the expensive items must not be the ones that start last.
        static const bool asc = getenv("AV1R_K1_ASC") != nullptr;         // (A/B switch: the ascending order of the first version)
        auto key = [&](const TxRec& r) {
            const int k = (int)r.txsz * 17 + std::min<int>(r.txtp, 16);
            return nosort ? 0 : (asc ? k : NKEY - 1 - k);
        };
        for (int i = 0; i < L.n_recs; i++)
            if (fw.tx[i].eob > 0) cnt[grp(fw.tx[i])][key(fw.tx[i]) + 1]++;
        for (int g = 0; g < 2; g++)
            for (int q = 0; q < NKEY; q++) cnt[g][q + 1] += cnt[g][q];
        const int n_small = (int)cnt[0][NKEY];
        for (int i = 0; i < L.n_recs; i++)
            if (fw.tx[i].eob > 0) {
                const int g = grp(fw.tx[i]);
                ord[(g ? n_small : 0) + cnt[g][key(fw.tx[i])]++] = (uint32_t)i;
            }
        dw.lay.n_order_small = n_small;
    }
    if (L.lf_device) {
        if (L.n_lfblk) memcpy(h + L.lfblk, fw.lf_blocks.data(), sizeof(LfBlk) * L.n_lfblk);
        const bool lf_frame = fw.fh.lf.level[0] || fw.fh.lf.level[1];
        for (int p = 0; p < (fw.mono ? 1 : 3) && lf_frame; p++)
            memcpy(h + L.lftx[p], fw.lf_tx[p].data(), (size_t)fw.plane_w4(p) * fw.plane_h4(p));
    } else {
        for (int p = 0; p < 3; p++)
            if (!fw.lf[p].empty()) memcpy(h + L.lf[p], fw.lf[p].data(), sizeof(LfEdge) * fw.lf[p].size());
    }
    if (!fw.cdef_idx.empty()) memcpy(h + L.cdef_idx, fw.cdef_idx.data(), fw.cdef_idx.size());
    if (!fw.skip_mi.empty()) memcpy(h + L.skip_mi, fw.skip_mi.data(), fw.skip_mi.size());
    for (int p = 0; p < 3; p++)
        if (!fw.lr[p].empty()) memcpy(h + L.lr[p], fw.lr[p].data(), sizeof(LrUnit) * fw.lr[p].size());
    if (L.n_inter) memcpy(h + L.inter, fw.inter.data(), sizeof(InterBlk) * L.n_inter);
    {   // quadrant list: record index | quadrant << 28 (bit 0 of the quadrant = right half, bit 1 = bottom half); the items of
        // small blocks (at most 16x16 luma samples) come first: K2 runs them as two-warp CTAs
        // Inside each of the two groups the items are ordered by *code path* (block shape, warped, compound, masked, OBMC), stable
        // in decode order: K2 is ~100 KB of straight-line filter code, a CTA walks ~40 KB of it, and the small-block kernel spent 5
        // warp-cycles per issued instruction waiting for instruction fetch (ncu stalled_no_instruction) when neighbouring CTAs --
        // resident on an SM together -- each took a different path through it.  Blocks are independent, any order is valid.
        uint32_t* it = (uint32_t*)(h + L.itiles);
        auto path_key = [](const InterBlk& b) -> int {
            const int lw = 31 - __builtin_clz((unsigned)b.w), lh = 31 - __builtin_clz((unsigned)b.h);    // 2 .. 7
            const int warped = b.warp[0] >= 0 || b.warp[1] >= 0;
            const int comp = b.ref[1] >= 0;
            const int masked = comp && (b.comp_type == COMPOUND_WEDGE || b.comp_type == COMPOUND_DIFFWTD);
            const int obmc = (b.obmc_above + b.obmc_left) > 0;
            static const bool off = getenv("AV1R_K2_NOSORT") != nullptr;   // (A/B switch of the experiment)
            if (off) return 0;
            // The most expensive paths get the lowest keys (large, warped, compound blocks first): a 64x64 compound warped
            // quadrant is ~100 us of work for its CTA on a full SM, and a kernel whose heaviest CTAs start last ends on them.
            static const bool asc = getenv("AV1R_K2_ASC") != nullptr;       // (A/B switch: the ascending order of the first version)
            const int k = ((((lw - 2) * 6 + (lh - 2)) * 2 + warped) * 2 + comp) * 2 * 2 + masked * 2 + obmc;   // < 36 * 16
            return asc ? k : 36 * 16 - 1 - k;
        };
        constexpr int NKEY = 36 * 16;
        int k = 0;
        for (int pass = 0; pass < 2; pass++) {   // small blocks (at most 16x16), then the rest
            uint32_t cnt[NKEY + 1] = {0};
            for (int i = 0; i < L.n_inter; i++) {
                const InterBlk& b = fw.inter[i];
                if ((b.w <= 16 && b.h <= 16) != (pass == 0)) continue;
                cnt[path_key(b) + 1] += pass == 0 ? 1 : ((b.w + 63) >> 6) * ((b.h + 63) >> 6);
            }
            for (int q = 0; q < NKEY; q++) cnt[q + 1] += cnt[q];
            const int base = k;
            for (int i = 0; i < L.n_inter; i++) {
                const InterBlk& b = fw.inter[i];
                if ((b.w <= 16 && b.h <= 16) != (pass == 0)) continue;
                uint32_t& pos = cnt[path_key(b)];
                if (pass == 0) {
                    it[base + pos++] = (uint32_t)i;
                    k++;
                } else {
                    const int qx = (b.w + 63) >> 6, qy = (b.h + 63) >> 6;
                    for (int y = 0; y < qy; y++)
                        for (int x = 0; x < qx; x++) {
                            it[base + pos++] = (uint32_t)i | ((uint32_t)(y * 2 + x) << 28);
                            k++;
                        }
                }
            }
            if (pass == 0) dw.lay.n_itiles_small = k;
        }
    }
    if (L.n_obmc) memcpy(h + L.obmc, fw.obmc.data(), sizeof(ObmcNb) * L.n_obmc);
    if (L.n_warps) memcpy(h + L.warps, fw.warps.data(), sizeof(WarpRec) * L.n_warps);
    if (!fw.pal.empty()) memcpy(h + L.pal, fw.pal.data(), fw.pal.size());
    // K3 plan (k3_plan.h): record order, unit table with neighbour dependencies, inter-intra blend links
    dw.lay.n_k3units = k3_plan_build(fw.tx.data(), L.n_recs, L.n_k3, dw.fp.subx, dw.fp.suby, dw.fp.sb128, dw.fp.mi_cols, dw.fp.mi_rows,
                                     (TxRec*)(h + L.recs), (uint32_t*)(h + L.k3order), (K3Unit*)(h + L.k3units), L.k3upos_on ? 1 : 0,
                                     L.k3upos_on ? (int32_t*)(h + L.k3upos) : nullptr);
}

// Pinned staging of one frame's work-lists, filled by the thread that parsed the frame (so the copy into pinned memory and the K3
// ordering run on the parser threads, not on the single thread that issues GPU work) and recycled through a pool.
struct HostArena {
    PinBuf pin;
    DevWork dw;
};
struct HostArenaPool {
    std::mutex m;
    std::vector<std::unique_ptr<HostArena>> free_;
    size_t hw = 0;
    std::unique_ptr<HostArena> get() {
        std::lock_guard<std::mutex> lk(m);
        if (free_.empty()) return std::make_unique<HostArena>();
        auto a = std::move(free_.back());
        free_.pop_back();
        return a;
    }
    void put(std::unique_ptr<HostArena> a) {
        std::lock_guard<std::mutex> lk(m);
        free_.push_back(std::move(a));
    }
    cudaError_t ensure(HostArena& a, size_t n) {
        size_t h;
        {
            std::lock_guard<std::mutex> lk(m);
            h = hw;
        }
        cudaError_t e = a.pin.ensure(n, &h);
        std::lock_guard<std::mutex> lk(m);
        hw = std::max(hw, h);
        return e;
    }
};

// Execution resources of one in-flight frame.
struct FrameSlot {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t fg_ev = nullptr;   // film-grain templates of this slot's frame are ready (prepared on the side stream)
    bool fg_prepared = false;
    PinBuf staging;
    DevBuf arena;        // submit path: the frame's work-lists
    DevBuf residual;
    DevBuf sync;         // K3 done flags + ticket
    DevBuf diffmask;     // K2 difference-weighted compound masks (luma-sized byte plane)
    DevBuf lfscratch;    // device-side deblocking edge classification: per-mi block index + the three LfEdge planes
    DevBuf grain_scratch;
    DevBuf cks_dev;
    PinBuf cks_host;     // 3 x uint64 (+ planes in parity mode)
    PinBuf planes_host;
    bool busy = false;
    std::shared_ptr<void> host_arena;   // pinned staging the queued H2D copy reads from (returned to the pool when the slot is reused)
    // frame buffers touched by the work queued on this slot: kept alive (out of the recycling pool)
    // until the slot's completion event has been waited on
    std::vector<std::shared_ptr<DevFrameBuf>> hold;
};

struct Pending {
    av1r_frame_result res;
    int slot = -1;                    // slot whose events/digests belong to this output (-1: none pending)
    std::shared_ptr<DevFrameBuf> shown;
    bool need_md5 = false;
    int w[3], h[3], bps;
    size_t host_off[3];
};

struct ClipFrame {
    bool show_existing = false;
    int show_slot = -1;
    FrameHdr fh;
    SeqHdr seq;
    int64_t pts = 0;
    DevWork dw;
    uint8_t* host = nullptr;   // pinned work-list arena (exact size): what every replay copies to the device inside the timed region
    DevBuf arena;              // resident mode only: device copy made once
    bool arena_valid = false;
    ~ClipFrame() { if (host) cudaFreeHost(host); }
};

// cudaEvent-based per-stage timer (profile replay only)
struct StageTimer {
    struct Span { int stage; cudaEvent_t a, b; int launches; };
    std::vector<Span> spans;
    cudaEvent_t cur = nullptr;
    void begin(cudaStream_t st) {
        if (cur) cudaEventDestroy(cur);
        cudaEventCreate(&cur);
        cudaEventRecord(cur, st);
    }
    void end(int stage, int launches, cudaStream_t st) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        spans.push_back({stage, cur, e, launches});
        cudaEventCreate(&cur);
        cudaEventRecord(cur, st);
    }
    void collect(av1r_stage_times* out) {
        for (auto& sp : spans) {
            float ms = 0;
            cudaEventElapsedTime(&ms, sp.a, sp.b);
            out->ms[sp.stage] += ms;
            out->launches[sp.stage] += sp.launches;
            cudaEventDestroy(sp.a);
            cudaEventDestroy(sp.b);
        }
        spans.clear();
        if (cur) cudaEventDestroy(cur);
        cur = nullptr;
    }
};

struct EngineImpl {
    av1r_config cfg;
    std::string err;
    StreamParser sp;
    bool opened = false;
    std::vector<cudaStream_t> streams;
    cudaStream_t side = nullptr;          // film-grain template preparation runs here, ahead of the frame it belongs to
    std::vector<std::unique_ptr<FrameSlot>> slots;
    int next_slot = 0, next_stream = 0;
    std::deque<Pending> pending;          // outputs in display order
    // reference slots of the GOP segment currently being issued (av1r_verify_* switches between segments)
    struct RefState { std::shared_ptr<DevFrameBuf> refs[8]; };
    RefState main_refs;
    RefState* rs = &main_refs;
    std::vector<std::shared_ptr<DevFrameBuf>> kept;   // keep_frames handles
    std::vector<std::shared_ptr<DevFrameBuf>> pool;   // recycled frame buffers
    DevBuf k3_prof;                                   // AV1R_K3_PROF: 16 cycle counters of the intra kernel
    DevBuf wedge_master;                              // 6 x 64 x 64 wedge master masks (inter-intra blends in K3)
    DevBuf k3_stuck;                                  // one int: watchdog word of the intra kernel (sticky until the next verify / replay)
    HostArenaPool host_pool;
    std::shared_ptr<void> make_host_arena(const FrameWork& fw, const SeqHdr& seq, std::string& e, int& rc);
    size_t hw_staging = 0, hw_arena = 0, hw_residual = 0, hw_sync = 0, hw_mask = 0, hw_lf = 0;   // high-water marks of the slot buffers
    int k3_ctas = 0;                                  // persistent CTAs per frame of the K3 unit kernel; 0 = default (AV1R_K3_CTAS)
    int k3_progressive = 2;                           // AV1R_K3_PROGRESSIVE: 0 whole-unit hand-over, 1 adaptive, 2 always cell-level (default)
    int k3_inter_mult = 6;                            // inter frames: CTAs = this x the unit wavefront (AV1R_K3_INTER_MULT)
    int k3_intra_run = 0;                             // consecutive frames without inter prediction issued so far
    int k3_since_intra = 1 << 20;                     // frames issued since the last frame without inter prediction (large: none yet)
    int64_t frames_decoded = 0;

    int wait_slot(FrameSlot& s);
    // Slot buffers of one kind grow together: a cudaFree / cudaMalloc pair drains the device, so when one slot needs a bigger buffer
    // every slot gets the new size in that one stall instead of each slot stalling the pipeline later when it meets its first big
    // frame (seen as 2x swings of the clip rate between otherwise identical runs).
    int ensure_slots(DevBuf FrameSlot::*member, FrameSlot& s, size_t n, size_t& hw);
    int finish_pending(Pending& p);
    std::shared_ptr<DevFrameBuf> get_frame(const DevFrameParams& fp);
    StageTimer* tm = nullptr;             // set during av1r_clip_profile
    int run_frame(FrameSlot& s, const DevWork& dw, const uint8_t* d_arena, std::shared_ptr<DevFrameBuf>& out_ref);
    int emit_output(FrameSlot* s, int slot_idx, const std::shared_ptr<DevFrameBuf>& frame, const FrameHdr& fh, const FilmGrainParams& fg,
                    const DevFrameParams& fp, int64_t pts, double parse_ms, bool existing);
    int decode_parsed(ParsedFrame& pf);
    int prepare_work(const FrameWork& fw, DevWork& dw);
    int acquire_slot(int& slot_idx);
    int exec_decoded(int slot_idx, const DevWork& dw, const uint8_t* d_arena, int64_t pts);
    int exec_show_existing(int slot_idx, const FrameHdr& fh, int show_slot, int64_t pts);
};

static std::atomic<long long> g_eprof_ns[20];   // written from the consumer and the parser threads
struct EProfPrinter {
    ~EProfPrinter() {
        if (!getenv("AV1R_PROFILE")) return;
        fprintf(stderr, "[engine prof] acquire %.1f prepare %.1f fill %.1f issue %.1f wait_parse %.1f drain %.1f replay_h2d %.1f replay_exec %.1f ms\n",
                    g_eprof_ns[0] * 1e-6, g_eprof_ns[1] * 1e-6, g_eprof_ns[2] * 1e-6, g_eprof_ns[3] * 1e-6, g_eprof_ns[4] * 1e-6, g_eprof_ns[5] * 1e-6,
                    g_eprof_ns[6] * 1e-6, g_eprof_ns[7] * 1e-6);
            fprintf(stderr, "[engine prof] host issue per stage: getframe %.1f itx %.1f inter %.1f intra %.1f deblock %.1f cdef %.1f lr %.1f emit %.1f ms\n",
                    g_eprof_ns[8] * 1e-6, g_eprof_ns[9] * 1e-6, g_eprof_ns[10] * 1e-6, g_eprof_ns[11] * 1e-6, g_eprof_ns[12] * 1e-6, g_eprof_ns[13] * 1e-6,
                    g_eprof_ns[14] * 1e-6, g_eprof_ns[15] * 1e-6);
    }
} g_eprof_printer;
extern "C" void av1r_debug_engine_prof(double* out20, int reset) {
    for (int i = 0; i < 20; i++) out20[i] = g_eprof_ns[i] * 1e-6;
    if (reset)
        for (auto& v : g_eprof_ns) v = 0;
}
#define EP_T() std::chrono::steady_clock::now()
#define EP_ADD(i, a) g_eprof_ns[i] += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - a).count()

std::shared_ptr<DevFrameBuf> EngineImpl::get_frame(const DevFrameParams& fp) {
    std::shared_ptr<DevFrameBuf> got;
    for (size_t i = 0; i < pool.size() && !got; i++) {
        auto& f = pool[i];
        if (f.use_count() == 1 && f->cw[0] == fp.cw[0] && f->ch[0] == fp.ch[0] && f->cw[1] == fp.cw[1] && f->ch[1] == fp.ch[1] &&
            f->bps == (fp.bd == 8 ? 1 : 2))
            got = f;
    }
    if (!got) {
        const size_t first_new = pool.size();
        if (!alloc_frames(fp, 8, pool, err)) return nullptr;
        got = pool[first_new];
    }
    for (int p = 0; p < 3; p++) { got->w[p] = fp.w[p]; got->h[p] = fp.h[p]; }
    return got;
}

int EngineImpl::ensure_slots(DevBuf FrameSlot::*member, FrameSlot& s, size_t n, size_t& hw) {
    if (n <= (s.*member).cap) return 0;
    const size_t want = std::max(align_up(n + n / 2, 1 << 20), hw);
    hw = want;
    CK(cudaDeviceSynchronize());
    for (auto& sl : slots) {
        DevBuf& b = (*sl).*member;
        if (b.cap >= want) continue;
        if (b.p) cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
        CK(cudaMalloc(&b.p, want));
        b.cap = want;
    }
    return 0;
}

int EngineImpl::wait_slot(FrameSlot& s) {
    if (s.busy) CK(cudaEventSynchronize(s.ev1));
    s.busy = false;
    s.hold.clear();
    s.host_arena.reset();
    return 0;
}

// Enqueue the reconstruction of one frame on slot s.  d_arena = device work-list arena.
int EngineImpl::run_frame(FrameSlot& s, const DevWork& dw, const uint8_t* d_arena, std::shared_ptr<DevFrameBuf>& out_ref) {
    const WorkLayout& L = dw.lay;
    const DevFrameParams& fp = dw.fp;
    cudaStream_t st = s.stream;
    const int since_intra = k3_since_intra;   // frames issued since the last frame without inter prediction
    k3_since_intra = L.n_inter == 0 ? 0 : k3_since_intra + 1;
    auto t_h = EP_T();
    auto recon = get_frame(fp);
    if (!recon) return AV1R_ENOMEM;
    s.hold.push_back(recon);
    EP_ADD(8, t_h);
    t_h = EP_T();
    // residual: unit-major tiles (devframe.h)
    DevResidual res;
    {
        const int units_x = (fp.cw[0] + 63) >> 6, units_y = (fp.ch[0] + 63) >> 6;
        int off = 0;
        for (int p = 0; p < 3; p++) {
            const int sx = p ? fp.subx : 0, sy = p ? fp.suby : 0;
            res.tw_log2[p] = 6 - sx;
            res.th_log2[p] = 6 - sy;
            res.plane_off[p] = off;
            off += 1 << (res.tw_log2[p] + res.th_log2[p]);
        }
        res.units_x = units_x;
        res.unit_elems = off;
        { int e_ = ensure_slots(&FrameSlot::residual, s, (size_t)units_x * units_y * off * sizeof(int16_t), hw_residual); if (e_) return e_; }
        res.base = (int16_t*)s.residual.p;
    }
    const TxRec* d_recs = (const TxRec*)(d_arena + L.recs);
    if (tm) tm->begin(st);
    CK(launch_itx(d_recs, (const uint32_t*)(d_arena + L.order), L.n_order, L.n_order_small, (const uint32_t*)(d_arena + L.coefs), res, fp, st));
    if (tm) tm->end(AV1R_ST_ITX, (L.n_order_small > 0) + (L.n_order > L.n_order_small), st);
    // Deblocking edge classification reads work-lists only: it is queued here, *before* the stream waits for the reference frames,
    // so that it runs while they are still being produced instead of sitting on the frame-to-frame dependency chain.
    const bool lf_run = dw.lf_on && (cfg.inloop_filters & 1);
    LfEdge* lf_edges_dev[3] = {nullptr, nullptr, nullptr};
    if (lf_run && L.lf_device) {
        // block list -> per-mi block index -> LfEdge planes (slot scratch)
        size_t off[4], o = align_up((size_t)fp.mi_cols * fp.mi_rows * sizeof(uint32_t), 256);
        for (int p = 0; p < 3; p++) { off[p] = o; o += align_up((size_t)fp.pw4[p] * fp.ph4[p] * sizeof(LfEdge), 256); }
        { int e_ = ensure_slots(&FrameSlot::lfscratch, s, o, hw_lf); if (e_) return e_; }
        LfClassify lc;
        lc.blks = (const LfBlk*)(d_arena + L.lfblk);
        lc.n_blks = L.n_lfblk;
        lc.mi_blk = (uint32_t*)s.lfscratch.p;
        for (int p = 0; p < 3; p++) {
            lc.lf_tx[p] = d_arena + L.lftx[p];
            lc.edges[p] = (LfEdge*)(s.lfscratch.p + off[p]);
            lc.plane_on[p] = dw.lf_plane_on[p];
            lf_edges_dev[p] = lc.edges[p];
        }
        lc.fp = fp;
        CK(launch_lf_classify(lc, st));
        if (tm) tm->end(AV1R_ST_DEBLOCK, 2, st);
    }
    EP_ADD(9, t_h);
    t_h = EP_T();
    if (L.n_inter > 0) {
        InterLaunch xl;
        xl.blks = (const InterBlk*)(d_arena + L.inter);
        xl.obmc = (const ObmcNb*)(d_arena + L.obmc);
        xl.warps = (const WarpRec*)(d_arena + L.warps);
        xl.n = L.n_inter;
        xl.tiles = (const uint32_t*)(d_arena + L.itiles);
        xl.n_tiles = L.n_itiles;
        xl.n_tiles_small = L.n_itiles_small;
        memset(xl.refs, 0, sizeof(xl.refs));
        memset(xl.ref_w, 0, sizeof(xl.ref_w));
        memset(xl.ref_h, 0, sizeof(xl.ref_h));
        for (int i = 0; i < 8; i++) xl.xscale[i] = xl.yscale[i] = 1 << 14;
        for (int i = 0; i < REFS_PER_FRAME; i++) {
            const int slot = dw.fh.ref_frame_idx[i];
            auto& rf = rs->refs[slot];
            if (!rf) { err = "inter frame references an empty slot"; return AV1R_EBITSTREAM; }
            if (rf->bps != (fp.bd == 8 ? 1 : 2)) {
                err = "reference frame bit depth differs from the current frame";
                return AV1R_EBITSTREAM;
            }
            if (!xl.refs[slot].p[0]) {
                xl.refs[slot] = rf->pl;
                // reference geometry + scale factors (spec 7.11.3.3): 1 << 14 = same size
                for (int p = 0; p < 3; p++) { xl.ref_w[slot][p] = rf->w[p]; xl.ref_h[slot][p] = rf->h[p]; }
                xl.xscale[slot] = (int)((((long long)rf->w[0] << 14) + fp.w[0] / 2) / fp.w[0]);
                xl.yscale[slot] = (int)((((long long)rf->h[0] << 14) + fp.h[0] / 2) / fp.h[0]);
                s.hold.push_back(rf);
                CK(cudaStreamWaitEvent(st, rf->ready, 0));   // produced on another stream
            }
        }
        xl.cur = recon->pl;
        xl.fp = fp;
        xl.mask_pitch = (uint32_t)align_up((size_t)fp.cw[0], 256);
        { int e_ = ensure_slots(&FrameSlot::diffmask, s, (size_t)xl.mask_pitch * fp.ch[0], hw_mask); if (e_) return e_; }
        xl.mask = s.diffmask.p;
        // (running the small-block launch on a second stream beside the large-block one was measured and is *slower*: c3 2425 ->
        // 2130, c1 6950 -> 5900 frames/s -- 64 streams share the 32 hardware queues, and unrelated frames start to wait for each other)
        CK(launch_inter(xl, st));
        if (tm) tm->end(AV1R_ST_INTER, (L.n_itiles_small > 0) + (L.n_itiles > L.n_itiles_small), st);
        CK(launch_inter_residual(d_recs, (const uint32_t*)(d_arena + L.order), L.n_order, recon->pl, res, fp, st));
        if (tm) tm->end(AV1R_ST_INTER, L.n_order > 0, st);
    }
    EP_ADD(10, t_h);
    t_h = EP_T();
    if (L.n_k3 > 0) {
        if (L.n_k3units == -2) { err = "intra block copy vector points at samples that are not reconstructed yet"; return AV1R_EBITSTREAM; }
        if (L.n_k3units < 0) { err = "a 64x64 unit holds more intra records than the intra kernel supports"; return AV1R_ENOSYS; }
        IntraLaunch il;
        il.recs = d_recs;
        il.order = (const uint32_t*)(d_arena + L.k3order);
        il.n = L.n_k3;
        il.units = (const K3Unit*)(d_arena + L.k3units);
        il.n_units = L.n_k3units;
        // persistent CTAs of this frame = width of the unit wavefront (+2 so that the prologue of the next units overlaps): measured
        // on c2 the frame latency does not improve beyond that, while every extra CTA is a spinning resident block that keeps the
        // frames in flight on other streams off the SMs (16 -> 64 CTAs per 1080p frame costs 45 % of the clip throughput)
        {
            const int UX = (fp.mi_cols + 15) >> 4, UY = (fp.mi_rows + 15) >> 4;
            il.ctas = k3_ctas > 0 ? k3_ctas : std::min(UY, (UX + 1) / 2) + 2;
            // inter frames: the few units that hold intra / inter-intra blocks are mostly independent of each other -> one round
            if (k3_ctas <= 0 && L.n_inter > 0) il.ctas = std::min(L.n_k3units, k3_inter_mult * il.ctas);
        }
        il.kind = dw.fh.allow_screen_content_tools ? 2 : (L.n_inter > 0 ? 1 : 0);
        il.load_tile = L.n_inter > 0;
        // sync block: [n_units x u64 progress words][n_units x int unit flags][ticket, stuck flag, pad]
        const size_t sync_bytes = (sizeof(unsigned long long) + sizeof(int)) * (size_t)L.n_k3units + 4 * sizeof(int);
        {   // sized for the frame geometry (one entry per 64x64 unit), not for this frame's unit count
            const size_t cap_bytes = (sizeof(unsigned long long) + sizeof(int)) * (size_t)(((fp.mi_cols + 15) >> 4) * ((fp.mi_rows + 15) >> 4)) + 4 * sizeof(int);
            int e_ = ensure_slots(&FrameSlot::sync, s, std::max(sync_bytes, cap_bytes), hw_sync);
            if (e_) return e_;
        }
        CK(cudaMemsetAsync(s.sync.p, 0, sync_bytes, st));
        il.uprog = (unsigned long long*)s.sync.p;
        il.uflags = (int*)(s.sync.p + sizeof(unsigned long long) * (size_t)L.n_k3units);
        il.ticket = il.uflags + L.n_k3units;
        il.upos = L.k3upos_on ? (const int32_t*)(d_arena + L.k3upos) : nullptr;
        il.stuck = (int*)k3_stuck.p;
        // Cell-level hand-over is the default: 1.6x lower frame latency and (since the wait loop got leaner) also the higher clip
        // rate with 16+ frames side by side (c2: 3220 vs 2960 frames/s).  AV1R_K3_PROGRESSIVE=1 selects the earlier adaptive policy
        // (whole-unit hand-over inside saturated runs of frames nothing predicts from), =0 whole-unit always.
        k3_intra_run = L.n_inter == 0 ? k3_intra_run + 1 : 0;
        bool saturated = false;
        {
            const int n = (int)slots.size();
            if (n > 8) {
                const FrameSlot& old = *slots[((next_slot - 1 - 8) % n + n) % n];
                saturated = old.busy && cudaEventQuery(old.ev1) == cudaErrorNotReady;
            }
        }
        il.progressive = k3_progressive == 1 ? !(k3_intra_run >= 4 && saturated) : (k3_progressive != 0);
        if (il.progressive && k3_ctas <= 0 && L.n_inter == 0) il.ctas = std::min(L.n_k3units, (il.ctas * 7 + 4) / 5);
        // A key frame at the head of a *long* GOP (at least 16 frames were issued since the previous intra frame, or it is the first
        // frame) starts a long dependency chain and is rare: three times as many CTAs take the next units' record / residual loads
        // off the wavefront (4K: 3.8 -> 3.2 ms per key frame; c3 +2.7 %, c1 +2.4 %).  They cost ~80 % more CTA-time per key frame,
        // so runs of intra frames (all-intra content) and short GOPs (the 8-frame segments of the batch, where it measured -7 %)
        // keep the narrow setting, the throughput optimum when many frames share the SMs (profiles/r2_k3_sweep.md).
        static const bool key_wide = getenv("AV1R_K3_KEYWIDE_OFF") == nullptr;   // (A/B switch)
        if (key_wide && k3_ctas <= 0 && L.n_inter == 0 && since_intra >= 16 && !L.k3upos_on)
            il.ctas = std::min(L.n_k3units, 3 * (std::min((fp.mi_rows + 15) >> 4, (((fp.mi_cols + 15) >> 4) + 1) / 2) + 2));
        // block-copy frames list their units in decode order: rows overlap only as far as the CTAs reach ahead in that order
        if (L.k3upos_on && k3_ctas <= 0) il.ctas = std::min(L.n_k3units, std::max(il.ctas, 2 * ((fp.mi_cols + 15) >> 4)));
        il.frame = recon->pl;
        il.res = res;
        il.fp = fp;
        il.wedge_master = wedge_master.p;
        il.pal = d_arena + L.pal;
        il.prof = (unsigned long long*)k3_prof.p;
        CK(launch_intra(il, st));
    }
    if (tm) tm->end(L.n_inter > 0 ? AV1R_ST_INTRA : AV1R_ST_INTRA_FRAME, L.n_k3 > 0 ? 1 : 0, st);
    EP_ADD(11, t_h);
    t_h = EP_T();
    std::shared_ptr<DevFrameBuf> cur = recon;
    if (lf_run) {
        LfLaunch ll;
        ll.frame = cur->pl;
        ll.fp = fp;
        for (int p = 0; p < 3; p++) ll.edges[p] = L.lf_device ? lf_edges_dev[p] : (const LfEdge*)(d_arena + L.lf[p]);
        for (int p = 0; p < 3; p++) ll.plane_on[p] = dw.lf_plane_on[p];
        CK(launch_deblock(ll, st));
        if (tm) tm->end(AV1R_ST_DEBLOCK, 2, st);
    }
    EP_ADD(12, t_h);
    t_h = EP_T();
    if (dw.cdef_on && (cfg.inloop_filters & 2)) {
        auto t_g = EP_T();
        auto dst = get_frame(fp);
        if (!dst) return AV1R_ENOMEM;
        s.hold.push_back(dst);
        if (getenv("AV1R_SLOWDBG")) { double us = std::chrono::duration<double, std::micro>(EP_T() - t_g).count(); if (us > 100) fprintf(stderr, "slow get_frame(cdef) %.0f us pool %zu\n", us, pool.size()); }
        CdefLaunch cl;
        cl.src = cur->pl;
        cl.dst = dst->pl;
        cl.cdef_idx = (const int8_t*)(d_arena + L.cdef_idx);
        cl.skip_mi = d_arena + L.skip_mi;
        cl.fp = fp;
        t_g = EP_T();
        CK(launch_cdef(cl, st));
        if (getenv("AV1R_SLOWDBG")) { double us = std::chrono::duration<double, std::micro>(EP_T() - t_g).count(); if (us > 100) fprintf(stderr, "slow launch_cdef %.0f us\n", us); }
        if (tm) tm->end(AV1R_ST_CDEF, 1, st);
        cur = dst;
    }
    EP_ADD(13, t_h);
    t_h = EP_T();
    std::shared_ptr<DevFrameBuf> deblocked = recon;
    const DevFrameParams& fpu = dw.fp_up;
    if (dw.fh.use_superres) {   // K6: stretch the CDEF output (and, for the stripe boundaries of loop restoration, the deblocked frame)
        SuperresLaunch sl;
        sl.planes = fp.mono ? 1 : 3;
        sl.bd = fp.bd;
        for (int p = 0; p < 3; p++) {
            const int sx = p ? fp.subx : 0;
            const int down_w = (dw.fh.frame_width + sx) >> sx, up_w = (dw.fh.upscaled_width + sx) >> sx;
            const int step = ((down_w << 14) + (up_w / 2)) / up_w;
            const int e = up_w * step - (down_w << 14);
            sl.step_x[p] = step;
            sl.initial_subpel_x[p] = ((-((up_w - down_w) << 13) + up_w / 2) / up_w + (1 << 7) - e / 2) & ((1 << 14) - 1);
            sl.up_w[p] = up_w;
            sl.h[p] = fpu.h[p];
            sl.src_cw[p] = fp.cw[p];
        }
        const bool need_db = dw.lr_on && (cfg.inloop_filters & 4) && cur != recon;
        auto up = get_frame(fpu);
        if (!up) return AV1R_ENOMEM;
        s.hold.push_back(up);
        sl.src = cur->pl;
        sl.dst = up->pl;
        CK(launch_superres(sl, st));
        if (need_db) {
            auto upd = get_frame(fpu);
            if (!upd) return AV1R_ENOMEM;
            s.hold.push_back(upd);
            sl.src = recon->pl;
            sl.dst = upd->pl;
            CK(launch_superres(sl, st));
            deblocked = upd;
        } else {
            deblocked = up;
        }
        if (tm) tm->end(AV1R_ST_SUPERRES, need_db ? 2 : 1, st);
        cur = up;
    }
    if (dw.lr_on && (cfg.inloop_filters & 4)) {
        auto t_g = EP_T();
        auto dst = get_frame(fpu);
        if (!dst) return AV1R_ENOMEM;
        s.hold.push_back(dst);
        if (getenv("AV1R_SLOWDBG")) { double us = std::chrono::duration<double, std::micro>(EP_T() - t_g).count(); if (us > 100) fprintf(stderr, "slow get_frame(lr) %.0f us pool %zu\n", us, pool.size()); }
        LrLaunch ll;
        ll.cdef = cur->pl;
        ll.deblocked = deblocked->pl;
        ll.dst = dst->pl;
        for (int p = 0; p < 3; p++) {
            ll.units[p] = (const LrUnitDev*)(d_arena + L.lr[p]);
            ll.lr_type[p] = dw.lr_has[p] ? dw.fh.lr_type[p] : 0;
            ll.unit_size[p] = dw.fh.lr_size[p];
            ll.unit_rows[p] = dw.lr_rows[p];
            ll.unit_cols[p] = dw.lr_cols[p];
        }
        ll.fp = fpu;
        t_g = EP_T();
        CK(launch_lr(ll, st));
        if (getenv("AV1R_SLOWDBG")) { double us = std::chrono::duration<double, std::micro>(EP_T() - t_g).count(); if (us > 100) fprintf(stderr, "slow launch_lr %.0f us\n", us); }
        if (tm) tm->end(AV1R_ST_LR, 1, st);
        cur = dst;
    }
    CK(cudaEventRecord(cur->ready, st));
    out_ref = cur;
    EP_ADD(14, t_h);
    return 0;
}

int EngineImpl::emit_output(FrameSlot* s, int slot_idx, const std::shared_ptr<DevFrameBuf>& frame, const FrameHdr& fh, const FilmGrainParams& fg,
                            const DevFrameParams& fp, int64_t pts, double parse_ms, bool existing) {
    cudaStream_t st = s->stream;
    std::shared_ptr<DevFrameBuf> shown = frame;
    s->hold.push_back(frame);
    if (existing) CK(cudaStreamWaitEvent(st, frame->ready, 0));
    if (tm) tm->begin(st);
    if (cfg.apply_grain && fg.apply_grain) {
        auto disp = get_frame(fp);
        if (!disp) return AV1R_ENOMEM;
        s->hold.push_back(disp);
        CK(s->grain_scratch.ensure(av1r_film_grain_scratch_bytes()));
        const void* src[3] = {frame->pl.p[0], frame->pl.p[1], frame->pl.p[2]};
        void* dst[3] = {disp->pl.p[0], disp->pl.p[1], disp->pl.p[2]};
        size_t sp_[3] = {frame->pl.pitch[0], frame->pl.pitch[1], frame->pl.pitch[2]};
        size_t dp_[3] = {disp->pl.pitch[0], disp->pl.pitch[1], disp->pl.pitch[2]};
        const int mc_id = sp.hp.seq.matrix_coefficients == 0;
        int rc;
        if (s->fg_prepared) {   // templates were prepared on the side stream while the frame was being reconstructed
            CK(cudaStreamWaitEvent(st, s->fg_ev, 0));
            s->fg_prepared = false;
        } else {
            rc = fg_launch_prepare((const av1r_film_grain_params*)&fg, fp.bd, fp.w[0], fp.h[0], fp.subx, fp.suby, fp.mono, mc_id, s->grain_scratch.p, st);
            if (rc) { err = av1r_stage_last_error(); return rc; }
        }
        rc = fg_launch_apply((const av1r_film_grain_params*)&fg, fp.bd, fp.w[0], fp.h[0], fp.subx, fp.suby, fp.mono, mc_id, src, sp_, dst, dp_,
                             s->grain_scratch.p, st);
        if (rc) {
            err = av1r_stage_last_error();
            return rc;
        }
        shown = disp;
        if (tm) tm->end(AV1R_ST_GRAIN, 2, st);
    }
    CK(s->cks_dev.ensure(64));
    CK(s->cks_host.ensure(64));
    const int np = fp.mono ? 1 : 3;
    {
        const void* src3[3] = {shown->pl.p[0], shown->pl.p[1], shown->pl.p[2]};
        const size_t pitch3[3] = {shown->pl.pitch[0], shown->pl.pitch[1], shown->pl.pitch[2]};
        const int w3[3] = {fp.w[0], fp.w[1], fp.w[2]}, h3[3] = {fp.h[0], fp.h[1], fp.h[2]};
        CK(launch_frame_checksum(src3, pitch3, w3, h3, np, fp.bd, (uint64_t*)s->cks_dev.p, st));
    }
    if (tm) tm->end(AV1R_ST_DIGEST, 1, st);
    CK(cudaMemcpyAsync(s->cks_host.p, s->cks_dev.p, 24, cudaMemcpyDeviceToHost, st));
    // watchdog word of the intra kernel: engine-wide and sticky, so a stuck *hidden* frame (ALTREF, or one shown later through
    // show_existing_frame) is reported with the next output instead of being lost
    CK(cudaMemcpyAsync(s->cks_host.p + 24, k3_stuck.p, 4, cudaMemcpyDeviceToHost, st));
    Pending pd;
    memset(&pd.res, 0, sizeof(pd.res));
    pd.res.struct_size = sizeof(pd.res);
    pd.res.pts = pts;
    pd.res.w = fp.w[0];
    pd.res.h = fp.h[0];
    pd.res.bpc = fp.bd;
    pd.res.layout = fp.mono ? 0 : (fp.subx && fp.suby ? 1 : (fp.subx ? 2 : 3));
    pd.res.frame_type = fh.frame_type;
    pd.res.shown_existing = existing;
    pd.res.host_parse_ms = (float)parse_ms;
    pd.res.frame_handle = -1;
    pd.slot = slot_idx;
    pd.shown = shown;
    pd.bps = fp.bd == 8 ? 1 : 2;
    for (int p = 0; p < 3; p++) { pd.w[p] = fp.w[p]; pd.h[p] = fp.h[p]; }
    pd.need_md5 = cfg.parity_md5 != 0;
    if (pd.need_md5) {
        size_t total = 0;
        for (int p = 0; p < np; p++) { pd.host_off[p] = total; total += (size_t)fp.w[p] * pd.bps * fp.h[p]; }
        CK(s->planes_host.ensure(total));
        for (int p = 0; p < np; p++)
            CK(cudaMemcpy2DAsync(s->planes_host.p + pd.host_off[p], (size_t)fp.w[p] * pd.bps, shown->pl.p[p], shown->pl.pitch[p],
                                 (size_t)fp.w[p] * pd.bps, fp.h[p], cudaMemcpyDeviceToHost, st));
    }
    if (cfg.keep_frames) {
        kept.push_back(shown);
        pd.res.frame_handle = (int64_t)kept.size() - 1;
    }
    pending.push_back(pd);
    return 0;
}

int EngineImpl::acquire_slot(int& slot_idx) {
    slot_idx = next_slot;
    FrameSlot& s = *slots[slot_idx];
    next_slot = (next_slot + 1) % (int)slots.size();
    // a slot can only be reused once the outputs that still read its pinned buffers were finalised
    for (auto& p : pending)
        if (p.slot == slot_idx) {
            int rc = finish_pending(p);
            if (rc) return rc;
        }
    int rc = wait_slot(s);
    if (rc) return rc;
    s.stream = streams[next_stream];
    next_stream = (next_stream + 1) % (int)streams.size();
    return 0;
}

static void upscaled_params(const FrameHdr& fh, const DevFrameParams& fp, DevFrameParams& up) {
    up = fp;
    if (!fh.use_superres) return;
    up.mi_cols = 2 * ((fh.upscaled_width + 7) >> 3);
    for (int p = 0; p < 3; p++) {
        const int sx = p ? fp.subx : 0;
        up.w[p] = (fh.upscaled_width + sx) >> sx;
        up.cw[p] = (up.mi_cols * 4) >> sx;
        up.pw4[p] = (up.mi_cols + sx) >> sx;
    }
}

static int prepare_work_seq(const SeqHdr& seq, const FrameWork& fw, DevWork& dw, std::string& err) {
    fill_params(seq, fw, dw.fp);
    upscaled_params(fw.fh, dw.fp, dw.fp_up);
    dw.fh = fw.fh;
    dw.lf_on = fw.fh.lf.level[0] || fw.fh.lf.level[1];
    dw.lf_plane_on[0] = dw.lf_on;
    dw.lf_plane_on[1] = dw.lf_on && fw.fh.lf.level[2];
    dw.lf_plane_on[2] = dw.lf_on && fw.fh.lf.level[3];
    dw.cdef_on = fw.fh.enable_cdef_frame;
    dw.lr_on = fw.fh.uses_lr;
    dw.parse_ms = fw.parse_ms;
    dw.coded_samples = fw.coded_samples;
    dw.coef_tokens = fw.coefs.size();
    plan_layout(fw, dw);
    return 0;
}

int EngineImpl::prepare_work(const FrameWork& fw, DevWork& dw) { return prepare_work_seq(sp.hp.seq, fw, dw, err); }

// Runs on the thread that parsed the frame: layout + copy of the work-lists into pinned memory.
std::shared_ptr<void> EngineImpl::make_host_arena(const FrameWork& fw, const SeqHdr& seq, std::string& e, int& rc) {
    auto a = host_pool.get();
    rc = prepare_work_seq(seq, fw, a->dw, e);
    if (rc) {
        host_pool.put(std::move(a));
        return nullptr;
    }
    if (host_pool.ensure(*a, a->dw.lay.total) != cudaSuccess) {
        e = "cudaHostAlloc(staging) failed";
        rc = AV1R_ENOMEM;
        host_pool.put(std::move(a));
        return nullptr;
    }
    fill_arena(fw, a->dw, a->pin.p);
    HostArenaPool* pool = &host_pool;
    return std::shared_ptr<void>(a.release(), [pool](void* p) { pool->put(std::unique_ptr<HostArena>((HostArena*)p)); });
}

int EngineImpl::exec_show_existing(int slot_idx, const FrameHdr& fh, int show_slot, int64_t pts) {
    FrameSlot& s = *slots[slot_idx];
    auto f = rs->refs[show_slot];
    if (!f) { err = "show_existing_frame of an empty slot"; return AV1R_EBITSTREAM; }
    DevFrameParams fp;
    FrameWork dummy;
    dummy.fh = fh;
    fill_params(sp.hp.seq, dummy, fp);
    {
        DevFrameParams up;
        upscaled_params(fh, fp, up);
        fp = up;
    }
    CK(cudaEventRecord(s.ev0, s.stream));
    int rc = emit_output(&s, slot_idx, f, fh, fh.fg, fp, pts, 0, true);
    if (rc) return rc;
    CK(cudaEventRecord(s.ev1, s.stream));
    s.busy = true;
    if (fh.frame_type == KEY_FRAME)
        for (int i = 0; i < 8; i++) rs->refs[i] = f;
    return 0;
}

// d_arena must already be (or be queued to become) valid on the slot's stream.
int EngineImpl::exec_decoded(int slot_idx, const DevWork& dw, const uint8_t* d_arena, int64_t pts) {
    FrameSlot& s = *slots[slot_idx];
    std::shared_ptr<DevFrameBuf> out;
    s.fg_prepared = false;
    if (!tm && dw.fh.show_frame && cfg.apply_grain && dw.fh.fg.apply_grain) {
        // grain templates depend on the header only: prepare them now on the side stream (a single-CTA, ~170 us latency-bound kernel)
        // instead of at the end of the frame's own chain.  (The stage profile keeps it inline so that its time is attributed.)
        CK(s.grain_scratch.ensure(av1r_film_grain_scratch_bytes()));
        const DevFrameParams& fpu = dw.fp_up;
        int grc = fg_launch_prepare((const av1r_film_grain_params*)&dw.fh.fg, fpu.bd, fpu.w[0], fpu.h[0], fpu.subx, fpu.suby, fpu.mono,
                                    sp.hp.seq.matrix_coefficients == 0, s.grain_scratch.p, side);
        if (grc) { err = av1r_stage_last_error(); return grc; }
        CK(cudaEventRecord(s.fg_ev, side));
        s.fg_prepared = true;
    }
    int rc = run_frame(s, dw, d_arena, out);
    if (rc) return rc;
    for (int i = 0; i < 8; i++)
        if ((dw.fh.refresh_frame_flags >> i) & 1) rs->refs[i] = out;
    if (dw.fh.show_frame) {
        auto t_e = EP_T();
        rc = emit_output(&s, slot_idx, out, dw.fh, dw.fh.fg, dw.fp_up, pts, dw.parse_ms, false);
        EP_ADD(15, t_e);
        if (rc) return rc;
    }
    CK(cudaEventRecord(s.ev1, s.stream));
    s.busy = true;
    frames_decoded++;
    return 0;
}


int EngineImpl::decode_parsed(ParsedFrame& pf) {
    int slot_idx;
    auto ta = EP_T();
    int rc = acquire_slot(slot_idx);
    EP_ADD(0, ta);
    if (rc) return rc;
    FrameSlot& s = *slots[slot_idx];
    sp.hp.seq = pf.seq;   // the frame's own sequence header (segments of different files alternate on one engine)
    if (pf.show_existing_slot >= 0) return exec_show_existing(slot_idx, pf.fh, pf.show_existing_slot, pf.pts);
    const FrameWork& fw = *pf.fw;
    if (pf.host) {   // staged by the parser thread
        HostArena& ha = *(HostArena*)pf.host.get();
        auto ti = EP_T();
        auto t1 = EP_T();
        { int e_ = ensure_slots(&FrameSlot::arena, s, ha.dw.lay.total, hw_arena); if (e_) return e_; }
        EP_ADD(16, t1);
        s.host_arena = pf.host;
        t1 = EP_T();
        CK(cudaEventRecord(s.ev0, s.stream));
        CK(cudaMemcpyAsync(s.arena.p, ha.pin.p, ha.dw.lay.total, cudaMemcpyHostToDevice, s.stream));
        EP_ADD(17, t1);
        rc = exec_decoded(slot_idx, ha.dw, s.arena.p, pf.pts);
        EP_ADD(3, ti);
        return rc;
    }
    DevWork dw;
    auto tp = EP_T();
    rc = prepare_work(fw, dw);
    EP_ADD(1, tp);
    if (rc) return rc;
    auto tf = EP_T();
    CK(s.staging.ensure(dw.lay.total, &hw_staging));
    { int e_ = ensure_slots(&FrameSlot::arena, s, dw.lay.total, hw_arena); if (e_) return e_; }
    fill_arena(fw, dw, s.staging.p);
    EP_ADD(2, tf);
    auto ti = EP_T();
    CK(cudaEventRecord(s.ev0, s.stream));
    CK(cudaMemcpyAsync(s.arena.p, s.staging.p, dw.lay.total, cudaMemcpyHostToDevice, s.stream));
    rc = exec_decoded(slot_idx, dw, s.arena.p, pf.pts);
    EP_ADD(3, ti);
    return rc;
}

int EngineImpl::finish_pending(Pending& p) {
    if (p.slot < 0) return 0;
    FrameSlot& s = *slots[p.slot];
    CK(cudaEventSynchronize(s.ev1));
    float ms = 0;
    cudaEventElapsedTime(&ms, s.ev0, s.ev1);
    p.res.device_ms = ms;
    memcpy(p.res.checksum, s.cks_host.p, 24);
    {   // the intra kernel's watchdog fired: a wait on a work-list dependency never completed -- report it, do not trust the digest
        int stuck = 0;
        memcpy(&stuck, s.cks_host.p + 24, 4);
        if (stuck) {
            p.res.status = AV1R_EIO;
            err = "intra kernel watchdog: a dependency wait made no progress (inconsistent work-list)";
        }
    }
    if (p.need_md5) {
        const int np = p.res.layout == 0 ? 1 : 3;
        for (int pl = 0; pl < np; pl++) {
            Md5 m;
            m.update(s.planes_host.p + p.host_off[pl], (size_t)p.w[pl] * p.bps * p.h[pl]);
            m.final(p.res.md5[pl]);
        }
    }
    s.busy = false;
    p.slot = -1;
    p.shown.reset();
    return 0;
}

// ------------------------------------------------------------------------------------------------
Engine::Engine() : impl_(new EngineImpl()) {}
Engine::~Engine() {
    if (impl_->opened) {
        cudaSetDevice(impl_->cfg.device);
        cudaDeviceSynchronize();
        if (impl_->k3_prof.p) {   // AV1R_K3_PROF: cycle totals of the intra kernel's phases
            unsigned long long c[16];
            cudaMemcpy(c, impl_->k3_prof.p, sizeof(c), cudaMemcpyDeviceToHost);
            static const char* names[16] = {"ticket", "prologue", "unit_wait", "halo", "res_wait", "dataflow", "writeback", "", "rec_fetch", "rec_wait",
                                            "rec_predict", "warp_idle", "records", "units", "", ""};
            fprintf(stderr, "[av1r k3 prof]");
            for (int i = 0; i < 14; i++)
                if (names[i][0]) fprintf(stderr, " %s=%llu", names[i], c[i]);
            fprintf(stderr, "\n");
        }
        for (auto& s : impl_->slots) {
            if (s->ev0) cudaEventDestroy(s->ev0);
            if (s->ev1) cudaEventDestroy(s->ev1);
            if (s->fg_ev) cudaEventDestroy(s->fg_ev);
        }
        impl_->slots.clear();
        impl_->pending.clear();
        impl_->kept.clear();
        impl_->pool.clear();
        for (auto& r : impl_->main_refs.refs) r.reset();
        for (auto st : impl_->streams) cudaStreamDestroy(st);
        if (impl_->side) cudaStreamDestroy(impl_->side);
    }
    delete impl_;
}
const std::string& Engine::error() const { return impl_->err; }

int Engine::open(const av1r_config& cfg) {
    EngineImpl& E = *impl_;
    std::string& err = E.err;
    E.cfg = cfg;
    if (E.cfg.streams <= 0) E.cfg.streams = 2;
    if (E.cfg.frames_in_flight <= 0) E.cfg.frames_in_flight = 8;
    // frame-level overlap uses many streams: give them their own hardware queues (default is 8, which caps
    // the number of concurrently running wavefront kernels).  Only effective before the context exists.
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        err = "no CUDA device available (this engine has no CPU fallback)";
        return AV1R_EIO;
    }
    if (cfg.device < 0 || cfg.device >= ndev) {
        err = "bad device ordinal";
        return AV1R_EINVAL;
    }
    CK(cudaSetDevice(cfg.device));
    E.streams.resize(E.cfg.streams);
    for (auto& st : E.streams) CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&E.side, cudaStreamNonBlocking));
    for (int i = 0; i < E.cfg.frames_in_flight; i++) {
        auto s = std::make_unique<FrameSlot>();
        CK(cudaEventCreate(&s->ev0));
        CK(cudaEventCreate(&s->ev1));
        CK(cudaEventCreateWithFlags(&s->fg_ev, cudaEventDisableTiming));
        E.slots.push_back(std::move(s));
    }
    if (const char* e = getenv("AV1R_K3_CTAS")) E.k3_ctas = std::max(0, atoi(e));
    if (const char* e = getenv("AV1R_K3_INTER_MULT")) E.k3_inter_mult = std::max(1, atoi(e));
    if (const char* e = getenv("AV1R_K3_PROGRESSIVE")) E.k3_progressive = std::min(2, std::max(0, atoi(e)));
    if (getenv("AV1R_K3_PROF")) {
        CK(E.k3_prof.ensure(16 * sizeof(unsigned long long)));
        CK(cudaMemset(E.k3_prof.p, 0, 16 * sizeof(unsigned long long)));
    }
    CK(E.k3_stuck.ensure(256));
    CK(cudaMemset(E.k3_stuck.p, 0, 256));
    E.sp.host_lf_edges = false;   // (streaming API parser) deblocking edges are classified on the device
    CK(E.wedge_master.ensure(6 * 64 * 64));
    CK(inter_copy_wedge_master(E.wedge_master.p, E.streams[0]));
    CK(cudaStreamSynchronize(E.streams[0]));
    E.opened = true;
    return 0;
}

int Engine::submit_tu(const uint8_t* data, size_t len, int64_t pts) {
    EngineImpl& E = *impl_;
    cudaSetDevice(E.cfg.device);
    std::vector<ParsedFrame> pfs;
    int prc = E.sp.parse_tu(data, len, pts, pfs);
    for (ParsedFrame& pf : pfs) {
        int rc = E.decode_parsed(pf);
        if (rc) return rc;
    }
    if (prc) {
        E.err = E.sp.err;
        return prc;
    }
    return 0;
}

int Engine::collect(av1r_frame_result* out, int cap, int* n) {
    EngineImpl& E = *impl_;
    cudaSetDevice(E.cfg.device);
    int k = 0;
    while (k < cap && !E.pending.empty()) {
        Pending& p = E.pending.front();
        if (p.slot >= 0) {
            FrameSlot& s = *E.slots[p.slot];
            if (cudaEventQuery(s.ev1) == cudaErrorNotReady) break;
            int rc = E.finish_pending(p);
            if (rc) { *n = k; return rc; }
        }
        out[k++] = p.res;
        E.pending.pop_front();
    }
    *n = k;
    return 0;
}

int Engine::flush() {
    EngineImpl& E = *impl_;
    cudaSetDevice(E.cfg.device);
    for (auto& p : E.pending) {
        int rc = E.finish_pending(p);
        if (rc) return rc;
    }
    if (cudaDeviceSynchronize() != cudaSuccess) {
        E.err = "device synchronize failed";
        return AV1R_EIO;
    }
    return 0;
}

int Engine::copy_frame(int64_t handle, int plane, void* dst, size_t dst_stride) {
    EngineImpl& E = *impl_;
    if (handle < 0 || handle >= (int64_t)E.kept.size() || !E.kept[handle] || plane < 0 || plane > 2) return AV1R_EINVAL;
    cudaSetDevice(E.cfg.device);
    auto& f = E.kept[handle];
    // visible size is not stored in the buffer: copy the coded width, callers clip
    cudaError_t e = cudaMemcpy2D(dst, dst_stride, f->pl.p[plane], f->pl.pitch[plane], std::min(dst_stride, (size_t)f->cw[plane] * f->bps),
                                 f->ch[plane], cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) {
        E.err = cudaGetErrorString(e);
        return AV1R_EIO;
    }
    return 0;
}

int Engine::release_frame(int64_t handle) {
    EngineImpl& E = *impl_;
    if (handle < 0 || handle >= (int64_t)E.kept.size()) return AV1R_EINVAL;
    E.kept[handle].reset();
    return 0;
}

int Engine::verify_file(const char* path, const av1r_config* cfg, av1r_report* out) {
    memset(out, 0, sizeof(*out));
    out->struct_size = sizeof(*out);
    out->first_bad_frame = -1;
    FILE* f = fopen(path, "rb");
    if (!f) {
        out->status = AV1R_ENOENT;
        snprintf(out->message, sizeof(out->message), "cannot open %s", path);
        return out->status;
    }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> buf(n > 0 ? n : 0);
    size_t got = n > 0 ? fread(buf.data(), 1, n, f) : 0;
    fclose(f);
    if ((long)got != n) {
        out->status = AV1R_EIO;
        snprintf(out->message, sizeof(out->message), "short read on %s", path);
        return out->status;
    }
    return verify_buffer(buf.data(), buf.size(), cfg, out, nullptr, 0);
}

// One helper thread with a FIFO of closures: the staging lane of a segment parser (merge of the tile lists, deblocking edges, copy
// into pinned memory run here while the parser thread is already in the next frame).  Persistent for the life of the worker: a
// std::async per temporal unit cost a thread creation plus a first CUDA call on a fresh thread per frame.
class SerialWorker {
public:
    explicit SerialWorker(int device) {
        th_ = std::thread([this, device] {
            cudaSetDevice(device);
            std::unique_lock<std::mutex> lk(m_);
            while (true) {
                cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;   // stop requested and nothing left
                auto f = std::move(q_.front());
                q_.pop_front();
                busy_ = true;
                lk.unlock();
                f();
                lk.lock();
                busy_ = false;
                if (q_.empty()) idle_.notify_all();
            }
        });
    }
    ~SerialWorker() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        th_.join();
    }
    void post(std::function<void()> f) {
        {
            std::lock_guard<std::mutex> lk(m_);
            q_.push_back(std::move(f));
        }
        cv_.notify_all();
    }
    void drain() {
        std::unique_lock<std::mutex> lk(m_);
        idle_.wait(lk, [&] { return q_.empty() && !busy_; });
    }

private:
    std::thread th_;
    std::mutex m_;
    std::condition_variable cv_, idle_;
    std::deque<std::function<void()>> q_;
    bool stop_ = false, busy_ = false;
};

// ---- verification of whole containers: one file on one GPU (av1r_verify_*), or a batch of files over several GPUs ------------
// One key-frame-delimited GOP segment: parsed by one host thread, issued to the GPU in order.
struct Segment {
    size_t tu0 = 0, tu1 = 0;                       // temporal units [tu0, tu1)
    std::vector<std::vector<ParsedFrame>> parsed;  // per TU
    std::vector<int> rc;
    std::vector<std::string> errs;
    std::mutex m;
    std::condition_variable cv;
    size_t n_done = 0;                              // TUs parsed so far
};

// Pre-scan of one container: sequence header, temporal units, where the independently decodable GOP segments start (temporal
// units whose first frame is a shown KEY_FRAME -- the only safe cut) and how many frames are shown before each unit.
// Header scan of one temporal unit (no tile data is read): keeps `scan`'s reference state current, reports whether the unit starts
// an independently decodable GOP segment (its first frame is a shown KEY_FRAME -- the only safe cut) and how many
// frames it shows.
static int scan_temporal_unit(HeaderParser& scan, const uint8_t* p, size_t n, bool* seg_start, int* shown_out, std::string& msg) {
    std::vector<ObuUnit> obus;
    if (!scan.split_obus(p, n, obus)) { msg = scan.error; return AV1R_EBITSTREAM; }
    bool first_frame = true, frame_open = false;   // frame_open: a frame header was seen and not all of its tiles yet
    int shown = 0;
    *seg_start = false;
    FrameHdr open_fh;
    auto tiles_after = [&](const uint8_t* q, size_t m) {   // tile group header at q: is the frame complete after this group?
        BitReader tb(q, m);
        TileGroupInfo tg;
        if (!scan.parse_tile_group_header(tb, open_fh, tg)) return true;   // the parser proper reports the error
        return tg.tg_end + 1 >= open_fh.tile_cols * open_fh.tile_rows;
    };
    for (const ObuUnit& u : obus) {
        if (u.type == OBU_SEQUENCE_HEADER) {
            if (!scan.parse_sequence_header(u.data, u.size)) { msg = scan.error; return AV1R_EBITSTREAM; }
        } else if (u.type == OBU_FRAME || (u.type == OBU_FRAME_HEADER && !frame_open)) {   // (a header while a frame is open is a copy)
            BitReader br(u.data, u.size);
            FrameHdr fh;
            if (!scan.parse_frame_header(br, fh, u.temporal_id, u.spatial_id)) { msg = scan.error; return AV1R_EBITSTREAM; }
            if (first_frame && !fh.show_existing_frame && fh.frame_type == KEY_FRAME && fh.show_frame) *seg_start = true;
            shown += fh.show_existing_frame || fh.show_frame;
            if (!fh.show_existing_frame) scan.reference_update(fh);
            else if (fh.frame_type == KEY_FRAME) {
                RefHdrState r = scan.refs[fh.frame_to_show_map_idx];
                for (auto& x : scan.refs) x = r;
            }
            first_frame = false;
            open_fh = fh;
            frame_open = !fh.show_existing_frame;
            if (u.type == OBU_FRAME) {
                br.byte_align();
                const size_t off = br.byte_pos();
                frame_open = off < u.size ? !tiles_after(u.data + off, u.size - off) : false;
            }
        } else if (u.type == OBU_TILE_GROUP && frame_open) {
            frame_open = !tiles_after(u.data, u.size);
        }
    }
    *shown_out = shown;
    return 0;
}

// Pre-scan of one container: sequence header, temporal units, where the GOP segments start and how many frames are shown before
// each unit.
int VerifyFile::prescan(std::string& msg) {
    std::string derr;
    if (!demux_buffer(data, len, dm, derr)) { msg = derr; return AV1R_EBITSTREAM; }
    if (!dm.config_obus.empty()) {
        std::vector<ObuUnit> obus;
        if (scan.split_obus(dm.config_obus.data(), dm.config_obus.size(), obus))
            for (auto& u : obus)
                if (u.type == OBU_SEQUENCE_HEADER) scan.parse_sequence_header(u.data, u.size);
    }
    starts.clear();
    start_seq.clear();
    frame_base.assign(dm.tus.size() + 1, 0);
    for (size_t i = 0; i < dm.tus.size(); i++) {
        bool seg = false;
        int shown = 0;
        const int rc = scan_temporal_unit(scan, data + dm.tus[i].offset, dm.tus[i].size, &seg, &shown, msg);
        if (rc) return rc;
        if (seg || i == 0) {
            starts.push_back(i);
            start_seq.push_back(scan.seq);
        }
        frame_base[i + 1] = frame_base[i] + shown;
    }
    if (!scan.seq.valid) { msg = "no sequence header"; return AV1R_EBITSTREAM; }
    if (starts.empty()) { msg = "no temporal units"; return AV1R_EBITSTREAM; }
    return 0;
}

void VerifyFile::init_report() {
    memset(rep, 0, sizeof(*rep));
    rep->struct_size = sizeof(*rep);
    rep->first_bad_frame = -1;
}

void VerifyFile::fail(int rc, int64_t frame, const std::string& msg) {
    std::lock_guard<std::mutex> lk(m);
    if (rep->status && rep->first_bad_frame >= 0 && (frame < 0 || frame >= rep->first_bad_frame)) return;   // keep the earliest
    rep->status = rc;
    rep->first_bad_frame = frame;
    snprintf(rep->message, sizeof(rep->message), "%s", msg.c_str());
}

int Engine::verify_buffer(const uint8_t* data, size_t len, const av1r_config* cfg, av1r_report* out, uint64_t* digests, int64_t cap_frames) {
    av1r_config c;
    av1r_default_config(&c);
    if (cfg) {
        size_t n = cfg->struct_size && cfg->struct_size < sizeof(c) ? cfg->struct_size : sizeof(c);
        memcpy(&c, cfg, n);
        c.struct_size = sizeof(c);
    }
    if (c.streams <= 2) c.streams = 16;
    if (c.frames_in_flight <= 8) c.frames_in_flight = 32;
    Engine eng;
    int rc = eng.open(c);
    if (rc) {
        memset(out, 0, sizeof(*out));
        out->struct_size = sizeof(*out);
        out->first_bad_frame = -1;
        out->status = rc;
        snprintf(out->message, sizeof(out->message), "%s", eng.error().c_str());
        return rc;
    }
    return eng.verify(data, len, out, digests, cap_frames);
}

// Verify a whole container with this (already open, reusable) engine.
int Engine::verify(const uint8_t* data, size_t len, av1r_report* out, uint64_t* digests, int64_t cap_frames) {
    VerifyFile vf;
    vf.data = data;
    vf.len = len;
    vf.rep = out;
    vf.digests = digests;
    vf.cap_frames = cap_frames;
    vf.init_report();
    auto t0 = std::chrono::steady_clock::now();
    std::string msg;
    int rc = vf.prescan(msg);
    if (rc) {
        out->status = rc;
        snprintf(out->message, sizeof(out->message), "%s", msg.c_str());
        return rc;
    }
    std::vector<VerifyFile*> files{&vf};
    std::vector<VerifyItem> items;
    for (size_t s = 0; s < vf.starts.size(); s++)
        items.push_back(VerifyItem{0, vf.starts[s], s + 1 < vf.starts.size() ? vf.starts[s + 1] : vf.dm.tus.size(), 0});
    int nthreads = 0;
    rc = verify_items(files, items, true, &nthreads);
    out->wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    out->frames_per_sec = out->wall_ms > 0 ? out->frames * 1000.0 / out->wall_ms : 0;
    if (!out->status && rc) {   // engine-level failure that no frame was blamed for
        out->status = rc;
        snprintf(out->message, sizeof(out->message), "%s", error().c_str());
    }
    if (!out->status)
        snprintf(out->message, sizeof(out->message), "ok: %lld frames, %zu GOP segments, %d parser threads", (long long)out->frames, items.size(), nthreads);
    return out->status;
}

// Parses and reconstructs a list of GOP segments (possibly of several files) on this engine's GPU.  Segments are parsed by
// host_threads workers in list order (bounded look-ahead), issued to the GPU by the calling thread in list order; every shown frame
// lands in its file's report / digest array at its display index.  stop_on_error: the first failing frame ends the run (single
// file); otherwise a failing segment only marks its own file and the remaining segments still run (batch).
int Engine::verify_items(std::vector<VerifyFile*>& files, const std::vector<VerifyItem>& items, bool stop_on_error, int* threads_used) {
    Engine& eng = *this;
    const av1r_config& c = impl_->cfg;
    EngineImpl& E = *impl_;
    cudaSetDevice(c.device);
    int nthreads = c.host_threads > 0 ? c.host_threads : (int)std::thread::hardware_concurrency();
    nthreads = std::max(1, std::min(nthreads, 32));
    if (threads_used) *threads_used = nthreads;
    int rc = eng.flush();
    if (rc) return rc;
    impl_->pending.clear();
    cudaMemset(impl_->k3_stuck.p, 0, 4);
    const size_t nseg = items.size();
    std::vector<std::unique_ptr<Segment>> segs(nseg);
    for (size_t s = 0; s < nseg; s++) {
        segs[s] = std::make_unique<Segment>();
        segs[s]->tu0 = items[s].tu0;
        segs[s]->tu1 = items[s].tu1;
        const size_t n = segs[s]->tu1 - segs[s]->tu0;
        segs[s]->parsed.resize(n);
        segs[s]->rc.assign(n, 0);
        segs[s]->errs.resize(n);
    }
    // ---- parser threads: claim segments in order; bounded look-ahead (frames parsed but not yet issued)
    std::atomic<size_t> next_seg{0};
    std::atomic<bool> abort_flag{false};
    std::mutex la_m;
    std::condition_variable la_cv;
    int64_t la_outstanding = 0;
    const int64_t la_limit = std::max<int64_t>(2 * nthreads, 8);
    // the segment the consumer is issuing is exempt from the look-ahead budget: otherwise the workers of later multi-TU segments
    // can use the budget up while the worker the consumer waits for is the one blocked on it (deadlock seen on the C5 batch)
    std::atomic<size_t> consumer_seg{0};
    std::mutex ready_m;                  // the consumer sleeps on ready_cv when no segment has a parsed unit for it
    std::condition_variable ready_cv;
    auto worker = [&]() {
        cudaSetDevice(E.cfg.device);   // pinned staging is allocated from this thread
        WorkerPool::nested_enabled() = nseg < (size_t)nthreads;   // enough segments to fill the cores: no tile / band helpers
        // with fewer segments than threads the staging of TU t overlaps the parse of TU t+1 on this worker's helper thread; when the
        // segments alone fill the cores a helper would only add time-slicing: stage inline
        std::unique_ptr<SerialWorker> helper;
        if (WorkerPool::nested_enabled()) helper = std::make_unique<SerialWorker>(E.cfg.device);
        while (!abort_flag.load()) {
            const size_t s = next_seg.fetch_add(1);
            if (s >= nseg) return;
            Segment& sg = *segs[s];
            const VerifyFile& vf = *files[items[s].file];
            StreamParser sp;
            sp.hp.seq = vf.seq_for(sg.tu0);
            sp.defer_finalize = true;   // merge of the tile lists runs in the staging step below, off the parse chain
            sp.host_lf_edges = false;   // deblocking edges are classified on the device
            auto publish = [&sg, &ready_cv](size_t t, std::vector<ParsedFrame>&& pfs, int prc, const std::string& perr) {
                {
                    std::lock_guard<std::mutex> lk(sg.m);
                    sg.parsed[t - sg.tu0] = std::move(pfs);
                    sg.rc[t - sg.tu0] = prc;
                    if (prc) sg.errs[t - sg.tu0] = perr;
                    sg.n_done++;
                }
                sg.cv.notify_all();
                ready_cv.notify_one();
            };
            bool failed = false;
            for (size_t t = sg.tu0; t < sg.tu1 && !abort_flag.load() && !failed; t++) {
                {
                    auto t_w = EP_T();
                    std::unique_lock<std::mutex> lk(la_m);
                    la_cv.wait(lk, [&] { return la_outstanding < la_limit || s <= consumer_seg.load() || abort_flag.load(); });
                    la_outstanding++;
                    EP_ADD(18, t_w);
                }
                auto pfs = std::make_shared<std::vector<ParsedFrame>>();
                const TemporalUnit& tu = vf.dm.tus[t];
                auto t_p = EP_T();
                const int prc = sp.parse_tu(vf.data + tu.offset, tu.size, ((int64_t)s << 32) | (int64_t)t, *pfs);
                EP_ADD(19, t_p);
                const std::string perr = sp.err;
                auto stage = [&E, publish, t, pfs, prc, perr]() {
                    int rc2 = prc;
                    std::string e2 = perr;
                    for (ParsedFrame& pf : *pfs)
                        if (pf.fw && !rc2) {
                            finalize_framework(*pf.fw);
                            int hrc = 0;
                            std::string herr;
                            pf.host = E.make_host_arena(*pf.fw, pf.seq, herr, hrc);
                            if (hrc) { rc2 = hrc; e2 = herr; }
                        }
                    publish(t, std::move(*pfs), rc2, e2);
                };
                if (helper) helper->post(stage);
                else stage();
                if (prc) failed = true;
            }
            if (helper) helper->drain();
            {   // mark the remaining TUs of a failed segment as done so the consumer never blocks
                std::lock_guard<std::mutex> lk(sg.m);
                sg.n_done = sg.tu1 - sg.tu0;
            }
            sg.cv.notify_all();
            ready_cv.notify_one();
        }
    };
    std::vector<std::thread> pool;
    for (int i = 0; i < nthreads; i++) pool.emplace_back(worker);
    // ---- consumer: issue GPU work segment by segment, TU by TU (display order)
    std::vector<av1r_frame_result> res(64);
    int64_t last_pts = -1;
    int same_pts = 0;
    int fatal = 0;          // engine-level error (CUDA fault): nothing more can run on this device
    bool stop = false;      // stop_on_error and a frame failed
    auto drain = [&](bool all) -> int {
        while (true) {
            int n = 0;
            if (all) {
                int r = eng.flush();
                if (r) return r;
            }
            int r = eng.collect(res.data(), (int)res.size(), &n);
            if (r) return r;
            for (int i = 0; i < n; i++) {
                const size_t s = (size_t)(res[i].pts >> 32), t = (size_t)(res[i].pts & 0xffffffff);
                VerifyFile& vf = *files[items[s].file];
                same_pts = res[i].pts == last_pts ? same_pts + 1 : 0;
                last_pts = res[i].pts;
                const int64_t idx = vf.frame_base[t] + same_pts;
                if (res[i].status) {   // a frame the device could not reconstruct consistently (intra kernel watchdog)
                    vf.fail(res[i].status, idx, E.err);
                    if (stop_on_error) stop = true;
                    continue;
                }
                std::lock_guard<std::mutex> lk(vf.m);
                if (vf.digests && idx < vf.cap_frames) memcpy(vf.digests + 3 * idx, res[i].checksum, 24);
                vf.rep->frames++;
                vf.rep->device_ms += res[i].device_ms;
                vf.rep->width = res[i].w;
                vf.rep->height = res[i].h;
                vf.rep->bit_depth = res[i].bpc;
            }
            if (n == 0) return 0;
        }
    };
    // Frames are issued as soon as they are parsed, from whichever active segment has one ready (every shown frame finds its place
    // through its pts, so the order between segments is free): a slow segment at the head of the list does not hold back the
    // frames other workers have already parsed, and a small look-ahead budget is enough however many segments there are.
    struct SegIssue {
        size_t next_t = 0;
        bool failed = false, done = false;
        std::unique_ptr<EngineImpl::RefState> refs;
    };
    std::vector<SegIssue> st(nseg);
    size_t lo = 0;   // first segment not yet fully issued
    while (lo < nseg && !fatal && !stop) {
        bool progressed = false;
        const size_t hi = std::min(nseg, next_seg.load());   // segments claimed by a worker so far
        for (size_t s = lo; s < hi && !fatal && !stop; s++) {
            SegIssue& si = st[s];
            if (si.done) continue;
            Segment& sg = *segs[s];
            VerifyFile& vf = *files[items[s].file];
            const size_t ntu = sg.tu1 - sg.tu0;
            while (si.next_t < ntu && !fatal && !stop) {
                const size_t t = si.next_t;
                std::vector<ParsedFrame> pfs;
                int prc;
                std::string msg;
                {
                    std::lock_guard<std::mutex> lk(sg.m);
                    if (sg.n_done <= t) break;          // not parsed yet: look at the other segments
                    pfs = std::move(sg.parsed[t]);
                    prc = sg.rc[t];
                    if (prc) msg = sg.errs[t];
                }
                si.next_t++;
                progressed = true;
                if (!si.refs) si.refs = std::make_unique<EngineImpl::RefState>();
                E.rs = si.refs.get();
                int frc = 0;
                if (!si.failed)
                    for (ParsedFrame& pf : pfs) {
                        if (pf.fw) {
                            std::lock_guard<std::mutex> lk(vf.m);
                            vf.rep->host_parse_ms += pf.fw->parse_ms;
                        }
                        int r = E.decode_parsed(pf);
                        if (r) { frc = r; msg = E.err; break; }
                    }
                E.rs = &E.main_refs;
                {
                    std::lock_guard<std::mutex> lk(la_m);
                    la_outstanding--;
                }
                la_cv.notify_all();
                if (si.failed) continue;          // rest of a failed segment: only hand the look-ahead budget back
                if (!frc && prc) frc = prc;
                if (frc) {
                    char tmp[600];
                    snprintf(tmp, sizeof(tmp), "temporal unit %zu: %s", sg.tu0 + t, msg.c_str());
                    vf.fail(frc, vf.frame_base[sg.tu0 + t], tmp);
                    if (frc == AV1R_EIO || frc == AV1R_ENOMEM) fatal = frc;
                    si.failed = true;
                    if (stop_on_error) stop = true;
                    continue;
                }
                auto td = EP_T();
                int r = drain(false);
                EP_ADD(5, td);
                if (r) fatal = r;
            }
            if (si.next_t == ntu) {
                si.done = true;
                si.refs.reset();                   // the segment's reference frames go back to the pool (slots still hold what is queued)
            }
        }
        while (lo < nseg && st[lo].done) lo++;
        {
            std::lock_guard<std::mutex> lk(la_m);
            consumer_seg.store(lo);
        }
        la_cv.notify_all();
        if (!progressed && lo < nseg) {
            // nothing ready anywhere: sleep until a worker publishes a temporal unit (or for a short while: the claim of a new
            // segment is not signalled)
            auto tw = EP_T();
            std::unique_lock<std::mutex> lk(ready_m);
            ready_cv.wait_for(lk, std::chrono::microseconds(200));
            EP_ADD(4, tw);
        }
    }
    E.rs = &E.main_refs;
    {   // under la_m: a worker between its predicate test and its block must not miss the wake-up
        std::lock_guard<std::mutex> lk(la_m);
        abort_flag.store(true);
    }
    la_cv.notify_all();
    for (auto& th : pool) th.join();
    int r = drain(true);
    if (!fatal && r) fatal = r;
    return fatal;
}

}  // namespace av1r

struct av1r_clip {
    std::vector<std::unique_ptr<av1r::ClipFrame>> frames;
    av1r_clip_info info;
    int resident = 0;   // 1: work-lists uploaded once, replays read them from HBM (no H2D inside the pass)
};

namespace av1r {

int Engine::clip_load(const uint8_t* const* tus, const size_t* lens, int n, av1r_clip** out) {
    EngineImpl& E = *impl_;
    std::string& err = E.err;
    cudaSetDevice(E.cfg.device);
    auto clip = std::make_unique<av1r_clip>();
    memset(&clip->info, 0, sizeof(clip->info));
    clip->info.struct_size = sizeof(clip->info);
    // ---- cut at shown key frames (header scan), then parse the GOP segments on all host cores; per-segment results are merged in
    // order, so the clip is the same as a sequential parse would give
    std::vector<int> starts;
    std::vector<SeqHdr> seg_seq;        // sequence header in force where each segment starts
    {
        HeaderParser scan;
        for (int i = 0; i < n; i++) {
            bool seg = false;
            int shown = 0;
            const int rc = scan_temporal_unit(scan, tus[i], lens[i], &seg, &shown, err);
            if (rc) return rc;
            if (seg || i == 0) {
                starts.push_back(i);
                seg_seq.push_back(scan.seq);
            }
        }
    }
    const int nseg = (int)starts.size();
    struct SegOut {
        std::vector<std::unique_ptr<ClipFrame>> frames;
        av1r_clip_info info;
        int rc = 0;
        std::string err;
    };
    std::vector<SegOut> outs(nseg);
    std::atomic<int> next{0};
    const int inloop = E.cfg.inloop_filters, apply_grain = E.cfg.apply_grain, device = E.cfg.device;
    auto worker = [&]() {
        cudaSetDevice(device);
        WorkerPool::nested_enabled() = nseg < (int)std::thread::hardware_concurrency();
        while (true) {
            const int sidx = next.fetch_add(1);
            if (sidx >= nseg) return;
            SegOut& so = outs[sidx];
            memset(&so.info, 0, sizeof(so.info));
            StreamParser parser;
            parser.hp.seq = seg_seq[sidx];
            parser.host_lf_edges = false;
            const int t1 = sidx + 1 < nseg ? starts[sidx + 1] : n;
            for (int t = starts[sidx]; t < t1 && !so.rc; t++) {
                std::vector<ParsedFrame> pfs;
                int rc = parser.parse_tu(tus[t], lens[t], t, pfs);
                if (rc) { so.rc = rc; so.err = parser.err; break; }
                for (ParsedFrame& pf : pfs) {
                    auto cf = std::make_unique<ClipFrame>();
                    cf->fh = pf.fh;
                    cf->seq = pf.seq;
                    cf->pts = pf.pts;
                    if (pf.show_existing_slot >= 0) {
                        cf->show_existing = true;
                        cf->show_slot = pf.show_existing_slot;
                        so.info.frames_shown++;
                    } else {
                        const FrameWork& fw = *pf.fw;
                        rc = prepare_work_seq(parser.hp.seq, fw, cf->dw, so.err);
                        if (rc) { so.rc = rc; break; }
                        if (cudaHostAlloc(&cf->host, align_up(cf->dw.lay.total, 4096), cudaHostAllocDefault) != cudaSuccess) {
                            so.rc = AV1R_ENOMEM;
                            so.err = "cudaHostAlloc(clip work-lists) failed";
                            break;
                        }
                        fill_arena(fw, cf->dw, cf->host);
                        av1r_clip_info& ci = so.info;
                        ci.frames_decoded++;
                        ci.frames_shown += fw.fh.show_frame;
                        ci.host_parse_ms += fw.parse_ms;
                        ci.worklist_bytes += cf->dw.lay.total;
                        ci.coded_samples += fw.coded_samples;
                        ci.coef_tokens += fw.coefs.size();
                        ci.tx_blocks += fw.tx.size();
                        ci.inter_samples += fw.inter_samples;
                        ci.inter_ref_samples += fw.inter_ref_samples;
                        ci.inter_blocks += fw.inter.size();
                        ci.obmc_neighbours += fw.obmc.size();
                        ci.lr_frames += cf->dw.lr_on && (inloop & 4);
                        ci.cdef_frames += cf->dw.cdef_on && (inloop & 2);
                        ci.deblock_frames += cf->dw.lf_on && (inloop & 1);
                        ci.grain_frames += fw.fh.show_frame && fw.fh.fg.apply_grain && apply_grain;
                        for (int i = 0; i < 24; i++) ci.tool_hist[i] += fw.tool_hist[i];
                        uint64_t isamp = 0;
                        for (const TxRec& r : fw.tx)
                            if (r.mode != TXM_INTER) isamp += (uint64_t)kTxW[r.txsz] * kTxH[r.txsz];
                        ci.intra_samples += isamp;
                        if (fw.inter.empty()) {
                            ci.intra_frame_samples += isamp;
                            ci.intra_frame_coded_samples += fw.coded_samples;
                            ci.intra_frame_tx_blocks += fw.tx.size();
                            ci.intra_frames++;
                        }
                        ci.width = fw.fh.upscaled_width;
                        ci.height = fw.fh.frame_height;
                        ci.bit_depth = parser.hp.seq.bit_depth;
                        const DevFrameParams& fp = cf->dw.fp_up;
                        ci.frame_bytes = 0;
                        for (int p = 0; p < (fp.mono ? 1 : 3); p++) ci.frame_bytes += (uint64_t)fp.w[p] * fp.h[p] * (fp.bd == 8 ? 1 : 2);
                    }
                    so.frames.push_back(std::move(cf));
                }
            }
        }
    };
    {
        const int nt = std::max(1, std::min(nseg, (int)std::thread::hardware_concurrency()));
        std::vector<std::thread> th;
        for (int i = 1; i < nt; i++) th.emplace_back(worker);
        worker();
        for (auto& t : th) t.join();
    }
    for (SegOut& so : outs) {
        if (so.rc) { err = so.err; return so.rc; }
        av1r_clip_info& d = clip->info;
        const av1r_clip_info& a = so.info;
        d.frames_decoded += a.frames_decoded; d.frames_shown += a.frames_shown; d.host_parse_ms += a.host_parse_ms;
        d.worklist_bytes += a.worklist_bytes; d.coded_samples += a.coded_samples; d.coef_tokens += a.coef_tokens;
        d.tx_blocks += a.tx_blocks; d.intra_samples += a.intra_samples; d.inter_samples += a.inter_samples;
        d.inter_ref_samples += a.inter_ref_samples; d.lr_frames += a.lr_frames; d.cdef_frames += a.cdef_frames;
        d.deblock_frames += a.deblock_frames; d.grain_frames += a.grain_frames; d.inter_blocks += a.inter_blocks;
        d.obmc_neighbours += a.obmc_neighbours;
        d.intra_frame_samples += a.intra_frame_samples; d.intra_frame_coded_samples += a.intra_frame_coded_samples;
        d.intra_frame_tx_blocks += a.intra_frame_tx_blocks; d.intra_frames += a.intra_frames;
        for (int i = 0; i < 24; i++) d.tool_hist[i] += a.tool_hist[i];
        if (a.frames_decoded) { d.width = a.width; d.height = a.height; d.bit_depth = a.bit_depth; d.frame_bytes = a.frame_bytes; }
        for (auto& cf : so.frames) clip->frames.push_back(std::move(cf));
    }
    *out = clip.release();
    return 0;
}

// One pass over the clip.  Default: every frame's work-lists travel host -> device inside the pass (one copy per frame from the
// clip's pinned arena into the slot's device arena, as in the streaming path: SURVEY 8d times the device path "from the first
// work-list H2D enqueue").  Resident mode (av1r_clip_set_resident): the lists are uploaded once and replays read them from HBM.
static int replay(EngineImpl& E, av1r_clip* clip) {
    std::string& err = E.err;
    for (auto& cf : clip->frames) {
        int slot_idx;
        auto t_q = EP_T();
        int rc = E.acquire_slot(slot_idx);
        EP_ADD(0, t_q);
        if (rc) return rc;
        FrameSlot& s = *E.slots[slot_idx];
        E.sp.hp.seq = cf->seq;
        if (cf->show_existing) {
            rc = E.exec_show_existing(slot_idx, cf->fh, cf->show_slot, cf->pts);
        } else {
            const uint8_t* d_arena;
            CK(cudaEventRecord(s.ev0, s.stream));
            if (clip->resident) {
                if (!cf->arena_valid) {
                    CK(cf->arena.ensure(cf->dw.lay.total));
                    CK(cudaMemcpy(cf->arena.p, cf->host, cf->dw.lay.total, cudaMemcpyHostToDevice));
                    cf->arena_valid = true;
                }
                d_arena = cf->arena.p;
            } else {
                auto t_a = EP_T();
                { int e_ = E.ensure_slots(&FrameSlot::arena, s, cf->dw.lay.total, E.hw_arena); if (e_) return e_; }
                if (E.tm) E.tm->begin(s.stream);
                CK(cudaMemcpyAsync(s.arena.p, cf->host, cf->dw.lay.total, cudaMemcpyHostToDevice, s.stream));
                if (E.tm) E.tm->end(AV1R_ST_H2D, 1, s.stream);
                d_arena = s.arena.p;
                EP_ADD(6, t_a);
            }
            auto t_x = EP_T();
            rc = E.exec_decoded(slot_idx, cf->dw, d_arena, cf->pts);
            EP_ADD(7, t_x);
        }
        if (rc) return rc;
    }
    return 0;
}

int Engine::clip_decode(av1r_clip* clip, uint64_t* cks, int cap, int* n_frames, float* device_ms) {
    return clip_decode_passes(clip, 1, cks, cap, n_frames, device_ms);
}

// `passes` replays of the clip enqueued back to back between ONE pair of fencing events: pass k + 1's first key frames start while
// pass k's last frames drain, as consecutive GOPs of a long file do in the streaming path (slots, frame buffers and reference events
// are recycled exactly as there).  Every pass must reproduce the digests of the first one.
int Engine::clip_decode_passes(av1r_clip* clip, int passes, uint64_t* cks, int cap, int* n_frames, float* device_ms) {
    EngineImpl& E = *impl_;
    std::string& err = E.err;
    if (passes < 1) return AV1R_EINVAL;
    cudaSetDevice(E.cfg.device);
    int rc = flush();
    if (rc) return rc;
    E.pending.clear();
    cudaMemset(E.k3_stuck.p, 0, 4);
    if (!clip->resident) {   // no slot arena may have to grow (cudaFree + cudaMalloc) inside the timed pass
        size_t mx = 0;
        for (auto& cf : clip->frames)
            if (!cf->show_existing) mx = std::max(mx, cf->dw.lay.total);
        if (!E.slots.empty()) { int e_ = E.ensure_slots(&FrameSlot::arena, *E.slots[0], mx, E.hw_arena); if (e_) return e_; }
    }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    std::vector<cudaEvent_t> done(E.streams.size());
    for (auto& d : done) CK(cudaEventCreateWithFlags(&d, cudaEventDisableTiming));
    // fence: every stream starts after e0 (recorded on stream 0)
    CK(cudaEventRecord(e0, E.streams[0]));
    for (size_t i = 1; i < E.streams.size(); i++) CK(cudaStreamWaitEvent(E.streams[i], e0, 0));
    for (int k = 0; k < passes; k++) {
        rc = replay(E, clip);
        if (rc) return rc;
    }
    for (size_t i = 1; i < E.streams.size(); i++) {
        CK(cudaEventRecord(done[i], E.streams[i]));
        CK(cudaStreamWaitEvent(E.streams[0], done[i], 0));
    }
    CK(cudaEventRecord(e1, E.streams[0]));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (device_ms) *device_ms = ms;
    for (auto& p : E.pending) {
        rc = E.finish_pending(p);
        if (rc) return rc;
    }
    const size_t per_pass = E.pending.size() / (size_t)passes;
    int k = 0;
    for (size_t i = 0; i < E.pending.size(); i++) {
        const auto& p = E.pending[i];
        if (i < per_pass) {
            if (cks && k < cap) memcpy(cks + 3 * k, p.res.checksum, 24);
            k++;
        } else if (memcmp(p.res.checksum, E.pending[i % per_pass].res.checksum, 24) != 0) {
            err = "clip replay: a later pass produced different digests than the first one (non-deterministic reconstruction)";
            E.pending.clear();
            return AV1R_EIO;
        }
    }
    E.pending.clear();
    if (n_frames) *n_frames = k;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    for (auto& d : done) cudaEventDestroy(d);
    return 0;
}

int Engine::clip_profile(av1r_clip* clip, av1r_stage_times* out) {
    EngineImpl& E = *impl_;
    cudaSetDevice(E.cfg.device);
    int rc = flush();
    if (rc) return rc;
    E.pending.clear();
    cudaMemset(E.k3_stuck.p, 0, 4);
    memset(out, 0, sizeof(*out));
    out->struct_size = sizeof(*out);
    // serialise on one stream so that the spans do not overlap
    std::vector<cudaStream_t> saved = E.streams;
    for (auto& st : E.streams) st = saved[0];
    StageTimer tm;
    E.tm = &tm;
    rc = replay(E, clip);
    E.tm = nullptr;
    E.streams = saved;
    if (rc) return rc;
    if (cudaDeviceSynchronize() != cudaSuccess) return AV1R_EIO;
    tm.collect(out);
    for (auto& p : E.pending) E.finish_pending(p);
    E.pending.clear();
    return 0;
}

}  // namespace av1r

extern "C" int av1r_clip_info_get(const av1r_clip* clip, av1r_clip_info* out) {
    if (!clip || !out) return AV1R_EINVAL;
    *out = clip->info;
    return 0;
}
extern "C" void av1r_clip_free(av1r_clip* clip) { delete clip; }
extern "C" int av1r_clip_set_resident(av1r_clip* clip, int resident) {
    if (!clip) return AV1R_EINVAL;
    clip->resident = resident != 0;
    return 0;
}
