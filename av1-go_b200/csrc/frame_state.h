// Host-side per-frame state produced by the sequential parse: mode info per block, the device
// work-lists, and what later frames need from this one (CDFs, segment ids, motion vectors).
#pragma once
#include <cstdint>
#include <cstring>
#include <deque>
#include <memory>
#include <string>
#include <vector>

#include "av1_consts.h"
#include "hdr.h"
#include "worklist.h"

namespace av1r {

#include "tables/tables_cdf.inc"   // struct CdfCtx + av1t_default_*_cdf

static_assert(sizeof(CdfCtx) == 2 * (AV1T_NCOEF_CDF + AV1T_NMODE_CDF), "CdfCtx layout");

struct Mv {
    int16_t row, col;
};

struct BlockInfo {
    uint16_t mi_row, mi_col;
    uint8_t bsize, skip, skip_mode, is_inter, segment_id, use_intrabc;
    uint8_t y_mode, uv_mode;
    int8_t angle_y, angle_uv;
    uint8_t use_filter_intra, fi_mode;
    int8_t cfl_alpha_u, cfl_alpha_v;
    uint8_t pal_size[2];
    uint16_t pal_colors[3][8];
    int8_t ref_frame[2];
    Mv mv[2];
    uint8_t interp_filter[2];
    uint8_t motion_mode, compound_type, comp_group_idx, compound_idx;
    uint8_t interintra, interintra_mode, wedge_interintra, wedge_index, wedge_sign, mask_type;
    uint8_t tx_size;           // luma TxSize (largest when var-tx)
    uint8_t qidx;              // qindex used by this block's residual
    int8_t delta_lf[4];
    uint8_t has_chroma;
    uint8_t num_mv_found, is_global_or_default;
    uint8_t lossless;
    uint8_t warp_valid;        // LocalValid (motion_mode == WARPED_CAUSAL)
    uint8_t lf_lvl[4];         // deblocking levels of this block: luma vertical edges, luma horizontal, U, V (spec 7.14.4)
    int32_t warp[6];           // LocalWarpParams
};

// Per-mi digest of a block for the loop-filter edge pre-pass (8 bytes, read sequentially instead of chasing BlockInfo pointers)
struct LfMi {
    uint8_t lvl[4];            // BlockInfo::lf_lvl
    uint8_t bsize;
    uint8_t filt_inside;       // !skip || intra: transform edges inside the block are filtered too
    uint8_t valid;             // 1 once a block of this frame covers the mi
    uint8_t pad;
};

// Motion vector + reference saved per 8x8 for later frames (spec 7.19), and the projected field (7.9)
struct SavedMv {
    Mv mv;
    int8_t ref;                // 0 = none
};
struct MfMv {
    Mv mv;
    int8_t ref_offset;         // 0 = invalid entry
};

enum { TOOL_INTER_BLOCKS, TOOL_COMPOUND_AVG, TOOL_COMPOUND_DIST, TOOL_COMPOUND_WEDGE, TOOL_COMPOUND_DIFFWTD, TOOL_INTERINTRA,
       TOOL_INTERINTRA_WEDGE, TOOL_OBMC, TOOL_LOCAL_WARP, TOOL_GLOBAL_WARP, TOOL_SKIP_MODE, TOOL_DUAL_FILTER, TOOL_TEMPORAL_MV,
       TOOL_INTRA_IN_INTER, TOOL_SUB8X8_CHROMA, TOOL_NEWMV, TOOL_VARTX_SPLIT, TOOL_SWITCHABLE_FILTER, TOOL_PALETTE, TOOL_INTRABC, TOOL_COUNT };

inline void cdf_load_defaults(CdfCtx& c, int base_q_idx) {
    int q = base_q_idx <= 20 ? 0 : base_q_idx <= 60 ? 1 : base_q_idx <= 120 ? 2 : 3;
    uint16_t* p = reinterpret_cast<uint16_t*>(&c);
    memcpy(p, av1t_default_coef_cdf[q], sizeof(uint16_t) * AV1T_NCOEF_CDF);
    memcpy(p + AV1T_NCOEF_CDF, av1t_default_mode_cdf, sizeof(uint16_t) * AV1T_NMODE_CDF);
}
inline void cdf_load_default_coefs(CdfCtx& c, int base_q_idx) {
    int q = base_q_idx <= 20 ? 0 : base_q_idx <= 60 ? 1 : base_q_idx <= 120 ? 2 : 3;
    memcpy(reinterpret_cast<uint16_t*>(&c), av1t_default_coef_cdf[q], sizeof(uint16_t) * AV1T_NCOEF_CDF);
}

// Loop-restoration parameters of one unit
struct LrUnit {
    uint8_t type;          // RESTORE_NONE / WIENER / SGRPROJ
    uint8_t sgr_set;
    int8_t wiener[2][3];   // [pass: 0 vertical, 1 horizontal][coef]
    int8_t sgr_xqd[2];
};

// Growable array of coefficient tokens without value-initialisation: the coefficient reader appends up to eob tokens per
// transform block through a raw pointer (std::vector::resize would zero-fill them first).
struct TokenBuf {
    uint32_t* p = nullptr;
    size_t n = 0, cap = 0;
    TokenBuf() = default;
    TokenBuf(const TokenBuf&) = delete;
    TokenBuf& operator=(const TokenBuf&) = delete;
    ~TokenBuf() { free(p); }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    void clear() { n = 0; }
    const uint32_t* begin() const { return p; }
    const uint32_t* end() const { return p + n; }
    uint32_t* tail(size_t extra) {   // room for `extra` more tokens; returns the write position (size is unchanged)
        if (n + extra > cap) {
            cap = std::max<size_t>(2 * cap, n + extra + 4096);
            p = static_cast<uint32_t*>(realloc(p, cap * sizeof(uint32_t)));
        }
        return p + n;
    }
};

// What one tile's parse produces.  Tiles are independent given the frame's initial CDFs, so they are parsed concurrently, each
// into its own TileOut; StreamParser merges the lists in tile order (offsets inside the records are rebased there).
struct TileOut {
    std::vector<TxRec> tx;
    TokenBuf coefs;
    std::vector<SbRange> sbs;
    std::vector<uint8_t> pal;
    std::vector<InterBlk> inter;
    std::vector<ObmcNb> obmc;
    std::vector<WarpRec> warps;         // local warp models of this tile (InterBlk::warp = 8 + index until merged)
    std::vector<LfBlk> lf_blocks;       // device-side deblocking edge classification: one record per block
    std::deque<BlockInfo> blocks;       // mode info storage; FrameWork::mi points into it
    uint64_t coded_samples = 0, coef_tokens = 0, tx_blocks = 0, inter_samples = 0, inter_ref_samples = 0;
    uint32_t tool_hist[24] = {0};
    CdfCtx end_cdf;
    int rc = 0;
    std::string err;
};

// Everything the device needs for one frame + what the next frames need from the parse.
struct FrameWork {
    FrameHdr fh;
    int bit_depth = 8, subx = 1, suby = 1, mono = 0, sb128 = 0;
    int mi_cols = 0, mi_rows = 0;
    // device work-lists
    std::vector<TxRec> tx;
    std::vector<uint32_t> coefs;
    std::vector<SbRange> sbs;
    std::vector<uint8_t> pal;
    std::vector<LfEdge> lf[3];          // per plane, (plane_h4 x plane_w4): host-side edge classification (oracle / tests)
    std::vector<LfBlk> lf_blocks;       // device-side edge classification: block list (see host_lf)
    bool host_lf = true;                // true: build_loopfilter_edges fills lf[] on the host; false: the engine ships lf_blocks + lf_tx
    std::vector<int8_t> cdef_idx;       // per 64x64 luma block, -1 = skip
    std::vector<uint8_t> skip_mi;       // per mi: block skip flag (CDEF 8x8 skip condition)
    std::vector<LrUnit> lr[3];
    int lr_cols[3] = {0, 0, 0}, lr_rows[3] = {0, 0, 0};
    std::vector<InterBlk> inter;        // K2 work-list
    std::vector<ObmcNb> obmc;
    std::vector<WarpRec> warps;         // [0..7] global models per reference frame (slot 0 unused), then local ones
    uint8_t gm_warp_valid[8] = {0};
    uint8_t ref_scaled[8] = {0};        // per reference frame (LAST .. ALTREF): its size differs from this frame's (spec 7.11.3.3 is_scaled)
    // motion state: what this frame's parse reads from earlier frames and what it leaves for later ones
    std::vector<uint8_t> prev_seg_ids;  // PrevSegmentIds (empty = all zero)
    std::vector<MfMv> mfmv;             // projected motion field per 8x8 (empty = no temporal candidates)
    std::vector<SavedMv> saved_mvs;     // per 8x8, filled at frame end
    int saved_order_hints[8] = {0};
    uint64_t inter_samples = 0, inter_ref_samples = 0;
    // tool histogram (blocks): see TOOL_* below; reported by the bench so the exercised tool set is visible
    uint32_t tool_hist[24] = {0};   // predicted samples / reference samples fetched (roofline model)
    // mode info (host only)
    std::vector<std::unique_ptr<TileOut>> tiles;   // per-tile outputs (own the BlockInfo storage)
    std::vector<BlockInfo*> mi;         // per mi -> block
    std::vector<LfMi> lf_mi;            // per mi: what the deblocking edge builder needs of the block (levels, size, skip && inter)
    std::vector<uint8_t> inter_tx;      // per mi: InterTxSizes
    std::vector<uint8_t> lf_tx[3];      // per plane 4x4: LoopfilterTxSizes
    std::vector<uint8_t> tx_types;      // per mi (luma)
    std::vector<uint8_t> seg_ids;       // per mi
    // statistics for the roofline model: coded samples A, coefficient bytes C
    uint64_t coded_samples = 0, coef_tokens = 0, tx_blocks = 0;
    double parse_ms = 0;
    // CDFs at the end of the context-update tile (saved into refreshed slots)
    CdfCtx end_cdf;
    bool have_end_cdf = false;
    bool finalized = true;              // false between the tile parse and finalize_framework() (stream_parser.h)

    void init(const SeqHdr& seq, const FrameHdr& h) {
        fh = h;
        bit_depth = seq.bit_depth;
        subx = seq.subsampling_x;
        suby = seq.subsampling_y;
        mono = seq.mono_chrome;
        sb128 = seq.use_128x128_superblock;
        mi_cols = h.mi_cols;
        mi_rows = h.mi_rows;
        // A FrameWork is recycled (stream_parser.cpp): only what the parse *reads before writing* is cleared.  mi must be null
        // ("not yet decoded in this frame"): every tile clears its own rectangle when it starts (TileDecoder::decode_tile; a tile
        // only ever looks at its own rectangle, and the 4 MB fill of a 4K frame leaves the frame-to-frame chain for the tile
        // threads); every other per-mi map is fully written by the parse of the frame before anything reads it (each block
        // writes its whole area), so resizing without a fill is enough.
        size_t n = (size_t)mi_cols * mi_rows;
        mi.resize(n);
        lf_mi.resize(n);   // written for every mi a block covers; the edge builder only reads covered mi
        inter_tx.resize(n);
        tx_types.resize(n);
        seg_ids.resize(n);
        skip_mi.resize(n);
        for (int p = 0; p < 3; p++) {
            int sx = p ? subx : 0, sy = p ? suby : 0;
            size_t pn = (size_t)((mi_cols + sx) >> sx) * ((mi_rows + sy) >> sy);
            lf_tx[p].resize(pn);
            lf[p].resize(pn);
            lr[p].clear();
            lr_rows[p] = lr_cols[p] = 0;
        }
        int c64 = (mi_cols + 15) >> 4, r64 = (mi_rows + 15) >> 4;
        cdef_idx.assign((size_t)c64 * r64, -1);
        tx.clear();
        coefs.clear();
        sbs.clear();
        pal.clear();
        inter.clear();
        obmc.clear();
        warps.clear();
        lf_blocks.clear();
        n_tiles_used = 0;
        prev_seg_ids.clear();
        mfmv.clear();
        saved_mvs.clear();
        memset(gm_warp_valid, 0, sizeof(gm_warp_valid));
        memset(ref_scaled, 0, sizeof(ref_scaled));
        memset(tool_hist, 0, sizeof(tool_hist));
        coded_samples = coef_tokens = tx_blocks = inter_samples = inter_ref_samples = 0;
        parse_ms = 0;
        have_end_cdf = false;
        finalized = true;
    }
    // per-tile outputs are recycled with the frame: hands out the next one, cleared but with its capacity kept
    TileOut& next_tile() {
        if (n_tiles_used == tiles.size()) tiles.push_back(std::make_unique<TileOut>());
        TileOut& t = *tiles[n_tiles_used++];
        t.tx.clear();
        t.coefs.clear();
        t.sbs.clear();
        t.pal.clear();
        t.inter.clear();
        t.obmc.clear();
        t.warps.clear();
        t.lf_blocks.clear();
        t.blocks.clear();
        t.coded_samples = t.coef_tokens = t.tx_blocks = t.inter_samples = t.inter_ref_samples = 0;
        memset(t.tool_hist, 0, sizeof(t.tool_hist));
        t.rc = 0;
        t.err.clear();
        return t;
    }
    size_t n_tiles_used = 0;
    int plane_w4(int p) const { int sx = p ? subx : 0; return (mi_cols + sx) >> sx; }
    int plane_h4(int p) const { int sy = p ? suby : 0; return (mi_rows + sy) >> sy; }
};

}  // namespace av1r
