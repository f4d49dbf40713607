// Host <-> device contract: the work-lists the sequential host parse emits and the sm_100a
// kernels consume.  POD only; shared by g++ (parser, oracle) and nvcc (kernels).
#pragma once
#include <stdint.h>

namespace av1r {

// One record per transform block, in decode order (which is also the only valid order for
// intra prediction inside a superblock).  32 bytes.
struct TxRec {
    uint16_t x4, y4;        // top-left in the plane, in units of 4 samples
    uint8_t plane;          // 0 Y, 1 U, 2 V
    uint8_t txsz;           // TxSize
    uint8_t txtp;           // TxType (16 = WHT, lossless)
    uint8_t mode;           // TXM_* below or intra PredMode 0..12
    uint16_t eob;           // 0 = no coded residual
    uint8_t qidx;           // qindex used for dequantisation (segment + delta-q applied)
    uint8_t flags;          // TXF_*
    uint32_t coef_off;      // first token of this block in the coefficient token array
    int8_t angle_delta;     // -3..3 (directional modes)
    uint8_t fi_mode;        // filter-intra mode 0..4 (mode == TXM_FILTER_INTRA)
    int16_t cfl_alpha;      // signed CfL alpha for this plane (mode == TXM_CFL)
    uint16_t cfl_max_w4;    // luma samples available to CfL, /4, absolute (MaxLumaW, MaxLumaH)
    uint16_t cfl_max_h4;
    uint32_t pal_off;       // palette: offset into the palette byte array (colours + index map)
    uint8_t qm_level;       // quantiser-matrix level (15 = flat)
    uint8_t seg_id;
    uint16_t ntok;          // number of (non-zero) coefficient tokens at coef_off (<= eob)
};
static_assert(sizeof(TxRec) == 32, "TxRec must stay 32 bytes");

// ---- inter prediction work-lists (K2) -------------------------------------------------------------
// One record per prediction rectangle: normally one per inter block (all planes); the chroma of a group of
// sub-8x8 luma blocks is predicted from each block's own motion, so such groups add chroma-only records.
struct InterBlk {
    uint16_t x, y;          // luma sample position of the rectangle
    uint8_t w, h;           // luma size (4..128)
    uint8_t planes;         // bit 0 luma, bit 1 chroma
    uint8_t bsize;          // MiSize of the block (wedge mask geometry)
    int8_t ref[2];          // reference *slot* (ref_frame_idx[refFrame - LAST_FRAME]); ref[1] = -1: single prediction
    uint8_t filt[2];        // interp_filter[0] (vertical), interp_filter[1] (horizontal)
    int16_t mv[2][2];       // [list][row, col] in 1/8 luma sample
    int16_t warp[2];        // index into the WarpRec list per list (-1: translational)
    uint8_t comp_type;      // COMPOUND_WEDGE / DIFFWTD / AVERAGE / INTRA / DISTANCE
    uint8_t wedge_index, wedge_sign, mask_type;
    uint8_t fwd_w, bck_w;   // distance weights (COMPOUND_DISTANCE)
    uint8_t interintra;     // 1: K2 leaves the clipped inter predictor, the wavefront kernel blends the intra part
    uint8_t obmc_above, obmc_left;   // neighbour counts in the ObmcNb list
    uint8_t obmc_chroma_above;       // chroma plane is >= 8x8: the above pass applies to chroma too
    uint16_t pad;
    uint32_t obmc_first;
};
static_assert(sizeof(InterBlk) == 40, "InterBlk must stay 40 bytes");

// One overlapped-motion neighbour (spec 7.11.3.10): predict the overlap area with the neighbour's motion and blend.
struct ObmcNb {
    uint16_t x4, y4;        // luma mi position where the overlap area starts
    uint8_t step4;          // neighbour extent along the edge, in mi units (clipped 2..16)
    int8_t ref;             // reference slot
    uint8_t filt[2];
    int16_t mv[2];
};
static_assert(sizeof(ObmcNb) == 12, "ObmcNb must stay 12 bytes");

// Warp model (global or local) with its shear decomposition (spec 7.11.3.6).
struct WarpRec {
    int32_t mat[6];
    int16_t alpha, beta, gamma, delta;
};
static_assert(sizeof(WarpRec) == 32, "WarpRec must stay 32 bytes");

enum : uint8_t {
    TXM_CFL = 13,           // DC_PRED followed by chroma-from-luma
    TXM_PALETTE = 14,
    TXM_FILTER_INTRA = 15,
    TXM_INTRABC = 16,       // intra block copy: predictor = the frame being decoded displaced by the block vector, which
                            // cfl_max_w4 / cfl_max_h4 hold (column, row as int16, 1/8 luma sample; spec 7.11.3.2 with use_intrabc)
    TXM_INTER = 255,        // no intra prediction: residual is added to the inter predictor
};
enum : uint8_t {
    TXF_HAVE_LEFT = 1, TXF_HAVE_ABOVE = 2, TXF_HAVE_ABOVE_RIGHT = 4, TXF_HAVE_BELOW_LEFT = 8,
    TXF_SMOOTH_EDGE = 16,   // get_filter_type(): a neighbour uses a smooth predictor
    TXF_LOSSLESS = 32,
    TXF_SB_FIRST = 64,      // first record of a superblock (wavefront bookkeeping)
    TXF_II = 128,           // inter-intra: (with an intra mode) blend the intra predictor of the whole block over the inter
                            // predictor already in the frame, cfl_alpha = ii_pack(); (with TXM_INTER) residual of such a block
};
// inter-intra parameters packed into TxRec::cfl_alpha: bit 0 wedge_interintra, bits 1..4 wedge_index, bits 5..6 interintra_mode,
// bits 7..11 MiSize
static inline int16_t ii_pack(int wedge, int wedge_index, int ii_mode, int bsize) {
    return (int16_t)(wedge | (wedge_index << 1) | (ii_mode << 5) | (bsize << 7));
}

// coefficient token: bits 0..9 position (row * min(txw,32) + col), bits 10..31 signed level
static inline uint32_t coef_token(int pos, int level) { return ((uint32_t)level << 10) | (uint32_t)pos; }
static inline int coef_token_pos(uint32_t t) { return (int)(t & 1023); }
static inline int coef_token_level(uint32_t t) { return (int32_t)t >> 10; }

// Per-superblock slice of the TxRec list (wavefront unit of the intra kernel).
// (a superblock is split into its 64x64 luma units: the unit is what the intra kernel keeps on chip)
struct SbRange {
    uint32_t first, count;  // TxRec indices of this 64x64 unit
    uint16_t sb_row, sb_col;  // in superblock units, absolute in the frame
    uint16_t tile_sb_col0;    // first superblock column of the tile (wavefront does not cross tiles)
    uint16_t tile_sb_col1;    // one past the last
    uint16_t tile_sb_row0;
    uint16_t ux, uy;          // unit position in 64-luma-sample units
    uint16_t pad[5];          // 32 bytes: the unit table of a work item is fetched with one bulk copy (16-byte granules)
};
static_assert(sizeof(SbRange) == 32, "SbRange must stay 32 bytes");

// One 64x64 luma unit that holds K3 (intra / palette / inter-intra) records: the intra kernel keeps the unit on chip and
// hands finished units to its neighbours.  Units are listed in wavefront order (superblock anti-diagonal, then decode order),
// so every dependency has a lower table index.
struct K3Unit {
    uint32_t first, count;  // positions in the K3 order list
    uint16_t ux, uy;        // unit position in 64-luma-sample units
    int32_t dep[5];         // table indices of the left, below-left, above-left, above, above-right units to wait for (-1: none)
};
static_assert(sizeof(K3Unit) == 32, "K3Unit must stay 32 bytes");
static constexpr int K3_UNIT_MAX_RECS = 640;    // per-record barrier capacity of the intra kernel per unit (4:2:0 worst case is 576)

// Per-4x4 loop-filter description, one byte pair per plane 4x4 unit and direction:
//   len: 0 = no edge here, else filter length 4 / 6 / 8 / 14 (13 for luma-wide is stored as 14)
//   lvl: filter level 0..63 to use for this edge
struct LfEdge {
    uint8_t len_v, lvl_v, len_h, lvl_h;
};
// What the deblocking edge classification needs of one block (spec 7.14.2 - 7.14.5).  With device-side classification the host ships
// this list (12 B per block) plus the per-4x4 transform-size maps instead of one LfEdge per 4x4 cell of every plane.
struct LfBlk {
    uint16_t mi_row, mi_col;
    uint8_t bsize;
    uint8_t filt_inside;       // !skip || intra: transform edges inside the block are filtered too
    uint8_t lvl[4];            // filter level per edge class: Y vertical, Y horizontal, U, V
    uint16_t pad;
};
static_assert(sizeof(LfBlk) == 12, "LfBlk must stay 12 bytes");

// Frame-level parameters every kernel needs (kept small; passed by value or via constant memory).
struct FrameParams {
    int32_t w[3], h[3];            // visible plane sizes (upscaled == coded here; superres handled separately)
    int32_t cw[3], ch[3];          // coded plane sizes rounded up to 8 luma samples (MiCols*4, MiRows*4)
    int32_t bd, subx, suby, mono;
    int32_t mi_cols, mi_rows, sb128;
    int32_t dq_dc[3], dq_ac[3];    // per-plane qindex deltas (DeltaQYDc ... DeltaQVAc)
    int32_t enable_edge_filter;
    int32_t lf_sharpness;
    int32_t cdef_damping, cdef_bits;
    int32_t cdef_y_pri[8], cdef_y_sec[8], cdef_uv_pri[8], cdef_uv_sec[8];
    int32_t lr_type[3], lr_size[3];
};

}  // namespace av1r
