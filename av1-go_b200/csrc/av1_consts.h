// AV1 enums and small derived tables (spec section 3 "Symbols" + section 9.3 conversion tables).
// Everything here is either an enum value fixed by the bitstream or computed from block
// geometry; the large normative tables live under tables/*.inc.
#pragma once
#include <cstdint>

namespace av1r {

enum BlockSize : uint8_t {
    BLOCK_4X4, BLOCK_4X8, BLOCK_8X4, BLOCK_8X8, BLOCK_8X16, BLOCK_16X8, BLOCK_16X16, BLOCK_16X32, BLOCK_32X16,
    BLOCK_32X32, BLOCK_32X64, BLOCK_64X32, BLOCK_64X64, BLOCK_64X128, BLOCK_128X64, BLOCK_128X128, BLOCK_4X16,
    BLOCK_16X4, BLOCK_8X32, BLOCK_32X8, BLOCK_16X64, BLOCK_64X16, BLOCK_SIZES_ALL, BLOCK_INVALID = 255
};
enum TxSize : uint8_t {
    TX_4X4, TX_8X8, TX_16X16, TX_32X32, TX_64X64, TX_4X8, TX_8X4, TX_8X16, TX_16X8, TX_16X32, TX_32X16, TX_32X64,
    TX_64X32, TX_4X16, TX_16X4, TX_8X32, TX_32X8, TX_16X64, TX_64X16, TX_SIZES_ALL
};
enum TxType : uint8_t {
    DCT_DCT, ADST_DCT, DCT_ADST, ADST_ADST, FLIPADST_DCT, DCT_FLIPADST, FLIPADST_FLIPADST, ADST_FLIPADST,
    FLIPADST_ADST, IDTX, V_DCT, H_DCT, V_ADST, H_ADST, V_FLIPADST, H_FLIPADST, TX_TYPES, WHT_WHT = 16
};
enum { TX_CLASS_2D = 0, TX_CLASS_HORIZ = 1, TX_CLASS_VERT = 2 };
enum Partition : uint8_t {
    PARTITION_NONE, PARTITION_HORZ, PARTITION_VERT, PARTITION_SPLIT, PARTITION_HORZ_A, PARTITION_HORZ_B,
    PARTITION_VERT_A, PARTITION_VERT_B, PARTITION_HORZ_4, PARTITION_VERT_4
};
enum PredMode : uint8_t {
    DC_PRED, V_PRED, H_PRED, D45_PRED, D135_PRED, D113_PRED, D157_PRED, D203_PRED, D67_PRED, SMOOTH_PRED,
    SMOOTH_V_PRED, SMOOTH_H_PRED, PAETH_PRED, UV_CFL_PRED, INTRA_MODES = 13,
    NEARESTMV = 13, NEARMV, GLOBALMV, NEWMV, NEAREST_NEARESTMV, NEAR_NEARMV, NEAREST_NEWMV, NEW_NEARESTMV,
    NEAR_NEWMV, NEW_NEARMV, GLOBAL_GLOBALMV, NEW_NEWMV
};
enum { SIMPLE_TRANSLATION = 0, OBMC_CAUSAL = 1, WARPED_CAUSAL = 2 };
enum { COMPOUND_WEDGE = 0, COMPOUND_DIFFWTD = 1, COMPOUND_AVERAGE = 2, COMPOUND_INTRA = 3, COMPOUND_DISTANCE = 4 };
enum { II_DC_PRED, II_V_PRED, II_H_PRED, II_SMOOTH_PRED };

static const uint8_t kBlockW[BLOCK_SIZES_ALL] = {4, 4, 8, 8, 8, 16, 16, 16, 32, 32, 32, 64, 64, 64, 128, 128, 4, 16, 8, 32, 16, 64};
static const uint8_t kBlockH[BLOCK_SIZES_ALL] = {4, 8, 4, 8, 16, 8, 16, 32, 16, 32, 64, 32, 64, 128, 64, 128, 16, 4, 32, 8, 64, 16};
static const uint8_t kBlockW4[BLOCK_SIZES_ALL] = {1, 1, 2, 2, 2, 4, 4, 4, 8, 8, 8, 16, 16, 16, 32, 32, 1, 4, 2, 8, 4, 16};
static const uint8_t kBlockH4[BLOCK_SIZES_ALL] = {1, 2, 1, 2, 4, 2, 4, 8, 4, 8, 16, 8, 16, 32, 16, 32, 4, 1, 8, 2, 16, 4};
static const uint8_t kBlockWLog2[BLOCK_SIZES_ALL] = {2, 2, 3, 3, 3, 4, 4, 4, 5, 5, 5, 6, 6, 6, 7, 7, 2, 4, 3, 5, 4, 6};
static const uint8_t kBlockHLog2[BLOCK_SIZES_ALL] = {2, 3, 2, 3, 4, 3, 4, 5, 4, 5, 6, 5, 6, 7, 6, 7, 4, 2, 5, 3, 6, 4};

static const uint8_t kTxW[TX_SIZES_ALL] = {4, 8, 16, 32, 64, 4, 8, 8, 16, 16, 32, 32, 64, 4, 16, 8, 32, 16, 64};
static const uint8_t kTxH[TX_SIZES_ALL] = {4, 8, 16, 32, 64, 8, 4, 16, 8, 32, 16, 64, 32, 16, 4, 32, 8, 64, 16};
static const uint8_t kTxWLog2[TX_SIZES_ALL] = {2, 3, 4, 5, 6, 2, 3, 3, 4, 4, 5, 5, 6, 2, 4, 3, 5, 4, 6};
static const uint8_t kTxHLog2[TX_SIZES_ALL] = {2, 3, 4, 5, 6, 3, 2, 4, 3, 5, 4, 6, 5, 4, 2, 5, 3, 6, 4};
static const uint8_t kTxSqr[TX_SIZES_ALL] = {TX_4X4, TX_8X8, TX_16X16, TX_32X32, TX_64X64, TX_4X4, TX_4X4, TX_8X8, TX_8X8,
                                             TX_16X16, TX_16X16, TX_32X32, TX_32X32, TX_4X4, TX_4X4, TX_8X8, TX_8X8, TX_16X16, TX_16X16};
static const uint8_t kTxSqrUp[TX_SIZES_ALL] = {TX_4X4, TX_8X8, TX_16X16, TX_32X32, TX_64X64, TX_8X8, TX_8X8, TX_16X16, TX_16X16,
                                               TX_32X32, TX_32X32, TX_64X64, TX_64X64, TX_16X16, TX_16X16, TX_32X32, TX_32X32, TX_64X64, TX_64X64};
static const uint8_t kSplitTx[TX_SIZES_ALL] = {TX_4X4, TX_4X4, TX_8X8, TX_16X16, TX_32X32, TX_4X4, TX_4X4, TX_8X8, TX_8X8,
                                               TX_16X16, TX_16X16, TX_32X32, TX_32X32, TX_4X8, TX_8X4, TX_8X16, TX_16X8, TX_16X32, TX_32X16};
static const uint8_t kMaxTxRect[BLOCK_SIZES_ALL] = {TX_4X4, TX_4X8, TX_8X4, TX_8X8, TX_8X16, TX_16X8, TX_16X16, TX_16X32, TX_32X16,
                                                    TX_32X32, TX_32X64, TX_64X32, TX_64X64, TX_64X64, TX_64X64, TX_64X64, TX_4X16,
                                                    TX_16X4, TX_8X32, TX_32X8, TX_16X64, TX_64X16};
static const uint8_t kMaxTxDepth[BLOCK_SIZES_ALL] = {0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4, 4, 4, 4, 2, 2, 3, 3, 4, 4};
// Adjusted_Tx_Size: 64-point dimensions code only the 32x32 low-frequency part
static const uint8_t kAdjTx[TX_SIZES_ALL] = {TX_4X4, TX_8X8, TX_16X16, TX_32X32, TX_32X32, TX_4X8, TX_8X4, TX_8X16, TX_16X8,
                                             TX_16X32, TX_32X16, TX_32X32, TX_32X32, TX_4X16, TX_16X4, TX_8X32, TX_32X8, TX_16X32, TX_32X16};
static const uint8_t kIntraModeCtx[13] = {0, 1, 2, 3, 4, 4, 4, 4, 3, 0, 1, 2, 0};
static const uint8_t kSizeGroup[BLOCK_SIZES_ALL] = {0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 0, 0, 1, 1, 2, 2};
static const uint8_t kModeToTxfm[14] = {DCT_DCT, ADST_DCT, DCT_ADST, DCT_DCT, ADST_ADST, ADST_DCT, DCT_ADST, DCT_ADST, ADST_DCT,
                                        ADST_ADST, ADST_DCT, DCT_ADST, ADST_ADST, DCT_DCT};
static const uint8_t kFilterIntraModeToIntraDir[5] = {DC_PRED, V_PRED, H_PRED, D157_PRED, DC_PRED};
// inverse ext-tx maps (symbol -> TxType) for the reduced sets
static const uint8_t kTxTypeIntraInvSet1[7] = {IDTX, DCT_DCT, V_DCT, H_DCT, ADST_ADST, ADST_DCT, DCT_ADST};
static const uint8_t kTxTypeIntraInvSet2[5] = {IDTX, DCT_DCT, ADST_ADST, ADST_DCT, DCT_ADST};
static const uint8_t kTxTypeInterInvSet1[16] = {IDTX, V_DCT, H_DCT, V_ADST, H_ADST, V_FLIPADST, H_FLIPADST, DCT_DCT, ADST_DCT,
                                                DCT_ADST, FLIPADST_DCT, DCT_FLIPADST, ADST_ADST, FLIPADST_FLIPADST, ADST_FLIPADST, FLIPADST_ADST};
static const uint8_t kTxTypeInterInvSet2[12] = {IDTX, V_DCT, H_DCT, DCT_DCT, ADST_DCT, DCT_ADST, FLIPADST_DCT, DCT_FLIPADST, ADST_ADST,
                                                FLIPADST_FLIPADST, ADST_FLIPADST, FLIPADST_ADST};
static const uint8_t kTxTypeInterInvSet3[2] = {IDTX, DCT_DCT};

static inline BlockSize block_from_wh(int w, int h) {
    for (int i = 0; i < BLOCK_SIZES_ALL; i++)
        if (kBlockW[i] == w && kBlockH[i] == h) return (BlockSize)i;
    return BLOCK_INVALID;
}
static inline BlockSize partition_subsize(int part, BlockSize bs) {
    int d = kBlockW[bs];
    switch (part) {
        case PARTITION_NONE: return bs;
        case PARTITION_HORZ: case PARTITION_HORZ_A: case PARTITION_HORZ_B: return block_from_wh(d, d / 2);
        case PARTITION_VERT: case PARTITION_VERT_A: case PARTITION_VERT_B: return block_from_wh(d / 2, d);
        case PARTITION_SPLIT: return block_from_wh(d / 2, d / 2);
        case PARTITION_HORZ_4: return block_from_wh(d, d / 4);
        case PARTITION_VERT_4: return block_from_wh(d / 4, d);
    }
    return BLOCK_INVALID;
}
static inline BlockSize plane_residual_size(BlockSize bs, int subx, int suby) {
    int w = kBlockW[bs] >> subx, h = kBlockH[bs] >> suby;
    if (w < 4) w = 4;
    if (h < 4) h = 4;
    return block_from_wh(w, h);
}
static inline int tx_class_of(int txtp) {
    switch (txtp) {
        case V_DCT: case V_ADST: case V_FLIPADST: return TX_CLASS_VERT;
        case H_DCT: case H_ADST: case H_FLIPADST: return TX_CLASS_HORIZ;
        default: return TX_CLASS_2D;
    }
}
static inline bool is_directional_mode(int m) { return m >= V_PRED && m <= D67_PRED; }

}  // namespace av1r
