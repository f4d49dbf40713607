// Launch descriptors of the intra wavefront kernel (K3).
#pragma once
#include <cuda_runtime.h>

#include "devframe.h"

struct av1r_film_grain_params;
namespace av1r {

struct IntraLaunch {            // passed by value
    const TxRec* recs;          // device, all records of the frame
    const uint32_t* order;      // device, K3 order: indices of the records K3 owns (intra, palette, inter-intra blend + residual), grouped by unit
    int n;                      // entries in order
    const K3Unit* units;        // device, unit table in wavefront order
    int n_units;
    int ctas;                   // persistent CTAs this frame may occupy (one unit per CTA at a time)
    int load_tile;              // 1: the frame already holds inter-predicted samples (inter frame) -> bring the unit in before predicting
    int progressive;            // 1: units hand their bottom row / right column over cell by cell (uprog), 0: whole units (uflags)
    int kind;                   // record kinds the frame may hold: 0 camera-content intra frame, 1 inter frame without screen-content tools
                                // (adds inter-intra blends + their residuals), 2 anything (adds palette and intra block copy)
    unsigned long long* uprog;  // device, n_units words, zeroed before launch: finished border cells of every unit (bit layout in intra.cu)
    int* uflags;                // device, n_units ints, zeroed before launch: unit done flags
    const int32_t* upos;        // device, one entry per 64x64 unit of the frame: its table index (frames with intra block copy; else null)
    int* ticket;                // device, one int, zeroed before launch
    int* stuck;                 // device, one int shared by every frame of the engine: raised (never cleared by the kernel) when a wait made no progress
    DevPlanes frame;
    DevResidual res;
    DevFrameParams fp;
    const uint8_t* wedge_master;   // device, 6 x 64 x 64 (inter-intra wedge blends)
    const uint8_t* pal;            // device, palette entries (colours + colour index maps)
    unsigned long long* prof;      // device, 16 cycle counters (AV1R_K3_PROF=1) or null
};

cudaError_t launch_intra(const IntraLaunch& L, cudaStream_t s);
cudaError_t launch_itx(const TxRec* recs, const uint32_t* order, int n, int n_small, const uint32_t* coefs, const DevResidual& res,
                       const DevFrameParams& fp, cudaStream_t s);

struct LfLaunch {
    DevPlanes frame;
    const LfEdge* edges[3];    // device, per plane [ph4][pw4]
    DevFrameParams fp;
    int plane_on[3];
};
cudaError_t launch_deblock(const LfLaunch& L, cudaStream_t s);
// Device-side edge classification (spec 7.14.2 - 7.14.5): scatter the block list into a per-mi block index, then one thread per
// 4x4 cell and plane derives filter length and level of its left / top edge.  Same result as the host's build_loopfilter_edges.
struct LfClassify {
    const LfBlk* blks;         // device
    int n_blks;
    uint32_t* mi_blk;          // device scratch, [mi_rows][mi_cols]: index into blks
    const uint8_t* lf_tx[3];   // device, per plane [ph4][pw4]: transform size of the cell
    LfEdge* edges[3];          // device, out
    int plane_on[3];
    DevFrameParams fp;
};
cudaError_t launch_lf_classify(const LfClassify& L, cudaStream_t s);

struct CdefLaunch {
    DevPlanes src, dst;
    const int8_t* cdef_idx;    // device, per 64x64
    const uint8_t* skip_mi;    // device, per mi
    DevFrameParams fp;
};
cudaError_t launch_cdef(const CdefLaunch& L, cudaStream_t s);

enum { RESTORE_NONE_D = 0, RESTORE_WIENER_D = 1, RESTORE_SGRPROJ_D = 2 };
struct LrUnitDev {             // same layout as the host LrUnit (frame_state.h)
    uint8_t type, sgr_set;
    int8_t wiener[2][3];
    int8_t sgr_xqd[2];
};
struct LrLaunch {
    DevPlanes cdef, deblocked, dst;
    const LrUnitDev* units[3];  // device, [unit_rows][unit_cols]
    int lr_type[3], unit_size[3], unit_rows[3], unit_cols[3];
    DevFrameParams fp;
};
cudaError_t launch_lr(const LrLaunch& L, cudaStream_t s);

// K6 super-resolution: stretches `src` (coded width) to `dst` (upscaled width), spec 7.16
struct SuperresLaunch {
    DevPlanes src, dst;
    int planes, bd;
    int up_w[3], h[3];            // visible size of the upscaled planes
    int src_cw[3];                // coded width of the source planes ((MiCols >> subX) * 4): the clamp range of the taps
    int step_x[3], initial_subpel_x[3];
};
cudaError_t launch_superres(const SuperresLaunch& L, cudaStream_t s);

cudaError_t launch_plane_checksum(const void* src, size_t pitch, int w, int h, int bpc, uint64_t* out_dev, cudaStream_t s);
// K8 film grain in two halves (filmgrain.cu): template preparation (depends on the header only) and application
int fg_launch_prepare(const struct ::av1r_film_grain_params* p, int bpc, int w, int h, int subx, int suby, int mono, int mc_identity, void* scratch,
                      cudaStream_t s);
int fg_launch_apply(const struct ::av1r_film_grain_params* p, int bpc, int w, int h, int subx, int suby, int mono, int mc_identity,
                    const void* const src[3], const size_t src_pitch[3], void* const dst[3], const size_t dst_pitch[3], void* scratch, cudaStream_t s);
cudaError_t launch_frame_checksum(const void* const src[3], const size_t pitch[3], const int w[3], const int h[3], int nplanes, int bpc,
                                  uint64_t* out_dev, cudaStream_t s);

}  // namespace av1r
