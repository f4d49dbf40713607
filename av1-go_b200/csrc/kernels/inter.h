// Launch descriptors of the inter prediction kernels (K2).
#pragma once
#include <cuda_runtime.h>

#include "devframe.h"

namespace av1r {

struct InterLaunch {
    const InterBlk* blks;      // device, n records
    const ObmcNb* obmc;        // device
    const WarpRec* warps;      // device
    int n;
    const uint32_t* tiles;     // device, n_tiles work items: record index | quadrant << 28 (one CTA per 64x64 luma quadrant of a record)
    int n_tiles;
    int n_tiles_small;         // the first n_tiles_small items belong to blocks of at most 16x16 luma samples
    DevPlanes refs[8];         // reference slots
    int ref_w[8][3], ref_h[8][3];   // visible size of each reference plane (clamp range of the taps)
    int xscale[8], yscale[8];  // spec 7.11.3.3 scale factors of each slot against this frame (1 << 14: same size)
    DevPlanes cur;             // frame being reconstructed
    uint8_t* mask;             // device, luma-sized byte plane: difference-weighted compound masks (COMPOUND_DIFFWTD blocks only)
    uint32_t mask_pitch;
    DevFrameParams fp;
};

cudaError_t launch_inter(const InterLaunch& L, cudaStream_t s);
// frame += residual for the plain inter transform blocks (order = indices of records with eob > 0)
cudaError_t launch_inter_residual(const TxRec* recs, const uint32_t* order, int n, const DevPlanes& cur, const DevResidual& res,
                                  const DevFrameParams& fp, cudaStream_t s);
// wedge master masks (6 x 64 x 64 bytes) for kernels outside inter.cu
cudaError_t inter_copy_wedge_master(uint8_t* dst_dev, cudaStream_t s);

}  // namespace av1r
