// Device-side frame / work-list descriptors shared by the reconstruction kernels and the engine.
#pragma once
#include <stdint.h>

#include "../worklist.h"

namespace av1r {

struct DevPlanes {
    uint8_t* p[3];          // device pointers
    uint32_t pitch[3];      // bytes
};

// int16 residual (K1 output, consumed by K2b/K3), stored *unit-major*: each 64x64 luma unit holds its Y tile followed by
// its U and V tiles contiguously (12 KB at 4:2:0), so that K3 brings a whole unit on chip with one bulk copy (TMA).
struct DevResidual {
    int16_t* base;
    int32_t units_x;        // units per row
    int32_t unit_elems;     // int16 elements per unit
    int32_t plane_off[3];   // element offset of each plane's tile inside a unit
    int32_t tw_log2[3], th_log2[3];   // tile size of each plane (6,6 / 5,5 at 4:2:0)
};
#ifdef __CUDACC__
// pointer to sample (x, y) of `plane`; rows of the same tile are (1 << tw_log2[plane]) elements apart
__device__ __forceinline__ int16_t* res_ptr(const DevResidual& r, int plane, int x, int y) {
    const int lw = r.tw_log2[plane], lh = r.th_log2[plane];
    const size_t unit = (size_t)(y >> lh) * r.units_x + (x >> lw);
    return r.base + unit * r.unit_elems + r.plane_off[plane] + ((y & ((1 << lh) - 1)) << lw) + (x & ((1 << lw) - 1));
}
#endif

// Per-frame kernel parameter block (passed by value; ~300 bytes)
struct DevFrameParams {
    int32_t cw[3], ch[3];   // coded plane sizes (MiCols*4 >> subx, ...)
    int32_t w[3], h[3];     // visible plane sizes
    int32_t bd, subx, suby, mono;
    int32_t mi_cols, mi_rows, sb128;
    int32_t dq_dc[3], dq_ac[3];
    int32_t enable_edge_filter;
    int32_t lf_sharpness;
    int32_t cdef_damping;
    int32_t cdef_y_pri[8], cdef_y_sec[8], cdef_uv_pri[8], cdef_uv_sec[8];
    int32_t pw4[3], ph4[3];
};

}  // namespace av1r
