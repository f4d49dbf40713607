// Device-side frame / work-list descriptors shared by the reconstruction kernels and the engine.
#pragma once
#include <stdint.h>

#include "../worklist.h"

namespace av1r {

struct DevPlanes {
    uint8_t* p[3];          // device pointers
    uint32_t pitch[3];      // bytes
};

struct DevResidual {        // int16 residual planes (K1 output, consumed by K2/K3)
    int16_t* p[3];
    uint32_t pitch[3];      // bytes
};

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
