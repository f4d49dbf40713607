// TEST INFRASTRUCTURE ONLY -- frame containers of the CPU oracle (always 16-bit samples).
#pragma once
#include <stdint.h>

#include <memory>
#include <vector>

#include "../av1-go_b200/csrc/frame_state.h"

namespace orc {

struct Plane {
    std::vector<uint16_t> d;
    int w = 0, h = 0, stride = 0;   // coded size (multiple of 8 luma samples)
    // 64 samples of margin to the right and below: a transform block that straddles the coded frame edge is reconstructed in
    // full (spec 7.11.2 / 7.13.3 write the whole block), and chroma-from-luma reads those luma samples (MaxLumaW / MaxLumaH)
    void alloc(int w_, int h_) {
        w = w_;
        h = h_;
        stride = w_ + 64;
        d.assign((size_t)stride * (h_ + 64), 0);
    }
    uint16_t& at(int x, int y) { return d[(size_t)y * stride + x]; }
    const uint16_t& at(int x, int y) const { return d[(size_t)y * stride + x]; }
};

struct FrameGeom {
    int bd, subx, suby, mono;
    int w[3], h[3];     // visible
    int cw[3], ch[3];   // coded (MiCols*4 >> subx ...)
    int dq_dc[3], dq_ac[3];
    int enable_edge_filter;
};

struct Frame {
    FrameGeom g;
    Plane p[3];
};

void reconstruct_frame(const av1r::FrameWork& fw, Frame& f, const Frame* const refs[8]);
void deblock_frame(const av1r::FrameWork& fw, Frame& f);
void cdef_frame(const av1r::FrameWork& fw, const Frame& in, Frame& out);
// super-resolution (spec 7.16): `out` must be allocated at the upscaled geometry
void upscale_frame(const av1r::FrameHdr& fh, const Frame& in, Frame& out);
void lr_frame(const av1r::FrameWork& fw, const Frame& deblocked, const Frame& cdef, Frame& out);

}  // namespace orc
