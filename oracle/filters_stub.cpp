// TEST INFRASTRUCTURE ONLY -- placeholders until the in-loop filter restatements land.
#include "oracle_frame.h"
namespace orc {
void deblock_frame(const av1r::FrameWork&, Frame&) {}
void cdef_frame(const av1r::FrameWork&, const Frame& in, Frame& out) { out = in; }
void lr_frame(const av1r::FrameWork&, const Frame&, const Frame& cdef, Frame& out) { out = cdef; }
}
