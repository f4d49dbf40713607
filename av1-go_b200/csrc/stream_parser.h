// Stream-level host driver: temporal unit -> OBUs -> frame headers -> per-tile parse ->
// FrameWork (device work-lists) + reference bookkeeping for everything the *parse* needs from
// earlier frames (CDFs, segment maps, motion fields).  No pixel work happens here.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "frame_state.h"
#include "obu.h"

namespace av1r {

struct ParsedFrame {
    std::shared_ptr<FrameWork> fw;   // null when show_existing_slot >= 0
    int show_existing_slot = -1;     // show_existing_frame: output the frame stored in this slot
    FrameHdr fh;
    SeqHdr seq;                      // sequence header in force for this frame (geometry / bit depth of show_existing outputs, grain matrix)
    int64_t pts = 0;
    std::shared_ptr<void> host;      // engine-side staging (pinned work-list arena), filled by the thread that parsed the frame
};

class StreamParser {
public:
    StreamParser();
    HeaderParser hp;
    std::string err;
    bool tile_threads = true;   // parse the tiles of a tile group concurrently on the process-wide worker pool
    // true: parse_tu returns frames whose work-lists are still per tile; the caller runs finalize_framework(*pf.fw) (on another
    // thread, typically) before it reads fw.tx / coefs / inter / lf.  Everything the parse of later frames needs is complete.
    bool defer_finalize = false;
    // false: the deblocking edges are classified on the device; frames carry FrameWork::lf_blocks (+ lf_tx) instead of lf[]
    bool host_lf_edges = true;
    // Parses one temporal unit; appends one ParsedFrame per frame (shown or not) in decode order.
    // Returns 0 or AV1R_E*.
    int parse_tu(const uint8_t* data, size_t len, int64_t pts, std::vector<ParsedFrame>& out);

private:
    CdfCtx slot_cdf_[NUM_REF_FRAMES];
    std::shared_ptr<FrameWork> slot_fw_[NUM_REF_FRAMES];   // parse-side state of the frame in each slot
    // frame being assembled (frame header seen, tile groups arriving)
    std::shared_ptr<FrameWork> cur_;
    FrameHdr cur_fh_;
    CdfCtx cur_init_cdf_;
    int tiles_done_ = 0;
    bool have_frame_ = false;
    std::vector<uint8_t> cur_hdr_bytes_;   // bytes of the active frame header (to recognise its redundant copies)
    int fail(int code, const std::string& m) { err = m; return code; }
    int parse_tu_inner(const uint8_t* data, size_t len, int64_t pts, std::vector<ParsedFrame>& out);
    int begin_frame(const FrameHdr& fh);
    void motion_field_estimation();
    int tile_group(const uint8_t* payload, size_t size, size_t offset);
    int finish_frame(int64_t pts, std::vector<ParsedFrame>& out);
};

std::shared_ptr<FrameWork> acquire_framework();
void cdf_clear_counters(CdfCtx& c);
// host pre-pass of the deblocking filter: per-4x4 edge length + level (spec 7.14.2 - 7.14.5)
void build_loopfilter_edges(FrameWork& fw);
// merge of the per-tile work-lists + deblocking edge classification (idempotent); see StreamParser::defer_finalize
void finalize_framework(FrameWork& fw);

}  // namespace av1r
