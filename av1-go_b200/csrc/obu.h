// OBU layer + sequence header + uncompressed frame header parser (host, sequential).
#pragma once
#include <string>
#include <vector>

#include "bits.h"
#include "hdr.h"

namespace av1r {

// Header-level state saved per reference slot (spec 7.20 reference frame update process).
struct RefHdrState {
    int valid = 0;
    int frame_id = 0, upscaled_width = 0, frame_width = 0, frame_height = 0, render_width = 0, render_height = 0;
    int mi_cols = 0, mi_rows = 0, frame_type = 0, order_hint = 0, bit_depth = 0, subsampling_x = 0, subsampling_y = 0;
    int showable_frame = 0;
    int saved_order_hints[8] = {0};
    int32_t saved_gm_params[8][6];
    int lf_ref_deltas[8], lf_mode_deltas[2];
    SegmentationParams seg;
    FilmGrainParams fg;
};

struct ObuUnit {
    int type, temporal_id, spatial_id;
    const uint8_t* data;   // payload
    size_t size;
};

struct TileGroupInfo {
    int tg_start, tg_end;
    size_t data_offset;    // offset of first tile's (size-prefixed) data inside payload
};

class HeaderParser {
public:
    SeqHdr seq;
    RefHdrState refs[NUM_REF_FRAMES];
    std::string error;
    int32_t prev_gm_params[8][6];

    HeaderParser();
    // Split a temporal unit into OBUs (has_size_field required, as in IVF/Section-5 streams).
    bool split_obus(const uint8_t* data, size_t len, std::vector<ObuUnit>& out);
    bool parse_sequence_header(const uint8_t* d, size_t n);
    // Parses the uncompressed header; on return br is positioned after it (not byte aligned).
    bool parse_frame_header(BitReader& br, FrameHdr& fh, int temporal_id, int spatial_id);
    bool parse_tile_group_header(BitReader& br, const FrameHdr& fh, TileGroupInfo& tg);
    // spec 7.20: store header-level state of the just-decoded frame into refreshed slots.
    void reference_update(const FrameHdr& fh);
    // spec 7.9.3 get_relative_dist (inline: the tile parser calls it for every temporal / reference MV candidate)
    int get_relative_dist(int a, int b) const {
        if (!seq.enable_order_hint) return 0;
        const int diff = a - b, m = 1 << (seq.order_hint_bits - 1);
        return (diff & (m - 1)) - (diff & m);
    }

private:
    bool fail(const char* msg) { error = msg; return false; }
    void setup_past_independence(FrameHdr& fh);
    void load_previous(FrameHdr& fh);
    bool frame_size(BitReader& br, FrameHdr& fh);
    void superres_params(BitReader& br, FrameHdr& fh);
    void compute_image_size(FrameHdr& fh);
    void render_size(BitReader& br, FrameHdr& fh);
    bool frame_size_with_refs(BitReader& br, FrameHdr& fh);
    void set_frame_refs(FrameHdr& fh, int last_frame_idx, int gold_frame_idx);
    bool tile_info(BitReader& br, FrameHdr& fh);
    void quantization_params(BitReader& br, FrameHdr& fh);
    void segmentation_params(BitReader& br, FrameHdr& fh);
    void loop_filter_params(BitReader& br, FrameHdr& fh);
    void cdef_params(BitReader& br, FrameHdr& fh);
    void lr_params(BitReader& br, FrameHdr& fh);
    void skip_mode_params(BitReader& br, FrameHdr& fh);
    void global_motion_params(BitReader& br, FrameHdr& fh);
    void film_grain_params(BitReader& br, FrameHdr& fh);
    int read_delta_q(BitReader& br);
    void read_global_param(BitReader& br, FrameHdr& fh, int type, int ref, int idx);
};

int get_qidx(const FrameHdr& fh, int ignore_deltas, int segment_id, int current_q_index);

}  // namespace av1r
