// Header-only scans: av1r_scan_headers / av1r_probe_* (no GPU).  Native stand-in for the
// `ffprobe -show_streams` child of /root/reference/internal/metadata/probe.go:145-153: fills the
// fields the daemon reads (Width/Height/BitDepth/CodecName -> HasAV1, probe.go:34-46,179).
#include <cstring>
#include <vector>

#include "../../include/av1r.h"
#include "../../include/av1r_stages.h"
#include "demux.h"
#include "obu.h"

using namespace av1r;

static_assert(sizeof(av1r_film_grain_params) == sizeof(FilmGrainParams), "film grain param layouts must match");

static void fill_info(const HeaderParser& hp, const FrameHdr& fh, int tu, av1r_frame_header_info& o) {
    memset(&o, 0, sizeof(o));
    o.struct_size = sizeof(o);
    o.tu_index = tu;
    o.frame_type = fh.frame_type;
    o.show_frame = fh.show_frame;
    o.showable_frame = fh.showable_frame;
    o.show_existing_frame = fh.show_existing_frame;
    o.frame_to_show_map_idx = fh.frame_to_show_map_idx;
    o.width = fh.frame_width;
    o.height = fh.frame_height;
    o.upscaled_width = fh.upscaled_width;
    o.bit_depth = hp.seq.bit_depth;
    o.subsampling_x = hp.seq.subsampling_x;
    o.subsampling_y = hp.seq.subsampling_y;
    o.mono_chrome = hp.seq.mono_chrome;
    o.matrix_coefficients = hp.seq.matrix_coefficients;
    o.refresh_frame_flags = fh.refresh_frame_flags;
    o.order_hint = fh.order_hint;
    o.primary_ref_frame = fh.primary_ref_frame;
    o.base_q_idx = fh.base_q_idx;
    o.tile_cols = fh.tile_cols;
    o.tile_rows = fh.tile_rows;
    o.use_128x128_superblock = hp.seq.use_128x128_superblock;
    for (int i = 0; i < 4; i++) o.lf_level[i] = fh.lf.level[i];
    o.cdef_enabled = fh.enable_cdef_frame;
    o.cdef_bits = fh.cdef_bits;
    for (int i = 0; i < 3; i++) o.lr_type[i] = fh.lr_type[i];
    o.tx_mode = fh.tx_mode;
    o.reduced_tx_set = fh.reduced_tx_set;
    o.header_bytes = (int)fh.header_bytes;
    memcpy(&o.film_grain, &fh.fg, sizeof(o.film_grain));
    o.error_resilient_mode = fh.error_resilient_mode;
    o.disable_cdf_update = fh.disable_cdf_update;
    o.disable_frame_end_update_cdf = fh.disable_frame_end_update_cdf;
    o.enable_order_hint = hp.seq.enable_order_hint;
    o.coded_lossless = fh.coded_lossless;
    o.segmentation_enabled = fh.seg.enabled;
    o.segmentation_update_map = fh.seg.update_map;
    o.segmentation_temporal_update = fh.seg.temporal_update;
    o.delta_q_present = fh.delta_q_present;
    o.delta_lf_present = fh.delta_lf_present;
}

// Walk one temporal unit; calls cb for every frame header (incl. show_existing_frame).
template <typename F>
static int scan_tu(HeaderParser& hp, const uint8_t* data, size_t len, F&& cb) {
    std::vector<ObuUnit> obus;
    if (!hp.split_obus(data, len, obus)) return AV1R_EBITSTREAM;
    bool seen_frame_header = false;
    FrameHdr fh;
    for (const ObuUnit& u : obus) {
        if (u.type == OBU_SEQUENCE_HEADER) {
            if (!hp.parse_sequence_header(u.data, u.size)) return AV1R_EBITSTREAM;
        } else if (u.type == OBU_TEMPORAL_DELIMITER) {
            seen_frame_header = false;
        } else if (u.type == OBU_FRAME_HEADER || u.type == OBU_FRAME || u.type == OBU_REDUNDANT_FRAME_HEADER) {
            if (seen_frame_header && u.type != OBU_FRAME) continue;  // redundant copy
            BitReader br(u.data, u.size);
            if (!hp.parse_frame_header(br, fh, u.temporal_id, u.spatial_id)) return AV1R_EBITSTREAM;
            br.byte_align();
            fh.header_bytes = br.byte_pos();
            cb(fh);
            if (fh.show_existing_frame) {
                if (fh.frame_type == KEY_FRAME) {
                    // frame loading process: the shown key frame refreshes every slot
                    RefHdrState r = hp.refs[fh.frame_to_show_map_idx];
                    for (int i = 0; i < NUM_REF_FRAMES; i++) hp.refs[i] = r;
                }
                seen_frame_header = false;
            } else {
                hp.reference_update(fh);
                seen_frame_header = (u.type == OBU_FRAME_HEADER);
                // a frame header OBU stays "active" until its tile groups are done; for a header
                // scan we only need to skip redundant copies, which the flag above does.
                if (u.type == OBU_FRAME) seen_frame_header = false;
            }
        } else if (u.type == OBU_TILE_GROUP) {
            seen_frame_header = false;  // conservative: next FRAME_HEADER is a new frame
        }
    }
    return 0;
}

extern "C" int av1r_scan_headers(const uint8_t* const* tus, const size_t* lens, int n_tus, av1r_frame_header_info* out,
                                 int cap, int* n) {
    if (!tus || !lens || !n) return AV1R_EINVAL;
    HeaderParser hp;
    int cnt = 0;
    for (int t = 0; t < n_tus; t++) {
        int rc = scan_tu(hp, tus[t], lens[t], [&](const FrameHdr& fh) {
            if (out && cnt < cap) fill_info(hp, fh, t, out[cnt]);
            cnt++;
        });
        if (rc) { *n = cnt; return rc; }
    }
    *n = cnt;
    return 0;
}

extern "C" int av1r_probe_buffer(const uint8_t* data, size_t len, av1r_stream_info* out) {
    if (!data || !out) return AV1R_EINVAL;
    uint32_t ss = out->struct_size ? out->struct_size : sizeof(*out);
    memset(out, 0, sizeof(*out));
    out->struct_size = ss;
    DemuxResult dm;
    std::string err;
    if (!demux_buffer(data, len, dm, err)) return AV1R_EBITSTREAM;
    HeaderParser hp;
    if (!dm.config_obus.empty()) {
        std::vector<ObuUnit> obus;
        if (hp.split_obus(dm.config_obus.data(), dm.config_obus.size(), obus))
            for (auto& u : obus)
                if (u.type == OBU_SEQUENCE_HEADER) hp.parse_sequence_header(u.data, u.size);
    }
    out->temporal_units = (int64_t)dm.tus.size();
    bool have = false;
    for (const TemporalUnit& tu : dm.tus) {
        int rc = scan_tu(hp, data + tu.offset, tu.size, [&](const FrameHdr& fh) {
            if (!have) {
                have = true;
                out->width = fh.upscaled_width;
                out->height = fh.frame_height;
            }
            if (fh.frame_type == KEY_FRAME && fh.show_frame && !fh.show_existing_frame) out->keyframes++;
        });
        if (rc) return rc;
    }
    if (!hp.seq.valid) return AV1R_EBITSTREAM;
    out->is_av1 = 1;
    out->bit_depth = hp.seq.bit_depth;
    out->profile = hp.seq.profile;
    out->subsampling_x = hp.seq.subsampling_x;
    out->subsampling_y = hp.seq.subsampling_y;
    out->mono_chrome = hp.seq.mono_chrome;
    out->film_grain_present = hp.seq.film_grain_params_present;
    return 0;
}

extern "C" int av1r_probe_file(const char* path, av1r_stream_info* out) {
    if (!path || !out) return AV1R_EINVAL;
    DemuxResult dm;
    std::string err;
    FILE* f = fopen(path, "rb");
    if (!f) return AV1R_ENOENT;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> buf(n > 0 ? n : 0);
    size_t got = n > 0 ? fread(buf.data(), 1, n, f) : 0;
    fclose(f);
    if ((long)got != n) return AV1R_EIO;
    return av1r_probe_buffer(buf.data(), buf.size(), out);
}
