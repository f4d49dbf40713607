#include "stream_parser.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <cstddef>

#include "../../include/av1r.h"
#include "parallel.h"
#include "tile.h"

namespace av1r {

void cdf_clear_counters(CdfCtx& c) {
    // every cdf is an array [nsyms+1]; the counter sits right after the *actual* symbols.
#define CLR(field, ns)                                                                  \
    {                                                                                   \
        uint16_t* p = reinterpret_cast<uint16_t*>(&c.field);                            \
        const size_t stride = sizeof(c.field) / sizeof(uint16_t);                       \
        (void)stride;                                                                   \
        p[ns] = 0;                                                                      \
    }
    uint16_t* base = reinterpret_cast<uint16_t*>(&c);
    auto clr_array = [&](size_t off_bytes, size_t total_bytes, int stride, auto ns_of) {
        uint16_t* p = base + off_bytes / 2;
        size_t n = total_bytes / 2 / stride;
        for (size_t k = 0; k < n; k++) p[k * stride + ns_of(k)] = 0;
    };
#define ARR(field, stride, ns_expr) \
    clr_array(offsetof(CdfCtx, field), sizeof(c.field), stride, [&](size_t k) -> int { (void)k; return ns_expr; })
    ARR(txb_skip, 3, 2); ARR(eob_extra, 3, 2); ARR(dc_sign, 3, 2);
    ARR(eob_pt_16, 6, 5); ARR(eob_pt_32, 7, 6); ARR(eob_pt_64, 8, 7); ARR(eob_pt_128, 9, 8);
    ARR(eob_pt_256, 10, 9); ARR(eob_pt_512, 11, 10); ARR(eob_pt_1024, 12, 11);
    ARR(coeff_base_eob, 4, 3); ARR(coeff_base, 5, 4); ARR(coeff_br, 5, 4);
    ARR(newmv, 3, 2); ARR(zeromv, 3, 2); ARR(refmv, 3, 2); ARR(drl, 3, 2);
    ARR(inter_compound_mode, 9, 8); ARR(compound_type, 3, 2); ARR(wedge_idx, 17, 16);
    ARR(interintra, 3, 2); ARR(wedge_interintra, 3, 2); ARR(interintra_mode, 5, 4);
    ARR(motion_mode, 4, 3); ARR(obmc, 3, 2);
    ARR(palette_y_size, 8, 7); ARR(palette_uv_size, 8, 7);
    ARR(palette_y_color_index, 9, (int)(k / 5) + 2); ARR(palette_uv_color_index, 9, (int)(k / 5) + 2);
    ARR(palette_y_mode, 3, 2); ARR(palette_uv_mode, 3, 2);
    ARR(comp_inter, 3, 2); ARR(single_ref, 3, 2); ARR(comp_ref_type, 3, 2); ARR(uni_comp_ref, 3, 2);
    ARR(comp_ref, 3, 2); ARR(comp_bwdref, 3, 2); ARR(txfm_partition, 3, 2); ARR(compound_index, 3, 2);
    ARR(comp_group_idx, 3, 2); ARR(skip_mode, 3, 2); ARR(skip, 3, 2); ARR(intra_inter, 3, 2);
#define NMV(pfx)                                                                                               \
    ARR(pfx##joints, 5, 4);                                                                                    \
    ARR(pfx##c0_classes, 12, 11); ARR(pfx##c0_class0_fp, 5, 4); ARR(pfx##c0_fp, 5, 4); ARR(pfx##c0_sign, 3, 2); \
    ARR(pfx##c0_class0_hp, 3, 2); ARR(pfx##c0_hp, 3, 2); ARR(pfx##c0_class0, 3, 2); ARR(pfx##c0_bits, 3, 2);    \
    ARR(pfx##c1_classes, 12, 11); ARR(pfx##c1_class0_fp, 5, 4); ARR(pfx##c1_fp, 5, 4); ARR(pfx##c1_sign, 3, 2); \
    ARR(pfx##c1_class0_hp, 3, 2); ARR(pfx##c1_hp, 3, 2); ARR(pfx##c1_class0, 3, 2); ARR(pfx##c1_bits, 3, 2);
    NMV(mv_) NMV(dv_)
    ARR(intrabc, 3, 2); ARR(seg_pred, 3, 2); ARR(seg_spatial, 9, 8);
    ARR(filter_intra, 3, 2); ARR(filter_intra_mode, 6, 5);
    ARR(switchable_restore, 4, 3); ARR(wiener_restore, 3, 2); ARR(sgrproj_restore, 3, 2);
    ARR(y_mode, 14, 13); ARR(uv_mode, 15, k < 13 ? 13 : 14);
    ARR(partition, 11, k < 4 ? 4 : (k < 16 ? 10 : 8));
    ARR(switchable_interp, 4, 3); ARR(kf_y_mode, 14, 13); ARR(angle_delta, 8, 7);
    ARR(tx_size, 4, k < 3 ? 2 : 3);
    ARR(delta_q, 5, 4); ARR(delta_lf_multi, 5, 4); ARR(delta_lf, 5, 4);
    ARR(intra_ext_tx, 17, (k / 52) == 1 ? 7 : ((k / 52) == 2 ? 5 : 16));
    ARR(inter_ext_tx, 17, (k / 4) == 1 ? 16 : ((k / 4) == 2 ? 12 : ((k / 4) == 3 ? 2 : 16)));
    ARR(cfl_sign, 9, 8); ARR(cfl_alpha, 17, 16);
#undef ARR
#undef NMV
#undef CLR
}

static std::atomic<long long> g_prof_ns[8];   // several parser threads add to these
struct ProfPrinter { ~ProfPrinter() { if (getenv("AV1R_PROFILE")) fprintf(stderr, "[prof] tiles %.1f merge %.1f lf %.1f wrap %.1f begin %.1f ms\n", g_prof_ns[0] * 1e-6, g_prof_ns[1] * 1e-6, g_prof_ns[2] * 1e-6, g_prof_ns[3] * 1e-6, g_prof_ns[4] * 1e-6); } } g_prof_printer;
extern "C" void av1r_debug_parse_prof(double* out5, int reset) {
    for (int i = 0; i < 5; i++) out5[i] = g_prof_ns[i] * 1e-6;
    if (reset)
        for (auto& v : g_prof_ns) v = 0;
}
#define PROF_T() std::chrono::steady_clock::now()
#define PROF_ADD(i, a) g_prof_ns[i] += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - a).count()

// FrameWork objects are recycled process-wide: a 4K frame's maps and lists are ~20 MB of vectors whose allocation (page faults) and
// growth would otherwise be paid on every frame.
namespace {
struct FrameWorkPool {
    std::mutex m;
    std::vector<FrameWork*> free_;
    ~FrameWorkPool() { for (auto* f : free_) delete f; }
};
FrameWorkPool g_fw_pool;
}  // namespace

std::shared_ptr<FrameWork> acquire_framework() {
    FrameWork* f = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_fw_pool.m);
        if (!g_fw_pool.free_.empty()) {
            f = g_fw_pool.free_.back();
            g_fw_pool.free_.pop_back();
        }
    }
    if (!f) f = new FrameWork();
    return std::shared_ptr<FrameWork>(f, [](FrameWork* p) {
        std::lock_guard<std::mutex> lk(g_fw_pool.m);
        if (g_fw_pool.free_.size() < 96) g_fw_pool.free_.push_back(p);
        else delete p;
    });
}

StreamParser::StreamParser() {
    for (auto& c : slot_cdf_) cdf_load_defaults(c, 0);
}

int StreamParser::begin_frame(const FrameHdr& fh) {
    auto tb = std::chrono::steady_clock::now();
    struct Fin { std::chrono::steady_clock::time_point t; ~Fin() { g_prof_ns[4] += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t).count(); } } fin{tb};
    // AV1 level 6.3 caps a picture at 16384 x 8704 and 35,651,584 samples: anything beyond is a corrupt header, not a frame to allocate
    if (fh.upscaled_width > 16384 || fh.frame_height > 8704 || (int64_t)fh.upscaled_width * fh.frame_height > 35651584 || fh.frame_width < 1 ||
        fh.frame_height < 1)
        return fail(AV1R_EBITSTREAM, "frame size outside the AV1 level limits");
    cur_ = acquire_framework();
    cur_->init(hp.seq, fh);
    cur_->host_lf = host_lf_edges;
    cur_fh_ = fh;
    tiles_done_ = 0;
    have_frame_ = true;
    if (getenv("AV1R_DEP_TRACE")) {   // decode-order dependency trace (tools/dep_trace.py): which earlier frames this one needs
        static thread_local int slot_no[8], frame_no = 0;
        if (fh.frame_type == 0 && fh.show_frame) frame_no = 0;
        fprintf(stderr, "dep frame %d type %d show %d tiles %dx%d primary %d refs", frame_no, fh.frame_type, fh.show_frame, fh.tile_cols, fh.tile_rows,
                fh.primary_ref_frame == PRIMARY_REF_NONE ? -1 : slot_no[fh.ref_frame_idx[fh.primary_ref_frame]]);
        if (!fh.frame_is_intra)
            for (int i = 0; i < 7; i++) fprintf(stderr, " %d", slot_no[fh.ref_frame_idx[i]]);
        fprintf(stderr, " mfmv %d endcdf %d\n", fh.use_ref_frame_mvs, !fh.disable_frame_end_update_cdf);
        for (int i = 0; i < 8; i++)
            if ((fh.refresh_frame_flags >> i) & 1) slot_no[i] = frame_no;
        frame_no++;
    }
    if (fh.primary_ref_frame == PRIMARY_REF_NONE) {
        cdf_load_defaults(cur_init_cdf_, fh.base_q_idx);
    } else {
        cur_init_cdf_ = slot_cdf_[fh.ref_frame_idx[fh.primary_ref_frame]];
    }
    // restoration unit arrays (tiles fill disjoint ranges concurrently)
    if (!fh.allow_intrabc)
        for (int plane = 0; plane < hp.seq.num_planes; plane++) {
            if (fh.lr_type[plane] == RESTORE_NONE) continue;
            const int sx = plane ? hp.seq.subsampling_x : 0, sy = plane ? hp.seq.subsampling_y : 0;
            const int unit_size = fh.lr_size[plane];
            auto count_units = [](int us, int fs) { return std::max((fs + (us >> 1)) / us, 1); };
            cur_->lr_rows[plane] = count_units(unit_size, (fh.frame_height + sy) >> sy);
            cur_->lr_cols[plane] = count_units(unit_size, (fh.upscaled_width + sx) >> sx);
            LrUnit z;
            memset(&z, 0, sizeof(z));
            cur_->lr[plane].assign((size_t)cur_->lr_rows[plane] * cur_->lr_cols[plane], z);
        }
    cur_->warps.resize(8);
    memset(cur_->warps.data(), 0, sizeof(WarpRec) * 8);
    // chroma formats / bit depths the device path is not built for are refused up front (soft outcome for the caller)
    if (!hp.seq.mono_chrome && !(hp.seq.subsampling_x == 1 && hp.seq.subsampling_y == 1))
        return fail(AV1R_ENOSYS, "4:4:4 / 4:2:2 streams are not supported yet");
    if (hp.seq.bit_depth > 10) return fail(AV1R_ENOSYS, "12-bit streams are not supported yet");
    // super-resolution: intra frames are reconstructed at the coded (downscaled) width and upscaled before loop restoration (K6);
    // an inter frame coded with superres predicts from references of a different width: scaled motion compensation (K2)
    if (!fh.frame_is_intra) {
        for (int i = 0; i < REFS_PER_FRAME; i++) {
            const RefHdrState& r = hp.refs[fh.ref_frame_idx[i]];
            if (!r.valid) return fail(AV1R_EBITSTREAM, "inter frame references an empty slot");
            // reference scaling (spec 7.11.3.3 / 5.9.7): a reference may be up to twice as large and down to 1/16 of the frame
            if (2 * fh.frame_width < r.upscaled_width || 2 * fh.frame_height < r.frame_height || fh.frame_width > 16 * r.upscaled_width ||
                fh.frame_height > 16 * r.frame_height)
                return fail(AV1R_EBITSTREAM, "reference frame size outside the allowed scaling range");
            const int xscale = (int)((((int64_t)r.upscaled_width << 14) + fh.frame_width / 2) / fh.frame_width);
            const int yscale = (int)((((int64_t)r.frame_height << 14) + fh.frame_height / 2) / fh.frame_height);
            cur_->ref_scaled[LAST_FRAME + i] = xscale != (1 << 14) || yscale != (1 << 14);
        }
        // global warp models of the 7 references (spec 7.11.3.6 validity)
        cur_->warps.resize(8);
        for (int ref = LAST_FRAME; ref <= ALTREF_FRAME; ref++) {
            WarpRec& w = cur_->warps[ref];
            memcpy(w.mat, fh.gm_params[ref], sizeof(w.mat));
            int16_t sh[4] = {0, 0, 0, 0};
            cur_->gm_warp_valid[ref] = (uint8_t)(fh.gm_type[ref] > GM_TRANSLATION ? setup_shear(w.mat, sh) : 0);
            w.alpha = sh[0]; w.beta = sh[1]; w.gamma = sh[2]; w.delta = sh[3];
        }
        memset(&cur_->warps[0], 0, sizeof(WarpRec));
        if (fh.use_ref_frame_mvs) motion_field_estimation();
    }
    // load_previous_segment_ids (spec 7.20)
    if (fh.seg.enabled && fh.primary_ref_frame != PRIMARY_REF_NONE) {
        const auto& prev = slot_fw_[fh.ref_frame_idx[fh.primary_ref_frame]];
        if (prev && prev->mi_cols == fh.mi_cols && prev->mi_rows == fh.mi_rows) cur_->prev_seg_ids = prev->seg_ids;
    }
    return 0;
}

// spec 7.9: project the motion vectors saved with earlier frames onto the current frame (8x8 granularity)
void StreamParser::motion_field_estimation() {
    const FrameHdr& fh = cur_fh_;
    FrameWork& fw = *cur_;
    const int w8 = fh.mi_cols >> 1, h8 = fh.mi_rows >> 1;
    fw.mfmv.assign((size_t)w8 * h8, MfMv{{0, 0}, 0});
    static const int kDivMult[32] = {0,    16384, 8192, 5461, 4096, 3276, 2730, 2340, 2048, 1820, 1638, 1489, 1365, 1260, 1170, 1092,
                                     1024, 963,   910,  862,  819,  780,  744,  712,  682,  655,  630,  606,  585,  564,  546,  528};
    // One source = one earlier frame whose saved motion vectors are projected.  The sources are chosen first (header data only),
    // then all of them are walked band by band in ONE parallel loop: a projected position never leaves the 8-row band of its
    // source (MAX_OFFSET_HEIGHT = 0), so bands are independent, and inside a band the sources are applied in the order of the
    // specification and each in raster order, so "the later projection wins" stays deterministic.
    struct Source {
        const FrameWork* sfw;
        int dst_sign;
        int roff[8];
        int64_t rfac[8];
        bool rok[8];
    };
    Source srcs[4];
    int n_srcs = 0;
    auto project = [&](int src, int dst_sign) -> int {
        const int src_idx = fh.ref_frame_idx[src - LAST_FRAME];
        const RefHdrState& r = hp.refs[src_idx];
        const auto& sfw = slot_fw_[src_idx];
        if (!sfw || r.mi_rows != fh.mi_rows || r.mi_cols != fh.mi_cols || r.frame_type == INTRA_ONLY_FRAME || r.frame_type == KEY_FRAME ||
            sfw->saved_mvs.empty())
            return 0;
        const int ref_to_cur = hp.get_relative_dist(fh.order_hints[src], fh.order_hint);
        // per reference of the source frame: offset, validity and the projection factor num * Div_Mult[den] (spec 7.9.3)
        Source& S = srcs[n_srcs++];
        S.sfw = sfw.get();
        S.dst_sign = dst_sign;
        for (int rf = 0; rf < 8; rf++) {
            S.roff[rf] = rf > INTRA_FRAME ? hp.get_relative_dist(fh.order_hints[src], r.saved_order_hints[rf]) : 0;
            S.rok[rf] = rf > INTRA_FRAME && std::abs(ref_to_cur) <= 31 && std::abs(S.roff[rf]) <= 31 && S.roff[rf] > 0;
            S.rfac[rf] = S.rok[rf] ? (int64_t)std::max(-31, std::min(31, ref_to_cur * dst_sign)) * kDivMult[std::min(31, S.roff[rf])] : 0;
        }
        return 1;
    };
    const int last_idx = fh.ref_frame_idx[0];
    const int last_alt = hp.refs[last_idx].saved_order_hints[ALTREF_FRAME];
    if (last_alt != fh.order_hints[GOLDEN_FRAME]) project(LAST_FRAME, -1);
    int ref_stamp = 1;
    if (hp.get_relative_dist(fh.order_hints[BWDREF_FRAME], fh.order_hint) > 0 && project(BWDREF_FRAME, 1)) ref_stamp--;
    if (hp.get_relative_dist(fh.order_hints[ALTREF2_FRAME], fh.order_hint) > 0 && project(ALTREF2_FRAME, 1)) ref_stamp--;
    if (hp.get_relative_dist(fh.order_hints[ALTREF_FRAME], fh.order_hint) > 0 && ref_stamp >= 0 && project(ALTREF_FRAME, 1)) ref_stamp--;
    if (ref_stamp >= 0) project(LAST2_FRAME, -1);
    if (!n_srcs) return;
    WorkerPool::get().parallel_for((h8 + 7) >> 3, [&](int band) {
        for (int si = 0; si < n_srcs; si++) {
            const Source& S = srcs[si];
            const int dst_sign = S.dst_sign;
            for (int row8 = band * 8; row8 < std::min(h8, band * 8 + 8); row8++)
                for (int col8 = 0; col8 < w8; col8++) {
                    const SavedMv& sm = S.sfw->saved_mvs[(size_t)row8 * w8 + col8];
                    if (sm.ref <= INTRA_FRAME || !S.rok[sm.ref]) continue;
                    const int ref_offset = S.roff[sm.ref];
                    const int64_t fac = S.rfac[sm.ref];
                    int proj[2];
                    const int mvc[2] = {sm.mv.row, sm.mv.col};
                    for (int i = 0; i < 2; i++) {
                        const int64_t v = (int64_t)mvc[i] * fac;
                        const int64_t sc = v >= 0 ? (v + 8192) >> 14 : -((-v + 8192) >> 14);
                        proj[i] = (int)std::max<int64_t>(-(1 << 14) + 1, std::min<int64_t>((1 << 14) - 1, sc));
                    }
                    auto pos = [&](int v8, int delta, int max8, int max_off8, bool& ok) {
                        const int base8 = (v8 >> 3) << 3;
                        const int off8 = delta >= 0 ? (delta >> 6) : -((-delta) >> 6);
                        v8 += dst_sign * off8;
                        if (v8 < 0 || v8 >= max8 || v8 < base8 - max_off8 || v8 >= base8 + 8 + max_off8) ok = false;
                        return v8;
                    };
                    bool ok = true;
                    const int py = pos(row8, proj[0], h8, 0, ok);
                    const int px = pos(col8, proj[1], w8, 8, ok);
                    if (!ok) continue;
                    MfMv& m = fw.mfmv[(size_t)py * w8 + px];
                    m.mv = sm.mv;
                    m.ref_offset = (int8_t)ref_offset;
                }
        }
    });
}

int StreamParser::tile_group(const uint8_t* payload, size_t size, size_t offset) {
    BitReader br(payload + offset, size - offset);
    TileGroupInfo tg;
    if (!hp.parse_tile_group_header(br, cur_fh_, tg)) return fail(AV1R_EBITSTREAM, hp.error);
    // tile groups cover the tiles of a frame in order, each tile exactly once (spec 5.11.1: tg_start == TileNum of the next tile)
    if (tg.tg_start != tiles_done_ || tg.tg_end < tg.tg_start || tg.tg_end >= cur_fh_.tile_cols * cur_fh_.tile_rows)
        return fail(AV1R_EBITSTREAM, "tile group does not continue at the next tile of the frame");
    size_t pos = offset + tg.data_offset;
    auto t0 = std::chrono::steady_clock::now();
    // ---- locate the tiles of this group, then parse them concurrently (each tile is an independent symbol stream)
    struct Task { const uint8_t* data; size_t size; int row, col, tile; };
    std::vector<Task> tasks;
    for (int tile = tg.tg_start; tile <= tg.tg_end; tile++) {
        size_t tile_size;
        if (tile == tg.tg_end) {
            tile_size = size - pos;
        } else {
            if (pos + cur_fh_.tile_size_bytes > size) return fail(AV1R_EBITSTREAM, "truncated tile size");
            tile_size = 0;
            for (int i = 0; i < cur_fh_.tile_size_bytes; i++) tile_size |= (size_t)payload[pos + i] << (8 * i);
            tile_size += 1;
            pos += cur_fh_.tile_size_bytes;
        }
        if (pos + tile_size > size) return fail(AV1R_EBITSTREAM, "tile exceeds tile group");
        tasks.push_back(Task{payload + pos, tile_size, tile / cur_fh_.tile_cols, tile % cur_fh_.tile_cols, tile});
        pos += tile_size;
    }
    FrameWork& fw = *cur_;
    const size_t first_out = fw.n_tiles_used;
    for (size_t i = 0; i < tasks.size(); i++) fw.next_tile();
    auto parse_one = [&](int i) {
        const Task& t = tasks[i];
        TileOut& to = *fw.tiles[first_out + i];
        TileDecoder td(hp.seq, hp, fw, to, cur_init_cdf_);
        to.rc = td.decode_tile(t.data, t.size, t.row, t.col);
        if (to.rc) to.err = td.err;
        if (t.tile == cur_fh_.context_update_tile_id) to.end_cdf = td.cdf;
    };
    if (tile_threads && tasks.size() > 1) WorkerPool::get().parallel_for((int)tasks.size(), parse_one);
    else for (size_t i = 0; i < tasks.size(); i++) parse_one((int)i);
    PROF_ADD(0, t0);
    // ---- what the *next* frame's parse needs is ready now (CDFs of the context-update tile); the merge of the per-tile work-lists
    // into the frame's lists is deferred to finalize_framework(), off the frame-to-frame critical path
    for (size_t i = 0; i < tasks.size(); i++) {
        TileOut& to = *fw.tiles[first_out + i];
        if (to.rc) return fail(to.rc, to.err);
        if (tasks[i].tile == cur_fh_.context_update_tile_id) {
            fw.end_cdf = to.end_cdf;
            fw.have_end_cdf = true;
        }
        tiles_done_++;
    }
    fw.finalized = false;
    cur_->parse_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

// Second half of a frame's host work: concatenates the per-tile work-lists in tile order (rebasing the offsets the records carry)
// and classifies the deblocking edges.  Nothing later frames' *parse* reads is touched here (that is mode info, saved motion
// vectors, segment ids and CDFs, all complete when parse_tu returns), so the verify path runs it on the staging thread while the
// next frame of the GOP is already being parsed.
void finalize_framework(FrameWork& fw) {
    if (fw.finalized) return;
    auto tm = PROF_T();
    for (size_t i = 0; i < fw.n_tiles_used; i++) {
        TileOut& to = *fw.tiles[i];
        const uint32_t coef_base = (uint32_t)fw.coefs.size(), pal_base = (uint32_t)fw.pal.size();
        const uint32_t tx_base = (uint32_t)fw.tx.size(), obmc_base = (uint32_t)fw.obmc.size();
        const int warp_base = (int)fw.warps.size() - 8;   // slots 0..7 hold the global models
        const size_t tx0 = fw.tx.size();
        fw.tx.insert(fw.tx.end(), to.tx.begin(), to.tx.end());
        for (size_t k = tx0; k < fw.tx.size(); k++) {
            fw.tx[k].coef_off += coef_base;
            if (fw.tx[k].mode == TXM_PALETTE) fw.tx[k].pal_off += pal_base;
        }
        fw.coefs.insert(fw.coefs.end(), to.coefs.begin(), to.coefs.end());
        fw.pal.insert(fw.pal.end(), to.pal.begin(), to.pal.end());
        const size_t sb0 = fw.sbs.size();
        fw.sbs.insert(fw.sbs.end(), to.sbs.begin(), to.sbs.end());
        for (size_t k = sb0; k < fw.sbs.size(); k++) fw.sbs[k].first += tx_base;
        const size_t in0 = fw.inter.size();
        fw.inter.insert(fw.inter.end(), to.inter.begin(), to.inter.end());
        for (size_t k = in0; k < fw.inter.size(); k++) {
            InterBlk& r = fw.inter[k];
            r.obmc_first += obmc_base;
            for (int l = 0; l < 2; l++)
                if (r.warp[l] >= 8) r.warp[l] = (int16_t)(r.warp[l] + warp_base);
        }
        fw.obmc.insert(fw.obmc.end(), to.obmc.begin(), to.obmc.end());
        fw.warps.insert(fw.warps.end(), to.warps.begin(), to.warps.end());
        fw.lf_blocks.insert(fw.lf_blocks.end(), to.lf_blocks.begin(), to.lf_blocks.end());
        fw.coded_samples += to.coded_samples;
        fw.coef_tokens += to.coef_tokens;
        fw.tx_blocks += to.tx_blocks;
        fw.inter_samples += to.inter_samples;
        fw.inter_ref_samples += to.inter_ref_samples;
        for (int k = 0; k < 24; k++) fw.tool_hist[k] += to.tool_hist[k];
    }
    PROF_ADD(1, tm);
    auto tl = PROF_T();
    if (fw.host_lf) build_loopfilter_edges(fw);
    PROF_ADD(2, tl);
    fw.parse_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tm).count();
    fw.finalized = true;
}

int StreamParser::finish_frame(int64_t pts, std::vector<ParsedFrame>& out) {
    auto t0 = std::chrono::steady_clock::now();
    auto tw = PROF_T();
    {
        FrameWork& fw = *cur_;
        // segment map kept for later frames (spec 7.4 decode frame wrapup)
        if (cur_fh_.seg.enabled && !cur_fh_.seg.update_map) {
            if (!fw.prev_seg_ids.empty()) fw.seg_ids = fw.prev_seg_ids;
            else std::fill(fw.seg_ids.begin(), fw.seg_ids.end(), 0);
        }
        // motion field motion vector storage (spec 7.19)
        for (int i = 0; i < 8; i++) fw.saved_order_hints[i] = cur_fh_.order_hints[i];
        if (!cur_fh_.frame_is_intra) {
            const int w8 = fw.mi_cols >> 1, h8 = fw.mi_rows >> 1;
            fw.saved_mvs.assign((size_t)w8 * h8, SavedMv{{0, 0}, 0});
            WorkerPool::get().parallel_for((h8 + 15) >> 4, [&](int job) {
            for (int row8 = job * 16; row8 < std::min(h8, job * 16 + 16); row8++)
                for (int col8 = 0; col8 < w8; col8++) {
                    const BlockInfo* b = fw.mi[(size_t)(row8 * 2 + 1) * fw.mi_cols + col8 * 2 + 1];
                    if (!b) continue;
                    SavedMv& sm = fw.saved_mvs[(size_t)row8 * w8 + col8];
                    for (int list = 0; list < 2; list++) {
                        const int r = b->ref_frame[list];
                        if (r <= INTRA_FRAME) continue;
                        if (hp.get_relative_dist(cur_fh_.order_hints[r], cur_fh_.order_hint) >= 0) continue;
                        if (std::abs(b->mv[list].row) > 4095 || std::abs(b->mv[list].col) > 4095) continue;
                        sm.mv = b->mv[list];
                        sm.ref = (int8_t)r;
                    }
                }
            });
        }
    }
    PROF_ADD(3, tw);
    cur_->parse_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    CdfCtx save = cur_init_cdf_;
    if (!cur_fh_.disable_frame_end_update_cdf && cur_->have_end_cdf) {
        save = cur_->end_cdf;
        cdf_clear_counters(save);
    }
    for (int i = 0; i < NUM_REF_FRAMES; i++)
        if ((cur_fh_.refresh_frame_flags >> i) & 1) {
            slot_cdf_[i] = save;
            slot_fw_[i] = cur_;
        }
    hp.reference_update(cur_fh_);
    ParsedFrame pf;
    pf.fw = cur_;
    pf.fh = cur_fh_;
    pf.seq = hp.seq;
    pf.pts = pts;
    if (!defer_finalize) finalize_framework(*cur_);
    out.push_back(pf);
    cur_.reset();
    have_frame_ = false;
    return 0;
}

int StreamParser::parse_tu(const uint8_t* data, size_t len, int64_t pts, std::vector<ParsedFrame>& out) {
    const int rc = parse_tu_inner(data, len, pts, out);
    if (rc) {   // a failed temporal unit never leaves a half-assembled frame behind: the next unit starts clean
        have_frame_ = false;
        cur_.reset();
    }
    return rc;
}

int StreamParser::parse_tu_inner(const uint8_t* data, size_t len, int64_t pts, std::vector<ParsedFrame>& out) {
    std::vector<ObuUnit> obus;
    if (!hp.split_obus(data, len, obus)) return fail(AV1R_EBITSTREAM, hp.error);
    for (const ObuUnit& u : obus) {
        switch (u.type) {
            case OBU_SEQUENCE_HEADER:
                if (!hp.parse_sequence_header(u.data, u.size)) return fail(AV1R_EBITSTREAM, hp.error);
                break;
            case OBU_TEMPORAL_DELIMITER:
                if (have_frame_) return fail(AV1R_EBITSTREAM, "temporal delimiter inside a frame: tile data of the previous frame is missing");
                break;
            case OBU_FRAME_HEADER:
            case OBU_REDUNDANT_FRAME_HEADER:
            case OBU_FRAME: {
                if (have_frame_) {
                    // only a copy of the active header may appear while tile groups are outstanding (spec 5.9.1 frame_header_copy)
                    if (u.type == OBU_FRAME) return fail(AV1R_EBITSTREAM, "frame OBU while another frame is open");
                    const size_t nb = cur_hdr_bytes_.size() > 1 ? cur_hdr_bytes_.size() - 1 : 0;   // last byte may differ in trailing bits
                    if (u.size < nb || memcmp(u.data, cur_hdr_bytes_.data(), nb) != 0)
                        return fail(AV1R_EBITSTREAM, "new frame header while tile data of the previous frame is missing");
                    break;   // redundant copy of the active header
                }
                if (u.type == OBU_REDUNDANT_FRAME_HEADER) return fail(AV1R_EBITSTREAM, "redundant frame header without an active frame");
                BitReader br(u.data, u.size);
                FrameHdr fh;
                if (!hp.parse_frame_header(br, fh, u.temporal_id, u.spatial_id)) return fail(AV1R_EBITSTREAM, hp.error);
                if (fh.show_existing_frame) {
                    ParsedFrame pf;
                    pf.show_existing_slot = fh.frame_to_show_map_idx;
                    pf.fh = fh;
                    pf.seq = hp.seq;
                    pf.pts = pts;
                    if (fh.frame_type == KEY_FRAME) {
                        // frame loading process (7.21): the shown key frame refreshes every slot
                        const int s = fh.frame_to_show_map_idx;
                        RefHdrState r = hp.refs[s];
                        CdfCtx c = slot_cdf_[s];
                        auto f = slot_fw_[s];
                        for (int i = 0; i < NUM_REF_FRAMES; i++) {
                            hp.refs[i] = r;
                            slot_cdf_[i] = c;
                            slot_fw_[i] = f;
                        }
                    }
                    out.push_back(pf);
                    break;
                }
                int rc = begin_frame(fh);
                if (rc) return rc;
                {
                    BitReader tmp = br;
                    tmp.byte_align();
                    cur_hdr_bytes_.assign(u.data, u.data + std::min(u.size, (size_t)tmp.byte_pos()));
                }
                if (u.type == OBU_FRAME) {
                    br.byte_align();
                    rc = tile_group(u.data, u.size, br.byte_pos());
                    if (rc) return rc;
                    if (tiles_done_ == cur_fh_.tile_cols * cur_fh_.tile_rows) {
                        rc = finish_frame(pts, out);
                        if (rc) return rc;
                    }
                }
                break;
            }
            case OBU_TILE_GROUP: {
                if (!have_frame_) return fail(AV1R_EBITSTREAM, "tile group without frame header");
                int rc = tile_group(u.data, u.size, 0);
                if (rc) return rc;
                if (tiles_done_ == cur_fh_.tile_cols * cur_fh_.tile_rows) {
                    rc = finish_frame(pts, out);
                    if (rc) return rc;
                }
                break;
            }
            case OBU_TILE_LIST:
                return fail(AV1R_ENOSYS, "large-scale tile lists are not supported");
            default:
                break;   // metadata, padding
        }
    }
    // a temporal unit holds whole frames: a frame header whose tile groups did not all arrive is a truncated / corrupt stream
    // (without this check the frame would silently vanish and the verifier would report one frame fewer as "ok")
    if (have_frame_) return fail(AV1R_EBITSTREAM, "temporal unit ends inside a frame: tile data missing");
    return 0;
}

void build_loopfilter_edges(FrameWork& fw) {
    const FrameHdr& fh = fw.fh;
    if (!fh.lf.level[0] && !fh.lf.level[1]) return;
    for (int plane = 0; plane < (fw.mono ? 1 : 3); plane++) {
        if (plane > 0 && !fh.lf.level[1 + plane]) continue;
        const int sx = plane ? fw.subx : 0, sy = plane ? fw.suby : 0;
        const int pw4 = fw.plane_w4(plane), ph4 = fw.plane_h4(plane);
        const int max_len = plane ? 8 : 16;
        const int li[2] = {plane == 0 ? 0 : plane + 1, plane == 0 ? 1 : plane + 1};   // index into BlockInfo::lf_lvl per pass
        const uint8_t* lf_tx = fw.lf_tx[plane].data();
        LfEdge* edges = fw.lf[plane].data();
        const int rows_per_job = 16;
        const int n_jobs = (ph4 + rows_per_job - 1) / rows_per_job;
        WorkerPool::get().parallel_for(n_jobs, [&](int job) {
            for (int r4 = job * rows_per_job; r4 < std::min(ph4, (job + 1) * rows_per_job); r4++) {
                const int row = r4 << sy;
                if (row >= fw.mi_rows) continue;
                // for sub-sampled planes the spec addresses mode info at the odd (bottom-right) luma mi
                const int mrow = std::min(fw.mi_rows - 1, row | sy);
                const int prow_v = mrow, prow_h = std::min(fw.mi_rows - 1, ((r4 - 1) << sy) | sy);
                BlockInfo* const* mi_ptr_row = &fw.mi[(size_t)mrow * fw.mi_cols];   // coverage test only (null = no block here)
                const LfMi* mi_row = &fw.lf_mi[(size_t)mrow * fw.mi_cols];
                const LfMi* mi_prev_row = r4 > 0 ? &fw.lf_mi[(size_t)prow_h * fw.mi_cols] : nullptr;
                BlockInfo* const* mi_prev_ptr_row = r4 > 0 ? &fw.mi[(size_t)prow_h * fw.mi_cols] : nullptr;
                (void)prow_v;
                const bool row_visible = row * 4 < fh.frame_height;
                for (int c4 = 0; c4 < pw4; c4++) {
                    const int col = c4 << sx;
                    if (col >= fw.mi_cols) continue;
                    const size_t idx = (size_t)r4 * pw4 + c4;
                    LfEdge e{0, 0, 0, 0};
                    const int mcol = std::min(fw.mi_cols - 1, col | sx);
                    if (row_visible && col * 4 < fh.frame_width && mi_ptr_row[mcol]) {
                        const LfMi& b = mi_row[mcol];
                        const int txsz = lf_tx[idx];
                        const int txw = kTxW[txsz], txh = kTxH[txsz];
                        const int bwp = std::max(4, kBlockW[b.bsize] >> sx), bhp = std::max(4, kBlockH[b.bsize] >> sy);
                        const int xp = c4 * 4, yp = r4 * 4;
                        const bool filt_inside = b.filt_inside;
                        if (c4 > 0 && (xp & (txw - 1)) == 0 && (filt_inside || (xp & (bwp - 1)) == 0)) {
                            int lvl = b.lvl[li[0]];
                            if (!lvl) {
                                const int pcol = std::min(fw.mi_cols - 1, ((c4 - 1) << sx) | sx);
                                if (mi_ptr_row[pcol]) lvl = mi_row[pcol].lvl[li[0]];
                            }
                            if (lvl) {
                                e.len_v = (uint8_t)std::min(max_len, std::min((int)kTxW[lf_tx[idx - 1]], txw));
                                e.lvl_v = (uint8_t)lvl;
                            }
                        }
                        if (r4 > 0 && (yp & (txh - 1)) == 0 && (filt_inside || (yp & (bhp - 1)) == 0)) {
                            int lvl = b.lvl[li[1]];
                            if (!lvl) {
                                if (mi_prev_ptr_row[mcol]) lvl = mi_prev_row[mcol].lvl[li[1]];
                            }
                            if (lvl) {
                                e.len_h = (uint8_t)std::min(max_len, std::min((int)kTxH[lf_tx[idx - pw4]], txh));
                                e.lvl_h = (uint8_t)lvl;
                            }
                        }
                    }
                    edges[idx] = e;
                }
            }
        });
    }
}

}  // namespace av1r
