// TEST INFRASTRUCTURE ONLY -- whole-stream CPU decoder = product host parser + scalar oracle
// reconstruction.  This is the checker for the CUDA path (and what validates the parser against
// libdav1d on this GPU-less build box).  Never linked into libav1r.so.
#include <stdint.h>
#include <string.h>

#include <memory>
#include <string>
#include <vector>

#include "../av1-go_b200/csrc/stream_parser.h"
#include "../include/av1r_stages.h"
#include "oracle_frame.h"

extern "C" void orc_film_grain(const void* g, int bd, int w, int h, int subx, int suby, int mono, int mc_identity,
                               const void* const in[3], const int in_stride[3], void* const out[3], const int out_stride[3]);

namespace orc {
using namespace av1r;

struct OutFrame {
    std::shared_ptr<Frame> f;
    int w, h, bd;
    double parse_ms;
    uint64_t coded_samples, coef_tokens, tx_blocks;
};

struct Stream {
    StreamParser sp;
    int inloop_filters = 7, apply_grain = 1;
    std::shared_ptr<Frame> slots[8];
    FilmGrainParams slot_fg[8];
    std::vector<OutFrame> out;
    std::string err;
    uint64_t tool_hist[24] = {0};
};

static std::shared_ptr<Frame> make_frame(const SeqHdr& seq, const FrameHdr& fh, bool upscaled = false) {
    auto f = std::make_shared<Frame>();
    FrameGeom& g = f->g;
    g.bd = seq.bit_depth;
    g.subx = seq.subsampling_x;
    g.suby = seq.subsampling_y;
    g.mono = seq.mono_chrome;
    for (int p = 0; p < 3; p++) {
        int sx = p ? g.subx : 0, sy = p ? g.suby : 0;
        g.w[p] = ((upscaled ? fh.upscaled_width : fh.frame_width) + sx) >> sx;
        g.h[p] = (fh.frame_height + sy) >> sy;
        g.cw[p] = ((upscaled ? 2 * ((fh.upscaled_width + 7) >> 3) : fh.mi_cols) * 4) >> sx;
        g.ch[p] = (fh.mi_rows * 4) >> sy;
        f->p[p].alloc(g.cw[p], g.ch[p]);
    }
    g.dq_dc[0] = fh.delta_q_y_dc; g.dq_ac[0] = 0;
    g.dq_dc[1] = fh.delta_q_u_dc; g.dq_ac[1] = fh.delta_q_u_ac;
    g.dq_dc[2] = fh.delta_q_v_dc; g.dq_ac[2] = fh.delta_q_v_ac;
    g.enable_edge_filter = seq.enable_intra_edge_filter;
    return f;
}

static std::shared_ptr<Frame> grain(const Stream& s, const std::shared_ptr<Frame>& in, const FilmGrainParams& fg, const FrameHdr& fh) {
    (void)fh;
    if (!s.apply_grain || !fg.apply_grain) return in;
    auto o = std::make_shared<Frame>(*in);
    const void* ip[3];
    void* op[3];
    int is[3], os[3];
    for (int p = 0; p < 3; p++) {
        ip[p] = in->p[p].d.data();
        op[p] = o->p[p].d.data();
        is[p] = in->p[p].stride * 2;
        os[p] = o->p[p].stride * 2;
    }
    // oracle frames are always 16-bit containers: run the grain at the stream's bit depth on uint16 data
    // (orc_film_grain reads uint8 when bd == 8, so widen through a temporary for 8-bit)
    const FrameGeom& g = in->g;
    if (g.bd == 8) {
        std::vector<uint8_t> i8[3], o8[3];
        for (int p = 0; p < 3; p++) {
            i8[p].resize((size_t)g.cw[p] * g.ch[p]);
            o8[p].resize(i8[p].size());
            for (int y = 0; y < g.ch[p]; y++)
                for (int x = 0; x < g.cw[p]; x++) i8[p][(size_t)y * g.cw[p] + x] = (uint8_t)in->p[p].at(x, y);
            ip[p] = i8[p].data();
            op[p] = o8[p].data();
            is[p] = os[p] = g.cw[p];
        }
        orc_film_grain(&fg, 8, g.w[0], g.h[0], g.subx, g.suby, g.mono, s.sp.hp.seq.matrix_coefficients == 0, ip, is, op, os);
        for (int p = 0; p < 3; p++)
            for (int y = 0; y < g.h[p]; y++)
                for (int x = 0; x < g.w[p]; x++) o->p[p].at(x, y) = o8[p][(size_t)y * g.cw[p] + x];
    } else {
        orc_film_grain(&fg, g.bd, g.w[0], g.h[0], g.subx, g.suby, g.mono, s.sp.hp.seq.matrix_coefficients == 0, ip, is, op, os);
    }
    return o;
}

}  // namespace orc

using namespace orc;

extern "C" void* orc_stream_open(int inloop_filters, int apply_grain) {
    Stream* s = new Stream();
    s->inloop_filters = inloop_filters;
    s->apply_grain = apply_grain;
    return s;
}
extern "C" void orc_stream_close(void* h) { delete (Stream*)h; }
extern "C" const char* orc_stream_error(void* h) { return ((Stream*)h)->err.c_str(); }

extern "C" int orc_stream_decode(void* h, const uint8_t* tu, size_t len) {
    Stream* s = (Stream*)h;
    std::vector<ParsedFrame> pfs;
    int rc = s->sp.parse_tu(tu, len, 0, pfs);
    int produced = 0;
    for (ParsedFrame& pf : pfs) {
        if (pf.show_existing_slot >= 0) {
            auto f = s->slots[pf.show_existing_slot];
            if (!f) { s->err = "show_existing_frame of empty slot"; return -74; }
            OutFrame o{grain(*s, f, pf.fh.fg, pf.fh), f->g.w[0], f->g.h[0], f->g.bd, 0, 0, 0, 0};
            s->out.push_back(o);
            produced++;
            if (pf.fh.frame_type == KEY_FRAME)
                for (int i = 0; i < 8; i++) s->slots[i] = f;
            continue;
        }
        const FrameWork& fw = *pf.fw;
        for (int i = 0; i < 24; i++) s->tool_hist[i] += fw.tool_hist[i];
        auto rec = make_frame(s->sp.hp.seq, pf.fh);
        const Frame* refs[8];
        for (int i = 0; i < 8; i++) refs[i] = s->slots[i].get();
        reconstruct_frame(fw, *rec, refs);
        std::shared_ptr<Frame> cur = rec;
        const bool do_db = (s->inloop_filters & 1) && (pf.fh.lf.level[0] || pf.fh.lf.level[1]);
        if (do_db) deblock_frame(fw, *cur);
        std::shared_ptr<Frame> deblocked = cur;
        if ((s->inloop_filters & 2) && pf.fh.enable_cdef_frame) {
            auto c = std::make_shared<Frame>(*cur);
            cdef_frame(fw, *cur, *c);
            cur = c;
        }
        if (pf.fh.use_superres) {   // spec 7.16: both the CDEF output and the deblocked frame (source of the LR stripe boundaries)
            auto up = make_frame(s->sp.hp.seq, pf.fh, true);
            upscale_frame(pf.fh, *cur, *up);
            if (deblocked != cur) {
                auto upd = make_frame(s->sp.hp.seq, pf.fh, true);
                upscale_frame(pf.fh, *deblocked, *upd);
                deblocked = upd;
            } else {
                deblocked = up;
            }
            cur = up;
        }
        if ((s->inloop_filters & 4) && pf.fh.uses_lr) {
            auto l = std::make_shared<Frame>(*cur);
            lr_frame(fw, *deblocked, *cur, *l);
            cur = l;
        }
        for (int i = 0; i < 8; i++)
            if ((pf.fh.refresh_frame_flags >> i) & 1) s->slots[i] = cur;
        if (pf.fh.show_frame) {
            OutFrame o{grain(*s, cur, pf.fh.fg, pf.fh), cur->g.w[0], cur->g.h[0], cur->g.bd, fw.parse_ms, fw.coded_samples, fw.coef_tokens, fw.tx_blocks};
            s->out.push_back(o);
            produced++;
        }
    }
    if (rc) { s->err = s->sp.err; return rc; }
    return produced;
}

extern "C" int orc_stream_num_frames(void* h) { return (int)((Stream*)h)->out.size(); }
extern "C" int orc_stream_frame_info(void* h, int idx, int* w, int* hh, int* bd, double* parse_ms, uint64_t* stats3) {
    Stream* s = (Stream*)h;
    if (idx < 0 || idx >= (int)s->out.size()) return -22;
    const OutFrame& o = s->out[idx];
    *w = o.w; *hh = o.h; *bd = o.bd;
    if (parse_ms) *parse_ms = o.parse_ms;
    if (stats3) { stats3[0] = o.coded_samples; stats3[1] = o.coef_tokens; stats3[2] = o.tx_blocks; }
    return 0;
}
extern "C" int orc_stream_frame_copy(void* h, int idx, int plane, uint16_t* dst, int dst_stride) {
    Stream* s = (Stream*)h;
    if (idx < 0 || idx >= (int)s->out.size()) return -22;
    const Frame& f = *s->out[idx].f;
    for (int y = 0; y < f.g.h[plane]; y++)
        memcpy(dst + (size_t)y * dst_stride, &f.p[plane].d[(size_t)y * f.p[plane].stride], sizeof(uint16_t) * f.g.w[plane]);
    return 0;
}

extern "C" void orc_stream_tool_hist(void* h, uint64_t* out24) { memcpy(out24, ((Stream*)h)->tool_hist, sizeof(uint64_t) * 24); }
