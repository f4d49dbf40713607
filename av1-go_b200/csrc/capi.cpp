// C ABI glue that needs no GPU (version / defaults).
#include <cstdio>
#include <cstring>

#include "../../include/av1r.h"
#include "../../include/av1r_stages.h"

extern "C" uint32_t av1r_abi_version(void) { return AV1R_ABI_VERSION; }

extern "C" void av1r_default_config(av1r_config* cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = sizeof(*cfg);
    cfg->device = 0;
    cfg->streams = 2;
    cfg->frames_in_flight = 8;
    cfg->parity_md5 = 0;
    cfg->apply_grain = 1;
    cfg->inloop_filters = 7;
    cfg->keep_frames = 0;
    cfg->host_threads = 0;
}

// ---- host-only: demux + sequential symbol parse of a whole container, no device work (measurement / diagnostics).
// This is the stage north_star asks to be timed separately; it mirrors the host half of av1r_verify_buffer
// (GOP segments on host_threads workers, tiles of a frame on the process-wide pool).
#include <atomic>
#include <chrono>
#include <thread>

#include "demux.h"
#include "k3_plan.h"
#include "parallel.h"
#include "stream_parser.h"

extern "C" int av1r_parse_buffer(const uint8_t* data, size_t len, int host_threads, int tile_threads, av1r_report* out) {
    using namespace av1r;
    if (!data || !out) return AV1R_EINVAL;
    memset(out, 0, sizeof(*out));
    out->struct_size = sizeof(*out);
    out->first_bad_frame = -1;
    DemuxResult dm;
    std::string derr;
    if (!demux_buffer(data, len, dm, derr)) {
        snprintf(out->message, sizeof(out->message), "%s", derr.c_str());
        return out->status = AV1R_EBITSTREAM;
    }
    auto t0 = std::chrono::steady_clock::now();
    // segment starts: temporal units whose first frame is a shown key frame
    HeaderParser scan;
    std::vector<size_t> starts;
    for (size_t i = 0; i < dm.tus.size(); i++) {
        std::vector<ObuUnit> obus;
        if (!scan.split_obus(data + dm.tus[i].offset, dm.tus[i].size, obus)) break;
        bool first = true;
        for (const ObuUnit& u : obus) {
            if (u.type == OBU_SEQUENCE_HEADER) scan.parse_sequence_header(u.data, u.size);
            else if (u.type == OBU_FRAME || u.type == OBU_FRAME_HEADER) {
                BitReader br(u.data, u.size);
                FrameHdr fh;
                if (!scan.parse_frame_header(br, fh, u.temporal_id, u.spatial_id)) break;
                if (first && !fh.show_existing_frame && fh.frame_type == KEY_FRAME && fh.show_frame) starts.push_back(i);
                if (!fh.show_existing_frame) scan.reference_update(fh);
                first = false;
            }
        }
    }
    if (starts.empty() || starts[0] != 0) starts.insert(starts.begin(), 0);
    const int nseg = (int)starts.size();
    int nthreads = host_threads > 0 ? host_threads : (int)std::thread::hardware_concurrency();
    nthreads = std::max(1, std::min(nthreads, nseg));
    std::atomic<int> next{0}, rc_all{0};
    std::atomic<long long> frames{0};
    std::vector<double> parse_ms(nseg, 0.0);
    auto worker = [&]() {
        WorkerPool::nested_enabled() = nseg < nthreads;
        while (true) {
            const int s = next.fetch_add(1);
            if (s >= nseg) return;
            StreamParser sp;
            sp.host_lf_edges = false;   // as on the verify path: deblocking edges are classified on the device
            sp.hp.seq = scan.seq;
            sp.tile_threads = tile_threads != 0;
            const size_t t1 = s + 1 < nseg ? starts[s + 1] : dm.tus.size();
            for (size_t t = starts[s]; t < t1; t++) {
                std::vector<ParsedFrame> pfs;
                const int rc = sp.parse_tu(data + dm.tus[t].offset, dm.tus[t].size, dm.tus[t].pts, pfs);
                for (auto& pf : pfs) {
                    if (pf.fw) parse_ms[s] += pf.fw->parse_ms;
                    if (pf.show_existing_slot >= 0 || pf.fh.show_frame) frames++;
                }
                if (rc) { rc_all = rc; return; }
            }
        }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < nthreads; i++) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
    out->wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    for (double v : parse_ms) out->host_parse_ms += v;
    out->frames = frames;
    out->frames_per_sec = out->wall_ms > 0 ? out->frames * 1000.0 / out->wall_ms : 0;
    out->status = rc_all;
    snprintf(out->message, sizeof(out->message), "parse only: %lld frames, %d GOP segments, %d segment threads, tile threads %s", (long long)out->frames, nseg,
             nthreads, tile_threads ? "on" : "off");
    return rc_all;
}

// Host-only clip statistics (no GPU): the counts the roofline model and the bench's "does this clip really contain its config's
// tools" gate are computed from -- same fields av1r_clip_load fills, minus the device-side ones (worklist_bytes stays 0).
extern "C" int av1r_parse_stats(const uint8_t* data, size_t len, av1r_clip_info* out) {
    using namespace av1r;
    if (!data || !out) return AV1R_EINVAL;
    memset(out, 0, sizeof(*out));
    out->struct_size = sizeof(*out);
    DemuxResult dm;
    std::string derr;
    if (!demux_buffer(data, len, dm, derr)) return AV1R_EBITSTREAM;
    StreamParser sp;
    if (!dm.config_obus.empty()) {
        std::vector<ObuUnit> obus;
        if (sp.hp.split_obus(dm.config_obus.data(), dm.config_obus.size(), obus))
            for (auto& u : obus)
                if (u.type == OBU_SEQUENCE_HEADER) sp.hp.parse_sequence_header(u.data, u.size);
    }
    for (const TemporalUnit& tu : dm.tus) {
        std::vector<ParsedFrame> pfs;
        const int rc = sp.parse_tu(data + tu.offset, tu.size, tu.pts, pfs);
        if (rc) return rc;
        for (ParsedFrame& pf : pfs) {
            if (pf.show_existing_slot >= 0) { out->frames_shown++; continue; }
            const FrameWork& fw = *pf.fw;
            out->frames_decoded++;
            out->frames_shown += fw.fh.show_frame;
            out->host_parse_ms += fw.parse_ms;
            out->coded_samples += fw.coded_samples;
            out->coef_tokens += fw.coefs.size();
            out->tx_blocks += fw.tx.size();
            out->inter_samples += fw.inter_samples;
            out->inter_ref_samples += fw.inter_ref_samples;
            out->inter_blocks += fw.inter.size();
            out->obmc_neighbours += fw.obmc.size();
            out->lr_frames += fw.fh.uses_lr != 0;
            out->cdef_frames += fw.fh.enable_cdef_frame != 0;
            out->deblock_frames += (fw.fh.lf.level[0] || fw.fh.lf.level[1]);
            out->grain_frames += fw.fh.show_frame && fw.fh.fg.apply_grain;
            for (int i = 0; i < 24; i++) out->tool_hist[i] += fw.tool_hist[i];
            uint64_t isamp = 0;
            for (const TxRec& r : fw.tx)
                if (r.mode != TXM_INTER) isamp += (uint64_t)kTxW[r.txsz] * kTxH[r.txsz];
            out->intra_samples += isamp;
            if (fw.inter.empty()) {
                out->intra_frame_samples += isamp;
                out->intra_frame_coded_samples += fw.coded_samples;
                out->intra_frame_tx_blocks += fw.tx.size();
                out->intra_frames++;
            }
            out->width = fw.fh.upscaled_width;
            out->height = fw.fh.frame_height;
            out->bit_depth = sp.hp.seq.bit_depth;
            const int bps = sp.hp.seq.bit_depth == 8 ? 1 : 2;
            const int sx = sp.hp.seq.subsampling_x, sy = sp.hp.seq.subsampling_y;
            out->frame_bytes = (uint64_t)fw.fh.upscaled_width * fw.fh.frame_height * bps;
            if (!sp.hp.seq.mono_chrome) out->frame_bytes += 2ull * ((fw.fh.upscaled_width + sx) >> sx) * ((fw.fh.frame_height + sy) >> sy) * bps;
        }
    }
    return 0;
}

// Debug / test entry point (host only, no GPU): parses every frame of a container, builds the intra kernel's plan (K3 order, unit
// table, neighbour dependencies -- k3_plan.h) and checks the invariants the kernel's freedom from deadlock rests on.
// Returns 0 and the number of frames / units checked, or AV1R_EINVAL with the first violation in msg.
extern "C" int av1r_debug_k3_check(const uint8_t* data, size_t len, long long* frames, long long* units_total, char* msg, size_t cap) {
    using namespace av1r;
    if (!data || !msg || cap == 0) return AV1R_EINVAL;
    msg[0] = 0;
    DemuxResult dm;
    std::string derr;
    if (!demux_buffer(data, len, dm, derr)) {
        snprintf(msg, cap, "%s", derr.c_str());
        return AV1R_EBITSTREAM;
    }
    StreamParser sp;
    if (!dm.config_obus.empty()) {
        std::vector<ObuUnit> obus;
        if (sp.hp.split_obus(dm.config_obus.data(), dm.config_obus.size(), obus))
            for (auto& u : obus)
                if (u.type == OBU_SEQUENCE_HEADER) sp.hp.parse_sequence_header(u.data, u.size);
    }
    long long nf = 0, nu_total = 0;
    for (const TemporalUnit& tu : dm.tus) {
        std::vector<ParsedFrame> pfs;
        const int rc = sp.parse_tu(data + tu.offset, tu.size, tu.pts, pfs);
        if (rc) {
            snprintf(msg, cap, "%s", sp.err.c_str());
            return rc;
        }
        for (ParsedFrame& pf : pfs) {
            if (!pf.fw) continue;
            const FrameWork& fw = *pf.fw;
            const int n_recs = (int)fw.tx.size();
            int n_k3 = 0;
            for (const TxRec& r : fw.tx) n_k3 += k3_owns(r);
            std::vector<TxRec> recs(fw.tx);
            std::vector<uint32_t> k3((size_t)std::max(1, n_k3));
            std::vector<K3Unit> units((size_t)((fw.mi_cols + 15) >> 4) * ((fw.mi_rows + 15) >> 4) + 1);
            const int nu = k3_plan_build(fw.tx.data(), n_recs, n_k3, fw.subx, fw.suby, fw.sb128, fw.mi_cols, fw.mi_rows, recs.data(), k3.data(), units.data(), fw.fh.allow_intrabc ? 1 : 0);
            const std::string bad = k3_plan_check(fw.tx.data(), n_recs, n_k3, fw.subx, fw.suby, fw.mi_cols, fw.mi_rows, recs.data(), k3.data(), units.data(), nu);
            if (!bad.empty()) {
                snprintf(msg, cap, "frame %lld: %s", nf, bad.c_str());
                return AV1R_EINVAL;
            }
            nf++;
            nu_total += nu;
        }
    }
    if (frames) *frames = nf;
    if (units_total) *units_total = nu_total;
    return 0;
}

// Plan-level self test of the block-copy rule (CPU tests): two 64x64 units in one row, each one luma 64x64 block-copy record; the
// vector points 64 samples to the left (forward == 0: the second unit copies from the first) or to the right (forward != 0: the
// first unit would copy from the second, which does not exist yet).  Returns k3_plan_build's result.
extern "C" int av1r_debug_k3_ibc_selftest(int forward) {
    using namespace av1r;
    TxRec tx[2];
    memset(tx, 0, sizeof(tx));
    for (int i = 0; i < 2; i++) {
        tx[i].x4 = (uint16_t)(16 * i);
        tx[i].txsz = TX_64X64;
        tx[i].mode = (uint8_t)(i == (forward ? 0 : 1) ? (int)TXM_INTRABC : (int)DC_PRED);
        tx[i].cfl_max_w4 = (uint16_t)(int16_t)(forward ? 64 * 8 : -64 * 8);
    }
    TxRec recs[2] = {tx[0], tx[1]};
    uint32_t k3[2];
    K3Unit units[4];
    return k3_plan_build(tx, 2, 2, 1, 1, 0, 32, 16, recs, k3, units, 1);
}

