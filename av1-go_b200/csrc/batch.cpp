// Multi-GPU verification inside the library: one engine (av1r_ctx equivalent) + one host thread per device, the batch's
// key-frame-delimited GOP segments assigned to the devices longest-first by coded bytes, no collective -- segments share no state
// (SURVEY 8e).  This is what the single-process Go daemon (/root/reference/cmd/av1d/main.go:311-349, one job at a time) calls to use
// a whole 8-GPU box for one job or for a queue of finished transcodes.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "../../include/av1r.h"
#include "engine.h"

using namespace av1r;

struct av1r_pool {
    std::vector<std::unique_ptr<Engine>> engines;
    std::vector<int> devices;
    av1r_config cfg;
    std::string err;
};

extern "C" int av1r_pool_open(const int* devices, int n_devices, const av1r_config* cfg, av1r_pool** out) {
    if (!out || n_devices <= 0 || n_devices > 64 || !devices) return AV1R_EINVAL;
    *out = nullptr;
    av1r_config c;
    av1r_default_config(&c);
    if (cfg) {
        size_t n = cfg->struct_size && cfg->struct_size < sizeof(c) ? cfg->struct_size : sizeof(c);
        memcpy(&c, cfg, n);
        c.struct_size = sizeof(c);
    }
    if (c.streams <= 2) c.streams = 16;
    if (c.frames_in_flight <= 8) c.frames_in_flight = 32;
    int total_threads = c.host_threads > 0 ? c.host_threads : (int)std::thread::hardware_concurrency();
    if (total_threads <= 0) total_threads = 4;
    auto pool = std::make_unique<av1r_pool>();
    pool->cfg = c;
    for (int i = 0; i < n_devices; i++)
        for (int j = 0; j < i; j++)
            if (devices[j] == devices[i]) return AV1R_EINVAL;   // one engine per device
    for (int i = 0; i < n_devices; i++) {
        av1r_config ci = c;
        ci.device = devices[i];
        // the host cores are shared by the devices: parser threads per engine = its share
        ci.host_threads = std::max(1, total_threads / n_devices);
        auto eng = std::make_unique<Engine>();
        const int rc = eng->open(ci);
        if (rc) {
            fprintf(stderr, "av1r_pool_open: device %d: %s\n", devices[i], eng->error().c_str());
            return rc;
        }
        pool->engines.push_back(std::move(eng));
        pool->devices.push_back(devices[i]);
    }
    *out = pool.release();
    return 0;
}

extern "C" void av1r_pool_close(av1r_pool* pool) { delete pool; }

extern "C" int av1r_pool_devices(const av1r_pool* pool) { return pool ? (int)pool->engines.size() : 0; }

// Longest-processing-time-first: items sorted by weight (coded bytes) descending, each to the least loaded device.
// Exposed for tests (host only): assignment[i] = device index of item i.
extern "C" int av1r_batch_assign(const uint64_t* weights, int n_items, int n_devices, int* assignment) {
    if (!weights || !assignment || n_items < 0 || n_devices <= 0) return AV1R_EINVAL;
    std::vector<int> order(n_items);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return weights[a] > weights[b]; });
    std::vector<uint64_t> load(n_devices, 0);
    for (int i : order) {
        int best = 0;
        for (int d = 1; d < n_devices; d++)
            if (load[d] < load[best]) best = d;
        assignment[i] = best;
        load[best] += weights[i];
    }
    return 0;
}

extern "C" int av1r_pool_verify_buffers(av1r_pool* pool, const uint8_t* const* data, const size_t* lens, int n_files, av1r_report* reports,
                                        uint64_t* const* digests, const int64_t* cap_frames, av1r_report* total) {
    if (!pool || !data || !lens || !reports || n_files < 0) return AV1R_EINVAL;
    const auto t0 = std::chrono::steady_clock::now();
    const int nd = (int)pool->engines.size();
    std::vector<std::unique_ptr<VerifyFile>> files(n_files);
    std::vector<VerifyFile*> fptr(n_files);
    for (int f = 0; f < n_files; f++) {
        files[f] = std::make_unique<VerifyFile>();
        VerifyFile& vf = *files[f];
        vf.data = data[f];
        vf.len = lens[f];
        vf.rep = &reports[f];
        vf.digests = digests ? digests[f] : nullptr;
        vf.cap_frames = (digests && cap_frames) ? cap_frames[f] : 0;
        vf.init_report();
        fptr[f] = &vf;
    }
    // ---- pre-scan (headers only) of all files, in parallel
    {
        std::atomic<int> next{0};
        int nt = std::max(1, std::min<int>((int)std::thread::hardware_concurrency(), n_files));
        auto scan_worker = [&]() {
            while (true) {
                const int f = next.fetch_add(1);
                if (f >= n_files) return;
                std::string msg;
                const int rc = files[f]->data ? files[f]->prescan(msg) : AV1R_EINVAL;
                if (rc) files[f]->fail(rc, -1, msg.empty() ? "null buffer" : msg);
            }
        };
        std::vector<std::thread> th;
        for (int i = 1; i < nt; i++) th.emplace_back(scan_worker);
        scan_worker();
        for (auto& t : th) t.join();
    }
    // ---- work items = GOP segments of every readable file; longest-first assignment by coded bytes
    std::vector<VerifyItem> items;
    for (int f = 0; f < n_files; f++) {
        VerifyFile& vf = *files[f];
        if (vf.rep->status) continue;
        for (size_t s = 0; s < vf.starts.size(); s++) {
            VerifyItem it{f, vf.starts[s], s + 1 < vf.starts.size() ? vf.starts[s + 1] : vf.dm.tus.size(), 0};
            for (size_t t = it.tu0; t < it.tu1; t++) it.bytes += vf.dm.tus[t].size;
            if (it.tu1 > it.tu0) items.push_back(it);
        }
    }
    std::vector<uint64_t> w(items.size());
    for (size_t i = 0; i < items.size(); i++) w[i] = items[i].bytes;
    std::vector<int> assign(items.size(), 0);
    av1r_batch_assign(w.data(), (int)items.size(), nd, assign.data());
    std::vector<std::vector<VerifyItem>> per_dev(nd);
    for (size_t i = 0; i < items.size(); i++) per_dev[assign[i]].push_back(items[i]);   // items are in (file, segment) order already
    // ---- one host thread per device
    std::vector<int> rcs(nd, 0);
    std::vector<std::thread> th;
    for (int d = 0; d < nd; d++)
        th.emplace_back([&, d]() {
            int used = 0;
            if (!per_dev[d].empty()) rcs[d] = pool->engines[d]->verify_items(fptr, per_dev[d], false, &used);
        });
    for (auto& t : th) t.join();
    const double wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    int first_rc = 0;
    for (int d = 0; d < nd; d++)
        if (rcs[d] && !first_rc) {
            first_rc = rcs[d];
            pool->err = pool->engines[d]->error();
        }
    av1r_report tot;
    memset(&tot, 0, sizeof(tot));
    tot.struct_size = sizeof(tot);
    tot.first_bad_frame = -1;
    for (int f = 0; f < n_files; f++) {
        av1r_report& r = reports[f];
        if (!r.status && first_rc && r.frames < files[f]->frame_base.back()) {   // a device died under this file
            r.status = first_rc;
            snprintf(r.message, sizeof(r.message), "device failure: %s", pool->err.c_str());
        }
        if (!r.status && r.frames != files[f]->frame_base.back()) {
            r.status = AV1R_EBITSTREAM;
            snprintf(r.message, sizeof(r.message), "decoded %lld of %lld frames", (long long)r.frames, (long long)files[f]->frame_base.back());
        }
        r.wall_ms = wall;
        r.frames_per_sec = wall > 0 ? r.frames * 1000.0 / wall : 0;
        if (!r.status)
            snprintf(r.message, sizeof(r.message), "ok: %lld frames, %zu GOP segments over %d GPU(s)", (long long)r.frames, files[f]->starts.size(), nd);
        tot.frames += r.frames;
        tot.host_parse_ms += r.host_parse_ms;
        tot.device_ms += r.device_ms;
        if (r.status && !tot.status) {
            tot.status = r.status;
            tot.first_bad_frame = f;   // for the aggregate: index of the first failing *file*
            snprintf(tot.message, sizeof(tot.message), "file %d: %.480s", f, r.message);
        }
        tot.width = r.width ? r.width : tot.width;
        tot.height = r.height ? r.height : tot.height;
        tot.bit_depth = r.bit_depth ? r.bit_depth : tot.bit_depth;
    }
    tot.wall_ms = wall;
    tot.frames_per_sec = wall > 0 ? tot.frames * 1000.0 / wall : 0;
    if (!tot.status)
        snprintf(tot.message, sizeof(tot.message), "ok: %d files, %lld frames, %zu GOP segments over %d GPU(s)", n_files, (long long)tot.frames, items.size(), nd);
    if (total) *total = tot;
    return tot.status;
}

static bool read_whole(const char* path, std::vector<uint8_t>& buf) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize(n > 0 ? (size_t)n : 0);
    const size_t got = n > 0 ? fread(buf.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    return (long)got == n;
}

extern "C" int av1r_pool_verify_files(av1r_pool* pool, const char* const* paths, int n_files, av1r_report* reports, av1r_report* total) {
    if (!pool || !paths || !reports || n_files < 0) return AV1R_EINVAL;
    std::vector<std::vector<uint8_t>> bufs(n_files);
    std::vector<const uint8_t*> ptrs(n_files, nullptr);
    std::vector<size_t> lens(n_files, 0);
    std::vector<int> unreadable(n_files, 0);
    for (int f = 0; f < n_files; f++) {
        if (paths[f] && read_whole(paths[f], bufs[f]) && !bufs[f].empty()) {
            ptrs[f] = bufs[f].data();
            lens[f] = bufs[f].size();
        } else {
            unreadable[f] = 1;
        }
    }
    const int rc = av1r_pool_verify_buffers(pool, ptrs.data(), lens.data(), n_files, reports, nullptr, nullptr, total);
    for (int f = 0; f < n_files; f++)
        if (unreadable[f]) {
            reports[f].status = AV1R_ENOENT;
            snprintf(reports[f].message, sizeof(reports[f].message), "cannot read %s", paths[f] ? paths[f] : "(null)");
            if (total && total->status == AV1R_EINVAL && total->first_bad_frame == f) total->status = AV1R_ENOENT;
        }
    return total ? total->status : rc;
}

// Convenience: open engines on `devices`, verify the files, close.  reports: n_files entries.
extern "C" int av1r_verify_batch(const char* const* paths, int n_files, const int* devices, int n_devices, const av1r_config* cfg,
                                 av1r_report* reports, av1r_report* total) {
    av1r_pool* pool = nullptr;
    const int rc = av1r_pool_open(devices, n_devices, cfg, &pool);
    if (rc) {
        for (int f = 0; f < n_files && reports; f++) {
            memset(&reports[f], 0, sizeof(reports[f]));
            reports[f].struct_size = sizeof(reports[f]);
            reports[f].status = rc;
            reports[f].first_bad_frame = -1;
            snprintf(reports[f].message, sizeof(reports[f].message), "cannot open the verification engines (rc %d)", rc);
        }
        return rc;
    }
    const int vrc = av1r_pool_verify_files(pool, paths, n_files, reports, total);
    av1r_pool_close(pool);
    return vrc;
}
