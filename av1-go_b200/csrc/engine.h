// Decode engine: one instance = one GPU (one av1r_ctx).  Host side parses OBUs / tiles into
// work-lists, device side reconstructs.  See DESIGN.md.
#pragma once
#include <cstdint>
#include <memory>
#include <string>

#include "../../include/av1r.h"
#include "../../include/av1r_stages.h"

#include <mutex>
#include <vector>

#include "demux.h"
#include "obu.h"

namespace av1r {

struct EngineImpl;

// One container of a verification run: what the pre-scan found and where its results go.  Shared by the engines of a batch
// (segments of one file may run on different GPUs), hence the mutex around the report.
struct VerifyFile {
    const uint8_t* data = nullptr;
    size_t len = 0;
    DemuxResult dm;
    HeaderParser scan;                  // carries the sequence header after prescan()
    std::vector<size_t> starts;         // temporal units at which an independently decodable GOP segment starts
    std::vector<SeqHdr> start_seq;      // sequence header in force at each of them
    std::vector<int64_t> frame_base;    // shown frames before each temporal unit (display index of its first shown frame)
    av1r_report* rep = nullptr;
    uint64_t* digests = nullptr;        // 3 x uint64 per shown frame, display order (may be null)
    int64_t cap_frames = 0;
    std::mutex m;
    int prescan(std::string& msg);
    const SeqHdr& seq_for(size_t tu) const {   // sequence header in force for the segment that holds temporal unit `tu`
        size_t k = 0;
        while (k + 1 < starts.size() && starts[k + 1] <= tu) k++;
        return start_seq.empty() ? scan.seq : start_seq[k];
    }
    void init_report();
    void fail(int rc, int64_t frame, const std::string& msg);
};
struct VerifyItem {
    int file;
    size_t tu0, tu1;                    // temporal units [tu0, tu1) of files[file]
    size_t bytes;                       // coded bytes (the assignment weight)
};

class Engine {
public:
    Engine();
    ~Engine();
    int open(const av1r_config& cfg);
    int submit_tu(const uint8_t* data, size_t len, int64_t pts);
    int collect(av1r_frame_result* out, int cap, int* n);
    int flush();
    int copy_frame(int64_t handle, int plane, void* dst, size_t dst_stride);
    int release_frame(int64_t handle);
    const std::string& error() const;
    // measurement: device-resident replay (include/av1r_stages.h)
    int clip_load(const uint8_t* const* tus, const size_t* lens, int n, struct ::av1r_clip** out);
    int clip_decode(struct ::av1r_clip* clip, uint64_t* cks, int cap, int* n_frames, float* device_ms);
    int clip_decode_passes(struct ::av1r_clip* clip, int passes, uint64_t* cks, int cap, int* n_frames, float* device_ms);
    int clip_profile(struct ::av1r_clip* clip, struct ::av1r_stage_times* out);
    int verify(const uint8_t* data, size_t len, av1r_report* out, uint64_t* digests, int64_t cap_frames);
    int verify_items(std::vector<VerifyFile*>& files, const std::vector<VerifyItem>& items, bool stop_on_error, int* threads_used);
    static int verify_file(const char* path, const av1r_config* cfg, av1r_report* out);
    static int verify_buffer(const uint8_t* data, size_t len, const av1r_config* cfg, av1r_report* out, uint64_t* digests,
                             int64_t cap_frames);

private:
    EngineImpl* impl_;
};

}  // namespace av1r
