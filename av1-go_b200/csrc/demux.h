// Container demux -> temporal units.  IVF (what the synthetic generator emits), raw Section-5
// OBU streams (split at temporal delimiters) and Matroska (the artefact the daemon really
// produces: `-f matroska`, /root/reference/internal/ffmpeg/transcode.go:143,
// `<base>.av1-tmp.mkv`, /root/reference/internal/daemon/daemon.go:86).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace av1r {

struct TemporalUnit {
    size_t offset, size;   // into the file buffer (or into `extra` when offset == SIZE_MAX)
    int64_t pts;
};

struct DemuxResult {
    std::vector<uint8_t> file;            // whole file
    std::vector<TemporalUnit> tus;
    std::vector<uint8_t> config_obus;     // Matroska CodecPrivate (av1C) config OBUs, may be empty
    std::string container;                // "ivf" | "obu" | "matroska"
};

bool demux_buffer(const uint8_t* data, size_t len, DemuxResult& out, std::string& err);
bool demux_file(const char* path, DemuxResult& out, std::string& err);

}  // namespace av1r
