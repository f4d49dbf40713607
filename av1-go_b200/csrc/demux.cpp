#include "demux.h"

#include <cstdio>
#include <cstring>

namespace av1r {

static uint32_t rd32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint64_t rd64(const uint8_t* p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

static bool demux_ivf(const uint8_t* d, size_t n, DemuxResult& out, std::string& err) {
    if (n < 32) { err = "ivf: truncated header"; return false; }
    size_t hl = d[6] | (d[7] << 8);
    if (memcmp(d + 8, "AV01", 4) != 0) { err = "ivf: fourcc is not AV01"; return false; }
    size_t pos = hl;
    while (pos + 12 <= n) {
        uint32_t sz = rd32(d + pos);
        int64_t pts = (int64_t)rd64(d + pos + 4);
        pos += 12;
        if (pos + sz > n) { err = "ivf: truncated frame"; return false; }
        out.tus.push_back({pos, sz, pts});
        pos += sz;
    }
    out.container = "ivf";
    return true;
}

// Raw low-overhead OBU stream: a temporal unit starts at every temporal delimiter OBU.
static bool demux_obu(const uint8_t* d, size_t n, DemuxResult& out, std::string& err) {
    size_t pos = 0, tu_start = 0;
    int64_t pts = 0;
    bool any = false;
    while (pos < n) {
        uint8_t h = d[pos];
        if (h & 0x80) { err = "obu: forbidden bit"; return false; }
        int type = (h >> 3) & 15, ext = (h >> 2) & 1, has_size = (h >> 1) & 1;
        if (!has_size) { err = "obu: stream without size fields"; return false; }
        size_t p = pos + 1 + ext;
        uint64_t sz = 0;
        int i = 0;
        for (; i < 8 && p < n; i++) {
            uint8_t b = d[p++];
            sz |= (uint64_t)(b & 0x7f) << (7 * i);
            if (!(b & 0x80)) break;
        }
        if (p + sz > n) { err = "obu: truncated"; return false; }
        if (type == 2 && any) {
            out.tus.push_back({tu_start, pos - tu_start, pts++});
            tu_start = pos;
        }
        any = true;
        pos = p + sz;
    }
    if (pos > tu_start) out.tus.push_back({tu_start, pos - tu_start, pts});
    out.container = "obu";
    return true;
}

// ---- Matroska (EBML) ------------------------------------------------------------------------
struct Ebml {
    const uint8_t* d;
    size_t n;
    bool read_id(size_t& pos, uint32_t& id) const {
        if (pos >= n) return false;
        uint8_t b = d[pos];
        int len = b & 0x80 ? 1 : b & 0x40 ? 2 : b & 0x20 ? 3 : b & 0x10 ? 4 : 0;
        if (!len || pos + len > n) return false;
        id = 0;
        for (int i = 0; i < len; i++) id = (id << 8) | d[pos + i];
        pos += len;
        return true;
    }
    bool read_size(size_t& pos, uint64_t& sz, bool& unknown) const {
        if (pos >= n) return false;
        uint8_t b = d[pos];
        int len = 0;
        for (int i = 0; i < 8; i++)
            if (b & (0x80 >> i)) { len = i + 1; break; }
        if (!len || pos + len > n) return false;
        uint64_t v = b & (0xff >> len);
        bool all1 = v == (uint64_t)(0xff >> len);
        for (int i = 1; i < len; i++) {
            v = (v << 8) | d[pos + i];
            if (d[pos + i] != 0xff) all1 = false;
        }
        pos += len;
        sz = v;
        unknown = all1;
        return true;
    }
    uint64_t read_uint(size_t pos, uint64_t sz) const {
        uint64_t v = 0;
        for (uint64_t i = 0; i < sz && i < 8; i++) v = (v << 8) | d[pos + i];
        return v;
    }
};

static bool demux_mkv(const uint8_t* d, size_t n, DemuxResult& out, std::string& err) {
    Ebml e{d, n};
    size_t pos = 0;
    uint32_t id;
    uint64_t sz;
    bool unk;
    if (!e.read_id(pos, id) || id != 0x1A45DFA3 || !e.read_size(pos, sz, unk)) { err = "mkv: no EBML header"; return false; }
    pos += sz;
    if (!e.read_id(pos, id) || id != 0x18538067 || !e.read_size(pos, sz, unk)) { err = "mkv: no Segment"; return false; }
    size_t seg_end = unk ? n : (pos + sz > n ? n : pos + (size_t)sz);
    int64_t av1_track = -1;
    int64_t pts_fallback = 0;
    while (pos < seg_end) {
        size_t el = pos;
        if (!e.read_id(pos, id) || !e.read_size(pos, sz, unk)) break;
        size_t end = unk ? seg_end : pos + (size_t)sz;
        if (end > seg_end) end = seg_end;
        if (id == 0x1654AE6B) {  // Tracks
            size_t p = pos;
            while (p < end) {
                uint32_t tid; uint64_t tsz; bool tu;
                if (!e.read_id(p, tid) || !e.read_size(p, tsz, tu)) break;
                size_t tend = p + (size_t)tsz;
                if (tid == 0xAE) {  // TrackEntry
                    int64_t num = -1;
                    bool is_av1 = false;
                    size_t cp_off = 0, cp_sz = 0;
                    size_t q = p;
                    while (q < tend) {
                        uint32_t fid; uint64_t fsz; bool fu;
                        if (!e.read_id(q, fid) || !e.read_size(q, fsz, fu)) break;
                        if (fid == 0xD7) num = (int64_t)e.read_uint(q, fsz);
                        else if (fid == 0x86) is_av1 = fsz == 5 && memcmp(d + q, "V_AV1", 5) == 0;
                        else if (fid == 0x63A2) { cp_off = q; cp_sz = (size_t)fsz; }
                        q += (size_t)fsz;
                    }
                    if (is_av1 && av1_track < 0) {
                        av1_track = num;
                        if (cp_sz > 4) out.config_obus.assign(d + cp_off + 4, d + cp_off + cp_sz);  // skip av1C 4-byte header
                    }
                }
                p = tend;
            }
        } else if (id == 0x1F43B675) {  // Cluster
            int64_t cluster_ts = 0;
            size_t p = pos;
            while (p < end) {
                size_t save = p;
                uint32_t cid; uint64_t csz; bool cu;
                if (!e.read_id(p, cid) || !e.read_size(p, csz, cu)) break;
                if (unk && (cid == 0x1F43B675 || cid == 0x1C53BB6B || cid == 0x1254C367)) { end = save; break; }
                size_t cend = p + (size_t)csz;
                if (cend > n) { err = "mkv: truncated cluster"; return false; }
                auto take_block = [&](size_t bp, size_t bend) -> bool {
                    uint64_t tn; bool tunk; size_t q = bp;
                    if (!e.read_size(q, tn, tunk) || q + 3 > bend) return false;
                    int16_t rel = (int16_t)((d[q] << 8) | d[q + 1]);
                    uint8_t flags = d[q + 2];
                    q += 3;
                    if ((int64_t)tn != av1_track) return true;
                    if (flags & 0x06) { err = "mkv: laced blocks unsupported"; return false; }
                    out.tus.push_back({q, bend - q, cluster_ts + rel});
                    return true;
                };
                if (cid == 0xE7) cluster_ts = (int64_t)e.read_uint(p, csz);
                else if (cid == 0xA3) { if (!take_block(p, cend)) { if (err.empty()) err = "mkv: bad SimpleBlock"; return false; } }
                else if (cid == 0xA0) {
                    size_t q = p;
                    while (q < cend) {
                        uint32_t gid; uint64_t gsz; bool gu;
                        if (!e.read_id(q, gid) || !e.read_size(q, gsz, gu)) break;
                        if (gid == 0xA1) { if (!take_block(q, q + (size_t)gsz)) { if (err.empty()) err = "mkv: bad Block"; return false; } }
                        q += (size_t)gsz;
                    }
                }
                p = cend;
            }
            if (unk) { pos = end; continue; }
        }
        (void)el;
        (void)pts_fallback;
        pos = end;
    }
    if (av1_track < 0) { err = "mkv: no V_AV1 track"; return false; }
    out.container = "matroska";
    return true;
}

bool demux_buffer(const uint8_t* data, size_t len, DemuxResult& out, std::string& err) {
    out.tus.clear();
    out.config_obus.clear();
    if (len >= 4 && memcmp(data, "DKIF", 4) == 0) return demux_ivf(data, len, out, err);
    if (len >= 4 && data[0] == 0x1A && data[1] == 0x45 && data[2] == 0xDF && data[3] == 0xA3) return demux_mkv(data, len, out, err);
    return demux_obu(data, len, out, err);
}

bool demux_file(const char* path, DemuxResult& out, std::string& err) {
    FILE* f = fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return false; }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.file.resize(n > 0 ? n : 0);
    if (n > 0 && fread(out.file.data(), 1, n, f) != (size_t)n) { fclose(f); err = "short read"; return false; }
    fclose(f);
    return demux_buffer(out.file.data(), out.file.size(), out, err);
}

}  // namespace av1r
