#include "demux.h"

#include <cstdio>
#include <cstring>

namespace av1r {

static uint32_t rd32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint64_t rd64(const uint8_t* p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

static bool demux_ivf(const uint8_t* d, size_t n, DemuxResult& out, std::string& err) {
    if (n < 32) { err = "ivf: truncated header"; return false; }
    size_t hl = d[6] | (d[7] << 8);
    if (hl < 32 || hl > n) { err = "ivf: bad header length"; return false; }
    if (memcmp(d + 8, "AV01", 4) != 0) { err = "ivf: fourcc is not AV01"; return false; }
    size_t pos = hl;
    while (pos + 12 <= n) {
        uint32_t sz = rd32(d + pos);
        int64_t pts = (int64_t)rd64(d + pos + 4);
        pos += 12;
        if (sz > n - pos) { err = "ivf: truncated frame"; return false; }
        if (sz == 0) { err = "ivf: empty frame"; return false; }   // a temporal unit holds at least a temporal delimiter
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
        if (p > n || sz > n - p) { err = "obu: truncated"; return false; }
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
// The file is untrusted (it is the possibly-broken output the verifier exists to catch): every element end is clamped to its
// parent's end and to the buffer, sizes are checked before they are added to a position, and a block is only recorded when its
// payload lies inside the buffer.
namespace {

struct Ebml {
    const uint8_t* d;
    size_t n;
    // element id at pos (1..4 bytes), pos advances; false at end / malformed
    bool read_id(size_t& pos, size_t end, uint32_t& id) const {
        if (pos >= end) return false;
        const uint8_t b = d[pos];
        const int len = b & 0x80 ? 1 : b & 0x40 ? 2 : b & 0x20 ? 3 : b & 0x10 ? 4 : 0;
        if (!len || end - pos < (size_t)len) return false;
        id = 0;
        for (int i = 0; i < len; i++) id = (id << 8) | d[pos + i];
        pos += len;
        return true;
    }
    // EBML variable-length integer (element size, track number, lace size); `unknown` = all value bits set
    bool read_vint(size_t& pos, size_t end, uint64_t& v, bool& unknown, int* nbytes = nullptr) const {
        if (pos >= end) return false;
        const uint8_t b = d[pos];
        int len = 0;
        for (int i = 0; i < 8; i++)
            if (b & (0x80 >> i)) { len = i + 1; break; }
        if (!len || end - pos < (size_t)len) return false;
        uint64_t x = b & (0xff >> len);
        bool all1 = x == (uint64_t)(0xff >> len);
        for (int i = 1; i < len; i++) {
            x = (x << 8) | d[pos + i];
            if (d[pos + i] != 0xff) all1 = false;
        }
        pos += len;
        v = x;
        unknown = all1;
        if (nbytes) *nbytes = len;
        return true;
    }
    // header of the child element at pos inside [pos, parent_end): id, payload range [pos, el_end).  An unknown-size child
    // extends to the parent's end.  false = no further well-formed child.
    bool child(size_t& pos, size_t parent_end, uint32_t& id, size_t& el_end, bool& unknown) const {
        uint64_t sz;
        if (!read_id(pos, parent_end, id) || !read_vint(pos, parent_end, sz, unknown)) return false;
        if (unknown) el_end = parent_end;
        else if (sz > (uint64_t)(parent_end - pos)) return false;   // child sticks out of its parent (or of the buffer)
        else el_end = pos + (size_t)sz;
        return true;
    }
    uint64_t read_uint(size_t pos, size_t end) const {   // big-endian unsigned of up to 8 bytes
        uint64_t v = 0;
        for (size_t i = pos; i < end && i < pos + 8; i++) v = (v << 8) | d[i];
        return v;
    }
};

enum : uint32_t {
    ID_EBML = 0x1A45DFA3, ID_SEGMENT = 0x18538067, ID_TRACKS = 0x1654AE6B, ID_TRACK_ENTRY = 0xAE, ID_TRACK_NUMBER = 0xD7,
    ID_CODEC_ID = 0x86, ID_CODEC_PRIVATE = 0x63A2, ID_CLUSTER = 0x1F43B675, ID_TIMESTAMP = 0xE7, ID_SIMPLE_BLOCK = 0xA3,
    ID_BLOCK_GROUP = 0xA0, ID_BLOCK = 0xA1, ID_CUES = 0x1C53BB6B, ID_TAGS = 0x1254C367, ID_SEEK_HEAD = 0x114D9B74,
    ID_INFO = 0x1549A966, ID_CHAPTERS = 0x1043A770, ID_ATTACHMENTS = 0x1941A469,
};

static bool is_top_level(uint32_t id) {
    return id == ID_CLUSTER || id == ID_CUES || id == ID_TAGS || id == ID_SEEK_HEAD || id == ID_INFO || id == ID_TRACKS ||
           id == ID_CHAPTERS || id == ID_ATTACHMENTS;
}

// (Simple)Block payload [bp, bend): track number, 16-bit relative time, flags, then one frame or a lace of frames.
static bool take_block(const Ebml& e, size_t bp, size_t bend, int64_t av1_track, int64_t cluster_ts, DemuxResult& out, std::string& err) {
    uint64_t tn;
    bool tunk;
    size_t q = bp;
    if (!e.read_vint(q, bend, tn, tunk) || bend - q < 3) { err = "mkv: truncated block header"; return false; }
    const int16_t rel = (int16_t)((e.d[q] << 8) | e.d[q + 1]);
    const uint8_t flags = e.d[q + 2];
    q += 3;
    if ((int64_t)tn != av1_track) return true;
    const int lacing = (flags >> 1) & 3;
    const int64_t pts = cluster_ts + rel;
    if (!lacing) {
        if (bend > q) out.tus.push_back({q, bend - q, pts});
        return true;
    }
    // laced block: frame count - 1, then the sizes of all frames but the last (Xiph 1, fixed 2, EBML 3)
    if (q >= bend) { err = "mkv: truncated lace header"; return false; }
    const int nfr = e.d[q++] + 1;
    std::vector<uint64_t> sizes((size_t)nfr, 0);
    if (lacing == 1) {
        for (int i = 0; i < nfr - 1; i++) {
            uint64_t s = 0;
            while (true) {
                if (q >= bend) { err = "mkv: truncated Xiph lace"; return false; }
                const uint8_t b = e.d[q++];
                s += b;
                if (b != 255) break;
            }
            sizes[i] = s;
        }
    } else if (lacing == 3) {
        int64_t prev = 0;
        for (int i = 0; i < nfr - 1; i++) {
            uint64_t v;
            bool u;
            int nb = 0;
            if (!e.read_vint(q, bend, v, u, &nb)) { err = "mkv: truncated EBML lace"; return false; }
            int64_t s;
            if (i == 0) s = (int64_t)v;
            else s = prev + ((int64_t)v - (((int64_t)1 << (7 * nb - 1)) - 1));   // signed difference to the previous size
            if (s < 0) { err = "mkv: negative EBML lace size"; return false; }
            sizes[i] = (uint64_t)s;
            prev = s;
        }
    }
    const size_t payload = bend - q;
    if (lacing == 2) {
        if (payload % (size_t)nfr) { err = "mkv: fixed lace does not divide the block"; return false; }
        for (auto& s : sizes) s = payload / (size_t)nfr;
    } else {
        uint64_t sum = 0;
        for (int i = 0; i < nfr - 1; i++) {
            if (sizes[i] > payload - sum) { err = "mkv: lace sizes exceed the block"; return false; }
            sum += sizes[i];
        }
        sizes[nfr - 1] = payload - sum;
    }
    for (int i = 0; i < nfr; i++) {
        if (sizes[i]) out.tus.push_back({q, (size_t)sizes[i], pts + i});
        q += (size_t)sizes[i];
    }
    return true;
}

}  // namespace

static bool demux_mkv(const uint8_t* d, size_t n, DemuxResult& out, std::string& err) {
    const Ebml e{d, n};
    size_t pos = 0, end;
    uint32_t id;
    bool unk;
    if (!e.child(pos, n, id, end, unk) || id != ID_EBML || unk) { err = "mkv: no EBML header"; return false; }
    pos = end;
    if (!e.child(pos, n, id, end, unk) || id != ID_SEGMENT) { err = "mkv: no Segment"; return false; }
    const size_t seg_end = end;
    int64_t av1_track = -1;
    while (pos < seg_end) {
        if (!e.child(pos, seg_end, id, end, unk)) break;    // trailing garbage / truncated element: keep what was found
        if (id == ID_TRACKS) {
            size_t p = pos, tend;
            uint32_t tid;
            bool tu;
            while (p < end && e.child(p, end, tid, tend, tu)) {
                if (tid == ID_TRACK_ENTRY) {
                    int64_t num = -1;
                    bool is_av1 = false;
                    size_t cp_off = 0, cp_end = 0, q = p, fend;
                    uint32_t fid;
                    bool fu;
                    while (q < tend && e.child(q, tend, fid, fend, fu)) {
                        if (fid == ID_TRACK_NUMBER) num = (int64_t)e.read_uint(q, fend);
                        else if (fid == ID_CODEC_ID) is_av1 = fend - q == 5 && memcmp(d + q, "V_AV1", 5) == 0;
                        else if (fid == ID_CODEC_PRIVATE) { cp_off = q; cp_end = fend; }
                        q = fend;
                    }
                    if (is_av1 && av1_track < 0 && num >= 0) {
                        av1_track = num;
                        if (cp_end - cp_off > 4) out.config_obus.assign(d + cp_off + 4, d + cp_end);   // skip the 4-byte av1C header
                    }
                }
                p = tend;
            }
        } else if (id == ID_CLUSTER) {
            int64_t cluster_ts = 0;
            size_t p = pos, cend;
            uint32_t cid;
            bool cu;
            while (p < end) {
                const size_t save = p;
                if (!e.child(p, end, cid, cend, cu)) {
                    if (!unk) { err = "mkv: truncated cluster"; return false; }
                    end = save;
                    break;
                }
                if (unk && is_top_level(cid)) { end = save; break; }   // an unknown-size cluster ends at the next top-level element
                if (cid == ID_TIMESTAMP) cluster_ts = (int64_t)e.read_uint(p, cend);
                else if (cid == ID_SIMPLE_BLOCK) { if (!take_block(e, p, cend, av1_track, cluster_ts, out, err)) return false; }
                else if (cid == ID_BLOCK_GROUP) {
                    size_t q = p, gend;
                    uint32_t gid;
                    bool gu;
                    while (q < cend && e.child(q, cend, gid, gend, gu)) {
                        if (gid == ID_BLOCK && !take_block(e, q, gend, av1_track, cluster_ts, out, err)) return false;
                        q = gend;
                    }
                }
                p = cend;
            }
        }
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
