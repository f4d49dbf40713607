// MSB-first bit reader for AV1 OBU headers (f(n), leb128, uvlc, su, ns, le).
// Written from the AV1 bitstream syntax (spec section 4.10); no reference-repo counterpart
// exists (the reference shells out to ffmpeg: /root/reference/internal/ffmpeg/transcode.go:195).
#pragma once
#include <cstddef>
#include <cstdint>

namespace av1r {

struct BitReader {
    const uint8_t* p = nullptr;
    size_t len = 0;     // bytes
    size_t pos = 0;     // bit position
    bool err = false;

    BitReader() {}
    BitReader(const uint8_t* d, size_t n) : p(d), len(n) {}

    inline uint32_t bit() {
        if ((pos >> 3) >= len) { err = true; return 0; }
        uint32_t b = (p[pos >> 3] >> (7 - (pos & 7))) & 1;
        pos++;
        return b;
    }
    inline uint32_t f(int n) {
        uint32_t v = 0;
        for (int i = 0; i < n; i++) v = (v << 1) | bit();
        return v;
    }
    inline int32_t su(int n) {
        int32_t v = (int32_t)f(n);
        int32_t sign = 1 << (n - 1);
        if (v & sign) v -= 2 * sign;
        return v;
    }
    inline uint32_t ns(uint32_t n) {
        int w = 0;
        uint32_t x = n;
        while (x) { x >>= 1; w++; }
        uint32_t m = (1u << w) - n;
        uint32_t v = f(w - 1);
        if (v < m) return v;
        uint32_t extra = bit();
        return (v << 1) - m + extra;
    }
    inline uint32_t uvlc() {
        int lz = 0;
        while (true) {
            if (bit()) break;
            lz++;
            if (lz >= 32 || err) return 0xFFFFFFFFu;
        }
        if (lz >= 32) return 0xFFFFFFFFu;
        return f(lz) + ((1u << lz) - 1);
    }
    inline uint32_t le(int nbytes) {
        uint32_t v = 0;
        for (int i = 0; i < nbytes; i++) v |= f(8) << (8 * i);
        return v;
    }
    inline uint64_t leb128() {
        uint64_t v = 0;
        for (int i = 0; i < 8; i++) {
            uint32_t b = f(8);
            v |= (uint64_t)(b & 0x7f) << (i * 7);
            if (!(b & 0x80)) break;
        }
        return v;
    }
    inline void byte_align() { pos = (pos + 7) & ~(size_t)7; }
    inline size_t byte_pos() const { return (pos + 7) >> 3; }
};

}  // namespace av1r
