// Multi-symbol adaptive arithmetic decoder (AV1 spec 8.2: init_symbol / read_symbol /
// read_bool / read_literal / exit_symbol) -- host side, sequential by nature.
// CDFs are stored inverted (32768 - cdf) with the adaptation counter right after the symbols.
#pragma once
#include <cstddef>
#include <cstdint>
#if defined(__SSE2__)
#include <emmintrin.h>
#define AV1R_MSAC_SSE2 1
#endif

namespace av1r {

struct Msac {
    const uint8_t* bptr = nullptr;
    const uint8_t* end = nullptr;
    uint64_t dif = 0;
    uint32_t rng = 0;
    int cnt = 0;
    bool update = true;

    void init(const uint8_t* data, size_t sz, bool disable_cdf_update) {
        bptr = data;
        end = data + sz;
        dif = ((uint64_t)1 << 63) - 1;
        rng = 0x8000;
        cnt = -15;
        update = !disable_cdf_update;
        refill();
    }
    // (forced inline, and the rare tail path works on a copy: a decoder held in a local variable never has its address taken, so
    // its window / range / count stay in registers across the byte and vector stores of the coefficient loop)
    __attribute__((always_inline)) inline void refill() {
        int s = 64 - 9 - (cnt + 15);
        if (__builtin_expect(s >= 0 && end - bptr >= 8, 1)) {
            // whole bytes that fit below the window: one big-endian 64-bit load instead of a byte loop
            uint64_t v;
            __builtin_memcpy(&v, bptr, 8);
            v = __builtin_bswap64(v);
            const int n = (s >> 3) + 1;                 // 1 .. 7 bytes (s <= 55)
            dif ^= (v >> (64 - 8 * n)) << (s - 8 * (n - 1));
            cnt += 8 * n;
            bptr += n;
            return;
        }
        Msac t = *this;
        t.refill_tail(s);
        *this = t;
    }
    __attribute__((noinline)) void refill_tail(int s) {
        for (; s >= 0 && bptr < end; s -= 8, bptr++) {
            dif ^= (uint64_t)bptr[0] << s;
            cnt += 8;
        }
        if (bptr >= end) cnt = 0x4000;  // out of data: keep shifting in the implicit padding
    }
    __attribute__((always_inline)) inline void normalize(uint64_t d, uint32_t r) {
        int sh = 15 - (31 - __builtin_clz(r));   // 16 - ilog(r)
        cnt -= sh;
        dif = ((d + 1) << sh) - 1;
        rng = r << sh;
        if (cnt < 0) refill();
    }
    // decode one symbol out of n with inverted cdf `c` (c[n-1] == 0, c[n] = counter); adapts.
    // n is a compile-time constant at every call site; small alphabets (the coefficient symbols, > 90 % of
    // all symbols) take a branch-free path: entropy-coded symbols are unpredictable by construction, so a
    // compare-and-count beats the data-dependent search loop.
    __attribute__((always_inline)) inline int symbol(uint16_t* c, int n) {
#ifdef AV1R_MSAC_SSE2
        if (n == 3 || n == 4) return symbol_v4(c, n);
        if (n >= 5 && n <= 16) return symbol_v16(c, n);
#endif
        const uint32_t r = rng;
        const uint32_t v16 = (uint32_t)(dif >> 48);
        const uint32_t r8 = r >> 8;
        int s;
        uint32_t u, v;
        if (n <= 4) {
            uint32_t t[4];
            t[0] = ((r8 * (uint32_t)(c[0] >> 6)) >> 1) + 4 * (uint32_t)(n - 1);
            t[1] = n > 2 ? ((r8 * (uint32_t)(c[1] >> 6)) >> 1) + 4 * (uint32_t)(n - 2) : 0;
            t[2] = n > 3 ? ((r8 * (uint32_t)(c[2] >> 6)) >> 1) + 4 * (uint32_t)(n - 3) : 0;
            t[3] = 0;
            s = (v16 < t[0]) + (n > 2 ? (v16 < t[1]) : 0) + (n > 3 ? (v16 < t[2]) : 0);
            u = s ? t[s - 1] : r;
            v = t[s];
        } else {
            v = r;
            s = -1;
            do {
                u = v;
                s++;
                v = ((r8 * (uint32_t)(c[s] >> 6)) >> 1) + 4 * (uint32_t)(n - 1 - s);
            } while (v16 < v);
        }
        normalize(dif - ((uint64_t)v << 48), u - v);
        if (update) adapt(c, s, n);
        return s;
    }
#ifdef AV1R_MSAC_SSE2
    // 3- and 4-symbol alphabets (coefficient base levels, base-range increments, end-of-block base: ~90 % of all symbols):
    // the three candidate split points, the comparison with the window and the probability adaptation run on four 16-bit lanes.
    // The 64-bit load/store covers c[0..3]: probabilities, the zero sentinel c[n-1] and (n == 3) the adaptation counter, which
    // the lane mask leaves untouched.
    __attribute__((always_inline)) inline int symbol_v4(uint16_t* c, int n) {
        const uint32_t r = rng;
        const uint32_t v16 = (uint32_t)(dif >> 48);
        const __m128i cv = _mm_loadl_epi64(reinterpret_cast<const __m128i*>(c));
        const __m128i en = n == 4 ? _mm_set_epi16(0, 0, 0, 0, 0, -1, -1, -1) : _mm_set_epi16(0, 0, 0, 0, 0, 0, -1, -1);   // lanes < n - 1
        const __m128i mp = n == 4 ? _mm_set_epi16(0, 0, 0, 0, 0, 4, 8, 12) : _mm_set_epi16(0, 0, 0, 0, 0, 0, 4, 8);        // 4 * (n - 1 - i)
        // ((r >> 8) * (c >> 6)) >> 1 == mulhi(r & 0xff00, (c >> 6) << 7)
        __m128i v = _mm_mulhi_epu16(_mm_set1_epi16((short)(r & 0xff00)), _mm_slli_epi16(_mm_srli_epi16(cv, 6), 7));
        v = _mm_and_si128(_mm_add_epi16(v, mp), en);
        // lanes with v <= window (unsigned): the first one is the symbol (lane n - 1 holds 0, so there always is one)
        const __m128i le = _mm_cmpeq_epi16(_mm_subs_epu16(v, _mm_set1_epi16((short)v16)), _mm_setzero_si128());
        const int s = __builtin_ctz((unsigned)_mm_movemask_epi8(le) | 0x100u) >> 1;
        // lanes s and s - 1 (lane -1 = the current range) by shifts of the 64-bit lane image: no store-to-load round trip
        const uint64_t vq = (uint64_t)_mm_cvtsi128_si64(v);
        const uint32_t vv = (uint32_t)(vq >> (16 * s)) & 0xffffu;
        const uint32_t u = (uint32_t)((((vq << 16) | r) >> (16 * s)) & 0xffffu);
        normalize(dif - ((uint64_t)vv << 48), u - vv);
        if (update) {
            const int cnt_ = c[n];
            const __m128i rate = _mm_cvtsi32_si128(3 + (cnt_ > 15) + (cnt_ > 31) + (n > 3 ? 2 : 1));
            const __m128i up = _mm_add_epi16(cv, _mm_srl_epi16(_mm_sub_epi16(_mm_set1_epi16((short)0x8000), cv), rate));
            const __m128i dn = _mm_sub_epi16(cv, _mm_srl_epi16(cv, rate));
            const __m128i lt = _mm_andnot_si128(le, en);                       // lanes i < s
            __m128i nv = _mm_or_si128(_mm_and_si128(lt, up), _mm_andnot_si128(lt, dn));
            nv = _mm_or_si128(_mm_and_si128(en, nv), _mm_andnot_si128(en, cv));   // sentinel / counter lanes keep their value
            _mm_storel_epi64(reinterpret_cast<__m128i*>(c), nv);
            c[n] = (uint16_t)(cnt_ + (cnt_ < 32));
        }
        return s;
    }
#endif
#ifdef AV1R_MSAC_SSE2
    // 5- to 16-symbol alphabets (partition, intra modes, transform types, end-of-block classes, motion vector classes ...): the same
    // scheme on two vectors of eight 16-bit lanes.  The loads and stores cover c[0..15], i.e. they reach past c[n] into whatever
    // follows this CDF (the next CDF of the context, or the pad behind it: tile.h); lanes >= n - 1 are written back unchanged.
    __attribute__((always_inline)) inline int symbol_v16(uint16_t* c, int n) {
        const uint32_t r = rng;
        const uint32_t v16 = (uint32_t)(dif >> 48);
        const __m128i c0 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(c)), c1 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(c + 8));
        const __m128i i0 = _mm_set_epi16(7, 6, 5, 4, 3, 2, 1, 0), i1 = _mm_set_epi16(15, 14, 13, 12, 11, 10, 9, 8);
        const __m128i nm1 = _mm_set1_epi16((short)(n - 1));
        const __m128i en0 = _mm_cmpgt_epi16(nm1, i0), en1 = _mm_cmpgt_epi16(nm1, i1);                     // lanes < n - 1
        const __m128i mp0 = _mm_slli_epi16(_mm_sub_epi16(nm1, i0), 2), mp1 = _mm_slli_epi16(_mm_sub_epi16(nm1, i1), 2);   // 4 * (n - 1 - i)
        const __m128i rr = _mm_set1_epi16((short)(r & 0xff00));
        __m128i v0 = _mm_mulhi_epu16(rr, _mm_slli_epi16(_mm_srli_epi16(c0, 6), 7));
        __m128i v1 = _mm_mulhi_epu16(rr, _mm_slli_epi16(_mm_srli_epi16(c1, 6), 7));
        v0 = _mm_and_si128(_mm_add_epi16(v0, mp0), en0);
        v1 = _mm_and_si128(_mm_add_epi16(v1, mp1), en1);
        const __m128i w = _mm_set1_epi16((short)v16), z = _mm_setzero_si128();
        const __m128i le0 = _mm_cmpeq_epi16(_mm_subs_epu16(v0, w), z), le1 = _mm_cmpeq_epi16(_mm_subs_epu16(v1, w), z);   // v <= window
        const unsigned mask = (unsigned)_mm_movemask_epi8(le0) | ((unsigned)_mm_movemask_epi8(le1) << 16);
        const int s = __builtin_ctz(mask | 0x80000000u) >> 1;     // lane n - 1 holds 0: there always is one before lane 16
        // split points below / at the symbol: lane s - 1 (the range itself for s == 0) and lane s
        alignas(16) uint16_t t[24];
        t[7] = (uint16_t)r;
        _mm_store_si128(reinterpret_cast<__m128i*>(t + 8), v0);
        _mm_storeu_si128(reinterpret_cast<__m128i*>(t + 16), v1);
        const uint32_t u = t[7 + s], vv = s < 16 ? t[8 + s] : 0;
        normalize(dif - ((uint64_t)vv << 48), u - vv);
        if (update) {
            const int cnt_ = c[n];
            const __m128i rate = _mm_cvtsi32_si128(5 + (cnt_ > 15) + (cnt_ > 31));
            const __m128i h = _mm_set1_epi16((short)0x8000);
            const __m128i up0 = _mm_add_epi16(c0, _mm_srl_epi16(_mm_sub_epi16(h, c0), rate)), dn0 = _mm_sub_epi16(c0, _mm_srl_epi16(c0, rate));
            const __m128i up1 = _mm_add_epi16(c1, _mm_srl_epi16(_mm_sub_epi16(h, c1), rate)), dn1 = _mm_sub_epi16(c1, _mm_srl_epi16(c1, rate));
            const __m128i lt0 = _mm_andnot_si128(le0, en0), lt1 = _mm_andnot_si128(le1, en1);           // lanes i < s
            __m128i n0 = _mm_or_si128(_mm_and_si128(lt0, up0), _mm_andnot_si128(lt0, dn0));
            __m128i n1 = _mm_or_si128(_mm_and_si128(lt1, up1), _mm_andnot_si128(lt1, dn1));
            n0 = _mm_or_si128(_mm_and_si128(en0, n0), _mm_andnot_si128(en0, c0));                        // other lanes keep their value
            n1 = _mm_or_si128(_mm_and_si128(en1, n1), _mm_andnot_si128(en1, c1));
            _mm_storeu_si128(reinterpret_cast<__m128i*>(c), n0);
            _mm_storeu_si128(reinterpret_cast<__m128i*>(c + 8), n1);
            c[n] = (uint16_t)(cnt_ + (cnt_ < 32));
        }
        return s;
    }
#endif
    __attribute__((always_inline)) static inline void adapt(uint16_t* c, int val, int n) {
        const int cnt_ = c[n];
        const int rate = 3 + (cnt_ > 15) + (cnt_ > 31) + (n > 3 ? 2 : 1);   // + Min(FloorLog2(n), 2)
        if (n <= 4) {
#pragma GCC unroll 3
            for (int i = 0; i < n - 1; i++) {
                const int up = c[i] + ((32768 - c[i]) >> rate), dn = c[i] - (c[i] >> rate);
                c[i] = (uint16_t)(i < val ? up : dn);
            }
        } else {
            for (int i = 0; i < n - 1; i++) {
                if (i < val) c[i] += (32768 - c[i]) >> rate;
                else c[i] -= c[i] >> rate;
            }
        }
        c[n] = (uint16_t)(cnt_ + (cnt_ < 32));
    }
    __attribute__((always_inline)) inline int bit() {   // read_bool / one bit of a literal: equiprobable, no adaptation
        const uint32_t r = rng;
        const uint32_t v = ((r >> 8) << 7) + 4;
        const uint64_t vw = (uint64_t)v << 48;
        int ret;
        uint64_t d = dif;
        uint32_t rn;
        if (d >= vw) { rn = r - v; d -= vw; ret = 0; } else { rn = v; ret = 1; }
        normalize(d, rn);
        return ret;
    }
    inline int literal(int n) {
        int x = 0;
        for (int i = 0; i < n; i++) x = 2 * x + bit();
        return x;
    }
    // non-adaptive symbol with an explicit 2-entry inverted cdf (split_or_horz / split_or_vert)
    inline int bool_icdf(uint32_t icdf0) {
        const uint32_t r = rng;
        const uint32_t v = ((r >> 8) * (icdf0 >> 6) >> 1) + 4;
        const uint64_t vw = (uint64_t)v << 48;
        uint64_t d = dif;
        uint32_t rn;
        int ret;
        if (d >= vw) { rn = r - v; d -= vw; ret = 0; } else { rn = v; ret = 1; }
        normalize(d, rn);
        return ret;
    }
};

}  // namespace av1r
