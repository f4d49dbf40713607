// 1-D inverse transforms of AV1 (spec 7.13.2): DCT4..64, ADST4/8/16, identity 4/8/16/32, WHT4.
// Integer butterflies with the normative Round2(a*cos128 +/- b*sin128, 12) rounding after every
// rotation.  Host+device: the sm_100a K1 kernel runs one of these per thread (one row, then one
// column, of a transform block) out of registers; the CPU oracle runs the same arithmetic in
// scalar loops.  Unit-pinned against libaom's av1_idct*/av1_iadst* (tests/test_itx.py).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define AV1R_HD __host__ __device__ __forceinline__
#else
#define AV1R_HD inline
#endif

namespace av1r {

// cos(i*pi/128) * 4096, i = 0..64
#define AV1R_COSPI_LIST                                                                                                 \
    4096, 4095, 4091, 4085, 4076, 4065, 4052, 4036, 4017, 3996, 3973, 3948, 3920, 3889, 3857, 3822, 3784, 3745, 3703,    \
        3659, 3612, 3564, 3513, 3461, 3406, 3349, 3290, 3229, 3166, 3102, 3035, 2967, 2896, 2824, 2751, 2675, 2598, 2520, \
        2440, 2359, 2276, 2191, 2106, 2019, 1931, 1842, 1751, 1660, 1567, 1474, 1380, 1285, 1189, 1092, 995, 897, 799,   \
        700, 601, 501, 401, 301, 201, 101, 0

template <int I> struct Cospi;
#define AV1R_CP(i) (Cospi<i>::v)
namespace cospi_detail {
constexpr int kTab[65] = {AV1R_COSPI_LIST};
}
template <int I> struct Cospi { static constexpr int v = cospi_detail::kTab[I]; };

AV1R_HD int32_t hbtf(int w0, int32_t a, int w1, int32_t b) {
    // Round2(w0*a + w1*b, 12); wrap-around arithmetic like the 32-bit reference decoders
    uint32_t s = (uint32_t)w0 * (uint32_t)a + (uint32_t)w1 * (uint32_t)b + 2048u;
    return (int32_t)s >> 12;
}

// (a, b) <- (a*c0 - b*c1, a*c1 + b*c0)  -- the "rotation" used in the first stage of each odd half
#define AV1R_ROT(a, b, c0, c1)                              \
    {                                                       \
        int32_t _x = hbtf(AV1R_CP(c0), a, -AV1R_CP(c1), b); \
        int32_t _y = hbtf(AV1R_CP(c1), a, AV1R_CP(c0), b);  \
        a = _x;                                             \
        b = _y;                                             \
    }
// (a, b) <- (a + b, a - b)
#define AV1R_BF(a, b)         \
    {                         \
        int32_t _s = a + b;   \
        int32_t _d = a - b;   \
        a = _s;               \
        b = _d;               \
    }
// (a, b) <- (b - a, b + a)
#define AV1R_BFR(a, b)        \
    {                         \
        int32_t _s = b - a;   \
        int32_t _d = b + a;   \
        a = _s;               \
        b = _d;               \
    }

// ---- DCT.  t[] holds the input in natural order; output in natural order, in place.
AV1R_HD void idct4_core(int32_t& x0, int32_t& x1, int32_t& x2, int32_t& x3) {
    // inputs: x0,x1,x2,x3 natural order
    int32_t s0 = hbtf(AV1R_CP(32), x0, AV1R_CP(32), x2);
    int32_t s1 = hbtf(AV1R_CP(32), x0, -AV1R_CP(32), x2);
    int32_t s2 = hbtf(AV1R_CP(48), x1, -AV1R_CP(16), x3);
    int32_t s3 = hbtf(AV1R_CP(16), x1, AV1R_CP(48), x3);
    x0 = s0 + s3;
    x1 = s1 + s2;
    x2 = s1 - s2;
    x3 = s0 - s3;
}

AV1R_HD void idct4(int32_t* t) { idct4_core(t[0], t[1], t[2], t[3]); }

AV1R_HD void idct8(int32_t* t) {
    // even half
    int32_t e0 = t[0], e1 = t[2], e2 = t[4], e3 = t[6];
    idct4_core(e0, e1, e2, e3);
    // odd half: a4=x1 a5=x5 a6=x3 a7=x7
    int32_t a4 = t[1], a5 = t[5], a6 = t[3], a7 = t[7];
    AV1R_ROT(a4, a7, 56, 8);
    AV1R_ROT(a5, a6, 24, 40);
    AV1R_BF(a4, a5);    // c4 = b4+b5, c5 = b4-b5
    AV1R_BFR(a6, a7);   // c6 = b7-b6, c7 = b7+b6
    int32_t d5 = hbtf(-AV1R_CP(32), a5, AV1R_CP(32), a6);
    int32_t d6 = hbtf(AV1R_CP(32), a5, AV1R_CP(32), a6);
    t[0] = e0 + a7; t[7] = e0 - a7;
    t[1] = e1 + d6; t[6] = e1 - d6;
    t[2] = e2 + d5; t[5] = e2 - d5;
    t[3] = e3 + a4; t[4] = e3 - a4;
}

// odd half of the 16-point DCT on u[0..7] = (x1, x9, x5, x13, x3, x11, x7, x15); result = f8..f15
AV1R_HD void idct16_odd(int32_t* u) {
    AV1R_ROT(u[0], u[7], 60, 4);
    AV1R_ROT(u[1], u[6], 28, 36);
    AV1R_ROT(u[2], u[5], 44, 20);
    AV1R_ROT(u[3], u[4], 12, 52);
    AV1R_BF(u[0], u[1]);
    AV1R_BFR(u[2], u[3]);
    AV1R_BF(u[4], u[5]);
    AV1R_BFR(u[6], u[7]);
    {   // d9 = btf(-16, c9, 48, c14); d14 = btf(48, c9, 16, c14); d10 = btf(-48, c10, -16, c13); d13 = btf(-16, c10, 48, c13)
        int32_t d9 = hbtf(-AV1R_CP(16), u[1], AV1R_CP(48), u[6]);
        int32_t d14 = hbtf(AV1R_CP(48), u[1], AV1R_CP(16), u[6]);
        int32_t d10 = hbtf(-AV1R_CP(48), u[2], -AV1R_CP(16), u[5]);
        int32_t d13 = hbtf(-AV1R_CP(16), u[2], AV1R_CP(48), u[5]);
        u[1] = d9; u[6] = d14; u[2] = d10; u[5] = d13;
    }
    {   // e8 = d8+d11; e9 = d9+d10; e10 = d9-d10; e11 = d8-d11; e12 = d15-d12; e13 = d14-d13; e14 = d13+d14; e15 = d12+d15
        int32_t e8 = u[0] + u[3], e11 = u[0] - u[3], e9 = u[1] + u[2], e10 = u[1] - u[2];
        int32_t e12 = u[7] - u[4], e15 = u[4] + u[7], e13 = u[6] - u[5], e14 = u[5] + u[6];
        u[0] = e8; u[1] = e9; u[6] = e14; u[7] = e15;
        u[2] = hbtf(-AV1R_CP(32), e10, AV1R_CP(32), e13);
        u[5] = hbtf(AV1R_CP(32), e10, AV1R_CP(32), e13);
        u[3] = hbtf(-AV1R_CP(32), e11, AV1R_CP(32), e12);
        u[4] = hbtf(AV1R_CP(32), e11, AV1R_CP(32), e12);
    }
}

AV1R_HD void idct16(int32_t* t) {
    int32_t e[8] = {t[0], t[2], t[4], t[6], t[8], t[10], t[12], t[14]};
    idct8(e);
    int32_t u[8] = {t[1], t[9], t[5], t[13], t[3], t[11], t[7], t[15]};
    idct16_odd(u);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        t[i] = e[i] + u[7 - i];
        t[15 - i] = e[i] - u[7 - i];
    }
}

// odd half of the 32-point DCT on u[0..15] = x(1,17,9,25,5,21,13,29,3,19,11,27,7,23,15,31); result h16..h31
AV1R_HD void idct32_odd(int32_t* u) {
    AV1R_ROT(u[0], u[15], 62, 2);
    AV1R_ROT(u[1], u[14], 30, 34);
    AV1R_ROT(u[2], u[13], 46, 18);
    AV1R_ROT(u[3], u[12], 14, 50);
    AV1R_ROT(u[4], u[11], 54, 10);
    AV1R_ROT(u[5], u[10], 22, 42);
    AV1R_ROT(u[6], u[9], 38, 26);
    AV1R_ROT(u[7], u[8], 6, 58);
    AV1R_BF(u[0], u[1]);  AV1R_BFR(u[2], u[3]);  AV1R_BF(u[4], u[5]);   AV1R_BFR(u[6], u[7]);
    AV1R_BF(u[8], u[9]);  AV1R_BFR(u[10], u[11]); AV1R_BF(u[12], u[13]); AV1R_BFR(u[14], u[15]);
    {   // stage 4 (indices +16): 17/30 (8,56), 18/29 (56,8 neg), 21/26 (40,24), 22/25 (24,40 neg)
        int32_t d17 = hbtf(-AV1R_CP(8), u[1], AV1R_CP(56), u[14]);
        int32_t d30 = hbtf(AV1R_CP(56), u[1], AV1R_CP(8), u[14]);
        int32_t d18 = hbtf(-AV1R_CP(56), u[2], -AV1R_CP(8), u[13]);
        int32_t d29 = hbtf(-AV1R_CP(8), u[2], AV1R_CP(56), u[13]);
        int32_t d21 = hbtf(-AV1R_CP(40), u[5], AV1R_CP(24), u[10]);
        int32_t d26 = hbtf(AV1R_CP(24), u[5], AV1R_CP(40), u[10]);
        int32_t d22 = hbtf(-AV1R_CP(24), u[6], -AV1R_CP(40), u[9]);
        int32_t d25 = hbtf(-AV1R_CP(40), u[6], AV1R_CP(24), u[9]);
        u[1] = d17; u[14] = d30; u[2] = d18; u[13] = d29; u[5] = d21; u[10] = d26; u[6] = d22; u[9] = d25;
    }
    {   // stage 5
        int32_t e16 = u[0] + u[3], e19 = u[0] - u[3], e17 = u[1] + u[2], e18 = u[1] - u[2];
        int32_t e20 = u[7] - u[4], e23 = u[4] + u[7], e21 = u[6] - u[5], e22 = u[5] + u[6];
        int32_t e24 = u[8] + u[11], e27 = u[8] - u[11], e25 = u[9] + u[10], e26 = u[9] - u[10];
        int32_t e28 = u[15] - u[12], e31 = u[12] + u[15], e29 = u[14] - u[13], e30 = u[13] + u[14];
        // stage 6
        int32_t f18 = hbtf(-AV1R_CP(16), e18, AV1R_CP(48), e29);
        int32_t f29 = hbtf(AV1R_CP(48), e18, AV1R_CP(16), e29);
        int32_t f19 = hbtf(-AV1R_CP(16), e19, AV1R_CP(48), e28);
        int32_t f28 = hbtf(AV1R_CP(48), e19, AV1R_CP(16), e28);
        int32_t f20 = hbtf(-AV1R_CP(48), e20, -AV1R_CP(16), e27);
        int32_t f27 = hbtf(-AV1R_CP(16), e20, AV1R_CP(48), e27);
        int32_t f21 = hbtf(-AV1R_CP(48), e21, -AV1R_CP(16), e26);
        int32_t f26 = hbtf(-AV1R_CP(16), e21, AV1R_CP(48), e26);
        // stage 7
        int32_t g16 = e16 + e23, g23 = e16 - e23, g17 = e17 + e22, g22 = e17 - e22;
        int32_t g18 = f18 + f21, g21 = f18 - f21, g19 = f19 + f20, g20 = f19 - f20;
        int32_t g24 = e31 - e24, g31 = e24 + e31, g25 = e30 - e25, g30 = e25 + e30;
        int32_t g26 = f29 - f26, g29 = f26 + f29, g27 = f28 - f27, g28 = f27 + f28;
        // stage 8
        u[0] = g16; u[1] = g17; u[2] = g18; u[3] = g19;
        u[4] = hbtf(-AV1R_CP(32), g20, AV1R_CP(32), g27);
        u[5] = hbtf(-AV1R_CP(32), g21, AV1R_CP(32), g26);
        u[6] = hbtf(-AV1R_CP(32), g22, AV1R_CP(32), g25);
        u[7] = hbtf(-AV1R_CP(32), g23, AV1R_CP(32), g24);
        u[8] = hbtf(AV1R_CP(32), g23, AV1R_CP(32), g24);
        u[9] = hbtf(AV1R_CP(32), g22, AV1R_CP(32), g25);
        u[10] = hbtf(AV1R_CP(32), g21, AV1R_CP(32), g26);
        u[11] = hbtf(AV1R_CP(32), g20, AV1R_CP(32), g27);
        u[12] = g28; u[13] = g29; u[14] = g30; u[15] = g31;
    }
}

AV1R_HD void idct32(int32_t* t) {
    int32_t e[16];
#pragma unroll
    for (int i = 0; i < 16; i++) e[i] = t[2 * i];
    idct16(e);
    int32_t u[16] = {t[1], t[17], t[9], t[25], t[5], t[21], t[13], t[29], t[3], t[19], t[11], t[27], t[7], t[23], t[15], t[31]};
    idct32_odd(u);
#pragma unroll
    for (int i = 0; i < 16; i++) {
        t[i] = e[i] + u[15 - i];
        t[31 - i] = e[i] - u[15 - i];
    }
}

// odd half of the 64-point DCT on u[0..31] (bit-reversed odd inputs); result = outputs 32..63 of the last stage
AV1R_HD void idct64_odd(int32_t* u) {
    // stage 2: rotations with angle 63 - 4*brev4(i)
    AV1R_ROT(u[0], u[31], 63, 1);
    AV1R_ROT(u[1], u[30], 31, 33);
    AV1R_ROT(u[2], u[29], 47, 17);
    AV1R_ROT(u[3], u[28], 15, 49);
    AV1R_ROT(u[4], u[27], 55, 9);
    AV1R_ROT(u[5], u[26], 23, 41);
    AV1R_ROT(u[6], u[25], 39, 25);
    AV1R_ROT(u[7], u[24], 7, 57);
    AV1R_ROT(u[8], u[23], 59, 5);
    AV1R_ROT(u[9], u[22], 27, 37);
    AV1R_ROT(u[10], u[21], 43, 21);
    AV1R_ROT(u[11], u[20], 11, 53);
    AV1R_ROT(u[12], u[19], 51, 13);
    AV1R_ROT(u[13], u[18], 19, 45);
    AV1R_ROT(u[14], u[17], 35, 29);
    AV1R_ROT(u[15], u[16], 3, 61);
    // stage 3
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        AV1R_BF(u[2 * k], u[2 * k + 1]);
        AV1R_BFR(u[2 * k + 2], u[2 * k + 3]);
    }
    // stage 4: pairs (1,30)(2,29) angle(4,60); (5,26)(6,25) (36,28); (9,22)(10,21) (20,44); (13,18)(14,17) (52,12)
#define AV1R_S4(i, j, ca, cb)                                        \
    {                                                                \
        int32_t _p = hbtf(-AV1R_CP(ca), u[i], AV1R_CP(cb), u[j]);    \
        int32_t _q = hbtf(AV1R_CP(cb), u[i], AV1R_CP(ca), u[j]);     \
        u[i] = _p;                                                   \
        u[j] = _q;                                                   \
    }
#define AV1R_S4N(i, j, ca, cb)                                       \
    {                                                                \
        int32_t _p = hbtf(-AV1R_CP(cb), u[i], -AV1R_CP(ca), u[j]);   \
        int32_t _q = hbtf(-AV1R_CP(ca), u[i], AV1R_CP(cb), u[j]);    \
        u[i] = _p;                                                   \
        u[j] = _q;                                                   \
    }
    AV1R_S4(1, 30, 4, 60);   AV1R_S4N(2, 29, 4, 60);
    AV1R_S4(5, 26, 36, 28);  AV1R_S4N(6, 25, 36, 28);
    AV1R_S4(9, 22, 20, 44);  AV1R_S4N(10, 21, 20, 44);
    AV1R_S4(13, 18, 52, 12); AV1R_S4N(14, 17, 52, 12);
    // stage 5: groups of 4: (a,b,c,d) -> (a+d, b+c, b-c, a-d) alternating with the mirrored form
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        int32_t a = u[k], b = u[k + 1], c = u[k + 2], d = u[k + 3];
        u[k] = a + d; u[k + 1] = b + c; u[k + 2] = b - c; u[k + 3] = a - d;
        a = u[k + 4]; b = u[k + 5]; c = u[k + 6]; d = u[k + 7];
        u[k + 4] = d - a; u[k + 5] = c - b; u[k + 6] = c + b; u[k + 7] = d + a;
    }
    // stage 6: (2,29)(3,28) angle (8,56); (4,27)(5,26) neg; (10,21)(11,20) angle (40,24); (12,19)(13,18) neg
    AV1R_S4(2, 29, 8, 56);   AV1R_S4(3, 28, 8, 56);
    AV1R_S4N(4, 27, 8, 56);  AV1R_S4N(5, 26, 8, 56);
    AV1R_S4(10, 21, 40, 24); AV1R_S4(11, 20, 40, 24);
    AV1R_S4N(12, 19, 40, 24); AV1R_S4N(13, 18, 40, 24);
    // stage 7: groups of 8
#pragma unroll
    for (int k = 0; k < 32; k += 16) {
        int32_t a[8];
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = u[k + i];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            u[k + i] = a[i] + a[7 - i];
            u[k + 7 - i] = a[i] - a[7 - i];
        }
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = u[k + 8 + i];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            u[k + 8 + i] = a[7 - i] - a[i];
            u[k + 15 - i] = a[7 - i] + a[i];
        }
    }
    // stage 8: (4..7 with 27..24) angle (16,48); (8..11 with 23..20) neg
    AV1R_S4(4, 27, 16, 48);  AV1R_S4(5, 26, 16, 48);  AV1R_S4(6, 25, 16, 48);  AV1R_S4(7, 24, 16, 48);
    AV1R_S4N(8, 23, 16, 48); AV1R_S4N(9, 22, 16, 48); AV1R_S4N(10, 21, 16, 48); AV1R_S4N(11, 20, 16, 48);
    // stage 9: groups of 16
    {
        int32_t a[16];
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = u[i];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            u[i] = a[i] + a[15 - i];
            u[15 - i] = a[i] - a[15 - i];
        }
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = u[16 + i];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            u[16 + i] = a[15 - i] - a[i];
            u[31 - i] = a[15 - i] + a[i];
        }
    }
    // stage 10: (8..15 with 23..16) by 32
#pragma unroll
    for (int i = 8; i < 16; i++) {
        int32_t p = hbtf(-AV1R_CP(32), u[i], AV1R_CP(32), u[31 - i]);
        int32_t q = hbtf(AV1R_CP(32), u[i], AV1R_CP(32), u[31 - i]);
        u[i] = p;
        u[31 - i] = q;
    }
#undef AV1R_S4
#undef AV1R_S4N
}

AV1R_HD void idct64(int32_t* t) {
    int32_t e[32];
#pragma unroll
    for (int i = 0; i < 32; i++) e[i] = t[2 * i];
    idct32(e);
    int32_t u[32] = {t[1],  t[33], t[17], t[49], t[9],  t[41], t[25], t[57], t[5],  t[37], t[21], t[53], t[13], t[45], t[29], t[61],
                     t[3],  t[35], t[19], t[51], t[11], t[43], t[27], t[59], t[7],  t[39], t[23], t[55], t[15], t[47], t[31], t[63]};
    idct64_odd(u);
#pragma unroll
    for (int i = 0; i < 32; i++) {
        t[i] = e[i] + u[31 - i];
        t[63 - i] = e[i] - u[31 - i];
    }
}

// ---- ADST
AV1R_HD void iadst4(int32_t* t) {
    const int32_t x0 = t[0], x1 = t[1], x2 = t[2], x3 = t[3];
    int32_t s0 = 1321 * x0, s1 = 2482 * x0, s2 = 3344 * x1, s3 = 3803 * x2, s4 = 1321 * x2, s5 = 2482 * x3, s6 = 3803 * x3;
    int32_t s7 = (x0 - x2) + x3;
    s0 = s0 + s3;
    s1 = s1 - s4;
    s3 = s2;
    s2 = 3344 * s7;
    s0 = s0 + s5;
    s1 = s1 - s6;
    int32_t y0 = s0 + s3, y1 = s1 + s3, y2 = s2, y3 = s0 + s1;
    y3 = y3 - s3;
    t[0] = (y0 + 2048) >> 12;
    t[1] = (y1 + 2048) >> 12;
    t[2] = (y2 + 2048) >> 12;
    t[3] = (y3 + 2048) >> 12;
}

// (a, b) <- (a*c0 + b*c1, a*c1 - b*c0)
#define AV1R_AROT(a, b, c0, c1)                             \
    {                                                       \
        int32_t _x = hbtf(AV1R_CP(c0), a, AV1R_CP(c1), b);  \
        int32_t _y = hbtf(AV1R_CP(c1), a, -AV1R_CP(c0), b); \
        a = _x;                                             \
        b = _y;                                             \
    }
// (a, b) <- (-a*c1 + b*c0, a*c0 + b*c1)   [btf(-c1,a,c0,b), btf(c0,a,c1,b)]
#define AV1R_AROTN(a, b, c0, c1)                            \
    {                                                       \
        int32_t _x = hbtf(-AV1R_CP(c1), a, AV1R_CP(c0), b); \
        int32_t _y = hbtf(AV1R_CP(c0), a, AV1R_CP(c1), b);  \
        a = _x;                                             \
        b = _y;                                             \
    }

AV1R_HD void iadst8(int32_t* t) {
    int32_t b0 = t[7], b1 = t[0], b2 = t[5], b3 = t[2], b4 = t[3], b5 = t[4], b6 = t[1], b7 = t[6];
    AV1R_AROT(b0, b1, 4, 60);
    AV1R_AROT(b2, b3, 20, 44);
    AV1R_AROT(b4, b5, 36, 28);
    AV1R_AROT(b6, b7, 52, 12);
    AV1R_BF(b0, b4); AV1R_BF(b1, b5); AV1R_BF(b2, b6); AV1R_BF(b3, b7);
    AV1R_AROT(b4, b5, 16, 48);
    AV1R_AROTN(b6, b7, 16, 48);
    AV1R_BF(b0, b2); AV1R_BF(b1, b3); AV1R_BF(b4, b6); AV1R_BF(b5, b7);
    AV1R_AROT(b2, b3, 32, 32);
    AV1R_AROT(b6, b7, 32, 32);
    t[0] = b0; t[1] = -b4; t[2] = b6; t[3] = -b2; t[4] = b3; t[5] = -b7; t[6] = b5; t[7] = -b1;
}

AV1R_HD void iadst16(int32_t* t) {
    int32_t b[16] = {t[15], t[0], t[13], t[2], t[11], t[4], t[9], t[6], t[7], t[8], t[5], t[10], t[3], t[12], t[1], t[14]};
    AV1R_AROT(b[0], b[1], 2, 62);
    AV1R_AROT(b[2], b[3], 10, 54);
    AV1R_AROT(b[4], b[5], 18, 46);
    AV1R_AROT(b[6], b[7], 26, 38);
    AV1R_AROT(b[8], b[9], 34, 30);
    AV1R_AROT(b[10], b[11], 42, 22);
    AV1R_AROT(b[12], b[13], 50, 14);
    AV1R_AROT(b[14], b[15], 58, 6);
#pragma unroll
    for (int i = 0; i < 8; i++) AV1R_BF(b[i], b[i + 8]);
    AV1R_AROT(b[8], b[9], 8, 56);
    AV1R_AROT(b[10], b[11], 40, 24);
    AV1R_AROTN(b[12], b[13], 8, 56);
    AV1R_AROTN(b[14], b[15], 40, 24);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        AV1R_BF(b[i], b[i + 4]);
        AV1R_BF(b[i + 8], b[i + 12]);
    }
    AV1R_AROT(b[4], b[5], 16, 48);
    AV1R_AROTN(b[6], b[7], 16, 48);
    AV1R_AROT(b[12], b[13], 16, 48);
    AV1R_AROTN(b[14], b[15], 16, 48);
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
        AV1R_BF(b[i], b[i + 2]);
        AV1R_BF(b[i + 1], b[i + 3]);
    }
    AV1R_AROT(b[2], b[3], 32, 32);
    AV1R_AROT(b[6], b[7], 32, 32);
    AV1R_AROT(b[10], b[11], 32, 32);
    AV1R_AROT(b[14], b[15], 32, 32);
    t[0] = b[0];   t[1] = -b[8];  t[2] = b[12];  t[3] = -b[4];  t[4] = b[6];   t[5] = -b[14]; t[6] = b[10];  t[7] = -b[2];
    t[8] = b[3];   t[9] = -b[11]; t[10] = b[15]; t[11] = -b[7]; t[12] = b[5];  t[13] = -b[13]; t[14] = b[9]; t[15] = -b[1];
}

// ---- identity
AV1R_HD void iidentity4(int32_t* t) {
#pragma unroll
    for (int i = 0; i < 4; i++) t[i] = (int32_t)(((int64_t)t[i] * 5793 + 2048) >> 12);
}
AV1R_HD void iidentity8(int32_t* t) {
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = t[i] * 2;
}
AV1R_HD void iidentity16(int32_t* t) {
#pragma unroll
    for (int i = 0; i < 16; i++) t[i] = (int32_t)(((int64_t)t[i] * 11586 + 2048) >> 12);
}
AV1R_HD void iidentity32(int32_t* t) {
#pragma unroll
    for (int i = 0; i < 32; i++) t[i] = t[i] * 4;
}

// ---- Walsh-Hadamard (lossless), shift = 2 for the row pass, 0 for the column pass
AV1R_HD void iwht4(int32_t* t, int shift) {
    int32_t a = t[0] >> shift, c = t[1] >> shift, d = t[2] >> shift, b = t[3] >> shift;
    a += c;
    d -= b;
    int32_t e = (a - d) >> 1;
    b = e - b;
    c = e - c;
    a -= b;
    d += c;
    t[0] = a; t[1] = b; t[2] = c; t[3] = d;
}

enum { ITX_DCT = 0, ITX_ADST = 1, ITX_FLIPADST = 2, ITX_IDENTITY = 3 };

// 1-D transform of length n (4..64) and kind k on t[0..n-1], in place.
AV1R_HD void itx_1d(int32_t* t, int n, int kind) {
    if (kind == ITX_DCT) {
        switch (n) {
            case 4: idct4(t); break;
            case 8: idct8(t); break;
            case 16: idct16(t); break;
            case 32: idct32(t); break;
            default: idct64(t); break;
        }
    } else if (kind == ITX_IDENTITY) {
        switch (n) {
            case 4: iidentity4(t); break;
            case 8: iidentity8(t); break;
            case 16: iidentity16(t); break;
            default: iidentity32(t); break;
        }
    } else {
        switch (n) {
            case 4: iadst4(t); break;
            case 8: iadst8(t); break;
            default: iadst16(t); break;
        }
    }
}

// vertical (column) and horizontal (row) 1-D kinds + flips of a 2-D TxType
AV1R_HD void txtp_decompose(int txtp, int& vkind, int& hkind, int& ud_flip, int& lr_flip) {
    // order: DCT_DCT, ADST_DCT, DCT_ADST, ADST_ADST, FLIPADST_DCT, DCT_FLIPADST, FLIPADST_FLIPADST, ADST_FLIPADST,
    //        FLIPADST_ADST, IDTX, V_DCT, H_DCT, V_ADST, H_ADST, V_FLIPADST, H_FLIPADST   (first = vertical)
    const int v[16] = {0, 1, 0, 1, 2, 0, 2, 1, 2, 3, 0, 3, 1, 3, 2, 3};
    const int h[16] = {0, 0, 1, 1, 0, 2, 2, 2, 1, 3, 3, 0, 3, 1, 3, 2};
    vkind = v[txtp & 15];
    hkind = h[txtp & 15];
    ud_flip = vkind == ITX_FLIPADST;
    lr_flip = hkind == ITX_FLIPADST;
}

}  // namespace av1r
