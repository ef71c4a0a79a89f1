// float_device.cuh — encoding/json's float64 text on the device.
//
// Replaces, for struct fields of kind GOFR_F_FLOAT64, what Responder.Respond reaches through json.Encoder.Encode
// (pkg/gofr/http/responder.go:40; Go 1.21 encoding/json floatEncoder, go.mod:3):
//     b = strconv.AppendFloat(b, f, fmt, -1, 64)   with fmt = 'e' if abs != 0 && (abs < 1e-6 || abs >= 1e21) else 'f',
//     then "e-0X" is cleaned to "e-X"; NaN and ±Inf are an UnsupportedValueError (Encode fails: the body stays empty).
// Precision -1 is the SHORTEST decimal that parses back to the same float64.  The digits come from the published Ryu
// algorithm (Ulf Adams, "Ryu: fast float-to-string conversion", PLDI 2018) restated here; its two tables of 128-bit powers
// of five are generated with exact integer arithmetic by scratch/gen/ryu_tables.py.  The oracle does NOT share this code: it
// finds the shortest digits by trying precisions with the C library's correctly rounded printf / strtod (oracle/orc_float.c),
// and the tests check both against Python's repr() (also shortest round-trip) on millions of values.
#pragma once
#include <stdint.h>

// included by value_device.cuh, inside serve_device.cuh (GOFR_HD comes from there)

namespace gofr {

#define RYU_TABLE(name, n) static const uint64_t name##_host[n][2]
#include "ryu_tables.inc"
#undef RYU_TABLE
#if defined(__CUDACC__)
#define RYU_TABLE(name, n) static __device__ const uint64_t name##_dev[n][2]
#include "ryu_tables.inc"
#undef RYU_TABLE
#endif

GOFR_HD const uint64_t* ryu_pow5_inv(uint32_t i) {
#if defined(__CUDA_ARCH__)
    return RYU_POW5_INV_SPLIT_dev[i];
#else
    return RYU_POW5_INV_SPLIT_host[i];
#endif
}
GOFR_HD const uint64_t* ryu_pow5(uint32_t i) {
#if defined(__CUDA_ARCH__)
    return RYU_POW5_SPLIT_dev[i];
#else
    return RYU_POW5_SPLIT_host[i];
#endif
}

GOFR_HD uint64_t umul128(uint64_t a, uint64_t b, uint64_t* hi) {
#if defined(__CUDA_ARCH__)
    *hi = __umul64hi(a, b);
    return a * b;
#else
    const unsigned __int128 p = (unsigned __int128)a * b;
    *hi = (uint64_t)(p >> 64);
    return (uint64_t)p;
#endif
}
// low 64 bits of (hi:lo) >> dist, 0 < dist < 64
GOFR_HD uint64_t shiftright128(uint64_t lo, uint64_t hi, uint32_t dist) { return (hi << (64 - dist)) | (lo >> dist); }

GOFR_HD uint64_t ryu_mul_shift(uint64_t m, const uint64_t* mul, int32_t j) {  // (m * mul) >> j, m < 2^55, j >= 64
    uint64_t high1, high0;
    const uint64_t low1 = umul128(m, mul[1], &high1);
    umul128(m, mul[0], &high0);
    const uint64_t sum = high0 + low1;
    if (sum < high0) ++high1;
    return shiftright128(sum, high1, (uint32_t)(j - 64));
}
GOFR_HD uint32_t ryu_pow5_factor(uint64_t v) {
    uint32_t c = 0;
    for (;;) {
        const uint64_t q = v / 5;
        if ((uint32_t)v - 5 * (uint32_t)q != 0) break;
        v = q;
        c++;
    }
    return c;
}
GOFR_HD bool ryu_multiple_of_pow5(uint64_t v, uint32_t p) { return ryu_pow5_factor(v) >= p; }
GOFR_HD bool ryu_multiple_of_pow2(uint64_t v, uint32_t p) { return (v & ((1ull << p) - 1)) == 0; }
GOFR_HD int32_t ryu_pow5bits(int32_t e) { return (int32_t)(((uint32_t)e * 1217359u) >> 19) + 1; }
GOFR_HD uint32_t ryu_log10_pow2(int32_t e) { return ((uint32_t)e * 78913u) >> 18; }
GOFR_HD uint32_t ryu_log10_pow5(int32_t e) { return ((uint32_t)e * 732923u) >> 20; }

// Shortest decimal of a finite non-zero binary floating-point number (sign stripped): value = digits * 10^exp10.
// MBITS / BIAS: 52 / 1023 for a float64, 23 / 127 for a float32 (strconv.AppendFloat(…, -1, 32): the shortest digits that
// identify the FLOAT32).  The core is the same for both — the interval around 4·m2 scaled by a power of five with the
// 125-bit tables, whose precision covers any m2 < 2^53 and every exponent of either format; the published float32 variant
// differs only in using narrower tables.
template <int MBITS = 52, int BIAS = 1023>
GOFR_HD void ryu_shortest(uint64_t ieee_mantissa, uint32_t ieee_exponent, uint64_t* digits, int32_t* exp10) {
    int32_t e2;
    uint64_t m2;
    if (ieee_exponent == 0) { e2 = 1 - BIAS - MBITS - 2; m2 = ieee_mantissa; }
    else { e2 = (int32_t)ieee_exponent - BIAS - MBITS - 2; m2 = (1ull << MBITS) | ieee_mantissa; }
    const bool accept = (m2 & 1) == 0;  // round-to-even: the interval's bounds are themselves valid
    const uint64_t mv = 4 * m2;
    const uint32_t mm_shift = ieee_mantissa != 0 || ieee_exponent <= 1;
    uint64_t vr, vp, vm;
    int32_t e10;
    bool vm_tz = false, vr_tz = false;
    if (e2 >= 0) {
        const uint32_t q = ryu_log10_pow2(e2) - (e2 > 3);
        e10 = (int32_t)q;
        const int32_t k = 125 + ryu_pow5bits((int32_t)q) - 1;
        const int32_t i = -e2 + (int32_t)q + k;
        const uint64_t* mul = ryu_pow5_inv(q);
        vr = ryu_mul_shift(4 * m2, mul, i);
        vp = ryu_mul_shift(4 * m2 + 2, mul, i);
        vm = ryu_mul_shift(4 * m2 - 1 - mm_shift, mul, i);
        if (q <= 21) {
            const uint32_t mv_mod5 = (uint32_t)mv - 5 * (uint32_t)(mv / 5);
            if (mv_mod5 == 0) vr_tz = ryu_multiple_of_pow5(mv, q);
            else if (accept) vm_tz = ryu_multiple_of_pow5(mv - 1 - mm_shift, q);
            else vp -= ryu_multiple_of_pow5(mv + 2, q);
        }
    } else {
        const uint32_t q = ryu_log10_pow5(-e2) - (-e2 > 1);
        e10 = (int32_t)q + e2;
        const int32_t i = -e2 - (int32_t)q;
        const int32_t k = ryu_pow5bits(i) - 125;
        const int32_t j = (int32_t)q - k;
        const uint64_t* mul = ryu_pow5((uint32_t)i);
        vr = ryu_mul_shift(4 * m2, mul, j);
        vp = ryu_mul_shift(4 * m2 + 2, mul, j);
        vm = ryu_mul_shift(4 * m2 - 1 - mm_shift, mul, j);
        if (q <= 1) {
            vr_tz = true;
            if (accept) vm_tz = mm_shift == 1;
            else --vp;
        } else if (q < 63) {
            vr_tz = ryu_multiple_of_pow2(mv, q);
        }
    }
    int32_t removed = 0;
    uint32_t last = 0;
    uint64_t out;
    if (vm_tz || vr_tz) {
        for (;;) {
            const uint64_t vp10 = vp / 10, vm10 = vm / 10;
            if (vp10 <= vm10) break;
            const uint32_t vm_mod = (uint32_t)vm - 10 * (uint32_t)vm10;
            const uint64_t vr10 = vr / 10;
            const uint32_t vr_mod = (uint32_t)vr - 10 * (uint32_t)vr10;
            vm_tz &= vm_mod == 0;
            vr_tz &= last == 0;
            last = vr_mod;
            vr = vr10; vp = vp10; vm = vm10;
            ++removed;
        }
        if (vm_tz) {
            for (;;) {
                const uint64_t vm10 = vm / 10;
                const uint32_t vm_mod = (uint32_t)vm - 10 * (uint32_t)vm10;
                if (vm_mod != 0) break;
                const uint64_t vp10 = vp / 10, vr10 = vr / 10;
                const uint32_t vr_mod = (uint32_t)vr - 10 * (uint32_t)vr10;
                vr_tz &= last == 0;
                last = vr_mod;
                vr = vr10; vp = vp10; vm = vm10;
                ++removed;
            }
        }
        if (vr_tz && last == 5 && vr % 2 == 0) last = 4;  // exactly ...50..0: round to even
        out = vr + ((vr == vm && (!accept || !vm_tz)) || last >= 5);
    } else {
        bool round_up = false;
        const uint64_t vp100 = vp / 100, vm100 = vm / 100;
        if (vp100 > vm100) {
            const uint64_t vr100 = vr / 100;
            const uint32_t vr_mod = (uint32_t)vr - 100 * (uint32_t)vr100;
            round_up = vr_mod >= 50;
            vr = vr100; vp = vp100; vm = vm100;
            removed += 2;
        }
        for (;;) {
            const uint64_t vp10 = vp / 10, vm10 = vm / 10;
            if (vp10 <= vm10) break;
            const uint64_t vr10 = vr / 10;
            const uint32_t vr_mod = (uint32_t)vr - 10 * (uint32_t)vr10;
            round_up = vr_mod >= 5;
            vr = vr10; vp = vp10; vm = vm10;
            ++removed;
        }
        out = vr + (vr == vm || round_up);
    }
    *digits = out;
    *exp10 = e10 + removed;
}

// The text encoding/json writes for a float64 (F32 == false) or a float32 (F32 == true: its bits in the low word) with
// the given bits: fills buf (<= 32 bytes), returns the length; 0 for NaN / ±Inf (UnsupportedValueError).
// float32 (floatEncoder with bits == 32): the format switch compares float32(abs) with float32(1e-6) and float32(1e21) —
// both of which have the shortest digits "1e-06" / "1e+21", so the decade of the shortest digits decides as for float64.
template <bool F32 = false>
GOFR_HD uint32_t json_float_text(uint64_t bits, uint8_t* buf) {
    const uint32_t ieee_exponent = F32 ? ((uint32_t)bits >> 23) & 0xFFu : (uint32_t)(bits >> 52) & 0x7FFu;
    const uint64_t ieee_mantissa = F32 ? (bits & 0x7FFFFFu) : (bits & ((1ull << 52) - 1));
    if (ieee_exponent == (F32 ? 0xFFu : 0x7FFu)) return 0;
    uint32_t n = 0;
    if (F32 ? (bits >> 31) & 1 : bits >> 63) buf[n++] = '-';
    if (ieee_exponent == 0 && ieee_mantissa == 0) { buf[n++] = '0'; return n; }
    uint64_t dig;
    int32_t e10;
    if (F32) ryu_shortest<23, 127>(ieee_mantissa, ieee_exponent, &dig, &e10);
    else ryu_shortest<52, 1023>(ieee_mantissa, ieee_exponent, &dig, &e10);
    uint8_t d[20];
    uint32_t nd = 0;
    while (dig) { d[nd++] = (uint8_t)('0' + (uint32_t)(dig % 10)); dig /= 10; }  // least significant first
    const int32_t sci = e10 + (int32_t)nd - 1;  // exponent of the first digit
    // abs < 1e-6 || abs >= 1e21  <=>  sci < -6 || sci >= 21 (the shortest digits denote the float's own decade)
    if (sci < -6 || sci >= 21) {
        buf[n++] = d[nd - 1];
        if (nd > 1) {
            buf[n++] = '.';
            for (uint32_t k = nd - 1; k-- > 0;) buf[n++] = d[k];
        }
        buf[n++] = 'e';
        uint32_t ae = (uint32_t)(sci < 0 ? -sci : sci);
        buf[n++] = sci < 0 ? '-' : '+';
        // %e prints at least two exponent digits; encoding/json then turns "e-0X" into "e-X" (and leaves "e+0X" alone —
        // which cannot occur here: positive exponents start at 21)
        if (ae >= 100) { buf[n++] = (uint8_t)('0' + ae / 100); ae %= 100; buf[n++] = (uint8_t)('0' + ae / 10); buf[n++] = (uint8_t)('0' + ae % 10); }
        else if (ae >= 10) { buf[n++] = (uint8_t)('0' + ae / 10); buf[n++] = (uint8_t)('0' + ae % 10); }
        else { if (sci >= 0) buf[n++] = '0'; buf[n++] = (uint8_t)('0' + ae); }
        return n;
    }
    if (sci < 0) {  // 0.000ddd
        buf[n++] = '0';
        buf[n++] = '.';
        for (int32_t k = 0; k < -sci - 1; k++) buf[n++] = '0';
        for (uint32_t k = nd; k-- > 0;) buf[n++] = d[k];
        return n;
    }
    // ddd[.ddd] or ddd000
    for (int32_t k = 0; k <= sci; k++) buf[n++] = (uint32_t)k < nd ? d[nd - 1 - (uint32_t)k] : (uint8_t)'0';
    if ((int32_t)nd > sci + 1) {
        buf[n++] = '.';
        for (uint32_t k = nd - 1 - (uint32_t)(sci + 1) + 1; k-- > 0;) buf[n++] = d[k];
    }
    return n;
}
GOFR_HD uint32_t json_float64_text(uint64_t bits, uint8_t* buf) { return json_float_text<false>(bits, buf); }
GOFR_HD uint32_t json_float32_text(uint32_t bits, uint8_t* buf) { return json_float_text<true>(bits, buf); }

}  // namespace gofr
