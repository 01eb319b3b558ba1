// dr_filter.cuh -- 2-bit "any seed?" pre-filter for the direct-repeat search (K1 fast path).
//
// searchCore (libcrispr.cpp:295-348) looks, for every window start j = 0, 8, 16, ... <= L-58, for the
// 8-mer read[j, j+8) at a position p with  j+49 <= p <= j+97  and  p+8 <= L-1.  A read in which no window
// has such a second occurrence can never reach scanRight / extendPreRepeat, so searchCore returns false
// for it.  This header answers exactly that question on a 2-bit recoding of the read:
//
//     code(byte) = (byte >> 1) & 3          A->0 C->1 T->2 G->3 ; N, IUPAC and lower case alias onto these
//
// Equal bytes give equal codes, so a byte-level seed is always a code-level seed: the filter can only
// over-report (aliasing, the excluded last base, windows/positions past the read end), never miss.  Reads it
// flags go to the exact kernel, which re-does the whole search on the bytes.
//
// Layout: base i of the read sits in bits [2i, 2i+2) of a little-endian word stream R[]; an 8-mer at a
// multiple of 8 is one aligned 16-bit half-word.  For a distance d, (R >> 2d) XOR R has a zero half-word h
// exactly when window j = 8h re-occurs at j+d, so one funnel shift + one XOR test two windows, and a packed
// unsigned 16-bit minimum (VIMNMX3.U16x2 on sm_100a) folds the tests: a half-word of the running minimum is
// zero iff some (window, distance) pair matched.
#pragma once
#include <stdint.h>

#include "dr_core.cuh"

namespace cb {

CB_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {        // low word of (hi:lo) >> (sh & 31)
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    sh &= 31;
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}

CB_HD uint32_t min3_u16x2(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    return __vimin3_u16x2(a, b, c);
#else
    auto mn = [](uint32_t x, uint32_t y) { return x < y ? x : y; };
    const uint32_t lo = mn(mn(a & 0xFFFFu, b & 0xFFFFu), c & 0xFFFFu);
    const uint32_t hi = mn(mn(a >> 16, b >> 16), c >> 16);
    return lo | (hi << 16);
#endif
}

// 4 bytes -> 8 bits (base k of the word in bits [2k, 2k+2))
CB_HD uint32_t code4(uint32_t w) { return (((w >> 1) & 0x03030303u) * 0x01041040u) >> 24; }
// 16 bytes -> 32 bits
CB_HD uint32_t pack16(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
    return code4(w0) | (code4(w1) << 8) | (code4(w2) << 16) | (code4(w3) << 24);
}

CB_HD bool has_zero_half(uint32_t x) { return (x & 0xFFFFu) == 0 || (x >> 16) == 0; }

// NW   : 32-bit words holding the read (16 bases each); R must have NW + 2 entries, the tail is look-ahead
//        (whatever follows the read in the batch, or zeros -- it can only add false positives)
// NWIN : words holding window starts, i.e. floor((max_len - 58) / 16) + 1
// DMIN..DMAX : seed distances, low_dr + low_spacer .. high_dr + high_spacer (49..97 with default options)
template <int NW, int NWIN, int DMIN, int DMAX>
CB_HD bool seed_filter(const uint32_t* R) {
    uint32_t acc0 = 0xFFFFFFFFu, acc1 = 0xFFFFFFFFu;
    constexpr int OLO = DMIN / 16, OHI = DMAX / 16;
#pragma unroll
    for (int phi = 0; phi < 16; ++phi) {
        uint32_t T[NW + 1];
#pragma unroll
        for (int k = OLO; k <= NW; ++k) T[k] = phi ? funnel_r(R[k], R[k + 1], 2 * phi) : R[k];
        uint32_t pend = 0xFFFFFFFFu;
        bool has_pend = false;
#pragma unroll
        for (int o = OLO; o <= OHI; ++o) {
            const int d = 16 * o + phi;
            if (d < DMIN || d > DMAX) continue;
#pragma unroll
            for (int k = 0; k < NWIN; ++k) {
                if (k + o > NW - 1) continue;               // every position of this word pair lies past the read
                const uint32_t x = R[k] ^ T[k + o];
                if (has_pend) {
                    if ((k & 1) == 0) acc0 = min3_u16x2(acc0, pend, x); else acc1 = min3_u16x2(acc1, pend, x);
                    has_pend = false;
                } else { pend = x; has_pend = true; }
            }
        }
        if (has_pend) acc0 = min3_u16x2(acc0, pend, pend);
    }
    return has_zero_half(acc0) || has_zero_half(acc1);
}

}  // namespace cb
