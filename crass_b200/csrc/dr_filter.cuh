// dr_filter.cuh -- 2-bit seed detection for the direct-repeat search (K1 fast path).
//
// searchCore (libcrispr.cpp:295-348) looks, for every window start j = 0, 8, 16, ... <= L-58, for the
// 8-mer read[j, j+8) at a position p with  j+49 <= p <= j+97  and  p+8 <= L-1.  A window without such a
// second occurrence can never reach scanRight / extendPreRepeat.  This header answers "which windows can
// have one?" on a 2-bit recoding of the read:
//
//     code(byte) = (byte >> 1) & 3          A->0 C->1 T->2 G->3 ; N, IUPAC and lower case alias onto these
//
// Equal bytes give equal codes, so a byte-level seed is always a code-level seed: the flags can only
// over-report (aliasing, the excluded last base, positions past the read end), never miss.  Flagged windows
// are re-examined on the bytes (find_left), everything after a confirmed seed runs the exact byte code of
// dr_core.cuh.
//
// Layout: base i of the read sits in bits [2i, 2i+2) of a little-endian word stream R[]; an 8-mer at a
// multiple of 8 is one aligned 16-bit half-word.  For a distance d, half-word h of (R >> 2d) equals half-word h of R
// exactly when window j = 8h re-occurs at j+d.  T + ~R is 0xFFFF per 16-bit lane exactly where T == R (modular add), so
// one VIADDMNMX.U16x2 (max(T + ~R, acc)) compares two windows at one distance AND folds the result into the running
// maximum of the window word: a half-word of acc[k] ends up 0xFFFF iff some distance matched for that window.
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

CB_HD int first_set(uint32_t x) {                                       // index of the lowest set bit, x != 0
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}

// max(a + b, c) per unsigned 16-bit lane, modular add: one VIADDMNMX.U16x2 on sm_100a.  With b = ~R the sum is 0xFFFF
// exactly where a == R, so "compare + fold into the running maximum" is a single instruction, and preparing b costs one
// LOP3 per window word (the min form with b = -R needed a per-lane two's complement, five instructions a word).
CB_HD uint32_t addmax_u16x2(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    return __viaddmax_u16x2(a, b, c);
#else
    auto mx = [](uint32_t x, uint32_t y) { return x > y ? x : y; };
    const uint32_t lo = mx(((a & 0xFFFFu) + (b & 0xFFFFu)) & 0xFFFFu, c & 0xFFFFu);
    const uint32_t hi = mx(((a >> 16) + (b >> 16)) & 0xFFFFu, c >> 16);
    return lo | (hi << 16);
#endif
}
CB_HD uint32_t max3_u16x2(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    return __vimax3_u16x2(a, b, c);
#else
    auto mx = [](uint32_t x, uint32_t y) { return x > y ? x : y; };
    const uint32_t lo = mx(mx(a & 0xFFFFu, b & 0xFFFFu), c & 0xFFFFu);
    const uint32_t hi = mx(mx(a >> 16, b >> 16), c >> 16);
    return lo | (hi << 16);
#endif
}
CB_HD uint32_t byte_perm(uint32_t x, uint32_t y, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, y, sel);
#else
    const uint64_t v = (uint64_t)x | ((uint64_t)y << 32);
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
#endif
}

// 4 bytes -> 8 bits in the TOP byte of the result (base k of the word in bits [24+2k, 24+2k+2)):
// (b & 6) is the 2-bit code shifted left by one, the multiplier gathers the four fields without carries
CB_HD uint32_t code4_top(uint32_t w) { return (w & 0x06060606u) * 0x00820820u; }
CB_HD uint32_t code4(uint32_t w) { return code4_top(w) >> 24; }
// 16 bytes -> 32 bits: the four top bytes are collected with three byte permutes
CB_HD uint32_t pack16(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
    const uint32_t lo = byte_perm(code4_top(w0), code4_top(w1), 0x0073u);
    const uint32_t hi = byte_perm(code4_top(w2), code4_top(w3), 0x0073u);
    return byte_perm(lo, hi, 0x5410u);
}

CB_HD bool has_full_half(uint32_t x) { return (x & 0xFFFFu) == 0xFFFFu || (x >> 16) == 0xFFFFu; }

// NW   : 32-bit words holding the read (16 bases each); R must have NW + 2 entries, the tail is look-ahead
//        (whatever follows the read in the batch, or zeros -- it can only add false positives)
// NWIN : words holding window starts, i.e. floor((max_len - 58) / 16) + 1
// DMIN..DMAX : seed distances, low_dr + low_spacer .. high_dr + high_spacer (49..97 with default options)
// acc[k] (k < NWIN): half-word h of acc[k] is 0xFFFF iff window 16k + 8h has a code-level seed.
template <int NW, int NWIN, int DMIN, int DMAX>
CB_HD void seed_flags(const uint32_t* R, uint32_t* acc) {
    constexpr int OLO = DMIN / 16, OHI = DMAX / 16;
    uint32_t notR[NWIN];                                        // ~window: T + ~R == 0xFFFF per 16-bit lane  <=>  T == window
#pragma unroll
    for (int k = 0; k < NWIN; ++k) { acc[k] = 0u; notR[k] = ~R[k]; }
#pragma unroll
    for (int phi = 0; phi < 16; ++phi) {
        uint32_t T[NW + 1];                                     // the stream shifted right by phi bases
#pragma unroll
        for (int k = OLO; k <= NW; ++k) T[k] = phi ? funnel_r(R[k], R[k + 1], 2 * phi) : R[k];
#pragma unroll
        for (int k = 0; k < NWIN; ++k) {
#pragma unroll
            for (int o = OLO; o <= OHI; ++o) {
                const int d = 16 * o + phi;
                if (d < DMIN || d > DMAX) continue;
                if (k + o > NW - 1) continue;                   // every position of this word pair lies past the read
                acc[k] = addmax_u16x2(T[k + o], notR[k], acc[k]);   // one VIADDMNMX.U16x2 per (window word, distance)
            }
        }
    }
}

template <int NWIN>
CB_HD bool any_flag(const uint32_t* acc) {
    uint32_t m = acc[0];
#pragma unroll
    for (int k = 1; k + 1 < NWIN; k += 2) m = max3_u16x2(m, acc[k], acc[k + 1]);
    if ((NWIN & 1) == 0) m = max3_u16x2(m, acc[NWIN - 1], acc[NWIN - 1]);
    return has_full_half(m);
}

template <int NWIN>
CB_HD uint32_t flag_mask(const uint32_t* acc) {                 // bit h <-> window 8h ; NWIN <= 16
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < NWIN; ++k) {
        if ((acc[k] & 0xFFFFu) == 0xFFFFu) m |= 1u << (2 * k);
        if ((acc[k] >> 16) == 0xFFFFu) m |= 2u << (2 * k);
    }
    return m;
}

template <int NW, int NWIN, int DMIN, int DMAX>
CB_HD bool seed_filter(const uint32_t* R) {
    uint32_t acc[NWIN];
    seed_flags<NW, NWIN, DMIN, DMAX>(R, acc);
    return any_flag<NWIN>(acc);
}

// ---- searchCore driven by the window flags -----------------------------------------------------------------------
// S[0 .. NW+1] is the packed read (dynamically indexable: shared or local memory), mask0 the flags of the windows
// 0, 8, 16, ... (phase 0).  Identical results to cb::search_core: windows without a flag cannot have a seed, flagged
// windows are checked on the bytes, and after a rejected candidate the window grid restarts at back()-1+8
// (libcrispr.cpp:390), for which the flags are recomputed on the re-phased stream.
//
// The search is written as a per-lane state machine with four straight-line stages (pick a flagged window, byte-level
// find_left, process the seed, recompute the flags) so that a warp that walks 32 candidates can execute every stage
// in lock-step: an implementation with early returns out of nested loops left the lanes diverged for good (ncu: 1.7
// active threads per instruction), this form re-converges after every stage.
template <int NW, int NWIN, int DMIN, int DMAX, class Seq>
struct PackedSearch {
    Seq s;                                                      // by value: a refilled lane moves on to another read (rebind)
    uint32_t L;
    const Params& o;
    const uint32_t* S;
    uint32_t* ss;
    const uint32_t cap;
    int se = -1, result = 0, pos = -1;
    uint32_t base = 0, mask = 0, j = 0, n_ss = 0, replen = 0;
    bool done = true, have = false, need_flags = false;

    CB_HD PackedSearch(const Seq& s_, uint32_t L_, const Params& o_, const uint32_t* S_, uint32_t* ss_, uint32_t cap_)
        : s(s_), L(L_), o(o_), S(S_), ss(ss_), cap(cap_) {}

    CB_HD void rebind(const Seq& s_, uint32_t L_) { s = s_; L = L_; }
    CB_HD void init(uint32_t mask0) {
        se = search_end(o, L);
        base = 0; mask = mask0; n_ss = 0; replen = 0; result = 0; pos = -1;
        done = se < 0; have = false; need_flags = false;
    }
    CB_HD void pick() {                                         // next flagged window of the current grid, in order
        have = false;
        if (done) return;
        if (!mask) { done = true; return; }
        const int h = first_set(mask);
        mask &= mask - 1;
        j = base + 8u * (uint32_t)h;
        if (j > (uint32_t)se) { done = true; return; }
        have = true;
    }
    CB_HD void find() {                                         // is the flag a real (byte-level) seed, and where is the leftmost one
        pos = -1;
        if (!have) return;
        uint32_t begin, end;
        window_text(o, L, j, begin, end);
        const int p = find_left(s, begin, end, j, o.window);
        if (p >= 0) pos = (int)(begin + (uint32_t)p);
    }
    CB_HD void seed() {                                         // scanRight / extendPreRepeat / qcFoundRepeats
        need_flags = false;
        if (pos < 0) return;
        bool advance = false; uint32_t nj = 0;
        const int r = process_seed(s, L, o, j, (uint32_t)pos, ss, n_ss, cap, replen, advance, nj);
        if (r != 0) { result = r; done = true; }
        else if (advance) {
            base = nj + 8u;                                     // j = back() - 1, then j += skips (libcrispr.cpp:390,295)
            if (base > (uint32_t)se) done = true; else need_flags = true;
        }
    }
    // seed() cut at the edit distance (dr_core.cuh: seed_begin / seed_feed): seed_start leaves `osa_pending` set and
    // the job in q.job_* when the verdict needs a similarity; the caller computes the distance when it suits the warp
    // and calls seed_resume.
    QcRun q;
    bool osa_pending = false;
    CB_HD void settle(int r, bool advance, uint32_t nj) {
        osa_pending = r == 2;
        if (r == 2) return;
        if (r != 0) { result = r; done = true; }
        else if (advance) {
            base = nj + 8u;
            if (base > (uint32_t)se) done = true; else need_flags = true;
        }
    }
    CB_HD void seed_start() {
        need_flags = false; osa_pending = false;
        if (pos < 0) return;
        bool advance = false; uint32_t nj = 0;
        const int r = seed_begin(s, L, o, j, (uint32_t)pos, ss, n_ss, cap, replen, advance, nj, q);
        settle(r, advance, nj);
    }
    CB_HD void seed_resume(int distance) {
        if (!osa_pending) return;
        bool advance = false; uint32_t nj = 0;
        const int r = seed_feed<Seq>(L, ss, n_ss, advance, nj, q, distance);
        settle(r, advance, nj);
    }
    CB_HD void reflag() {                                       // the window grid moved: flags of the re-phased stream
        if (!need_flags) return;
        const uint32_t q = base >> 4, sh = (base & 15u) * 2u;
        uint32_t Q[NW + 2];
#pragma unroll
        for (int k = 0; k < NW + 2; ++k) {
            const uint32_t lo = (q + k < (uint32_t)(NW + 2)) ? S[q + k] : 0u;
            const uint32_t hi = (q + k + 1 < (uint32_t)(NW + 2)) ? S[q + k + 1] : 0u;
            Q[k] = funnel_r(lo, hi, sh);
        }
        uint32_t acc[NWIN];
        seed_flags<NW, NWIN, DMIN, DMAX>(Q, acc);
        mask = flag_mask<NWIN>(acc);
    }
};

// the staged form driven for one read (what a lane of k_dr_exact_staged goes through, without the waiting)
template <int NW, int NWIN, int DMIN, int DMAX, class Seq>
CB_HD int search_core_staged(const Seq& s, uint32_t L, const Params& o, const uint32_t* S, uint32_t mask0,
                             uint32_t* ss, uint32_t cap, uint32_t& n_ss, uint32_t& replen) {
    PackedSearch<NW, NWIN, DMIN, DMAX, Seq> st(s, L, o, S, ss, cap);
    st.init(mask0);
    for (;;) {
        st.pick();
        if (!st.have) break;
        st.find();
        st.seed_start();
        while (st.osa_pending) st.seed_resume(edit_distance(s, st.q.job_a0, st.q.job_n, st.q.job_b0, st.q.job_m));
        st.reflag();
    }
    n_ss = st.result == 1 ? st.n_ss : 0;
    replen = st.replen;
    return st.result;
}

template <int NW, int NWIN, int DMIN, int DMAX, class Seq>
CB_HD int search_core_packed(const Seq& s, uint32_t L, const Params& o, const uint32_t* S, uint32_t mask0,
                             uint32_t* ss, uint32_t cap, uint32_t& n_ss, uint32_t& replen) {
    PackedSearch<NW, NWIN, DMIN, DMAX, Seq> st(s, L, o, S, ss, cap);
    st.init(mask0);
    for (;;) {
        st.pick();
        if (!st.have) break;
        st.find();
        st.seed();
        st.reflag();
    }
    n_ss = st.result == 1 ? st.n_ss : 0;
    replen = st.replen;
    return st.result;
}

}  // namespace cb
