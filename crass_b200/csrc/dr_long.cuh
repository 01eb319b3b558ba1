// dr_long.cuh -- K1 for long reads: one WARP per read (device code only).
//
// searchCore on a 1-10 kb read is ~120-1250 windows; the reference walks them one by one.  Here the warp
//   1. recodes the read to 2 bits per base into shared memory (coalesced 128-bit loads, cb::pack16),
//   2. evaluates the seed flags of 32 segments x 12 windows per round with the same register kernel as the short-read
//      filter (cb::seed_flags<13,6,49,97>: one VIADDMNMX.U16x2 per window word and distance),
//   3. walks the flagged windows in order; a flag is confirmed on the bytes (warp-parallel find_left) and the candidate
//      array is handled by warp-cooperative forms of scanRight / extendPreRepeat / qcFoundRepeats: the control flow is
//      uniform across the warp (every lane holds the same scalars), the inner loops -- positions of a text search, the
//      repeats voting on a column, the spacer/repeat edit distances -- are spread over the lanes,
//   4. when a candidate is rejected the window grid restarts at back()-1+8 (libcrispr.cpp:390) and the flags are
//      recomputed on the re-phased stream from there.
// Results are bit-identical to cb::search_core (dr_core.cuh); tests/test_gpu_parity.py::test_long_reads* compare them
// with the oracle read by read.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dr_core.cuh"
#include "dr_filter.cuh"

namespace cbl {

using cb::Params;

constexpr uint32_t kFull = 0xFFFFFFFFu;

struct GSeq {                                   // bytes of the read through the read-only path
    const uint8_t* p;
    __device__ __forceinline__ uint8_t operator[](uint32_t i) const { return __ldg(p + i); }
};

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// leftmost p in [0, tl-8] with s[b+p, b+p+8) == s[pat, pat+8), or -1; 32 positions per round
__device__ __forceinline__ int warp_find_left8(const GSeq& s, uint32_t b, uint32_t e, uint32_t pat) {
    if (e <= b) return -1;
    const uint32_t tl = e - b;
    if (tl < 8) return -1;
    uint32_t k0 = 0, k1 = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { k0 |= (uint32_t)s[pat + i] << (8 * i); k1 |= (uint32_t)s[pat + 4 + i] << (8 * i); }
    for (uint32_t p0 = 0; p0 + 8 <= tl; p0 += 32) {
        const uint32_t p = p0 + lane_id();
        bool ok = p + 8 <= tl;
        if (ok) {
            uint32_t w0 = 0, w1 = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) { w0 |= (uint32_t)s[b + p + i] << (8 * i); w1 |= (uint32_t)s[b + p + 4 + i] << (8 * i); }
            ok = (w0 == k0) && (w1 == k1);
        }
        const uint32_t bal = __ballot_sync(kFull, ok);
        if (bal) return (int)(p0 + (uint32_t)__ffs((int)bal) - 1u);
    }
    return -1;
}

// scanRight (libcrispr.cpp:170-263), uniform control, lane 0 appends to ss
__device__ __forceinline__ void warp_scan_right(const GSeq& s, uint32_t L, uint32_t* ss, uint32_t& n_ss, uint32_t cap, uint32_t pat,
                                                uint32_t min_spacer, uint32_t scan_range) {
    const uint32_t w = 8;
    uint32_t last = ss[n_ss - 2], second_last = ss[n_ss - 4];
    uint32_t spacing = last - second_last;
    for (;;) {
        const uint32_t cand = last + spacing;
        uint32_t begin = cand - scan_range;
        uint32_t end = cand + w + scan_range;
        const uint32_t min_begin = last + w + min_spacer;
        if (begin < min_begin) begin = min_begin;
        if (begin > L - 1) break;
        if (end > L) end = L;
        if (begin >= end) break;
        const int pos = warp_find_left8(s, begin, end, pat);
        if (pos < 0) break;
        if (n_ss + 2 > cap) break;
        const uint32_t st = begin + (uint32_t)pos;
        uint32_t en = st + w - 1;
        if (en >= L) en = L - 1;
        if (lane_id() == 0) { ss[n_ss] = st; ss[n_ss + 1] = en; }
        n_ss += 2;
        second_last = last;
        last = st;
        spacing = last - second_last;
        if (spacing < min_spacer + w) break;
    }
    __syncwarp();
}

__device__ __forceinline__ bool vote_passes(bool valid, uint8_t c, int cut_off, int* acc) {
    acc[0] += __popc(__ballot_sync(kFull, valid && c == 'A'));
    acc[1] += __popc(__ballot_sync(kFull, valid && c == 'C'));
    acc[2] += __popc(__ballot_sync(kFull, valid && c == 'G'));
    acc[3] += __popc(__ballot_sync(kFull, valid && c == 'T'));
    (void)cut_off;
    return true;
}

// extendPreRepeat (libcrispr.cpp:520-772): the repeats vote lane-parallel, the control flow is uniform
__device__ __forceinline__ uint32_t warp_extend(const GSeq& s, uint32_t L, uint32_t* ss, uint32_t n_ss, uint32_t window, uint32_t min_spacer) {
    const uint32_t lane = lane_id();
    const uint32_t num_repeats = n_ss / 2;
    uint32_t rep = window;
    int cut_off = (int)num_repeats - 1;
    if (2 > cut_off) cut_off = 2;
    const uint32_t first = ss[0], last = ss[n_ss - 2];
    uint32_t msp = 0xFFFFFFFFu;
    for (uint32_t i = 2 + 2 * lane; i < n_ss; i += 64) msp = min(msp, ss[i] - ss[i - 2]);
    msp = __reduce_min_sync(kFull, msp);
    uint32_t right = 0;
    uint32_t max_right = msp - min_spacer;
    uint32_t idx_end = n_ss;
    while (max_right > 0) {
        if (last + window + right >= L) idx_end -= 2;
        int acc[4] = {0, 0, 0, 0};
        const uint32_t lim = min(idx_end, n_ss);
        for (uint32_t k0 = 0; k0 < lim; k0 += 64) {
            const uint32_t k = k0 + 2 * lane;
            const bool valid = k < lim && ss[k] + rep < L;      // the reference stops at the first repeat whose column is past the end;
            uint8_t c = 0;                                      // starts ascend, so that is the same as skipping all of them
            if (valid) c = s[ss[k] + rep];
            vote_passes(valid, c, cut_off, acc);
        }
        if (acc[0] >= cut_off || acc[1] >= cut_off || acc[2] >= cut_off || acc[3] >= cut_off) { rep++; max_right--; right++; }
        else break;
    }
    uint32_t left = 0;
    const int test_for_negative = (int)(msp - rep);
    const uint32_t max_left = test_for_negative >= 0 ? (uint32_t)test_for_negative : 0;
    uint32_t idx_start = 0;
    while (left < max_left) {
        if ((int)first - (int)left <= 0) idx_start += 2;
        int acc[4] = {0, 0, 0, 0};
        for (uint32_t k0 = idx_start; k0 < n_ss; k0 += 64) {
            const uint32_t k = k0 + 2 * lane;
            bool valid = k < n_ss;
            uint32_t at = 0;
            if (valid) { at = ss[k] - left - 1; valid = at < L; }
            uint8_t c = 0;
            if (valid) c = s[at];
            vote_passes(valid, c, cut_off, acc);
        }
        if (acc[0] >= cut_off || acc[1] >= cut_off || acc[2] >= cut_off || acc[3] >= cut_off) { rep++; left++; }
        else break;
    }
    __syncwarp();
    for (uint32_t k = 2 * lane; k + 1 < n_ss; k += 64) {
        const uint32_t a = ss[k], b = ss[k + 1];
        ss[k] = a < left ? 0 : a - left;
        ss[k + 1] = (b + right >= L) ? L - 1 : b + right;
    }
    __syncwarp();
    return rep;
}

// qcFoundRepeats (libcrispr.cpp:869-1029): cheap tests uniform, the 2(n-2) edit distances spread over the lanes, the float
// sums taken in the reference's order from the per-pair similarities parked in `sims`
__device__ __forceinline__ int warp_qc(const GSeq& s, uint32_t L, const uint32_t* ss, uint32_t n_ss, int min_spacer, int max_spacer, float* sims) {
    const uint32_t n = n_ss / 2;
    if (n < 2) return -1;
    if (n < 3) return cb::qc_found_repeats(s, L, ss, n_ss, min_spacer, max_spacer);   // one distance: every lane computes it
    const uint32_t lane = lane_id();
    const uint32_t r0 = ss[0];
    const uint32_t rl = cb::substr_len(L, r0, ss[1] - ss[0] + 1);
    if (cb::low_complexity(s, r0, rl)) return 0;
    const uint32_t nsp = n - 1;
    int min_len = 10000000, max_len = 0;
    float ssl = 0.0f, rsl = 0.0f;
    uint32_t prev_len = 0;
    for (uint32_t i = 0; i < nsp; ++i) {
        uint32_t st, len;
        cb::spacer_at(ss, L, i, st, len);
        if ((int)len < min_len) min_len = (int)len;
        if ((int)len > max_len) max_len = (int)len;
        if (i > 0) {
            ssl = cb::f_add(ssl, cb::f_sub((float)prev_len, (float)len));
            rsl = cb::f_add(rsl, cb::f_sub((float)rl, (float)prev_len));
        }
        prev_len = len;
    }
    const float nc = (float)(nsp - 1);
    if (min_len < min_spacer) return 0;
    if (max_len > max_spacer) return 0;
    float a_ssl = cb::f_div(ssl, nc); if (a_ssl < 0) a_ssl = -a_ssl;
    float a_rsl = cb::f_div(rsl, nc); if (a_rsl < 0) a_rsl = -a_rsl;
    if ((int)a_ssl > 12) return 0;
    if ((int)a_rsl > 30) return 0;
    if (rl > (uint32_t)cb::kMaxEdit || max_len > cb::kMaxEdit) return -1;
    // item 2i = sim(repeat, spacer i), item 2i+1 = sim(spacer i, spacer i+1), i = 0 .. nsp-2
    const uint32_t n_items = 2 * (nsp - 1);
    for (uint32_t t = lane; t < n_items; t += 32) {
        const uint32_t i = t >> 1;
        uint32_t st0, len0;
        cb::spacer_at(ss, L, i, st0, len0);
        float v;
        if ((t & 1) == 0) v = cb::similarity(s, r0, rl, st0, len0);
        else {
            uint32_t st1, len1;
            cb::spacer_at(ss, L, i + 1, st1, len1);
            v = cb::similarity(s, st0, len0, st1, len1);
        }
        sims[t] = v;
    }
    __syncwarp();
    float rs = 0.0f, sp = 0.0f;
    for (uint32_t i = 0; i + 1 < nsp; ++i) {
        rs = cb::f_add(rs, sims[2 * i]);
        float ss_diff = 0.0f;
        ss_diff = cb::f_add(ss_diff, sims[2 * i + 1]);
        sp = cb::f_add(sp, ss_diff);
    }
    __syncwarp();
    sp = cb::f_div(sp, nc);
    rs = cb::f_div(rs, nc);
    if ((double)sp > 0.82) return 0;
    if ((double)rs > 0.82) return 0;
    return 1;
}

// process_seed (dr_core.cuh), warp-cooperative
__device__ __forceinline__ int warp_process_seed(const GSeq& s, uint32_t L, const Params& o, uint32_t j, uint32_t p, uint32_t* ss,
                                                 uint32_t& n_ss, uint32_t cap, uint32_t& replen, bool& advance, uint32_t& next_j, float* sims) {
    const uint32_t w = 8;
    if (lane_id() == 0) {
        ss[0] = j; ss[1] = min(j + w - 1, L - 1);
        ss[2] = p; ss[3] = min(p + w - 1, L - 1);
    }
    n_ss = 4;
    __syncwarp();
    warp_scan_right(s, L, ss, n_ss, cap, j, o.low_spacer, o.scan_range);
    advance = false;
    if (n_ss / 2 >= o.min_repeats) {
        const uint32_t len = warp_extend(s, L, ss, n_ss, w, o.low_spacer);
        replen = len;
        if (len >= o.low_dr && len <= o.high_dr) {
            const int q = warp_qc(s, L, ss, n_ss, (int)o.low_spacer, (int)o.high_spacer, sims);
            if (q == 1) return 1;
            if (q < 0) return q;
        }
        advance = true;
        next_j = ss[n_ss - 1] - 1;
    }
    n_ss = 0;
    return 0;
}

}  // namespace cbl
