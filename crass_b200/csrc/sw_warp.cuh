// sw_warp.cuh -- K6, one warp per found read: the alignments of sw_core.cuh as a lane wavefront.
//
// The thread-per-read form (sw_core.cuh, kept as k_update_start_stops_thread and as the code the CPU fuzz runs) leaves
// 4 of 32 lanes busy on the bench workload (ncu r1e: thread_inst_executed_per_inst_executed 4.1): most found reads need
// no alignment at all, some need two, and their sizes differ.  Here a warp owns the read.  The DR runs across the lanes,
// CPL matrix columns per lane (1, 2 or 4: DRs up to 32, 64, 128 bytes), the read runs down the rows, and lane l works
// on row t - l at step t: its left neighbour finished that row one step earlier, so the value to the left of a lane's
// first column arrives by one shuffle per step and the diagonal one is the previous step's left value.  A lane keeps the
// row above for its columns in registers (score + end of the predecessor walk, see sw_core.cuh) -- no matrix, no shared
// or local memory.  Every cell sees the same operands in the same order as in the reference (three IEEE double additions
// and findMax's comparison tree), so scores and tie-breaks are bit-identical; the first maximum in row-major order is
// recovered by a lexicographic (value, -row, -column) warp reduction.
// The similarity test behind every alignment (levenstheinDistance of the two returned strings) runs on all lanes alike
// in the bit-parallel form (the DR side is at most 64 bytes there; the distance is symmetric in its arguments), lane 0
// falls back to the row DP otherwise.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sw_core.cuh"

namespace cbw {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ double shfl_up_d(double v, int d) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_up_sync(kFull, lo, d); hi = __shfl_up_sync(kFull, hi, d);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_xor_d(double v, int m) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(kFull, lo, m); hi = __shfl_xor_sync(kFull, hi, m);
    return __hiloint2double(hi, lo);
}

// levenstheinDistance(read[a0, a0+al), bb[0, bl)) with the DR side as the bit vector (bl <= 64, distance is symmetric).
// -1 if a byte outside A/C/G/T/N occurs or bl > 64.  Uniform over the warp: every lane computes the same value.
template <class Seq>
__device__ int osa_bitpar_dr_vs_read(const uint8_t* bb, uint32_t bl, const Seq& a, uint32_t a0, uint32_t al) {
    if (bl == 0 || al == 0 || bl > 64) return -1;
    uint64_t p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0;
    bool bad = false;
    for (uint32_t i = 0; i < bl; ++i) {
        const int k = cb::osa_slot(bb[i]);
        const uint64_t bit = 1ull << i;
        bad |= k < 0;
        p0 |= k == 0 ? bit : 0; p1 |= k == 1 ? bit : 0; p2 |= k == 2 ? bit : 0; p3 |= k == 3 ? bit : 0; p4 |= k == 4 ? bit : 0;
    }
    const uint64_t top = 1ull << (bl - 1);
    uint64_t vp = bl >= 64 ? ~0ull : ((1ull << bl) - 1ull), vn = 0, d0 = 0, pm_prev = 0;
    int score = (int)bl;
    for (uint32_t j = 0; j < al; ++j) {
        const int k = cb::osa_slot(a[a0 + j]);
        bad |= k < 0;
        const uint64_t pm = k == 0 ? p0 : k == 1 ? p1 : k == 2 ? p2 : k == 3 ? p3 : k == 4 ? p4 : 0;
        uint64_t tr = ((((~d0) & pm) << 1) & pm_prev) & ~3ull;          // rows 1 and 2 never transpose
        if (j < 2) tr = 0;                                              // columns 1 and 2 never transpose
        d0 = (((pm & vp) + vp) ^ vp) | pm | vn | tr;
        const uint64_t hp = vn | ~(d0 | vp);
        const uint64_t hn = d0 & vp;
        score += (hp & top) ? 1 : 0;
        score -= (hn & top) ? 1 : 0;
        const uint64_t x = (hp << 1) | 1ull;
        vp = (hn << 1) | ~(d0 | x);
        vn = d0 & x;
        pm_prev = pm;
    }
    return bad ? -1 : score;
}

// the similarity test of smithWaterman (SmithWaterman.cpp:291-303); all lanes call it and get the same answer
template <class Seq>
__device__ bool sw_similarity_ok_warp(const Seq& a, const uint8_t* bb, const cb::SwResult& r, double similarity) {
    int d = osa_bitpar_dr_vs_read(bb + r.b_pos, r.b_len, a, r.a_pos, r.a_len);
    if (d < 0) {                                         // uniform: every lane saw the same bytes
        if ((threadIdx.x & 31u) == 0) d = cb::osa_distance_read_vs_dr(a, r.a_pos, r.a_len, bb + r.b_pos, r.b_len);
        d = __shfl_sync(kFull, d, 0);
    }
    const double sim_ld = __dsub_rn(1.0, __ddiv_rn((double)d, (double)r.a_len));
    return sim_ld >= similarity;
}

// smithWaterman(read, DR, ..., start, len, similarity) by one warp; bb = the DR in shared memory.  r is the same on
// every lane.  Needs len >= 1, 1 <= lb <= 32 * CPL, start + len <= la.
template <int CPL, class Seq>
__device__ void smith_waterman_warp(const Seq& a, uint32_t la, const uint8_t* bb, uint32_t lb, int start, int len,
                                    double similarity, cb::SwResult& r) {
    const int lane = (int)(threadIdx.x & 31u);
    const int m = (int)lb;
    const int nlanes = (m + CPL - 1) / CPL;
    const int j0 = lane * CPL + 1;                       // my first column (1-based)
    double up[CPL];
    uint32_t oup[CPL];
    uint8_t bcol[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) { up[c] = 0; oup[c] = 0; bcol[c] = (j0 + c <= m) ? bb[j0 + c - 1] : 0; }
    double diag_in = 0, last_v = 0, best_v = -1;
    uint32_t odiag_in = 0, last_o = 0, best_o = 0;
    int best_i = 0, best_j = 0;
    const int steps = len + nlanes - 1;
    for (int t = 0; t < steps; ++t) {
        double lv = shfl_up_d(last_v, 1);
        uint32_t lo = __shfl_up_sync(kFull, last_o, 1);
        if (lane == 0) { lv = 0; lo = 0; }               // column 0
        const int i = t - lane + 1;
        if (i >= 1 && i <= len && lane < nlanes) {
            const uint8_t ai = a[(uint32_t)(i - 1 + start)];
            double left = lv, diag = diag_in;
            uint32_t o_left = lo, o_diag = odiag_in;
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const int j = j0 + c;
                if (j <= m) {
                    int index;
                    const double v = cb::sw_find_max(__dadd_rn(diag, ai == bcol[c] ? 1.2 : -1.0), __dadd_rn(up[c], -1.0),
                                                     __dadd_rn(left, -1.0), 0.0, index);
                    uint32_t o = ((uint32_t)i << 8) | (uint32_t)j;
                    if (index == 0) { if (i > 1 && j > 1) o = o_diag; }
                    else if (index == 1) { if (i > 1) o = oup[c]; }
                    else if (index == 2) { if (j > 1) o = o_left; }
                    if (v > best_v) { best_v = v; best_i = i; best_j = j; best_o = o; }
                    diag = up[c]; o_diag = oup[c];
                    left = v; o_left = o;
                    up[c] = v; oup[c] = o;
                }
            }
            diag_in = lv; odiag_in = lo;
            last_v = left; last_o = o_left;
        }
    }
    // first maximum in row-major order: greatest value, then smallest row, then smallest column
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const double ov = shfl_xor_d(best_v, d);
        const int oi = __shfl_xor_sync(kFull, best_i, d), oj = __shfl_xor_sync(kFull, best_j, d);
        const uint32_t oo = __shfl_xor_sync(kFull, best_o, d);
        const bool take = ov > best_v || (ov == best_v && (oi < best_i || (oi == best_i && oj < best_j)));
        if (take) { best_v = ov; best_i = oi; best_j = oj; best_o = oo; }
    }
    int ci = (int)(best_o >> 8) - 1, cj = (int)(best_o & 255u) - 1;
    if (cj < 0) cj = 0;
    if (ci < 0) ci = 0;
    r.start_align = ci + start;
    r.end_align = r.start_align + best_i - ci - 1;
    r.a_pos = (uint32_t)(ci + start);
    r.a_len = (uint32_t)(best_i - ci + start);
    if (r.a_len > la - r.a_pos) r.a_len = la - r.a_pos;
    r.b_pos = (uint32_t)cj;
    r.b_len = (uint32_t)(best_j - cj);
    if (r.b_len > lb - r.b_pos) r.b_len = lb - r.b_pos;
    r.ok = 1;
    if (similarity != 0 && !sw_similarity_ok_warp(a, bb, r, similarity)) {
        r.ok = 0; r.start_align = 0; r.end_align = 0;
        r.a_pos = r.a_len = r.b_pos = r.b_len = 0;
    }
}

template <class Seq>
__device__ __forceinline__ void smith_waterman_warp_any(const Seq& a, uint32_t la, const uint8_t* bb, uint32_t lb, int start,
                                                        int len, double similarity, cb::SwResult& r) {
    if (lb <= 32) smith_waterman_warp<1>(a, la, bb, lb, start, len, similarity, r);
    else if (lb <= 64) smith_waterman_warp<2>(a, la, bb, lb, start, len, similarity, r);
    else smith_waterman_warp<4>(a, la, bb, lb, start, len, similarity, r);
}

// update_start_stops of sw_core.cuh by one warp.  bb: kMaxSwDr + 1 bytes of shared memory owned by the warp.
// Scalar bookkeeping is done by all lanes alike; lane 0 writes.  Returns the status on every lane.
template <class Seq, class DrSeq>
__device__ uint8_t update_start_stops_warp(const Seq& s, uint32_t L, const uint32_t* ss_in, uint32_t n_in, int front_offset,
                                           const DrSeq& dr, uint32_t dr_len, uint32_t low_spacer, uint8_t* bb, uint32_t* out,
                                           uint32_t& n_out) {
    const int lane = (int)(threadIdx.x & 31u);
    n_out = 0;
    if (n_in < 2 || (n_in & 1)) return cb::kUssBadList;
    if (dr_len > (uint32_t)cb::kMaxSwDr || dr_len == 0) return cb::kUssDrTooLong;
    if (L >= cb::kMaxSwRead) return cb::kUssReadTooLong;
    const int dr_length = (int)dr_len;
    uint32_t first_start = 0, last_end = 0;
    bool past = false;
    for (uint32_t k = 2 * (uint32_t)lane; k < n_in; k += 64) {             // ReadHolder.cpp:392-437, one repeat per lane
        int usable = dr_length - 1;
        uint32_t st = ss_in[k];
        if (front_offset >= (int)st) { usable = dr_length - (front_offset - (int)st) - 1; st = 0; }
        else st -= (uint32_t)front_offset;
        past |= st >= L;
        uint32_t en = st + (uint32_t)usable;
        if (en >= L) en = L - 1;
        if (k == 0) first_start = st;
        if (k + 2 == n_in) last_end = en;
    }
    if (__any_sync(kFull, past)) return cb::kUssPastRead;                  // nothing written
    first_start = __shfl_sync(kFull, first_start, 0);
    last_end = __shfl_sync(kFull, last_end, (int)(((n_in - 2) / 2) & 31u));
    for (uint32_t j = (uint32_t)lane; j < dr_len; j += 32) bb[j] = dr[j];
    __syncwarp();
    cb::SwResult r;
    bool front = false, back = false;
    uint32_t front_end = 0, back_start = 0, back_end = 0;
    if (first_start > low_spacer) {                                        // :443-481
        // alignment first, the caller's cheap tests next, the edit distance only for the few that pass (see sw_core.cuh)
        smith_waterman_warp_any(s, L, bb, dr_len, 0, (int)(first_start - low_spacer), 0.0, r);
        if (r.end_align != 0 && r.end_align - r.start_align >= 4) {
            const int at = cb::bytes_find(bb, dr_len, bb + r.b_pos, r.b_len, true);
            if (at >= 0 && (uint32_t)at + r.b_len == dr_len && r.start_align == 0 && sw_similarity_ok_warp(s, bb, r, 0.85)) { front = true; front_end = (uint32_t)r.end_align; }
        }
    }
    const uint32_t end_dist = L - last_end;                                // :483-510
    if (end_dist > low_spacer) {
        smith_waterman_warp_any(s, L, bb, dr_len, (int)(last_end + low_spacer), (int)(end_dist - low_spacer), 0.0, r);
        if (r.end_align != 0 && r.end_align - r.start_align >= 4) {
            if ((int)L - 1 == r.end_align && cb::bytes_find(bb, dr_len, bb + r.b_pos, r.b_len, false) == 0 && sw_similarity_ok_warp(s, bb, r, 0.85)) {
                int diff = (int)r.a_len - (int)r.b_len;
                if (diff < 0) diff = -diff;
                back = true;
                back_start = (uint32_t)(r.start_align + diff);
                back_end = (uint32_t)r.end_align;
                if (back_end >= L) back_end = L - 1;
            }
        }
    }
    const uint32_t base = front ? 2u : 0u;
    for (uint32_t k = 2 * (uint32_t)lane; k < n_in; k += 64) {             // the shifted repeats again, now to their final place
        int usable = dr_length - 1;
        uint32_t st = ss_in[k];
        if (front_offset >= (int)st) { usable = dr_length - (front_offset - (int)st) - 1; st = 0; }
        else st -= (uint32_t)front_offset;
        uint32_t en = st + (uint32_t)usable;
        if (en >= L) en = L - 1;
        out[base + k] = st; out[base + k + 1] = en;
    }
    if (lane == 0) {
        if (front) { out[0] = 0; out[1] = front_end; }
        if (back) { out[base + n_in] = back_start; out[base + n_in + 1] = back_end; }
    }
    n_out = n_in + base + (back ? 2u : 0u);
    return cb::kUssOk;
}

}  // namespace cbw
