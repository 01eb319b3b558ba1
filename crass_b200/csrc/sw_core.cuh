// sw_core.cuh -- partial-DR recovery: the first consumer of the path's start/stop lists.
//
// `__host__ __device__`, free of warp intrinsics, so that tests/hostsim can compile the very same source with g++
// and fuzz it against the parity oracle without a GPU.  The product never runs it on the CPU.
//
// Behavioural contract (bit-exact with the reference, /root/reference/src/crass/...):
//   sw_find_max          == findMax                       SmithWaterman.cpp:68-131
//   smith_waterman       == smithWaterman(7 arguments)    SmithWaterman.cpp:151-308
//   update_start_stops   == ReadHolder::updateStartStops  ReadHolder.cpp:382-511
//
// The reference fills three (search length + 1) x (DR length + 1) matrices (score, and the predecessor of every cell
// as two int matrices) and walks the predecessors back from the first maximum in row-major order.  Here only one row
// is kept: next to its score every cell carries where that walk would END if it started in the cell -- the walk stops
// at the first cell whose predecessor is itself or lies in row/column 0, so the end of a cell's walk is its own
// position in those cases and its predecessor's end otherwise.  That is the same function of the same predecessor
// choices, computed forwards, in O(DR length) memory instead of O(search length x DR length).
// Scores are IEEE doubles (match 1.2, mismatch -1, gap -1, SmithWaterman.h:60-63) added in the reference's order: sums
// of 1.2 are not exact, so which of two paths is "greater" depends on the order of the additions and must not change.
#pragma once
#include <stdint.h>

#include "dr_core.cuh"

namespace cb {

constexpr int kMaxSwDr = 127;              // longest DR (matrix columns); status kUssDrTooLong above
constexpr uint32_t kMaxSwRead = 1u << 16;  // the edit-distance rows are 16 bit; status kUssReadTooLong above

enum UssStatus : uint8_t {
    kUssOk = 0,
    kUssBadList = 1,        // fewer than one repeat or an odd number of entries
    kUssDrTooLong = 2,
    kUssPastRead = 3,       // a shifted start lies at or past the end of the read (the reference logs "Something wrong
                            // with front offset!" for > and then reads out of bounds; nothing is written here)
    kUssReadTooLong = 4,
};

CB_HD double d_add(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}

// value of the cell and which argument won: 0 = a (diagonal), 1 = b (up), 2 = c (left), 3 = d (zero)
CB_HD double sw_find_max(double a, double b, double c, double d, int& index) {
    if (b > a) {
        if (c > d) { if (c > b) { index = 2; return c; } index = 1; return b; }
        if (d > b) { index = 3; return d; }
        index = 1; return b;
    }
    if (c > d) { if (c > a) { index = 2; return c; } index = 0; return a; }
    if (d > a) { index = 3; return d; }
    index = 0; return a;
}

// levenstheinDistance(a, b) with a = s[a0, a0+n) (n < 65536) and b = bb[0, m) (m <= kMaxSwDr): the same recurrence
// as dr_core.cuh's osa_distance, rows over b, 16-bit cells.
template <class Seq>
CB_HD_NOINLINE int osa_distance_read_vs_dr(const Seq& s, uint32_t a0, uint32_t n, const uint8_t* bb, uint32_t m) {
    if (n == 0) return (int)m;
    if (m == 0) return (int)n;
    uint16_t r0[kMaxSwDr + 1], r1[kMaxSwDr + 1], r2[kMaxSwDr + 1];
    uint16_t* prev2 = r0; uint16_t* prev = r1; uint16_t* cur = r2;
    for (uint32_t j = 0; j <= m; ++j) prev[j] = (uint16_t)j;
    uint8_t a_prev = 0;
    for (uint32_t i = 1; i <= n; ++i) {
        const uint8_t s_i = s[a0 + i - 1];
        cur[0] = (uint16_t)i;
        for (uint32_t j = 1; j <= m; ++j) {
            const uint8_t t_j = bb[j - 1];
            int cell = prev[j] + 1;
            const int left = cur[j - 1] + 1;
            const int diag = prev[j - 1] + (s_i != t_j);
            if (left < cell) cell = left;
            if (diag < cell) cell = diag;
            if (i > 2 && j > 2) {
                const int trans = prev2[j - 2] + 1 + (a_prev != t_j) + (s_i != bb[j - 2]);
                if (trans < cell) cell = trans;
            }
            cur[j] = (uint16_t)cell;
        }
        a_prev = s_i;
        uint16_t* t = prev2; prev2 = prev; prev = cur; cur = t;
    }
    return (int)prev[m];
}

struct SwResult;
template <class Seq> CB_HD bool sw_similarity_ok(const Seq& a, const uint8_t* bb, const SwResult& r, double similarity);

struct SwResult {
    int ok;                       // 0 = rejected by the similarity test (both strings empty, start = end = 0)
    int start_align, end_align;   // *aStartAlign, *aEndAlign
    uint32_t a_pos, a_len;        // sp.first  = seqA.substr(a_pos, a_len)
    uint32_t b_pos, b_len;        // sp.second = seqB.substr(b_pos, b_len)
};

// a = the read (la bytes), bb = the DR (lb <= kMaxSwDr bytes, local copy); searches a[start, start+len).
// Needs len >= 1, lb >= 1, start + len <= la.
template <class Seq>
CB_HD_NOINLINE void smith_waterman(const Seq& a, uint32_t la, const uint8_t* bb, uint32_t lb, int start, int len,
                                   double similarity, SwResult& r) {
    double row[kMaxSwDr + 1];          // row[j] = M[i-1][j] until cell (i, j) is done, M[i][j] afterwards
    uint32_t org[kMaxSwDr + 1];        // end of the predecessor walk from that cell: (i << 8) | j
    const int m = (int)lb;
    for (int j = 0; j <= m; ++j) { row[j] = 0; org[j] = 0; }
    double matrix_max = -1;
    int i_max = 0, j_max = 0;
    uint32_t o_max = 0;
    for (int i = 1; i <= len; ++i) {
        const uint8_t ai = a[(uint32_t)(i - 1 + start)];
        double diag = 0, left = 0;     // M[i-1][j-1], M[i][j-1]
        uint32_t o_diag = 0, o_left = 0;
        for (int j = 1; j <= m; ++j) {
            const double up = row[j];
            const uint32_t o_up = org[j];
            int index;
            const double v = sw_find_max(d_add(diag, ai == bb[j - 1] ? 1.2 : -1.0), d_add(up, -1.0), d_add(left, -1.0), 0.0, index);
            const uint32_t self = ((uint32_t)i << 8) | (uint32_t)j;
            uint32_t o = self;
            if (index == 0) { if (i > 1 && j > 1) o = o_diag; }
            else if (index == 1) { if (i > 1) o = o_up; }
            else if (index == 2) { if (j > 1) o = o_left; }
            if (v > matrix_max) { matrix_max = v; i_max = i; j_max = j; o_max = o; }
            diag = up; o_diag = o_up;
            left = v; o_left = o;
            row[j] = v; org[j] = o;
        }
    }
    int ci = (int)(o_max >> 8) - 1, cj = (int)(o_max & 255u) - 1;
    if (cj < 0) cj = 0;
    if (ci < 0) ci = 0;
    r.start_align = ci + start;
    r.end_align = r.start_align + i_max - ci - 1;
    // sp.first's LENGTH carries the search offset as well (SmithWaterman.cpp:282); substr clips at the end of the read
    r.a_pos = (uint32_t)(ci + start);
    r.a_len = (uint32_t)(i_max - ci + start);
    if (r.a_len > la - r.a_pos) r.a_len = la - r.a_pos;
    r.b_pos = (uint32_t)cj;
    r.b_len = (uint32_t)(j_max - cj);
    if (r.b_len > lb - r.b_pos) r.b_len = lb - r.b_pos;
    r.ok = 1;
    if (similarity != 0 && !sw_similarity_ok(a, bb, r, similarity)) {
        r.ok = 0; r.start_align = 0; r.end_align = 0;
        r.a_pos = r.a_len = r.b_pos = r.b_len = 0;
    }
}

// the test at the end of smithWaterman (SmithWaterman.cpp:291-303): 1 - distance / len(first string) >= similarity
template <class Seq>
CB_HD bool sw_similarity_ok(const Seq& a, const uint8_t* bb, const SwResult& r, double similarity) {
    const int d = osa_distance_read_vs_dr(a, r.a_pos, r.a_len, bb + r.b_pos, r.b_len);
    const double sim_ld = 1.0 - ((double)d / (double)r.a_len);
    return sim_ld >= similarity;
}

// position of the first (last = false) or last occurrence of nd[0, nl) in h[0, hl), -1 if none; nl >= 1
CB_HD int bytes_find(const uint8_t* h, uint32_t hl, const uint8_t* nd, uint32_t nl, bool last) {
    int r = -1;
    if (nl > hl) return -1;
    for (uint32_t i = 0; i + nl <= hl; ++i) {
        bool eq = true;
        for (uint32_t k = 0; k < nl; ++k) if (h[i + k] != nd[k]) { eq = false; break; }
        if (eq) { r = (int)i; if (!last) break; }
    }
    return r;
}

// One read.  ss_in: its n_in start/stop entries; out: room for n_in + 4.  dr: the consensus DR of the read's group
// (any memory), front_offset: where the read's repeat starts inside it (WorkHorse.cpp:1347).
template <class Seq, class DrSeq>
CB_HD uint8_t update_start_stops(const Seq& s, uint32_t L, const uint32_t* ss_in, uint32_t n_in, int front_offset,
                                 const DrSeq& dr, uint32_t dr_len, uint32_t low_spacer, uint32_t* out, uint32_t& n_out) {
    n_out = 0;
    if (n_in < 2 || (n_in & 1)) return kUssBadList;
    if (dr_len > (uint32_t)kMaxSwDr || dr_len == 0) return kUssDrTooLong;
    if (L >= kMaxSwRead) return kUssReadTooLong;
    const int dr_length = (int)dr_len;
    // ReadHolder.cpp:392-437: shift every repeat, stretch it to the DR's length, clamp the end to the read
    uint32_t first_start = 0, last_end = 0;
    for (uint32_t k = 0; k < n_in; k += 2) {
        int usable = dr_length - 1;
        uint32_t st = ss_in[k];
        if (front_offset >= (int)st) { usable = dr_length - (front_offset - (int)st) - 1; st = 0; }
        else st -= (uint32_t)front_offset;
        if (st >= L) { n_out = 0; return kUssPastRead; }
        uint32_t en = st + (uint32_t)usable;
        if (en >= L) en = L - 1;
        out[2 + k] = st; out[3 + k] = en;
        if (k == 0) first_start = st;
        last_end = en;
    }
    uint8_t bb[kMaxSwDr + 1];
    for (uint32_t j = 0; j < dr_len; ++j) bb[j] = dr[j];
    uint32_t base = 2, n = n_in;
    // The caller's tests on the alignment (:446-451, :498-502) are cheap and nearly always fail for a flank without a
    // partial repeat, while the similarity test inside smithWaterman costs an edit distance.  A failed similarity test
    // zeroes start/end, which fails the caller's tests as well, so the order does not matter: align first
    // (similarity 0 = "just return the substrings"), test, and only then pay for the distance.
    SwResult r;
    if (first_start > low_spacer) {                                   // :443-481: a partial repeat in front of the first one
        smith_waterman(s, L, bb, dr_len, 0, (int)(first_start - low_spacer), 0.0, r);
        if (r.end_align != 0 && r.end_align - r.start_align >= 4) {
            const int at = bytes_find(bb, dr_len, bb + r.b_pos, r.b_len, true);
            if (at >= 0 && (uint32_t)at + r.b_len == dr_len && r.start_align == 0 && sw_similarity_ok(s, bb, r, 0.85)) {
                out[0] = 0; out[1] = (uint32_t)r.end_align;
                base = 0; n += 2;
            }
        }
    }
    const uint32_t end_dist = L - last_end;                            // :483-510: ... and behind the last one
    if (end_dist > low_spacer) {
        smith_waterman(s, L, bb, dr_len, (int)(last_end + low_spacer), (int)(end_dist - low_spacer), 0.0, r);
        if (r.end_align != 0 && r.end_align - r.start_align >= 4) {
            if ((int)L - 1 == r.end_align && bytes_find(bb, dr_len, bb + r.b_pos, r.b_len, false) == 0 && sw_similarity_ok(s, bb, r, 0.85)) {
                int diff = (int)r.a_len - (int)r.b_len;
                if (diff < 0) diff = -diff;
                uint32_t en = (uint32_t)r.end_align;
                if (en >= L) en = L - 1;                               // startStopsAdd clamps the end only
                out[base + n] = (uint32_t)(r.start_align + diff); out[base + n + 1] = en;
                n += 2;
            }
        }
    }
    if (base) for (uint32_t k = 0; k < n; ++k) out[k] = out[k + 2];
    n_out = n;
    return kUssOk;
}

}  // namespace cb
