// dr_core.cuh -- per-read direct-repeat search logic shared by every K1 kernel.
//
// Everything here is `__host__ __device__` and free of warp intrinsics so that the very same
// source can be compiled by g++ into the test-only host simulator (tests/hostsim) and checked
// against the parity oracle without a GPU.  The product never runs this code on the CPU.
//
// Behavioural contract (bit-exact with the reference, /root/reference/src/crass/...):
//   find_left          == PatternMatcher::bmpSearch            PatternMatcher.cpp:26-59
//   scan_right         == scanRight                            libcrispr.cpp:170-263
//   extend_pre_repeat  == extendPreRepeat                      libcrispr.cpp:520-772
//   osa_distance       == PatternMatcher::levenstheinDistance  PatternMatcher.cpp:111-195
//   similarity         == PatternMatcher::getStringSimilarity  PatternMatcher.cpp:197-204
//   low_complexity     == isRepeatLowComplexity                libcrispr.cpp:1031-1069
//   qc_found_repeats   == qcFoundRepeats + testSpacer*         libcrispr.cpp:773-1029
//   search_core        == searchCore                           libcrispr.cpp:265-395
// Unsigned wrap-around of the reference's `unsigned int` arithmetic is reproduced on purpose.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CB_HD __host__ __device__ __forceinline__
#define CB_HD_NOINLINE __host__ __device__ __noinline__
#else
#define CB_HD inline
#define CB_HD_NOINLINE inline
#endif

namespace cb {

struct Params {
    uint32_t low_dr, high_dr, low_spacer, high_spacer, window, min_repeats, kmer_clust, scan_range;
};

// maximum string length the per-thread edit-distance rows can hold; repeat <= high_dr and
// spacers <= high_spacer after the (pure, re-ordered) length tests, so 2*... is never needed.
constexpr int kMaxEdit = 255;

// ---- byte accessors ---------------------------------------------------------------------------
struct PtrSeq {                       // plain pointer (global memory through L1, shared memory, or host)
    const uint8_t* p;
    CB_HD uint8_t operator[](uint32_t i) const { return p[i]; }
};

// ---- float helpers: IEEE single ops in the reference's order, never contracted ------------------
CB_HD float f_add(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
CB_HD float f_sub(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
CB_HD float f_div(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
CB_HD float one_minus(float q) {      // (float)(1.0 - (double)q)
#if defined(__CUDA_ARCH__)
    return __double2float_rn(__dsub_rn(1.0, (double)q));
#else
    return (float)(1.0 - (double)q);
#endif
}

// ---- bmpSearch == leftmost exact occurrence --------------------------------------------------------
// pattern = s[pat .. pat+w); text = s[b .. e).  Returns index relative to b, or -1.
template <class Seq>
CB_HD int find_left(const Seq& s, uint32_t b, uint32_t e, uint32_t pat, uint32_t w) {
    if (e <= b || w == 0) return -1;
    uint32_t tl = e - b;
    if (w > tl) return -1;
    const uint32_t kw = w < 8 ? w : 8;                   // rolling key over the first min(w,8) bytes
    uint64_t key = 0, win = 0;
    for (uint32_t i = 0; i < kw; ++i) key |= (uint64_t)s[pat + i] << (8 * i);
    for (uint32_t i = 0; i + 1 < kw; ++i) win |= (uint64_t)s[b + i] << (8 * (i + 1));
    const uint32_t top = 8 * (kw - 1);
    for (uint32_t p = 0; p + w <= tl; ++p) {
        win = (win >> 8) | ((uint64_t)s[b + p + kw - 1] << top);
        if (win == key) {
            bool ok = true;
            for (uint32_t i = kw; i < w; ++i) if (s[b + p + i] != s[pat + i]) { ok = false; break; }
            if (ok) return (int)p;
        }
    }
    return -1;
}

CB_HD void ss_add(uint32_t* ss, uint32_t& n, uint32_t L, uint32_t i, uint32_t j) {
    // ReadHolder::startStopsAdd (ReadHolder.cpp:263-297): only the end is clamped
    ss[n++] = i;
    if (j >= L) j = L - 1;
    ss[n++] = j;
}

// ---- scanRight ---------------------------------------------------------------------------------------
template <class Seq>
CB_HD void scan_right(const Seq& s, uint32_t L, uint32_t* ss, uint32_t& n_ss, uint32_t cap,
                      uint32_t pat, uint32_t w, uint32_t min_spacer, uint32_t scan_range) {
    uint32_t last = ss[n_ss - 2], second_last = ss[n_ss - 4];
    uint32_t spacing = last - second_last;
    for (;;) {
        uint32_t cand = last + spacing;
        uint32_t begin = cand - scan_range;
        uint32_t end = cand + w + scan_range;
        uint32_t min_begin = last + w + min_spacer;
        if (begin < min_begin) begin = min_begin;
        if (begin > L - 1) break;
        if (end > L) end = L;
        if (begin >= end) break;
        int pos = find_left(s, begin, end, pat, w);
        if (pos < 0) break;
        if (n_ss + 2 > cap) break;                       // cap is sized so that this cannot happen
        ss_add(ss, n_ss, L, begin + (uint32_t)pos, begin + (uint32_t)pos + w - 1);
        second_last = last;
        last = begin + (uint32_t)pos;
        spacing = last - second_last;
        if (spacing < min_spacer + w) break;
    }
}

CB_HD void vote(uint8_t c, int& a, int& cc, int& g, int& t) {
    // only upper-case A/C/G/T vote (libcrispr.cpp:630-644)
    a += (c == 'A'); cc += (c == 'C'); g += (c == 'G'); t += (c == 'T');
}

// ---- extendPreRepeat -----------------------------------------------------------------------------------
template <class Seq>
CB_HD uint32_t extend_pre_repeat(const Seq& s, uint32_t L, uint32_t* ss, uint32_t n_ss, uint32_t window, uint32_t min_spacer) {
    uint32_t num_repeats = n_ss / 2;
    uint32_t rep = window;
    int cut_off = (int)num_repeats - 1;
    if (2 > cut_off) cut_off = 2;
    const uint32_t first = ss[0], last = ss[n_ss - 2];
    uint32_t msp = ss[2] - ss[0];
    for (uint32_t i = 4; i < n_ss; i += 2) {
        uint32_t cur = ss[i] - ss[i - 2];
        if (cur < msp) msp = cur;
    }
    uint32_t right = 0;
    uint32_t max_right = msp - min_spacer;
    uint32_t idx_end = n_ss;
    int cA = 0, cC = 0, cG = 0, cT = 0;
    while (max_right > 0) {
        if (last + window + right >= L) idx_end -= 2;     // cumulative (libcrispr.cpp:614-616)
        for (uint32_t k = 0; k < idx_end && k < n_ss; k += 2) {
            if (ss[k] + rep >= L) break;                  // (:624-627)
            vote(s[ss[k] + rep], cA, cC, cG, cT);
        }
        if (cA >= cut_off || cC >= cut_off || cG >= cut_off || cT >= cut_off) {
            rep++; max_right--; right++; cA = cC = cG = cT = 0;
        } else break;
    }
    cA = cC = cG = cT = 0;
    uint32_t left = 0;
    int test_for_negative = (int)(msp - rep);             // no min_spacer term (:674)
    uint32_t max_left = test_for_negative >= 0 ? (uint32_t)test_for_negative : 0;
    uint32_t idx_start = 0;
    while (left < max_left) {
        if ((int)first - (int)left <= 0) idx_start += 2;  // cumulative (:700-704)
        for (uint32_t k = idx_start; k < n_ss; k += 2) {
            uint32_t at = ss[k] - left - 1;
            if (at < L) vote(s[at], cA, cC, cG, cT);
        }
        if (cA >= cut_off || cC >= cut_off || cG >= cut_off || cT >= cut_off) {
            rep++; left++; cA = cC = cG = cT = 0;
        } else break;
    }
    for (uint32_t k = 0; k + 1 < n_ss; k += 2) {          // (:741-768)
        ss[k] = ss[k] < left ? 0 : ss[k] - left;
        ss[k + 1] = (ss[k + 1] + right >= L) ? L - 1 : ss[k + 1] + right;
    }
    return rep;
}

// ---- modified edit distance (OSA with the i>2 && j>2 transposition guard) ---------------------------------
// a = s[a0 .. a0+n), b = s[b0 .. b0+m); n, m <= kMaxEdit.  Three rolling rows in thread-local memory.
template <class Seq>
CB_HD_NOINLINE int osa_distance(const Seq& s, uint32_t a0, uint32_t n, uint32_t b0, uint32_t m) {
    if (n == 0) return (int)m;
    if (m == 0) return (int)n;
    uint8_t r0[kMaxEdit + 1], r1[kMaxEdit + 1], r2[kMaxEdit + 1], bb[kMaxEdit + 1];
    uint8_t* prev2 = r0; uint8_t* prev = r1; uint8_t* cur = r2;
    for (uint32_t j = 0; j <= m; ++j) prev[j] = (uint8_t)j;
    for (uint32_t j = 0; j < m; ++j) bb[j] = s[b0 + j];
    uint8_t a_prev = 0;
    for (uint32_t i = 1; i <= n; ++i) {
        const uint8_t s_i = s[a0 + i - 1];
        cur[0] = (uint8_t)i;
        for (uint32_t j = 1; j <= m; ++j) {
            const uint8_t t_j = bb[j - 1];
            int cell = prev[j] + 1;
            int left = cur[j - 1] + 1;
            int diag = prev[j - 1] + (s_i != t_j);
            if (left < cell) cell = left;
            if (diag < cell) cell = diag;
            if (i > 2 && j > 2) {
                int trans = prev2[j - 2] + 1 + (a_prev != t_j) + (s_i != bb[j - 2]);
                if (trans < cell) cell = trans;
            }
            cur[j] = (uint8_t)cell;
        }
        a_prev = s_i;
        uint8_t* t = prev2; prev2 = prev; prev = cur; cur = t;
    }
    return (int)prev[m];
}

// Bit-parallel form of the same distance (Hyyro's Damerau bit-vector recurrence, one 64-bit column per text byte)
// for the common case: both strings upper-case A/C/G/T only and the row string at most 64 long.  The reference's
// transposition term is D[i-2][j-2] + 1 + [a(i-1) != b(j)] + [a(i) != b(j-1)], which can only improve a cell when both
// cross comparisons match, i.e. it is the restricted (OSA) transposition; the guard i > 2 && j > 2 clears the
// transposition bits of row 2 and of column 2.  Returns -1 when the fast form does not apply (caller falls back).
// byte -> slot of the match-vector table: A C G T N get their own, anything else is "unsupported"
CB_HD int osa_slot(uint8_t c) {
    return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : c == 'N' ? 4 : -1;
}

// Loop bounds are made warp-uniform on the device (maximum over the lanes that are here together) and the bodies are
// predicated instead: lanes with different string lengths then stay converged through both loops.
CB_HD uint32_t uniform_bound(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __reduce_max_sync(__activemask(), v);
#else
    return v;
#endif
}

template <class Seq>
CB_HD int osa_distance_bitpar(const Seq& s, uint32_t a0, uint32_t n, uint32_t b0, uint32_t m) {
    const bool usable = !(n == 0 || m == 0 || n > 64);
    if (!usable) { n = 0; m = 0; }
    // match vectors, built as two 32-bit halves (a variable 64-bit shift and five 64-bit selects per byte were a fifth of
    // the long-read kernel's instructions)
    uint32_t l0 = 0, l1 = 0, l2 = 0, l3 = 0, l4 = 0, h0 = 0, h1 = 0, h2 = 0, h3 = 0, h4 = 0;
    bool bad = !usable;
    const uint32_t n_loop = uniform_bound(n);
    const uint32_t n_lo = n_loop < 32 ? n_loop : 32;
    for (uint32_t i = 0; i < n_lo; ++i) {
        if (i < n) {
            const int k = osa_slot(s[a0 + i]);
            const uint32_t bit = 1u << i;
            bad |= k < 0;
            l0 |= k == 0 ? bit : 0; l1 |= k == 1 ? bit : 0; l2 |= k == 2 ? bit : 0; l3 |= k == 3 ? bit : 0; l4 |= k == 4 ? bit : 0;
        }
    }
    for (uint32_t i = 32; i < n_loop; ++i) {
        if (i < n) {
            const int k = osa_slot(s[a0 + i]);
            const uint32_t bit = 1u << (i - 32);
            bad |= k < 0;
            h0 |= k == 0 ? bit : 0; h1 |= k == 1 ? bit : 0; h2 |= k == 2 ? bit : 0; h3 |= k == 3 ? bit : 0; h4 |= k == 4 ? bit : 0;
        }
    }
    const uint64_t p0 = l0 | ((uint64_t)h0 << 32), p1 = l1 | ((uint64_t)h1 << 32), p2 = l2 | ((uint64_t)h2 << 32),
                   p3 = l3 | ((uint64_t)h3 << 32), p4 = l4 | ((uint64_t)h4 << 32);
    const uint64_t top = n ? 1ull << (n - 1) : 0;
    uint64_t vp = n >= 64 ? ~0ull : ((1ull << n) - 1ull), vn = 0, d0 = 0, pm_prev = 0;
    int score = (int)n;
    const uint32_t m_loop = uniform_bound(m);
    for (uint32_t j = 0; j < m_loop; ++j) {
        if (j < m) {
            const int k = osa_slot(s[b0 + j]);
            bad |= k < 0;
            const uint64_t pm = k == 0 ? p0 : k == 1 ? p1 : k == 2 ? p2 : k == 3 ? p3 : k == 4 ? p4 : 0;
            uint64_t tr = ((((~d0) & pm) << 1) & pm_prev) & ~3ull;      // rows 1 and 2 never transpose
            if (j < 2) tr = 0;                                          // columns 1 and 2 never transpose
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
    }
    return bad ? -1 : score;
}

template <class Seq>
CB_HD int edit_distance(const Seq& s, uint32_t a0, uint32_t n, uint32_t b0, uint32_t m) {
    const int fast = osa_distance_bitpar(s, a0, n, b0, m);
    return fast >= 0 ? fast : osa_distance(s, a0, n, b0, m);
}

template <class Seq>
CB_HD float similarity(const Seq& s, uint32_t a0, uint32_t n, uint32_t b0, uint32_t m) {
    if (n < 3 || m < 3) return 0.0f;
    float max_length = (float)(n > m ? n : m);
    float d = (float)edit_distance(s, a0, n, b0, m);
    return one_minus(f_div(d, max_length));
}

template <class Seq>
CB_HD bool low_complexity(const Seq& s, uint32_t r0, uint32_t len) {
    int c = 0, g = 0, a = 0, t = 0, n = 0;
    const int cut_off = (int)((double)(int)len * 0.75);
    for (uint32_t i = 0; i < len; ++i) {
        const uint8_t ch = s[r0 + i] & 0xDF;              // fold case for the four letters only
        const uint8_t raw = s[r0 + i];
        const bool letter = (raw >= 'A' && raw <= 'Z') || (raw >= 'a' && raw <= 'z');
        if (letter && ch == 'C') c++;
        else if (letter && ch == 'T') t++;
        else if (letter && ch == 'A') a++;
        else if (letter && ch == 'G') g++;
        else n++;
    }
    return a > cut_off || t > cut_off || g > cut_off || c > cut_off || n > cut_off;
}

CB_HD uint32_t substr_len(uint32_t L, uint32_t pos, uint32_t n) { uint32_t r = L - pos; return n < r ? n : r; }

// internal spacer i of a start/stop list: starts after repeat i, ends before repeat i+1
CB_HD void spacer_at(const uint32_t* ss, uint32_t L, uint32_t i, uint32_t& st, uint32_t& len) {
    st = ss[2 * i + 1] + 1;
    const int raw = (int)(ss[2 * i + 2] - st);
    len = st <= L ? (raw < 0 ? L - st : substr_len(L, st, (uint32_t)raw)) : 0;
}

// ---- qcFoundRepeats -----------------------------------------------------------------------------------------
// All tests are pure, so the cheap ones run first; the result is the conjunction, as in the reference.
// Returns 1 pass, 0 fail, -1 where the reference would throw.
template <class Seq>
CB_HD int qc_found_repeats(const Seq& s, uint32_t L, const uint32_t* ss, uint32_t n_ss, int min_spacer, int max_spacer) {
    const uint32_t n = n_ss / 2;
    if (n < 2) return -1;
    const uint32_t r0 = ss[0];
    const uint32_t rl = substr_len(L, r0, ss[1] - ss[0] + 1);
    if (low_complexity(s, r0, rl)) return 0;
    if (n >= 3) {
        // getAllSpacerStrings (ReadHolder.cpp:199-239) == the n-1 internal spacers; a negative int length
        // handed to substr() becomes a huge count, i.e. "rest of the read" (getNextSpacer :929-933)
        const uint32_t nsp = n - 1;
        int min_len = 10000000, max_len = 0;
        float ssl = 0.0f, rsl = 0.0f;
        uint32_t prev_len = 0;
        for (uint32_t i = 0; i < nsp; ++i) {
            uint32_t st, len;
            spacer_at(ss, L, i, st, len);
            if ((int)len < min_len) min_len = (int)len;
            if ((int)len > max_len) max_len = (int)len;
            if (i > 0) {                                   // pair (i-1, i)
                ssl = f_add(ssl, f_sub((float)prev_len, (float)len));
                rsl = f_add(rsl, f_sub((float)rl, (float)prev_len));
            }
            prev_len = len;
        }
        const float nc = (float)(nsp - 1);
        if (min_len < min_spacer) return 0;
        if (max_len > max_spacer) return 0;
        float a_ssl = f_div(ssl, nc); if (a_ssl < 0) a_ssl = -a_ssl;
        float a_rsl = f_div(rsl, nc); if (a_rsl < 0) a_rsl = -a_rsl;
        if ((int)a_ssl > 12) return 0;
        if ((int)a_rsl > 30) return 0;
        if (rl > (uint32_t)kMaxEdit || max_len > kMaxEdit) return -1;
        float rs = 0.0f, sp = 0.0f;
        uint32_t st0, len0;
        spacer_at(ss, L, 0, st0, len0);
        for (uint32_t i = 0; i + 1 < nsp; ++i) {
            uint32_t st1, len1;
            spacer_at(ss, L, i + 1, st1, len1);
            rs = f_add(rs, similarity(s, r0, rl, st0, len0));
            float ss_diff = 0.0f;
            ss_diff = f_add(ss_diff, similarity(s, st0, len0, st1, len1));
            sp = f_add(sp, ss_diff);
            st0 = st1; len0 = len1;
        }
        sp = f_div(sp, nc);
        rs = f_div(rs, nc);
        if ((double)sp > 0.82) return 0;
        if ((double)rs > 0.82) return 0;
        return 1;
    }
    // two repeats: spacerStringAt(0) is ONE BASE SHORT and its length is computed in unsigned
    // arithmetic (ReadHolder.cpp:102-147)
    const uint32_t st = ss[1] + 1;
    const uint32_t en = ss[2] - 1;
    if (st > L) return -1;
    const uint32_t sl = substr_len(L, st, en - st);
    if ((int)sl < min_spacer) return 0;
    if ((int)sl > max_spacer) return 0;
    int diff = (int)sl - (int)rl; if (diff < 0) diff = -diff;
    if (diff > 30) return 0;
    if (rl > (uint32_t)kMaxEdit || sl > (uint32_t)kMaxEdit) return -1;
    const float sim = similarity(s, r0, rl, st, sl);
    if ((double)sim > 0.82) return 0;
    return 1;
}

// ---- qcFoundRepeats in resumable form ---------------------------------------------------------------------------
// The same function cut at its only expensive step, the edit distance behind getStringSimilarity: qc_start runs every
// test that needs no distance and either settles the result or posts the first similarity job; the caller computes the
// job's distance whenever it suits it (the staged K1 kernel lets the lanes of a warp wait for each other so that the
// bit-parallel recurrence runs with full warps) and hands it to qc_feed, which adds the similarity to the float sums in
// the reference's order and posts the next job or the verdict.  Return values: 1 pass, 0 fail, -1 where the reference
// would throw, 2 = a job is posted in q.job_*.
struct QcRun {
    uint32_t nsp, i, r0, rl, st0, len0, st1, len1;
    uint32_t job_a0, job_n, job_b0, job_m;
    float rs, sp;
    bool is_short, second;                                // second: the posted job is sim(spacer i, spacer i+1)
};

CB_HD int qc_next(const uint32_t* ss, uint32_t L, QcRun& q) {
    while (q.i + 1 < q.nsp) {
        if (!q.second) {                                   // rs += sim(repeat, spacer i)
            if (q.rl >= 3 && q.len0 >= 3) { q.job_a0 = q.r0; q.job_n = q.rl; q.job_b0 = q.st0; q.job_m = q.len0; return 2; }
            q.rs = f_add(q.rs, 0.0f);
            q.second = true;
        } else {                                           // sp += 0 + sim(spacer i, spacer i+1)
            spacer_at(ss, L, q.i + 1, q.st1, q.len1);
            if (q.len0 >= 3 && q.len1 >= 3) { q.job_a0 = q.st0; q.job_n = q.len0; q.job_b0 = q.st1; q.job_m = q.len1; return 2; }
            q.sp = f_add(q.sp, f_add(0.0f, 0.0f));
            q.st0 = q.st1; q.len0 = q.len1; q.i++; q.second = false;
        }
    }
    const float nc = (float)(q.nsp - 1);
    const float sp = f_div(q.sp, nc), rs = f_div(q.rs, nc);
    if ((double)sp > 0.82) return 0;
    if ((double)rs > 0.82) return 0;
    return 1;
}

template <class Seq>
CB_HD int qc_start(const Seq& s, uint32_t L, const uint32_t* ss, uint32_t n_ss, int min_spacer, int max_spacer, QcRun& q) {
    const uint32_t n = n_ss / 2;
    if (n < 2) return -1;
    q.r0 = ss[0];
    q.rl = substr_len(L, q.r0, ss[1] - ss[0] + 1);
    if (low_complexity(s, q.r0, q.rl)) return 0;
    if (n >= 3) {
        q.nsp = n - 1;
        int min_len = 10000000, max_len = 0;
        float ssl = 0.0f, rsl = 0.0f;
        uint32_t prev_len = 0;
        for (uint32_t i = 0; i < q.nsp; ++i) {
            uint32_t st, len;
            spacer_at(ss, L, i, st, len);
            if ((int)len < min_len) min_len = (int)len;
            if ((int)len > max_len) max_len = (int)len;
            if (i > 0) {
                ssl = f_add(ssl, f_sub((float)prev_len, (float)len));
                rsl = f_add(rsl, f_sub((float)q.rl, (float)prev_len));
            }
            prev_len = len;
        }
        const float nc = (float)(q.nsp - 1);
        if (min_len < min_spacer) return 0;
        if (max_len > max_spacer) return 0;
        float a_ssl = f_div(ssl, nc); if (a_ssl < 0) a_ssl = -a_ssl;
        float a_rsl = f_div(rsl, nc); if (a_rsl < 0) a_rsl = -a_rsl;
        if ((int)a_ssl > 12) return 0;
        if ((int)a_rsl > 30) return 0;
        if (q.rl > (uint32_t)kMaxEdit || max_len > kMaxEdit) return -1;
        q.is_short = false; q.second = false; q.i = 0; q.rs = 0.0f; q.sp = 0.0f;
        spacer_at(ss, L, 0, q.st0, q.len0);
        return qc_next(ss, L, q);
    }
    const uint32_t st = ss[1] + 1;
    const uint32_t en = ss[2] - 1;
    if (st > L) return -1;
    const uint32_t sl = substr_len(L, st, en - st);
    if ((int)sl < min_spacer) return 0;
    if ((int)sl > max_spacer) return 0;
    int diff = (int)sl - (int)q.rl; if (diff < 0) diff = -diff;
    if (diff > 30) return 0;
    if (q.rl > (uint32_t)kMaxEdit || sl > (uint32_t)kMaxEdit) return -1;
    if (q.rl < 3 || sl < 3) return 1;                       // similarity 0
    q.is_short = true;
    q.job_a0 = q.r0; q.job_n = q.rl; q.job_b0 = st; q.job_m = sl;
    return 2;
}

CB_HD int qc_feed(const uint32_t* ss, uint32_t L, QcRun& q, int distance) {
    const float max_length = (float)(q.job_n > q.job_m ? q.job_n : q.job_m);
    const float sim = one_minus(f_div((float)distance, max_length));
    if (q.is_short) return (double)sim > 0.82 ? 0 : 1;
    if (!q.second) { q.rs = f_add(q.rs, sim); q.second = true; }
    else { q.sp = f_add(q.sp, f_add(0.0f, sim)); q.st0 = q.st1; q.len0 = q.len1; q.i++; q.second = false; }
    return qc_next(ss, L, q);
}

// ---- candidate handling shared by every K1 variant: a verified seed (j, p) ------------------------------------
// On entry ss is empty.  Returns 1 (array accepted: ss/n_ss/replen hold the result), 0 (rejected; if
// advance is set the window cursor must jump to *next_j = back()-1, libcrispr.cpp:390), <0 error.
template <class Seq>
CB_HD int process_seed(const Seq& s, uint32_t L, const Params& o, uint32_t j, uint32_t p,
                       uint32_t* ss, uint32_t& n_ss, uint32_t cap, uint32_t& replen, bool& advance, uint32_t& next_j) {
    const uint32_t w = o.window;
    n_ss = 0;
    ss_add(ss, n_ss, L, j, j + w - 1);
    ss_add(ss, n_ss, L, p, p + w - 1);
    scan_right(s, L, ss, n_ss, cap, j, w, o.low_spacer, o.scan_range);
    advance = false;
    if (n_ss / 2 >= o.min_repeats) {
        const uint32_t len = extend_pre_repeat(s, L, ss, n_ss, w, o.low_spacer);
        replen = len;
        if (len >= o.low_dr && len <= o.high_dr) {
            const int q = qc_found_repeats(s, L, ss, n_ss, (int)o.low_spacer, (int)o.high_spacer);
            if (q == 1) return 1;
            if (q < 0) return q;
        }
        advance = true;
        next_j = ss[n_ss - 1] - 1;
    }
    n_ss = 0;
    return 0;
}

// process_seed cut at the same place: seed_begin returns 1 / 0 / <0 like process_seed, or 2 with a similarity job
// posted in q; seed_feed takes the job's distance and returns 1 / 0 / <0 / 2 likewise.  On 0, `advance` / `next_j` are
// set as by process_seed.
template <class Seq>
CB_HD int seed_settle(int verdict, uint32_t* ss, uint32_t& n_ss, bool& advance, uint32_t& next_j) {
    if (verdict == 1 || verdict < 0 || verdict == 2) return verdict;
    advance = true;
    next_j = ss[n_ss - 1] - 1;
    n_ss = 0;
    return 0;
}

template <class Seq>
CB_HD int seed_begin(const Seq& s, uint32_t L, const Params& o, uint32_t j, uint32_t p, uint32_t* ss, uint32_t& n_ss,
                     uint32_t cap, uint32_t& replen, bool& advance, uint32_t& next_j, QcRun& q) {
    const uint32_t w = o.window;
    n_ss = 0;
    ss_add(ss, n_ss, L, j, j + w - 1);
    ss_add(ss, n_ss, L, p, p + w - 1);
    scan_right(s, L, ss, n_ss, cap, j, w, o.low_spacer, o.scan_range);
    advance = false;
    if (n_ss / 2 < o.min_repeats) { n_ss = 0; return 0; }
    const uint32_t len = extend_pre_repeat(s, L, ss, n_ss, w, o.low_spacer);
    replen = len;
    int verdict = 0;
    if (len >= o.low_dr && len <= o.high_dr) verdict = qc_start(s, L, ss, n_ss, (int)o.low_spacer, (int)o.high_spacer, q);
    return seed_settle<Seq>(verdict, ss, n_ss, advance, next_j);
}

template <class Seq>
CB_HD int seed_feed(uint32_t L, uint32_t* ss, uint32_t& n_ss, bool& advance, uint32_t& next_j, QcRun& q, int distance) {
    return seed_settle<Seq>(qc_feed(ss, L, q, distance), ss, n_ss, advance, next_j);
}

CB_HD uint32_t window_skips(const Params& o) {
    uint32_t skips = o.low_dr - (2 * o.window - 1);
    if (skips < 1) skips = 1;
    return skips;
}
CB_HD int search_end(const Params& o, uint32_t L) {
    return (int)(L - o.low_dr - o.low_spacer - o.window - 1);
}
CB_HD void window_text(const Params& o, uint32_t L, uint32_t j, uint32_t& begin, uint32_t& end) {
    begin = j + o.low_dr + o.low_spacer;
    end = j + o.high_dr + o.high_spacer + o.window;
    if (end >= L) end = L - 1;                            // the last base is never in the seed text
    if (end < begin) end = begin;
}

// ---- searchCore, one thread per read (generic: any parameters, any length) --------------------------------------
// start_j lets a caller that already knows that no window before start_j has a seed skip them.
template <class Seq>
CB_HD int search_core(const Seq& s, uint32_t L, const Params& o, uint32_t* ss, uint32_t cap, uint32_t& n_ss,
                      uint32_t& replen, uint32_t start_j = 0) {
    n_ss = 0; replen = 0;
    const uint32_t skips = window_skips(o);
    const int se = search_end(o, L);
    if (se < 0) return 0;
    for (uint32_t j = start_j; j <= (uint32_t)se; j = j + skips) {
        uint32_t begin, end;
        window_text(o, L, j, begin, end);
        const int pos = find_left(s, begin, end, j, o.window);
        if (pos >= 0) {
            bool advance; uint32_t nj;
            const int r = process_seed(s, L, o, j, begin + (uint32_t)pos, ss, n_ss, cap, replen, advance, nj);
            if (r != 0) return r;
            if (advance) j = nj;
        }
    }
    n_ss = 0;
    return 0;
}

// capacity (entries) of a start/stop list that scan_right can never overflow for a read of length L
CB_HD uint32_t ss_capacity(const Params& o, uint32_t L) {
    uint32_t unit = o.window + o.low_spacer;
    if (unit == 0) unit = 1;
    return 2 * (L / unit + 4);
}

}  // namespace cb
