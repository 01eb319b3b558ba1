/* crass_oracle.c -- TEST INFRASTRUCTURE ONLY (see crass_oracle.h).
 *
 * Plain-C restatement of the reference's read-scanning hot path.  Each function cites the
 * reference lines it follows.  Integer types mirror the reference's (unsigned wrap-around is
 * part of the behaviour that has to be reproduced).  Parity is pinned against the compiled
 * reference by tests/test_oracle_vs_ref.py and against golden dumps in tests/golden/.
 */
#include "crass_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <time.h>
#include <zlib.h>

/* ------------------------------------------------------------------------------------------ */
/* small containers                                                                            */
/* ------------------------------------------------------------------------------------------ */
typedef struct { char* p; size_t n, cap; } sbuf;

static void sb_reserve(sbuf* b, size_t extra) {
    if (b->n + extra + 1 > b->cap) {
        size_t c = b->cap ? b->cap : 256;
        while (c < b->n + extra + 1) c *= 2;
        b->p = (char*)realloc(b->p, c);
        b->cap = c;
    }
}
static void sb_add(sbuf* b, const void* s, size_t n) { sb_reserve(b, n); memcpy(b->p + b->n, s, n); b->n += n; b->p[b->n] = 0; }
static void sb_adds(sbuf* b, const char* s) { sb_add(b, s, strlen(s)); }
static void sb_addu(sbuf* b, unsigned long long v) { char t[32]; int n = snprintf(t, sizeof t, "%llu", v); sb_add(b, t, (size_t)n); }
static void sb_addi(sbuf* b, long long v) { char t[32]; int n = snprintf(t, sizeof t, "%lld", v); sb_add(b, t, (size_t)n); }
static void sb_addc(sbuf* b, char c) { sb_add(b, &c, 1); }

/* string -> int map, open addressing, keys owned */
typedef struct { char* key; uint32_t len; int val; } smap_ent;
typedef struct { smap_ent* e; size_t cap, n; } smap;

static uint64_t hash_bytes(const void* p, size_t n) {
    const uint8_t* s = (const uint8_t*)p; uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= s[i]; h *= 1099511628211ull; }
    return h;
}
static void smap_init(smap* m) { m->cap = 1024; m->n = 0; m->e = (smap_ent*)calloc(m->cap, sizeof(smap_ent)); }
static void smap_free(smap* m) { for (size_t i = 0; i < m->cap; ++i) free(m->e[i].key); free(m->e); m->e = NULL; }
static smap_ent* smap_slot(smap* m, const void* k, uint32_t len) {
    size_t i = hash_bytes(k, len) & (m->cap - 1);
    while (m->e[i].key && !(m->e[i].len == len && memcmp(m->e[i].key, k, len) == 0)) i = (i + 1) & (m->cap - 1);
    return &m->e[i];
}
static int* smap_find(smap* m, const void* k, uint32_t len) { smap_ent* s = smap_slot(m, k, len); return s->key ? &s->val : NULL; }
static int* smap_put(smap* m, const void* k, uint32_t len, int val) {
    if ((m->n + 1) * 2 > m->cap) {
        smap old = *m; m->cap = old.cap * 2; m->e = (smap_ent*)calloc(m->cap, sizeof(smap_ent));
        for (size_t i = 0; i < old.cap; ++i) if (old.e[i].key) *smap_slot(m, old.e[i].key, old.e[i].len) = old.e[i];
        free(old.e);
    }
    smap_ent* s = smap_slot(m, k, len);
    if (!s->key) { s->key = (char*)malloc(len + 1); memcpy(s->key, k, len); s->key[len] = 0; s->len = len; m->n++; }
    s->val = val;
    return &s->val;
}

/* ------------------------------------------------------------------------------------------ */
/* PatternMatcher                                                                              */
/* ------------------------------------------------------------------------------------------ */
int orc_find_left(const uint8_t* s, uint32_t b, uint32_t e, const uint8_t* pat, uint32_t w) {
    /* bmpSearch (PatternMatcher.cpp:26-59): bad-character Boyer-Moore; returns the leftmost
     * occurrence of pat in text = s[b,e) (index relative to b) or -1; empty text/pattern -> -1. */
    if (e <= b || w == 0) return -1;
    uint32_t tl = e - b;
    if (w > tl) return -1;
    for (uint32_t p = 0; p + w <= tl; ++p)
        if (memcmp(s + b + p, pat, w) == 0) return (int)p;
    return -1;
}

int orc_edit_distance(const uint8_t* a, uint32_t la, const uint8_t* b, uint32_t lb) {
    /* levenstheinDistance (PatternMatcher.cpp:111-195) */
    int n = (int)la, m = (int)lb;
    if (n == 0) return m;
    if (m == 0) return n;
    int* d = (int*)malloc(sizeof(int) * (size_t)(n + 1) * (size_t)(m + 1));
#define D(i, j) d[(size_t)(i) * (size_t)(m + 1) + (size_t)(j)]
    for (int i = 0; i <= n; ++i) D(i, 0) = i;
    for (int j = 0; j <= m; ++j) D(0, j) = j;
    for (int i = 1; i <= n; ++i) {
        uint8_t s_i = a[i - 1];
        for (int j = 1; j <= m; ++j) {
            uint8_t t_j = b[j - 1];
            int cost = (s_i == t_j) ? 0 : 1;
            int above = D(i - 1, j), left = D(i, j - 1), diag = D(i - 1, j - 1);
            int cell = above + 1;
            if (left + 1 < cell) cell = left + 1;
            if (diag + cost < cell) cell = diag + cost;
            if (i > 2 && j > 2) {                          /* :181-186 */
                int trans = D(i - 2, j - 2) + 1;
                if (a[i - 2] != t_j) trans++;
                if (s_i != b[j - 2]) trans++;
                if (cell > trans) cell = trans;
            }
            D(i, j) = cell;
        }
    }
    int r = D(n, m);
#undef D
    free(d);
    return r;
}

float orc_similarity(const uint8_t* a, uint32_t la, const uint8_t* b, uint32_t lb) {
    /* getStringSimilarity (PatternMatcher.cpp:197-204): float division, double subtraction,
     * rounded back to float by the return type. */
    float max_length = (float)(la > lb ? la : lb);
    if (la < 3 || lb < 3) return 0;
    float edit_distance = (float)orc_edit_distance(a, la, b, lb);
    return (float)(1.0 - (double)(edit_distance / max_length));
}

int orc_low_complexity(const uint8_t* r, uint32_t len) {
    /* isRepeatLowComplexity (libcrispr.cpp:1031-1069) */
    int c = 0, g = 0, a = 0, t = 0, n = 0;
    int cut_off = (int)((int)len * 0.75);
    for (uint32_t i = 0; i < len; ++i) {
        switch (r[i]) {
            case 'c': case 'C': c++; break;
            case 't': case 'T': t++; break;
            case 'a': case 'A': a++; break;
            case 'g': case 'G': g++; break;
            default: n++; break;
        }
    }
    return (a > cut_off) || (t > cut_off) || (g > cut_off) || (c > cut_off) || (n > cut_off);
}

/* ------------------------------------------------------------------------------------------ */
/* ReadHolder helpers                                                                          */
/* ------------------------------------------------------------------------------------------ */
static void ss_add(uint32_t* ss, uint32_t* n, uint32_t L, uint32_t i, uint32_t j) {
    /* startStopsAdd (ReadHolder.cpp:263-297): only the END is clamped to L-1 */
    ss[(*n)++] = i;
    if (j >= (uint32_t)(int)L) j = L - 1;
    ss[(*n)++] = j;
}

static const char comp_tab[128] = {
    /* SeqUtils.cpp:51-60 : identity except letters; 'U'->'A', 'u'->'a', '`'(96)->64 */
    0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31,
    32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 63,
    64, 'T', 'V', 'G', 'H', 'E', 'F', 'C', 'D', 'I', 'J', 'M', 'L', 'K', 'N', 'O',
    'P', 'Q', 'Y', 'S', 'A', 'A', 'B', 'W', 'X', 'R', 'Z', 91, 92, 93, 94, 95,
    64, 't', 'v', 'g', 'h', 'e', 'f', 'c', 'd', 'i', 'j', 'm', 'l', 'k', 'n', 'o',
    'p', 'q', 'y', 's', 'a', 'a', 'b', 'w', 'x', 'r', 'z', 123, 124, 125, 126, 127
};

void orc_revcomp(const uint8_t* in, uint32_t len, uint8_t* out) {
    /* reverseComplement (SeqUtils.cpp:62-87); bytes >= 0x80 are undefined in the reference
     * (negative table index) and cannot reach it through kseq (isgraph filter) */
    for (uint32_t i = 0; i < len; ++i) out[len - 1 - i] = (uint8_t)comp_tab[in[i] & 127];
}

static int bytes_less(const uint8_t* a, uint32_t la, const uint8_t* b, uint32_t lb) {
    /* std::string operator< : unsigned-char lexicographic, shorter prefix first */
    uint32_t m = la < lb ? la : lb;
    int c = memcmp(a, b, m);
    if (c) return c < 0;
    return la < lb;
}

int orc_dr_lowlexi(uint8_t* seq, uint32_t L, uint32_t* ss, uint32_t n_ss, uint8_t* dr_out, uint32_t* dr_len, int* was_lowlexi) {
    /* ReadHolder::DRLowLexi (ReadHolder.cpp:513-591) */
    int num_repeats = (int)(n_ss / 2);
    uint32_t idx;
    if (num_repeats == 1) idx = 0;
    else if (num_repeats == 2) {
        if (ss[0] == 0) idx = 2;
        else if (ss[n_ss - 1] == L) idx = 0;
        else {
            int lenA = (int)(ss[1] - ss[0]), lenB = (int)(ss[3] - ss[2]);
            idx = (lenA > lenB) ? 0 : 2;
        }
    } else idx = 2;
    /* repeatStringAt: substr(start, end-start+1), clipped to the string end */
    uint32_t st = ss[idx], ln = ss[idx + 1] - ss[idx] + 1;
    if (st > L) return -1;
    if (ln > L - st) ln = L - st;
    uint8_t* rc = (uint8_t*)malloc(ln + 1);
    orc_revcomp(seq + st, ln, rc);
    if (bytes_less(seq + st, ln, rc, ln)) {
        memcpy(dr_out, seq + st, ln);
        *dr_len = ln; *was_lowlexi = 1;
        free(rc);
        return 0;
    }
    memcpy(dr_out, rc, ln);
    *dr_len = ln; *was_lowlexi = 0;
    free(rc);
    /* reverseComplementSeq (ReadHolder.cpp:593-610) + reverseStartStops (:321-380) */
    uint8_t* tmp = (uint8_t*)malloc(L + 1);
    orc_revcomp(seq, L, tmp);
    memcpy(seq, tmp, L);
    free(tmp);
    uint32_t* t = (uint32_t*)malloc(sizeof(uint32_t) * (n_ss + 1));
    int true_start_offset = (int)L - (int)ss[n_ss - 1] - 1;
    uint32_t prev_pos_fixed = (uint32_t)true_start_offset, prev_pos_orig = ss[n_ss - 1];
    for (uint32_t k = 0; k < n_ss; ++k) {
        uint32_t cur = ss[n_ss - 1 - k];
        uint32_t gap = prev_pos_orig - cur;
        prev_pos_fixed += gap;
        t[k] = prev_pos_fixed;
        prev_pos_orig = cur;
    }
    memcpy(ss, t, sizeof(uint32_t) * n_ss);
    free(t);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* libcrispr: scanRight / extendPreRepeat / qcFoundRepeats / searchCore                        */
/* ------------------------------------------------------------------------------------------ */
void orc_scan_right(const uint8_t* s, uint32_t L, uint32_t* ss, uint32_t* n_ss, uint32_t cap,
                    const uint8_t* pat, uint32_t w, uint32_t min_spacer, uint32_t scan_range) {
    /* scanRight (libcrispr.cpp:170-263) */
    uint32_t n = *n_ss;
    uint32_t last = ss[n - 2], second_last = ss[n - 4];
    uint32_t spacing = last - second_last;
    for (;;) {
        int cand = (int)(last + spacing);
        uint32_t begin = (uint32_t)cand - scan_range;
        uint32_t end = (uint32_t)cand + w + scan_range;
        uint32_t min_begin = last + w + min_spacer;
        if (begin < min_begin) begin = min_begin;
        if (begin > L - 1) break;
        if (end > L) end = L;
        if (begin >= end) break;
        int pos = orc_find_left(s, begin, end, pat, w);
        if (pos < 0) break;
        if (*n_ss + 2 > cap) break;                         /* caller sizes cap >= 2*(L/(w+minSp)+2) */
        ss_add(ss, n_ss, L, begin + (uint32_t)pos, begin + (uint32_t)pos + w - 1);
        second_last = last;
        last = begin + (uint32_t)pos;
        spacing = last - second_last;
        if (spacing < min_spacer + w) break;
    }
}

uint32_t orc_extend_pre_repeat(const uint8_t* s, uint32_t L, uint32_t* ss, uint32_t n_ss, int window, int min_spacer) {
    /* extendPreRepeat (libcrispr.cpp:520-772) */
    uint32_t num_repeats = n_ss / 2;
    uint32_t rep = (uint32_t)window;                       /* RH_RepeatLength */
    int cut_off = (int)num_repeats - 1;
    if (2 > cut_off) cut_off = 2;
    uint32_t first = ss[0], last = ss[n_ss - 2];
    uint32_t msp = ss[2] - ss[0];
    for (uint32_t i = 4; i < n_ss; i += 2) {
        uint32_t cur = ss[i] - ss[i - 2];
        if (cur < msp) msp = cur;
    }
    uint32_t right = 0;
    uint32_t max_right = msp - (uint32_t)min_spacer;
    uint32_t idx_end = n_ss;
    int cA = 0, cC = 0, cG = 0, cT = 0;
    while (max_right > 0) {
        if ((last + (uint32_t)window + right) >= (uint32_t)(int)L) idx_end -= 2;     /* :614-616, cumulative */
        for (uint32_t k = 0; k < idx_end; k += 2) {
            if (k >= n_ss) break;                          /* reference would read out of bounds; unreachable from searchCore */
            if (ss[k] + rep >= L) { k = idx_end; }         /* :624-627 */
            else {
                switch (s[ss[k] + rep]) { case 'A': cA++; break; case 'C': cC++; break; case 'G': cG++; break; case 'T': cT++; break; }
            }
        }
        if (cA >= cut_off || cC >= cut_off || cG >= cut_off || cT >= cut_off) {
            rep++; max_right--; right++; cA = cC = cG = cT = 0;
        } else break;
    }
    cA = cC = cG = cT = 0;
    uint32_t left = 0;
    int test_for_negative = (int)(msp - rep);               /* :674, no minSpacer term */
    uint32_t max_left = (test_for_negative >= 0) ? (uint32_t)test_for_negative : 0;
    uint32_t idx_start = 0;
    while (left < max_left) {
        if ((int)first - (int)left <= 0) idx_start += 2;    /* :700-704, cumulative */
        for (uint32_t k = idx_start; k < n_ss; k += 2) {
            uint32_t at = ss[k] - left - 1;
            if (at < L) {                                   /* reference indexes unchecked */
                switch (s[at]) { case 'A': cA++; break; case 'C': cC++; break; case 'G': cG++; break; case 'T': cT++; break; }
            }
        }
        if (cA >= cut_off || cC >= cut_off || cG >= cut_off || cT >= cut_off) {
            rep++; left++; cA = cC = cG = cT = 0;
        } else break;
    }
    for (uint32_t k = 0; k + 1 < n_ss; k += 2) {            /* :741-768 */
        if (ss[k] < left) ss[k] = 0; else ss[k] -= left;
        if (ss[k + 1] + right >= L) ss[k + 1] = L - 1; else ss[k + 1] += right;
    }
    return rep;
}

/* std::string::substr(pos, n) length for pos <= L */
static uint32_t substr_len(uint32_t L, uint32_t pos, size_t n) { size_t r = (size_t)(L - pos); return (uint32_t)(n < r ? n : r); }

int orc_qc_found_repeats(const uint8_t* s, uint32_t L, const uint32_t* ss, uint32_t n_ss, int min_spacer, int max_spacer) {
    /* qcFoundRepeats (libcrispr.cpp:869-1029) + testSpacer* (:773-867) */
    uint32_t n = n_ss / 2;
    if (n < 2) return -1;                                   /* reference throws */
    /* repeatStringAt(0) (ReadHolder.cpp:77-100) */
    uint32_t r0 = ss[0], rl = substr_len(L, ss[0], (size_t)(ss[1] - ss[0] + 1));
    const uint8_t* repeat = s + r0;
    if (orc_low_complexity(repeat, rl)) return 0;
    int single_compare_index = 0;
    int is_short = (2 > (n - 1));
    if (!is_short) {
        /* getAllSpacerStrings (ReadHolder.cpp:199-239) == the n-1 internal spacers at their true length */
        uint32_t nsp = n - 1;
        uint32_t* sp_st = (uint32_t*)malloc(sizeof(uint32_t) * nsp);
        uint32_t* sp_ln = (uint32_t*)malloc(sizeof(uint32_t) * nsp);
        for (uint32_t i = 0; i < nsp; ++i) {
            uint32_t start_cut = ss[2 * i + 1] + 1;
            int length = (int)(ss[2 * i + 2] - start_cut);  /* getNextSpacer :929-933 */
            sp_st[i] = start_cut;
            sp_ln[i] = start_cut <= L ? substr_len(L, start_cut, (size_t)length) : 0;
        }
        float ave_ss_len = 0.0f, ave_rs_len = 0.0f, ave_ss = 0.0f, ave_rs = 0.0f;
        int min_len = 10000000, max_len = 0, num_compared = 0;
        for (uint32_t i = 0; i + 1 < nsp; ++i) {
            num_compared++;
            ave_rs += orc_similarity(repeat, rl, s + sp_st[i], sp_ln[i]);
            float ss_diff = 0;
            ss_diff += orc_similarity(s + sp_st[i], sp_ln[i], s + sp_st[i + 1], sp_ln[i + 1]);
            ave_ss += ss_diff;
            ave_ss_len += ((float)sp_ln[i] - (float)sp_ln[i + 1]);
            ave_rs_len += ((float)rl - (float)sp_ln[i]);
        }
        for (uint32_t i = 0; i < nsp; ++i) {
            if ((int)sp_ln[i] < min_len) min_len = (int)sp_ln[i];
            if ((int)sp_ln[i] > max_len) max_len = (int)sp_ln[i];
        }
        free(sp_st); free(sp_ln);
        if (num_compared == 0) { is_short = 1; single_compare_index = 1; }
        else {
            ave_ss /= (float)num_compared;
            ave_rs /= (float)num_compared;
            ave_ss_len /= (float)num_compared; if (ave_ss_len < 0) ave_ss_len = -ave_ss_len;
            ave_rs_len /= (float)num_compared; if (ave_rs_len < 0) ave_rs_len = -ave_rs_len;
            if (min_len < min_spacer) return 0;
            if (max_len > max_spacer) return 0;
            if ((double)ave_ss > 0.82) return 0;
            if ((double)ave_rs > 0.82) return 0;
            if ((int)ave_ss_len > 12) return 0;
            if ((int)ave_rs_len > 30) return 0;
        }
    }
    if (is_short) {
        if (single_compare_index % 2 != 0) return -1;       /* spacerStringAt throws on odd index */
        /* spacerStringAt (ReadHolder.cpp:102-147): ONE BASE SHORT, unsigned length */
        uint32_t st = ss[single_compare_index + 1] + 1;
        uint32_t en = ss[single_compare_index + 2] - 1;
        if (st > L) return -1;
        uint32_t sl = substr_len(L, st, (size_t)(uint32_t)(en - st));
        int spacer_len = (int)sl;
        if (spacer_len < min_spacer) return 0;
        if (spacer_len > max_spacer) return 0;
        float sim = orc_similarity(repeat, rl, s + st, sl);
        if ((double)sim > 0.82) return 0;
        int diff = abs((int)sl - (int)rl);
        if (diff > 30) return 0;
    }
    return 1;
}

int orc_search_core(const uint8_t* s, uint32_t L, const orc_params* o, uint32_t* ss, uint32_t ss_cap,
                    uint32_t* n_ss, uint32_t* replen) {
    /* searchCore (libcrispr.cpp:265-395) */
    *n_ss = 0; *replen = 0;
    uint32_t w = o->window;
    uint32_t skips = o->low_dr - (2 * w - 1);
    if (skips < 1) skips = 1;
    int search_end = (int)(L - o->low_dr - o->low_spacer - w - 1);
    if (search_end < 0) return 0;
    if (ss_cap < 8) return -2;
    for (uint32_t j = 0; j <= (uint32_t)search_end; j = j + skips) {
        uint32_t begin = j + o->low_dr + o->low_spacer;
        uint32_t end = j + o->high_dr + o->high_spacer + w;
        if (end >= L) end = L - 1;
        if (end < begin) end = begin;
        int pos = orc_find_left(s, begin, end, s + j, w);
        if (pos >= 0) {
            ss_add(ss, n_ss, L, j, j + w - 1);
            uint32_t found = begin + (uint32_t)pos;
            ss_add(ss, n_ss, L, found, found + w - 1);
            orc_scan_right(s, L, ss, n_ss, ss_cap, s + j, w, o->low_spacer, 24);
        }
        if ((*n_ss / 2) >= o->min_repeats) {
            uint32_t len = orc_extend_pre_repeat(s, L, ss, *n_ss, (int)w, (int)o->low_spacer);
            *replen = len;
            if (len >= o->low_dr && len <= o->high_dr) {
                if (orc_qc_found_repeats(s, L, ss, *n_ss, (int)o->low_spacer, (int)o->high_spacer) == 1) return 1;
            }
            j = ss[*n_ss - 1] - 1;
        }
        *n_ss = 0;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Aho-Corasick, first match only                                                              */
/* ------------------------------------------------------------------------------------------ */
struct orc_ac {
    uint32_t nstates, nsyms;
    uint8_t symv[256];          /* byte -> symbol id, 0 = not in any pattern (acism.c:36-42) */
    int32_t* go;                /* nstates * nsyms, full DFA (failure links resolved) */
    uint32_t* out_len;          /* longest pattern ending at this state (0 = none) */
};

orc_ac* orc_ac_create(const uint8_t* const* pats, const uint32_t* lens, uint32_t n) {
    orc_ac* ac = (orc_ac*)calloc(1, sizeof(orc_ac));
    uint32_t ns = 1;
    for (uint32_t i = 0; i < n; ++i) for (uint32_t k = 0; k < lens[i]; ++k)
        if (!ac->symv[pats[i][k]]) ac->symv[pats[i][k]] = (uint8_t)(ns++);
    ac->nsyms = ns;
    size_t total = 1;
    for (uint32_t i = 0; i < n; ++i) total += lens[i];
    int32_t* go = (int32_t*)malloc(sizeof(int32_t) * total * ns);
    memset(go, 0xff, sizeof(int32_t) * total * ns);
    uint32_t* out = (uint32_t*)calloc(total, sizeof(uint32_t));
    uint32_t* depth = (uint32_t*)calloc(total, sizeof(uint32_t));
    uint32_t nst = 1;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t st = 0;
        for (uint32_t k = 0; k < lens[i]; ++k) {
            uint32_t sy = ac->symv[pats[i][k]];
            if (go[(size_t)st * ns + sy] < 0) { go[(size_t)st * ns + sy] = (int32_t)nst; depth[nst] = depth[st] + 1; nst++; }
            st = (uint32_t)go[(size_t)st * ns + sy];
        }
        if (lens[i] > 0) out[st] = depth[st];
    }
    /* BFS: failure links, output = longest terminal suffix, complete the goto function */
    uint32_t* fail = (uint32_t*)calloc(nst, sizeof(uint32_t));
    uint32_t* queue = (uint32_t*)malloc(sizeof(uint32_t) * nst);
    uint32_t qh = 0, qt = 0;
    for (uint32_t sy = 1; sy < ns; ++sy) {
        int32_t t = go[sy];
        if (t < 0) go[sy] = 0; else { fail[t] = 0; queue[qt++] = (uint32_t)t; }
    }
    go[0] = 0;
    while (qh < qt) {
        uint32_t st = queue[qh++];
        if (!out[st]) out[st] = out[fail[st]];
        go[(size_t)st * ns + 0] = 0;
        for (uint32_t sy = 1; sy < ns; ++sy) {
            int32_t t = go[(size_t)st * ns + sy];
            if (t < 0) go[(size_t)st * ns + sy] = go[(size_t)fail[st] * ns + sy];
            else { fail[t] = (uint32_t)go[(size_t)fail[st] * ns + sy]; queue[qt++] = (uint32_t)t; }
        }
    }
    free(fail); free(queue); free(depth);
    ac->go = go; ac->out_len = out; ac->nstates = nst;
    return ac;
}

int orc_ac_first_match(const orc_ac* ac, const uint8_t* text, uint32_t tlen, int* end, int* plen) {
    /* acism_more with a callback that returns 1 (acism.c:26-104; libcrispr.cpp:441): the first
     * reported match is the one with the smallest end offset, longest pattern on ties. */
    uint32_t st = 0, ns = ac->nsyms;
    for (uint32_t i = 0; i < tlen; ++i) {
        uint32_t sy = ac->symv[text[i]];
        if (!sy) { st = 0; continue; }
        st = (uint32_t)ac->go[(size_t)st * ns + sy];
        if (ac->out_len[st]) { *end = (int)(i + 1); *plen = (int)ac->out_len[st]; return 1; }
    }
    return 0;
}

void orc_ac_destroy(orc_ac* ac) { if (ac) { free(ac->go); free(ac->out_len); free(ac); } }

/* ------------------------------------------------------------------------------------------ */
/* kseq                                                                                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct { char* s; size_t l, m; } kstr;      /* s == NULL until first written, as kstring_t */
typedef struct {
    uint8_t* buf; size_t n, pos;
    kstr name, comment, seq, qual;
    int last_char;
    int is_eof;                                       /* kstream's flag: a read of its 4096-byte buffer came back short */
    int read_error, stale_done;                       /* the input ended with a FAILED read (a damaged archive), see kp_getc */
} kparser;

/* kstream refills a 4096-byte buffer and learns about the end of the input from a SHORT read: when the input's size is a
 * multiple of 4096 the flag is still clear after the last byte has been consumed, and the first ks_getuntil called there
 * returns an empty string instead of -1 (kseq.cpp:71-92) -- one extra record with an empty name and an empty sequence when
 * the last byte is a header character.  With the whole input in memory that state is one bit. */
/* A failed gzread (-1, a damaged archive) is taken for a short read by kstream: end = -1 sets is_eof, but only end == 0 makes
 * ks_getc return -1 (kseq.cpp:55-69) -- so the ks_getc that meets the failure hands out buf[0] once more before the stream ends:
 * the first byte of the last good 4096-byte chunk, or of the failing one if zlib had copied part of it before it met the damage
 * (kp_open leaves that byte at buf[n]); ks_getuntil (kseq.cpp:84-96) copies nothing in that state and leaves begin = 1. */
static int kp_getc(kparser* k) {
    if (k->pos < k->n) return (int)(signed char)k->buf[k->pos++];
    if (k->read_error && !k->stale_done && !k->is_eof && k->n >= 4096) { k->stale_done = 1; k->is_eof = 1; return (int)(signed char)k->buf[k->n]; }
    k->is_eof = 1;
    return -1;
}

static void ks_put(kstr* s, size_t need) { if (need > s->m) { s->m = need * 2 + 16; s->s = (char*)realloc(s->s, s->m); } }

/* ks_getuntil (kseq.cpp:71-147): delimiter 0 = isspace, otherwise a literal byte */
static int kp_getuntil(kparser* k, int delimiter, kstr* str, int* dret) {
    if (dret) *dret = 0;
    str->l = 0;
    if (k->pos >= k->n) {
        if (k->is_eof) return -1;
        k->is_eof = 1;                                /* the refill that finds nothing happens inside this call */
        k->stale_done = 1;
        ks_put(str, 1);
        str->s[0] = 0;
        return 0;
    }
    size_t i = k->pos;
    if (delimiter > 1) { while (i < k->n && k->buf[i] != (uint8_t)delimiter) ++i; }
    else { while (i < k->n && !isspace(k->buf[i])) ++i; }
    ks_put(str, i - k->pos + 1);
    memcpy(str->s, k->buf + k->pos, i - k->pos);
    str->l = i - k->pos;
    if (i < k->n) { if (dret) *dret = (int)(signed char)k->buf[i]; k->pos = i + 1; } else { k->pos = i; k->is_eof = 1; k->stale_done = 1; }
    str->s[str->l] = 0;
    return (int)str->l;
}

/* kseq_read (kseq.cpp:171-225) */
static int kp_read(kparser* k) {
    int c;
    if (k->last_char == 0) {
        while ((c = kp_getc(k)) != -1 && c != '>' && c != '@') {}
        if (c == -1) return -1;
        k->last_char = c;
    }
    k->comment.l = k->seq.l = k->qual.l = 0;
    if (kp_getuntil(k, 0, &k->name, &c) < 0) return -1;
    if (c != '\n') kp_getuntil(k, '\n', &k->comment, 0);
    while ((c = kp_getc(k)) != -1 && c != '>' && c != '+' && c != '@') {
        if (isgraph(c)) { ks_put(&k->seq, k->seq.l + 2); k->seq.s[k->seq.l++] = (char)c; }
    }
    if (c == '>' || c == '@') k->last_char = c;
    ks_put(&k->seq, k->seq.l + 2);
    k->seq.s[k->seq.l] = 0;
    if (c != '+') return (int)k->seq.l;
    ks_put(&k->qual, k->seq.l + 2);
    while ((c = kp_getc(k)) != -1 && c != '\n') {}
    if (c == -1) return -2;
    while ((c = kp_getc(k)) != -1 && k->qual.l < k->seq.l) {
        if (c >= 33 && c <= 127) k->qual.s[k->qual.l++] = (char)c;
    }
    k->qual.s[k->qual.l] = 0;
    k->last_char = 0;
    if (k->seq.l != k->qual.l) return -2;
    return (int)k->seq.l;
}

static int kp_open(kparser* k, const char* path) {
    memset(k, 0, sizeof *k);
    gzFile fp = gzopen(path, "r");
    if (!fp) return -1;
    size_t cap = 1 << 20;
    k->buf = (uint8_t*)malloc(cap);
    for (;;) {                                        /* 4096 bytes per call, as kstream reads (kseq.cpp:44,60-66): what a damaged archive
                                                         yields depends on it (zlib drops what a failing call had inflated) */
        if (cap - k->n < (1 << 19)) { cap *= 2; k->buf = (uint8_t*)realloc(k->buf, cap); }
        if (k->n >= 4096) k->buf[k->n] = k->buf[k->n - 4096];   /* kstream has ONE buffer: a failing call leaves its first byte or overwrites it */
        int r = gzread(fp, k->buf + k->n, 4096);
        if (r < 0) { k->read_error = 1; break; }
        k->n += (size_t)r;
        if (r < 4096) break;                          /* is_eof: kstream does not read again */
    }
    gzclose(fp);
    k->is_eof = k->n % 4096 != 0;                     /* the short read has happened by the time the last byte is consumed */
    return 0;
}
static void kp_close(kparser* k) { free(k->buf); free(k->name.s); free(k->comment.s); free(k->seq.s); free(k->qual.s); }

char* orc_kseq_dump(const char* path) {
    kparser k;
    if (kp_open(&k, path)) return NULL;
    sbuf o = {0, 0, 0};
    int l;
    while ((l = kp_read(&k)) >= 0) {
        sb_adds(&o, k.name.s); sb_addc(&o, '\t');
        sb_adds(&o, k.comment.s ? k.comment.s : "\x01"); sb_addc(&o, '\t');
        sb_adds(&o, k.seq.s); sb_addc(&o, '\t');
        sb_adds(&o, k.qual.s ? k.qual.s : "\x01"); sb_addc(&o, '\n');
    }
    sb_adds(&o, "#ret="); sb_addi(&o, l); sb_addc(&o, '\n');
    kp_close(&k);
    return o.p;
}

/* ------------------------------------------------------------------------------------------ */
/* read store (ReadMap / StringCheck / lookupTable restated on flat arrays)                    */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int token, phase, lowlexi, is_fasta;
    uint32_t replen, n_ss, L, qual_hash;
    uint32_t* ss;
    char *seq, *header, *comment;
    uint64_t order;             /* insertion order, to keep ReadList order inside a token */
} held_read;

typedef struct {
    held_read* r; size_t n, cap;
    smap s2t;                   /* StringCheck::mS2T_map */
    char** t2s; uint32_t* t2s_len; int next_token;   /* mT2S_map; first token is 2 (StringCheck.cpp:51) */
    size_t t2s_cap;
    smap patterns_hash, reads_found;
} store;

static void store_init(store* st) {
    memset(st, 0, sizeof *st);
    smap_init(&st->s2t); smap_init(&st->patterns_hash); smap_init(&st->reads_found);
    st->next_token = 1;
}

static uint32_t fnv1a32(const char* s) { uint32_t h = 2166136261u; for (; s && *s; ++s) { h ^= (uint8_t)*s; h *= 16777619u; } return h; }

static char* dupn(const char* s, size_t n) { char* p = (char*)malloc(n + 1); memcpy(p, s, n); p[n] = 0; return p; }

/* addReadHolder (libcrispr.cpp:1119-1162) */
static void store_add(store* st, const kparser* k, const uint32_t* ss, uint32_t n_ss, uint32_t replen, int phase) {
    if (st->n == st->cap) { st->cap = st->cap ? st->cap * 2 : 1024; st->r = (held_read*)realloc(st->r, st->cap * sizeof(held_read)); }
    held_read* h = &st->r[st->n];
    memset(h, 0, sizeof *h);
    h->L = (uint32_t)k->seq.l;
    h->seq = dupn(k->seq.s, k->seq.l);
    h->header = dupn(k->name.s, k->name.l);
    h->comment = k->comment.s ? dupn(k->comment.s, strlen(k->comment.s)) : dupn("", 0);
    h->is_fasta = k->qual.s ? 0 : 1;                     /* setQual clears RH_IsFasta (ReadHolder.h:272-276) */
    h->qual_hash = fnv1a32(k->qual.s ? k->qual.s : "");
    h->ss = (uint32_t*)malloc(sizeof(uint32_t) * (n_ss + 2));
    memcpy(h->ss, ss, sizeof(uint32_t) * n_ss);
    h->n_ss = n_ss; h->replen = replen; h->phase = phase; h->order = st->n;
    uint8_t* dr = (uint8_t*)malloc(h->L + 2); uint32_t dr_len = 0;
    orc_dr_lowlexi((uint8_t*)h->seq, h->L, h->ss, h->n_ss, dr, &dr_len, &h->lowlexi);
    int* tok = smap_find(&st->s2t, dr, dr_len);
    if (!tok) {
        st->next_token++;
        if ((size_t)st->next_token >= st->t2s_cap) {
            st->t2s_cap = st->t2s_cap ? st->t2s_cap * 2 : 1024;
            st->t2s = (char**)realloc(st->t2s, st->t2s_cap * sizeof(char*));
            st->t2s_len = (uint32_t*)realloc(st->t2s_len, st->t2s_cap * sizeof(uint32_t));
        }
        st->t2s[st->next_token] = dupn((char*)dr, dr_len);
        st->t2s_len[st->next_token] = dr_len;
        tok = smap_put(&st->s2t, dr, dr_len, st->next_token);
    }
    h->token = *tok;
    free(dr);
    st->n++;
}

static void store_free(store* st) {
    for (size_t i = 0; i < st->n; ++i) { free(st->r[i].ss); free(st->r[i].seq); free(st->r[i].header); free(st->r[i].comment); }
    free(st->r);
    for (int t = 2; t <= st->next_token; ++t) free(st->t2s[t]);
    free(st->t2s); free(st->t2s_len);
    smap_free(&st->s2t); smap_free(&st->patterns_hash); smap_free(&st->reads_found);
}

/* ------------------------------------------------------------------------------------------ */
/* createNonRedundantSet                                                                       */
/* ------------------------------------------------------------------------------------------ */
typedef struct { char* s; uint32_t len; } str_t;
typedef struct { int gid; int* tokens; size_t n, cap; } group_t;

static void laurenize11(const uint8_t* km, uint8_t* out) {
    /* laurenize (SeqUtils.cpp:89-97): min(seq, revcomp(seq)); ties -> revcomp (identical anyway) */
    uint8_t rc[11];
    orc_revcomp(km, 11, rc);
    if (bytes_less(km, 11, rc, 11)) memcpy(out, km, 11); else memcpy(out, rc, 11);
}

static int contains(const char* hay, uint32_t hl, const char* needle, uint32_t nl) {
    if (nl > hl) return 0;
    if (nl == 0) return 1;
    for (uint32_t i = 0; i + nl <= hl; ++i) if (memcmp(hay + i, needle, nl) == 0) return 1;
    return 0;
}

static int cmp_len(const void* a, const void* b) {
    const str_t* x = (const str_t*)a; const str_t* y = (const str_t*)b;
    return (x->len > y->len) - (x->len < y->len);
}

/* drs[i] is the string of token i+2.  Emits "G token gid" lines (gid order, cluster order) and the
 * non-redundant patterns (survivors then their reverse complements per group), as the reference
 * restatement in refshim does; survivors inside one group may come out in a different order than
 * std::sort leaves them (unstable sort on equal lengths) -- consumers compare them as sorted sets. */
static void non_redundant(const uint8_t* const* drs, const uint32_t* lens, uint32_t n, int min_count,
                          sbuf* glines, str_t** out_pats, size_t* out_n) {
    const uint32_t K = 11;                                  /* crassDefines.h:66 */
    smap k2gid; smap_init(&k2gid);
    group_t* groups = NULL; size_t ng = 0, capg = 0;
    int next_gid = 1;
    for (uint32_t t = 0; t < n; ++t) {                      /* clusterDRReads (WorkHorse.cpp:1404-1637) in token order */
        int num_mers = (int)lens[t] - (int)K + 1;
        uint8_t (*homeless)[11] = (uint8_t(*)[11])malloc((size_t)(num_mers > 0 ? num_mers : 1) * 11);
        int nh = 0, group = 0;
        int* gc_gid = (int*)malloc(sizeof(int) * (size_t)(num_mers > 0 ? num_mers : 1));
        int* gc_cnt = (int*)malloc(sizeof(int) * (size_t)(num_mers > 0 ? num_mers : 1));
        int ngc = 0;
        for (int i = 0; i < num_mers; ++i) {
            uint8_t km[11];
            laurenize11(drs[t] + i, km);
            int* g = smap_find(&k2gid, km, K);
            if (!g) { memcpy(homeless[nh++], km, K); }
            else if (0 == group) {
                int k; for (k = 0; k < ngc; ++k) if (gc_gid[k] == *g) break;
                if (k == ngc) { gc_gid[ngc] = *g; gc_cnt[ngc] = 1; ngc++; }
                else { gc_cnt[k]++; if (min_count <= gc_cnt[k]) group = *g; }
            }
        }
        if (0 == group) {
            group = next_gid++;
            if (ng == capg) { capg = capg ? capg * 2 : 64; groups = (group_t*)realloc(groups, capg * sizeof(group_t)); }
            groups[ng].gid = group; groups[ng].tokens = NULL; groups[ng].n = groups[ng].cap = 0; ng++;
        }
        group_t* G = &groups[group - 1];                    /* gids are dense, 1-based, creation order */
        if (G->n == G->cap) { G->cap = G->cap ? G->cap * 2 : 8; G->tokens = (int*)realloc(G->tokens, G->cap * sizeof(int)); }
        G->tokens[G->n++] = (int)t + 2;
        for (int i = 0; i < nh; ++i) smap_put(&k2gid, homeless[i], K, group);
        free(homeless); free(gc_gid); free(gc_cnt);
    }
    str_t* pats = NULL; size_t np = 0, capp = 0;
    for (size_t g = 0; g < ng; ++g) {                       /* createNonRedundantSet (WorkHorse.cpp:648-709) */
        group_t* G = &groups[g];
        str_t* v = (str_t*)malloc(sizeof(str_t) * G->n);
        for (size_t i = 0; i < G->n; ++i) {
            if (glines) { sb_adds(glines, "G\t"); sb_addi(glines, G->tokens[i]); sb_addc(glines, '\t'); sb_addi(glines, G->gid); sb_addc(glines, '\n'); }
            v[i].s = (char*)drs[G->tokens[i] - 2]; v[i].len = lens[G->tokens[i] - 2];
        }
        qsort(v, G->n, sizeof(str_t), cmp_len);             /* removeRedundantRepeats (:612-645) */
        uint8_t* dead = (uint8_t*)calloc(G->n ? G->n : 1, 1);
        for (size_t i = 0; i < G->n; ++i) {
            if (dead[i] || v[i].len == 0) continue;
            uint8_t* rc = (uint8_t*)malloc(v[i].len + 1);
            orc_revcomp((const uint8_t*)v[i].s, v[i].len, rc);
            for (size_t j = i + 1; j < G->n; ++j) {
                if (dead[j] || v[j].len == 0) continue;
                if (contains(v[j].s, v[j].len, v[i].s, v[i].len) || contains(v[j].s, v[j].len, (char*)rc, v[i].len)) dead[j] = 1;
            }
            free(rc);
        }
        size_t first = np;
        for (size_t i = 0; i < G->n; ++i) {
            if (dead[i] || v[i].len == 0) continue;
            if (np == capp) { capp = capp ? capp * 2 : 64; pats = (str_t*)realloc(pats, capp * sizeof(str_t)); }
            pats[np].s = dupn(v[i].s, v[i].len); pats[np].len = v[i].len; np++;
        }
        size_t last = np;
        for (size_t i = first; i < last; ++i) {
            if (np == capp) { capp = capp ? capp * 2 : 64; pats = (str_t*)realloc(pats, capp * sizeof(str_t)); }
            pats[np].s = (char*)malloc(pats[i].len + 1); pats[np].len = pats[i].len;
            orc_revcomp((const uint8_t*)pats[i].s, pats[i].len, (uint8_t*)pats[np].s);
            pats[np].s[pats[np].len] = 0; np++;
        }
        free(dead); free(v); free(G->tokens);
    }
    free(groups);
    smap_free(&k2gid);
    *out_pats = pats; *out_n = np;
}

char* orc_non_redundant(const uint8_t* const* drs, const uint32_t* lens, uint32_t n, int min_count) {
    sbuf o = {0, 0, 0};
    str_t* pats; size_t np;
    sb_adds(&o, "");
    non_redundant(drs, lens, n, min_count, &o, &pats, &np);
    for (size_t i = 0; i < np; ++i) { sb_adds(&o, "P\t"); sb_add(&o, pats[i].s, pats[i].len); sb_addc(&o, '\n'); free(pats[i].s); }
    free(pats);
    return o.p;
}

/* ------------------------------------------------------------------------------------------ */
/* whole path                                                                                  */
/* ------------------------------------------------------------------------------------------ */
static int cmp_held(const void* a, const void* b) {
    const held_read* x = (const held_read*)a; const held_read* y = (const held_read*)b;
    if (x->token != y->token) return x->token < y->token ? -1 : 1;
    return (x->order > y->order) - (x->order < y->order);
}
static int cmp_strp(const void* a, const void* b) {
    const str_t* x = (const str_t*)a; const str_t* y = (const str_t*)b;
    uint32_t m = x->len < y->len ? x->len : y->len;
    int c = memcmp(x->s, y->s, m);
    if (c) return c;
    return (x->len > y->len) - (x->len < y->len);
}

static double now_ms(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec / 1e6; }

char* orc_run_files(const char* const* paths, uint32_t n_paths, const orc_params* p, int phases, double* timings_ms) {
    store st; store_init(&st);
    sbuf o = {0, 0, 0};
    int max_len = 0;
    uint32_t ss_cap = 4096; uint32_t* ss = (uint32_t*)malloc(sizeof(uint32_t) * ss_cap);
    double t0 = now_ms();
    for (uint32_t f = 0; f < n_paths; ++f) {                /* searchFile (libcrispr.cpp:68-166) */
        kparser k;
        if (kp_open(&k, paths[f])) { sb_adds(&o, "E\tcannot open\n"); return o.p; }
        int l;
        while ((l = kp_read(&k)) >= 0) {
            if (l > max_len) max_len = l;
            uint32_t need = 2 * ((uint32_t)l / 4 + 4);
            if (need > ss_cap) { ss_cap = need * 2; ss = (uint32_t*)realloc(ss, sizeof(uint32_t) * ss_cap); }
            uint32_t n_ss = 0, replen = 0;
            if (orc_search_core((const uint8_t*)k.seq.s, (uint32_t)l, p, ss, ss_cap, &n_ss, &replen) == 1) {
                /* patternsHash[repeatStringAt(0)] is taken AFTER addReadHolder copied the holder, i.e. on
                 * the un-flipped temporary (libcrispr.cpp:136-137) */
                uint32_t r0 = ss[0], rl = substr_len((uint32_t)l, ss[0], (size_t)(ss[1] - ss[0] + 1));
                smap_put(&st.patterns_hash, k.seq.s + r0, rl, 1);
                store_add(&st, &k, ss, n_ss, replen, 1);
                smap_put(&st.reads_found, k.name.s, (uint32_t)k.name.l, 1);
            }
        }
        kp_close(&k);
    }
    double t1 = now_ms();
    size_t n_found_p1 = st.reads_found.n;
    /* createNonRedundantSet */
    uint32_t ntok = (uint32_t)(st.next_token - 1);
    sbuf glines = {0, 0, 0}; sb_adds(&glines, "");
    str_t* pats = NULL; size_t np = 0;
    non_redundant((const uint8_t* const*)(st.t2s + 2), st.t2s_len + 2, ntok, (int)p->kmer_clust, &glines, &pats, &np);
    double t2 = now_ms();
    if (phases >= 2 && np > 0) {                            /* findSingletons (libcrispr.cpp:444-518) */
        const uint8_t** pp = (const uint8_t**)malloc(sizeof(uint8_t*) * np);
        uint32_t* pl = (uint32_t*)malloc(sizeof(uint32_t) * np);
        for (size_t i = 0; i < np; ++i) { pp[i] = (const uint8_t*)pats[i].s; pl[i] = pats[i].len; }
        orc_ac* ac = orc_ac_create(pp, pl, (uint32_t)np);
        for (uint32_t f = 0; f < n_paths; ++f) {
            kparser k;
            if (kp_open(&k, paths[f])) continue;
            int l;
            while ((l = kp_read(&k)) >= 0) {
                int end, plen;
                if (!orc_ac_first_match(ac, (const uint8_t*)k.seq.s, (uint32_t)l, &end, &plen)) continue;
                if (smap_find(&st.reads_found, k.name.s, (uint32_t)k.name.l)) continue;      /* on_match :411 */
                uint32_t dr_end = (uint32_t)(end - 1);
                if (dr_end >= (uint32_t)l) dr_end = (uint32_t)l - 1;
                uint32_t one[2]; uint32_t n1 = 0;
                ss_add(one, &n1, (uint32_t)l, dr_end - ((uint32_t)plen - 1), dr_end);
                store_add(&st, &k, one, 2, 0, 2);
            }
            kp_close(&k);
        }
        orc_ac_destroy(ac); free(pp); free(pl);
    }
    double t3 = now_ms();
    if (timings_ms) { timings_ms[0] = t1 - t0; timings_ms[1] = t2 - t1; timings_ms[2] = t3 - t2; }

    sb_adds(&o, "# crass-dump v1\n");
    sb_adds(&o, "M\t"); sb_addi(&o, max_len); sb_addc(&o, '\t'); sb_addu(&o, n_found_p1); sb_addc(&o, '\t'); sb_addu(&o, st.patterns_hash.n); sb_addc(&o, '\n');
    sb_add(&o, glines.p, glines.n); free(glines.p);
    qsort(pats, np, sizeof(str_t), cmp_strp);
    for (size_t i = 0; i < np; ++i) { sb_adds(&o, "P\t"); sb_add(&o, pats[i].s, pats[i].len); sb_addc(&o, '\n'); }
    {
        str_t* hk = (str_t*)malloc(sizeof(str_t) * (st.patterns_hash.n + 1)); size_t nh = 0;
        for (size_t i = 0; i < st.patterns_hash.cap; ++i) if (st.patterns_hash.e[i].key) { hk[nh].s = st.patterns_hash.e[i].key; hk[nh].len = st.patterns_hash.e[i].len; nh++; }
        qsort(hk, nh, sizeof(str_t), cmp_strp);
        for (size_t i = 0; i < nh; ++i) { sb_adds(&o, "H\t"); sb_add(&o, hk[i].s, hk[i].len); sb_addc(&o, '\n'); }
        free(hk);
    }
    qsort(st.r, st.n, sizeof(held_read), cmp_held);
    size_t i = 0;
    while (i < st.n) {
        size_t j = i; while (j < st.n && st.r[j].token == st.r[i].token) ++j;
        int t = st.r[i].token;
        sb_adds(&o, "T\t"); sb_addi(&o, t); sb_addc(&o, '\t'); sb_add(&o, st.t2s[t], st.t2s_len[t]); sb_addc(&o, '\t'); sb_addu(&o, j - i); sb_addc(&o, '\n');
        for (size_t k = i; k < j; ++k) {
            held_read* h = &st.r[k];
            sb_adds(&o, "R\t"); sb_addi(&o, t); sb_addc(&o, '\t'); sb_addi(&o, h->phase); sb_addc(&o, '\t'); sb_adds(&o, h->header); sb_addc(&o, '\t');
            sb_addi(&o, h->lowlexi); sb_addc(&o, '\t'); sb_addu(&o, h->replen); sb_addc(&o, '\t');
            for (uint32_t q = 0; q < h->n_ss; ++q) { if (q) sb_addc(&o, ','); sb_addu(&o, h->ss[q]); }
            sb_addc(&o, '\t'); sb_add(&o, h->seq, h->L); sb_addc(&o, '\t'); sb_adds(&o, h->comment); sb_addc(&o, '\t');
            sb_addi(&o, h->is_fasta); sb_addc(&o, '\t'); sb_addu(&o, h->qual_hash); sb_addc(&o, '\n');
        }
        i = j;
    }
    for (size_t q = 0; q < np; ++q) free(pats[q].s);
    free(pats); free(ss);
    store_free(&st);
    return o.p;
}

uint64_t orc_phase1_batch(const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads, const orc_params* p, uint8_t* found) {
    uint32_t ss_cap = 8192; uint32_t* ss = (uint32_t*)malloc(sizeof(uint32_t) * ss_cap);
    uint64_t nf = 0;
    for (uint32_t i = 0; i < n_reads; ++i) {
        uint32_t L = (uint32_t)(offsets[i + 1] - offsets[i]);
        uint32_t need = 2 * (L / 4 + 4);
        if (need > ss_cap) { ss_cap = need * 2; ss = (uint32_t*)realloc(ss, sizeof(uint32_t) * ss_cap); }
        uint32_t n_ss, replen;
        int f = orc_search_core(bases + offsets[i], L, p, ss, ss_cap, &n_ss, &replen) == 1;
        if (found) found[i] = (uint8_t)f;
        nf += (uint64_t)f;
    }
    free(ss);
    return nf;
}

uint64_t orc_phase2_batch(const orc_ac* ac, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads, uint8_t* found) {
    uint64_t nf = 0;
    for (uint32_t i = 0; i < n_reads; ++i) {
        int end, plen;
        int f = orc_ac_first_match(ac, bases + offsets[i], (uint32_t)(offsets[i + 1] - offsets[i]), &end, &plen);
        if (found) found[i] = (uint8_t)f;
        nf += (uint64_t)f;
    }
    return nf;
}

/* ------------------------------------------------------------------------------------------ */
/* Partial-DR recovery: smithWaterman + ReadHolder::updateStartStops          */
/* ------------------------------------------------------------------------------------------ */
static double sw_find_max(double a, double b, double c, double d, int* index) {
    /* findMax (SmithWaterman.cpp:68-131): ties go to a, then b, then ... exactly as the nested tests do */
    if (b > a) {
        if (c > d) { if (c > b) { *index = 2; return c; } *index = 1; return b; }
        if (d > b) { *index = 3; return d; }
        *index = 1; return b;
    }
    if (c > d) { if (c > a) { *index = 2; return c; } *index = 0; return a; }
    if (d > a) { *index = 3; return d; }
    *index = 0; return a;
}

int orc_smith_waterman(const uint8_t* a, uint32_t la, const uint8_t* b, uint32_t lb, int a_start_search, int a_search_len,
                       double similarity, int* a_start_align, int* a_end_align,
                       uint32_t* a_ret_pos, uint32_t* a_ret_len, uint32_t* b_ret_pos, uint32_t* b_ret_len) {
    /* smithWaterman (SmithWaterman.cpp:151-308).  Scores are doubles (match 1.2, mismatch -1, gap -1,
     * SmithWaterman.h:60-63) and are added in the reference's order, so equalities between cells are the reference's.
     * Returns 1 with the two substrings (as positions/lengths in a and b), 0 when the similarity test rejects.
     * Needs a_search_len >= 1, lb >= 1 and a_start_search + a_search_len <= la (the reference reads out of bounds
     * otherwise). */
    const int n = a_search_len, m = (int)lb;
    const size_t w = (size_t)m + 1;
    double* M = (double*)calloc((size_t)(n + 1) * w, sizeof(double));
    int* Ii = (int*)calloc((size_t)(n + 1) * w, sizeof(int));
    int* Ij = (int*)calloc((size_t)(n + 1) * w, sizeof(int));
    double matrix_max = -1;
    int i_max = 0, j_max = 0;
    for (int i = 1; i <= n; ++i) {
        for (int j = 1; j <= m; ++j) {
            int index = -1;
            const double sim = (a[i - 1 + a_start_search] == b[j - 1]) ? 1.2 : -1;
            const double v = sw_find_max(M[(size_t)(i - 1) * w + (j - 1)] + sim, M[(size_t)(i - 1) * w + j] + -1,
                                         M[(size_t)i * w + (j - 1)] + -1, 0, &index);
            M[(size_t)i * w + j] = v;
            if (v > matrix_max) { matrix_max = v; i_max = i; j_max = j; }
            int pi = i, pj = j;
            if (index == 0) { pi = i - 1; pj = j - 1; } else if (index == 1) { pi = i - 1; } else if (index == 2) { pj = j - 1; }
            Ii[(size_t)i * w + j] = pi; Ij[(size_t)i * w + j] = pj;
        }
    }
    int ci = i_max, cj = j_max;
    int ni = Ii[(size_t)ci * w + cj], nj = Ij[(size_t)ci * w + cj];
    while (nj != 0 && ni != 0 && (ci != ni || cj != nj)) {
        ci = ni; cj = nj;
        ni = Ii[(size_t)ci * w + cj]; nj = Ij[(size_t)ci * w + cj];
    }
    free(M); free(Ii); free(Ij);
    ci--; cj--;
    if (cj < 0) cj = 0;
    if (ci < 0) ci = 0;
    *a_start_align = ci + a_start_search;
    *a_end_align = *a_start_align + i_max - ci - 1;
    /* a_ret = seqA.substr(current_i + aStartSearch, i_max - current_i + aStartSearch): the LENGTH carries the search
     * offset as well (:282), std::string::substr clips it at the end of the read */
    uint32_t ap = (uint32_t)(ci + a_start_search), al = (uint32_t)(i_max - ci + a_start_search);
    if (al > la - ap) al = la - ap;
    uint32_t bp = (uint32_t)cj, bl = (uint32_t)(j_max - cj);
    if (bl > lb - bp) bl = lb - bp;
    if (similarity != 0) {
        double sim_ld = 1.0 - (orc_edit_distance(a + ap, al, b + bp, bl) / (double)al);
        if (!(sim_ld >= similarity)) {
            *a_start_align = 0; *a_end_align = 0;
            *a_ret_pos = *a_ret_len = *b_ret_pos = *b_ret_len = 0;
            return 0;
        }
    }
    *a_ret_pos = ap; *a_ret_len = al; *b_ret_pos = bp; *b_ret_len = bl;
    return 1;
}

static int64_t bytes_find(const uint8_t* h, uint32_t hl, const uint8_t* nd, uint32_t nl, int last) {
    /* std::string::find / rfind of a non-empty needle; -1 = npos */
    int64_t r = -1;
    if (nl > hl) return -1;
    for (uint32_t i = 0; i + nl <= hl; ++i)
        if (!memcmp(h + i, nd, nl)) { r = i; if (!last) break; }
    return r;
}

int orc_update_start_stops(const uint8_t* seq, uint32_t L, uint32_t* ss, uint32_t* n_ss, uint32_t cap, int front_offset,
                           const uint8_t* dr, uint32_t dr_len, uint32_t low_spacer) {
    /* ReadHolder::updateStartStops (ReadHolder.cpp:382-511): shift every repeat to the consensus DR, then look for one
     * more, partial, repeat in front of the first and behind the last one (similarity cut-off 0.85, at least 4 bases,
     * crassDefines.h:81-82).  n_ss must be even and >= 2; returns -2 if cap is too small, -3 if a shifted start lies
     * past the read (the reference only logs that case and then reads out of bounds). */
    const int DR_length = (int)dr_len;
    uint32_t n = *n_ss;
    if (n < 2 || (n & 1)) return -1;
    for (uint32_t k = 0; k < n; k += 2) {
        int usable_length = DR_length - 1;
        if (front_offset >= (int)ss[k]) {
            int below = front_offset - (int)ss[k];
            usable_length = DR_length - below - 1;
            ss[k] = 0;
        } else ss[k] -= (uint32_t)front_offset;
        if (ss[k] >= L) return -3;
        ss[k + 1] = ss[k] + (uint32_t)usable_length;
        if (ss[k + 1] >= L) ss[k + 1] = L - 1;
    }
    if (ss[0] > low_spacer) {                                                   /* :443-481 front */
        int ps = 0, pe = 0;
        uint32_t ap, al, bp, bl;
        orc_smith_waterman(seq, L, dr, dr_len, 0, (int)(ss[0] - low_spacer), 0.85, &ps, &pe, &ap, &al, &bp, &bl);
        if (pe != 0 && pe - ps >= 4) {
            int64_t at = bytes_find(dr, dr_len, dr + bp, bl, 1);                /* DR->rfind(sp.second) */
            if (at >= 0 && (uint64_t)at + bl == dr_len && ps == 0) {
                if (n + 2 > cap) return -2;
                memmove(ss + 2, ss, n * sizeof(uint32_t));
                ss[0] = 0; ss[1] = (uint32_t)pe;
                n += 2;
            }
        }
    }
    uint32_t end_dist = L - ss[n - 1];                                          /* :483-510 back */
    if (end_dist > low_spacer) {
        int ps = 0, pe = 0;
        uint32_t ap, al, bp, bl;
        orc_smith_waterman(seq, L, dr, dr_len, (int)(ss[n - 1] + low_spacer), (int)(end_dist - low_spacer), 0.85, &ps, &pe, &ap, &al, &bp, &bl);
        if (pe != 0 && pe - ps >= 4) {
            if ((int)L - 1 == pe && bytes_find(dr, dr_len, dr + bp, bl, 0) == 0) {
                if (n + 2 > cap) return -2;
                int diff = (int)al - (int)bl;
                if (diff < 0) diff = -diff;
                ss_add(ss, &n, L, (uint32_t)(ps + diff), (uint32_t)pe);
            }
        }
    }
    *n_ss = n;
    return 0;
}

/* ======================================================================================================================
 * Consensus DR of a group: ksw_align + Aligner (reference: src/crass/ksw.c, src/crass/Aligner.cpp)
 * ====================================================================================================================== */
#define ORC_KSW_XSTOP  0x20000
#define ORC_KSW_XSUBO  0x40000
#define ORC_KSW_XSTART 0x80000
#define ORC_KSW_MAXQ   512                     /* query length the fixed work arrays take (DRs are < 100) */

typedef struct { int score, te, qe, score2, te2, tb, qb; } orc_kswr;

static int orc_sat_add16(int a, int b) { int s = a + b; return s > 32767 ? 32767 : s < -32768 ? -32768 : s; }   /* _mm_adds_epi16 */
static int orc_sat_subu16(int a, int b) {                                                                           /* _mm_subs_epu16 */
    const unsigned ua = (unsigned)a & 0xFFFFu, ub = (unsigned)b & 0xFFFFu;
    return (int)(int16_t)(ua > ub ? ua - ub : 0u);
}
static int orc_max(int a, int b) { return a > b ? a : b; }

/* ksw_i16 (ksw.c:219-322): eight 16-bit lanes, query position k = j + lane * slen sits in lane `lane` of vector j */
static orc_kswr orc_ksw_i16(int qlen, const uint8_t* query, int tlen, const uint8_t* target, const int8_t* mat, int gapo, int gape, int xtra) {
    orc_kswr r = {0, -1, -1, -1, -1, -1, -1};
    const int p = 8, slen = (qlen + p - 1) / p, gapoe = gapo + gape;
    const int minsc = (xtra & ORC_KSW_XSUBO) ? (xtra & 0xffff) : 0x10000;
    const int endsc = (xtra & ORC_KSW_XSTOP) ? (xtra & 0xffff) : 0x10000;
    static __thread int16_t bufs[4][ORC_KSW_MAXQ + 8];
    int16_t *H0 = bufs[0], *H1 = bufs[1], *E = bufs[2], *Hmax = bufs[3];
    int te = -1, gmax = 0, qmax = 0;
    int n_b = 0, m_b = 0;
    uint64_t* b = NULL;
    for (int a = 0; a < 25; ++a) if (mat[a] > qmax) qmax = mat[a];                   /* q->max */
    if (qlen <= 0 || qlen > ORC_KSW_MAXQ) return r;
    memset(H0, 0, sizeof(int16_t) * (size_t)slen * 8);
    memset(E, 0, sizeof(int16_t) * (size_t)slen * 8);
    memset(Hmax, 0, sizeof(int16_t) * (size_t)slen * 8);
    for (int i = 0; i < tlen; ++i) {
        int f[8] = {0}, mx[8] = {0}, h[8], imax = 0;
        const int8_t* ma = mat + target[i] * 5;
        h[0] = 0;
        for (int l = 1; l < 8; ++l) h[l] = H0[(slen - 1) * 8 + l - 1];               /* _mm_slli_si128(H0[slen-1], 2) */
        for (int j = 0; j < slen; ++j) {
            for (int l = 0; l < 8; ++l) {
                const int k = j + l * slen;
                const int sc = k >= qlen ? 0 : ma[query[k]];
                int hh = orc_sat_add16(h[l], sc);
                int e = E[j * 8 + l];
                hh = orc_max(hh, e);
                hh = orc_max(hh, f[l]);
                mx[l] = orc_max(mx[l], hh);
                H1[j * 8 + l] = (int16_t)hh;
                hh = orc_sat_subu16(hh, gapoe);
                e = orc_sat_subu16(e, gape);
                e = orc_max(e, hh);
                E[j * 8 + l] = (int16_t)e;
                f[l] = orc_max(orc_sat_subu16(f[l], gape), hh);
                h[l] = H0[j * 8 + l];
            }
        }
        for (int k = 0; k < 16; ++k) {                                              /* the lazy-F loop */
            int stop = 0;
            for (int l = 7; l > 0; --l) f[l] = f[l - 1];
            f[0] = 0;
            for (int j = 0; j < slen; ++j) {
                int any = 0;
                for (int l = 0; l < 8; ++l) {
                    int hh = orc_max(H1[j * 8 + l], f[l]);
                    H1[j * 8 + l] = (int16_t)hh;
                    hh = orc_sat_subu16(hh, gapoe);
                    f[l] = orc_sat_subu16(f[l], gape);
                    if (f[l] > hh) any = 1;
                }
                if (!any) { stop = 1; break; }
            }
            if (stop) break;
        }
        for (int l = 0; l < 8; ++l) imax = orc_max(imax, mx[l]);
        if (imax >= minsc) {
            if (n_b == 0 || (int32_t)b[n_b - 1] + 1 != i) {
                if (n_b == m_b) { m_b = m_b ? m_b << 1 : 8; b = (uint64_t*)realloc(b, 8 * (size_t)m_b); }
                b[n_b++] = (uint64_t)imax << 32 | (uint32_t)i;
            } else if ((int)(b[n_b - 1] >> 32) < imax) b[n_b - 1] = (uint64_t)imax << 32 | (uint32_t)i;
        }
        if (imax > gmax) {
            gmax = imax; te = i;
            memcpy(Hmax, H1, sizeof(int16_t) * (size_t)slen * 8);
            if (gmax >= endsc) break;
        }
        { int16_t* t = H1; H1 = H0; H0 = t; }
    }
    r.score = gmax; r.te = te;
    {
        int max = -1;
        for (int i = 0; i < slen * 8; ++i)
            if ((int)(uint16_t)Hmax[i] > max) { max = (uint16_t)Hmax[i]; r.qe = i / 8 + i % 8 * slen; }
        if (b) {
            const int w = (r.score + qmax - 1) / qmax;
            const int low = te - w, high = te + w;
            for (int i = 0; i < n_b; ++i) {
                const int e = (int32_t)b[i];
                if ((e < low || e > high) && (uint32_t)(b[i] >> 32) > (uint32_t)r.score2) { r.score2 = (int)(b[i] >> 32); r.te2 = e; }
            }
        }
    }
    free(b);
    return r;
}

static void orc_revseq(int l, uint8_t* s) { for (int i = 0; i < l >> 1; ++i) { const uint8_t t = s[i]; s[i] = s[l - 1 - i]; s[l - 1 - i] = t; } }

static void orc_aligner_matrix(int8_t* mat) {                                        /* Aligner.h:119-131 */
    int k = 0;
    for (int i = 0; i < 4; ++i) { for (int j = 0; j < 4; ++j) mat[k++] = i == j ? 1 : -3; mat[k++] = 0; }
    for (int j = 0; j < 5; ++j) mat[k++] = 0;
}

/* ksw_align (ksw.c:330-354): forward pass, then -- for the start positions -- the same kernel on the reversed prefixes */
static orc_kswr orc_ksw_align_codes(int qlen, uint8_t* query, int tlen, uint8_t* target, const int8_t* mat, int gapo, int gape, int xtra) {
    orc_kswr r = orc_ksw_i16(qlen, query, tlen, target, mat, gapo, gape, xtra);
    if ((xtra & ORC_KSW_XSTART) == 0 || ((xtra & ORC_KSW_XSUBO) && r.score < (xtra & 0xffff))) return r;
    orc_revseq(r.qe + 1, query); orc_revseq(r.te + 1, target);
    const orc_kswr rr = orc_ksw_i16(r.qe + 1, query, tlen, target, mat, gapo, gape, ORC_KSW_XSTOP | r.score);
    orc_revseq(r.qe + 1, query); orc_revseq(r.te + 1, target);
    if (r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
    return r;
}

int orc_ksw_align(const uint8_t* query, int qlen, const uint8_t* target, int tlen, int xtra, int* out7) {
    int8_t mat[25];
    orc_aligner_matrix(mat);
    uint8_t* q = (uint8_t*)malloc((size_t)qlen + 1);
    uint8_t* t = (uint8_t*)malloc((size_t)tlen + 1);
    memcpy(q, query, (size_t)qlen); memcpy(t, target, (size_t)tlen);
    const orc_kswr r = orc_ksw_align_codes(qlen, q, tlen, t, mat, 5, 2, xtra);
    free(q); free(t);
    out7[0] = r.score; out7[1] = r.te; out7[2] = r.qe; out7[3] = r.score2; out7[4] = r.te2; out7[5] = r.tb; out7[6] = r.qb;
    return 0;
}

static uint8_t orc_nt4(uint8_t c) {                                                 /* Aligner::seq_nt4_table (Aligner.cpp:41-58) */
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 4; }
}
static int orc_cov_row(uint8_t c) {                                                 /* CHAR_TO_INDEX - 1 (Aligner.cpp:61-70): everything but C, G, T counts as A */
    switch (c) { case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 0; }
}

enum { ORC_AL_REVERSED = 1, ORC_AL_FAILED = 2, ORC_AL_EQUAL = 4 };

/* Aligner::getOffsetAgainstMaster (Aligner.cpp:263-362) */
static int orc_offset_against_master(const uint8_t* slave, int slen, const uint8_t* master_codes, int mlen, const int8_t* mat, int* flags) {
    const int xtra = ORC_KSW_XSTART | ORC_KSW_XSUBO | 5;
    if (slen <= 0) { *flags |= ORC_AL_EQUAL; return 0; }                            /* (the reference reads before its arrays here) */
    uint8_t* fw = (uint8_t*)malloc((size_t)slen + 1);
    uint8_t* rv = (uint8_t*)malloc((size_t)slen + 1);
    uint8_t* rc = (uint8_t*)malloc((size_t)slen + 1);
    uint8_t* mt = (uint8_t*)malloc((size_t)mlen + 1);
    orc_revcomp(slave, (uint32_t)slen, rc);
    for (int i = 0; i < slen; ++i) { fw[i] = orc_nt4(slave[i]); rv[i] = orc_nt4(rc[i]); }
    memcpy(mt, master_codes, (size_t)mlen);
    const orc_kswr f = orc_ksw_align_codes(slen, fw, mlen, mt, mat, 5, 2, xtra);
    const orc_kswr v = orc_ksw_align_codes(slen, rv, mlen, mt, mat, 5, 2, xtra);
    free(fw); free(rv); free(rc); free(mt);
    if (v.score == f.score) { *flags |= ORC_AL_EQUAL; return 0; }
    orc_kswr best = f;
    if (v.score > f.score) { best = v; *flags |= ORC_AL_REVERSED; }
    if (slen / 2 > best.score) { *flags |= ORC_AL_FAILED; return 0; }
    if (best.score < 5) { *flags |= ORC_AL_FAILED; return 0; }
    return best.tb - best.qb;
}

/* first repeat of the list whose length is dr_len, as the reference's unbounded loops look for it (Aligner.cpp:372-376); -1 = none */
static int orc_first_full_repeat(const uint32_t* ss, uint32_t n_ss, int dr_len) {
    for (uint32_t k = 0; k + 1 < n_ss; k += 2) if ((int)ss[k + 1] - (int)ss[k] == dr_len - 1) return (int)k;
    return -1;
}

/* Aligner::placeReadsInCoverageArray (Aligner.cpp:364-417) for one read; `rev`: the read and its list were reverse complemented
 * (ReadHolder::reverseComplementSeq, ReadHolder.cpp:593-608) */
static int orc_place_read(const uint8_t* seq, uint32_t L, const uint32_t* ss_in, uint32_t n_ss, int rev, int dr_len, int dr_place,
                          uint32_t array_len, int32_t* coverage) {
    uint8_t* s = (uint8_t*)malloc(L + 1);
    uint32_t* ss = (uint32_t*)malloc(sizeof(uint32_t) * (n_ss + 1));
    if (rev) {
        orc_revcomp(seq, L, s);
        for (uint32_t k = 0; k < n_ss; ++k) ss[k] = L - 1 - ss_in[n_ss - 1 - k];
    } else { memcpy(s, seq, L); memcpy(ss, ss_in, sizeof(uint32_t) * n_ss); }
    int rc = 0;
    int k = orc_first_full_repeat(ss, n_ss, dr_len);
    if (k < 0) rc = 1;                                                               /* the reference would run off the list */
    else {
        do {
            if ((int)ss[k + 1] - (int)ss[k] == dr_len - 1) {
                const int start_pos = dr_place - (int)ss[k];
                for (uint32_t i = 0; i < L; ++i) {
                    const int at = (int)i + start_pos;
                    if (at < 0 || at >= (int)array_len) { rc = 2; continue; }       /* logged as memory corruption by the reference, then written anyway */
                    coverage[(size_t)orc_cov_row(s[i]) * array_len + (uint32_t)at]++;
                }
            }
            k += 2;
            if (k >= (int)(n_ss / 2) * 2) break;
        } while ((int)ss[k + 1] - (int)ss[k] == dr_len - 1);
    }
    free(s); free(ss);
    return rc;
}

int orc_consensus_group(const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads, const uint32_t* read_dr,
                        const uint32_t* ss_offsets, const uint32_t* ss_pool, const uint8_t* dr_bytes, const uint32_t* dr_offsets,
                        uint32_t n_drs, uint32_t array_len, int32_t* dr_place, uint8_t* dr_flags, int32_t* zone,
                        uint8_t* consensus, float* conservation, int32_t* coverage) {
    int8_t mat[25];
    orc_aligner_matrix(mat);
    int status = 0;
    memset(coverage, 0, sizeof(int32_t) * 4 * (size_t)array_len);
    const int mlen = (int)(dr_offsets[1] - dr_offsets[0]);
    uint8_t* master = (uint8_t*)malloc((size_t)mlen + 1);
    for (int i = 0; i < mlen; ++i) master[i] = orc_nt4(dr_bytes[dr_offsets[0] + i]);
    dr_place[0] = (int32_t)(array_len * 0.5);                                        /* CRASS_DEF_CONS_ARRAY_START */
    dr_flags[0] = 0;
    for (uint32_t d = 1; d < n_drs; ++d) {
        const uint8_t* slave = dr_bytes + dr_offsets[d];
        const int slen = (int)(dr_offsets[d + 1] - dr_offsets[d]);
        int flags = 0;
        int off = orc_offset_against_master(slave, slen, master, mlen, mat, &flags);
        if (flags & ORC_AL_EQUAL) {                                                  /* extendSlaveDR (Aligner.cpp:420-452): two more bases on either side, from the first read that has them */
            uint8_t ext[ORC_KSW_MAXQ + 8];
            int ext_len = 0;
            for (uint32_t i = 0; i < n_reads && !ext_len; ++i) {
                if (read_dr[i] != d) continue;
                const uint32_t* ss = ss_pool + ss_offsets[i];
                const uint32_t n_ss = ss_offsets[i + 1] - ss_offsets[i];
                const uint32_t L = (uint32_t)(offsets[i + 1] - offsets[i]);
                const int k = orc_first_full_repeat(ss, n_ss, slen);
                if (k < 0) continue;
                if ((int)ss[k] - 2 < 0 || (int)ss[k + 1] + 2 > (int)L) continue;
                uint32_t st = ss[k] - 2, ln = (uint32_t)slen + 4;                    /* std::string::substr clamps at the end of the read */
                if (st + ln > L) ln = L - st;
                memcpy(ext, bases + offsets[i] + st, ln);
                ext_len = (int)ln;
            }
            flags = 0;
            off = orc_offset_against_master(ext, ext_len, master, mlen, mat, &flags);
            if (flags & ORC_AL_EQUAL) flags |= ORC_AL_FAILED;
        }
        if (flags & ORC_AL_FAILED) flags &= ~ORC_AL_REVERSED;                     /* alignSlave returns before it turns the reads round */
        dr_flags[d] = (uint8_t)flags;
        dr_place[d] = (flags & ORC_AL_FAILED) ? -1 : dr_place[0] + off;
    }
    free(master);
    /* the reads of the master and of every placed slave go into the coverage array */
    for (uint32_t i = 0; i < n_reads; ++i) {
        const uint32_t d = read_dr[i];
        if (d >= n_drs || (dr_place[d] < 0 && d != 0)) continue;
        const int rc = orc_place_read(bases + offsets[i], (uint32_t)(offsets[i + 1] - offsets[i]), ss_pool + ss_offsets[i],
                                      ss_offsets[i + 1] - ss_offsets[i], (dr_flags[d] & ORC_AL_REVERSED) != 0,
                                      (int)(dr_offsets[d + 1] - dr_offsets[d]), dr_place[d], array_len, coverage);
        if (rc) status = rc;
    }
    /* calculateDRZone (Aligner.cpp:456-484): the master's own interval */
    zone[0] = dr_place[0];
    zone[1] = dr_place[0] + mlen - 1;
    /* generateConsensus (Aligner.cpp:147-246) */
    int num_gt_zero = 0;
    for (uint32_t j = 0; j < array_len; ++j) {
        static const char alphabet[4] = {'A', 'C', 'G', 'T'};
        int max_count = 0;
        float total = 0.0f;
        consensus[j] = 'N';
        for (int k = 0; k < 4; ++k) {
            const int c = coverage[(size_t)k * array_len + j];
            total += (float)c;
            if (c > max_count) { max_count = c; consensus[j] = (uint8_t)alphabet[k]; }
        }
        if (total > 2) { conservation[j] = (float)max_count / total; num_gt_zero++; }
        else conservation[j] = 0;
    }
    int zs = zone[0], ze = zone[1];
    const int n = (int)array_len;
    if (num_gt_zero >= 2) {
        while (zs > 0 && zs <= n) { if (conservation[zs - 1] < 0.55f) zs++; else break; }      /* (sic: the zone shrinks while its neighbour is poor) */
        while (ze < n - 1 && ze >= -1) { if (conservation[ze + 1] < 0.55f) ze--; else break; }
    }
    while (zs > 0 && zs <= n) { if (conservation[zs - 1] >= 0.55f) zs--; else break; }
    while (ze < n - 1 && ze >= -1) { if (conservation[ze + 1] >= 0.55f) ze++; else break; }
    zone[0] = zs; zone[1] = ze;
    return status;
}

void orc_free(void* p) { free(p); }
