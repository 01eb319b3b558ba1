// ac_build.cpp -- compiles the non-redundant DR patterns into what kernel K2 needs.
//
// The reference builds a mischasan/aho-corasick interleaved state matrix (acism_create,
// src/aho-corasick/acism_create.c:72-132) and stops at the first callback (on_match returns 1,
// libcrispr.cpp:441; acism.c:86-87), i.e. it only ever needs, per read, the match with the smallest
// end offset and, among those, the longest pattern.  Two device structures answer that question:
//
//  (1) fast path, all patterns >= 23 bytes (DRs are >= lowDRsize, so always with default options):
//      * q-gram filter: every occurrence [a, a+len) of a pattern contains the read-aligned 16-mer that starts at
//        8*ceil(a/8) (8i <= a+7 and 8i+16 <= a+23 <= a+len), so a read can only match if one of its 16-mers at
//        offsets 0, 8, 16, ... is a 16-mer of some pattern.  16-mers are 2-bit codes ((byte>>1)&3: equal bytes
//        give equal codes, the test can only over-report).  That 16-mer starts at pattern offset 8*ceil(a/8) - a,
//        i.e. at one of the offsets 0..7, so only those eight 16-mers of every pattern are keys (not all len-15 of
//        them: 2.5x fewer keys, and the share of reads that go to the key table falls with them).  Two levels: a
//        2^19/2^20-bit bitmap staged in shared memory, and the exact open-addressing key table in global memory
//        probed on bitmap hits.
//      * start table: first 16-mer of every pattern -> chain of the patterns that begin with it.  The candidate
//        kernel slides over the read, looks every 16-mer up and verifies the chained patterns byte by byte; the
//        earliest end (longest pattern on ties) over all verified occurrences is exactly acism's first callback.
//      Both are built in one linear pass over the patterns (no trie, no sorting): O(#pattern bytes).
//  (2) generic path (some pattern < 23 bytes, or CRASS_B200_K2=generic): a dense DFA (goto function with the failure
//      links folded in) with one "longest pattern ending here" value per state, one table load per base:
//          entry = table[state * stride + (sym - 1)]      next = entry & 0xFFFFFF, out_len = entry >> 24
//      symv maps a byte to 1..n_syms-1, or 0 for bytes that occur in no pattern (those reset the scan to the root
//      exactly like acism.c:36-42).  Built lazily (ensure_dfa) because it is the expensive part.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>

#include "internal.h"

namespace cbh {

Automaton::~Automaton() { free_device_tables(this); }

uint32_t qgram_hash(uint32_t code, uint32_t bits) { return (code * 0x9E3779B1u) >> (32 - bits); }

// The tables themselves are built on the device (k_ac_build in cluster.cuh, launched when a context first uses the
// matcher): one thread per (pattern, window) sets the bitmap bits and claims the key-table / start-table slots with
// atomics -- the same open-addressing scheme, so look-ups probe exactly as they would in a table filled sequentially.
// What the host decides here are the sizes.  CRASS_B200_AC_BUILD=host fills the tables on the host as well (the round-1
// path; tests compare the two).
static void plan_filter_and_starts(Automaton* A) {
    A->q_bits = 0;
    A->q_bitmap.clear(); A->q_keys.clear(); A->s_keys.clear(); A->s_head.clear(); A->p_next.clear();
    A->tables_on_host = false;
    if (A->min_pattern_len < 23) return;                             // no aligned 16-mer guaranteed: K2 walks the DFA instead
    const uint32_t n = A->n_patterns;
    // the key table is sized for the number of 16-mer POSITIONS (an upper bound of the distinct codes), so that the
    // codes can be inserted in one pass without sorting
    const size_t positions = (size_t)n * 8;                          // pattern offsets 0..7 (every pattern is >= 23 bytes)
    uint32_t tbits = 4;
    while (((size_t)1 << tbits) < positions * 2 + 2) ++tbits;
    A->q_table_bits = tbits;
    // keys up to which the 64 KB bitmap is used (128 KB above).  Measured on config 5 (tools/bench_ac_sweep.py): even at
    // 160 k keys (20 k patterns, 27 % of the bits set) two resident CTAs with 64 KB each beat one with 128 KB.
    uint32_t small_max = 1000000;
    if (const char* e = getenv("CRASS_B200_QGRAM_SMALL_MAX")) small_max = (uint32_t)strtoul(e, nullptr, 10);
    A->q_bits = positions <= small_max ? 19 : 20;
    if (const char* e = getenv("CRASS_B200_QGRAM_BITS")) { const int b = atoi(e); if (b >= 15 && b <= 20) A->q_bits = (uint32_t)b; }
    // A half-size copy for the 2-bit-stream filter (k_ac_filter_packed): its CTA is the bitmap + 40 KB of tiles, so 32 KB
    // instead of 64 KB lets three CTAs instead of two live on an SM, which is worth more than the bits while the set is
    // sparse (config 5, 50 M reads: 100 patterns 1.80 -> 1.63 ms, 1 000: 2.33 -> 2.15 ms, 3 000: level, 20 000: worse).
    // h18 = h19 >> 1 for the multiplicative hash, so the copy is the OR of neighbouring bits (set during insertion).
    A->q_bits_small = 0;
    A->q_bitmap_small.clear();
    uint32_t fold_max = 16384;
    if (const char* e = getenv("CRASS_B200_QGRAM_FOLD_MAX")) fold_max = (uint32_t)strtoul(e, nullptr, 10);
    if (A->q_bits == 19 && positions <= fold_max) A->q_bits_small = 18;
    // two hash functions behind the bitmap while that keeps it sparse (kernels.cuh: qgram_second_bit)
    uint32_t two_max = 100000;
    if (const char* e = getenv("CRASS_B200_QGRAM_TWO_MAX")) two_max = (uint32_t)strtoul(e, nullptr, 10);
    A->q_hashes = positions <= two_max ? 2 : 1;
    uint32_t sbits = 4;
    while (((size_t)1 << sbits) < (size_t)n * 2 + 2) ++sbits;
    A->s_bits = sbits;
    A->s_ones_head = 0xFFFFFFFFu;
    A->q_has_ones = 0;
}

static void fill_filter_and_starts_on_host(Automaton* A) {
    const uint32_t n = A->n_patterns;
    const uint8_t* bytes = A->p_bytes.data();
    const uint32_t* offs = A->p_offs.data();
    const uint32_t tbits = A->q_table_bits, sbits = A->s_bits;
    A->q_keys.assign((size_t)1 << tbits, 0xFFFFFFFFu);               // 0xFFFFFFFF = empty; the all-G 16-mer is kept in q_has_ones
    A->q_bitmap.assign((size_t)1 << (A->q_bits - 5), 0);
    if (A->q_bits_small) A->q_bitmap_small.assign((size_t)1 << (A->q_bits_small - 5), 0);
    A->s_keys.assign((size_t)1 << sbits, 0xFFFFFFFFu);
    A->s_head.assign((size_t)1 << sbits, 0xFFFFFFFFu);
    A->p_next.assign(n, 0xFFFFFFFFu);
    A->s_ones_head = 0xFFFFFFFFu;
    A->q_has_ones = 0;
    const uint32_t tmask = ((uint32_t)1 << tbits) - 1, smask = ((uint32_t)1 << sbits) - 1;
    uint32_t distinct = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t len = offs[i + 1] - offs[i];
        uint32_t code = 0;
        for (uint32_t k = 0; k < len; ++k) {
            code = (code >> 2) | ((uint32_t)((bytes[offs[i] + k] >> 1) & 3) << 30);   // base k of the window in bits [2k,2k+2)
            if (k < 15) continue;
            if (k > 22) break;                                       // window starts 0..7 only (see the header comment)
            const uint32_t h = qgram_hash(code, A->q_bits);
            A->q_bitmap[h >> 5] |= 1u << (h & 31);
            if (A->q_bits_small) { const uint32_t hs = h >> 1; A->q_bitmap_small[hs >> 5] |= 1u << (hs & 31); }
            if (A->q_hashes > 1) {                                               // second hash of the Bloom filter (kernels.cuh: qgram_second_bit)
                const uint32_t h2 = (code * 0xC2B2AE35u) >> (32 - A->q_bits);
                A->q_bitmap[h2 >> 5] |= 1u << (h2 & 31);
                if (A->q_bits_small) { const uint32_t hs = h2 >> 1; A->q_bitmap_small[hs >> 5] |= 1u << (hs & 31); }
            }
            if (code == 0xFFFFFFFFu) { distinct += !A->q_has_ones; A->q_has_ones = 1; }
            else {
                uint32_t slot = (code * 0x85EBCA6Bu) >> (32 - tbits);
                while (A->q_keys[slot] != 0xFFFFFFFFu && A->q_keys[slot] != code) slot = (slot + 1) & tmask;
                if (A->q_keys[slot] != code) { A->q_keys[slot] = code; ++distinct; }
            }
            if (k == 15) {                                           // the pattern's first 16-mer: chain it in the start table
                if (code == 0xFFFFFFFFu) { A->p_next[i] = A->s_ones_head; A->s_ones_head = i; }
                else {
                    uint32_t slot = (code * 0x85EBCA6Bu) >> (32 - sbits);
                    while (A->s_keys[slot] != 0xFFFFFFFFu && A->s_keys[slot] != code) slot = (slot + 1) & smask;
                    A->s_keys[slot] = code;
                    A->p_next[i] = A->s_head[slot];
                    A->s_head[slot] = i;
                }
            }
        }
    }
    A->q_count = distinct;
    A->tables_on_host = true;
}

static void build_filter_and_starts(Automaton* A) {
    plan_filter_and_starts(A);
    const char* sel = getenv("CRASS_B200_AC_BUILD");
    if (A->q_bits && sel && !strcmp(sel, "host")) fill_filter_and_starts_on_host(A);
}

int build_automaton(const uint8_t* bytes, const uint32_t* offs, uint32_t n, Automaton** out) {
    if (n == 0) return fail(CRASS_B200_EINVAL, "ac_build: empty pattern set (the reference guards this case in WorkHorse.cpp:373)");
    uint32_t lo = 0xffffffffu, hi = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t len = offs[i + 1] - offs[i];
        if (len == 0) return fail(CRASS_B200_EINVAL, "ac_build: empty pattern");
        if (len > 255) return fail(CRASS_B200_EINVAL, "ac_build: pattern longer than 255 bytes");
        lo = std::min(lo, len); hi = std::max(hi, len);
    }
    if ((size_t)offs[n] + 1 >= (1u << 24)) return fail(CRASS_B200_EINVAL, "ac_build: more than 2^24 automaton states");
    Automaton* A = new Automaton();
    A->min_pattern_len = lo;
    A->max_pattern_len = hi;
    A->n_patterns = n;
    A->p_bytes.resize((size_t)offs[n] + 16);                         // 16 bytes of slack for word-wise device reads
    memcpy(A->p_bytes.data(), bytes, offs[n]);
    memset(A->p_bytes.data() + offs[n], 0, 16);
    A->p_offs.assign(offs, offs + n + 1);
    build_filter_and_starts(A);
    static std::atomic<uint64_t> next_serial{1};
    A->serial = next_serial++;
    *out = A;
    return 0;
}

// byte -> symbol map of the dense DFA (1..n_syms-1 in order of first appearance, 0 = in no pattern); only the generic K2
// path and the introspection calls need it, so it is not part of every build
void ensure_symbols(Automaton* A) {
    if (A->n_syms) return;
    memset(A->symv, 0, sizeof A->symv);
    uint32_t ns = 1;
    const size_t total = A->p_offs[A->n_patterns];
    for (size_t k = 0; k < total; ++k)
        if (!A->symv[A->p_bytes[k]]) A->symv[A->p_bytes[k]] = (uint8_t)(ns++);
    A->n_syms = ns;
    uint32_t stride = 1;
    while (stride < ns - 1) stride <<= 1;
    A->stride = stride;
}

void ensure_dfa(Automaton* A) {
    if (A->has_dfa) return;
    ensure_symbols(A);
    const uint32_t n = A->n_patterns;
    const uint8_t* bytes = A->p_bytes.data();
    const uint32_t* offs = A->p_offs.data();
    const uint32_t real = A->n_syms - 1;                            // symbols that have transitions
    const uint32_t stride = A->stride;
    const size_t total = (size_t)offs[n] + 1;
    // trie in insertion order
    std::vector<int32_t> child(total * real, -1);
    std::vector<uint16_t> depth(total, 0), term(total, 0);
    uint32_t nst = 1;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t st = 0;
        for (uint32_t k = offs[i]; k < offs[i + 1]; ++k) {
            const uint32_t sy = A->symv[bytes[k]] - 1;
            int32_t& c = child[(size_t)st * real + sy];
            if (c < 0) { c = (int32_t)nst; depth[nst] = (uint16_t)(depth[st] + 1); nst++; }
            st = (uint32_t)c;
        }
        term[st] = depth[st];
    }
    // breadth-first renumbering + failure links + longest-output propagation + full goto function
    std::vector<uint32_t> order;                                    // BFS order of old ids
    order.reserve(nst);
    std::vector<uint32_t> newid(nst, 0), fail_old(nst, 0);
    std::vector<uint32_t> go((size_t)nst * real, 0);                 // old ids
    order.push_back(0);
    for (size_t qh = 0; qh < order.size(); ++qh) {
        const uint32_t st = order[qh];
        newid[st] = (uint32_t)qh;
        if (st != 0 && !term[st]) term[st] = term[fail_old[st]];
        for (uint32_t sy = 0; sy < real; ++sy) {
            const int32_t c = child[(size_t)st * real + sy];
            const uint32_t via_fail = st == 0 ? 0 : go[(size_t)fail_old[st] * real + sy];
            if (c < 0) go[(size_t)st * real + sy] = via_fail;
            else {
                go[(size_t)st * real + sy] = (uint32_t)c;
                fail_old[c] = via_fail;
                order.push_back((uint32_t)c);
            }
        }
    }
    A->n_states = nst;
    A->table.assign((size_t)nst * stride, 0);
    A->out_len.assign(nst, 0);
    for (uint32_t st = 0; st < nst; ++st) {
        const uint32_t id = newid[st];
        A->out_len[id] = term[st];
        for (uint32_t sy = 0; sy < real; ++sy) {
            const uint32_t to = go[(size_t)st * real + sy];
            A->table[(size_t)id * stride + sy] = newid[to] | ((uint32_t)term[to] << 24);
        }
    }
    A->has_dfa = true;
}

}  // namespace cbh
