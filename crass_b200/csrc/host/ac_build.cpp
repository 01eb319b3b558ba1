// ac_build.cpp -- compiles the non-redundant DR patterns into the automaton kernel K2 walks.
//
// The reference builds a mischasan/aho-corasick interleaved state matrix (acism_create,
// src/aho-corasick/acism_create.c:72-132) and stops at the first callback (on_match returns 1,
// libcrispr.cpp:441; acism.c:86-87), i.e. it only ever needs, per read, the match with the smallest
// end offset and, among those, the longest pattern.  That is what a dense DFA (goto function with the
// failure links folded in) with one "longest pattern ending here" value per state answers with a
// single table load per base:
//     entry = table[state * stride + (sym - 1)]      next = entry & 0xFFFFFF, out_len = entry >> 24
// symv maps a byte to 1..n_syms-1, or 0 for bytes that occur in no pattern (those reset the scan to the
// root exactly like acism.c:36-42).  States are numbered breadth-first so that the shallow states a random
// read keeps visiting form a contiguous prefix that the kernel stages in shared memory.
#include <string.h>

#include <algorithm>
#include <atomic>
#include <queue>

#include "internal.h"

namespace cbh {

void build_qgram_filter(Automaton* A, const uint8_t* bytes, const uint32_t* offs, uint32_t n);

Automaton::~Automaton() { free_device_tables(this); }

int build_automaton(const uint8_t* bytes, const uint32_t* offs, uint32_t n, Automaton** out) {
    if (n == 0) return fail(CRASS_B200_EINVAL, "ac_build: empty pattern set (the reference guards this case in WorkHorse.cpp:373)");
    Automaton* A = new Automaton();
    memset(A->symv, 0, sizeof A->symv);
    uint32_t ns = 1;
    size_t total = 1;
    A->min_pattern_len = 0xffffffffu;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t len = offs[i + 1] - offs[i];
        if (len == 0) { delete A; return fail(CRASS_B200_EINVAL, "ac_build: empty pattern"); }
        if (len > 255) { delete A; return fail(CRASS_B200_EINVAL, "ac_build: pattern longer than 255 bytes"); }
        if (len < A->min_pattern_len) A->min_pattern_len = len;
        if (len > A->max_pattern_len) A->max_pattern_len = len;
        total += len;
        for (uint32_t k = offs[i]; k < offs[i + 1]; ++k)
            if (!A->symv[bytes[k]]) A->symv[bytes[k]] = (uint8_t)(ns++);
    }
    if (total >= (1u << 24)) { delete A; return fail(CRASS_B200_EINVAL, "ac_build: more than 2^24 automaton states"); }
    A->n_syms = ns;
    A->n_patterns = n;
    const uint32_t real = ns - 1;                                   // symbols that have transitions
    uint32_t stride = 1;
    while (stride < real) stride <<= 1;
    A->stride = stride;

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
    build_qgram_filter(A, bytes, offs, n);
    static std::atomic<uint64_t> next_serial{1};
    A->serial = next_serial++;
    *out = A;
    return 0;
}

// ---- q-gram pre-filter of kernel K2 ----------------------------------------------------------------------
// Every pattern is at least 23 bytes long (DRs are >= lowDRsize), so any occurrence [a, a+len) in a read contains the
// read-aligned 16-mer that starts at 8*ceil(a/8): 8i <= a+7 and 8i+16 <= a+23 <= a+len.  The kernel therefore only has to
// look up the 16-mers at offsets 0, 8, 16, ... of a read in the set of all 16-mers of all patterns (2-bit codes,
// (byte>>1)&3, so that equal bytes give equal codes and the test can only over-report).  Two levels:
//   bitmap  2^bits-bit Bloom-style bitmap (one multiplicative hash), staged in shared memory by every CTA
//   keys    open-addressing table of the exact 32-bit codes in global memory (L2 resident), probed only on bitmap hits
uint32_t qgram_hash(uint32_t code, uint32_t bits) { return (code * 0x9E3779B1u) >> (32 - bits); }

void build_qgram_filter(Automaton* A, const uint8_t* bytes, const uint32_t* offs, uint32_t n) {
    A->q_bits = 0;
    A->q_bitmap.clear(); A->q_keys.clear();
    if (A->min_pattern_len < 23) return;                            // no guarantee of an aligned 16-mer: K2 uses the plain scan
    std::vector<uint32_t> codes;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t len = offs[i + 1] - offs[i];
        uint32_t code = 0;
        for (uint32_t k = 0; k < len; ++k) {
            code = (code >> 2) | ((uint32_t)((bytes[offs[i] + k] >> 1) & 3) << 30);   // base k of the window in bits [2k,2k+2)
            if (k >= 15) codes.push_back(code);
        }
    }
    std::sort(codes.begin(), codes.end());
    codes.erase(std::unique(codes.begin(), codes.end()), codes.end());
    A->q_count = (uint32_t)codes.size();
    A->q_bits = codes.size() <= 40000 ? 19 : 20;                     // 64 KB or 128 KB of shared memory
    A->q_bitmap.assign((size_t)1 << (A->q_bits - 5), 0);
    uint32_t tbits = 4;
    while (((size_t)1 << tbits) < codes.size() * 2 + 2) ++tbits;
    A->q_table_bits = tbits;
    A->q_keys.assign((size_t)1 << tbits, 0xFFFFFFFFu);               // 0xFFFFFFFF = empty; the all-G 16-mer is kept in q_has_ones
    A->q_has_ones = 0;
    for (uint32_t c : codes) {
        const uint32_t h = qgram_hash(c, A->q_bits);
        A->q_bitmap[h >> 5] |= 1u << (h & 31);
        if (c == 0xFFFFFFFFu) { A->q_has_ones = 1; continue; }
        uint32_t slot = (c * 0x85EBCA6Bu) >> (32 - tbits);
        while (A->q_keys[slot] != 0xFFFFFFFFu) slot = (slot + 1) & (((uint32_t)1 << tbits) - 1);
        A->q_keys[slot] = c;
    }
}

}  // namespace cbh
