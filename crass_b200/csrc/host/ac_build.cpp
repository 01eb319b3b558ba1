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

#include <queue>

#include "internal.h"

namespace cbh {

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
    *out = A;
    return 0;
}

}  // namespace cbh
