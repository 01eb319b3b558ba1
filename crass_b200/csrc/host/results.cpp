// results.cpp -- host replay of device hits into the containers the reference fills, and the
// step between the two phases.
//
//   add_read_holder     addReadHolder               libcrispr.cpp:1119-1162
//   dr_lowlexi          ReadHolder::DRLowLexi       ReadHolder.cpp:513-591 (+ reverseComplementSeq :593-610,
//                                                   reverseStartStops :321-380)
//   reverse_complement  reverseComplement/comp_tab  SeqUtils.cpp:51-87
//   non_redundant_set   WorkHorse::createNonRedundantSet / clusterDRReads / removeRedundantRepeats
//                                                   WorkHorse.cpp:648-709, 1404-1637, 612-645, 78-86
// The GPU kernels decide WHICH reads hit and WHERE; everything here is O(hits) bookkeeping that has
// to happen in read order on one thread because token numbers are handed out by first appearance.
#include <limits.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <functional>
#include <sstream>
#include <thread>
#include <unordered_map>

#include "internal.h"

namespace cbh {

namespace {
thread_local std::string g_err;
// complement table of the reference: identity outside the letters, IUPAC aware, 'U'->'A',
// and the reference's own oddity '`' (96) -> '@' (64)
struct CompTab {
    uint8_t t[128];
    CompTab() {
        for (int i = 0; i < 128; ++i) t[i] = (uint8_t)i;
        const char* from = "ABCDEFGHIJKLMNOPQRSTUVWXYZ";
        const char* to   = "TVGHEFCDIJMLKNOPQYSAABWXRZ";
        for (int i = 0; i < 26; ++i) { t[(int)from[i]] = (uint8_t)to[i]; t[(int)from[i] + 32] = (uint8_t)(to[i] + 32); }
        t[96] = 64;
    }
};
const CompTab kComp;
}  // namespace

void set_error(const std::string& msg) { g_err = msg; }
int fail(int code, const std::string& msg) { g_err = msg; return code; }
const char* last_error_cstr() { return g_err.c_str(); }

void reverse_complement(const uint8_t* in, size_t n, uint8_t* out) {
    for (size_t i = 0; i < n; ++i) out[n - 1 - i] = kComp.t[in[i] & 127];
}

std::string reverse_complement(const std::string& s) {
    std::string r(s.size(), '\0');
    reverse_complement((const uint8_t*)s.data(), s.size(), (uint8_t*)&r[0]);
    return r;
}

Results::~Results() {
    for (auto& kv : reads) for (HeldRead* h : kv.second) delete h;
}

size_t Results::num_reads() const {
    size_t n = 0;
    for (auto& kv : reads) n += kv.second.size();
    return n;
}

static std::string repeat_string_at(const HeldRead& h, size_t i) {
    // ReadHolder::repeatStringAt: substr(start, end - start + 1)
    const uint32_t st = h.ss[i];
    if (st > h.seq.size()) return std::string();
    return h.seq.substr(st, (size_t)(h.ss[i + 1] - h.ss[i] + 1));
}

std::string dr_lowlexi(HeldRead& h) {
    const size_t n_rep = h.ss.size() / 2;
    size_t idx;
    if (n_rep == 1) idx = 0;
    else if (n_rep == 2) {
        if (h.ss.front() == 0) idx = 2;                                   // first repeat is a partial
        else if (h.ss.back() == (uint32_t)h.seq.size()) idx = 0;          // never true: ends are clamped to L-1
        else idx = ((int)(h.ss[1] - h.ss[0]) > (int)(h.ss[3] - h.ss[2])) ? 0 : 2;
    } else idx = 2;
    std::string dr = repeat_string_at(h, idx);
    std::string rc = reverse_complement(dr);
    if (dr < rc) { h.was_lowlexi = true; return dr; }
    // flip the read and mirror the coordinates (palindromes take this branch too)
    h.seq = reverse_complement(h.seq);
    const uint32_t L = (uint32_t)h.seq.size();
    std::vector<uint32_t> m(h.ss.size());
    for (size_t i = 0; i < h.ss.size(); ++i) m[i] = L - 1 - h.ss[h.ss.size() - 1 - i];
    h.ss.swap(m);
    h.was_lowlexi = false;
    return rc;
}

void add_read_holder(Results& r, HeldRead* h) {
    std::string dr = dr_lowlexi(*h);
    auto it = r.s2t.find(dr);
    int tok;
    if (it == r.s2t.end()) {
        tok = ++r.next_free_token;                                        // first token is 2
        r.s2t[dr] = tok;
        r.t2s.push_back(dr);
    } else tok = it->second;
    h->token = tok;
    r.reads[tok].push_back(h);
}

// ---- the step between the phases -------------------------------------------------------------------
namespace {
const size_t kClusterKmer = 11;                                            // CRASS_DEF_KMER_SIZE, crassDefines.h:66

std::string low_lexi_kmer(const std::string& dr, size_t pos) {            // laurenize (SeqUtils.cpp:89-97)
    std::string k = dr.substr(pos, kClusterKmer);
    std::string rc = reverse_complement(k);
    return k < rc ? k : rc;
}

}  // namespace

std::vector<std::string> non_redundant_set(const std::vector<std::string>& drs, int min_count,
                                           std::vector<std::pair<int, int> >* groups_out) {
    // (1) greedy k-mer clustering in token order.  A DR joins the first group that reaches min_count shared
    //     11-mers while walking its k-mers left to right (the test is only made from a group's second hit on);
    //     otherwise it founds a new group.  K-mers never seen before are then given to the chosen group.
    //     K-mers made of A/C/G/T only (practically all of them) are handled as 22-bit integers: with A<C<G<T packed
    //     big-endian, min(forward, reverse complement) as numbers is the lexicographic minimum laurenize() takes.
    //     Anything else goes through the string map; the two key spaces cannot collide because the canonical form of
    //     a k-mer is routed by its own bytes.
#ifdef CB_PROFILE_NR
    auto t_mark = std::chrono::steady_clock::now();
#define CB_NR_MARK(what) do { const auto now_ = std::chrono::steady_clock::now(); \
        fprintf(stderr, "non_redundant_set: %-10s %.3f ms\n", what, std::chrono::duration<double, std::milli>(now_ - t_mark).count()); t_mark = now_; } while (0)
#else
#define CB_NR_MARK(what) do {} while (0)
#endif
    std::unordered_map<std::string, int> kmer_group_str;
    size_t total_kmers = 0;
    for (const std::string& d : drs) if (d.size() >= kClusterKmer) total_kmers += d.size() - kClusterKmer + 1;
    size_t tsize = 1024;
    while (tsize < total_kmers * 2 + 16) tsize <<= 1;
    // the table is kept across calls (fresh multi-megabyte vectors cost more in page faults than the clustering itself);
    // only the slots this call touches are reset on the way out
    static thread_local std::vector<uint32_t> tkey_store;
    static thread_local std::vector<int> tval_store;
    if (tkey_store.size() < tsize) { tkey_store.assign(tsize, 0xFFFFFFFFu); tval_store.assign(tsize, INT_MAX); }
    tsize = tkey_store.size();
    uint32_t* const tkey = tkey_store.data();           // plain pointers: TLS lookups are not free inside a shared object
    int* const tval = tval_store.data();
    std::vector<std::vector<size_t> > touched_by;       // per worker: the slots it claimed
    struct Reset {
        uint32_t* k; int* v; std::vector<std::vector<size_t> >& t;
        ~Reset() { for (auto& l : t) for (size_t s : l) { k[s] = 0xFFFFFFFFu; v[s] = INT_MAX; } }
    } reset_on_exit{tkey, tval, touched_by};
    static const int8_t kCode[256] = {
#define X4 -1, -1, -1, -1
#define X16 X4, X4, X4, X4
        X16, X16, X16, X16,
        -1, 0, -1, 1, -1, -1, -1, 2, X4, X4, -1, -1, -1, -1, 3, -1, -1, -1, X4, X4,     // 'A'=65 'C'=67 'G'=71 'T'=84
        X16, X16, X16, X16, X16, X16, X16, X16, X16, X16
#undef X16
#undef X4
    };
    std::vector<std::vector<int> > members;                               // group id - 1 -> tokens
    std::vector<uint32_t> unseen_int;
    std::vector<std::string> unseen_str;
    std::vector<std::pair<int, int> > counts;                             // (group, shared so far)
    // pass A (no dependencies, streams through the strings): canonical integer key of every k-mer, kStr for the rare
    // k-mers that need the string map
    const uint32_t kStr = 0xFFFFFFFFu;
    std::vector<uint32_t> keys(total_kmers);
    std::vector<size_t> koff(drs.size() + 1, 0);
    for (size_t t = 0; t < drs.size(); ++t)
        koff[t + 1] = koff[t] + (drs[t].size() >= kClusterKmer ? drs[t].size() - kClusterKmer + 1 : 0);
    // passes A and B are cut into contiguous DR ranges for a few worker threads when the list is long (a merged
    // multi-rank list, or a deep sample); the result does not depend on the cut
    const unsigned n_workers = total_kmers > (1u << 17) ? std::min<unsigned>(8, std::max<unsigned>(1, std::thread::hardware_concurrency())) : 1;
    auto for_dr_ranges = [&](const std::function<void(unsigned, size_t, size_t)>& body) {
        if (n_workers <= 1) { body(0, 0, drs.size()); return; }
        std::vector<std::thread> pool;
        size_t t0 = 0;
        for (unsigned w = 0; w < n_workers; ++w) {                      // equal shares of k-mers, not of DRs
            const size_t want = total_kmers * (w + 1) / n_workers;
            size_t t1 = (size_t)(std::upper_bound(koff.begin(), koff.end(), want) - koff.begin()) - 1;
            if (w + 1 == n_workers) t1 = drs.size();
            if (t1 < t0) t1 = t0;
            pool.emplace_back(body, w, t0, t1);
            t0 = t1;
        }
        for (auto& th : pool) th.join();
    };
    std::atomic<bool> any_str{false};
    CB_NR_MARK("setup");
    for_dr_ranges([&](unsigned, size_t t_begin, size_t t_end) {
        const uint32_t kmask = (1u << (2 * kClusterKmer)) - 1u;
        for (size_t t = t_begin; t < t_end; ++t) {
            const std::string& dr = drs[t];
            size_t w = koff[t];
            uint32_t fw = 0, rc = 0;
            int valid = 0;                                               // trailing run of A/C/G/T bytes
            for (size_t p = 0; p < dr.size(); ++p) {
                const int c = kCode[(uint8_t)dr[p]];
                if (c < 0) valid = 0;
                else { valid++; fw = ((fw << 2) | (uint32_t)c) & kmask; rc = (rc >> 2) | ((uint32_t)(3 - c) << (2 * (kClusterKmer - 1))); }
                if (p + 1 < kClusterKmer) continue;
                uint32_t key = kStr;
                if (valid >= (int)kClusterKmer) key = fw < rc ? fw : rc;
                else {
                    const std::string km = low_lexi_kmer(dr, p + 1 - kClusterKmer);
                    uint32_t k2 = 0; bool acgt = true;                   // e.g. a 'U' whose reverse complement is all A/C/G/T
                    for (char ch : km) { const int c2 = kCode[(uint8_t)ch]; if (c2 < 0) { acgt = false; break; } k2 = (k2 << 2) | (uint32_t)c2; }
                    if (acgt) key = k2;
                }
                if (key == kStr) any_str.store(true, std::memory_order_relaxed);
                keys[w++] = key;
            }
        }
    });
    CB_NR_MARK("pass A");
    // pass B (no dependencies either): first[q] = index of the first DR, in token order, that contains k-mer q.
    // A k-mer is "seen globally" for DR t exactly when first[q] < t, and its group is the group of that first DR:
    // the reference hands its homeless k-mers to the DR's group when the DR is done (WorkHorse.cpp:1611-1617).
    // One hash probe per k-mer with the probes prefetched a fixed distance ahead.
    //   Workers claim a slot with a compare-and-swap on its key and lower its value with an atomic minimum, so the
    //   table ends up holding min(t) per k-mer whatever the interleaving; untouched slots hold (empty, INT_MAX).
    std::vector<uint32_t> first(total_kmers);
    touched_by.assign(n_workers, std::vector<size_t>());
    for_dr_ranges([&](unsigned w, size_t t_begin, size_t t_end) {
        const size_t kAhead = 24;
        const size_t q_end = koff[t_end];
        size_t t = t_begin;
        for (size_t q = koff[t_begin]; q < q_end; ++q) {
            if (q + kAhead < q_end && keys[q + kAhead] != kStr) {
                const size_t s = (size_t)(keys[q + kAhead] * 0x9E3779B1u) & (tsize - 1);
                __builtin_prefetch(&tkey[s]); __builtin_prefetch(&tval[s]);
            }
            while (q >= koff[t + 1]) ++t;
            const uint32_t key = keys[q];
            if (key == kStr) continue;
            size_t s = (size_t)(key * 0x9E3779B1u) & (tsize - 1);
            for (;;) {
                uint32_t cur = __atomic_load_n(&tkey[s], __ATOMIC_RELAXED);
                if (cur == key) break;
                if (cur == 0xFFFFFFFFu) {
                    if (__atomic_compare_exchange_n(&tkey[s], &cur, key, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) { touched_by[w].push_back(s); break; }
                    if (cur == key) break;
                }
                s = (s + 1) & (tsize - 1);
            }
            int seen = __atomic_load_n(&tval[s], __ATOMIC_RELAXED);
            while ((int)t < seen && !__atomic_compare_exchange_n(&tval[s], &seen, (int)t, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
            first[q] = (uint32_t)s;                                      // the slot for now; resolved below
        }
    });
    for_dr_ranges([&](unsigned, size_t t_begin, size_t t_end) {
        for (size_t q = koff[t_begin]; q < koff[t_end]; ++q)
            if (keys[q] != kStr) first[q] = (uint32_t)tval[first[q]];
    });
    for (size_t t = 0; any_str.load() && t < drs.size(); ++t)             // the rare k-mers with other letters
        for (size_t q = koff[t]; q < koff[t + 1]; ++q)
            if (keys[q] == kStr) {
                auto ins = kmer_group_str.emplace(low_lexi_kmer(drs[t], q - koff[t]), (int)t);
                first[q] = (uint32_t)ins.first->second;
            }
    CB_NR_MARK("pass B");
    // pass C (the order-dependent greedy walk, now on small sequential arrays only)
    std::vector<int> group_of(drs.size(), 0);
    for (size_t t = 0; t < drs.size(); ++t) {
        counts.clear();
        int group = 0;
        for (size_t q = koff[t]; q < koff[t + 1] && !group; ++q) {
            if (first[q] >= t) continue;                                 // never seen before this DR
            const int known = group_of[first[q]];
            auto c2 = std::find_if(counts.begin(), counts.end(), [&](const std::pair<int, int>& p) { return p.first == known; });
            if (c2 == counts.end()) counts.push_back(std::make_pair(known, 1));
            else if (++c2->second >= min_count) group = known;
        }
        if (!group) { members.emplace_back(); group = (int)members.size(); }
        group_of[t] = group;
        members[group - 1].push_back((int)t + 2);
    }
    (void)unseen_int; (void)unseen_str;
    // (2) per group: drop every variant that contains a shorter surviving variant (either strand), then emit
    //     the survivors followed by their reverse complements.
    //     Groups are independent, so they are spread over a few worker threads; the output order (group id, then
    //     survivors, then their reverse complements) does not depend on the thread count.
    CB_NR_MARK("pass C");
    if (groups_out)
        for (size_t g = 0; g < members.size(); ++g)
            for (int tok : members[g]) groups_out->push_back(std::make_pair(tok, (int)g + 1));
    std::vector<std::vector<std::string> > survivors(members.size());
    auto reduce_group = [&](size_t g) {
        std::vector<const std::string*> v;
        for (int tok : members[g]) v.push_back(&drs[tok - 2]);
        std::stable_sort(v.begin(), v.end(), [](const std::string* a, const std::string* b) { return a->size() < b->size(); });
        std::vector<char> dead(v.size(), 0);
        for (size_t i = 0; i < v.size(); ++i) {
            if (dead[i] || v[i]->empty()) continue;
            const std::string& a = *v[i];
            const std::string rc = reverse_complement(a);
            for (size_t j = i + 1; j < v.size(); ++j) {
                if (dead[j] || v[j]->empty()) continue;
                const std::string& b = *v[j];
                if (b.size() == a.size()) {                               // same length: containment is equality
                    if (b == a || b == rc) dead[j] = 1;
                } else if (memmem(b.data(), b.size(), a.data(), a.size()) || memmem(b.data(), b.size(), rc.data(), rc.size())) dead[j] = 1;
            }
        }
        for (size_t i = 0; i < v.size(); ++i) if (!dead[i] && !v[i]->empty()) survivors[g].push_back(*v[i]);
    };
    size_t work = 0;
    for (auto& m : members) work += m.size() * m.size();
    unsigned n_threads = work > 200000 ? std::min<unsigned>(8, std::max<unsigned>(1, std::thread::hardware_concurrency())) : 1;
    if (n_threads <= 1) {
        for (size_t g = 0; g < members.size(); ++g) reduce_group(g);
    } else {
        std::atomic<size_t> next{0};
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < n_threads; ++t)
            pool.emplace_back([&]() { for (size_t g; (g = next++) < members.size();) reduce_group(g); });
        for (auto& th : pool) th.join();
    }
    CB_NR_MARK("reduce");
    std::vector<std::string> out;
    for (size_t g = 0; g < members.size(); ++g) {
        for (const std::string& s : survivors[g]) out.push_back(s);
        for (const std::string& s : survivors[g]) out.push_back(reverse_complement(s));
    }
    CB_NR_MARK("emit");
#undef CB_NR_MARK
    return out;
}

// ---- "crass-dump v1" (tests/dumpfmt.py documents the format; the oracle emits the same text) ----------
static uint32_t fnv1a32(const std::string& s) {
    uint32_t h = 2166136261u;
    for (unsigned char c : s) { h ^= c; h *= 16777619u; }
    return h;
}

std::string dump_results(Results& r, int max_read_len) {
    std::ostringstream os;
    os << "# crass-dump v1\n";
    os << "M\t" << max_read_len << "\t" << r.n_found_phase1 << "\t" << r.patterns_hash.size() << "\n";
    for (auto& tg : r.token_groups) os << "G\t" << tg.first << "\t" << tg.second << "\n";
    std::vector<std::string> p(r.non_redundant);
    std::sort(p.begin(), p.end());
    for (auto& s : p) os << "P\t" << s << "\n";
    for (auto& kv : r.patterns_hash) os << "H\t" << kv.first << "\n";
    for (auto& kv : r.reads) {
        os << "T\t" << kv.first << "\t" << r.t2s[kv.first - 2] << "\t" << kv.second.size() << "\n";
        for (HeldRead* h : kv.second) {
            os << "R\t" << kv.first << "\t" << h->phase << "\t" << h->header << "\t" << (h->was_lowlexi ? 1 : 0) << "\t"
               << h->repeat_len << "\t";
            for (size_t i = 0; i < h->ss.size(); ++i) { if (i) os << ","; os << h->ss[i]; }
            os << "\t" << h->seq << "\t" << h->comment << "\t" << (h->is_fasta ? 1 : 0) << "\t" << fnv1a32(h->qual) << "\n";
        }
    }
    return os.str();
}

}  // namespace cbh
