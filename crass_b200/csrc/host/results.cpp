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
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
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
const uint32_t kStrKey = 0xFFFFFFFFu;                                      // "this k-mer goes through the string map"

std::string low_lexi_kmer(std::string_view dr, size_t pos) {              // laurenize (SeqUtils.cpp:89-97)
    std::string k(dr.substr(pos, kClusterKmer));
    std::string rc = reverse_complement(k);
    return k < rc ? k : rc;
}

const int8_t kCode[256] = {
#define X4 -1, -1, -1, -1
#define X16 X4, X4, X4, X4
    X16, X16, X16, X16,
    -1, 0, -1, 1, -1, -1, -1, 2, X4, X4, -1, -1, -1, -1, 3, -1, -1, -1, X4, X4,         // 'A'=65 'C'=67 'G'=71 'T'=84
    X16, X16, X16, X16, X16, X16, X16, X16, X16, X16
#undef X16
#undef X4
};

// the workers of one call meet here between the passes; waits are short, so they spin politely
struct SpinBarrier {
    explicit SpinBarrier(unsigned n) : n_(n) {}
    void wait() {
        if (n_ <= 1) return;
        const unsigned g = gen_.load(std::memory_order_acquire);
        if (count_.fetch_add(1, std::memory_order_acq_rel) + 1 == n_) {
            count_.store(0, std::memory_order_relaxed);
            gen_.fetch_add(1, std::memory_order_release);
        } else {
            while (gen_.load(std::memory_order_acquire) == g) std::this_thread::yield();
        }
    }
    unsigned n_;
    std::atomic<unsigned> count_{0}, gen_{0};
};

// Helper threads that outlive the call: starting eight threads costs about as much as the passes they run.
// run(n, job) executes job(1..n-1) on the helpers and job(0) on the caller and returns when all are through;
// parallel regions of different callers are serialised.
class WorkerPool {
public:
    static WorkerPool& instance() { static WorkerPool p; return p; }
    void run(unsigned n, const std::function<void(unsigned)>& job) {
        if (n <= 1) { job(0); return; }
        after_fork();
        std::lock_guard<std::mutex> one_region(region_);
        {
            std::lock_guard<std::mutex> l(m_);
            while (threads_->size() < n - 1) { const unsigned id = (unsigned)threads_->size(); threads_->emplace_back([this, id]() { loop(id); }); }
            job_ = &job; n_active_ = n - 1; pending_ = n - 1; ++generation_; failure_ = nullptr;
        }
        posted_.store(generation_, std::memory_order_release);
        wake_.notify_all();
        std::exception_ptr mine;
        try { job(0); } catch (...) { mine = std::current_exception(); }   // the helpers still have to finish before the caller unwinds
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [this]() { return pending_ == 0; });
        std::exception_ptr theirs = failure_;
        failure_ = nullptr;
        l.unlock();
        if (mine) std::rethrow_exception(mine);
        if (theirs) std::rethrow_exception(theirs);                     // e.g. bad_alloc inside a helper: the caller turns it into ENOMEM
    }
    // A caller that knows a parallel region is coming within the next fraction of a millisecond (the GPU is still busy
    // with the kernels that feed it) gets the helpers out of their sleep early: they spin for a job until the deadline.
    void prewake(unsigned n, unsigned spin_us) {
        if (n <= 1) return;
        after_fork();
        {
            std::lock_guard<std::mutex> l(m_);
            while (threads_->size() < n - 1) { const unsigned id = (unsigned)threads_->size(); threads_->emplace_back([this, id]() { loop(id); }); }
            spin_until_ = std::chrono::steady_clock::now() + std::chrono::microseconds(spin_us);
            ++prewake_;
        }
        wake_.notify_all();
    }
    ~WorkerPool() {
        if (getpid() != pid_) return;                                // a forked child never had the threads
        { std::lock_guard<std::mutex> l(m_); stop_ = true; }
        wake_.notify_all();
        for (auto& t : *threads_) t.join();
        delete threads_;
    }
private:
    WorkerPool() : threads_(new std::vector<std::thread>()), pid_(getpid()) {}
    // Helper threads do not survive fork(): in a child (Python multiprocessing with the fork start method) the pool
    // starts over with fresh threads and fresh synchronisation objects.  The parent's thread handles and whatever
    // state its mutexes were in at the time of the fork are abandoned, not destroyed.
    void after_fork() {
        if (getpid() == pid_) return;
        threads_ = new std::vector<std::thread>();
        new (&region_) std::mutex(); new (&m_) std::mutex();
        new (&wake_) std::condition_variable(); new (&done_) std::condition_variable();
        job_ = nullptr; n_active_ = pending_ = 0; generation_ = prewake_ = 0; posted_.store(0); stop_ = false; failure_ = nullptr;
        pid_ = getpid();
    }
    void loop(unsigned id) {
        uint64_t seen = 0, seen_prewake = 0;
        for (;;) {
            const std::function<void(unsigned)>* job = nullptr;
            {
                std::unique_lock<std::mutex> l(m_);
                wake_.wait(l, [&]() { return stop_ || (generation_ != seen && id < n_active_) || prewake_ != seen_prewake; });
                if (stop_) return;
                if (!(generation_ != seen && id < n_active_)) {          // woken ahead of a job: poll for it without the lock
                    seen_prewake = prewake_;
                    const auto until = spin_until_;
                    const uint64_t was = generation_;
                    l.unlock();
                    while (posted_.load(std::memory_order_acquire) == was && std::chrono::steady_clock::now() < until) {
#if defined(__x86_64__)
                        __builtin_ia32_pause();
#endif
                    }
                    continue;                                            // back to the wait: a posted job passes it at once
                }
                seen = generation_;
                seen_prewake = prewake_;
                job = job_;
            }
            std::exception_ptr err;
            try { (*job)(id + 1); } catch (...) { err = std::current_exception(); }
            std::lock_guard<std::mutex> l(m_);
            if (err && !failure_) failure_ = err;
            if (--pending_ == 0) done_.notify_one();
        }
    }
    std::mutex region_, m_;
    std::condition_variable wake_, done_;
    std::vector<std::thread>* threads_;
    pid_t pid_;
    std::exception_ptr failure_;
    const std::function<void(unsigned)>* job_ = nullptr;
    unsigned n_active_ = 0, pending_ = 0;
    uint64_t generation_ = 0, prewake_ = 0;
    std::atomic<uint64_t> posted_{0};                    // == generation_, readable without the lock
    std::chrono::steady_clock::time_point spin_until_;
    bool stop_ = false;
};

unsigned cluster_workers(size_t total_kmers) {
    unsigned n = 1;
    if (total_kmers > (1u << 15)) n = std::min<unsigned>(8, std::max<unsigned>(1, std::thread::hardware_concurrency()));
    if (const char* e = getenv("CRASS_B200_HOST_THREADS")) { const int v = atoi(e); if (v >= 1 && v <= 64) n = (unsigned)v; }
    return n;
}

// small per-worker open-addressing map (k-mer key -> head of a chain of survivors) for the substring reduction
struct HeadTable {
    std::vector<uint32_t> key;
    std::vector<int> head;
    size_t mask = 0;
    void reset(size_t want) {
        size_t cap = 16;
        while (cap < 2 * want + 2) cap <<= 1;
        key.assign(cap, kStrKey);
        head.assign(cap, -1);
        mask = cap - 1;
    }
    size_t slot(uint32_t k) const {
        size_t s = (size_t)(k * 0x9E3779B1u) & mask;
        while (key[s] != kStrKey && key[s] != k) s = (s + 1) & mask;
        return s;
    }
};
}  // namespace

void parallel_run(unsigned n, const std::function<void(unsigned)>& job) { WorkerPool::instance().run(n, job); }
unsigned host_threads() {                                // helpers for per-read work (holders of a replay): every core up to 16
    if (getenv("CRASS_B200_HOST_THREADS")) return cluster_workers((size_t)1 << 20);
    return std::min<unsigned>(16, std::max<unsigned>(1, std::thread::hardware_concurrency()));
}

void prewake_cluster_workers(unsigned spin_us) { WorkerPool::instance().prewake(cluster_workers((size_t)1 << 20), spin_us); }

std::vector<std::string> non_redundant_set(const std::vector<std::string>& drs, int min_count,
                                           std::vector<std::pair<int, int> >* groups_out) {
    return non_redundant_set(std::vector<std::string_view>(drs.begin(), drs.end()), min_count, groups_out, nullptr);
}

std::vector<std::string> non_redundant_set(const std::vector<std::string_view>& drs, int min_count,
                                           std::vector<std::pair<int, int> >* groups_out, const ClusterPre* pre_in) {
    // (1) greedy k-mer clustering in token order.  A DR joins the first group that reaches min_count shared
    //     11-mers while walking its k-mers left to right (the test is only made from a group's second hit on);
    //     otherwise it founds a new group.  K-mers never seen before are then given to the chosen group.
    //     K-mers made of A/C/G/T only (practically all of them) are handled as 22-bit integers: with A<C<G<T packed
    //     big-endian, min(forward, reverse complement) as numbers is the lexicographic minimum laurenize() takes.
    //     Anything else goes through the string map; the two key spaces cannot collide because the canonical form of
    //     a k-mer is routed by its own bytes.
    // (2) per group: drop every variant that contains a shorter variant (either strand), then emit the survivors
    //     followed by their reverse complements.
    // The work is cut into passes; all but the short order-dependent one (C) run on a few worker threads that are
    // started once per call.  Nothing in the result depends on the number of workers.
#ifdef CB_PROFILE_NR
    auto t_mark = std::chrono::steady_clock::now();
#define CB_NR_MARK(what) do { const auto now_ = std::chrono::steady_clock::now(); \
        fprintf(stderr, "non_redundant_set: %-10s %.3f ms\n", what, std::chrono::duration<double, std::milli>(now_ - t_mark).count()); t_mark = now_; } while (0)
#else
#define CB_NR_MARK(what) do {} while (0)
#endif
    const size_t n_dr = drs.size();
    std::vector<size_t> koff(n_dr + 1, 0);                                 // k-mers of DR t: [koff[t], koff[t+1])
    for (size_t t = 0; t < n_dr; ++t)
        koff[t + 1] = koff[t] + (drs[t].size() >= kClusterKmer ? drs[t].size() - kClusterKmer + 1 : 0);
    const size_t total_kmers = koff[n_dr];
    // passes A and B may come from the GPU (K5, kernels.cuh): keys[] and, for every k-mer with an integer key, the first
    // DR that holds it.  Only arrays of exactly this list are accepted.
    const ClusterPre* const pre = pre_in && pre_in->total == total_kmers ? pre_in : nullptr;
    size_t tsize = 1024;
    while (!pre && tsize < total_kmers * 2 + 16) tsize <<= 1;
    // the work arrays are kept across calls (fresh multi-megabyte vectors cost more in page faults than the clustering
    // itself); of the hash table only the slots this call touches are reset on the way out, so that untouched slots
    // always hold (empty, INT_MAX)
    struct Slot { uint32_t key; int val; };             // key and value share a cache line: one miss per probe
    static thread_local std::vector<Slot> table_store;
    static thread_local std::vector<uint32_t> keys_store, first_store, runc_store;
    if (table_store.size() < tsize) table_store.assign(tsize, Slot{kStrKey, INT_MAX});
    if (keys_store.size() < total_kmers) {
        keys_store.resize(total_kmers + total_kmers / 4);
        first_store.resize(keys_store.size());
        runc_store.resize(keys_store.size());
    }
    uint32_t* const keys = pre ? const_cast<uint32_t*>(pre->keys) : keys_store.data();   // pass A: canonical key per k-mer (read-only when given)
    uint32_t* const first = pre ? pre->first : first_store.data();         // pass B: first DR (token order) holding it; B2: runs
    uint32_t* const runc = runc_store.data();                              // pass B2: length of each run
    std::vector<uint32_t> n_runs(n_dr, 0);
    tsize = table_store.size();
    Slot* const table = table_store.data();             // plain pointers: TLS lookups are not free inside a shared object

    const unsigned n_workers = cluster_workers(total_kmers);
    std::vector<size_t> cut(n_workers + 1, n_dr);                          // worker w owns DRs [cut[w], cut[w+1]): equal k-mer shares
    cut[0] = 0;
    for (unsigned w = 1; w < n_workers; ++w) {
        const size_t want = total_kmers * w / n_workers;
        cut[w] = std::max(cut[w - 1], (size_t)(std::upper_bound(koff.begin(), koff.end(), want) - koff.begin()) - 1);
    }
    std::vector<std::vector<size_t> > touched_by(n_workers);               // per worker: the table slots it claimed
    struct Reset {
        Slot* tab; std::vector<std::vector<size_t> >& t;
        ~Reset() { for (auto& l : t) for (size_t s : l) tab[s] = Slot{kStrKey, INT_MAX}; }
    } reset_on_exit{table, touched_by};
    std::atomic<bool> any_str{false};
    std::vector<std::vector<std::pair<uint32_t, uint32_t> > > str_pos(n_workers);   // per worker: (DR, k-mer) that go through the string map
    std::vector<std::vector<int> > members;                               // group id - 1 -> tokens
    std::vector<std::vector<std::string> > survivors, survivors_rc;
    std::vector<size_t> schedule;                                          // groups, largest first
    std::atomic<size_t> next_group{0};
    SpinBarrier barrier(n_workers);
    CB_NR_MARK("setup");

    // pass A (no dependencies, streams through the strings): canonical integer key of every k-mer, kStrKey for the
    // rare k-mers that need the string map
    auto pass_a = [&](unsigned worker, size_t t_begin, size_t t_end) {
        const uint32_t kmask = (1u << (2 * kClusterKmer)) - 1u;
        for (size_t t = t_begin; t < t_end; ++t) {
            const std::string_view dr = drs[t];
            size_t w = koff[t];
            uint32_t fw = 0, rc = 0;
            int valid = 0;                                               // trailing run of A/C/G/T bytes
            for (size_t p = 0; p < dr.size(); ++p) {
                const int c = kCode[(uint8_t)dr[p]];
                if (c < 0) valid = 0;
                else { valid++; fw = ((fw << 2) | (uint32_t)c) & kmask; rc = (rc >> 2) | ((uint32_t)(3 - c) << (2 * (kClusterKmer - 1))); }
                if (p + 1 < kClusterKmer) continue;
                uint32_t key = kStrKey;
                if (valid >= (int)kClusterKmer) key = fw < rc ? fw : rc;
                else {
                    const std::string km = low_lexi_kmer(dr, p + 1 - kClusterKmer);
                    uint32_t k2 = 0; bool acgt = true;                   // e.g. a 'U' whose reverse complement is all A/C/G/T
                    for (char ch : km) { const int c2 = kCode[(uint8_t)ch]; if (c2 < 0) { acgt = false; break; } k2 = (k2 << 2) | (uint32_t)c2; }
                    if (acgt) key = k2;
                }
                if (key == kStrKey) { any_str.store(true, std::memory_order_relaxed); str_pos[worker].push_back(std::make_pair((uint32_t)t, (uint32_t)w)); }
                keys[w++] = key;
            }
        }
    };
    // pass B (no dependencies either): first[q] = index of the first DR, in token order, that contains k-mer q.
    // A k-mer is "seen globally" for DR t exactly when first[q] < t, and its group is the group of that first DR:
    // the reference hands its homeless k-mers to the DR's group when the DR is done (WorkHorse.cpp:1611-1617).
    // One hash probe per k-mer with the probes prefetched a fixed distance ahead.  Workers claim a slot with a
    // compare-and-swap on its key and lower its value with an atomic minimum, so the table ends up holding min(t)
    // per k-mer whatever the interleaving; first[] holds the slot until every worker is through (pass B2).
    auto pass_b = [&](unsigned w, size_t t_begin, size_t t_end) {
        const size_t kAhead = 24;
        const size_t q_end = koff[t_end];
        size_t t = t_begin;
        for (size_t q = koff[t_begin]; q < q_end; ++q) {
            if (q + kAhead < q_end && keys[q + kAhead] != kStrKey) {
                const size_t s = (size_t)(keys[q + kAhead] * 0x9E3779B1u) & (tsize - 1);
                __builtin_prefetch(&table[s]);
            }
            while (q >= koff[t + 1]) ++t;
            const uint32_t key = keys[q];
            if (key == kStrKey) continue;
            size_t s = (size_t)(key * 0x9E3779B1u) & (tsize - 1);
            for (;;) {
                uint32_t cur = __atomic_load_n(&table[s].key, __ATOMIC_RELAXED);
                if (cur == key) break;
                if (cur == kStrKey) {
                    if (__atomic_compare_exchange_n(&table[s].key, &cur, key, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) { touched_by[w].push_back(s); break; }
                    if (cur == key) break;
                }
                s = (s + 1) & (tsize - 1);
            }
            int seen = __atomic_load_n(&table[s].val, __ATOMIC_RELAXED);
            while ((int)t < seen && !__atomic_compare_exchange_n(&table[s].val, &seen, (int)t, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
            first[q] = (uint32_t)s;
        }
    };
    // the rare k-mers with other letters go through a string map, on one thread between B and B2
    auto resolve_str = [&]() {
        std::unordered_map<std::string, int> kmer_first_str;
        for (const auto& list : str_pos)                                  // worker ranges are in DR order
            for (const auto& tq : list)
                first[tq.second] = (uint32_t)kmer_first_str.emplace(low_lexi_kmer(drs[tq.first], tq.second - koff[tq.first]), (int)tq.first).first->second;
    };
    // pass B2: slot -> first DR, and in the same sweep the k-mers of every DR are folded into runs (first DR f, length):
    // consecutive already-seen k-mers that were first seen in the same DR.  K-mers new to the DR (first >= t) change no
    // tally in the walk below, so runs reach across them.  The runs overwrite first[] in place.
    auto pass_b2 = [&](size_t t_begin, size_t t_end) {
        const size_t kAhead = 24, q_end = koff[t_end];
        for (size_t t = t_begin; t < t_end; ++t) {
            size_t out = koff[t];
            for (size_t q = koff[t]; q < koff[t + 1]; ++q) {
                if (!pre && q + kAhead < q_end && keys[q + kAhead] != kStrKey) __builtin_prefetch(&table[first[q + kAhead]]);
                const uint32_t f = keys[q] != kStrKey && !pre ? (uint32_t)table[first[q]].val : first[q];
                if (f >= t) continue;                                    // never seen before this DR
                if (out > koff[t] && first[out - 1] == f) runc[out - 1]++;
                else { first[out] = f; runc[out] = 1; ++out; }
            }
            n_runs[t] = (uint32_t)(out - koff[t]);
        }
    };
    // pass C (one thread): the order-dependent greedy walk over the runs.  The reference bumps a group's tally once per
    // shared k-mer and tests it against min_count from the group's second hit on; a run of c k-mers of one group
    // therefore wins as soon as the tally it leaves behind is >= min_count (and >= 2 if it opened the tally).
    std::vector<int> group_of;                                            // DR t -> group id (1-based)
    std::vector<uint8_t> dead;                                            // pass D result when it ran on the GPU
    bool use_dead = false;
    auto pass_c = [&]() {
        group_of.assign(n_dr, 0);
        std::vector<std::pair<int, int> > counts;                         // (group, shared so far)
        for (size_t t = 0; t < n_dr; ++t) {
            counts.clear();
            int group = 0;
            for (size_t j = koff[t], je = koff[t] + n_runs[t]; j < je && !group; ++j) {
                const int known = group_of[first[j]];
                const int c = (int)runc[j];
                auto c2 = std::find_if(counts.begin(), counts.end(), [&](const std::pair<int, int>& p) { return p.first == known; });
                if (c2 == counts.end()) {
                    counts.push_back(std::make_pair(known, c));
                    if (c >= 2 && c >= min_count) group = known;
                } else if ((c2->second += c) >= min_count) group = known;
            }
            if (!group) { members.emplace_back(); group = (int)members.size(); }
            group_of[t] = group;
            members[group - 1].push_back((int)t + 2);
        }
        // pass D may run on the GPU as well (K5, k_cl_reduce): it hands back one "dead" flag per DR
        if (pre && pre->device_reduce) { dead.assign(n_dr, 0); use_dead = pre->device_reduce(group_of.data(), n_dr, dead.data()); }
        if (groups_out)
            for (size_t g = 0; g < members.size(); ++g)
                for (int tok : members[g]) groups_out->push_back(std::make_pair(tok, (int)g + 1));
        survivors.resize(members.size());
        survivors_rc.resize(members.size());
        schedule.resize(members.size());
        for (size_t g = 0; g < members.size(); ++g) schedule[g] = g;
        std::stable_sort(schedule.begin(), schedule.end(), [&](size_t a, size_t b) { return members[a].size() > members[b].size(); });
    };
    // pass D (groups are independent): removeRedundantRepeats.  Variants are visited shortest first; one dies when an
    // earlier one is contained in it on either strand (containment is transitive, so only survivors need to be
    // remembered).  A variant a inside b shows its first 11-mer at the position where it starts, and its reverse
    // complement shows the same canonical 11-mer where it ends, so b's own k-mer keys from pass A find every candidate
    // in a small map keyed by the survivors' first k-mers; candidates are confirmed by comparing bytes.
    // Survivors whose first k-mer holds a byte outside A/C/G/T cannot be found that way and are searched for one by
    // one (there are few); groups with a variant shorter than a k-mer take the plain pairwise search throughout.
    auto reduce_group = [&](size_t g, HeadTable& map, HeadTable& full) {
        std::vector<uint32_t> first_key;                                  // survivor j -> canonical key of its first k-mer
        std::vector<int> v;
        bool plain = false;
        for (int tok : members[g]) {
            const size_t t = (size_t)tok - 2;
            v.push_back((int)t);
            if (drs[t].size() < kClusterKmer) plain = true;
        }
        std::stable_sort(v.begin(), v.end(), [&](int a, int b) { return drs[a].size() < drs[b].size(); });
        std::vector<std::string>& out = survivors[g];
        std::vector<std::string>& out_rc = survivors_rc[g];
        if (plain) {
            for (int tb : v) {
                const std::string_view b = drs[tb];
                if (b.empty()) continue;
                bool dead = false;
                for (size_t j = 0; j < out.size() && !dead; ++j) {
                    const std::string& a = out[j];
                    const std::string& rc = out_rc[j];
                    if (b.size() == a.size()) dead = b == a || b == rc;   // same length: containment is equality
                    else dead = memmem(b.data(), b.size(), a.data(), a.size()) || memmem(b.data(), b.size(), rc.data(), rc.size());
                }
                if (!dead) { out.emplace_back(b); out_rc.push_back(reverse_complement(out.back())); }
            }
            return;
        }
        // survivors enter the k-mer map only once the walk has moved on to longer variants (equal lengths cannot
        // contain each other); among equal lengths only identity on either strand matters, answered by a hash of the
        // whole string
        map.reset(v.size());
        full.reset(2 * v.size());
        std::vector<int> chain;                                           // survivor j -> next survivor with the same first k-mer
        std::vector<int> odd;                                             // survivors that are not in the map, shortest first
        size_t n_mapped = 0, cur_len = 0;
        auto hash_of = [](std::string_view x) {
            uint64_t h = 1469598103934665603ull;
            for (unsigned char c : x) { h ^= c; h *= 1099511628211ull; }
            return (uint32_t)(h ^ (h >> 32)) & 0x7FFFFFFFu;              // never the empty marker
        };
        for (int tb : v) {
            const std::string_view b = drs[tb];
            const uint32_t* kb = &keys[koff[tb]];
            const size_t nk = koff[tb + 1] - koff[tb];
            if (b.size() != cur_len) {
                for (; n_mapped < out.size(); ++n_mapped) {
                    bool acgt = true;
                    for (size_t i = 0; i < kClusterKmer; ++i) acgt = acgt && kCode[(uint8_t)out[n_mapped][i]] >= 0;
                    if (!acgt) { odd.push_back((int)n_mapped); continue; }
                    const size_t s = map.slot(first_key[n_mapped]);
                    chain[n_mapped] = map.key[s] == kStrKey ? -1 : map.head[s];
                    map.key[s] = first_key[n_mapped];
                    map.head[s] = (int)n_mapped;
                }
                cur_len = b.size();
            }
            bool dead = false;
            const uint32_t hb = hash_of(b);
            for (size_t s = (size_t)(hb * 0x9E3779B1u) & full.mask; full.key[s] != kStrKey && !dead; s = (s + 1) & full.mask) {
                if (full.key[s] != hb) continue;
                const int j = full.head[s] >> 1;                          // bit 0: the entry stands for the reverse complement
                dead = b == ((full.head[s] & 1) ? out_rc[j] : out[j]);
            }
            for (size_t o = 0; o < odd.size() && !dead; ++o) {
                const std::string& a = out[odd[o]];
                if (a.size() >= b.size()) break;                          // equal lengths were settled above
                dead = memmem(b.data(), b.size(), a.data(), a.size()) || memmem(b.data(), b.size(), out_rc[odd[o]].data(), a.size());
            }
            for (size_t p = 0; p < nk && !dead; ++p) {
                if (kb[p] == kStrKey) continue;
                const size_t s = map.slot(kb[p]);
                if (map.key[s] == kStrKey) continue;
                for (int j = map.head[s]; j >= 0 && !dead; j = chain[j]) {
                    const std::string& a = out[j];
                    const size_t la = a.size();
                    if (p + la <= b.size() && memcmp(b.data() + p, a.data(), la) == 0) dead = true;
                    else if (p + kClusterKmer >= la && memcmp(b.data() + (p + kClusterKmer - la), out_rc[j].data(), la) == 0) dead = true;
                }
            }
            if (dead) continue;
            const int j = (int)out.size();
            out.emplace_back(b);
            out_rc.push_back(reverse_complement(out.back()));
            first_key.push_back(kb[0]);
            chain.push_back(-1);
            for (int strand = 0; strand < 2; ++strand) {
                const uint32_t h = strand ? hash_of(out_rc[j]) : hb;
                size_t s = (size_t)(h * 0x9E3779B1u) & full.mask;
                while (full.key[s] != kStrKey) s = (s + 1) & full.mask;
                full.key[s] = h;
                full.head[s] = 2 * j + strand;
            }
        }
    };
    // A pass that throws (bad_alloc) must not leave the other workers waiting at the next barrier: the failure is noted,
    // every worker still walks through all the barriers with the remaining passes skipped, and the caller rethrows.
    std::atomic<bool> failed{false};
    auto guarded = [&](auto&& pass) {
        if (failed.load(std::memory_order_relaxed)) return;
        try { pass(); } catch (...) { failed.store(true); }
    };
    auto worker = [&](unsigned w) {
        guarded([&]() {
            if (pre) {                                                    // A and B came from the GPU: only list the string-keyed k-mers
                if (pre->n_string_keys && pre->string_tq) {               // ... which the GPU may have listed already
                    if (w == 0) {
                        for (uint32_t i = 0; i < pre->n_string_keys; ++i) {
                            const uint32_t t = pre->string_tq[2 * i], q = pre->string_tq[2 * i + 1];
                            if (t < n_dr && q >= koff[t] && q < koff[t + 1] && keys[q] == kStrKey) str_pos[0].push_back(std::make_pair(t, q));
                        }
                        std::sort(str_pos[0].begin(), str_pos[0].end());   // DR order, as the sequential map needs it
                        any_str.store(true, std::memory_order_relaxed);
                    }
                } else if (pre->n_string_keys)
                    for (size_t t = cut[w]; t < cut[w + 1]; ++t)
                        for (size_t q = koff[t]; q < koff[t + 1]; ++q)
                            if (keys[q] == kStrKey) { any_str.store(true, std::memory_order_relaxed); str_pos[w].push_back(std::make_pair((uint32_t)t, (uint32_t)q)); }
            } else {
                pass_a(w, cut[w], cut[w + 1]);
                pass_b(w, cut[w], cut[w + 1]);
            }
        });
        barrier.wait();
        if (w == 0) CB_NR_MARK("pass A+B");
        if (any_str.load()) {
            if (w == 0) guarded(resolve_str);
            barrier.wait();
        }
        guarded([&]() { pass_b2(cut[w], cut[w + 1]); });
        barrier.wait();
        if (w == 0) { CB_NR_MARK("pass B2"); guarded(pass_c); CB_NR_MARK("pass C"); }
        barrier.wait();
        guarded([&]() {
            HeadTable map, full;
            for (size_t i; (i = next_group.fetch_add(1)) < schedule.size();) {
                const size_t g = schedule[i];
                if (!use_dead) { reduce_group(g, map, full); continue; }
                std::vector<int> v;                                       // the group's survivors, shortest first (stable)
                for (int tok : members[g]) if (!dead[(size_t)tok - 2] && !drs[(size_t)tok - 2].empty()) v.push_back(tok - 2);
                std::stable_sort(v.begin(), v.end(), [&](int a, int b) { return drs[a].size() < drs[b].size(); });
                for (int t : v) { survivors[g].emplace_back(drs[t]); survivors_rc[g].push_back(reverse_complement(survivors[g].back())); }
            }
        });
    };
    WorkerPool::instance().run(n_workers, worker);
    if (failed.load()) throw std::bad_alloc();
    CB_NR_MARK("reduce");
    std::vector<std::string> out;
    for (size_t g = 0; g < members.size(); ++g) {
        for (std::string& s : survivors[g]) out.push_back(std::move(s));
        for (std::string& s : survivors_rc[g]) out.push_back(std::move(s));
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
    if (r.lazy_kmer_clust) {                             // the whole-path drivers leave this to whoever wants the text
        r.token_groups.clear();
        r.non_redundant = non_redundant_set(r.t2s, r.lazy_kmer_clust, &r.token_groups);
        r.lazy_kmer_clust = 0;
    }
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
