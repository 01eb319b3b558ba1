// internal.h -- host-side structures behind the opaque handles of include/crass_b200.h
#pragma once
#include <stdint.h>
#include <stddef.h>

#include <functional>
#include <memory>
#include <utility>
#include <map>
#include <string>
#include <string_view>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../../include/crass_b200.h"

namespace cbh {

// ---- error reporting (thread-local message behind crass_b200_last_error) ----------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

// ---- pinned host memory hooks (implemented next to the CUDA runtime; plain malloc without a device)
// host memory for read bases: page-locked when a device exists and `want_pinned` (cudaHostAlloc), else ordinary memory (2 MB
// aligned, huge pages asked for); register_host page-locks ordinary memory after the fact (cudaHostRegister)
void* alloc_host(size_t bytes, bool* pinned, bool want_pinned = true);
void free_host(void* p, bool pinned, bool registered = false);
bool register_host(void* p, size_t bytes);

// std::vector whose resize() does not value-initialise: the per-read arrays of a 10 M-read batch are 440 MB, and zero-filling
// them on one thread cost more than the worker threads need to fill them (30 ms of a 350 ms run)
template <class T>
struct NoInitAlloc : std::allocator<T> {
    template <class U> struct rebind { using other = NoInitAlloc<U>; };
    NoInitAlloc() = default;
    template <class U> NoInitAlloc(const NoInitAlloc<U>&) {}
    template <class U, class... A> void construct(U* p, A&&... a) {
        if constexpr (sizeof...(A) == 0) ::new ((void*)p) U; else ::new ((void*)p) U(std::forward<A>(a)...);
    }
};
template <class T> using PodVec = std::vector<T, NoInitAlloc<T> >;

// ---- a parsed read set: the record stream kseq_read() hands to searchFile ------------------------
struct Batch {
    uint8_t* bases = nullptr;           // all reads back to back (pinned when a device exists)
    size_t bases_cap = 0;
    bool pinned = false;                // from cudaHostAlloc
    bool registered = false;            // ordinary memory, page-locked later (pin_now)
    bool want_pinned = false;           // what reserve_bases asks alloc_host for: page-locking costs more than one copy from ordinary memory,
                                        // so only an engine that reuses its pooled batches asks for it
    void pin_now() { if (bases && !pinned && !registered) registered = register_host(bases, bases_cap); }
    PodVec<uint64_t> offsets;           // n+1
    PodVec<char> name_pool;             // NUL-terminated strings
    PodVec<uint64_t> name_off;
    PodVec<char> text_pool;             // comments and qualities, NUL-terminated
    PodVec<int64_t> comment_off;        // -1: seq->comment.s == NULL ; else offset of the string searchFile sees
    PodVec<int64_t> qual_off;           // -1: seq->qual.s == NULL    ; (may be a STALE string of an earlier record)
    uint32_t max_len = 0;
    int parse_status = -1;              // value of the kseq_read() call that ended the loop
    // A range of a streamed input keeps the bases where the parser's pieces wrote them (each piece at its own file offset
    // inside `bases`, so nothing is copied together on the host): `offsets` are then the positions in the back-to-back layout
    // the DEVICE gets (one copy per segment), and the records [rec0, next rec0) of a segment lie on the host from host0 on.
    // Empty = the host layout is back to back too.
    struct Segment { uint64_t rec0, dev0, host0; };
    std::vector<Segment> segs;
    uint8_t* compact = nullptr;         // back-to-back host copy of a segmented batch, made when somebody asks for one
    uint32_t n() const { return (uint32_t)(offsets.size() - 1); }
    const uint8_t* read_ptr(size_t i) const {
        if (segs.empty()) return bases + offsets[i];
        size_t lo = 0, hi = segs.size();                              // last segment with rec0 <= i
        while (hi - lo > 1) { const size_t mid = (lo + hi) / 2; if (segs[mid].rec0 <= i) lo = mid; else hi = mid; }
        return bases + segs[lo].host0 + (offsets[i] - segs[lo].dev0);
    }
    const uint8_t* contiguous();        // `bases` when back to back, else `compact` (built on first use; not thread safe)
    ~Batch();
    void reserve_bases(size_t need);
    // empty again, keeping every buffer (the page-locked base buffer above all: allocating one costs more than parsing
    // a file into it, so the engine hands released batches to the next parse)
    void reset() {
        offsets.clear(); name_pool.clear(); name_off.clear(); text_pool.clear(); comment_off.clear(); qual_off.clear();
        max_len = 0; parse_status = -1;
        segs.clear();
        if (compact) { free(compact); compact = nullptr; }
    }
};

// parses into `reuse` when given (which the caller keeps owning, also on failure), else into a new Batch
int parse_file(const char* path, Batch** out, Batch* reuse = nullptr);
// the same record stream handed out range by range (about range_bytes of the input each; each range ends on a true record
// start and carries kseq's stale comment / quality into the next): 1 = a range was parsed into `reuse`, 0 = the stream had
// ended, < 0 = the error code.  The batch of the LAST range carries the parse status of the stream, the others 0.
struct ParseStream;
ParseStream* parse_stream_open(const char* path, size_t range_bytes);
size_t parse_stream_size(const ParseStream* s);
int parse_stream_next(ParseStream* s, Batch* reuse);
void parse_stream_close(ParseStream* s);

// ---- the containers the reference fills (ReadMap / StringCheck / lookupTable) ----------------------
struct HeldRead {                       // the fields of ReadHolder the path sets (ReadHolder.h:440-451)
    int token = 0;
    int phase = 0;
    bool was_lowlexi = false;           // RH_WasLowLexi
    bool is_fasta = true;               // RH_IsFasta
    uint32_t repeat_len = 0;            // RH_RepeatLength
    std::vector<uint32_t> ss;           // RH_StartStops
    std::string seq, header, comment, qual;
};

struct Results {
    // StringCheck: first token is 2 (StringCheck.cpp:46-55)
    int next_free_token = 1;
    std::map<std::string, int> s2t;
    std::vector<std::string> t2s;       // index = token - 2
    // ReadMap: token -> reads in insertion order
    std::map<int, std::vector<HeldRead*> > reads;
    std::map<std::string, bool> patterns_hash, reads_found;
    // hash indexes in front of the ordered containers (replay only; whoever rewrites the containers clears them)
    std::unordered_map<std::string, int> s2t_index;
    std::unordered_set<std::string> patterns_index, found_index;
    std::vector<std::vector<HeldRead*>*> reads_index;
    size_t n_found_phase1 = 0;
    std::vector<std::string> non_redundant;          // last computed pattern list
    std::vector<std::pair<int, int> > token_groups;  // (token, group id) in group order
    int lazy_kmer_clust = 0;            // != 0: non_redundant / token_groups have not been computed yet (done by the first dump)
    ~Results();
    size_t num_reads() const;
};

void reverse_complement(const uint8_t* in, size_t n, uint8_t* out);
std::string reverse_complement(const std::string& s);
// ReadHolder::DRLowLexi on a held read; returns the token string
std::string dr_lowlexi(HeldRead& h);
void add_read_holder(Results& r, HeldRead* h);
// WorkHorse::createNonRedundantSet on tokens 2..; fills groups (token, gid) when not NULL
std::vector<std::string> non_redundant_set(const std::vector<std::string>& drs, int min_count,
                                           std::vector<std::pair<int, int> >* groups);
// passes A and B computed elsewhere (K5 on the GPU) for exactly the list handed to non_redundant_set: the canonical key
// of every 11-mer in token order (0xFFFFFFFF = goes through the string map) and, for integer keys, the first DR holding it
struct ClusterPre {
    const uint32_t* keys;
    uint32_t* first;            // overwritten (folded into runs)
    size_t total;               // number of 11-mers
    uint32_t n_string_keys;     // how many keys are 0xFFFFFFFF
    const uint32_t* string_tq;  // optional: their (DR, k-mer index) pairs, n_string_keys of them in any order; NULL = scan keys[]
    // optional: pass D elsewhere too.  Given the group of every DR (1-based), fills dead[t] = 1 for every DR that holds an
    // earlier (shorter, or equally long with a smaller t) DR of its group on either strand; returns false to decline.
    std::function<bool(const int* group_of, size_t n, uint8_t* dead)> device_reduce;
};
std::vector<std::string> non_redundant_set(const std::vector<std::string_view>& drs, int min_count,
                                           std::vector<std::pair<int, int> >* groups, const ClusterPre* pre = nullptr);
// runs job(0..n-1) on the library's persistent helper threads (job(0) on the caller) and returns when all are through; an
// exception thrown by a job is rethrown on the caller.  host_threads(): how many the host passes may use
// (CRASS_B200_HOST_THREADS, default min(8, hardware threads)).
void parallel_run(unsigned n, const std::function<void(unsigned)>& job);
unsigned host_threads();
// wakes the helper threads of non_redundant_set ahead of a call that is about to come (they poll for it for spin_us)
void prewake_cluster_workers(unsigned spin_us);
// the DR tokens of a host copy of a token block (include/crass_b200.h) as views into it, in first-appearance order
std::vector<std::string_view> block_views(const void* block, uint32_t cap, uint32_t stride, uint32_t* count, uint32_t* flags);
std::string dump_results(Results& r, int max_read_len);

// ---- multi-pattern automaton (dense DFA, failure links resolved) ---------------------------------------
struct Automaton {
    uint32_t n_states = 0, n_syms = 0;          // n_syms includes symbol 0 = "byte not in any pattern"
    uint8_t symv[256] = {0};
    std::vector<uint32_t> table;                // n_states * stride entries: next_state | (out_len << 24)?  see ac_build.cpp
    uint32_t stride = 0;                        // entries per state (n_syms rounded up to a power of two)
    uint32_t min_pattern_len = 0, max_pattern_len = 0, n_patterns = 0;
    std::vector<uint16_t> out_len;              // longest pattern ending in the state (0 = none)
    bool has_dfa = false;                       // table/out_len are built lazily by ensure_dfa()
    // the patterns themselves (+16 bytes of slack), kept for the verify kernel and the lazy DFA
    std::vector<uint8_t> p_bytes;
    std::vector<uint32_t> p_offs;
    // q-gram pre-filter + pattern-start table (ac_build.cpp): empty when the shortest pattern is below 23 bytes
    uint32_t q_bits = 0, q_table_bits = 0, q_count = 0, q_has_ones = 0;
    std::vector<uint32_t> q_bitmap, q_keys;
    uint32_t q_hashes = 1;                      // hash functions behind the bitmap (1 or 2)
    uint32_t q_bits_small = 0;                  // 0 = none; else a folded copy of the bitmap with fewer bits (same hash family)
    std::vector<uint32_t> q_bitmap_small;
    uint32_t s_bits = 0, s_ones_head = 0xFFFFFFFFu;
    std::vector<uint32_t> s_keys, s_head, p_next;
    bool tables_on_host = false;                // false: bitmap / key table / start table are built on the device (k_ac_build)
    // device copies, owned by the context that uploaded them
    void* d_table = nullptr;
    void* d_out_len = nullptr;
    void* d_q_bitmap = nullptr;
    void* d_q_keys = nullptr;
    int device = -1;
    uint64_t serial = 0;                        // unique per build; contexts key their device copy on it
    ~Automaton();
};
int build_automaton(const uint8_t* bytes, const uint32_t* offs, uint32_t n, Automaton** out);
void ensure_symbols(Automaton* a);              // symv / n_syms / stride, on first use
void ensure_dfa(Automaton* a);                  // builds table/out_len on first use (generic K2 path, introspection)
void free_device_tables(Automaton* a);          // implemented in the CUDA TU

}  // namespace cbh

struct crass_b200_batch { cbh::Batch b; };
struct crass_b200_results { cbh::Results r; };
struct crass_b200_ac { cbh::Automaton a; };
