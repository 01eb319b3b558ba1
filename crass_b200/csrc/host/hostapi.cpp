// hostapi.cpp -- the host-only entry points of include/crass_b200.h: feed path, replay containers, the
// step between the phases, and the whole-path driver that strings the kernels together the way
// WorkHorse::parseSeqFiles does (WorkHorse.cpp:321-414).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <exception>
#include <sstream>
#include <string>
#include <thread>
#include <string_view>
#include <unordered_set>
#include <vector>

#include "internal.h"

using namespace cbh;

namespace {

char* dup_cstr(const std::string& s) {
    char* p = (char*)malloc(s.size() + 1);
    if (!p) return nullptr;
    memcpy(p, s.data(), s.size());
    p[s.size()] = 0;
    return p;
}

std::vector<std::string> split_lines(const char* text) {
    std::vector<std::string> v;
    if (!text) return v;
    const char* p = text;
    while (*p) {
        const char* e = strchr(p, '\n');
        if (!e) e = p + strlen(p);
        if (e > p) v.push_back(std::string(p, e));
        p = *e ? e + 1 : e;
    }
    return v;
}

// the (header, comment, qual) triple searchFile copies into the ReadHolder (libcrispr.cpp:112-131)
void fill_holder(HeldRead& h, const Batch& b, uint32_t i) {
    h.seq.assign((const char*)b.read_ptr(i), (size_t)(b.offsets[i + 1] - b.offsets[i]));
    h.header = b.name_pool.data() + b.name_off[i];
    if (b.comment_off[i] >= 0) h.comment = b.text_pool.data() + b.comment_off[i];
    if (b.qual_off[i] >= 0) { h.qual = b.text_pool.data() + b.qual_off[i]; h.is_fasta = false; }
}

int check_hits(const Batch& b, const crass_b200_hit* hits, uint32_t n_hits) {
    for (uint32_t k = 0; k < n_hits; ++k) {
        if (hits[k].read_index >= b.n()) return fail(CRASS_B200_EINVAL, "hit refers to a read outside the batch");
        if (k && hits[k].read_index < hits[k - 1].read_index) return fail(CRASS_B200_EINVAL, "hits must be sorted by read_index (replay is order sensitive)");
        if (hits[k].n_ss < 2 || (hits[k].n_ss & 1)) return fail(CRASS_B200_EINVAL, "malformed start/stop list");
    }
    return 0;
}

}  // namespace

extern "C" {

// ---- feed path -----------------------------------------------------------------------------------------
int crass_b200_parse_file(const char* path, crass_b200_batch** out) {
    if (!path || !out) return fail(CRASS_B200_EINVAL, "NULL argument");
    Batch* b = nullptr;
    if (int r = parse_file(path, &b)) return r;
    crass_b200_batch* h = new crass_b200_batch();
    // move the parsed content into the handle
    std::swap(h->b.bases, b->bases); std::swap(h->b.bases_cap, b->bases_cap); std::swap(h->b.pinned, b->pinned); std::swap(h->b.registered, b->registered);
    h->b.offsets.swap(b->offsets); h->b.name_pool.swap(b->name_pool); h->b.name_off.swap(b->name_off);
    h->b.text_pool.swap(b->text_pool); h->b.comment_off.swap(b->comment_off); h->b.qual_off.swap(b->qual_off);
    h->b.max_len = b->max_len; h->b.parse_status = b->parse_status;
    delete b;
    *out = h;
    return 0;
}

struct crass_b200_parse_stream { cbh::ParseStream* s; };

int crass_b200_parse_stream_open(const char* path, uint64_t range_bytes, crass_b200_parse_stream** out) {
    if (!path || !out) return fail(CRASS_B200_EINVAL, "NULL argument");
    cbh::ParseStream* s = parse_stream_open(path, (size_t)range_bytes);
    if (!s) return CRASS_B200_EIO;
    *out = new crass_b200_parse_stream{s};
    return 0;
}

int crass_b200_parse_stream_next(crass_b200_parse_stream* s, crass_b200_batch** out) {
    if (!s || !out) return fail(CRASS_B200_EINVAL, "NULL argument");
    crass_b200_batch* h = new crass_b200_batch();
    const int got = parse_stream_next(s->s, &h->b);
    if (got <= 0) { delete h; *out = nullptr; return got; }
    *out = h;
    return 1;
}

void crass_b200_parse_stream_close(crass_b200_parse_stream* s) {
    if (!s) return;
    parse_stream_close(s->s);
    delete s;
}

int crass_b200_batch_from_memory(const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads, const char* const* names,
                                 crass_b200_batch** out) {
    if (!offsets || !out || (n_reads && !bases)) return fail(CRASS_B200_EINVAL, "NULL argument");
    crass_b200_batch* h = new crass_b200_batch();
    Batch& b = h->b;
    try {
        b.offsets.assign(offsets, offsets + n_reads + 1);
        b.reserve_bases((size_t)offsets[n_reads] + 16);
        memcpy(b.bases, bases, (size_t)offsets[n_reads]);
        char tmp[32];
        for (uint32_t i = 0; i < n_reads; ++i) {
            const char* nm = names ? names[i] : tmp;
            if (!names) snprintf(tmp, sizeof tmp, "r%010u", i);
            b.name_off.push_back(b.name_pool.size());
            b.name_pool.insert(b.name_pool.end(), nm, nm + strlen(nm) + 1);
            b.comment_off.push_back(-1);
            b.qual_off.push_back(-1);
            const uint64_t L = offsets[i + 1] - offsets[i];
            if (L > b.max_len) b.max_len = (uint32_t)L;
        }
    } catch (std::exception& e) {
        delete h;
        return fail(CRASS_B200_ENOMEM, e.what());
    }
    *out = h;
    return 0;
}

void crass_b200_batch_destroy(crass_b200_batch* b) { delete b; }
uint32_t crass_b200_batch_num_reads(const crass_b200_batch* b) { return b ? b->b.n() : 0; }
uint32_t crass_b200_batch_max_read_len(const crass_b200_batch* b) { return b ? b->b.max_len : 0; }
int crass_b200_batch_parse_status(const crass_b200_batch* b) { return b ? b->b.parse_status : -1; }
const uint8_t* crass_b200_batch_bases(const crass_b200_batch* b) {
    if (!b) return nullptr;
    try { return const_cast<Batch&>(b->b).contiguous(); }           // (a streamed range is copied back to back on first use)
    catch (std::exception&) { fail(CRASS_B200_ENOMEM, "batch_bases"); return nullptr; }
}
const uint8_t* crass_b200_batch_read(const crass_b200_batch* b, uint32_t i, uint32_t* len) {
    if (!b || i >= b->b.n()) { if (len) *len = 0; return nullptr; }
    if (len) *len = (uint32_t)(b->b.offsets[i + 1] - b->b.offsets[i]);
    return b->b.read_ptr(i);
}
const uint64_t* crass_b200_batch_offsets(const crass_b200_batch* b) { return b ? b->b.offsets.data() : nullptr; }
const char* crass_b200_batch_name(const crass_b200_batch* b, uint32_t i) {
    return (b && i < b->b.n()) ? b->b.name_pool.data() + b->b.name_off[i] : nullptr;
}
const char* crass_b200_batch_comment(const crass_b200_batch* b, uint32_t i, int* has) {
    if (!b || i >= b->b.n()) { if (has) *has = 0; return nullptr; }
    const int64_t o = b->b.comment_off[i];
    if (has) *has = o >= 0;
    return o >= 0 ? b->b.text_pool.data() + o : "";
}
const char* crass_b200_batch_qual(const crass_b200_batch* b, uint32_t i, int* has) {
    if (!b || i >= b->b.n()) { if (has) *has = 0; return nullptr; }
    const int64_t o = b->b.qual_off[i];
    if (has) *has = o >= 0;
    return o >= 0 ? b->b.text_pool.data() + o : "";
}

// ---- replay ----------------------------------------------------------------------------------------------
int crass_b200_results_create(crass_b200_results** out) {
    if (!out) return fail(CRASS_B200_EINVAL, "NULL argument");
    *out = new crass_b200_results();
    return 0;
}
void crass_b200_results_destroy(crass_b200_results* r) { delete r; }

// Replay is the serial end of the path, so it is cut in two: building the holders (string copies, DRLowLexi with its
// reverse complements) is per-read work and runs on the helper threads; only the container updates, whose order is the
// result, stay on the calling thread.
namespace {
struct Built { HeldRead* h; std::string token, raw0; };

void build_holders(const Batch& b, const crass_b200_hit* hits, const uint32_t* which, uint32_t n, const uint32_t* ss_pool, int phase,
                   std::vector<Built>& out) {
    out.resize(n);
    const unsigned workers = n >= 2048 ? std::min<unsigned>(cbh::host_threads(), 16) : 1;
    cbh::parallel_run(workers, [&](unsigned w) {
        const uint32_t lo = (uint32_t)((uint64_t)n * w / workers), hi = (uint32_t)((uint64_t)n * (w + 1) / workers);
        for (uint32_t k = lo; k < hi; ++k) {
            const crass_b200_hit& ht = hits[which ? which[k] : k];
            HeldRead* h = new HeldRead();
            fill_holder(*h, b, ht.read_index);
            h->ss.assign(ss_pool + ht.ss_offset, ss_pool + ht.ss_offset + ht.n_ss);
            h->repeat_len = phase == 1 ? ht.repeat_len : 0;
            h->phase = phase;
            Built& o = out[k];
            o.h = h;
            if (phase == 1) {                                        // patternsHash takes repeatStringAt(0) of the un-flipped temporary holder
                const uint32_t st = h->ss[0];
                if (st <= h->seq.size()) o.raw0 = h->seq.substr(st, (size_t)(h->ss[1] - h->ss[0] + 1));
            }
            o.token = dr_lowlexi(*h);
        }
    });
}

void insert_holder(Results& r, Built& o) {                           // addReadHolder's container half (libcrispr.cpp:1119-1162)
    // the ordered maps are what the dump walks; a hash index in front of them answers the 99 % of look-ups that hit
    auto hit = r.s2t_index.find(o.token);
    int tok;
    if (hit == r.s2t_index.end()) {
        auto it = r.s2t.find(o.token);                               // (the index is rebuilt lazily: containers filled by other calls)
        if (it == r.s2t.end()) {
            tok = ++r.next_free_token;                               // first token is 2
            r.s2t[o.token] = tok;
            r.t2s.push_back(o.token);
        } else tok = it->second;
        r.s2t_index.emplace(o.token, tok);
    } else tok = hit->second;
    o.h->token = tok;
    if ((size_t)tok >= r.reads_index.size()) r.reads_index.resize((size_t)tok + 64, nullptr);
    std::vector<HeldRead*>*& list = r.reads_index[(size_t)tok];
    if (!list) list = &r.reads[tok];                                 // std::map nodes do not move
    list->push_back(o.h);
}
}  // namespace

int crass_b200_results_add_phase1(crass_b200_results* rh, const crass_b200_batch* bh, const crass_b200_hit* hits, uint32_t n_hits,
                                  const uint32_t* ss_pool) {
    if (!rh || !bh || (n_hits && (!hits || !ss_pool))) return fail(CRASS_B200_EINVAL, "NULL argument");
    Results& r = rh->r;
    const Batch& b = bh->b;
    if (int e = check_hits(b, hits, n_hits)) return e;
    try {
        std::vector<Built> built;
        static const bool trace = getenv("CRASS_B200_TRACE_REPLAY") != nullptr;
        const auto tb0 = std::chrono::steady_clock::now();
        build_holders(b, hits, nullptr, n_hits, ss_pool, 1, built);
        if (trace) fprintf(stderr, "[crass_b200]     phase-1 replay: %u holders built in %.2f ms\n", n_hits,
                           std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tb0).count());
        // searchFile's loop body for a hit (libcrispr.cpp:134-139) fills three containers that do not know of each other: the
        // ReadMap / StringCheck pair, patternsHash and readsFound.  Each is filled in hit order, on a thread of its own.
        std::exception_ptr err_a, err_b, err_c;
        auto fill_reads = [&]() {
            try { for (uint32_t k = 0; k < n_hits; ++k) insert_holder(r, built[k]); } catch (...) { err_a = std::current_exception(); }
        };
        auto fill_patterns = [&]() {
            try {
                for (uint32_t k = 0; k < n_hits; ++k)
                    if (r.patterns_index.insert(built[k].raw0).second) r.patterns_hash[built[k].raw0] = true;
            } catch (...) { err_b = std::current_exception(); }
        };
        auto fill_found = [&]() {
            try {
                for (uint32_t k = 0; k < n_hits; ++k)
                    if (r.found_index.insert(built[k].h->header).second)
                        r.reads_found.emplace_hint(r.reads_found.end(), built[k].h->header, true);   // (a no-op for a header that is there already)
            } catch (...) { err_c = std::current_exception(); }
        };
        if (n_hits >= 1024 && cbh::host_threads() >= 3) cbh::parallel_run(3, [&](unsigned w) { if (w == 0) fill_reads(); else if (w == 1) fill_patterns(); else fill_found(); });
        else { fill_reads(); fill_patterns(); fill_found(); }
        if (err_a) std::rethrow_exception(err_a);
        if (err_b) std::rethrow_exception(err_b);
        if (err_c) std::rethrow_exception(err_c);
    } catch (std::exception& ex) {
        return fail(CRASS_B200_ENOMEM, std::string("results_add_phase1: ") + ex.what());
    }
    r.n_found_phase1 = r.reads_found.size();
    return 0;
}

int crass_b200_results_add_phase2(crass_b200_results* rh, const crass_b200_batch* bh, const crass_b200_hit* hits, uint32_t n_hits,
                                  const uint32_t* ss_pool) {
    return crass_b200_results_add_phase2_ranges(rh, 1, &bh, &hits, &n_hits, &ss_pool);
}

// findSingletons' replay for several batches at once (the ranges of a streamed file, the files of a run), in the order given:
// readsFound is only read in phase 2, so the header test and the holders of ALL the hits are made on the worker threads in one
// go, and only the container inserts (which number new tokens in order) walk the hits one by one.
int crass_b200_results_add_phase2_ranges(crass_b200_results* rh, uint32_t n_ranges, const crass_b200_batch* const* bhs,
                                         const crass_b200_hit* const* hits, const uint32_t* n_hits, const uint32_t* const* ss_pools) {
    if (!rh || (n_ranges && (!bhs || !hits || !n_hits || !ss_pools))) return fail(CRASS_B200_EINVAL, "NULL argument");
    Results& r = rh->r;
    for (uint32_t f = 0; f < n_ranges; ++f) {
        if (!bhs[f] || (n_hits[f] && (!hits[f] || !ss_pools[f]))) return fail(CRASS_B200_EINVAL, "NULL argument");
        if (int e = check_hits(bhs[f]->b, hits[f], n_hits[f])) return e;
    }
    try {
        static const bool trace = getenv("CRASS_B200_TRACE_REPLAY") != nullptr;
        const auto tb0 = std::chrono::steady_clock::now();
        // on_match (libcrispr.cpp:408-442): readsFound is tested by HEADER and never written here, so the test can be made
        // for all hits up front
        if (r.found_index.size() != r.reads_found.size()) {                  // containers filled by other calls: index them now
            r.found_index.clear();
            for (const auto& kv : r.reads_found) r.found_index.insert(kv.first);
        }
        // all hits of all batches as one list: one pass over it on the worker threads for the header test, one for the holders
        size_t total = 0;
        std::vector<size_t> first(n_ranges + 1, 0);
        for (uint32_t f = 0; f < n_ranges; ++f) { total += n_hits[f]; first[f + 1] = total; }
        const unsigned workers = total >= 2048 ? std::min<unsigned>(cbh::host_threads(), 16) : 1;
        struct Ref { uint32_t f, k; };
        std::vector<std::vector<Ref> > part(workers);
        cbh::parallel_run(workers, [&](unsigned w) {
            const size_t lo = total * w / workers, hi = total * (w + 1) / workers;
            uint32_t f = 0;
            std::string name;
            for (size_t g = lo; g < hi; ++g) {
                while (g >= first[f + 1]) ++f;
                const uint32_t k = (uint32_t)(g - first[f]);
                const Batch& b = bhs[f]->b;
                name.assign(b.name_pool.data() + b.name_off[hits[f][k].read_index]);
                if (r.found_index.find(name) == r.found_index.end()) part[w].push_back(Ref{f, k});
            }
        });
        std::vector<Ref> take;
        for (unsigned w = 0; w < workers; ++w) take.insert(take.end(), part[w].begin(), part[w].end());
        std::vector<Built> built(take.size());
        const unsigned bw = take.size() >= 2048 ? workers : 1;
        cbh::parallel_run(bw, [&](unsigned w) {
            const size_t lo = take.size() * w / bw, hi = take.size() * (w + 1) / bw;
            for (size_t g = lo; g < hi; ++g) {
                const Batch& b = bhs[take[g].f]->b;
                const crass_b200_hit& ht = hits[take[g].f][take[g].k];
                HeldRead* h = new HeldRead();
                fill_holder(*h, b, ht.read_index);
                h->ss.assign(ss_pools[take[g].f] + ht.ss_offset, ss_pools[take[g].f] + ht.ss_offset + ht.n_ss);
                h->repeat_len = 0;
                h->phase = 2;
                built[g].h = h;
                built[g].token = dr_lowlexi(*h);
            }
        });
        const auto tb1 = std::chrono::steady_clock::now();
        const size_t taken = built.size();
        for (Built& o : built) insert_holder(r, o);
        if (trace) fprintf(stderr, "[crass_b200]     phase-2 replay: %zu of %zu hits new (%u batches), holders built in %.2f ms, inserted in %.2f ms\n", taken, total, n_ranges,
                           std::chrono::duration<double, std::milli>(tb1 - tb0).count(),
                           std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tb1).count());
    } catch (std::exception& ex) {
        return fail(CRASS_B200_ENOMEM, std::string("results_add_phase2: ") + ex.what());
    }
    return 0;
}

uint32_t crass_b200_results_num_tokens(const crass_b200_results* r) { return r ? (uint32_t)r->r.t2s.size() : 0; }
uint32_t crass_b200_results_num_reads(const crass_b200_results* r) { return r ? (uint32_t)r->r.num_reads() : 0; }

char* crass_b200_results_dr_list(const crass_b200_results* r) {
    std::string s;
    if (r) for (const std::string& d : r->r.t2s) { s += d; s += '\n'; }
    return dup_cstr(s);
}

int crass_b200_results_adopt_tokens(crass_b200_results* rh, const char* dr_list_all_ranks) {
    if (!rh) return fail(CRASS_B200_EINVAL, "NULL argument");
    Results& r = rh->r;
    // global first-appearance order over the rank-ordered concatenation == the sequential token order
    std::vector<std::string> all = split_lines(dr_list_all_ranks);
    std::map<std::string, int> s2t;
    std::vector<std::string> t2s;
    for (const std::string& d : all) if (s2t.find(d) == s2t.end()) { s2t[d] = (int)t2s.size() + 2; t2s.push_back(d); }
    for (const std::string& d : r.t2s) if (s2t.find(d) == s2t.end()) return fail(CRASS_B200_EINVAL, "local DR missing from the gathered list");
    std::map<int, std::vector<HeldRead*> > reads;
    for (auto& kv : r.reads) {
        const int nt = s2t[r.t2s[kv.first - 2]];
        for (HeldRead* h : kv.second) { h->token = nt; reads[nt].push_back(h); }
    }
    r.reads.swap(reads);
    r.s2t.swap(s2t);
    r.s2t_index.clear(); r.reads_index.clear();                     // they pointed into the old containers
    r.t2s.swap(t2s);
    r.next_free_token = (int)r.t2s.size() + 1;
    return 0;
}

char* crass_b200_results_non_redundant(crass_b200_results* rh, uint32_t kmer_clust, uint32_t* n_patterns) {
    if (!rh) return nullptr;
    Results& r = rh->r;
    r.token_groups.clear();
    r.lazy_kmer_clust = 0;
    r.non_redundant = non_redundant_set(r.t2s, (int)kmer_clust, &r.token_groups);
    if (n_patterns) *n_patterns = (uint32_t)r.non_redundant.size();
    std::string s;
    for (const std::string& p : r.non_redundant) { s += p; s += '\n'; }
    return dup_cstr(s);
}

char* crass_b200_results_dump(crass_b200_results* r, int max_read_len) {
    if (!r) return nullptr;
    return dup_cstr(dump_results(r->r, max_read_len));
}

char* crass_b200_dr_list_from_hits(const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads, const crass_b200_hit* hits,
                                   uint32_t n_hits, const uint32_t* ss_pool) {
    if ((n_hits && (!bases || !offsets || !hits || !ss_pool))) { fail(CRASS_B200_EINVAL, "NULL argument"); return nullptr; }
    std::unordered_set<std::string> seen;
    seen.reserve(4096);
    std::string out, rc;
    for (uint32_t k = 0; k < n_hits; ++k) {
        if (k + 8 < n_hits) __builtin_prefetch(bases + offsets[hits[k + 8].read_index]);
        const crass_b200_hit& h = hits[k];
        if (h.read_index >= n_reads || h.n_ss < 2) { fail(CRASS_B200_EINVAL, "malformed hit"); return nullptr; }
        const uint8_t* s = bases + offsets[h.read_index];
        const uint32_t L = (uint32_t)(offsets[h.read_index + 1] - offsets[h.read_index]);
        const uint32_t* ss = ss_pool + h.ss_offset;
        // representative repeat exactly as ReadHolder::DRLowLexi picks it (ReadHolder.cpp:513-566)
        uint32_t idx;
        const uint32_t n_rep = h.n_ss / 2;
        if (n_rep == 1) idx = 0;
        else if (n_rep == 2) {
            if (ss[0] == 0) idx = 2;
            else if (ss[3] == L) idx = 0;
            else idx = ((int)(ss[1] - ss[0]) > (int)(ss[3] - ss[2])) ? 0 : 2;
        } else idx = 2;
        uint32_t st = ss[idx], ln = ss[idx + 1] - ss[idx] + 1;
        if (st > L) st = L;
        if (ln > L - st) ln = L - st;
        std::string dr((const char*)s + st, ln);
        rc.resize(ln);
        reverse_complement((const uint8_t*)dr.data(), ln, (uint8_t*)&rc[0]);
        const std::string& tok = dr < rc ? dr : rc;
        if (seen.insert(tok).second) { out += tok; out += '\n'; }
    }
    return dup_cstr(out);
}

char* crass_b200_dr_list_from_tokens(const uint8_t* records, uint32_t stride, const crass_b200_hit* hits, uint32_t n_hits) {
    if (n_hits && (!records || !hits)) { fail(CRASS_B200_EINVAL, "NULL argument"); return nullptr; }
    std::vector<uint32_t> order(n_hits);
    for (uint32_t i = 0; i < n_hits; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return hits[a].read_index < hits[b].read_index; });
    std::unordered_set<std::string> seen;
    std::string out;
    for (uint32_t i = 0; i < n_hits; ++i) {
        const uint8_t* rec = records + (size_t)order[i] * stride;
        std::string t((const char*)rec + 2, rec[0]);
        if (seen.insert(t).second) { out += t; out += '\n'; }
    }
    return dup_cstr(out);
}

char* crass_b200_dr_list_from_unique(const uint8_t* records, uint32_t stride, const uint32_t* first_read, uint32_t n) {
    if (n && (!records || !first_read)) { fail(CRASS_B200_EINVAL, "NULL argument"); return nullptr; }
    std::vector<uint32_t> order(n);
    for (uint32_t i = 0; i < n; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return first_read[a] < first_read[b]; });
    std::string out;
    for (uint32_t i = 0; i < n; ++i) {
        const uint8_t* rec = records + (size_t)order[i] * stride;
        out.append((const char*)rec + 2, rec[0]);
        out += '\n';
    }
    return dup_cstr(out);
}

}  // extern "C"

// the DR tokens of a host copy of a token block, as views into it, in order-key (first-appearance) order
std::vector<std::string_view> cbh::block_views(const void* block, uint32_t cap, uint32_t stride, uint32_t* count, uint32_t* flags) {
    const uint8_t* p = (const uint8_t*)block;
    uint32_t hdr[2];
    memcpy(hdr, p, sizeof hdr);
    if (count) *count = hdr[0];
    if (flags) *flags = hdr[1];
    const uint32_t n = hdr[0] < cap ? hdr[0] : cap;
    const uint8_t* recs = p + 16;
    std::vector<uint64_t> order(n);                                             // (order key, slot), sorted as one integer
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t key;
        memcpy(&key, recs + (size_t)i * stride + stride - 4, 4);
        order[i] = ((uint64_t)key << 32) | i;
    }
    std::sort(order.begin(), order.end());
    std::vector<std::string_view> v;
    v.reserve(n);
    for (uint32_t i = 0; i < n; ++i) {
        const uint8_t* rec = recs + (size_t)(uint32_t)order[i] * stride;
        const uint32_t ln = rec[0] + 6u <= stride ? rec[0] : stride - 6;
        if (ln) v.push_back(std::string_view((const char*)rec + 2, ln));
    }
    return v;
}

extern "C" {

char* crass_b200_dr_list_from_block(const void* block, uint32_t cap, uint32_t stride, uint32_t* count, uint32_t* flags) {
    if (!block || stride < 12) { fail(CRASS_B200_EINVAL, "bad token block"); return nullptr; }
    std::string out;
    for (const std::string_view& d : block_views(block, cap, stride, count, flags)) { out.append(d); out += '\n'; }
    return dup_cstr(out);
}

char* crass_b200_non_redundant_patterns_from_block(const void* block, uint32_t cap, uint32_t stride, uint32_t kmer_clust,
                                                   uint32_t* count, uint32_t* flags, uint32_t* n_patterns) {
    if (!block || stride < 12) { fail(CRASS_B200_EINVAL, "bad token block"); return nullptr; }
    const std::vector<std::string> nr = non_redundant_set(block_views(block, cap, stride, count, flags), (int)kmer_clust, nullptr);
    if (n_patterns) *n_patterns = (uint32_t)nr.size();
    std::string out;
    for (const std::string& p : nr) { out += p; out += '\n'; }
    return dup_cstr(out);
}

int crass_b200_ac_build_from_block(const void* block, uint32_t cap, uint32_t stride, uint32_t kmer_clust, crass_b200_ac** out,
                                   uint32_t* count, uint32_t* flags, uint32_t* n_patterns) {
    if (!block || stride < 12 || !out) return fail(CRASS_B200_EINVAL, "bad token block");
    *out = nullptr;
    uint32_t c = 0, f = 0;
    const std::vector<std::string_view> drs = block_views(block, cap, stride, &c, &f);
    if (count) *count = c;
    if (flags) *flags = f;
    if (n_patterns) *n_patterns = 0;
    if (f || c > cap || drs.empty()) return 0;                                  // overflowed or empty: no matcher, the caller looks at count/flags
    const std::vector<std::string> nr = non_redundant_set(drs, (int)kmer_clust, nullptr);
    if (n_patterns) *n_patterns = (uint32_t)nr.size();
    std::vector<uint8_t> bytes;
    std::vector<uint32_t> offs(1, 0);
    for (const std::string& p : nr) { bytes.insert(bytes.end(), p.begin(), p.end()); offs.push_back((uint32_t)bytes.size()); }
    if (bytes.empty()) return 0;
    return crass_b200_ac_build(bytes.data(), offs.data(), (uint32_t)nr.size(), out);
}

char* crass_b200_merge_dr_lists(const char* concatenated) {
    if (!concatenated) return dup_cstr(std::string());
    std::unordered_set<std::string_view> seen;                                  // views into the caller's text
    const std::string_view all(concatenated);
    seen.reserve(all.size() / 24 + 16);
    std::string out;
    out.reserve(all.size());
    for (size_t p = 0; p < all.size();) {
        size_t e = all.find('\n', p);
        if (e == std::string_view::npos) e = all.size();
        const std::string_view d = all.substr(p, e - p);
        if (!d.empty() && seen.insert(d).second) { out.append(d); out += '\n'; }
        p = e + 1;
    }
    return dup_cstr(out);
}

char* crass_b200_non_redundant_set(const char* dr_list, uint32_t kmer_clust) {
    std::vector<std::string> drs = split_lines(dr_list);
    std::vector<std::pair<int, int> > groups;
    std::vector<std::string> nr = non_redundant_set(drs, (int)kmer_clust, &groups);
    std::ostringstream os;
    for (auto& g : groups) os << "G\t" << g.first << "\t" << g.second << "\n";
    for (auto& p : nr) os << "P\t" << p << "\n";
    return dup_cstr(os.str());
}

int crass_b200_ac_build_from_dr_list(const char* dr_list, uint32_t kmer_clust, crass_b200_ac** out, uint32_t* n_patterns) {
    if (!dr_list || !out) return fail(CRASS_B200_EINVAL, "NULL argument");
#ifdef CB_PROFILE_NR
    const auto t_in = std::chrono::steady_clock::now();
    struct Tail { std::chrono::steady_clock::time_point t0; ~Tail() {
        fprintf(stderr, "non_redundant_set: whole-call %.3f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count()); } } tail{t_in};
#endif
    std::vector<std::string_view> lines;                                        // views into the caller's text
    const std::string_view all(dr_list);
    lines.reserve(all.size() / 24 + 16);
    for (size_t p = 0; p < all.size();) {
        size_t e = all.find('\n', p);
        if (e == std::string_view::npos) e = all.size();
        if (e > p) lines.push_back(all.substr(p, e - p));
        p = e + 1;
    }
#ifdef CB_PROFILE_NR
    fprintf(stderr, "non_redundant_set: split %.3f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_in).count());
#endif
    const std::vector<std::string> nr = non_redundant_set(lines, (int)kmer_clust, nullptr);
#ifdef CB_PROFILE_NR
    const auto t_nr = std::chrono::steady_clock::now();
    struct Tail2 { std::chrono::steady_clock::time_point t0; ~Tail2() {
        fprintf(stderr, "non_redundant_set: build %.3f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count()); } } tail2{t_nr};
#endif
    if (n_patterns) *n_patterns = (uint32_t)nr.size();
    std::vector<uint8_t> bytes;
    std::vector<uint32_t> offs(1, 0);
    size_t total = 0;
    for (const std::string& p : nr) total += p.size();
    bytes.reserve(total + 1);
    offs.reserve(nr.size() + 1);
    for (const std::string& p : nr) { bytes.insert(bytes.end(), p.begin(), p.end()); offs.push_back((uint32_t)bytes.size()); }
    if (bytes.empty()) bytes.push_back(0);
    return crass_b200_ac_build(bytes.data(), offs.data(), (uint32_t)nr.size(), out);
}

char* crass_b200_non_redundant_patterns(const char* dr_list, uint32_t kmer_clust, uint32_t* n_patterns) {
    if (!dr_list) { fail(CRASS_B200_EINVAL, "NULL argument"); return nullptr; }
    std::vector<std::string_view> lines;
    const std::string_view all(dr_list);
    lines.reserve(all.size() / 24 + 16);
    for (size_t p = 0; p < all.size();) {
        size_t e = all.find('\n', p);
        if (e == std::string_view::npos) e = all.size();
        if (e > p) lines.push_back(all.substr(p, e - p));
        p = e + 1;
    }
    const std::vector<std::string> nr = non_redundant_set(lines, (int)kmer_clust, nullptr);
    if (n_patterns) *n_patterns = (uint32_t)nr.size();
    std::string out;
    out.reserve(all.size());
    for (const std::string& p : nr) { out += p; out += '\n'; }
    return dup_cstr(out);
}

int crass_b200_ac_build_from_pattern_list(const char* patterns, crass_b200_ac** out, uint32_t* n_patterns) {
    if (!patterns || !out) return fail(CRASS_B200_EINVAL, "NULL argument");
    const std::string_view all(patterns);
    std::vector<uint8_t> bytes;
    std::vector<uint32_t> offs(1, 0);
    bytes.reserve(all.size() + 1);
    for (size_t p = 0; p < all.size();) {
        size_t e = all.find('\n', p);
        if (e == std::string_view::npos) e = all.size();
        if (e > p) { bytes.insert(bytes.end(), all.begin() + p, all.begin() + e); offs.push_back((uint32_t)bytes.size()); }
        p = e + 1;
    }
    if (n_patterns) *n_patterns = (uint32_t)offs.size() - 1;
    if (bytes.empty()) bytes.push_back(0);
    return crass_b200_ac_build(bytes.data(), offs.data(), (uint32_t)offs.size() - 1, out);
}

// ---- the whole path -------------------------------------------------------------------------------------------
int crass_b200_run_files(crass_b200_ctx* ctx, const char* const* paths, uint32_t n_paths, const crass_b200_params* params,
                         int phases, crass_b200_results** out, int* max_read_len) {
    if (!ctx || !paths || !params || !out) return fail(CRASS_B200_EINVAL, "NULL argument");
    crass_b200_results* res = nullptr;
    if (int r = crass_b200_results_create(&res)) return r;
    std::vector<crass_b200_batch*> batches(n_paths, nullptr);
    std::vector<std::vector<uint8_t> > found(n_paths);
    int rc = 0, max_len = 0;
    auto cleanup = [&]() { for (auto* b : batches) crass_b200_batch_destroy(b); };
    for (uint32_t f = 0; f < n_paths && !rc; ++f) {                          // phase 1: searchFile per file
        rc = crass_b200_parse_file(paths[f], &batches[f]);
        if (rc) break;
        const Batch& b = batches[f]->b;
        if ((int)b.max_len > max_len) max_len = (int)b.max_len;
        found[f].assign(b.n() + 1, 0);
        crass_b200_hit* hits = nullptr; uint32_t nh = 0, np = 0; uint32_t* pool = nullptr;
        rc = crass_b200_batch_upload(ctx, b.bases, b.offsets.data(), b.n());
        if (!rc) rc = crass_b200_dr_search_resident(ctx, params, found[f].data(), &hits, &nh, &pool, &np);
        if (!rc) rc = crass_b200_results_add_phase1(res, batches[f], hits, nh, pool);
        free(hits); free(pool);
    }
    if (!rc) {
        uint32_t n_pat = 0;
        char* pats = crass_b200_results_non_redundant(res, params->kmer_clust, &n_pat);   // createNonRedundantSet
        if (phases >= 2 && n_pat > 0) {                                      // WorkHorse.cpp:373 guards the empty set
            std::vector<uint8_t> bytes; std::vector<uint32_t> offs(1, 0);
            for (const std::string& p : res->r.non_redundant) { bytes.insert(bytes.end(), p.begin(), p.end()); offs.push_back((uint32_t)bytes.size()); }
            crass_b200_ac* ac = nullptr;
            rc = crass_b200_ac_build(bytes.data(), offs.data(), n_pat, &ac);
            for (uint32_t f = 0; f < n_paths && !rc; ++f) {                  // phase 2: findSingletons per file
                const Batch& b = batches[f]->b;
                crass_b200_hit* hits = nullptr; uint32_t nh = 0, np = 0; uint32_t* pool = nullptr;
                if (n_paths == 1) rc = crass_b200_ac_scan_resident(ctx, ac, 1, nullptr, &hits, &nh, &pool, &np);   // batch still resident
                else rc = crass_b200_ac_scan(ctx, ac, b.bases, b.offsets.data(), b.n(), found[f].data(), nullptr, &hits, &nh, &pool, &np);
                if (!rc) rc = crass_b200_results_add_phase2(res, batches[f], hits, nh, pool);
                free(hits); free(pool);
            }
            crass_b200_ac_destroy(ac);
        }
        free(pats);
    }
    cleanup();
    if (rc) { crass_b200_results_destroy(res); return rc; }
    if (max_read_len) *max_read_len = max_len;
    *out = res;
    return 0;
}

}  // extern "C"
