// capi.cu -- the extern "C" boundary (include/crass_b200.h): contexts, launches, host<->device plumbing.
//
// No CPU fallback lives here: every compute entry point needs a CUDA device and fails with
// CRASS_B200_ENODEVICE otherwise.  Host-side bookkeeping (parser, replay, clustering, automaton
// construction) is in host/*.cpp.
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <sstream>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../include/crass_b200.h"
#include "host/internal.h"
#include "kernels.cuh"
#include "nccl_dl.h"

namespace cbh { const char* last_error_cstr(); }

#define CB_STR2(x) #x
#define CB_STR(x) CB_STR2(x)

namespace {

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            std::ostringstream _s;                                                                       \
            _s << #expr << " failed: " << cudaGetErrorString(_e) << " (" << __FILE__ << ":" << __LINE__ << ")"; \
            return cbh::fail(_e == cudaErrorMemoryAllocation ? CRASS_B200_ENOMEM : CRASS_B200_ECUDA, _s.str()); \
        }                                                                                                \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        CUDA_TRY(cudaMalloc(&p, want));
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};

struct PinnedBuf {                                  // grow-only page-locked host staging
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        CUDA_TRY(cudaHostAlloc(&p, want, cudaHostAllocDefault));
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};

int g_device_count = -2;    // -2: not probed
Nccl g_nccl;

int probe_devices() {
    if (g_device_count == -2) {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        g_device_count = (e == cudaSuccess) ? n : 0;
        if (e != cudaSuccess) (void)cudaGetLastError();
    }
    return g_device_count;
}

cb::Params to_core(const crass_b200_params& p) {
    cb::Params o;
    o.low_dr = p.low_dr; o.high_dr = p.high_dr; o.low_spacer = p.low_spacer; o.high_spacer = p.high_spacer;
    o.window = p.window; o.min_repeats = p.min_repeats; o.kmer_clust = p.kmer_clust; o.scan_range = p.scan_range;
    return o;
}

int validate_params(const crass_b200_params* p) {
    if (!p) return cbh::fail(CRASS_B200_EINVAL, "params is NULL");
    if (p->window < 1 || p->window > 32) return cbh::fail(CRASS_B200_EINVAL, "window must be in 1..32 (the reference CLI clamps it to 6..9)");
    if (p->min_repeats < 1) return cbh::fail(CRASS_B200_EINVAL, "min_repeats must be >= 1 (0 is undefined behaviour in the reference)");
    if (p->high_dr > (uint32_t)cb::kMaxEdit || p->high_spacer > (uint32_t)cb::kMaxEdit)
        return cbh::fail(CRASS_B200_EINVAL, "high_dr / high_spacer above 255 are not supported");
    if (p->low_dr > p->high_dr || p->low_spacer > p->high_spacer) return cbh::fail(CRASS_B200_EINVAL, "low bound above high bound");
    return 0;
}

}  // namespace

struct crass_b200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    uint64_t launches = 0;
    uint64_t last_candidates = 0;
    uint64_t hits_per_64k = 0, pool_per_64k = 0;         // the most records / pool words per 65 536 reads any host-form call needed
    DevBuf d_bases, d_offsets, d_found, d_skip, d_hits, d_pool, d_counters, d_scratch, d_error, d_misc, d_symv;
    uint32_t* h_counters = nullptr;   // pinned, 8 words
    // resident batch (crass_b200_batch_upload)
    uint32_t res_n_reads = 0, res_max_len = 0;
    uint64_t res_n_bases = 0;
    bool res_valid = false, res_found_valid = false;
    DevBuf d_found_p1, d_cand, d_tokens, d_tok_table, d_tok_unique, d_ac_table, d_ac_symv, d_ac_bitmap, d_ac_keys;
    DevBuf d_ac_skeys, d_ac_shead, d_ac_pnext, d_ac_poffs, d_ac_pbytes;
    DevBuf d_cand_counts;                // K1 fast path: per chunk, sizes of the two candidate lists + queue head
    // K1 fast path, chunk pipeline: the exact kernel of chunk i runs on `side` beside the filter kernel of chunk i+1
    cudaStream_t side = nullptr;         // highest priority: its few CTAs take the room the filter grid leaves on every SM
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_chunk[16] = {};
    DevBuf d_ac_bitmap_small;            // folded copy of the q-gram bitmap for k_ac_filter_packed (0 bytes when unused)
    DevBuf d_cand_mask;                  // K2 fast path: per candidate, the aligned 16-mers that can belong to an occurrence
    // K5 (clustering passes A/B on the token block): device arrays and their pinned host mirrors
    DevBuf d_cl_ckeys;
    void* comm = nullptr;                // ncclComm_t of the one-process-per-GPU mode (crass_b200_ctx_comm_init)
    int comm_world = 0, comm_rank = 0;
    DevBuf d_cons;                       // K7 (consensus DR): every device array of a call, carved from one allocation
    DevBuf d_ticket;                     // work counter of the long-read kernel
    DevBuf d_long_list;                  // mixed batches: the reads the short-read filter leaves to the long-read kernel
    DevBuf d_cl_order, d_cl_koff, d_cl_keys, d_cl_first, d_cl_tab, d_cl_info;
    DevBuf d_cl_group, d_cl_chain, d_cl_next, d_cl_odd, d_cl_dead, d_cl_str;
    PinnedBuf h_cl_block, h_cl_order, h_cl_keys, h_cl_first, h_cl_info, h_cl_group, h_cl_dead, h_cl_str;
    PinnedBuf h_ac_stage;                // matcher tables on their way to the device
    DevBuf d_ac_ones;                    // {some pattern holds the all-ones 16-mer, head of the chain of patterns that begin with it}
    DevBuf d_cl_tail;                    // every array of cluster.cuh's ClusterTail, carved from one allocation
    PinnedBuf h_cl_pat;                  // the pattern set on its way back from the device
    cudaEvent_t ev_ac_ready = nullptr;   // recorded on `stream` behind the build of the matcher tables; scans on other streams wait for it
    cudaEvent_t ev_scan = nullptr;       // recorded behind the last scan; a new matcher's tables wait for it before they overwrite the old ones
    bool scan_recorded = false;
    uint32_t cl_last_patterns = 0, cl_last_bytes = 0;   // sizes of the last device-built pattern set (how much the next call copies back unasked)
    bool stage_busy = false;             // h_ac_stage is the source of copies that may not have run yet (ev_ac_ready tells)
    // the 2-bit stream of the batch, written by k_dr_filter and read by k_ac_filter_packed (crass_b200_ctx_keep_packed)
    DevBuf d_packed;
    uint64_t keep_packed_bases = 0;      // caller opt-in for the *_dev calls: capacity in bases, 0 = off
    const void* packed_src = nullptr;    // d_bases / n_reads the stream was made from
    uint32_t packed_reads = 0;
    bool packed_valid = false;
    bool packed_internal = false;        // set around the resident calls: the context owns the batch, reuse is always safe
    DevBuf d_rank, d_hits_sorted;        // crass_b200_sort_hits_dev: per-chunk prefix counts, read-ordered copy of the hits
    uint64_t ac_serial = 0;              // build serial of the automaton currently held in d_ac_*
    uint64_t ac_dfa_serial = 0;          // ... and of the dense DFA (generic K2 path), uploaded lazily
    // K4 token output of the next dr_search launches (crass_b200_ctx_set_token_output / host forms)
    uint8_t* tok_ptr = nullptr;
    uint32_t tok_stride = 0;
    std::string last_dr_list;
};

namespace cbh {

void* alloc_host(size_t bytes, bool* pinned, bool want_pinned) {
    *pinned = false;
    if (want_pinned && probe_devices() > 0) {
        void* p = nullptr;
        if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess) { *pinned = true; return p; }
        (void)cudaGetLastError();
    }
    if (bytes >= ((size_t)8 << 20)) {                            // large and ordinary: 2 MB aligned, huge pages if the kernel grants them
        void* p = nullptr;
        if (posix_memalign(&p, (size_t)2 << 20, bytes) == 0) { madvise(p, bytes, MADV_HUGEPAGE); return p; }
    }
    return malloc(bytes);
}

void free_host(void* p, bool pinned, bool registered) {
    if (!p) return;
    if (pinned) { cudaFreeHost(p); return; }
    if (registered) { if (cudaHostUnregister(p) != cudaSuccess) (void)cudaGetLastError(); }
    free(p);
}

bool register_host(void* p, size_t bytes) {
    if (!p || probe_devices() <= 0) return false;
    if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess) return true;
    (void)cudaGetLastError();
    return false;
}

void free_device_tables(Automaton*) {}                     // device copies are owned by the contexts (d_ac_*)

}  // namespace cbh

// ======================================================================================================
extern "C" {

const char* crass_b200_last_error(void) { return cbh::last_error_cstr(); }
int crass_b200_abi_version(void) { return CRASS_B200_ABI_VERSION; }
const char* crass_b200_build_info(void) {
    return "crass_b200 hot path; CUDA " CB_STR(CUDART_VERSION) "; sm_100a; kernels: K1 dr_filter_warp (2-bit seeds, VIADDMNMX) | dr_filter (TMA tiles) + dr_exact_staged, dr_long; "
           "K2 ac_build (tables on the device) + ac_filter[_packed|_long] (16-mer q-gram, two-hash bitmap) + ac_verify_mask|_warp; K4 tokens, K4b/K4c token blocks; "
           "K5 createNonRedundantSet on the device (cl_rank .. cl_walk .. cl_dead_packed .. cl_emit); hit ordering; K6 update_start_stops (warp Smith-Waterman); "
           "K7 ksw_align (8 threads per alignment) + consensus groups; generic dr_search / ac_scan; edit_distance";
}
int crass_b200_device_count(void) { return probe_devices(); }

void crass_b200_default_params(crass_b200_params* p) {
    p->low_dr = 23; p->high_dr = 47; p->low_spacer = 26; p->high_spacer = 50;      // crassDefines.h:121-124
    p->window = 8; p->min_repeats = 2; p->kmer_clust = 6; p->scan_range = 24;        // :56,:91,:67 ; libcrispr.cpp:347
}

int crass_b200_ctx_create(int device, crass_b200_ctx** out) {
    if (!out) return cbh::fail(CRASS_B200_EINVAL, "out is NULL");
    if (probe_devices() <= 0) return cbh::fail(CRASS_B200_ENODEVICE, "no CUDA device: crass_b200 has no CPU execution path");
    if (device < 0 || device >= probe_devices()) return cbh::fail(CRASS_B200_EINVAL, "bad device ordinal");
    CUDA_TRY(cudaSetDevice(device));
    crass_b200_ctx* c = new crass_b200_ctx();
    c->device = device;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    int prio_lo = 0, prio_hi = 0;
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CUDA_TRY(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, prio_hi));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_ac_ready, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_scan, cudaEventDisableTiming));
    for (cudaEvent_t& e : c->ev_chunk) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CUDA_TRY(cudaHostAlloc((void**)&c->h_counters, 8 * sizeof(uint32_t), cudaHostAllocDefault));
    *out = c;
    return 0;
}

void crass_b200_ctx_destroy(crass_b200_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    DevBuf* bufs[] = {&c->d_bases, &c->d_offsets, &c->d_found, &c->d_skip, &c->d_hits, &c->d_pool, &c->d_counters,
                      &c->d_scratch, &c->d_error, &c->d_misc, &c->d_symv, &c->d_found_p1, &c->d_cand, &c->d_tokens, &c->d_tok_table, &c->d_tok_unique, &c->d_ac_table, &c->d_ac_symv,
                      &c->d_ac_bitmap, &c->d_ac_keys, &c->d_ac_skeys, &c->d_ac_shead, &c->d_ac_pnext, &c->d_ac_poffs, &c->d_ac_pbytes,
                      &c->d_rank, &c->d_hits_sorted, &c->d_cand_counts, &c->d_cand_mask, &c->d_ac_bitmap_small, &c->d_packed,
                      &c->d_cl_order, &c->d_cl_koff, &c->d_cl_keys, &c->d_cl_first, &c->d_cl_tab, &c->d_cl_info,
                      &c->d_cl_group, &c->d_cl_chain, &c->d_cl_next, &c->d_cl_odd, &c->d_cl_dead, &c->d_cl_str, &c->d_ac_ones, &c->d_cl_tail, &c->d_cl_ckeys, &c->d_ticket, &c->d_cons, &c->d_long_list};
    for (DevBuf* b : bufs) b->release();
    for (PinnedBuf* b : {&c->h_cl_block, &c->h_cl_order, &c->h_cl_keys, &c->h_cl_first, &c->h_cl_info, &c->h_ac_stage,
                         &c->h_cl_group, &c->h_cl_dead, &c->h_cl_str, &c->h_cl_pat}) b->release();
    if (c->h_counters) cudaFreeHost(c->h_counters);
    if (c->side) { cudaStreamSynchronize(c->side); cudaStreamDestroy(c->side); }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_ac_ready) cudaEventDestroy(c->ev_ac_ready);
    if (c->ev_scan) cudaEventDestroy(c->ev_scan);
    for (cudaEvent_t e : c->ev_chunk) if (e) cudaEventDestroy(e);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    cudaStreamDestroy(c->stream);
    delete c;
}

int crass_b200_ctx_device(const crass_b200_ctx* c) { return c ? c->device : -1; }
uint64_t crass_b200_ctx_launch_count(const crass_b200_ctx* c) { return c ? c->launches : 0; }
uint64_t crass_b200_ctx_last_candidates(const crass_b200_ctx* c) { return c ? c->last_candidates : 0; }
int crass_b200_ctx_set_token_output(crass_b200_ctx* c, void* d_tokens, uint32_t stride) {
    if (!c) return cbh::fail(CRASS_B200_EINVAL, "ctx is NULL");
    if (d_tokens && stride < 8) return cbh::fail(CRASS_B200_EINVAL, "token stride too small");
    c->tok_ptr = (uint8_t*)d_tokens; c->tok_stride = d_tokens ? stride : 0;
    return 0;
}
const char* crass_b200_ctx_last_dr_list(const crass_b200_ctx* c) { return c ? c->last_dr_list.c_str() : ""; }
int crass_b200_ctx_keep_packed_bases(crass_b200_ctx* c, uint64_t n_bases) {
    if (!c) return cbh::fail(CRASS_B200_EINVAL, "ctx is NULL");
    c->keep_packed_bases = n_bases > 1 ? n_bases : (n_bases ? 1 : 0);
    if (!n_bases) c->packed_valid = false;
    return 0;
}
int crass_b200_ctx_keep_packed(crass_b200_ctx* c, int on) {
    if (!c) return cbh::fail(CRASS_B200_EINVAL, "ctx is NULL");
    c->keep_packed_bases = on ? 1 : 0;
    if (!on) c->packed_valid = false;
    return 0;
}

// hit records come off the device in slot order; replay wants read order.  Read indices are distinct, so three or four
// stable 8-bit counting passes over (index, slot) pairs do it in a fraction of the time of a comparison sort.
void crass_b200_sort_hits(crass_b200_hit* hits, uint32_t n) {
    if (!hits || n < 2) return;
    if (n < 512) {
        std::sort(hits, hits + n, [](const crass_b200_hit& a, const crass_b200_hit& b) { return a.read_index < b.read_index; });
        return;
    }
    // sort (read index, slot) pairs, then move every record once; the work arrays are kept per thread
    static thread_local std::vector<uint64_t> ka, kb;
    static thread_local std::vector<crass_b200_hit> tmp;
    if (ka.size() < n) { ka.resize(n + n / 4); kb.resize(ka.size()); tmp.resize(ka.size()); }
    uint64_t* src = ka.data();
    uint64_t* dst = kb.data();
    uint32_t top = 0;
    for (uint32_t i = 0; i < n; ++i) { top |= hits[i].read_index; src[i] = ((uint64_t)hits[i].read_index << 32) | i; }
    for (uint32_t shift = 0; shift < 32 && (top >> shift); shift += 8) {       // 256 write streams stay inside L1
        uint32_t count[257] = {0};
        for (uint32_t i = 0; i < n; ++i) count[((src[i] >> (32 + shift)) & 255u) + 1]++;
        for (uint32_t b = 0; b < 256; ++b) count[b + 1] += count[b];
        for (uint32_t i = 0; i < n; ++i) dst[count[(src[i] >> (32 + shift)) & 255u]++] = src[i];
        std::swap(src, dst);
    }
    memcpy(tmp.data(), hits, sizeof(crass_b200_hit) * (size_t)n);
    for (uint32_t i = 0; i < n; ++i) hits[i] = tmp[(uint32_t)src[i]];
}

int crass_b200_unique_tokens_dev(crass_b200_ctx* c, const crass_b200_hit* d_hits, uint32_t n_hits, const void* d_tokens, uint32_t stride,
                                 void* d_out_tokens, uint32_t* d_out_first_read, uint32_t* d_out_count, void* stream_v) {
    if (!c || !d_out_count) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    if (stride < 8 || (stride & 3)) return cbh::fail(CRASS_B200_EINVAL, "token stride must be a multiple of 4, at least 8");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream_v;
    CUDA_TRY(cudaMemsetAsync(d_out_count, 0, sizeof(uint32_t), st));
    if (n_hits == 0) return 0;
    if (!d_hits || !d_tokens || !d_out_tokens || !d_out_first_read) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    uint32_t cap = 1024;
    while (cap < 2u * n_hits) cap <<= 1;
    if (int r = c->d_tok_table.reserve((size_t)cap * 2 * sizeof(uint32_t))) return r;
    uint32_t* rep = c->d_tok_table.as<uint32_t>();
    uint32_t* first_read = rep + cap;
    CUDA_TRY(cudaMemsetAsync(rep, 0xFF, (size_t)cap * 2 * sizeof(uint32_t), st));
    cbk::k_token_dedupe<<<(n_hits + 255) / 256, 256, 0, st>>>(d_hits, n_hits, (const uint8_t*)d_tokens, stride, rep, first_read, cap - 1);
    cbk::k_token_compact<<<(cap + 255) / 256, 256, 0, st>>>(rep, first_read, cap, (const uint8_t*)d_tokens, stride, (uint8_t*)d_out_tokens,
                                                           d_out_first_read, d_out_count);
    c->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int crass_b200_sort_hits_dev(crass_b200_ctx* c, const uint8_t* d_found, uint32_t n_reads, const crass_b200_hit* d_hits,
                             const uint32_t* d_n_hits, uint32_t max_hits, crass_b200_hit* d_sorted, void* stream_v) {
    if (!c || !d_found || !d_hits || !d_n_hits || !d_sorted) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    if (((uintptr_t)d_found) & 15) return cbh::fail(CRASS_B200_EINVAL, "found flags must be 16-byte aligned");
    if (n_reads == 0 || max_hits == 0) return 0;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream_v;
    const uint32_t n_chunks = (n_reads + cbk::kRankChunk - 1) / cbk::kRankChunk;
    if (int r = c->d_rank.reserve((size_t)n_chunks * sizeof(uint32_t))) return r;
    uint32_t* chunks = c->d_rank.as<uint32_t>();
    cbk::k_rank_chunks<<<n_chunks, 64, 0, st>>>(d_found, n_reads, chunks);
    cbk::k_rank_scan<<<1, 1024, 0, st>>>(chunks, n_chunks);
    cbk::k_rank_scatter<<<(max_hits + 255) / 256, 256, 0, st>>>(d_found, chunks, d_hits, d_n_hits, max_hits, d_sorted);
    c->launches += 3;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---- K4b/K4c in block form ------------------------------------------------------------------------------
}  // extern "C"
namespace {
template <class Src>
int dedupe_into_block(crass_b200_ctx* c, const Src& src, uint32_t n_slots, uint32_t stride, void* d_block, uint32_t cap, cudaStream_t st) {
    CUDA_TRY(cudaMemsetAsync(d_block, 0, cbk::kTokenBlockHeader, st));
    if (n_slots == 0) return 0;
    uint32_t table = 1024;
    while (table < 2u * n_slots) table <<= 1;
    if (int r = c->d_tok_table.reserve((size_t)table * 2 * sizeof(uint32_t))) return r;
    uint32_t* rep = c->d_tok_table.as<uint32_t>();
    uint32_t* first_read = rep + table;
    CUDA_TRY(cudaMemsetAsync(rep, 0xFF, (size_t)table * 2 * sizeof(uint32_t), st));
    cbk::k_block_dedupe<Src><<<(n_slots + 255) / 256, 256, 0, st>>>(src, n_slots, rep, first_read, table - 1);
    cbk::k_block_compact<Src><<<(table + 255) / 256, 256, 0, st>>>(src, rep, first_read, table, (uint8_t*)d_block, cap, stride);
    c->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return 0;
}
}  // namespace
extern "C" {

size_t crass_b200_token_block_bytes(uint32_t cap, uint32_t stride) { return cbk::kTokenBlockHeader + (size_t)cap * stride; }

int crass_b200_unique_tokens_block_dev(crass_b200_ctx* c, const crass_b200_hit* d_hits, uint32_t n_hits, const void* d_tokens, uint32_t stride,
                                       void* d_block, uint32_t cap, void* stream_v) {
    if (!c || !d_block || (n_hits && (!d_hits || !d_tokens))) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    if (stride < 12 || (stride & 3) || cap == 0) return cbh::fail(CRASS_B200_EINVAL, "token stride must be a multiple of 4, at least 12; cap > 0");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream_v;
    cbk::HitTokens src{d_hits, (const uint8_t*)d_tokens, stride};
    return dedupe_into_block(c, src, n_hits, stride, d_block, cap, st);
}

int crass_b200_comm_unique_id(uint8_t id[128]) {
    if (!id) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    if (!g_nccl.load()) return cbh::fail(CRASS_B200_ECUDA, "libnccl.so.2 could not be loaded (CRASS_B200_NCCL_LIB names another)");
    Nccl::UniqueId u;
    if (const int rc = g_nccl.GetUniqueId(&u)) return cbh::fail(CRASS_B200_ECUDA, std::string("ncclGetUniqueId failed: ") + g_nccl.why(rc));
    memcpy(id, u.internal, 128);
    return 0;
}

int crass_b200_ctx_comm_init(crass_b200_ctx* c, const uint8_t id[128], int rank, int world) {
    if (!c || !id || world < 1 || rank < 0 || rank >= world) return cbh::fail(CRASS_B200_EINVAL, "bad argument");
    if (!g_nccl.load()) return cbh::fail(CRASS_B200_ECUDA, "libnccl.so.2 could not be loaded (CRASS_B200_NCCL_LIB names another)");
    CUDA_TRY(cudaSetDevice(c->device));
    if (c->comm) { g_nccl.CommDestroy(c->comm); c->comm = nullptr; c->comm_world = 0; }
    Nccl::UniqueId u;
    memcpy(u.internal, id, 128);
    if (const int rc = g_nccl.CommInitRank(&c->comm, world, u, rank)) { c->comm = nullptr; return cbh::fail(CRASS_B200_ECUDA, std::string("ncclCommInitRank failed: ") + g_nccl.why(rc)); }
    c->comm_world = world; c->comm_rank = rank;
    return 0;
}

int crass_b200_ctx_comm_world(const crass_b200_ctx* c) { return c ? c->comm_world : 0; }

int crass_b200_exchange_tokens_dev(crass_b200_ctx* c, const crass_b200_hit* d_hits, uint32_t n_hits, const void* d_tokens, uint32_t stride,
                                   void* d_send, uint32_t cap, void* d_recv, uint32_t shard_reads, void* d_merged, uint32_t out_cap,
                                   void* stream_v) {
    if (!c || !d_send || !d_recv || !d_merged) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    if (!c->comm) return cbh::fail(CRASS_B200_EINVAL, "no communicator: call crass_b200_ctx_comm_init first");
    if (int r = crass_b200_unique_tokens_block_dev(c, d_hits, n_hits, d_tokens, stride, d_send, cap, stream_v)) return r;
    const size_t nb = crass_b200_token_block_bytes(cap, stride);
    if (const int rc = g_nccl.AllGather(d_send, d_recv, nb, /*ncclUint8*/ 1, c->comm, (cudaStream_t)stream_v))
        return cbh::fail(CRASS_B200_ECUDA, std::string("ncclAllGather failed: ") + g_nccl.why(rc));
    return crass_b200_merge_token_blocks_dev(c, d_recv, (uint32_t)c->comm_world, cap, stride, shard_reads, d_merged, out_cap, stream_v);
}

int crass_b200_merge_token_blocks_dev(crass_b200_ctx* c, const void* d_blocks, uint32_t n_ranks, uint32_t cap, uint32_t stride,
                                      uint32_t shard_reads, void* d_out_block, uint32_t out_cap, void* stream_v) {
    if (!c || !d_blocks || !d_out_block) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    if (stride < 12 || (stride & 3) || cap == 0 || out_cap == 0 || n_ranks == 0 || n_ranks > 1024)
        return cbh::fail(CRASS_B200_EINVAL, "bad block geometry");
    if ((uint64_t)n_ranks * shard_reads > 0xFFFFFFFFull) return cbh::fail(CRASS_B200_EINVAL, "n_ranks * shard_reads must fit 32 bits");
    if ((uint64_t)n_ranks * cap > 0x7FFFFFFFull) return cbh::fail(CRASS_B200_EINVAL, "n_ranks * cap too large");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream_v;
    cbk::GatheredBlocks src{(const uint8_t*)d_blocks, cap, stride, shard_reads};
    if (int r = dedupe_into_block(c, src, n_ranks * cap, stride, d_out_block, out_cap, st)) return r;
    cbk::k_block_flags<<<1, 1024, 0, st>>>(src, n_ranks, (uint8_t*)d_out_block);
    c->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---- K5 + host passes: the step between the phases from a token block on the device --------------------------------
}  // extern "C"
namespace {
int ensure_ac_on_device(crass_b200_ctx* c, crass_b200_ac* ac);
const uint32_t kClusterDeviceMax = 32768;            // the rank kernel is O(n^2); longer lists take the host passes

int cluster_block_host_passes(crass_b200_ctx* c, const void* d_block, uint32_t cap, uint32_t stride, uint32_t kmer_clust,
                  std::vector<std::string>* patterns, uint32_t* count, uint32_t* flags, cudaStream_t st) {
    patterns->clear();
    // CRASS_B200_TRACE=1: stage times of this call on stderr
    static const bool trace = getenv("CRASS_B200_TRACE") != nullptr;
    auto t_mark = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (!trace) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "cluster_block: %-18s %.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_mark).count());
        t_mark = now;
    };
    const size_t block_bytes = cbk::kTokenBlockHeader + (size_t)cap * stride;
    const size_t max_kmers = (size_t)cap * (stride - 16);
    size_t tab = 1024;
    while (tab < 2 * max_kmers) tab <<= 1;
    if (int r = c->d_cl_order.reserve((size_t)cap * 4)) return r;
    if (int r = c->d_cl_ckeys.reserve((size_t)cap * 4 + 16)) return r;
    if (int r = c->d_cl_koff.reserve(((size_t)cap + 1025) * 4)) return r;
    if (int r = c->d_cl_keys.reserve(max_kmers * 4 + 16)) return r;
    if (int r = c->d_cl_first.reserve(max_kmers * 4 + 16)) return r;
    if (int r = c->d_cl_tab.reserve(tab * 8)) return r;
    if (int r = c->d_cl_info.reserve(16)) return r;
    const uint32_t kStrListCap = 16384;                                        // (DR, k-mer) pairs of string-keyed 11-mers listed by K5
    if (int r = c->d_cl_str.reserve((size_t)kStrListCap * 8)) return r;
    if (int r = c->h_cl_str.reserve((size_t)kStrListCap * 8)) return r;
    if (int r = c->h_cl_block.reserve(block_bytes)) return r;
    if (int r = c->h_cl_order.reserve((size_t)cap * 4)) return r;
    if (int r = c->h_cl_info.reserve(16)) return r;
    cbk::ClusterArrays a{(const uint8_t*)d_block, cap, stride, c->d_cl_order.as<uint32_t>(), c->d_cl_koff.as<uint32_t>(),
                         c->d_cl_keys.as<uint32_t>(), c->d_cl_first.as<uint32_t>(), c->d_cl_tab.as<uint32_t>(),
                         c->d_cl_tab.as<uint32_t>() + tab, (uint32_t)(tab - 1), c->d_cl_info.as<uint32_t>(),
                         c->d_cl_str.as<uint32_t>(), kStrListCap, c->d_cl_ckeys.as<uint32_t>(), kClusterDeviceMax};
    CUDA_TRY(cudaMemsetAsync(c->d_cl_info.p, 0, 16, st));
    CUDA_TRY(cudaMemsetAsync(c->d_cl_tab.p, 0xFF, tab * 8, st));
    cbk::k_cl_gather<<<(cap + 255) / 256, 256, 0, st>>>(a);
    cbk::k_cl_rank<<<(cap + cbk::kClRankThreads) / cbk::kClRankThreads, cbk::kClRankThreads * cbk::kClRankParts, 0, st>>>(a);
    cbk::k_rank_scan<<<1, 1024, 0, st>>>(a.koff, cap + 1, a.info);                  // info[0] = n, written by k_cl_rank
    cbk::k_cl_keys<<<(cap * 32 + cbk::kClKeysThreads - 1) / cbk::kClKeysThreads, cbk::kClKeysThreads, 0, st>>>(a);
    cbk::k_cl_first<<<c->sm_count * 4, 256, 0, st>>>(a);
    c->launches += 5;
    CUDA_TRY(cudaGetLastError());
    cbh::prewake_cluster_workers(1000);                        // the host passes start in a few hundred microseconds
    // two round trips: the sizes first, then exactly the records and array entries that exist
    CUDA_TRY(cudaMemcpyAsync(c->h_cl_block.p, d_block, cbk::kTokenBlockHeader, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(c->h_cl_info.p, a.info, 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    mark("kernels + sizes");
    const uint8_t* hb = c->h_cl_block.as<uint8_t>();
    uint32_t hdr[2];
    memcpy(hdr, hb, sizeof hdr);
    if (count) *count = hdr[0];
    if (flags) *flags = hdr[1];
    if (hdr[1] || hdr[0] > cap || hdr[0] == 0) return 0;                       // overflowed / empty: the caller looks at count and flags
    const uint32_t n = c->h_cl_info.as<uint32_t>()[0], total = c->h_cl_info.as<uint32_t>()[1], n_str = c->h_cl_info.as<uint32_t>()[2];
    if (int r = c->h_cl_keys.reserve((size_t)total * 4 + 16)) return r;
    if (int r = c->h_cl_first.reserve((size_t)total * 4 + 16)) return r;
    CUDA_TRY(cudaMemcpyAsync(c->h_cl_block.as<uint8_t>() + cbk::kTokenBlockHeader, (const uint8_t*)d_block + cbk::kTokenBlockHeader,
                             (size_t)hdr[0] * stride, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(c->h_cl_order.p, a.order, (size_t)hdr[0] * 4, cudaMemcpyDeviceToHost, st));
    if (total && n == hdr[0] && n <= kClusterDeviceMax) {
        CUDA_TRY(cudaMemcpyAsync(c->h_cl_keys.p, a.keys, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(c->h_cl_first.p, a.first, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
        if (n_str && n_str <= kStrListCap) CUDA_TRY(cudaMemcpyAsync(c->h_cl_str.p, a.str_tq, (size_t)n_str * 8, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    mark("copies");
    std::vector<std::string_view> drs;
    bool device_ok = n == hdr[0] && n <= kClusterDeviceMax;
    if (device_ok) {
        drs.reserve(n);
        const uint32_t* order = c->h_cl_order.as<uint32_t>();
        for (uint32_t t = 0; t < n && device_ok; ++t) {
            if (order[t] >= n) { device_ok = false; break; }
            const uint8_t* rec = hb + cbk::kTokenBlockHeader + (size_t)order[t] * stride;
            const uint32_t ln = rec[0] + 6u <= stride ? rec[0] : stride - 6;
            if (!ln) device_ok = false;                                     // an empty token would shift the numbering
            drs.push_back(std::string_view((const char*)rec + 2, ln));
        }
    }
    if (!device_ok) {                                                          // host passes on the copied block
        *patterns = cbh::non_redundant_set(cbh::block_views(hb, cap, stride, nullptr, nullptr), (int)kmer_clust, nullptr, nullptr);
        return 0;
    }
    cbh::ClusterPre pre{c->h_cl_keys.as<uint32_t>(), c->h_cl_first.as<uint32_t>(), total, n_str,
                        n_str && n_str <= kStrListCap ? c->h_cl_str.as<uint32_t>() : nullptr, nullptr};
    // pass D on the device as well: the host walk (pass C) sends the group of every DR, the kernels send back the flags
    size_t ctab = 1024;
    while (ctab < 2 * (size_t)cap) ctab <<= 1;
    int reduce_rc = 0;
    // (measured: no faster than the threaded host reduction at 6.5 k .. 16 k variants -- the chains are walked by pointer
    // chasing -- so it is opt-in: CRASS_B200_CLUSTER=device-reduce)
    const char* dsel = getenv("CRASS_B200_CLUSTER");
    if (dsel && !strcmp(dsel, "device-reduce")) pre.device_reduce = [&](const int* group_of, size_t n_dr, uint8_t* dead) -> bool {
        if (n_dr != n) return false;
        auto body = [&]() -> int {
            if (int r = c->d_cl_group.reserve((size_t)cap * 4)) return r;
            if (int r = c->d_cl_chain.reserve(ctab * 12)) return r;
            if (int r = c->d_cl_next.reserve((size_t)cap * 4)) return r;
            if (int r = c->d_cl_odd.reserve((size_t)cap * 4)) return r;
            if (int r = c->d_cl_dead.reserve((size_t)cap + 16)) return r;
            if (int r = c->h_cl_group.reserve((size_t)cap * 4)) return r;
            if (int r = c->h_cl_dead.reserve((size_t)cap + 16)) return r;
            memcpy(c->h_cl_group.p, group_of, (size_t)n * 4);
            cbk::ReduceArrays ra{a, c->d_cl_group.as<uint32_t>(), c->d_cl_chain.as<unsigned long long>(),
                                 (uint32_t*)(c->d_cl_chain.as<unsigned long long>() + ctab), (uint32_t)(ctab - 1),
                                 c->d_cl_next.as<uint32_t>(), c->d_cl_odd.as<uint32_t>(), c->d_cl_dead.as<uint8_t>()};
            CUDA_TRY(cudaMemcpyAsync(c->d_cl_group.p, c->h_cl_group.p, (size_t)n * 4, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemsetAsync(c->d_cl_chain.p, 0xFF, ctab * 12, st));
            cbk::k_cl_chain<<<(n + 127) / 128, 128, 0, st>>>(ra);
            cbk::k_cl_reduce<<<(n + 127) / 128, 128, 0, st>>>(ra);
            c->launches += 2;
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaMemcpyAsync(c->h_cl_dead.p, ra.dead, n, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            memcpy(dead, c->h_cl_dead.p, n);
            return 0;
        };
        reduce_rc = body();
        return reduce_rc == 0;
    };
    mark("views");
    *patterns = cbh::non_redundant_set(drs, (int)kmer_clust, nullptr, &pre);
    mark("host passes");
    return reduce_rc;
}

// The whole step on the device (K5 + cluster.cuh): one host synchronisation, after which the host holds the pattern set
// (bytes + offsets) and its sizes.  *handled = false: the kernels declined (see the flags in cluster.cuh) or the list is
// longer than the O(n^2) rank kernel should see; the caller takes the host passes.
struct PatternSet {
    std::vector<uint8_t> bytes;
    std::vector<uint32_t> offs;                       // n + 1
    uint32_t n() const { return offs.empty() ? 0u : (uint32_t)offs.size() - 1; }
    void from_strings(const std::vector<std::string>& v) {
        bytes.clear(); offs.assign(1, 0);
        for (const std::string& p : v) { bytes.insert(bytes.end(), p.begin(), p.end()); offs.push_back((uint32_t)bytes.size()); }
    }
};

int cluster_block_device(crass_b200_ctx* c, const void* d_block, uint32_t cap, uint32_t stride, uint32_t kmer_clust,
                         PatternSet* ps, uint32_t* count, uint32_t* flags, bool* handled, cudaStream_t st) {
    *handled = false;
    static const bool trace = getenv("CRASS_B200_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    const size_t max_kmers = (size_t)cap * (stride - 16);
    size_t tab = 1024;
    while (tab < 2 * max_kmers) tab <<= 1;
    const uint32_t kStrListCap = 16384;                // string-keyed 11-mers the device resolves (a power of two)
    if (int r = c->d_cl_order.reserve((size_t)cap * 4)) return r;
    if (int r = c->d_cl_ckeys.reserve((size_t)cap * 4 + 16)) return r;
    if (int r = c->d_cl_koff.reserve(((size_t)cap + 1025) * 4)) return r;
    if (int r = c->d_cl_keys.reserve(max_kmers * 4 + 16)) return r;
    if (int r = c->d_cl_first.reserve(max_kmers * 4 + 16)) return r;
    if (int r = c->d_cl_tab.reserve(tab * 8)) return r;
    if (int r = c->d_cl_info.reserve(cbk::kInfoWords * 4)) return r;
    if (int r = c->d_cl_str.reserve((size_t)16384 * 8)) return r;
    if (int r = c->h_cl_block.reserve(cbk::kTokenBlockHeader + (size_t)cap * stride)) return r;
    if (int r = c->h_cl_info.reserve(cbk::kInfoWords * 4)) return r;
    // the tail's arrays, carved from one allocation
    size_t at = 0;
    auto carve = [&](size_t bytes) { const size_t o = at; at += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_lens = carve((size_t)cap * 4), o_impure = carve((size_t)cap), o_runc = carve(max_kmers * 4 + 16), o_nruns = carve((size_t)cap * 4),
                 o_group = carve((size_t)cap * 4), o_gnum = carve(((size_t)cap + 2) * 4), o_gstart = carve(((size_t)cap + 2) * 4),
                 o_gfill = carve((size_t)cap * 4), o_members = carve((size_t)cap * 4), o_sorted = carve((size_t)cap * 4),
                 o_alive = carve(((size_t)cap + 2) * 4), o_alive1 = carve(((size_t)cap + 2) * 4), o_listed = carve((size_t)cap * 4), o_plen = carve((2 * (size_t)cap + 2) * 4), o_psrc = carve(2 * (size_t)cap * 4),
                 o_pbytes = carve(2 * (size_t)cap * (stride - 6) + 32), o_canon = carve((size_t)kStrListCap * 12), o_packed = carve((size_t)cap * 32),
                 o_str_rep = carve((size_t)kStrListCap * 8), o_str_min = carve((size_t)kStrListCap * 8), o_str_slot = carve((size_t)kStrListCap * 4),
                 o_spacked = carve((size_t)cap * 32), o_slens = carve((size_t)cap * 4);
    if (int r = c->d_cl_tail.reserve(at)) return r;
    uint8_t* tb = c->d_cl_tail.as<uint8_t>();
    cbk::ClusterArrays a{(const uint8_t*)d_block, cap, stride, c->d_cl_order.as<uint32_t>(), c->d_cl_koff.as<uint32_t>(),
                         c->d_cl_keys.as<uint32_t>(), c->d_cl_first.as<uint32_t>(), c->d_cl_tab.as<uint32_t>(),
                         c->d_cl_tab.as<uint32_t>() + tab, (uint32_t)(tab - 1), c->d_cl_info.as<uint32_t>(),
                         c->d_cl_str.as<uint32_t>(), kStrListCap, c->d_cl_ckeys.as<uint32_t>(), kClusterDeviceMax};
    cbk::ClusterTail t{a, (uint32_t*)(tb + o_lens), tb + o_impure, (uint32_t*)(tb + o_runc), (uint32_t*)(tb + o_nruns), (uint32_t*)(tb + o_group),
                       (uint32_t*)(tb + o_gnum), (uint32_t*)(tb + o_gstart), (uint32_t*)(tb + o_gfill), (uint32_t*)(tb + o_members),
                       (uint32_t*)(tb + o_sorted), (uint32_t*)(tb + o_alive), (uint32_t*)(tb + o_alive1), (uint32_t*)(tb + o_listed), (uint32_t*)(tb + o_plen), (uint32_t*)(tb + o_psrc),
                       tb + o_pbytes, tb + o_canon, (uint32_t*)(tb + o_str_rep), (uint32_t*)(tb + o_str_min), (uint32_t*)(tb + o_str_slot),
                       (ulonglong4*)(tb + o_spacked), (uint32_t*)(tb + o_slens), (ulonglong4*)(tb + o_packed), kmer_clust};
    CUDA_TRY(cudaMemsetAsync(c->d_cl_info.p, 0, cbk::kInfoWords * 4, st));
    CUDA_TRY(cudaMemsetAsync(c->d_cl_tab.p, 0xFF, tab * 8, st));
    CUDA_TRY(cudaMemsetAsync(t.plen, 0, (2 * (size_t)cap + 2) * 4, st));
    CUDA_TRY(cudaMemsetAsync(t.str_rep, 0xFF, (size_t)kStrListCap * 16, st));           // str_rep and str_min lie back to back
    const uint32_t per_dr128 = (cap + 127) / 128, per_dr256 = (cap + 1 + 255) / 256, warp_per_dr = (cap + 1 + 3) / 4;
    cbk::k_cl_gather<<<(cap + 255) / 256, 256, 0, st>>>(a);
    cbk::k_cl_rank<<<(cap + cbk::kClRankThreads) / cbk::kClRankThreads, cbk::kClRankThreads * cbk::kClRankParts, 0, st>>>(a);
    cbk::k_rank_scan<<<1, 1024, 0, st>>>(a.koff, cap + 1, a.info);                  // info[0] = n, written by k_cl_rank
    cbk::k_cl_keys<<<(cap * 32 + cbk::kClKeysThreads - 1) / cbk::kClKeysThreads, cbk::kClKeysThreads, 0, st>>>(a);
    cbk::k_cl_first<<<c->sm_count * 4, 256, 0, st>>>(a);
    cbk::k_cl_str_canon<<<kStrListCap / 128, 128, 0, st>>>(t);
    cbk::k_cl_str_insert<<<kStrListCap / 128, 128, 0, st>>>(t);
    cbk::k_cl_str_first<<<kStrListCap / 128, 128, 0, st>>>(t);
    cbk::k_cl_runs<<<per_dr128, 128, 0, st>>>(t);
    cbk::k_cl_walk<<<per_dr128, 128, 0, st>>>(t);
    cbk::k_cl_founders<<<per_dr256, 256, 0, st>>>(t);
    cbk::k_rank_scan<<<1, 1024, 0, st>>>(t.gnum, cap + 1, a.info);
    cbk::k_cl_hist<<<per_dr256, 256, 0, st>>>(t);
    cbk::k_rank_scan<<<1, 1024, 0, st>>>(t.gstart, cap + 1, a.info);
    cbk::k_cl_members<<<per_dr256, 256, 0, st>>>(t);
    cbk::k_cl_group_sort<<<warp_per_dr, 128, 0, st>>>(t);
    if (stride <= 70) {                                     // tokens of at most 64 bases: compared as 2-bit codes, in two rounds
        cbk::k_cl_dead_packed<1><<<warp_per_dr, 128, 0, st>>>(t);
        cbk::k_rank_scan<<<1, 1024, 0, st>>>(t.alive1, cap + 1, a.info);
        cbk::k_cl_dead_list<<<per_dr256, 256, 0, st>>>(t);
        cbk::k_cl_dead_packed<2><<<warp_per_dr, 128, 0, st>>>(t);
        c->launches += 3;
    }
    else cbk::k_cl_dead<<<warp_per_dr, 128, 4 * stride, st>>>(t);
    cbk::k_rank_scan<<<1, 1024, 0, st>>>(t.alive, cap + 1, a.info);
    cbk::k_cl_place<<<per_dr256, 256, 0, st>>>(t);
    cbk::k_rank_scan<<<1, 1024, 0, st>>>(t.plen, 2 * cap + 1, a.info + cbk::kInfoPatterns);
    cbk::k_cl_emit<<<(2 * cap + 127) / 128, 128, 0, st>>>(t);
    c->launches += 21;
    CUDA_TRY(cudaGetLastError());
    // one round trip: header, info record, and as much of the pattern set as the last call needed (twice that, in fact)
    const size_t all_offs = 2 * (size_t)cap + 1, all_bytes = 2 * (size_t)cap * (stride - 6) + 16;
    const size_t spec_offs = std::min<size_t>(all_offs, std::max<size_t>(1024, 2 * (size_t)c->cl_last_patterns) + 1);
    const size_t spec_bytes = std::min<size_t>(all_bytes, std::max<size_t>(32768, 2 * (size_t)c->cl_last_bytes));
    const size_t h_offs_at = 0, h_bytes_at = (all_offs * 4 + 255) & ~(size_t)255;
    if (int r = c->h_cl_pat.reserve(h_bytes_at + all_bytes)) return r;
    CUDA_TRY(cudaMemcpyAsync(c->h_cl_block.p, d_block, cbk::kTokenBlockHeader, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(c->h_cl_info.p, a.info, cbk::kInfoWords * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(c->h_cl_pat.as<uint8_t>() + h_offs_at, t.plen, spec_offs * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(c->h_cl_pat.as<uint8_t>() + h_bytes_at, t.pbytes, spec_bytes, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    uint32_t hdr[2];
    memcpy(hdr, c->h_cl_block.p, sizeof hdr);
    if (count) *count = hdr[0];
    if (flags) *flags = hdr[1];
    if (hdr[1] || hdr[0] > cap || hdr[0] == 0) { *handled = true; ps->bytes.clear(); ps->offs.clear(); return 0; }   // overflowed / empty: the caller looks at count and flags
    const uint32_t* info = c->h_cl_info.as<uint32_t>();
    if (trace) fprintf(stderr, "cluster_block: on the device: %u DRs, %u groups, %u patterns (%u bytes), %u string-keyed k-mers, flags %u, %.3f ms\n",
                       info[cbk::kInfoN], info[cbk::kInfoGroups], info[cbk::kInfoPatterns], info[cbk::kInfoBytes], info[cbk::kInfoStr],
                       info[cbk::kInfoFlags], std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    if (info[cbk::kInfoN] != hdr[0] || info[cbk::kInfoFlags]) return 0;            // declined
    const uint32_t np = info[cbk::kInfoPatterns], nbytes = info[cbk::kInfoBytes];
    if ((size_t)np + 1 > spec_offs || nbytes > spec_bytes) {                       // more than guessed: fetch the rest
        if ((size_t)np + 1 > spec_offs)
            CUDA_TRY(cudaMemcpyAsync(c->h_cl_pat.as<uint8_t>() + h_offs_at + spec_offs * 4, t.plen + spec_offs, ((size_t)np + 1 - spec_offs) * 4, cudaMemcpyDeviceToHost, st));
        if (nbytes > spec_bytes)
            CUDA_TRY(cudaMemcpyAsync(c->h_cl_pat.as<uint8_t>() + h_bytes_at + spec_bytes, t.pbytes + spec_bytes, nbytes - spec_bytes, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    c->cl_last_patterns = np; c->cl_last_bytes = nbytes;
    const uint32_t* ho = (const uint32_t*)(c->h_cl_pat.as<uint8_t>() + h_offs_at);
    ps->offs.assign(ho, ho + np + 1);
    ps->bytes.assign(c->h_cl_pat.as<uint8_t>() + h_bytes_at, c->h_cl_pat.as<uint8_t>() + h_bytes_at + nbytes);
    *handled = true;
    return 0;
}

int cluster_block(crass_b200_ctx* c, const void* d_block, uint32_t cap, uint32_t stride, uint32_t kmer_clust,
                  PatternSet* ps, uint32_t* count, uint32_t* flags, cudaStream_t st) {
    // CRASS_B200_CLUSTER: (unset) everything on the device; "device-passes" = K5's first passes on the device, the rest on the
    // host (the round-1 default); "device-reduce" = that plus pass D on the device
    const char* sel = getenv("CRASS_B200_CLUSTER");
    if (!sel || (strcmp(sel, "device-passes") && strcmp(sel, "device-reduce"))) {
        bool handled = false;
        if (int r = cluster_block_device(c, d_block, cap, stride, kmer_clust, ps, count, flags, &handled, st)) return r;
        if (handled) return 0;
    }
    std::vector<std::string> nr;
    if (int r = cluster_block_host_passes(c, d_block, cap, stride, kmer_clust, &nr, count, flags, st)) return r;
    ps->from_strings(nr);
    return 0;
}
}  // namespace
extern "C" {

int crass_b200_cluster_block_dev(crass_b200_ctx* c, const void* d_block, uint32_t cap, uint32_t stride, uint32_t kmer_clust,
                                 crass_b200_ac** out, uint32_t* count, uint32_t* flags, uint32_t* n_patterns, void* stream_v) {
    if (!c || !d_block || !out) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    if (stride < 28 || (stride & 3) || cap == 0) return cbh::fail(CRASS_B200_EINVAL, "bad block geometry");
    *out = nullptr;
    if (n_patterns) *n_patterns = 0;
    CUDA_TRY(cudaSetDevice(c->device));
    PatternSet ps;
    if (int r = cluster_block(c, d_block, cap, stride, kmer_clust, &ps, count, flags, (cudaStream_t)stream_v)) return r;
    if (n_patterns) *n_patterns = ps.n();
    if (ps.n() == 0) return 0;
    if (int r = crass_b200_ac_build(ps.bytes.data(), ps.offs.data(), ps.n(), out)) return r;
    // the tables are built on this context's device straight away (k_ac_build): the scan that follows finds them ready
    if (int r = ensure_ac_on_device(c, *out)) { crass_b200_ac_destroy(*out); *out = nullptr; return r; }
    return 0;
}

char* crass_b200_cluster_block_patterns_dev(crass_b200_ctx* c, const void* d_block, uint32_t cap, uint32_t stride, uint32_t kmer_clust,
                                            uint32_t* count, uint32_t* flags, uint32_t* n_patterns, void* stream_v) {
    if (!c || !d_block || stride < 28 || (stride & 3) || cap == 0) { cbh::fail(CRASS_B200_EINVAL, "bad argument"); return nullptr; }
    if (cudaSetDevice(c->device) != cudaSuccess) { cbh::fail(CRASS_B200_ECUDA, "cudaSetDevice failed"); return nullptr; }
    PatternSet ps;
    if (cluster_block(c, d_block, cap, stride, kmer_clust, &ps, count, flags, (cudaStream_t)stream_v)) return nullptr;
    if (n_patterns) *n_patterns = ps.n();
    char* text = (char*)malloc(ps.bytes.size() + ps.n() + 1);
    if (!text) { cbh::fail(CRASS_B200_ENOMEM, "malloc"); return nullptr; }
    char* w = text;
    for (uint32_t i = 0; i < ps.n(); ++i) { memcpy(w, ps.bytes.data() + ps.offs[i], ps.offs[i + 1] - ps.offs[i]); w += ps.offs[i + 1] - ps.offs[i]; *w++ = '\n'; }
    *w = 0;
    return text;
}

// ---- K7: consensus DR of DR groups ---------------------------------------------------------------------------------------
int crass_b200_ksw_align(crass_b200_ctx* c, const uint8_t* pool, uint64_t pool_bytes, const crass_b200_ksw_job* jobs, uint32_t n_jobs,
                         crass_b200_ksw_result* out) {
    if (!c || !pool || !jobs || !out) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    static_assert(sizeof(crass_b200_ksw_job) == sizeof(cbk::KswJob) && sizeof(crass_b200_ksw_result) == sizeof(cbk::KswResult), "layout");
    for (uint32_t i = 0; i < n_jobs; ++i)
        if ((uint64_t)jobs[i].q_off + jobs[i].q_len > pool_bytes || (uint64_t)jobs[i].t_off + jobs[i].t_len > pool_bytes)
            return cbh::fail(CRASS_B200_EINVAL, "ksw job reaches past the pool");
    if (n_jobs == 0) return 0;
    CUDA_TRY(cudaSetDevice(c->device));
    size_t at = 0;
    auto carve = [&](size_t bytes) { const size_t o = at; at += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_pool = carve(pool_bytes + 16), o_jobs = carve((size_t)n_jobs * sizeof(cbk::KswJob)), o_res = carve((size_t)n_jobs * sizeof(cbk::KswResult));
    if (int r = c->d_cons.reserve(at)) return r;
    uint8_t* d = c->d_cons.as<uint8_t>();
    CUDA_TRY(cudaMemcpyAsync(d + o_pool, pool, pool_bytes, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(d + o_jobs, jobs, (size_t)n_jobs * sizeof(cbk::KswJob), cudaMemcpyHostToDevice, c->stream));
    cbk::k_ksw_align<<<(n_jobs * 8 + 127) / 128, 128, 0, c->stream>>>(d + o_pool, (const cbk::KswJob*)(d + o_jobs), nullptr, n_jobs, (cbk::KswResult*)(d + o_res));
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, d + o_res, (size_t)n_jobs * sizeof(cbk::KswResult), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

int crass_b200_consensus_groups(crass_b200_ctx* c, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads,
                                const uint32_t* read_dr, const uint32_t* ss_offsets, const uint32_t* ss_pool,
                                const uint8_t* dr_bytes, const uint32_t* dr_offsets, uint32_t n_drs,
                                const uint32_t* group_first_dr, uint32_t n_groups, uint32_t array_len,
                                int32_t* dr_place, uint8_t* dr_flags, int32_t* zone, uint8_t* consensus, float* conservation,
                                int32_t* coverage, uint32_t* status) {
    if (!c || !offsets || !dr_bytes || !dr_offsets || !group_first_dr || !dr_place || !dr_flags || !zone || !consensus || !conservation || !coverage)
        return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    if (n_reads && (!bases || !read_dr || !ss_offsets || !ss_pool)) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    if (n_groups == 0 || n_drs == 0 || array_len == 0 || group_first_dr[0] != 0 || group_first_dr[n_groups] != n_drs)
        return cbh::fail(CRASS_B200_EINVAL, "bad group layout");
    const uint32_t kExtStride = 8 * cbk::kKswMaxSlen;                          // the longest query the alignment kernel takes
    std::vector<uint32_t> dr_group(n_drs);
    for (uint32_t g = 0; g < n_groups; ++g) {
        if (group_first_dr[g] >= group_first_dr[g + 1]) return cbh::fail(CRASS_B200_EINVAL, "empty group");
        for (uint32_t d = group_first_dr[g]; d < group_first_dr[g + 1]; ++d) {
            dr_group[d] = g;
            const uint32_t len = dr_offsets[d + 1] - dr_offsets[d];
            if (len == 0 || len + 4 > kExtStride) return cbh::fail(CRASS_B200_EINVAL, "DR length not in 1..124");
        }
    }
    for (uint32_t i = 0; i < n_reads; ++i) if (read_dr[i] >= n_drs) return cbh::fail(CRASS_B200_EINVAL, "read_dr out of range");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const uint64_t n_bases = n_reads ? offsets[n_reads] : 0;
    const uint32_t n_ss = n_reads ? ss_offsets[n_reads] : 0;
    const uint32_t dr_total = dr_offsets[n_drs];
    const int xtra = 0x80000 | 0x40000 | 5;                                     // KSW_XSTART | KSW_XSUBO | AL_minAlignmentScore (Aligner.h:105-123)
    // round-1 jobs: DR d against its master on both strands (a master's own pair stays empty)
    std::vector<cbk::KswJob> jobs(2 * (size_t)n_drs);
    for (uint32_t d = 0; d < n_drs; ++d) {
        const uint32_t m = group_first_dr[dr_group[d]];
        cbk::KswJob jb{dr_offsets[d], d == m ? 0u : dr_offsets[d + 1] - dr_offsets[d], dr_offsets[m], dr_offsets[m + 1] - dr_offsets[m], 0u, xtra};
        jobs[2 * d] = jb; jb.q_rc = 1; jobs[2 * d + 1] = jb;
    }
    size_t at = 0;
    auto carve = [&](size_t bytes) { const size_t o = at; at += (bytes + 255) & ~(size_t)255; return o; };
    const size_t cols = (size_t)n_groups * array_len;
    const size_t o_bases = carve(n_bases + 16), o_offs = carve(((size_t)n_reads + 1) * 8), o_rdr = carve((size_t)n_reads * 4 + 4),
                 o_sso = carve(((size_t)n_reads + 1) * 4), o_ssp = carve((size_t)n_ss * 4 + 4), o_pool = carve((size_t)dr_total + (size_t)n_drs * kExtStride + 16),
                 o_dro = carve(((size_t)n_drs + 1) * 4), o_drg = carve((size_t)n_drs * 4), o_gf = carve(((size_t)n_groups + 1) * 4),
                 o_place = carve((size_t)n_drs * 4), o_flags = carve(n_drs), o_extr = carve((size_t)n_drs * 4), o_extl = carve((size_t)n_drs * 4),
                 o_jobs = carve(jobs.size() * sizeof(cbk::KswJob)), o_jobs2 = carve(jobs.size() * sizeof(cbk::KswJob)), o_j2dr = carve((size_t)n_drs * 4),
                 o_res = carve(jobs.size() * sizeof(cbk::KswResult)), o_cnt = carve(256),
                 o_cov = carve(cols * 16), o_cons = carve(cols), o_conserv = carve(cols * 4), o_zone = carve((size_t)n_groups * 8), o_good = carve((size_t)n_groups * 4);
    if (int r = c->d_cons.reserve(at)) return r;
    uint8_t* d = c->d_cons.as<uint8_t>();
    const uint64_t zero_off = 0;
    if (n_reads) {
        CUDA_TRY(cudaMemcpyAsync(d + o_bases, bases, n_bases, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d + o_offs, offsets, ((size_t)n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d + o_rdr, read_dr, (size_t)n_reads * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d + o_sso, ss_offsets, ((size_t)n_reads + 1) * 4, cudaMemcpyHostToDevice, st));
        if (n_ss) CUDA_TRY(cudaMemcpyAsync(d + o_ssp, ss_pool, (size_t)n_ss * 4, cudaMemcpyHostToDevice, st));
    } else CUDA_TRY(cudaMemcpyAsync(d + o_offs, &zero_off, 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d + o_pool, dr_bytes, dr_total, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d + o_dro, dr_offsets, ((size_t)n_drs + 1) * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d + o_drg, dr_group.data(), (size_t)n_drs * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d + o_gf, group_first_dr, ((size_t)n_groups + 1) * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d + o_jobs, jobs.data(), jobs.size() * sizeof(cbk::KswJob), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(d + o_cnt, 0, 256, st));
    CUDA_TRY(cudaMemsetAsync(d + o_cov, 0, cols * 16, st));
    CUDA_TRY(cudaMemsetAsync(d + o_good, 0, (size_t)n_groups * 4, st));
    uint32_t* cnt = (uint32_t*)(d + o_cnt);                                    // [0] round-2 slaves, [1] status bits
    cbk::ConsArrays a{d + o_bases, (const uint64_t*)(d + o_offs), n_reads, (const uint32_t*)(d + o_rdr), (const uint32_t*)(d + o_sso),
                      (const uint32_t*)(d + o_ssp), d + o_pool, (const uint32_t*)(d + o_dro), n_drs, (const uint32_t*)(d + o_drg),
                      (const uint32_t*)(d + o_gf), n_groups, array_len, dr_total, kExtStride, (int32_t*)(d + o_place), d + o_flags,
                      (uint32_t*)(d + o_extr), (uint32_t*)(d + o_extl), (cbk::KswJob*)(d + o_jobs2), cnt, (uint32_t*)(d + o_j2dr),
                      (int32_t*)(d + o_cov), d + o_cons, (float*)(d + o_conserv), (int32_t*)(d + o_zone), (uint32_t*)(d + o_good)};
    cbk::KswResult* res = (cbk::KswResult*)(d + o_res);
    const uint32_t n_jobs = 2 * n_drs;
    cbk::k_ksw_align<<<(n_jobs * 8 + 127) / 128, 128, 0, st>>>(d + o_pool, (const cbk::KswJob*)(d + o_jobs), nullptr, n_jobs, res);
    cbk::k_cons_decide<<<(n_drs + 127) / 128, 128, 0, st>>>(a, res, nullptr, nullptr, n_drs, 1);
    // slaves whose forward and reverse scores are equal: extendSlaveDR, then once more (lists built on the device, launches sized for all)
    if (n_reads) cbk::k_cons_ext_find<<<(n_reads + 127) / 128, 128, 0, st>>>(a);
    cbk::k_cons_ext_build<<<(n_drs + 127) / 128, 128, 0, st>>>(a, xtra);
    cbk::k_ksw_align<<<(n_jobs * 8 + 127) / 128, 128, 0, st>>>(d + o_pool, a.jobs2, cnt, n_jobs, res);
    cbk::k_cons_decide<<<(n_drs + 127) / 128, 128, 0, st>>>(a, res, a.job2_dr, cnt, n_drs, 2);
    if (n_reads) cbk::k_cons_place<<<(n_reads * 32 + 127) / 128, 128, 0, st>>>(a, cnt + 1);
    cbk::k_cons_columns<<<(uint32_t)((cols + 255) / 256), 256, 0, st>>>(a);
    cbk::k_cons_zone<<<(n_groups + 63) / 64, 64, 0, st>>>(a);
    c->launches += 7 + (n_reads ? 2 : 0);
    CUDA_TRY(cudaGetLastError());
    uint32_t h_cnt[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(dr_place, d + o_place, (size_t)n_drs * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(dr_flags, d + o_flags, n_drs, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(zone, d + o_zone, (size_t)n_groups * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(consensus, d + o_cons, cols, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(conservation, d + o_conserv, cols * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(coverage, d + o_cov, cols * 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(h_cnt, cnt, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (status) *status = h_cnt[1];
    return 0;
}

// ---- K1 ------------------------------------------------------------------------------------------------
int crass_b200_dr_search_dev(crass_b200_ctx* c, const uint8_t* d_bases, const uint64_t* d_offsets, uint32_t n_reads,
                             uint32_t max_read_len, const crass_b200_params* params, uint8_t* d_found,
                             crass_b200_hit* d_hits, uint32_t hits_cap, uint32_t* d_ss_pool, uint32_t ss_cap,
                             uint32_t* d_counters, void* stream_v) {
    if (!c) return cbh::fail(CRASS_B200_EINVAL, "ctx is NULL");
    if (int r = validate_params(params)) return r;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream_v;
    const cb::Params o = to_core(*params);
    CUDA_TRY(cudaMemsetAsync(d_counters, 0, 4 * sizeof(uint32_t), st));
    c->packed_valid = false;                                   // a 2-bit stream of an earlier batch is stale from here on
    if (n_reads == 0) return 0;
    if (int r = c->d_error.reserve(sizeof(int))) return r;
    CUDA_TRY(cudaMemsetAsync(c->d_error.p, 0, sizeof(int), st));
    if (c->tok_ptr && c->tok_stride < params->high_dr + 2) return cbh::fail(CRASS_B200_EINVAL, "token stride must be at least high_dr + 2");
    cbk::HitSink sink{d_hits, hits_cap, d_ss_pool, ss_cap, d_counters, c->tok_ptr, c->tok_stride};
    const uint32_t cap = cb::ss_capacity(o, max_read_len);
    const int threads = 128;
    // Fast path: 2-bit seed filter over every read + exact search on the few candidates.  Needs the default
    // window geometry (8-mers every 8 bases, seed distances 49..97) and reads that fit a thread's registers.
    const char* force = getenv("CRASS_B200_K1");
    // the 2-bit kernels recode every base anyway; on request they leave that stream in HBM for the singleton scan
    auto keep_stream = [&](uint32_t** keep) -> int {
        *keep = nullptr;
        if (!(c->keep_packed_bases || c->packed_internal)) return 0;
        // the stream's size: the bases of the batch when the caller has said how many (crass_b200_ctx_keep_packed_bases; the
        // resident calls know), else the bound n_reads * max_read_len -- which one long read among millions of short ones
        // makes absurd: above 2^35 bases the stream is not kept (phase 2 then reads the bytes)
        uint64_t bases_bound = (uint64_t)n_reads * max_read_len;
        if (c->packed_internal && c->res_n_bases) bases_bound = c->res_n_bases;
        else if (c->keep_packed_bases > 1) bases_bound = std::min<uint64_t>(bases_bound, c->keep_packed_bases);
        if (bases_bound > ((uint64_t)1 << 35)) { c->packed_valid = false; return 0; }
        const size_t words = (size_t)(bases_bound >> 4) + 256;
        const bool grew = words * sizeof(uint32_t) > c->d_packed.cap;
        if (int r = c->d_packed.reserve(words * sizeof(uint32_t))) return r;
        if (grew) CUDA_TRY(cudaMemsetAsync(c->d_packed.p, 0, c->d_packed.cap, st));       // look-ahead words past the batch are defined
        *keep = c->d_packed.as<uint32_t>();
        c->packed_src = d_bases; c->packed_reads = n_reads; c->packed_valid = true;
        return 0;
    };
    const bool fast_ok = o.window == 8 && cb::window_skips(o) == 8 && o.low_dr + o.low_spacer == 49 &&
                         o.high_dr + o.high_spacer == 97 && max_read_len <= 304 && (((uintptr_t)d_bases) & 15) == 0;
    if (fast_ok && !(force && !strcmp(force, "generic"))) {
        if (!d_found) { if (int r = c->d_found.reserve((size_t)n_reads + 16)) return r; d_found = c->d_found.as<uint8_t>(); }
        if (int r = c->d_cand.reserve(((size_t)n_reads + 16) * sizeof(uint32_t))) return r;
        uint32_t* cand = c->d_cand.as<uint32_t>();
        // The batch is cut into chunks of whole tiles; the filter kernels follow each other on the caller's stream, the exact
        // kernel of chunk i runs on the context's high-priority side stream while the filter works on chunk i+1.  The filter
        // is bound by the integer pipe, the exact kernel by the latency of a few long dependent chains, so side by side
        // they cost little more than the filter alone.  (CRASS_B200_K1_CHUNKS=1: one filter launch, then one exact launch.)
        const uint32_t kChunkAlign = 1024;                                        // a multiple of both tile sizes
        uint32_t n_chunks = 1;                    // measured (tools/k1_variants.py, 10 M x 150 bp): 1 chunk 0.63 ms, 2: 0.71, 4: 0.78
        if (const char* e = getenv("CRASS_B200_K1_CHUNKS")) n_chunks = (uint32_t)std::min(16, std::max(1, atoi(e)));
        uint32_t per_chunk = ((n_reads + n_chunks - 1) / n_chunks + kChunkAlign - 1) / kChunkAlign * kChunkAlign;
        n_chunks = (n_reads + per_chunk - 1) / per_chunk;
        if (int r = c->d_cand_counts.reserve(16 * 4 * sizeof(uint32_t))) return r;    // per chunk: [0] several flagged windows, [1] a single one, [2] queue head
        uint32_t* cand_counts = c->d_cand_counts.as<uint32_t>();
        CUDA_TRY(cudaMemsetAsync(cand_counts, 0, 16 * 4 * sizeof(uint32_t), st));
        uint32_t* keep = nullptr;
        if (int r = keep_stream(&keep)) return r;
        const int se = (int)max_read_len - 58;
        const int nwin = se < 0 ? 1 : se / 16 + 1;
        int* d_err = c->d_error.as<int>();
        const char* fsel = getenv("CRASS_B200_K1F");                               // "tma": CTA tiles staged by bulk copies; default: warp tiles
        const char* esel = getenv("CRASS_B200_K1E");                               // "lockstep": 32 candidates per warp task; "refill": lanes refilled; default: staged
        const bool f_tma = fsel && !strcmp(fsel, "tma");
        const bool e_lockstep = esel && !strcmp(esel, "lockstep");
        const bool e_refill = esel && !strcmp(esel, "refill");
        // CTAs per SM: the filter leaves room for the exact kernel's CTAs when the two run side by side
        int f_ctas = n_chunks > 1 ? 12 : 16, e_ctas = e_lockstep ? 8 : (n_chunks > 1 ? 2 : 4);
        if (const char* e = getenv("CRASS_B200_K1F_CTAS")) f_ctas = std::max(1, atoi(e));
        if (const char* e = getenv("CRASS_B200_K1E_CTAS")) e_ctas = std::max(1, atoi(e));
        uint32_t quorum = cbk::kStageMin, refill_min = cbk::kRefillMin;
        if (const char* e = getenv("CRASS_B200_K1_QUORUM")) quorum = (uint32_t)std::min(32, std::max(1, atoi(e)));
        if (const char* e = getenv("CRASS_B200_K1_REFILL")) refill_min = (uint32_t)std::min(32, std::max(1, atoi(e)));
        const bool piped = n_chunks > 1;
        if (piped) {
            CUDA_TRY(cudaEventRecord(c->ev_fork, st));
            CUDA_TRY(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
        }
#define CB_FAST(NW, NWIN)                                                                                                           \
    do {                                                                                                                            \
        const size_t fsmem = cbk::dr_filter_smem_bytes<NW>();                                                                       \
        const size_t esmem = cbk::dr_exact_smem_bytes<NW>();                                                                        \
        const size_t ssmem = cbk::dr_staged_smem_bytes<NW>();                                                                       \
        int per_sm = 1;                                                                                                             \
        if (f_tma) {                                                                                                                \
            CUDA_TRY(cudaFuncSetAttribute(cbk::k_dr_filter<NW, NWIN, 49, 97>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem)); \
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cbk::k_dr_filter<NW, NWIN, 49, 97>, cbk::kFilterTile, fsmem)); \
        } else {                                                                                                                    \
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cbk::k_dr_filter_warp<NW, NWIN, 49, 97>, cbk::kFwWarps * 32, 0)); \
        }                                                                                                                           \
        per_sm = std::max(1, std::min(per_sm, f_ctas));                                                                             \
        CUDA_TRY(cudaFuncSetAttribute(cbk::k_dr_exact_packed<NW, NWIN, 49, 97>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esmem)); \
        CUDA_TRY(cudaFuncSetAttribute(cbk::k_dr_exact_refill<NW, NWIN, 49, 97>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esmem)); \
        CUDA_TRY(cudaFuncSetAttribute(cbk::k_dr_exact_staged<NW, NWIN, 49, 97>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssmem)); \
        for (uint32_t ch = 0; ch < n_chunks; ++ch) {                                                                                \
            const uint32_t r_begin = ch * per_chunk, r_end = std::min(n_reads, r_begin + per_chunk);                                \
            cbk::CandRegion region{cand, r_begin, r_end, cand_counts + 4 * ch};                                                     \
            if (f_tma) {                                                                                                            \
                const uint32_t t0 = r_begin / cbk::kFilterTile, t1 = (r_end + cbk::kFilterTile - 1) / cbk::kFilterTile;             \
                const int pblocks = (int)std::min<uint32_t>(t1 - t0, (uint32_t)(c->sm_count * per_sm));     /* persistent: one wave */ \
                cbk::k_dr_filter<NW, NWIN, 49, 97><<<pblocks, cbk::kFilterTile, fsmem, st>>>(d_bases, d_offsets, n_reads, t0, t1, d_found, region, keep); \
            } else {                                                                                                                \
                const uint32_t w_tiles = (r_end - r_begin + 31) / 32;                                                               \
                const int pblocks = (int)((w_tiles + cbk::kFwWarps - 1) / cbk::kFwWarps);                    /* one tile per warp */ \
                cbk::k_dr_filter_warp<NW, NWIN, 49, 97><<<pblocks, cbk::kFwWarps * 32, 0, st>>>(d_bases, d_offsets, n_reads, r_begin, r_end, d_found, region, keep); \
            }                                                                                                                       \
            cudaStream_t est = st;                                                                                                  \
            if (piped) {                                                                                                            \
                CUDA_TRY(cudaEventRecord(c->ev_chunk[ch], st));                                                                     \
                CUDA_TRY(cudaStreamWaitEvent(c->side, c->ev_chunk[ch], 0));                                                         \
                est = c->side;                                                                                                      \
            }                                                                                                                       \
            const int eblocks = c->sm_count * e_ctas;                                                                               \
            if (e_lockstep) cbk::k_dr_exact_packed<NW, NWIN, 49, 97><<<eblocks, cbk::kExactThreads, esmem, est>>>(d_bases, d_offsets, n_reads, region, o, d_found, sink, d_err); \
            else if (e_refill) cbk::k_dr_exact_refill<NW, NWIN, 49, 97><<<eblocks, cbk::kExactThreads, esmem, est>>>(d_bases, d_offsets, n_reads, region, o, d_found, sink, d_err); \
            else cbk::k_dr_exact_staged<NW, NWIN, 49, 97><<<eblocks, cbk::kExactThreads, ssmem, est>>>(d_bases, d_offsets, n_reads, region, o, d_found, sink, d_err, quorum, refill_min); \
        }                                                                                                                           \
    } while (0)
        if (max_read_len <= 112) { if (nwin <= 3) CB_FAST(7, 3); else CB_FAST(7, 4); }
        else if (max_read_len <= 160) { if (nwin <= 6) CB_FAST(10, 6); else CB_FAST(10, 7); }
        else if (max_read_len <= 256) { if (nwin <= 12) CB_FAST(16, 12); else CB_FAST(16, 13); }
        else { if (nwin <= 15) CB_FAST(19, 15); else CB_FAST(19, 16); }
#undef CB_FAST
        if (piped) {
            CUDA_TRY(cudaEventRecord(c->ev_join, c->side));
            CUDA_TRY(cudaStreamWaitEvent(st, c->ev_join, 0));
        }
        c->launches += 2 * n_chunks;
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    // Long reads (config 3): one warp per read, 2-bit window flags in shared memory, warp-cooperative candidate handling.
    const bool geometry_ok = o.window == 8 && cb::window_skips(o) == 8 && o.low_dr + o.low_spacer == 49 && o.high_dr + o.high_spacer == 97;
    if (geometry_ok && max_read_len > 304 && max_read_len <= 65536 && (((uintptr_t)d_bases) & 15) == 0 && !(force && !strcmp(force, "generic"))) {
        // Some read is longer than the 304 bases of the thread-per-read path.  Both paths are enqueued and sort it out on the
        // device without a host round trip: in a batch of mostly short reads the warp filter takes everything up to 304 bases
        // and lists the longer reads for the warp-per-read kernel; in a batch of mostly long reads (mean above 400 bases) the
        // filter steps aside and the warp-per-read kernel takes all of it, as before.
        const char* msel = getenv("CRASS_B200_K1_MIXED");
        const bool mixed = !(msel && !strcmp(msel, "0"));
        uint32_t* long_list = nullptr;
        uint32_t* long_count = nullptr;
        if (mixed) {
            if (!d_found) { if (int r = c->d_found.reserve((size_t)n_reads + 16)) return r; d_found = c->d_found.as<uint8_t>(); }
            if (int r = c->d_cand.reserve(((size_t)n_reads + 16) * sizeof(uint32_t))) return r;
            if (int r = c->d_long_list.reserve(((size_t)n_reads + 16) * sizeof(uint32_t))) return r;
            if (int r = c->d_cand_counts.reserve(16 * 4 * sizeof(uint32_t))) return r;
            uint32_t* cand_counts = c->d_cand_counts.as<uint32_t>();
            CUDA_TRY(cudaMemsetAsync(cand_counts, 0, 16 * 4 * sizeof(uint32_t), st));
            long_list = c->d_long_list.as<uint32_t>();
            long_count = cand_counts + 3;
            uint32_t* keep_m = nullptr;
            if (int r = keep_stream(&keep_m)) return r;
            cbk::CandRegion region{c->d_cand.as<uint32_t>(), 0u, n_reads, cand_counts};
            const size_t ssmem = cbk::dr_staged_smem_bytes<19>();
            CUDA_TRY(cudaFuncSetAttribute(cbk::k_dr_exact_staged<19, 16, 49, 97>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssmem));
            const uint32_t w_tiles = (n_reads + 31) / 32;
            cbk::k_dr_filter_warp<19, 16, 49, 97><<<(w_tiles + cbk::kFwWarps - 1) / cbk::kFwWarps, cbk::kFwWarps * 32, 0, st>>>(
                d_bases, d_offsets, n_reads, 0u, n_reads, d_found, region, keep_m, long_list, long_count);
            cbk::k_dr_exact_staged<19, 16, 49, 97><<<c->sm_count * 4, cbk::kExactThreads, ssmem, st>>>(
                d_bases, d_offsets, n_reads, region, o, d_found, sink, c->d_error.as<int>(), cbk::kStageMin, cbk::kRefillMin);
            c->launches += 2;
            CUDA_TRY(cudaGetLastError());
        }
        const uint32_t words = max_read_len / 16 + 32;
        const size_t smem = (size_t)cbk::kLongWarps * words * sizeof(uint32_t);
        CUDA_TRY(cudaFuncSetAttribute(cbk::k_dr_long, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 1;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cbk::k_dr_long, cbk::kLongWarps * 32, smem));
        const uint32_t want_blocks = (n_reads + cbk::kLongWarps - 1) / cbk::kLongWarps;
        const int blocks = (int)std::min<uint32_t>(want_blocks, (uint32_t)(c->sm_count * std::max(per_sm, 1)));
        if (int r = c->d_scratch.reserve((size_t)blocks * cbk::kLongWarps * 2 * cap * sizeof(uint32_t))) return r;
        uint32_t* keep = nullptr;
        if (int r = keep_stream(&keep)) return r;
        if (int r = c->d_ticket.reserve(16)) return r;
        CUDA_TRY(cudaMemsetAsync(c->d_ticket.p, 0, 16, st));
        cbk::k_dr_long<<<blocks, cbk::kLongWarps * 32, smem, st>>>(d_bases, d_offsets, n_reads, o, d_found, sink, c->d_scratch.as<uint32_t>(), cap,
                                                                   c->d_error.as<int>(), words, keep, c->d_ticket.as<uint32_t>(), long_list, long_count);
        c->launches++;
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    if (cap <= 32) {
        int blocks = (int)std::min<uint64_t>(((uint64_t)n_reads + threads - 1) / threads, (uint64_t)c->sm_count * 64);
        cbk::k_dr_search_generic<32><<<blocks, threads, 0, st>>>(d_bases, d_offsets, n_reads, o, d_found, sink, nullptr, 0, c->d_error.as<int>());
    } else {
        int blocks = (int)std::min<uint64_t>(((uint64_t)n_reads + threads - 1) / threads, (uint64_t)c->sm_count * 16);
        if (int r = c->d_scratch.reserve((size_t)blocks * threads * cap * sizeof(uint32_t))) return r;
        cbk::k_dr_search_generic<0><<<blocks, threads, 0, st>>>(d_bases, d_offsets, n_reads, o, d_found, sink, c->d_scratch.as<uint32_t>(), cap, c->d_error.as<int>());
    }
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // extern "C"

namespace {

// shared tail of the two host-facing search calls: run `launch` with growing output buffers until nothing
// overflows, then bring the hits back sorted by read index.
template <class Launch>
int run_with_outputs(crass_b200_ctx* c, uint32_t n_reads, uint64_t n_bases, uint8_t* found_host, Launch launch,
                     crass_b200_hit** hits, uint32_t* n_hits, uint32_t** ss_pool, uint32_t* n_ss_pool, uint32_t token_stride = 0) {
    // first guess: a quarter of the reads hit; later calls start from what this context has seen plus a quarter, so a
    // sample where phase 2 recruits most reads re-runs one kernel once and not once per batch
    uint32_t hits_cap = std::max<uint32_t>(4096, n_reads / 4 + 16);
    uint32_t pool_cap = hits_cap * 6;
    hits_cap = (uint32_t)std::min<uint64_t>((uint64_t)n_reads + 16, std::max<uint64_t>(hits_cap, (c->hits_per_64k * n_reads >> 16) * 5 / 4 + 1024));
    pool_cap = (uint32_t)std::min<uint64_t>(0xFFFFFFF0ull, std::max<uint64_t>(pool_cap, (c->pool_per_64k * n_reads >> 16) * 5 / 4 + 4096));
    if (int r = c->d_counters.reserve(8 * sizeof(uint32_t))) return r;
    if (int r = c->d_found.reserve((size_t)n_reads + 16)) return r;
    // the "reference would throw" flag belongs to this call: a flag left by an earlier search must not fail a scan
    if (c->d_error.p) CUDA_TRY(cudaMemsetAsync(c->d_error.p, 0, sizeof(int), c->stream));
    for (int attempt = 0; attempt < 3; ++attempt) {
        if (int r = c->d_hits.reserve((size_t)hits_cap * sizeof(crass_b200_hit))) return r;
        if (int r = c->d_pool.reserve((size_t)pool_cap * sizeof(uint32_t))) return r;
        if (token_stride) {
            if (int r = c->d_tokens.reserve((size_t)hits_cap * token_stride)) return r;
            c->tok_ptr = c->d_tokens.as<uint8_t>(); c->tok_stride = token_stride;
        }
        const int lr = launch(hits_cap, pool_cap);
        if (token_stride) { c->tok_ptr = nullptr; c->tok_stride = 0; }
        if (lr) return lr;
        CUDA_TRY(cudaMemcpyAsync(c->h_counters, c->d_counters.p, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        if (!c->h_counters[2]) break;
        hits_cap = c->h_counters[0] + 16;
        pool_cap = c->h_counters[1] + 16;
        if (attempt == 2) return cbh::fail(CRASS_B200_EOVERFLOW, "hit buffers overflowed three times");
    }
    int err = 0;
    if (c->d_error.p) {
        CUDA_TRY(cudaMemcpyAsync(&err, c->d_error.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        if (err) return cbh::fail(CRASS_B200_EINVAL, "kernel reported a condition where the reference throws");
    }
    const uint32_t nh = c->h_counters[0], np = c->h_counters[1];
    c->last_candidates = c->h_counters[3];
    if (n_reads) {
        c->hits_per_64k = std::max<uint64_t>(c->hits_per_64k, (((uint64_t)nh << 16) + n_reads - 1) / n_reads);
        c->pool_per_64k = std::max<uint64_t>(c->pool_per_64k, (((uint64_t)np << 16) + n_reads - 1) / n_reads);
    }
    crass_b200_hit* h = (crass_b200_hit*)malloc(sizeof(crass_b200_hit) * (size_t)(nh ? nh : 1));
    uint32_t* p = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(np ? np : 1));
    if (!h || !p) { free(h); free(p); return cbh::fail(CRASS_B200_ENOMEM, "malloc"); }
    if (nh) {                                                  // the hit records travel in read order (what replay consumes)
        if (int r = c->d_hits_sorted.reserve(sizeof(crass_b200_hit) * (size_t)nh)) { free(h); free(p); return r; }
        if (int r = crass_b200_sort_hits_dev(c, c->d_found.as<uint8_t>(), n_reads, c->d_hits.as<crass_b200_hit>(), c->d_counters.as<uint32_t>(), nh,
                                             c->d_hits_sorted.as<crass_b200_hit>(), c->stream)) { free(h); free(p); return r; }
        CUDA_TRY(cudaMemcpyAsync(h, c->d_hits_sorted.p, sizeof(crass_b200_hit) * (size_t)nh, cudaMemcpyDeviceToHost, c->stream));
    }
    if (np) CUDA_TRY(cudaMemcpyAsync(p, c->d_pool.p, sizeof(uint32_t) * (size_t)np, cudaMemcpyDeviceToHost, c->stream));
    if (found_host) CUDA_TRY(cudaMemcpyAsync(found_host, c->d_found.p, n_reads, cudaMemcpyDeviceToHost, c->stream));
    if (token_stride) {
        // K4b: distinct tokens + the first read that carries each, de-duplicated on the device; only those come back
        c->last_dr_list.clear();
        if (nh) {
            const size_t rec_bytes = (size_t)nh * token_stride;
            if (int r = c->d_tok_unique.reserve(rec_bytes + (size_t)nh * sizeof(uint32_t) + 16)) { free(h); free(p); return r; }
            uint8_t* d_ut = c->d_tok_unique.as<uint8_t>();
            uint32_t* d_fr = (uint32_t*)(d_ut + rec_bytes);
            uint32_t* d_cnt = d_fr + nh;
            if (int r = crass_b200_unique_tokens_dev(c, c->d_hits.as<crass_b200_hit>(), nh, c->d_tokens.p, token_stride, d_ut, d_fr, d_cnt, c->stream)) { free(h); free(p); return r; }
            uint32_t nu = 0;
            CUDA_TRY(cudaMemcpyAsync(&nu, d_cnt, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            std::vector<uint8_t> ut((size_t)nu * token_stride);
            std::vector<uint32_t> fr(nu);
            CUDA_TRY(cudaMemcpyAsync(ut.data(), d_ut, ut.size(), cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaMemcpyAsync(fr.data(), d_fr, (size_t)nu * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            std::vector<uint32_t> order(nu);
            for (uint32_t i = 0; i < nu; ++i) order[i] = i;
            std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return fr[a] < fr[b]; });
            for (uint32_t i = 0; i < nu; ++i) {
                const uint8_t* rec = ut.data() + (size_t)order[i] * token_stride;
                c->last_dr_list.append((const char*)rec + 2, rec[0]);
                c->last_dr_list += '\n';
            }
        }
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *hits = h; *n_hits = nh; *ss_pool = p; *n_ss_pool = np;
    (void)n_bases;
    return 0;
}

int upload_batch(crass_b200_ctx* c, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads, uint64_t n_bases) {
    c->res_valid = false; c->res_found_valid = false; c->packed_valid = false;
    if (int r = c->d_bases.reserve(n_bases + 64)) return r;
    if (int r = c->d_offsets.reserve(((size_t)n_reads + 1) * sizeof(uint64_t))) return r;
    CUDA_TRY(cudaMemcpyAsync(c->d_bases.p, bases, n_bases, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->d_offsets.p, offsets, ((size_t)n_reads + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    return 0;
}

uint32_t max_len_of(const uint64_t* offsets, uint32_t n_reads) {
    uint64_t m = 0;
    for (uint32_t i = 0; i < n_reads; ++i) m = std::max<uint64_t>(m, offsets[i + 1] - offsets[i]);
    return (uint32_t)m;
}

}  // namespace

extern "C" {

int crass_b200_dr_search(crass_b200_ctx* c, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads,
                         const crass_b200_params* params, uint8_t* found, crass_b200_hit** hits, uint32_t* n_hits,
                         uint32_t** ss_pool, uint32_t* n_ss_pool) {
    if (!c) return cbh::fail(CRASS_B200_EINVAL, "ctx is NULL");
    if (!hits || !n_hits || !ss_pool || !n_ss_pool) return cbh::fail(CRASS_B200_EINVAL, "output pointer is NULL");
    if (int r = validate_params(params)) return r;
    CUDA_TRY(cudaSetDevice(c->device));
    const uint64_t n_bases = n_reads ? offsets[n_reads] : 0;
    if (n_reads && offsets[0] != 0) return cbh::fail(CRASS_B200_EINVAL, "offsets[0] must be 0");
    const uint32_t max_len = max_len_of(offsets, n_reads);
    if (int r = upload_batch(c, bases, offsets, n_reads, n_bases)) return r;
    auto launch = [&](uint32_t hits_cap, uint32_t pool_cap) {
        return crass_b200_dr_search_dev(c, c->d_bases.as<uint8_t>(), c->d_offsets.as<uint64_t>(), n_reads, max_len, params,
                                        c->d_found.as<uint8_t>(), c->d_hits.as<crass_b200_hit>(), hits_cap,
                                        c->d_pool.as<uint32_t>(), pool_cap, c->d_counters.as<uint32_t>(), c->stream);
    };
    return run_with_outputs(c, n_reads, n_bases, found, launch, hits, n_hits, ss_pool, n_ss_pool);
}

int crass_b200_batch_upload(crass_b200_ctx* c, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads) {
    if (!c || !offsets || (n_reads && !bases)) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    if (n_reads && offsets[0] != 0) return cbh::fail(CRASS_B200_EINVAL, "offsets[0] must be 0");
    const uint64_t n_bases = n_reads ? offsets[n_reads] : 0;
    if (int r = upload_batch(c, bases, offsets, n_reads, n_bases)) return r;
    c->res_n_reads = n_reads; c->res_n_bases = n_bases; c->res_max_len = max_len_of(offsets, n_reads);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->res_valid = true;
    return 0;
}

int crass_b200_dr_search_resident(crass_b200_ctx* c, const crass_b200_params* params, uint8_t* found, crass_b200_hit** hits,
                                  uint32_t* n_hits, uint32_t** ss_pool, uint32_t* n_ss_pool) {
    if (!c || !c->res_valid) return cbh::fail(CRASS_B200_EINVAL, "no resident batch: call crass_b200_batch_upload first");
    if (!hits || !n_hits || !ss_pool || !n_ss_pool) return cbh::fail(CRASS_B200_EINVAL, "output pointer is NULL");
    if (int r = validate_params(params)) return r;
    CUDA_TRY(cudaSetDevice(c->device));
    const uint32_t n_reads = c->res_n_reads;
    auto launch = [&](uint32_t hits_cap, uint32_t pool_cap) {
        return crass_b200_dr_search_dev(c, c->d_bases.as<uint8_t>(), c->d_offsets.as<uint64_t>(), n_reads, c->res_max_len, params,
                                        c->d_found.as<uint8_t>(), c->d_hits.as<crass_b200_hit>(), hits_cap,
                                        c->d_pool.as<uint32_t>(), pool_cap, c->d_counters.as<uint32_t>(), c->stream);
    };
    const uint32_t tstride = (params->high_dr + 2 + 15) & ~15u;
    c->packed_internal = true;                                 // the context owns the batch: keep its 2-bit stream for phase 2
    const int rs = run_with_outputs(c, n_reads, c->res_n_bases, found, launch, hits, n_hits, ss_pool, n_ss_pool, tstride);
    c->packed_internal = false;
    if (rs) return rs;
    if (int r = c->d_found_p1.reserve((size_t)n_reads + 16)) return r;
    if (n_reads) CUDA_TRY(cudaMemcpyAsync(c->d_found_p1.p, c->d_found.p, n_reads, cudaMemcpyDeviceToDevice, c->stream));
    c->res_found_valid = true;
    return 0;
}

// ---- K2 ------------------------------------------------------------------------------------------------
}  // extern "C"
namespace { int ensure_ac_on_device(crass_b200_ctx* c, crass_b200_ac* ac); }
extern "C" {
int crass_b200_ac_build(const uint8_t* pat_bytes, const uint32_t* pat_offsets, uint32_t n_patterns, crass_b200_ac** out) {
    if (!pat_bytes || !pat_offsets || !out) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    cbh::Automaton* a = nullptr;
    if (int r = cbh::build_automaton(pat_bytes, pat_offsets, n_patterns, &a)) return r;
    // crass_b200_ac is a thin wrapper; move the automaton in
    crass_b200_ac* h = new crass_b200_ac();
    h->a = *a;                     // no device copies exist yet, so a plain member-wise copy is safe
    delete a;
    *out = h;
    return 0;
}

void crass_b200_ac_destroy(crass_b200_ac* ac) { delete ac; }
int crass_b200_ac_upload(crass_b200_ctx* c, const crass_b200_ac* ac) {
    if (!c || !ac) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    return ensure_ac_on_device(c, const_cast<crass_b200_ac*>(ac));
}
// introspection of the dense DFA (built on first use: the fast path never needs it)
uint32_t crass_b200_ac_num_states(const crass_b200_ac* ac) {
    if (!ac) return 0;
    cbh::ensure_dfa(&const_cast<crass_b200_ac*>(ac)->a);
    return ac->a.n_states;
}
uint32_t crass_b200_ac_num_symbols(const crass_b200_ac* ac) {
    if (!ac) return 0;
    cbh::ensure_symbols(&const_cast<crass_b200_ac*>(ac)->a);
    return ac->a.n_syms;
}
char* crass_b200_ac_pattern_text(const crass_b200_ac* ac, uint32_t* n_patterns) {
    if (!ac) { cbh::fail(CRASS_B200_EINVAL, "NULL argument"); return nullptr; }
    const cbh::Automaton& a = ac->a;
    if (n_patterns) *n_patterns = a.n_patterns;
    char* text = (char*)malloc((size_t)a.p_offs[a.n_patterns] + a.n_patterns + 1);
    if (!text) { cbh::fail(CRASS_B200_ENOMEM, "malloc"); return nullptr; }
    char* w = text;
    for (uint32_t i = 0; i < a.n_patterns; ++i) {
        const uint32_t len = a.p_offs[i + 1] - a.p_offs[i];
        memcpy(w, a.p_bytes.data() + a.p_offs[i], len);
        w += len;
        *w++ = '\n';
    }
    *w = 0;
    return text;
}
uint64_t crass_b200_ac_table_bytes(const crass_b200_ac* ac) {
    if (!ac) return 0;
    cbh::ensure_dfa(&const_cast<crass_b200_ac*>(ac)->a);
    return (uint64_t)ac->a.table.size() * 4;
}

}  // extern "C"

namespace {
int ensure_ac_on_device(crass_b200_ctx* c, crass_b200_ac* ac) {
    // The device copy lives in context-owned, grow-only buffers (no cudaMalloc/cudaFree per pattern set); it is
    // refreshed whenever a different automaton (by build serial) is used with this context.  Only the pattern bytes and
    // offsets travel: bitmap, key table and start table are built on the device (k_ac_build), unless the matcher carries
    // host-built tables (CRASS_B200_AC_BUILD=host).  Nothing here waits on the host: the work is ordered on c->stream
    // behind the last scan (ev_scan) and scans order themselves behind it (ev_ac_ready).
    cbh::Automaton& a = ac->a;
    if (c->ac_serial == a.serial && a.serial != 0) return 0;
    c->ac_dfa_serial = 0;
    c->ac_serial = 0;
    if (c->scan_recorded) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_scan, 0));
    if (a.q_bits) {                                   // filter + pattern-start table (fast path)
        const size_t bm_words = (size_t)1 << (a.q_bits - 5), bms_words = a.q_bits_small ? (size_t)1 << (a.q_bits_small - 5) : 0;
        const size_t t_words = (size_t)1 << a.q_table_bits, s_words = (size_t)1 << a.s_bits;
        if (int r = c->d_ac_bitmap.reserve(bm_words * 4)) return r;
        if (int r = c->d_ac_bitmap_small.reserve(bms_words * 4 + 16)) return r;
        if (int r = c->d_ac_keys.reserve(t_words * 4)) return r;
        if (int r = c->d_ac_skeys.reserve(s_words * 4)) return r;
        if (int r = c->d_ac_shead.reserve(s_words * 4)) return r;
        if (int r = c->d_ac_pnext.reserve((size_t)a.n_patterns * 4 + 16)) return r;
        if (int r = c->d_ac_ones.reserve(16)) return r;
        const uint32_t ones_host[2] = {a.q_has_ones, a.s_ones_head};
        struct Up { DevBuf* d; const void* h; size_t bytes; } ups[] = {
            {&c->d_ac_poffs, a.p_offs.data(), a.p_offs.size() * sizeof(uint32_t)},
            {&c->d_ac_pbytes, a.p_bytes.data(), a.p_bytes.size()},
            {&c->d_ac_bitmap, a.q_bitmap.data(), a.q_bitmap.size() * sizeof(uint32_t)},
            {&c->d_ac_bitmap_small, a.q_bitmap_small.data(), a.q_bitmap_small.size() * sizeof(uint32_t)},
            {&c->d_ac_keys, a.q_keys.data(), a.q_keys.size() * sizeof(uint32_t)},
            {&c->d_ac_skeys, a.s_keys.data(), a.s_keys.size() * sizeof(uint32_t)},
            {&c->d_ac_shead, a.s_head.data(), a.s_head.size() * sizeof(uint32_t)},
            {&c->d_ac_pnext, a.p_next.data(), a.p_next.size() * sizeof(uint32_t)},
            {&c->d_ac_ones, ones_host, sizeof ones_host},
        };
        const size_t n_up = a.tables_on_host ? sizeof ups / sizeof ups[0] : 2;
        // staged through one page-locked buffer: copies from pageable vectors are synchronous and slow for small tables.
        // The buffer may still be feeding the previous matcher's copies, hence the event before it is written again.
        size_t total = 0;
        for (size_t i = 0; i < n_up; ++i) total += (ups[i].bytes + 63) & ~(size_t)63;
        total += 64;                                  // the initial value of d_ac_ones
        if (c->stage_busy) { CUDA_TRY(cudaEventSynchronize(c->ev_ac_ready)); c->stage_busy = false; }
        if (int r = c->h_ac_stage.reserve(total)) return r;
        size_t at = 0;
        for (size_t i = 0; i < n_up; ++i) {
            const Up& u = ups[i];
            if (!u.bytes) continue;
            if (int r = u.d->reserve(u.bytes + 16)) return r;
            memcpy(c->h_ac_stage.as<uint8_t>() + at, u.h, u.bytes);
            CUDA_TRY(cudaMemcpyAsync(u.d->p, c->h_ac_stage.as<uint8_t>() + at, u.bytes, cudaMemcpyHostToDevice, c->stream));
            at += (u.bytes + 63) & ~(size_t)63;
        }
        if (!a.tables_on_host) {
            CUDA_TRY(cudaMemsetAsync(c->d_ac_bitmap.p, 0, bm_words * 4, c->stream));
            if (bms_words) CUDA_TRY(cudaMemsetAsync(c->d_ac_bitmap_small.p, 0, bms_words * 4, c->stream));
            CUDA_TRY(cudaMemsetAsync(c->d_ac_keys.p, 0xFF, t_words * 4, c->stream));
            CUDA_TRY(cudaMemsetAsync(c->d_ac_skeys.p, 0xFF, s_words * 4, c->stream));
            CUDA_TRY(cudaMemsetAsync(c->d_ac_shead.p, 0xFF, s_words * 4, c->stream));
            const uint32_t ones_init[2] = {0u, 0xFFFFFFFFu};
            memcpy(c->h_ac_stage.as<uint8_t>() + at, ones_init, sizeof ones_init);      // 
            CUDA_TRY(cudaMemcpyAsync(c->d_ac_ones.p, c->h_ac_stage.as<uint8_t>() + at, sizeof ones_init, cudaMemcpyHostToDevice, c->stream));
            cbk::MatcherTables m{c->d_ac_pbytes.as<uint8_t>(), c->d_ac_poffs.as<uint32_t>(), a.n_patterns,
                                 c->d_ac_bitmap.as<uint32_t>(), a.q_bits, a.q_hashes, c->d_ac_bitmap_small.as<uint32_t>(), a.q_bits_small,
                                 c->d_ac_keys.as<uint32_t>(), a.q_table_bits, c->d_ac_skeys.as<uint32_t>(), c->d_ac_shead.as<uint32_t>(),
                                 a.s_bits, c->d_ac_pnext.as<uint32_t>(), c->d_ac_ones.as<uint32_t>()};
            cbk::k_ac_build<<<(a.n_patterns * 8 + 255) / 256, 256, 0, c->stream>>>(m);
            c->launches++;
            CUDA_TRY(cudaGetLastError());
        }
        c->stage_busy = true;
    }
    CUDA_TRY(cudaEventRecord(c->ev_ac_ready, c->stream));
    c->ac_serial = a.serial;
    return 0;
}

int ensure_dfa_on_device(crass_b200_ctx* c, crass_b200_ac* ac) {      // generic K2 path only
    cbh::Automaton& a = ac->a;
    if (c->ac_dfa_serial == a.serial && a.serial != 0) return 0;
    cbh::ensure_dfa(&a);
    if (c->scan_recorded) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_scan, 0));
    if (int r = c->d_ac_table.reserve(a.table.size() * sizeof(uint32_t))) return r;
    if (int r = c->d_ac_symv.reserve(256)) return r;
    CUDA_TRY(cudaMemcpyAsync(c->d_ac_table.p, a.table.data(), a.table.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->d_ac_symv.p, a.symv, 256, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->ac_dfa_serial = a.serial;
    return 0;
}
}  // namespace

extern "C" {

int crass_b200_ac_scan_dev(crass_b200_ctx* c, const crass_b200_ac* ac_c, const uint8_t* d_bases, const uint64_t* d_offsets,
                           uint32_t n_reads, uint32_t max_read_len, const uint8_t* d_skip, uint8_t* d_found,
                           crass_b200_hit* d_hits, uint32_t hits_cap, uint32_t* d_ss_pool, uint32_t ss_cap,
                           uint32_t* d_counters, void* stream_v) {
    (void)max_read_len;
    if (!c || !ac_c) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    crass_b200_ac* ac = const_cast<crass_b200_ac*>(ac_c);
    CUDA_TRY(cudaSetDevice(c->device));
    if (int r = ensure_ac_on_device(c, ac)) return r;
    cudaStream_t st = (cudaStream_t)stream_v;
    if (st != c->stream) CUDA_TRY(cudaStreamWaitEvent(st, c->ev_ac_ready, 0));            // the tables are built on c->stream
    CUDA_TRY(cudaMemsetAsync(d_counters, 0, 4 * sizeof(uint32_t), st));
    if (n_reads == 0) return 0;
    cbk::HitSink sink{d_hits, hits_cap, d_ss_pool, ss_cap, d_counters, nullptr, 0};
    // a later matcher's tables must not overwrite these before the scan is through
    auto scan_enqueued = [&]() -> int { CUDA_TRY(cudaEventRecord(c->ev_scan, st)); c->scan_recorded = true; return 0; };
    // Fast path: 16-mer q-gram filter over every read + automaton walk over the few candidates.
    const char* force = getenv("CRASS_B200_K2");
    if (ac->a.q_bits && max_read_len <= 304 && (((uintptr_t)d_bases) & 15) == 0 && !(force && !strcmp(force, "generic"))) {
        if (!d_found) { if (int r = c->d_found.reserve((size_t)n_reads + 16)) return r; d_found = c->d_found.as<uint8_t>(); }
        if (int r = c->d_cand.reserve(((size_t)n_reads + 16) * sizeof(uint32_t))) return r;
        if (int r = c->d_cand_mask.reserve(((size_t)n_reads + 16) * sizeof(uint64_t))) return r;
        uint32_t* cand = c->d_cand.as<uint32_t>();
        uint64_t* cmask = c->d_cand_mask.as<uint64_t>();
        cbk::QgramFilter q{c->d_ac_bitmap.as<uint32_t>(), c->d_ac_keys.as<uint32_t>(), ac->a.q_bits, ac->a.q_table_bits, ac->a.q_hashes, c->d_ac_ones.as<uint32_t>()};
        const uint32_t n_tiles = (n_reads + cbk::kAcTile - 1) / cbk::kAcTile;
        const size_t bm_bytes = ((size_t)1 << ac->a.q_bits) / 8;
#define CB_ACF(NW)                                                                                                              \
    do {                                                                                                                        \
        const size_t smem = bm_bytes + (size_t)(cbk::kAcTile * NW + NW + 8) * sizeof(uint32_t);                                 \
        CUDA_TRY(cudaFuncSetAttribute(cbk::k_ac_filter<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
        const int per_sm = std::max<int>(1, (int)((size_t)220 * 1024 / smem));                                                  \
        const int fblocks = (int)std::min<uint32_t>(n_tiles, (uint32_t)(c->sm_count * std::min(per_sm, 8)));                    \
        cbk::k_ac_filter<NW><<<fblocks, cbk::kAcTile, smem, st>>>(d_bases, d_offsets, n_reads, q, d_skip, d_found, cand, cmask, d_counters); \
    } while (0)
        // the 2-bit stream the direct-repeat search of this very batch left behind: a quarter of the bytes, no recoding
        const char* fsel = getenv("CRASS_B200_K2F");
        const bool use_packed = c->packed_valid && c->packed_src == (const void*)d_bases && c->packed_reads == n_reads &&
                                (c->keep_packed_bases || c->packed_internal) && !(fsel && !strcmp(fsel, "bytes"));
        // the 2-bit-stream form takes the folded (half-size) bitmap when the matcher carries one: three CTAs per SM instead of two
        cbk::QgramFilter qp = q;
        if (ac->a.q_bits_small) { qp.bitmap = c->d_ac_bitmap_small.as<uint32_t>(); qp.bits = ac->a.q_bits_small; }
        const size_t bm_bytes_p = ((size_t)1 << qp.bits) / 8;
#define CB_ACP(NW)                                                                                                              \
    do {                                                                                                                        \
        const size_t smem = bm_bytes_p + 2 * (size_t)cbk::ac_packed_words<NW>() * sizeof(uint32_t);                             \
        CUDA_TRY(cudaFuncSetAttribute(cbk::k_ac_filter_packed<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        int per_sm = 1;                                                                                                         \
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cbk::k_ac_filter_packed<NW>, cbk::kAcPackedTile, smem)); \
        const uint32_t p_tiles = (n_reads + cbk::kAcPackedTile - 1) / cbk::kAcPackedTile;                                       \
        const int pblocks = (int)std::min<uint32_t>(p_tiles, (uint32_t)(c->sm_count * std::max(per_sm, 1)));                    \
        cbk::k_ac_filter_packed<NW><<<pblocks, cbk::kAcPackedTile, smem, st>>>(c->d_packed.as<uint32_t>(), d_offsets, n_reads, qp, d_skip, d_found, cand, cmask, d_counters); \
    } while (0)
        if (use_packed) {
            if (max_read_len <= 112) CB_ACP(7);
            else if (max_read_len <= 160) CB_ACP(10);
            else if (max_read_len <= 256) CB_ACP(16);
            else CB_ACP(19);
        }
        else if (max_read_len <= 112) CB_ACF(7);
        else if (max_read_len <= 160) CB_ACF(10);
        else if (max_read_len <= 256) CB_ACF(16);
        else CB_ACF(19);
#undef CB_ACF
#undef CB_ACP
        CUDA_TRY(cudaGetLastError());
        cbk::PatternStarts ps{c->d_ac_pbytes.as<uint8_t>(), c->d_ac_poffs.as<uint32_t>(), c->d_ac_skeys.as<uint32_t>(),
                              c->d_ac_shead.as<uint32_t>(), c->d_ac_pnext.as<uint32_t>(), ac->a.s_bits, c->d_ac_ones.as<uint32_t>(),
                              ac->a.min_pattern_len};
        // one warp per candidate: a single thread walking a read's dependent table probes is latency-bound even at 150 bp
        // (CRASS_B200_K2V=list selects the thread-per-candidate form for comparison)
        const char* vsel = getenv("CRASS_B200_K2V");
        if (vsel && !strcmp(vsel, "list")) cbk::k_ac_verify_list<<<c->sm_count * 8, 128, 0, st>>>(d_bases, d_offsets, cand, ps, d_found, sink);
        else if (!(vsel && !strcmp(vsel, "warp"))) {
            // default: only the starts the filter's 16-mer hits allow (CRASS_B200_K2V=warp: all starts, for comparison)
            int per_sm = 16;
            if (const char* e = getenv("CRASS_B200_K2V_CTAS")) per_sm = std::max(1, atoi(e));
            cbk::k_ac_verify_mask<<<c->sm_count * per_sm, 128, 0, st>>>(d_bases, d_offsets, cand, cmask, ps, d_found, sink);
        }
        else {
            // 28 registers per thread: 16 CTAs of 4 warps fill an SM; the kernel is a chain of dependent L2 probes, so warps in
            // flight are what it runs on (CRASS_B200_K2V_CTAS = CTAs per SM, for measurements)
            int per_sm = 16;
            if (const char* e = getenv("CRASS_B200_K2V_CTAS")) per_sm = std::max(1, atoi(e));
            cbk::k_ac_verify_warp<<<c->sm_count * per_sm, 128, 0, st>>>(d_bases, d_offsets, cand, nullptr, ps, d_found, sink);
        }
        c->launches += 2;
        CUDA_TRY(cudaGetLastError());
        return scan_enqueued();
    }
    if (ac->a.q_bits && max_read_len > 304 && (((uintptr_t)d_bases) & 15) == 0 && !(force && !strcmp(force, "generic"))) {
        // long reads: warp-per-read q-gram filter, then the same verify kernel over the candidates
        if (!d_found) { if (int r = c->d_found.reserve((size_t)n_reads + 16)) return r; d_found = c->d_found.as<uint8_t>(); }
        if (int r = c->d_cand.reserve(((size_t)n_reads + 16) * sizeof(uint32_t))) return r;
        uint32_t* cand = c->d_cand.as<uint32_t>();
        if (int r = c->d_cand_mask.reserve(((size_t)n_reads + 16) * sizeof(uint32_t))) return r;
        uint32_t* cand_from = getenv("CRASS_B200_K2V_FROM0") ? nullptr : c->d_cand_mask.as<uint32_t>();   // per candidate: where the verification may start
        cbk::QgramFilter q{c->d_ac_bitmap.as<uint32_t>(), c->d_ac_keys.as<uint32_t>(), ac->a.q_bits, ac->a.q_table_bits, ac->a.q_hashes, c->d_ac_ones.as<uint32_t>()};
        const size_t smem = ((size_t)1 << ac->a.q_bits) / 8;
        const char* fsel = getenv("CRASS_B200_K2F");
        const bool use_packed = c->packed_valid && c->packed_src == (const void*)d_bases && c->packed_reads == n_reads &&
                                (c->keep_packed_bases || c->packed_internal) && !(fsel && !strcmp(fsel, "bytes"));
        CUDA_TRY(cudaFuncSetAttribute(cbk::k_ac_filter_long<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(cbk::k_ac_filter_long<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 1;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cbk::k_ac_filter_long<false>, cbk::kAcLongThreads, smem));
        const uint32_t want_blocks = (n_reads + (cbk::kAcLongThreads / 32) - 1) / (cbk::kAcLongThreads / 32);
        const int blocks = (int)std::min<uint32_t>(want_blocks, (uint32_t)(c->sm_count * std::max(per_sm, 1)));
        if (use_packed) cbk::k_ac_filter_long<true><<<blocks, cbk::kAcLongThreads, smem, st>>>((const uint8_t*)c->d_packed.p, d_offsets, n_reads, q, d_skip, d_found, cand, c->d_cand_mask.as<uint32_t>(), d_counters);
        else cbk::k_ac_filter_long<false><<<blocks, cbk::kAcLongThreads, smem, st>>>(d_bases, d_offsets, n_reads, q, d_skip, d_found, cand, c->d_cand_mask.as<uint32_t>(), d_counters);
        CUDA_TRY(cudaGetLastError());
        cbk::PatternStarts ps{c->d_ac_pbytes.as<uint8_t>(), c->d_ac_poffs.as<uint32_t>(), c->d_ac_skeys.as<uint32_t>(),
                              c->d_ac_shead.as<uint32_t>(), c->d_ac_pnext.as<uint32_t>(), ac->a.s_bits, c->d_ac_ones.as<uint32_t>(),
                              ac->a.min_pattern_len};
        cbk::k_ac_verify_warp<<<c->sm_count * 16, 128, 0, st>>>(d_bases, d_offsets, cand, cand_from, ps, d_found, sink);      // 16 CTAs per SM
        c->launches += 2;
        CUDA_TRY(cudaGetLastError());
        return scan_enqueued();
    }
    if (int r = ensure_dfa_on_device(c, ac)) return r;
    uint32_t stride_log2 = 0;
    while ((1u << stride_log2) < ac->a.stride) ++stride_log2;
    const int threads = 256;
    int blocks = (int)std::min<uint64_t>(((uint64_t)n_reads + threads - 1) / threads, (uint64_t)c->sm_count * 32);
    cbk::k_ac_scan_generic<<<blocks, threads, 0, st>>>(d_bases, d_offsets, n_reads, c->d_ac_table.as<uint32_t>(), stride_log2,
                                                       c->d_ac_symv.as<uint8_t>(), d_skip, d_found, sink);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return scan_enqueued();
}

int crass_b200_ac_scan(crass_b200_ctx* c, const crass_b200_ac* ac, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads,
                       const uint8_t* skip, uint8_t* found, crass_b200_hit** hits, uint32_t* n_hits, uint32_t** ss_pool,
                       uint32_t* n_ss_pool) {
    if (!c || !ac) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    if (!hits || !n_hits || !ss_pool || !n_ss_pool) return cbh::fail(CRASS_B200_EINVAL, "output pointer is NULL");
    CUDA_TRY(cudaSetDevice(c->device));
    const uint64_t n_bases = n_reads ? offsets[n_reads] : 0;
    const uint32_t max_len = max_len_of(offsets, n_reads);
    if (int r = upload_batch(c, bases, offsets, n_reads, n_bases)) return r;
    if (skip) {
        if (int r = c->d_skip.reserve((size_t)n_reads + 16)) return r;
        CUDA_TRY(cudaMemcpyAsync(c->d_skip.p, skip, n_reads, cudaMemcpyHostToDevice, c->stream));
    }
    auto launch = [&](uint32_t hits_cap, uint32_t pool_cap) {
        return crass_b200_ac_scan_dev(c, ac, c->d_bases.as<uint8_t>(), c->d_offsets.as<uint64_t>(), n_reads, max_len,
                                      skip ? c->d_skip.as<uint8_t>() : nullptr, c->d_found.as<uint8_t>(),
                                      c->d_hits.as<crass_b200_hit>(), hits_cap, c->d_pool.as<uint32_t>(), pool_cap,
                                      c->d_counters.as<uint32_t>(), c->stream);
    };
    return run_with_outputs(c, n_reads, n_bases, found, launch, hits, n_hits, ss_pool, n_ss_pool);
}

int crass_b200_ac_scan_resident(crass_b200_ctx* c, const crass_b200_ac* ac, int skip_found, uint8_t* found, crass_b200_hit** hits,
                                uint32_t* n_hits, uint32_t** ss_pool, uint32_t* n_ss_pool) {
    if (!c || !ac) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    if (!c->res_valid) return cbh::fail(CRASS_B200_EINVAL, "no resident batch: call crass_b200_batch_upload first");
    if (skip_found && !c->res_found_valid) return cbh::fail(CRASS_B200_EINVAL, "skip_found needs a preceding crass_b200_dr_search_resident");
    if (!hits || !n_hits || !ss_pool || !n_ss_pool) return cbh::fail(CRASS_B200_EINVAL, "output pointer is NULL");
    CUDA_TRY(cudaSetDevice(c->device));
    const uint32_t n_reads = c->res_n_reads;
    auto launch = [&](uint32_t hits_cap, uint32_t pool_cap) {
        return crass_b200_ac_scan_dev(c, ac, c->d_bases.as<uint8_t>(), c->d_offsets.as<uint64_t>(), n_reads, c->res_max_len,
                                      skip_found ? c->d_found_p1.as<uint8_t>() : nullptr, c->d_found.as<uint8_t>(),
                                      c->d_hits.as<crass_b200_hit>(), hits_cap, c->d_pool.as<uint32_t>(), pool_cap,
                                      c->d_counters.as<uint32_t>(), c->stream);
    };
    c->packed_internal = true;
    const int rs = run_with_outputs(c, n_reads, c->res_n_bases, found, launch, hits, n_hits, ss_pool, n_ss_pool);
    c->packed_internal = false;
    return rs;
}

// ---- K3 ------------------------------------------------------------------------------------------------
int crass_b200_edit_distance_batch(crass_b200_ctx* c, const uint8_t* bytes, uint64_t n_bytes, const uint32_t* a_off,
                                   const uint32_t* a_len, const uint32_t* b_off, const uint32_t* b_len, uint32_t n_pairs,
                                   int32_t* out_dist, float* out_sim) {
    if (!c) return cbh::fail(CRASS_B200_EINVAL, "ctx is NULL");
    CUDA_TRY(cudaSetDevice(c->device));
    if (n_pairs == 0) return 0;
    for (uint32_t i = 0; i < n_pairs; ++i) {
        if (a_len[i] > (uint32_t)cb::kMaxEdit || b_len[i] > (uint32_t)cb::kMaxEdit) return cbh::fail(CRASS_B200_EINVAL, "string longer than 255");
        if ((uint64_t)a_off[i] + a_len[i] > n_bytes || (uint64_t)b_off[i] + b_len[i] > n_bytes) return cbh::fail(CRASS_B200_EINVAL, "pair out of range");
    }
    const size_t idx_bytes = (size_t)n_pairs * sizeof(uint32_t);
    if (int r = c->d_bases.reserve(n_bytes + 64)) return r;
    if (int r = c->d_misc.reserve(idx_bytes * 6)) return r;
    uint8_t* m = c->d_misc.as<uint8_t>();
    CUDA_TRY(cudaMemcpyAsync(c->d_bases.p, bytes, n_bytes, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(m + 0 * idx_bytes, a_off, idx_bytes, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(m + 1 * idx_bytes, a_len, idx_bytes, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(m + 2 * idx_bytes, b_off, idx_bytes, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(m + 3 * idx_bytes, b_len, idx_bytes, cudaMemcpyHostToDevice, c->stream));
    const int threads = 128;
    const int blocks = (int)((n_pairs + threads - 1) / threads);
    cbk::k_edit_distance<<<blocks, threads, 0, c->stream>>>(c->d_bases.as<uint8_t>(), (uint32_t*)(m + 0 * idx_bytes), (uint32_t*)(m + 1 * idx_bytes),
                                                             (uint32_t*)(m + 2 * idx_bytes), (uint32_t*)(m + 3 * idx_bytes), n_pairs,
                                                             (int32_t*)(m + 4 * idx_bytes), (float*)(m + 5 * idx_bytes));
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out_dist, m + 4 * idx_bytes, idx_bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(out_sim, m + 5 * idx_bytes, idx_bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---- K6: partial-DR recovery ----------------------------------------------------------------------------------
int crass_b200_update_start_stops_dev(crass_b200_ctx* c, const uint8_t* d_bases, const uint64_t* d_offsets,
                                      const uint8_t* d_dr_bytes, const uint32_t* d_dr_offsets,
                                      const crass_b200_uss_job* d_jobs, uint32_t n_jobs, const uint32_t* d_ss_in,
                                      uint32_t low_spacer, uint32_t* d_ss_out, uint32_t* d_n_out, uint8_t* d_status, void* stream) {
    if (!c) return cbh::fail(CRASS_B200_EINVAL, "ctx is NULL");
    if (!d_bases || !d_offsets || !d_dr_bytes || !d_dr_offsets || !d_jobs || !d_ss_in || !d_ss_out || !d_n_out || !d_status)
        return cbh::fail(CRASS_B200_EINVAL, "update_start_stops: NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    if (n_jobs == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const char* sel = getenv("CRASS_B200_K6");
    if (sel && !strcmp(sel, "thread")) {                              // one thread per read: kept for comparison
        const int blocks = (int)((n_jobs + cbk::kUssThreads - 1) / cbk::kUssThreads);
        cbk::k_update_start_stops_thread<<<blocks, cbk::kUssThreads, 0, st>>>(d_bases, d_offsets, d_dr_bytes, d_dr_offsets, d_jobs, n_jobs,
                                                                              d_ss_in, low_spacer, d_ss_out, d_n_out, d_status);
    } else {                                                          // one warp per read
        const int blocks = (int)((n_jobs + cbk::kUssWarps - 1) / cbk::kUssWarps);
        cbk::k_update_start_stops<<<blocks, cbk::kUssWarps * 32, 0, st>>>(d_bases, d_offsets, d_dr_bytes, d_dr_offsets, d_jobs, n_jobs,
                                                                          d_ss_in, low_spacer, d_ss_out, d_n_out, d_status);
    }
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int crass_b200_update_start_stops(crass_b200_ctx* c, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads,
                                  const uint8_t* dr_bytes, const uint32_t* dr_offsets, uint32_t n_drs,
                                  const crass_b200_uss_job* jobs, uint32_t n_jobs, const uint32_t* ss_in, uint32_t n_ss_in,
                                  uint32_t low_spacer, uint32_t* ss_out, uint32_t ss_out_cap, uint32_t* n_out, uint8_t* status) {
    if (!c) return cbh::fail(CRASS_B200_EINVAL, "ctx is NULL");
    if (n_jobs == 0) return 0;
    if (!bases || !offsets || !dr_bytes || !dr_offsets || !jobs || !ss_in || !ss_out || !n_out || !status)
        return cbh::fail(CRASS_B200_EINVAL, "update_start_stops: NULL argument");
    for (uint32_t i = 0; i < n_jobs; ++i) {
        const crass_b200_uss_job& j = jobs[i];
        if (j.read >= n_reads || j.dr >= n_drs) return cbh::fail(CRASS_B200_EINVAL, "update_start_stops: job names a read or DR that does not exist");
        if ((uint64_t)j.ss_offset + j.n_ss > n_ss_in) return cbh::fail(CRASS_B200_EINVAL, "update_start_stops: start/stop list out of range");
        if ((uint64_t)j.out_offset + j.n_ss + 4 > ss_out_cap) return cbh::fail(CRASS_B200_EINVAL, "update_start_stops: ss_out too small (n_ss + 4 entries per job)");
    }
    CUDA_TRY(cudaSetDevice(c->device));
    const uint64_t n_bases = offsets[n_reads];
    const size_t dr_total = dr_offsets[n_drs];
    auto up4 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const size_t o_dro = up4(dr_total), o_jobs = o_dro + up4((size_t)(n_drs + 1) * 4), o_in = o_jobs + up4((size_t)n_jobs * sizeof(crass_b200_uss_job)),
                 o_out = o_in + up4((size_t)n_ss_in * 4), o_n = o_out + up4((size_t)ss_out_cap * 4), o_st = o_n + up4((size_t)n_jobs * 4),
                 total = o_st + up4(n_jobs);
    if (int r = c->d_bases.reserve(n_bases + 64)) return r;
    if (int r = c->d_offsets.reserve((size_t)(n_reads + 1) * sizeof(uint64_t))) return r;
    if (int r = c->d_misc.reserve(total)) return r;
    c->res_valid = false; c->res_found_valid = false; c->packed_valid = false;      // the context's batch buffers are reused
    uint8_t* m = c->d_misc.as<uint8_t>();
    CUDA_TRY(cudaMemcpyAsync(c->d_bases.p, bases, n_bases, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->d_offsets.p, offsets, (size_t)(n_reads + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(m, dr_bytes, dr_total, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(m + o_dro, dr_offsets, (size_t)(n_drs + 1) * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(m + o_jobs, jobs, (size_t)n_jobs * sizeof(crass_b200_uss_job), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(m + o_in, ss_in, (size_t)n_ss_in * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemsetAsync(m + o_out, 0, (size_t)ss_out_cap * 4, c->stream));
    if (int r = crass_b200_update_start_stops_dev(c, c->d_bases.as<uint8_t>(), c->d_offsets.as<uint64_t>(), m, (const uint32_t*)(m + o_dro),
                                                  (const crass_b200_uss_job*)(m + o_jobs), n_jobs, (const uint32_t*)(m + o_in), low_spacer,
                                                  (uint32_t*)(m + o_out), (uint32_t*)(m + o_n), m + o_st, c->stream)) return r;
    CUDA_TRY(cudaMemcpyAsync(ss_out, m + o_out, (size_t)ss_out_cap * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(n_out, m + o_n, (size_t)n_jobs * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(status, m + o_st, n_jobs, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---- KAT entry points -----------------------------------------------------------------------------------
int crass_b200_scan_right(crass_b200_ctx* c, const uint8_t* seq, uint32_t len, uint32_t* ss, uint32_t* n_ss, uint32_t ss_cap,
                          const uint8_t* pattern, uint32_t pattern_len, uint32_t min_spacer, uint32_t scan_range) {
    if (!c || !seq || !ss || !n_ss || !pattern) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    if (*n_ss < 4 || *n_ss > ss_cap) return cbh::fail(CRASS_B200_EINVAL, "need at least two repeats in ss");
    CUDA_TRY(cudaSetDevice(c->device));
    if (int r = c->d_bases.reserve((size_t)len + pattern_len + 64)) return r;
    if (int r = c->d_misc.reserve(((size_t)ss_cap + 4) * sizeof(uint32_t))) return r;
    uint32_t* d_ss = c->d_misc.as<uint32_t>() + 4;
    uint32_t* d_n = c->d_misc.as<uint32_t>();
    CUDA_TRY(cudaMemcpyAsync(c->d_bases.p, seq, len, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->d_bases.as<uint8_t>() + len, pattern, pattern_len, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(d_ss, ss, *n_ss * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(d_n, n_ss, sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    cbk::k_scan_right_one<<<1, 1, 0, c->stream>>>(c->d_bases.as<uint8_t>(), len, d_ss, d_n, ss_cap, pattern_len, min_spacer, scan_range);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(n_ss, d_n, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMemcpyAsync(ss, d_ss, *n_ss * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

int crass_b200_extend_pre_repeat(crass_b200_ctx* c, const uint8_t* seq, uint32_t len, uint32_t* ss, uint32_t n_ss,
                                 uint32_t window, uint32_t min_spacer, uint32_t* repeat_len) {
    if (!c || !seq || !ss || !repeat_len) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    if (n_ss < 4 || (n_ss & 1)) return cbh::fail(CRASS_B200_EINVAL, "need at least two repeats in ss");
    CUDA_TRY(cudaSetDevice(c->device));
    if (int r = c->d_bases.reserve((size_t)len + 64)) return r;
    if (int r = c->d_misc.reserve(((size_t)n_ss + 4) * sizeof(uint32_t))) return r;
    uint32_t* d_ss = c->d_misc.as<uint32_t>() + 4;
    uint32_t* d_r = c->d_misc.as<uint32_t>();
    CUDA_TRY(cudaMemcpyAsync(c->d_bases.p, seq, len, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(d_ss, ss, n_ss * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    cbk::k_extend_one<<<1, 1, 0, c->stream>>>(c->d_bases.as<uint8_t>(), len, d_ss, n_ss, window, min_spacer, d_r);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(repeat_len, d_r, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(ss, d_ss, n_ss * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

int crass_b200_qc_found_repeats(crass_b200_ctx* c, const uint8_t* seq, uint32_t len, const uint32_t* ss, uint32_t n_ss,
                                int min_spacer, int max_spacer, int* result) {
    if (!c || !seq || !ss || !result) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    if (n_ss < 4 || (n_ss & 1)) return cbh::fail(CRASS_B200_EINVAL, "need at least two repeats in ss");
    CUDA_TRY(cudaSetDevice(c->device));
    if (int r = c->d_bases.reserve((size_t)len + 64)) return r;
    if (int r = c->d_misc.reserve(((size_t)n_ss + 4) * sizeof(uint32_t))) return r;
    uint32_t* d_ss = c->d_misc.as<uint32_t>() + 4;
    int* d_r = c->d_misc.as<int>();
    CUDA_TRY(cudaMemcpyAsync(c->d_bases.p, seq, len, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(d_ss, ss, n_ss * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    cbk::k_qc_one<<<1, 1, 0, c->stream>>>(c->d_bases.as<uint8_t>(), len, d_ss, n_ss, min_spacer, max_spacer, d_r);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(result, d_r, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

void crass_b200_free(void* p) { free(p); }

}  // extern "C"
