// engine.cu -- the whole read-scanning path on one or more GPUs of one box, behind the C-ABI (include/crass_b200.h).
//
// What WorkHorse::parseSeqFiles does with libcrispr (WorkHorse.cpp:321-414): searchFile for every file, then
// createNonRedundantSet, then findSingletons for every file -- one caller, one ReadMap.  Here a file's reads are cut
// into contiguous shards, one per device; every device has its own context, stream and host thread and
// runs phase 1 on its shard (K1 + K4 tokens); the shards' distinct DR tokens meet on the first device through ONE
// all-gather of fixed-size token blocks (NCCL when the devices are distinct and libnccl.so.2 can be loaded, peer copies
// otherwise) and are merged there in first-appearance order (K4c), which is the token numbering a sequential run
// produces (StringCheck.cpp:46-55); createNonRedundantSet runs once (K5 + host passes), every device takes the matcher
// and scans its shard (K2).  The hit records of all shards come back to the CALLING thread as one list in global read
// order, so the containers are filled exactly as a single-GPU (or the reference's single-threaded) run fills them:
// the readsFound test stays on the host, keyed by header (libcrispr.cpp:411), duplicates across shards included.
//
// Feed path: a file is parsed once (host/parser.cpp, worker threads) into page-locked memory; the shards are copied
// host-to-device in slices that alternate between two copy streams, so that the next slice's descriptor work overlaps
// the running copy, and stay resident in HBM (bytes, offsets, phase-1 flags, the 2-bit stream K1 leaves behind) until
// the engine is told to release the file -- findSingletons neither parses nor uploads a second time.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <exception>
#include <memory>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <map>
#include <unistd.h>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/crass_b200.h"
#include "host/internal.h"
#include "nccl_dl.h"

namespace {

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

#define ENG_CUDA(expr)                                                                                              \
    do {                                                                                                            \
        cudaError_t _e = (expr);                                                                                    \
        if (_e != cudaSuccess)                                                                                      \
            return cbh::fail(_e == cudaErrorMemoryAllocation ? CRASS_B200_ENOMEM : CRASS_B200_ECUDA,                \
                             std::string(#expr) + " failed: " + cudaGetErrorString(_e));                            \
    } while (0)

struct DBuf {                                             // grow-only device buffer of the lane's device
    void* p = nullptr; size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        const size_t want = bytes + bytes / 8 + 256;
        ENG_CUDA(cudaMalloc(&p, want));
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};
struct HBuf {                                             // grow-only page-locked host buffer
    void* p = nullptr; size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        const size_t want = bytes + bytes / 8 + 256;
        ENG_CUDA(cudaHostAlloc(&p, want, cudaHostAllocDefault));
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};

// bytes per K4 token record / token-block record: length, orientation, token, 4-byte order key
uint32_t token_stride_for(const crass_b200_params& p) { return std::max<uint32_t>(64, (p.high_dr + 6 + 3) & ~3u); }

struct Shard {                                            // one device's contiguous part of one file
    uint32_t r0 = 0, r1 = 0;                              // global read range
    uint64_t b0 = 0, b1 = 0;                              // its bytes in the batch
    DBuf d_bases, d_offsets, d_found;                     // resident in HBM between the phases
    bool resident = false, searched = false;
    uint32_t n() const { return r1 - r0; }
};

struct HitList {                                          // one shard's hit records in read order, indices local to the shard
    std::vector<crass_b200_hit> hits;
    std::vector<uint32_t> pool;
};

struct Lane {                                             // one device: context, streams, scratch
    int device = 0;
    int index = 0;
    crass_b200_ctx* ctx = nullptr;
    cudaStream_t stream = nullptr, copy[2] = {nullptr, nullptr};
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};
    // copies from ORDINARY host memory (an engine's first run, CRASS_B200_PIN): slices are memcpy'd into small page-locked
    // staging buffers by a few threads and sent from there -- what the driver does for a pageable source, on more than one core
    static constexpr int kStageBufs = 8;
    static constexpr size_t kStageBytes = (size_t)4 << 20;
    HBuf stage;                                           // kStageBufs x kStageBytes
    cudaEvent_t ev_stage[kStageBufs] = {};
    bool stage_used[kStageBufs] = {};
    DBuf d_hits, d_sorted, d_pool, d_cnt, d_tokens, d_found2, d_block, d_recv, d_merged;
    HBuf h_cnt, h_offsets, h_hits, h_pool;
    uint64_t h2d_bytes = 0, d2h_bytes = 0;
    uint64_t hits_per_64k = 0, pool_per_64k = 0;          // the most records / pool words per 65 536 reads any call needed so far
    void* comm = nullptr;                                 // ncclComm_t
    int rc = 0;
    std::string err;
    std::vector<std::unique_ptr<Shard> > spare;           // shards of released files: their HBM buffers serve the next file
};

struct FileState {
    std::string path;
    crass_b200_batch* batch = nullptr;
    std::vector<std::unique_ptr<Shard> > shards;          // one per lane
    uint64_t resident_bytes = 0;
    double parse_ms = 0;
};

}  // namespace

struct crass_b200_engine {
    std::vector<std::unique_ptr<Lane> > lanes;
    std::vector<std::unique_ptr<FileState> > files;
    std::map<std::string, std::vector<std::string> > ranged;  // files searched range by range: the names their ranges are resident under
    std::vector<crass_b200_batch*> batch_pool;            // released batches: their buffers (page-locked bases) serve the next parse
    std::mutex pool_mu;                                   // ... taken by the parsing thread of a streamed run while another thread searches
    int pin_policy = 1;                                   // CRASS_B200_PIN: 0 never, 1 auto (from an engine's second run on), 2 always
    std::atomic<int> runs_done{0};
    size_t stream_bytes = (size_t)128 << 20;              // range size of the streamed feed (CRASS_B200_STREAM_MB; 0 = whole files)
    Nccl nccl;
    bool use_nccl = false;
    uint32_t block_cap = 16384;                           // token-block capacity per shard (grows on overflow)
    uint32_t tok_stride = 64;                             // of the most recent search
    size_t resident_budget = 0;                           // bytes of read data a device keeps between the phases
    bool trace = false;
    // last call's stage times (ms) for crass_b200_engine_stats
    double t_parse = 0, t_phase1 = 0, t_exchange = 0, t_phase2 = 0, t_replay = 0;
};

namespace {

// runs fn(lane) on one host thread per device and collects the first failure
template <class F>
int for_each_lane(crass_b200_engine* e, F fn) {
    const size_t n = e->lanes.size();
    for (auto& l : e->lanes) { l->rc = 0; l->err.clear(); }
    auto body = [&](size_t i) {
        Lane& l = *e->lanes[i];
        try {
            if (cudaSetDevice(l.device) != cudaSuccess) { l.rc = CRASS_B200_ECUDA; l.err = "cudaSetDevice failed"; return; }
            l.rc = fn(l);
            if (l.rc) l.err = crass_b200_last_error();
        } catch (std::exception& ex) { l.rc = CRASS_B200_ENOMEM; l.err = ex.what(); }
    };
    if (n == 1) body(0);
    else {
        std::vector<std::thread> th;
        for (size_t i = 1; i < n; ++i) th.emplace_back(body, i);
        body(0);
        for (auto& t : th) t.join();
    }
    for (auto& l : e->lanes) if (l->rc) return cbh::fail(l->rc, "device " + std::to_string(l->device) + ": " + l->err);
    return 0;
}

void retire_batch(crass_b200_engine* e, crass_b200_batch* b) {
    if (!b) return;
    std::lock_guard<std::mutex> g(e->pool_mu);
    if (e->batch_pool.size() < 64) e->batch_pool.push_back(b);                 // (a streamed file is many small batches)
    else crass_b200_batch_destroy(b);
}

crass_b200_batch* take_batch(crass_b200_engine* e, size_t want_bytes) {       // the pooled batch whose base buffer fits best
    std::lock_guard<std::mutex> g(e->pool_mu);
    int best = -1;
    for (size_t i = 0; i < e->batch_pool.size(); ++i) {
        const size_t cap = e->batch_pool[i]->b.bases_cap;
        if (best < 0) { best = (int)i; continue; }
        const size_t bc = e->batch_pool[(size_t)best]->b.bases_cap;
        if ((cap >= want_bytes && (bc < want_bytes || cap < bc)) || (cap < want_bytes && bc < want_bytes && cap > bc)) best = (int)i;
    }
    // Page-locking the base buffers is worth it for an engine that is used again (its buffers are pooled): locking 1.6 GB costs
    // well over a second, a run on it 0.1 s.  The FIRST run of an engine therefore parses into ordinary memory (copied to the
    // device through the driver's staging buffers); from the second run on, buffers are page-locked as they are taken from the pool.
    const bool pin = e->pin_policy == 2 || (e->pin_policy == 1 && e->runs_done.load() >= 1);
    crass_b200_batch* h;
    if (best < 0) h = new crass_b200_batch();
    else { h = e->batch_pool[(size_t)best]; e->batch_pool.erase(e->batch_pool.begin() + best); }
    h->b.want_pinned = pin;
    if (pin) h->b.pin_now();
    return h;
}

// kseq-compatible parse of `path` into a batch from the engine's pool (a fresh one the first time)
int parse_pooled(crass_b200_engine* e, const char* path, crass_b200_batch** out) {
    crass_b200_batch* h = take_batch(e, (size_t)-1);                          // whole files: the largest buffer there is
    cbh::Batch* got = nullptr;
    const int rc = cbh::parse_file(path, &got, &h->b);
    if (rc) { retire_batch(e, h); return rc; }
    *out = h;
    return 0;
}

std::unique_ptr<Shard> take_shard(Lane& l) {
    if (l.spare.empty()) return std::unique_ptr<Shard>(new Shard());
    std::unique_ptr<Shard> s = std::move(l.spare.back());
    l.spare.pop_back();
    s->resident = s->searched = false;
    return s;
}

void make_shards(crass_b200_engine* e, FileState& fs) {               // contiguous shards of (almost) equal read counts
    const cbh::Batch& b = fs.batch->b;
    const uint32_t n = b.n(), G = (uint32_t)e->lanes.size();
    for (uint32_t g = 0; g < G; ++g) {
        std::unique_ptr<Shard> s = take_shard(*e->lanes[g]);
        s->r0 = (uint32_t)((uint64_t)n * g / G); s->r1 = (uint32_t)((uint64_t)n * (g + 1) / G);
        s->b0 = b.offsets[s->r0]; s->b1 = b.offsets[s->r1];
        fs.shards.push_back(std::move(s));
    }
}

int search_parsed(crass_b200_engine* e, std::unique_ptr<FileState> fs, const crass_b200_params* params,
                  const crass_b200_batch** batch_out, crass_b200_hit** hits_out, uint32_t* n_hits, uint32_t** pool_out, uint32_t* n_pool);

FileState* find_file(crass_b200_engine* e, const char* path) {
    for (auto& f : e->files) if (f->path == path) return f.get();
    return nullptr;
}

// H2D of one shard: slices alternate between the lane's two copy streams; the compute stream waits for both
int upload_shard(Lane& l, const cbh::Batch& b, Shard& s) {
    const uint64_t nbytes = s.b1 - s.b0;
    if (int r = s.d_bases.reserve(nbytes + 64)) return r;
    if (int r = s.d_offsets.reserve(((size_t)s.n() + 1) * sizeof(uint64_t))) return r;
    if (int r = s.d_found.reserve((size_t)s.n() + 16)) return r;
    if (int r = l.h_offsets.reserve(((size_t)s.n() + 1) * sizeof(uint64_t))) return r;
    uint64_t* ho = l.h_offsets.as<uint64_t>();
    const uint64_t* src = b.offsets.data() + s.r0;
    for (uint32_t i = 0; i <= s.n(); ++i) ho[i] = src[i] - s.b0;                 // shard-relative offsets
    ENG_CUDA(cudaMemcpyAsync(s.d_offsets.p, ho, ((size_t)s.n() + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, l.copy[0]));
    const uint64_t slice = (uint64_t)64 << 20;
    int which = 0;
    struct Slice { uint8_t* dst; const uint8_t* src; uint64_t len; };
    std::vector<Slice> work;
    const bool staged = !b.pinned && !b.registered && nbytes >= ((uint64_t)16 << 20) && !getenv("CRASS_B200_NO_STAGING");
    if (staged) {
        if (int r = l.stage.reserve((size_t)Lane::kStageBufs * Lane::kStageBytes)) return r;
        for (int k = 0; k < Lane::kStageBufs; ++k) if (!l.ev_stage[k]) ENG_CUDA(cudaEventCreateWithFlags(&l.ev_stage[k], cudaEventDisableTiming));
    }
    // back to back on the device; on the host a streamed range lies in segments (cbh::Batch::segs), each copied to its place
    const size_t n_seg = b.segs.empty() ? 1 : b.segs.size();
    for (size_t k = 0; k < n_seg; ++k) {
        const uint64_t dev0 = b.segs.empty() ? 0 : b.segs[k].dev0, host0 = b.segs.empty() ? 0 : b.segs[k].host0;
        const uint64_t dev1 = k + 1 < n_seg ? b.segs[k + 1].dev0 : b.offsets.back();
        const uint64_t lo = std::max<uint64_t>(dev0, s.b0), hi = std::min<uint64_t>(dev1, s.b1);
        if (staged) {
            for (uint64_t at = lo; at < hi; at += Lane::kStageBytes)
                work.push_back(Slice{s.d_bases.as<uint8_t>() + (at - s.b0), b.bases + host0 + (at - dev0), std::min<uint64_t>(Lane::kStageBytes, hi - at)});
            continue;
        }
        for (uint64_t at = lo; at < hi; at += slice, which ^= 1) {
            const uint64_t len = std::min(slice, hi - at);
            ENG_CUDA(cudaMemcpyAsync(s.d_bases.as<uint8_t>() + (at - s.b0), b.bases + host0 + (at - dev0), len, cudaMemcpyHostToDevice, l.copy[which]));
        }
    }
    if (staged && !work.empty()) {
        constexpr int kThreads = 4, kPer = Lane::kStageBufs / kThreads;
        std::atomic<int> bad{0};
        auto run = [&](int t) {
            if (cudaSetDevice(l.device) != cudaSuccess) { bad.store(1); return; }
            for (size_t j = (size_t)t, k = 0; j < work.size() && !bad.load(); j += kThreads, ++k) {
                const int buf = t * kPer + (int)(k % kPer);
                uint8_t* st = l.stage.as<uint8_t>() + (size_t)buf * Lane::kStageBytes;
                if (l.stage_used[buf] && cudaEventSynchronize(l.ev_stage[buf]) != cudaSuccess) { bad.store(1); return; }   // its last copy has left
                memcpy(st, work[j].src, (size_t)work[j].len);
                if (cudaMemcpyAsync(work[j].dst, st, (size_t)work[j].len, cudaMemcpyHostToDevice, l.copy[t & 1]) != cudaSuccess ||
                    cudaEventRecord(l.ev_stage[buf], l.copy[t & 1]) != cudaSuccess) { bad.store(1); return; }
                l.stage_used[buf] = true;
            }
        };
        std::thread helpers[kThreads - 1];
        bool started[kThreads - 1] = {};
        for (int t = 1; t < kThreads; ++t) {
            try { helpers[t - 1] = std::thread(run, t); started[t - 1] = true; } catch (...) {}   // (no thread to be had: its share is done below)
        }
        run(0);
        for (int t = 1; t < kThreads; ++t) { if (started[t - 1]) helpers[t - 1].join(); else run(t); }
        if (bad.load()) { (void)cudaGetLastError(); return cbh::fail(CRASS_B200_ECUDA, "staged copy to the device failed"); }
    }
    for (int k = 0; k < 2; ++k) {
        ENG_CUDA(cudaEventRecord(l.ev_copy[k], l.copy[k]));
        ENG_CUDA(cudaStreamWaitEvent(l.stream, l.ev_copy[k], 0));
    }
    l.h2d_bytes += nbytes + ((uint64_t)s.n() + 1) * 8;
    s.resident = true;
    return 0;
}

// runs `launch(hits_cap, pool_cap)` on the lane (a K1 or K2 call that fills d_hits / d_pool / d_cnt and whose flags are in
// d_flags), grows the buffers once if the counters say so, and brings the records back in read order
template <class Launch>
int collect_hits(Lane& l, uint32_t n_reads, const uint8_t* d_flags, Launch launch, HitList& out, uint32_t kTokStride) {
    const bool want_tokens = kTokStride != 0;
    // first guess: a quarter of the reads hit; afterwards what this lane has seen, plus a quarter (a sample where phase 2
    // recruits most reads overflows once, on its first range, and never again)
    uint32_t hits_cap = std::max<uint32_t>(4096, n_reads / 4 + 1024), pool_cap = hits_cap * 6;
    hits_cap = (uint32_t)std::min<uint64_t>((uint64_t)n_reads + 16, std::max<uint64_t>(hits_cap, (l.hits_per_64k * n_reads >> 16) * 5 / 4 + 1024));
    pool_cap = (uint32_t)std::min<uint64_t>(0xFFFFFFF0ull, std::max<uint64_t>(pool_cap, (l.pool_per_64k * n_reads >> 16) * 5 / 4 + 4096));
    if (int r = l.d_cnt.reserve(8 * sizeof(uint32_t))) return r;
    if (int r = l.h_cnt.reserve(8 * sizeof(uint32_t))) return r;
    uint32_t* hc = l.h_cnt.as<uint32_t>();
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (int r = l.d_hits.reserve((size_t)hits_cap * sizeof(crass_b200_hit))) return r;
        if (int r = l.d_pool.reserve((size_t)pool_cap * sizeof(uint32_t))) return r;
        if (want_tokens) {
            if (int r = l.d_tokens.reserve((size_t)hits_cap * kTokStride)) return r;
            if (int r = crass_b200_ctx_set_token_output(l.ctx, l.d_tokens.p, kTokStride)) return r;
        }
        const int lr = launch(hits_cap, pool_cap);
        if (want_tokens) crass_b200_ctx_set_token_output(l.ctx, nullptr, 0);
        if (lr) return lr;
        ENG_CUDA(cudaMemcpyAsync(hc, l.d_cnt.p, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, l.stream));
        ENG_CUDA(cudaStreamSynchronize(l.stream));
        if (!hc[2]) break;
        if (attempt == 1) return cbh::fail(CRASS_B200_EOVERFLOW, "hit buffers overflowed after being sized from the counters");
        hits_cap = hc[0] + 16; pool_cap = hc[1] + 16;                             // the counters include what did not fit
    }
    const uint32_t nh = hc[0], np = hc[1];
    if (n_reads) {
        l.hits_per_64k = std::max<uint64_t>(l.hits_per_64k, (((uint64_t)nh << 16) + n_reads - 1) / n_reads);
        l.pool_per_64k = std::max<uint64_t>(l.pool_per_64k, (((uint64_t)np << 16) + n_reads - 1) / n_reads);
    }
    out.hits.resize(nh); out.pool.resize(np);
    if (nh) {
        if (int r = l.d_sorted.reserve((size_t)nh * sizeof(crass_b200_hit))) return r;
        if (int r = crass_b200_sort_hits_dev(l.ctx, d_flags, n_reads, l.d_hits.as<crass_b200_hit>(), l.d_cnt.as<uint32_t>(), nh,
                                             l.d_sorted.as<crass_b200_hit>(), l.stream)) return r;
        if (int r = l.h_hits.reserve((size_t)nh * sizeof(crass_b200_hit))) return r;
        if (int r = l.h_pool.reserve((size_t)np * sizeof(uint32_t) + 16)) return r;
        ENG_CUDA(cudaMemcpyAsync(l.h_hits.p, l.d_sorted.p, (size_t)nh * sizeof(crass_b200_hit), cudaMemcpyDeviceToHost, l.stream));
        ENG_CUDA(cudaMemcpyAsync(l.h_pool.p, l.d_pool.p, (size_t)np * sizeof(uint32_t), cudaMemcpyDeviceToHost, l.stream));
        ENG_CUDA(cudaStreamSynchronize(l.stream));
        memcpy(out.hits.data(), l.h_hits.p, (size_t)nh * sizeof(crass_b200_hit));
        memcpy(out.pool.data(), l.h_pool.p, (size_t)np * sizeof(uint32_t));
        l.d2h_bytes += (uint64_t)nh * sizeof(crass_b200_hit) + (uint64_t)np * 4 + 16;
    }
    return 0;
}

// the shards' lists -> one list in global read order (shards are contiguous and in order, so this is a concatenation)
void concat_hits(const std::vector<std::unique_ptr<Shard> >& shards, const std::vector<HitList>& parts,
                 std::vector<crass_b200_hit>& hits, std::vector<uint32_t>& pool) {
    size_t nh = 0, np = 0;
    for (const HitList& p : parts) { nh += p.hits.size(); np += p.pool.size(); }
    hits.clear(); pool.clear();
    hits.reserve(nh); pool.reserve(np);
    for (size_t g = 0; g < parts.size(); ++g) {
        const uint32_t pool0 = (uint32_t)pool.size(), r0 = shards[g]->r0;
        for (crass_b200_hit h : parts[g].hits) { h.read_index += r0; h.ss_offset += pool0; hits.push_back(h); }
        pool.insert(pool.end(), parts[g].pool.begin(), parts[g].pool.end());
    }
}

int max_len_of(const cbh::Batch& b) { return (int)b.max_len; }

void drop_residency_if_needed(crass_b200_engine* e, FileState* keep) {
    // read data a device keeps between the phases is bounded; the oldest files give way first (they are uploaded again
    // from the host batch when their turn in phase 2 comes)
    uint64_t total = 0;
    for (auto& f : e->files) total += f->resident_bytes;
    for (auto& f : e->files) {
        if (total <= e->resident_budget) break;
        if (f.get() == keep || !f->resident_bytes) continue;
        for (size_t g = 0; g < f->shards.size(); ++g) {
            cudaSetDevice(e->lanes[g]->device);
            f->shards[g]->d_bases.release(); f->shards[g]->d_offsets.release();
            f->shards[g]->resident = false;
        }
        total -= f->resident_bytes;
        f->resident_bytes = 0;
    }
}

}  // namespace

extern "C" {

int crass_b200_engine_create(const int* devices, uint32_t n_devices, crass_b200_engine** out) {
    if (!out || (n_devices && !devices)) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    const int have = crass_b200_device_count();
    if (have <= 0) return cbh::fail(CRASS_B200_ENODEVICE, "no CUDA device: crass_b200 has no CPU execution path");
    std::vector<int> devs(devices, devices + n_devices);
    if (devs.empty()) devs.push_back(0);
    if (devs.size() > 64) return cbh::fail(CRASS_B200_EINVAL, "more than 64 devices");
    std::unique_ptr<crass_b200_engine> e(new crass_b200_engine());
    e->trace = getenv("CRASS_B200_TRACE") != nullptr;
    size_t min_mem = (size_t)-1;
    bool distinct = true;
    for (size_t i = 0; i < devs.size(); ++i) {
        if (devs[i] < 0 || devs[i] >= have) return cbh::fail(CRASS_B200_EINVAL, "bad device ordinal");
        for (size_t j = 0; j < i; ++j) if (devs[j] == devs[i]) distinct = false;   // the same GPU twice: shards share it (tests)
        std::unique_ptr<Lane> l(new Lane());
        l->device = devs[i]; l->index = (int)i;
        ENG_CUDA(cudaSetDevice(devs[i]));
        if (int r = crass_b200_ctx_create(devs[i], &l->ctx)) return r;
        crass_b200_ctx_keep_packed(l->ctx, 1);
        ENG_CUDA(cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) {
            ENG_CUDA(cudaStreamCreateWithFlags(&l->copy[k], cudaStreamNonBlocking));
            ENG_CUDA(cudaEventCreateWithFlags(&l->ev_copy[k], cudaEventDisableTiming));
        }
        size_t fr = 0, tot = 0;
        ENG_CUDA(cudaMemGetInfo(&fr, &tot));
        min_mem = std::min(min_mem, tot);
        e->lanes.push_back(std::move(l));
    }
    e->resident_budget = min_mem / 2;
    if (const char* v = getenv("CRASS_B200_RESIDENT_MB")) e->resident_budget = (size_t)strtoull(v, nullptr, 10) << 20;
    if (const char* v = getenv("CRASS_B200_PIN")) e->pin_policy = !strcmp(v, "never") ? 0 : !strcmp(v, "always") ? 2 : 1;
    if (const char* v = getenv("CRASS_B200_STREAM_MB")) e->stream_bytes = (size_t)strtoull(v, nullptr, 10) << 20;
    if (const char* v = getenv("CRASS_B200_STREAM_BYTES")) e->stream_bytes = (size_t)strtoull(v, nullptr, 10);   // (tests: ranges of a few KB)
    // peer access for the block gather without NCCL
    for (size_t i = 0; i < devs.size(); ++i)
        for (size_t j = 0; j < devs.size(); ++j) {
            if (devs[i] == devs[j]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devs[i], devs[j]);
            if (can) { cudaSetDevice(devs[i]); if (cudaDeviceEnablePeerAccess(devs[j], 0) != cudaSuccess) (void)cudaGetLastError(); }
        }
    const char* xsel = getenv("CRASS_B200_EXCHANGE");                            // "peer": never NCCL
    if (devs.size() > 1 && distinct && !(xsel && !strcmp(xsel, "peer")) && e->nccl.load()) {
        std::vector<void*> comms(devs.size(), nullptr);
        const int rc = e->nccl.CommInitAll(comms.data(), (int)devs.size(), devs.data());
        if (rc == 0) {
            for (size_t i = 0; i < devs.size(); ++i) e->lanes[i]->comm = comms[i];
            e->use_nccl = true;
        } else if (e->trace) fprintf(stderr, "[crass_b200] ncclCommInitAll failed (%d): token blocks travel by peer copies\n", rc);
    }
    if (e->trace) fprintf(stderr, "[crass_b200] engine: %zu device(s), exchange by %s\n", devs.size(), e->use_nccl ? "NCCL all-gather" : "peer copies");
    *out = e.release();
    return 0;
}

void crass_b200_engine_destroy(crass_b200_engine* e) {
    if (!e) return;
    for (auto& f : e->files) {
        for (size_t g = 0; g < f->shards.size(); ++g) {
            cudaSetDevice(e->lanes[g]->device);
            f->shards[g]->d_bases.release(); f->shards[g]->d_offsets.release(); f->shards[g]->d_found.release();
        }
        crass_b200_batch_destroy(f->batch);
    }
    for (crass_b200_batch* b : e->batch_pool) crass_b200_batch_destroy(b);
    for (auto& l : e->lanes) {
        cudaSetDevice(l->device);
        for (auto& s : l->spare) { s->d_bases.release(); s->d_offsets.release(); s->d_found.release(); }
        if (l->stream) cudaStreamSynchronize(l->stream);
        if (l->comm && e->nccl.CommDestroy) e->nccl.CommDestroy(l->comm);
        for (DBuf* b : {&l->d_hits, &l->d_sorted, &l->d_pool, &l->d_cnt, &l->d_tokens, &l->d_found2, &l->d_block, &l->d_recv, &l->d_merged}) b->release();
        for (HBuf* b : {&l->h_cnt, &l->h_offsets, &l->h_hits, &l->h_pool}) b->release();
        for (int k = 0; k < 2; ++k) { if (l->ev_copy[k]) cudaEventDestroy(l->ev_copy[k]); if (l->copy[k]) cudaStreamDestroy(l->copy[k]); }
        for (int k = 0; k < Lane::kStageBufs; ++k) if (l->ev_stage[k]) cudaEventDestroy(l->ev_stage[k]);
        l->stage.release();
        if (l->stream) cudaStreamDestroy(l->stream);
        crass_b200_ctx_destroy(l->ctx);
    }
    delete e;
}

uint32_t crass_b200_engine_num_devices(const crass_b200_engine* e) { return e ? (uint32_t)e->lanes.size() : 0; }
int crass_b200_engine_uses_nccl(const crass_b200_engine* e) { return e && e->use_nccl ? 1 : 0; }

void crass_b200_engine_transfer_bytes(const crass_b200_engine* e, uint64_t* h2d, uint64_t* d2h) {
    uint64_t a = 0, b = 0;
    if (e) for (auto& l : e->lanes) { a += l->h2d_bytes; b += l->d2h_bytes; }
    if (h2d) *h2d = a;
    if (d2h) *d2h = b;
}

uint64_t crass_b200_engine_launch_count(const crass_b200_engine* e) {
    uint64_t n = 0;
    if (e) for (auto& l : e->lanes) n += crass_b200_ctx_launch_count(l->ctx);
    return n;
}

void crass_b200_engine_stage_ms(const crass_b200_engine* e, double* parse, double* phase1, double* exchange, double* phase2) {
    if (parse) *parse = e ? e->t_parse : 0;
    if (phase1) *phase1 = e ? e->t_phase1 : 0;
    if (exchange) *exchange = e ? e->t_exchange : 0;
    if (phase2) *phase2 = e ? e->t_phase2 : 0;
}

void crass_b200_engine_release_file(crass_b200_engine* e, const char* path) {
    if (!e || !path) return;
    auto rg = e->ranged.find(path);
    if (rg != e->ranged.end()) {                                               // a file searched range by range: all its ranges
        const std::vector<std::string> names = rg->second;
        e->ranged.erase(rg);
        for (const std::string& nm : names) crass_b200_engine_release_file(e, nm.c_str());
    }
    for (size_t i = 0; i < e->files.size(); ++i) {
        if (e->files[i]->path != path) continue;
        FileState* f = e->files[i].get();
        for (size_t g = 0; g < f->shards.size(); ++g) {
            Lane& l = *e->lanes[g];
            if (l.spare.size() < 64) { l.spare.push_back(std::move(f->shards[g])); continue; }
            cudaSetDevice(l.device);
            f->shards[g]->d_bases.release(); f->shards[g]->d_offsets.release(); f->shards[g]->d_found.release();
        }
        retire_batch(e, f->batch);
        e->files.erase(e->files.begin() + (long)i);
        return;
    }
}

// searchFile (libcrispr.cpp:68-166) for one file on all devices of the engine
int crass_b200_engine_search_file(crass_b200_engine* e, const char* path, const crass_b200_params* params,
                                  const crass_b200_batch** batch_out, crass_b200_hit** hits_out, uint32_t* n_hits,
                                  uint32_t** pool_out, uint32_t* n_pool) {
    if (!e || !path || !params || !hits_out || !n_hits || !pool_out || !n_pool) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    crass_b200_engine_release_file(e, path);                                  // searching a file again starts over
    std::unique_ptr<FileState> fs(new FileState());
    fs->path = path;
    double t0 = now_ms();
    if (int r = parse_pooled(e, path, &fs->batch)) return r;
    fs->parse_ms = now_ms() - t0;
    e->t_parse += fs->parse_ms;
    return search_parsed(e, std::move(fs), params, batch_out, hits_out, n_hits, pool_out, n_pool);
}

}  // extern "C"
namespace {
// searchFile from the parsed batch on: shards, copy in, K1 on every device, one hit list in read order
int search_parsed(crass_b200_engine* e, std::unique_ptr<FileState> fs, const crass_b200_params* params,
                  const crass_b200_batch** batch_out, crass_b200_hit** hits_out, uint32_t* n_hits, uint32_t** pool_out, uint32_t* n_pool) {
    const std::string path_s = fs->path;
    const char* path = path_s.c_str();
    double t0 = now_ms();
    const cbh::Batch& b = fs->batch->b;
    const uint32_t G = (uint32_t)e->lanes.size();
    make_shards(e, *fs);
    FileState* f = fs.get();
    e->files.push_back(std::move(fs));
    drop_residency_if_needed(e, f);
    std::vector<HitList> parts(G);
    t0 = now_ms();
    const int max_len = max_len_of(b);
    e->tok_stride = token_stride_for(*params);
    const int rc = for_each_lane(e, [&](Lane& l) -> int {
        Shard& s = *f->shards[l.index];
        if (s.n() == 0) { s.searched = true; l.h_cnt.reserve(32); if (l.h_cnt.p) l.h_cnt.as<uint32_t>()[0] = 0; return 0; }
        if (int r = upload_shard(l, b, s)) return r;
        crass_b200_ctx_keep_packed_bases(l.ctx, std::max<uint64_t>(2, s.b1 - s.b0));      // the 2-bit stream of exactly this shard
        auto launch = [&](uint32_t hits_cap, uint32_t pool_cap) {
            return crass_b200_dr_search_dev(l.ctx, s.d_bases.as<uint8_t>(), s.d_offsets.as<uint64_t>(), s.n(), (uint32_t)max_len, params,
                                            s.d_found.as<uint8_t>(), l.d_hits.as<crass_b200_hit>(), hits_cap, l.d_pool.as<uint32_t>(), pool_cap,
                                            l.d_cnt.as<uint32_t>(), l.stream);
        };
        if (int r = collect_hits(l, s.n(), s.d_found.as<uint8_t>(), launch, parts[l.index], e->tok_stride)) return r;
        s.searched = true;
        return 0;
    });
    if (rc) { crass_b200_engine_release_file(e, path); return rc; }
    e->t_phase1 += now_ms() - t0;
    f->resident_bytes = 0;
    for (auto& s : f->shards) f->resident_bytes = std::max<uint64_t>(f->resident_bytes, s->d_bases.cap + s->d_offsets.cap);
    std::vector<crass_b200_hit> hits; std::vector<uint32_t> pool;
    concat_hits(f->shards, parts, hits, pool);
    crass_b200_hit* h = (crass_b200_hit*)malloc(sizeof(crass_b200_hit) * (hits.size() ? hits.size() : 1));
    uint32_t* p = (uint32_t*)malloc(sizeof(uint32_t) * (pool.size() ? pool.size() : 1));
    if (!h || !p) { free(h); free(p); return cbh::fail(CRASS_B200_ENOMEM, "malloc"); }
    if (!hits.empty()) memcpy(h, hits.data(), hits.size() * sizeof(crass_b200_hit));
    if (!pool.empty()) memcpy(p, pool.data(), pool.size() * sizeof(uint32_t));
    *hits_out = h; *n_hits = (uint32_t)hits.size(); *pool_out = p; *n_pool = (uint32_t)pool.size();
    if (batch_out) *batch_out = f->batch;
    return 0;
}

// a queue between two stages of the streamed run
template <class T>
struct StageQueue {
    std::mutex mu; std::condition_variable cv; std::deque<T> q; bool closed = false;
    void push(T v) { { std::lock_guard<std::mutex> g(mu); q.push_back(std::move(v)); } cv.notify_one(); }
    void close() { { std::lock_guard<std::mutex> g(mu); closed = true; } cv.notify_all(); }
    bool pop(T& v) {
        std::unique_lock<std::mutex> g(mu);
        cv.wait(g, [&] { return !q.empty() || closed; });
        if (q.empty()) return false;
        v = std::move(q.front()); q.pop_front();
        return true;
    }
};

struct SearchedRange {
    const crass_b200_batch* batch = nullptr;
    crass_b200_hit* hits = nullptr; uint32_t nh = 0;
    uint32_t* pool = nullptr; uint32_t np = 0;
};

// phase 1 of a streamed run: this thread parses range after range (worker threads inside the parser), a second thread copies
// each parsed range to the devices and runs K1 on it, a third hands the hits of each searched range, in file order, to
// `consume` (the replay into the containers).  The ranges stay resident under the names left in range_paths.
template <class Consume>
int stream_phase1(crass_b200_engine* e, cbh::ParseStream* first_ps, const char* const* paths, uint32_t n_paths, const crass_b200_params* params,
                  Consume consume, int* max_len, std::vector<std::string>& range_paths) {
    StageQueue<std::unique_ptr<FileState> > parsed;
    StageQueue<SearchedRange> searched;
    std::atomic<int> fail_rc{0};
    std::string fail_msg;
    std::mutex fail_mu;
    auto note_failure = [&](int rc) {
        std::lock_guard<std::mutex> g(fail_mu);
        if (!fail_rc.load()) { fail_msg = crass_b200_last_error(); fail_rc.store(rc); }
    };
    const double t_run = now_ms();
    auto mark = [&](const char* what, uint32_t i, double t0) {                 // CRASS_B200_TRACE: the timeline of the stages
        if (e->trace) fprintf(stderr, "[crass_b200]   %-7s range %2u: %8.1f -> %8.1f ms\n", what, i, t0 - t_run, now_ms() - t_run);
    };
    std::thread searcher([&]() {
        std::unique_ptr<FileState> fs;
        uint32_t i = 0;
        while (parsed.pop(fs)) {
            if (fail_rc.load()) { retire_batch(e, fs->batch); continue; }
            try {                                                               // (nothing may leave a thread as an exception)
                SearchedRange r;
                const double t0 = now_ms();
                const int rc = search_parsed(e, std::move(fs), params, &r.batch, &r.hits, &r.nh, &r.pool, &r.np);
                mark("search", i++, t0);
                if (rc) { note_failure(rc); continue; }
                searched.push(r);
            } catch (std::exception& ex) { cbh::fail(CRASS_B200_ENOMEM, std::string("streamed search: ") + ex.what()); note_failure(CRASS_B200_ENOMEM); }
        }
        searched.close();
    });
    std::thread replayer([&]() {
        SearchedRange r;
        uint64_t first_read = 0;
        uint32_t i = 0;
        while (searched.pop(r)) {
            if (!fail_rc.load()) {
                try {
                    const double t0 = now_ms();
                    const int rc = consume(r, first_read);
                    e->t_replay += now_ms() - t0;
                    mark("replay", i++, t0);
                    if (rc) note_failure(rc);
                } catch (std::exception& ex) { cbh::fail(CRASS_B200_ENOMEM, std::string("streamed replay: ") + ex.what()); note_failure(CRASS_B200_ENOMEM); }
            }
            first_read += crass_b200_batch_num_reads(r.batch);
            free(r.hits); free(r.pool);
        }
    });
    // the files one after the other (each a kseq stream of its own: nothing is carried from file to file), their ranges
    // numbered through; first_ps, if given, is the already opened stream of paths[0]
    uint32_t i = 0;
    for (uint32_t f = 0; f < n_paths && !fail_rc.load(); ++f) {
        cbh::ParseStream* ps = f == 0 && first_ps ? first_ps : cbh::parse_stream_open(paths[f], e->stream_bytes);
        if (!ps) { note_failure(CRASS_B200_EIO); break; }
        for (; !fail_rc.load(); ++i) {
            crass_b200_batch* h = take_batch(e, e->stream_bytes);
            const double t0 = now_ms();
            const int got = cbh::parse_stream_next(ps, &h->b);
            e->t_parse += now_ms() - t0;
            mark("parse", i, t0);
            if (got <= 0) { retire_batch(e, h); if (got < 0) note_failure(got); break; }
            *max_len = std::max(*max_len, (int)h->b.max_len);
            std::unique_ptr<FileState> fs(new FileState());
            fs->path = std::string(paths[f]) + "#" + std::to_string(i);
            fs->batch = h;
            range_paths.push_back(fs->path);
            parsed.push(std::move(fs));
        }
        if (!(f == 0 && first_ps)) cbh::parse_stream_close(ps);
    }
    parsed.close();
    searcher.join();
    replayer.join();
    const int rc = fail_rc.load();
    if (rc) cbh::fail(rc, fail_msg);
    return rc;
}

int run_streamed(crass_b200_engine* e, cbh::ParseStream* first_ps, const char* const* paths, uint32_t n_paths, const crass_b200_params* params,
                 int phases, crass_b200_results* res, int* max_len) {
    std::vector<std::string> range_paths;
    int rc = stream_phase1(e, first_ps, paths, n_paths, params, [&](const SearchedRange& r, uint64_t) {
        return crass_b200_results_add_phase1(res, r.batch, r.hits, r.nh, r.pool);
    }, max_len, range_paths);
    const double t_p1 = now_ms();
    if (!rc && phases >= 2) {
        // the containers' token list (filled in read order above) is the sequential numbering
        crass_b200_ac* ac = nullptr;
        uint32_t n_pat = 0;
        const double t0 = now_ms();
        char* pats = crass_b200_results_non_redundant(res, params->kmer_clust, &n_pat);
        free(pats);
        if (n_pat) {
            std::vector<uint8_t> bytes; std::vector<uint32_t> offs(1, 0);
            for (const std::string& p : res->r.non_redundant) { bytes.insert(bytes.end(), p.begin(), p.end()); offs.push_back((uint32_t)bytes.size()); }
            rc = crass_b200_ac_build(bytes.data(), offs.data(), n_pat, &ac);
        }
        e->t_exchange += now_ms() - t0;
        if (e->trace) fprintf(stderr, "[crass_b200]   clustering + matcher %.1f ms (%u patterns)\n", now_ms() - t0, n_pat);
        // K2 over every resident range first (a fraction of a millisecond each), then ONE replay of all their matches
        std::vector<const crass_b200_batch*> bs; std::vector<crass_b200_hit*> hs; std::vector<uint32_t> nhs; std::vector<uint32_t*> pools;
        const double tf = now_ms();
        for (size_t f = 0; f < range_paths.size() && !rc && ac; ++f) {
            const crass_b200_batch* b = nullptr;
            crass_b200_hit* hits = nullptr; uint32_t nh = 0, np = 0; uint32_t* pool = nullptr;
            rc = crass_b200_engine_find_singletons(e, range_paths[f].c_str(), ac, 1, &b, &hits, &nh, &pool, &np);
            if (rc) { free(hits); free(pool); break; }
            bs.push_back(b); hs.push_back(hits); nhs.push_back(nh); pools.push_back(pool);
        }
        const double t1 = now_ms();
        if (!rc && !bs.empty()) {
            rc = crass_b200_results_add_phase2_ranges(res, (uint32_t)bs.size(), bs.data(), hs.data(), nhs.data(), pools.data());
            e->t_replay += now_ms() - t1;
        }
        if (e->trace) fprintf(stderr, "[crass_b200]   phase 2: scan of %zu ranges %.1f ms, replay %.1f ms\n", bs.size(), t1 - tf, now_ms() - t1);
        for (size_t f = 0; f < hs.size(); ++f) { free(hs[f]); free(pools[f]); }
        crass_b200_ac_destroy(ac);
    } else if (!rc) {
        res->r.lazy_kmer_clust = (int)params->kmer_clust;
    }
    const double t_p2 = now_ms();
    for (const std::string& p : range_paths) crass_b200_engine_release_file(e, p.c_str());
    if (e->trace) fprintf(stderr, "[crass_b200]   after phase 1: clustering + phase 2 + replay %.1f ms, release %.1f ms\n", t_p2 - t_p1, now_ms() - t_p2);
    return rc;
}
}  // namespace
extern "C" {

// findSingletons (libcrispr.cpp:444-518) for one file: every device scans its shard with the given matcher.  skip_found != 0
// leaves out the reads phase 1 flagged on the device (their headers are in readsFound anyway); 0 scans every read, which is
// what a caller needs whose readsFound table may hold other headers than this engine's phase 1 put there.
int crass_b200_engine_find_singletons(crass_b200_engine* e, const char* path, const crass_b200_ac* ac, int skip_found,
                                      const crass_b200_batch** batch_out, crass_b200_hit** hits_out, uint32_t* n_hits,
                                      uint32_t** pool_out, uint32_t* n_pool) {
    if (!e || !path || !ac || !hits_out || !n_hits || !pool_out || !n_pool) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    FileState* f = find_file(e, path);
    if (!f) {                                                                 // not searched through this engine: parse and shard now
        std::unique_ptr<FileState> fs(new FileState());
        fs->path = path;
        const double t0 = now_ms();
        if (int r = parse_pooled(e, path, &fs->batch)) return r;
        e->t_parse += now_ms() - t0;
        make_shards(e, *fs);
        f = fs.get();
        e->files.push_back(std::move(fs));
    }
    const cbh::Batch& b = f->batch->b;
    const uint32_t G = (uint32_t)e->lanes.size();
    if (ac->a.min_pattern_len < 23) cbh::ensure_dfa(&const_cast<crass_b200_ac*>(ac)->a);   // built once, before the lanes share it
    std::vector<HitList> parts(G);
    const double t0 = now_ms();
    const int max_len = max_len_of(b);
    const int rc = for_each_lane(e, [&](Lane& l) -> int {
        Shard& s = *f->shards[l.index];
        if (s.n() == 0) return 0;
        if (!s.resident) { if (int r = upload_shard(l, b, s)) return r; }
        const bool skip = skip_found && s.searched;
        if (int r = l.d_found2.reserve((size_t)s.n() + 16)) return r;
        auto launch = [&](uint32_t hits_cap, uint32_t pool_cap) {
            return crass_b200_ac_scan_dev(l.ctx, ac, s.d_bases.as<uint8_t>(), s.d_offsets.as<uint64_t>(), s.n(), (uint32_t)max_len,
                                          skip ? s.d_found.as<uint8_t>() : nullptr, l.d_found2.as<uint8_t>(), l.d_hits.as<crass_b200_hit>(),
                                          hits_cap, l.d_pool.as<uint32_t>(), pool_cap, l.d_cnt.as<uint32_t>(), l.stream);
        };
        return collect_hits(l, s.n(), l.d_found2.as<uint8_t>(), launch, parts[l.index], 0);
    });
    if (rc) return rc;
    e->t_phase2 += now_ms() - t0;
    std::vector<crass_b200_hit> hits; std::vector<uint32_t> pool;
    concat_hits(f->shards, parts, hits, pool);
    crass_b200_hit* h = (crass_b200_hit*)malloc(sizeof(crass_b200_hit) * (hits.size() ? hits.size() : 1));
    uint32_t* p = (uint32_t*)malloc(sizeof(uint32_t) * (pool.size() ? pool.size() : 1));
    if (!h || !p) { free(h); free(p); return cbh::fail(CRASS_B200_ENOMEM, "malloc"); }
    if (!hits.empty()) memcpy(h, hits.data(), hits.size() * sizeof(crass_b200_hit));
    if (!pool.empty()) memcpy(p, pool.data(), pool.size() * sizeof(uint32_t));
    *hits_out = h; *n_hits = (uint32_t)hits.size(); *pool_out = p; *n_pool = (uint32_t)pool.size();
    if (batch_out) *batch_out = f->batch;
    return 0;
}

// searchFile, streamed: what crass_b200_engine_search_file does, but the file goes through the devices range by range
// (stream_phase1) and `fn` gets the hits of each range, in file order, on ONE helper thread while later ranges are still being
// parsed and searched.  The caller is inside this call the whole time, so its containers are only ever touched by that one
// thread.  A non-zero return of fn ends the run with that code.  The ranges stay resident for
// crass_b200_engine_find_singletons_ranges and go with crass_b200_engine_release_file(path).
int crass_b200_engine_search_file_ranges(crass_b200_engine* e, const char* path, const crass_b200_params* params,
                                         crass_b200_range_fn fn, void* user, int* max_read_len) {
    if (!e || !path || !params || !fn) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    crass_b200_engine_release_file(e, path);
    cbh::ParseStream* ps = cbh::parse_stream_open(path, e->stream_bytes);       // 0: the whole file is one range
    if (!ps) return CRASS_B200_EIO;
    int max_len = 0;
    std::vector<std::string> range_paths;
    const int rc = stream_phase1(e, ps, &path, 1, params, [&](const SearchedRange& r, uint64_t first_read) {
        return fn(user, r.batch, r.hits, r.nh, r.pool, r.np, first_read);
    }, &max_len, range_paths);
    cbh::parse_stream_close(ps);
    e->ranged[path] = range_paths;
    if (rc) { crass_b200_engine_release_file(e, path); return rc; }
    if (max_read_len) *max_read_len = max_len;
    return 0;
}

// findSingletons over the ranges crass_b200_engine_search_file_ranges left resident (a file it has not seen is searched as a
// whole, parsed and copied in now): a helper thread runs K2 range after range, `fn` is called on THIS thread in file order.
int crass_b200_engine_find_singletons_ranges(crass_b200_engine* e, const char* path, const crass_b200_ac* ac, int skip_found,
                                             crass_b200_range_fn fn, void* user) {
    if (!e || !path || !ac || !fn) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    std::vector<std::string> names;
    auto it = e->ranged.find(path);
    if (it != e->ranged.end()) names = it->second; else names.push_back(path);
    StageQueue<SearchedRange> scanned;
    int scan_rc = 0; std::string scan_err;
    std::thread scanner([&]() {
        try {
            for (const std::string& nm : names) {
                SearchedRange r;
                scan_rc = crass_b200_engine_find_singletons(e, nm.c_str(), ac, skip_found, &r.batch, &r.hits, &r.nh, &r.pool, &r.np);
                if (scan_rc) { scan_err = crass_b200_last_error(); break; }
                scanned.push(r);
            }
        } catch (std::exception& ex) { scan_rc = CRASS_B200_ENOMEM; scan_err = std::string("streamed scan: ") + ex.what(); }
        scanned.close();
    });
    int rc = 0;
    uint64_t first_read = 0;
    SearchedRange r;
    while (scanned.pop(r)) {
        if (!rc) rc = fn(user, r.batch, r.hits, r.nh, r.pool, r.np, first_read);
        first_read += crass_b200_batch_num_reads(r.batch);
        free(r.hits); free(r.pool);
    }
    scanner.join();
    if (scan_rc) return cbh::fail(scan_rc, scan_err);
    return rc;
}

// The exchange step for the files searched so far: every device de-duplicates the DR tokens of its most recent search (K4b)
// into a token block, the blocks meet on the first device (NCCL all-gather, or peer copies), are merged there in
// first-appearance order (K4c) and clustered (K5 + host passes): *ac_out is the matcher for phase 2 (NULL when no DR was
// found), *patterns_out (optional, malloc'd) the non-redundant set as '\n'-separated text.
// Valid for ONE searched file (the token records of a lane belong to its last search); run_files_multi below falls back
// to the host containers' token list when there are several files.
int crass_b200_engine_exchange(crass_b200_engine* e, const char* path, uint32_t kmer_clust, crass_b200_ac** ac_out,
                               uint32_t* n_variants, uint32_t* n_patterns) {
    if (!e || !path || !ac_out) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    *ac_out = nullptr;
    FileState* f = find_file(e, path);
    if (!f) return cbh::fail(CRASS_B200_EINVAL, "file was not searched through this engine");
    const uint32_t G = (uint32_t)e->lanes.size();
    const double t0 = now_ms();
    uint32_t shard_reads = 1;
    for (auto& s : f->shards) shard_reads = std::max(shard_reads, s->n());
    const uint32_t kTokStride = e->tok_stride;
    for (int round = 0; round < 8; ++round) {
        const uint32_t cap = e->block_cap;
        const size_t nb = crass_b200_token_block_bytes(cap, kTokStride);
        const uint32_t out_cap = cap * std::min<uint32_t>(G, 4);
        // K4b on every device
        int rc = for_each_lane(e, [&](Lane& l) -> int {
            if (int r = l.d_block.reserve(nb)) return r;
            if (l.index == 0 || e->use_nccl) { if (int r = l.d_recv.reserve(nb * G)) return r; }
            const uint32_t nh = f->shards[l.index]->n() ? l.h_cnt.as<uint32_t>()[0] : 0;     // hits of the lane's last search
            return crass_b200_unique_tokens_block_dev(l.ctx, l.d_hits.as<crass_b200_hit>(), nh, l.d_tokens.p, kTokStride, l.d_block.p, cap, l.stream);
        });
        if (rc) return rc;
        Lane& root = *e->lanes[0];
        if (G > 1 && e->use_nccl) {
            // one all-gather of the fixed-size blocks over NVLink; group semantics: one thread issues all ranks' calls
            e->nccl.GroupStart();
            int nrc = 0;
            for (auto& l : e->lanes) {
                cudaSetDevice(l->device);
                const int r1 = e->nccl.AllGather(l->d_block.p, l->d_recv.p, nb, /*ncclUint8*/ 1, l->comm, l->stream);
                if (r1 && !nrc) nrc = r1;
            }
            const int r2 = e->nccl.GroupEnd();
            if (nrc || r2) return cbh::fail(CRASS_B200_ECUDA, std::string("ncclAllGather failed: ") + (e->nccl.GetErrorString ? e->nccl.GetErrorString(nrc ? nrc : r2) : "?"));
        } else if (G > 1) {
            for (auto& l : e->lanes) { cudaSetDevice(l->device); ENG_CUDA(cudaStreamSynchronize(l->stream)); }
            cudaSetDevice(root.device);
            for (auto& l : e->lanes)
                ENG_CUDA(cudaMemcpyPeerAsync(root.d_recv.as<uint8_t>() + nb * (size_t)l->index, root.device, l->d_block.p, l->device, nb, root.stream));
        }
        cudaSetDevice(root.device);
        const void* merged = root.d_block.p;
        uint32_t mcap = cap;
        if (G > 1) {
            if (int r = root.d_merged.reserve(crass_b200_token_block_bytes(out_cap, kTokStride))) return r;
            if (int r = crass_b200_merge_token_blocks_dev(root.ctx, root.d_recv.p, G, cap, kTokStride, shard_reads, root.d_merged.p, out_cap, root.stream)) return r;
            merged = root.d_merged.p; mcap = out_cap;
        }
        uint32_t count = 0, flags = 0, npat = 0;
        if (int r = crass_b200_cluster_block_dev(root.ctx, merged, mcap, kTokStride, kmer_clust, ac_out, &count, &flags, &npat, root.stream)) return r;
        if (flags & 2) return cbh::fail(CRASS_B200_EINVAL, "token stride too small for the DR lengths in use");
        if ((flags & 1) || count > mcap) { e->block_cap *= 2; continue; }       // a block overflowed: larger blocks, once more
        if (n_variants) *n_variants = count;
        if (n_patterns) *n_patterns = npat;
        e->t_exchange += now_ms() - t0;
        return 0;
    }
    return cbh::fail(CRASS_B200_EOVERFLOW, "token blocks kept overflowing");
}

// WorkHorse::parseSeqFiles on the engine's devices: searchFile for every path, createNonRedundantSet, findSingletons for
// every path; *out holds the containers a single sequential run would have filled.
static int run_files_impl(crass_b200_engine* e, const char* const* paths, uint32_t n_paths, const crass_b200_params* params,
                          int phases, crass_b200_results** out, int* max_read_len);

int crass_b200_engine_run_files(crass_b200_engine* e, const char* const* paths, uint32_t n_paths, const crass_b200_params* params,
                                int phases, crass_b200_results** out, int* max_read_len) {
    if (!e || !paths || !params || !out) return cbh::fail(CRASS_B200_EINVAL, "NULL argument");
    const int rc = run_files_impl(e, paths, n_paths, params, phases, out, max_read_len);
    e->runs_done.fetch_add(1);                                                // (an engine that is used again page-locks its pooled buffers)
    return rc;
}

static int run_files_impl(crass_b200_engine* e, const char* const* paths, uint32_t n_paths, const crass_b200_params* params,
                          int phases, crass_b200_results** out, int* max_read_len) {
    e->t_parse = e->t_phase1 = e->t_exchange = e->t_phase2 = e->t_replay = 0;
    crass_b200_results* res = nullptr;
    if (int r = crass_b200_results_create(&res)) return r;
    int rc = 0, max_len = 0;
    const char* xsel = getenv("CRASS_B200_EXCHANGE");
    // A large file is STREAMED (the reference's O(1)-memory kseq loop turned into a pipeline): the parser hands
    // out ranges of about stream_bytes, each ending on a true record start; while its worker threads parse range i+1, a second
    // thread copies range i to the devices and runs K1 on it, and a third replays the hits of range i-1 into the containers.
    // The ranges then behave like the files of a multi-file run (token order from the containers, K2 over the resident ranges).
    // Several files (paired-end reads) go through the same pipeline one after the other, whatever their sizes.
    if (e->stream_bytes && n_paths >= 1) {
        for (uint32_t f = 0; f < n_paths; ++f)
            if (!paths[f] || (strcmp(paths[f], "-") != 0 && access(paths[f], R_OK) != 0)) {
                crass_b200_results_destroy(res);
                return cbh::fail(CRASS_B200_EIO, std::string("cannot open ") + (paths[f] ? paths[f] : "(null)"));
            }
        cbh::ParseStream* ps = cbh::parse_stream_open(paths[0], e->stream_bytes);
        if (!ps) { crass_b200_results_destroy(res); return CRASS_B200_EIO; }
        if (n_paths >= 2 || cbh::parse_stream_size(ps) >= 2 * e->stream_bytes) {
            rc = run_streamed(e, ps, paths, n_paths, params, phases, res, &max_len);
            cbh::parse_stream_close(ps);
            if (e->trace)
                fprintf(stderr, "[crass_b200] engine run (streamed): parse %.1f ms, phase 1 (H2D + K1 + D2H) %.1f ms, clustering %.1f ms, phase 2 %.1f ms, replay %.1f ms (stages overlap)\n",
                        e->t_parse, e->t_phase1, e->t_exchange, e->t_phase2, e->t_replay);
            if (rc) { crass_b200_results_destroy(res); return rc; }
            if (max_read_len) *max_read_len = max_len;
            *out = res;
            return 0;
        }
        cbh::parse_stream_close(ps);
    }
    // One file (the common case): the token exchange on the devices numbers the tokens of the whole input, so nothing the
    // devices do next waits for the host containers -- the phase-1 hits are replayed on a helper thread while the devices
    // exchange, cluster and scan; only the phase-2 replay (which tests readsFound) has to come after it.
    if (n_paths == 1 && phases >= 2 && !(xsel && !strcmp(xsel, "host"))) {
        const crass_b200_batch* b = nullptr;
        crass_b200_hit* hits = nullptr; uint32_t nh = 0, np = 0; uint32_t* pool = nullptr;
        rc = crass_b200_engine_search_file(e, paths[0], params, &b, &hits, &nh, &pool, &np);
        if (!rc) {
            max_len = (int)crass_b200_batch_max_read_len(b);
            int replay_rc = 0; std::string replay_err; double replay_ms = 0;
            std::thread replay([&]() {
                const double t0 = now_ms();
                replay_rc = crass_b200_results_add_phase1(res, b, hits, nh, pool);
                if (replay_rc) replay_err = crass_b200_last_error();
                replay_ms = now_ms() - t0;
            });
            crass_b200_ac* ac = nullptr;
            uint32_t nv = 0, n_pat = 0;
            crass_b200_hit* hits2 = nullptr; uint32_t nh2 = 0, np2 = 0; uint32_t* pool2 = nullptr;
            rc = crass_b200_engine_exchange(e, paths[0], params->kmer_clust, &ac, &nv, &n_pat);
            if (!rc && ac) rc = crass_b200_engine_find_singletons(e, paths[0], ac, 1, &b, &hits2, &nh2, &pool2, &np2);
            replay.join();
            e->t_replay += replay_ms;
            if (!rc && replay_rc) rc = cbh::fail(replay_rc, replay_err);
            if (!rc && ac) {
                const double t0 = now_ms();
                rc = crass_b200_results_add_phase2(res, b, hits2, nh2, pool2);
                e->t_replay += now_ms() - t0;
            }
            res->r.lazy_kmer_clust = (int)params->kmer_clust;                 // groups / pattern list of the dump: computed when asked for
            free(hits2); free(pool2);
            crass_b200_ac_destroy(ac);
        }
        free(hits); free(pool);
    } else {
    for (uint32_t f = 0; f < n_paths && !rc; ++f) {
        const crass_b200_batch* b = nullptr;
        crass_b200_hit* hits = nullptr; uint32_t nh = 0, np = 0; uint32_t* pool = nullptr;
        rc = crass_b200_engine_search_file(e, paths[f], params, &b, &hits, &nh, &pool, &np);
        if (!rc) {
            max_len = std::max(max_len, (int)crass_b200_batch_max_read_len(b));
            const double t0 = now_ms();
            rc = crass_b200_results_add_phase1(res, b, hits, nh, pool);
            e->t_replay += now_ms() - t0;
        }
        free(hits); free(pool);
    }
    if (!rc && phases >= 2) {
        // several files: the containers' token list (filled in read order above) is the sequential numbering
        crass_b200_ac* ac = nullptr;
        uint32_t n_pat = 0;
        const double t0 = now_ms();
        char* pats = crass_b200_results_non_redundant(res, params->kmer_clust, &n_pat);
        free(pats);
        if (n_pat) {
            std::vector<uint8_t> bytes; std::vector<uint32_t> offs(1, 0);
            for (const std::string& p : res->r.non_redundant) { bytes.insert(bytes.end(), p.begin(), p.end()); offs.push_back((uint32_t)bytes.size()); }
            rc = crass_b200_ac_build(bytes.data(), offs.data(), n_pat, &ac);
        }
        e->t_exchange += now_ms() - t0;
        if (!rc && ac) {                                                      // WorkHorse.cpp:373 guards the empty set
            for (uint32_t f = 0; f < n_paths && !rc; ++f) {
                const crass_b200_batch* b = nullptr;
                crass_b200_hit* hits = nullptr; uint32_t nh = 0, np = 0; uint32_t* pool = nullptr;
                rc = crass_b200_engine_find_singletons(e, paths[f], ac, 1, &b, &hits, &nh, &pool, &np);
                if (!rc) {
                    const double t1 = now_ms();
                    rc = crass_b200_results_add_phase2(res, b, hits, nh, pool);
                    e->t_replay += now_ms() - t1;
                }
                free(hits); free(pool);
            }
        }
        crass_b200_ac_destroy(ac);
    } else if (!rc) {
        res->r.lazy_kmer_clust = (int)params->kmer_clust;
    }
    }
    for (uint32_t f = 0; f < n_paths; ++f) crass_b200_engine_release_file(e, paths[f]);
    if (e->trace)
        fprintf(stderr, "[crass_b200] engine run: parse %.1f ms, phase 1 (H2D + K1 + D2H) %.1f ms, exchange + clustering %.1f ms, phase 2 %.1f ms, replay %.1f ms\n",
                e->t_parse, e->t_phase1, e->t_exchange, e->t_phase2, e->t_replay);
    if (rc) { crass_b200_results_destroy(res); return rc; }
    if (max_read_len) *max_read_len = max_len;
    *out = res;
    return 0;
}

int crass_b200_run_files_multi(const int* devices, uint32_t n_devices, const char* const* paths, uint32_t n_paths,
                               const crass_b200_params* params, int phases, crass_b200_results** out, int* max_read_len) {
    crass_b200_engine* e = nullptr;
    if (int r = crass_b200_engine_create(devices, n_devices, &e)) return r;
    const int rc = crass_b200_engine_run_files(e, paths, n_paths, params, phases, out, max_read_len);
    crass_b200_engine_destroy(e);
    return rc;
}

}  // extern "C"
