// kernels.cuh -- sm_100a kernels of the read-scanning hot path (device code only).
//
//   K1  k_dr_search_generic   searchCore for any parameter set / read length, one thread per read
//   K2  k_ac_scan_generic     first-match multi-pattern scan, one thread per read
//   K3  k_edit_distance       batched modified edit distance + similarity, one thread per pair
//   KAT k_scan_right_one / k_extend_one : single-read entry points for the reference's unit-test vectors
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/crass_b200.h"
#include "dr_core.cuh"
#include "dr_filter.cuh"
#include "dr_long.cuh"
#include "sw_core.cuh"
#include "sw_warp.cuh"

namespace cbk {

using cb::Params;

struct HitSink {
    crass_b200_hit* hits;
    uint32_t hits_cap;
    uint32_t* pool;
    uint32_t pool_cap;
    uint32_t* counters;      // [0] hits, [1] pool entries, [2] overflow flag, [3] exact-path reads
    uint8_t* tokens;         // optional: one DR token record per hit slot (K4), NULL = off
    uint32_t token_stride;
};

// global-memory byte accessor through the read-only path
struct GmemSeq {
    const uint8_t* p;
    __device__ __forceinline__ uint8_t operator[](uint32_t i) const { return __ldg(p + i); }
};

__device__ __forceinline__ uint32_t emit_hit(const HitSink& sink, uint32_t read_index, const uint32_t* ss, uint32_t n_ss, uint32_t replen) {
    const uint32_t slot = atomicAdd(&sink.counters[0], 1u);
    const uint32_t off = atomicAdd(&sink.counters[1], n_ss);
    if (slot < sink.hits_cap && off + n_ss <= sink.pool_cap) {
        crass_b200_hit h;
        h.read_index = read_index; h.n_ss = n_ss; h.ss_offset = off; h.repeat_len = replen;
        sink.hits[slot] = h;
        for (uint32_t i = 0; i < n_ss; ++i) sink.pool[off + i] = ss[i];
        return slot;
    }
    sink.counters[2] = 1u;
    return 0xFFFFFFFFu;
}

// ---- K4: the low-lexi DR token of a hit (ReadHolder::DRLowLexi, ReadHolder.cpp:513-591), computed where the hit is
// found so that the multi-GPU merge / pattern-set construction never has to touch the reads on the host.
// Record layout (stride bytes): [0] length, [1] 1 = read kept its orientation (RH_WasLowLexi), [2..] the token.
__constant__ uint8_t c_comp_tab[128] = {                       // SeqUtils.cpp:51-60
    0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31,
    32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 63,
    64, 'T', 'V', 'G', 'H', 'E', 'F', 'C', 'D', 'I', 'J', 'M', 'L', 'K', 'N', 'O', 'P', 'Q', 'Y', 'S', 'A', 'A', 'B', 'W', 'X', 'R', 'Z', 91, 92, 93, 94, 95,
    64, 't', 'v', 'g', 'h', 'e', 'f', 'c', 'd', 'i', 'j', 'm', 'l', 'k', 'n', 'o', 'p', 'q', 'y', 's', 'a', 'a', 'b', 'w', 'x', 'r', 'z', 123, 124, 125, 126, 127};

// complement of a base: the four letters that make up nearly every read are settled in registers; the table (constant
// memory, one serialised access per distinct index in a warp) is only read for the other IUPAC letters
__device__ __forceinline__ uint8_t comp_byte(uint8_t c) {
    if (c == 'A') return 'T';
    if (c == 'T') return 'A';
    if (c == 'C') return 'G';
    if (c == 'G') return 'C';
    return c_comp_tab[c & 127];
}

template <class Seq>
__device__ __noinline__ void emit_token(uint8_t* __restrict__ rec, uint32_t stride, const Seq& s, uint32_t L, const uint32_t* ss, uint32_t n_ss) {
    const uint32_t n_rep = n_ss / 2;
    uint32_t idx;
    if (n_rep == 1) idx = 0;
    else if (n_rep == 2) {
        if (ss[0] == 0) idx = 2;
        else if (ss[3] == L) idx = 0;
        else idx = ((int)(ss[1] - ss[0]) > (int)(ss[3] - ss[2])) ? 0 : 2;
    } else idx = 2;
    uint32_t st = ss[idx], ln = ss[idx + 1] - ss[idx] + 1;
    if (st > L) st = L;
    if (ln > L - st) ln = L - st;
    if (ln > stride - 2) ln = stride - 2;                        // cannot happen: stride >= high_dr + 2
    // forward < reverse complement ?  (std::string operator<, unsigned bytes; equal -> take the reverse complement)
    bool fwd_less = false;
    for (uint32_t i = 0; i < ln; ++i) {
        const uint8_t a = s[st + i], b = comp_byte(s[st + ln - 1 - i]);
        if (a != b) { fwd_less = a < b; break; }
    }
    rec[0] = (uint8_t)ln;
    rec[1] = fwd_less ? 1 : 0;
    for (uint32_t i = 0; i < ln; ++i) rec[2 + i] = fwd_less ? s[st + i] : comp_byte(s[st + ln - 1 - i]);
}

// ---- K4b: distinct DR tokens of a hit list, with the smallest read index that carries each --------------------
// (what StringCheck::addString's first-appearance numbering needs).  Open-addressing table keyed by the token bytes:
// rep[i] = hit slot of the first thread that claimed entry i, first_read[i] = min read index over all equal tokens.
__device__ __forceinline__ uint32_t token_hash(const uint8_t* rec) {
    uint32_t h = 2166136261u;
    const uint32_t n = rec[0];
    for (uint32_t i = 0; i < n; ++i) { h ^= rec[2 + i]; h *= 16777619u; }
    return h ^ (h >> 15);
}

__device__ __forceinline__ bool token_equal(const uint8_t* a, const uint8_t* b) {
    if (a[0] != b[0]) return false;
    const uint32_t n = a[0];
    for (uint32_t i = 0; i < n; ++i) if (a[2 + i] != b[2 + i]) return false;
    return true;
}

__global__ void __launch_bounds__(256)
k_token_dedupe(const crass_b200_hit* __restrict__ hits, uint32_t n_hits, const uint8_t* __restrict__ tokens, uint32_t stride,
               uint32_t* __restrict__ rep, uint32_t* __restrict__ first_read, uint32_t table_mask) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_hits) return;
    const uint8_t* mine = tokens + (size_t)k * stride;
    const uint32_t read = hits[k].read_index;
    uint32_t i = token_hash(mine) & table_mask;
    for (;;) {
        uint32_t cur = atomicCAS(&rep[i], 0xFFFFFFFFu, k);
        if (cur == 0xFFFFFFFFu) cur = k;
        if (cur == k || token_equal(mine, tokens + (size_t)cur * stride)) { atomicMin(&first_read[i], read); return; }
        i = (i + 1) & table_mask;
    }
}

__global__ void __launch_bounds__(256)
k_token_compact(const uint32_t* __restrict__ rep, const uint32_t* __restrict__ first_read, uint32_t table_size,
                const uint8_t* __restrict__ tokens, uint32_t stride, uint8_t* __restrict__ out_tokens,
                uint32_t* __restrict__ out_first_read, uint32_t* __restrict__ out_count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= table_size) return;
    const uint32_t r = rep[i];
    if (r == 0xFFFFFFFFu) return;
    const uint32_t j = atomicAdd(out_count, 1u);
    out_first_read[j] = first_read[i];
    const uint8_t* src = tokens + (size_t)r * stride;
    uint8_t* dst = out_tokens + (size_t)j * stride;
    for (uint32_t b = 0; b < stride; b += 4) *reinterpret_cast<uint32_t*>(dst + b) = *reinterpret_cast<const uint32_t*>(src + b);
}

// ---- hit records into read order on the device -----------------------------------------------------------------
// Every read has at most one hit and found[r] != 0 marks exactly the reads that have one, so the position of a hit in
// read order is the number of marked reads before it: per-chunk counts, one exclusive scan, then every hit counts the
// marked reads of its own chunk in front of it and moves its record there.  (The flags are 10 MB for 10 M reads and
// stay in L2 between the three launches.)
constexpr uint32_t kRankChunk = 1024;                // reads per chunk

__device__ __forceinline__ uint32_t count_marked16(const uint4& v) {                   // 16 flag bytes, each 0 or 1
    return __popc(v.x & 0x01010101u) + __popc(v.y & 0x01010101u) + __popc(v.z & 0x01010101u) + __popc(v.w & 0x01010101u);
}

__global__ void __launch_bounds__(64)
k_rank_chunks(const uint8_t* __restrict__ found, uint32_t n_reads, uint32_t* __restrict__ chunk_count) {
    const uint32_t c = blockIdx.x;                                                      // one chunk per CTA, 16 bytes per thread
    const uint32_t r = c * kRankChunk + threadIdx.x * 16u;
    uint32_t s = 0;
    if (r + 16 <= n_reads) s = count_marked16(*reinterpret_cast<const uint4*>(found + r));
    else for (uint32_t i = r; i < n_reads; ++i) s += found[i] & 1u;
    s = __reduce_add_sync(0xFFFFFFFFu, s);
    __shared__ uint32_t w[2];
    if ((threadIdx.x & 31u) == 0) w[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) chunk_count[c] = w[0] + w[1];
}

constexpr uint32_t kScanItems = 8;                   // consecutive entries per thread: 8192 per pass of the one CTA

__global__ void __launch_bounds__(1024)
k_rank_scan(uint32_t* __restrict__ chunk_count, uint32_t n_chunks, const uint32_t* __restrict__ n_dev = nullptr) {   // in place: counts -> exclusive prefix
    if (n_dev) n_chunks = min(n_chunks, *n_dev + 1u);                                 // only the part that is in use (+ the total)
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const bool aligned = (reinterpret_cast<uintptr_t>(chunk_count) & 15u) == 0;
    for (uint32_t base = 0; base < n_chunks; base += 1024 * kScanItems) {
        const uint32_t i0 = base + threadIdx.x * kScanItems;
        uint32_t v[kScanItems];
        if (aligned && i0 + kScanItems <= n_chunks) {
            const uint4 a = *reinterpret_cast<const uint4*>(chunk_count + i0), b = *reinterpret_cast<const uint4*>(chunk_count + i0 + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (uint32_t k = 0; k < kScanItems; ++k) v[k] = i0 + k < n_chunks ? chunk_count[i0 + k] : 0u;
        }
        uint32_t mine = 0;
#pragma unroll
        for (uint32_t k = 0; k < kScanItems; ++k) mine += v[k];
        uint32_t x = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d); if ((threadIdx.x & 31) >= d) x += y; }
        if ((threadIdx.x & 31u) == 31u) warp_sum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t s = warp_sum[threadIdx.x];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, s, d); if ((int)threadIdx.x >= d) s += y; }
            warp_sum[threadIdx.x] = s;                                                  // inclusive over the warps
        }
        __syncthreads();
        uint32_t before = carry + (threadIdx.x >= 32 ? warp_sum[(threadIdx.x >> 5) - 1] : 0u) + x - mine;
        const uint32_t after_me = before + mine;
#pragma unroll
        for (uint32_t k = 0; k < kScanItems; ++k) { const uint32_t t = v[k]; v[k] = before; before += t; }
        if (aligned && i0 + kScanItems <= n_chunks) {
            *reinterpret_cast<uint4*>(chunk_count + i0) = make_uint4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<uint4*>(chunk_count + i0 + 4) = make_uint4(v[4], v[5], v[6], v[7]);
        } else {
#pragma unroll
            for (uint32_t k = 0; k < kScanItems; ++k) if (i0 + k < n_chunks) chunk_count[i0 + k] = v[k];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry = after_me;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
k_rank_scatter(const uint8_t* __restrict__ found, const uint32_t* __restrict__ chunk_before, const crass_b200_hit* __restrict__ hits,
               const uint32_t* __restrict__ n_hits_dev, uint32_t max_hits, crass_b200_hit* __restrict__ sorted) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_hits = min(*n_hits_dev, max_hits);
    if (k >= n_hits) return;
    const crass_b200_hit h = hits[k];
    const uint32_t c0 = (h.read_index / kRankChunk) * kRankChunk;
    uint32_t rank = chunk_before[h.read_index / kRankChunk];
    uint32_t r = c0;
    for (; r + 16 <= h.read_index; r += 16) rank += count_marked16(*reinterpret_cast<const uint4*>(found + r));
    for (; r < h.read_index; ++r) rank += found[r] & 1u;
    if (rank < max_hits) sorted[rank] = h;
}

// ---- K5: the data-parallel passes of createNonRedundantSet on the token block ---------------------------
// Input: a token block (K4b/K4c output, records in arbitrary order with their first-appearance key).  Output, all in
// token order t = rank of the key: order[t] = record slot, koff[t] = offset of DR t's 11-mers, keys[q] = canonical
// 22-bit key of every 11-mer (min(forward, reverse complement) with A<C<G<T, = laurenize, SeqUtils.cpp:89-97;
// 0xFFFFFFFF for the rare 11-mers whose canonical form holds another letter: the host sends those through a string map)
// and first[q] = the first DR in token order that contains the 11-mer (what clusterDRReads' k-mer map answers,
// WorkHorse.cpp:1404-1637).  The order-dependent walk over these arrays stays on the host (results.cpp, pass C).
constexpr uint32_t kTokenBlockHeader = 16;          // bytes in front of the records of a token block (see K4b/K4c below)
constexpr uint32_t kClKmer = 11;
constexpr uint32_t kClStr = 0xFFFFFFFFu;

__device__ __forceinline__ int cl_code(uint8_t c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }

struct ClusterArrays {
    const uint8_t* block;        // token block
    uint32_t cap, stride;
    uint32_t* order;             // [cap]
    uint32_t* koff;              // [cap + 1]
    uint32_t* keys;              // [cap * (stride - 16)]
    uint32_t* first;             // same size
    uint32_t* tab_key;           // open-addressing table: key -> smallest t
    uint32_t* tab_val;
    uint32_t tab_mask;
    uint32_t* info;              // [0] n, [1] total k-mers, [2] k-mers that need the string map
    uint32_t* str_tq;            // (t, q) of the first str_cap of those, in no particular order
    uint32_t str_cap;
    uint32_t* ckeys;             // [cap] the order keys gathered from the records (k_cl_gather), what k_cl_rank reads
    uint32_t max_n;              // lists longer than this are left to the host (the rank kernel is O(n^2)): the kernels then see n = 0
    __device__ uint32_t n() const { const uint32_t k = min(*reinterpret_cast<const uint32_t*>(block), cap); return k > max_n ? 0u : k; }
    __device__ const uint8_t* rec(uint32_t slot) const { return block + kTokenBlockHeader + (size_t)slot * stride; }
    __device__ uint32_t key_of(uint32_t slot) const { return *reinterpret_cast<const uint32_t*>(rec(slot) + stride - 4); }
    __device__ uint32_t len_of(uint32_t slot) const { const uint32_t l = rec(slot)[0]; return l + 6u <= stride ? l : stride - 6u; }
};

// token position of every record = number of records with a smaller key (keys are distinct: one token per first read).
// O(n^2), so it is spread wide: a CTA ranks 64 records, eight threads per record each counting an eighth of every
// 512-key tile (read as 16-byte vectors from shared memory).
__global__ void __launch_bounds__(256)
k_cl_gather(ClusterArrays a) {                       // the keys sit 64 bytes apart in the block: read them once, not once per CTA
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i < a.n()) a.ckeys[i] = a.key_of(i);
}

constexpr uint32_t kClRankThreads = 64;              // records per CTA
constexpr uint32_t kClRankParts = 8;

__global__ void __launch_bounds__(kClRankThreads * kClRankParts)
k_cl_rank(ClusterArrays a) {
    __shared__ __align__(16) uint32_t tile[kClRankThreads * kClRankParts];
    __shared__ uint32_t partial[kClRankParts][kClRankThreads];
    const uint32_t n = a.n();
    const uint32_t il = threadIdx.x % kClRankThreads, part = threadIdx.x / kClRankThreads;
    const uint32_t i = blockIdx.x * kClRankThreads + il;
    if (blockIdx.x * kClRankThreads >= n) {                          // CTAs past the list only clear their part of koff
        if (part == 0 && i <= a.cap) a.koff[i] = 0;
        if (i == 0 && part == 0) a.info[0] = n;
        return;
    }
    const uint32_t mine = i < n ? a.ckeys[i] : 0u;
    uint32_t rank = 0;
    for (uint32_t j0 = 0; j0 < n; j0 += kClRankThreads * kClRankParts) {
        __syncthreads();
        tile[threadIdx.x] = j0 + threadIdx.x < n ? a.ckeys[j0 + threadIdx.x] : 0xFFFFFFFFu;   // padding never counts
        __syncthreads();
        const uint4* t4 = reinterpret_cast<const uint4*>(tile + part * kClRankThreads);
#pragma unroll
        for (uint32_t j = 0; j < kClRankThreads / 4; ++j) {
            const uint4 k = t4[j];
            rank += (k.x < mine ? 1u : 0u) + (k.y < mine ? 1u : 0u) + (k.z < mine ? 1u : 0u) + (k.w < mine ? 1u : 0u);
        }
    }
    partial[part][il] = rank;
    __syncthreads();
    if (part != 0) return;
    rank = 0;
#pragma unroll
    for (uint32_t k = 0; k < kClRankParts; ++k) rank += partial[k][il];
    if (i < n) {
        a.order[rank] = i;
        const uint32_t len = a.len_of(i);
        a.koff[rank] = len >= kClKmer ? len - kClKmer + 1 : 0u;     // counts; k_rank_scan turns them into offsets
    } else if (i <= a.cap) {
        a.koff[i] = 0;
    }
    if (i == 0) a.info[0] = n;
}

// pass A + B: keys of DR t and the hash table key -> min t.  One warp per DR, one lane per 11-mer.
constexpr uint32_t kClKeysThreads = 128;

__global__ void __launch_bounds__(kClKeysThreads)
k_cl_keys(ClusterArrays a) {
    const uint32_t n = a.n();
    const uint32_t t = (blockIdx.x * kClKeysThreads + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (t == 0 && lane == 0) a.info[1] = a.koff[n];
    if (t >= n) return;
    const uint32_t slot = a.order[t];
    const uint8_t* dr = a.rec(slot) + 2;
    const uint32_t q0 = a.koff[t], nk = a.koff[t + 1] - q0;
    for (uint32_t p = lane; p < nk; p += 32) {
        const uint8_t* k = dr + p;                                   // the window dr[p, p+11)
        uint32_t fw = 0, rc = 0;
        bool valid = true;
#pragma unroll
        for (uint32_t i = 0; i < kClKmer; ++i) {
            const int c = cl_code(k[i]);
            valid = valid && c >= 0;
            fw = (fw << 2) | (uint32_t)(c & 3);
            rc |= (uint32_t)((3 - c) & 3) << (2 * i);
        }
        uint32_t key = kClStr;
        if (valid) key = fw < rc ? fw : rc;
        else {                                                       // another letter in the window: canonical form on the bytes
            int cmp = 0;                                             // k vs its reverse complement, as unsigned bytes
            for (uint32_t i = 0; i < kClKmer && !cmp; ++i) {
                const uint8_t x = k[i], y = c_comp_tab[k[kClKmer - 1 - i] & 127];
                cmp = x < y ? -1 : x > y ? 1 : 0;
            }
            uint32_t k2 = 0;
            bool acgt = true;                                        // e.g. a 'U' whose reverse complement is all A/C/G/T
            for (uint32_t i = 0; i < kClKmer && acgt; ++i) {
                const int c2 = cl_code(cmp < 0 ? k[i] : c_comp_tab[k[kClKmer - 1 - i] & 127]);
                if (c2 < 0) acgt = false; else k2 = (k2 << 2) | (uint32_t)c2;
            }
            if (acgt) key = k2;
        }
        const uint32_t q = q0 + p;
        a.keys[q] = key;
        if (key == kClStr) {                                         // the host resolves these; tell it where they are
            const uint32_t i = atomicAdd(&a.info[2], 1u);
            if (i < a.str_cap) { a.str_tq[2 * i] = t; a.str_tq[2 * i + 1] = q; }
            a.first[q] = kClStr;
        } else {
            uint32_t s = (key * 0x9E3779B1u) & a.tab_mask;
            for (;;) {
                const uint32_t cur = atomicCAS(&a.tab_key[s], kClStr, key);
                if (cur == kClStr || cur == key) break;
                s = (s + 1) & a.tab_mask;
            }
            atomicMin(&a.tab_val[s], t);
            a.first[q] = s;
        }
    }
}

// pass B2, first half: slot -> first DR
__global__ void __launch_bounds__(256)
k_cl_first(ClusterArrays a) {
    const uint32_t total = a.koff[a.n()];
    for (uint32_t q = blockIdx.x * 256 + threadIdx.x; q < total; q += gridDim.x * 256)
        if (a.keys[q] != kClStr) a.first[q] = a.tab_val[a.first[q]];
}

// pass D (removeRedundantRepeats, WorkHorse.cpp:612-645): DR b is redundant when an earlier DR of its group -- shorter, or
// equally long with a smaller token position -- is inside it on either strand.  (The reference only remembers survivors;
// containment is transitive, so "any earlier DR" gives the same answer and needs no order of evaluation.)  A DR a inside
// b shows its first 11-mer where it starts, its reverse complement shows the same canonical 11-mer where it ends: the
// candidates for b come from chains keyed by (group, first key) and are confirmed on the bytes.  DRs without an integer
// first key (another letter in the first window, or shorter than a k-mer) are few and are tried against every b.
struct ReduceArrays {
    ClusterArrays a;
    const uint32_t* group;       // [n] 1-based group of DR t (from the host walk)
    unsigned long long* ckey;    // chain table: (group << 32 | first key) -> head
    uint32_t* chead;
    uint32_t cmask;
    uint32_t* next;              // [n] next DR with the same (group, first key)
    uint32_t* odd;               // [n] DRs that are in no chain; info[3] = how many
    uint8_t* dead;               // [n] out
};

__global__ void __launch_bounds__(128)
k_cl_chain(ReduceArrays r) {
    const uint32_t n = r.a.n();
    const uint32_t t = blockIdx.x * 128 + threadIdx.x;
    if (t >= n) return;
    r.dead[t] = 0;
    const uint32_t slot = r.a.order[t];
    const uint32_t len = r.a.len_of(slot);
    const uint32_t q = r.a.koff[t];
    // a chain key must be the plain integer key of an all-A/C/G/T window: with another letter in it the canonical form of
    // the reverse complement need not be the same (the complement table is not an involution, 'U' -> 'A' -> 'T')
    bool acgt = len >= kClKmer;
    for (uint32_t i = 0; acgt && i < kClKmer; ++i) acgt = cl_code(r.a.rec(slot)[2 + i]) >= 0;
    if (!acgt) { r.odd[atomicAdd(&r.a.info[3], 1u)] = t; return; }
    const unsigned long long k = ((unsigned long long)r.group[t] << 32) | r.a.keys[q];
    uint32_t s = (uint32_t)((k * 0x9E3779B97F4A7C15ull) >> 32) & r.cmask;
    for (;;) {
        const unsigned long long cur = atomicCAS(&r.ckey[s], ~0ull, k);
        if (cur == ~0ull || cur == k) break;
        s = (s + 1) & r.cmask;
    }
    r.next[t] = atomicExch(&r.chead[s], t);
}

__device__ __forceinline__ bool cl_inside(const uint8_t* b, uint32_t at, const uint8_t* a, uint32_t la, bool revcomp) {
    if (!revcomp) { for (uint32_t i = 0; i < la; ++i) if (b[at + i] != a[i]) return false; }
    else { for (uint32_t i = 0; i < la; ++i) if (b[at + i] != c_comp_tab[a[la - 1 - i] & 127]) return false; }
    return true;
}

__global__ void __launch_bounds__(128)
k_cl_reduce(ReduceArrays r) {
    const uint32_t n = r.a.n();
    const uint32_t tb = blockIdx.x * 128 + threadIdx.x;
    if (tb >= n) return;
    const uint32_t sb = r.a.order[tb];
    const uint8_t* b = r.a.rec(sb) + 2;
    const uint32_t lb = r.a.len_of(sb), g = r.group[tb];
    const uint32_t q0 = r.a.koff[tb], nk = r.a.koff[tb + 1] - q0;
    bool dead = false;
    for (uint32_t p = 0; p < nk && !dead; ++p) {
        const uint32_t key = r.a.keys[q0 + p];
        if (key == kClStr) continue;
        const unsigned long long k = ((unsigned long long)g << 32) | key;
        uint32_t s = (uint32_t)((k * 0x9E3779B97F4A7C15ull) >> 32) & r.cmask;
        while (r.ckey[s] != ~0ull && r.ckey[s] != k) s = (s + 1) & r.cmask;
        if (r.ckey[s] == ~0ull) continue;
        for (uint32_t ta = r.chead[s]; ta != 0xFFFFFFFFu && !dead; ta = r.next[ta]) {
            if (ta == tb) continue;
            const uint32_t sa = r.a.order[ta];
            const uint32_t la = r.a.len_of(sa);
            if (!(la < lb || (la == lb && ta < tb))) continue;                 // only earlier DRs count
            const uint8_t* a = r.a.rec(sa) + 2;
            if (p + la <= lb && cl_inside(b, p, a, la, false)) dead = true;
            else if (p + kClKmer >= la && cl_inside(b, p + kClKmer - la, a, la, true)) dead = true;
        }
    }
    const uint32_t n_odd = r.a.info[3];
    for (uint32_t i = 0; i < n_odd && !dead; ++i) {
        const uint32_t ta = r.odd[i];
        if (ta == tb || r.group[ta] != g) continue;
        const uint32_t sa = r.a.order[ta];
        const uint32_t la = r.a.len_of(sa);
        if (!(la < lb || (la == lb && ta < tb))) continue;
        const uint8_t* a = r.a.rec(sa) + 2;
        for (uint32_t at = 0; at + la <= lb && !dead; ++at) dead = cl_inside(b, at, a, la, false) || cl_inside(b, at, a, la, true);
    }
    r.dead[tb] = dead ? 1 : 0;
}

// ---- K4b/K4c in block form: the unit of the multi-GPU exchange -----------------------------------------------
// A token block is 16 header bytes (u32 count, u32 flags, 8 spare) followed by `cap` records of `stride` bytes; the
// last four bytes of a record hold its order key (read index of first appearance).  count may exceed cap: the block
// then holds the first cap records that arrived and the caller retries with a larger one.

struct HitTokens {                                   // source = the token records of a hit list (K4b)
    const crass_b200_hit* hits;
    const uint8_t* tokens;
    uint32_t stride;
    __device__ bool valid(uint32_t) const { return true; }
    __device__ const uint8_t* rec(uint32_t k) const { return tokens + (size_t)k * stride; }
    __device__ uint32_t order(uint32_t k) const { return hits[k].read_index; }
};

struct GatheredBlocks {                              // source = the blocks of all ranks, rank-major (K4c)
    const uint8_t* base;
    uint32_t cap, stride, shard_reads;
    __device__ size_t block_bytes() const { return kTokenBlockHeader + (size_t)cap * stride; }
    __device__ const uint8_t* block(uint32_t r) const { return base + r * block_bytes(); }
    __device__ bool valid(uint32_t k) const { return k % cap < *reinterpret_cast<const uint32_t*>(block(k / cap)); }
    __device__ const uint8_t* rec(uint32_t k) const { return block(k / cap) + kTokenBlockHeader + (size_t)(k % cap) * stride; }
    // shards are contiguous read ranges in rank order, so (rank, first read in shard) is the global first appearance
    __device__ uint32_t order(uint32_t k) const { return (k / cap) * shard_reads + *reinterpret_cast<const uint32_t*>(rec(k) + stride - 4); }
};

template <class Src>
__global__ void __launch_bounds__(256)
k_block_dedupe(Src src, uint32_t n_slots, uint32_t* __restrict__ rep, uint32_t* __restrict__ first_read, uint32_t table_mask) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_slots || !src.valid(k)) return;
    const uint8_t* mine = src.rec(k);
    const uint32_t ord = src.order(k);
    uint32_t i = token_hash(mine) & table_mask;
    for (;;) {
        uint32_t cur = atomicCAS(&rep[i], 0xFFFFFFFFu, k);
        if (cur == 0xFFFFFFFFu) cur = k;
        if (cur == k || token_equal(mine, src.rec(cur))) { atomicMin(&first_read[i], ord); return; }
        i = (i + 1) & table_mask;
    }
}

template <class Src>
__global__ void __launch_bounds__(256)
k_block_compact(Src src, const uint32_t* __restrict__ rep, const uint32_t* __restrict__ first_read, uint32_t table_size,
                uint8_t* __restrict__ out_block, uint32_t out_cap, uint32_t stride) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= table_size) return;
    const uint32_t r = rep[i];
    if (r == 0xFFFFFFFFu) return;
    const uint32_t j = atomicAdd(reinterpret_cast<uint32_t*>(out_block), 1u);
    if (j >= out_cap) return;
    const uint8_t* s = src.rec(r);
    if (s[0] + 6u > stride) atomicOr(reinterpret_cast<uint32_t*>(out_block) + 1, 2u);      // token would run into its order key
    uint8_t* dst = out_block + kTokenBlockHeader + (size_t)j * stride;
    for (uint32_t b = 0; b + 4 < stride; b += 4) *reinterpret_cast<uint32_t*>(dst + b) = *reinterpret_cast<const uint32_t*>(s + b);
    *reinterpret_cast<uint32_t*>(dst + stride - 4) = first_read[i];
}

// flags word of the merged block: bit 0 = some rank's own block had overflowed
__global__ void k_block_flags(GatheredBlocks src, uint32_t n_ranks, uint8_t* __restrict__ out_block) {
    if (threadIdx.x < n_ranks && *reinterpret_cast<const uint32_t*>(src.block(threadIdx.x)) > src.cap)
        atomicOr(reinterpret_cast<uint32_t*>(out_block) + 1, 1u);
}

// ---- K1 generic --------------------------------------------------------------------------------------
// LOCAL_SS > 0: the start/stop list lives in thread-local memory (short reads); otherwise in a slice of
// ss_scratch (ss_cap entries per thread).
template <int LOCAL_SS>
__global__ void __launch_bounds__(128)
k_dr_search_generic(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offsets, uint32_t n_reads, Params o,
                    uint8_t* __restrict__ found, HitSink sink, uint32_t* __restrict__ ss_scratch, uint32_t ss_cap,
                    int* __restrict__ error_flag) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nthreads = gridDim.x * blockDim.x;
    uint32_t local_ss[LOCAL_SS > 0 ? LOCAL_SS : 1];
    uint32_t* ss = LOCAL_SS > 0 ? local_ss : ss_scratch + (size_t)tid * ss_cap;
    const uint32_t cap = LOCAL_SS > 0 ? (uint32_t)LOCAL_SS : ss_cap;
    for (uint32_t r = tid; r < n_reads; r += nthreads) {
        const uint64_t b = offsets[r];
        const uint32_t L = (uint32_t)(offsets[r + 1] - b);
        GmemSeq s{bases + b};
        uint32_t n_ss = 0, replen = 0;
        const int f = cb::search_core(s, L, o, ss, cap, n_ss, replen);
        if (f < 0) *error_flag = f;
        if (found) found[r] = (f == 1);
        if (f == 1) {
            const uint32_t slot = emit_hit(sink, r, ss, n_ss, replen);
            if (sink.tokens && slot != 0xFFFFFFFFu) emit_token(sink.tokens + (size_t)slot * sink.token_stride, sink.token_stride, s, L, ss, n_ss);
        }
    }
}

// ---- K1 fast path, stage 1: 2-bit seed filter ------------------------------------------------------------
// One thread per read, 128 reads per tile.  The tile's bytes are contiguous in the batch: the CTA streams them
// once with coalesced 128-bit loads, recodes 16 bases -> one 32-bit word on the fly and keeps only the packed
// words in shared memory (4x smaller than the bytes).  Each thread then realigns its read out of the packed
// tile with funnel shifts and runs cb::seed_filter entirely in registers.  Reads with a possible seed are
// appended to cand_list; every read gets found[r] = 0 (the exact kernel overwrites the hits with 1).
constexpr int kFilterTile = 128;

__device__ __forceinline__ uint4 ldg_stream128(const uint8_t* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// ---- TMA tile staging (cp.async.bulk + mbarrier) shared by the two filter kernels ----------------------------------
// A tile = TILE consecutive reads = one contiguous byte range of the batch.  One elected thread issues a 1-D bulk copy
// (UBLKCP) of that range into shared memory and arms an mbarrier with the byte count; the copy of tile i+1 is in flight
// while the CTA computes on tile i.  Consumers recode the bytes to 2 bits (16 bases -> one word) cooperatively.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// bytes of tile `tile` that a bulk copy may fetch: [a0, a0 + tma_bytes), 16-byte granular, never past the batch end
template <int TILE, int NW>
__device__ __forceinline__ void tile_extent(const uint64_t* __restrict__ offsets, uint32_t n_reads, uint64_t n_bases, uint32_t tile,
                                            uint64_t& a0, uint32_t& want_bytes, uint32_t& tma_bytes) {
    const uint32_t r0 = tile * TILE;
    const uint32_t r1 = min(r0 + (uint32_t)TILE, n_reads);
    const uint64_t lo = offsets[r0], hi = offsets[r1];
    a0 = lo & ~(uint64_t)15;
    constexpr uint32_t kMax = (uint32_t)(TILE * NW + NW + 8) * 16u;
    uint32_t want = (uint32_t)(((hi - a0 + 15) >> 4) + NW + 4) * 16u;        // + look-ahead for the last read's realignment
    if (want > kMax) want = kMax;
    want_bytes = want;
    const uint64_t n16 = n_bases & ~(uint64_t)15;
    tma_bytes = a0 >= n16 ? 0u : (uint32_t)((n16 - a0) < (uint64_t)want ? (n16 - a0) : (uint64_t)want);
}

template <int TILE, int NW>
__device__ __forceinline__ void tile_issue(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offsets, uint32_t n_reads,
                                           uint64_t n_bases, uint32_t tile, uint8_t* buf, uint64_t* bar) {
    uint64_t a0; uint32_t want, tma_bytes;
    tile_extent<TILE, NW>(offsets, n_reads, n_bases, tile, a0, want, tma_bytes);
    if (tma_bytes) {
        mbar_arrive_expect_tx(bar, tma_bytes);
        tma_load_1d(buf, bases + a0, tma_bytes, bar);
    } else {
        mbar_arrive(bar);
    }
}

// cooperative 2-bit recoding of the staged bytes into `packed` (one word per 16 bases); vectors the bulk copy could not
// cover (ragged end of the batch) are fetched with guarded byte loads
// If `keep` is given, the words also go to the batch-wide 2-bit stream keep[w] = recode(bases[16w, 16w+16)) so that the
// singleton scan of the same batch can read a quarter of the bytes and skip the recoding (k_ac_filter_packed).
template <int TILE>
__device__ __forceinline__ void tile_pack(const uint8_t* __restrict__ bases, uint64_t n_bases, uint64_t a0, uint32_t want_bytes,
                                          uint32_t tma_bytes, const uint8_t* buf, uint32_t* packed, uint32_t* __restrict__ keep = nullptr) {
    const uint32_t nvec = want_bytes >> 4, nt = tma_bytes >> 4;
    if (keep) keep += a0 >> 4;
    for (uint32_t v = threadIdx.x; v < nvec; v += TILE) {
        uint32_t w;
        if (v < nt) {
            const uint4 x = *reinterpret_cast<const uint4*>(buf + 16u * v);
            w = cb::pack16(x.x, x.y, x.z, x.w);
        } else {
            const uint64_t at = a0 + 16ull * v;
            uint32_t q[4] = {0, 0, 0, 0};
            for (int i = 0; i < 16; ++i)
                if (at + i < n_bases) q[i >> 2] |= (uint32_t)__ldg(bases + at + i) << (8 * (i & 3));
            w = cb::pack16(q[0], q[1], q[2], q[3]);
        }
        packed[v] = w;
        if (keep) keep[v] = w;
    }
}

template <int NW>
constexpr size_t dr_filter_smem_bytes() { return (size_t)(kFilterTile * NW + NW + 8) * 20; }   // 16 B/vector bytes + 4 B/vector packed

// Where a filter launch leaves its candidate reads and the exact kernel of the same chunk picks them up.  Two lists
// in one array: reads with several flagged windows (nearly all reads that carry an array) fill [lo, hi) from the front,
// reads with a single flagged window (mostly chance 8-mer repeats) from the back.
struct CandRegion {
    uint32_t* list;
    uint32_t lo, hi;
    uint32_t* counts;          // [0] front entries, [1] back entries, [2] queue head of k_dr_exact_refill
};

template <int NW, int NWIN, int DMIN, int DMAX>
__global__ void __launch_bounds__(kFilterTile)
k_dr_filter(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offsets, uint32_t n_reads,
            uint32_t tile_begin, uint32_t tile_end, uint8_t* __restrict__ found, CandRegion cand,
            uint32_t* __restrict__ keep_packed) {
    constexpr int kWords = kFilterTile * NW + NW + 8;          // packed words a tile can need (+ look-ahead + realignment)
    extern __shared__ __align__(128) uint8_t dyn_smem[];       // dr_filter_smem_bytes<NW>() bytes
    uint8_t* buf = dyn_smem;                                   // the tile's bytes, written by the bulk copy
    uint32_t* sm = reinterpret_cast<uint32_t*>(dyn_smem + kWords * 16);   // the same tile, 2 bits per base
    __shared__ uint64_t full;
    const uint32_t n_tiles = tile_end;                         // this launch covers the tiles [tile_begin, tile_end) of the batch
    const uint64_t n_bases = offsets[n_reads];
    if (threadIdx.x == 0) mbar_init(&full, 1);
    __syncthreads();
    if (threadIdx.x == 0 && tile_begin + blockIdx.x < n_tiles) tile_issue<kFilterTile, NW>(bases, offsets, n_reads, n_bases, tile_begin + blockIdx.x, buf, &full);
    uint32_t parity = 0;
    for (uint32_t tile = tile_begin + blockIdx.x; tile < n_tiles; tile += gridDim.x, parity ^= 1u) {
        const uint32_t r0 = tile * kFilterTile;
        const uint32_t r1 = min(r0 + (uint32_t)kFilterTile, n_reads);
        uint64_t a0; uint32_t want, tma_bytes;
        tile_extent<kFilterTile, NW>(offsets, n_reads, n_bases, tile, a0, want, tma_bytes);
        mbar_wait(&full, parity);                               // the tile's bytes have landed
        tile_pack<kFilterTile>(bases, n_bases, a0, want, tma_bytes, buf, sm, keep_packed);
        __syncthreads();                                        // packed tile complete, byte buffer free again
        if (threadIdx.x == 0 && tile + gridDim.x < n_tiles)
            tile_issue<kFilterTile, NW>(bases, offsets, n_reads, n_bases, tile + gridDim.x, buf, &full);   // overlaps the compute below
        const uint32_t r = r0 + threadIdx.x;
        if (r < r1) {
            const uint32_t b = (uint32_t)(offsets[r] - a0);     // base offset of the read inside the packed tile
            const uint32_t wi = b >> 4, sh = (b & 15u) * 2u;
            uint32_t R[NW + 2];
#pragma unroll
            for (int k = 0; k < NW + 2; ++k) R[k] = cb::funnel_r(sm[wi + k], sm[wi + k + 1], sh);
            uint32_t acc[NWIN];
            cb::seed_flags<NW, NWIN, DMIN, DMAX>(R, acc);
            found[r] = 0;
            if (cb::any_flag<NWIN>(acc)) {
                if (__popc(cb::flag_mask<NWIN>(acc)) >= 2) cand.list[cand.lo + atomicAdd(&cand.counts[0], 1u)] = r;
                else cand.list[cand.hi - 1u - atomicAdd(&cand.counts[1], 1u)] = r;
            }
        }
        __syncthreads();                                        // everyone is done with the packed tile
    }
}

// ---- K1 fast path, stage 1, warp tiles (the default) ------------------------------------------------------------
// The same filter without a staged copy of the bytes: a tile is the 32 consecutive reads of one WARP.  The lanes fetch the
// tile's 16-byte vectors straight from global memory (coalesced, streaming), recode them on the way and keep only the
// 2-bit words in a warp-private slice of shared memory (1.3 KB at 150 bp); then every lane realigns its own read out of
// that slice and evaluates the window flags in registers as before.  No CTA barrier, no byte buffer: 5.4 KB of shared
// memory and 32 registers per 128-thread CTA, so all 64 warps of an SM are resident (the CTA-tile kernel above stages
// 16 + 4 bytes per vector and stops at 8 CTAs), which is what an ALU-bound kernel with ~540 instructions per read needs
// to keep the pipe fed.  Latency is hidden by the other warps instead of by a bulk copy in flight.
constexpr int kFwWarps = 4;

// a batch whose mean read length is above 400 bases is handled by the warp-per-read kernel as a whole
__device__ __forceinline__ bool batch_is_mostly_long(const uint64_t* __restrict__ offsets, uint32_t n_reads) {
    return offsets[n_reads] > 400ull * n_reads;
}

__device__ __noinline__ uint4 ragged_vector(const uint8_t* __restrict__ bases, uint64_t at, uint64_t n_bases) {
    uint32_t q[4] = {0, 0, 0, 0};
    for (int i = 0; i < 16; ++i)
        if (at + i < n_bases) q[i >> 2] |= (uint32_t)__ldg(bases + at + i) << (8 * (i & 3));
    return make_uint4(q[0], q[1], q[2], q[3]);
}

template <int NW, int NWIN, int DMIN, int DMAX>
__global__ void __launch_bounds__(kFwWarps * 32, (NW <= 10 ? 16 : 8))
k_dr_filter_warp(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offsets, uint32_t n_reads,
                 uint32_t r_begin, uint32_t r_end, uint8_t* __restrict__ found, CandRegion cand,
                 uint32_t* __restrict__ keep_packed, uint32_t* __restrict__ long_list = nullptr, uint32_t* __restrict__ long_count = nullptr) {
    constexpr int kWords = 32 * NW + NW + 8;                   // a warp tile's packed words (+ look-ahead + realignment)
    __shared__ uint32_t sm_all[kFwWarps][kWords];
    uint32_t* sm = sm_all[threadIdx.x >> 5];
    const uint32_t lane = threadIdx.x & 31u;
    // one tile per warp, no loop: nothing but the read's own registers is live while the flags are evaluated
    const uint32_t r0 = r_begin + ((blockIdx.x * kFwWarps + (threadIdx.x >> 5)) << 5);
    if (r0 >= r_end) return;
    // MIXED batches (long_list given: some read is longer than the NW words a lane holds).  Reads that long go onto long_list
    // for the warp-per-read kernel; everything else stays here -- one 305 bp read used to send ten million 150 bp reads to
    // the warp-per-read kernel.  A batch of mostly long reads is that kernel's as a whole: this one steps aside.
    if (long_list && batch_is_mostly_long(offsets, n_reads)) return;
    const uint32_t r1 = min(r0 + 32u, r_end);
    const uint32_t r = r0 + lane;
    uint32_t b;                                                 // base offset of the lane's read inside the packed tile
    bool too_long = false;
    {
        const uint64_t my_off = offsets[min(r, r1 - 1u)];       // one coalesced load; the tile's extent comes by shuffle
        const uint64_t a0 = __shfl_sync(0xFFFFFFFFu, my_off, 0) & ~(uint64_t)15;
        const uint64_t last_off = __shfl_sync(0xFFFFFFFFu, my_off, (int)(r1 - 1u - r0));
        b = (uint32_t)(my_off - a0);
        // the last read's realignment touches the words [wi, wi + NW + 2]
        uint32_t nvec = (uint32_t)((last_off - a0) >> 4) + NW + 3;
        uint32_t* keep = keep_packed ? keep_packed + (a0 >> 4) : nullptr;
        const uint64_t n_bases = offsets[n_reads];
        if (long_list) {
            const uint64_t tile_end = offsets[r1];
            uint64_t nxt = __shfl_down_sync(0xFFFFFFFFu, my_off, 1);
            if (r + 1u >= r1) nxt = tile_end;
            too_long = r < r1 && nxt - my_off > (uint64_t)(16 * NW);
            if (nvec > (uint32_t)kWords) {
                // the tile does not fit its words (a long read inside it): its short reads go to the exact kernel unfiltered
                // (it computes the flags of every candidate itself), and the 2-bit stream of the tile's bytes is written here
                if (keep) {
                    const uint64_t nv = (tile_end - a0 + 15) >> 4;
                    for (uint64_t v = lane; v < nv; v += 32) {
                        const uint64_t at = a0 + 16ull * v;
                        const uint4 x = at + 16 <= n_bases ? ldg_stream128(bases + at) : ragged_vector(bases, at, n_bases);
                        keep[v] = cb::pack16(x.x, x.y, x.z, x.w);
                    }
                }
                if (r < r1) {
                    found[r] = 0;
                    if (too_long) long_list[atomicAdd(long_count, 1u)] = r;
                    else cand.list[cand.lo + atomicAdd(&cand.counts[0], 1u)] = r;
                }
                return;
            }
        }
        if (nvec > (uint32_t)kWords) nvec = kWords;
        const uint64_t avail = (n_bases - a0) >> 4;             // whole vectors the batch still holds from a0 on
        const uint32_t n_whole = avail < (uint64_t)nvec ? (uint32_t)avail : nvec;
        const uint8_t* src = bases + a0 + 16u * lane;
#pragma unroll 1
        for (uint32_t v0 = lane; v0 < n_whole; v0 += 128u, src += 2048) {
            uint4 x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)                         // four independent 128-bit loads in flight per lane
                if (v0 + 32u * u < n_whole) x[u] = ldg_stream128(src + 512 * u);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (v0 + 32u * u < n_whole) {
                    const uint32_t w = cb::pack16(x[u].x, x[u].y, x[u].z, x[u].w);
                    sm[v0 + 32u * u] = w;
                    if (keep) keep[v0 + 32u * u] = w;
                }
            }
        }
        for (uint32_t v = n_whole + lane; v < nvec; v += 32u) {     // ragged end of the batch (last tile only)
            const uint4 x = ragged_vector(bases, a0 + 16ull * v, n_bases);
            const uint32_t w = cb::pack16(x.x, x.y, x.z, x.w);
            sm[v] = w;
            if (keep) keep[v] = w;
        }
    }
    __syncwarp();
    if (r >= r1) return;
    if (too_long) { found[r] = 0; long_list[atomicAdd(long_count, 1u)] = r; return; }    // (a long read in a tile that still fits)
    const uint32_t wi = b >> 4, sh = (b & 15u) * 2u;
    uint32_t R[NW + 2];
#pragma unroll
    for (int k = 0; k < NW + 2; ++k) R[k] = cb::funnel_r(sm[wi + k], sm[wi + k + 1], sh);
    uint32_t acc[NWIN];
    cb::seed_flags<NW, NWIN, DMIN, DMAX>(R, acc);
    found[r] = 0;
    if (cb::any_flag<NWIN>(acc)) {
        if (__popc(cb::flag_mask<NWIN>(acc)) >= 2) cand.list[cand.lo + atomicAdd(&cand.counts[0], 1u)] = r;
        else cand.list[cand.hi - 1u - atomicAdd(&cand.counts[1], 1u)] = r;
    }
}

// ---- K1 fast path, stage 2: exact searchCore on the candidate list ---------------------------------------------
// One thread per candidate read.  The thread fetches its read once with 128-bit loads, keeps both the bytes and their
// 2-bit recoding in its own shared-memory slot, recomputes the window flags and then runs cb::search_core_packed: only
// flagged windows are looked at on the bytes, and the flags are recomputed on the re-phased stream whenever a rejected
// candidate moves the window grid.  (Byte reads straight from global memory were 45 % of this kernel's stall samples.)
constexpr int kExactThreads = 128;

struct SmemSeq {                                                 // byte accessor into the thread's shared-memory copy
    const uint8_t* p;
    __device__ __forceinline__ uint8_t operator[](uint32_t i) const { return p[i]; }
};

template <int NW>
constexpr size_t dr_exact_smem_bytes() { return (size_t)kExactThreads * ((((NW + 4) | 1) + (((NW + 3) * 4) | 1)) * 4); }

template <int NW, int NWIN, int DMIN, int DMAX>
__global__ void __launch_bounds__(kExactThreads)
k_dr_exact_packed(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offsets, uint32_t n_reads,
                  CandRegion cand, Params o, uint8_t* __restrict__ found, HitSink sink, int* __restrict__ error_flag) {
    constexpr int kSlot = (NW + 4) | 1;                         // odd strides: conflict-free per-thread slots
    constexpr int kByteSlot = ((NW + 3) * 4) | 1;               // words holding the NW + 3 byte vectors of the read
    extern __shared__ uint32_t sm[];                            // dr_exact_smem_bytes<NW>()
    const uint32_t n_front = cand.counts[0], n_back = cand.counts[1];      // the two lists of the filter kernel
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&sink.counters[3], n_front + n_back);   // the chunks of a batch add up
    const uint64_t n_bases = offsets[n_reads];
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t tasks_front = (n_front + 31u) >> 5, tasks = tasks_front + ((n_back + 31u) >> 5);
    uint32_t* S = sm + threadIdx.x * kSlot;
    uint32_t* B = sm + kExactThreads * kSlot + threadIdx.x * kByteSlot;
    uint32_t ss[32];
    // every warp takes 32 candidates of one list at a time and walks the stages of cb::PackedSearch in lock-step
    for (uint32_t task = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; task < tasks; task += n_warps) {
        const bool front = task < tasks_front;
        const uint32_t i = ((front ? task : task - tasks_front) << 5) + (threadIdx.x & 31u);
        const bool active = i < (front ? n_front : n_back);
        const uint32_t r = active ? cand.list[front ? cand.lo + i : cand.hi - 1u - i] : 0u;
        const uint64_t b = active ? offsets[r] : 0ull;
        const uint32_t L = active ? (uint32_t)(offsets[r + 1] - b) : 0u;
        SmemSeq s{reinterpret_cast<const uint8_t*>(B) + (uint32_t)(b & 15u)};
        uint32_t mask0 = 0;
        if (active) {
            const uint64_t a0 = b & ~(uint64_t)15;
            const uint32_t sh = (uint32_t)(b & 15u) * 2u;
            uint32_t prev = 0;
            uint32_t R[NW + 2];
#pragma unroll
            for (int v = 0; v < NW + 3; ++v) {
                const uint64_t at = a0 + 16ull * v;
                uint4 x = make_uint4(0, 0, 0, 0);
                if (at + 16 <= n_bases) x = __ldg(reinterpret_cast<const uint4*>(bases + at));
                else {
                    uint32_t q[4] = {0, 0, 0, 0};
                    for (int t = 0; t < 16; ++t)
                        if (at + t < n_bases) q[t >> 2] |= (uint32_t)__ldg(bases + at + t) << (8 * (t & 3));
                    x = make_uint4(q[0], q[1], q[2], q[3]);
                }
                B[4 * v] = x.x; B[4 * v + 1] = x.y; B[4 * v + 2] = x.z; B[4 * v + 3] = x.w;
                const uint32_t w = cb::pack16(x.x, x.y, x.z, x.w);
                if (v > 0) R[v - 1] = cb::funnel_r(prev, w, sh);
                prev = w;
            }
#pragma unroll
            for (int k = 0; k < NW + 2; ++k) S[k] = R[k];
            uint32_t acc[NWIN];
            cb::seed_flags<NW, NWIN, DMIN, DMAX>(R, acc);
            mask0 = cb::flag_mask<NWIN>(acc);
        }
        cb::PackedSearch<NW, NWIN, DMIN, DMAX, SmemSeq> st(s, L, o, S, ss, 32u);
        st.init(mask0);
        if (!active) st.done = true;
        __syncwarp();
        for (;;) {
            st.pick();
            if (!__any_sync(0xFFFFFFFFu, st.have)) break;
            st.find();
            __syncwarp();
            st.seed();
            __syncwarp();
            st.reflag();
            __syncwarp();
        }
        const int f = st.result;
        if (f < 0) *error_flag = f;
        if (f == 1) {
            found[r] = 1;
            const uint32_t slot = emit_hit(sink, r, ss, st.n_ss, st.replen);
            if (sink.tokens && slot != 0xFFFFFFFFu) emit_token(sink.tokens + (size_t)slot * sink.token_stride, sink.token_stride, s, L, ss, st.n_ss);
        }
        __syncwarp();
    }
}

// ---- K1 fast path, stage 2 with lane refill (the default) ---------------------------------------------------------
// The same per-lane state machine, but a lane that is through with its candidate takes the next one from a queue instead
// of idling until the slowest lane of its group of 32 is done: the warp loops over the stages (pick / find / seed /
// re-flag) and, whenever at least kRefillMin lanes are free, runs one fetch stage in which they load new candidates.
// A candidate with a rejected array and a second seed further on no longer holds 31 finished lanes hostage, so a few
// warps per SM get through the list -- which leaves the rest of the SM to the filter kernel of the next chunk running
// beside it (capi.cu pipelines the chunks on two streams).
constexpr uint32_t kRefillMin = 8;

template <int NW, int NWIN, int DMIN, int DMAX>
__global__ void __launch_bounds__(kExactThreads)
k_dr_exact_refill(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offsets, uint32_t n_reads,
                  CandRegion cand, Params o, uint8_t* __restrict__ found, HitSink sink, int* __restrict__ error_flag) {
    constexpr int kSlot = (NW + 4) | 1;
    constexpr int kByteSlot = ((NW + 3) * 4) | 1;
    extern __shared__ uint32_t sm[];                            // dr_exact_smem_bytes<NW>()
    const uint32_t n_front = cand.counts[0], n_back = cand.counts[1], total = n_front + n_back;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&sink.counters[3], total);
    const uint64_t n_bases = offsets[n_reads];
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t* S = sm + threadIdx.x * kSlot;
    uint32_t* B = sm + kExactThreads * kSlot + threadIdx.x * kByteSlot;
    uint32_t ss[32];
    cb::PackedSearch<NW, NWIN, DMIN, DMAX, SmemSeq> st(SmemSeq{reinterpret_cast<const uint8_t*>(B)}, 0u, o, S, ss, 32u);
    bool busy = false, exhausted = false;
    uint32_t r = 0;
    for (;;) {
        const uint32_t free_mask = __ballot_sync(0xFFFFFFFFu, !busy);
        if (!exhausted && (free_mask == 0xFFFFFFFFu || __popc(free_mask) >= kRefillMin)) {
            const uint32_t want = __popc(free_mask);
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&cand.counts[2], want);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            exhausted = base + want >= total;
            const uint32_t q = base + __popc(free_mask & ((1u << lane) - 1u));
            if (!busy && q < total) {
                r = q < n_front ? cand.list[cand.lo + q] : cand.list[cand.hi - 1u - (q - n_front)];
                const uint64_t b = offsets[r];
                const uint32_t L = (uint32_t)(offsets[r + 1] - b);
                const uint64_t a0 = b & ~(uint64_t)15;
                const uint32_t sh = (uint32_t)(b & 15u) * 2u;
                uint32_t prev = 0;
                uint32_t R[NW + 2];
#pragma unroll
                for (int v = 0; v < NW + 3; ++v) {
                    const uint64_t at = a0 + 16ull * v;
                    uint4 x;
                    if (at + 16 <= n_bases) x = __ldg(reinterpret_cast<const uint4*>(bases + at));
                    else x = ragged_vector(bases, at, n_bases);
                    B[4 * v] = x.x; B[4 * v + 1] = x.y; B[4 * v + 2] = x.z; B[4 * v + 3] = x.w;
                    const uint32_t w = cb::pack16(x.x, x.y, x.z, x.w);
                    if (v > 0) R[v - 1] = cb::funnel_r(prev, w, sh);
                    prev = w;
                }
#pragma unroll
                for (int k = 0; k < NW + 2; ++k) S[k] = R[k];
                uint32_t acc[NWIN];
                cb::seed_flags<NW, NWIN, DMIN, DMAX>(R, acc);
                st.rebind(SmemSeq{reinterpret_cast<const uint8_t*>(B) + (uint32_t)(b & 15u)}, L);
                st.init(cb::flag_mask<NWIN>(acc));
                busy = true;
            }
            __syncwarp();
        }
        if (!__any_sync(0xFFFFFFFFu, busy)) {
            if (exhausted) break;
            continue;
        }
        st.pick();
        st.find();
        __syncwarp();
        st.seed();
        __syncwarp();
        st.reflag();
        __syncwarp();
        if (busy && st.done) {                                  // this lane's read is settled: report, then the slot is free
            const int f = st.result;
            if (f < 0) *error_flag = f;
            if (f == 1) {
                found[r] = 1;
                const uint32_t slot = emit_hit(sink, r, ss, st.n_ss, st.replen);
                if (sink.tokens && slot != 0xFFFFFFFFu) emit_token(sink.tokens + (size_t)slot * sink.token_stride, sink.token_stride, st.s, st.L, ss, st.n_ss);
            }
            busy = false;
        }
        __syncwarp();
    }
}

// ---- K1 fast path, stage 2, staged (the default) -------------------------------------------------------------------
// ncu on the two kernels above: 11 of 32 lanes active per instruction, 40 % of the instructions in the bit-parallel edit
// distance at 8 lanes -- the lanes of a warp reach the expensive steps at different times.  Here a lane's search is cut
// into stages (fetch a candidate | pick a flagged window + byte-level find | scanRight + extendPreRepeat + the tests of
// qcFoundRepeats that need no distance | one edit distance | re-flag after a rejected array) and the warp runs a stage
// only when at least kStageMin lanes are waiting for it, or when nothing else can move; lanes waiting for another stage
// sit the round out.  The float sums of qcFoundRepeats are fed in the reference's order (dr_core.cuh: qc_start /
// qc_feed), so the verdicts are bit-identical.  The start/stop list lives in shared memory instead of the thread's
// stack: with four CTAs on an SM the stacks of the other kernels did not fit the L1 that the shared memory leaves.
constexpr uint32_t kStageMin = 16;                              // default quorum (CRASS_B200_K1_QUORUM overrides it for measurements)
constexpr int kStagedSs = 33;                                   // odd stride: conflict-free per-thread lists of 32 entries

template <int NW>
constexpr size_t dr_staged_smem_bytes() { return (size_t)kExactThreads * ((((NW + 4) | 1) + (((NW + 3) * 4) | 1) + kStagedSs) * 4); }

template <int NW, int NWIN, int DMIN, int DMAX>
__global__ void __launch_bounds__(kExactThreads)
k_dr_exact_staged(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offsets, uint32_t n_reads,
                  CandRegion cand, Params o, uint8_t* __restrict__ found, HitSink sink, int* __restrict__ error_flag,
                  uint32_t quorum, uint32_t refill_min) {
    constexpr int kSlot = (NW + 4) | 1;
    constexpr int kByteSlot = ((NW + 3) * 4) | 1;
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    enum : uint32_t { FREE = 0, PICK = 1, SEED = 2, OSA = 3, FLAG = 4 };
    extern __shared__ uint32_t sm[];                            // dr_staged_smem_bytes<NW>()
    const uint32_t n_front = cand.counts[0], n_back = cand.counts[1], total = n_front + n_back;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&sink.counters[3], total);
    const uint64_t n_bases = offsets[n_reads];
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t* S = sm + threadIdx.x * kSlot;
    uint32_t* B = sm + kExactThreads * kSlot + threadIdx.x * kByteSlot;
    uint32_t* ss = sm + kExactThreads * (kSlot + kByteSlot) + threadIdx.x * kStagedSs;
    cb::PackedSearch<NW, NWIN, DMIN, DMAX, SmemSeq> st(SmemSeq{reinterpret_cast<const uint8_t*>(B)}, 0u, o, S, ss, 32u);
    uint32_t state = FREE, r = 0;
    bool exhausted = false;
    // where a lane goes after a seed stage (scan/extend/cheap tests, or one more distance)
    auto after_seed = [&]() {
        if (st.osa_pending) { state = OSA; return; }
        if (st.done) {
            const int f = st.result;
            if (f < 0) *error_flag = f;
            if (f == 1) {
                found[r] = 1;
                const uint32_t slot = emit_hit(sink, r, ss, st.n_ss, st.replen);
                if (sink.tokens && slot != 0xFFFFFFFFu) emit_token(sink.tokens + (size_t)slot * sink.token_stride, sink.token_stride, st.s, st.L, ss, st.n_ss);
            }
            state = FREE;
            return;
        }
        state = st.need_flags ? FLAG : PICK;
    };
    for (;;) {
        const uint32_t m_free = __ballot_sync(FULL, state == FREE);
        if (!exhausted && (m_free == FULL || __popc(m_free) >= refill_min)) {
            // candidates are taken from the queue as lanes fall free, never ahead of need: a warp that reserved a block of
            // 64 in advance (tried) sits on work that idle warps could have had, and the kernel got 0.06 ms slower
            const uint32_t take = __popc(m_free);
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&cand.counts[2], take);
            base = __shfl_sync(FULL, base, 0);
            exhausted = base + take >= total;
            const uint32_t rank = __popc(m_free & ((1u << lane) - 1u));
            const uint32_t q = base + rank;
            if (state == FREE && q < total) {
                r = q < n_front ? cand.list[cand.lo + q] : cand.list[cand.hi - 1u - (q - n_front)];
                const uint64_t b = offsets[r];
                const uint32_t L = (uint32_t)(offsets[r + 1] - b);
                const uint64_t a0 = b & ~(uint64_t)15;
                const uint32_t sh = (uint32_t)(b & 15u) * 2u;
                uint32_t prev = 0;
                uint32_t R[NW + 2];
#pragma unroll
                for (int v = 0; v < NW + 3; ++v) {
                    const uint64_t at = a0 + 16ull * v;
                    uint4 x;
                    if (at + 16 <= n_bases) x = __ldg(reinterpret_cast<const uint4*>(bases + at));
                    else x = ragged_vector(bases, at, n_bases);
                    B[4 * v] = x.x; B[4 * v + 1] = x.y; B[4 * v + 2] = x.z; B[4 * v + 3] = x.w;
                    const uint32_t w = cb::pack16(x.x, x.y, x.z, x.w);
                    if (v > 0) R[v - 1] = cb::funnel_r(prev, w, sh);
                    prev = w;
                }
#pragma unroll
                for (int k = 0; k < NW + 2; ++k) S[k] = R[k];
                uint32_t acc[NWIN];
                cb::seed_flags<NW, NWIN, DMIN, DMAX>(R, acc);
                st.rebind(SmemSeq{reinterpret_cast<const uint8_t*>(B) + (uint32_t)(b & 15u)}, L);
                st.init(cb::flag_mask<NWIN>(acc));
                state = PICK;
            }
            __syncwarp();
        }
        // which stages have a quorum; without one, the fullest stage runs alone (the tail of the list, or a lone lane
        // that everybody else is waiting behind)
        const uint32_t c_pick = __popc(__ballot_sync(FULL, state == PICK)), c_seed = __popc(__ballot_sync(FULL, state == SEED)),
                       c_osa = __popc(__ballot_sync(FULL, state == OSA)), c_flag = __popc(__ballot_sync(FULL, state == FLAG));
        if (!(c_pick | c_seed | c_osa | c_flag)) {
            if (exhausted) break;
            continue;
        }
        uint32_t forced = FREE;
        if (c_pick < quorum && c_seed < quorum && c_osa < quorum && c_flag < quorum) {
            forced = PICK; uint32_t best = c_pick;
            if (c_seed > best) { forced = SEED; best = c_seed; }
            if (c_osa > best) { forced = OSA; best = c_osa; }
            if (c_flag > best) { forced = FLAG; best = c_flag; }
        }
        if (forced == PICK || __popc(__ballot_sync(FULL, state == PICK)) >= quorum) {
            if (state == PICK) {
                do { st.pick(); if (!st.have) break; st.find(); } while (st.pos < 0);   // flagged windows until one holds on the bytes
                if (!st.have) state = FREE;                     // no window left: the read has no array
                else state = SEED;
            }
            __syncwarp();
        }
        if (forced == SEED || __popc(__ballot_sync(FULL, state == SEED)) >= quorum) {
            if (state == SEED) { st.seed_start(); after_seed(); }
            __syncwarp();
        }
        if (forced == OSA || __popc(__ballot_sync(FULL, state == OSA)) >= quorum) {
            if (state == OSA) {
                st.seed_resume(cb::edit_distance(st.s, st.q.job_a0, st.q.job_n, st.q.job_b0, st.q.job_m));
                after_seed();
            }
            __syncwarp();
        }
        if (forced == FLAG || __popc(__ballot_sync(FULL, state == FLAG)) >= quorum) {
            if (state == FLAG) { st.reflag(); state = PICK; }
            __syncwarp();
        }
    }
}

// ---- K1 for long reads: one warp per read (dr_long.cuh) -------------------------------------------------------
constexpr int kLongWarps = 4;

__global__ void __launch_bounds__(kLongWarps * 32, 10)
k_dr_long(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offsets, uint32_t n_reads, Params o,
          uint8_t* __restrict__ found, HitSink sink, uint32_t* __restrict__ ss_scratch, uint32_t ss_cap, int* __restrict__ error_flag,
          uint32_t words_per_warp, uint32_t* __restrict__ keep, uint32_t* __restrict__ ticket,
          const uint32_t* __restrict__ long_list = nullptr, const uint32_t* __restrict__ long_count = nullptr) {
    extern __shared__ uint32_t long_smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    uint32_t* S = long_smem + (size_t)warp * words_per_warp;           // the read, 2 bits per base, aligned to its start, zero padded
    const uint32_t gw = blockIdx.x * kLongWarps + warp;
    uint32_t* ss = ss_scratch + (size_t)gw * 2 * ss_cap;
    float* sims = reinterpret_cast<float*>(ss + ss_cap);
    const uint64_t n_bases = offsets[n_reads];
    // reads are handed out one at a time in order (a ticket per warp): lengths of 1-10 kb and the few reads that carry an
    // array make a fixed assignment end with most warps idle
    // mixed batches: only the reads the short-read filter listed (unless the batch is mostly long reads: then all of it)
    const bool listed = long_list && !batch_is_mostly_long(offsets, n_reads);
    const uint32_t n_mine = listed ? *long_count : n_reads;
    for (;;) {
        uint32_t r = 0;
        if (lane == 0) r = atomicAdd(ticket, 1u);
        r = __shfl_sync(cbl::kFull, r, 0);
        if (r >= n_mine) break;
        if (listed) r = long_list[r];
        const uint64_t b = offsets[r];
        const uint32_t L = (uint32_t)(offsets[r + 1] - b);
        const int se = cb::search_end(o, L);
        if (se < 0 && !keep) { if (lane == 0 && found) found[r] = 0; continue; }       // too short for an array (but phase 2 may want its words)
        cbl::GSeq s{bases + b};
        const uint64_t a0 = b & ~(uint64_t)15;
        const uint32_t shb = (uint32_t)(b & 15u);
        const uint32_t nvec = (shb + L + 15) >> 4;                      // 16-byte vectors of the batch grid that hold the read
        const uint32_t nW = (L + 15) >> 4;                              // words of the read-aligned stream
        // word v of the grid = recode of bases[a0 + 16v, +16) (0 past the read's last vector); stream word k is cut out of grid
        // words k and k+1, the second of which is the next lane's: four vectors per lane in flight, neighbours by shuffle
        uint32_t carry = 0;                                             // grid word v0 - 1 of the round before
        for (uint32_t v0 = 0; v0 <= nW; v0 += 128) {
            uint4 x[4];
            uint32_t w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t v = v0 + 32u * u + lane;
                const uint64_t at = a0 + 16ull * v;
                x[u] = make_uint4(0, 0, 0, 0);
                if (v < nvec && at + 16 <= n_bases) x[u] = ldg_stream128(bases + at);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t v = v0 + 32u * u + lane;
                const uint64_t at = a0 + 16ull * v;
                if (v < nvec && at + 16 > n_bases) x[u] = ragged_vector(bases, at, n_bases);      // the batch's last bytes
                w[u] = v < nvec ? cb::pack16(x[u].x, x[u].y, x[u].z, x[u].w) : 0u;
                if (keep && v < nvec) keep[(a0 >> 4) + v] = w[u];      // the batch-wide 2-bit stream for the singleton scan
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t v = v0 + 32u * u + lane;
                uint32_t prev = __shfl_up_sync(cbl::kFull, w[u], 1);
                const uint32_t edge = __shfl_sync(cbl::kFull, u ? w[u ? u - 1 : 0] : carry, 31);
                if (lane == 0) prev = u ? edge : carry;
                if (v >= 1 && v <= nW) S[v - 1] = cb::funnel_r(prev, w[u], 2 * shb);
            }
            carry = __shfl_sync(cbl::kFull, w[3], 31);
        }
        for (uint32_t k = nW + lane; k < nW + 24; k += 32) S[k] = 0u;
        __syncwarp();
        if (se < 0) { if (lane == 0 && found) found[r] = 0; continue; }
        uint32_t base = 0, n_ss = 0, replen = 0;
        int result = 0;
        for (;;) {                                                      // one iteration per phase of the window grid
            const uint32_t q = base >> 4, sh = 2 * (base & 15u);
            const uint32_t nwin = ((uint32_t)se - base) / 8 + 1;
            const uint32_t nseg = (nwin + 11) / 12;
            bool restart = false, finished = false;
            for (uint32_t seg0 = 0; seg0 < nseg && !restart && !finished; seg0 += 32) {
                const uint32_t seg = seg0 + lane;
                uint32_t mask = 0;
                if (seg < nseg) {
                    uint32_t Q[15];
                    const uint32_t i0 = q + 6 * seg;
#pragma unroll
                    for (int i = 0; i < 15; ++i) Q[i] = cb::funnel_r(S[i0 + i], S[i0 + i + 1], sh);
                    uint32_t acc[6];
                    cb::seed_flags<13, 6, 49, 97>(Q, acc);
                    mask = cb::flag_mask<6>(acc);
                    const uint32_t nvalid = nwin - 12 * seg;
                    if (nvalid < 12) mask &= (1u << nvalid) - 1u;
                }
                uint32_t bal = __ballot_sync(cbl::kFull, mask != 0);
                while (bal && !restart && !finished) {
                    const int f = __ffs((int)bal) - 1;
                    const uint32_t m = __shfl_sync(cbl::kFull, mask, f);
                    const int h = __ffs((int)m) - 1;
                    if ((int)lane == f) mask &= mask - 1;
                    const uint32_t j = base + 8u * (12u * (seg0 + (uint32_t)f) + (uint32_t)h);
                    uint32_t begin, end;
                    cb::window_text(o, L, j, begin, end);
                    const int pos = cbl::warp_find_left8(s, begin, end, j);
                    if (pos >= 0) {
                        bool advance = false; uint32_t nj = 0;
                        const int rr = cbl::warp_process_seed(s, L, o, j, begin + (uint32_t)pos, ss, n_ss, ss_cap, replen, advance, nj, sims);
                        if (rr != 0) { result = rr; finished = true; }
                        else if (advance) {
                            base = nj + 8u;
                            if (base > (uint32_t)se) finished = true; else restart = true;
                        }
                    }
                    bal = __ballot_sync(cbl::kFull, mask != 0);
                }
            }
            if (!restart) break;
        }
        if (lane == 0) {
            if (found) found[r] = result == 1;
            if (result < 0) *error_flag = result;
            if (result == 1) {
                const uint32_t slot = emit_hit(sink, r, ss, n_ss, replen);
                if (sink.tokens && slot != 0xFFFFFFFFu) emit_token(sink.tokens + (size_t)slot * sink.token_stride, sink.token_stride, s, L, ss, n_ss);
            }
        }
        __syncwarp();
    }
}

// ---- K2 generic --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_ac_scan_generic(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offsets, uint32_t n_reads,
                  const uint32_t* __restrict__ table, uint32_t stride_log2, const uint8_t* __restrict__ symv_g,
                  const uint8_t* __restrict__ skip, uint8_t* __restrict__ found, HitSink sink) {
    __shared__ uint8_t symv[256];
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) symv[i] = symv_g[i];
    __syncthreads();
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nthreads = gridDim.x * blockDim.x;
    for (uint32_t r = tid; r < n_reads; r += nthreads) {
        if (skip && skip[r]) { if (found) found[r] = 0; continue; }
        const uint64_t b = offsets[r];
        const uint32_t L = (uint32_t)(offsets[r + 1] - b);
        const uint8_t* s = bases + b;
        uint32_t st = 0, end = 0, plen = 0;
        for (uint32_t i = 0; i < L; ++i) {
            const uint32_t sy = symv[__ldg(s + i)];
            if (!sy) { st = 0; continue; }
            const uint32_t e = __ldg(table + ((size_t)st << stride_log2) + (sy - 1));
            st = e & 0xFFFFFFu;
            if (e >> 24) { end = i + 1; plen = e >> 24; break; }
        }
        if (found) found[r] = plen != 0;
        if (plen) {
            // on_match (libcrispr.cpp:420-437): DR_end = textpos-1 clamped to L-1; start = DR_end-(len-1)
            uint32_t dr_end = end - 1;
            if (dr_end >= L) dr_end = L - 1;
            uint32_t ss[2] = { dr_end - (plen - 1), dr_end };
            emit_hit(sink, r, ss, 2, 0);
        }
    }
}

// ---- K2 fast path, stage 1: 16-mer q-gram filter -----------------------------------------------------------
// Same tile staging as k_dr_filter (coalesced 128-bit loads, 2-bit recoding into shared memory, one thread per
// read).  Every pattern is >= 23 bytes, so an occurrence always contains the read-aligned 16-mer at a multiple of 8
// (see host/ac_build.cpp); the thread looks those 16-mers up in a shared-memory bitmap and, on a bitmap hit, in
// the exact key table in global memory.  Reads with a confirmed pattern 16-mer go to the automaton kernel.
struct QgramFilter {
    const uint32_t* bitmap;     // 2^bits bits
    const uint32_t* keys;       // 2^table_bits slots, 0xFFFFFFFF = empty
    uint32_t bits, table_bits;
    uint32_t hashes;            // 1 or 2 hash functions behind the bitmap (two while the key set is small enough to keep it sparse)
    const uint32_t* ones;       // [0] != 0: some pattern holds the all-ones 16-mer (which cannot be a table key), [1] head of the chain of patterns that begin with it
};

constexpr int kAcTile = 1024;     // big CTAs: the 64-128 KB bitmap is staged once per CTA, so threads/CTA sets the occupancy

__device__ __forceinline__ bool qgram_member(const QgramFilter& q, uint32_t code) {
    if (code == 0xFFFFFFFFu) return __ldg(q.ones) != 0;
    const uint32_t mask = (1u << q.table_bits) - 1u;
    uint32_t slot = (code * 0x85EBCA6Bu) >> (32 - q.table_bits);
    for (;;) {
        const uint32_t k = __ldg(q.keys + slot);
        if (k == code) return true;
        if (k == 0xFFFFFFFFu) return false;
        slot = (slot + 1) & mask;
    }
}

// The bitmap is a Bloom filter with TWO hash functions; the second bit is only looked at where the first one is set (a few
// per cent of the look-ups), so it costs next to nothing and squares the share of 16-mers that go on to the key table in L2:
// at 4 000 patterns (32 k keys in 2^19 bits) from one probe per read to a quarter of one.  Above 100 k keys the second
// bit fills the bitmap faster than it filters (20 000 patterns: 1.89 -> 1.97 ms per 20 M reads), so the matcher then keeps one.
constexpr uint32_t kQgramHash2 = 0xC2B2AE35u;
__device__ __forceinline__ bool qgram_second_bit(const uint32_t* bm, uint32_t code, uint32_t hshift) {
    const uint32_t h = (code * kQgramHash2) >> hshift;
    return ((bm[h >> 5] >> (h & 31u)) & 1u) != 0;
}

template <int NW>
__global__ void __launch_bounds__(kAcTile)
k_ac_filter(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offsets, uint32_t n_reads, QgramFilter q,
            const uint8_t* __restrict__ skip, uint8_t* __restrict__ found, uint32_t* __restrict__ cand_list,
            uint64_t* __restrict__ cand_mask, uint32_t* __restrict__ counters) {
    constexpr int kWords = kAcTile * NW + NW + 8;
    extern __shared__ uint32_t dsm[];
    uint32_t* bm = dsm;                                          // bitmap, (1 << bits) / 32 words
    uint32_t* sm = dsm + (1u << (q.bits - 5));                   // packed tile
    for (uint32_t i = threadIdx.x; i < (1u << (q.bits - 5)); i += kAcTile) bm[i] = __ldg(q.bitmap + i);
    const uint32_t n_tiles = (n_reads + kAcTile - 1) / kAcTile;
    const uint64_t n_bases = offsets[n_reads];
    const uint32_t hshift = 32 - q.bits;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t r0 = tile * kAcTile;
        const uint32_t r1 = min(r0 + (uint32_t)kAcTile, n_reads);
        const uint64_t lo = offsets[r0], hi = offsets[r1];
        const uint64_t a0 = lo & ~(uint64_t)15;
        uint32_t nvec = (uint32_t)((hi - a0 + 15) >> 4) + NW + 4;
        if (nvec > (uint32_t)kWords) nvec = kWords;
        __syncthreads();
        for (uint32_t v = threadIdx.x; v < nvec; v += kAcTile) {
            const uint64_t at = a0 + 16ull * v;
            uint32_t w;
            if (at + 16 <= n_bases) {
                const uint4 x = ldg_stream128(bases + at);
                w = cb::pack16(x.x, x.y, x.z, x.w);
            } else {
                uint32_t qq[4] = {0, 0, 0, 0};
                for (int i = 0; i < 16; ++i)
                    if (at + i < n_bases) qq[i >> 2] |= (uint32_t)__ldg(bases + at + i) << (8 * (i & 3));
                w = cb::pack16(qq[0], qq[1], qq[2], qq[3]);
            }
            sm[v] = w;
        }
        __syncthreads();
        const uint32_t r = r0 + threadIdx.x;
        if (r < r1) {
            const uint64_t rb = offsets[r];
            const uint32_t L = (uint32_t)(offsets[r + 1] - rb);
            const uint32_t b = (uint32_t)(rb - a0);
            const uint32_t wi = b >> 4, sh = (b & 15u) * 2u;
            uint32_t R[NW + 1];
#pragma unroll
            for (int k = 0; k < NW + 1; ++k) R[k] = cb::funnel_r(sm[wi + k], sm[wi + k + 1], sh);
            // 16-mer i starts at base 8i: even i is word i/2, odd i straddles two words
            uint64_t hits = 0;
#pragma unroll
            for (int i = 0; i < 2 * NW - 1; ++i) {
                const uint32_t code = (i & 1) ? cb::funnel_r(R[i >> 1], R[(i >> 1) + 1], 16) : R[i >> 1];
                const uint32_t h = (code * 0x9E3779B1u) >> hshift;
                const uint32_t bit = (bm[h >> 5] >> (h & 31u)) & 1u;
                hits |= (uint64_t)bit << i;
            }
            const int n_kmers = L >= 16 ? (int)((L - 16) >> 3) + 1 : 0;        // 16-mers that lie inside the read
            if (n_kmers < 64) hits &= (1ull << n_kmers) - 1ull;
            bool cand = false;
            uint64_t mask = 0;                                   // the confirmed 16-mer + the bitmap hits behind it (see k_ac_verify_mask)
            while (hits && !cand) {
                const int i = __ffsll((long long)hits) - 1;
                hits &= hits - 1;
                const uint32_t w0 = cb::funnel_r(sm[wi + (i >> 1)], sm[wi + (i >> 1) + 1], sh);
                const uint32_t w1 = cb::funnel_r(sm[wi + (i >> 1) + 1], sm[wi + (i >> 1) + 2], sh);
                const uint32_t code = (i & 1) ? cb::funnel_r(w0, w1, 16) : w0;
                if (q.hashes > 1 && !qgram_second_bit(bm, code, hshift)) continue;
                cand = qgram_member(q, code);
                if (cand) mask = hits | (1ull << i);
            }
            found[r] = 0;
            if (cand && !(skip && skip[r])) {
                const uint32_t slot = atomicAdd(&counters[3], 1u);
                cand_list[slot] = r;
                cand_mask[slot] = mask;
            }
        }
    }
}

// ---- K2 fast path, stage 1, on the 2-bit stream phase 1 left behind ----------------------------------------------------
// When the direct-repeat search of the same batch ran through k_dr_filter, every base of the batch already exists as
// 2 bits in HBM (tile_pack's `keep` stream).  This form of the filter never touches the bytes: one elected thread
// moves a tile's packed words into shared memory with a 1-D bulk copy (a quarter of the bytes, no recoding
// instructions), double-buffered so that the copy of tile i+1 runs under the look-ups of tile i.  The look-ups are
// those of k_ac_filter, so the candidate set is identical.
constexpr int kAcPackedTile = 512;                   // reads per tile = threads per CTA

template <int NW>
__host__ __device__ constexpr int ac_packed_words() { return (kAcPackedTile * NW + NW + 16 + 3) & ~3; }   // per buffer, 16-byte granular

template <int NW>
__device__ __forceinline__ void ac_packed_issue(const uint32_t* __restrict__ packed, const uint64_t* __restrict__ offsets, uint32_t n_reads,
                                                uint32_t tile, uint32_t* buf, uint64_t* bar) {
    const uint32_t r0 = tile * kAcPackedTile;
    const uint32_t r1 = min(r0 + (uint32_t)kAcPackedTile, n_reads);
    const uint64_t a0 = offsets[r0] & ~(uint64_t)63;                                    // word index a0/16 is a multiple of 4
    uint32_t nw = ((uint32_t)((offsets[r1] - a0 + 15) >> 4) + NW + 4 + 3) & ~3u;
    if (nw > (uint32_t)ac_packed_words<NW>()) nw = ac_packed_words<NW>();
    mbar_arrive_expect_tx(bar, nw * 4u);
    tma_load_1d(buf, packed + (a0 >> 4), nw * 4u, bar);
}

template <int NW>
__global__ void __launch_bounds__(kAcPackedTile)
k_ac_filter_packed(const uint32_t* __restrict__ packed, const uint64_t* __restrict__ offsets, uint32_t n_reads, QgramFilter q,
                   const uint8_t* __restrict__ skip, uint8_t* __restrict__ found, uint32_t* __restrict__ cand_list, uint64_t* __restrict__ cand_mask,
                   uint32_t* __restrict__ counters) {
    constexpr int kWords = ac_packed_words<NW>();
    extern __shared__ __align__(128) uint32_t dsm[];
    uint32_t* tiles = dsm;                                       // two packed tiles (bulk-copy targets, 16-byte aligned)
    uint32_t* bm = dsm + 2 * kWords;                             // bitmap, (1 << bits) / 32 words
    __shared__ uint64_t full[2];
    for (uint32_t i = threadIdx.x; i < (1u << (q.bits - 5)); i += kAcPackedTile) bm[i] = __ldg(q.bitmap + i);
    if (threadIdx.x == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); }
    __syncthreads();
    const uint32_t n_tiles = (n_reads + kAcPackedTile - 1) / kAcPackedTile;
    const uint32_t hshift = 32 - q.bits;
    if (threadIdx.x == 0 && blockIdx.x < n_tiles) ac_packed_issue<NW>(packed, offsets, n_reads, blockIdx.x, tiles, &full[0]);
    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t cur = it & 1u;
        // the other buffer was last read in the previous iteration, which ended with a CTA barrier
        if (threadIdx.x == 0 && tile + gridDim.x < n_tiles)
            ac_packed_issue<NW>(packed, offsets, n_reads, tile + gridDim.x, tiles + (cur ^ 1u) * kWords, &full[cur ^ 1u]);
        const uint32_t* sm = tiles + cur * kWords;
        const uint32_t r0 = tile * kAcPackedTile;
        const uint32_t r1 = min(r0 + (uint32_t)kAcPackedTile, n_reads);
        const uint64_t a0 = offsets[r0] & ~(uint64_t)63;
        const uint32_t r = r0 + threadIdx.x;
        uint64_t rb = 0;
        uint32_t L = 0;
        if (r < r1) { rb = offsets[r]; L = (uint32_t)(offsets[r + 1] - rb); }
        mbar_wait(&full[cur], (it >> 1) & 1u);                   // this tile's words have landed
        if (r < r1) {
            const uint32_t b = (uint32_t)(rb - a0);
            const uint32_t wi = b >> 4, sh = (b & 15u) * 2u;
            uint32_t R[NW + 1];
#pragma unroll
            for (int k = 0; k < NW + 1; ++k) R[k] = cb::funnel_r(sm[wi + k], sm[wi + k + 1], sh);
            uint64_t hits = 0;                                   // 16-mer i starts at base 8i: even i is word i/2, odd i straddles
#pragma unroll
            for (int k = 0; k < 2 * NW - 1; ++k) {
                const uint32_t code = (k & 1) ? cb::funnel_r(R[k >> 1], R[(k >> 1) + 1], 16) : R[k >> 1];
                const uint32_t h = (code * 0x9E3779B1u) >> hshift;
                hits |= (uint64_t)((bm[h >> 5] >> (h & 31u)) & 1u) << k;
            }
            const int n_kmers = L >= 16 ? (int)((L - 16) >> 3) + 1 : 0;        // 16-mers that lie inside the read
            if (n_kmers < 64) hits &= (1ull << n_kmers) - 1ull;
            bool cand = false;
            uint64_t mask = 0;
            while (hits && !cand) {
                const int k = __ffsll((long long)hits) - 1;
                hits &= hits - 1;
                const uint32_t w0 = cb::funnel_r(sm[wi + (k >> 1)], sm[wi + (k >> 1) + 1], sh);
                const uint32_t w1 = cb::funnel_r(sm[wi + (k >> 1) + 1], sm[wi + (k >> 1) + 2], sh);
                const uint32_t code = (k & 1) ? cb::funnel_r(w0, w1, 16) : w0;
                if (q.hashes > 1 && !qgram_second_bit(bm, code, hshift)) continue;
                cand = qgram_member(q, code);
                if (cand) mask = hits | (1ull << k);
            }
            found[r] = 0;
            if (cand && !(skip && skip[r])) {
                const uint32_t slot = atomicAdd(&counters[3], 1u);
                cand_list[slot] = r;
                cand_mask[slot] = mask;
            }
        }
        __syncthreads();                                         // everyone is done with this buffer
    }
}

// ---- K2 fast path for long reads: the same 16-mer filter, one warp per read -------------------------------------------
// Lanes take consecutive 16-byte vectors of the read (coalesced), recode them to one word each, fetch the neighbour's
// word with a shuffle and test the two read-aligned 16-mers that start inside their vector (offsets == read start mod 8).
constexpr int kAcLongThreads = 256;

template <bool PACKED>                               // PACKED: `bases` is the 2-bit stream of the batch (u32 words) instead of the bytes
__global__ void __launch_bounds__(kAcLongThreads)
k_ac_filter_long(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offsets, uint32_t n_reads, QgramFilter q,
                 const uint8_t* __restrict__ skip, uint8_t* __restrict__ found, uint32_t* __restrict__ cand_list,
                 uint32_t* __restrict__ cand_from, uint32_t* __restrict__ counters) {
    extern __shared__ uint32_t dsm[];
    uint32_t* bm = dsm;
    for (uint32_t i = threadIdx.x; i < (1u << (q.bits - 5)); i += kAcLongThreads) bm[i] = __ldg(q.bitmap + i);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gw = (blockIdx.x * kAcLongThreads + threadIdx.x) >> 5, nw = (gridDim.x * kAcLongThreads) >> 5;
    const uint64_t n_bases = offsets[n_reads];
    const uint32_t hshift = 32 - q.bits;
    const uint64_t n_words = (n_bases + 15) >> 4;                  // words of the batch-wide stream
    for (uint32_t r = gw; r < n_reads; r += nw) {
        if (lane == 0) found[r] = 0;
        if (skip && skip[r]) continue;
        const uint64_t b = offsets[r];
        const uint32_t L = (uint32_t)(offsets[r + 1] - b);
        if (L < 16) continue;
        // A lane takes FOUR consecutive words of the stream (one 16-byte load of the 2-bit stream, or four of the bytes) and the
        // first word of the next lane: eight read-aligned 16-mers per lane and round, 1 984 bases per warp and round -- with
        // one word per lane a 5 kb read was eleven dependent trips to DRAM.  Groups start at a multiple of four words, so up to
        // three words in front of the read ride along (their 16-mers fail the range test below).
        const uint64_t w_read = b >> 4;                             // word that holds the read's first base
        const uint64_t w_first = w_read & ~(uint64_t)3;
        const int32_t x0 = (int32_t)((int64_t)(w_first << 4) - (int64_t)b);   // first base of word w_first, relative to the read start (<= 0)
        const int32_t d = (int32_t)((uint32_t)(b & 7u));            // read-aligned 16-mers start at stream offsets == b mod 8
        const uint32_t n_groups = (uint32_t)((((b + L + 15) >> 4) - w_first + 3) >> 2);
        bool cand = false;
        uint32_t g_hit = 0;
        for (uint32_t g0 = 0; g0 < n_groups && !cand; g0 += 31) {
            const uint32_t g = g0 + lane;
            uint32_t w[5] = {0, 0, 0, 0, 0};
            if (g < n_groups + 1) {                                 // (the group past the last one only lends its first word)
                const uint64_t wi = w_first + 4ull * g;
                if (PACKED) {
                    if (wi + 4 <= n_words) {
                        const uint4 x = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(bases) + wi));
                        w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w;
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) if (wi + k < n_words) w[k] = __ldg(reinterpret_cast<const uint32_t*>(bases) + wi + k);
                    }
                } else {
                    uint4 x[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t at = (wi + k) << 4;
                        x[k] = make_uint4(0, 0, 0, 0);
                        if (at + 16 <= n_bases) x[k] = ldg_stream128(bases + at);
                        else if (at < n_bases) x[k] = ragged_vector(bases, at, n_bases);
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) w[k] = cb::pack16(x[k].x, x[k].y, x[k].z, x[k].w);
                }
            }
            w[4] = __shfl_down_sync(0xFFFFFFFFu, w[0], 1);
            bool hit = false;
            if (lane < 31 && g < n_groups) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const int32_t x = x0 + 16 * (int32_t)(4u * g + (uint32_t)k) + d + 8 * t;       // start of the 16-mer, relative to the read start
                        if (x >= 0 && x + 16 <= (int32_t)L) {
                            const uint32_t code = cb::funnel_r(w[k], w[k + 1], 2u * (uint32_t)(d + 8 * t));
                            const uint32_t h = (code * 0x9E3779B1u) >> hshift;
                            if (((bm[h >> 5] >> (h & 31u)) & 1u) && (q.hashes < 2 || qgram_second_bit(bm, code, hshift))) hit = hit || qgram_member(q, code);
                        }
                    }
                }
            }
            cand = __any_sync(0xFFFFFFFFu, hit);
            g_hit = g0;
        }
        if (cand && lane == 0) {
            const uint32_t at = atomicAdd(&counters[3], 1u);
            cand_list[at] = r;
            // No aligned 16-mer before this round is a key, and an occurrence that starts at a holds the aligned 16-mer at
            // a ... a + 7: the verification may start seven bases before the round's first 16-mer.
            const int32_t first_x = x0 + 64 * (int32_t)g_hit + d;
            cand_from[at] = first_x > 7 ? (uint32_t)(first_x - 7) : 0u;
        }
    }
}

// ---- K2 fast path, stage 2: exact verification of the candidate reads ---------------------------------------------
// Slides over the read, looks the 16-mer starting at every position up in the pattern-start table (first 16-mer of each
// pattern -> chain of patterns) and compares the chained patterns byte by byte.  Over all verified occurrences it keeps
// the smallest end offset and, on ties, the longest pattern: exactly the first callback acism delivers
// (acism.c:26-104 with on_match returning 1).  A byte that occurs in no pattern can never be inside an occurrence,
// which is all the "reset to ROOT" rule of acism.c:36-42 means for the first match.
struct PatternStarts {
    const uint8_t* p_bytes;
    const uint32_t* p_offs;
    const uint32_t* s_keys;
    const uint32_t* s_head;
    const uint32_t* p_next;
    uint32_t s_bits;
    const uint32_t* ones;       // see QgramFilter
    uint32_t min_len;
};

__global__ void __launch_bounds__(128)
k_ac_verify_list(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offsets, const uint32_t* __restrict__ cand_list,
                 PatternStarts ps, uint8_t* __restrict__ found, HitSink sink) {
    const uint32_t n_cand = sink.counters[3];
    const uint32_t nthreads = gridDim.x * blockDim.x;
    const uint32_t smask = (1u << ps.s_bits) - 1u;
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cand; c += nthreads) {
        const uint32_t r = cand_list[c];
        const uint64_t b = offsets[r];
        const uint32_t L = (uint32_t)(offsets[r + 1] - b);
        const uint8_t* s = bases + b;
        uint32_t best_end = 0xFFFFFFFFu, best_len = 0, code = 0;
        for (uint32_t i = 0; i < L; ++i) {
            code = (code >> 2) | ((uint32_t)((__ldg(s + i) >> 1) & 3u) << 30);
            if (i < 15) continue;
            const uint32_t p = i - 15;                               // start of this 16-mer
            if (best_len && p + ps.min_len > best_end) break;        // later starts cannot end earlier
            uint32_t pi;
            if (code == 0xFFFFFFFFu) pi = __ldg(ps.ones + 1);
            else {
                uint32_t slot = (code * 0x85EBCA6Bu) >> (32 - ps.s_bits);
                for (;;) {
                    const uint32_t k = __ldg(ps.s_keys + slot);
                    if (k == code) { pi = __ldg(ps.s_head + slot); break; }
                    if (k == 0xFFFFFFFFu) { pi = 0xFFFFFFFFu; break; }
                    slot = (slot + 1) & smask;
                }
            }
            for (; pi != 0xFFFFFFFFu; pi = __ldg(ps.p_next + pi)) {
                const uint32_t po = __ldg(ps.p_offs + pi), len = __ldg(ps.p_offs + pi + 1) - po;
                if (p + len > L) continue;
                const uint32_t end = p + len;
                if (end > best_end || (end == best_end && len <= best_len)) continue;
                bool same = true;
                for (uint32_t k = 0; k < len; ++k)
                    if (__ldg(s + p + k) != __ldg(ps.p_bytes + po + k)) { same = false; break; }
                if (same) { best_end = end; best_len = len; }
            }
        }
        if (best_len) {
            uint32_t dr_end = best_end - 1;                          // on_match (libcrispr.cpp:420-437)
            if (dr_end >= L) dr_end = L - 1;
            uint32_t ss[2] = { dr_end - (best_len - 1), dr_end };
            found[r] = 1;
            emit_hit(sink, r, ss, 2, 0);
        }
    }
}

// Same verification with one WARP per candidate (long reads: a single thread walking 10 kb of dependent table probes
// is latency-bound).  Lanes take 32 consecutive start positions per round; every lane keeps its own best occurrence,
// the warp minimum of (end, -len) is the answer.  The scan stops once no later start can end earlier.
__global__ void __launch_bounds__(128)
k_ac_verify_warp(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offsets, const uint32_t* __restrict__ cand_list,
                 const uint32_t* __restrict__ cand_from, PatternStarts ps, uint8_t* __restrict__ found, HitSink sink) {
    const uint32_t n_cand = sink.counters[3];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const uint32_t smask = (1u << ps.s_bits) - 1u;
    for (uint32_t c = gw; c < n_cand; c += nw) {
        const uint32_t r = cand_list[c];
        const uint64_t b = offsets[r];
        const uint32_t L = (uint32_t)(offsets[r + 1] - b);
        const uint8_t* s = bases + b;
        uint32_t best = 0xFFFFFFFFu;                                  // (end << 8) | (255 - len): min == earliest end, longest pattern
        for (uint32_t p0 = cand_from ? cand_from[c] : 0u; p0 + 16 <= L; p0 += 32) {      // (the long-read filter knows where nothing can start)
            const uint32_t warp_best = __reduce_min_sync(0xFFFFFFFFu, best);
            if (warp_best != 0xFFFFFFFFu && p0 + ps.min_len > (warp_best >> 8)) break;
            const uint32_t p = p0 + lane;
            if (p + 16 <= L) {
                uint32_t code = 0;
#pragma unroll
                for (int k = 0; k < 16; ++k) code |= (uint32_t)((__ldg(s + p + k) >> 1) & 3u) << (2 * k);
                uint32_t pi;
                if (code == 0xFFFFFFFFu) pi = __ldg(ps.ones + 1);
                else {
                    uint32_t slot = (code * 0x85EBCA6Bu) >> (32 - ps.s_bits);
                    for (;;) {
                        const uint32_t k = __ldg(ps.s_keys + slot);
                        if (k == code) { pi = __ldg(ps.s_head + slot); break; }
                        if (k == 0xFFFFFFFFu) { pi = 0xFFFFFFFFu; break; }
                        slot = (slot + 1) & smask;
                    }
                }
                for (; pi != 0xFFFFFFFFu; pi = __ldg(ps.p_next + pi)) {
                    const uint32_t po = __ldg(ps.p_offs + pi), len = __ldg(ps.p_offs + pi + 1) - po;
                    if (p + len > L) continue;
                    const uint32_t key = ((p + len) << 8) | (255u - len);
                    if (key >= best) continue;
                    // a chained pattern shares the lane's 16-mer code, so it nearly always matches: no early exit, the byte
                    // loads of all positions are independent and overlap instead of forming a chain of 2 x len latencies
                    uint32_t diff = 0;
#pragma unroll 8
                    for (uint32_t k = 0; k < len; ++k) diff |= (uint32_t)__ldg(s + p + k) ^ (uint32_t)__ldg(ps.p_bytes + po + k);
                    if (diff == 0) best = key;
                }
            }
        }
        best = __reduce_min_sync(0xFFFFFFFFu, best);
        if (best != 0xFFFFFFFFu && lane == 0) {
            const uint32_t best_end = best >> 8, best_len = 255u - (best & 255u);
            uint32_t dr_end = best_end - 1;                          // on_match (libcrispr.cpp:420-437)
            if (dr_end >= L) dr_end = L - 1;
            uint32_t ss[2] = { dr_end - (best_len - 1), dr_end };
            found[r] = 1;
            emit_hit(sink, r, ss, 2, 0);
        }
    }
}

// Verification guided by the filter (reads up to 304 bp): an occurrence that starts at a contains the read-aligned 16-mer
// number i = ceil(a / 8), and that 16-mer is a key of the filter.  The filter hands over, per candidate, the set of its
// aligned 16-mers that can be keys (bit i: the one it confirmed in the key table and every bitmap hit behind it), so only
// the eight starts 8i-7 .. 8i of every set bit have to be looked at: one round of the warp (4 bits x 8 starts) for a
// typical candidate instead of the five rounds over all 150 starts k_ac_verify_warp makes.  Same answer: the minimum of
// (end, -length) over a superset of the occurrences' starts.
__global__ void __launch_bounds__(128)
k_ac_verify_mask(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offsets, const uint32_t* __restrict__ cand_list,
                 const uint64_t* __restrict__ cand_mask, PatternStarts ps, uint8_t* __restrict__ found, HitSink sink) {
    const uint32_t n_cand = sink.counters[3];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const uint32_t smask = (1u << ps.s_bits) - 1u;
    for (uint32_t c = gw; c < n_cand; c += nw) {
        const uint32_t r = cand_list[c];
        uint64_t mask = cand_mask[c];
        const uint64_t b = offsets[r];
        const uint32_t L = (uint32_t)(offsets[r + 1] - b);
        const uint8_t* s = bases + b;
        uint32_t best = 0xFFFFFFFFu;                                  // (end << 8) | (255 - len): min == earliest end, longest pattern
        while (mask) {
            const uint32_t warp_best = __reduce_min_sync(0xFFFFFFFFu, best);
            const int i_lo = __ffsll((long long)mask) - 1;            // no start of this or a later round lies before 8 * i_lo - 7
            if (warp_best != 0xFFFFFFFFu && (uint32_t)(8 * i_lo > 7 ? 8 * i_lo - 7 : 0) + ps.min_len > (warp_best >> 8)) break;
            int i = -1;                                               // my bit: the (lane / 8)-th lowest one
            uint64_t m = mask;
#pragma unroll
            for (uint32_t q = 0; q < 4; ++q) {
                const int lowest = m ? __ffsll((long long)m) - 1 : -1;
                if (q == (lane >> 3)) i = lowest;
                m &= m - 1;                                           // (0 & -1 == 0)
            }
            mask = m;
            const int p_signed = 8 * i - 7 + (int)(lane & 7u);
            if (i >= 0 && p_signed >= 0 && (uint32_t)p_signed + 16 <= L) {
                const uint32_t p = (uint32_t)p_signed;
                uint32_t code = 0;
#pragma unroll
                for (int k = 0; k < 16; ++k) code |= (uint32_t)((__ldg(s + p + k) >> 1) & 3u) << (2 * k);
                uint32_t pi;
                if (code == 0xFFFFFFFFu) pi = __ldg(ps.ones + 1);
                else {
                    uint32_t slot = (code * 0x85EBCA6Bu) >> (32 - ps.s_bits);
                    for (;;) {
                        const uint32_t k = __ldg(ps.s_keys + slot);
                        if (k == code) { pi = __ldg(ps.s_head + slot); break; }
                        if (k == 0xFFFFFFFFu) { pi = 0xFFFFFFFFu; break; }
                        slot = (slot + 1) & smask;
                    }
                }
                for (; pi != 0xFFFFFFFFu; pi = __ldg(ps.p_next + pi)) {
                    const uint32_t po = __ldg(ps.p_offs + pi), len = __ldg(ps.p_offs + pi + 1) - po;
                    if (p + len > L) continue;
                    const uint32_t key = ((p + len) << 8) | (255u - len);
                    if (key >= best) continue;
                    uint32_t diff = 0;
#pragma unroll 8
                    for (uint32_t k = 0; k < len; ++k) diff |= (uint32_t)__ldg(s + p + k) ^ (uint32_t)__ldg(ps.p_bytes + po + k);
                    if (diff == 0) best = key;
                }
            }
        }
        best = __reduce_min_sync(0xFFFFFFFFu, best);
        if (best != 0xFFFFFFFFu && lane == 0) {
            const uint32_t best_end = best >> 8, best_len = 255u - (best & 255u);
            uint32_t dr_end = best_end - 1;                          // on_match (libcrispr.cpp:420-437)
            if (dr_end >= L) dr_end = L - 1;
            uint32_t ss[2] = { dr_end - (best_len - 1), dr_end };
            found[r] = 1;
            emit_hit(sink, r, ss, 2, 0);
        }
    }
}

// ---- K2 generic helper: automaton walk over a candidate list (kept for pattern sets the verify kernel cannot take) -----
__global__ void __launch_bounds__(128)
k_ac_scan_list(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offsets, const uint32_t* __restrict__ cand_list,
               const uint32_t* __restrict__ table, uint32_t stride_log2, const uint8_t* __restrict__ symv_g,
               uint8_t* __restrict__ found, HitSink sink) {
    __shared__ uint8_t symv[256];
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) symv[i] = symv_g[i];
    __syncthreads();
    const uint32_t n_cand = sink.counters[3];
    const uint32_t nthreads = gridDim.x * blockDim.x;
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cand; c += nthreads) {
        const uint32_t r = cand_list[c];
        const uint64_t b = offsets[r];
        const uint32_t L = (uint32_t)(offsets[r + 1] - b);
        const uint8_t* s = bases + b;
        uint32_t st = 0, end = 0, plen = 0;
        for (uint32_t i = 0; i < L; ++i) {
            const uint32_t sy = symv[__ldg(s + i)];
            if (!sy) { st = 0; continue; }
            const uint32_t e = __ldg(table + ((size_t)st << stride_log2) + (sy - 1));
            st = e & 0xFFFFFFu;
            if (e >> 24) { end = i + 1; plen = e >> 24; break; }
        }
        if (plen) {
            uint32_t dr_end = end - 1;
            if (dr_end >= L) dr_end = L - 1;
            uint32_t ss[2] = { dr_end - (plen - 1), dr_end };
            found[r] = 1;
            emit_hit(sink, r, ss, 2, 0);
        }
    }
}

// ---- K3 ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_edit_distance(const uint8_t* __restrict__ bytes, const uint32_t* __restrict__ a_off, const uint32_t* __restrict__ a_len,
                const uint32_t* __restrict__ b_off, const uint32_t* __restrict__ b_len, uint32_t n_pairs,
                int32_t* __restrict__ out_dist, float* __restrict__ out_sim) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    GmemSeq s{bytes};
    out_dist[i] = cb::edit_distance(s, a_off[i], a_len[i], b_off[i], b_len[i]);
    out_sim[i] = cb::similarity(s, a_off[i], a_len[i], b_off[i], b_len[i]);
}

// ---- K6: partial-DR recovery (ReadHolder::updateStartStops + smithWaterman) ------------------------------------------
// The jobs are the found reads of the DR groups (1-2 % of a metagenome, most of a CRISPR-rich sample).  Each read's
// repeats are shifted to the group's consensus DR and two ends-free alignments look for a partial repeat in the flanks.
// k_update_start_stops: one WARP per read, the alignment as a lane wavefront in registers (sw_warp.cuh).
// k_update_start_stops_thread: one thread per read (sw_core.cuh, the source the CPU fuzz compiles); kept for comparison
// (CRASS_B200_K6=thread): 4 of 32 lanes busy on the bench workload because the alignments differ in size from read to read.
constexpr int kUssThreads = 64;
constexpr int kUssWarps = 4;

__global__ void __launch_bounds__(kUssThreads)
k_update_start_stops_thread(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offsets, const uint8_t* __restrict__ dr_bytes,
                            const uint32_t* __restrict__ dr_offsets, const crass_b200_uss_job* __restrict__ jobs, uint32_t n_jobs,
                            const uint32_t* __restrict__ ss_in, uint32_t low_spacer, uint32_t* __restrict__ ss_out,
                            uint32_t* __restrict__ n_out, uint8_t* __restrict__ status) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_jobs) return;
    const crass_b200_uss_job job = jobs[i];
    const uint64_t lo = offsets[job.read];
    const uint32_t L = (uint32_t)(offsets[job.read + 1] - lo);
    GmemSeq s{bases + lo};
    GmemSeq dr{dr_bytes + dr_offsets[job.dr]};
    uint32_t n = 0;
    status[i] = cb::update_start_stops(s, L, ss_in + job.ss_offset, job.n_ss, job.front_offset, dr,
                                       dr_offsets[job.dr + 1] - dr_offsets[job.dr], low_spacer, ss_out + job.out_offset, n);
    n_out[i] = n;
}

__global__ void __launch_bounds__(kUssWarps * 32)
k_update_start_stops(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offsets, const uint8_t* __restrict__ dr_bytes,
                     const uint32_t* __restrict__ dr_offsets, const crass_b200_uss_job* __restrict__ jobs, uint32_t n_jobs,
                     const uint32_t* __restrict__ ss_in, uint32_t low_spacer, uint32_t* __restrict__ ss_out,
                     uint32_t* __restrict__ n_out, uint8_t* __restrict__ status) {
    __shared__ uint8_t dr_sm[kUssWarps][cb::kMaxSwDr + 1];
    const uint32_t w = threadIdx.x >> 5;
    const uint32_t i = blockIdx.x * kUssWarps + w;
    if (i >= n_jobs) return;                                         // whole warps leave together
    const crass_b200_uss_job job = jobs[i];
    const uint64_t lo = offsets[job.read];
    const uint32_t L = (uint32_t)(offsets[job.read + 1] - lo);
    GmemSeq s{bases + lo};
    GmemSeq dr{dr_bytes + dr_offsets[job.dr]};
    uint32_t n = 0;
    const uint8_t st = cbw::update_start_stops_warp(s, L, ss_in + job.ss_offset, job.n_ss, job.front_offset, dr,
                                                    dr_offsets[job.dr + 1] - dr_offsets[job.dr], low_spacer, dr_sm[w],
                                                    ss_out + job.out_offset, n);
    if ((threadIdx.x & 31u) == 0) { status[i] = st; n_out[i] = n; }
}

// ---- KAT entry points --------------------------------------------------------------------------------
__global__ void k_scan_right_one(const uint8_t* __restrict__ seq_and_pat, uint32_t L, uint32_t* ss, uint32_t* n_ss, uint32_t cap,
                                 uint32_t w, uint32_t min_spacer, uint32_t scan_range) {
    GmemSeq s{seq_and_pat};
    uint32_t n = *n_ss;
    cb::scan_right(s, L, ss, n, cap, L, w, min_spacer, scan_range);
    *n_ss = n;
}

__global__ void k_extend_one(const uint8_t* __restrict__ seq, uint32_t L, uint32_t* ss, uint32_t n_ss, uint32_t window,
                             uint32_t min_spacer, uint32_t* replen) {
    GmemSeq s{seq};
    *replen = cb::extend_pre_repeat(s, L, ss, n_ss, window, min_spacer);
}

__global__ void k_qc_one(const uint8_t* __restrict__ seq, uint32_t L, const uint32_t* ss, uint32_t n_ss, int min_spacer, int max_spacer,
                         int* result) {
    GmemSeq s{seq};
    *result = cb::qc_found_repeats(s, L, ss, n_ss, min_spacer, max_spacer);
}

}  // namespace cbk

#include "cluster.cuh"
#include "consensus.cuh"
