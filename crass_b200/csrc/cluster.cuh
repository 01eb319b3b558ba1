// cluster.cuh -- the rest of createNonRedundantSet on the device, and the matcher tables built where the
// patterns already are.  Included by kernels.cuh after K5's first kernels (k_cl_rank / k_cl_keys / k_cl_first), whose
// arrays it continues from:
//
//   k_cl_str_canon / _insert / _first the rare 11-mers with a letter outside A/C/G/T: first DR holding each (the host's
//                                     string map, results.cpp resolve_str), through a table keyed by their canonical bytes
//   k_cl_runs                         pass B2: the k-mers of every DR folded into runs (first DR, count)
//   k_cl_walk                         pass C: clusterDRReads' greedy, order-dependent group assignment
//                                     (WorkHorse.cpp:1542-1625).  DR t only ever looks at groups of DRs before it, so the
//                                     walk is a dependency graph, not a chain: one thread per DR, each waiting for exactly
//                                     the group words it reads; a group is named by the DR that founded it until
//   k_cl_founders (+ scan)            numbers the groups in founding order
//   k_cl_hist / k_cl_members / k_cl_group_sort     the members of every group in (length, token order) order
//   k_cl_dead                         pass D: removeRedundantRepeats (WorkHorse.cpp:612-645) -- a DR goes when an earlier
//                                     member of its group is inside it on either strand
//   k_cl_place / k_cl_emit            survivors, then their reverse complements, group by group (WorkHorse.cpp:690-697)
//   k_ac_build                        bitmap, key table and pattern-start table of K2 from the pattern bytes
//
// Everything the host learns comes back in one 64-byte info record; anything these kernels do not handle (a letter whose
// complement is not an involution, an empty token, too many string-keyed k-mers, a walk that does not settle) raises a
// flag there and the caller takes the host passes instead.
#pragma once

namespace cbk {

enum : uint32_t {
    kInfoN = 0, kInfoKmers = 1, kInfoStr = 2, kInfoOdd = 3, kInfoFlags = 4, kInfoPatterns = 5, kInfoMinLen = 6, kInfoMaxLen = 7,
    kInfoBytes = 8, kInfoTicket = 9, kInfoGroups = 10, kInfoWords = 16
};
enum : uint32_t { kClFlagLetter = 1, kClFlagWalk = 2, kClFlagEmpty = 4, kClFlagStr = 8 };

struct ClusterTail {
    ClusterArrays a;
    uint32_t* lens;      // [cap]      length of DR t
    uint8_t* impure;     // [cap]      1 = DR t holds a letter outside A/C/G/T
    uint32_t* runc;      // [k-mers]   length of each run (the first DR of a run overwrites a.first in place)
    uint32_t* nruns;     // [cap]
    uint32_t* group;     // [cap]      founder of DR t's group + 1; 0 = not known yet
    uint32_t* gnum;      // [cap + 1]  founder flags -> exclusive scan = group number of a founder; [n] = number of groups
    uint32_t* gstart;    // [cap + 1]  members per group -> exclusive scan
    uint32_t* gfill;     // [cap]
    uint32_t* members;   // [cap]      DRs group by group, no order inside a group
    uint32_t* sorted;    // [cap]      DRs in (group, length, token position) order
    uint32_t* alive;     // [cap + 1]  survivor flags in that order -> exclusive scan
    uint32_t* alive1;    // [cap + 1]  the same after the first round of pass D
    uint32_t* listed;    // [cap]      sorted positions that survived the first round
    uint32_t* plen;      // [2 cap + 1] pattern lengths -> offsets
    uint32_t* psrc;      // [2 cap]    pattern -> DR t, bit 31 = reverse complement
    uint8_t* pbytes;     // pattern bytes (+ 16 bytes of zeroed slack)
    uint8_t* canon;      // [str_cap * 12] canonical bytes of the string-keyed k-mers
    uint32_t* str_rep;   // [2 str_cap] open-addressing table over the canonical bytes: entry that claimed the slot ...
    uint32_t* str_min;   // [2 str_cap] ... and the smallest DR holding that k-mer
    uint32_t* str_slot;  // [str_cap]   slot of entry i
    ulonglong4* spacked; // [cap]      packed[] in sorted order
    uint32_t* slens;     // [cap]      lens[] in sorted order; bit 31 = the DR holds a letter outside A/C/G/T (code matches must be confirmed on the bytes)
    ulonglong4* packed;  // [cap]      2-bit codes ((byte >> 1) & 3) of DR t: x, y = bases 0..31, 32..63; z, w = the same of its reverse complement
    uint32_t min_count;
    __device__ const uint8_t* dr(uint32_t t) const { return a.rec(a.order[t]) + 2; }
};

__device__ __forceinline__ uint32_t ld_now(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }

// laurenize on the bytes (SeqUtils.cpp:89-97): the smaller of the k-mer and its reverse complement
__device__ __forceinline__ void cl_canon(const uint8_t* k, uint8_t* out) {
    int cmp = 0;
    for (uint32_t i = 0; i < kClKmer && !cmp; ++i) {
        const uint8_t x = k[i], y = c_comp_tab[k[kClKmer - 1 - i] & 127];
        cmp = x < y ? -1 : x > y ? 1 : 0;
    }
    for (uint32_t i = 0; i < kClKmer; ++i) out[i] = cmp < 0 ? k[i] : c_comp_tab[k[kClKmer - 1 - i] & 127];
}

__global__ void __launch_bounds__(128)
k_cl_str_canon(ClusterTail c) {
    const uint32_t m = c.a.info[kInfoStr];
    const uint32_t i = blockIdx.x * 128 + threadIdx.x;
    if (i == 0 && m > c.a.str_cap) atomicOr(&c.a.info[kInfoFlags], kClFlagStr);
    if (i >= min(m, c.a.str_cap)) return;
    const uint32_t t = c.a.str_tq[2 * i], q = c.a.str_tq[2 * i + 1];
    uint8_t out[12];
    cl_canon(c.dr(t) + (q - c.a.koff[t]), out);
    out[11] = 0;
    for (uint32_t b = 0; b < 12; ++b) c.canon[(size_t)i * 12 + b] = out[b];
}

// (k-mer -> first DR) for the string-keyed k-mers: an open-addressing table keyed by the twelve canonical bytes; a slot
// belongs to the entry that claimed it, later entries compare their bytes with the owner's (written by the kernel before)
__global__ void __launch_bounds__(128)
k_cl_str_insert(ClusterTail c) {
    const uint32_t m = min(c.a.info[kInfoStr], c.a.str_cap);
    const uint32_t i = blockIdx.x * 128 + threadIdx.x;
    if (i >= m) return;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(c.canon);
    const uint32_t w0 = w[3 * i], w1 = w[3 * i + 1], w2 = w[3 * i + 2];
    const uint32_t mask = 2 * c.a.str_cap - 1;
    uint32_t s = ((w0 * 0x9E3779B1u) ^ (w1 * 0x85EBCA6Bu) ^ (w2 * 0xC2B2AE35u));
    s = (s ^ (s >> 15)) & mask;
    for (;;) {
        uint32_t cur = atomicCAS(&c.str_rep[s], 0xFFFFFFFFu, i);
        if (cur == 0xFFFFFFFFu) cur = i;
        if (cur == i || (w[3 * cur] == w0 && w[3 * cur + 1] == w1 && w[3 * cur + 2] == w2)) break;
        s = (s + 1) & mask;
    }
    atomicMin(&c.str_min[s], c.a.str_tq[2 * i]);
    c.str_slot[i] = s;
}

__global__ void __launch_bounds__(128)
k_cl_str_first(ClusterTail c) {
    const uint32_t m = min(c.a.info[kInfoStr], c.a.str_cap);
    const uint32_t i = blockIdx.x * 128 + threadIdx.x;
    if (i < m) c.a.first[c.a.str_tq[2 * i + 1]] = c.str_min[c.str_slot[i]];
}

// pass B2.  first[q] >= t means "never seen before this DR": such k-mers change no tally in the walk, so runs reach across them.
__global__ void __launch_bounds__(128)
k_cl_runs(ClusterTail c) {
    const uint32_t n = c.a.n();
    const uint32_t t = blockIdx.x * 128 + threadIdx.x;
    if (t >= n) return;
    const uint32_t slot = c.a.order[t];
    const uint32_t len = c.a.len_of(slot);
    const uint8_t* dr = c.a.rec(slot) + 2;
    c.lens[t] = len;
    c.group[t] = 0;
    uint32_t flags = len ? 0u : (uint32_t)kClFlagEmpty;
    unsigned long long f0 = 0, f1 = 0, r0 = 0, r1 = 0;
    bool impure = false;
    for (uint32_t i = 0; i < len; ++i) {
        const uint8_t b = dr[i];
        if (b >= 128 || c_comp_tab[c_comp_tab[b & 127]] != b) flags |= kClFlagLetter;   // 'U' -> 'A' -> 'T': containment on both strands is no longer transitive
        impure |= cl_code(b) < 0;
        if (i < 64) {
            const unsigned long long cf = (b >> 1) & 3u, cr = (c_comp_tab[b & 127] >> 1) & 3u;
            const uint32_t j = len - 1 - i;                      // where the complement of base i sits in the reverse complement
            if (i < 32) f0 |= cf << (2 * i); else f1 |= cf << (2 * (i - 32));
            if (j < 32) r0 |= cr << (2 * j); else if (j < 64) r1 |= cr << (2 * (j - 32));
        }
    }
    c.packed[t] = make_ulonglong4(f0, f1, r0, r1);
    c.impure[t] = impure ? 1 : 0;
    if (flags) atomicOr(&c.a.info[kInfoFlags], flags);
    const uint32_t q0 = c.a.koff[t], q1 = c.a.koff[t + 1];
    uint32_t out = q0;
    for (uint32_t q = q0; q < q1; ++q) {
        const uint32_t f = c.a.first[q];
        if (f >= t) continue;
        if (out > q0 && c.a.first[out - 1] == f) c.runc[out - 1]++;
        else { c.a.first[out] = f; c.runc[out] = 1; ++out; }
    }
    c.nruns[t] = out - q0;
}

// pass C.  CTAs take their 128 DRs in ticket order, so every DR a thread can wait for belongs to a CTA that has started.
// A lane never blocks inside an iteration (it looks at one run, and if that run's group is not known yet it simply
// comes back), so the lanes of a warp cannot starve each other.
constexpr uint32_t kClWalkSpins = 1u << 20;

__global__ void __launch_bounds__(128)
k_cl_walk(ClusterTail c) {
    __shared__ uint32_t base_s;
    if (threadIdx.x == 0) base_s = atomicAdd(&c.a.info[kInfoTicket], 1u) * 128u;
    __syncthreads();
    const uint32_t n = c.a.n();
    const uint32_t t = base_s + threadIdx.x;
    bool done = t >= n;
    uint32_t q0 = 0, nr = 0, j = 0, spins = 0;
    if (!done) { q0 = c.a.koff[t]; nr = c.nruns[t]; }
    while (!__all_sync(0xFFFFFFFFu, done)) {
        uint32_t mine = 0;
        if (done) {
        } else if (j < nr) {
            const uint32_t g = ld_now(&c.group[c.a.first[q0 + j]]);
            if (g == 0) {
                if (++spins > kClWalkSpins) { atomicOr(&c.a.info[kInfoFlags], kClFlagWalk); mine = t + 1; }
            } else {
                // the tally of group g after this run, and whether an earlier run opened it
                uint32_t tally = c.runc[q0 + j];
                bool opened = false;
                for (uint32_t i = 0; i < j; ++i)
                    if (ld_now(&c.group[c.a.first[q0 + i]]) == g) { tally += c.runc[q0 + i]; opened = true; }
                if (opened ? tally >= c.min_count : (tally >= 2 && tally >= c.min_count)) mine = g;
                ++j;
            }
        } else mine = t + 1;                                     // no group reached min_count: the DR founds one
        if (mine) {
            *reinterpret_cast<volatile uint32_t*>(&c.group[t]) = mine;
            done = true;
        }
    }
}

__global__ void __launch_bounds__(256)
k_cl_founders(ClusterTail c) {
    const uint32_t n = c.a.n();
    const uint32_t t = blockIdx.x * 256 + threadIdx.x;
    if (t == 0) c.a.info[kInfoMinLen] = 0xFFFFFFFFu;
    if (t > c.a.cap) return;
    c.gnum[t] = t < n && c.group[t] == t + 1 ? 1u : 0u;
    c.gstart[t] = 0;
    if (t < c.a.cap) c.gfill[t] = 0;
}

__device__ __forceinline__ uint32_t cl_group_of(const ClusterTail& c, uint32_t t) { return c.gnum[c.group[t] - 1]; }

__global__ void __launch_bounds__(256)
k_cl_hist(ClusterTail c) {
    const uint32_t n = c.a.n();
    const uint32_t t = blockIdx.x * 256 + threadIdx.x;
    if (t == 0) c.a.info[kInfoGroups] = c.gnum[n];
    if (t < n) atomicAdd(&c.gstart[cl_group_of(c, t)], 1u);
}

__global__ void __launch_bounds__(256)
k_cl_members(ClusterTail c) {
    const uint32_t n = c.a.n();
    const uint32_t t = blockIdx.x * 256 + threadIdx.x;
    if (t >= n) return;
    const uint32_t g = cl_group_of(c, t);
    c.members[c.gstart[g] + atomicAdd(&c.gfill[g], 1u)] = t;
}

// position inside the group = members that come before (shorter, or as long with a smaller token position): the order of
// the stable sort by length in removeRedundantRepeats.  One warp per DR.
__global__ void __launch_bounds__(128)
k_cl_group_sort(ClusterTail c) {
    const uint32_t n = c.a.n();
    const uint32_t t = (blockIdx.x * 128 + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (t >= n) return;
    const uint32_t g = cl_group_of(c, t);
    const uint32_t gs = c.gstart[g], ge = c.gstart[g + 1], len = c.lens[t];
    uint32_t before = 0;
    for (uint32_t i = gs + lane; i < ge; i += 32) {
        const uint32_t m = c.members[i], lm = c.lens[m];
        before += (lm < len || (lm == len && m < t)) ? 1u : 0u;
    }
    before = __reduce_add_sync(0xFFFFFFFFu, before);
    if (lane == 0) { c.sorted[gs + before] = t; c.spacked[gs + before] = c.packed[t]; c.slens[gs + before] = len | (c.impure[t] ? 0x80000000u : 0u); }
}

// pass D on the 2-bit codes (DRs up to 64 bases: every default-geometry token).  b's code sits in registers and slides by one
// base per step; a lane holds one earlier member a (code, code of its reverse complement, mask of its length) and compares
// 128 bits per position and strand.  Equal bytes give equal codes, so the codes can only over-report ('N' codes like 'G'):
// a code match is confirmed on the bytes before it counts.
__device__ __forceinline__ bool cl_bytes_at(const uint8_t* b, uint32_t at, const uint8_t* a, uint32_t la, bool revcomp) {
    if (!revcomp) { for (uint32_t i = 0; i < la; ++i) if (b[at + i] != a[i]) return false; }
    else { for (uint32_t i = 0; i < la; ++i) if (b[at + i] != c_comp_tab[a[la - 1 - i] & 127]) return false; }
    return true;
}

// Most members die, and most of those to one of the shortest members of their group, so the test runs in two rounds:
// PHASE 1 tries only the first kClDeadFirst members of the group; what survives that (a superset of the real survivors) is
// listed (k_cl_dead_list, after a scan of the flags), and PHASE 2 tries every listed earlier member on the listed DRs only.
constexpr uint32_t kClDeadFirst = 64;

template <int PHASE>
__global__ void __launch_bounds__(128)
k_cl_dead_packed(ClusterTail c) {
    const uint32_t n = c.a.n();
    const uint32_t s = (blockIdx.x * 128 + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (s == n && lane == 0) c.alive[n] = 0;
    if (s >= n) return;
    if (PHASE == 2 && c.alive1[s + 1] == c.alive1[s]) { if (lane == 0) c.alive[s] = 0; return; }     // died in phase 1 (alive1 = scan of its flags)
    const uint32_t tb = c.sorted[s];
    const uint32_t lb = c.slens[s] & 0x7FFFFFFFu;
    const bool b_impure = (c.slens[s] >> 31) != 0;
    const ulonglong4 pb = c.spacked[s];
    const uint32_t gs = c.gstart[cl_group_of(c, tb)];
    // phase 1: positions [gs, min(s, gs + 64)) of the sorted order; phase 2: entries [alive1[gs], alive1[s]) of the list
    const uint32_t j_begin = PHASE == 1 ? gs : c.alive1[gs];
    const uint32_t j_end = PHASE == 1 ? min(s, gs + kClDeadFirst) : c.alive1[s];
    bool dead = false;
    for (uint32_t j0 = j_begin; j0 < j_end; j0 += 32) {
        const uint32_t j = j0 + lane;
        if (j < j_end) {
            const uint32_t i = PHASE == 1 ? j : c.listed[j];
            const uint32_t la = c.slens[i] & 0x7FFFFFFFu;        // <= lb by the order
            const bool confirm = b_impure || (c.slens[i] >> 31) != 0;   // A, C, G, T have four different codes: nothing to confirm between two such DRs
            const ulonglong4 pa = c.spacked[i];
            const unsigned long long m0 = la >= 32 ? ~0ull : (1ull << (2 * la)) - 1ull;
            const unsigned long long m1 = la <= 32 ? 0ull : la >= 64 ? ~0ull : (1ull << (2 * (la - 32))) - 1ull;
            unsigned long long b0 = pb.x, b1 = pb.y;
            for (uint32_t at = 0; at + la <= lb && !dead; ++at) {
                const bool fw = (((b0 ^ pa.x) & m0) | ((b1 ^ pa.y) & m1)) == 0;
                const bool rc = (((b0 ^ pa.z) & m0) | ((b1 ^ pa.w) & m1)) == 0;
                if ((fw || rc) && confirm) {
                    const uint8_t* bg = c.dr(tb);
                    const uint8_t* ag = c.dr(c.sorted[i]);
                    dead = (fw && cl_bytes_at(bg, at, ag, la, false)) || (rc && cl_bytes_at(bg, at, ag, la, true));
                } else dead = fw || rc;
                b0 = (b0 >> 2) | (b1 << 62);
                b1 >>= 2;
            }
        }
        if (__any_sync(0xFFFFFFFFu, dead)) { dead = true; break; }
    }
    if (lane == 0) {
        if (PHASE == 1) { c.alive1[s] = dead ? 0u : 1u; if (s + 1 == n) c.alive1[n] = 0; }
        else c.alive[s] = dead ? 0u : 1u;
    }
}

__global__ void __launch_bounds__(256)
k_cl_dead_list(ClusterTail c) {
    const uint32_t n = c.a.n();
    const uint32_t s = blockIdx.x * 256 + threadIdx.x;
    if (s < n && c.alive1[s + 1] != c.alive1[s]) c.listed[c.alive1[s]] = s;
}

// pass D on the bytes (tokens longer than 64 bases: wider DR bounds than the default).  One warp per DR b, lanes over the members in front of it; containment is transitive on both strands (for the
// letters k_cl_runs lets through), so "an earlier member" and the reference's "an earlier survivor" are the same test.
__global__ void __launch_bounds__(128)
k_cl_dead(ClusterTail c) {
    extern __shared__ uint8_t cl_b_smem[];
    const uint32_t n = c.a.n();
    const uint32_t s = (blockIdx.x * 128 + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (s == n && lane == 0) c.alive[n] = 0;
    if (s >= n) return;
    uint8_t* b = cl_b_smem + (threadIdx.x >> 5) * c.a.stride;
    const uint32_t tb = c.sorted[s];
    const uint32_t lb = c.lens[tb];
    const uint8_t* bg = c.dr(tb);
    for (uint32_t i = lane; i < lb; i += 32) b[i] = bg[i];
    __syncwarp();
    const uint32_t gs = c.gstart[cl_group_of(c, tb)];
    bool dead = false;
    for (uint32_t i0 = gs; i0 < s; i0 += 32) {
        const uint32_t i = i0 + lane;
        if (i < s) {
            const uint32_t ta = c.sorted[i];
            const uint32_t la = c.lens[ta];                      // <= lb by the order
            const uint8_t* a = c.dr(ta);
            const uint8_t a0 = a[0], ar0 = c_comp_tab[a[la - 1] & 127];
            for (uint32_t at = 0; at + la <= lb && !dead; ++at) {
                if (b[at] == a0) {
                    uint32_t k = 1;
                    while (k < la && b[at + k] == a[k]) ++k;
                    dead = k == la;
                }
                if (!dead && b[at] == ar0) {
                    uint32_t k = 1;
                    while (k < la && b[at + k] == c_comp_tab[a[la - 1 - k] & 127]) ++k;
                    dead = k == la;
                }
            }
        }
        if (__any_sync(0xFFFFFFFFu, dead)) { dead = true; break; }
    }
    if (lane == 0) c.alive[s] = dead ? 0u : 1u;
}

// survivors of group g go to patterns [2 S(gs), 2 S(gs) + m), their reverse complements to the m slots behind them
// (S = survivors in front, m = survivors of the group)
__global__ void __launch_bounds__(256)
k_cl_place(ClusterTail c) {
    const uint32_t n = c.a.n();
    const uint32_t s = blockIdx.x * 256 + threadIdx.x;
    if (s == 0) c.a.info[kInfoPatterns] = 2 * c.alive[n];
    if (s >= n || c.alive[s + 1] == c.alive[s]) return;           // alive[] holds the exclusive scan of the flags by now
    const uint32_t t = c.sorted[s];
    const uint32_t g = cl_group_of(c, t);
    const uint32_t s0 = c.alive[c.gstart[g]], m = c.alive[c.gstart[g + 1]] - s0;
    const uint32_t at = 2 * s0 + (c.alive[s] - s0);
    c.plen[at] = c.plen[at + m] = c.lens[t];
    c.psrc[at] = t;
    c.psrc[at + m] = t | 0x80000000u;
}

__global__ void __launch_bounds__(128)
k_cl_emit(ClusterTail c) {
    const uint32_t np = c.a.info[kInfoPatterns];
    const uint32_t p = blockIdx.x * 128 + threadIdx.x;
    if (p == 0) {
        const uint32_t total = c.plen[np];
        c.a.info[kInfoBytes] = total;
        for (uint32_t i = 0; i < 16; ++i) c.pbytes[total + i] = 0;
    }
    if (p >= np) return;
    const uint32_t src = c.psrc[p], t = src & 0x7FFFFFFFu;
    const uint32_t len = c.lens[t];
    const uint8_t* dr = c.dr(t);
    uint8_t* out = c.pbytes + c.plen[p];
    if (src >> 31) for (uint32_t i = 0; i < len; ++i) out[i] = c_comp_tab[dr[len - 1 - i] & 127];
    else for (uint32_t i = 0; i < len; ++i) out[i] = dr[i];
    atomicMin(&c.a.info[kInfoMinLen], len);
    atomicMax(&c.a.info[kInfoMaxLen], len);
}

// ---- matcher tables of K2 (host/ac_build.cpp describes them) built on the device ----------------------------------------
// One thread per (pattern, window 0..7): sets the bitmap bits, claims the key-table slot of the window's 16-mer and, for
// window 0, chains the pattern into the start table.  Open addressing with compare-and-swap ends in the same kind of table
// a sequential fill produces (a look-up probes until it meets its key or an empty slot), and the chains of the start table
// are walked without early exit, so their order does not matter.  The all-ones code (sixteen G-coded bases) is the empty
// marker and cannot be a key: ones[0] says whether some pattern has it, ones[1] heads the chain of patterns that begin
// with it.
struct MatcherTables {
    const uint8_t* pbytes;
    const uint32_t* poffs;
    uint32_t n_patterns;
    uint32_t* bitmap;  uint32_t bits;  uint32_t hashes;
    uint32_t* bitmap_small;  uint32_t bits_small;       // 0 = none
    uint32_t* keys;  uint32_t table_bits;
    uint32_t* s_keys;  uint32_t* s_head;  uint32_t s_bits;
    uint32_t* p_next;
    uint32_t* ones;
};

__global__ void __launch_bounds__(256)
k_ac_build(MatcherTables m) {
    const uint32_t i = (blockIdx.x * 256 + threadIdx.x) >> 3, w = threadIdx.x & 7u;
    if (i >= m.n_patterns) return;
    const uint8_t* p = m.pbytes + m.poffs[i] + w;
    uint32_t code = 0;
#pragma unroll
    for (uint32_t k = 0; k < 16; ++k) code |= (uint32_t)((p[k] >> 1) & 3) << (2 * k);
    const uint32_t h = (code * 0x9E3779B1u) >> (32 - m.bits);
    atomicOr(&m.bitmap[h >> 5], 1u << (h & 31));
    if (m.bits_small) { const uint32_t hs = h >> (m.bits - m.bits_small); atomicOr(&m.bitmap_small[hs >> 5], 1u << (hs & 31)); }
    if (m.hashes > 1) {                                                          // second hash of the Bloom filter (qgram_second_bit)
        const uint32_t h2 = (code * kQgramHash2) >> (32 - m.bits);
        atomicOr(&m.bitmap[h2 >> 5], 1u << (h2 & 31));
        if (m.bits_small) { const uint32_t hs = h2 >> (m.bits - m.bits_small); atomicOr(&m.bitmap_small[hs >> 5], 1u << (hs & 31)); }
    }
    if (code == 0xFFFFFFFFu) {
        m.ones[0] = 1;
        if (w == 0) m.p_next[i] = atomicExch(&m.ones[1], i);
        return;
    }
    const uint32_t tmask = (1u << m.table_bits) - 1u;
    uint32_t slot = (code * 0x85EBCA6Bu) >> (32 - m.table_bits);
    for (;;) {
        const uint32_t cur = atomicCAS(&m.keys[slot], 0xFFFFFFFFu, code);
        if (cur == 0xFFFFFFFFu || cur == code) break;
        slot = (slot + 1) & tmask;
    }
    if (w == 0) {
        const uint32_t smask = (1u << m.s_bits) - 1u;
        uint32_t s = (code * 0x85EBCA6Bu) >> (32 - m.s_bits);
        for (;;) {
            const uint32_t cur = atomicCAS(&m.s_keys[s], 0xFFFFFFFFu, code);
            if (cur == 0xFFFFFFFFu || cur == code) break;
            s = (s + 1) & smask;
        }
        m.p_next[i] = atomicExch(&m.s_head[s], i);
    }
}

}  // namespace cbk
