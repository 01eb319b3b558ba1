// consensus.cuh -- K7: the consensus ("true") DR of a group on the device (second half).
//
// WorkHorse::parseGroupedDRs (WorkHorse.cpp:1135-1171) takes the longest DR of a group as the master, aligns every other DR
// (slave) of the group against it on both strands with ksw_align (Aligner::getOffsetAgainstMaster, Aligner.cpp:263-362), lays
// all reads of the master and of the placed slaves into a coverage array at the offsets the alignments give
// (placeReadsInCoverageArray, :364-417) and reads consensus, conservation and the DR zone off it (generateConsensus, :147-246).
//
//   k_ksw_align       ksw_align (ksw.c:330-354) with the Aligner's scores.  The reference runs the striped SSE2 kernel ksw_i16
//                     (ksw.c:219-322): eight 16-bit lanes, query position j + lane*slen in lane `lane` of vector j, a lazy-F
//                     loop that stops as soon as no lane can still improve.  Its results depend on that layout in the corner
//                     cases (E is not refreshed by the lazy loop), so the layout is kept: EIGHT THREADS per alignment, one per
//                     SSE lane, the byte shifts of the vectors are shuffles inside the group, the horizontal maximum and the
//                     movemask test are group votes.  Four alignments per warp; a job is one (query, strand, target) triple
//                     and runs the forward pass and, for the start positions, the pass over the reversed prefixes.
//   k_cons_decide     forward against reverse score, the two failure tests, offset = tb - qb
//   k_cons_ext_find / k_cons_ext_build   extendSlaveDR (:420-452) for slaves whose two scores are equal: two more bases on
//                     either side from the first read that has them, then k_ksw_align / k_cons_decide once more
//   k_cons_place      one warp per read: every full-length repeat of the read's DR puts the read into the coverage rows
//                     (reads of reversed slaves as their reverse complement with mirrored start/stop lists)
//   k_cons_columns / k_cons_zone         consensus base, conservation, and the zone trimmed and grown at the 0.55 cut-off
#pragma once

namespace cbk {

constexpr int kKswMaxSlen = 16;                      // query length up to 128 (DRs are below 100)
constexpr int kKswMaxTarget = 4096;                  // (any length works; a bound against runaway jobs)

struct KswJob {
    uint32_t q_off, q_len;                           // query letters in the pool
    uint32_t t_off, t_len;                           // target letters in the pool
    uint32_t q_rc;                                   // 1 = align the reverse complement of the query
    int32_t xtra;                                    // KSW_X* flags | threshold (ksw.h:6-9)
};
struct KswResult { int32_t score, te, qe, score2, te2, tb, qb, pad; };

__device__ __forceinline__ int nt4_code(uint8_t c) {                   // Aligner::seq_nt4_table (Aligner.cpp:41-58)
    const uint8_t u = c & 0xDF;
    return (c < 128 && u == 'A') ? 0 : (c < 128 && u == 'C') ? 1 : (c < 128 && u == 'G') ? 2 : (c < 128 && u == 'T') ? 3 : 4;
}
__device__ __forceinline__ int ksw_score(int a, int b) { return (a == 4 || b == 4) ? 0 : (a == b ? 1 : -3); }   // Aligner.h:119-131
__device__ __forceinline__ int subu16(int a, int b) { return a > b ? a - b : 0; }                                  // _mm_subs_epu16 on values >= 0

struct KswSeqs {                                     // how a pass sees the two sequences
    const uint8_t* q; int qlen; bool q_rc;           // forward pass: the query as given (or its reverse complement)
    const uint8_t* t; int tlen;
    int q_flip, t_flip;                              // second pass: positions <= flip are mirrored (revseq of the prefixes), -1 = off
    __device__ int qcode(int k) const {
        if (q_flip >= 0) k = q_flip - k;
        return q_rc ? nt4_code(c_comp_tab[q[qlen - 1 - k] & 127]) : nt4_code(q[k]);
    }
    __device__ int tcode(int i) const {
        if (i <= t_flip) i = t_flip - i;
        return nt4_code(t[i]);
    }
};

// one pass of ksw_i16 by the eight lanes of a group; every lane returns the same result
__device__ KswResult ksw_i16_lanes(uint32_t lane, uint32_t gmask, const KswSeqs& s, int qlen, int xtra) {
    KswResult r{0, -1, -1, -1, -1, -1, -1, 0};
    const int slen = (qlen + 7) >> 3, gapoe = 7, gape = 2;
    const int minsc = (xtra & 0x40000) ? (xtra & 0xffff) : 0x10000;
    const int endsc = (xtra & 0x20000) ? (xtra & 0xffff) : 0x10000;
    short Ha[kKswMaxSlen], Hb[kKswMaxSlen], E[kKswMaxSlen], Hmax[kKswMaxSlen];
    signed char qc[kKswMaxSlen];
    for (int j = 0; j < slen; ++j) {
        const int k = j + (int)lane * slen;
        qc[j] = (signed char)(k < qlen ? s.qcode(k) : -1);
        Ha[j] = Hb[j] = E[j] = Hmax[j] = 0;
    }
    short* H0 = Ha;
    short* H1 = Hb;
    int gmax = 0, te = -1;
    for (int i = 0; i < s.tlen; ++i) {
        const int tc = s.tcode(i);
        int f = 0, mx = 0;
        int h = __shfl_up_sync(gmask, (int)H0[slen - 1], 1, 8);
        if (lane == 0) h = 0;
        for (int j = 0; j < slen; ++j) {
            h += qc[j] < 0 ? 0 : ksw_score(tc, qc[j]);
            int e = E[j];
            h = max(h, e);
            h = max(h, f);
            mx = max(mx, h);
            H1[j] = (short)h;
            h = subu16(h, gapoe);
            e = max(subu16(e, gape), h);
            E[j] = (short)e;
            f = max(subu16(f, gape), h);
            h = H0[j];
        }
        bool open = true;
        for (int k = 0; k < 16 && open; ++k) {                           // the lazy-F loop
            f = __shfl_up_sync(gmask, f, 1, 8);
            if (lane == 0) f = 0;
            for (int j = 0; j < slen; ++j) {
                int hh = max((int)H1[j], f);
                H1[j] = (short)hh;
                hh = subu16(hh, gapoe);
                f = subu16(f, gape);
                if (!(__ballot_sync(gmask, f > hh) & gmask)) { open = false; break; }
            }
        }
        int imax = mx;
        imax = max(imax, __shfl_xor_sync(gmask, imax, 4, 8));
        imax = max(imax, __shfl_xor_sync(gmask, imax, 2, 8));
        imax = max(imax, __shfl_xor_sync(gmask, imax, 1, 8));
        if (imax > gmax) {
            gmax = imax; te = i;
            for (int j = 0; j < slen; ++j) Hmax[j] = H1[j];
            if (gmax >= endsc) break;
        }
        short* t = H1; H1 = H0; H0 = t;
    }
    r.score = gmax; r.te = te;
    // qe: first maximum of Hmax in the order of the vector memory, i = 8 j + lane
    int best = -1, best_i = 0x7fffffff;
    for (int j = 0; j < slen; ++j)
        if ((int)Hmax[j] > best) { best = Hmax[j]; best_i = 8 * j + (int)lane; }
    for (int d = 4; d; d >>= 1) {
        const int ob = __shfl_xor_sync(gmask, best, d, 8), oi = __shfl_xor_sync(gmask, best_i, d, 8);
        if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
    }
    r.qe = best_i / 8 + best_i % 8 * slen;
    // score2 / te2 stay -1: the reference's copy of ksw.c compares the candidates with (uint32_t)r.score2 = 0xFFFFFFFF
    // (ksw.c:204,314), so its list of sub-optimal columns never yields one
    (void)minsc;
    return r;
}

__global__ void __launch_bounds__(128)
k_ksw_align(const uint8_t* __restrict__ pool, const KswJob* __restrict__ jobs, const uint32_t* __restrict__ n_pairs_dev, uint32_t n_jobs,
            KswResult* __restrict__ out) {
    if (n_pairs_dev) n_jobs = min(n_jobs, 2u * *n_pairs_dev);                // a list built on the device: (forward, reverse) pairs
    const uint32_t job = (blockIdx.x * 128 + threadIdx.x) >> 3, lane = threadIdx.x & 7u;
    if (job >= n_jobs) return;
    const uint32_t gmask = 0xFFu << (threadIdx.x & 24u);
    const KswJob jb = jobs[job];
    KswResult r{0, -1, -1, -1, -1, -1, -1, 0};
    if (jb.q_len >= 1 && jb.q_len <= 8 * kKswMaxSlen && jb.t_len <= kKswMaxTarget) {
        KswSeqs s{pool + jb.q_off, (int)jb.q_len, jb.q_rc != 0, pool + jb.t_off, (int)jb.t_len, -1, -1};
        r = ksw_i16_lanes(lane, gmask, s, (int)jb.q_len, jb.xtra);
        if ((jb.xtra & 0x80000) && !((jb.xtra & 0x40000) && r.score < (jb.xtra & 0xffff))) {
            s.q_flip = r.qe; s.t_flip = r.te;                            // revseq of the two prefixes; the target keeps its full length
            const KswResult rr = ksw_i16_lanes(lane, gmask, s, r.qe + 1, 0x20000 | r.score);
            if (r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
        }
    } else r.pad = 1;                                                    // not a job this kernel takes
    if (lane == 0) out[job] = r;
}

// ---- the group around it --------------------------------------------------------------------------------------------------
enum : uint32_t { kAlReversed = 1, kAlFailed = 2, kAlEqual = 4 };

struct ConsArrays {
    const uint8_t* bases;  const uint64_t* offsets;  uint32_t n_reads;   // the reads of all groups
    const uint32_t* read_dr;                                             // DR (global index) each read hangs on
    const uint32_t* ss_offsets;  const uint32_t* ss_pool;
    uint8_t* pool;                                                       // DR letters, then one slot of ext_stride bytes per DR for the extended strings
    const uint32_t* dr_offsets;  uint32_t n_drs;
    const uint32_t* dr_group;                                            // group of DR d
    const uint32_t* group_first;  uint32_t n_groups;                     // first DR of a group = its master
    uint32_t array_len, ext_base, ext_stride;
    int32_t* dr_place;  uint8_t* dr_flags;
    uint32_t* ext_read;  uint32_t* ext_len;                              // per DR: first read that can lend the extension, length of the extended string
    KswJob* jobs2;  uint32_t* n_jobs2;                                   // second round (extended slaves): job list and its length
    uint32_t* job2_dr;
    int32_t* coverage;  uint8_t* consensus;  float* conservation;  int32_t* zone;  uint32_t* n_good;
    __device__ uint32_t dr_len(uint32_t d) const { return dr_offsets[d + 1] - dr_offsets[d]; }
    __device__ bool is_master(uint32_t d) const { return group_first[dr_group[d]] == d; }
};

// results of the jobs (2 s, 2 s + 1) = (forward, reverse) of the s-th entry of `dr_of` (round 1: every DR, masters idle; round 2: job2_dr)
__global__ void __launch_bounds__(128)
k_cons_decide(ConsArrays c, const KswResult* __restrict__ res, const uint32_t* __restrict__ dr_of, const uint32_t* __restrict__ n_dev, uint32_t n, int round) {
    if (n_dev) n = min(n, *n_dev);
    const uint32_t s = blockIdx.x * 128 + threadIdx.x;
    if (s >= n) return;
    const uint32_t d = dr_of ? dr_of[s] : s;
    const uint32_t g = c.dr_group[d];
    const int master_at = (int)(c.array_len * 0.5);                      // CRASS_DEF_CONS_ARRAY_START
    if (c.is_master(d)) { c.dr_place[d] = master_at; c.dr_flags[d] = 0; c.ext_read[d] = 0xFFFFFFFFu; return; }
    const int qlen = round == 1 ? (int)c.dr_len(d) : (int)c.ext_len[d];
    const KswResult f = res[2 * s], v = res[2 * s + 1];
    uint32_t flags = 0;
    int off = 0;
    if (qlen == 0 || v.score == f.score) flags |= kAlEqual;
    else {
        const KswResult& best = v.score > f.score ? v : f;
        if (v.score > f.score) flags |= kAlReversed;
        if (qlen / 2 > best.score || best.score < 5) flags |= kAlFailed;
        else off = best.tb - best.qb;
    }
    if (round == 1 && (flags & kAlEqual)) {                              // settled by the second round
        c.dr_flags[d] = (uint8_t)kAlEqual; c.dr_place[d] = -1; c.ext_read[d] = 0xFFFFFFFFu;
        return;
    }
    if (flags & kAlEqual) flags |= kAlFailed;
    if (flags & kAlFailed) flags &= ~kAlReversed;                        // alignSlave returns before it turns the reads round
    c.dr_flags[d] = (uint8_t)flags;
    c.dr_place[d] = (flags & kAlFailed) ? -1 : master_at + off;
    if (round == 1) c.ext_read[d] = 0xFFFFFFFFu;
    (void)g;
}

__device__ __forceinline__ int first_full_repeat(const uint32_t* ss, uint32_t n_ss, int dr_len) {
    for (uint32_t k = 0; k + 1 < n_ss; k += 2) if ((int)ss[k + 1] - (int)ss[k] == dr_len - 1) return (int)k;
    return -1;
}

__global__ void __launch_bounds__(128)
k_cons_ext_find(ConsArrays c) {
    const uint32_t i = blockIdx.x * 128 + threadIdx.x;
    if (i >= c.n_reads) return;
    const uint32_t d = c.read_dr[i];
    if (d >= c.n_drs || !(c.dr_flags[d] & kAlEqual) || c.dr_place[d] != -1) return;
    const uint32_t* ss = c.ss_pool + c.ss_offsets[i];
    const uint32_t n_ss = c.ss_offsets[i + 1] - c.ss_offsets[i];
    const int L = (int)(c.offsets[i + 1] - c.offsets[i]);
    const int k = first_full_repeat(ss, n_ss, (int)c.dr_len(d));
    if (k < 0) return;
    if ((int)ss[k] - 2 < 0 || (int)ss[k + 1] + 2 > L) return;
    atomicMin(&c.ext_read[d], i);
}

__global__ void __launch_bounds__(128)
k_cons_ext_build(ConsArrays c, int xtra) {
    const uint32_t d = blockIdx.x * 128 + threadIdx.x;
    if (d >= c.n_drs || c.is_master(d) || !(c.dr_flags[d] & kAlEqual) || c.dr_place[d] != -1) return;
    uint32_t ln = 0;
    const uint32_t i = c.ext_read[d];
    uint8_t* ext = c.pool + c.ext_base + (size_t)d * c.ext_stride;
    if (i != 0xFFFFFFFFu) {
        const uint32_t* ss = c.ss_pool + c.ss_offsets[i];
        const uint32_t n_ss = c.ss_offsets[i + 1] - c.ss_offsets[i];
        const uint32_t L = (uint32_t)(c.offsets[i + 1] - c.offsets[i]);
        const int k = first_full_repeat(ss, n_ss, (int)c.dr_len(d));
        const uint32_t st = ss[k] - 2;
        ln = c.dr_len(d) + 4;
        if (st + ln > L) ln = L - st;                                    // std::string::substr stops at the end of the read
        if (ln > c.ext_stride) ln = c.ext_stride;
        for (uint32_t b = 0; b < ln; ++b) ext[b] = c.bases[c.offsets[i] + st + b];
    }
    c.ext_len[d] = ln;
    const uint32_t s = atomicAdd(c.n_jobs2, 1u);
    c.job2_dr[s] = d;
    const uint32_t m = c.group_first[c.dr_group[d]];
    const KswJob jb{(uint32_t)(c.ext_base + (size_t)d * c.ext_stride), ln, c.dr_offsets[m], c.dr_len(m), 0u, xtra};
    c.jobs2[2 * s] = jb;
    KswJob jr = jb; jr.q_rc = 1;
    c.jobs2[2 * s + 1] = jr;
}

// placeReadsInCoverageArray for one read, one warp
__global__ void __launch_bounds__(128)
k_cons_place(ConsArrays c, uint32_t* __restrict__ status) {
    const uint32_t i = (blockIdx.x * 128 + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (i >= c.n_reads) return;
    const uint32_t d = c.read_dr[i];
    if (d >= c.n_drs) return;
    const int place = c.dr_place[d];
    if (place < 0) return;                                               // a slave that could not be placed
    const bool rev = (c.dr_flags[d] & kAlReversed) != 0;
    const int dr_len = (int)c.dr_len(d);
    const uint32_t* ss_in = c.ss_pool + c.ss_offsets[i];
    const uint32_t n_ss = c.ss_offsets[i + 1] - c.ss_offsets[i];
    const uint8_t* seq = c.bases + c.offsets[i];
    const uint32_t L = (uint32_t)(c.offsets[i + 1] - c.offsets[i]);
    int32_t* cov = c.coverage + (size_t)c.dr_group[d] * 4 * c.array_len;
    auto ss = [&](uint32_t k) -> int { return rev ? (int)(L - 1 - ss_in[n_ss - 1 - k]) : (int)ss_in[k]; };   // reverseStartStops (ReadHolder.cpp:321-380)
    uint32_t k = 0;
    while (k + 1 < n_ss && ss(k + 1) - ss(k) != dr_len - 1) k += 2;
    if (k + 1 >= n_ss) { if (lane == 0) atomicOr(status, 1u); return; }    // no full-length repeat: the reference would run off the list
    do {
        if (ss(k + 1) - ss(k) == dr_len - 1) {
            const int start_pos = place - ss(k);
            for (uint32_t b = lane; b < L; b += 32) {
                const uint8_t ch = rev ? c_comp_tab[seq[L - 1 - b] & 127] : seq[b];
                const int at = (int)b + start_pos;
                if (at < 0 || at >= (int)c.array_len) { atomicOr(status, 2u); continue; }
                const uint8_t u = ch & 0xDF;
                const int row = (ch < 128 && u == 'C') ? 1 : (ch < 128 && u == 'G') ? 2 : (ch < 128 && u == 'T') ? 3 : 0;   // CHAR_TO_INDEX (Aligner.cpp:61-70)
                atomicAdd(&cov[(size_t)row * c.array_len + (uint32_t)at], 1);
            }
        }
        k += 2;
        if (k >= (n_ss / 2) * 2) break;
    } while (ss(k + 1) - ss(k) == dr_len - 1);
}

__global__ void __launch_bounds__(256)
k_cons_columns(ConsArrays c) {
    const uint32_t x = blockIdx.x * 256 + threadIdx.x;
    if (x >= c.n_groups * c.array_len) return;
    const uint32_t g = x / c.array_len, j = x % c.array_len;
    const int32_t* cov = c.coverage + (size_t)g * 4 * c.array_len;
    int max_count = 0;
    float total = 0.0f;
    uint8_t cons = 'N';
    const char alphabet[4] = {'A', 'C', 'G', 'T'};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int v = cov[(size_t)k * c.array_len + j];
        total = __fadd_rn(total, (float)v);
        if (v > max_count) { max_count = v; cons = (uint8_t)alphabet[k]; }
    }
    c.consensus[x] = cons;
    float cv = 0.0f;
    if (total > 2.0f) { cv = __fdiv_rn((float)max_count, total); atomicAdd(&c.n_good[g], 1u); }
    c.conservation[x] = cv;
}

__global__ void __launch_bounds__(64)
k_cons_zone(ConsArrays c) {
    const uint32_t g = blockIdx.x * 64 + threadIdx.x;
    if (g >= c.n_groups) return;
    const uint32_t m = c.group_first[g];
    const float* cons = c.conservation + (size_t)g * c.array_len;
    const int n = (int)c.array_len;
    int zs = c.dr_place[m], ze = c.dr_place[m] + (int)c.dr_len(m) - 1;  // calculateDRZone (Aligner.cpp:456-484)
    if (c.n_good[g] >= 2) {
        while (zs > 0 && zs <= n) { if (cons[zs - 1] < 0.55f) zs++; else break; }      // (sic: the zone shrinks while its neighbour is poor)
        while (ze < n - 1 && ze >= -1) { if (cons[ze + 1] < 0.55f) ze--; else break; }
    }
    while (zs > 0 && zs <= n) { if (cons[zs - 1] >= 0.55f) zs--; else break; }
    while (ze < n - 1 && ze >= -1) { if (cons[ze + 1] >= 0.55f) ze++; else break; }
    c.zone[2 * g] = zs; c.zone[2 * g + 1] = ze;
}

}  // namespace cbk
