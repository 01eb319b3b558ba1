// hostsim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the product's __host__ __device__ per-read logic (crass_b200/csrc/*.cuh) with g++ so
// that the exact code the CUDA kernels execute can be fuzzed against the parity oracle on a box
// without a GPU.  This library is loaded by tests/ only; the product (libcrass_b200.so) has no
// CPU execution path and does not link or load it.
#include <stdint.h>
#include <string.h>
#include <vector>

#include "../../crass_b200/csrc/dr_core.cuh"
#include "../../crass_b200/csrc/dr_filter.cuh"
#include "../../crass_b200/csrc/sw_core.cuh"

using namespace cb;

static Params to_params(const uint32_t* p) {
    Params o;
    o.low_dr = p[0]; o.high_dr = p[1]; o.low_spacer = p[2]; o.high_spacer = p[3];
    o.window = p[4]; o.min_repeats = p[5]; o.kmer_clust = p[6]; o.scan_range = 24;
    return o;
}

extern "C" {

int hs_search_core(const uint8_t* seq, uint32_t len, const uint32_t* params, uint32_t* ss, uint32_t ss_cap,
                   uint32_t* n_ss, uint32_t* replen) {
    Params o = to_params(params);
    PtrSeq s{seq};
    uint32_t cap = ss_capacity(o, len);
    if (cap > ss_cap) return -2;
    uint32_t n = 0, rl = 0;
    int r = search_core(s, len, o, ss, cap, n, rl);
    *n_ss = n; *replen = rl;
    return r;
}

void hs_scan_right(const uint8_t* seq, uint32_t len, uint32_t* ss, uint32_t* n_ss, uint32_t cap,
                   const uint8_t* pat, uint32_t w, uint32_t min_spacer, uint32_t scan_range) {
    std::vector<uint8_t> buf(seq, seq + len);          // pattern appended behind the read, as the KAT entry does
    buf.insert(buf.end(), pat, pat + w);
    PtrSeq s{buf.data()};
    uint32_t n = *n_ss;
    scan_right(s, len, ss, n, cap, len, w, min_spacer, scan_range);
    *n_ss = n;
}

uint32_t hs_extend_pre_repeat(const uint8_t* seq, uint32_t len, uint32_t* ss, uint32_t n_ss, int window, int min_spacer) {
    PtrSeq s{seq};
    return extend_pre_repeat(s, len, ss, n_ss, (uint32_t)window, (uint32_t)min_spacer);
}

int hs_qc_found_repeats(const uint8_t* seq, uint32_t len, const uint32_t* ss, uint32_t n_ss, int min_spacer, int max_spacer) {
    PtrSeq s{seq};
    return qc_found_repeats(s, len, ss, n_ss, min_spacer, max_spacer);
}

int hs_edit_distance(const uint8_t* a, uint32_t la, const uint8_t* b, uint32_t lb) {
    std::vector<uint8_t> buf(a, a + la);
    buf.insert(buf.end(), b, b + lb);
    buf.push_back(0);
    PtrSeq s{buf.data()};
    return edit_distance(s, 0, la, la, lb);
}

float hs_similarity(const uint8_t* a, uint32_t la, const uint8_t* b, uint32_t lb) {
    std::vector<uint8_t> buf(a, a + la);
    buf.insert(buf.end(), b, b + lb);
    buf.push_back(0);
    PtrSeq s{buf.data()};
    return similarity(s, 0, la, la, lb);
}

int hs_low_complexity(const uint8_t* a, uint32_t la) {
    PtrSeq s{a};
    return low_complexity(s, 0, la) ? 1 : 0;
}

}  // extern "C"

// 2-bit seed flags exactly as the K1 filter kernel computes them: pack 16 bytes per word (zero padded), realign
// to an arbitrary base offset `shift_bases` (as a read inside a tile is), then seed_flags<NW,NWIN,49,97>.
// mode 0: returns the filter decision; mode 1: runs search_core_packed (the exact candidate kernel's logic) and
// returns found, filling ss / n_ss / replen; mode 2: the same through the staged form (seed cut at the edit distance).
template <int NW, int NWIN>
static int run_packed(const uint32_t* R, const uint8_t* seq, uint32_t len, int mode, uint32_t* ss, uint32_t* n_ss, uint32_t* replen) {
    uint32_t acc[NWIN];
    seed_flags<NW, NWIN, 49, 97>(R, acc);
    const bool cand = any_flag<NWIN>(acc);
    if (mode == 0) return cand ? 1 : 0;
    *n_ss = 0; *replen = 0;
    if (!cand) return 0;
    Params o; o.low_dr = 23; o.high_dr = 47; o.low_spacer = 26; o.high_spacer = 50; o.window = 8; o.min_repeats = 2; o.kmer_clust = 6; o.scan_range = 24;
    PtrSeq s{seq};
    uint32_t S[NW + 4];
    for (int k = 0; k < NW + 2; ++k) S[k] = R[k];
    S[NW + 2] = S[NW + 3] = 0;
    uint32_t n = 0, rl = 0;
    const int r = mode == 2 ? search_core_staged<NW, NWIN, 49, 97>(s, len, o, S, flag_mask<NWIN>(acc), ss, 32, n, rl)
                            : search_core_packed<NW, NWIN, 49, 97>(s, len, o, S, flag_mask<NWIN>(acc), ss, 32, n, rl);
    *n_ss = n; *replen = rl;
    return r;
}

extern "C" {

int hs_packed(const uint8_t* seq, uint32_t len, uint32_t shift_bases, int nw, const uint8_t* tail, uint32_t tail_len, int mode,
              uint32_t* ss, uint32_t* n_ss, uint32_t* replen) {
    std::vector<uint8_t> buf(shift_bases, (uint8_t)'G');
    buf.insert(buf.end(), seq, seq + len);
    buf.insert(buf.end(), tail, tail + tail_len);       // what follows the read in the batch
    buf.resize(buf.size() + 16 * 40, 0);
    std::vector<uint32_t> packed(buf.size() / 16);
    for (size_t v = 0; v < packed.size(); ++v) {
        uint32_t w[4];
        memcpy(w, buf.data() + 16 * v, 16);
        packed[v] = pack16(w[0], w[1], w[2], w[3]);
    }
    uint32_t R[40];
    const uint32_t wi = shift_bases >> 4, sh = (shift_bases & 15) * 2;
    for (int k = 0; k < nw + 2; ++k) R[k] = funnel_r(packed[wi + k], packed[wi + k + 1], sh);
    const int se = (int)len - 58;
    const int nwin = se < 0 ? 1 : se / 16 + 1;
#define HS_RUN(NW, NWIN) return run_packed<NW, NWIN>(R, seq, len, mode, ss, n_ss, replen)
    if (nw == 7 && nwin <= 3) HS_RUN(7, 3);
    if (nw == 7) HS_RUN(7, 4);
    if (nw == 10 && nwin <= 6) HS_RUN(10, 6);
    if (nw == 10) HS_RUN(10, 7);
    if (nw == 16 && nwin <= 12) HS_RUN(16, 12);
    if (nw == 16) HS_RUN(16, 13);
    if (nw == 19 && nwin <= 15) HS_RUN(19, 15);
    if (nw == 19) HS_RUN(19, 16);
#undef HS_RUN
    return -9;
}

}  // extern "C"

// partial-DR recovery (sw_core.cuh): the device code of k_update_start_stops, one read per call
extern "C" {

int hs_smith_waterman(const uint8_t* a, uint32_t la, const uint8_t* b, uint32_t lb, int start, int len, double similarity,
                      int* start_align, int* end_align, uint32_t* a_pos, uint32_t* a_len, uint32_t* b_pos, uint32_t* b_len) {
    if (lb > (uint32_t)kMaxSwDr) return -1;
    PtrSeq s{a};
    SwResult r;
    smith_waterman(s, la, b, lb, start, len, similarity, r);
    *start_align = r.start_align; *end_align = r.end_align;
    *a_pos = r.a_pos; *a_len = r.a_len; *b_pos = r.b_pos; *b_len = r.b_len;
    return r.ok;
}

// out must hold *n_ss + 4 entries; returns the UssStatus
int hs_update_start_stops(const uint8_t* seq, uint32_t len, const uint32_t* ss, uint32_t n_ss, int front_offset,
                          const uint8_t* dr, uint32_t dr_len, uint32_t low_spacer, uint32_t* out, uint32_t* n_out) {
    PtrSeq s{seq}, d{dr};
    uint32_t n = 0;
    const uint8_t st = update_start_stops(s, len, ss, n_ss, front_offset, d, dr_len, low_spacer, out, n);
    *n_out = n;
    return (int)st;
}

}  // extern "C"
