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
    return osa_distance(s, 0, la, la, lb);
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

// 2-bit seed pre-filter exactly as the K1 filter kernel runs it: pack 16 bytes per word (zero padded), realign
// to an arbitrary base offset `shift_bases` (as a read inside a tile is), then seed_filter<NW,NWIN,49,97>.
int hs_seed_filter(const uint8_t* seq, uint32_t len, uint32_t shift_bases, int nw, const uint8_t* tail, uint32_t tail_len) {
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
    if (nw == 7 && nwin <= 3) return seed_filter<7, 3, 49, 97>(R);
    if (nw == 7) return seed_filter<7, 4, 49, 97>(R);
    if (nw == 10 && nwin <= 6) return seed_filter<10, 6, 49, 97>(R);
    if (nw == 10) return seed_filter<10, 7, 49, 97>(R);
    if (nw == 16 && nwin <= 12) return seed_filter<16, 12, 49, 97>(R);
    if (nw == 16) return seed_filter<16, 13, 49, 97>(R);
    if (nw == 19 && nwin <= 15) return seed_filter<19, 15, 49, 97>(R);
    if (nw == 19) return seed_filter<19, 16, 49, 97>(R);
    return -1;
}

}  // extern "C"
