/* crass_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded CPU restatement of the reference's read-scanning hot path
 * (ctSkennerton/crass 1.0.1: src/crass/libcrispr.cpp, PatternMatcher.cpp, ReadHolder.cpp,
 * SeqUtils.cpp, kseq.cpp, src/aho-corasick/acism.c, and WorkHorse::createNonRedundantSet).
 * It exists to check the CUDA product path; nothing in crass_b200/ may include, link or
 * call it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / the CPU arm.
 *
 * Parity is PINNED: tests/test_oracle_vs_ref.py fuzzes every function here against the
 * compiled, unmodified reference (oracle/_ref/libcrass_ref.so, built by oracle/Makefile),
 * tests/test_oracle_golden.py checks it against the reference's own Catch vectors
 * (src/test/test_libcrispr.cpp, restated in tests/golden/catch_vectors.json) and against
 * dumps produced by the reference on its bundled read sets (tests/golden/ *.dump.gz).
 */
#ifndef CRASS_ORACLE_H
#define CRASS_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint32_t low_dr;       /* options.lowDRsize          default 23 */
    uint32_t high_dr;      /* options.highDRsize         default 47 */
    uint32_t low_spacer;   /* options.lowSpacerSize      default 26 */
    uint32_t high_spacer;  /* options.highSpacerSize     default 50 */
    uint32_t window;       /* options.searchWindowLength default 8  */
    uint32_t min_repeats;  /* options.minNumRepeats      default 2  */
    uint32_t kmer_clust;   /* options.kmer_clust_size    default 6  */
} orc_params;

/* leftmost p in [b, e-w] with s[p,p+w)==pat, else -1   (== bmpSearch, PatternMatcher.cpp:26-59) */
int orc_find_left(const uint8_t* s, uint32_t b, uint32_t e, const uint8_t* pat, uint32_t w);
/* libcrispr.cpp:170-263 */
void orc_scan_right(const uint8_t* s, uint32_t L, uint32_t* ss, uint32_t* n_ss, uint32_t cap,
                    const uint8_t* pat, uint32_t w, uint32_t min_spacer, uint32_t scan_range);
/* libcrispr.cpp:520-772 ; returns the repeat length, rewrites ss in place */
uint32_t orc_extend_pre_repeat(const uint8_t* s, uint32_t L, uint32_t* ss, uint32_t n_ss, int window, int min_spacer);
/* PatternMatcher.cpp:111-195 (OSA distance with the i>2&&j>2 transposition guard) */
int orc_edit_distance(const uint8_t* a, uint32_t la, const uint8_t* b, uint32_t lb);
/* PatternMatcher.cpp:197-204 */
float orc_similarity(const uint8_t* a, uint32_t la, const uint8_t* b, uint32_t lb);
/* libcrispr.cpp:1031-1069 */
int orc_low_complexity(const uint8_t* a, uint32_t la);
/* libcrispr.cpp:869-1029 */
int orc_qc_found_repeats(const uint8_t* s, uint32_t L, const uint32_t* ss, uint32_t n_ss, int min_spacer, int max_spacer);
/* libcrispr.cpp:265-395 ; returns 1/0, -2 when ss_cap is too small */
int orc_search_core(const uint8_t* s, uint32_t L, const orc_params* p, uint32_t* ss, uint32_t ss_cap,
                    uint32_t* n_ss, uint32_t* replen);
/* SeqUtils.cpp:51-87 */
void orc_revcomp(const uint8_t* in, uint32_t len, uint8_t* out);
/* ReadHolder.cpp:513-610,321-380 ; seq and ss are rewritten when the read is flipped */
int orc_dr_lowlexi(uint8_t* seq, uint32_t L, uint32_t* ss, uint32_t n_ss, uint8_t* dr_out, uint32_t* dr_len, int* was_lowlexi);

/* Aho-Corasick with the reference's "first callback wins" use (acism.c:26-104, libcrispr.cpp:408-442) */
typedef struct orc_ac orc_ac;
orc_ac* orc_ac_create(const uint8_t* const* pats, const uint32_t* lens, uint32_t n);
int orc_ac_first_match(const orc_ac* ac, const uint8_t* text, uint32_t tlen, int* end, int* plen);
void orc_ac_destroy(orc_ac* ac);

/* kseq.cpp:171-225 record stream as searchFile sees it; same text format as ref_kseq_dump */
char* orc_kseq_dump(const char* path);
/* WorkHorse.cpp:612-709,1404-1637 ; same text format as ref_non_redundant */
char* orc_non_redundant(const uint8_t* const* drs, const uint32_t* lens, uint32_t n, int min_count);
/* searchFile* -> createNonRedundantSet -> findSingletons*, "crass-dump v1" text (see tests/dumpfmt.py) */
char* orc_run_files(const char* const* paths, uint32_t n_paths, const orc_params* p, int phases, double* timings_ms);
/* the two per-read hot loops alone over an in-memory batch (for the CPU baseline timing);
 * returns number of reads found.  found[] (n_reads bytes) may be NULL. */
uint64_t orc_phase1_batch(const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads, const orc_params* p, uint8_t* found);
uint64_t orc_phase2_batch(const orc_ac* ac, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads, uint8_t* found);
/* smithWaterman (SmithWaterman.cpp:151-308) and ReadHolder::updateStartStops (ReadHolder.cpp:382-511) */
int orc_smith_waterman(const uint8_t* a, uint32_t la, const uint8_t* b, uint32_t lb, int a_start_search, int a_search_len,
                       double similarity, int* a_start_align, int* a_end_align,
                       uint32_t* a_ret_pos, uint32_t* a_ret_len, uint32_t* b_ret_pos, uint32_t* b_ret_len);
int orc_update_start_stops(const uint8_t* seq, uint32_t L, uint32_t* ss, uint32_t* n_ss, uint32_t cap, int front_offset,
                           const uint8_t* dr, uint32_t dr_len, uint32_t low_spacer);
/* Consensus DR of a group (second half).
 * orc_ksw_align: ksw_align (ksw.c:330-354) on nt4 codes with the Aligner's scores (match 1, mismatch -3, ambiguous 0, gap open 5,
 * extend 2; Aligner.h:105-131), 16-bit form: a lane-by-lane restatement of the striped SSE2 kernel ksw_i16 (ksw.c:219-322), whose
 * lazy-F loop and saturating arithmetic decide the corner cases.  out7 = score, te, qe, score2, te2, tb, qb.
 * orc_consensus_group: Aligner::setMasterDR / alignSlave / generateConsensus (Aligner.cpp:72-246) with getOffsetAgainstMaster
 * (:263-362), placeReadsInCoverageArray (:364-417), extendSlaveDR (:420-452), calculateDRZone (:456-484); same interface as
 * ref_consensus_group (oracle/refshim/ref_shim.cpp). */
int orc_ksw_align(const uint8_t* query, int qlen, const uint8_t* target, int tlen, int xtra, int* out7);
int orc_consensus_group(const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads, const uint32_t* read_dr,
                        const uint32_t* ss_offsets, const uint32_t* ss_pool, const uint8_t* dr_bytes, const uint32_t* dr_offsets,
                        uint32_t n_drs, uint32_t array_len, int32_t* dr_place, uint8_t* dr_flags, int32_t* zone,
                        uint8_t* consensus, float* conservation, int32_t* coverage);
void orc_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
