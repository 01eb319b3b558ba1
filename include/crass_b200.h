/* crass_b200.h -- C-ABI of the B200-native read-scanning hot path of crass.
 *
 * This is the drop-in boundary: plain pointers and sizes, no STL, no torch types.  The C++
 * shim that a crass maintainer links instead of src/crass/libcrispr.cpp (see INTEGRATION.md,
 * crass_b200/csrc/dropin/libcrispr_b200.cpp) is the only caller the reference needs; the Python
 * package (crass_b200/) binds the same symbols through ctypes for tests and bench.py.
 *
 * Reference interfaces replaced (all /root/reference/src/crass/...):
 *   searchFile()      libcrispr.h:74-80  / libcrispr.cpp:68-166    -> crass_b200_parse_file + crass_b200_dr_search
 *   searchCore()      libcrispr.h:82-84  / libcrispr.cpp:265-395   -> crass_b200_dr_search[_dev]      (kernel K1)
 *   scanRight()       libcrispr.h:94-97  / libcrispr.cpp:170-263   -> inside K1; crass_b200_scan_right (KAT entry)
 *   extendPreRepeat() libcrispr.h:99-101 / libcrispr.cpp:520-772   -> inside K1; crass_b200_extend_pre_repeat (KAT entry)
 *   qcFoundRepeats()  libcrispr.h:113-115/ libcrispr.cpp:869-1029  -> inside K1
 *   PatternMatcher::levenstheinDistance / getStringSimilarity  PatternMatcher.cpp:111-204 -> crass_b200_edit_distance_batch (K3)
 *   findSingletons()  libcrispr.h:86-92  / libcrispr.cpp:444-518   -> crass_b200_ac_build + crass_b200_ac_scan[_dev] (kernel K2)
 *   acism_create/acism_scan  src/aho-corasick/acism.h:34,52-57     -> crass_b200_ac_build / crass_b200_ac_scan
 *   addReadHolder() + ReadHolder::DRLowLexi  libcrispr.cpp:1119-1162, ReadHolder.cpp:513-610 -> crass_b200_results_* (host replay)
 *   WorkHorse::createNonRedundantSet  WorkHorse.cpp:612-709,1404-1637 -> crass_b200_non_redundant_set
 *   kseq_read()       kseq.cpp:171-225                              -> crass_b200_parse_file
 *   ReadHolder::updateStartStops + smithWaterman  ReadHolder.cpp:382-511, SmithWaterman.cpp:151-308
 *                                                                   -> crass_b200_update_start_stops[_dev] (kernel K6)
 *
 * Conventions
 *   - every function returns 0 on success or a negative crass_b200_status; crass_b200_last_error()
 *     gives the message of the last failure on the calling thread.
 *   - "_dev" entry points take DEVICE pointers and a CUDA stream (void* = cudaStream_t, NULL = default
 *     stream) and only enqueue work; the others take HOST pointers and do the H2D / D2H copies.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     CRASS_B200_ENODEVICE.
 *   - a read batch is byte-packed: bases[] holds the reads back to back exactly as kseq delivers them
 *     (any byte 33..126, no case folding), offsets[n_reads+1] are uint64 byte offsets (offsets[0]==0).
 */
#ifndef CRASS_B200_H
#define CRASS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRASS_B200_ABI_VERSION 1

typedef enum {
    CRASS_B200_OK = 0,
    CRASS_B200_EINVAL = -1,      /* bad argument / unsupported parameter combination */
    CRASS_B200_ENODEVICE = -2,   /* no CUDA device / driver */
    CRASS_B200_ECUDA = -3,       /* a CUDA call failed (message has the CUDA error string) */
    CRASS_B200_ENOMEM = -4,
    CRASS_B200_EIO = -5,         /* cannot open / read input file */
    CRASS_B200_EOVERFLOW = -6    /* caller-provided output capacity too small (needed size is reported) */
} crass_b200_status;

/* the searched fields of the reference's `options` struct (crassDefines.h:140-170) */
typedef struct {
    uint32_t low_dr;        /* lowDRsize           (23) */
    uint32_t high_dr;       /* highDRsize          (47) */
    uint32_t low_spacer;    /* lowSpacerSize       (26) */
    uint32_t high_spacer;   /* highSpacerSize      (50) */
    uint32_t window;        /* searchWindowLength  (8), CLI clamps it to 6..9 */
    uint32_t min_repeats;   /* minNumRepeats       (2) */
    uint32_t kmer_clust;    /* kmer_clust_size     (6) */
    uint32_t scan_range;    /* hard-coded 24 in libcrispr.cpp:347 */
} crass_b200_params;

void crass_b200_default_params(crass_b200_params* p);

/* one found read of phase 1 (searchCore returned true) or phase 2 (first automaton match) */
typedef struct {
    uint32_t read_index;    /* index into the batch */
    uint32_t n_ss;          /* number of start/stop entries (2 per repeat) */
    uint32_t ss_offset;     /* first entry in the start/stop pool */
    uint32_t repeat_len;    /* RH_RepeatLength (0 for phase-2 hits, as in the reference) */
} crass_b200_hit;

typedef struct crass_b200_ctx crass_b200_ctx;     /* one per GPU: stream, staging, workspaces */
typedef struct crass_b200_ac crass_b200_ac;       /* compiled multi-pattern automaton */
typedef struct crass_b200_batch crass_b200_batch; /* parsed reads (host) */
typedef struct crass_b200_results crass_b200_results; /* ReadMap / StringCheck / lookupTable mirror */

const char* crass_b200_last_error(void);
int crass_b200_abi_version(void);
const char* crass_b200_build_info(void);
int crass_b200_device_count(void);

int crass_b200_ctx_create(int device, crass_b200_ctx** out);
void crass_b200_ctx_destroy(crass_b200_ctx* ctx);
int crass_b200_ctx_device(const crass_b200_ctx* ctx);
/* number of kernel launches issued through this context so far (bench.py's gpu_launches) */
uint64_t crass_b200_ctx_launch_count(const crass_b200_ctx* ctx);
/* number of reads the most recent host-form search/scan sent to the exact (candidate) path; equals the number
 * of reads when the generic kernels ran */
uint64_t crass_b200_ctx_last_candidates(const crass_b200_ctx* ctx);
/* K4: make the following crass_b200_dr_search_dev launches also write, for the hit stored in slot k of d_hits, its
 * low-lexi DR token (ReadHolder::DRLowLexi) to d_tokens + k*stride: byte 0 = length, byte 1 = 1 if the read keeps its
 * orientation, bytes 2.. = the token.  stride >= high_dr + 2.  NULL switches it off. */
int crass_b200_ctx_set_token_output(crass_b200_ctx* ctx, void* d_tokens, uint32_t stride);
/* Reuse of phase 1's work in phase 2 for the *_dev calls.  The direct-repeat filter recodes every base of the batch to
 * 2 bits; with on != 0 the next crass_b200_dr_search_dev launches leave that stream in HBM (n_reads * max_read_len / 4
 * bytes, owned by ctx) and crass_b200_ac_scan_dev launches on the SAME d_bases pointer and n_reads read it instead of
 * the bytes (a quarter of the traffic, no recoding).  The caller promises that the bases are not modified between the
 * two launches.  Results are identical either way.  The resident host-buffer calls do this on their own. */
int crass_b200_ctx_keep_packed(crass_b200_ctx* ctx, int on);
/* the same with the number of bases of the batches that follow (0 = off): the stream is then sized exactly, which matters for
 * batches of mixed lengths (n_reads * max_read_len is far too much when one read in millions is long) */
int crass_b200_ctx_keep_packed_bases(crass_b200_ctx* ctx, uint64_t n_bases);
/* the distinct tokens of the most recent crass_b200_dr_search_resident in read order, '\n'-separated (owned by ctx) */
const char* crass_b200_ctx_last_dr_list(const crass_b200_ctx* ctx);
/* the *_dev entry points leave hit records in device slot order; this puts a host copy into read order (what the
 * host-buffer calls return and replay requires), in place */
void crass_b200_sort_hits(crass_b200_hit* hits, uint32_t n_hits);
/* the same on the device, before the copy: d_found are the flags the search launch wrote (16-byte aligned; exactly the
 * reads with a non-zero flag have a hit record, at most one each), *d_n_hits the hit counter on the device (counters[0]),
 * max_hits the capacity of d_sorted.  A hit's place in read order is the number of flagged reads before it. */
int crass_b200_sort_hits_dev(crass_b200_ctx* ctx, const uint8_t* d_found, uint32_t n_reads, const crass_b200_hit* d_hits,
                             const uint32_t* d_n_hits, uint32_t max_hits, crass_b200_hit* d_sorted, void* stream);
/* K4b: de-duplicate the token records of d_hits[0..n_hits) on the device.  Writes the distinct records to d_out_tokens
 * (same stride), the smallest read index carrying each to d_out_first_read, and their number to d_out_count; all three
 * must hold n_hits entries in the worst case.  Order is arbitrary: sort by first_read for first-appearance order
 * (crass_b200_dr_list_from_unique does that on the host). */
int crass_b200_unique_tokens_dev(crass_b200_ctx* ctx, const crass_b200_hit* d_hits, uint32_t n_hits, const void* d_tokens,
                                 uint32_t stride, void* d_out_tokens, uint32_t* d_out_first_read, uint32_t* d_out_count, void* stream);
char* crass_b200_dr_list_from_unique(const uint8_t* records, uint32_t stride, const uint32_t* first_read, uint32_t n);
/* K4b/K4c in block form -- the unit of the multi-GPU exchange (stands in for the single StringCheck
 * that numbers DR tokens by first appearance, StringCheck.cpp:46-55, libcrispr.cpp:1119-1162).
 * A token block is 16 header bytes {u32 count, u32 flags, 8 spare} followed by cap records of stride bytes (stride a
 * multiple of 4, >= high_dr + 6): byte 0 = token length, byte 1 = orientation, bytes 2.. = token, last 4 bytes = order
 * key (read index of first appearance).  count may exceed cap; the block then holds cap of the records and the caller
 * retries with a larger block.  flags bit 0: a gathered block had overflowed; bit 1: a token did not fit its record.
 *   unique_tokens_block : distinct tokens of d_hits[0..n_hits) -> d_block (order key = smallest read index)
 *   merge_token_blocks  : d_blocks = n_ranks blocks of the same geometry back to back in rank order (what one
 *                         all-gather delivers); shards are contiguous read ranges of at most shard_reads reads, so the
 *                         merged order key rank*shard_reads + key reproduces the first-appearance order of one
 *                         sequential run.  Output: one block of out_cap records.
 *   dr_list_from_block  : host copy of a block -> '\n'-separated DR list in order-key order (malloc'd) */
size_t crass_b200_token_block_bytes(uint32_t cap, uint32_t stride);
int crass_b200_unique_tokens_block_dev(crass_b200_ctx* ctx, const crass_b200_hit* d_hits, uint32_t n_hits, const void* d_tokens,
                                       uint32_t stride, void* d_block, uint32_t cap, void* stream);
int crass_b200_merge_token_blocks_dev(crass_b200_ctx* ctx, const void* d_blocks, uint32_t n_ranks, uint32_t cap, uint32_t stride,
                                      uint32_t shard_reads, void* d_out_block, uint32_t out_cap, void* stream);
char* crass_b200_dr_list_from_block(const void* block, uint32_t cap, uint32_t stride, uint32_t* count, uint32_t* flags);
/* the same list from token records copied back by the caller: records[k] belongs to hits[k] (unsorted, as on the device) */
char* crass_b200_dr_list_from_tokens(const uint8_t* records, uint32_t stride, const crass_b200_hit* hits, uint32_t n_hits);

/* ---- phase 1: direct-repeat search (kernel K1) -------------------------------------------------
 * Device-resident form.  Outputs (all device memory, caller-allocated):
 *   d_found[n_reads]            1 if searchCore would return true for the read, else 0
 *   d_hits[hits_cap], d_ss_pool[ss_cap]   hit records in arbitrary order + their start/stop entries
 *   d_counters[4]               [0]=number of hits, [1]=number of pool entries used,
 *                               [2]=1 if a capacity was exceeded (hits beyond capacity are dropped
 *                                   but still counted, so the caller can re-run with larger buffers)
 *                               [3]=number of reads that took the exact (candidate) path
 * The counters are zeroed by the call (on the stream). */
int crass_b200_dr_search_dev(crass_b200_ctx* ctx, const uint8_t* d_bases, const uint64_t* d_offsets,
                             uint32_t n_reads, uint32_t max_read_len, const crass_b200_params* params,
                             uint8_t* d_found, crass_b200_hit* d_hits, uint32_t hits_cap,
                             uint32_t* d_ss_pool, uint32_t ss_cap, uint32_t* d_counters, void* stream);

/* Host form: copies the batch in (one asynchronous copy on the context's stream; the copy is 98 % of this call, so the
 * multi-device engine below is where transfers are sliced over two copy streams), runs K1, copies the hits out,
 * sorted by read_index.  hits/ss_pool are malloc'd by the library (free with crass_b200_free).
 * found may be NULL. */
int crass_b200_dr_search(crass_b200_ctx* ctx, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads,
                         const crass_b200_params* params, uint8_t* found,
                         crass_b200_hit** hits, uint32_t* n_hits, uint32_t** ss_pool, uint32_t* n_ss_pool);

/* Resident form: upload a host batch once, then run both phases on it without copying it again.
 *   crass_b200_batch_upload       H2D of bases+offsets into context-owned device buffers (replaces the previous one)
 *   crass_b200_dr_search_resident K1 on the resident batch; remembers the found flags on the device
 *   crass_b200_ac_scan_resident   K2 on the resident batch; skip_found != 0 skips the reads phase 1 flagged
 * Output conventions as in crass_b200_dr_search. */
int crass_b200_batch_upload(crass_b200_ctx* ctx, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads);
int crass_b200_dr_search_resident(crass_b200_ctx* ctx, const crass_b200_params* params, uint8_t* found,
                                  crass_b200_hit** hits, uint32_t* n_hits, uint32_t** ss_pool, uint32_t* n_ss_pool);
int crass_b200_ac_scan_resident(crass_b200_ctx* ctx, const crass_b200_ac* ac, int skip_found, uint8_t* found,
                                crass_b200_hit** hits, uint32_t* n_hits, uint32_t** ss_pool, uint32_t* n_ss_pool);

/* ---- phase 2: singleton scan (kernel K2) ---------------------------------------------------------
 * patterns: n_patterns byte strings back to back, pat_offsets[n_patterns+1]. */
int crass_b200_ac_build(const uint8_t* pat_bytes, const uint32_t* pat_offsets, uint32_t n_patterns, crass_b200_ac** out);
void crass_b200_ac_destroy(crass_b200_ac* ac);
/* copy the automaton into the context's device buffers now (otherwise the first scan does it) */
int crass_b200_ac_upload(crass_b200_ctx* ctx, const crass_b200_ac* ac);
uint32_t crass_b200_ac_num_states(const crass_b200_ac* ac);
uint32_t crass_b200_ac_num_symbols(const crass_b200_ac* ac);
uint64_t crass_b200_ac_table_bytes(const crass_b200_ac* ac);
/* the patterns the matcher was built from, '\n'-separated in build order (malloc'd, crass_b200_free) */
char* crass_b200_ac_pattern_text(const crass_b200_ac* ac, uint32_t* n_patterns);

/* d_skip[n_reads] may be NULL; reads with d_skip[i]!=0 are not scanned (the readsFound test of
 * on_match, libcrispr.cpp:411, for reads already found in phase 1).  For every scanned read with a
 * match a hit {read_index, n_ss=2, repeat_len=0} is appended and ss_pool gets (start, end) exactly as
 * on_match computes them (libcrispr.cpp:420-437).  d_found[n_reads] (may be NULL) gets 1/0. */
int crass_b200_ac_scan_dev(crass_b200_ctx* ctx, const crass_b200_ac* ac, const uint8_t* d_bases, const uint64_t* d_offsets,
                           uint32_t n_reads, uint32_t max_read_len, const uint8_t* d_skip, uint8_t* d_found,
                           crass_b200_hit* d_hits, uint32_t hits_cap, uint32_t* d_ss_pool, uint32_t ss_cap,
                           uint32_t* d_counters, void* stream);

int crass_b200_ac_scan(crass_b200_ctx* ctx, const crass_b200_ac* ac, const uint8_t* bases, const uint64_t* offsets,
                       uint32_t n_reads, const uint8_t* skip, uint8_t* found,
                       crass_b200_hit** hits, uint32_t* n_hits, uint32_t** ss_pool, uint32_t* n_ss_pool);

/* ---- K3: batched modified edit distance (PatternMatcher::levenstheinDistance) --------------------
 * pair i compares a = bytes[a_off[i] .. a_off[i]+a_len[i]) with b likewise; all HOST pointers.
 * out_dist[n_pairs] int32, out_sim[n_pairs] float (getStringSimilarity, bit-exact). */
int crass_b200_edit_distance_batch(crass_b200_ctx* ctx, const uint8_t* bytes, uint64_t n_bytes,
                                   const uint32_t* a_off, const uint32_t* a_len, const uint32_t* b_off, const uint32_t* b_len,
                                   uint32_t n_pairs, int32_t* out_dist, float* out_sim);

/* ---- K6: partial-DR recovery, batched (first consumer of the path's start/stop lists) -------------
 * ReadHolder::updateStartStops (ReadHolder.cpp:382-511) with the 7-argument smithWaterman (SmithWaterman.cpp:151-308,
 * similarity cut-off CRASS_DEF_PARTIAL_SIM_CUT_OFF 0.85, minimum length CRASS_DEF_MIN_PARTIAL_LENGTH 4).  WorkHorse
 * calls it once per read of a DR group with the group's consensus DR (WorkHorse.cpp:1347); here one call takes the
 * jobs of any number of groups.  Job i rewrites the read's list ss_in[ss_offset, ss_offset + n_ss) into
 * ss_out[out_offset, ...) -- reserve n_ss + 4 entries: a partial repeat may be added at either end -- and reports
 * n_out[i] entries and status[i]: 0 ok; 1 list empty or odd; 2 DR empty or longer than 127; 3 a shifted start lies
 * at or past the end of the read (the reference logs "Something wrong with front offset!" and reads out of bounds);
 * 4 read of 65536 bases or more.  For status != 0 nothing is written and n_out[i] = 0. */
typedef struct crass_b200_uss_job {
    uint32_t read;          /* index of the read in offsets */
    uint32_t ss_offset;     /* first entry of its start/stop list in ss_in */
    uint32_t n_ss;          /* number of entries (even, >= 2) */
    int32_t  front_offset;  /* dr_aligner.offset(token) - dr_aligner.getDRZoneStart(), may be negative */
    uint32_t dr;            /* index of the group's consensus DR in dr_offsets */
    uint32_t out_offset;    /* first entry of the new list in ss_out */
} crass_b200_uss_job;

/* device pointers + stream: only enqueues */
int crass_b200_update_start_stops_dev(crass_b200_ctx* ctx, const uint8_t* d_bases, const uint64_t* d_offsets,
                                      const uint8_t* d_dr_bytes, const uint32_t* d_dr_offsets,
                                      const crass_b200_uss_job* d_jobs, uint32_t n_jobs, const uint32_t* d_ss_in,
                                      uint32_t low_spacer, uint32_t* d_ss_out, uint32_t* d_n_out, uint8_t* d_status, void* stream);

/* host pointers: validates the jobs, copies, runs, copies back, synchronises */
int crass_b200_update_start_stops(crass_b200_ctx* ctx, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads,
                                  const uint8_t* dr_bytes, const uint32_t* dr_offsets, uint32_t n_drs,
                                  const crass_b200_uss_job* jobs, uint32_t n_jobs, const uint32_t* ss_in, uint32_t n_ss_in,
                                  uint32_t low_spacer, uint32_t* ss_out, uint32_t ss_out_cap, uint32_t* n_out, uint8_t* status);

/* ---- known-answer entry points for the two functions the reference's own unit tests pin ----------
 * (src/test/test_libcrispr.cpp).  They run the SAME device code K1 uses, one read per call. */
int crass_b200_scan_right(crass_b200_ctx* ctx, const uint8_t* seq, uint32_t len, uint32_t* ss, uint32_t* n_ss, uint32_t ss_cap,
                          const uint8_t* pattern, uint32_t pattern_len, uint32_t min_spacer, uint32_t scan_range);
int crass_b200_extend_pre_repeat(crass_b200_ctx* ctx, const uint8_t* seq, uint32_t len, uint32_t* ss, uint32_t n_ss,
                                 uint32_t window, uint32_t min_spacer, uint32_t* repeat_len);
/* qcFoundRepeats (libcrispr.h:113-115) on one read: *result = 1 pass, 0 fail, -1 where the reference throws */
int crass_b200_qc_found_repeats(crass_b200_ctx* ctx, const uint8_t* seq, uint32_t len, const uint32_t* ss, uint32_t n_ss,
                                int min_spacer, int max_spacer, int* result);

/* ---- feed path: kseq-compatible FASTA/FASTQ(.gz) parser ------------------------------------------ */
int crass_b200_parse_file(const char* path, crass_b200_batch** out);      /* "-" = stdin */
/* The same record stream handed out range by range -- the streamed feed (the reference's loop over kseq_read holds one record
 * at a time, libcrispr.cpp:96-131; here a range of about range_bytes of the input is one batch, so that parsing, copying,
 * K1 and the replay of successive ranges overlap: crass_b200_engine_run_files does that for a file of two ranges or more and for
 * every run over several files, which go through the pipeline one after the other).
 * Every range ends on a true record start and inherits kseq's stale comment / quality strings from the one before it;
 * crass_b200_parse_stream_next returns 1 and a batch (parse status 0 while more follows, the stream's final status in the
 * last one), 0 when the stream had ended, a negative error code otherwise. */
typedef struct crass_b200_parse_stream crass_b200_parse_stream;
int crass_b200_parse_stream_open(const char* path, uint64_t range_bytes, crass_b200_parse_stream** out);
int crass_b200_parse_stream_next(crass_b200_parse_stream* s, crass_b200_batch** out);
void crass_b200_parse_stream_close(crass_b200_parse_stream* s);
int crass_b200_batch_from_memory(const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads,
                                 const char* const* names /* may be NULL: r%010u */, crass_b200_batch** out);
void crass_b200_batch_destroy(crass_b200_batch* b);
uint32_t crass_b200_batch_num_reads(const crass_b200_batch* b);
uint32_t crass_b200_batch_max_read_len(const crass_b200_batch* b);
int crass_b200_batch_parse_status(const crass_b200_batch* b);             /* last kseq_read return: -1 EOF, -2 truncated */
const uint8_t* crass_b200_batch_bases(const crass_b200_batch* b);          /* all reads back to back (offsets below index it); a range
                                                                              of a streamed file keeps its reads in segments and is
                                                                              copied together on the first call: prefer _batch_read */
const uint8_t* crass_b200_batch_read(const crass_b200_batch* b, uint32_t i, uint32_t* len);   /* seq->seq.s / seq->seq.l of record i */
const uint64_t* crass_b200_batch_offsets(const crass_b200_batch* b);
/* record fields exactly as searchFile sees them (stale comment/qual buffers of kseq included);
 * has_comment / has_qual mirror (seq->comment.s != NULL) / (seq->qual.s != NULL) */
const char* crass_b200_batch_name(const crass_b200_batch* b, uint32_t i);
const char* crass_b200_batch_comment(const crass_b200_batch* b, uint32_t i, int* has_comment);
const char* crass_b200_batch_qual(const crass_b200_batch* b, uint32_t i, int* has_qual);

/* ---- host replay: the containers the reference fills --------------------------------------------- */
int crass_b200_results_create(crass_b200_results** out);
void crass_b200_results_destroy(crass_b200_results* r);
/* replays phase-1 hits in read order: addReadHolder (DRLowLexi, token), patternsHash, readsFound */
int crass_b200_results_add_phase1(crass_b200_results* r, const crass_b200_batch* b,
                                  const crass_b200_hit* hits, uint32_t n_hits, const uint32_t* ss_pool);
/* replays phase-2 hits in read order: on_match (header not in readsFound -> addReadHolder) */
int crass_b200_results_add_phase2(crass_b200_results* r, const crass_b200_batch* b,
                                  const crass_b200_hit* hits, uint32_t n_hits, const uint32_t* ss_pool);
/* the same for several batches at once, in the order given (the ranges of a streamed file): the header test and the holders of
 * all hits are made on the worker threads, only the container inserts walk the hits in order */
int crass_b200_results_add_phase2_ranges(crass_b200_results* r, uint32_t n_batches, const crass_b200_batch* const* batches,
                                         const crass_b200_hit* const* hits, const uint32_t* n_hits, const uint32_t* const* ss_pools);
uint32_t crass_b200_results_num_tokens(const crass_b200_results* r);
uint32_t crass_b200_results_num_reads(const crass_b200_results* r);
/* the distinct low-lexi DR strings in token order (token = index + 2), '\n'-separated; malloc'd */
char* crass_b200_results_dr_list(const crass_b200_results* r);
/* merges DR strings discovered by other shards in front of / behind this one (multi-GPU): the list is
 * the concatenation over ranks in rank order; tokens are renumbered by first appearance */
int crass_b200_results_adopt_tokens(crass_b200_results* r, const char* dr_list_all_ranks);
/* WorkHorse::createNonRedundantSet on the current token set; patterns '\n'-separated; malloc'd */
char* crass_b200_results_non_redundant(crass_b200_results* r, uint32_t kmer_clust, uint32_t* n_patterns);
/* "crass-dump v1" text of the whole state (same format the oracle emits); malloc'd */
char* crass_b200_results_dump(crass_b200_results* r, int max_read_len);

/* the distinct low-lexi DR strings (ReadHolder::DRLowLexi) of a sorted hit list in first-appearance order,
 * '\n'-separated, without building any container: what a shard contributes to the multi-GPU merge; malloc'd */
char* crass_b200_dr_list_from_hits(const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads,
                                   const crass_b200_hit* hits, uint32_t n_hits, const uint32_t* ss_pool);
/* first-appearance de-duplication of the rank-ordered concatenation of such lists; malloc'd */
char* crass_b200_merge_dr_lists(const char* concatenated);

/* clustering step alone on an ordered DR list ('\n'-separated, token order); "G"/"P" lines; malloc'd */
char* crass_b200_non_redundant_set(const char* dr_list, uint32_t kmer_clust);
/* the same clustering followed by crass_b200_ac_build on its pattern list, without the text round trip: the step
 * WorkHorse.cpp:367-379 takes between the phases (createNonRedundantSet, then findSingletons builds its matcher,
 * libcrispr.cpp:455-470).  *n_patterns (optional) = size of the non-redundant set; an empty set is EINVAL. */
int crass_b200_ac_build_from_dr_list(const char* dr_list, uint32_t kmer_clust, crass_b200_ac** out, uint32_t* n_patterns);
/* K5: the same step from a token block that is still on the device (the output of unique_tokens_block_dev or
 * merge_token_blocks_dev).  All of createNonRedundantSet runs as kernels (K5, csrc/cluster.cuh): token order, the canonical
 * key of every 11-mer and the first DR holding it (clusterDRReads' k-mer map, WorkHorse.cpp:1404-1637), the
 * order-dependent group walk (:1542-1625) as a dependency graph, removeRedundantRepeats (:612-645) and the emission of
 * survivors + reverse complements (:690-697); the pattern bytes come back in the one host synchronisation of the call and
 * the matcher's tables are built on the device (k_ac_build) before the call returns, so a scan can follow at once.
 * Lists the kernels decline (a letter with a non-involutive complement such as 'U', more than 32768 variants, ...) take
 * the host passes; CRASS_B200_CLUSTER=device-passes selects the round-1 split (first passes on the device, rest on the host).
 * *count / *flags as in dr_list_from_block; *out stays NULL (return 0) for an overflowed or empty block.
 * The second form returns the pattern set as '\n'-separated text instead (malloc'd). */
int crass_b200_cluster_block_dev(crass_b200_ctx* ctx, const void* d_block, uint32_t cap, uint32_t stride, uint32_t kmer_clust,
                                 crass_b200_ac** out, uint32_t* count, uint32_t* flags, uint32_t* n_patterns, void* stream);
char* crass_b200_cluster_block_patterns_dev(crass_b200_ctx* ctx, const void* d_block, uint32_t cap, uint32_t stride,
                                            uint32_t kmer_clust, uint32_t* count, uint32_t* flags, uint32_t* n_patterns, void* stream);
/* the two halves separately, for drivers that cluster on one rank and hand the pattern set to the others: the
 * non-redundant set as '\n'-separated text in the reference's order (per group: survivors, then their reverse
 * complements; malloc'd), and the matcher built from such a text */
char* crass_b200_non_redundant_patterns(const char* dr_list, uint32_t kmer_clust, uint32_t* n_patterns);
/* both straight from a host copy of a token block (no intermediate text): *count / *flags as in dr_list_from_block.
 * ac_build_from_block leaves *out NULL (and returns 0) when the block had overflowed or holds no token. */
char* crass_b200_non_redundant_patterns_from_block(const void* block, uint32_t cap, uint32_t stride, uint32_t kmer_clust,
                                                   uint32_t* count, uint32_t* flags, uint32_t* n_patterns);
int crass_b200_ac_build_from_block(const void* block, uint32_t cap, uint32_t stride, uint32_t kmer_clust, crass_b200_ac** out,
                                   uint32_t* count, uint32_t* flags, uint32_t* n_patterns);
int crass_b200_ac_build_from_pattern_list(const char* patterns, crass_b200_ac** out, uint32_t* n_patterns);

/* ---- K7: consensus ("true") DR of DR groups (second half) ----------------------------------------------
 * Replaces, for the groups WorkHorse::parseGroupedDRs (WorkHorse.cpp:1135-1171) walks, Aligner::setMasterDR / alignSlave /
 * generateConsensus (Aligner.cpp:72-246) and the ksw_align (ksw.c:330-354) under them.
 *
 * crass_b200_ksw_align: a batch of ksw_align calls with the Aligner's scoring (match 1, mismatch -3, ambiguous 0, gap open 5,
 * extend 2; Aligner.h:105-131), 16-bit kernel.  Sequences are LETTERS in `pool` (nt4-coded on the device as
 * Aligner::prepareSequenceForAlignment does); q_rc != 0 aligns the reverse complement of the query.  All HOST pointers.
 *
 * crass_b200_consensus_groups: n_groups groups at once.  DRs group_first_dr[g] .. group_first_dr[g+1]-1 belong to group g, the
 * first of them is its master (findMasterDR: the longest).  reads: the reads hanging on those DRs, back to back (offsets[n_reads
 * + 1]), read_dr[i] = DR index, start/stop list of read i = ss_pool[ss_offsets[i] .. ss_offsets[i+1]).  array_len =
 * CRASS_DEF_CONS_ARRAY_RL_MULTIPLIER * max read length.  Out: dr_place[n_drs] (AL_Offsets, -1 = the slave could not be placed),
 * dr_flags[n_drs] (bit 0: slave and its reads were reverse complemented, bit 1: alignment failed, bit 2: forward and reverse
 * scores stayed equal), and per group zone[2] (AL_ZoneStart/End after generateConsensus), consensus[array_len],
 * conservation[array_len], coverage[4 * array_len] (rows A C G T).  *status: bit 0 = a read without a full-length repeat was
 * skipped, bit 1 = a read reached past the array (the reference writes out of bounds in both cases).  All HOST pointers. */
typedef struct { uint32_t q_off, q_len, t_off, t_len, q_rc; int32_t xtra; } crass_b200_ksw_job;
typedef struct { int32_t score, te, qe, score2, te2, tb, qb, status; } crass_b200_ksw_result;
int crass_b200_ksw_align(crass_b200_ctx* ctx, const uint8_t* pool, uint64_t pool_bytes, const crass_b200_ksw_job* jobs, uint32_t n_jobs,
                         crass_b200_ksw_result* out);
int crass_b200_consensus_groups(crass_b200_ctx* ctx, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads,
                                const uint32_t* read_dr, const uint32_t* ss_offsets, const uint32_t* ss_pool,
                                const uint8_t* dr_bytes, const uint32_t* dr_offsets, uint32_t n_drs,
                                const uint32_t* group_first_dr, uint32_t n_groups, uint32_t array_len,
                                int32_t* dr_place, uint8_t* dr_flags, int32_t* zone, uint8_t* consensus, float* conservation,
                                int32_t* coverage, uint32_t* status);

/* ---- one process per GPU: the exchange of SURVEY.md 8e with the collective enqueued by the library -----------------------
 * crass_b200_comm_unique_id: an NCCL unique id (libnccl is dlopen'ed); rank 0 makes one and hands it to the other ranks by
 * whatever means the driver has (a file, MPI, torch.distributed).  crass_b200_ctx_comm_init: ncclCommInitRank for this
 * context's device (collective: every rank calls it).  crass_b200_exchange_tokens_dev: K4b on this rank's hits, ONE
 * ncclAllGather of the fixed-size token blocks, K4c over the gathered blocks -- three enqueues on the caller's stream, no host
 * synchronisation, no other stream; d_merged then holds the merged block of all ranks in first-appearance order of the
 * rank-ordered shards (= the sequential token order, StringCheck.cpp:46-55), ready for crass_b200_cluster_block_dev on every
 * rank.  d_send: token_block_bytes(cap), d_recv: world x that, d_merged: token_block_bytes(out_cap). */
int crass_b200_comm_unique_id(uint8_t id[128]);
int crass_b200_ctx_comm_init(crass_b200_ctx* ctx, const uint8_t id[128], int rank, int world);
int crass_b200_ctx_comm_world(const crass_b200_ctx* ctx);          /* 0 = no communicator */
int crass_b200_exchange_tokens_dev(crass_b200_ctx* ctx, const crass_b200_hit* d_hits, uint32_t n_hits, const void* d_tokens, uint32_t stride,
                                   void* d_send, uint32_t cap, void* d_recv, uint32_t shard_reads, void* d_merged, uint32_t out_cap,
                                   void* stream);

/* ---- whole path, one call: searchFile* -> createNonRedundantSet -> findSingletons* --------------- */
int crass_b200_run_files(crass_b200_ctx* ctx, const char* const* paths, uint32_t n_paths,
                         const crass_b200_params* params, int phases, crass_b200_results** out, int* max_read_len);

/* ---- the same on one or more GPUs of one box, one caller, one set of containers -----------------------
 * An engine owns one context, stream and host thread per device.  A file's reads are cut into contiguous shards, one per
 * device; phase 1 and phase 2 run on the shards side by side; the shards' distinct DR tokens meet on the first device in
 * ONE all-gather of fixed-size token blocks (NCCL when the devices are distinct and libnccl.so.2 can be loaded -- it is
 * dlopen'ed, not linked -- peer copies otherwise; CRASS_B200_EXCHANGE=peer|host selects) and are merged there in
 * first-appearance order, the numbering StringCheck gives a sequential run (StringCheck.cpp:46-55).  Hit records come
 * back to the CALLING thread as one list in global read order with batch-wide read indices, so the caller fills its
 * single ReadMap exactly as after a one-GPU run (WorkHorse.cpp:321-414; readsFound is tested on the host by header,
 * libcrispr.cpp:411, which also covers headers that repeat across shards).  The same device may be named more than once
 * (the shards then share it): that is how the multi-device logic is tested on a box with one GPU.
 * A searched file stays parsed (host memory; page-locked once the engine reuses its pooled buffers, CRASS_B200_PIN) and resident in HBM (bytes, offsets, phase-1 flags, 2-bit stream)
 * until released, so findSingletons neither parses nor uploads it again; CRASS_B200_RESIDENT_MB bounds the bytes a device
 * keeps (default: half its memory), older files are uploaded again from the host copy. */
typedef struct crass_b200_engine crass_b200_engine;
int crass_b200_engine_create(const int* devices, uint32_t n_devices, crass_b200_engine** out);
void crass_b200_engine_destroy(crass_b200_engine* e);
uint32_t crass_b200_engine_num_devices(const crass_b200_engine* e);
int crass_b200_engine_uses_nccl(const crass_b200_engine* e);
/* searchFile (libcrispr.h:74-80): parse, shard, copy in, K1 on every device; hits / ss_pool are malloc'd (crass_b200_free),
 * *batch stays owned by the engine until crass_b200_engine_release_file (or a new search of the same path). */
int crass_b200_engine_search_file(crass_b200_engine* e, const char* path, const crass_b200_params* params,
                                  const crass_b200_batch** batch, crass_b200_hit** hits, uint32_t* n_hits,
                                  uint32_t** ss_pool, uint32_t* n_ss_pool);
/* the step between the phases for the ONE file searched last: K4b on every device, all-gather, K4c merge and
 * createNonRedundantSet (K5 + host passes) on the first device; *ac = the matcher for phase 2 (NULL: no DR found) */
int crass_b200_engine_exchange(crass_b200_engine* e, const char* path, uint32_t kmer_clust, crass_b200_ac** ac,
                               uint32_t* n_variants, uint32_t* n_patterns);
/* findSingletons (libcrispr.h:86-92): K2 on every device over the resident shards of `path` (parsed and copied in now if it
 * was not searched through this engine); skip_found != 0 leaves out the reads this engine's phase 1 flagged */
int crass_b200_engine_find_singletons(crass_b200_engine* e, const char* path, const crass_b200_ac* ac, int skip_found,
                                      const crass_b200_batch** batch, crass_b200_hit** hits, uint32_t* n_hits,
                                      uint32_t** ss_pool, uint32_t* n_ss_pool);
void crass_b200_engine_release_file(crass_b200_engine* e, const char* path);
/* The same two calls STREAMED, for a caller that fills its own containers (the drop-in libcrispr, whose ReadMap holds the
 * reference's ReadHolder objects): the file goes through the devices in ranges of about CRASS_B200_STREAM_MB (each ending
 * on a record start), and `fn` receives the hits of each range in file order -- batch / hits / ss_pool are valid during the
 * call only, first_read is the file index of the range's read 0, a non-zero return ends the run with that code.  For
 * search_file_ranges fn runs on one helper thread of the engine while this call parses and searches later ranges (the caller
 * is blocked in the call, so its containers see one thread); for find_singletons_ranges it runs on the calling thread while a
 * helper runs K2 on the next range.  Replaces the kseq loop of libcrispr.cpp:96-131 and the second one of :487-513. */
typedef int (*crass_b200_range_fn)(void* user, const crass_b200_batch* batch, const crass_b200_hit* hits, uint32_t n_hits,
                                   const uint32_t* ss_pool, uint32_t n_ss_pool, uint64_t first_read);
int crass_b200_engine_search_file_ranges(crass_b200_engine* e, const char* path, const crass_b200_params* params,
                                         crass_b200_range_fn fn, void* user, int* max_read_len);
int crass_b200_engine_find_singletons_ranges(crass_b200_engine* e, const char* path, const crass_b200_ac* ac, int skip_found,
                                             crass_b200_range_fn fn, void* user);
/* WorkHorse::parseSeqFiles: searchFile* -> createNonRedundantSet -> findSingletons* on the engine's devices (streamed: parsing
 * of range i+1, copy + K1 of range i and the replay of range i-1 overlap, across file boundaries too) */
int crass_b200_engine_run_files(crass_b200_engine* e, const char* const* paths, uint32_t n_paths, const crass_b200_params* params,
                                int phases, crass_b200_results** out, int* max_read_len);
int crass_b200_run_files_multi(const int* devices, uint32_t n_devices, const char* const* paths, uint32_t n_paths,
                               const crass_b200_params* params, int phases, crass_b200_results** out, int* max_read_len);
/* bookkeeping for bench.py: bytes copied host-to-device / device-to-host so far, kernel launches so far, and the stage
 * times (ms) of the most recent crass_b200_engine_run_files */
void crass_b200_engine_transfer_bytes(const crass_b200_engine* e, uint64_t* h2d, uint64_t* d2h);
uint64_t crass_b200_engine_launch_count(const crass_b200_engine* e);
void crass_b200_engine_stage_ms(const crass_b200_engine* e, double* parse, double* phase1, double* exchange, double* phase2);

void crass_b200_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* CRASS_B200_H */
