// libcrispr_b200.cpp -- drop-in replacement for crass's src/crass/libcrispr.cpp.
//
// Same free functions, same signatures (src/crass/libcrispr.h:74-125), same side effects on the caller's
// ReadMap / StringCheck / lookupTables, same return values and exceptions -- but the two per-read hot loops run
// on a B200 through the C-ABI of include/crass_b200.h.  A crass maintainer compiles THIS file instead of
// libcrispr.cpp (against crass's own headers) and links libcrass_b200.so; WorkHorse, NodeManager and the XML writer
// are untouched.  See INTEGRATION.md.  It is built here by oracle/Makefile (target `ref`, output
// oracle/_ref/libcrass_dropin.so) only to prove the claim: tests/test_dropin.py runs the reference's harness and
// the reference's own Catch tests against it.
//
// What stays on the host, exactly as in the reference: ReadHolder construction, addReadHolder (DRLowLexi is the
// reference's own ReadHolder method), the container updates in read order, progress lines, logging.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "libcrispr.h"          // crass's header (found through -I<crass>/src/crass)
#include "LoggerSimp.h"
#include "Exception.h"
#include "StlExt.h"
#include "config.h"

#include "crass_b200.h"         // -I<repo>/include

namespace {

// One context for the single-read entry points, one engine for the two file-level ones.  CRASS_B200_DEVICES=0,1,...
// (default: CRASS_B200_DEVICE or 0) names the GPUs the reads are sharded over; the hit records come back to this thread
// as one list in read order whatever their number, so the containers below are filled as by a one-GPU run.
struct Engine {
    crass_b200_ctx* ctx = nullptr;
    crass_b200_engine* eng = nullptr;
    ~Engine() { if (eng) crass_b200_engine_destroy(eng); if (ctx) crass_b200_ctx_destroy(ctx); }
};
Engine& the_engine() { static Engine e; return e; }

std::vector<int> device_list() {
    std::vector<int> devs;
    if (const char* l = getenv("CRASS_B200_DEVICES")) {
        std::stringstream ss(l);
        std::string tok;
        while (std::getline(ss, tok, ',')) if (!tok.empty()) devs.push_back(atoi(tok.c_str()));
    }
    if (devs.empty()) { const char* dev = getenv("CRASS_B200_DEVICE"); devs.push_back(dev ? atoi(dev) : 0); }
    return devs;
}

crass_b200_ctx* engine() {
    Engine& e = the_engine();
    if (!e.ctx) {
        if (crass_b200_ctx_create(device_list()[0], &e.ctx))
            throw crispr::exception(__FILE__, __LINE__, __PRETTY_FUNCTION__, crass_b200_last_error());
    }
    return e.ctx;
}

crass_b200_engine* file_engine() {
    Engine& e = the_engine();
    if (!e.eng) {
        const std::vector<int> devs = device_list();
        if (crass_b200_engine_create(devs.data(), (uint32_t)devs.size(), &e.eng))
            throw crispr::exception(__FILE__, __LINE__, __PRETTY_FUNCTION__, crass_b200_last_error());
    }
    return e.eng;
}

void check(int rc, const char* file, int line, const char* fn) {
    if (rc) throw crispr::exception(file, line, fn, crass_b200_last_error());
}
#define B200_TRY(expr) check((expr), __FILE__, __LINE__, __PRETTY_FUNCTION__)

crass_b200_params to_params(const options& o) {
    crass_b200_params p;
    crass_b200_default_params(&p);
    p.low_dr = o.lowDRsize; p.high_dr = o.highDRsize; p.low_spacer = o.lowSpacerSize; p.high_spacer = o.highSpacerSize;
    p.window = o.searchWindowLength; p.min_repeats = o.minNumRepeats; p.kmer_clust = (uint32_t)o.kmer_clust_size;
    return p;
}

struct HitGuard {
    crass_b200_hit* hits = nullptr; uint32_t* pool = nullptr; uint32_t n = 0, np = 0;
    ~HitGuard() { crass_b200_free(hits); crass_b200_free(pool); }
};

// the holder searchFile / on_match build from a kseq record (libcrispr.cpp:112-131, 425-435)
void fill_holder(ReadHolder& h, const crass_b200_batch* b, uint32_t i) {
    uint32_t len = 0;
    const uint8_t* seq = crass_b200_batch_read(b, i, &len);
    h.setSequence(std::string((const char*)seq, (size_t)len));
    h.setHeader(crass_b200_batch_name(b, i));
    int has = 0;
    const char* c = crass_b200_batch_comment(b, i, &has);
    if (has) h.setComment(c);
    const char* q = crass_b200_batch_qual(b, i, &has);
    if (has) h.setQual(q);
}

void progress(const char* who, int count, time_t& start) {
    time_t now; time(&now);
    std::cout << "\r[" << PACKAGE_NAME << "_" << who << "]: " << "Processed " << count << " ...";
    std::cout << difftime(now, start) << " sec" << std::flush;
}

void one_read(const ReadHolder& h, const uint32_t* extra, size_t n_extra, std::vector<uint8_t>& bases, uint64_t offs[2]) {
    std::string s = const_cast<ReadHolder&>(h).getSeq();
    bases.assign(s.begin(), s.end());
    (void)extra; (void)n_extra;
    offs[0] = 0; offs[1] = bases.size();
}



// The progress lines of one file.  The reference prints one before every 100 000th read of a FILE as it reads it (its
// log_counter is per call, libcrispr.cpp:91,99-109; :495-496 in phase 2) showing the running total over all files, and one
// more when the file ends.  Here the reads arrive range by range: a line that falls exactly on the end of a range is held back
// until the next range shows that the file goes on.
struct Ticker {
    const char* who; int base; time_t* start;
    uint64_t seen = 0;                                              // reads of this file handed over so far
    bool held = false;
    Ticker(const char* w, int b, time_t* s) : who(w), base(b), start(s) {}
    // reads [seen, seen + n) arrive; `replay(upto)` handles the hits of the reads below local index upto
    template <class Replay>
    void range(uint32_t n, Replay replay) {
        if (held && n) { progress(who, base + (int)seen, *start); held = false; }
        uint32_t done = 0;
        while (done < n) {
            const uint64_t to_tick = CRASS_DEF_READ_COUNTER_LOGGER - seen % CRASS_DEF_READ_COUNTER_LOGGER;
            const uint32_t upto = (uint32_t)std::min<uint64_t>(n, (uint64_t)done + to_tick);
            replay(upto);
            seen += upto - done;
            done = upto;
            if (seen % CRASS_DEF_READ_COUNTER_LOGGER == 0) {
                if (done < n) progress(who, base + (int)seen, *start); else held = true;
            }
        }
    }
};

struct Phase1Sink {
    ReadMap* reads; StringCheck* strings; lookupTable* patterns; lookupTable* found;
    Ticker tick;
    std::string failure;                                            // what() of an exception the containers threw
    Phase1Sink(ReadMap* r, StringCheck* s, lookupTable* p, lookupTable* f, int base, time_t* t)
        : reads(r), strings(s), patterns(p), found(f), tick("patternFinder", base, t) {}
};

int phase1_range(void* user, const crass_b200_batch* batch, const crass_b200_hit* hits, uint32_t n_hits, const uint32_t* pool,
                 uint32_t, uint64_t) {
    Phase1Sink& k = *(Phase1Sink*)user;
    try {
        uint32_t at = 0;
        k.tick.range(crass_b200_batch_num_reads(batch), [&](uint32_t upto) {
            for (; at < n_hits && hits[at].read_index < upto; ++at) {           // hits come sorted by read index
                const crass_b200_hit& ht = hits[at];
                ReadHolder tmp_holder;
                fill_holder(tmp_holder, batch, ht.read_index);
                for (uint32_t i = 0; i + 1 < ht.n_ss; i += 2) tmp_holder.startStopsAdd(pool[ht.ss_offset + i], pool[ht.ss_offset + i + 1]);
                tmp_holder.setRepeatLength((int)ht.repeat_len);
                addReadHolder(k.reads, k.strings, tmp_holder);
                (*k.patterns)[tmp_holder.repeatStringAt(0)] = true;
                (*k.found)[tmp_holder.getHeader()] = true;
            }
        });
    } catch (crispr::exception& e) {                                            // must not leave the engine's thread
        k.failure = e.what();
        return CRASS_B200_EINVAL;
    }
    return 0;
}

struct Phase2Sink {
    ReadMap* reads; StringCheck* strings; lookupTable* found;
    Ticker tick;
    std::string failure;                                            // what() of an exception the containers threw
    Phase2Sink(ReadMap* r, StringCheck* s, lookupTable* f, int base, time_t* t) : reads(r), strings(s), found(f), tick("singletonFinder", base, t) {}
};

int phase2_range(void* user, const crass_b200_batch* batch, const crass_b200_hit* hits, uint32_t n_hits, const uint32_t* pool,
                 uint32_t, uint64_t) {
    Phase2Sink& k = *(Phase2Sink*)user;
    try {
        uint32_t at = 0;
        k.tick.range(crass_b200_batch_num_reads(batch), [&](uint32_t upto) {
            for (; at < n_hits && hits[at].read_index < upto; ++at) {           // on_match (libcrispr.cpp:408-442)
                const crass_b200_hit& ht = hits[at];
                const char* name = crass_b200_batch_name(batch, ht.read_index);
                if (k.found->find(name) != k.found->end()) continue;
                ReadHolder tmp_holder;
                fill_holder(tmp_holder, batch, ht.read_index);
                tmp_holder.startStopsAdd(pool[ht.ss_offset], pool[ht.ss_offset + 1]);
                addReadHolder(k.reads, k.strings, tmp_holder);
            }
        });
    } catch (crispr::exception& e) {                                            // must not unwind through the engine (its helper thread is running)
        k.failure = e.what();
        return CRASS_B200_EINVAL;
    }
    return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
int searchFile(const char* inputFastq, const options& opts, ReadMap* mReads, StringCheck* mStringCheck,
               lookupTable& patternsHash, lookupTable& readsFound, time_t& time_start) {
    static int read_counter = 0;
    crass_b200_params p = to_params(opts);
    Phase1Sink sink(mReads, mStringCheck, &patternsHash, &readsFound, read_counter, &time_start);
    int max_read_length = 0;
    // The kseq loop of libcrispr.cpp:96-131 as a pipeline: worker threads parse range i+1 of the file while range i is copied
    // to the devices and searched (K1 on every device) and the hits of range i-1 are replayed, in read order, into the caller's
    // containers by phase1_range.  The ranges stay parsed and resident in HBM for findSingletons.
    if (crass_b200_engine_search_file_ranges(file_engine(), inputFastq, &p, phase1_range, &sink, &max_read_length)) {
        if (!sink.failure.empty()) {
            std::cerr << sink.failure << std::endl;
            throw crispr::exception(__FILE__, __LINE__, __PRETTY_FUNCTION__, "Fatal error in search algorithm!");
        }
        const std::string why = crass_b200_last_error();
        if (why.find("cannot open") != std::string::npos) {          // getFileHandle() exits on an unopenable file (SeqUtils.cpp:112-122)
            std::cerr << PACKAGE_NAME << " : [ERROR] Could not open FASTQ " << inputFastq << " for reading." << std::endl;
            exit(1);
        }
        std::cerr << why << std::endl;
        throw crispr::exception(__FILE__, __LINE__, __PRETTY_FUNCTION__, "Fatal error in search algorithm!");
    }
    read_counter += (int)sink.tick.seen;
    logInfo("finished processing file:" << inputFastq, 1);
    progress("patternFinder", read_counter, time_start);
    logInfo("So far " << mReads->size() << " direct repeat variants have been found from " << read_counter << " reads", 2);
    return max_read_length;
}

void findSingletons(const char* inputFastq, const options& opts, std::vector<std::string>* nonRedundantPatterns,
                    lookupTable& readsFound, ReadMap* mReads, StringCheck* mStringCheck, time_t& startTime) {
    (void)opts;
    static int read_counter = 0;
    std::vector<uint8_t> bytes;
    std::vector<uint32_t> offs(1, 0);
    for (std::vector<std::string>::iterator it = nonRedundantPatterns->begin(); it != nonRedundantPatterns->end(); ++it) {
        bytes.insert(bytes.end(), it->begin(), it->end());
        offs.push_back((uint32_t)bytes.size());
    }
    crass_b200_ac* ac = NULL;
    B200_TRY(crass_b200_ac_build(bytes.data(), offs.data(), (uint32_t)nonRedundantPatterns->size(), &ac));
    Phase2Sink sink(mReads, mStringCheck, &readsFound, read_counter, &startTime);
    // K2 over the ranges searchFile left in HBM (no second parse, no second copy; a file this engine has not seen is parsed
    // and copied in now), the matches of one range replayed while the next is scanned.  Every read is scanned: readsFound is the
    // CALLER's table, keyed by header, and may hold names the device flags know nothing about, so that test stays on the host
    // (libcrispr.cpp:411).
    int rc = crass_b200_engine_find_singletons_ranges(file_engine(), inputFastq, ac, 0, phase2_range, &sink);
    crass_b200_ac_destroy(ac);
    if (rc && !sink.failure.empty()) throw crispr::exception(__FILE__, __LINE__, __PRETTY_FUNCTION__, sink.failure.c_str());
    if (rc) {
        const std::string why = crass_b200_last_error();
        if (why.find("cannot open") != std::string::npos) {
            std::cerr << PACKAGE_NAME << " : [ERROR] Could not open FASTQ " << inputFastq << " for reading." << std::endl;
            exit(1);
        }
        throw crispr::exception(__FILE__, __LINE__, __PRETTY_FUNCTION__, why.c_str());
    }
    read_counter += (int)sink.tick.seen;
    progress("singletonFinder", read_counter, startTime);
    crass_b200_engine_release_file(file_engine(), inputFastq);      // phase 2 is the last use of a file (WorkHorse.cpp:381-399)
}

// ---- single-read entry points kept for source compatibility (and used by crass's own unit tests) -----------------
int searchCore(ReadHolder& tmpHolder, const options& opts) {
    std::vector<uint8_t> bases; uint64_t offs[2];
    one_read(tmpHolder, NULL, 0, bases, offs);
    if (bases.size() < opts.lowDRsize + opts.lowSpacerSize + opts.searchWindowLength + 1) {
        logWarn("Read " << tmpHolder.getHeader() << " is too short. With current parameters, the minimum length must be "
                        << opts.lowDRsize + opts.lowSpacerSize + opts.searchWindowLength + 1 << "bp (read is " << bases.size() << "bp)", 3);
        return false;
    }
    crass_b200_params p = to_params(opts);
    HitGuard hg;
    B200_TRY(crass_b200_dr_search(engine(), bases.data(), offs, 1, &p, NULL, &hg.hits, &hg.n, &hg.pool, &hg.np));
    tmpHolder.clearStartStops();
    if (!hg.n) return false;
    for (uint32_t i = 0; i + 1 < hg.hits[0].n_ss; i += 2) tmpHolder.startStopsAdd(hg.pool[hg.hits[0].ss_offset + i], hg.pool[hg.hits[0].ss_offset + i + 1]);
    tmpHolder.setRepeatLength((int)hg.hits[0].repeat_len);
    return true;
}

int scanRight(ReadHolder& tmp_holder, std::string& pattern, unsigned int minSpacerLength, unsigned int scanRange) {
    std::string seq = tmp_holder.getSeq();
    StartStopList ss = tmp_holder.getStartStopList();
    const uint32_t before = (uint32_t)ss.size();
    const uint32_t cap = 2 * ((uint32_t)seq.size() / 4 + 8);
    ss.resize(cap);
    uint32_t n = before;
    B200_TRY(crass_b200_scan_right(engine(), (const uint8_t*)seq.data(), (uint32_t)seq.size(), ss.data(), &n, cap,
                                   (const uint8_t*)pattern.data(), (uint32_t)pattern.size(), minSpacerLength, scanRange));
    for (uint32_t i = before; i + 1 < n; i += 2) tmp_holder.startStopsAdd(ss[i], ss[i + 1]);
    return (int)tmp_holder.back();
}

unsigned int extendPreRepeat(ReadHolder& tmp_holder, int searchWindowLength, int minSpacerLength) {
    std::string seq = tmp_holder.getSeq();
    StartStopList ss = tmp_holder.getStartStopList();
    uint32_t replen = 0;
    B200_TRY(crass_b200_extend_pre_repeat(engine(), (const uint8_t*)seq.data(), (uint32_t)seq.size(), ss.data(), (uint32_t)ss.size(),
                                          (uint32_t)searchWindowLength, (uint32_t)minSpacerLength, &replen));
    tmp_holder.clearStartStops();
    for (size_t i = 0; i + 1 < ss.size(); i += 2) tmp_holder.startStopsAdd(ss[i], ss[i + 1]);
    tmp_holder.setRepeatLength((int)replen);
    return replen;
}

// ---- host-side helpers of the reference's interface that are not on the per-read path -------------------------------
bool testSpacerLength(int minSpacerLength, int maxSpacerLength, int minAllowedSpacerLength, int maxAllowedSpacerLength) {
    return !(minSpacerLength < minAllowedSpacerLength) && !(maxSpacerLength > maxAllowedSpacerLength);
}
bool testSpacerRepeatSimilarity(float similarity) { return !(similarity > CRASS_DEF_SPACER_OR_REPEAT_MAX_SIMILARITY); }
bool testSpacerSpacerSimilarity(float similarity) { return !(similarity > CRASS_DEF_SPACER_OR_REPEAT_MAX_SIMILARITY); }
bool testSpacerSpacerLengthDiff(int difference) { return !(difference > CRASS_DEF_SPACER_TO_SPACER_LENGTH_DIFF); }
bool testRepeatSpacerLengthDiff(int difference) { return !(difference > CRASS_DEF_SPACER_TO_REPEAT_LENGTH_DIFF); }

bool isRepeatLowComplexity(std::string& repeat) {                         // used by WorkHorse.cpp:1196
    int counts[5] = {0, 0, 0, 0, 0};
    for (std::string::iterator it = repeat.begin(); it != repeat.end(); ++it) {
        switch (*it) {
            case 'c': case 'C': counts[0]++; break;
            case 't': case 'T': counts[1]++; break;
            case 'a': case 'A': counts[2]++; break;
            case 'g': case 'G': counts[3]++; break;
            default: counts[4]++; break;
        }
    }
    const int cut_off = static_cast<int>(static_cast<int>(repeat.length()) * CRASS_DEF_LOW_COMPLEXITY_THRESHHOLD);
    for (int i = 0; i < 5; ++i) if (counts[i] > cut_off) return true;
    return false;
}

bool drHasHighlyAbundantKmers(std::string& directRepeat, float& maxFrequency) {   // used by WorkHorse.cpp:1206
    std::map<std::string, int> kmer_counter;
    const size_t kmer_length = 3;
    const size_t max_index = directRepeat.length() - kmer_length;
    int total_count = 0, max_count = 0;
    for (size_t i = 0; i < max_index; i++) {
        std::string kmer = directRepeat.substr(i, kmer_length);
        addOrIncrement(kmer_counter, kmer);
        total_count++;
    }
    for (std::map<std::string, int>::iterator it = kmer_counter.begin(); it != kmer_counter.end(); ++it)
        if (it->second > max_count) max_count = it->second;
    maxFrequency = static_cast<float>(max_count) / static_cast<float>(total_count);
    return maxFrequency > CRASS_DEF_KMER_MAX_ABUNDANCE_CUTOFF;
}

bool drHasHighlyAbundantKmers(std::string& directRepeat) {
    float max;
    return drHasHighlyAbundantKmers(directRepeat, max);
}

bool qcFoundRepeats(ReadHolder& tmp_holder, int minSpacerLength, int maxSpacerLength) {
    if (tmp_holder.numRepeats() < 2)
        throw crispr::exception(__FILE__, __LINE__, __PRETTY_FUNCTION__, "The vector holding the repeat indexes has less than 2 repeats!");
    std::string seq = tmp_holder.getSeq();
    StartStopList ss = tmp_holder.getStartStopList();
    int result = 0;
    B200_TRY(crass_b200_qc_found_repeats(engine(), (const uint8_t*)seq.data(), (uint32_t)seq.size(), ss.data(), (uint32_t)ss.size(),
                                         minSpacerLength, maxSpacerLength, &result));
    return result == 1;
}

void addReadHolder(ReadMap* mReads, StringCheck* mStringCheck, ReadHolder& tmpReadholder) {
    ReadHolder* candidate = new ReadHolder(tmpReadholder);
    std::string dr_lowlexi;
    try {
        dr_lowlexi = candidate->DRLowLexi();
    } catch (crispr::exception& e) {
        std::cerr << e.what() << std::endl;
        throw crispr::exception(__FILE__, __LINE__, __PRETTY_FUNCTION__, "Cannot obtain read in lowlexi form");
    }
    StringToken st = mStringCheck->getToken(dr_lowlexi);
    if (0 == st) {
        st = mStringCheck->addString(dr_lowlexi);
        (*mReads)[st] = new ReadList();
    }
    (*mReads)[st]->push_back(candidate);
}
