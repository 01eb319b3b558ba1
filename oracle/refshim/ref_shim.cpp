// ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (parity oracle), never part of the product.
//
// Thin extern "C" wrappers around the UNMODIFIED reference hot path, compiled from the
// sources where they lie under /root/reference by oracle/Makefile into
// oracle/_ref/libcrass_ref.so.  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load that library.
//
// Everything in here calls the reference's own functions:
//   searchFile / searchCore / scanRight / extendPreRepeat / qcFoundRepeats /
//   findSingletons / addReadHolder        (src/crass/libcrispr.cpp)
//   PatternMatcher::bmpSearch / levenstheinDistance / getStringSimilarity
//   ReadHolder::DRLowLexi, reverseComplement, kseq_read, acism_create / acism_scan
// except the step between the two phases (WorkHorse::createNonRedundantSet,
// src/crass/WorkHorse.cpp:612-709,1404-1637), which cannot be compiled here because
// WorkHorse.h pulls in Xerces-C; it is restated below on the same STL containers.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>
#include <map>
#include <sstream>
#include <iostream>
#include <algorithm>
#include <chrono>

#include "libcrispr.h"
#include "LoggerSimp.h"
#include "Exception.h"
#include "PatternMatcher.h"
#include "ReadHolder.h"
#include "SeqUtils.h"
#include "StringCheck.h"
#include "kseq.h"

#include "SmithWaterman.h"
#include "Aligner.h"
#include "Types.h"
extern "C" {
#include "ksw.h"
}

extern "C" {
#include "../aho-corasick/msutil.h"
#include "../aho-corasick/acism.h"
}

namespace {

struct CoutSilencer {                                      // the progress lines of libcrispr; CRASS_REF_SHOW_PROGRESS=1 lets them through
    std::streambuf* old;
    std::ostringstream sink;
    bool on;
    CoutSilencer() : old(nullptr), on(getenv("CRASS_REF_SHOW_PROGRESS") == nullptr) { if (on) old = std::cout.rdbuf(sink.rdbuf()); }
    ~CoutSilencer() { if (on) std::cout.rdbuf(old); }
};

bool g_inited = false;

void fill_options(options& o, const uint32_t* p) {
    // defaults as in src/crass/crass.cpp:430-460, then the six searched fields + kmer_clust_size
    o.logLevel = 0;
    o.reportStats = false;
    o.lowDRsize = p[0];
    o.highDRsize = p[1];
    o.lowSpacerSize = p[2];
    o.highSpacerSize = p[3];
    o.searchWindowLength = p[4];
    o.minNumRepeats = p[5];
    o.kmer_clust_size = (int)p[6];
    o.output_fastq = "./";
    o.delim = "\t";
    o.logToScreen = false;
    o.coverageBins = -1;
    o.graphColourType = BLUE_RED;
    o.layoutAlgorithm = "neato";
    o.longDescription = false;
    o.showSingles = false;
    o.cNodeKmerLength = 7;
    o.covCutoff = 3;
}

void holder_from(ReadHolder& h, const char* seq, uint32_t len, const uint32_t* ss, uint32_t n_ss) {
    h.setSequence(std::string(seq, len));
    h.setHeader("r");
    for (uint32_t i = 0; i + 1 < n_ss; i += 2) h.startStopsAdd(ss[i], ss[i + 1]);
}

int copy_ss(ReadHolder& h, uint32_t* ss, uint32_t cap, uint32_t* n_ss) {
    StartStopList l = h.getStartStopList();
    *n_ss = (uint32_t)l.size();
    if (l.size() > cap) return -2;
    for (size_t i = 0; i < l.size(); ++i) ss[i] = l[i];
    return 0;
}

// ---- restatement of the inter-phase step on the reference's containers ----
bool sort_len_asc(const std::string& a, const std::string& b) { return a.length() < b.length(); }
bool include_substring(const std::string& a, const std::string& b) {          // WorkHorse.cpp:78-86
    if (std::string::npos != b.find(a)) return true;
    if (std::string::npos != b.find(reverseComplement(a))) return true;
    return false;
}
bool is_not_empty(const std::string& a) { return !a.empty(); }

void remove_redundant(std::vector<std::string>& v) {                           // WorkHorse.cpp:612-645
    std::sort(v.begin(), v.end(), sort_len_asc);
    for (size_t i = 0; i < v.size(); ++i) {
        if (v[i].empty()) continue;
        for (size_t j = i + 1; j < v.size(); ++j) {
            if (v[j].empty()) continue;
            if (include_substring(v[i], v[j])) v[j].clear();
        }
    }
    std::vector<std::string>::iterator e = std::partition(v.begin(), v.end(), is_not_empty);
    v.erase(e, v.end());
}

struct ClusterState {
    std::map<std::string, int> k2gid;
    std::map<int, std::vector<int> > dr2gid;
    int next_gid;
    ClusterState() : next_gid(1) {}
};

void cluster_dr(ClusterState& cs, StringCheck& sc, int token, int min_count) { // WorkHorse.cpp:1404-1637
    const int K = 11;                                                          // crassDefines.h:66
    std::string dr = sc.getString(token);
    int num_mers = (int)dr.length() - K + 1;
    std::vector<std::string> homeless;
    std::map<int, int> group_count;
    int group = 0;
    for (int i = 0; i < num_mers; ++i) {
        std::string km = laurenize(dr.substr(i, K));
        std::map<std::string, int>::iterator it = cs.k2gid.find(km);
        if (it == cs.k2gid.end()) {
            homeless.push_back(km);
        } else if (0 == group) {
            std::map<int, int>::iterator g = group_count.find(it->second);
            if (g == group_count.end()) {
                group_count[it->second] = 1;
            } else {
                g->second++;
                if (min_count <= g->second) group = it->second;
            }
        }
    }
    if (0 == group) group = cs.next_gid++;
    cs.dr2gid[group].push_back(token);
    for (size_t i = 0; i < homeless.size(); ++i) cs.k2gid[homeless[i]] = group;
}

std::vector<std::string> restated_non_redundant_set(ReadMap& reads, StringCheck& sc, int min_count,
                                                    std::vector<std::pair<int, int> >* token_groups) {
    ClusterState cs;
    for (ReadMapIterator it = reads.begin(); it != reads.end(); ++it) cluster_dr(cs, sc, it->first, min_count);
    std::vector<std::string> out;
    for (std::map<int, std::vector<int> >::iterator g = cs.dr2gid.begin(); g != cs.dr2gid.end(); ++g) {
        std::vector<std::string> v;
        for (size_t i = 0; i < g->second.size(); ++i) {
            v.push_back(sc.getString(g->second[i]));
            if (token_groups) token_groups->push_back(std::make_pair(g->second[i], g->first));
        }
        remove_redundant(v);
        std::vector<std::string> rc;
        for (size_t i = 0; i < v.size(); ++i) rc.push_back(reverseComplement(v[i]));
        out.insert(out.end(), v.begin(), v.end());
        out.insert(out.end(), rc.begin(), rc.end());
    }
    return out;
}

}  // namespace
// the reference's own clusterDRReads / removeRedundantRepeats / createNonRedundantSet (workhorse_pin.cpp)
std::vector<std::string> workhorse_non_redundant_set(ReadMap& reads, StringCheck& sc, options& o, std::vector<std::pair<int, int> >* token_groups);
namespace {

// The step between the phases runs the reference's own WorkHorse code; CRASS_REF_CLUSTER=restated selects the restatement
// above instead (tests compare the two).
std::vector<std::string> non_redundant_set(ReadMap& reads, StringCheck& sc, options& o, std::vector<std::pair<int, int> >* token_groups) {
    const char* sel = getenv("CRASS_REF_CLUSTER");
    if (sel && !strcmp(sel, "restated")) return restated_non_redundant_set(reads, sc, o.kmer_clust_size, token_groups);
    return workhorse_non_redundant_set(reads, sc, o, token_groups);
}

uint32_t fnv1a(const std::string& s) {
    uint32_t h = 2166136261u;
    for (size_t i = 0; i < s.size(); ++i) { h ^= (unsigned char)s[i]; h *= 16777619u; }
    return h;
}

void dump_reads(std::ostringstream& os, ReadMap& reads, StringCheck& sc, std::map<ReadHolder*, int>& phase_of) {
    for (ReadMapIterator it = reads.begin(); it != reads.end(); ++it) {
        if (!it->second) continue;
        os << "T\t" << it->first << "\t" << sc.getString(it->first) << "\t" << it->second->size() << "\n";
        for (ReadListIterator r = it->second->begin(); r != it->second->end(); ++r) {
            ReadHolder* h = *r;
            os << "R\t" << it->first << "\t" << phase_of[h] << "\t" << h->getHeader() << "\t"
               << (h->getLowLexi() ? 1 : 0) << "\t" << h->getRepeatLength() << "\t";
            StartStopList l = h->getStartStopList();
            for (size_t i = 0; i < l.size(); ++i) { if (i) os << ","; os << l[i]; }
            os << "\t" << h->getSeq() << "\t" << h->getComment() << "\t" << (h->getIsFasta() ? 1 : 0)
               << "\t" << fnv1a(h->getQual()) << "\n";
        }
    }
}

char* dup_string(const std::string& s) {
    char* p = (char*)malloc(s.size() + 1);
    memcpy(p, s.data(), s.size());
    p[s.size()] = 0;
    return p;
}

struct RefAc {
    char* conc;
    MEMREF* pattv;
    int npatts;
    ACISM* psp;
};

struct FirstMatch { int strnum; int textpos; int calls; };
int first_match_cb(int strnum, int textpos, void* ctx) {
    FirstMatch* f = (FirstMatch*)ctx;
    f->strnum = strnum; f->textpos = textpos; f->calls++;
    return 1;                                                                  // libcrispr.cpp:441
}

}  // namespace

extern "C" {

void ref_init(void) {
    if (!g_inited) {
        intialiseGlobalLogger("", 0);
        g_inited = true;
    }
}

const char* ref_version(void) { return PACKAGE_NAME " " PACKAGE_VERSION " (reference, compiled for oracle/_ref)"; }

// params: [lowDR, highDR, lowSpacer, highSpacer, window, minRepeats, kmerClust]
int ref_search_core(const char* seq, uint32_t len, const uint32_t* params, uint32_t* ss, uint32_t ss_cap,
                    uint32_t* n_ss, uint32_t* replen) {
    ref_init();
    options o; fill_options(o, params);
    try {
        ReadHolder h;
        h.setSequence(std::string(seq, len));
        h.setHeader("r");
        int found = searchCore(h, o);
        *replen = h.getRepeatLength();
        if (copy_ss(h, ss, ss_cap, n_ss)) return -2;
        return found ? 1 : 0;
    } catch (crispr::exception& e) { return -1; } catch (std::exception& e) { return -3; }
}

int ref_scan_right(const char* seq, uint32_t len, uint32_t* ss, uint32_t* n_ss, uint32_t ss_cap,
                   const char* pat, uint32_t patlen, uint32_t min_spacer, uint32_t scan_range) {
    ref_init();
    try {
        ReadHolder h; holder_from(h, seq, len, ss, *n_ss);
        std::string p(pat, patlen);
        int r = scanRight(h, p, min_spacer, scan_range);
        if (copy_ss(h, ss, ss_cap, n_ss)) return -2;
        return r;
    } catch (crispr::exception& e) { return -1; }
}

int ref_extend_pre_repeat(const char* seq, uint32_t len, uint32_t* ss, uint32_t n_ss, int window, int min_spacer) {
    ref_init();
    try {
        ReadHolder h; holder_from(h, seq, len, ss, n_ss);
        uint32_t n2 = 0;
        int r = (int)extendPreRepeat(h, window, min_spacer);
        if (copy_ss(h, ss, n_ss, &n2)) return -2;
        return r;
    } catch (crispr::exception& e) { return -1; }
}

int ref_qc_found_repeats(const char* seq, uint32_t len, const uint32_t* ss, uint32_t n_ss, int min_spacer, int max_spacer) {
    ref_init();
    try {
        ReadHolder h; holder_from(h, seq, len, ss, n_ss);
        return qcFoundRepeats(h, min_spacer, max_spacer) ? 1 : 0;
    } catch (crispr::exception& e) { return -1; } catch (std::exception& e) { return -3; }
}

int ref_bmp_search(const char* text, uint32_t tlen, const char* pat, uint32_t plen) {
    return PatternMatcher::bmpSearch(std::string(text, tlen), std::string(pat, plen));
}

int ref_edit_distance(const char* a, uint32_t alen, const char* b, uint32_t blen) {
    std::string s(a, alen), t(b, blen);
    return PatternMatcher::levenstheinDistance(s, t);
}

float ref_similarity(const char* a, uint32_t alen, const char* b, uint32_t blen) {
    std::string s(a, alen), t(b, blen);
    return PatternMatcher::getStringSimilarity(s, t);
}

int ref_low_complexity(const char* a, uint32_t alen) {
    std::string s(a, alen);
    return isRepeatLowComplexity(s) ? 1 : 0;
}

void ref_revcomp(const char* a, uint32_t alen, char* out) {
    std::string r = reverseComplement(std::string(a, alen));
    memcpy(out, r.data(), r.size() < alen ? r.size() : alen);
}

// DRLowLexi on a holder with the given SS; writes the oriented read, mirrored SS, the DR string.
int ref_dr_lowlexi(const char* seq, uint32_t len, uint32_t* ss, uint32_t n_ss, char* seq_out, char* dr_out,
                   uint32_t* dr_len, int* was_lowlexi) {
    ref_init();
    try {
        ReadHolder h; holder_from(h, seq, len, ss, n_ss);
        std::string dr = h.DRLowLexi();
        uint32_t n2 = 0;
        copy_ss(h, ss, n_ss, &n2);
        std::string s = h.getSeq();
        memcpy(seq_out, s.data(), s.size());
        memcpy(dr_out, dr.data(), dr.size());
        *dr_len = (uint32_t)dr.size();
        *was_lowlexi = h.getLowLexi() ? 1 : 0;
        return 0;
    } catch (crispr::exception& e) { return -1; } catch (std::exception& e) { return -3; }
}

// ---- acism, driven exactly like findSingletons does (libcrispr.cpp:452-469) ----
void* ref_ac_create(const char* const* pats, const uint32_t* lens, uint32_t n) {
    std::string conc;
    for (uint32_t i = 0; i < n; ++i) { conc += std::string(pats[i], lens[i]); conc += "\n"; }
    RefAc* a = new RefAc;
    a->conc = new char[conc.size() + 1];
    std::copy(conc.begin(), conc.end(), a->conc);
    a->conc[conc.size()] = '\0';
    a->conc[conc.size() - 1] = '\0';
    a->pattv = refsplit(a->conc, '\n', &a->npatts);
    a->psp = acism_create(a->pattv, a->npatts);
    return a;
}

// returns 1 and (end = textpos, len = pattern length) of the FIRST callback, 0 when no match
int ref_ac_first_match(void* handle, const char* text, uint32_t tlen, int* end, int* plen) {
    RefAc* a = (RefAc*)handle;
    FirstMatch f = { -1, -1, 0 };
    MEMREF t = { text, tlen };
    (void)acism_scan(a->psp, t, (ACISM_ACTION*)first_match_cb, &f);
    if (!f.calls) return 0;
    *end = f.textpos;
    *plen = (int)a->pattv[f.strnum].len;
    return 1;
}

void ref_ac_destroy(void* handle) {
    RefAc* a = (RefAc*)handle;
    acism_destroy(a->psp);
    free(a->pattv);
    delete[] a->conc;
    delete a;
}

// ---- kseq: the record stream exactly as searchFile sees it (incl. stale comment/qual buffers) ----
char* ref_kseq_dump(const char* path) {
    gzFile fp = gzopen(path, "r");
    if (!fp) return NULL;
    kseq_t* seq = kseq_init(fp);
    std::ostringstream os;
    int l;
    while ((l = kseq_read(seq)) >= 0) {
        os << seq->name.s << "\t" << (seq->comment.s ? seq->comment.s : "\x01") << "\t" << seq->seq.s << "\t"
           << (seq->qual.s ? seq->qual.s : "\x01") << "\n";
    }
    os << "#ret=" << l << "\n";
    kseq_destroy(seq);
    gzclose(fp);
    return dup_string(os.str());
}

// ---- whole path: searchFile* -> (restated) createNonRedundantSet -> findSingletons* ----
// phases: 1 = stop after phase 1, 2 = run both.  timings_ms[0..2] = phase1, cluster, phase2 (may be NULL).
char* ref_run_files(const char* const* paths, uint32_t n_paths, const uint32_t* params, int phases, double* timings_ms) {
    ref_init();
    CoutSilencer quiet;
    options o; fill_options(o, params);
    ReadMap reads;
    StringCheck sc;
    lookupTable patterns, found;
    std::ostringstream os;
    std::map<ReadHolder*, int> phase_of;
    int max_len = 0;
    time_t t0; time(&t0);
    typedef std::chrono::steady_clock clk;
    try {
        clk::time_point a = clk::now();
        for (uint32_t i = 0; i < n_paths; ++i) {
            int m = searchFile(paths[i], o, &reads, &sc, patterns, found, t0);
            if (m > max_len) max_len = m;
        }
        clk::time_point b = clk::now();
        for (ReadMapIterator it = reads.begin(); it != reads.end(); ++it)
            for (ReadListIterator r = it->second->begin(); r != it->second->end(); ++r) phase_of[*r] = 1;
        size_t n_p1 = found.size();
        std::vector<std::pair<int, int> > token_groups;
        std::vector<std::string> nr = non_redundant_set(reads, sc, o, &token_groups);
        clk::time_point c = clk::now();
        if (phases >= 2 && nr.size() > 0) {
            time(&t0);
            for (uint32_t i = 0; i < n_paths; ++i) findSingletons(paths[i], o, &nr, found, &reads, &sc, t0);
        }
        clk::time_point d = clk::now();
        if (timings_ms) {
            timings_ms[0] = std::chrono::duration<double, std::milli>(b - a).count();
            timings_ms[1] = std::chrono::duration<double, std::milli>(c - b).count();
            timings_ms[2] = std::chrono::duration<double, std::milli>(d - c).count();
        }
        for (ReadMapIterator it = reads.begin(); it != reads.end(); ++it)
            for (ReadListIterator r = it->second->begin(); r != it->second->end(); ++r)
                if (!phase_of.count(*r)) phase_of[*r] = 2;
        os << "# crass-dump v1\n";
        os << "M\t" << max_len << "\t" << n_p1 << "\t" << patterns.size() << "\n";
        for (size_t i = 0; i < token_groups.size(); ++i) os << "G\t" << token_groups[i].first << "\t" << token_groups[i].second << "\n";
        std::vector<std::string> nrs(nr);
        std::sort(nrs.begin(), nrs.end());
        for (size_t i = 0; i < nrs.size(); ++i) os << "P\t" << nrs[i] << "\n";
        for (lookupTable::iterator it = patterns.begin(); it != patterns.end(); ++it) os << "H\t" << it->first << "\n";
        dump_reads(os, reads, sc, phase_of);
    } catch (crispr::exception& e) {
        os << "E\t" << e.what() << "\n";
    }
    for (ReadMapIterator it = reads.begin(); it != reads.end(); ++it) {
        if (!it->second) continue;
        for (ReadListIterator r = it->second->begin(); r != it->second->end(); ++r) delete *r;
        delete it->second;
    }
    return dup_string(os.str());
}

// which code runs the step between the phases in this build / environment
const char* ref_cluster_impl(void) {
    const char* sel = getenv("CRASS_REF_CLUSTER");
    return (sel && !strcmp(sel, "restated")) ? "restatement (ref_shim.cpp)" : "reference WorkHorse.cpp:612-709,1404-1637 (compiled from the reference's own text)";
}

// clustering + non-redundant set alone, on an ordered list of DR token strings (tokens 2,3,...)
char* ref_non_redundant(const char* const* drs, const uint32_t* lens, uint32_t n, int min_count) {
    StringCheck sc;
    ReadMap reads;
    for (uint32_t i = 0; i < n; ++i) { int t = sc.addString(std::string(drs[i], lens[i])); reads[t] = NULL; }
    std::vector<std::pair<int, int> > tg;
    options o;
    uint32_t dflt[7] = {23, 47, 26, 50, 8, 2, (uint32_t)min_count};
    fill_options(o, dflt);
    CoutSilencer quiet;
    std::vector<std::string> nr = non_redundant_set(reads, sc, o, &tg);
    std::ostringstream os;
    for (size_t i = 0; i < tg.size(); ++i) os << "G\t" << tg[i].first << "\t" << tg[i].second << "\n";
    for (size_t i = 0; i < nr.size(); ++i) os << "P\t" << nr[i] << "\n";
    return dup_string(os.str());
}

// ---- partial-DR recovery: the consumers of the path's start/stop lists ----
// smithWaterman (SmithWaterman.cpp:151) as updateStartStops calls it; the two returned strings come back as lengths
// plus the position of the second one's first/last occurrence in b (what the caller tests with find/rfind).
int ref_smith_waterman(const char* a, uint32_t la, const char* b, uint32_t lb, int a_start_search, int a_search_len,
                       double similarity, int* a_start_align, int* a_end_align, uint32_t* first_len, uint32_t* second_len,
                       int* second_find, int* second_rfind) {
    ref_init();
    try {
        std::string sa(a, la), sb(b, lb);
        stringPair sp = smithWaterman(sa, sb, a_start_align, a_end_align, a_start_search, a_search_len, similarity);
        *first_len = (uint32_t)sp.first.length();
        *second_len = (uint32_t)sp.second.length();
        *second_find = (int)sb.find(sp.second);
        *second_rfind = (int)sb.rfind(sp.second);
        return sp.first.empty() && sp.second.empty() ? 0 : 1;
    } catch (crispr::exception& e) { return -1; } catch (std::exception& e) { return -3; }
}

// ReadHolder::updateStartStops (ReadHolder.cpp:382) on a holder with the given start/stop list
int ref_update_start_stops(const char* seq, uint32_t len, uint32_t* ss, uint32_t* n_ss, uint32_t cap, int front_offset,
                           const char* dr, uint32_t dr_len, uint32_t low_spacer) {
    ref_init();
    try {
        ReadHolder h; holder_from(h, seq, len, ss, *n_ss);
        const uint32_t defaults[7] = {23, 47, low_spacer, 50, 8, 2, 6};
        options o; fill_options(o, defaults);
        std::string d(dr, dr_len);
        h.updateStartStops(front_offset, &d, &o);
        if (copy_ss(h, ss, cap, n_ss)) return -2;
        return 0;
    } catch (crispr::exception& e) { return -1; } catch (std::exception& e) { return -3; }
}

// ---- consensus DR of a group (second half): the reference's own Aligner (Aligner.cpp:72-418) and the
// ksw_align it calls (ksw.c:330), driven the way WorkHorse::parseGroupedDRs does (WorkHorse.cpp:1166-1171, 750-773):
// setMasterDR on DR 0, alignSlave for every other DR of the group, generateConsensus.
// reads: n_reads sequences back to back with their start/stop lists; read_dr[i] = index of the DR (token) the read hangs on.
// Out: dr_place[d] = AL_Offsets of DR d (-1: the slave could not be placed), dr_flags[d] bit 0 = the slave and its reads were
// reverse complemented, zone[2], consensus / conservation [array_len], coverage[4 * array_len] (rows A C G T).
int ref_ksw_align(const uint8_t* query, int qlen, const uint8_t* target, int tlen, int xtra, int* out7) {
    // query / target are nt4 codes (0..4); mat / gaps as the Aligner sets them up (Aligner.h:105-131)
    int8_t mat[25]; int k = 0;
    for (int i = 0; i < 4; ++i) { for (int j = 0; j < 4; ++j) mat[k++] = i == j ? 1 : -3; mat[k++] = 0; }
    for (int j = 0; j < 5; ++j) mat[k++] = 0;
    std::vector<uint8_t> q(query, query + qlen), t(target, target + tlen);
    q.push_back(0); t.push_back(0);
    kswq_t* prof = 0;
    kswr_t r = ksw_align(qlen, q.data(), tlen, t.data(), 5, mat, 5, 2, xtra, &prof);
    free(prof);
    out7[0] = r.score; out7[1] = r.te; out7[2] = r.qe; out7[3] = r.score2; out7[4] = r.te2; out7[5] = r.tb; out7[6] = r.qb;
    return 0;
}

int ref_consensus_group(const char* bases, const uint64_t* offsets, uint32_t n_reads, const uint32_t* read_dr,
                        const uint32_t* ss_offsets, const uint32_t* ss_pool, const char* dr_bytes, const uint32_t* dr_offsets,
                        uint32_t n_drs, uint32_t array_len, int32_t* dr_place, uint8_t* dr_flags, int32_t* zone,
                        char* consensus, float* conservation, int32_t* coverage) {
    ref_init();
    CoutSilencer quiet;
    ReadMap reads;
    StringCheck sc;
    std::vector<StringToken> tok(n_drs);
    try {
        for (uint32_t d = 0; d < n_drs; ++d) {
            tok[d] = sc.addString(std::string(dr_bytes + dr_offsets[d], dr_offsets[d + 1] - dr_offsets[d]));
            reads[tok[d]] = new ReadList();
        }
        for (uint32_t i = 0; i < n_reads; ++i) {
            ReadHolder* h = new ReadHolder();
            holder_from(*h, bases + offsets[i], (uint32_t)(offsets[i + 1] - offsets[i]), ss_pool + ss_offsets[i], ss_offsets[i + 1] - ss_offsets[i]);
            reads[tok[read_dr[i]]]->push_back(h);
        }
        Aligner al((int)array_len, &reads, &sc);
        al.setMasterDR(tok[0]);
        dr_place[0] = al.offset(tok[0]);
        dr_flags[0] = 0;
        for (uint32_t d = 1; d < n_drs; ++d) {
            StringToken t = tok[d];
            al.alignSlave(t);
            dr_flags[d] = t != tok[d] ? 1 : 0;
            dr_place[d] = al.offset(t);
        }
        al.generateConsensus();
        zone[0] = al.getDRZoneStart(); zone[1] = al.getDRZoneEnd();
        const char alphabet[4] = {'A', 'C', 'G', 'T'};
        for (uint32_t j = 0; j < array_len; ++j) {
            consensus[j] = al.consensusAt((int)j);
            conservation[j] = al.conservationAt((int)j);
            for (int k = 0; k < 4; ++k) coverage[(size_t)k * array_len + j] = al.coverageAt((int)j, alphabet[k]);
        }
    } catch (crispr::exception& e) { return -1; } catch (std::exception& e) { return -3; }
    for (ReadMapIterator it = reads.begin(); it != reads.end(); ++it)
        if (it->second) { for (size_t i = 0; i < it->second->size(); ++i) delete (*it->second)[i]; delete it->second; }
    return 0;
}

void ref_free(void* p) { free(p); }

}  // extern "C"
