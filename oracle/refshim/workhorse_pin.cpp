// workhorse_pin.cpp -- TEST INFRASTRUCTURE ONLY (compiled into oracle/_ref/libcrass_ref.so and libcrass_dropin.so).
//
// The step between the phases (clusterDRReads / removeRedundantRepeats / createNonRedundantSet) executed by the
// REFERENCE'S OWN CODE: the function bodies are cut verbatim out of /root/reference/src/crass/WorkHorse.cpp at build time
// (gen_workhorse_excerpt.py -> oracle/_ref/gen/workhorse_excerpt.inc) and compiled here as members of a WorkHorse shell
// that holds exactly the members those bodies touch (WorkHorse.h:185-199).  WorkHorse.cpp as a translation unit cannot be
// compiled in this image (WorkHorse.h pulls in the Xerces-C headers through writer.h); these three functions use none.
#include <algorithm>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "config.h"
#include "crassDefines.h"
#include "Exception.h"
#include "LoggerSimp.h"
#include "ReadHolder.h"
#include "SeqUtils.h"
#include "StlExt.h"
#include "StringCheck.h"
#include "Types.h"

class WorkHorse {                                   // shell: see WorkHorse.h:104-122 for the declarations, :185-199 for the members
public:
    void removeRedundantRepeats(Vecstr& repeatVector);
    Vecstr* createNonRedundantSet(GroupKmerMap& groupKmerCountsMap, int& nextFreeGID);
    bool clusterDRReads(StringToken DRToken, int* nextFreeGID, std::map<std::string, int>* k2GIDMap, GroupKmerMap* groupKmerCountsMap);

    ReadMap mReads;
    options* mOpts;
    StringCheck mStringCheck;
    std::map<int, bool> mGroupMap;
    DR_Cluster_Map mDR2GIDMap;
};

#include "workhorse_excerpt.inc"

// createNonRedundantSet as WorkHorse::parseSeqFiles calls it (WorkHorse.cpp:367-370) on the harness's containers;
// token_groups gets (token, group id) in group order
std::vector<std::string> workhorse_non_redundant_set(ReadMap& reads, StringCheck& sc, options& o, std::vector<std::pair<int, int> >* token_groups) {
    WorkHorse wh;
    wh.mReads = reads;                              // the keys are what counts (read lists are not touched)
    wh.mStringCheck = sc;
    wh.mOpts = &o;
    GroupKmerMap group_kmer_counts_map;
    int next_free_GID = 1;                          // WorkHorse.cpp:369
    Vecstr* nr = wh.createNonRedundantSet(group_kmer_counts_map, next_free_GID);
    std::vector<std::string> out(nr->begin(), nr->end());
    delete nr;
    if (token_groups)
        for (DR_Cluster_MapIterator g = wh.mDR2GIDMap.begin(); g != wh.mDR2GIDMap.end(); ++g)
            if (g->second)
                for (DR_ClusterIterator t = g->second->begin(); t != g->second->end(); ++t) token_groups->push_back(std::make_pair((int)*t, g->first));
    for (DR_Cluster_MapIterator g = wh.mDR2GIDMap.begin(); g != wh.mDR2GIDMap.end(); ++g) delete g->second;
    for (GroupKmerMap::iterator k = group_kmer_counts_map.begin(); k != group_kmer_counts_map.end(); ++k) delete k->second;
    return out;
}
