// TEST INFRASTRUCTURE ONLY (see oracle_common.h).  CPU restatement of the DBoW2 transform the reference runs per (key)frame
// (Frame::ComputeBoW src/Frame.cc:738-745, KeyFrame::ComputeBoW):
//   TemplatedVocabulary::transform(features, BowVector&, FeatureVector&, levelsup)   Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1197
//   TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup)          :1218-1259  (tree descent, first minimum wins)
//   BowVector::addWeight / addIfNotExist / normalize                                   BowVector.cpp:35-88
//   FeatureVector::addFeature                                                          FeatureVector.cpp:30-45
//   L1Scoring::score                                                                   ScoringObject.cpp:23-68
// Pinned by oracle/_ref, which compiles the reference's own DBoW2 sources (tests/test_ref_pins_oracle_cpu.py).
#include "oracle_common.h"

#include <algorithm>
#include <cmath>
#include <map>
#include <vector>

extern "C" {

// Vocabulary as flat arrays: children of node n = children[childStart[n] .. childStart[n+1]) (node ids, in the order the vocabulary lists them);
// desc [nNodes][32]; weight [nNodes]; wordId [nNodes] (-1 for inner nodes).  weighting: 0 TF_IDF, 1 TF, 2 IDF, 3 BINARY; norm: 0 none, 1 L1, 2 L2.
// Outputs in std::map order: words (id ascending, value), feature-vector entries (node id ascending, then feature index in insertion order).
int orbo_bow_transform(int L, int weighting, int norm, const int* childStart, const int* children, const uint8_t* desc, const double* weight, const int* wordId,
                       int N, const uint8_t* feat, int levelsup, int* outWord, double* outValue, int capWords, int* fvNode, int* fvFeature, int* nFeat) {
    std::map<unsigned, double> v;
    std::map<unsigned, std::vector<unsigned>> fv;
    auto dist = [](const uint8_t* a, const uint8_t* b) { int d = 0; for (int i = 0; i < 32; ++i) d += __builtin_popcount(a[i] ^ b[i]); return d; };
    const int nid_level = L - levelsup;
    for (int i = 0; i < N; ++i) {
        const uint8_t* f = feat + (size_t)i * 32;
        unsigned nid = 0, final_id = 0;
        int current_level = 0;
        do {
            ++current_level;
            const int a = childStart[final_id], b = childStart[final_id + 1];
            final_id = children[a];
            double best_d = dist(f, desc + (size_t)final_id * 32);
            for (int c = a + 1; c < b; ++c) {
                const unsigned id = children[c];
                const double d = dist(f, desc + (size_t)id * 32);
                if (d < best_d) { best_d = d; final_id = id; }
            }
            if (current_level == nid_level) nid = final_id;
        } while (childStart[final_id + 1] > childStart[final_id]);
        const unsigned wid = (unsigned)wordId[final_id];
        const double w = weight[final_id];
        if (w > 0) {
            if (weighting == 0 || weighting == 1) v[wid] += w;          // addWeight (a new entry starts from 0.0 + w == w)
            else if (!v.count(wid)) v[wid] = w;                          // addIfNotExist
            fv[nid].push_back(i);
        }
    }
    if ((weighting == 0 || weighting == 1) && !v.empty() && norm == 0) {
        const double nd = (double)v.size();
        for (auto& kv : v) kv.second /= nd;
    }
    if (norm) {
        double nrm = 0.0;
        if (norm == 1) for (auto& kv : v) nrm += std::fabs(kv.second);
        else { for (auto& kv : v) nrm = std::fma(kv.second, kv.second, nrm); nrm = std::sqrt(nrm); }   // one FMA per term in the reference's -O3 -march=native build (checked against oracle/_ref)
        if (nrm > 0.0) for (auto& kv : v) kv.second /= nrm;
    }
    int n = 0;
    for (auto& kv : v) { if (n < capWords) { outWord[n] = (int)kv.first; outValue[n] = kv.second; } ++n; }
    int m = 0;
    for (auto& kv : fv) for (unsigned fi : kv.second) { if (m < N) { fvNode[m] = (int)kv.first; fvFeature[m] = (int)fi; } ++m; }
    *nFeat = m;
    return n;
}

double orbo_bow_score_l1(int n1, const int* id1, const double* v1, int n2, const int* id2, const double* v2) {
    int i = 0, j = 0;
    double score = 0;
    while (i < n1 && j < n2) {
        if (id1[i] == id2[j]) { score += std::fabs(v1[i] - v2[j]) - std::fabs(v1[i]) - std::fabs(v2[j]); ++i; ++j; }
        else if (id1[i] < id2[j]) i = (int)(std::lower_bound(id1 + i, id1 + n1, id2[j]) - id1);
        else j = (int)(std::lower_bound(id2 + j, id2 + n2, id1[i]) - id2);
    }
    return -score / 2.0;
}

}  // extern "C"
