// TEST INFRASTRUCTURE ONLY -- C entry point over the REFERENCE's own DBoW2: Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h (header-only template,
// included where it lies) instantiated as ORBVocabulary = TemplatedVocabulary<FORB::TDescriptor, FORB> (include/ORBVocabulary.h:30), linked with the
// reference's BowVector.cpp / FeatureVector.cpp / ScoringObject.cpp / FORB.cpp / DUtils/Random.cpp (compiled by oracle/Makefile from $(REF)).
// The vocabulary tree comes in as flat arrays (tests build synthetic trees; the 145 MB ORBvoc.txt is not shipped); this file only fills
// m_nodes / m_words and flattens the BowVector / FeatureVector that transform() returns (Frame::ComputeBoW, src/Frame.cc:738-745, levelsup = 4).
#include "DBoW2/FORB.h"
#include "DBoW2/TemplatedVocabulary.h"

namespace {
typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;
struct Voc : ORBVocabulary {
    Voc(int k, int L, DBoW2::WeightingType w, DBoW2::ScoringType s) : ORBVocabulary(k, L, w, s) {}
    using ORBVocabulary::m_nodes;
    using ORBVocabulary::m_words;
};
}  // namespace

extern "C" {

// nodes: parent[n] (root = node 0, parent[0] ignored), descriptor[n][32], weight[n] (doubles; used for leaves), children are appended to their parent in
// node-id order; a node without children is a word, word ids are assigned in node-id order (as TemplatedVocabulary::createWords / loadFromTextFile do).
void* ref_bow_create(int k, int L, int weighting, int scoring, int nNodes, const int* parent, const uint8_t* desc, const double* weight) {
    Voc* v = new Voc(k, L, (DBoW2::WeightingType)weighting, (DBoW2::ScoringType)scoring);
    v->m_nodes.resize(nNodes);
    for (int i = 0; i < nNodes; ++i) {
        v->m_nodes[i].id = i;
        v->m_nodes[i].weight = weight[i];
        v->m_nodes[i].descriptor = cv::Mat(1, 32, CV_8UC1, (void*)(desc + (size_t)i * 32), 32).clone();
        if (i > 0) { v->m_nodes[i].parent = parent[i]; v->m_nodes[parent[i]].children.push_back(i); }
    }
    for (int i = 0; i < nNodes; ++i)
        if (i > 0 && v->m_nodes[i].isLeaf()) { v->m_nodes[i].word_id = (DBoW2::WordId)v->m_words.size(); v->m_words.push_back(&v->m_nodes[i]); }
    return v;
}
void ref_bow_destroy(void* h) { delete (Voc*)h; }

// transform(features, BowVector, FeatureVector, levelsup): outputs in map order -- words (id, value) ascending id; feature vector entries
// (node id, feature index) ascending node id, then insertion (= feature) order.  Returns the number of words; *nFeat the number of entries.
int ref_bow_transform(void* h, int N, const uint8_t* desc, int levelsup, int* wordId, double* wordValue, int capWords, int* fvNode, int* fvFeature, int* nFeat) {
    Voc* v = (Voc*)h;
    std::vector<cv::Mat> features(N);
    for (int i = 0; i < N; ++i) features[i] = cv::Mat(1, 32, CV_8UC1, (void*)(desc + (size_t)i * 32), 32);
    DBoW2::BowVector bv;
    DBoW2::FeatureVector fv;
    v->transform(features, bv, fv, levelsup);
    int n = 0;
    for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++n)
        if (n < capWords) { wordId[n] = (int)it->first; wordValue[n] = it->second; }
    int m = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it)
        for (size_t j = 0; j < it->second.size(); ++j, ++m)
            if (m < N) { fvNode[m] = (int)it->first; fvFeature[m] = (int)it->second[j]; }
    *nFeat = m;
    return n;
}

// L1Scoring::score (ScoringObject.cpp) of two flattened BowVectors -- what KeyFrameDatabase compares
double ref_bow_score(void* h, int n1, const int* id1, const double* v1, int n2, const int* id2, const double* v2) {
    DBoW2::BowVector a, b;
    for (int i = 0; i < n1; ++i) a.addWeight(id1[i], v1[i]);
    for (int i = 0; i < n2; ++i) b.addWeight(id2[i], v2[i]);
    return ((Voc*)h)->score(a, b);
}

}  // extern "C"
