/*
 * oracle/ref_dbow2_capi.cc -- C entry points around the REFERENCE's own DBoW2 (S/Thirdparty/DBoW2/DBoW2/{TemplatedVocabulary.h, FORB.cpp,
 * BowVector.cpp, FeatureVector.cpp, ScoringObject.cpp} compiled unmodified against oracle/dbowshim by oracle/Makefile into
 * oracle/_ref/libref_dbow2.so): load a vocabulary with the reference's own loadFromTextFile (the ORBvoc.txt format) and run
 * transform(features, BowVector, FeatureVector, levelsup) as Frame::ComputeBoW does (S/src/Frame.cc:395-402).
 *
 * TEST INFRASTRUCTURE ONLY (tests/test_oracle_vs_reference.py).
 */
#include <cstring>
#include <vector>
#include "FORB.h"
#include "TemplatedVocabulary.h"

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;      /* S/include/ORBVocabulary.h */

extern "C" {

void *ref_dbow2_load_text(const char *path)
{
    ORBVocabulary *v = new ORBVocabulary();
    if (!v->loadFromTextFile(path)) { delete v; return nullptr; }
    return v;
}
void ref_dbow2_destroy(void *v) { delete (ORBVocabulary *)v; }
int ref_dbow2_size(void *v) { return (int)((ORBVocabulary *)v)->size(); }

/* returns the BowVector size; bow_ids / bow_vals ascending by word; fv as CSR (nodes ascending, items in push order) */
int ref_dbow2_transform(void *vp, int N, const unsigned char *desc, int levelsup, int *bow_ids, double *bow_vals, int *fv_nodes, int *fv_start, int *fv_items,
                        int *fv_count)
{
    ORBVocabulary *v = (ORBVocabulary *)vp;
    std::vector<cv::Mat> feats(N);
    for (int i = 0; i < N; i++) { feats[i].create(1, 32, CV_8U); std::memcpy(feats[i].ptr<unsigned char>(), desc + (size_t)32 * i, 32); }
    DBoW2::BowVector bv; DBoW2::FeatureVector fv;
    v->transform(feats, bv, fv, levelsup);
    int nb = 0;
    for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++nb) { bow_ids[nb] = (int)it->first; bow_vals[nb] = it->second; }
    int nf = 0, pos = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++nf) {
        fv_nodes[nf] = (int)it->first; fv_start[nf] = pos;
        for (size_t k = 0; k < it->second.size(); k++) fv_items[pos++] = (int)it->second[k];
    }
    fv_start[nf] = pos;
    *fv_count = nf;
    return nb;
}

}
