// host_shim_test.cc -- drives the C++ drop-in classes (ORBextractor / ORBmatcher / Optimizer with the reference's
// signatures) on binary inputs written by tests/test_host_shim_gpu.py and dumps their outputs for comparison with
// the oracle.  Usage: host_shim_test <dir>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>
#include "ORBextractor.h"
#include "ORBmatcher.h"
#include "Optimizer.h"

using namespace iORB_SLAM;

template <typename T> std::vector<T> rd(const std::string &p)
{
    std::ifstream f(p, std::ios::binary | std::ios::ate);
    if (!f) { fprintf(stderr, "cannot open %s\n", p.c_str()); exit(2); }
    const size_t n = (size_t)f.tellg();
    std::vector<T> v(n / sizeof(T));
    f.seekg(0); f.read((char *)v.data(), n);
    return v;
}
template <typename T> void wr(const std::string &p, const std::vector<T> &v)
{
    std::ofstream f(p, std::ios::binary);
    f.write((const char *)v.data(), v.size() * sizeof(T));
}

static void fill_frame(Frame &F, ORBextractor &ex, const cv::Mat &img)
{
    ex(img, cv::Mat(), F.mvKeys, F.mDescriptors);
    F.N = (int)F.mvKeys.size();
    F.mvKeysUn = F.mvKeys;                          // no distortion
    F.mvuRight.assign(F.N, -1.f);
    F.mvpMapPoints.assign(F.N, nullptr);
    F.mvbOutlier.assign(F.N, false);
    F.mvScaleFactors = ex.GetScaleFactors();
    F.mvInvLevelSigma2 = ex.GetInverseScaleSigmaSquares();
}

static void dump_frame(const std::string &d, const std::string &tag, const Frame &F)
{
    std::vector<float> k; std::vector<int> o; std::vector<unsigned char> desc;
    for (int i = 0; i < F.N; i++) {
        const cv::KeyPoint &p = F.mvKeys[i];
        k.push_back(p.pt.x); k.push_back(p.pt.y); k.push_back(p.angle); k.push_back(p.response); k.push_back(p.size);
        o.push_back(p.octave);
        desc.insert(desc.end(), F.mDescriptors.ptr(i), F.mDescriptors.ptr(i) + 32);
    }
    wr(d + "/" + tag + "_kp.bin", k); wr(d + "/" + tag + "_oct.bin", o); wr(d + "/" + tag + "_desc.bin", desc);
}

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    const std::string d = argv[1];
    const std::vector<int> dims = rd<int>(d + "/dims.bin");            // w, h, nfeatures
    const int w = dims[0], h = dims[1];
    ORBextractor ex(dims[2], 1.2f, 8, 20, 7);
    Frame Last, Cur;
    for (int f = 0; f < 2; f++) {
        std::vector<unsigned char> px = rd<unsigned char>(d + (f ? "/frame1.bin" : "/frame0.bin"));
        cv::Mat img(h, w, CV_8UC1);
        memcpy(img.data, px.data(), px.size());
        fill_frame(f ? Cur : Last, ex, img);
    }
    dump_frame(d, "last", Last); dump_frame(d, "cur", Cur);
    ex.FetchImagePyramid();
    std::vector<unsigned char> lvl3(ex.mvImagePyramid[3].data, ex.mvImagePyramid[3].data + (size_t)ex.mvImagePyramid[3].rows * ex.mvImagePyramid[3].cols);
    wr(d + "/pyr3.bin", lvl3);

    // tracking: last-frame map points -> SearchByProjection(Cur, Last) -> PoseOptimization
    const std::vector<float> K4 = rd<float>(d + "/K4.bin"), bounds = rd<float>(d + "/bounds.bin"), Tcw = rd<float>(d + "/Tcw.bin"), Xw = rd<float>(d + "/Xw.bin");
    const std::vector<unsigned char> valid = rd<unsigned char>(d + "/valid.bin");
    Frame::fx = K4[0]; Frame::fy = K4[1]; Frame::cx = K4[2]; Frame::cy = K4[3];
    Frame::mnMinX = bounds[0]; Frame::mnMinY = bounds[1]; Frame::mnMaxX = bounds[2]; Frame::mnMaxY = bounds[3];
    std::vector<MapPoint> mps(Last.N);
    for (int i = 0; i < Last.N; i++) {
        mps[i].mnId = i;
        for (int c = 0; c < 3; c++) mps[i].mWorldPos.at<float>(c) = Xw[3 * i + c];
        memcpy(mps[i].mDescriptor.ptr(0), Last.mDescriptors.ptr(i), 32);
        if (valid[i]) Last.mvpMapPoints[i] = &mps[i];
    }
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) Cur.mTcw.at<float>(r, c) = Tcw[4 * r + c];
    ORBmatcher matcher(0.9, true);
    const int nmatches = matcher.SearchByProjection(Cur, Last, 15, true);
    std::vector<int> fm(Cur.N, -1);
    for (int k = 0; k < Cur.N; k++) if (Cur.mvpMapPoints[k]) fm[k] = (int)Cur.mvpMapPoints[k]->mnId;
    wr(d + "/fm.bin", fm);
    const std::vector<float> T0 = rd<float>(d + "/T0.bin");
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) Cur.mTcw.at<float>(r, c) = T0[4 * r + c];
    const int ninl = Optimizer::PoseOptimization(&Cur);
    std::vector<float> Tout; std::vector<unsigned char> outl;
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) Tout.push_back(Cur.mTcw.at<float>(r, c));
    for (int k = 0; k < Cur.N; k++) outl.push_back(Cur.mvbOutlier[k] ? 1 : 0);
    wr(d + "/Tpose.bin", Tout); wr(d + "/outlier.bin", outl);
    wr(d + "/counts.bin", std::vector<int>{nmatches, ninl, ORBmatcher::DescriptorDistance(Cur.mDescriptors.row(0), Cur.mDescriptors.row(1))});

    // local BA on a mock map: KF 0 has mnId 0 (fixed, local), KF 1 is only a fixed camera (not covisible)
    const std::vector<int> bdims = rd<int>(d + "/ba_dims.bin");        // K, P, E
    const int K = bdims[0], P = bdims[1], E = bdims[2];
    const std::vector<float> bposes = rd<float>(d + "/ba_poses.bin"), bpts = rd<float>(d + "/ba_points.bin"), buv = rd<float>(d + "/ba_uv.bin"),
                             bw = rd<float>(d + "/ba_w.bin");
    const std::vector<int> bkf = rd<int>(d + "/ba_kf.bin"), bpt = rd<int>(d + "/ba_pt.bin"), boct = rd<int>(d + "/ba_oct.bin");
    const std::vector<double> bintr = rd<double>(d + "/ba_intr.bin");
    std::vector<KeyFrame> kfs(K);
    std::vector<MapPoint> pts(P);
    Map map;
    for (int k = 0; k < K; k++) {
        kfs[k].mnId = k; kfs[k].fx = (float)bintr[0]; kfs[k].fy = (float)bintr[1]; kfs[k].cx = (float)bintr[2]; kfs[k].cy = (float)bintr[3];
        for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) kfs[k].Tcw.at<float>(r, c) = bposes[16 * k + 4 * r + c];
        kfs[k].mvInvLevelSigma2 = ex.GetInverseScaleSigmaSquares();
        map.mvKFs.push_back(&kfs[k]);
    }
    for (int p = 0; p < P; p++) { pts[p].mnId = p; pts[p].nObs = 0; for (int c = 0; c < 3; c++) pts[p].mWorldPos.at<float>(c) = bpts[3 * p + c]; map.mvMPs.push_back(&pts[p]); }
    for (int e = 0; e < E; e++) {
        KeyFrame &kf = kfs[bkf[e]];
        const size_t idx = kf.mvKeysUn.size();
        kf.mvKeysUn.push_back(cv::KeyPoint(buv[2 * e], buv[2 * e + 1], 31.f, 0.f, 0.f, boct[e]));
        kf.mvuRight.push_back(-1.f);
        kf.mvpMapPoints.push_back(&pts[bpt[e]]);
        pts[bpt[e]].mObservations[&kf] = idx; pts[bpt[e]].nObs++;
    }
    // current KF = last one; covisible = all but KF 1
    KeyFrame *cur = &kfs[K - 1];
    for (int k = 0; k < K - 1; k++) if (k != 1) cur->mvpOrderedConnectedKeyFrames.push_back(&kfs[k]);
    bool stop = false;
    Optimizer::LocalBundleAdjustment(cur, &stop, &map);
    std::vector<float> oposes, opts; std::vector<int> nobs;
    for (int k = 0; k < K; k++) for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) oposes.push_back(kfs[k].Tcw.at<float>(r, c));
    for (int p = 0; p < P; p++) { for (int c = 0; c < 3; c++) opts.push_back(pts[p].mWorldPos.at<float>(c)); nobs.push_back(pts[p].nObs); }
    wr(d + "/ba_out_poses.bin", oposes); wr(d + "/ba_out_points.bin", opts); wr(d + "/ba_out_nobs.bin", nobs);
    // MapPoint::ComputeDistinctiveDescriptors for the whole mock map in one call: keyframe descriptors from a file, one row per observation
    {
        const std::vector<unsigned char> kd = rd<unsigned char>(d + "/ba_desc.bin");      // [E, 32] in edge order
        std::vector<size_t> row_of_edge(E);
        std::vector<int> nrow(K, 0);
        for (int e = 0; e < E; e++) row_of_edge[e] = nrow[bkf[e]]++;
        for (int k = 0; k < K; k++) kfs[k].mDescriptors = cv::Mat(std::max(nrow[k], 1), 32, CV_8U);
        for (int e = 0; e < E; e++) memcpy(kfs[bkf[e]].mDescriptors.ptr((int)row_of_edge[e]), &kd[32 * (size_t)e], 32);
        kfs[2].mbBad = true;                                                              // observations in a bad keyframe are left out (MapPoint.cc:262-266)
        pts[3].mbBad = true;                                                              // a bad point keeps its descriptor (:251-253)
        std::vector<MapPoint *> all;
        for (int p = 0; p < P; p++) { memset(pts[p].mDescriptor.ptr(0), 0xAB, 32); all.push_back(&pts[p]); }
        ORBmatcher::ComputeDistinctiveDescriptors(all);
        std::vector<unsigned char> chosen;
        for (int p = 0; p < P; p++) chosen.insert(chosen.end(), pts[p].mDescriptor.ptr(0), pts[p].mDescriptor.ptr(0) + 32);
        wr(d + "/distinctive.bin", chosen);
    }
    printf("host shim ok: %d/%d keypoints, %d matches, %d inliers\n", Last.N, Cur.N, nmatches, ninl);
    return 0;
}
