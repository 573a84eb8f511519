// host_essential_graph_test.cc -- drives Optimizer::OptimizeEssentialGraph of the C++ drop-in on a mock map written by tests/test_host_shim_gpu.py
// (keyframe poses, spanning tree, covisibility weights, earlier loop edges, loop connections, corrected / non-corrected Sim3, map points) and dumps the
// corrected keyframe poses and map points.  Usage: host_essential_graph_test <dir>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>
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
template <typename T> void wr(const std::string &p, const std::vector<T> &v) { std::ofstream f(p, std::ios::binary); f.write((const char *)v.data(), v.size() * sizeof(T)); }

static g2o::Sim3 sim3_of(const double *s)
{
    g2o::Sim3 g;
    g.rotation().x() = s[0]; g.rotation().y() = s[1]; g.rotation().z() = s[2]; g.rotation().w() = s[3];
    g.translation()[0] = s[4]; g.translation()[1] = s[5]; g.translation()[2] = s[6]; g.scale() = s[7];
    return g;
}

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    const std::string d = argv[1];
    const std::vector<float> poses = rd<float>(d + "/eg_poses.bin"), pts = rd<float>(d + "/eg_points.bin");
    const std::vector<int> hdr = rd<int>(d + "/eg_hdr.bin"), parent = rd<int>(d + "/eg_parent.bin"), cov = rd<int>(d + "/eg_cov.bin"), loops = rd<int>(d + "/eg_loopedges.bin"),
                           conn = rd<int>(d + "/eg_loopconn.bin"), cidx = rd<int>(d + "/eg_corr_idx.bin"), pref = rd<int>(d + "/eg_point_ref.bin"),
                           pcorr = rd<int>(d + "/eg_point_corr.bin");
    const std::vector<double> corr = rd<double>(d + "/eg_corr.bin"), noncorr = rd<double>(d + "/eg_noncorr.bin");
    const int K = hdr[0], loopKF = hdr[1], curKF = hdr[2], fixScale = hdr[3];
    std::vector<KeyFrame> kfs(K);
    Map map;
    for (int k = 0; k < K; k++) {
        kfs[k].mnId = k;
        for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) kfs[k].Tcw.at<float>(r, c) = poses[16 * k + 4 * r + c];
        map.mvKFs.push_back(&kfs[k]);
    }
    for (int k = 0; k < K; k++) if (parent[k] >= 0) { kfs[k].mpParent = &kfs[parent[k]]; kfs[parent[k]].mspChildrens.insert(&kfs[k]); }
    for (size_t e = 0; e + 2 < cov.size() + 1 && e < cov.size(); e += 3) { kfs[cov[e]].mConnectedKeyFrameWeights[&kfs[cov[e + 1]]] = cov[e + 2]; kfs[cov[e + 1]].mConnectedKeyFrameWeights[&kfs[cov[e]]] = cov[e + 2]; }
    for (size_t e = 0; e < loops.size(); e += 2) { kfs[loops[e]].mspLoopEdges.insert(&kfs[loops[e + 1]]); kfs[loops[e + 1]].mspLoopEdges.insert(&kfs[loops[e]]); }
    std::map<KeyFrame *, std::set<KeyFrame *>> LoopConnections;
    for (size_t e = 0; e < conn.size(); e += 2) LoopConnections[&kfs[conn[e]]].insert(&kfs[conn[e + 1]]);
    LoopClosing::KeyFrameAndPose Corrected, NonCorrected;
    for (size_t i = 0; i < cidx.size(); i++) { Corrected[&kfs[cidx[i]]] = sim3_of(&corr[8 * i]); NonCorrected[&kfs[cidx[i]]] = sim3_of(&noncorr[8 * i]); }
    const int P = (int)pref.size();
    std::vector<MapPoint> mps(P);
    for (int p = 0; p < P; p++) {
        mps[p] = MapPoint();
        mps[p].mWorldPos = cv::Mat(3, 1, CV_32F);
        for (int c = 0; c < 3; c++) mps[p].mWorldPos.at<float>(c) = pts[3 * p + c];
        mps[p].mpRefKF = &kfs[pref[p]];
        if (pcorr[p] >= 0) { mps[p].mnCorrectedByKF = curKF; mps[p].mnCorrectedReference = pcorr[p]; }
        map.mvMPs.push_back(&mps[p]);
    }
    const bool bFixScale = fixScale != 0;
    Optimizer::OptimizeEssentialGraph(&map, &kfs[loopKF], &kfs[curKF], NonCorrected, Corrected, LoopConnections, bFixScale);
    std::vector<float> op, ox;
    for (int k = 0; k < K; k++) for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) op.push_back(kfs[k].Tcw.at<float>(r, c));
    for (int p = 0; p < P; p++) for (int c = 0; c < 3; c++) ox.push_back(mps[p].mWorldPos.at<float>(c));
    wr(d + "/eg_out_poses.bin", op); wr(d + "/eg_out_points.bin", ox);
    printf("essential graph host shim ok: %d keyframes, %d points\n", K, P);
    return 0;
}
