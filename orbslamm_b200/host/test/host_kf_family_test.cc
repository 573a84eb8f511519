// host_kf_family_test.cc -- drives the KeyFrame / Sim3 members of the C++ drop-in ORBmatcher (SearchByProjection(KF, Scw), Fuse x2, SearchBySim3,
// SearchByProjection(Frame, KF)) on binary inputs written by tests/test_host_shim_gpu.py and dumps what they did to the mock map.
// Usage: host_kf_family_test <dir>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>
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

static std::string D;
static std::vector<float> sf, inv2;

static void load_kf(KeyFrame &K, const std::string &tag)
{
    const std::vector<float> xy = rd<float>(D + "/" + tag + "_xy.bin"), ang = rd<float>(D + "/" + tag + "_angle.bin"), T = rd<float>(D + "/" + tag + "_Tcw.bin"),
                             K4 = rd<float>(D + "/K4.bin"), gb = rd<float>(D + "/grid_bounds.bin");
    const std::vector<int> oct = rd<int>(D + "/" + tag + "_octave.bin");
    const std::vector<unsigned char> desc = rd<unsigned char>(D + "/" + tag + "_desc.bin");
    K.N = (int)oct.size();
    K.mvKeysUn.resize(K.N);
    for (int i = 0; i < K.N; i++) K.mvKeysUn[i] = cv::KeyPoint(xy[2 * i], xy[2 * i + 1], 31.f, ang[i], 0.f, oct[i]);
    K.mDescriptors = cv::Mat(K.N, 32, CV_8U);
    memcpy(K.mDescriptors.data, desc.data(), desc.size());
    K.mvpMapPoints.assign(K.N, nullptr);
    K.fx = K4[0]; K.fy = K4[1]; K.cx = K4[2]; K.cy = K4[3];
    K.mvScaleFactors = sf; K.mvInvLevelSigma2 = inv2; K.mfLogScaleFactor = std::log(sf[1]);
    K.mnMinX = (int)gb[0]; K.mnMinY = (int)gb[1]; K.mnMaxX = (int)gb[2]; K.mnMaxY = (int)gb[3];       // KeyFrame::KeyFrame(Frame&): float -> int
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) K.Tcw.at<float>(r, c) = T[4 * r + c];
    Frame::fx = K4[0]; Frame::fy = K4[1]; Frame::cx = K4[2]; Frame::cy = K4[3];
    Frame::mnMinX = gb[0]; Frame::mnMinY = gb[1]; Frame::mnMaxX = gb[2]; Frame::mnMaxY = gb[3];
}

static void load_points(std::vector<MapPoint> &mps, const std::string &tag, long id0)
{
    const std::vector<float> X = rd<float>(D + "/" + tag + "_Xw.bin"), Nn = rd<float>(D + "/" + tag + "_normal.bin"), mn = rd<float>(D + "/" + tag + "_mfmin.bin"),
                             mx = rd<float>(D + "/" + tag + "_mfmax.bin");
    const std::vector<unsigned char> desc = rd<unsigned char>(D + "/" + tag + "_desc.bin");
    mps.assign(mn.size(), MapPoint());
    for (size_t i = 0; i < mps.size(); i++) {
        MapPoint &p = mps[i];
        p = MapPoint();
        p.mnId = id0 + (long)i;
        for (int k = 0; k < 3; k++) { p.mWorldPos.at<float>(k) = X[3 * i + k]; p.mNormalVector.at<float>(k) = Nn[3 * i + k]; }
        p.mfMinDistance = mn[i]; p.mfMaxDistance = mx[i];
        memcpy(p.mDescriptor.ptr(0), desc.data() + 32 * i, 32);
    }
}

static cv::Mat mat44(const std::vector<float> &T) { cv::Mat m(4, 4, CV_32F); memcpy(m.data, T.data(), 64); return m; }

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    D = argv[1];
    sf = rd<float>(D + "/sf.bin"); inv2 = rd<float>(D + "/inv2.bin");
    const std::vector<unsigned char> skip = rd<unsigned char>(D + "/skip.bin"), held = rd<unsigned char>(D + "/held.bin");
    const std::vector<float> Scw = rd<float>(D + "/Scw.bin");
    const int M = (int)skip.size();

    {   // SearchByProjection(pKF, Scw, vpPoints, vpMatched, 10)
        KeyFrame K; load_kf(K, "kf");
        std::vector<MapPoint> mps; load_points(mps, "pts", 0);
        MapPoint holder; holder.mnId = 1000000;
        std::vector<MapPoint *> vp(M), matched(K.N, nullptr);
        for (int i = 0; i < M; i++) { vp[i] = &mps[i]; mps[i].mbBad = skip[i] != 0; }
        for (int k = 0; k < K.N; k++) if (held[k]) matched[k] = &holder;
        ORBmatcher matcher(0.75, true);
        const int n = matcher.SearchByProjection(&K, mat44(Scw), vp, matched, 10);
        std::vector<int> out(K.N + 1, -1);
        for (int k = 0; k < K.N; k++) out[k] = matched[k] == &holder ? -2 : (matched[k] ? (int)matched[k]->mnId : -1);
        out[K.N] = n;
        wr(D + "/out_search_kf_sim3.bin", out);
    }
    for (int variant = 0; variant < 2; variant++) {   // Fuse(pKF, vpMapPoints, 3) and Fuse(pKF, Scw, vpPoints, 4, vpReplacePoint)
        KeyFrame K; load_kf(K, "kf");
        std::vector<MapPoint> mps; load_points(mps, "pts", 0);
        std::vector<MapPoint> holders(K.N);
        for (int k = 0; k < K.N; k++) { holders[k].mnId = 1000000 + k; holders[k].nObs = 1000; if (held[k]) { K.mvpMapPoints[k] = &holders[k]; holders[k].mObservations[&K] = k; } }
        std::vector<MapPoint *> vp(M), repl(M, nullptr);
        for (int i = 0; i < M; i++) { vp[i] = &mps[i]; mps[i].mbBad = skip[i] != 0; }
        ORBmatcher matcher(0.6, true);
        const int n = variant == 0 ? matcher.Fuse(&K, vp, 3.0f) : matcher.Fuse(&K, mat44(Scw), vp, 4.0f, repl);
        std::vector<int> out;
        for (int i = 0; i < M; i++) {                  // per point: bad, replaced-by id, index in the keyframe, vpReplacePoint id
            out.push_back(mps[i].mbBad ? 1 : 0); out.push_back(mps[i].mpReplaced ? (int)mps[i].mpReplaced->mnId : -1);
            out.push_back(mps[i].GetIndexInKeyFrame(&K)); out.push_back(repl[i] ? (int)repl[i]->mnId : -1);
        }
        for (int k = 0; k < K.N; k++) out.push_back(K.mvpMapPoints[k] ? (int)K.mvpMapPoints[k]->mnId : -1);
        out.push_back(n);
        wr(D + (variant == 0 ? "/out_fuse_kf.bin" : "/out_fuse_sim3.bin"), out);
    }
    {   // SearchByProjection(CurrentFrame, pKF, sAlreadyFound, 10, 100)
        KeyFrame Kc; load_kf(Kc, "kf");                 // the current frame has the keyframe case's features
        Frame F;
        F.N = Kc.N; F.mvKeysUn = Kc.mvKeysUn; F.mvKeys = Kc.mvKeysUn; F.mDescriptors = Kc.mDescriptors; F.mvuRight.assign(F.N, -1.f);
        F.mvpMapPoints.assign(F.N, nullptr); F.mvbOutlier.assign(F.N, false); F.mvScaleFactors = sf; F.mvInvLevelSigma2 = inv2; F.mfLogScaleFactor = std::log(sf[1]);
        F.mTcw = Kc.Tcw.clone();
        const std::vector<float> fb = rd<float>(D + "/frame_bounds.bin");
        Frame::mnMinX = fb[0]; Frame::mnMinY = fb[1]; Frame::mnMaxX = fb[2]; Frame::mnMaxY = fb[3];
        MapPoint holder; holder.mnId = 1000000;
        for (int k = 0; k < F.N; k++) if (held[k]) F.mvpMapPoints[k] = &holder;
        std::vector<MapPoint> mps; load_points(mps, "pts", 0);
        const std::vector<unsigned char> has = rd<unsigned char>(D + "/has.bin");
        const std::vector<float> kang = rd<float>(D + "/reloc_angle.bin");
        KeyFrame K; K.N = M; K.mvKeysUn.resize(M); K.mvpMapPoints.assign(M, nullptr);
        std::set<MapPoint *> found;
        for (int i = 0; i < M; i++) { K.mvKeysUn[i].angle = kang[i]; if (has[i]) K.mvpMapPoints[i] = &mps[i]; if (skip[i]) found.insert(&mps[i]); }
        ORBmatcher matcher(0.9, true);
        const int n = matcher.SearchByProjection(F, &K, found, 10.f, 100);
        std::vector<int> out(F.N + 1, -1);
        for (int k = 0; k < F.N; k++) out[k] = F.mvpMapPoints[k] == &holder ? -2 : (F.mvpMapPoints[k] ? (int)F.mvpMapPoints[k]->mnId : -1);
        out[F.N] = n;
        wr(D + "/out_search_frame_kf.bin", out);
    }
    {   // SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, 7.5)
        const std::vector<float> gb = rd<float>(D + "/s3_grid_bounds.bin");
        wr(D + "/grid_bounds.bin", gb);
        KeyFrame K1, K2; load_kf(K1, "s3_kf1"); load_kf(K2, "s3_kf2");
        std::vector<MapPoint> p1, p2; load_points(p1, "s3_pts1", 0); load_points(p2, "s3_pts2", 100000);
        const std::vector<unsigned char> has1 = rd<unsigned char>(D + "/s3_has1.bin"), has2 = rd<unsigned char>(D + "/s3_has2.bin");
        for (int i = 0; i < K1.N; i++) if (has1[i]) { K1.mvpMapPoints[i] = &p1[i]; p1[i].mObservations[&K1] = i; }
        for (int i = 0; i < K2.N; i++) if (has2[i]) { K2.mvpMapPoints[i] = &p2[i]; p2[i].mObservations[&K2] = i; }
        const std::vector<int> m12 = rd<int>(D + "/s3_m12.bin");
        std::vector<MapPoint *> vm(K1.N, nullptr);
        for (int i = 0; i < K1.N; i++) if (m12[i] >= 0) vm[i] = &p2[m12[i]];
        const std::vector<float> srt = rd<float>(D + "/s3_srt.bin");       // s12, R12 (9), t12 (3)
        cv::Mat R(3, 3, CV_32F), t(3, 1, CV_32F);
        memcpy(R.data, &srt[1], 36); memcpy(t.data, &srt[10], 12);
        ORBmatcher matcher(0.75, true);
        const float s12 = srt[0];
        const int n = matcher.SearchBySim3(&K1, &K2, vm, s12, R, t, 7.5f);
        std::vector<int> out(K1.N + 1, -1);
        for (int i = 0; i < K1.N; i++) out[i] = vm[i] ? (int)(vm[i]->mnId - 100000) : -1;
        out[K1.N] = n;
        wr(D + "/out_search_by_sim3.bin", out);
        // Optimizer::OptimizeSim3 on those matches (LoopClosing::ComputeSim3 runs the two back to back)
        const std::vector<double> s0 = rd<double>(D + "/s3_init.bin");
        g2o::Sim3 g;
        g.rotation().x() = s0[0]; g.rotation().y() = s0[1]; g.rotation().z() = s0[2]; g.rotation().w() = s0[3];
        g.translation()[0] = s0[4]; g.translation()[1] = s0[5]; g.translation()[2] = s0[6]; g.scale() = s0[7];
        const int nin = Optimizer::OptimizeSim3(&K1, &K2, vm, g, 10.f, false);
        std::vector<double> so = {g.rotation().x(), g.rotation().y(), g.rotation().z(), g.rotation().w(), g.translation()[0], g.translation()[1],
                                  g.translation()[2], g.scale(), (double)nin};
        for (int i = 0; i < K1.N; i++) so.push_back(vm[i] ? (double)(vm[i]->mnId - 100000) : -1.0);
        wr(D + "/out_optimize_sim3.bin", so);
    }
    {   // SearchByBoW x2, SearchForTriangulation on a keyframe pair with vocabulary nodes per feature
        wr(D + "/grid_bounds.bin", rd<float>(D + "/bow_grid_bounds.bin"));
        const std::vector<int> node1 = rd<int>(D + "/bow_node1.bin"), node2 = rd<int>(D + "/bow_node2.bin");
        const std::vector<unsigned char> has1 = rd<unsigned char>(D + "/bow_has1.bin"), has2 = rd<unsigned char>(D + "/bow_has2.bin"),
                                         tri1 = rd<unsigned char>(D + "/bow_tri1.bin"), tri2 = rd<unsigned char>(D + "/bow_tri2.bin");
        const std::vector<float> F12 = rd<float>(D + "/bow_F12.bin"), ls2 = rd<float>(D + "/bow_ls2.bin");
        for (int ori = 0; ori < 2; ori++) {
            KeyFrame K1, K2; load_kf(K1, "bow_kf1"); load_kf(K2, "bow_kf2");
            K1.mvLevelSigma2 = ls2; K2.mvLevelSigma2 = ls2; K1.mvuRight.assign(K1.N, -1.f); K2.mvuRight.assign(K2.N, -1.f);
            for (int i = 0; i < K1.N; i++) K1.mFeatVec.addFeature((unsigned)node1[i], (unsigned)i);
            for (int i = 0; i < K2.N; i++) K2.mFeatVec.addFeature((unsigned)node2[i], (unsigned)i);
            std::vector<MapPoint> p1(K1.N), p2(K2.N);
            for (int i = 0; i < K1.N; i++) { p1[i].mnId = i; if (has1[i]) K1.mvpMapPoints[i] = &p1[i]; }
            for (int i = 0; i < K2.N; i++) { p2[i].mnId = i; if (has2[i]) K2.mvpMapPoints[i] = &p2[i]; }
            std::vector<int> out;
            {   // keyframe - keyframe
                std::vector<MapPoint *> m12;
                ORBmatcher matcher(0.75, ori != 0);
                const int n = matcher.SearchByBoW(&K1, &K2, m12);
                for (int i = 0; i < K1.N; i++) out.push_back(m12[i] ? (int)m12[i]->mnId : -1);
                out.push_back(n);
            }
            {   // keyframe - frame (the frame has keyframe 2's features)
                Frame F;
                F.N = K2.N; F.mvKeys = K2.mvKeysUn; F.mvKeysUn = K2.mvKeysUn; F.mDescriptors = K2.mDescriptors; F.mFeatVec = K2.mFeatVec;
                std::vector<MapPoint *> mf;
                ORBmatcher matcher(0.7, ori != 0);
                const int n = matcher.SearchByBoW(&K1, F, mf);
                for (int k = 0; k < F.N; k++) out.push_back(mf[k] ? (int)mf[k]->mnId : -1);
                out.push_back(n);
            }
            {   // triangulation candidates: features without a map point
                for (int i = 0; i < K1.N; i++) K1.mvpMapPoints[i] = tri1[i] ? &p1[i] : nullptr;
                for (int i = 0; i < K2.N; i++) K2.mvpMapPoints[i] = tri2[i] ? &p2[i] : nullptr;
                cv::Mat F(3, 3, CV_32F);
                memcpy(F.data, F12.data(), 36);
                std::vector<std::pair<size_t, size_t>> pairs;
                ORBmatcher matcher(0.6, ori != 0);
                const int n = matcher.SearchForTriangulation(&K1, &K2, F, pairs, false);
                std::vector<int> m(K1.N, -1);
                for (auto &pr : pairs) m[pr.first] = (int)pr.second;
                out.insert(out.end(), m.begin(), m.end());
                out.push_back(n);
            }
            wr(D + (ori ? "/out_bow_ori1.bin" : "/out_bow_ori0.bin"), out);
        }
    }
    {   // SearchForInitialization, called twice like the initialiser does (the second call starts from the updated vbPrevMatched)
        wr(D + "/grid_bounds.bin", rd<float>(D + "/init_bounds.bin"));
        KeyFrame A, B; load_kf(A, "init_f1"); load_kf(B, "init_f2");
        Frame F1, F2;
        F1.N = A.N; F1.mvKeysUn = A.mvKeysUn; F1.mvKeys = A.mvKeysUn; F1.mDescriptors = A.mDescriptors;
        F2.N = B.N; F2.mvKeysUn = B.mvKeysUn; F2.mvKeys = B.mvKeysUn; F2.mDescriptors = B.mDescriptors;
        std::vector<cv::Point2f> prev(F1.N);
        for (int i = 0; i < F1.N; i++) prev[i] = F1.mvKeysUn[i].pt;
        std::vector<int> out, m12;
        ORBmatcher matcher(0.9, true);
        for (int call = 0; call < 2; call++) {
            const int n = matcher.SearchForInitialization(F1, F2, prev, m12, 100);
            out.insert(out.end(), m12.begin(), m12.end());
            out.push_back(n);
        }
        wr(D + "/out_init.bin", out);
    }
    printf("kf family host shim ok\n");
    return 0;
}
