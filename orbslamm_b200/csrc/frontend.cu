// frontend.cu -- fused per-frame front end (orbf_*): ORB extraction -> SearchByProjection(Cur, Last) ->
// PoseOptimization for a batch of independent camera streams in ONE host-buffer call.  It composes the three public
// handles (orbx / orbm / orbo) on one CUDA stream: keypoints and descriptors never leave HBM between the stages, the
// image upload is chunked and overlapped with the extraction, and only what Tracking::TrackWithMotionModel
// (S/src/Tracking.cc:912-973) needs on the host comes back: the frame's keypoints / descriptors, the matches
// (Frame::mvpMapPoints), the optimised pose and the outlier flags (Frame::mvbOutlier).
#include <mutex>
#include "common.cuh"

using namespace orbs;

struct orbf_handle {
    orbx_handle *ex; orbm_handle *mt; orbo_handle *po;
    int device;
    cudaStream_t stream;
    std::mutex mu;
    StagePool pool;
};

extern "C" {

int orbf_create(orbf_handle **out, orbx_handle *ex, orbm_handle *mt, orbo_handle *po, int device)
{
    ORBS_REQUIRE(out && ex && mt && po, ORBS_E_INVALID, "null argument");
    *out = nullptr;
    ORBS_CUDA(cudaSetDevice(device));
    orbf_handle *h = new orbf_handle();
    h->ex = ex; h->mt = mt; h->po = po; h->device = device;
    h->stream = (cudaStream_t)orbx_stream(ex);
    int rc;
    if ((rc = orbm_set_stream(mt, h->stream)) || (rc = orbo_set_stream(po, h->stream))) { delete h; return rc; }
    *out = h;
    return ORBS_OK;
}

int orbf_destroy(orbf_handle *h)
{
    if (!h) return ORBS_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    h->pool.release();
    delete h;
    return ORBS_OK;
}

int orbf_track_frames(orbf_handle *h, const uint8_t *images, int n_frames, int width, int height, int stride, size_t frame_stride,
                      const float *K4, const float *scale_factors, const float *inv_level_sigma2, int nlevels,
                      const float *q_Xw, const int32_t *q_octave, const float *q_angle, const uint8_t *q_desc, const uint8_t *q_valid,
                      const int32_t *q_counts, int q_slab, float th_proj, int th_dist, int check_ori,
                      float *Tcw, float *kp_xy, float *kp_angle, float *kp_response, int32_t *kp_octave, float *kp_size, uint8_t *desc,
                      int cap, int32_t *counts, int32_t *feat_match, int32_t *nmatches, uint8_t *outlier, int32_t *n_inliers)
{
    ORBS_REQUIRE(h && images && K4 && scale_factors && inv_level_sigma2 && q_Xw && q_octave && q_angle && q_desc && q_valid && q_counts && Tcw &&
                 counts && feat_match && nmatches && outlier && n_inliers, ORBS_E_INVALID, "null argument");
    ORBS_REQUIRE(n_frames > 0 && width > 0 && height > 0 && q_slab > 0 && nlevels > 0, ORBS_E_INVALID, "non-positive size");
    std::lock_guard<std::mutex> lk(h->mu);
    ORBS_CUDA(cudaSetDevice(h->device));
    int slab = 0, rc;
    if ((rc = orbx_max_keypoints(h->ex, width, height, &slab))) return rc;
    ORBS_REQUIRE(cap >= slab, ORBS_E_CAPACITY, "output capacity must be at least orbx_max_keypoints()");
    cudaStream_t st = h->stream;
    // 1. extraction (chunked upload overlapped with compute; results stay on the device)
    if ((rc = orbx_extract_host_async(h->ex, images, n_frames, width, height, stride, frame_stride))) return rc;
    orbx_device_view v;
    if ((rc = orbx_device_results(h->ex, &v))) return rc;
    // 2. last-frame map points -> device
    Stager S(&h->pool, st, ORBS_MEM_HOST);
    const size_t nq = (size_t)n_frames * q_slab, nf = (size_t)n_frames * v.slab;
    const float *d_Xw = S.in(q_Xw, nq * 3); const int32_t *d_qoct = S.in(q_octave, nq); const float *d_qang = S.in(q_angle, nq);
    const uint8_t *d_qdesc = S.in(q_desc, nq * 32); const int32_t *d_qcnt = S.in(q_counts, n_frames);
    uint8_t *d_qvalid = const_cast<uint8_t *>(S.in(q_valid, nq));
    const float *d_sf = S.in(scale_factors, nlevels), *d_ils = S.in(inv_level_sigma2, nlevels);
    float *d_T = S.inout(Tcw, (size_t)n_frames * 16);
    float *d_uv = S.scratch<float>(nq * 2), *d_rad = S.scratch<float>(nq);
    int32_t *d_mn = S.scratch<int32_t>(nq), *d_mx = S.scratch<int32_t>(nq);
    int32_t *d_fm = S.scratch<int32_t>(nf), *d_nm = S.scratch<int32_t>(n_frames), *d_ninl = S.scratch<int32_t>(n_frames);
    uint8_t *d_out = S.scratch<uint8_t>(nf);
    if (S.rc) return S.rc;
    ORBS_CUDA(cudaMemsetAsync(d_fm, 0xff, nf * sizeof(int32_t), st));              // every feature free (-1)
    const float bounds[4] = {0.f, 0.f, (float)width, (float)height};                // Frame.cc:456-463 (no distortion)
    // 3. project -> search -> pose optimisation, all on device pointers
    if ((rc = orbm_project_last_frame(h->mt, n_frames, d_T, K4, bounds, d_sf, nlevels, d_Xw, d_qoct, d_qcnt, q_slab, th_proj, d_qvalid, d_uv, d_rad,
                                      d_mn, d_mx, ORBS_MEM_DEVICE))) return rc;
    if ((rc = orbm_search_by_projection(h->mt, n_frames, bounds, v.kp_xy, v.kp_octave, v.kp_angle, v.desc, v.counts, v.slab, d_qvalid, d_uv, d_rad,
                                        d_mn, d_mx, d_qang, d_qdesc, d_qcnt, q_slab, th_dist, 0.f, check_ori, d_fm, d_nm, ORBS_MEM_DEVICE))) return rc;
    if ((rc = orbo_pose_optimization_matched(h->po, n_frames, d_T, K4, v.kp_xy, v.kp_octave, v.counts, v.slab, d_fm, d_Xw, d_qcnt, q_slab, d_ils,
                                             nlevels, d_out, d_ninl, nullptr, ORBS_MEM_DEVICE))) return rc;
    // 4. results -> host
    if ((rc = orbx_download(h->ex, kp_xy, kp_angle, kp_response, kp_octave, kp_size, desc, cap, counts))) return rc;
    ORBS_CUDA(cudaMemcpy2DAsync(feat_match, (size_t)cap * 4, d_fm, (size_t)v.slab * 4, (size_t)v.slab * 4, n_frames, cudaMemcpyDeviceToHost, st));
    ORBS_CUDA(cudaMemcpy2DAsync(outlier, (size_t)cap, d_out, (size_t)v.slab, (size_t)v.slab, n_frames, cudaMemcpyDeviceToHost, st));
    ORBS_CUDA(cudaMemcpyAsync(nmatches, d_nm, n_frames * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    ORBS_CUDA(cudaMemcpyAsync(n_inliers, d_ninl, n_frames * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    return S.finish();          // copies Tcw back and synchronises
}

}  // extern "C"
