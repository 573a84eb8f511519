/*
 * oracle/ref_extractor_capi.cc -- C entry points around the REFERENCE's own ORBextractor class
 * (/root/reference/SingleRobotScenario/src/ORBextractor.cc, compiled unmodified against oracle/cvshim by
 * oracle/Makefile into oracle/_ref/libref_orbextractor.so).
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/test_oracle_vs_reference.py to validate the restatement in orb_oracle.c
 * against the reference's object code.  It does not exist on the GPU box unless oracle/_ref/ was built here.
 */
#include <cstring>
#include <vector>
#include "ORBextractor.h"

extern "C" {

void ref_bump_reset(void);

void *ref_orbx_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th)
{
    return new iORB_SLAM::ORBextractor(nfeatures, scale_factor, nlevels, ini_th, min_th);
}

void ref_orbx_destroy(void *h) { delete (iORB_SLAM::ORBextractor *)h; }

/* ORBextractor::operator() (ORBextractor.cc:1043-1105).  Returns the number of keypoints (<= cap are written). */
int ref_orbx_extract(void *h, const unsigned char *img, int w, int hgt, int stride, int cap,
                     float *x, float *y, float *angle, float *response, int *octave, float *size, unsigned char *desc)
{
    iORB_SLAM::ORBextractor &E = *(iORB_SLAM::ORBextractor *)h;
    cv::Mat image(hgt, w, CV_8UC1, (void *)img, (size_t)stride), mask, d;
    std::vector<cv::KeyPoint> kps;
    ref_bump_reset();
    E(image, mask, kps, d);
    const int n = (int)kps.size();
    for (int i = 0; i < n && i < cap; i++) {
        x[i] = kps[i].pt.x; y[i] = kps[i].pt.y; angle[i] = kps[i].angle; response[i] = kps[i].response;
        octave[i] = kps[i].octave; size[i] = kps[i].size;
        std::memcpy(desc + (size_t)32 * i, d.ptr(i), 32);
    }
    return n;
}

/* getters (ORBextractor.h:63-85) and the public pyramid */
int ref_orbx_tables(void *h, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2)
{
    iORB_SLAM::ORBextractor &E = *(iORB_SLAM::ORBextractor *)h;
    const int n = E.GetLevels();
    std::vector<float> a = E.GetScaleFactors(), b = E.GetInverseScaleFactors(), c = E.GetScaleSigmaSquares(), d = E.GetInverseScaleSigmaSquares();
    for (int i = 0; i < n; i++) { scale[i] = a[i]; inv_scale[i] = b[i]; sigma2[i] = c[i]; inv_sigma2[i] = d[i]; }
    return n;
}

int ref_orbx_pyramid_level(void *h, int level, int *w, int *hgt, unsigned char *out, int out_stride)
{
    iORB_SLAM::ORBextractor &E = *(iORB_SLAM::ORBextractor *)h;
    if (level < 0 || level >= (int)E.mvImagePyramid.size()) return -1;
    const cv::Mat &m = E.mvImagePyramid[level];
    *w = m.cols; *hgt = m.rows;
    if (out) for (int r = 0; r < m.rows; r++) std::memcpy(out + (size_t)r * out_stride, m.ptr(r), m.cols);
    return 0;
}

}
