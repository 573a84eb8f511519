// se3.cuh -- fp64 SE3 / projection-edge arithmetic of the g2o types the reference optimizer uses.
// Restates g2o::SE3Quat (types/se3quat.h) on Eigen::Quaterniond semantics, EdgeSE3ProjectXYZ(OnlyPose)
// (types/types_six_dof_expmap.{h,cpp}) and RobustKernelHuber (core/robust_kernel_impl.cpp:78-91).
#pragma once
#include <cuda_runtime.h>

namespace orbs {

struct Se3 { double q[4]; /* x y z w */ double t[3]; };

__host__ __device__ inline void quat_from_R(const double R[9], double q[4])        // Eigen::Quaterniond(Matrix3d)
{
    double t = R[0] + R[4] + R[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (R[7] - R[5]) * t; q[1] = (R[2] - R[6]) * t; q[2] = (R[3] - R[1]) * t;
    } else {
        // Eigen picks i = the largest diagonal entry, j = (i + 1) % 3, k = (j + 1) % 3; written out per case so that R and q are never indexed by a
        // runtime value (which would put them in local memory on the device -- this sits on the one-thread stretch of every LM trial)
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > (i == 1 ? R[4] : R[0])) i = 2;
        if (i == 0) {            // j = 1, k = 2
            t = sqrt(R[0] - R[4] - R[8] + 1.0);
            q[0] = 0.5 * t; t = 0.5 / t;
            q[3] = (R[7] - R[5]) * t; q[1] = (R[3] + R[1]) * t; q[2] = (R[6] + R[2]) * t;
        } else if (i == 1) {     // j = 2, k = 0
            t = sqrt(R[4] - R[8] - R[0] + 1.0);
            q[1] = 0.5 * t; t = 0.5 / t;
            q[3] = (R[2] - R[6]) * t; q[2] = (R[7] + R[5]) * t; q[0] = (R[1] + R[3]) * t;
        } else {                 // j = 0, k = 1
            t = sqrt(R[8] - R[0] - R[4] + 1.0);
            q[2] = 0.5 * t; t = 0.5 / t;
            q[3] = (R[3] - R[1]) * t; q[0] = (R[2] + R[6]) * t; q[1] = (R[5] + R[7]) * t;
        }
    }
}

__host__ __device__ inline void quat_normalize_pos(double q[4])                     // SE3Quat::normalizeRotation
{
    if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
    const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}

__host__ __device__ inline void quat_rotate(const double q[4], const double v[3], double out[3])   // Eigen _transformVector
{
    double uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    out[0] = v[0] + q[3] * uv[0] + (q[1] * uv[2] - q[2] * uv[1]);
    out[1] = v[1] + q[3] * uv[1] + (q[2] * uv[0] - q[0] * uv[2]);
    out[2] = v[2] + q[3] * uv[2] + (q[0] * uv[1] - q[1] * uv[0]);
}

__host__ __device__ inline void quat_mul(const double a[4], const double b[4], double o[4])
{
    o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    o[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    o[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}

__host__ __device__ inline void quat_to_R(const double q[4], double R[9])           // Eigen toRotationMatrix
{
    const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
    const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
    const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
    const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

__host__ __device__ inline void se3_from_Tcw(const float *T, Se3 &s)                // Converter::toSE3Quat, Converter.cc:37-47
{
    double R[9];
    #pragma unroll
    for (int i = 0; i < 9; i++) R[i] = (double)T[4 * (i / 3) + i % 3];
    quat_from_R(R, s.q);
    quat_normalize_pos(s.q);
    s.t[0] = T[3]; s.t[1] = T[7]; s.t[2] = T[11];
}

__host__ __device__ inline void se3_to_Tcw(const Se3 &s, float *T)                  // Converter::toCvMat(SE3Quat), Converter.cc:49-72
{
    double R[9];
    quat_to_R(s.q, R);
    #pragma unroll
    for (int i = 0; i < 9; i++) T[4 * (i / 3) + i % 3] = (float)R[i];
#pragma unroll
    for (int r = 0; r < 3; r++) T[4 * r + 3] = (float)s.t[r];
    T[12] = T[13] = T[14] = 0.f; T[15] = 1.f;
}

__host__ __device__ inline void se3_map(const Se3 &s, const double X[3], double out[3])
{
    quat_rotate(s.q, X, out);
    out[0] += s.t[0]; out[1] += s.t[1]; out[2] += s.t[2];
}

// SE3Quat::exp(update), update = (omega, upsilon); se3quat.h:223-257 (keeps the theta < 1e-5 branch R = I + W + W^2)
__host__ __device__ inline void se3_exp(const double u[6], Se3 &out)
{
    const double w[3] = {u[0], u[1], u[2]};
    const double theta = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    const double O[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double O2[9], R[9], V[9];
    #pragma unroll
    for (int i = 0; i < 9; i++) O2[i] = O[3 * (i / 3)] * O[i % 3] + O[3 * (i / 3) + 1] * O[3 + i % 3] + O[3 * (i / 3) + 2] * O[6 + i % 3];
    if (theta < 0.00001) {
        #pragma unroll
        for (int i = 0; i < 9; i++) { R[i] = (i % 4 == 0 ? 1.0 : 0.0) + O[i] + O2[i]; V[i] = R[i]; }
    } else {
        const double a = sin(theta) / theta, b = (1 - cos(theta)) / (theta * theta), c = (theta - sin(theta)) / pow(theta, 3.0);
        #pragma unroll
        for (int i = 0; i < 9; i++) {
            R[i] = (i % 4 == 0 ? 1.0 : 0.0) + a * O[i] + b * O2[i];
            V[i] = (i % 4 == 0 ? 1.0 : 0.0) + b * O[i] + c * O2[i];
        }
    }
    quat_from_R(R, out.q);
    #pragma unroll
    for (int r = 0; r < 3; r++) out.t[r] = V[3 * r] * u[3] + V[3 * r + 1] * u[4] + V[3 * r + 2] * u[5];
    quat_normalize_pos(out.q);
}

__host__ __device__ inline void se3_mul(const Se3 &a, const Se3 &b, Se3 &o)         // SE3Quat::operator*, se3quat.h:104-110
{
    double rt[3];
    quat_rotate(a.q, b.t, rt);
    Se3 r;
    r.t[0] = a.t[0] + rt[0]; r.t[1] = a.t[1] + rt[1]; r.t[2] = a.t[2] + rt[2];
    quat_mul(a.q, b.q, r.q);
    quat_normalize_pos(r.q);
    o = r;
}

// RobustKernelHuber::robustify -> (rho, rho')
__host__ __device__ inline void huber(double e, double delta, double dsqr, double &rho0, double &rho1)
{
    if (e <= dsqr) { rho0 = e; rho1 = 1.; }
    else { const double s = sqrt(e); rho0 = 2 * s * delta - dsqr; rho1 = delta / s; }
}

// reprojection error, obs - project(Xc)   (computeError, types_six_dof_expmap.h:90-95)
__host__ __device__ inline void reproj_error(const double Xc[3], const double in[4], double u, double v, double e[2])
{
    const double px = Xc[0] / Xc[2], py = Xc[1] / Xc[2];
    e[0] = u - (px * in[0] + in[2]);
    e[1] = v - (py * in[1] + in[3]);
}

// EdgeSE3ProjectXYZOnlyPose::linearizeOplus, types_six_dof_expmap.cpp:266-288
__host__ __device__ inline void jac_pose_only(const double Xc[3], double fx, double fy, double Jp[12])
{
    const double x = Xc[0], y = Xc[1], invz = 1.0 / Xc[2], invz_2 = invz * invz;
    Jp[0] = x * y * invz_2 * fx; Jp[1] = -(1 + (x * x * invz_2)) * fx; Jp[2] = y * invz * fx;
    Jp[3] = -invz * fx; Jp[4] = 0; Jp[5] = x * invz_2 * fx;
    Jp[6] = (1 + y * y * invz_2) * fy; Jp[7] = -x * y * invz_2 * fy; Jp[8] = -x * invz * fy;
    Jp[9] = 0; Jp[10] = -invz * fy; Jp[11] = y * invz_2 * fy;
}

// EdgeSE3ProjectXYZ::linearizeOplus, types_six_dof_expmap.cpp:103-139: Jl (2x3, point) and Jp (2x6, pose)
__host__ __device__ inline void jac_binary(const double Xc[3], const double R[9], double fx, double fy, double Jl[6], double Jp[12])
{
    const double x = Xc[0], y = Xc[1], z = Xc[2], z_2 = z * z;
    const double tmp[6] = {fx, 0, -x / z * fx, 0, fy, -y / z * fy};
    const double s = -1. / z;
    #pragma unroll
    for (int r = 0; r < 2; r++) {
        const double st0 = s * tmp[3 * r], st1 = s * tmp[3 * r + 1], st2 = s * tmp[3 * r + 2];
        #pragma unroll
        for (int c = 0; c < 3; c++) Jl[3 * r + c] = st0 * R[c] + st1 * R[3 + c] + st2 * R[6 + c];
    }
    Jp[0] = x * y / z_2 * fx; Jp[1] = -(1 + (x * x / z_2)) * fx; Jp[2] = y / z * fx;
    Jp[3] = -1. / z * fx; Jp[4] = 0; Jp[5] = x / z_2 * fx;
    Jp[6] = (1 + y * y / z_2) * fy; Jp[7] = -x * y / z_2 * fy; Jp[8] = -x / z * fy;
    Jp[9] = 0; Jp[10] = -1. / z * fy; Jp[11] = y / z_2 * fy;
}

// Matrix3d::inverse(): cofactors / determinant
__host__ __device__ inline void inv3(const double *A, double *I)
{
    const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
    const double det = A[0] * c00 + A[1] * c01 + A[2] * c02, id = 1.0 / det;
    I[0] = c00 * id; I[1] = (A[2] * A[7] - A[1] * A[8]) * id; I[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    I[3] = c01 * id; I[4] = (A[0] * A[8] - A[2] * A[6]) * id; I[5] = (A[2] * A[3] - A[0] * A[5]) * id;
    I[6] = c02 * id; I[7] = (A[1] * A[6] - A[0] * A[7]) * id; I[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}

}  // namespace orbs
