/*
 * ba_oracle.c — CPU restatement (fp64) of COLMAP's bundle-adjustment hot path.
 *
 * TEST INFRASTRUCTURE ONLY: nothing in the product may call, link or import this file.
 *
 * Follows, in the reference tree:
 *   residual + analytic Jacobians  src/colmap/estimators/cost_functions/reprojection_error.h:62-210
 *   quaternion rotation + Jacobian  src/colmap/estimators/cost_functions/quaternion_utils.h:105-153
 *   camera models                   src/colmap/sensor/models_jacobian.h:139-398, models.h:281-285 (depth guard)
 *   problem semantics               src/colmap/estimators/bundle_adjustment_ceres.cc:270-565,688-888
 *   solver options                  ibid. :102-116,202-212
 * and, for everything inside ceres::Solve — an EXTERNAL dependency that is not under /root/reference
 * (find_package(Ceres) without version; Ubuntu 24.04 docker image ships Ceres 2.2.0) — the published
 * algorithm of Ceres' TrustRegionMinimizer + LevenbergMarquardtStrategy + SchurEliminator +
 * ConjugateGradientsSolver/SCHUR_JACOBI and EigenQuaternionManifold, restated from its documentation:
 * Jacobi column scaling 1/(1+||J_i||), LM diagonal sqrt(clamp(diag(J'J),1e-6,1e32)/radius), step quality
 * rho = cost change / model cost change, radius /= max(1/3, 1-(2rho-1)^3) on success and /= 2,4,8.. on
 * failure, gradient max-norm convergence through the manifold Plus, CG termination on the Q-tolerance.
 *
 * PARITY STATUS: **parity unpinned** against Ceres iterates (no Ceres here, no test in the reference stores
 * Ceres outputs).  Pinned by the reference's own tests and checked in tests/test_ba_cpu.py: residual known
 * answers (reprojection_error_test.cc:41-71), analytic Jacobian == finite differences
 * (reprojection_error_test.cc:211-323), ground-truth recovery tolerances and residual counts of
 * bundle_adjustment_test.cc:303-411, constants bit-identical.
 *
 * Every loop over observations / points is OpenMP-parallel (per-thread partial sums for the camera-side quantities,
 * one thread per point for the point-side ones), because this file is also the CPU baseline bench.py times beside
 * the GPU path: a serial oracle would flatter the GPU numbers.
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC (Makefile).
 */
#define _GNU_SOURCE
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/b200_bundle_adjustment.h"

#define MAXDK 16             /* variable intrinsics of one camera (RAD_TAN_THIN_PRISM_FISHEYE has 16 parameters) */
#define JS 6                  /* column of the sensor_from_rig tangent block in a Jacobian row */
#define JI 12                 /* column of the intrinsics block */
#define JC (12 + MAXDK)       /* [pose 6 | sensor 6 | intrinsics] */

/* ------------------------------------------------------------------------- camera models */
static int wide_num_params(int id);
static int model_num_params(int id) {
  switch (id) { case 0: return 3; case 1: case 2: case 8: return 4; case 3: case 9: return 5; default: return wide_num_params(id); }
}
/* FisheyeProjectionWithJac (models_jacobian.h:44-83): (a, b) -> (atan r / r)(a, b); J = d(out) / d(a, b), row-major */
static void fisheye_with_jac(double a, double b, double* fa, double* fb, double J[4]) {
  const double r2 = a * a + b * b, r = sqrt(r2);
  if (r < 2.220446049250313e-16) { *fa = a; *fb = b; J[0] = 1; J[1] = 0; J[2] = 0; J[3] = 1; return; }
  const double theta = atan(r), s = theta / r, g = (r / (1.0 + r2) - theta) / (r2 * r);
  *fa = s * a; *fb = s * b;
  J[0] = s + a * a * g; J[1] = a * b * g; J[2] = J[1]; J[3] = s + b * b * g;
}
/* parameter groups (focal, principal point, extra): models.h:462-520 */
static int param_group(int id, int k) { /* 0 focal, 1 pp, 2 extra, 3 metadata (never refined) */
  if (id == 17) return 3;   /* EQUIRECTANGULAR {width, height}: sensor metadata (bundle_adjustment_ceres.cc:435-441) */
  const int single = (id == 0 || id == 2 || id == 3 || id == 8 || id == 9 || id == 12 || id == 14);   /* f, cx, cy, ... */
  if (single) return k == 0 ? 0 : (k < 3 ? 1 : 2);
  return k < 2 ? 0 : (k < 4 ? 1 : 2);                                                                /* fx, fy, cx, cy, ... */
}

/* ImgFromCamWithJac (models_jacobian.h:139-398).  Returns 0 if the depth guard fails. */
static int img_from_cam(int id, const double* q, double u, double v, double w, double* x, double* y, double* Jp /*2xP*/,
                        double* Juvw /*2x3*/) {
  if (!(w >= 2.220446049250313e-16)) return 0; /* HasProjectableDepth, check_cheirality = true */
  const double iw = 1.0 / w, uu = u * iw, vv = v * iw;
  if (id == 0) {
    const double f = q[0];
    *x = f * uu + q[1]; *y = f * vv + q[2];
    const double fi = f * iw;
    Juvw[0] = fi; Juvw[1] = 0; Juvw[2] = -fi * uu; Juvw[3] = 0; Juvw[4] = fi; Juvw[5] = -fi * vv;
    Jp[0] = uu; Jp[1] = 1; Jp[2] = 0; Jp[3] = vv; Jp[4] = 0; Jp[5] = 1;
  } else if (id == 1) {
    *x = q[0] * uu + q[2]; *y = q[1] * vv + q[3];
    Juvw[0] = q[0] * iw; Juvw[1] = 0; Juvw[2] = -q[0] * iw * uu; Juvw[3] = 0; Juvw[4] = q[1] * iw; Juvw[5] = -q[1] * iw * vv;
    Jp[0] = uu; Jp[1] = 0; Jp[2] = 1; Jp[3] = 0; Jp[4] = 0; Jp[5] = vv; Jp[6] = 0; Jp[7] = 1;
  } else if (id == 2) {
    const double f = q[0], k = q[3];
    const double uu2 = uu * uu, vv2 = vv * vv, r2 = uu2 + vv2, kr2 = k * r2, alpha = 1.0 + kr2;
    const double xd = alpha * uu, yd = alpha * vv;
    *x = f * xd + q[1]; *y = f * yd + q[2];
    const double two_k = 2.0 * k, fi = f * iw, beta = 1.0 + 3.0 * kr2, cross = two_k * uu * vv;
    Juvw[0] = fi * (alpha + two_k * uu2); Juvw[1] = fi * cross; Juvw[2] = -fi * uu * beta;
    Juvw[3] = fi * cross; Juvw[4] = fi * (alpha + two_k * vv2); Juvw[5] = -fi * vv * beta;
    Jp[0] = xd; Jp[1] = 1; Jp[2] = 0; Jp[3] = f * uu * r2; Jp[4] = yd; Jp[5] = 0; Jp[6] = 1; Jp[7] = f * vv * r2;
  } else if (id == 8 || id == 9) { /* SIMPLE_RADIAL_FISHEYE / RADIAL_FISHEYE (models_jacobian.h:725-851) */
    const int P = id == 8 ? 4 : 5;
    const double f = q[0], k1 = q[3], k2 = id == 9 ? q[4] : 0.0;
    double fa, fb, Jf[4];
    fisheye_with_jac(uu, vv, &fa, &fb, Jf);
    const double t2 = fa * fa + fb * fb, t4 = t2 * t2, radial = k1 * t2 + k2 * t4;
    const double xd = fa + fa * radial, yd = fb + fb * radial;
    *x = f * xd + q[1]; *y = f * yd + q[2];
    const double dr = k1 + 2.0 * k2 * t2;
    const double D[4] = {1.0 + radial + 2.0 * fa * fa * dr, 2.0 * fa * fb * dr, 2.0 * fa * fb * dr, 1.0 + radial + 2.0 * fb * fb * dr};
    double Jab[4];
    for (int r = 0; r < 2; ++r) for (int c = 0; c < 2; ++c) Jab[2 * r + c] = f * (D[2 * r] * Jf[c] + D[2 * r + 1] * Jf[2 + c]);
    for (int r = 0; r < 2; ++r) {
      Juvw[3 * r] = Jab[2 * r] * iw; Juvw[3 * r + 1] = Jab[2 * r + 1] * iw;
      Juvw[3 * r + 2] = -(Jab[2 * r] * uu + Jab[2 * r + 1] * vv) * iw;
    }
    Jp[0] = xd; Jp[1] = 1; Jp[2] = 0; Jp[3] = f * fa * t2;
    Jp[P] = yd; Jp[P + 1] = 0; Jp[P + 2] = 1; Jp[P + 3] = f * fb * t2;
    if (id == 9) { Jp[4] = f * fa * t4; Jp[P + 4] = f * fb * t4; }
  } else {
    const double f = q[0], k1 = q[3], k2 = q[4];
    const double uu2 = uu * uu, vv2 = vv * vv, r2 = uu2 + vv2, r4 = r2 * r2, radial = k1 * r2 + k2 * r4;
    const double xd = uu * (1.0 + radial), yd = vv * (1.0 + radial);
    *x = f * xd + q[1]; *y = f * yd + q[2];
    const double dr = k1 + 2.0 * k2 * r2, cross = 2.0 * uu * vv * dr;
    const double a00 = f * (1.0 + radial + 2.0 * uu2 * dr), a01 = f * cross, a10 = f * cross,
                 a11 = f * (1.0 + radial + 2.0 * vv2 * dr);
    Juvw[0] = a00 * iw; Juvw[1] = a01 * iw; Juvw[2] = -(a00 * uu + a01 * vv) * iw;
    Juvw[3] = a10 * iw; Juvw[4] = a11 * iw; Juvw[5] = -(a10 * uu + a11 * vv) * iw;
    Jp[0] = xd; Jp[1] = 1; Jp[2] = 0; Jp[3] = f * uu * r2; Jp[4] = f * uu * r4;
    Jp[5] = yd; Jp[6] = 0; Jp[7] = 1; Jp[8] = f * vv * r2; Jp[9] = f * vv * r4;
  }
  return 1;
}

/* ---- camera models with more than five parameters (groundwork for SURVEY 8f-4): ImgFromCam of OPENCV, OPENCV_FISHEYE,
 * FULL_OPENCV, FOV, THIN_PRISM_FISHEYE (sensor/models.h:1513-1885, 2180-2238) written over complex numbers; the Jacobians
 * come from the complex-step method  df/dx = Im f(x + ih) / h  (exact to rounding for analytic f, h = 1e-40) - an
 * independent route from the product's dual numbers (colmap_b200/csrc/ba_models.cuh) and from the reference's
 * hand-derived formulas (sensor/models_jacobian.h:401-1565). */
typedef double complex cplx;
static int wide_num_params(int id) {
  switch (id) { case 4: case 5: return 8; case 6: case 10: return 12; case 7: case 13: return 5; case 11: return 16; case 12: case 15: return 4;
                case 14: return 3; case 16: return 6; case 17: return 2; default: return -1; }
}
static void c_fisheye(cplx a, cplx b, cplx* fa, cplx* fb) {
  const cplx r = csqrt(a * a + b * b);
  if (creal(r) > 2.220446049250313e-16) { const cplx s = catan(r) / r; *fa = a * s; *fb = b * s; } else { *fa = a; *fb = b; }
}
static int c_project_wide(int id, const cplx* q, cplx u, cplx v, cplx w, cplx* x, cplx* y) {
  if (id == 12 || id == 13) { /* division models (models.h:2406-2530) */
    const cplx k = q[id == 12 ? 3 : 4], dsq = w * w - 4.0 * k * (u * u + v * v);
    if (creal(dsq) < 0) return 0;
    const cplx r = 2.0 / (w + csqrt(dsq));
    *x = q[0] * r * u + q[id == 12 ? 1 : 2]; *y = q[id == 12 ? 0 : 1] * r * v + q[id == 12 ? 2 : 3];
    return 1;
  }
  if (!(creal(w) >= 2.220446049250313e-16)) return 0;
  if (id == 16) { /* EUCM (models.h:2758-2792) */
    const cplx rho2 = q[5] * (u * u + v * v) + w * w;
    if (creal(rho2) < 0) return 0;
    const cplx den = q[4] * csqrt(rho2) + (1.0 - q[4]) * w;
    if (!(creal(den) >= 2.220446049250313e-16)) return 0;
    *x = q[0] * u / den + q[2]; *y = q[1] * v / den + q[3];
    return 1;
  }
  const cplx a = u / w, b = v / w;
  if (id == 14 || id == 15) { /* (SIMPLE_)FISHEYE (models.h:2608-2720) */
    cplx fa, fb; c_fisheye(a, b, &fa, &fb);
    *x = q[0] * fa + q[id == 14 ? 1 : 2]; *y = q[id == 14 ? 0 : 1] * fb + q[id == 14 ? 2 : 3];
    return 1;
  }
  if (id == 11) { /* RAD_TAN_THIN_PRISM_FISHEYE (models.h:2300-2376) */
    cplx fa, fb; c_fisheye(a, b, &fa, &fb);
    const cplx t2 = fa * fa + fb * fb;
    const cplx radial = 1.0 + t2 * (q[4] + t2 * (q[5] + t2 * (q[6] + t2 * (q[7] + t2 * (q[8] + t2 * q[9])))));
    const cplx px = radial * fa, py = radial * fb, r2 = px * px + py * py;
    *x = q[0] * (px + 2.0 * q[11] * px * py + q[10] * (r2 + 2.0 * px * px) + q[12] * r2 + q[13] * r2 * r2) + q[2];
    *y = q[1] * (py + 2.0 * q[10] * px * py + q[11] * (r2 + 2.0 * py * py) + q[14] * r2 + q[15] * r2 * r2) + q[3];
    return 1;
  }
  if (id == 4) {
    const cplx r2 = a * a + b * b, radial = q[4] * r2 + q[5] * r2 * r2;
    *x = q[0] * (a + a * radial + 2.0 * q[6] * a * b + q[7] * (r2 + 2.0 * a * a)) + q[2];
    *y = q[1] * (b + b * radial + 2.0 * q[7] * a * b + q[6] * (r2 + 2.0 * b * b)) + q[3];
  } else if (id == 5) {
    cplx fa, fb; c_fisheye(a, b, &fa, &fb);
    const cplx t2 = fa * fa + fb * fb;
    const cplx radial = t2 * (q[4] + t2 * (q[5] + t2 * (q[6] + t2 * q[7])));
    *x = q[0] * fa * (1.0 + radial) + q[2]; *y = q[1] * fb * (1.0 + radial) + q[3];
  } else if (id == 6) {
    const cplx r2 = a * a + b * b;
    const cplx radial = (1.0 + r2 * (q[4] + r2 * (q[5] + r2 * q[8]))) / (1.0 + r2 * (q[9] + r2 * (q[10] + r2 * q[11])));
    *x = q[0] * (a * radial + 2.0 * q[6] * a * b + q[7] * (r2 + 2.0 * a * a)) + q[2];
    *y = q[1] * (b * radial + 2.0 * q[7] * a * b + q[6] * (r2 + 2.0 * b * b)) + q[3];
  } else if (id == 7) {
    const cplx om = q[4], r2 = a * a + b * b, o2 = om * om;
    cplx f;
    if (creal(o2) < 1e-4) f = o2 * r2 / 3.0 - o2 / 12.0 + 1.0;
    else if (creal(r2) < 1e-4) { const cplx t = ctan(om / 2.0); f = -2.0 * t * (4.0 * r2 * t * t - 3.0) / (3.0 * om); }
    else { const cplx r = csqrt(r2); f = catan(r * 2.0 * ctan(om / 2.0)) / (r * om); }
    *x = q[0] * a * f + q[2]; *y = q[1] * b * f + q[3];
  } else if (id == 10) {
    cplx fa, fb; c_fisheye(a, b, &fa, &fb);
    const cplx r2 = fa * fa + fb * fb;
    const cplx radial = r2 * (q[4] + r2 * (q[5] + r2 * (q[8] + r2 * q[9])));
    *x = q[0] * (fa + fa * radial + 2.0 * q[6] * fa * fb + q[7] * (r2 + 2.0 * fa * fa) + q[10] * r2) + q[2];
    *y = q[1] * (fb + fb * radial + 2.0 * q[7] * fa * fb + q[6] * (r2 + 2.0 * fb * fb) + q[11] * r2) + q[3];
  } else return 0;
  return 1;
}
/* EQUIRECTANGULAR (models.h:2853-2875) uses atan2, which has no complex-step form: value in real arithmetic, Jacobians by
 * Richardson-extrapolated central differences */
static int equirect(const double* q, const double* p, double* xy) {
  const double hor = sqrt(p[0] * p[0] + p[2] * p[2]);
  if (hor + fabs(p[1]) < 2.220446049250313e-16) return 0;
  xy[0] = (atan2(p[0], p[2]) / (2.0 * M_PI) + 0.5) * q[0];
  xy[1] = (0.5 - atan2(-p[1], hor) / M_PI) * q[1];
  return 1;
}
int ba_oracle_project_wide(int id, const double* params, const double* uvw, double* xy, double* J_uvw, double* J_params) {
  const int P = wide_num_params(id);
  if (P < 0) return -1;
  if (id == 17) {
    if (!equirect(params, uvw, xy)) return 0;
    double in[5] = {uvw[0], uvw[1], uvw[2], params[0], params[1]};
    for (int k = 0; k < 5; ++k) {
      double d[2][2];
      for (int lvl = 0; lvl < 2; ++lvl) {
        const double h = (lvl ? 0.5e-4 : 1e-4) * fmax(1.0, fabs(in[k])), keep = in[k];
        double a[2], b[2];
        in[k] = keep + h; equirect(in + 3, in, a);
        in[k] = keep - h; equirect(in + 3, in, b);
        in[k] = keep;
        d[lvl][0] = (a[0] - b[0]) / (2 * h); d[lvl][1] = (a[1] - b[1]) / (2 * h);
      }
      for (int r = 0; r < 2; ++r) {
        const double v = (4.0 * d[1][r] - d[0][r]) / 3.0;
        if (k < 3) J_uvw[3 * r + k] = v; else J_params[2 * r + k - 3] = v;
      }
    }
    return 1;
  }
  const double h = 1e-40;
  cplx in[3 + 16], x, y;
  for (int k = 0; k < 3; ++k) in[k] = uvw[k];
  for (int k = 0; k < P; ++k) in[3 + k] = params[k];
  if (!c_project_wide(id, in + 3, in[0], in[1], in[2], &x, &y)) return 0;
  xy[0] = creal(x); xy[1] = creal(y);
  for (int k = 0; k < 3 + P; ++k) {
    const cplx keep = in[k];
    in[k] = keep + h * I;
    c_project_wide(id, in + 3, in[0], in[1], in[2], &x, &y);
    in[k] = keep;
    if (k < 3) { J_uvw[k] = cimag(x) / h; J_uvw[3 + k] = cimag(y) / h; }
    else { J_params[k - 3] = cimag(x) / h; J_params[P + k - 3] = cimag(y) / h; }
  }
  return 1;
}

/* QuaternionRotatePointWithJac (quaternion_utils.h:105-153), q = (x,y,z,w) */
static void quat_rotate_jac(const double* q, const double* p, double out[3], double J[12]) {
  const double qx = q[0], qy = q[1], qz = q[2], qw = q[3], px = p[0], py = p[1], pz = p[2];
  const double qx_py = qx * py, qx_pz = qx * pz, qy_px = qy * px, qy_pz = qy * pz, qz_px = qz * px, qz_py = qz * py;
  const double c0 = qy_pz - qz_py, c1 = qz_px - qx_pz, c2 = qx_py - qy_px;
  const double d0 = qy * c2 - qz * c1, d1 = qz * c0 - qx * c2, d2 = qx * c1 - qy * c0;
  out[0] = px + 2.0 * (qw * c0 + d0); out[1] = py + 2.0 * (qw * c1 + d1); out[2] = pz + 2.0 * (qw * c2 + d2);
  if (J) {
    const double qx_px = qx * px, qy_py = qy * py, qz_pz = qz * pz, qw_px = qw * px, qw_py = qw * py, qw_pz = qw * pz;
    J[0] = 2.0 * (qy_py + qz_pz); J[1] = 2.0 * (-2.0 * qy_px + qx_py + qw_pz); J[2] = 2.0 * (-2.0 * qz_px - qw_py + qx_pz); J[3] = 2.0 * (-qz_py + qy_pz);
    J[4] = 2.0 * (qy_px - 2.0 * qx_py - qw_pz); J[5] = 2.0 * (qx_px + qz_pz); J[6] = 2.0 * (qw_px - 2.0 * qz_py + qy_pz); J[7] = 2.0 * (qz_px - qx_pz);
    J[8] = 2.0 * (qz_px + qw_py - 2.0 * qx_pz); J[9] = 2.0 * (-qw_px + qz_py - 2.0 * qy_pz); J[10] = 2.0 * (qx_px + qy_py); J[11] = 2.0 * (-qy_px + qx_py);
  }
}
static void quat_to_R(const double* q, double R[9]) { /* Eigen::Quaterniond::toRotationMatrix */
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z, twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x,
               tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
/* EigenQuaternionManifold::Plus: q_new = [sin|d| d/|d|, cos|d|] * q */
static void quat_plus(const double* q, const double* d, double* out) {
  const double n = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (n == 0.0) { memcpy(out, q, 32); return; }
  const double s = sin(n) / n, dw = cos(n), dx = s * d[0], dy = s * d[1], dz = s * d[2];
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  out[0] = dw * x + dx * w + dy * z - dz * y;
  out[1] = dw * y - dx * z + dy * w + dz * x;
  out[2] = dw * z + dx * y - dy * x + dz * w;
  out[3] = dw * w - dx * x - dy * y - dz * z;
}

/* ------------------------------------------------------------------------- exported building blocks */
/* AnalyticalReprojErrorCostFunction::Evaluate (reprojection_error.h:69-135): residual (2), J_point 2x3,
 * J_pose 2x7 (ambient), J_params 2xP; returns 0 when the depth guard zeroes everything. */
int ba_oracle_reproj(int model_id, const double* point, const double* pose, const double* params, const double* xy,
                     double* res, double* J_point, double* J_pose, double* J_params) {
  double pc[3], Jq[12], Juvw[6], Jp[10];
  const int P = model_num_params(model_id);
  quat_rotate_jac(pose, point, pc, Jq);
  pc[0] += pose[4]; pc[1] += pose[5]; pc[2] += pose[6];
  double x, y;
  if (!img_from_cam(model_id, params, pc[0], pc[1], pc[2], &x, &y, Jp, Juvw)) {
    res[0] = res[1] = 0;
    if (J_point) memset(J_point, 0, 48);
    if (J_pose) memset(J_pose, 0, 112);
    if (J_params) memset(J_params, 0, 16 * P);
    return 0;
  }
  res[0] = x - xy[0]; res[1] = y - xy[1];
  if (J_point) {
    double R[9]; quat_to_R(pose, R);
    for (int r = 0; r < 2; ++r)
      for (int c = 0; c < 3; ++c) J_point[3 * r + c] = Juvw[3 * r] * R[c] + Juvw[3 * r + 1] * R[3 + c] + Juvw[3 * r + 2] * R[6 + c];
  }
  if (J_pose) {
    for (int r = 0; r < 2; ++r) {
      for (int c = 0; c < 4; ++c) J_pose[7 * r + c] = Juvw[3 * r] * Jq[c] + Juvw[3 * r + 1] * Jq[4 + c] + Juvw[3 * r + 2] * Jq[8 + c];
      for (int c = 0; c < 3; ++c) J_pose[7 * r + 4 + c] = Juvw[3 * r + c];
    }
  }
  if (J_params) memcpy(J_params, Jp, 16 * P);
  return 1;
}
void ba_oracle_quat_plus(const double* q, const double* d, double* out) { quat_plus(q, d, out); }


/* ImgFromCamWithJac of ANY model: hand-written Jacobians for the <= 5-parameter family, complex step for the others */
static int img_from_cam_any(int id, const double* q, double u, double v, double w, double* x, double* y, double* Jp /*2xP*/,
                            double* Juvw /*2x3*/) {
  if (wide_num_params(id) < 0) return img_from_cam(id, q, u, v, w, x, y, Jp, Juvw);
  const double uvw[3] = {u, v, w};
  double xy[2];
  if (ba_oracle_project_wide(id, q, uvw, xy, Juvw, Jp) != 1) return 0;
  *x = xy[0]; *y = xy[1];
  return 1;
}
/* RigReprojErrorCostFunctor / ReprojErrorCostFunctor with analytic derivatives (reprojection_error.h:62-135,344-420):
 * p_cam = R(q_s) (R(q_r) X + t_r) + t_s with sensor == NULL meaning the identity (trivial frame / reference sensor).
 * J_rig, J_sensor: 2x7 ambient (qx qy qz qw tx ty tz), J_params 2xP, J_point 2x3.  Returns 0 behind the camera. */
int ba_oracle_reproj_rig(int model_id, const double* point, const double* rig, const double* sensor, const double* params,
                         const double* xy, double* res, double* J_point, double* J_rig, double* J_sensor, double* J_params) {
  const int P = model_num_params(model_id);
  double pr[3], Jqr[12], pc[3], Jqs[12], Rs[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Rr[9];
  quat_rotate_jac(rig, point, pr, Jqr);
  pr[0] += rig[4]; pr[1] += rig[5]; pr[2] += rig[6];
  if (sensor) {
    quat_rotate_jac(sensor, pr, pc, Jqs);
    pc[0] += sensor[4]; pc[1] += sensor[5]; pc[2] += sensor[6];
    quat_to_R(sensor, Rs);
  } else { pc[0] = pr[0]; pc[1] = pr[1]; pc[2] = pr[2]; memset(Jqs, 0, sizeof(Jqs)); }
  double x, y, Juvw[6], Jp[2 * MAXDK];
  if (!img_from_cam_any(model_id, params, pc[0], pc[1], pc[2], &x, &y, Jp, Juvw)) {
    res[0] = res[1] = 0;
    if (J_point) memset(J_point, 0, 48);
    if (J_rig) memset(J_rig, 0, 112);
    if (J_sensor) memset(J_sensor, 0, 112);
    if (J_params) memset(J_params, 0, 16 * P);
    return 0;
  }
  res[0] = x - xy[0]; res[1] = y - xy[1];
  double A[6];   /* d(x, y) / d(p_rig) = Juvw * R_s */
  for (int r = 0; r < 2; ++r) for (int c = 0; c < 3; ++c) A[3 * r + c] = Juvw[3 * r] * Rs[c] + Juvw[3 * r + 1] * Rs[3 + c] + Juvw[3 * r + 2] * Rs[6 + c];
  quat_to_R(rig, Rr);
  if (J_point) for (int r = 0; r < 2; ++r) for (int c = 0; c < 3; ++c) J_point[3 * r + c] = A[3 * r] * Rr[c] + A[3 * r + 1] * Rr[3 + c] + A[3 * r + 2] * Rr[6 + c];
  if (J_rig) for (int r = 0; r < 2; ++r) {
    for (int c = 0; c < 4; ++c) J_rig[7 * r + c] = A[3 * r] * Jqr[c] + A[3 * r + 1] * Jqr[4 + c] + A[3 * r + 2] * Jqr[8 + c];
    for (int c = 0; c < 3; ++c) J_rig[7 * r + 4 + c] = A[3 * r + c];
  }
  if (J_sensor) for (int r = 0; r < 2; ++r) {
    for (int c = 0; c < 4; ++c) J_sensor[7 * r + c] = sensor ? Juvw[3 * r] * Jqs[c] + Juvw[3 * r + 1] * Jqs[4 + c] + Juvw[3 * r + 2] * Jqs[8 + c] : 0.0;
    for (int c = 0; c < 3; ++c) J_sensor[7 * r + 4 + c] = sensor ? Juvw[3 * r + c] : 0.0;
  }
  if (J_params) memcpy(J_params, Jp, 16 * P);
  return 1;
}

/* ------------------------------------------------------------------------- loss (Ceres LossFunction + Corrector) */
static void loss_eval(int type, double a, double s, double rho[3]) {
  const double b = a * a, c = 1.0 / b;
  if (type == B200BA_LOSS_SOFT_L1) {
    const double sum = 1.0 + s * c, tmp = sqrt(sum);
    rho[0] = 2.0 * b * (tmp - 1.0); rho[1] = fmax(2.2250738585072014e-308, 1.0 / tmp); rho[2] = -(c * rho[1]) / (2.0 * sum);
  } else if (type == B200BA_LOSS_CAUCHY) {
    const double sum = 1.0 + s * c, inv = 1.0 / sum;
    rho[0] = b * log(sum); rho[1] = fmax(2.2250738585072014e-308, inv); rho[2] = -c * (inv * inv);
  } else if (type == B200BA_LOSS_HUBER) {
    if (s > b) { const double r = sqrt(s); rho[0] = 2.0 * a * r - b; rho[1] = fmax(2.2250738585072014e-308, a / r); rho[2] = -rho[1] / (2.0 * s); }
    else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
  } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
}

/* ------------------------------------------------------------------------- flattened problem */
typedef struct {
  const b200ba_options* o;
  b200ba_problem* p;
  int nvp;            /* variable poses */
  int* pose_off;      /* [num_poses] camera-side offset or -1 */
  uint8_t* pose_mask; /* [num_poses] 6-bit active tangent dims */
  int* cam_off;       /* [num_cameras] or -1 */
  int* sens_off;      /* [num_sensors] camera-side offset of the sensor_from_rig block or -1 */
  int* cam_nvar;
  int (*cam_var)[MAXDK];
  int* pt_var;        /* [num_points] index among variable points or -1 */
  int64_t nvpt;
  int nc;             /* camera-side dimension */
  int64_t nobs;       /* effective observations */
  int64_t* obs;       /* original observation ids, sorted by (variable point index, then constant-point obs) */
  int64_t* pt_start;  /* [nvpt+1] ranges into obs */
  int64_t nobs_var;   /* observations of variable points (first in obs) */
  /* linearisation */
  double *r, *Jc, *Jp; /* per effective obs: r[2], Jc[2*JC], Jp[6] (already loss-corrected, unscaled) */
  double* cost_obs;    /* per effective obs: 1/2 rho at the linearisation point */
  double cost_delta;   /* sum over obs of (candidate cost - linearisation cost), filled by linearize(want_jac=0) */
  double* scale_c;    /* [nc] jacobi scaling */
  double* scale_p;    /* [3*nvpt] */
} ba_flat;

static int cmp_obs(const void* a, const void* b, void* ctx) {
  const ba_flat* F = (const ba_flat*)ctx;
  const int64_t oa = *(const int64_t*)a, ob = *(const int64_t*)b;
  const int pa = F->pt_var[F->p->obs_point_idx[oa]], pb = F->pt_var[F->p->obs_point_idx[ob]];
  const int64_t ka = pa < 0 ? INT64_MAX : pa, kb = pb < 0 ? INT64_MAX : pb;
  if (ka != kb) return ka < kb ? -1 : 1;
  return oa < ob ? -1 : (oa > ob ? 1 : 0);
}

static int flatten(ba_flat* F) {
  const b200ba_options* o = F->o;
  b200ba_problem* p = F->p;
  F->pose_off = (int*)malloc(sizeof(int) * p->num_poses);
  F->pose_mask = (uint8_t*)malloc(p->num_poses);
  F->cam_off = (int*)malloc(sizeof(int) * p->num_cameras);
  F->sens_off = (int*)malloc(sizeof(int) * (p->num_sensors + 1));
  F->cam_nvar = (int*)calloc(p->num_cameras, sizeof(int));
  F->cam_var = (int(*)[MAXDK])calloc(p->num_cameras, sizeof(int[MAXDK]));
  F->pt_var = (int*)malloc(sizeof(int) * p->num_points);
  /* which blocks appear in at least one observation */
  uint8_t* pose_used = (uint8_t*)calloc(p->num_poses, 1);
  uint8_t* cam_used = (uint8_t*)calloc(p->num_cameras, 1);
  uint8_t* pt_used = (uint8_t*)calloc(p->num_points, 1);
  for (int64_t i = 0; i < p->num_observations; ++i) {
    pose_used[p->obs_pose_idx[i]] = 1; cam_used[p->obs_camera_idx[i]] = 1; pt_used[p->obs_point_idx[i]] = 1;
  }
  int off = 0;
  F->nvp = 0;
  for (int i = 0; i < p->num_poses; ++i) {
    const int cst = !o->refine_rig_from_world || (p->pose_constant && p->pose_constant[i]) || !pose_used[i];
    if (cst) { F->pose_off[i] = -1; F->pose_mask[i] = 0; continue; }
    uint8_t m = 0x3f;
    if (o->constant_rig_from_world_rotation) m &= ~0x07;
    if (p->pose_fixed_translation_dim && p->pose_fixed_translation_dim[i] >= 0) m &= ~(1u << (3 + p->pose_fixed_translation_dim[i]));
    F->pose_off[i] = off; F->pose_mask[i] = m; off += 6; F->nvp++;
  }
  {  /* sensor_from_rig blocks (ParameterizeRigsAndFrames, bundle_adjustment_ceres.cc:478-539): variable iff refined,
        not constant, and its camera is observed */
    uint8_t* sens_used = (uint8_t*)calloc(p->num_sensors + 1, 1);
    for (int c = 0; c < p->num_cameras; ++c) if (cam_used[c] && p->camera_sensor_idx && p->camera_sensor_idx[c] >= 0) sens_used[p->camera_sensor_idx[c]] = 1;
    for (int k = 0; k < p->num_sensors; ++k) {
      const int cst = !o->refine_sensor_from_rig || (p->sensor_constant && p->sensor_constant[k]) || !sens_used[k];
      F->sens_off[k] = cst ? -1 : off;
      if (!cst) off += 6;
    }
    free(sens_used);
  }
  for (int c = 0; c < p->num_cameras; ++c) {
    const int id = p->camera_model_id[c], P = model_num_params(id);
    if (P < 0) return -2;
    int nv = 0;
    if (!(p->camera_constant && p->camera_constant[c]) && cam_used[c])
      for (int k = 0; k < P; ++k) {
        const int g = param_group(id, k);
        const int refine = g == 0 ? o->refine_focal_length : (g == 1 ? o->refine_principal_point : (g == 2 ? o->refine_extra_params : 0));
        if (refine) F->cam_var[c][nv++] = k;
      }
    F->cam_nvar[c] = nv;
    F->cam_off[c] = nv ? off : -1;
    off += nv;
  }
  F->nc = off;
  F->nvpt = 0;
  for (int64_t i = 0; i < p->num_points; ++i) {
    const int cst = !o->refine_points3D || (p->point_constant && p->point_constant[i]) || !pt_used[i];
    F->pt_var[i] = cst ? -1 : (int)F->nvpt++;
  }
  /* effective observations: connected to >= 1 variable block (BundleAdjustmentSummary::num_residuals) */
  F->obs = (int64_t*)malloc(sizeof(int64_t) * (p->num_observations ? p->num_observations : 1));
  F->nobs = 0;
  for (int64_t i = 0; i < p->num_observations; ++i)
    if (F->pose_off[p->obs_pose_idx[i]] >= 0 || F->cam_off[p->obs_camera_idx[i]] >= 0 || F->pt_var[p->obs_point_idx[i]] >= 0 ||
        (p->camera_sensor_idx && p->camera_sensor_idx[p->obs_camera_idx[i]] >= 0 && F->sens_off[p->camera_sensor_idx[p->obs_camera_idx[i]]] >= 0))
      F->obs[F->nobs++] = i;
  qsort_r(F->obs, F->nobs, sizeof(int64_t), cmp_obs, F);
  F->pt_start = (int64_t*)calloc(F->nvpt + 1, sizeof(int64_t));
  F->nobs_var = 0;
  for (int64_t j = 0; j < F->nobs; ++j) {
    const int pv = F->pt_var[p->obs_point_idx[F->obs[j]]];
    if (pv >= 0) { F->pt_start[pv + 1]++; F->nobs_var++; }
  }
  for (int64_t k = 0; k < F->nvpt; ++k) F->pt_start[k + 1] += F->pt_start[k];
  F->r = (double*)calloc(2 * (F->nobs + 1), sizeof(double));
  F->cost_obs = (double*)calloc(F->nobs + 1, sizeof(double));
  F->Jc = (double*)calloc(2 * JC * (F->nobs + 1), sizeof(double));
  F->Jp = (double*)calloc(6 * (F->nobs + 1), sizeof(double));
  F->scale_c = (double*)malloc(sizeof(double) * (F->nc + 1));
  F->scale_p = (double*)malloc(sizeof(double) * (3 * F->nvpt + 1));
  free(pose_used); free(cam_used); free(pt_used);
  return 0;
}

/* residuals (+ Jacobians in tangent space, loss-corrected) at parameters (poses, cams, points); returns cost */
static double linearize(ba_flat* F, const double* poses, const double* sens, const double* cams, const double* pts, int want_jac) {
  const b200ba_problem* p = F->p;
  const b200ba_options* o = F->o;
  double cost = 0.0, delta = 0.0;
#pragma omp parallel for reduction(+ : cost, delta) schedule(static)
  for (int64_t j = 0; j < F->nobs; ++j) {
    const int64_t i = F->obs[j];
    const int pi = p->obs_pose_idx[i], ci = p->obs_camera_idx[i], ti = p->obs_point_idx[i];
    const int id = p->camera_model_id[ci];
    const int si = p->camera_sensor_idx ? p->camera_sensor_idx[ci] : -1;
    double res[2], Jpt[6], Jps[14], Jss[14], Jpr[2 * MAXDK];
    ba_oracle_reproj_rig(id, pts + 3 * ti, poses + 7 * pi, si >= 0 ? sens + 7 * si : NULL, cams + p->camera_param_offset[ci],
                         p->obs_xy + 2 * i, res, Jpt, Jps, Jss, Jpr);
    const double s = res[0] * res[0] + res[1] * res[1];
    double rho[3];
    loss_eval(o->loss_function_type, o->loss_function_scale, s, rho);
    cost += 0.5 * rho[0];
    if (want_jac) F->cost_obs[j] = 0.5 * rho[0]; else delta += 0.5 * rho[0] - F->cost_obs[j];
    double* Jc = F->Jc + 2 * JC * j;
    double* Jp = F->Jp + 6 * j;
    if (want_jac) {
      memset(Jc, 0, sizeof(double) * 2 * JC);
      for (int blk = 0; blk < 2; ++blk) {   /* rig_from_world pose, then sensor_from_rig: quaternion (x) R^3 tangent */
        const int boff = blk == 0 ? F->pose_off[pi] : (si >= 0 ? F->sens_off[si] : -1);
        if (boff < 0) continue;
        const double* q = blk == 0 ? poses + 7 * pi : sens + 7 * si;
        const double* Ja = blk == 0 ? Jps : Jss;
        const uint8_t m = blk == 0 ? F->pose_mask[pi] : 0x3f;
        const int jo = blk == 0 ? 0 : JS;
        /* J_quat (2x4) * PlusJacobian (4x3) of EigenQuaternionManifold at delta = 0 */
        const double PJ[12] = {q[3], q[2], -q[1], -q[2], q[3], q[0], q[1], -q[0], q[3], -q[0], -q[1], -q[2]};
        for (int r = 0; r < 2; ++r) {
          for (int c = 0; c < 3; ++c) {
            double v = 0;
            for (int k = 0; k < 4; ++k) v += Ja[7 * r + k] * PJ[3 * k + c];
            Jc[JC * r + jo + c] = ((m >> c) & 1) ? v : 0.0;
          }
          for (int c = 0; c < 3; ++c) Jc[JC * r + jo + 3 + c] = ((m >> (3 + c)) & 1) ? Ja[7 * r + 4 + c] : 0.0;
        }
      }
      if (F->cam_off[ci] >= 0) {
        const int P = model_num_params(id);
        for (int r = 0; r < 2; ++r)
          for (int k = 0; k < F->cam_nvar[ci]; ++k) Jc[JC * r + JI + k] = Jpr[P * r + F->cam_var[ci][k]];
      }
      if (F->pt_var[ti] >= 0) memcpy(Jp, Jpt, 48); else memset(Jp, 0, 48);
    }
    /* Corrector (Triggs): ceres/corrector.cc */
    double rs = 1.0;
    if (o->loss_function_type != B200BA_LOSS_TRIVIAL) {
      const double sqrt_rho1 = sqrt(rho[1]);
      double alpha_sq_norm = 0.0;
      rs = sqrt_rho1;
      if (s != 0.0 && rho[2] > 0.0) {
        const double D = 1.0 + 2.0 * s * rho[2] / rho[1];
        const double alpha = 1.0 - sqrt(D);
        rs = sqrt_rho1 / (1 - alpha);
        alpha_sq_norm = alpha / s;
      }
      if (want_jac) {
        for (int c = 0; c < JC; ++c) {
          const double rj = res[0] * Jc[c] + res[1] * Jc[JC + c];
          Jc[c] = sqrt_rho1 * (Jc[c] - alpha_sq_norm * res[0] * rj);
          Jc[JC + c] = sqrt_rho1 * (Jc[JC + c] - alpha_sq_norm * res[1] * rj);
        }
        for (int c = 0; c < 3; ++c) {
          const double rj = res[0] * Jp[c] + res[1] * Jp[3 + c];
          Jp[c] = sqrt_rho1 * (Jp[c] - alpha_sq_norm * res[0] * rj);
          Jp[3 + c] = sqrt_rho1 * (Jp[3 + c] - alpha_sq_norm * res[1] * rj);
        }
      }
    }
    if (want_jac) { F->r[2 * j] = rs * res[0]; F->r[2 * j + 1] = rs * res[1]; }
  }
  F->cost_delta = delta;
  return cost;
}

static int chol3_inv(const double A[9], double inv[9]) { /* symmetric 3x3 inverse */
  const double a = A[0], b = A[1], c = A[2], d = A[4], e = A[5], f = A[8];
  const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
  const double det = a * c00 + b * c01 + c * c02;
  if (!(det > 0)) return 0;
  const double id = 1.0 / det;
  inv[0] = c00 * id; inv[1] = c01 * id; inv[2] = c02 * id;
  inv[3] = inv[1]; inv[4] = (a * f - c * c) * id; inv[5] = (b * c - a * e) * id;
  inv[6] = inv[2]; inv[7] = inv[5]; inv[8] = (a * d - b * b) * id;
  return 1;
}
static int chol_solve(int n, double* A, double* b) { /* in place LL^T, dense */
  for (int j = 0; j < n; ++j) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0)) return 0;
    d = sqrt(d); A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
      A[(size_t)i * n + j] = s / d;
    }
  }
  for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= A[(size_t)i * n + k] * b[k]; b[i] = s / A[(size_t)i * n + i]; }
  for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < n; ++k) s -= A[(size_t)k * n + i] * b[k]; b[i] = s / A[(size_t)i * n + i]; }
  return 1;
}

/* camera-side blocks of observation j: rig_from_world pose (6), sensor_from_rig (6), intrinsics (nv); off < 0 = constant.
 * jo = first column of the block inside a Jacobian row */
typedef struct { int off[3], n[3], jo[3]; } obs_blk;
static inline void obs_blocks(const ba_flat* F, int64_t j, obs_blk* B) {
  const int64_t i = F->obs[j];
  const int ci = F->p->obs_camera_idx[i];
  const int si = F->p->camera_sensor_idx ? F->p->camera_sensor_idx[ci] : -1;
  B->off[0] = F->pose_off[F->p->obs_pose_idx[i]]; B->n[0] = 6; B->jo[0] = 0;
  B->off[1] = si >= 0 ? F->sens_off[si] : -1; B->n[1] = 6; B->jo[1] = JS;
  B->off[2] = F->cam_off[ci]; B->n[2] = F->cam_nvar[ci]; B->jo[2] = JI;
}
/* y (2) = Jc_scaled(j) * x */
static inline void jc_times(const ba_flat* F, int64_t j, const double* x, double y[2]) {
  obs_blk B; obs_blocks(F, j, &B);
  const double* Jc = F->Jc + 2 * JC * j;
  y[0] = y[1] = 0;
  for (int w = 0; w < 3; ++w)
    if (B.off[w] >= 0) for (int c = 0; c < B.n[w]; ++c) { const double v = F->scale_c[B.off[w] + c] * x[B.off[w] + c]; y[0] += Jc[B.jo[w] + c] * v; y[1] += Jc[JC + B.jo[w] + c] * v; }
}
/* out += Jc_scaled(j)^T u */
static inline void jct_times_add(const ba_flat* F, int64_t j, const double u[2], double* out) {
  obs_blk B; obs_blocks(F, j, &B);
  const double* Jc = F->Jc + 2 * JC * j;
  for (int w = 0; w < 3; ++w)
    if (B.off[w] >= 0) for (int c = 0; c < B.n[w]; ++c) out[B.off[w] + c] += F->scale_c[B.off[w] + c] * (Jc[B.jo[w] + c] * u[0] + Jc[JC + B.jo[w] + c] * u[1]);
}

typedef struct {
  double* Hpp_inv; /* [9*nvpt] (H_pp + D_p^2)^-1 in scaled space */
  double* Dc2;     /* [nc] */
  double* Dp2;     /* [3*nvpt] */
} ba_lin;

/* S x = (H_cc + D_c^2) x - H_cp (H_pp + D_p^2)^-1 H_pc x, matrix-free over the observations */
static void schur_apply(const ba_flat* F, const ba_lin* L, const double* x, double* out) {
  for (int i = 0; i < F->nc; ++i) out[i] = L->Dc2[i] * x[i];
#pragma omp parallel
  {
    double* loc = (double*)calloc(F->nc + 1, sizeof(double));
#pragma omp for schedule(static) nowait
    for (int64_t k = 0; k < F->nvpt; ++k) {
      double z[3] = {0, 0, 0};
      for (int64_t j = F->pt_start[k]; j < F->pt_start[k + 1]; ++j) {
        double y[2]; jc_times(F, j, x, y);
        const double* Jp = F->Jp + 6 * j;
        for (int c = 0; c < 3; ++c) z[c] += F->scale_p[3 * k + c] * (Jp[c] * y[0] + Jp[3 + c] * y[1]);
      }
      const double* Hi = L->Hpp_inv + 9 * k;
      double w[3];
      for (int c = 0; c < 3; ++c) w[c] = Hi[3 * c] * z[0] + Hi[3 * c + 1] * z[1] + Hi[3 * c + 2] * z[2];
      for (int64_t j = F->pt_start[k]; j < F->pt_start[k + 1]; ++j) {
        double y[2]; jc_times(F, j, x, y);
        const double* Jp = F->Jp + 6 * j;
        double u[2];
        for (int r = 0; r < 2; ++r) {
          double t = 0;
          for (int c = 0; c < 3; ++c) t += Jp[3 * r + c] * F->scale_p[3 * k + c] * w[c];
          u[r] = y[r] - t;
        }
        jct_times_add(F, j, u, loc);
      }
    }
#pragma omp for schedule(static) nowait
    for (int64_t j = F->nobs_var; j < F->nobs; ++j) { double y[2]; jc_times(F, j, x, y); jct_times_add(F, j, y, loc); }
#pragma omp critical
    for (int i = 0; i < F->nc; ++i) out[i] += loc[i];
    free(loc);
  }
}

/* exact diagonal blocks of S per camera-side parameter block (SCHUR_JACOBI preconditioner), inverted */
static void schur_jacobi(const ba_flat* F, const ba_lin* L, const int* blk_start, int nblk, double* Minv /* packed */,
                         const int* blk_pack) {
  double* M = (double*)calloc((size_t)blk_pack[nblk] + 1, sizeof(double));
  /* exact: M_b = Hcc_bb + Dc2_b - sum_p V_b(p)^T Hpp_inv(p) V_b(p),  V_b(p) = sum_{o in p, b in o} Jp_o^T Jc_o,b */
  const size_t msize = (size_t)blk_pack[nblk] + 1;
#pragma omp parallel
  {
  double* Mloc = (double*)calloc(msize, sizeof(double));   /* per-thread partial blocks, joined below */
#pragma omp for schedule(static) nowait
  for (int64_t j = 0; j < F->nobs; ++j) {
    obs_blk B; obs_blocks(F, j, &B);
    const double* Jc = F->Jc + 2 * JC * j;
    for (int which = 0; which < 3; ++which) {
      const int off = B.off[which], n = B.n[which], jo = B.jo[which];
      if (off < 0) continue;
      int b = 0; /* block index by binary search */
      { int lo = 0, hi = nblk - 1; while (lo < hi) { const int mid = (lo + hi + 1) / 2; if (blk_start[mid] <= off) lo = mid; else hi = mid - 1; } b = lo; }
      double* Mb = Mloc + blk_pack[b];
      for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c)
          Mb[r * n + c] += F->scale_c[off + r] * F->scale_c[off + c] * (Jc[jo + r] * Jc[jo + c] + Jc[JC + jo + r] * Jc[JC + jo + c]);
    }
  }
  /* point terms */
#pragma omp for schedule(static) nowait
  for (int64_t k = 0; k < F->nvpt; ++k) {
    const int64_t s = F->pt_start[k], e2 = F->pt_start[k + 1];
    const double* Hi = L->Hpp_inv + 9 * k;
    /* distinct blocks among this point's observations */
    for (int64_t j = s; j < e2; ++j) {
      for (int which = 0; which < 3; ++which) {
        obs_blk B; obs_blocks(F, j, &B);
        const int off = B.off[which], n = B.n[which];
        if (off < 0) continue;
        /* only handle the block at its first occurrence within the point */
        int first = 1;
        for (int64_t j2 = s; j2 < j && first; ++j2) { obs_blk B2; obs_blocks(F, j2, &B2); if (B2.off[which] == off) first = 0; }
        if (!first) continue;
        double V[3 * MAXDK];
        memset(V, 0, sizeof(V));
        for (int64_t j2 = j; j2 < e2; ++j2) {
          obs_blk B2; obs_blocks(F, j2, &B2);
          if (B2.off[which] != off) continue;
          const double* Jc = F->Jc + 2 * JC * j2; const double* Jp = F->Jp + 6 * j2; const int jo = B2.jo[which];
          for (int a = 0; a < 3; ++a)
            for (int c = 0; c < n; ++c)
              V[a * n + c] += F->scale_p[3 * k + a] * F->scale_c[off + c] * (Jp[a] * Jc[jo + c] + Jp[3 + a] * Jc[JC + jo + c]);
        }
        int b; { int lo = 0, hi = nblk - 1; while (lo < hi) { const int mid = (lo + hi + 1) / 2; if (blk_start[mid] <= off) lo = mid; else hi = mid - 1; } b = lo; }
        double* Mb = Mloc + blk_pack[b];
        for (int r = 0; r < n; ++r)
          for (int c = 0; c < n; ++c) {
            double t = 0;
            for (int a = 0; a < 3; ++a) for (int a2 = 0; a2 < 3; ++a2) t += V[a * n + r] * Hi[3 * a + a2] * V[a2 * n + c];
            Mb[r * n + c] -= t;
          }
      }
    }
  }
#pragma omp critical
  for (size_t i = 0; i < msize; ++i) M[i] += Mloc[i];
  free(Mloc);
  }
  for (int b = 0; b < nblk; ++b) { const int n = blk_start[b + 1] - blk_start[b]; for (int r = 0; r < n; ++r) M[blk_pack[b] + r * n + r] += L->Dc2[blk_start[b] + r]; }
  /* invert each block (Cholesky based) */
#pragma omp parallel for schedule(static)
  for (int b = 0; b < nblk; ++b) {
    const int n = blk_start[b + 1] - blk_start[b];
    double A[MAXDK * MAXDK], col[MAXDK];
    for (int c = 0; c < n; ++c) {
      memcpy(A, M + blk_pack[b], sizeof(double) * n * n);
      for (int r = 0; r < n; ++r) col[r] = r == c ? 1.0 : 0.0;
      if (!chol_solve(n, A, col)) { for (int r = 0; r < n; ++r) col[r] = r == c ? 1.0 : 0.0; }
      for (int r = 0; r < n; ++r) Minv[blk_pack[b] + r * n + c] = col[r];
    }
  }
  free(M);
}

/* gradient max norm through the manifold: || x - Plus(x, -g) ||_inf */
static double gradient_max_norm(const ba_flat* F, const double* poses, const double* sens, const double* gc, const double* gp) {
  double m = 0;
  for (int i = 0; i < F->p->num_sensors; ++i) {
    const int off = F->sens_off[i];
    if (off < 0) continue;
    double d[3] = {-gc[off], -gc[off + 1], -gc[off + 2]}, qn[4];
    quat_plus(sens + 7 * i, d, qn);
    for (int k = 0; k < 4; ++k) m = fmax(m, fabs(qn[k] - sens[7 * i + k]));
    for (int k = 3; k < 6; ++k) m = fmax(m, fabs(gc[off + k]));
  }
  for (int i = 0; i < F->p->num_poses; ++i) {
    const int off = F->pose_off[i];
    if (off < 0) continue;
    double d[3] = {-gc[off], -gc[off + 1], -gc[off + 2]}, qn[4];
    quat_plus(poses + 7 * i, d, qn);
    for (int k = 0; k < 4; ++k) m = fmax(m, fabs(qn[k] - poses[7 * i + k]));
    for (int k = 3; k < 6; ++k) m = fmax(m, fabs(gc[off + k]));
  }
  for (int c = 0; c < F->p->num_cameras; ++c) if (F->cam_off[c] >= 0) for (int k = 0; k < F->cam_nvar[c]; ++k) m = fmax(m, fabs(gc[F->cam_off[c] + k]));
  for (int64_t k = 0; k < 3 * F->nvpt; ++k) m = fmax(m, fabs(gp[k]));
  return m;
}

static void apply_step(const ba_flat* F, const double* poses, const double* sens, const double* cams, const double* pts, const double* dc, const double* dp,
                       double* nposes, double* nsens, double* ncams, double* npts, int64_t ncamparams) {
  const b200ba_problem* p = F->p;
  memcpy(nsens, sens, sizeof(double) * 7 * p->num_sensors);
  for (int i = 0; i < p->num_sensors; ++i) {
    const int off = F->sens_off[i];
    if (off < 0) continue;
    quat_plus(sens + 7 * i, dc + off, nsens + 7 * i);
    for (int k = 0; k < 3; ++k) nsens[7 * i + 4 + k] = sens[7 * i + 4 + k] + dc[off + 3 + k];
  }
  memcpy(nposes, poses, sizeof(double) * 7 * p->num_poses);
  memcpy(ncams, cams, sizeof(double) * ncamparams);
  memcpy(npts, pts, sizeof(double) * 3 * p->num_points);
  for (int i = 0; i < p->num_poses; ++i) {
    const int off = F->pose_off[i];
    if (off < 0) continue;
    double d[6];
    for (int k = 0; k < 6; ++k) d[k] = ((F->pose_mask[i] >> k) & 1) ? dc[off + k] : 0.0;
    quat_plus(poses + 7 * i, d, nposes + 7 * i);
    for (int k = 0; k < 3; ++k) nposes[7 * i + 4 + k] = poses[7 * i + 4 + k] + d[3 + k];
  }
  for (int c = 0; c < p->num_cameras; ++c)
    if (F->cam_off[c] >= 0)
      for (int k = 0; k < F->cam_nvar[c]; ++k) ncams[p->camera_param_offset[c] + F->cam_var[c][k]] += dc[F->cam_off[c] + k];
  for (int64_t i = 0; i < p->num_points; ++i)
    if (F->pt_var[i] >= 0) for (int k = 0; k < 3; ++k) npts[3 * i + k] += dp[3 * (int64_t)F->pt_var[i] + k];
}

int ba_oracle_solve(const b200ba_options* o, b200ba_problem* p, b200ba_summary* sum) {
  memset(sum, 0, sizeof(*sum));
  sum->termination_type = B200BA_FAILURE;
  ba_flat F; memset(&F, 0, sizeof(F)); F.o = o; F.p = p;
  if (flatten(&F) != 0) return -2;
  sum->num_residuals = (int)(2 * F.nobs);
  int64_t ncamparams = 0;
  for (int c = 0; c < p->num_cameras; ++c) { const int e = p->camera_param_offset[c] + model_num_params(p->camera_model_id[c]); if (e > ncamparams) ncamparams = e; }
  int neff = 0;
  for (int i = 0; i < p->num_poses; ++i) if (F.pose_off[i] >= 0) neff += __builtin_popcount(F.pose_mask[i]);
  for (int c = 0; c < p->num_cameras; ++c) neff += F.cam_nvar[c];
  for (int k = 0; k < p->num_sensors; ++k) if (F.sens_off[k] >= 0) neff += 6;
  neff += (int)(3 * F.nvpt);
  sum->num_effective_parameters = neff;
  int lst = o->linear_solver_type;
  if (lst == B200BA_AUTO) {   /* config.NumImages() decides (bundle_adjustment_ceres.cc:131,204-210) */
    int nimg = p->num_config_images;
    if (nimg <= 0) {
      unsigned char* used = (unsigned char*)calloc((size_t)p->num_poses + 1, 1);
      for (int64_t i = 0; i < p->num_observations; ++i) used[p->obs_pose_idx[i]] = 1;
      for (int i = 0; i < p->num_poses; ++i) nimg += used[i];
      free(used);
    }
    lst = nimg <= 50 ? B200BA_DENSE_SCHUR : (nimg <= 1000 ? B200BA_SPARSE_SCHUR : B200BA_ITERATIVE_SCHUR);
  }
  sum->linear_solver_type_used = lst;
  if (F.nobs == 0 || neff == 0) { sum->termination_type = B200BA_CONVERGENCE; return 0; }

  const int nc = F.nc; const int64_t np3 = 3 * F.nvpt;
  double* poses = (double*)malloc(sizeof(double) * 7 * p->num_poses); memcpy(poses, p->poses, sizeof(double) * 7 * p->num_poses);
  double* cams = (double*)malloc(sizeof(double) * (ncamparams + 1)); memcpy(cams, p->camera_params, sizeof(double) * ncamparams);
  double* pts = (double*)malloc(sizeof(double) * 3 * p->num_points); memcpy(pts, p->points, sizeof(double) * 3 * p->num_points);
  /* ParameterizeRigsAndFrames normalises quaternions (bundle_adjustment_ceres.cc:514) */
  for (int i = 0; i < p->num_poses; ++i) if (F.pose_off[i] >= 0) { double* q = poses + 7 * i; const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]); for (int k = 0; k < 4; ++k) q[k] /= n; }
  double* sens = (double*)malloc(sizeof(double) * (7 * p->num_sensors + 1)); double* nsens = (double*)malloc(sizeof(double) * (7 * p->num_sensors + 1));
  if (p->num_sensors) memcpy(sens, p->sensor_from_rig, sizeof(double) * 7 * p->num_sensors);
  for (int i = 0; i < p->num_sensors; ++i) if (F.sens_off[i] >= 0) { double* q = sens + 7 * i; const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]); for (int k = 0; k < 4; ++k) q[k] /= n; }
  double* nposes = (double*)malloc(sizeof(double) * 7 * p->num_poses);
  double* ncams = (double*)malloc(sizeof(double) * (ncamparams + 1));
  double* npts = (double*)malloc(sizeof(double) * 3 * p->num_points);
  double* gc = (double*)malloc(sizeof(double) * (nc + 1)); double* gp = (double*)malloc(sizeof(double) * (np3 + 1));
  double* dc = (double*)malloc(sizeof(double) * (nc + 1)); double* dp = (double*)malloc(sizeof(double) * (np3 + 1));
  double* rhs = (double*)malloc(sizeof(double) * (nc + 1));
  ba_lin L; L.Hpp_inv = (double*)malloc(sizeof(double) * (9 * F.nvpt + 1)); L.Dc2 = (double*)malloc(sizeof(double) * (nc + 1)); L.Dp2 = (double*)malloc(sizeof(double) * (np3 + 1));
  double* diag_c = (double*)malloc(sizeof(double) * (nc + 1)); double* diag_p = (double*)malloc(sizeof(double) * (np3 + 1));
  /* camera-side parameter blocks */
  int nblk = 0; int* blk_start = (int*)malloc(sizeof(int) * (p->num_poses + p->num_sensors + p->num_cameras + 2)); int* blk_pack = (int*)malloc(sizeof(int) * (p->num_poses + p->num_sensors + p->num_cameras + 2));
  { int pk = 0;
    for (int i = 0; i < p->num_poses; ++i) if (F.pose_off[i] >= 0) { blk_start[nblk] = F.pose_off[i]; blk_pack[nblk] = pk; pk += 36; nblk++; }
    for (int i = 0; i < p->num_sensors; ++i) if (F.sens_off[i] >= 0) { blk_start[nblk] = F.sens_off[i]; blk_pack[nblk] = pk; pk += 36; nblk++; }
    for (int c = 0; c < p->num_cameras; ++c) if (F.cam_off[c] >= 0) { blk_start[nblk] = F.cam_off[c]; blk_pack[nblk] = pk; pk += F.cam_nvar[c] * F.cam_nvar[c]; nblk++; }
    blk_start[nblk] = nc; blk_pack[nblk] = pk; }

  double cost = linearize(&F, poses, sens, cams, pts, 1);
  sum->initial_cost = cost;
  double radius = o->initial_trust_region_radius, decrease_factor = 2.0;
  int iter = 0, have_scale = 0;
  sum->termination_type = B200BA_NO_CONVERGENCE;
  for (;;) {
    /* jacobi scaling from the first Jacobian */
    if (!have_scale) {
      for (int i = 0; i < nc; ++i) F.scale_c[i] = 0; for (int64_t i = 0; i < np3; ++i) F.scale_p[i] = 0;
      /* camera side: per-thread partial sums joined afterwards; point side: one thread per point (observations are
       * grouped by point) */
#pragma omp parallel
      {
        double* loc = (double*)calloc(nc + 1, sizeof(double));
#pragma omp for schedule(static) nowait
        for (int64_t j = 0; j < F.nobs; ++j) {
          obs_blk B; obs_blocks(&F, j, &B); const double* Jc = F.Jc + 2 * JC * j;
          for (int w = 0; w < 3; ++w)
            if (B.off[w] >= 0) for (int c = 0; c < B.n[w]; ++c) loc[B.off[w] + c] += Jc[B.jo[w] + c] * Jc[B.jo[w] + c] + Jc[JC + B.jo[w] + c] * Jc[JC + B.jo[w] + c];
        }
#pragma omp critical
        for (int i = 0; i < nc; ++i) F.scale_c[i] += loc[i];
        free(loc);
      }
#pragma omp parallel for schedule(static)
      for (int64_t k = 0; k < F.nvpt; ++k)
        for (int64_t j = F.pt_start[k]; j < F.pt_start[k + 1]; ++j) {
          const double* Jp = F.Jp + 6 * j;
          for (int c = 0; c < 3; ++c) F.scale_p[3 * k + c] += Jp[c] * Jp[c] + Jp[3 + c] * Jp[3 + c];
        }
      for (int i = 0; i < nc; ++i) F.scale_c[i] = o->jacobi_scaling ? 1.0 / (1.0 + sqrt(F.scale_c[i])) : 1.0;
      for (int64_t i = 0; i < np3; ++i) F.scale_p[i] = o->jacobi_scaling ? 1.0 / (1.0 + sqrt(F.scale_p[i])) : 1.0;
      have_scale = 1;
    }
    /* gradient (scaled space) and diag(J_s^T J_s) */
    memset(gc, 0, sizeof(double) * nc); memset(gp, 0, sizeof(double) * np3);
    memset(diag_c, 0, sizeof(double) * nc); memset(diag_p, 0, sizeof(double) * np3);
#pragma omp parallel
    {
      double* lg = (double*)calloc(2 * (size_t)nc + 2, sizeof(double)); double* ld = lg + nc + 1;
#pragma omp for schedule(static) nowait
      for (int64_t j = 0; j < F.nobs; ++j) {
        jct_times_add(&F, j, F.r + 2 * j, lg);
        obs_blk B; obs_blocks(&F, j, &B); const double* Jc = F.Jc + 2 * JC * j;
        for (int w = 0; w < 3; ++w)
          if (B.off[w] >= 0) for (int c = 0; c < B.n[w]; ++c) { const double sc = F.scale_c[B.off[w] + c]; ld[B.off[w] + c] += sc * sc * (Jc[B.jo[w] + c] * Jc[B.jo[w] + c] + Jc[JC + B.jo[w] + c] * Jc[JC + B.jo[w] + c]); }
      }
#pragma omp critical
      for (int i = 0; i < nc; ++i) { gc[i] += lg[i]; diag_c[i] += ld[i]; }
      free(lg);
    }
#pragma omp parallel for schedule(static)
    for (int64_t pv = 0; pv < F.nvpt; ++pv)
      for (int64_t j = F.pt_start[pv]; j < F.pt_start[pv + 1]; ++j) {
        const double* Jp = F.Jp + 6 * j;
        for (int c = 0; c < 3; ++c) {
          const double s = F.scale_p[3 * pv + c];
          gp[3 * pv + c] += s * (Jp[c] * F.r[2 * j] + Jp[3 + c] * F.r[2 * j + 1]);
          diag_p[3 * pv + c] += s * s * (Jp[c] * Jp[c] + Jp[3 + c] * Jp[3 + c]);
        }
      }
    if (iter == 0 || 1) {
      /* gradient tolerance is tested on the unscaled gradient */
      double* ugc = dc; double* ugp = dp;
      for (int i = 0; i < nc; ++i) ugc[i] = gc[i] / F.scale_c[i];
      for (int64_t i = 0; i < np3; ++i) ugp[i] = gp[i] / F.scale_p[i];
      if (gradient_max_norm(&F, poses, sens, ugc, ugp) <= o->gradient_tolerance) { sum->termination_type = B200BA_CONVERGENCE; break; }
    }
    /* inner loop: retry with smaller radius until a step is accepted (Jacobian unchanged) */
    int accepted = 0;
    while (!accepted) {
      if (iter >= o->max_num_iterations) goto done;
      ++iter;
      for (int i = 0; i < nc; ++i) L.Dc2[i] = fmin(fmax(diag_c[i], o->min_lm_diagonal), o->max_lm_diagonal) / radius;
      for (int64_t i = 0; i < np3; ++i) L.Dp2[i] = fmin(fmax(diag_p[i], o->min_lm_diagonal), o->max_lm_diagonal) / radius;
      /* point blocks */
      int ok = 1;
#pragma omp parallel for schedule(static) reduction(&& : ok)
      for (int64_t k = 0; k < F.nvpt; ++k) {
        double H[9] = {0};
        for (int64_t j = F.pt_start[k]; j < F.pt_start[k + 1]; ++j) {
          const double* Jp = F.Jp + 6 * j;
          for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) H[3 * a + b] += F.scale_p[3 * k + a] * F.scale_p[3 * k + b] * (Jp[a] * Jp[b] + Jp[3 + a] * Jp[3 + b]);
        }
        for (int a = 0; a < 3; ++a) H[4 * a] += L.Dp2[3 * k + a];
        if (!chol3_inv(H, L.Hpp_inv + 9 * k)) ok = 0;
      }
      /* reduced right-hand side: -g_c + H_cp Hpp^-1 g_p */
      for (int i = 0; i < nc; ++i) rhs[i] = -gc[i];
#pragma omp parallel
      {
        double* loc = (double*)calloc(nc + 1, sizeof(double));
#pragma omp for schedule(static) nowait
        for (int64_t k = 0; k < F.nvpt; ++k) {
          const double* Hi = L.Hpp_inv + 9 * k; double w[3];
          for (int c = 0; c < 3; ++c) w[c] = Hi[3 * c] * gp[3 * k] + Hi[3 * c + 1] * gp[3 * k + 1] + Hi[3 * c + 2] * gp[3 * k + 2];
          for (int64_t j = F.pt_start[k]; j < F.pt_start[k + 1]; ++j) {
            const double* Jp = F.Jp + 6 * j; double u[2];
            for (int r = 0; r < 2; ++r) { double t = 0; for (int c = 0; c < 3; ++c) t += Jp[3 * r + c] * F.scale_p[3 * k + c] * w[c]; u[r] = t; }
            jct_times_add(&F, j, u, loc);
          }
        }
#pragma omp critical
        for (int i = 0; i < nc; ++i) rhs[i] += loc[i];
        free(loc);
      }
      /* reduced solve */
      memset(dc, 0, sizeof(double) * nc);
      if (ok && nc > 0) {
        if (lst != B200BA_ITERATIVE_SCHUR) {
          double* S = (double*)malloc(sizeof(double) * (size_t)nc * nc); double* e = (double*)calloc(nc, sizeof(double)); double* col = (double*)malloc(sizeof(double) * nc);
          for (int i = 0; i < nc; ++i) { e[i] = 1.0; schur_apply(&F, &L, e, col); e[i] = 0.0; for (int r = 0; r < nc; ++r) S[(size_t)r * nc + i] = col[r]; }
          memcpy(dc, rhs, sizeof(double) * nc);
          ok = chol_solve(nc, S, dc);
          free(S); free(e); free(col);
        } else {
          /* preconditioned CG, SCHUR_JACOBI, Q-tolerance termination (ceres/conjugate_gradients_solver) */
          double* Minv = (double*)malloc(sizeof(double) * (blk_pack[nblk] + 1));
          schur_jacobi(&F, &L, blk_start, nblk, Minv, blk_pack);
          double *r = (double*)malloc(sizeof(double) * nc), *z = (double*)malloc(sizeof(double) * nc), *pv = (double*)malloc(sizeof(double) * nc), *q = (double*)malloc(sizeof(double) * nc);
          memcpy(r, rhs, sizeof(double) * nc);
          double norm_b = 0; for (int i = 0; i < nc; ++i) norm_b += rhs[i] * rhs[i]; norm_b = sqrt(norm_b);
          double rho = 1.0, Q0 = 0.0;
          if (norm_b > 0) for (int it = 1; it <= o->max_linear_solver_iterations; ++it) {
            for (int b = 0; b < nblk; ++b) { const int n = blk_start[b + 1] - blk_start[b]; const double* Mb = Minv + blk_pack[b];
              for (int rr = 0; rr < n; ++rr) { double t = 0; for (int c = 0; c < n; ++c) t += Mb[rr * n + c] * r[blk_start[b] + c]; z[blk_start[b] + rr] = t; } }
            const double last_rho = rho; rho = 0; for (int i = 0; i < nc; ++i) rho += r[i] * z[i];
            if (it == 1) memcpy(pv, z, sizeof(double) * nc); else { const double beta = rho / last_rho; for (int i = 0; i < nc; ++i) pv[i] = z[i] + beta * pv[i]; }
            schur_apply(&F, &L, pv, q);
            double pq = 0; for (int i = 0; i < nc; ++i) pq += pv[i] * q[i];
            if (!(pq > 0)) break;
            const double alpha = rho / pq;
            for (int i = 0; i < nc; ++i) { dc[i] += alpha * pv[i]; r[i] -= alpha * q[i]; }
            sum->num_linear_solver_iterations++;
            double Q1 = 0; for (int i = 0; i < nc; ++i) Q1 += dc[i] * (rhs[i] + r[i]); Q1 = -1.0 * Q1;
            const double zeta = it * (Q1 - Q0) / Q1;
            if (zeta < o->eta) break;
            Q0 = Q1;
            double nr = 0; for (int i = 0; i < nc; ++i) nr += r[i] * r[i];
            if (sqrt(nr) <= 1e-300) break;
          }
          free(Minv); free(r); free(z); free(pv); free(q);
        }
      }
      /* back-substitution: dp = Hpp^-1 (-g_p - H_pc dc) */
#pragma omp parallel for schedule(static)
      for (int64_t k = 0; k < F.nvpt; ++k) {
        double t[3] = {-gp[3 * k], -gp[3 * k + 1], -gp[3 * k + 2]};
        for (int64_t j = F.pt_start[k]; j < F.pt_start[k + 1]; ++j) {
          double y[2]; jc_times(&F, j, dc, y); const double* Jp = F.Jp + 6 * j;
          for (int c = 0; c < 3; ++c) t[c] -= F.scale_p[3 * k + c] * (Jp[c] * y[0] + Jp[3 + c] * y[1]);
        }
        const double* Hi = L.Hpp_inv + 9 * k;
        for (int c = 0; c < 3; ++c) dp[3 * k + c] = Hi[3 * c] * t[0] + Hi[3 * c + 1] * t[1] + Hi[3 * c + 2] * t[2];
      }
      /* model cost change = -(J d)^T (r + J d / 2) */
      double model = 0;
#pragma omp parallel for schedule(static) reduction(+ : model)
      for (int64_t j = 0; j < F.nobs; ++j) {
        double y[2]; jc_times(&F, j, dc, y);
        if (j < F.nobs_var) { const int64_t pv = F.pt_var[p->obs_point_idx[F.obs[j]]]; const double* Jp = F.Jp + 6 * j;
          for (int r = 0; r < 2; ++r) for (int c = 0; c < 3; ++c) y[r] += Jp[3 * r + c] * F.scale_p[3 * pv + c] * dp[3 * pv + c]; }
        model -= y[0] * (F.r[2 * j] + 0.5 * y[0]) + y[1] * (F.r[2 * j + 1] + 0.5 * y[1]);
      }
      double rho_q = 0.0, new_cost = cost;
      if (ok && model > 0) {
        /* undo the scaling, apply */
        double* udc = (double*)malloc(sizeof(double) * (nc + 1)); double* udp = (double*)malloc(sizeof(double) * (np3 + 1));
        for (int i = 0; i < nc; ++i) udc[i] = dc[i] * F.scale_c[i];
        for (int64_t i = 0; i < np3; ++i) udp[i] = dp[i] * F.scale_p[i];
        apply_step(&F, poses, sens, cams, pts, udc, udp, nposes, nsens, ncams, npts, ncamparams);
        free(udc); free(udp);
        new_cost = linearize(&F, nposes, nsens, ncams, npts, 0);
        /* cost change summed residual by residual: same quantity as cost - new_cost without the cancellation */
        rho_q = (-F.cost_delta) / model;
      }
      if (getenv("BA_ORACLE_VERBOSE")) fprintf(stderr, "[oracle] it %d cost %.12g new %.12g model %.6g rho %.6g radius %.4g pcg_total %d\n", iter, cost, new_cost, model, rho_q, radius, sum->num_linear_solver_iterations);
      if (ok && model > 0 && rho_q > o->min_relative_decrease) {
        accepted = 1;
        sum->num_successful_steps++;
        memcpy(poses, nposes, sizeof(double) * 7 * p->num_poses); memcpy(cams, ncams, sizeof(double) * ncamparams); memcpy(pts, npts, sizeof(double) * 3 * p->num_points);
        memcpy(sens, nsens, sizeof(double) * 7 * p->num_sensors);
        const double cost_change = -F.cost_delta;
        cost = linearize(&F, poses, sens, cams, pts, 1);
        const double t = 2.0 * rho_q - 1.0;
        radius = radius / fmax(1.0 / 3.0, 1.0 - t * t * t);
        radius = fmin(o->max_trust_region_radius, radius);
        decrease_factor = 2.0;
        if (fabs(cost_change) <= o->function_tolerance * cost) { sum->termination_type = B200BA_CONVERGENCE; goto done; }
      } else {
        sum->num_unsuccessful_steps++;
        radius = radius / decrease_factor;
        decrease_factor *= 2.0;
        if (radius < o->min_trust_region_radius) { sum->termination_type = B200BA_CONVERGENCE; goto done; }
      }
    }
  }
done:
  sum->final_cost = cost;
  /* write back only variable blocks; constants stay bit-identical */
  for (int i = 0; i < p->num_poses; ++i) if (F.pose_off[i] >= 0) memcpy(p->poses + 7 * i, poses + 7 * i, 56);
  for (int i = 0; i < p->num_sensors; ++i) if (F.sens_off[i] >= 0) memcpy(p->sensor_from_rig + 7 * i, sens + 7 * i, 56);
  free(sens); free(nsens); free(F.sens_off);
  for (int c = 0; c < p->num_cameras; ++c) if (F.cam_off[c] >= 0) for (int k = 0; k < F.cam_nvar[c]; ++k) { const int idx = p->camera_param_offset[c] + F.cam_var[c][k]; p->camera_params[idx] = cams[idx]; }
  for (int64_t i = 0; i < p->num_points; ++i) if (F.pt_var[i] >= 0) memcpy(p->points + 3 * i, pts + 3 * i, 24);
  free(poses); free(cams); free(pts); free(nposes); free(ncams); free(npts); free(gc); free(gp); free(dc); free(dp); free(rhs);
  free(L.Hpp_inv); free(L.Dc2); free(L.Dp2); free(diag_c); free(diag_p); free(blk_start); free(blk_pack);
  free(F.pose_off); free(F.pose_mask); free(F.cam_off); free(F.cam_nvar); free(F.cam_var); free(F.pt_var); free(F.obs); free(F.pt_start);
  free(F.cost_obs); free(F.r); free(F.Jc); free(F.Jp); free(F.scale_c); free(F.scale_p);
  return 0;
}
