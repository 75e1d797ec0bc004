// ba_models.cuh — camera models with MORE than five parameters, as formulas only, differentiated by forward-mode dual
// numbers that live in registers (groundwork for SURVEY.md section 8f rank 4; the LM kernels of bundle_adjustment.cu
// still take the <= 5-parameter models through their hand-written Jacobians, so nothing here is on the solve path yet —
// it is reachable through the CPU test hook b200ba_test_project_wide only).
//
// Reference behaviour: ImgFromCam of OPENCV, OPENCV_FISHEYE, FULL_OPENCV, FOV, THIN_PRISM_FISHEYE
// (src/colmap/sensor/models.h:1513-1600, 1625-1683, 1713-1781, 1809-1885, 2180-2238; FisheyeFromNormal :429-438; depth
// guard :281-285).  The reference carries ~1200 lines of hand-derived Jacobians for these models
// (sensor/models_jacobian.h:401-1565); here the derivative falls out of the value computation, the same source works on
// the host and on the device, and a new model is ten lines.
#pragma once
#include <math.h>

#ifndef BA_HD
#define BA_HD __host__ __device__ __forceinline__
#endif

template <int N>
struct BaDual {
  double v;
  double d[N];
  BA_HD BaDual() : v(0.0) { for (int i = 0; i < N; ++i) d[i] = 0.0; }
  BA_HD BaDual(double x) : v(x) { for (int i = 0; i < N; ++i) d[i] = 0.0; }   // constant
  BA_HD static BaDual variable(double x, int k) { BaDual r(x); r.d[k] = 1.0; return r; }
};
template <int N> BA_HD BaDual<N> operator+(const BaDual<N>& a, const BaDual<N>& b) { BaDual<N> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> BA_HD BaDual<N> operator-(const BaDual<N>& a, const BaDual<N>& b) { BaDual<N> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> BA_HD BaDual<N> operator-(const BaDual<N>& a) { BaDual<N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
template <int N> BA_HD BaDual<N> operator*(const BaDual<N>& a, const BaDual<N>& b) { BaDual<N> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int N> BA_HD BaDual<N> operator/(const BaDual<N>& a, const BaDual<N>& b) {
  BaDual<N> r; const double ib = 1.0 / b.v; r.v = a.v * ib;
  for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
  return r;
}
template <int N> BA_HD BaDual<N> ba_sqrt(const BaDual<N>& a) { BaDual<N> r; r.v = sqrt(a.v); const double s = 0.5 / r.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * s; return r; }
template <int N> BA_HD BaDual<N> ba_atan(const BaDual<N>& a) { BaDual<N> r; r.v = atan(a.v); const double s = 1.0 / (1.0 + a.v * a.v); for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * s; return r; }
template <int N> BA_HD BaDual<N> ba_tan(const BaDual<N>& a) { BaDual<N> r; r.v = tan(a.v); const double s = 1.0 + r.v * r.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * s; return r; }
BA_HD double ba_sqrt(double a) { return sqrt(a); }
BA_HD double ba_atan(double a) { return atan(a); }
BA_HD double ba_tan(double a) { return tan(a); }
template <int N> BA_HD double ba_value(const BaDual<N>& a) { return a.v; }
BA_HD double ba_value(double a) { return a; }

BA_HD int ba_wide_model_num_params(int id) { return id == 4 ? 8 : (id == 5 ? 8 : (id == 6 ? 12 : (id == 7 ? 5 : (id == 10 ? 12 : -1)))); }

// equidistant fisheye mapping of the normalised point (identity for r -> 0)
template <typename S>
BA_HD void ba_fisheye_from_normal(const S& a, const S& b, S* fa, S* fb) {
  const S r = ba_sqrt(a * a + b * b);
  if (ba_value(r) > 2.220446049250313e-16) { const S s = ba_atan(r) / r; *fa = a * s; *fb = b * s; }
  else { *fa = a; *fb = b; }
}

// q: parameters (as S), (u, v, w): point in the camera frame; false if the depth guard fails
template <typename S>
BA_HD bool ba_project_wide(int id, const S* q, const S& u, const S& v, const S& w, S* x, S* y) {
  if (!(ba_value(w) >= 2.220446049250313e-16)) return false;
  const S a = u / w, b = v / w;
  const S two(2.0), one(1.0);
  if (id == 4) {          // OPENCV: fx fy cx cy k1 k2 p1 p2
    const S a2 = a * a, ab = a * b, b2 = b * b, r2 = a2 + b2, radial = q[4] * r2 + q[5] * r2 * r2;
    const S du = a * radial + two * q[6] * ab + q[7] * (r2 + two * a2);
    const S dv = b * radial + two * q[7] * ab + q[6] * (r2 + two * b2);
    *x = q[0] * (a + du) + q[2]; *y = q[1] * (b + dv) + q[3];
  } else if (id == 5) {   // OPENCV_FISHEYE: fx fy cx cy k1 k2 k3 k4
    S fa, fb; ba_fisheye_from_normal(a, b, &fa, &fb);
    const S t2 = fa * fa + fb * fb, t4 = t2 * t2, t6 = t4 * t2, t8 = t4 * t4;
    const S radial = q[4] * t2 + q[5] * t4 + q[6] * t6 + q[7] * t8;
    *x = q[0] * (fa + fa * radial) + q[2]; *y = q[1] * (fb + fb * radial) + q[3];
  } else if (id == 6) {   // FULL_OPENCV: fx fy cx cy k1 k2 p1 p2 k3 k4 k5 k6
    const S a2 = a * a, ab = a * b, b2 = b * b, r2 = a2 + b2, r4 = r2 * r2, r6 = r4 * r2;
    const S radial = (one + q[4] * r2 + q[5] * r4 + q[8] * r6) / (one + q[9] * r2 + q[10] * r4 + q[11] * r6);
    const S xd = a * radial + two * q[6] * ab + q[7] * (r2 + two * a2);
    const S yd = b * radial + two * q[7] * ab + q[6] * (r2 + two * b2);
    *x = q[0] * xd + q[2]; *y = q[1] * yd + q[3];
  } else if (id == 7) {   // FOV: fx fy cx cy omega
    const S omega = q[4], r2 = a * a + b * b, o2 = omega * omega;
    S factor;
    if (ba_value(o2) < 1e-4) factor = (o2 * r2) / S(3.0) - o2 / S(12.0) + one;
    else if (ba_value(r2) < 1e-4) { const S t = ba_tan(omega / two); factor = (S(-2.0) * t * (S(4.0) * r2 * t * t - S(3.0))) / (S(3.0) * omega); }
    else { const S r = ba_sqrt(r2); factor = ba_atan(r * two * ba_tan(omega / two)) / (r * omega); }
    *x = q[0] * (a * factor) + q[2]; *y = q[1] * (b * factor) + q[3];
  } else if (id == 10) {  // THIN_PRISM_FISHEYE: fx fy cx cy k1 k2 p1 p2 k3 k4 sx1 sy1
    S fa, fb; ba_fisheye_from_normal(a, b, &fa, &fb);
    const S a2 = fa * fa, ab = fa * fb, b2 = fb * fb, r2 = a2 + b2, r4 = r2 * r2, r6 = r4 * r2, r8 = r6 * r2;
    const S radial = q[4] * r2 + q[5] * r4 + q[8] * r6 + q[9] * r8;
    const S du = fa * radial + two * q[6] * ab + q[7] * (r2 + two * a2) + q[10] * r2;
    const S dv = fb * radial + two * q[7] * ab + q[6] * (r2 + two * b2) + q[11] * r2;
    *x = q[0] * (fa + du) + q[2]; *y = q[1] * (fb + dv) + q[3];
  } else {
    return false;
  }
  return true;
}

// value + Jacobians wrt (u, v, w) [2x3] and the P parameters [2xP], all in one pass of duals with 3 + P derivative slots
template <int P>
BA_HD bool ba_project_wide_with_jac(int id, const double* params, double u, double v, double w, double* xy, double* J_uvw,
                                    double* J_params) {
  typedef BaDual<3 + P> D;
  D q[P];
  for (int k = 0; k < P; ++k) q[k] = D::variable(params[k], 3 + k);
  D x, y;
  if (!ba_project_wide<D>(id, q, D::variable(u, 0), D::variable(v, 1), D::variable(w, 2), &x, &y)) return false;
  xy[0] = x.v; xy[1] = y.v;
  for (int c = 0; c < 3; ++c) { J_uvw[c] = x.d[c]; J_uvw[3 + c] = y.d[c]; }
  for (int k = 0; k < P; ++k) { J_params[k] = x.d[3 + k]; J_params[P + k] = y.d[3 + k]; }
  return true;
}
