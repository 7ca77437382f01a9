// detmath.cuh -- float64 device math built from + - * / sqrt and floor only.
//
// The 6x6 solve, the pose increment and the convergence checkers of the ICP loop run on the GPU inside the last
// block of the normal-equation kernel.  They use this small self-contained libm so that the result is a pure function
// of the (exactly accumulated) integer sums: IEEE-754 add/mul/div/sqrt are correctly rounded on sm_100a, and the
// library is built with -fmad=false, so no step depends on CUDA's transcendental approximations.
//
// Algorithms (libpointmatcher behaviour per SURVEY.md Appendix A.5 / A.7):
//   sincos : Cody-Waite reduction by pi/2 (33-bit head + tail), forward Taylor sums (<= 11 terms, stop when a term no
//            longer changes the sum; factorial ratios applied as reciprocal multiplications)
//   atan2  : three half-angle reductions + 11-term alternating series, first quadrant only
//   jacobi : cyclic Jacobi eigen-decomposition of a symmetric NxN matrix
//   solve6 : LLT with a relative pivot test standing in for fullPivHouseholderQr(A).isInvertible(); minimal-norm
//            eigen fallback when the test fails (PointToPlaneErrorMinimizer::compute_in_place)
#pragma once

#include <float.h>

namespace aicp {

// correctly rounded reciprocals of the series denominators (compile-time constants: the same values a run-time IEEE
// division produces, which is what the oracle executes)
#define AICP_R(x) (1.0 / (double)(x))
__device__ static const double kInvSin[12] = {0.0, AICP_R(2 * 3), AICP_R(4 * 5), AICP_R(6 * 7), AICP_R(8 * 9), AICP_R(10 * 11), AICP_R(12 * 13),
                                     AICP_R(14 * 15), AICP_R(16 * 17), AICP_R(18 * 19), AICP_R(20 * 21), AICP_R(22 * 23)};
__device__ static const double kInvCos[12] = {0.0, AICP_R(1 * 2), AICP_R(3 * 4), AICP_R(5 * 6), AICP_R(7 * 8), AICP_R(9 * 10), AICP_R(11 * 12),
                                     AICP_R(13 * 14), AICP_R(15 * 16), AICP_R(17 * 18), AICP_R(19 * 20), AICP_R(21 * 22)};
__device__ static const double kInvOdd[12] = {0.0, AICP_R(3), AICP_R(5), AICP_R(7), AICP_R(9), AICP_R(11), AICP_R(13), AICP_R(15), AICP_R(17),
                                     AICP_R(19), AICP_R(21), AICP_R(23)};
#undef AICP_R

__device__ inline void det_sincos(double x, double* s, double* c) {
  const double two_over_pi = 0.63661977236758138;
  const double pio2_hi = 1.5707963267341256;
  const double pio2_lo = 6.0771005065061922e-11;
  double kd = floor(x * two_over_pi + 0.5);
  double r = (x - kd * pio2_hi) - kd * pio2_lo;
  double r2 = r * r;
  double term = r, ss = r;
  for (int n = 1; n <= 11; ++n) {
    double inv = kInvSin[n];
    term = (term * r2) * inv;
    term = -term;
    double ns = ss + term;
    if (ns == ss) break;
    ss = ns;
  }
  double cterm = 1.0, cc = 1.0;
  for (int n = 1; n <= 11; ++n) {
    double inv = kInvCos[n];
    cterm = (cterm * r2) * inv;
    cterm = -cterm;
    double nc = cc + cterm;
    if (nc == cc) break;
    cc = nc;
  }
  long long k = (long long)kd;
  int quad = (int)(((k % 4) + 4) % 4);
  if (quad == 0) { *s = ss; *c = cc; }
  else if (quad == 1) { *s = cc; *c = -ss; }
  else if (quad == 2) { *s = -ss; *c = -cc; }
  else { *s = -cc; *c = ss; }
}

__device__ inline double det_atan01(double z) {
  double u = z;
  for (int h = 0; h < 3; ++h) u = u / (1.0 + sqrt(1.0 + u * u));
  double u2 = u * u, p = u, sum = u;
  for (int n = 1; n <= 11; ++n) {
    double inv = kInvOdd[n];
    p = p * u2;
    p = -p;
    double ns = sum + p * inv;
    if (ns == sum) break;
    sum = ns;
  }
  return 8.0 * sum;
}

__device__ inline double det_atan2_pos(double y, double x) {
  const double pio2 = 1.5707963267948966;
  if (y == 0.0 && x == 0.0) return 0.0;
  if (y <= x) return det_atan01(y / x);
  return pio2 - det_atan01(x / y);
}

template <int N>
__device__ inline void det_jacobi(double (&a)[N][N], double (&v)[N][N]) {
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0;
#pragma unroll
    for (int p = 0; p < N - 1; ++p)
#pragma unroll
      for (int q = p + 1; q < N; ++q) off = off + a[p][q] * a[p][q];
    if (off == 0.0) break;
#pragma unroll
    for (int p = 0; p < N - 1; ++p) {
#pragma unroll
      for (int q = p + 1; q < N; ++q) {
        double apq = a[p][q];
        if (apq != 0.0) {
          double app = a[p][p], aqq = a[q][q];
          double theta = (aqq - app) / (2.0 * apq);
          double t;
          if (theta >= 0.0) t = 1.0 / (theta + sqrt(theta * theta + 1.0));
          else t = -1.0 / (-theta + sqrt(theta * theta + 1.0));
          double c = 1.0 / sqrt(t * t + 1.0);
          double s = t * c;
          a[p][p] = app - t * apq;
          a[q][q] = aqq + t * apq;
          a[p][q] = 0.0;
          a[q][p] = 0.0;
#pragma unroll
          for (int r = 0; r < N; ++r) {
            if (r != p && r != q) {
              double arp = a[r][p], arq = a[r][q];
              double nrp = c * arp - s * arq;
              double nrq = s * arp + c * arq;
              a[r][p] = nrp; a[p][r] = nrp;
              a[r][q] = nrq; a[q][r] = nrq;
            }
          }
#pragma unroll
          for (int r = 0; r < N; ++r) {
            double vrp = v[r][p], vrq = v[r][q];
            v[r][p] = c * vrp - s * vrq;
            v[r][q] = s * vrp + c * vrq;
          }
        }
      }
    }
  }
}

__device__ inline double fixed128_to_double(long long hi, unsigned long long lo) {
  bool neg = hi < 0;
  unsigned long long mh = (unsigned long long)hi, ml = lo;
  if (neg) { ml = ~ml + 1ull; mh = ~mh + (ml == 0ull ? 1ull : 0ull); }
  double v = (double)mh * 18446744073709551616.0 + (double)ml;
  v = v * (1.0 / 1073741824.0);
  return neg ? -v : v;
}

// minimal-norm solution of a (numerically) singular A x = b through the eigen-decomposition (the fallback of
// PointToPlaneErrorMinimizer::compute_in_place)
static __device__ __noinline__ void det_solve6_minimal_norm(const double (&A)[6][6], const double (&b)[6], double* x) {
  const double rtol = 6.0 * (double)FLT_EPSILON;
  double a[6][6], v[6][6];
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) a[i][j] = A[i][j];
  det_jacobi<6>(a, v);
  double lmax = 0.0;
  for (int i = 0; i < 6; ++i) { double l = fabs(a[i][i]); if (l > lmax) lmax = l; }
  for (int i = 0; i < 6; ++i) x[i] = 0.0;
  for (int e = 0; e < 6; ++e) {
    double l = a[e][e];
    if (!(l > rtol * lmax)) continue;
    double proj = 0.0;
    for (int i = 0; i < 6; ++i) proj = proj + v[i][e] * b[i];
    double coef = proj / l;
    for (int i = 0; i < 6; ++i) x[i] = x[i] + coef * v[i][e];
  }
}

// sums: 21 upper-triangle entries of A (row-major, i <= j) then 6 entries of g; solves A x = -g.
// returns 1 (LLT) or 2 (minimal-norm fallback)
__device__ inline int det_solve6(const long long* sum_hi, const unsigned long long* sum_lo, double* x) {
  double A[6][6], b[6];
  int s = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j) { double v = fixed128_to_double(sum_hi[s], sum_lo[s]); A[i][j] = v; A[j][i] = v; ++s; }
  for (int i = 0; i < 6; ++i) { b[i] = -fixed128_to_double(sum_hi[s], sum_lo[s]); ++s; }
  const double rtol = 6.0 * (double)FLT_EPSILON;
  double L[6][6];
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) L[i][j] = 0.0;
  bool ok = true;
  for (int j = 0; j < 6 && ok; ++j) {
    double d = A[j][j];
    for (int k = 0; k < j; ++k) d = d - L[j][k] * L[j][k];
    if (!(d > rtol * A[j][j]) || !(d > 0.0)) { ok = false; break; }
    double ljj = sqrt(d);
    L[j][j] = ljj;
    for (int i = j + 1; i < 6; ++i) {
      double v = A[i][j];
      for (int k = 0; k < j; ++k) v = v - L[i][k] * L[j][k];
      L[i][j] = v / ljj;
    }
  }
  if (ok) {
    double y[6];
    for (int i = 0; i < 6; ++i) {
      double v = b[i];
      for (int k = 0; k < i; ++k) v = v - L[i][k] * y[k];
      y[i] = v / L[i][i];
    }
    for (int i = 5; i >= 0; --i) {
      double v = y[i];
      for (int k = i + 1; k < 6; ++k) v = v - L[k][i] * x[k];
      x[i] = v / L[i][i];
    }
    return 1;
  }
  det_solve6_minimal_norm(A, b, x);
  return 2;
}

// warp-cooperative det_solve6: lane s < 28 passes sum s (already converted: fixed128_to_double) in `v`; every lane returns the
// same x.  The LLT runs with one ROW of A and L per lane (lanes 0..5), so the matrices live in 12 registers per lane
// instead of 72 in one thread (which spilled 2.9 KB under the loop kernel's 64-register cap and made the solve the longest
// serial section of an iteration).  Each scalar is produced by exactly the operation sequence of det_solve6 -- same
// operands, same order -- so the result is bit-identical; lanes that hold no row compute along on garbage.
__device__ inline int det_solve6_warp(double v, double* x) {
  const unsigned int FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  const int i = lane < 6 ? lane : 5;
  double A[6], L[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const int a = i < j ? i : j, c = i < j ? j : i;                 // A is symmetric: entry (min, max) of the upper triangle
    A[j] = __shfl_sync(FULL, v, a * 6 - a * (a - 1) / 2 + (c - a));
    L[j] = 0.0;
  }
  const double b = -__shfl_sync(FULL, v, 21 + i);
  const double rtol = 6.0 * (double)FLT_EPSILON;
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    if (!ok) break;                                                 // warp-uniform
    // lane j: d = A[j][j] - sum_k L[j][k]^2; lanes i > j: v = A[i][j] - sum_k L[i][k] L[j][k] -- the same expression
    double acc = A[j];
#pragma unroll
    for (int k = 0; k < j; ++k) acc = acc - L[k] * __shfl_sync(FULL, L[k], j);
    const double ajj = __shfl_sync(FULL, A[j], j);
    const double d = __shfl_sync(FULL, acc, j);
    if (!(d > rtol * ajj) || !(d > 0.0)) { ok = false; break; }
    const double ljj = sqrt(d);
    L[j] = i == j ? ljj : (i > j ? acc / ljj : 0.0);
  }
  if (ok) {
    double y[6], yv = b;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      y[k] = __shfl_sync(FULL, yv / L[k], k);                       // lane k holds b[k] - sum_{m<k} L[k][m] y[m] and L[k][k]
      if (i > k) yv = yv - L[k] * y[k];
    }
#pragma unroll
    for (int r = 5; r >= 0; --r) {
      double vv = y[r];
#pragma unroll
      for (int k = r + 1; k < 6; ++k) vv = vv - __shfl_sync(FULL, L[r], k) * x[k];      // L[k][r] lives in lane k
      x[r] = vv / __shfl_sync(FULL, L[r], r);
    }
    return 1;
  }
  // singular system (rare): every lane rebuilds A and b, the minimal-norm solution is computed redundantly
  double Af[6][6], bf[6];
  int s = 0;
  for (int r = 0; r < 6; ++r)
    for (int c = r; c < 6; ++c) { const double e = __shfl_sync(FULL, v, s); Af[r][c] = e; Af[c][r] = e; ++s; }
  for (int r = 0; r < 6; ++r) { bf[r] = -__shfl_sync(FULL, v, s); ++s; }
  det_solve6_minimal_norm(Af, bf, x);
  return 2;
}

// rotation vector + translation -> column-major float 4x4 (AngleAxis; zero vector -> identity rotation)
__device__ inline void det_pose_increment(const double* x, float* dT) {
  double wx = x[0], wy = x[1], wz = x[2];
  double th2 = (wx * wx + wy * wy) + wz * wz;
  double th = sqrt(th2);
  double R[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  if (th > 0.0) {
    double ux = wx / th, uy = wy / th, uz = wz / th;
    double s, c;
    det_sincos(th, &s, &c);
    double omc = 1.0 - c;
    R[0][0] = c + (ux * ux) * omc;        R[0][1] = (ux * uy) * omc - uz * s;   R[0][2] = (ux * uz) * omc + uy * s;
    R[1][0] = (uy * ux) * omc + uz * s;   R[1][1] = c + (uy * uy) * omc;        R[1][2] = (uy * uz) * omc - ux * s;
    R[2][0] = (uz * ux) * omc - uy * s;   R[2][1] = (uz * uy) * omc + ux * s;   R[2][2] = c + (uz * uz) * omc;
  }
  for (int c4 = 0; c4 < 3; ++c4) {
    for (int r = 0; r < 3; ++r) dT[c4 * 4 + r] = (float)R[r][c4];
    dT[c4 * 4 + 3] = 0.f;
  }
  dT[12] = (float)x[3]; dT[13] = (float)x[4]; dT[14] = (float)x[5]; dT[15] = 1.f;
}

// quaternion (w,x,y,z) of the rotation block of a column-major float 4x4 (Eigen's branch structure), float64
__device__ inline void det_quat_from_T(const float* T, double* q) {
#define AICP_M(r, c) ((double)T[(c) * 4 + (r)])
  double tr = (AICP_M(0, 0) + AICP_M(1, 1)) + AICP_M(2, 2);
  if (tr > 0.0) {
    double t = sqrt(tr + 1.0);
    q[0] = 0.5 * t;
    t = 0.5 / t;
    q[1] = (AICP_M(2, 1) - AICP_M(1, 2)) * t;
    q[2] = (AICP_M(0, 2) - AICP_M(2, 0)) * t;
    q[3] = (AICP_M(1, 0) - AICP_M(0, 1)) * t;
  } else {
    int i = 0;
    if (AICP_M(1, 1) > AICP_M(0, 0)) i = 1;
    if (AICP_M(2, 2) > AICP_M(i, i)) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    double t = sqrt(((AICP_M(i, i) - AICP_M(j, j)) - AICP_M(k, k)) + 1.0);
    q[1 + i] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (AICP_M(k, j) - AICP_M(j, k)) * t;
    q[1 + j] = (AICP_M(j, i) + AICP_M(i, j)) * t;
    q[1 + k] = (AICP_M(k, i) + AICP_M(i, k)) * t;
  }
#undef AICP_M
}

__device__ inline double det_quat_angular_distance(const double* a, const double* b) {
  double w = ((a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]) + a[3] * b[3];
  double vx = ((a[1] * b[0] - a[0] * b[1]) - a[2] * b[3]) + a[3] * b[2];
  double vy = ((a[2] * b[0] - a[0] * b[2]) - a[3] * b[1]) + a[1] * b[3];
  double vz = ((a[3] * b[0] - a[0] * b[3]) - a[1] * b[2]) + a[2] * b[1];
  double vn = sqrt((vx * vx + vy * vy) + vz * vz);
  return 2.0 * det_atan2_pos(vn, fabs(w));
}

}  // namespace aicp
