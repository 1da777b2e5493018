// kjx_small_math.cuh — register-resident d x d math for the small-state-dimension path (d = 2, 3).
//
// Everything here is __host__ __device__ so the same code is unit-tested on the CPU
// (tests/host/host_emulation.cu) and runs inside the sm_100a kernels (kjx_small_kernels.cuh).
//
//  * transition<>        closed-form A(dt) = expm(F dt)       priors.py:163-175 (Matern-3/2), 238-255 (Matern-5/2)
//  * filter_predict/update   one step of SDEGP.kalman_filter  sde_gp.py:276-296  (reference order of operations)
//  * smoother_step       one step of the RTS smoother         sde_gp.py:351-361
//  * FSum / fold / combine / apply   Sarkka & Garcia-Fernandez parallel-scan elements for the filter
//  * SSum / element / combine / apply   same for the smoother            
//
// H is the first unit vector for these priors (priors.py:160, 235), so H m = m[0], H P H^T = P[0][0],
// H P = P[0][:].  Covariances are kept exactly symmetric (upper triangle computed, mirrored): the
// reference does not symmetrise (sde_gp.py:296) but its asymmetry is O(eps), far below the 1e-9 gate.
#pragma once
#include "kjx_sites.cuh"

namespace kjx {

template <typename T> KJX_HD T rsqrt_(T x) {
#ifdef __CUDA_ARCH__
  return rsqrt(x);
#else
  return T(1) / sqrt(x);
#endif
}

template <int KIND> struct BlockDim;
template <> struct BlockDim<KJX_BLOCK_MATERN32> { static constexpr int D = 2; };
template <> struct BlockDim<KJX_BLOCK_MATERN52> { static constexpr int D = 3; };

// ---------------------------------------------------------------------------------------------
// A(dt): priors.py:172-174 and priors.py:246-255 (same expression order as the reference)
template <typename T, int KIND> struct Transition;

template <typename T> struct Transition<T, KJX_BLOCK_MATERN32> {
  static KJX_HD void eval(T dt, T lam, T (&A)[2][2]) {
    const T e = exp(-dt * lam);
    A[0][0] = e * (dt * lam + T(1));
    A[0][1] = e * (dt * T(1));
    A[1][0] = e * (dt * (-(lam * lam)));
    A[1][1] = e * (dt * (-lam) + T(1));
  }
};

template <typename T> struct Transition<T, KJX_BLOCK_MATERN52> {
  static KJX_HD void eval(T dt, T lam, T (&A)[3][3]) {
    const T dtlam = dt * lam;
    const T e = exp(-dtlam);
    const T lam2 = lam * lam;
    A[0][0] = e * (dt * (lam * (T(0.5) * dtlam + T(1))) + T(1));
    A[0][1] = e * (dt * (dtlam + T(1)));
    A[0][2] = e * (dt * (T(0.5) * dt));
    A[1][0] = e * (dt * (T(-0.5) * dtlam * lam2));
    A[1][1] = e * (dt * (lam * (T(1) - dtlam)) + T(1));
    A[1][2] = e * (dt * (T(1) - T(0.5) * dtlam));
    A[2][0] = e * (dt * (lam2 * lam * (T(0.5) * dtlam - T(1))));
    A[2][1] = e * (dt * (lam2 * (dtlam - T(3))));
    A[2][2] = e * (dt * (lam * (T(0.5) * dtlam - T(2))) + T(1));
  }
};

// ---------------------------------------------------------------------------------------------
// tiny dense helpers (fully unrolled; all indices compile-time so everything stays in registers)
template <typename T, int D> KJX_HD void mat_vec(const T (&A)[D][D], const T (&x)[D], T (&y)[D]) {
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T s = A[i][0] * x[0];
#pragma unroll
    for (int j = 1; j < D; ++j) s += A[i][j] * x[j];
    y[i] = s;
  }
}
template <typename T, int D> KJX_HD void mat_mul(const T (&A)[D][D], const T (&B)[D][D], T (&C)[D][D]) {
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      T s = A[i][0] * B[0][j];
#pragma unroll
      for (int k = 1; k < D; ++k) s += A[i][k] * B[k][j];
      C[i][j] = s;
    }
}
// C = A B^T + S, only the upper triangle computed then mirrored (result symmetric by construction)
template <typename T, int D>
KJX_HD void mat_mul_nt_sym_add(const T (&A)[D][D], const T (&B)[D][D], const T (&S)[D][D], T (&C)[D][D]) {
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i; j < D; ++j) {
      T s = S[i][j];
#pragma unroll
      for (int k = 0; k < D; ++k) s += A[i][k] * B[j][k];
      C[i][j] = s;
      C[j][i] = s;
    }
}
// X = A (P - Pinf) A^T + Pinf  (sde_gp.py:278 / :354), P and Pinf symmetric
template <typename T, int D>
KJX_HD void predict_cov(const T (&A)[D][D], const T (&P)[D][D], const T (&Pinf)[D][D], T (&Pp)[D][D]) {
  T X[D][D], Y[D][D];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) X[i][j] = P[i][j] - Pinf[i][j];
  mat_mul<T, D>(A, X, Y);
  mat_mul_nt_sym_add<T, D>(Y, A, Pinf, Pp);
}

// ---------------------------------------------------------------------------------------------
// Kalman filter step (sde_gp.py:276-300), H = e0
template <typename T, int D> struct FiltState { T m[D]; T P[D][D]; };

template <typename T, int D>
KJX_HD void filter_predict(const T (&A)[D][D], const T (&Pinf)[D][D], FiltState<T, D>& s) {
  T mp[D], Pp[D][D];
  mat_vec<T, D>(A, s.m, mp);                       // :277
  predict_cov<T, D>(A, s.P, Pinf, Pp);             // :278
#pragma unroll
  for (int i = 0; i < D; ++i) {
    s.m[i] = mp[i];
#pragma unroll
    for (int j = 0; j < D; ++j) s.P[i][j] = Pp[i][j];
  }
}
// s holds the predicted (m_, P_) on entry, the filtered (m, P) on exit
template <typename T, int D>
KJX_HD void filter_update(FiltState<T, D>& s, T site_mean, T site_var) {
  const T S = s.P[0][0] + site_var;                // :292
  const T Sinv = inv1(S);                          // :294 solve(S, HP): 1x1 Cholesky, NaN if S < 0
  const T innov = site_mean - s.m[0];
  T K[D], HP[D];
#pragma unroll
  for (int i = 0; i < D; ++i) { HP[i] = s.P[0][i]; K[i] = HP[i] * Sinv; }
#pragma unroll
  for (int i = 0; i < D; ++i) s.m[i] += K[i] * innov;    // :295
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i; j < D; ++j) {                  // :296
      const T v = s.P[i][j] - K[i] * HP[j];
      s.P[i][j] = v;
      s.P[j][i] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// Solve Pp X = B with the Cholesky factorisation of Pp (utils.py:25-30). A non-positive pivot gives
// NaN/inf exactly like a failed LAPACK factorisation propagated by jaxlib.
template <typename T, int D>
KJX_HD void chol_solve(const T (&Pp)[D][D], const T (&B)[D][D], T (&X)[D][D]) {
  T L[D][D], rinv[D];
#pragma unroll
  for (int j = 0; j < D; ++j) {
    T piv = Pp[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) piv -= L[j][k] * L[j][k];
    rinv[j] = rsqrt_(piv);
    L[j][j] = piv * rinv[j];
#pragma unroll
    for (int i = j + 1; i < D; ++i) {
      T v = Pp[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
      L[i][j] = v * rinv[j];
    }
  }
#pragma unroll
  for (int c = 0; c < D; ++c) {
    T z[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      T v = B[i][c];
#pragma unroll
      for (int k = 0; k < i; ++k) v -= L[i][k] * z[k];
      z[i] = v * rinv[i];
    }
#pragma unroll
    for (int i = D - 1; i >= 0; --i) {
      T v = z[i];
#pragma unroll
      for (int k = i + 1; k < D; ++k) v -= L[k][i] * X[k][c];
      X[i][c] = v * rinv[i];
    }
  }
}

// Quantities shared by the RTS step and the smoother scan element for filtered (mf, Pf) and the
// transition A to the NEXT step: m_pred, P_pred, and Gt = solve(P_pred, A Pf) (sde_gp.py:352-359).
template <typename T, int D> struct SmoothGain { T mp[D]; T Pp[D][D]; T Gt[D][D]; T AP[D][D]; };

template <typename T, int D>
KJX_HD void smoother_gain(const T (&A)[D][D], const T (&Pinf)[D][D], const T (&mf)[D], const T (&Pf)[D][D],
                          SmoothGain<T, D>& g) {
  mat_vec<T, D>(A, mf, g.mp);                      // :352
  mat_mul<T, D>(A, Pf, g.AP);                      // :353
  predict_cov<T, D>(A, Pf, Pinf, g.Pp);            // :354
  chol_solve<T, D>(g.Pp, g.AP, g.Gt);              // :359
}

// RTS step (sde_gp.py:360-361): (ms, Ps) is the smoothed state at n+1 on entry, at n on exit
template <typename T, int D>
KJX_HD void smoother_step(const SmoothGain<T, D>& g, const T (&mf)[D], const T (&Pf)[D][D], FiltState<T, D>& s) {
  T dm[D], dP[D][D], W[D][D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    dm[i] = s.m[i] - g.mp[i];
#pragma unroll
    for (int j = 0; j < D; ++j) dP[i][j] = s.P[i][j] - g.Pp[i][j];
  }
  // W = G dP with G = Gt^T :  W[i][j] = sum_k Gt[k][i] dP[k][j]
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      T v = g.Gt[0][i] * dP[0][j];
#pragma unroll
      for (int k = 1; k < D; ++k) v += g.Gt[k][i] * dP[k][j];
      W[i][j] = v;
    }
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T v = mf[i];
#pragma unroll
    for (int k = 0; k < D; ++k) v += g.Gt[k][i] * dm[k];
    s.m[i] = v;                                    // :360
  }
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i; j < D; ++j) {                  // :361  P = Pf + (G dP) G^T
      T v = Pf[i][j];
#pragma unroll
      for (int k = 0; k < D; ++k) v += W[i][k] * g.Gt[k][j];
      s.P[i][j] = v;
      s.P[j][i] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// Filter scan element / chunk summary: given the state x_in at the chunk start,
//   p(x_end | x_in, data in chunk) = N(A x_in + b, C),   p(data in chunk | x_in) ∝ N_info(x_in; eta, J).
template <typename T, int D> struct FSum {
  T A[D][D]; T b[D]; T C[D][D]; T eta[D]; T J[D][D];
  static constexpr int NS = D * D + D + D * (D + 1) / 2 + D + D * (D + 1) / 2;
};

template <typename T, int D> KJX_HD void fsum_identity(FSum<T, D>& s) {
#pragma unroll
  for (int i = 0; i < D; ++i) {
    s.b[i] = T(0); s.eta[i] = T(0);
#pragma unroll
    for (int j = 0; j < D; ++j) { s.A[i][j] = (i == j) ? T(1) : T(0); s.C[i][j] = T(0); s.J[i][j] = T(0); }
  }
}
template <typename T, int D> KJX_HD void fsum_pack(const FSum<T, D>& s, T* v) {
  int n = 0;
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) v[n++] = s.A[i][j];
#pragma unroll
  for (int i = 0; i < D; ++i) v[n++] = s.b[i];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i; j < D; ++j) v[n++] = s.C[i][j];
#pragma unroll
  for (int i = 0; i < D; ++i) v[n++] = s.eta[i];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i; j < D; ++j) v[n++] = s.J[i][j];
}
template <typename T, int D> KJX_HD void fsum_unpack(const T* v, FSum<T, D>& s) {
  int n = 0;
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) s.A[i][j] = v[n++];
#pragma unroll
  for (int i = 0; i < D; ++i) s.b[i] = v[n++];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i; j < D; ++j) { s.C[i][j] = v[n]; s.C[j][i] = v[n]; ++n; }
#pragma unroll
  for (int i = 0; i < D; ++i) s.eta[i] = v[n++];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i; j < D; ++j) { s.J[i][j] = v[n]; s.J[j][i] = v[n]; ++n; }
}

// Fold one more time step (transition F, then observation (ytilde, R) unless masked) onto a running
// summary: O(d^3) for the predict, O(d^2) rank-one for the update (f = 1, H = e0).
template <typename T, int D>
KJX_HD void fsum_fold_step(FSum<T, D>& s, const T (&F)[D][D], const T (&Pinf)[D][D], T ytilde, T R, bool masked) {
  T A2[D][D], b2[D], C2[D][D];
  mat_mul<T, D>(F, s.A, A2);
  mat_vec<T, D>(F, s.b, b2);
  predict_cov<T, D>(F, s.C, Pinf, C2);             // F C F^T + Q with Q = Pinf - F Pinf F^T
  if (!masked) {
    const T sv = C2[0][0] + R;
    const T sinv = inv1(sv);
    const T res = ytilde - b2[0];
    T a[D], K[D], c[D];
#pragma unroll
    for (int i = 0; i < D; ++i) { a[i] = A2[0][i]; c[i] = C2[0][i]; K[i] = c[i] * sinv; }
#pragma unroll
    for (int i = 0; i < D; ++i) {
      s.eta[i] += a[i] * (res * sinv);
#pragma unroll
      for (int j = i; j < D; ++j) { const T v = s.J[i][j] + a[i] * a[j] * sinv; s.J[i][j] = v; s.J[j][i] = v; }
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
      s.b[i] = b2[i] + K[i] * res;
#pragma unroll
      for (int j = 0; j < D; ++j) s.A[i][j] = A2[i][j] - K[i] * a[j];
#pragma unroll
      for (int j = i; j < D; ++j) { const T v = C2[i][j] - K[i] * c[j]; s.C[i][j] = v; s.C[j][i] = v; }
    }
  } else {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      s.b[i] = b2[i];
#pragma unroll
      for (int j = 0; j < D; ++j) { s.A[i][j] = A2[i][j]; s.C[i][j] = C2[i][j]; }
    }
  }
}

// Inverse of a general d x d matrix by Gauss-Jordan elimination with partial pivoting (predicated
// row swaps, static indices).  Used on M = I + C_i J_j whose eigenvalues are real and >= 1.
template <typename T, int D> KJX_HD void inv_gj(T (&M)[D][D], T (&R)[D][D]) {
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) R[i][j] = (i == j) ? T(1) : T(0);
#pragma unroll
  for (int k = 0; k < D; ++k) {
#pragma unroll
    for (int r = k + 1; r < D; ++r) {
      const bool sw = fabs(M[r][k]) > fabs(M[k][k]);
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const T a = M[k][c], b = M[r][c];
        M[k][c] = sw ? b : a; M[r][c] = sw ? a : b;
        const T p = R[k][c], q = R[r][c];
        R[k][c] = sw ? q : p; R[r][c] = sw ? p : q;
      }
    }
    const T pinv = T(1) / M[k][k];
#pragma unroll
    for (int c = 0; c < D; ++c) { M[k][c] *= pinv; R[k][c] *= pinv; }
#pragma unroll
    for (int r = 0; r < D; ++r) {
      if (r == k) continue;
      const T f = M[r][k];
#pragma unroll
      for (int c = 0; c < D; ++c) { M[r][c] -= f * M[k][c]; R[r][c] -= f * R[k][c]; }
    }
  }
}

// Inverse of M = I + C J (C, J symmetric positive semi-definite: real eigenvalues >= 1, det >= 1) in closed form —
// adjugate over determinant, one reciprocal.  The scan's combines and boundary applications are pure latency (one warp
// walks a dependent chain of them), and Gauss-Jordan's three dependent pivot searches + reciprocals were a third of each;
// the forward error of the closed form is of the same order (condition number x epsilon) for these matrices, and the
// scan only sets the state entering a chunk — the recursions inside the chunks are the reference's.
template <typename T, int D> struct InvSmall {
  static KJX_HD void run(T (&M)[D][D], T (&R)[D][D]) { inv_gj<T, D>(M, R); }
};
template <typename T> struct InvSmall<T, 2> {
  static KJX_HD void run(T (&M)[2][2], T (&R)[2][2]) {
    const T rdet = T(1) / (M[0][0] * M[1][1] - M[0][1] * M[1][0]);
    R[0][0] = M[1][1] * rdet; R[0][1] = -M[0][1] * rdet;
    R[1][0] = -M[1][0] * rdet; R[1][1] = M[0][0] * rdet;
  }
};
template <typename T> struct InvSmall<T, 3> {
  static KJX_HD void run(T (&M)[3][3], T (&R)[3][3]) {
    const T c00 = M[1][1] * M[2][2] - M[1][2] * M[2][1];
    const T c01 = M[1][2] * M[2][0] - M[1][0] * M[2][2];
    const T c02 = M[1][0] * M[2][1] - M[1][1] * M[2][0];
    const T rdet = T(1) / (M[0][0] * c00 + M[0][1] * c01 + M[0][2] * c02);
    R[0][0] = c00 * rdet; R[1][0] = c01 * rdet; R[2][0] = c02 * rdet;
    R[0][1] = (M[0][2] * M[2][1] - M[0][1] * M[2][2]) * rdet;
    R[1][1] = (M[0][0] * M[2][2] - M[0][2] * M[2][0]) * rdet;
    R[2][1] = (M[0][1] * M[2][0] - M[0][0] * M[2][1]) * rdet;
    R[0][2] = (M[0][1] * M[1][2] - M[0][2] * M[1][1]) * rdet;
    R[1][2] = (M[0][2] * M[1][0] - M[0][0] * M[1][2]) * rdet;
    R[2][2] = (M[0][0] * M[1][1] - M[0][1] * M[1][0]) * rdet;
  }
};

// out = i (earlier) ⊗ j (later)          (Sarkka & Garcia-Fernandez 2021, Lemma 8)
template <typename T, int D>
KJX_HDN void fsum_combine(const FSum<T, D>& si, const FSum<T, D>& sj, FSum<T, D>& out) {
  T M[D][D], Minv[D][D];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      T v = (i == j) ? T(1) : T(0);
#pragma unroll
      for (int k = 0; k < D; ++k) v += si.C[i][k] * sj.J[k][j];
      M[i][j] = v;
    }
  InvSmall<T, D>::run(M, Minv);
  T T1[D][D], T2[D][D], W[D][D], u[D], w[D];
  mat_mul<T, D>(sj.A, Minv, T1);                   // T1 = A_j M^-1
  // u = b_i + C_i eta_j ; b = T1 u + b_j
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T v = si.b[i];
#pragma unroll
    for (int k = 0; k < D; ++k) v += si.C[i][k] * sj.eta[k];
    u[i] = v;
  }
  mat_vec<T, D>(T1, u, w);
  FSum<T, D> r;
#pragma unroll
  for (int i = 0; i < D; ++i) r.b[i] = w[i] + sj.b[i];
  mat_mul<T, D>(T1, si.A, r.A);                    // A = A_j M^-1 A_i
  mat_mul<T, D>(T1, si.C, W);                      // W = A_j M^-1 C_i
  mat_mul_nt_sym_add<T, D>(W, sj.A, sj.C, r.C);    // C = W A_j^T + C_j
  // T2 = A_i^T M^-T  (T2[i][j] = sum_k A_i[k][i] Minv[j][k])
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      T v = si.A[0][i] * Minv[j][0];
#pragma unroll
      for (int k = 1; k < D; ++k) v += si.A[k][i] * Minv[j][k];
      T2[i][j] = v;
    }
  // eta = T2 (eta_j - J_j b_i) + eta_i
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T v = sj.eta[i];
#pragma unroll
    for (int k = 0; k < D; ++k) v -= sj.J[i][k] * si.b[k];
    u[i] = v;
  }
  mat_vec<T, D>(T2, u, w);
#pragma unroll
  for (int i = 0; i < D; ++i) r.eta[i] = w[i] + si.eta[i];
  // J = T2 J_j A_i + J_i  (symmetric)
  mat_mul<T, D>(T2, sj.J, W);
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i; j < D; ++j) {
      T v = si.J[i][j];
#pragma unroll
      for (int k = 0; k < D; ++k) v += W[i][k] * si.A[k][j];
      r.J[i][j] = v; r.J[j][i] = v;
    }
  out = r;
}

// Filtered state after a chunk with summary s, entered with state (m, P):
//   m' = A (I + P J)^-1 (m + P eta) + b ,  P' = A (I + P J)^-1 P A^T + C
template <typename T, int D>
KJX_HDN void fsum_apply(const FSum<T, D>& s, FiltState<T, D>& x) {
  T M[D][D], Minv[D][D], T1[D][D], W[D][D], u[D], w[D];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      T v = (i == j) ? T(1) : T(0);
#pragma unroll
      for (int k = 0; k < D; ++k) v += x.P[i][k] * s.J[k][j];
      M[i][j] = v;
    }
  InvSmall<T, D>::run(M, Minv);
  mat_mul<T, D>(s.A, Minv, T1);
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T v = x.m[i];
#pragma unroll
    for (int k = 0; k < D; ++k) v += x.P[i][k] * s.eta[k];
    u[i] = v;
  }
  mat_vec<T, D>(T1, u, w);
  mat_mul<T, D>(T1, x.P, W);
  T Pn[D][D];
  mat_mul_nt_sym_add<T, D>(W, s.A, s.C, Pn);
#pragma unroll
  for (int i = 0; i < D; ++i) {
    x.m[i] = w[i] + s.b[i];
#pragma unroll
    for (int j = 0; j < D; ++j) x.P[i][j] = Pn[i][j];
  }
}

// ---------------------------------------------------------------------------------------------
// "Backward information" form of the smoother boundary (two-filter identity).  The (eta, J) part of
// the combination of all filter elements AFTER step k is the likelihood of every later observation as
// a function of x_k:  p(y_{k+1:N} | x_k) ∝ exp(-x'Jx/2 + eta'x).  Multiplying it into the filtered
// N(m_k, P_k) gives the smoothed marginal of step k — mathematically identical to what the RTS
// recursion (sde_gp.py:360-361) delivers there — so the smoother needs no scan of its own when it
// runs right after the filter scan: a SUFFIX scan over the same elements provides its boundary states.
template <typename T, int D> struct Info { T eta[D]; T J[D][D]; };

template <typename T, int D> KJX_HD void info_zero(Info<T, D>& f) {
#pragma unroll
  for (int i = 0; i < D; ++i) {
    f.eta[i] = T(0);
#pragma unroll
    for (int j = 0; j < D; ++j) f.J[i][j] = T(0);
  }
}
template <typename T, int D> KJX_HD void info_of(const FSum<T, D>& s, Info<T, D>& f) {
#pragma unroll
  for (int i = 0; i < D; ++i) {
    f.eta[i] = s.eta[i];
#pragma unroll
    for (int j = 0; j < D; ++j) f.J[i][j] = s.J[i][j];
  }
}
// information of (si ⊗ later) about the state entering si, given only the information part of `later`
template <typename T, int D>
KJX_HDN void fsum_combine_info(const FSum<T, D>& si, const Info<T, D>& fj, Info<T, D>& out) {
  T M[D][D], Minv[D][D], T2[D][D], W[D][D], u[D], w[D];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      T v = (i == j) ? T(1) : T(0);
#pragma unroll
      for (int k = 0; k < D; ++k) v += si.C[i][k] * fj.J[k][j];
      M[i][j] = v;
    }
  InvSmall<T, D>::run(M, Minv);
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      T v = si.A[0][i] * Minv[j][0];
#pragma unroll
      for (int k = 1; k < D; ++k) v += si.A[k][i] * Minv[j][k];
      T2[i][j] = v;
    }
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T v = fj.eta[i];
#pragma unroll
    for (int k = 0; k < D; ++k) v -= fj.J[i][k] * si.b[k];
    u[i] = v;
  }
  mat_vec<T, D>(T2, u, w);
  mat_mul<T, D>(T2, fj.J, W);
  Info<T, D> r;
#pragma unroll
  for (int i = 0; i < D; ++i) r.eta[i] = w[i] + si.eta[i];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i; j < D; ++j) {
      T v = si.J[i][j];
#pragma unroll
      for (int k = 0; k < D; ++k) v += W[i][k] * si.A[k][j];
      r.J[i][j] = v; r.J[j][i] = v;
    }
  out = r;
}
// smoothed (m, P) of a step from its filtered (m, P) and the information of everything after it:
//   P_s = (I + P J)^-1 P ,  m_s = (I + P J)^-1 (m + P eta)
template <typename T, int D>
KJX_HDN void info_smooth(const Info<T, D>& f, FiltState<T, D>& x) {
  T M[D][D], Minv[D][D], u[D], w[D], Pn[D][D];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      T v = (i == j) ? T(1) : T(0);
#pragma unroll
      for (int k = 0; k < D; ++k) v += x.P[i][k] * f.J[k][j];
      M[i][j] = v;
    }
  InvSmall<T, D>::run(M, Minv);
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T v = x.m[i];
#pragma unroll
    for (int k = 0; k < D; ++k) v += x.P[i][k] * f.eta[k];
    u[i] = v;
  }
  mat_vec<T, D>(Minv, u, w);
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i; j < D; ++j) {
      T v = Minv[i][0] * x.P[0][j];
#pragma unroll
      for (int k = 1; k < D; ++k) v += Minv[i][k] * x.P[k][j];
      Pn[i][j] = v; Pn[j][i] = v;
    }
#pragma unroll
  for (int i = 0; i < D; ++i) {
    x.m[i] = w[i];
#pragma unroll
    for (int j = 0; j < D; ++j) x.P[i][j] = Pn[i][j];
  }
}

// ---------------------------------------------------------------------------------------------
// Smoother scan element / chunk summary: smoothed(start of chunk) = E smoothed(next chunk) + g, cov E P E^T + L
template <typename T, int D> struct SSum {
  T E[D][D]; T g[D]; T L[D][D];
  static constexpr int NS = D * D + D + D * (D + 1) / 2;
};
template <typename T, int D> KJX_HD void ssum_identity(SSum<T, D>& s) {
#pragma unroll
  for (int i = 0; i < D; ++i) {
    s.g[i] = T(0);
#pragma unroll
    for (int j = 0; j < D; ++j) { s.E[i][j] = (i == j) ? T(1) : T(0); s.L[i][j] = T(0); }
  }
}
template <typename T, int D> KJX_HD void ssum_pack(const SSum<T, D>& s, T* v) {
  int n = 0;
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) v[n++] = s.E[i][j];
#pragma unroll
  for (int i = 0; i < D; ++i) v[n++] = s.g[i];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i; j < D; ++j) v[n++] = s.L[i][j];
}
template <typename T, int D> KJX_HD void ssum_unpack(const T* v, SSum<T, D>& s) {
  int n = 0;
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) s.E[i][j] = v[n++];
#pragma unroll
  for (int i = 0; i < D; ++i) s.g[i] = v[n++];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i; j < D; ++j) { s.L[i][j] = v[n]; s.L[j][i] = v[n]; ++n; }
}

// Element of one step from the gain quantities: E = G, g = mf - G m_pred, L = Pf - G (A Pf)
template <typename T, int D>
KJX_HD void ssum_element(const SmoothGain<T, D>& g, const T (&mf)[D], const T (&Pf)[D][D], SSum<T, D>& e) {
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T v = mf[i];
#pragma unroll
    for (int k = 0; k < D; ++k) { e.E[i][k] = g.Gt[k][i]; v -= g.Gt[k][i] * g.mp[k]; }
    e.g[i] = v;
  }
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i; j < D; ++j) {
      T v = Pf[i][j];
#pragma unroll
      for (int k = 0; k < D; ++k) v -= g.Gt[k][i] * g.AP[k][j];
      e.L[i][j] = v; e.L[j][i] = v;
    }
}
// terminal element of the series (step N-1): E = 0, g = mf, L = Pf
template <typename T, int D>
KJX_HD void ssum_terminal(const T (&mf)[D], const T (&Pf)[D][D], SSum<T, D>& e) {
#pragma unroll
  for (int i = 0; i < D; ++i) {
    e.g[i] = mf[i];
#pragma unroll
    for (int j = 0; j < D; ++j) { e.E[i][j] = T(0); e.L[i][j] = Pf[i][j]; }
  }
}
// out = i (earlier) ⊗ j (later):  E = E_i E_j, g = E_i g_j + g_i, L = E_i L_j E_i^T + L_i
template <typename T, int D>
KJX_HDN void ssum_combine(const SSum<T, D>& si, const SSum<T, D>& sj, SSum<T, D>& out) {
  SSum<T, D> r;
  T w[D], W[D][D];
  mat_mul<T, D>(si.E, sj.E, r.E);
  mat_vec<T, D>(si.E, sj.g, w);
#pragma unroll
  for (int i = 0; i < D; ++i) r.g[i] = w[i] + si.g[i];
  mat_mul<T, D>(si.E, sj.L, W);
  mat_mul_nt_sym_add<T, D>(W, si.E, si.L, r.L);
  out = r;
}
// smoothed state at the chunk start from the smoothed state x entering from the right
template <typename T, int D>
KJX_HDN void ssum_apply(const SSum<T, D>& s, FiltState<T, D>& x) {
  T w[D], W[D][D], Pn[D][D];
  mat_vec<T, D>(s.E, x.m, w);
  mat_mul<T, D>(s.E, x.P, W);
  mat_mul_nt_sym_add<T, D>(W, s.E, s.L, Pn);
#pragma unroll
  for (int i = 0; i < D; ++i) {
    x.m[i] = w[i] + s.g[i];
#pragma unroll
    for (int j = 0; j < D; ++j) x.P[i][j] = Pn[i][j];
  }
}

}  // namespace kjx
