// bfgs.h -- the 6-DoF solve of GICP: inner optimiser (BFGS exactly as the
// reference drives it, or Gauss-Newton) and the outer convergence loop.
//
// Reference: estimateRigidTransformationBFGS gicp.hpp:217-287 (driver),
// pcl::BFGS (pcl/registration/bfgs.h, PCL 1.10 -- GSL vector_bfgs2 + Fletcher
// line search; parameters set at gicp.hpp:253-258), outer loop gicp.hpp:445-583.
//
// Everything is templated on a Backend that supplies the two data-parallel
// steps, so the SAME control code runs
//   * inside the persistent cooperative kernel (every thread of every CTA
//     executes this scalar code redundantly on bitwise-identical reduced sums;
//     the Backend methods are grid-wide collectives), and
//   * on the host thread in host-driven mode (Backend methods launch kernels).
// Backend concept:
//   int  correspond(const float* T34, const double* R9);   // K4: returns #correspondences m
//   void fdf(const double* x, double* f, double* g6);       // K5: objective + gradient at x
//   int  gn(const double* x, double* f, double* b6, double* H21);   // GN normal equations
#pragma once

#include "hd.h"

namespace lb {

enum { BFGS_RUNNING = -1, BFGS_SUCCESS = 0, BFGS_NO_PROGRESS = 1 };

struct SolveStats {
  int n_evals;        // objective evaluations (fused f+g passes over the correspondences)
  int n_inner;        // inner (BFGS / GN) iterations
};

template <class Backend>
struct Bfgs6 {
  // parameters (gicp.hpp:253-258; pcl::BFGS::Parameters defaults for the rest)
  static constexpr int N = 6;
  int bracket_iters, section_iters, order;
  double rho, sigma, tau1, tau2, tau3, step_size;
  // state
  Backend* be;
  double f, gradient[N];
  double delta_f, fp0;
  double x0[N], dx0[N], dg0[N], g0[N], p[N];
  double pnorm, g0norm;
  double f_alpha, df_alpha, x_alpha[N], g_alpha[N];
  double f_cache_key, df_cache_key, x_cache_key, g_cache_key;
  int n_evals;

  LB_HD static double dot(const double* a, const double* b) {
    double s = 0.0;
    for (int i = 0; i < N; i++) s += a[i] * b[i];
    return s;
  }
  LB_HD static double norm(const double* a) { return sqrt(dot(a, a)); }

  LB_HD void init(Backend* b) {
    be = b;
    bracket_iters = 100; section_iters = 100; order = 3;
    rho = 0.01; sigma = 0.01; tau1 = 9; tau2 = 0.05; tau3 = 0.5; step_size = 1;
    n_evals = 0;
  }

  LB_HD void evaluate(const double* x, double* fo, double* go) { be->fdf(x, fo, go); n_evals++; }

  LB_HD void move_to(double alpha) {
    for (int i = 0; i < N; i++) x_alpha[i] = x0[i] + alpha * p[i];
    x_cache_key = alpha;
  }
  LB_HD double slope() const { return dot(g_alpha, p); }

  // The reference evaluates f-only (operator()) and gradient-only (df) passes
  // separately; the fused kernel returns both from one pass with identical
  // arithmetic, so values are the same and only the number of passes drops.
  LB_HD double apply_f(double alpha) {
    if (alpha == f_cache_key) return f_alpha;
    move_to(alpha);
    evaluate(x_alpha, &f_alpha, g_alpha);
    f_cache_key = alpha;
    g_cache_key = alpha;
    return f_alpha;
  }
  LB_HD double apply_df(double alpha) {
    if (alpha == df_cache_key) return df_alpha;
    move_to(alpha);
    if (alpha != g_cache_key) {
      evaluate(x_alpha, &f_alpha, g_alpha);
      f_cache_key = alpha;
      g_cache_key = alpha;
    }
    df_alpha = slope();
    df_cache_key = alpha;
    return df_alpha;
  }
  LB_HD void apply_fdf(double alpha, double* fo, double* dfo) {
    if (alpha == f_cache_key && alpha == df_cache_key) { *fo = f_alpha; *dfo = df_alpha; return; }
    if (alpha == f_cache_key || alpha == df_cache_key) {
      *fo = apply_f(alpha);
      *dfo = apply_df(alpha);
      return;
    }
    move_to(alpha);
    evaluate(x_alpha, &f_alpha, g_alpha);
    f_cache_key = alpha;
    g_cache_key = alpha;
    df_alpha = slope();
    df_cache_key = alpha;
    *fo = f_alpha; *dfo = df_alpha;
  }

  LB_HD int minimize_init(double* x) {
    delta_f = 0;
    evaluate(x, &f, gradient);
    for (int i = 0; i < N; i++) { x0[i] = x[i]; g0[i] = gradient[i]; }
    g0norm = norm(g0);
    for (int i = 0; i < N; i++) p[i] = gradient[i] * -1 / g0norm;
    pnorm = norm(p);
    fp0 = -g0norm;
    for (int i = 0; i < N; i++) { x_alpha[i] = x0[i]; g_alpha[i] = g0[i]; }
    x_cache_key = 0; f_alpha = f; f_cache_key = 0; g_cache_key = 0;
    df_alpha = slope(); df_cache_key = 0;
    return -2;
  }

  LB_HD static double poly_eval4(const double* c, double x) {
    if (x * x <= 1.0) {
      double v = c[3];
      v = v * x + c[2];
      v = v * x + c[1];
      v = v * x + c[0];
      return v;
    }
    double v = c[0];
    double inv_x = 1.0 / x;
    for (int i = 1; i < 4; i++) v = v * inv_x + c[i];
    return pow(x, 3.0) * v;
  }
  LB_HD static void check_extremum(const double* c, double x, double* xmin, double* fmin) {
    double y = poly_eval4(c, x);
    if (y < *fmin) { *xmin = x; *fmin = y; }
  }

  LB_HD static double interpolate(double a, double fa, double fpa, double b, double fb, double fpb,
                                  double xmin, double xmax, int order) {
    double y, ymin, ymax, fmin;
    ymin = (xmin - a) / (b - a);
    ymax = (xmax - a) / (b - a);
    if (ymin > ymax) { double tmp = ymin; ymin = ymax; ymax = tmp; }
    const double inf = (double)INFINITY;
    if (order > 2 && !(fpb != fpb) && fpb != inf) {
      fpa = fpa * (b - a);
      fpb = fpb * (b - a);
      double eta = 3 * (fb - fa) - 2 * fpa - fpb;
      double xi = fpa + fpb - 2 * (fb - fa);
      double c[4] = {fa, fpa, eta, xi};
      y = ymin;
      fmin = poly_eval4(c, ymin);
      check_extremum(c, ymax, &y, &fmin);
      double p0 = c[1], p1 = 2 * c[2], p2 = 3 * c[3];
      double a2 = 2 * p2;
      double disc = (p1 * p1) - (4 * p0 * p2);
      if (0.0 < disc) {
        double dr = sqrt(disc);
        double y0 = (-p1 - dr) / a2;
        double y1 = (-p1 + dr) / a2;
        if (y0 > y1) { double tmp = y0; y0 = y1; y1 = tmp; }
        if (y0 > ymin && y0 < ymax) check_extremum(c, y0, &y, &fmin);
        if (y1 > ymin && y1 < ymax) check_extremum(c, y1, &y, &fmin);
      } else if (0.0 == disc) {
        double y0 = -p1 / a2;
        if (y0 > ymin && y0 < ymax) check_extremum(c, y0, &y, &fmin);
      }
    } else {
      fpa = fpa * (b - a);
      double fl = fa + ymin * (fpa + ymin * (fb - fa - fpa));
      double fh = fa + ymax * (fpa + ymax * (fb - fa - fpa));
      double c = 2 * (fb - fa - fpa);
      y = ymin; fmin = fl;
      if (fh < fmin) { y = ymax; fmin = fh; }
      if (c > a) {  // sic: pcl::BFGS compares the curvature with a (GSL: > 0)
        double z = -fpa / c;
        if (z > ymin && z < ymax) {
          double fz = fa + z * (fpa + z * (fb - fa - fpa));
          if (fz < fmin) { y = z; fmin = fz; }
        }
      }
    }
    return a + y * (b - a);
  }

  LB_HD int line_search(double alpha1, double* alpha_new) {
    double f0, fp0_, falpha, falpha_prev, fpalpha, fpalpha_prev, delta, alpha_next;
    double alpha = alpha1, alpha_prev = 0.0;
    double a, b, fa, fb, fpa, fpb;
    const double nan_ = (double)NAN;
    int i = 0;
    apply_fdf(0.0, &f0, &fp0_);
    falpha_prev = f0;
    fpalpha_prev = fp0_;
    a = 0.0; b = alpha;
    fa = f0; fb = 0.0;
    fpa = fp0_; fpb = 0.0;
    while (i++ < bracket_iters) {
      falpha = apply_f(alpha);
      if (falpha > f0 + alpha * rho * fp0_ || falpha >= falpha_prev) {
        a = alpha_prev; fa = falpha_prev; fpa = fpalpha_prev;
        b = alpha; fb = falpha; fpb = nan_;
        break;
      }
      fpalpha = apply_df(alpha);
      if (fabs(fpalpha) <= -sigma * fp0_) { *alpha_new = alpha; return BFGS_SUCCESS; }
      if (fpalpha >= 0) {
        a = alpha; fa = falpha; fpa = fpalpha;
        b = alpha_prev; fb = falpha_prev; fpb = fpalpha_prev;
        break;
      }
      delta = alpha - alpha_prev;
      {
        double lower = alpha + delta;
        double upper = alpha + tau1 * delta;
        alpha_next = interpolate(alpha_prev, falpha_prev, fpalpha_prev, alpha, falpha, fpalpha, lower, upper, order);
      }
      alpha_prev = alpha;
      falpha_prev = falpha;
      fpalpha_prev = fpalpha;
      alpha = alpha_next;
    }
    while (i++ < section_iters) {
      delta = b - a;
      {
        double lower = a + tau2 * delta;
        double upper = b - tau3 * delta;
        alpha = interpolate(a, fa, fpa, b, fb, fpb, lower, upper, order);
      }
      falpha = apply_f(alpha);
      if ((a - alpha) * fpa <= 2.220446049250313e-16) return BFGS_NO_PROGRESS;
      if (falpha > f0 + rho * alpha * fp0_ || falpha >= fa) {
        b = alpha; fb = falpha; fpb = nan_;
      } else {
        fpalpha = apply_df(alpha);
        if (fabs(fpalpha) <= -sigma * fp0_) { *alpha_new = alpha; return BFGS_SUCCESS; }
        if (((b - a) >= 0 && fpalpha >= 0) || ((b - a) <= 0 && fpalpha <= 0)) {
          b = a; fb = fa; fpb = fpa;
          a = alpha; fa = falpha; fpa = fpalpha;
        } else {
          a = alpha; fa = falpha; fpa = fpalpha;
        }
      }
    }
    return BFGS_SUCCESS;
  }

  LB_HD int minimize_one_step(double* x) {
    double alpha = 0.0, alpha1;
    double f0 = f;
    if (pnorm == 0.0 || g0norm == 0.0 || fp0 == 0) return BFGS_NO_PROGRESS;
    if (delta_f < 0) {
      double del = fmax(-delta_f, 10 * 2.220446049250313e-16 * fabs(f0));
      alpha1 = fmin(1.0, 2.0 * del / (-fp0));
    } else {
      alpha1 = fabs(step_size);
    }
    int status = line_search(alpha1, &alpha);
    if (status != BFGS_SUCCESS) return status;
    // updatePosition
    {
      double fa_, dfa_;
      apply_fdf(alpha, &fa_, &dfa_);
      f = f_alpha;
      for (int i = 0; i < N; i++) { x[i] = x_alpha[i]; gradient[i] = g_alpha[i]; }
    }
    delta_f = f - f0;
    {
      double dxg, dgg, dxdg, dgnorm, A, B;
      for (int i = 0; i < N; i++) dx0[i] = x[i] - x0[i];
      for (int i = 0; i < N; i++) dg0[i] = gradient[i] - g0[i];
      dxg = dot(dx0, gradient);
      dgg = dot(dg0, gradient);
      dxdg = dot(dx0, dg0);
      dgnorm = norm(dg0);
      if (dxdg != 0) {
        B = dxg / dxdg;
        A = -(1.0 + dgnorm * dgnorm / dxdg) * B + dgg / dxdg;
      } else {
        B = 0; A = 0;
      }
      for (int i = 0; i < N; i++) p[i] = -A * dx0[i];
      for (int i = 0; i < N; i++) p[i] += gradient[i];
      for (int i = 0; i < N; i++) p[i] += -B * dg0[i];
    }
    for (int i = 0; i < N; i++) { g0[i] = gradient[i]; x0[i] = x[i]; }
    g0norm = norm(g0);
    pnorm = norm(p);
    double dir = (dot(p, gradient) > 0) ? -1.0 : 1.0;
    {
      double sc = dir / pnorm;
      for (int i = 0; i < N; i++) p[i] *= sc;
    }
    pnorm = norm(p);
    fp0 = dot(p, g0);
    // changeDirection
    for (int i = 0; i < N; i++) { x_alpha[i] = x0[i]; g_alpha[i] = g0[i]; }
    x_cache_key = 0.0; f_cache_key = 0.0; g_cache_key = 0.0;
    df_alpha = slope(); df_cache_key = 0.0;
    return BFGS_SUCCESS;
  }

  LB_HD int test_gradient(double eps) const { return (norm(gradient) < eps) ? BFGS_SUCCESS : BFGS_RUNNING; }
};

// estimateRigidTransformationBFGS (gicp.hpp:234-287): x in/out.  returns 0 ok.
template <class Backend>
LB_HD int solve_bfgs(Backend& be, double* x, int max_inner, SolveStats& st) {
  Bfgs6<Backend> s;
  s.init(&be);
  const double gradient_tol = 1e-2;
  int inner = 0;
  int result = s.minimize_init(x);
  result = BFGS_RUNNING;
  do {
    inner++;
    result = s.minimize_one_step(x);
    if (result) break;
    result = s.test_gradient(gradient_tol);
  } while (result == BFGS_RUNNING && inner < max_inner);
  st.n_evals += s.n_evals;
  st.n_inner += inner;
  if (result == BFGS_NO_PROGRESS || result == BFGS_SUCCESS || inner == max_inner) return 0;
  return -1;
}

// Gauss-Newton inner solve (north_star; SURVEY App. A.5 -- not in the reference).
template <class Backend>
LB_HD int solve_gn(Backend& be, double* x, int max_inner, SolveStats& st) {
  for (int it = 0; it < max_inner; it++) {
    double f, b[6], H[21], d[6];
    if (be.gn(x, &f, b, H) != 0) return -1;
    st.n_evals++;
    st.n_inner++;
    if (solve6_neg(H, b, d) != 0) return -1;
    double mx = 0;
    for (int a = 0; a < 6; a++) { x[a] += d[a]; double v = fabs(d[a]); if (v > mx) mx = v; }
    if (mx < 1e-6) break; /* below the float32 resolution of T(x) */
  }
  return 0;
}

struct OuterParams {
  double rotation_epsilon, transformation_epsilon;
  int max_iterations, max_inner_iterations;
  int optimizer;  // 0 BFGS, 1 GN
};

struct OuterResult {
  float final_T[16];   // row-major 4x4 = previous * guess (gicp.hpp:583)
  float prev_T[16];    // previous_transformation_ (= transformation_ once converged): the guess-free increment
  int nr_iterations, converged, n_corr;
  double delta;
  SolveStats st;
};

// computeTransformation outer loop (gicp.hpp:445-583), split at the correspondence step so that the same code
// drives (a) the single-kernel / host loops below and (b) the stream-ordered execution, where the correspondence +
// moment kernel and the solve kernel of one outer iteration are separate launches (gicp_kernels.cuh).
struct OuterState {
  float T[12];      // transformation_ (reset to I by align())
  float prev[12];   // previous_transformation_
  int nr, converged, m_last, done;   // done: the while loop has been left (converged, or a caught exception)
  double delta;
  SolveStats st;
};

LB_HD void outer_init(OuterState& s) {
  const float I[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  for (int i = 0; i < 12; i++) { s.T[i] = I[i]; s.prev[i] = I[i]; }
  s.nr = 0; s.converged = 0; s.m_last = 0; s.done = 0;
  s.delta = 0.;
  s.st.n_evals = 0; s.st.n_inner = 0;
}

// R = rot(double(transformation_) * double(guess))   gicp.hpp:450-460
LB_HD void outer_rotation(const OuterState& s, const float* guess, double* R) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double acc = 0.0;
      for (int k = 0; k < 3; k++) acc += (double)s.T[i * 4 + k] * (double)guess[k * 4 + j];
      acc += (double)s.T[i * 4 + 3] * (double)guess[12 + j];
      R[i * 3 + j] = acc;
    }
}

// everything of one outer iteration after the correspondence step returned m pairs (gicp.hpp:509-569)
template <class Backend>
LB_HD void outer_step(OuterState& s, Backend& be, const OuterParams& P, int m) {
  s.m_last = m;
  for (int i = 0; i < 12; i++) s.prev[i] = s.T[i];
  if (m < 4) { s.done = 1; return; }  // NotEnoughPointsException -> caught -> break (gicp.hpp:225-233,542-547)
  double x[6];
  state_from_transform(s.T, x);
  int rc = (P.optimizer == 1) ? solve_gn(be, x, P.max_inner_iterations, s.st)
                              : solve_bfgs(be, x, P.max_inner_iterations, s.st);
  if (rc != 0) { s.done = 1; return; }  // SolverDidntConvergeException
  apply_state(x, s.T);
  double delta = 0.;
  for (int k = 0; k < 4; k++)
    for (int l = 0; l < 4; l++) {
      double ratio = (k < 3 && l < 3) ? 1. / P.rotation_epsilon : 1. / P.transformation_epsilon;
      double d = (k < 3) ? (double)(s.prev[k * 4 + l] - s.T[k * 4 + l]) : 0.0;
      double c_delta = ratio * fabs(d);
      if (c_delta > delta) delta = c_delta;
    }
  s.delta = delta;
  s.nr++;
  if (s.nr >= P.max_iterations || delta < 1) {
    s.converged = 1; s.done = 1;
    for (int i = 0; i < 12; i++) s.prev[i] = s.T[i];
  }
}

// final = previous * guess (float 4x4 product, Eigen coefficient order)   gicp.hpp:583
LB_HD void outer_finish(const OuterState& s, const float* guess, OuterResult& out) {
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      float a0 = (i < 3) ? s.prev[i * 4 + 0] : 0.f, a1 = (i < 3) ? s.prev[i * 4 + 1] : 0.f;
      float a2 = (i < 3) ? s.prev[i * 4 + 2] : 0.f, a3 = (i < 3) ? s.prev[i * 4 + 3] : 1.f;
      float v = a0 * guess[0 * 4 + j];
      v = v + a1 * guess[1 * 4 + j];
      v = v + a2 * guess[2 * 4 + j];
      v = v + a3 * guess[3 * 4 + j];
      out.final_T[i * 4 + j] = v;
    }
  for (int i = 0; i < 12; i++) out.prev_T[i] = s.prev[i];
  out.prev_T[12] = 0.f; out.prev_T[13] = 0.f; out.prev_T[14] = 0.f; out.prev_T[15] = 1.f;
  out.nr_iterations = s.nr;
  out.converged = s.converged;
  out.n_corr = s.m_last;
  out.delta = s.delta;
  out.st = s.st;
}

// guess: row-major 4x4 float.
template <class Backend>
LB_HD void gicp_outer_loop(Backend& be, const OuterParams& P, const float* guess, OuterResult& out) {
  OuterState s;
  outer_init(s);
  while (!s.done) {
    double R[9];
    outer_rotation(s, guess, R);
    int m = be.correspond(s.T, R);
    outer_step(s, be, P, m);
  }
  outer_finish(s, guess, out);
}

// The inner solve's objective in moment form (hd.h "moment form of the objective"): the Backend methods fdf / gn of
// one outer iteration evaluated from the 74 reduced moments, no pass over the points.
struct MomentObjective {
  double mom[MOM_N];
  float A0[12];      // the transform the moments were taken about (= the one the correspondences were searched with)
  int m;
  LB_HD void fdf(const double* x, double* f, double* g) const {
    Trig t;
    trig_compute(x, t);
    float A[12];
    apply_state_trig(x, t, A);
    double s13[13];
    moment_sums13(mom, A0, A, s13);
    objective_finish_trig(s13, m, t, f, g);
  }
  LB_HD int gn(const double* x, double* f, double* b, double* H) const {
    Trig t;
    trig_compute(x, t);
    float A[12];
    apply_state_trig(x, t, A);
    double dP[9], dT[9], dS[9], s28[28];
    r_derivatives_trig(t, dP, dT, dS);
    moment_gn28(mom, A0, A, dP, dT, dS, s28);
    *f = s28[0] / (double)m;
    for (int e = 0; e < 6; e++) b[e] = s28[1 + e];
    for (int e = 0; e < 21; e++) H[e] = s28[7 + e];
    return 0;
  }
};

}  // namespace lb
