/*
 * bfgs_oracle.c -- restatement of pcl::BFGS<Functor> (pcl/registration/bfgs.h,
 * PCL 1.10) for the CPU oracle.  TEST INFRASTRUCTURE ONLY.
 *
 * PCL is NOT in the reference tree (find_package(PCL 1.7) in
 * point_cloud_odometry/CMakeLists.txt:10; de-facto 1.10.0 on ROS noetic,
 * docker/Dockerfile:1).  pcl::BFGS is an Eigen port of GSL's
 * multimin/vector_bfgs2.c + linear_minimize.c (Fletcher's line search with
 * bracketing + sectioning and cubic/quadratic interpolation).  The call site
 * that fixes how it is driven is gicp.hpp:250-271:
 *   parameters sigma=0.01 rho=0.01 tau1=9 tau2=0.05 tau3=0.5 order=3,
 *   minimizeInit(x); do { minimizeOneStep(x); testGradient(1e-2) } ...
 * "parity unpinned": no reference test stores a BFGS trajectory.
 *
 * Documented PCL-specific details that are mirrored on purpose:
 *   - interpolate(): the quadratic branch tests `c > a` (PCL) where GSL has
 *     `c > 0`;
 *   - the derivative quadratic is solved in closed form
 *     (PCL's PolynomialSolver<Scalar,2> specialisation);
 *   - Eigen::poly_eval uses Horner for |x|<=1 and the reversed form otherwise.
 */
#include "lb_oracle.h"

#include <float.h>
#include <math.h>
#include <string.h>

static double dotn(const double* a, const double* b, int n) {
  double s = 0.0;
  for (int i = 0; i < n; i++) s += a[i] * b[i];
  return s;
}
static double normn(const double* a, int n) { return sqrt(dotn(a, a, n)); }

void og_bfgs_init_params(og_bfgs* b, og_functor fn) {
  memset(b, 0, sizeof(*b));
  b->fn = fn; b->n = fn.n;
  b->bracket_iters = 100; b->section_iters = 100;
  b->rho = 0.01; b->sigma = 0.01; b->tau1 = 9; b->tau2 = 0.05; b->tau3 = 0.5;
  b->step_size = 1; b->order = 3;
}

/* Eigen::poly_eval for a degree-3 polynomial c[0] + c[1] x + c[2] x^2 + c[3] x^3 */
static double poly_eval4(const double c[4], double x) {
  if (x * x <= 1.0) {
    double v = c[3];
    v = v * x + c[2];
    v = v * x + c[1];
    v = v * x + c[0];
    return v;
  } else {
    double v = c[0];
    double inv_x = 1.0 / x;
    for (int i = 1; i < 4; i++) v = v * inv_x + c[i];
    return pow(x, 3.0) * v;
  }
}

static void check_extremum(const double c[4], double x, double* xmin, double* fmin) {
  double y = poly_eval4(c, x);
  if (y < *fmin) { *xmin = x; *fmin = y; }
}

static void move_to(og_bfgs* b, double alpha) {
  for (int i = 0; i < b->n; i++) b->x_alpha[i] = b->x0[i] + alpha * b->p[i];
  b->x_cache_key = alpha;
}
static double slope(const og_bfgs* b) { return dotn(b->g_alpha, b->p, b->n); }

static double apply_f(og_bfgs* b, double alpha) {
  if (alpha == b->f_cache_key) return b->f_alpha;
  move_to(b, alpha);
  b->f_alpha = b->fn.f(b->fn.ctx, b->x_alpha); b->n_f++;
  b->f_cache_key = alpha;
  return b->f_alpha;
}
static double apply_df(og_bfgs* b, double alpha) {
  if (alpha == b->df_cache_key) return b->df_alpha;
  move_to(b, alpha);
  if (alpha != b->g_cache_key) {
    b->fn.df(b->fn.ctx, b->x_alpha, b->g_alpha); b->n_df++;
    b->g_cache_key = alpha;
  }
  b->df_alpha = slope(b);
  b->df_cache_key = alpha;
  return b->df_alpha;
}
static void apply_fdf(og_bfgs* b, double alpha, double* f, double* df) {
  if (alpha == b->f_cache_key && alpha == b->df_cache_key) {
    *f = b->f_alpha; *df = b->df_alpha; return;
  }
  if (alpha == b->f_cache_key || alpha == b->df_cache_key) {
    *f = apply_f(b, alpha);
    *df = apply_df(b, alpha);
    return;
  }
  move_to(b, alpha);
  b->fn.fdf(b->fn.ctx, b->x_alpha, &b->f_alpha, b->g_alpha); b->n_fdf++;
  b->f_cache_key = alpha;
  b->g_cache_key = alpha;
  b->df_alpha = slope(b);
  b->df_cache_key = alpha;
  *f = b->f_alpha; *df = b->df_alpha;
}
static void update_position(og_bfgs* b, double alpha, double* x, double* f, double* g) {
  double fa, dfa;
  apply_fdf(b, alpha, &fa, &dfa);
  *f = b->f_alpha;
  memcpy(x, b->x_alpha, sizeof(double) * (size_t)b->n);
  memcpy(g, b->g_alpha, sizeof(double) * (size_t)b->n);
}
static void change_direction(og_bfgs* b) {
  memcpy(b->x_alpha, b->x0, sizeof(double) * (size_t)b->n);
  b->x_cache_key = 0.0;
  b->f_cache_key = 0.0;
  memcpy(b->g_alpha, b->g0, sizeof(double) * (size_t)b->n);
  b->g_cache_key = 0.0;
  b->df_alpha = slope(b);
  b->df_cache_key = 0.0;
}

int og_bfgs_minimize_init(og_bfgs* b, double* x) {
  int n = b->n;
  b->delta_f = 0;
  memset(b->dx, 0, sizeof(b->dx));
  b->fn.fdf(b->fn.ctx, x, &b->f, b->gradient); b->n_fdf++;
  memcpy(b->x0, x, sizeof(double) * (size_t)n);
  memcpy(b->g0, b->gradient, sizeof(double) * (size_t)n);
  b->g0norm = normn(b->g0, n);
  for (int i = 0; i < n; i++) b->p[i] = b->gradient[i] * -1 / b->g0norm;
  b->pnorm = normn(b->p, n);
  b->fp0 = -b->g0norm;
  memcpy(b->x_alpha, b->x0, sizeof(double) * (size_t)n); b->x_cache_key = 0;
  b->f_alpha = b->f; b->f_cache_key = 0;
  memcpy(b->g_alpha, b->g0, sizeof(double) * (size_t)n); b->g_cache_key = 0;
  b->df_alpha = slope(b); b->df_cache_key = 0;
  return OG_BFGS_NOT_STARTED;
}

static double interpolate(double a, double fa, double fpa, double b, double fb, double fpb,
                          double xmin, double xmax, int order) {
  double y, alpha, ymin, ymax, fmin;
  ymin = (xmin - a) / (b - a);
  ymax = (xmax - a) / (b - a);
  if (ymin > ymax) { double tmp = ymin; ymin = ymax; ymax = tmp; }

  if (order > 2 && !(fpb != fpb) && fpb != INFINITY) {
    fpa = fpa * (b - a);
    fpb = fpb * (b - a);
    double eta = 3 * (fb - fa) - 2 * fpa - fpb;
    double xi = fpa + fpb - 2 * (fb - fa);
    double c[4] = {fa, fpa, eta, xi};
    double y0, y1;
    y = ymin;
    fmin = poly_eval4(c, ymin);
    check_extremum(c, ymax, &y, &fmin);
    {
      /* derivative c1 + 2 c2 y + 3 c3 y^2: PCL closed-form quadratic solver */
      double p0 = c[1], p1 = 2 * c[2], p2 = 3 * c[3];
      double a2 = 2 * p2;
      double disc = (p1 * p1) - (4 * p0 * p2);
      if (0.0 < disc) {
        double dr = sqrt(disc);
        y0 = (-p1 - dr) / a2;
        y1 = (-p1 + dr) / a2;
        if (y0 > y1) { double tmp = y0; y0 = y1; y1 = tmp; }
        if (y0 > ymin && y0 < ymax) check_extremum(c, y0, &y, &fmin);
        if (y1 > ymin && y1 < ymax) check_extremum(c, y1, &y, &fmin);
      } else if (0.0 == disc) {
        y0 = -p1 / a2;
        if (y0 > ymin && y0 < ymax) check_extremum(c, y0, &y, &fmin);
      }
      /* complex roots: hasRealRoot == false, nothing to check */
    }
  } else {
    fpa = fpa * (b - a);
    double fl = fa + ymin * (fpa + ymin * (fb - fa - fpa));
    double fh = fa + ymax * (fpa + ymax * (fb - fa - fpa));
    double c = 2 * (fb - fa - fpa); /* curvature */
    y = ymin; fmin = fl;
    if (fh < fmin) { y = ymax; fmin = fh; }
    if (c > a) { /* sic: PCL compares with a (GSL: c > 0) */
      double z = -fpa / c;
      if (z > ymin && z < ymax) {
        double f = fa + z * (fpa + z * (fb - fa - fpa));
        if (f < fmin) { y = z; fmin = f; }
      }
    }
  }
  alpha = a + y * (b - a);
  return alpha;
}

static int line_search(og_bfgs* s, double rho, double sigma, double tau1, double tau2,
                       double tau3, int order, double alpha1, double* alpha_new) {
  double f0, fp0, falpha, falpha_prev, fpalpha, fpalpha_prev, delta, alpha_next;
  double alpha = alpha1, alpha_prev = 0.0;
  double a, b, fa, fb, fpa, fpb;
  int i = 0;

  apply_fdf(s, 0.0, &f0, &fp0);
  falpha_prev = f0;
  fpalpha_prev = fp0;
  a = 0.0; b = alpha;
  fa = f0; fb = 0.0;
  fpa = fp0; fpb = 0.0;

  /* bracketing */
  while (i++ < s->bracket_iters) {
    falpha = apply_f(s, alpha);
    if (falpha > f0 + alpha * rho * fp0 || falpha >= falpha_prev) {
      a = alpha_prev; fa = falpha_prev; fpa = fpalpha_prev;
      b = alpha; fb = falpha; fpb = NAN;
      break;
    }
    fpalpha = apply_df(s, alpha);
    if (fabs(fpalpha) <= -sigma * fp0) { *alpha_new = alpha; return OG_BFGS_SUCCESS; }
    if (fpalpha >= 0) {
      a = alpha; fa = falpha; fpa = fpalpha;
      b = alpha_prev; fb = falpha_prev; fpb = fpalpha_prev;
      break;
    }
    delta = alpha - alpha_prev;
    {
      double lower = alpha + delta;
      double upper = alpha + tau1 * delta;
      alpha_next = interpolate(alpha_prev, falpha_prev, fpalpha_prev, alpha, falpha, fpalpha,
                               lower, upper, order);
    }
    alpha_prev = alpha;
    falpha_prev = falpha;
    fpalpha_prev = fpalpha;
    alpha = alpha_next;
  }
  /* sectioning of bracket [a,b] */
  while (i++ < s->section_iters) {
    delta = b - a;
    {
      double lower = a + tau2 * delta;
      double upper = b - tau3 * delta;
      alpha = interpolate(a, fa, fpa, b, fb, fpb, lower, upper, order);
    }
    falpha = apply_f(s, alpha);
    if ((a - alpha) * fpa <= DBL_EPSILON) return OG_BFGS_NO_PROGRESS; /* roundoff prevents progress */
    if (falpha > f0 + rho * alpha * fp0 || falpha >= fa) {
      b = alpha; fb = falpha; fpb = NAN;
    } else {
      fpalpha = apply_df(s, alpha);
      if (fabs(fpalpha) <= -sigma * fp0) { *alpha_new = alpha; return OG_BFGS_SUCCESS; }
      if (((b - a) >= 0 && fpalpha >= 0) || ((b - a) <= 0 && fpalpha <= 0)) {
        b = a; fb = fa; fpb = fpa;
        a = alpha; fa = falpha; fpa = fpalpha;
      } else {
        a = alpha; fa = falpha; fpa = fpalpha;
      }
    }
  }
  return OG_BFGS_SUCCESS;
}

int og_bfgs_minimize_one_step(og_bfgs* b, double* x) {
  int n = b->n;
  double alpha = 0.0, alpha1;
  double f0 = b->f;
  if (b->pnorm == 0.0 || b->g0norm == 0.0 || b->fp0 == 0) {
    memset(b->dx, 0, sizeof(b->dx));
    return OG_BFGS_NO_PROGRESS;
  }
  if (b->delta_f < 0) {
    double del = fmax(-b->delta_f, 10 * DBL_EPSILON * fabs(f0));
    alpha1 = fmin(1.0, 2.0 * del / (-b->fp0));
  } else {
    alpha1 = fabs(b->step_size);
  }
  int status = line_search(b, b->rho, b->sigma, b->tau1, b->tau2, b->tau3, b->order, alpha1, &alpha);
  if (status != OG_BFGS_SUCCESS) return status;

  update_position(b, alpha, x, &b->f, b->gradient);
  b->delta_f = b->f - f0;

  {
    double dxg, dgg, dxdg, dgnorm, A, B;
    for (int i = 0; i < n; i++) { b->dx0[i] = x[i] - b->x0[i]; b->dx[i] = b->dx0[i]; }
    for (int i = 0; i < n; i++) b->dg0[i] = b->gradient[i] - b->g0[i];
    dxg = dotn(b->dx0, b->gradient, n);
    dgg = dotn(b->dg0, b->gradient, n);
    dxdg = dotn(b->dx0, b->dg0, n);
    dgnorm = normn(b->dg0, n);
    if (dxdg != 0) {
      B = dxg / dxdg;
      A = -(1.0 + dgnorm * dgnorm / dxdg) * B + dgg / dxdg;
    } else {
      B = 0; A = 0;
    }
    for (int i = 0; i < n; i++) b->p[i] = -A * b->dx0[i];
    for (int i = 0; i < n; i++) b->p[i] += b->gradient[i];
    for (int i = 0; i < n; i++) b->p[i] += -B * b->dg0[i];
  }
  memcpy(b->g0, b->gradient, sizeof(double) * (size_t)n);
  memcpy(b->x0, x, sizeof(double) * (size_t)n);
  b->g0norm = normn(b->g0, n);
  b->pnorm = normn(b->p, n);

  double dir = (dotn(b->p, b->gradient, n) > 0) ? -1.0 : 1.0;
  {
    double sc = dir / b->pnorm;
    for (int i = 0; i < n; i++) b->p[i] *= sc;
  }
  b->pnorm = normn(b->p, n);
  b->fp0 = dotn(b->p, b->g0, n);
  change_direction(b);
  return OG_BFGS_SUCCESS;
}

int og_bfgs_test_gradient(const og_bfgs* b, double eps) {
  if (eps < 0) return OG_BFGS_NEG_GRAD_EPS;
  return (normn(b->gradient, b->n) < eps) ? OG_BFGS_SUCCESS : OG_BFGS_RUNNING;
}

/* ---- self-test helper: quadratic f = 0.5 x'Ax - b'x -------------------- */
typedef struct { const double* A; const double* b; int n; } quad_ctx;
static void quad_fdf(void* c, const double* x, double* f, double* g) {
  quad_ctx* q = (quad_ctx*)c;
  double ff = 0;
  for (int i = 0; i < q->n; i++) {
    double ax = 0;
    for (int j = 0; j < q->n; j++) ax += q->A[i * q->n + j] * x[j];
    g[i] = ax - q->b[i];
    ff += 0.5 * x[i] * ax - q->b[i] * x[i];
  }
  *f = ff;
}
static double quad_f(void* c, const double* x) { double f, g[OG_BFGS_MAXN]; quad_fdf(c, x, &f, g); return f; }
static void quad_df(void* c, const double* x, double* g) { double f; quad_fdf(c, x, &f, g); }

int og_bfgs_minimize_quadratic(const double* A, const double* bvec, int n, double* x,
                               int max_iters, double grad_tol) {
  quad_ctx q = {A, bvec, n};
  og_functor fn = {quad_f, quad_df, quad_fdf, &q, n};
  og_bfgs s;
  og_bfgs_init_params(&s, fn);
  og_bfgs_minimize_init(&s, x);
  int it = 0, result;
  do {
    it++;
    result = og_bfgs_minimize_one_step(&s, x);
    if (result) break;
    result = og_bfgs_test_gradient(&s, grad_tol);
  } while (result == OG_BFGS_RUNNING && it < max_iters);
  return it;
}
