// dmvae_mpc.cu - batched MPC path tracker.
//
// The reference turns every generated waypoint set into a drivable trajectory with a model-predictive controller
// (MPC/MPC_Tracking.py, driven by Distribution.py:91-105): per 15-25 ms time step one SLSQP solve over a 20 x 2
// control sequence with a 30-step bicycle-model rollout, finite-difference gradients, about a second of host time per
// step and 200-600 steps per trajectory - the real wall-clock owner of results/GeneratedData.  Trajectories are
// independent, so here ONE THREAD tracks ONE trajectory from its first step to its last, all in float64 as the
// reference:
//
//   mpc_prepare_kernel   PathInterpolator._create_interpolators (:103-221) reduced to what the controller reads: the
//                        not-a-knot cubic interpolants of the velocity knots (scipy interp1d(kind='cubic'), evaluated
//                        as piecewise cubics and extrapolated with their end pieces), the start heading, the 1 ms scan
//                        that decides where the end velocity is taken, the end heading; PathTracker.__init__
//                        (:435-441) for the initial state.
//   mpc_track_kernel     PathTracker.step (:454-493) in a loop: the reference window (:465-478 with get_reference
//                        :224-252 and get_reference_heading :254-277), the optimal control problem of solve_mpc
//                        (:311-415), the Euler step of the bicycle model with the first control (:484-486).
//
// The optimiser is NOT SLSQP.  The cost of solve_mpc reads only heading and speed, whose dynamics
//   theta' = theta + dt / L * v * tan(delta),  v' = v + dt * a
// make it a small optimal control problem in the state (theta, v, previous a, previous delta): it is solved by
// differential dynamic programming with exact second derivatives (a Newton method on the control sequence, one 2 x 2
// box-constrained quadratic problem per stage, Tassa et al.'s control-limited DDP), warm-started from the previous
// step's shifted solution and iterated until no control moves by more than cfg.tol (1e-6; the method converges
// quadratically, so the controls are then good to ~1e-11).  SLSQP in the reference stops at ftol = 1e-6,
// i.e. within ~1e-3 of the same minimiser in the controls; the parity tests hold this kernel to 1e-7 against that
// minimiser (oracle/mpc_oracle.py solve_exact) and to the early-stopping noise against the reference's own runs.
//
// One quirk of the reference is kept because it moves the minimiser: its bounds list is [accel] * 20 + [steer] * 20
// against a variable vector that interleaves (a, delta) per row (:390-398), so rows 10..19 have |a| <= 0.5 (the
// steering bound) - braking harder than 0.5 m/s^2 is only available in the first half of the control horizon.
#include <math.h>

#include "dmvae_launch.h"

namespace dmvae {

// per-trajectory persistent state in the workspace, structure of arrays over the n trajectories (field f of
// trajectory j at ws[f * n + j]): knots, the two sets of piecewise-cubic coefficients, scalars, previous solution
__host__ __device__ inline int mpc_f_knots(const MpcCfg&) { return 0; }
__host__ __device__ inline int mpc_f_cx(const MpcCfg& c) { return c.n_way; }
__host__ __device__ inline int mpc_f_cy(const MpcCfg& c) { return c.n_way + 4 * (c.n_way - 1); }
__host__ __device__ inline int mpc_f_scal(const MpcCfg& c) { return c.n_way + 8 * (c.n_way - 1); }
enum MpcScalar { MS_T_END = 0, MS_START_THETA, MS_END_VX, MS_END_VY, MS_END_THETA, MS_LAST_A, MS_LAST_D, MS_HAVE_LAST, MS_COUNT };
__host__ __device__ inline int mpc_f_warm(const MpcCfg& c) { return mpc_f_scal(c) + MS_COUNT; }
__host__ __device__ inline int mpc_fields(const MpcCfg& c) { return mpc_f_warm(c) + 2 * c.blocks; }

size_t mpc_workspace_bytes(const MpcCfg& c, long long n) { return (size_t)mpc_fields(c) * (size_t)n * sizeof(double); }

constexpr double MPC_WRAP = -2.8;   // MPC_Tracking.py:202: headings below -2.8 rad move up by 2 pi
__device__ __forceinline__ double wrap_heading(double th) { return th >= MPC_WRAP ? th : th + 2.0 * M_PI; }

struct Knots {
  const double* ws;
  long long n, j;
  int nw, f_cx, f_cy;
  __device__ __forceinline__ double knot(int i) const { return ws[(size_t)i * n + j]; }
  // both interpolants at time t; cur: interval cursor (times mostly increase from call to call)
  __device__ void eval(double t, int& cur, double& vx, double& vy) const {
    while (cur < nw - 2 && t >= knot(cur + 1)) ++cur;
    while (cur > 0 && t < knot(cur)) --cur;
    const double s = t - knot(cur);
    const double* cx = ws + (size_t)(f_cx + 4 * cur) * n + j;
    const double* cy = ws + (size_t)(f_cy + 4 * cur) * n + j;
    vx = cx[0] + s * (cx[n] + s * (cx[2 * n] + s * cx[3 * n]));
    vy = cy[0] + s * (cy[n] + s * (cy[2 * n] + s * cy[3 * n]));
  }
};

// Not-a-knot cubic interpolant of (t_i, y_i), i < nw (nw >= 4; a parabola / a line for 3 / 2 points) as one cubic per interval:
// c0 + c1 s + c2 s^2 + c3 s^3, s = t - t_i, written to out[(4 i + k) * n].  m: scratch of MPC_MAX_WAY doubles x 3.
__device__ void notaknot(const double* t, const double* y, int nw, double* out, long long n, double* m, double* cp, double* dp) {
  auto h = [&](int i) { return t[i + 1] - t[i]; };
  auto d = [&](int i) { return (y[i + 1] - y[i]) / h(i); };
  // fewer than four points: the reference falls back to interp1d(kind='quadratic') for three (one parabola through
  // them) and kind='linear' for two (MPC_Tracking.py:126-137, :173-178)
  if (nw == 2) {
    out[0] = y[0];
    out[(size_t)n] = d(0);
    out[(size_t)2 * n] = 0.0;
    out[(size_t)3 * n] = 0.0;
    return;
  }
  if (nw == 3) {
    const double dd = (d(1) - d(0)) / (t[2] - t[0]);   // second divided difference
    out[0] = y[0];
    out[(size_t)n] = d(0) - dd * h(0);
    out[(size_t)2 * n] = dd;
    out[(size_t)3 * n] = 0.0;
    out[(size_t)4 * n] = y[1];
    out[(size_t)5 * n] = d(0) + dd * h(0);
    out[(size_t)6 * n] = dd;
    out[(size_t)7 * n] = 0.0;
    return;
  }
  // second derivatives m_1 .. m_{nw-2} from the tridiagonal system with m_0 and m_{nw-1} eliminated through the
  // not-a-knot conditions (third derivative continuous across t_1 and t_{nw-2})
  const int N = nw - 2;   // unknowns
  for (int k = 0; k < N; ++k) {   // row of m_{k+1}
    const int i = k + 1;
    double lo = h(i - 1), di = 2.0 * (h(i - 1) + h(i)), up = h(i);
    const double rhs = 6.0 * (d(i) - d(i - 1));
    if (i == 1) {          // m_0 = (1 + h0 / h1) m_1 - (h0 / h1) m_2
      const double r = h(0) / h(1);
      di += h(0) * (1.0 + r);
      up -= h(0) * r;
      lo = 0.0;
    }
    if (i == nw - 2) {     // m_{nw-1} = (1 + h_{nw-2} / h_{nw-3}) m_{nw-2} - (h_{nw-2} / h_{nw-3}) m_{nw-3}
      const double r = h(nw - 2) / h(nw - 3);
      di += h(nw - 2) * (1.0 + r);
      lo -= h(nw - 2) * r;
      up = 0.0;
    }
    if (k == 0) {
      cp[0] = up / di;
      dp[0] = rhs / di;
    } else {
      const double den = di - lo * cp[k - 1];
      cp[k] = up / den;
      dp[k] = (rhs - lo * dp[k - 1]) / den;
    }
  }
  m[N] = dp[N - 1];
  for (int k = N - 2; k >= 0; --k) m[k + 1] = dp[k] - cp[k] * m[k + 2];
  {
    const double r0 = h(0) / h(1), r1 = h(nw - 2) / h(nw - 3);
    m[0] = (1.0 + r0) * m[1] - r0 * m[2];
    m[nw - 1] = (1.0 + r1) * m[nw - 2] - r1 * m[nw - 3];
  }
  for (int i = 0; i < nw - 1; ++i) {
    const double hi = h(i);
    out[(size_t)(4 * i + 0) * n] = y[i];
    out[(size_t)(4 * i + 1) * n] = d(i) - hi * (2.0 * m[i] + m[i + 1]) / 6.0;
    out[(size_t)(4 * i + 2) * n] = 0.5 * m[i];
    out[(size_t)(4 * i + 3) * n] = (m[i + 1] - m[i]) / (6.0 * hi);
  }
}

// one thread per trajectory
__global__ void __launch_bounds__(128) mpc_prepare_kernel(const MpcCfg c, const void* __restrict__ way, const double* __restrict__ init,
                                                          double* __restrict__ ws, long long n, double* __restrict__ state,
                                                          int* __restrict__ status, double* __restrict__ profile) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int nw = c.n_way;
  double t[MPC_MAX_WAY], vx[MPC_MAX_WAY], vy[MPC_MAX_WAY], m[MPC_MAX_WAY], cp[MPC_MAX_WAY], dp[MPC_MAX_WAY];
  double t_last, t_mid, scan_stop, t_first;
  bool ok = true;
  // velocity knots (:146-168).  float32 waypoints: the time differences, their halves and the knot times are float32
  // values (NumPy keeps a float32 array float32 when it meets a Python float), the position differences float64.
  const double vx0 = init[j * 5 + 3], vy0 = init[j * 5 + 4];
  if (c.way_f32) {
    const float* w = static_cast<const float*>(way) + (size_t)j * nw * 3;
    for (int i = 0; i + 1 < nw; ++i) {
      float dtf = __fsub_rn(w[(i + 1) * 3 + 2], w[i * 3 + 2]);
      ok = ok && dtf > 0.f;
      if (dtf == 0.f) dtf = 1e-6f;
      t[i + 1] = (double)__fadd_rn(w[i * 3 + 2], __fmul_rn(dtf, 0.5f));
      if (nw == 2) {
        // two waypoints: the reference's position interpolant is interp1d(kind='linear'), which stays in the waypoints'
        // float32 (the spline kinds evaluate in float64): position difference and quotient are float32 operations
        vx[i + 1] = (double)__fdiv_rn(__fsub_rn(w[(i + 1) * 3], w[i * 3]), dtf);
        vy[i + 1] = (double)__fdiv_rn(__fsub_rn(w[(i + 1) * 3 + 1], w[i * 3 + 1]), dtf);
      } else {
        vx[i + 1] = ((double)w[(i + 1) * 3] - (double)w[i * 3]) / (double)dtf;
        vy[i + 1] = ((double)w[(i + 1) * 3 + 1] - (double)w[i * 3 + 1]) / (double)dtf;
      }
    }
    t_first = (double)w[2];
    t_last = (double)w[(nw - 1) * 3 + 2];
    t_mid = (double)__fmul_rn(__fadd_rn(w[(nw - 1) * 3 + 2], w[(nw - 2) * 3 + 2]), 0.5f);
    scan_stop = (double)__fadd_rn(w[(nw - 1) * 3 + 2], 0.001f);
  } else {
    const double* w = static_cast<const double*>(way) + (size_t)j * nw * 3;
    for (int i = 0; i + 1 < nw; ++i) {
      double dtd = w[(i + 1) * 3 + 2] - w[i * 3 + 2];
      ok = ok && dtd > 0.0;
      if (dtd == 0.0) dtd = 1e-6;
      t[i + 1] = w[i * 3 + 2] + dtd / 2;
      vx[i + 1] = (w[(i + 1) * 3] - w[i * 3]) / dtd;
      vy[i + 1] = (w[(i + 1) * 3 + 1] - w[i * 3 + 1]) / dtd;
    }
    t_first = w[2];
    t_last = w[(nw - 1) * 3 + 2];
    t_mid = (w[(nw - 1) * 3 + 2] + w[(nw - 2) * 3 + 2]) / 2;
    scan_stop = w[(nw - 1) * 3 + 2] + 0.001;
  }
  t[0] = 0.0;
  vx[0] = vx0;
  vy[0] = vy0;
  ok = ok && t[1] > 0.0;   // the knot times themselves must increase for the interpolant
  ok = ok && t_last < 1e6 && t_first > -1e6;   // finite (an infinite last time would make the 1 ms heading scan endless)
  status[j] = ok ? 0 : 1;  // 1: waypoint times do not increase strictly (the reference raises, :118-119)
  // initial state (:435-441): heading wrapped, speed = |(vx, vy)|
  {
    double th = init[j * 5 + 2];
    if (th < MPC_WRAP) th += 2.0 * M_PI;
    state[j * 4 + 0] = init[j * 5 + 0];
    state[j * 4 + 1] = init[j * 5 + 1];
    state[j * 4 + 2] = th;
    state[j * 4 + 3] = sqrt(vx0 * vx0 + vy0 * vy0);
  }
  double* scal = ws + (size_t)mpc_f_scal(c) * n + j;
  scal[(size_t)MS_LAST_A * n] = 0.0;
  scal[(size_t)MS_LAST_D * n] = 0.0;
  scal[(size_t)MS_HAVE_LAST * n] = 0.0;
  for (int k = 0; k < 2 * c.blocks; ++k) ws[(size_t)(mpc_f_warm(c) + k) * n + j] = 0.0;
  if (!ok) return;
  for (int i = 0; i < nw; ++i) ws[(size_t)i * n + j] = t[i];
  notaknot(t, vx, nw, ws + (size_t)mpc_f_cx(c) * n + j, n, m, cp, dp);
  notaknot(t, vy, nw, ws + (size_t)mpc_f_cy(c) * n + j, n, m, cp, dp);
  const Knots K{ws, n, j, nw, mpc_f_cx(c), mpc_f_cy(c)};
  int cur = 0;
  double ax, ay;
  K.eval(t_first, cur, ax, ay);
  const double start_theta = wrap_heading(atan2(ay, ax));   // :198-202
  // :204-218: does the heading leave a 45 degree cone around the start heading anywhere on the 1 ms grid
  // np.arange(0, t_last + 0.001, 0.001)?  Only whether, not where: the end velocity does not depend on the time found.
  const long long grid = (long long)ceil(scan_stop / 0.001);
  bool turned = false;
  cur = 0;
  for (long long i = 0; i < grid; ++i) {
    K.eval((double)i * 0.001, cur, ax, ay);
    const double th = wrap_heading(atan2(ay, ax));
    if (fabs(th - start_theta) > 45 * M_PI / 180) {
      turned = true;
      break;
    }
  }
  double evx, evy;
  K.eval(turned ? t_mid : t_last, cur, evx, evy);
  const double end_theta = wrap_heading(atan2(evy, evx));
  scal[(size_t)MS_T_END * n] = t_last;
  scal[(size_t)MS_START_THETA * n] = start_theta;
  scal[(size_t)MS_END_VX * n] = evx;
  scal[(size_t)MS_END_VY * n] = evy;
  scal[(size_t)MS_END_THETA * n] = end_theta;
  if (profile != nullptr) {
    profile[j * 5 + 0] = start_theta;
    profile[j * 5 + 1] = evx;
    profile[j * 5 + 2] = evy;
    profile[j * 5 + 3] = end_theta;
    profile[j * 5 + 4] = t_last;
  }
}

// reference window of one controller call (:465-478)
__device__ void mpc_window(const MpcCfg& c, const Knots& K, const double t_end, const double start_theta, const double evx,
                           const double evy, const double end_theta, const double current_time, const double dt, int& cur,
                           double* thr, double* vr) {
  double held = 0.0;
  for (int i = 0; i <= c.horizon; ++i) {
    const double t = current_time + i * dt;
    double vx, vy;
    // heading of (vx, vy).  Where the end velocity stands in, get_reference_heading recomputes atan2(evy, evx) and wraps
    // it: that is end_theta, which the wrap leaves alone
    double heading = end_theta;
    if (t <= t_end) {   // get_reference (:235-243)
      K.eval(t, cur, vx, vy);
      heading = atan2(vy, vx);
      if (fabs(heading - start_theta) > 90 * M_PI / 180) {
        vx = evx;
        vy = evy;
        heading = end_theta;
      }
    } else {            // beyond the last waypoint: straight on with the end velocity (:246-250)
      vx = evx;
      vy = evy;
    }
    const double v = sqrt(vx * vx + vy * vy);
    if (v >= 0.1) held = wrap_heading(heading);   // :264-273 (beyond the end: the wrapped end heading, wrapped again)
    thr[i] = held;
    vr[i] = v;
  }
}

// minimise 0.5 d^T Q d + g^T d over lo <= d <= hi (2 variables, Q positive definite); fr[i]: d_i ended strictly inside
__device__ void box_qp2(double q00, double q01, double q11, double g0, double g1, const double lo[2], const double hi[2], double d[2],
                        bool fr[2]) {
  const double det = q00 * q11 - q01 * q01;
  const double u0 = -(q11 * g0 - q01 * g1) / det, u1 = -(q00 * g1 - q01 * g0) / det;
  if (u0 >= lo[0] && u0 <= hi[0] && u1 >= lo[1] && u1 <= hi[1]) {
    d[0] = u0; d[1] = u1;
    fr[0] = fr[1] = true;
    return;
  }
  // the minimum lies on the boundary: the best of the four edges, each a clipped one-dimensional minimum
  double best = INFINITY;
  auto value = [&](double a, double b) { return 0.5 * (q00 * a * a + 2.0 * q01 * a * b + q11 * b * b) + g0 * a + g1 * b; };
  auto clip = [](double x, double l, double h) { return fmin(fmax(x, l), h); };
  for (int e = 0; e < 4; ++e) {
    double a, b;
    if (e < 2) {
      a = e == 0 ? lo[0] : hi[0];
      b = clip(-(g1 + q01 * a) / q11, lo[1], hi[1]);
    } else {
      b = e == 2 ? lo[1] : hi[1];
      a = clip(-(g0 + q01 * b) / q00, lo[0], hi[0]);
    }
    const double f = value(a, b);
    if (f < best) {
      best = f;
      d[0] = a; d[1] = b;
    }
  }
  fr[0] = d[0] > lo[0] && d[0] < hi[0];
  fr[1] = d[1] > lo[1] && d[1] < hi[1];
}

// Per-thread working set of the solver (local memory).  It is what bounds the kernel: the first build kept 7 KB per
// thread (rollouts, gains and controls for the largest horizons in float64), 460 MB for the resident threads of a B200,
// and streamed it through HBM (110 GB of DRAM traffic for 20 steps of 65 536 trajectories).  Now: sized by the
// instantiation (HOR, BLK), no stored rollouts (the forward sweep recomputes the nominal states with the arithmetic
// that produced them, the backward sweep walks them back), feedback gains in float32 (they shape the Newton step, not
// the fixed point it converges to): 2.4 KB per thread at (30, 20).
template <int HOR, int BLK>
struct MpcLocal {
  double thr[HOR + 1], vr[HOR + 1];
  double a[BLK], dl[BLK], tau[BLK], an[BLK], dn[BLK], taun[BLK];
  double kff[BLK][2];
  float kfb[BLK][2][4];
};

__device__ __forceinline__ double mpc_alim(const MpcCfg& c, int k) { return 2 * k < c.blocks ? c.max_accel : c.max_steer; }

// One controller call: minimise the objective of solve_mpc (:329-373) over the control rows S.a / S.dl (in: warm
// start, out: solution).  Returns the iterations used.
template <int HOR, int BLK>
__device__ int mpc_solve(const MpcCfg& c, const double dt, const double th0, const double v0, const bool have_last, const double la,
                         const double ld, MpcLocal<HOR, BLK>& S) {
  const int N = c.horizon, M = c.blocks;
  const double cdt = dt / c.wheelbase;
  const double qth2 = 2.0 * c.q_theta, qv2 = 2.0 * c.q_v;
  // nominal rollout of the warm start: its cost and its last state (thN, vN); tau = tan of every row's steering
  double cost = 0.0, thN = th0, vN = v0;
  for (int i = 0; i < N; ++i) {
    const int k = i < M ? i : M - 1;
    if (i < M) {
      S.tau[i] = tan(S.dl[i]);
      if (i > 0 || have_last) {
        const double da = S.a[i] - (i == 0 ? la : S.a[i - 1]), dd = S.dl[i] - (i == 0 ? ld : S.dl[i - 1]);
        cost += c.r_accel * da * da + c.r_steer * dd * dd;
      }
    }
    const double eth = thN - S.thr[i], ev = vN - S.vr[i];
    cost += c.q_theta * eth * eth + c.q_v * ev * ev;
    thN = thN + cdt * vN * S.tau[k];
    vN = vN + dt * S.a[k];
  }
  {
    const double eth = thN - S.thr[N], ev = vN - S.vr[N];
    cost += c.q_theta * eth * eth + c.q_v * ev * ev;
  }
  int it = 0;
  for (; it < c.max_iter; ++it) {
    // ---- backward sweep: gradient g and Hessian H of the cost-to-go in x = (theta, v, previous a, previous delta).
    // The nominal states are walked back from the last one (v_i = v_{i+1} - dt a, theta_i = theta_{i+1} - cdt v_i tau):
    // equal to the forward values to rounding, which moves the fixed point by ~1e-15.
    double th = thN, v = vN;
    double maxd = 0.0;   // largest feed-forward change of this sweep
    double g[4] = {qth2 * (th - S.thr[N]), qv2 * (v - S.vr[N]), 0.0, 0.0};
    double H[4][4] = {{qth2, 0, 0, 0}, {0, qv2, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
    for (int i = N - 1; i >= M; --i) {   // rows beyond the control horizon: the last control row is held (:337-339)
      const double tau = S.tau[M - 1], sig = 1.0 + tau * tau;
      v = v - dt * S.a[M - 1];
      th = th - cdt * v * tau;
      const double a01 = cdt * tau, a03 = cdt * v * sig, a12 = dt;
      // A = [[1, a01, 0, a03], [0, 1, a12, 0], [0, 0, 1, 0], [0, 0, 0, 1]];  HA = H A,  H' = A^T HA
      double HA[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        HA[r][0] = H[r][0];
        HA[r][1] = a01 * H[r][0] + H[r][1];
        HA[r][2] = a12 * H[r][1] + H[r][2];
        HA[r][3] = a03 * H[r][0] + H[r][3];
      }
      const double g0 = g[0];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        H[0][q] = HA[0][q];
        H[1][q] = a01 * HA[0][q] + HA[1][q];
        H[2][q] = a12 * HA[1][q] + HA[2][q];
        H[3][q] = a03 * HA[0][q] + HA[3][q];
      }
      // second derivatives of theta' = theta + cdt v tan(q): d2/dv dq = cdt sig, d2/dq2 = 2 cdt v sig tau
      H[1][3] += g0 * cdt * sig;
      H[3][1] += g0 * cdt * sig;
      H[3][3] += g0 * 2.0 * cdt * v * sig * tau;
      H[0][0] += qth2;
      H[1][1] += qv2;
      const double n0 = g[0], n1 = a01 * g[0] + g[1], n2 = a12 * g[1] + g[2], n3 = a03 * g[0] + g[3];
      g[0] = n0 + qth2 * (th - S.thr[i]);
      g[1] = n1 + qv2 * (v - S.vr[i]);
      g[2] = n2;
      g[3] = n3;
    }
    for (int i = M - 1; i >= 0; --i) {
      const double w = (i == 0 && !have_last) ? 0.0 : 1.0;
      const double ra2 = 2.0 * w * c.r_accel, rs2 = 2.0 * w * c.r_steer;
      const double pa = i == 0 ? la : S.a[i - 1], pd = i == 0 ? ld : S.dl[i - 1];
      const double tau = S.tau[i], sig = 1.0 + tau * tau;
      v = v - dt * S.a[i];
      th = th - cdt * v * tau;
      const double ct = cdt * tau, b = cdt * v * sig;
      const double lth = qth2 * (th - S.thr[i]), lv = qv2 * (v - S.vr[i]);
      const double lua = ra2 * (S.a[i] - pa), lud = rs2 * (S.dl[i] - pd);
      // columns of d x' / d u: a -> (0, dt, 1, 0), delta -> (b, 0, 0, 1)
      double Ha[4], Hd[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        Ha[r] = dt * H[r][1] + H[r][2];
        Hd[r] = b * H[r][0] + H[r][3];
      }
      const double Qu[2] = {lua + dt * g[1] + g[2], lud + b * g[0] + g[3]};
      double q00 = ra2 + dt * Ha[1] + Ha[2];
      const double q01 = b * Ha[0] + Ha[3];
      double q11 = rs2 + b * Hd[0] + Hd[3] + g[0] * 2.0 * cdt * v * sig * tau;
      double Qux[2][4] = {{Ha[0], ct * Ha[0] + Ha[1], -ra2, 0.0}, {Hd[0], ct * Hd[0] + Hd[1] + g[0] * cdt * sig, 0.0, -rs2}};
      const double Qx[4] = {lth + g[0], lv + ct * g[0] + g[1], -lua, -lud};
      double Qxx[4][4] = {{qth2 + H[0][0], ct * H[0][0] + H[0][1], 0, 0},
                          {ct * H[0][0] + H[0][1], qv2 + ct * ct * H[0][0] + 2.0 * ct * H[0][1] + H[1][1], 0, 0},
                          {0, 0, ra2, 0},
                          {0, 0, 0, rs2}};
      // keep the 2 x 2 problem positive definite (it is, by the increment weights, in every case seen; a guard)
      if (!(q00 > 0.0) || !(q00 * q11 - q01 * q01 > 1e-12 * q00 * q11)) {
        const double shift = fabs(q00) + fabs(q11) + fabs(q01) + 1e-6;
        q00 += shift;
        q11 += shift;
      }
      const double alim = mpc_alim(c, i);
      const double lo[2] = {-alim - S.a[i], -c.max_steer - S.dl[i]}, hi[2] = {alim - S.a[i], c.max_steer - S.dl[i]};
      double d[2];
      bool fr[2];
      box_qp2(q00, q01, q11, Qu[0], Qu[1], lo, hi, d, fr);
      double K[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
      if (fr[0] && fr[1]) {
        const double det = q00 * q11 - q01 * q01;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          K[0][q] = -(q11 * Qux[0][q] - q01 * Qux[1][q]) / det;
          K[1][q] = -(q00 * Qux[1][q] - q01 * Qux[0][q]) / det;
        }
      } else if (fr[0]) {
#pragma unroll
        for (int q = 0; q < 4; ++q) K[0][q] = -Qux[0][q] / q00;
      } else if (fr[1]) {
#pragma unroll
        for (int q = 0; q < 4; ++q) K[1][q] = -Qux[1][q] / q11;
      }
      S.kff[i][0] = d[0];
      S.kff[i][1] = d[1];
      maxd = fmax(maxd, fmax(fabs(d[0]), fabs(d[1])));
      // the gains are kept (and, for consistency, used below) in float32
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        S.kfb[i][0][q] = (float)K[0][q];
        S.kfb[i][1][q] = (float)K[1][q];
        K[0][q] = (double)S.kfb[i][0][q];
        K[1][q] = (double)S.kfb[i][1][q];
      }
      // cost-to-go: g = Qx + K^T (Quu d + Qu) + Qux^T d,  H = Qxx + K^T Quu K + K^T Qux + Qux^T K
      const double t0 = q00 * d[0] + q01 * d[1] + Qu[0], t1 = q01 * d[0] + q11 * d[1] + Qu[1];
#pragma unroll
      for (int q = 0; q < 4; ++q) g[q] = Qx[q] + K[0][q] * t0 + K[1][q] * t1 + Qux[0][q] * d[0] + Qux[1][q] * d[1];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const double k0 = q00 * K[0][r] + q01 * K[1][r], k1 = q01 * K[0][r] + q11 * K[1][r];   // (Quu K)[:, r]
#pragma unroll
        for (int q = 0; q < 4; ++q)
          H[r][q] = Qxx[r][q] + k0 * K[0][q] + k1 * K[1][q] + K[0][r] * Qux[0][q] + K[1][r] * Qux[1][q] + Qux[0][r] * K[0][q] +
                    Qux[1][r] * K[1][q];
      }
    }
    // a Newton step three orders below the tolerance: nothing left to do (and no step of that size could be told
    // from rounding in the cost; without this test a converged lane would walk through every halving below while the
    // rest of its warp waits - measured: 6 of 32 lanes active in the forward sweep)
    if (maxd < 1e-3 * c.tol) {
      ++it;
      break;
    }
    // ---- forward sweep with a halving step: the new controls, their rollout and its cost; the nominal rollout is
    // recomputed alongside (the same arithmetic that produced it: the same values)
    double alpha = 1.0, maxdu = 0.0, cost_n = 0.0, thn = th0, vn = v0;
    bool accepted = false;
    for (int ls = 0; ls < 8; ++ls, alpha *= 0.5) {
      double dth = 0.0, dv = 0.0, dpa = 0.0, dpd = 0.0;   // deviation of the new rollout from the nominal one
      double thm = th0, vm = v0;                          // nominal
      thn = th0;
      vn = v0;
      cost_n = 0.0;
      maxdu = 0.0;
      for (int i = 0; i < N; ++i) {
        const int k = i < M ? i : M - 1;
        if (i < M) {
          double ua = S.a[i] + alpha * S.kff[i][0] + (double)S.kfb[i][0][0] * dth + (double)S.kfb[i][0][1] * dv +
                      (double)S.kfb[i][0][2] * dpa + (double)S.kfb[i][0][3] * dpd;
          double ud = S.dl[i] + alpha * S.kff[i][1] + (double)S.kfb[i][1][0] * dth + (double)S.kfb[i][1][1] * dv +
                      (double)S.kfb[i][1][2] * dpa + (double)S.kfb[i][1][3] * dpd;
          const double alim = mpc_alim(c, i);
          ua = fmin(fmax(ua, -alim), alim);
          ud = fmin(fmax(ud, -c.max_steer), c.max_steer);
          if (i > 0 || have_last) {
            const double da = ua - (i == 0 ? la : S.an[i - 1]), dd = ud - (i == 0 ? ld : S.dn[i - 1]);
            cost_n += c.r_accel * da * da + c.r_steer * dd * dd;
          }
          S.an[i] = ua;
          S.dn[i] = ud;
          S.taun[i] = tan(ud);
          dpa = ua - S.a[i];     // the next stage sees this row as its previous control
          dpd = ud - S.dl[i];
          maxdu = fmax(maxdu, fmax(fabs(dpa), fabs(dpd)));
        }
        const double eth = thn - S.thr[i], ev = vn - S.vr[i];
        cost_n += c.q_theta * eth * eth + c.q_v * ev * ev;
        thn = thn + cdt * vn * S.taun[k];
        vn = vn + dt * S.an[k];
        thm = thm + cdt * vm * S.tau[k];
        vm = vm + dt * S.a[k];
        dth = thn - thm;
        dv = vn - vm;
      }
      {
        const double eth = thn - S.thr[N], ev = vn - S.vr[N];
        cost_n += c.q_theta * eth * eth + c.q_v * ev * ev;
      }
      if (cost_n <= cost + 1e-13 * fabs(cost)) {   // the two costs are summed in different orders: rounding is no rejection
        accepted = true;
        break;
      }
    }
    if (!accepted) break;   // no step size lowers the cost any more: converged as far as float64 goes
    for (int k = 0; k < M; ++k) {
      S.a[k] = S.an[k];
      S.dl[k] = S.dn[k];
      S.tau[k] = S.taun[k];
    }
    thN = thn;
    vN = vn;
    cost = cost_n;
    if (maxdu < c.tol) {
      ++it;
      break;
    }
  }
  return it;
}

template <int HOR, int BLK>
__global__ void __launch_bounds__(64, 8) mpc_track_kernel(const MpcCfg c, double* __restrict__ ws, const long long n, const double dt,
                                                       const int* __restrict__ n_steps, const int* __restrict__ status, const int step_begin,
                                                       const int step_count, double* __restrict__ state, double* __restrict__ states_out,
                                                       double* __restrict__ controls_out, const long long out_rows,
                                                       int* __restrict__ iters_out) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  if (status != nullptr && status[j] != 0) return;
  const int last_step = min(step_begin + step_count, n_steps[j]);
  if (step_begin >= last_step && !(step_begin == 0 && states_out != nullptr)) return;
  MpcLocal<HOR, BLK> S;
  const Knots K{ws, n, j, c.n_way, mpc_f_cx(c), mpc_f_cy(c)};
  double* scal = ws + (size_t)mpc_f_scal(c) * n + j;
  const double t_end = scal[(size_t)MS_T_END * n], start_theta = scal[(size_t)MS_START_THETA * n];
  const double evx = scal[(size_t)MS_END_VX * n], evy = scal[(size_t)MS_END_VY * n], end_theta = scal[(size_t)MS_END_THETA * n];
  double la = scal[(size_t)MS_LAST_A * n], ld = scal[(size_t)MS_LAST_D * n];
  bool have_last = scal[(size_t)MS_HAVE_LAST * n] != 0.0;
  for (int k = 0; k < c.blocks; ++k) {
    S.a[k] = ws[(size_t)(mpc_f_warm(c) + 2 * k) * n + j];
    S.dl[k] = ws[(size_t)(mpc_f_warm(c) + 2 * k + 1) * n + j];
  }
  double x = state[j * 4], y = state[j * 4 + 1], th = state[j * 4 + 2], v = state[j * 4 + 3];
  if (step_begin == 0 && states_out != nullptr) {
    double* o = states_out + (size_t)j * out_rows * 4;
    o[0] = x; o[1] = y; o[2] = th; o[3] = v;
  }
  int cur = 0;
  long long iters = 0;
  for (int s = step_begin; s < last_step; ++s) {
    const double current_time = s * dt;   // run_simulation: current_time = i * self.dt (:513)
    mpc_window(c, K, t_end, start_theta, evx, evy, end_theta, current_time, dt, cur, S.thr, S.vr);
    iters += mpc_solve(c, dt, th, v, have_last, la, ld, S);
    la = S.a[0];
    ld = S.dl[0];
    have_last = true;
    // the first control drives the vehicle for one step (:484-486; the controls are clipped inside the dynamics, :55-56)
    const double ua = fmin(fmax(la, -c.max_accel), c.max_accel), ud = fmin(fmax(ld, -c.max_steer), c.max_steer);
    const double dx = v * cos(th), dy = v * sin(th), dth = v * tan(ud) / c.wheelbase;
    x += dx * dt;
    y += dy * dt;
    th += dth * dt;
    v += ua * dt;
    if (states_out != nullptr) {
      double* o = states_out + ((size_t)j * out_rows + (size_t)(s + 1)) * 4;
      o[0] = x; o[1] = y; o[2] = th; o[3] = v;
    }
    if (controls_out != nullptr) {
      double* o = controls_out + ((size_t)j * (out_rows - 1) + (size_t)s) * 2;
      o[0] = la; o[1] = ld;
    }
    // warm start of the next call: the solution moved up by one row (rows keep within their own limits: a later row's
    // acceleration limit is never wider than an earlier one's)
    for (int k = 0; k + 1 < c.blocks; ++k) {
      S.a[k] = S.a[k + 1];
      S.dl[k] = S.dl[k + 1];
    }
  }
  state[j * 4] = x; state[j * 4 + 1] = y; state[j * 4 + 2] = th; state[j * 4 + 3] = v;
  scal[(size_t)MS_LAST_A * n] = la;
  scal[(size_t)MS_LAST_D * n] = ld;
  scal[(size_t)MS_HAVE_LAST * n] = have_last ? 1.0 : 0.0;
  for (int k = 0; k < c.blocks; ++k) {
    ws[(size_t)(mpc_f_warm(c) + 2 * k) * n + j] = S.a[k];
    ws[(size_t)(mpc_f_warm(c) + 2 * k + 1) * n + j] = S.dl[k];
  }
  if (iters_out != nullptr) iters_out[j] = (step_begin == 0 ? 0 : iters_out[j]) + (int)iters;
}

// reference windows at given times, for tests of the interpolation path alone: out (n, n_times, horizon + 1, 2)
__global__ void __launch_bounds__(128) mpc_window_kernel(const MpcCfg c, const double* __restrict__ ws, const long long n, const double dt,
                                                         const double* __restrict__ times, const int n_times, const int* __restrict__ status,
                                                         double* __restrict__ out) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n || (status != nullptr && status[j] != 0)) return;
  const Knots K{ws, n, j, c.n_way, mpc_f_cx(c), mpc_f_cy(c)};
  const double* scal = ws + (size_t)mpc_f_scal(c) * n + j;
  double thr[MPC_MAX_HOR + 1], vr[MPC_MAX_HOR + 1];
  int cur = 0;
  for (int k = 0; k < n_times; ++k) {
    mpc_window(c, K, scal[(size_t)MS_T_END * n], scal[(size_t)MS_START_THETA * n], scal[(size_t)MS_END_VX * n], scal[(size_t)MS_END_VY * n],
               scal[(size_t)MS_END_THETA * n], times[k], dt, cur, thr, vr);
    double* o = out + ((size_t)j * n_times + k) * (size_t)(c.horizon + 1) * 2;
    for (int i = 0; i <= c.horizon; ++i) {
      o[2 * i] = thr[i];
      o[2 * i + 1] = vr[i];
    }
  }
}

cudaError_t launch_mpc_prepare(const MpcCfg& c, const void* way, const double* init, double* ws, long long n, double* state, int* status,
                               double* profile, cudaStream_t stream) {
  mpc_prepare_kernel<<<(unsigned int)((n + 127) / 128), 128, 0, stream>>>(c, way, init, ws, n, state, status, profile);
  return cudaGetLastError();
}

cudaError_t launch_mpc_track(const MpcCfg& c, double* ws, long long n, double dt, const int* n_steps, const int* status, int step_begin,
                             int step_count, double* state, double* states_out, double* controls_out, long long out_rows, int* iters_out,
                             cudaStream_t stream) {
  // the reference's own horizons (Distribution.py:98-99) have their own, smaller instantiation
  auto kernel = c.horizon <= 30 && c.blocks <= 20 ? mpc_track_kernel<30, 20> : mpc_track_kernel<MPC_MAX_HOR, MPC_MAX_BLK>;
  kernel<<<(unsigned int)((n + 63) / 64), 64, 0, stream>>>(c, ws, n, dt, n_steps, status, step_begin, step_count, state, states_out, controls_out,
                                                          out_rows, iters_out);
  return cudaGetLastError();
}

cudaError_t launch_mpc_windows(const MpcCfg& c, const double* ws, long long n, double dt, const double* times, int n_times, const int* status,
                               double* out, cudaStream_t stream) {
  mpc_window_kernel<<<(unsigned int)((n + 127) / 128), 128, 0, stream>>>(c, ws, n, dt, times, n_times, status, out);
  return cudaGetLastError();
}

}  // namespace dmvae
