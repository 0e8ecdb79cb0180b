/*
 * oracle_c.c -- plain-C fp64 restatement of the reference algorithm.  TEST INFRASTRUCTURE ONLY.
 *
 * The twin of oracle/minsnap_np.py + oracle/flight_np.py + oracle/freebody.py, written so the checker can
 * fly thousands of whole missions in seconds (large parity samples for the CUDA path) and so bench.py can
 * quote an optimised multi-core CPU number next to the NumPy-style one.  Never linked into libuavb.so and
 * never imported by the product package.  Pinned by tests/test_oracle_c.py against the NumPy oracle and the
 * reference-generated goldens (tests/golden/).
 *
 * Citations: ms = /root/reference/uav_ac/planning/minimum_snap.py, ctl = uav_ac/control/controller.py,
 * quad = uav_ac/quadrotor/quad.py, main = uav_ac/main.py, sim = uav_ac/simulation/mujoco_sim.py.
 * The rigid-body step restates MuJoCo's documented Euler update (parity unpinned, see oracle/freebody.py).
 *
 * Build: oracle/build.py  (gcc -O2 -fPIC -shared -pthread, output oracle/_build/liboracle_c.so)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NC 8 /* ms:28 */

/* ------------------------------------------------------------------------------------------ planner */

/* ms:258-286: k-th derivative of the ascending monomial basis at t */
static void basis_row(int order, double t, double* out) {
  for (int i = 0; i < NC; ++i) {
    double fall = 1.0;
    int e = i;
    for (int k = 0; k < order; ++k) {
      fall *= (double)e;
      if (e > 0) --e;
    }
    out[i] = fall * pow(t, (double)e);
  }
}

/* ms:311-321 */
void oracle_segment_times(const double* w, int S, double velocity, double factor, double* T) {
  for (int i = 0; i < S; ++i) {
    const double dx = w[3 * (i + 1)] - w[3 * i], dy = w[3 * (i + 1) + 1] - w[3 * i + 1], dz = w[3 * (i + 1) + 2] - w[3 * i + 2];
    T[i] = sqrt(dx * dx + dy * dy + dz * dz) / velocity;
    if (i == 0 || i == S - 1) T[i] *= factor;
  }
}

/* Dense LU with partial pivoting (what np.linalg.solve / LAPACK gesv does), n x n matrix, nrhs right-hand sides,
 * both row-major, solution overwrites B.  Returns 0, or -1 when a pivot is exactly zero (LinAlgError). */
static int lu_solve(double* A, double* B, int n, int nrhs) {
  for (int k = 0; k < n; ++k) {
    int p = k;
    double best = fabs(A[k * n + k]);
    for (int r = k + 1; r < n; ++r)
      if (fabs(A[r * n + k]) > best) { best = fabs(A[r * n + k]); p = r; }
    if (best == 0.0) return -1;
    if (p != k) {
      for (int c = 0; c < n; ++c) { double t = A[k * n + c]; A[k * n + c] = A[p * n + c]; A[p * n + c] = t; }
      for (int c = 0; c < nrhs; ++c) { double t = B[k * nrhs + c]; B[k * nrhs + c] = B[p * nrhs + c]; B[p * nrhs + c] = t; }
    }
    const double inv = 1.0 / A[k * n + k];
    for (int r = k + 1; r < n; ++r) {
      const double f = A[r * n + k] * inv;
      if (f == 0.0) continue;
      A[r * n + k] = f;
      for (int c = k + 1; c < n; ++c) A[r * n + c] -= f * A[k * n + c];
      for (int c = 0; c < nrhs; ++c) B[r * nrhs + c] -= f * B[k * nrhs + c];
    }
  }
  for (int k = n - 1; k >= 0; --k) {
    for (int c = 0; c < nrhs; ++c) {
      double s = B[k * nrhs + c];
      for (int j = k + 1; j < n; ++j) s -= A[k * n + j] * B[j * nrhs + c];
      B[k * nrhs + c] = s / A[k * n + k];
    }
  }
  return 0;
}

/* ms:138-255: assemble [[Q, A^T], [A, 0]] [c; lam] = [0; b] in the reference's row order and solve it.
 * coeffs_out [8S][3] (row 8 i + j = coefficient of t^j of spline i), times_out [S].  Returns 0 or -1 (singular). */
int oracle_minsnap_solve(const double* w, int S, double velocity, double factor, double* coeffs_out, double* times_out) {
  const int n = 8 * S, m = 6 * S + 2, N = n + m;
  double* T = times_out;
  oracle_segment_times(w, S, velocity, factor, T);
  double* K = (double*)calloc((size_t)N * N, sizeof(double));
  double* rhs = (double*)calloc((size_t)N * 3, sizeof(double));
  double row[NC];
  if (!K || !rhs) { free(K); free(rhs); return -2; }
  /* Q: ms:155-169 */
  for (int s = 0; s < S; ++s)
    for (int r = 4; r < 8; ++r)
      for (int c = 4; c < 8; ++c) {
        const double fr = (double)(r * (r - 1) * (r - 2) * (r - 3)), fc = (double)(c * (c - 1) * (c - 2) * (c - 3));
        const int e = r + c - 7;
        K[(8 * s + r) * N + 8 * s + c] = fr * fc * pow(T[s], (double)e) / (double)e;
      }
  /* A and b: ms:171-255 */
  int r = 0;
#define PUT(col0, sign)                                                   \
  for (int j = 0; j < NC; ++j) {                                          \
    K[(n + r) * N + (col0) + j] += (sign) * row[j];                       \
    K[((col0) + j) * N + n + r] += (sign) * row[j];                       \
  }
  for (int i = 0; i < S; ++i) {                      /* position at t = 0 */
    basis_row(0, 0.0, row);
    PUT(8 * i, 1.0)
    for (int a = 0; a < 3; ++a) rhs[(n + r) * 3 + a] = w[3 * i + a];
    ++r;
  }
  for (int i = 0; i < S; ++i) {                      /* position at t = T_i */
    basis_row(0, T[i], row);
    PUT(8 * i, 1.0)
    for (int a = 0; a < 3; ++a) rhs[(n + r) * 3 + a] = w[3 * (i + 1) + a];
    ++r;
  }
  for (int k = 1; k <= 3; ++k) { basis_row(k, 0.0, row); PUT(0, 1.0) ++r; }                 /* start at rest */
  for (int k = 1; k <= 3; ++k) { basis_row(k, T[S - 1], row); PUT(8 * (S - 1), 1.0) ++r; }  /* end at rest */
  for (int s = 1; s < S; ++s)                        /* continuity of derivatives 1..4 */
    for (int k = 1; k <= 4; ++k) {
      basis_row(k, T[s - 1], row);
      PUT(8 * (s - 1), 1.0)
      basis_row(k, 0.0, row);
      PUT(8 * s, -1.0)
      ++r;
    }
#undef PUT
  const int rc = lu_solve(K, rhs, N, 3);
  if (rc == 0) memcpy(coeffs_out, rhs, sizeof(double) * (size_t)n * 3);
  free(K);
  free(rhs);
  return rc;
}

/* len(np.arange(0, T, dt)) = ceil(T / dt) (ms:104) */
int oracle_sample_count(double T, double dt) {
  const double n = ceil(T / dt);
  return n > 0.0 ? (int)n : 0;
}

/* ms:97-136: (N, 11) table [pos3 vel3 acc3 yaw spline_id]; returns the number of rows written (<= max_rows). */
int oracle_sample_table(const double* coeffs, const double* T, int S, double dt, double* table, int max_rows) {
  int n = 0;
  double p[NC], v[NC], a[NC];
  for (int i = 0; i < S; ++i) {
    const int cnt = oracle_sample_count(T[i], dt);
    for (int j = 0; j < cnt && n < max_rows; ++j, ++n) {
      const double t = (double)j * dt;
      basis_row(0, t, p); basis_row(1, t, v); basis_row(2, t, a);
      double* o = table + (size_t)n * 11;
      for (int ax = 0; ax < 3; ++ax) {
        double sp = 0, sv = 0, sa = 0;
        for (int k = 0; k < NC; ++k) {
          const double c = coeffs[(8 * i + k) * 3 + ax];
          sp += p[k] * c; sv += v[k] * c; sa += a[k] * c;
        }
        o[ax] = sp; o[3 + ax] = sv; o[6 + ax] = sa;
      }
      o[9] = 0.0; o[10] = (double)i;
    }
  }
  /* ms:126-136: heading of valid rows, np.unwrap over them, hold-last-valid, first-valid look-ahead */
  const double two_pi = 2.0 * M_PI;
  int first = -1;
  for (int r = 0; r < n; ++r) {
    const double vx = table[(size_t)r * 11 + 3], vy = table[(size_t)r * 11 + 4];
    if (sqrt(vx * vx + vy * vy) >= 1e-3) { first = r; break; }
  }
  if (first < 0) return n;
  double prev_raw = atan2(table[(size_t)first * 11 + 4], table[(size_t)first * 11 + 3]), hold = prev_raw;
  for (int r = 0; r < n; ++r) {
    const double vx = table[(size_t)r * 11 + 3], vy = table[(size_t)r * 11 + 4];
    if (r > first && sqrt(vx * vx + vy * vy) >= 1e-3) {
      const double raw = atan2(vy, vx), dd = raw - prev_raw;
      double ddmod = fmod(dd + M_PI, two_pi);
      if (ddmod < 0) ddmod += two_pi;
      ddmod -= M_PI;
      if (ddmod == -M_PI && dd > 0) ddmod = M_PI;
      hold += dd + ((fabs(dd) < M_PI) ? 0.0 : (ddmod - dd));
      prev_raw = raw;
    }
    table[(size_t)r * 11 + 9] = hold;
  }
  return n;
}

/* ------------------------------------------------------------------------------------------ vehicle + controller */

typedef struct oracle_vehicle {  /* same field order as struct uavb_vehicle (include/uavb.h) */
  double g, dt, mass, inertia[3], arm, kf, kappa, min_thrust, max_thrust, tau_rise, tau_fall;
  double max_ascent, max_descent, max_speed_xy, max_horiz_accel, max_tilt;
  double gains[11]; /* kp_xy kd_xy kp_z kd_z ki_z kp_roll kp_pitch kp_yaw kp_p kp_q kp_r */
  double integral_limit;
  double ground_on, ground_z; /* optional unilateral floor (not in uavb_vehicle: uavb_rollout_args.ground_on / ground_z); 0 = free body */
} oracle_vehicle;

static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
static double pymod(double a, double b) { double r = fmod(a, b); if (r != 0.0 && ((r < 0) != (b < 0))) r += b; return r; }

/* quad:133-155 */
static void quat_to_rot(const double* q, double* R) {
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double q0 = q[0] / n, q1 = q[1] / n, q2 = q[2] / n, q3 = q[3] / n;
  R[0] = 1 - 2 * (q2 * q2 + q3 * q3); R[1] = 2 * (q1 * q2 - q0 * q3); R[2] = 2 * (q1 * q3 + q0 * q2);
  R[3] = 2 * (q1 * q2 + q0 * q3); R[4] = 1 - 2 * (q1 * q1 + q3 * q3); R[5] = 2 * (q2 * q3 - q0 * q1);
  R[6] = 2 * (q1 * q3 - q0 * q2); R[7] = 2 * (q2 * q3 + q0 * q1); R[8] = 1 - 2 * (q1 * q1 + q2 * q2);
}

typedef struct {
  double integral, thrust_cmd, pqr_cmd[3];
} ctrl_state;

/* main:47-61 with ctl:26-168 */
static void outer_loop(const oracle_vehicle* v, const double* X, const double* row, double dt_outer, ctrl_state* cs) {
  double R[9];
  quat_to_rot(X + 3, R);
  /* altitude ctl:26-56 */
  const double climb = clampd(row[5], -v->max_ascent, v->max_descent);
  const double e = row[2] - X[2], ed = climb - X[9];
  cs->integral = clampd(cs->integral + e * dt_outer, -v->integral_limit, v->integral_limit);
  double acc = v->gains[2] * e + v->gains[4] * cs->integral + v->gains[3] * ed + row[8] - v->g;
  acc /= R[8];
  const double c = clampd(-v->mass * acc, 4 * v->min_thrust, 4 * v->max_thrust);
  cs->thrust_cmd = c;
  /* lateral ctl:58-97 */
  double vxd = row[3], vyd = row[4];
  const double vmag = sqrt(vxd * vxd + vyd * vyd);
  if (vmag > v->max_speed_xy) { vxd = vxd / vmag * v->max_speed_xy; vyd = vyd / vmag * v->max_speed_xy; }
  double ax = v->gains[0] * (row[0] - X[0]) + v->gains[1] * (vxd - X[7]) + row[6];
  double ay = v->gains[0] * (row[1] - X[1]) + v->gains[1] * (vyd - X[8]) + row[7];
  const double amag = sqrt(ax * ax + ay * ay);
  if (amag > v->max_horiz_accel) { ax = ax / amag * v->max_horiz_accel; ay = ay / amag * v->max_horiz_accel; }
  const double accz = -c / v->mass;
  const double bx = clampd(ax / accz, -v->max_tilt, v->max_tilt), by = clampd(ay / accz, -v->max_tilt, v->max_tilt);
  /* roll / pitch ctl:132-154 */
  const double bdx = v->gains[5] * (bx - R[2]), bdy = v->gains[6] * (by - R[5]);
  const double pc = (R[3] * bdx - R[0] * bdy) / R[8], qc = (R[4] * bdx - R[1] * bdy) / R[8];
  /* yaw ctl:156-168, Euler angles from the raw quaternion quad:189-213 */
  const double q0 = X[3], q1 = X[4], q2 = X[5], q3 = X[6];
  const double phi = atan2(2 * (q0 * q1 + q2 * q3), 1 - 2 * (q1 * q1 + q2 * q2));
  const double theta = asin(clampd(2 * (q0 * q2 - q3 * q1), -1.0, 1.0));
  const double psi = atan2(2 * (q0 * q3 + q1 * q2), 1 - 2 * (q2 * q2 + q3 * q3));
  const double err = pymod(pymod(row[9], 2 * M_PI) - psi + M_PI, 2 * M_PI) - M_PI;
  cs->pqr_cmd[0] = pc; cs->pqr_cmd[1] = qc;
  cs->pqr_cmd[2] = (v->gains[7] * err * cos(theta) - qc * sin(phi)) / cos(phi);
}

/* main:42-44: ctl:115-130, quad:88-122 */
static void inner_loop(const oracle_vehicle* v, const double* X, const ctrl_state* cs, double a_rise, double a_fall, double* omega) {
  const double* I = v->inertia;
  const double* w = X + 10;
  const double Iw[3] = {I[0] * w[0], I[1] * w[1], I[2] * w[2]};
  const double M[3] = {I[0] * v->gains[8] * (cs->pqr_cmd[0] - w[0]) + (w[1] * Iw[2] - w[2] * Iw[1]),
                       I[1] * v->gains[9] * (cs->pqr_cmd[1] - w[1]) + (w[2] * Iw[0] - w[0] * Iw[2]),
                       I[2] * v->gains[10] * (cs->pqr_cmd[2] - w[2]) + (w[0] * Iw[1] - w[1] * Iw[0])};
  const double c_bar = clampd(cs->thrust_cmd, 4 * v->min_thrust, 4 * v->max_thrust);
  const double pb = M[0] / v->arm, qb = M[1] / v->arm, rb = -M[2] / v->kappa;
  const double mf[4] = {(pb + qb + rb) / 4, (-pb + qb - rb) / 4, (-pb - qb + rb) / 4, (pb - qb - rb) / 4};
  const double coll = c_bar / 4;
  double s = 1.0;
  for (int i = 0; i < 4; ++i) {
    double lim = 1.0;
    if (mf[i] > 0) lim = (v->max_thrust - coll) / mf[i];
    else if (mf[i] < 0) lim = (v->min_thrust - coll) / mf[i];
    if (lim < s) s = lim;
  }
  s = clampd(s, 0.0, 1.0);
  for (int i = 0; i < 4; ++i) {
    const double f = clampd(coll + s * mf[i], v->min_thrust, v->max_thrust);
    const double cmd = sqrt(f / v->kf);
    omega[i] += ((cmd > omega[i]) ? a_rise : a_fall) * (cmd - omega[i]);
  }
}

/* sim:232-255 + MuJoCo Euler free-joint step (oracle/freebody.py) */
static void freebody_step(const oracle_vehicle* v, double* X, const double* omega, const double* R, const double* wind) {
  const double f[4] = {v->kf * omega[0] * omega[0], v->kf * omega[1] * omega[1], v->kf * omega[2] * omega[2], v->kf * omega[3] * omega[3]};
  const double thrust = f[0] + f[1] + f[2] + f[3];
  const double tau[3] = {v->arm * (f[0] + f[3] - f[1] - f[2]), v->arm * (f[0] + f[1] - f[2] - f[3]), v->kappa * (-f[0] + f[1] - f[2] + f[3])};
  double F[3] = {R[2] * -thrust, R[5] * -thrust, R[8] * -thrust};
  if (wind) { F[0] += wind[0]; F[1] += wind[1]; F[2] += wind[2]; }
  const double* I = v->inertia;
  double* w = X + 10;
  const double Iw[3] = {I[0] * w[0], I[1] * w[1], I[2] * w[2]};
  const double wd[3] = {(tau[0] - (w[1] * Iw[2] - w[2] * Iw[1])) / I[0], (tau[1] - (w[2] * Iw[0] - w[0] * Iw[2])) / I[1],
                        (tau[2] - (w[0] * Iw[1] - w[1] * Iw[0])) / I[2]};
  X[7] += v->dt * (F[0] / v->mass); X[8] += v->dt * (F[1] / v->mass); X[9] += v->dt * (v->g + F[2] / v->mass);
  w[0] += v->dt * wd[0]; w[1] += v->dt * wd[1]; w[2] += v->dt * wd[2];
  X[0] += v->dt * X[7]; X[1] += v->dt * X[8]; X[2] += v->dt * X[9];
  const double wn = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  double ax = 1, ay = 0, az = 0, ang = 0;
  if (wn >= 1e-15) { ax = w[0] / wn; ay = w[1] / wn; az = w[2] / wn; ang = v->dt * wn; }
  const double sn = sin(0.5 * ang), cs = cos(0.5 * ang);
  const double b0 = cs, b1 = sn * ax, b2 = sn * ay, b3 = sn * az;
  double* q = X + 3;
  const double qn = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double a0 = q[0] / qn, a1 = q[1] / qn, a2 = q[2] / qn, a3 = q[3] / qn;
  double n0 = a0 * b0 - a1 * b1 - a2 * b2 - a3 * b3, n1 = a0 * b1 + a1 * b0 + a2 * b3 - a3 * b2;
  double n2 = a0 * b2 - a1 * b3 + a2 * b0 + a3 * b1, n3 = a0 * b3 + a1 * b2 - a2 * b1 + a3 * b0;
  const double nn = sqrt(n0 * n0 + n1 * n1 + n2 * n2 + n3 * n3);
  q[0] = n0 / nn; q[1] = n1 / nn; q[2] = n2 / nn; q[3] = n3 / nn;
}

/* Headless mission (integration test :26-31 + main:37-61) for ONE drone on a sampled table.
 * out_metrics[8] = final_dist, collision, rmse, mean_err, max_err, status(0), first_hit, periods; X_out[13]; omega_out[4].
 * log_out (optional): state after every log_stride ticks, [n_ticks / log_stride][13]. */
void oracle_closed_loop(const oracle_vehicle* v, const double* table, int n_rows, const double* start, int freq, int n_ticks,
                        const double* obstacles, int n_obs, const double* goal, const double* wind, int thrust_frame_lag,
                        double* out_metrics, double* X_out, double* omega_out, int log_stride, double* log_out) {
  double X[13] = {0}, omega[4] = {0}, R_stale[9], R_now[9];
  X[0] = start[0]; X[1] = start[1]; X[2] = start[2]; X[3] = 1.0;
  ctrl_state cs = {0.0, 0.0, {0.0, 0.0, 0.0}};
  const double dt_outer = v->dt * freq;
  const double a_rise = 1 - exp(-v->dt / v->tau_rise), a_fall = 1 - exp(-v->dt / v->tau_fall);
  int idx = 0, collided = 0, first_hit = -1, periods = 0, n_log = 0;
  double sum_e = 0, sum_e2 = 0, max_e = 0;
  const double* row = table;
  quat_to_rot(X + 3, R_stale);
  for (int k = 0; k < n_ticks; ++k) {
    if (k % freq == 0) {
      row = table + (size_t)idx * 11;
      outer_loop(v, X, row, dt_outer, &cs);
      idx = idx + 1 < n_rows ? idx + 1 : n_rows - 1;
    }
    inner_loop(v, X, &cs, a_rise, a_fall, omega);
    quat_to_rot(X + 3, R_now);
    freebody_step(v, X, omega, thrust_frame_lag ? R_stale : R_now, wind);
    memcpy(R_stale, R_now, sizeof(R_now));
    if (v->ground_on != 0.0 && X[2] > v->ground_z) { /* below the floor (NED): back onto it, no downward velocity */
      X[2] = v->ground_z;
      if (X[9] > 0.0) X[9] = 0.0;
    }
    if (!collided)
      for (int b = 0; b < n_obs; ++b) {
        const double* q = obstacles + 6 * b;
        if (q[0] <= X[0] && X[0] <= q[1] && q[2] <= X[1] && X[1] <= q[3] && q[4] <= X[2] && X[2] <= q[5]) { collided = 1; first_hit = k; break; }
      }
    if (log_out && log_stride > 0 && (k + 1) % log_stride == 0) memcpy(log_out + (size_t)(n_log++) * 13, X, sizeof(X));
    if ((k + 1) % freq == 0) {
      const double ex = X[0] - row[0], ey = X[1] - row[1], ez = X[2] - row[2];
      const double e2 = ex * ex + ey * ey + ez * ez, e = sqrt(e2);
      sum_e += e; sum_e2 += e2; if (e > max_e) max_e = e;
      ++periods;
    }
  }
  double fd = 0.0;
  if (goal) fd = sqrt((X[0] - goal[0]) * (X[0] - goal[0]) + (X[1] - goal[1]) * (X[1] - goal[1]) + (X[2] - goal[2]) * (X[2] - goal[2]));
  out_metrics[0] = fd; out_metrics[1] = collided; out_metrics[2] = periods ? sqrt(sum_e2 / periods) : 0.0;
  out_metrics[3] = periods ? sum_e / periods : 0.0; out_metrics[4] = max_e; out_metrics[5] = 0.0;
  out_metrics[6] = first_hit; out_metrics[7] = periods;
  if (X_out) memcpy(X_out, X, sizeof(X));
  if (omega_out) memcpy(omega_out, omega, sizeof(omega));
}

/* ------------------------------------------------------------------------------------------ batches (pthreads) */
#include <pthread.h>

typedef struct {
  int tid, n_threads, B;
  /* closed loop */
  const oracle_vehicle* vehicles; int veh_stride; const double* table; int n_rows; const double* start; int freq, n_ticks;
  const double* obstacles; int n_obs; const double* goal; const double* wind; int lag; double* metrics_out; double* X_out;
  /* solves */
  const double* w; const double* velocity; int S; double factor; double* coeffs_out; double* times_out; int bad;
} job_t;

static void* fly_worker(void* arg) {
  job_t* j = (job_t*)arg;
  for (int b = j->tid; b < j->B; b += j->n_threads)
    oracle_closed_loop(j->vehicles + (size_t)j->veh_stride * b, j->table, j->n_rows, j->start, j->freq, j->n_ticks, j->obstacles, j->n_obs, j->goal,
                       j->wind ? j->wind + 3 * (size_t)b : NULL, j->lag, j->metrics_out + 8 * (size_t)b, j->X_out ? j->X_out + 13 * (size_t)b : NULL,
                       NULL, 0, NULL);
  return NULL;
}

static void* solve_worker(void* arg) {
  job_t* j = (job_t*)arg;
  for (int b = j->tid; b < j->B; b += j->n_threads)
    if (oracle_minsnap_solve(j->w + (size_t)b * (j->S + 1) * 3, j->S, j->velocity[b], j->factor, j->coeffs_out + (size_t)b * 24 * j->S,
                             j->times_out + (size_t)b * j->S))
      ++j->bad;
  return NULL;
}

static void run_jobs(job_t* proto, void* (*fn)(void*)) {
  int n = proto->n_threads < 1 ? 1 : (proto->n_threads > 256 ? 256 : proto->n_threads);
  pthread_t th[256];
  job_t jobs[256];
  for (int t = 0; t < n; ++t) { jobs[t] = *proto; jobs[t].tid = t; jobs[t].n_threads = n; jobs[t].bad = 0; }
  for (int t = 1; t < n; ++t) pthread_create(&th[t], NULL, fn, &jobs[t]);
  fn(&jobs[0]);
  for (int t = 1; t < n; ++t) pthread_join(th[t], NULL);
  proto->bad = 0;
  for (int t = 0; t < n; ++t) proto->bad += jobs[t].bad;
}

/* B drones on one shared table with per-drone vehicles (Monte-Carlo), drones striped over n_threads threads.
 * vehicles [B] (veh_stride 1) or a single vehicle (veh_stride 0); metrics_out [B][8]; X_out [B][13] or NULL. */
void oracle_closed_loop_batch(const oracle_vehicle* vehicles, int veh_stride, int B, const double* table, int n_rows, const double* start,
                              int freq, int n_ticks, const double* obstacles, int n_obs, const double* goal, const double* wind /* [B][3] or NULL */,
                              int thrust_frame_lag, double* metrics_out, double* X_out, int n_threads) {
  job_t j;
  memset(&j, 0, sizeof(j));
  j.n_threads = n_threads; j.B = B; j.vehicles = vehicles; j.veh_stride = veh_stride; j.table = table; j.n_rows = n_rows; j.start = start;
  j.freq = freq; j.n_ticks = n_ticks; j.obstacles = obstacles; j.n_obs = n_obs; j.goal = goal; j.wind = wind; j.lag = thrust_frame_lag;
  j.metrics_out = metrics_out; j.X_out = X_out;
  run_jobs(&j, fly_worker);
}

/* B independent solves (config 2), missions striped over n_threads threads.  Returns the number of singular systems. */
int oracle_minsnap_solve_batch(const double* w, const double* velocity, int B, int S, double factor, double* coeffs_out, double* times_out,
                               int n_threads) {
  job_t j;
  memset(&j, 0, sizeof(j));
  j.n_threads = n_threads; j.B = B; j.w = w; j.velocity = velocity; j.S = S; j.factor = factor; j.coeffs_out = coeffs_out; j.times_out = times_out;
  run_jobs(&j, solve_worker);
  return j.bad;
}
