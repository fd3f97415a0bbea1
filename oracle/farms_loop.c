/*
 * farms_loop.c -- the farms_mujoco per-iteration data plane in C, fp64.
 *
 * TEST INFRASTRUCTURE ONLY (see the header of mjstep_oracle.c): bench.py's cpu_baseline /
 * --impl reference legs time it, tests/test_oracle_farms_loop.py checks it bit-for-bit against
 * the NumPy restatement (oracle/farms_oracle.py), which stays the line-by-line checker.  Nothing
 * under farms_mujoco_b200/ links or calls it.
 *
 * Why it exists: the reference runs this glue as compiled code -- drag_forces / SwimmingHandler.
 * step as Cython `nogil` C loops built with -O3 (farms_mujoco/swimming/drag.pyx:152-268,389-411,
 * setup.py:60-61), cycontacts2data as Cython (sensors/sensors.pyx:140-190), the gathers as a
 * dozen NumPy kernels (simulation/physics.py:435-524) -- so a CPU baseline that spends 77 % of
 * its time in per-link Python quaternion calls (round 1) flatters the GPU.  One call of
 * orc_farms_run() = K iterations of Simulation.run for ONE environment in the reference's order
 * (task.py:168-186: sensors -> swimming callback -> control, then mj_step, simulation.py:156),
 * with no interpreter in the loop.
 *
 * Each function names the farms_oracle.py function it restates and the reference lines behind it.
 */
#define _POSIX_C_SOURCE 200809L
#define _DEFAULT_SOURCE
#include <math.h>
#include <string.h>
#include <time.h>
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#include "../include/farms_b200.h"
#include "oracle.h"

void orc_step(const FbModel *m, OrcData *d);
void orc_contact_force(const FbModel *m, const OrcData *d, int i, double *result);

/* farms_core column layout (farms_mujoco_b200/layout.py) */
enum { L_COM_POS = 0, L_COM_QUAT = 3, L_URDF_POS = 7, L_URDF_QUAT = 10, L_LIN = 14, L_ANG = 17, L_COLS = 20 };
enum { C_REACTION = 0, C_FRICTION = 3, C_TOTAL = 6, C_POSITION = 9, C_COLS = 12 };

typedef struct FarmsUnits { double meters, newtons, torques, velocity, angular_velocity; } FarmsUnits;

static FarmsUnits units_of(const FbFarms *f) {
  /* farms_core SimulationUnitScaling (units.py): derived factors of (meters, seconds, kilograms) */
  FarmsUnits u;
  u.meters = f->meters;
  u.velocity = f->meters/f->seconds;
  u.angular_velocity = 1.0/f->seconds;
  u.newtons = f->kilograms*f->meters/(f->seconds*f->seconds);
  u.torques = f->kilograms*f->meters*f->meters/(f->seconds*f->seconds);
  return u;
}

/* ---- xyzw quaternions (farms_oracle.py: quat_conj / quat_mult / quat_rot) ---------------- */
static void q_conj(const double *q, double *r) { r[0] = -q[0]; r[1] = -q[1]; r[2] = -q[2]; r[3] = q[3]; }
static void q_mult(const double *a, const double *b, double *r) {
  const double x0 = a[0], y0 = a[1], z0 = a[2], w0 = a[3], x1 = b[0], y1 = b[1], z1 = b[2], w1 = b[3];
  r[0] = w0*x1 + x0*w1 + y0*z1 - z0*y1;
  r[1] = w0*y1 - x0*z1 + y0*w1 + z0*x1;
  r[2] = w0*z1 + x0*y1 - y0*x1 + z0*w1;
  r[3] = w0*w1 - x0*x1 - y0*y1 - z0*z1;
}
static void q_rot(const double *v, const double *q, double *out) {
  double v4[4] = {v[0], v[1], v[2], 0.0}, qc[4], t[4], r[4];
  q_conj(q, qc);
  q_mult(q, v4, t);
  q_mult(t, qc, r);
  out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
}

/* ---- physics2data (farms_oracle.py: physics2data; physics.py:527-545) --------------------- */
static void links_row(const OrcData *d, const FbFarms *f, const FarmsUnits *u, double *row) {
  /* physicslinks2data physics.py:449-466 + physicslinksvelsensors2data :435-446 */
  for (int l = 0; l < f->n_links; l++) {
    const int b = f->link_body[l];
    double *r = row + (size_t)l*L_COLS;
    const double *q = d->xquat + 4*b;
    for (int k = 0; k < 3; k++) {
      r[L_URDF_POS + k] = d->xpos[3*b + k]/u->meters;
      r[L_COM_POS + k] = d->xipos[3*b + k]/u->meters;
      r[L_LIN + k] = d->body_linvel[3*b + k]/u->velocity;
      r[L_ANG + k] = d->body_angvel[3*b + k]/u->angular_velocity;
    }
    /* wxyz -> xyzw; the CoM orientation column takes the BODY quaternion too (:463-466) */
    r[L_URDF_QUAT] = q[1]; r[L_URDF_QUAT + 1] = q[2]; r[L_URDF_QUAT + 2] = q[3]; r[L_URDF_QUAT + 3] = q[0];
    r[L_COM_QUAT] = q[1]; r[L_COM_QUAT + 1] = q[2]; r[L_COM_QUAT + 2] = q[3]; r[L_COM_QUAT + 3] = q[0];
  }
}

static void joints_row(const FbModel *m, const OrcData *d, const FbFarms *f, const FarmsUnits *u, double *row) {
  /* physicsjointssensors2data :481-497, physicsjoints2data :500-507, physicsactuators2data :510-524
   * (accumulating: the caller zeroes a ring row before it is revisited) */
  (void)m;
  const double itorques = 1.0/u->torques;
  for (int j = 0; j < f->n_joints; j++) {
    double *r = row + (size_t)j*f->joint_cols;
    if (f->joint_jntid[j] >= 0) r[f->col_joint_limit_force] = d->jnt_limit_force[f->joint_jntid[j]]/u->torques;
    r[f->col_joint_position] = d->qpos[f->joint_qposadr[j]];
    r[f->col_joint_velocity] = d->qvel[f->joint_dofadr[j]]/u->angular_velocity;
    if (f->joint_act_position[j] >= 0) r[f->col_joint_torque] += d->actuator_force[f->joint_act_position[j]]*itorques;
    if (f->joint_act_velocity[j] >= 0) r[f->col_joint_torque] += d->actuator_force[f->joint_act_velocity[j]]*itorques;
    if (f->joint_act_torque[j] >= 0) r[f->col_joint_torque] += d->actuator_force[f->joint_act_torque[j]]*itorques;
  }
}

static void contacts_row(const FbModel *m, const OrcData *d, const FbFarms *f, const FarmsUnits *u,
                         double *row, double *norm_sum /* [n_contacts] scratch */) {
  /* cycontacts2data sensors.pyx:140-190 (+ store_forces :20-52, postprocess_contacts :113-137) */
  for (int s = 0; s < f->n_contacts; s++) norm_sum[s] = 0.0;
  for (int i = 0; i < d->ncon; i++) {
    const int c = d->con_cand[i];
    const double *frame = d->con_frame + 9*i, *pos = d->con_pos + 3*i;
    for (int key = 0; key < 4; key++) {
      const int sx = f->cand_sensor[4*c + key];
      if (sx < 0) continue;
      const double sign = (key & 1) ? 1.0 : -1.0;      /* (g1,g2) -1, (g2,g1) +1, (g1,-1) -1, (g2,-1) +1 */
      double ft[6], total[3];
      orc_contact_force(m, d, i, ft);                  /* mj_contactForce, sensors.pyx:70 */
      double *r = row + (size_t)sx*C_COLS;
      for (int k = 0; k < 3; k++) {
        const double reaction = sign*ft[0]*frame[k];
        const double friction = sign*ft[1]*frame[3 + k] + sign*ft[2]*frame[6 + k];
        total[k] = reaction + friction;
        r[C_REACTION + k] += reaction; r[C_FRICTION + k] += friction; r[C_TOTAL + k] += total[k];
      }
      const double norm = sqrt(total[0]*total[0] + total[1]*total[1] + total[2]*total[2]);
      for (int k = 0; k < 3; k++) r[C_POSITION + k] += norm*pos[k];
      norm_sum[sx] += norm;
    }
  }
  for (int s = 0; s < f->n_contacts; s++) {
    double *r = row + (size_t)s*C_COLS;
    if (norm_sum[s] > 0) for (int k = 0; k < 3; k++) r[C_POSITION + k] /= norm_sum[s];
    for (int k = 0; k < 9; k++) r[k] *= 1.0/u->newtons;
    for (int k = 0; k < 3; k++) r[C_POSITION + k] *= 1.0/u->meters;
  }
}

/* ---- swimming (farms_oracle.py: drag_forces, SwimmingHandlerOracle.step; drag.pyx:152-268,389-411) */
static void drag_forces(const double *link, double *xf, const double *coef /* [2][3] */, const FbFarms *f,
                        double mass, double height, double density, double gravity) {
  const double pos_z = link[L_COM_POS + 2], surface = f->water_sph ? 1e8 : f->water_surface;
  if (pos_z > surface) return;                                    /* :192-194, row untouched */
  double global2urdf[4], com2urdf[4], urdf2com[4], lin[3], ang[3], buoyancy[3] = {0, 0, 0}, wv[3];
  q_conj(link + L_URDF_QUAT, global2urdf);
  q_mult(global2urdf, link + L_COM_QUAT, com2urdf);
  q_conj(com2urdf, urdf2com);
  q_rot(link + L_LIN, global2urdf, lin);
  q_rot(link + L_ANG, global2urdf, ang);
  if (f->water_buoyancy && mass > 0 && pos_z < surface) {        /* compute_buoyancy :111-149 */
    double frac = (surface - pos_z > 0 ? surface - pos_z : 0)/height;
    if (frac > 1) frac = 1;
    const double lift[3] = {0.0, 0.0, -1000*mass*gravity/density*frac};
    q_rot(lift, global2urdf, buoyancy);
  }
  q_rot(f->water_velocity, global2urdf, wv);
  double force[3], torque[3];
  for (int k = 0; k < 3; k++) {
    const double v = lin[k] - wv[k], w = ang[k];
    const double sv = v > 0 ? 1.0 : (v < 0 ? -1.0 : 0.0), sw = w > 0 ? 1.0 : (w < 0 ? -1.0 : 0.0);
    force[k] = sv*v*v*f->water_viscosity*coef[k] + buoyancy[k];   /* compute_force :66-88 */
    torque[k] = sw*w*w*coef[3 + k];                               /* compute_torque :91-108 */
  }
  q_rot(force, urdf2com, xf);
  q_rot(torque, urdf2com, xf + 3);
}

static void swimming_step(const FbFarms *f, const double *links, double *xfrc) {
  if (!f->water_drag) return;
  for (int i = 0; i < f->n_swim; i++)
    drag_forces(links + (size_t)f->swim_links_index[i]*L_COLS, xfrc + (size_t)f->swim_xfrc_index[i]*6,
                f->swim_coefficients + 6*i, f, f->swim_mass[i], f->swim_height[i], f->swim_density[i], -9.81);
}

/* downstream swimming callback (farms_oracle.py: apply_xfrc) */
static void apply_xfrc(const FbModel *m, OrcData *d, const FbFarms *f, const FarmsUnits *u, const double *xfrc) {
  memset(d->xfrc_applied, 0, sizeof(double)*6*(size_t)m->nbody);
  for (int k = 0; k < f->n_xfrc; k++) {
    const int b = f->xfrc_body[k];
    const double *R = d->xmat + 9*b, *w = xfrc + 6*(size_t)k;
    for (int r = 0; r < 3; r++) {
      d->xfrc_applied[6*b + r] = (R[3*r]*w[0] + R[3*r + 1]*w[1] + R[3*r + 2]*w[2])*u->newtons;
      d->xfrc_applied[6*b + 3 + r] = (R[3*r]*w[3] + R[3*r + 1]*w[4] + R[3*r + 2]*w[5])*u->torques;
    }
  }
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9*ts.tv_nsec;
}

/* before_step of one iteration (task.py:168-186): sensors -> swimming callback.  Ring row
 * `iteration % buffer_size`; contacts / joints rows are zeroed first (the reference's `+=`
 * columns are only right on the first visit of a row; a ring needs the reset).
 * The xfrc row is NOT zeroed: drag_forces leaves it untouched above the surface (D-6). */
void orc_farms_sensors(const FbModel *m, OrcData *d, const FbFarms *f, long long iteration, int buffer_size,
                       double *links, double *joints, double *contacts, double *xfrc, double *norm_scratch,
                       int swimming, double *stage_s /* [4] or NULL: physics2data, swimming, control, mj_step */) {
  const FarmsUnits u = units_of(f);
  const size_t row = (size_t)(iteration % buffer_size);
  double *rl = links + row*f->n_links*L_COLS, *rj = joints + row*f->n_joints*f->joint_cols;
  double *rc = contacts + row*f->n_contacts*C_COLS, *rx = xfrc + row*f->n_xfrc*6;
  const double t0 = stage_s ? now_s() : 0.0;
  memset(rj, 0, sizeof(double)*(size_t)f->n_joints*f->joint_cols);
  memset(rc, 0, sizeof(double)*(size_t)f->n_contacts*C_COLS);
  links_row(d, f, &u, rl);
  joints_row(m, d, f, &u, rj);
  contacts_row(m, d, f, &u, rc, norm_scratch);
  const double t1 = stage_s ? now_s() : 0.0;
  if (swimming && f->n_swim > 0) {
    swimming_step(f, rl, rx);
    apply_xfrc(m, d, f, &u, rx);
  }
  if (stage_s) { stage_s[0] += t1 - t0; stage_s[1] += now_s() - t1; }
}

/* n_iterations x (before_step, travelling-wave control, mj_step), starting at iteration it0. */
void orc_farms_run(const FbModel *m, OrcData *d, const FbFarms *f, const FbWaveController *wc, double env_phase,
                   long long it0, int n_iterations, int buffer_size,
                   double *links, double *joints, double *contacts, double *xfrc, double *norm_scratch,
                   int swimming, double *stage_s) {
  for (int i = 0; i < n_iterations; i++) {
    const long long it = it0 + i;
    orc_farms_sensors(m, d, f, it, buffer_size, links, joints, contacts, xfrc, norm_scratch, swimming, stage_s);
    const double t2 = stage_s ? now_s() : 0.0;
    if (wc) {
      /* step_joints_control_position (task.py:309-321) with the controller of include/farms_b200.h */
      const double t = (double)it*m->timestep;
      for (int j = 0; j < wc->n; j++) {
        if (wc->actuator[j] < 0) continue;
        d->ctrl[wc->actuator[j]] = (wc->offset ? wc->offset[j] : 0.0)
            + wc->amplitude[j]*sin(2*M_PI*wc->frequency[j]*t - wc->phase_lag[j] + env_phase);
      }
    }
    const double t3 = stage_s ? now_s() : 0.0;
    orc_step(m, d);
    if (stage_s) { stage_s[2] += t3 - t2; stage_s[3] += now_s() - t3; }
  }
}
