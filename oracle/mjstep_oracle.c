/*
 * mjstep_oracle.c -- fp64 CPU restatement of the MuJoCo subset that the
 * FARMS MJCF exercises.  TEST INFRASTRUCTURE ONLY: nothing in the product
 * path (farms_mujoco_b200/) may link, import or call this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY UNPINNED: the algorithm lives in MuJoCo, a third-party dependency of
 * the reference that is absent from /root/reference and from this image
 * (requirements.txt:8 lists dm_control with no version; mjcf.py:1366 accepts
 * both sides of MuJoCo 3.2.3).  The reference has no tests or golden vectors
 * (SURVEY.md section 4).  This file restates MuJoCo's published algorithm
 * (engine_forward.c / engine_core_smooth.c / engine_core_constraint.c /
 * engine_collision_primitive.c, as summarised in SURVEY.md Appendix A) and is
 * pinned only by analytic known answers and self-consistency checks
 * (tests/test_oracle_*.py).
 *
 * Reference call sites this stands in for:
 *   farms_mujoco/simulation/simulation.py:53   Physics.from_mjcf_model
 *   farms_mujoco/simulation/simulation.py:156,175  env.step -> mj_step
 *   farms_mujoco/simulation/task.py:137        physics.reset -> mj_forward
 *   farms_mujoco/sensors/sensors.pyx:10,70     mj_contactForce
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../include/farms_b200.h"
#include "oracle.h"

#define MJ_MINVAL 1e-15
#define MJ_MINIMP 0.0001
#define MJ_MAXIMP 0.9999

/* ---------------------------------------------------------------- vec/quat */
static void cross3(double *r, const double *a, const double *b) {
  double x = a[1]*b[2] - a[2]*b[1], y = a[2]*b[0] - a[0]*b[2], z = a[0]*b[1] - a[1]*b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static double dot3(const double *a, const double *b) { return a[0]*b[0] + a[1]*b[1] + a[2]*b[2]; }
static double dot6(const double *a, const double *b) {
  return a[0]*b[0] + a[1]*b[1] + a[2]*b[2] + a[3]*b[3] + a[4]*b[4] + a[5]*b[5];
}
static double normalize3(double *v) {
  double n = sqrt(dot3(v, v));
  if (n < MJ_MINVAL) { v[0] = 1; v[1] = 0; v[2] = 0; return n; }
  v[0] /= n; v[1] /= n; v[2] /= n; return n;
}
static void normalize4(double *q) {
  double n = sqrt(q[0]*q[0] + q[1]*q[1] + q[2]*q[2] + q[3]*q[3]);
  if (n < MJ_MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
static void mulquat(double *r, const double *a, const double *b) {
  double w = a[0]*b[0] - a[1]*b[1] - a[2]*b[2] - a[3]*b[3];
  double x = a[0]*b[1] + a[1]*b[0] + a[2]*b[3] - a[3]*b[2];
  double y = a[0]*b[2] - a[1]*b[3] + a[2]*b[0] + a[3]*b[1];
  double z = a[0]*b[3] + a[1]*b[2] - a[2]*b[1] + a[3]*b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static void quat2mat(double *m, const double *q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w*w + x*x - y*y - z*z; m[1] = 2*(x*y - w*z);         m[2] = 2*(x*z + w*y);
  m[3] = 2*(x*y + w*z);         m[4] = w*w - x*x + y*y - z*z; m[5] = 2*(y*z - w*x);
  m[6] = 2*(x*z - w*y);         m[7] = 2*(y*z + w*x);         m[8] = w*w - x*x - y*y + z*z;
}
static void mulmatvec3(double *r, const double *m, const double *v) {
  double x = m[0]*v[0] + m[1]*v[1] + m[2]*v[2];
  double y = m[3]*v[0] + m[4]*v[1] + m[5]*v[2];
  double z = m[6]*v[0] + m[7]*v[1] + m[8]*v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static void axisangle2quat(double *q, const double *axis, double angle) {
  double s = sin(0.5*angle);
  q[0] = cos(0.5*angle); q[1] = axis[0]*s; q[2] = axis[1]*s; q[3] = axis[2]*s;
}

/* ------------------------------------------------------ spatial (c-frame) */
/* mju_mulInertVec: 10-number inertia times [ang; lin] */
static void mul_inert_vec(double *r, const double *i, const double *v) {
  r[0] = i[0]*v[0] + i[3]*v[1] + i[4]*v[2] - i[8]*v[4] + i[7]*v[5];
  r[1] = i[3]*v[0] + i[1]*v[1] + i[5]*v[2] + i[8]*v[3] - i[6]*v[5];
  r[2] = i[4]*v[0] + i[5]*v[1] + i[2]*v[2] - i[7]*v[3] + i[6]*v[4];
  r[3] = i[8]*v[1] - i[7]*v[2] + i[9]*v[3];
  r[4] = i[6]*v[2] - i[8]*v[0] + i[9]*v[4];
  r[5] = i[7]*v[0] - i[6]*v[1] + i[9]*v[5];
}
/* mju_crossMotion: vel x_m v */
static void cross_motion(double *r, const double *vel, const double *v) {
  r[0] = -vel[2]*v[1] + vel[1]*v[2];
  r[1] =  vel[2]*v[0] - vel[0]*v[2];
  r[2] = -vel[1]*v[0] + vel[0]*v[1];
  r[3] = -vel[2]*v[4] + vel[1]*v[5] - vel[5]*v[1] + vel[4]*v[2];
  r[4] =  vel[2]*v[3] - vel[0]*v[5] + vel[5]*v[0] - vel[3]*v[2];
  r[5] = -vel[1]*v[3] + vel[0]*v[4] - vel[4]*v[0] + vel[3]*v[1];
}
/* mju_crossForce: vel x* f */
static void cross_force(double *r, const double *vel, const double *f) {
  r[0] = -vel[2]*f[1] + vel[1]*f[2] - vel[5]*f[4] + vel[4]*f[5];
  r[1] =  vel[2]*f[0] - vel[0]*f[2] + vel[5]*f[3] - vel[3]*f[5];
  r[2] = -vel[1]*f[0] + vel[0]*f[1] - vel[4]*f[3] + vel[3]*f[4];
  r[3] = -vel[2]*f[4] + vel[1]*f[5];
  r[4] =  vel[2]*f[3] - vel[0]*f[5];
  r[5] = -vel[1]*f[3] + vel[0]*f[4];
}

/* --------------------------------------------------------------- A.1 + A.2 */
static int body_rootid(const FbModel *m, int b) {
  while (b > 0 && m->body_parentid[b] > 0) b = m->body_parentid[b];
  return b;
}

void orc_kinematics(const FbModel *m, OrcData *d) {
  memset(d->xpos, 0, 3*sizeof(double));
  d->xquat[0] = 1; d->xquat[1] = d->xquat[2] = d->xquat[3] = 0;
  quat2mat(d->xmat, d->xquat);
  memset(d->xipos, 0, 3*sizeof(double));
  quat2mat(d->ximat, d->xquat);
  for (int b = 1; b < m->nbody; b++) {
    int p = m->body_parentid[b], jid = m->body_jntid[b];
    double pos[3], quat[4], tmp[3];
    if (jid >= 0 && m->jnt_type[jid] == FB_JNT_FREE) {
      int adr = m->jnt_qposadr[jid];
      normalize4(d->qpos + adr + 3);           /* mj_kinematics normalises qpos in place */
      memcpy(pos, d->qpos + adr, 3*sizeof(double));
      memcpy(quat, d->qpos + adr + 3, 4*sizeof(double));
      memcpy(d->xanchor + 3*jid, pos, 3*sizeof(double));
      d->xaxis[3*jid] = 0; d->xaxis[3*jid+1] = 0; d->xaxis[3*jid+2] = 1;
    } else {
      mulmatvec3(tmp, d->xmat + 9*p, m->body_pos + 3*b);
      for (int k = 0; k < 3; k++) pos[k] = d->xpos[3*p+k] + tmp[k];
      mulquat(quat, d->xquat + 4*p, m->body_quat + 4*b);
      if (jid >= 0) {
        int adr = m->jnt_qposadr[jid];
        double mat[9];
        quat2mat(mat, quat);
        mulmatvec3(tmp, mat, m->jnt_pos + 3*jid);
        for (int k = 0; k < 3; k++) d->xanchor[3*jid+k] = pos[k] + tmp[k];
        mulmatvec3(d->xaxis + 3*jid, mat, m->jnt_axis + 3*jid);
        if (m->jnt_type[jid] == FB_JNT_HINGE) {
          double qloc[4], qnew[4];
          axisangle2quat(qloc, m->jnt_axis + 3*jid, d->qpos[adr] - m->qpos0[adr]);
          mulquat(qnew, quat, qloc);
          memcpy(quat, qnew, sizeof(qnew));
          quat2mat(mat, quat);
          mulmatvec3(tmp, mat, m->jnt_pos + 3*jid);
          for (int k = 0; k < 3; k++) pos[k] = d->xanchor[3*jid+k] - tmp[k];
        } else { /* slide */
          for (int k = 0; k < 3; k++) pos[k] += d->xaxis[3*jid+k]*(d->qpos[adr] - m->qpos0[adr]);
        }
      }
    }
    normalize4(quat);
    memcpy(d->xpos + 3*b, pos, sizeof(pos));
    memcpy(d->xquat + 4*b, quat, sizeof(quat));
    quat2mat(d->xmat + 9*b, quat);
    mulmatvec3(tmp, d->xmat + 9*b, m->body_ipos + 3*b);
    for (int k = 0; k < 3; k++) d->xipos[3*b+k] = pos[k] + tmp[k];
    double iq[4];
    mulquat(iq, quat, m->body_iquat + 4*b);
    quat2mat(d->ximat + 9*b, iq);
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_bodyid[g];
    double tmp[3], q[4];
    mulmatvec3(tmp, d->xmat + 9*b, m->geom_pos + 3*g);
    for (int k = 0; k < 3; k++) d->geom_xpos[3*g+k] = d->xpos[3*b+k] + tmp[k];
    mulquat(q, d->xquat + 4*b, m->geom_quat + 4*g);
    quat2mat(d->geom_xmat + 9*g, q);
  }
}

void orc_com_pos(const FbModel *m, OrcData *d) {
  int nb = m->nbody;
  /* subtree_com: mass-weighted, children first */
  double *mass = (double *)calloc(nb, sizeof(double));
  for (int b = 0; b < nb; b++) {
    mass[b] = m->body_mass[b];
    for (int k = 0; k < 3; k++) d->subtree_com[3*b+k] = m->body_mass[b]*d->xipos[3*b+k];
  }
  for (int b = nb - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    mass[p] += mass[b];
    for (int k = 0; k < 3; k++) d->subtree_com[3*p+k] += d->subtree_com[3*b+k];
  }
  for (int b = 0; b < nb; b++) {
    if (mass[b] < MJ_MINVAL) {
      memcpy(d->subtree_com + 3*b, d->xipos + 3*b, 3*sizeof(double));
    } else {
      for (int k = 0; k < 3; k++) d->subtree_com[3*b+k] /= mass[b];
    }
  }
  free(mass);
  /* cinert: mju_inertCom about subtree_com[root] in world axes */
  memset(d->cinert, 0, 10*sizeof(double));
  for (int b = 1; b < nb; b++) {
    const double *com = d->subtree_com + 3*body_rootid(m, b);
    const double *mat = d->ximat + 9*b, *in = m->body_inertia + 3*b;
    double dif[3], mb = m->body_mass[b], *r = d->cinert + 10*b;
    for (int k = 0; k < 3; k++) dif[k] = d->xipos[3*b+k] - com[k];
    /* mat * diag(in) * mat' */
    r[0] = mat[0]*mat[0]*in[0] + mat[1]*mat[1]*in[1] + mat[2]*mat[2]*in[2];
    r[1] = mat[3]*mat[3]*in[0] + mat[4]*mat[4]*in[1] + mat[5]*mat[5]*in[2];
    r[2] = mat[6]*mat[6]*in[0] + mat[7]*mat[7]*in[1] + mat[8]*mat[8]*in[2];
    r[3] = mat[0]*mat[3]*in[0] + mat[1]*mat[4]*in[1] + mat[2]*mat[5]*in[2];
    r[4] = mat[0]*mat[6]*in[0] + mat[1]*mat[7]*in[1] + mat[2]*mat[8]*in[2];
    r[5] = mat[3]*mat[6]*in[0] + mat[4]*mat[7]*in[1] + mat[5]*mat[8]*in[2];
    r[0] += mb*(dif[1]*dif[1] + dif[2]*dif[2]);
    r[1] += mb*(dif[0]*dif[0] + dif[2]*dif[2]);
    r[2] += mb*(dif[0]*dif[0] + dif[1]*dif[1]);
    r[3] -= mb*dif[0]*dif[1];
    r[4] -= mb*dif[0]*dif[2];
    r[5] -= mb*dif[1]*dif[2];
    r[6] = mb*dif[0]; r[7] = mb*dif[1]; r[8] = mb*dif[2]; r[9] = mb;
  }
  /* cdof */
  for (int j = 0; j < m->njnt; j++) {
    int b = m->jnt_bodyid[j], da = m->jnt_dofadr[j];
    const double *com = d->subtree_com + 3*body_rootid(m, b);
    double off[3];
    for (int k = 0; k < 3; k++) off[k] = com[k] - d->xanchor[3*j+k];
    if (m->jnt_type[j] == FB_JNT_FREE) {
      for (int i = 0; i < 3; i++) {
        double *c = d->cdof + 6*(da + i);
        memset(c, 0, 6*sizeof(double));
        c[3+i] = 1;
      }
      for (int i = 0; i < 3; i++) {
        double *c = d->cdof + 6*(da + 3 + i);
        double ax[3] = { d->xmat[9*b+i], d->xmat[9*b+3+i], d->xmat[9*b+6+i] };
        memcpy(c, ax, sizeof(ax));
        cross3(c + 3, ax, off);
      }
    } else if (m->jnt_type[j] == FB_JNT_HINGE) {
      double *c = d->cdof + 6*da;
      memcpy(c, d->xaxis + 3*j, 3*sizeof(double));
      cross3(c + 3, d->xaxis + 3*j, off);
    } else {
      double *c = d->cdof + 6*da;
      c[0] = c[1] = c[2] = 0;
      memcpy(c + 3, d->xaxis + 3*j, 3*sizeof(double));
    }
  }
}

/* -------------------------------------------------------------------- A.3 */
void orc_crb(const FbModel *m, OrcData *d) {
  int nb = m->nbody, nv = m->nv;
  memcpy(d->crb, d->cinert, 10*nb*sizeof(double));
  for (int b = nb - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    if (p > 0) for (int k = 0; k < 10; k++) d->crb[10*p+k] += d->crb[10*b+k];
  }
  memset(d->qM, 0, m->nM*sizeof(double));
  for (int i = 0; i < nv; i++) {
    double buf[6];
    int adr = m->dof_Madr[i];
    d->qM[adr] = m->dof_armature[i];
    mul_inert_vec(buf, d->crb + 10*m->dof_bodyid[i], d->cdof + 6*i);
    for (int j = i; j >= 0; j = m->dof_parentid[j]) d->qM[adr++] += dot6(d->cdof + 6*j, buf);
  }
}

/* in-place sparse L'DL over dof_parentid chains (mj_factorI) */
void orc_factor(const FbModel *m, double *qLD, double *qLDiagInv) {
  int nv = m->nv;
  for (int k = nv - 1; k >= 0; k--) {
    int Madr_kk = m->dof_Madr[k], Madr_ki = Madr_kk + 1, i = m->dof_parentid[k];
    while (i >= 0) {
      double tmp = qLD[Madr_ki]/qLD[Madr_kk];
      int cnt = 0;
      for (int j = i; j >= 0; j = m->dof_parentid[j]) cnt++;
      for (int t = 0; t < cnt; t++) qLD[m->dof_Madr[i] + t] -= tmp*qLD[Madr_ki + t];
      qLD[Madr_ki] = tmp;
      i = m->dof_parentid[i];
      Madr_ki++;
    }
  }
  for (int k = 0; k < nv; k++) qLDiagInv[k] = 1.0/qLD[m->dof_Madr[k]];
}

/* x <- inv(L'DL) x (mj_solveLD) */
void orc_solve(const FbModel *m, const double *qLD, const double *qLDiagInv, double *x) {
  int nv = m->nv;
  for (int k = nv - 1; k >= 0; k--) {
    int adr = m->dof_Madr[k] + 1;
    for (int i = m->dof_parentid[k]; i >= 0; i = m->dof_parentid[i]) x[i] -= qLD[adr++]*x[k];
  }
  for (int k = 0; k < nv; k++) x[k] *= qLDiagInv[k];
  for (int k = 0; k < nv; k++) {
    int adr = m->dof_Madr[k] + 1;
    for (int i = m->dof_parentid[k]; i >= 0; i = m->dof_parentid[i]) x[k] -= qLD[adr++]*x[i];
  }
}

/* res = M * vec with the sparse qM (mj_mulM) */
void orc_mulM(const FbModel *m, const double *qM, double *res, const double *vec) {
  int nv = m->nv;
  for (int i = 0; i < nv; i++) {
    int adr = m->dof_Madr[i];
    res[i] = qM[adr]*vec[i];
    int a = adr + 1;
    for (int j = m->dof_parentid[i]; j >= 0; j = m->dof_parentid[j], a++) res[i] += qM[a]*vec[j];
  }
  for (int i = 0; i < nv; i++) {
    int a = m->dof_Madr[i] + 1;
    for (int j = m->dof_parentid[i]; j >= 0; j = m->dof_parentid[j], a++) res[j] += qM[a]*vec[i];
  }
}

/* dense copy of the sparse mass matrix, row-major nv x nv (tests) */
void orc_fullM(const FbModel *m, const double *qM, double *dst) {
  int nv = m->nv;
  memset(dst, 0, (size_t)nv*nv*sizeof(double));
  for (int i = 0; i < nv; i++) {
    int adr = m->dof_Madr[i];
    for (int j = i; j >= 0; j = m->dof_parentid[j], adr++) {
      dst[i*nv + j] = qM[adr];
      dst[j*nv + i] = qM[adr];
    }
  }
}

/* point Jacobian (mj_jac): jacp/jacr are 3 x nv row-major */
static void jac_point(const FbModel *m, const OrcData *d, double *jacp, double *jacr,
                      const double *point, int body) {
  int nv = m->nv;
  if (jacp) memset(jacp, 0, 3*nv*sizeof(double));
  if (jacr) memset(jacr, 0, 3*nv*sizeof(double));
  const double *com = d->subtree_com + 3*body_rootid(m, body);
  double off[3] = { point[0] - com[0], point[1] - com[1], point[2] - com[2] };
  /* last dof of the nearest ancestor (or self) that has dofs */
  int b = body;
  while (b > 0 && m->body_dofnum[b] == 0) b = m->body_parentid[b];
  if (b == 0) return;
  int i = m->body_dofadr[b] + m->body_dofnum[b] - 1;
  for (; i >= 0; i = m->dof_parentid[i]) {
    const double *c = d->cdof + 6*i;
    if (jacr) { jacr[i] = c[0]; jacr[nv+i] = c[1]; jacr[2*nv+i] = c[2]; }
    if (jacp) {
      double t[3];
      cross3(t, c, off);
      jacp[i] = c[3] + t[0]; jacp[nv+i] = c[4] + t[1]; jacp[2*nv+i] = c[5] + t[2];
    }
  }
}

/* -------------------------------------------------------------------- A.6 */
static void make_frame(double *f) {
  normalize3(f);
  if (sqrt(dot3(f+3, f+3)) < 0.5) {
    f[3] = f[4] = f[5] = 0;
    if (f[1] < 0.5 && f[1] > -0.5) f[4] = 1; else f[5] = 1;
  }
  double t = dot3(f, f+3);
  for (int k = 0; k < 3; k++) f[3+k] -= t*f[k];
  normalize3(f+3);
  cross3(f+6, f, f+3);
}

/* mjc_PlaneCylinder: point k (0: the lowest rim point, 1: the rim point straight above it at the
 * other end, 2 / 3: the two points that make a triangle with it on the lower rim) of a cylinder
 * (axis, x axis, radius, half length; centre at distance dist0 from the plane along n).  Returns
 * the point's distance from the plane; p = the point relative to the centre. */
static double plane_cylinder_point(const double *n, const double *axis_in, const double *xaxis, double radius,
                                   double half, int k, double dist0, double *p) {
  double axis[3] = { axis_in[0], axis_in[1], axis_in[2] }, vec[3], vec1[3];
  double prjaxis = dot3(n, axis);
  if (prjaxis > 0) { for (int i = 0; i < 3; i++) axis[i] = -axis[i]; prjaxis = -prjaxis; }
  for (int i = 0; i < 3; i++) vec[i] = axis[i]*prjaxis - n[i];       /* -normal without its part along the axis */
  double len = sqrt(dot3(vec, vec));
  if (len >= MJ_MINVAL) { for (int i = 0; i < 3; i++) vec[i] *= radius/len; }
  else { for (int i = 0; i < 3; i++) vec[i] = xaxis[i]*radius; }     /* disk parallel to the plane */
  double prjvec = dot3(vec, n);
  for (int i = 0; i < 3; i++) axis[i] *= half;
  prjaxis *= half;
  if (k == 0) { for (int i = 0; i < 3; i++) p[i] = vec[i] + axis[i]; return dist0 + prjaxis + prjvec; }
  if (k == 1) { for (int i = 0; i < 3; i++) p[i] = vec[i] - axis[i]; return dist0 - prjaxis + prjvec; }
  cross3(vec1, vec, axis);
  normalize3(vec1);
  for (int i = 0; i < 3; i++) p[i] = (k == 2 ? 1.0 : -1.0)*vec1[i]*radius*sqrt(3.0)/2 + axis[i] - 0.5*vec[i];
  return dist0 + prjaxis - 0.5*prjvec;
}

void orc_collision(const FbModel *m, OrcData *d) {
  d->ncon = 0;
  int box_geom = -1, box_count = 0;
  for (int c = 0; c < m->ncand; c++) {
    int g1 = m->cand_geom1[c], g2 = m->cand_geom2[c], end = m->cand_end[c];
    const double *pm = d->geom_xmat + 9*g1, *pp = d->geom_xpos + 3*g1;
    const double *gm = d->geom_xmat + 9*g2, *gp = d->geom_xpos + 3*g2;
    double n[3] = { pm[2], pm[5], pm[8] };
    if (end == 20) {
      /* sphere-sphere of an explicit <pair> (mjc_SphereSphere): normal from geom1 to geom2, the
       * point half way between the two surfaces */
      double r1 = m->geom_size[3*g1], r2 = m->geom_size[3*g2];
      double dif[3] = { gp[0]-pp[0], gp[1]-pp[1], gp[2]-pp[2] };
      double cdist = normalize3(dif);
      if (cdist > m->cand_margin[c] + r1 + r2) continue;
      int i = d->ncon++;
      double sdist = cdist - r1 - r2;
      d->con_cand[i] = c;
      d->con_dist[i] = sdist;
      for (int k = 0; k < 3; k++) d->con_pos[3*i+k] = pp[k] + dif[k]*(r1 + 0.5*sdist);
      double *f = d->con_frame + 9*i;
      memcpy(f, dif, sizeof(dif));
      f[3] = f[4] = f[5] = 0;
      make_frame(f);
      d->con_efc_address[i] = -1;
      continue;
    }
    if (end >= 11) {
      /* plane-cylinder (mjc_PlaneCylinder): up to four contacts, each inside the margin on its own
       * (the first is the lowest, so MuJoCo's early return never drops another) */
      const double axis[3] = { gm[2], gm[5], gm[8] }, xaxis[3] = { gm[0], gm[3], gm[6] };
      double dif[3] = { gp[0]-pp[0], gp[1]-pp[1], gp[2]-pp[2] }, p[3];
      double cdist = plane_cylinder_point(n, axis, xaxis, m->geom_size[3*g2], m->geom_size[3*g2+1], end - 11,
                                          dot3(dif, n), p);
      if (cdist > m->cand_margin[c]) continue;
      int i = d->ncon++;
      d->con_cand[i] = c;
      d->con_dist[i] = cdist;
      for (int k = 0; k < 3; k++) d->con_pos[3*i+k] = gp[k] + p[k] - n[k]*0.5*cdist;
      double *f = d->con_frame + 9*i;
      memcpy(f, n, sizeof(n));
      f[3] = f[4] = f[5] = 0;
      make_frame(f);
      d->con_efc_address[i] = -1;
      continue;
    }
    if (end == 10) {
      /* plane-ellipsoid (mjc_PlaneConvex): the support point of the ellipsoid along -normal,
       * centre + R (s o normalize(s o R'(-n))) (mjc_support, ellipsoid case); one contact while the
       * point is inside the margin */
      const double *sz = m->geom_size + 3*g2;
      double w[3], sup[3];
      for (int k = 0; k < 3; k++) w[k] = -(gm[k]*n[0] + gm[3+k]*n[1] + gm[6+k]*n[2])*sz[k];
      normalize3(w);
      for (int k = 0; k < 3; k++) w[k] *= sz[k];
      for (int k = 0; k < 3; k++) sup[k] = gp[k] + gm[3*k]*w[0] + gm[3*k+1]*w[1] + gm[3*k+2]*w[2];
      double dif[3] = { sup[0]-pp[0], sup[1]-pp[1], sup[2]-pp[2] };
      double sdist = dot3(dif, n);
      if (sdist > m->cand_margin[c]) continue;
      int i = d->ncon++;
      d->con_cand[i] = c;
      d->con_dist[i] = sdist;
      for (int k = 0; k < 3; k++) d->con_pos[3*i+k] = sup[k] - n[k]*0.5*sdist;
      double *f = d->con_frame + 9*i;
      memcpy(f, n, sizeof(n));
      f[3] = f[4] = f[5] = 0;
      make_frame(f);
      d->con_efc_address[i] = -1;
      continue;
    }
    if (end >= 2) {
      /* plane-box (mjc_PlaneBox): corner end-2 (bit 0: x, 1: y, 2: z) of the half-sizes; kept when
       * it is below the box centre along the normal and inside the margin; at most 4 per pair */
      if (g2 != box_geom || (c > 0 && m->cand_geom1[c-1] != g1)) { box_geom = g2; box_count = 0; }
      int i8 = end - 2;
      double vec[3] = { (i8 & 1 ? 1 : -1)*m->geom_size[3*g2], (i8 & 2 ? 1 : -1)*m->geom_size[3*g2+1],
                        (i8 & 4 ? 1 : -1)*m->geom_size[3*g2+2] };
      double corner[3];
      for (int k = 0; k < 3; k++) corner[k] = gm[3*k]*vec[0] + gm[3*k+1]*vec[1] + gm[3*k+2]*vec[2];
      double dif[3] = { gp[0]-pp[0], gp[1]-pp[1], gp[2]-pp[2] };
      double ldist = dot3(n, corner), bdist = dot3(dif, n) + ldist;
      if (bdist > m->cand_margin[c] || ldist > 0 || box_count >= 4) continue;
      box_count++;
      int i = d->ncon++;
      d->con_cand[i] = c;
      d->con_dist[i] = bdist;
      for (int k = 0; k < 3; k++) d->con_pos[3*i+k] = gp[k] + corner[k] - n[k]*0.5*bdist;
      double *f = d->con_frame + 9*i;
      memcpy(f, n, sizeof(n));
      f[3] = f[4] = f[5] = 0;
      make_frame(f);
      d->con_efc_address[i] = -1;
      continue;
    }
    double axis[3] = { gm[2], gm[5], gm[8] };
    double centre[3], radius = m->geom_size[3*g2];
    double margin = m->cand_margin[c];
    for (int k = 0; k < 3; k++) centre[k] = gp[k] + end*m->geom_size[3*g2+1]*axis[k];
    double tmp[3] = { centre[0]-pp[0], centre[1]-pp[1], centre[2]-pp[2] };
    double cdist = dot3(tmp, n);
    if (cdist > margin + radius) continue;
    int i = d->ncon++;
    double dist = cdist - radius;
    d->con_cand[i] = c;
    d->con_dist[i] = dist;
    for (int k = 0; k < 3; k++) d->con_pos[3*i+k] = centre[k] - n[k]*(radius + 0.5*dist);
    double *f = d->con_frame + 9*i;
    memcpy(f, n, sizeof(n));
    if (end != 0) memcpy(f+3, axis, sizeof(axis)); else f[3] = f[4] = f[5] = 0;
    make_frame(f);
    d->con_efc_address[i] = -1;
  }
}

/* -------------------------------------------------------------------- A.7 */
static double get_impedance(const double *solimp, double pos_minus_margin) {
  double dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  dmin = fmin(MJ_MAXIMP, fmax(MJ_MINIMP, dmin));
  dmax = fmin(MJ_MAXIMP, fmax(MJ_MINIMP, dmax));
  width = fmax(MJ_MINVAL, width);
  mid = fmin(MJ_MAXIMP, fmax(MJ_MINIMP, mid));
  power = fmax(1.0, power);
  if (dmin == dmax || width <= MJ_MINVAL) return 0.5*(dmin + dmax);
  double x = fabs(pos_minus_margin)/width, y;
  if (x >= 1) return dmax;
  if (x <= 0) return dmin;
  if (power == 1) y = x;
  else if (x <= mid) y = pow(x, power)/pow(mid, power - 1);
  else y = 1 - pow(1 - x, power)/pow(1 - mid, power - 1);
  return dmin + y*(dmax - dmin);
}

static void add_row(const FbModel *m, OrcData *d, const double *jac, double pos, double margin,
                    double diag_approx, const double *solref_in, const double *solimp,
                    int type, int id) {
  int nv = m->nv, r = d->nefc++;
  memcpy(d->efc_J + (size_t)r*nv, jac, nv*sizeof(double));
  d->efc_pos[r] = pos;
  d->efc_margin[r] = margin;
  d->efc_type[r] = type;
  d->efc_id[r] = id;
  double solref[2] = { solref_in[0], solref_in[1] };
  if (solref[0] > 0) solref[0] = fmax(solref[0], 2*m->timestep);   /* refsafe */
  double dmax = fmin(MJ_MAXIMP, fmax(MJ_MINIMP, solimp[1]));
  double imp = get_impedance(solimp, pos - margin);
  d->efc_R[r] = fmax(MJ_MINVAL, (1 - imp)*diag_approx/imp);
  double K, B;
  if (solref[0] > 0) {
    K = 1/fmax(MJ_MINVAL, dmax*dmax*solref[0]*solref[0]*solref[1]*solref[1]);
    B = 2/fmax(MJ_MINVAL, dmax*solref[0]);
  } else {
    K = -solref[0]/fmax(MJ_MINVAL, dmax*dmax);
    B = -solref[1]/fmax(MJ_MINVAL, dmax);
  }
  d->efc_KBIP[4*r] = K; d->efc_KBIP[4*r+1] = B; d->efc_KBIP[4*r+2] = imp; d->efc_KBIP[4*r+3] = 0;
}

void orc_make_constraint(const FbModel *m, OrcData *d) {
  int nv = m->nv;
  d->nefc = 0;
  double *jac = (double *)calloc(nv, sizeof(double));
  double *jacp = (double *)calloc(3*nv, sizeof(double));
  double *jacp1 = (double *)calloc(3*nv, sizeof(double));
  for (int j = 0; j < m->njnt; j++) d->jnt_limit_row[j] = -1;
  /* joint limits (mj_instantiateLimit) */
  for (int j = 0; j < m->njnt; j++) {
    if (!m->jnt_limited[j] || m->jnt_type[j] == FB_JNT_FREE) continue;
    double value = d->qpos[m->jnt_qposadr[j]], margin = m->jnt_margin[j];
    for (int side = -1; side <= 1; side += 2) {
      double dist = side*(m->jnt_range[2*j + (side+1)/2] - value);
      if (dist < margin) {
        memset(jac, 0, nv*sizeof(double));
        jac[m->jnt_dofadr[j]] = -side;
        if (d->jnt_limit_row[j] < 0) d->jnt_limit_row[j] = d->nefc;
        add_row(m, d, jac, dist, margin, m->dof_invweight0[m->jnt_dofadr[j]],
                m->jnt_solref + 2*j, m->jnt_solimp + 5*j, ORC_CNSTR_LIMIT, j);
      }
    }
  }
  /* contacts, pyramidal condim 3 (mj_instantiateContact) */
  for (int i = 0; i < d->ncon; i++) {
    int c = d->con_cand[i];
    double includemargin = m->cand_margin[c] - m->cand_gap[c];
    if (d->con_dist[i] >= includemargin) { d->con_efc_address[i] = -1; continue; }
    int b2 = m->geom_bodyid[m->cand_geom2[c]], b1 = m->geom_bodyid[m->cand_geom1[c]];
    double mu = m->cand_friction[c];
    jac_point(m, d, jacp, NULL, d->con_pos + 3*i, b2);   /* plane candidates: body1 is the world, zero */
    if (b1 > 0) {
      /* explicit pair between two bodies of the tree (mj_jacDifPair): jac2 - jac1 */
      jac_point(m, d, jacp1, NULL, d->con_pos + 3*i, b1);
      for (int v = 0; v < 3*nv; v++) jacp[v] -= jacp1[v];
    }
    const double *f = d->con_frame + 9*i;
    double tran = m->body_invweight0[2*b1] + m->body_invweight0[2*b2];
    d->con_efc_address[i] = d->nefc;
    int first = d->nefc;
    for (int k = 1; k < 3; k++) {
      for (int sgn = 1; sgn >= -1; sgn -= 2) {
        for (int v = 0; v < nv; v++) {
          double jn = f[0]*jacp[v] + f[1]*jacp[nv+v] + f[2]*jacp[2*nv+v];
          double jt = f[3*k]*jacp[v] + f[3*k+1]*jacp[nv+v] + f[3*k+2]*jacp[2*nv+v];
          jac[v] = jn + sgn*mu*jt;
        }
        add_row(m, d, jac, d->con_dist[i], includemargin, tran + mu*mu*tran,
                m->cand_solref + 2*c, m->cand_solimp + 5*c, ORC_CNSTR_CONTACT_PYRAMIDAL, i);
      }
    }
    /* mj_makeImpedance: R[1] = R[0]/impratio, contact.mu = friction*sqrt(R[1]/R[0]);
     * pyramidal edges all get Rpy = 2 contact.mu^2 R[1] */
    double impratio = fmax(MJ_MINVAL, m->impratio);
    double R1 = d->efc_R[first]/impratio;
    double cmu = mu*sqrt(R1/d->efc_R[first]);
    double Rpy = fmax(MJ_MINVAL, 2*cmu*cmu*R1);
    for (int k = 0; k < 4; k++) d->efc_R[first+k] = Rpy;
  }
  for (int r = 0; r < d->nefc; r++) d->efc_D[r] = 1/d->efc_R[r];
  free(jac);
  free(jacp);
  free(jacp1);
}

/* -------------------------------------------------------------------- A.4 */
void orc_com_vel(const FbModel *m, OrcData *d) {
  memset(d->cvel, 0, 6*sizeof(double));
  for (int b = 1; b < m->nbody; b++) {
    double cvel[6], tmp[6];
    memcpy(cvel, d->cvel + 6*m->body_parentid[b], sizeof(cvel));
    int jid = m->body_jntid[b];
    if (jid >= 0) {
      int da = m->jnt_dofadr[jid];
      if (m->jnt_type[jid] == FB_JNT_FREE) {
        memset(d->cdof_dot + 6*da, 0, 18*sizeof(double));
        for (int i = 0; i < 3; i++) for (int k = 0; k < 6; k++) cvel[k] += d->cdof[6*(da+i)+k]*d->qvel[da+i];
        for (int i = 3; i < 6; i++) cross_motion(d->cdof_dot + 6*(da+i), cvel, d->cdof + 6*(da+i));
        for (int i = 3; i < 6; i++) for (int k = 0; k < 6; k++) cvel[k] += d->cdof[6*(da+i)+k]*d->qvel[da+i];
      } else {
        cross_motion(tmp, cvel, d->cdof + 6*da);
        memcpy(d->cdof_dot + 6*da, tmp, sizeof(tmp));
        for (int k = 0; k < 6; k++) cvel[k] += d->cdof[6*da+k]*d->qvel[da];
      }
    }
    memcpy(d->cvel + 6*b, cvel, sizeof(cvel));
  }
}

/* RNE bias forces, flg_acc = 0 */
void orc_rne(const FbModel *m, OrcData *d, double *result) {
  int nb = m->nbody, nv = m->nv;
  double *cacc = (double *)calloc(6*nb, sizeof(double));
  double *cfrc = (double *)calloc(6*nb, sizeof(double));
  for (int k = 0; k < 3; k++) cacc[3+k] = -m->gravity[k];
  for (int b = 1; b < nb; b++) {
    int p = m->body_parentid[b], da = m->body_dofadr[b];
    double tmp[6], tmp1[6], tmp2[6];
    memcpy(cacc + 6*b, cacc + 6*p, 6*sizeof(double));
    for (int i = 0; i < m->body_dofnum[b]; i++)
      for (int k = 0; k < 6; k++) cacc[6*b+k] += d->cdof_dot[6*(da+i)+k]*d->qvel[da+i];
    mul_inert_vec(tmp, d->cinert + 10*b, cacc + 6*b);
    mul_inert_vec(tmp1, d->cinert + 10*b, d->cvel + 6*b);
    cross_force(tmp2, d->cvel + 6*b, tmp1);
    for (int k = 0; k < 6; k++) cfrc[6*b+k] = tmp[k] + tmp2[k];
  }
  for (int b = nb - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    if (p > 0) for (int k = 0; k < 6; k++) cfrc[6*p+k] += cfrc[6*b+k];
  }
  for (int i = 0; i < nv; i++) result[i] = dot6(d->cdof + 6*i, cfrc + 6*m->dof_bodyid[i]);
  free(cacc);
  free(cfrc);
}

/* -------------------------------------------------------------------- A.5 */
void orc_passive(const FbModel *m, OrcData *d) {
  memset(d->qfrc_passive, 0, m->nv*sizeof(double));
  for (int j = 0; j < m->njnt; j++) {
    if (m->jnt_type[j] == FB_JNT_FREE) continue;
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    d->qfrc_passive[da] -= m->jnt_stiffness[j]*(d->qpos[qa] - d->qpos_spring[qa]);
  }
  for (int i = 0; i < m->nv; i++) d->qfrc_passive[i] -= m->dof_damping[i]*d->qvel[i];
}

void orc_actuation(const FbModel *m, OrcData *d) {
  memset(d->qfrc_actuator, 0, m->nv*sizeof(double));
  for (int a = 0; a < m->nu; a++) {
    int j = m->actuator_trnid[a];
    double gear = m->actuator_gear[a];
    double length = gear*d->qpos[m->jnt_qposadr[j]], velocity = gear*d->qvel[m->jnt_dofadr[j]];
    double ctrl = d->ctrl[a];
    if (m->actuator_ctrllimited[a])
      ctrl = fmin(m->actuator_ctrlrange[2*a+1], fmax(m->actuator_ctrlrange[2*a], ctrl));
    const double *g = m->actuator_gainprm + 3*a, *bp = m->actuator_biasprm + 3*a;
    double force = g[0]*ctrl + bp[0] + bp[1]*length + bp[2]*velocity;
    if (m->actuator_forcelimited[a])
      force = fmin(m->actuator_forcerange[2*a+1], fmax(m->actuator_forcerange[2*a], force));
    d->actuator_force[a] = force;
    d->qfrc_actuator[m->jnt_dofadr[j]] += gear*force;
  }
}

/* xfrc_applied (world force, torque at xipos) -> generalised (mj_xfrcAccumulate) */
static void xfrc_accumulate(const FbModel *m, OrcData *d, double *qfrc) {
  int nv = m->nv;
  double *jacp = (double *)calloc(3*nv, sizeof(double));
  double *jacr = (double *)calloc(3*nv, sizeof(double));
  for (int b = 1; b < m->nbody; b++) {
    const double *x = d->xfrc_applied + 6*b;
    if (x[0] == 0 && x[1] == 0 && x[2] == 0 && x[3] == 0 && x[4] == 0 && x[5] == 0) continue;
    jac_point(m, d, jacp, jacr, d->xipos + 3*b, b);
    for (int v = 0; v < nv; v++)
      for (int k = 0; k < 3; k++) qfrc[v] += jacp[k*nv+v]*x[k] + jacr[k*nv+v]*x[3+k];
  }
  free(jacp);
  free(jacr);
}

/* -------------------------------------------------------------------- A.8 */
typedef struct { double alpha; int row; } Breakpoint;
static int cmp_break(const void *a, const void *b) {
  double x = ((const Breakpoint *)a)->alpha, y = ((const Breakpoint *)b)->alpha;
  return (x > y) - (x < y);
}

/* dense Cholesky solve H x = b (H overwritten); returns 0 on success */
static int chol_solve(double *H, double *x, int n) {
  for (int j = 0; j < n; j++) {
    double s = H[j*n+j];
    for (int k = 0; k < j; k++) s -= H[j*n+k]*H[j*n+k];
    if (s <= 0) return -1;
    H[j*n+j] = sqrt(s);
    for (int i = j+1; i < n; i++) {
      double t = H[i*n+j];
      for (int k = 0; k < j; k++) t -= H[i*n+k]*H[j*n+k];
      H[i*n+j] = t/H[j*n+j];
    }
  }
  for (int i = 0; i < n; i++) {
    double t = x[i];
    for (int k = 0; k < i; k++) t -= H[i*n+k]*x[k];
    x[i] = t/H[i*n+i];
  }
  for (int i = n-1; i >= 0; i--) {
    double t = x[i];
    for (int k = i+1; k < n; k++) t -= H[k*n+i]*x[k];
    x[i] = t/H[i*n+i];
  }
  return 0;
}

/* Primal Newton with exact line search on the convex piecewise-quadratic cost
 *   1/2 (a-a0)' M (a-a0) + sum_i 1/2 D_i min(0, J_i a - aref_i)^2
 * (limit and pyramidal rows share s_i).  The minimiser
 * is unique, so any convergent method gives MuJoCo's answer. */
void orc_solve_constraints(const FbModel *m, OrcData *d) {
  int nv = m->nv, ne = d->nefc;
  memcpy(d->qacc, d->qacc_smooth, nv*sizeof(double));
  memset(d->qfrc_constraint, 0, nv*sizeof(double));
  d->solver_niter = 0;
  if (ne == 0) return;
  double *r = (double *)calloc(ne, sizeof(double)), *jp = (double *)calloc(ne, sizeof(double));
  double *grad = (double *)calloc(nv, sizeof(double)), *Ma = (double *)calloc(nv, sizeof(double));
  double *p = (double *)calloc(nv, sizeof(double)), *Mp = (double *)calloc(nv, sizeof(double));
  double *H = (double *)calloc((size_t)nv*nv, sizeof(double));
  Breakpoint *bp = (Breakpoint *)calloc(ne, sizeof(Breakpoint));
  double scale = 1.0/(m->meaninertia*(nv > 1 ? nv : 1));
  for (int iter = 0; iter < 200; iter++) {
    d->solver_niter = iter + 1;
    for (int i = 0; i < ne; i++) {
      double s = -d->efc_aref[i];
      for (int v = 0; v < nv; v++) s += d->efc_J[(size_t)i*nv+v]*d->qacc[v];
      r[i] = s;
    }
    orc_mulM(m, d->qM, Ma, d->qacc);
    double gnorm = 0;
    for (int v = 0; v < nv; v++) grad[v] = Ma[v] - d->qfrc_smooth[v];
    for (int i = 0; i < ne; i++) if (r[i] < 0)
      for (int v = 0; v < nv; v++) grad[v] += d->efc_J[(size_t)i*nv+v]*d->efc_D[i]*r[i];
    for (int v = 0; v < nv; v++) gnorm += grad[v]*grad[v];
    {
      double ref = 0;
      for (int v = 0; v < nv; v++) ref += Ma[v]*Ma[v] + d->qfrc_smooth[v]*d->qfrc_smooth[v];
      if (sqrt(gnorm) <= 1e-13*sqrt(ref) || scale*sqrt(gnorm) < 1e-300) break;
    }
    orc_fullM(m, d->qM, H);
    for (int i = 0; i < ne; i++) if (r[i] < 0) {
      const double *J = d->efc_J + (size_t)i*nv;
      for (int a = 0; a < nv; a++) if (J[a] != 0)
        for (int b = 0; b < nv; b++) H[a*nv+b] += d->efc_D[i]*J[a]*J[b];
    }
    for (int v = 0; v < nv; v++) p[v] = -grad[v];
    if (chol_solve(H, p, nv)) break;
    /* exact line search: phi'(alpha) = g0 + alpha*pMp + sum D_i jp_i min(0, r_i + alpha jp_i) */
    orc_mulM(m, d->qM, Mp, p);
    double pMp = 0, g0 = 0;
    for (int v = 0; v < nv; v++) { pMp += p[v]*Mp[v]; g0 += p[v]*(Ma[v] - d->qfrc_smooth[v]); }
    double c0 = g0, c1 = pMp;   /* phi'(alpha) = c0 + c1*alpha on the current segment */
    int nb = 0;
    for (int i = 0; i < ne; i++) {
      double s = 0;
      for (int v = 0; v < nv; v++) s += d->efc_J[(size_t)i*nv+v]*p[v];
      jp[i] = s;
      /* active on (0, eps): r<0, or r==0 and decreasing */
      if (r[i] < 0 || (r[i] == 0 && s < 0)) { c0 += d->efc_D[i]*s*r[i]; c1 += d->efc_D[i]*s*s; }
      if (s != 0) {
        double a = -r[i]/s;
        if (a > 0) { bp[nb].alpha = a; bp[nb].row = i; nb++; }
      }
    }
    qsort(bp, nb, sizeof(Breakpoint), cmp_break);
    double alpha = 0;
    int k = 0;
    for (;;) {
      double next = (k < nb) ? bp[k].alpha : INFINITY;
      double root = (c1 > 0) ? -c0/c1 : INFINITY;
      if (root <= next) { alpha = root; break; }
      if (k >= nb) { alpha = next; break; }
      int i = bp[k].row;
      /* crossing r_i + alpha jp_i = 0 toggles the row: active after the
       * breakpoint iff jp_i < 0 (residual decreasing through 0) */
      double sgn = (jp[i] < 0) ? 1.0 : -1.0;
      c0 += sgn*d->efc_D[i]*jp[i]*r[i];
      c1 += sgn*d->efc_D[i]*jp[i]*jp[i];
      k++;
    }
    if (!isfinite(alpha)) alpha = 1;
    double step2 = 0, a2 = 0;
    for (int v = 0; v < nv; v++) {
      d->qacc[v] += alpha*p[v];
      step2 += alpha*alpha*p[v]*p[v];
      a2 += d->qacc[v]*d->qacc[v];
    }
    if (step2 <= 1e-30*(a2 + 1e-300)) break;
  }
  for (int i = 0; i < ne; i++) {
    double s = -d->efc_aref[i];
    for (int v = 0; v < nv; v++) s += d->efc_J[(size_t)i*nv+v]*d->qacc[v];
    d->efc_force[i] = (s < 0) ? -d->efc_D[i]*s : 0;
    for (int v = 0; v < nv; v++) d->qfrc_constraint[v] += d->efc_J[(size_t)i*nv+v]*d->efc_force[i];
  }
  free(r); free(jp); free(grad); free(Ma); free(p); free(Mp); free(H); free(bp);
}

/* ------------------------------------------------------------------- A.11 */
void orc_contact_force(const FbModel *m, const OrcData *d, int i, double *result) {
  /* mj_contactForce, pyramidal condim 3 */
  memset(result, 0, 6*sizeof(double));
  int adr = d->con_efc_address[i];
  if (adr < 0) return;
  double mu = m->cand_friction[d->con_cand[i]];
  const double *f = d->efc_force + adr;
  result[0] = f[0] + f[1] + f[2] + f[3];
  result[1] = (f[0] - f[1])*mu;
  result[2] = (f[2] - f[3])*mu;
}

static void sensors(const FbModel *m, OrcData *d) {
  for (int b = 0; b < m->nbody; b++) {
    const double *cv = d->cvel + 6*b, *com = d->subtree_com + 3*body_rootid(m, b);
    double off[3] = { d->xipos[3*b]-com[0], d->xipos[3*b+1]-com[1], d->xipos[3*b+2]-com[2] }, t[3];
    cross3(t, cv, off);
    for (int k = 0; k < 3; k++) { d->body_angvel[3*b+k] = cv[k]; d->body_linvel[3*b+k] = cv[3+k] + t[k]; }
  }
  for (int j = 0; j < m->njnt; j++)
    d->jnt_limit_force[j] = d->jnt_limit_row[j] >= 0 ? d->efc_force[d->jnt_limit_row[j]] : 0.0;
  for (int i = 0; i < d->ncon; i++) orc_contact_force(m, d, i, d->con_force + 6*i);
}

/* ------------------------------------------------------------------- A.0 */
void orc_forward(const FbModel *m, OrcData *d) {
  int nv = m->nv;
  orc_kinematics(m, d);
  orc_com_pos(m, d);
  orc_crb(m, d);
  memcpy(d->qLD, d->qM, m->nM*sizeof(double));
  orc_factor(m, d->qLD, d->qLDiagInv);
  orc_collision(m, d);
  orc_make_constraint(m, d);
  orc_com_vel(m, d);
  orc_passive(m, d);
  /* referenceConstraint: aref = -B*(J qvel) - K*imp*(pos - margin) */
  for (int i = 0; i < d->nefc; i++) {
    double vel = 0;
    for (int v = 0; v < nv; v++) vel += d->efc_J[(size_t)i*nv+v]*d->qvel[v];
    d->efc_aref[i] = -d->efc_KBIP[4*i+1]*vel
                     - d->efc_KBIP[4*i]*d->efc_KBIP[4*i+2]*(d->efc_pos[i] - d->efc_margin[i]);
  }
  orc_rne(m, d, d->qfrc_bias);
  orc_actuation(m, d);
  for (int v = 0; v < nv; v++)
    d->qfrc_smooth[v] = d->qfrc_passive[v] - d->qfrc_bias[v] + d->qfrc_actuator[v];
  xfrc_accumulate(m, d, d->qfrc_smooth);
  memcpy(d->qacc_smooth, d->qfrc_smooth, nv*sizeof(double));
  orc_solve(m, d->qLD, d->qLDiagInv, d->qacc_smooth);
  orc_solve_constraints(m, d);
  sensors(m, d);
}

/* A.10 semi-implicit Euler with implicit joint damping */
void orc_euler(const FbModel *m, OrcData *d) {
  int nv = m->nv;
  double h = m->timestep;
  int damped = 0;
  for (int i = 0; i < nv; i++) if (m->dof_damping[i] > 0) damped = 1;
  double *qacc = (double *)calloc(nv, sizeof(double));
  if (damped) {
    double *MhB = (double *)calloc(m->nM, sizeof(double)), *inv = (double *)calloc(nv, sizeof(double));
    memcpy(MhB, d->qM, m->nM*sizeof(double));
    for (int i = 0; i < nv; i++) MhB[m->dof_Madr[i]] += h*m->dof_damping[i];
    orc_factor(m, MhB, inv);
    for (int i = 0; i < nv; i++) qacc[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
    orc_solve(m, MhB, inv, qacc);
    free(MhB); free(inv);
  } else {
    memcpy(qacc, d->qacc, nv*sizeof(double));
  }
  for (int i = 0; i < nv; i++) d->qvel[i] += h*qacc[i];
  for (int j = 0; j < m->njnt; j++) {
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == FB_JNT_FREE) {
      for (int k = 0; k < 3; k++) d->qpos[qa+k] += h*d->qvel[da+k];
      double w[3] = { d->qvel[da+3], d->qvel[da+4], d->qvel[da+5] }, qrot[4], qn[4];
      double angle = h*normalize3(w);
      axisangle2quat(qrot, w, angle);
      normalize4(d->qpos + qa + 3);
      mulquat(qn, d->qpos + qa + 3, qrot);
      normalize4(qn);
      memcpy(d->qpos + qa + 3, qn, sizeof(qn));
    } else {
      d->qpos[qa] += h*d->qvel[da];
    }
  }
  d->time += h;
  free(qacc);
}

void orc_step(const FbModel *m, OrcData *d) {
  orc_forward(m, d);
  orc_euler(m, d);
}

void orc_step_n(const FbModel *m, OrcData *d, int n) {
  for (int i = 0; i < n; i++) orc_step(m, d);
}
