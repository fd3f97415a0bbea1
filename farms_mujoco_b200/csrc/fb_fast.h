/*
 * fb_fast.h -- the environment-per-thread step: ONE CUDA thread advances one
 * environment while no joint limit or contact is active.
 *
 * Every environment runs the same model, so a warp of 32 environments executes
 * the same tree walk with the same (uniform) indices: no divergence, no
 * shuffles, no barriers, every lane busy.  Forward dynamics is the
 * articulated-body recursion (three sweeps over the bodies), which needs no
 * mass matrix and therefore a per-environment working set small enough to keep
 * 64 environments per SM resident in shared memory.  Implicit joint damping of
 * the Euler integrator, (M + h D) x = f, is the same recursion
 * with h*damping added to the joint-space diagonal, so the result is the x the
 * team path (fb_device.h: CRB + L'DL, as MuJoCo does it) computes, up to fp32
 * rounding.
 *
 * Spatial vectors are [angular; linear] in WORLD axes, taken about the body's
 * own reference point (its joint anchor); moving a quantity from a body to its
 * parent is a pure translation by r = anchor(body) - anchor(parent), built from
 * the parent rotation and model constants (never from differences of world
 * positions), which keeps every number at link scale in fp32.
 *
 * As soon as pass 1 of a step sees a joint limit violated or a body inside the
 * conservative contact bound of a plane, the environment stops here with its
 * state untouched and is appended to the pending list; the team kernel
 * (fb_device.h) takes it over from that step with the full constraint solver.
 *
 * Shared-memory layout: element i of a per-environment array lives at
 * (off + i)*BLK + thread (DevFastLayout), i.e. consecutive lanes hit consecutive
 * banks.
 *
 * Reference anchors: step order mj_step via farms_mujoco/simulation/
 * simulation.py:156; log rows farms_mujoco/simulation/physics.py:435-524;
 * drag/buoyancy farms_mujoco/swimming/drag.pyx:152-268,389-411; hook order
 * farms_mujoco/simulation/task.py:168-186.
 */
#ifndef FB_FAST_H_
#define FB_FAST_H_

#include "fb_device.h"

#ifdef FB_HOST_EMU
FB_DEV void fb_st4(float *p, float a, float b, float c, float d) { p[0] = a; p[1] = b; p[2] = c; p[3] = d; }
FB_DEV void fb_st2(float *p, float a, float b) { p[0] = a; p[1] = b; }
#else
FB_DEV void fb_st4(float *p, float a, float b, float c, float d) {
  *reinterpret_cast<float4 *>(p) = make_float4(a, b, c, d);
}
FB_DEV void fb_st2(float *p, float a, float b) { *reinterpret_cast<float2 *>(p) = make_float2(a, b); }
#endif

/* articulated inertia about a point, world axes:  [ A  H ] [w]   A, M symmetric
 *                                                 [ H' M ] [v]   (xx yy zz xy xz yz) */
struct ArtInertia {
  float A[6], H[9], M[6];
};

FB_DEV void sym_mul(const float *S, const float *x, float *y) {
  y[0] = S[0]*x[0] + S[3]*x[1] + S[4]*x[2];
  y[1] = S[3]*x[0] + S[1]*x[1] + S[5]*x[2];
  y[2] = S[4]*x[0] + S[5]*x[1] + S[2]*x[2];
}
/* y = H x, y = H' x */
FB_DEV void h_mul(const float *H, const float *x, float *y) {
  y[0] = H[0]*x[0] + H[1]*x[1] + H[2]*x[2];
  y[1] = H[3]*x[0] + H[4]*x[1] + H[5]*x[2];
  y[2] = H[6]*x[0] + H[7]*x[1] + H[8]*x[2];
}
FB_DEV void ht_mul(const float *H, const float *x, float *y) {
  y[0] = H[0]*x[0] + H[3]*x[1] + H[6]*x[2];
  y[1] = H[1]*x[0] + H[4]*x[1] + H[7]*x[2];
  y[2] = H[2]*x[0] + H[5]*x[1] + H[8]*x[2];
}
/* S -= k * a a' (symmetric rank one) */
FB_DEV void sym_rank1(float *S, const float *a, float k) {
  float b0 = k*a[0], b1 = k*a[1], b2 = k*a[2];
  S[0] -= b0*a[0]; S[1] -= b1*a[1]; S[2] -= b2*a[2];
  S[3] -= b0*a[1]; S[4] -= b0*a[2]; S[5] -= b1*a[2];
}

/* Move an articulated inertia and a force from a point to the point -r away
 * (child anchor -> parent anchor, r = child - parent):
 *   H' = H + r~ M,  A' = A - H r~ + r~ H''  ,  n' = n + r x f          */
FB_DEV void art_shift(ArtInertia &I, float *p, const float *r) {
  float Mc[3][3] = {{I.M[0], I.M[3], I.M[4]}, {I.M[3], I.M[1], I.M[5]}, {I.M[4], I.M[5], I.M[2]}};
  float Hn[9];
FB_UNROLL
  for (int j = 0; j < 3; j++) {
    float c[3];
    v_cross(r, Mc[j], c);                       /* column j of r~ M */
    Hn[j] = I.H[j] + c[0]; Hn[3 + j] = I.H[3 + j] + c[1]; Hn[6 + j] = I.H[6 + j] + c[2];
  }
  float T[9];                                   /* T = -H r~ + r~ Hn' */
FB_UNROLL
  for (int i = 0; i < 3; i++) {
    float hr[3];
    v_cross(I.H + 3*i, r, hr);                  /* row i of H r~ */
    T[3*i] = -hr[0]; T[3*i+1] = -hr[1]; T[3*i+2] = -hr[2];
  }
FB_UNROLL
  for (int j = 0; j < 3; j++) {
    float rh[3];
    v_cross(r, Hn + 3*j, rh);                   /* column j of r~ Hn' */
    T[j] += rh[0]; T[3 + j] += rh[1]; T[6 + j] += rh[2];
  }
  I.A[0] += T[0]; I.A[1] += T[4]; I.A[2] += T[8];
  I.A[3] += 0.5f*(T[1] + T[3]); I.A[4] += 0.5f*(T[2] + T[6]); I.A[5] += 0.5f*(T[5] + T[7]);
FB_UNROLL
  for (int k = 0; k < 9; k++) I.H[k] = Hn[k];
  float c[3];
  v_cross(r, p + 3, c);
  p[0] += c[0]; p[1] += c[1]; p[2] += c[2];
}

/* solve the symmetric positive definite 6x6 system K x = b in registers (L D L') */
FB_DEV void solve6(float K[6][6], const float *b, float *x) {
  float dinv[6];
FB_UNROLL
  for (int j = 0; j < 6; j++) {
    float d = K[j][j];
FB_UNROLL
    for (int k = 0; k < j; k++) d -= K[j][k]*K[j][k]*K[k][k];
    K[j][j] = d;
    dinv[j] = 1.0f/d;
FB_UNROLL
    for (int i = j + 1; i < 6; i++) {
      float v = K[i][j];
FB_UNROLL
      for (int k = 0; k < j; k++) v -= K[i][k]*K[j][k]*K[k][k];
      K[i][j] = v*dinv[j];
    }
  }
  float y[6];
FB_UNROLL
  for (int i = 0; i < 6; i++) {
    float v = b[i];
FB_UNROLL
    for (int k = 0; k < i; k++) v -= K[i][k]*y[k];
    y[i] = v;
  }
FB_UNROLL
  for (int i = 5; i >= 0; i--) {
    float v = y[i]*dinv[i];
FB_UNROLL
    for (int k = i + 1; k < 6; k++) v -= K[k][i]*x[k];
    x[i] = v;
  }
}

template <int BLK> struct FbFast {
  const FbParams &P;
  const DevModel &m;
  float *s;           /* shared floats, already offset by the thread index */
  const int env;
  float env_phase;
  const float *g_ctrl, *g_spring;

  FB_MEM FbFast(const FbParams &P_, float *s_, int env_)
      : P(P_), m(P_.m), s(s_), env(env_) {
    env_phase = P.env_phase[env];
    g_ctrl = P.ctrl + (size_t)env*(m.nu > 0 ? m.nu : 1);
    g_spring = P.qpos_spring + (size_t)env*m.nq;
  }

  FB_MEM float &S(int off, int i) const { return s[(size_t)(off + i)*BLK]; }

  FB_MEM Quat quat_of(int b) const {
    Quat q = {S(m.X.quat, 4*b), S(m.X.quat, 4*b+1), S(m.X.quat, 4*b+2), S(m.X.quat, 4*b+3)};
    return q;
  }

  /* r = anchor(b) - anchor(parent) in world axes.  qp/Rp: orientation of the parent
   * (identity for the world), Rb: rotation of b, dq: joint coordinate minus qpos0.
   * The anchor is fixed in the frame of b BEFORE its joint rotation (MuJoCo xanchor). */
  FB_MEM void anchor_offset(int b, int jtype, int jid, Quat qp, const float *Rp, const float *Rb,
                            float dq, float *r) const {
    m_rot(Rp, MF(ft_dpos, 3*b), MF(ft_dpos, 3*b+1), MF(ft_dpos, 3*b+2), r);
    if (jid >= 0 && jtype != FB_JNT_FREE) {
      if (m.X.any_jpos) {
        Quat bq = {MF(body_quat, 4*b), MF(body_quat, 4*b+1), MF(body_quat, 4*b+2), MF(body_quat, 4*b+3)};
        float Rpre[9], t[3];
        q_mat(q_normalize(q_mul(qp, bq)), Rpre);
        m_rot(Rpre, MF(jnt_pos, 3*jid), MF(jnt_pos, 3*jid+1), MF(jnt_pos, 3*jid+2), t);
        r[0] += t[0]; r[1] += t[1]; r[2] += t[2];
      }
      if (jtype == FB_JNT_SLIDE) {
        float ax[3];
        m_rot(Rb, MF(jnt_axis, 3*jid), MF(jnt_axis, 3*jid+1), MF(jnt_axis, 3*jid+2), ax);
        r[0] += ax[0]*dq; r[1] += ax[1]*dq; r[2] += ax[2]*dq;
      }
    }
  }

  /* sum of gear*force of the joint's actuators and the farms joint_torque sum */
  FB_MEM void actuation(int jid, int fj, float q, float qd, float time, int store_ctrl,
                        float *tau, float *trq_log) const {
    int a0 = MI(jnt_actstart, jid), a1 = MI(jnt_actstart, jid + 1);
    int ap = -1, av = -1, at = -1;
    if (fj >= 0) { ap = MI(fj_actpos, fj); av = MI(fj_actvel, fj); at = MI(fj_acttrq, fj); }
    float tsum = 0.f, lsum = 0.f;
    for (int t = a0; t < a1; t++) {
      int a = MI(act_sorted, t);
      float gear = MF(act_gear, a);
      int w = MI(ft_actwc, a);
      float c;
      if (w >= 0) {
        float ph = 6.283185307179586f*MF(wc_freq, w)*time - MF(wc_lag, w) + env_phase;
        c = MF(wc_off, w) + MF(wc_amp, w)*sinf(ph);
        if (store_ctrl) P.ctrl[(size_t)env*m.nu + a] = c;
      } else {
        c = g_ctrl[a];
      }
      if (MI(act_ctrllimited, a)) c = fminf(MF(act_ctrlrange, 2*a+1), fmaxf(MF(act_ctrlrange, 2*a), c));
      float f = MF(act_gain, a)*c + MF(act_bias, 3*a) + MF(act_bias, 3*a+1)*(gear*q) + MF(act_bias, 3*a+2)*(gear*qd);
      if (MI(act_forcelimited, a)) f = fminf(MF(act_forcerange, 2*a+1), fmaxf(MF(act_forcerange, 2*a), f));
      tsum += gear*f;
      if (a == ap || a == av || a == at) lsum += f;
    }
    *tau = tsum;
    *trq_log = lsum;
  }

  FB_MEM void load_state() {
    const float *gq = P.qpos + (size_t)env*m.nq, *gv = P.qvel + (size_t)env*m.nv;
    const float *gx = P.xfrc_applied + (size_t)env*6*m.nbody;
    for (int i = 0; i < m.nq; i++) S(m.X.qpos, i) = gq[i];
    for (int i = 0; i < m.nv; i++) S(m.X.qvel, i) = gv[i];
    for (int i = 0; i < 6*m.nbody; i++) S(m.X.wrench, i) = gx[i];
  }

  FB_MEM void store_state(long long iteration) {
    float *gq = P.qpos + (size_t)env*m.nq, *gv = P.qvel + (size_t)env*m.nv;
    float *gx = P.xfrc_applied + (size_t)env*6*m.nbody;
    for (int i = 0; i < m.nq; i++) gq[i] = S(m.X.qpos, i);
    for (int i = 0; i < m.nv; i++) gv[i] = S(m.X.qvel, i);
    for (int i = 0; i < 6*m.nbody; i++) gx[i] = S(m.X.wrench, i);
    P.iteration[env] = iteration;
  }

  /* ---- pass 1: poses, velocities, links rows, constraint detection.  Returns 1
   * when a limit or a plane bound is active (the step must not be taken here). */
  FB_MEM int pass_poses(float *row_links) {
    const int nb = m.nbody;
    int active = 0;
    for (int b = 1; b < nb; b++) {
      const int p = MI(body_parent, b), jid = MI(body_jnt, b);
      const int jtype = jid >= 0 ? MI(jnt_type, jid) : -1;
      Quat qp = {1.f, 0.f, 0.f, 0.f};
      float op[3] = {0.f, 0.f, 0.f}, vp[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (p > 0) {
        qp = quat_of(p);
FB_UNROLL
        for (int k = 0; k < 3; k++) op[k] = S(m.X.org, 3*p + k);
FB_UNROLL
        for (int k = 0; k < 6; k++) vp[k] = S(m.X.vel, 6*p + k);
      }
      Quat q;
      float o[3], xpos[3], v[6], R[9];
      if (jtype == FB_JNT_FREE) {
        const int qa = MI(jnt_qposadr, jid), da = MI(jnt_dofadr, jid);
        Quat qq = {S(m.X.qpos, qa+3), S(m.X.qpos, qa+4), S(m.X.qpos, qa+5), S(m.X.qpos, qa+6)};
        q = q_normalize(qq);
        S(m.X.qpos, qa+3) = q.w; S(m.X.qpos, qa+4) = q.x; S(m.X.qpos, qa+5) = q.y; S(m.X.qpos, qa+6) = q.z;
        q_mat(q, R);
FB_UNROLL
        for (int k = 0; k < 3; k++) { o[k] = S(m.X.qpos, qa + k); xpos[k] = o[k]; v[3 + k] = S(m.X.qvel, da + k); }
        m_rot(R, S(m.X.qvel, da+3), S(m.X.qvel, da+4), S(m.X.qvel, da+5), v);
      } else {
        float Rp[9], r[3], dq = 0.f, qd = 0.f;
        q_mat(qp, Rp);
        Quat bq = {MF(body_quat, 4*b), MF(body_quat, 4*b+1), MF(body_quat, 4*b+2), MF(body_quat, 4*b+3)};
        q = q_mul(qp, bq);
        float ja[3] = {0.f, 0.f, 0.f};
        if (jid >= 0) {
          const int qa = MI(jnt_qposadr, jid), da = MI(jnt_dofadr, jid);
          const float qj = S(m.X.qpos, qa);
          qd = S(m.X.qvel, da);
          dq = qj - MF(jnt_qpos0, jid);
          ja[0] = MF(jnt_axis, 3*jid); ja[1] = MF(jnt_axis, 3*jid+1); ja[2] = MF(jnt_axis, 3*jid+2);
          if (jtype == FB_JNT_HINGE) {
            float sn, cs;
            fb_sincos(0.5f*dq, &sn, &cs);
            Quat ql = {cs, ja[0]*sn, ja[1]*sn, ja[2]*sn};
            q = q_mul(q, ql);
          }
          if (MI(jnt_limited, jid)) {
            const float margin = MF(jnt_margin, jid);
            if (qj - MF(jnt_range, 2*jid) < margin || MF(jnt_range, 2*jid+1) - qj < margin) active = 1;
          }
        }
        q = q_normalize(q);
        q_mat(q, R);
        anchor_offset(b, jtype, jid, qp, Rp, R, dq, r);
        float ax[3] = {0.f, 0.f, 0.f}, cr[3];
        if (jid >= 0) m_rot(R, ja[0], ja[1], ja[2], ax);
        v_cross(vp, r, cr);                      /* w_parent x r */
FB_UNROLL
        for (int k = 0; k < 3; k++) {
          o[k] = op[k] + r[k];
          v[k] = vp[k] + (jtype == FB_JNT_HINGE ? ax[k]*qd : 0.f);
          v[3 + k] = vp[3 + k] + cr[k] + (jtype == FB_JNT_SLIDE ? ax[k]*qd : 0.f);
          xpos[k] = o[k];
        }
        if (jid >= 0 && m.X.any_jpos) {
          float t[3];
          m_rot(R, MF(jnt_pos, 3*jid), MF(jnt_pos, 3*jid+1), MF(jnt_pos, 3*jid+2), t);
          xpos[0] -= t[0]; xpos[1] -= t[1]; xpos[2] -= t[2];
        }
      }
      S(m.X.quat, 4*b) = q.w; S(m.X.quat, 4*b+1) = q.x; S(m.X.quat, 4*b+2) = q.y; S(m.X.quat, 4*b+3) = q.z;
FB_UNROLL
      for (int k = 0; k < 3; k++) S(m.X.org, 3*b + k) = o[k];
FB_UNROLL
      for (int k = 0; k < 6; k++) S(m.X.vel, 6*b + k) = v[k];
      /* conservative plane bound */
      for (int t = MI(ft_chkstart, b); t < MI(ft_chkstart, b + 1); t++)
        if (MF(ft_chk, 4*t)*xpos[0] + MF(ft_chk, 4*t+1)*xpos[1] + MF(ft_chk, 4*t+2)*xpos[2] < MF(ft_chk, 4*t+3)) active = 1;
      /* links row: physics.py:449-466 + :435-446 */
      const int l = MI(ft_link, b);
      if (l >= 0) {
        float h[3], cr[3];
        m_rot(R, MF(ft_hloc, 3*b), MF(ft_hloc, 3*b+1), MF(ft_hloc, 3*b+2), h);   /* com - anchor */
        v_cross(v, h, cr);
        const float im = m.inv_meters, iv = m.inv_velocity, iw = m.inv_angvel;
        float *row = row_links + 20*l;
        fb_st4(row, (o[0] + h[0])*im, (o[1] + h[1])*im, (o[2] + h[2])*im, q.x);
        fb_st4(row + 4, q.y, q.z, q.w, xpos[0]*im);
        fb_st4(row + 8, xpos[1]*im, xpos[2]*im, q.x, q.y);
        fb_st4(row + 12, q.z, q.w, (v[3] + cr[0])*iv, (v[4] + cr[1])*iv);
        fb_st4(row + 16, (v[5] + cr[2])*iv, v[0]*iw, v[1]*iw, v[2]*iw);
      }
    }
    return active;
  }

  /* ---- pass 2: leaves -> root, articulated inertias and bias forces */
  FB_MEM void pass_inertia(float time, float *aroot, int store_ctrl) {
    const int nb = m.nbody;
    const float hdt = m.timestep;
    ArtInertia C;        /* carry from child b+1 */
    float pc[6];
FB_UNROLL
    for (int k = 0; k < 6; k++) { C.A[k] = 0.f; C.M[k] = 0.f; pc[k] = 0.f; }
FB_UNROLL
    for (int k = 0; k < 9; k++) C.H[k] = 0.f;
    for (int b = nb - 1; b >= 1; b--) {
      const int p = MI(body_parent, b), jid = MI(body_jnt, b), flags = MI(ft_flags, b);
      const int jtype = jid >= 0 ? MI(jnt_type, jid) : -1;
      const Quat q = quat_of(b);
      float R[9], v[6], fx[6];
      q_mat(q, R);
FB_UNROLL
      for (int k = 0; k < 6; k++) { v[k] = S(m.X.vel, 6*b + k); fx[k] = S(m.X.wrench, 6*b + k); }
      /* rigid-body inertia about the anchor */
      const float mass = MF(body_mass, b);
      float h[3], Ib[6], Iw[6];
      m_rot(R, MF(ft_hloc, 3*b), MF(ft_hloc, 3*b+1), MF(ft_hloc, 3*b+2), h);
FB_UNROLL
      for (int k = 0; k < 6; k++) Ib[k] = MF(ft_inertia, 6*b + k);
      {
        /* Iw = R Ib R' */
        float T[9];
FB_UNROLL
        for (int i = 0; i < 3; i++) {
          float ri[3] = {R[3*i], R[3*i+1], R[3*i+2]}, t[3];
          sym_mul(Ib, ri, t);
          T[3*i] = t[0]; T[3*i+1] = t[1]; T[3*i+2] = t[2];       /* row i of R Ib */
        }
        Iw[0] = T[0]*R[0] + T[1]*R[1] + T[2]*R[2];
        Iw[1] = T[3]*R[3] + T[4]*R[4] + T[5]*R[5];
        Iw[2] = T[6]*R[6] + T[7]*R[7] + T[8]*R[8];
        Iw[3] = T[0]*R[3] + T[1]*R[4] + T[2]*R[5];
        Iw[4] = T[0]*R[6] + T[1]*R[7] + T[2]*R[8];
        Iw[5] = T[3]*R[6] + T[4]*R[7] + T[5]*R[8];
      }
      ArtInertia I;
      const float hh = h[0]*h[0] + h[1]*h[1] + h[2]*h[2];
      I.A[0] = Iw[0] + mass*(hh - h[0]*h[0]); I.A[1] = Iw[1] + mass*(hh - h[1]*h[1]);
      I.A[2] = Iw[2] + mass*(hh - h[2]*h[2]);
      I.A[3] = Iw[3] - mass*h[0]*h[1]; I.A[4] = Iw[4] - mass*h[0]*h[2]; I.A[5] = Iw[5] - mass*h[1]*h[2];
      I.H[0] = 0.f; I.H[1] = -mass*h[2]; I.H[2] = mass*h[1];
      I.H[3] = mass*h[2]; I.H[4] = 0.f; I.H[5] = -mass*h[0];
      I.H[6] = -mass*h[1]; I.H[7] = mass*h[0]; I.H[8] = 0.f;
      I.M[0] = mass; I.M[1] = mass; I.M[2] = mass; I.M[3] = 0.f; I.M[4] = 0.f; I.M[5] = 0.f;
      /* bias force  v x* (I v) - applied wrench (force F, torque T at the com) */
      float pA[6];
      {
        float wxh[3], mom[3], L[3], t[3], c1[3], c2[3];
        v_cross(v, h, wxh);
FB_UNROLL
        for (int k = 0; k < 3; k++) mom[k] = mass*(v[3 + k] + wxh[k]);
        sym_mul(Iw, v, L);
        v_cross(h, mom, t);
FB_UNROLL
        for (int k = 0; k < 3; k++) L[k] += t[k];
        v_cross(v, L, c1);
        v_cross(v + 3, mom, c2);
        v_cross(h, fx, t);                        /* h x F */
FB_UNROLL
        for (int k = 0; k < 3; k++) pA[k] = c1[k] + c2[k] - (fx[3 + k] + t[k]);
        v_cross(v, mom, c1);
FB_UNROLL
        for (int k = 0; k < 3; k++) pA[3 + k] = c1[k] - fx[k];
      }
      if (flags & FT_ADD_CARRY) {
FB_UNROLL
        for (int k = 0; k < 6; k++) { I.A[k] += C.A[k]; I.M[k] += C.M[k]; pA[k] += pc[k]; }
FB_UNROLL
        for (int k = 0; k < 9; k++) I.H[k] += C.H[k];
      }
      if (flags & FT_HAS_SLOT) {
        const int so = m.X.slots + 27*MI(ft_slot, b);
FB_UNROLL
        for (int k = 0; k < 6; k++) { I.A[k] += S(so, k); I.M[k] += S(so, 15 + k); pA[k] += S(so, 21 + k); }
FB_UNROLL
        for (int k = 0; k < 9; k++) I.H[k] += S(so, 6 + k);
      }
      float dq = 0.f;
      if (jtype == FB_JNT_FREE) {
        /* floating root: I a + pA = 0 for the (gravity-free frame) acceleration */
        float K[6][6], rhs[6];
        K[0][0] = I.A[0]; K[1][1] = I.A[1]; K[2][2] = I.A[2];
        K[1][0] = I.A[3]; K[2][0] = I.A[4]; K[2][1] = I.A[5];
        K[3][3] = I.M[0]; K[4][4] = I.M[1]; K[5][5] = I.M[2];
        K[4][3] = I.M[3]; K[5][3] = I.M[4]; K[5][4] = I.M[5];
FB_UNROLL
        for (int i = 0; i < 3; i++)
FB_UNROLL
          for (int j = 0; j < 3; j++) K[3 + j][i] = I.H[3*i + j];
FB_UNROLL
        for (int k = 0; k < 6; k++) rhs[k] = -pA[k];
        solve6(K, rhs, aroot);
        continue;
      }
      if (jid >= 0) {
        const int qa = MI(jnt_qposadr, jid), da = MI(jnt_dofadr, jid);
        const float qj = S(m.X.qpos, qa), qd = S(m.X.qvel, da);
        dq = qj - MF(jnt_qpos0, jid);
        float ax[3], U[6], c[6], tau, trq;
        m_rot(R, MF(jnt_axis, 3*jid), MF(jnt_axis, 3*jid+1), MF(jnt_axis, 3*jid+2), ax);
        actuation(jid, MI(ft_fj, b), qj, qd, time, store_ctrl, &tau, &trq);
        const float stiff = MF(jnt_stiffness, jid), damp = MF(dof_damping, da);
        if (stiff != 0.f) tau -= stiff*(qj - g_spring[qa]);
        tau -= damp*qd;
        float d, u;
        if (jtype == FB_JNT_HINGE) {
          sym_mul(I.A, ax, U);
          ht_mul(I.H, ax, U + 3);
          d = ax[0]*U[0] + ax[1]*U[1] + ax[2]*U[2];
          u = tau - (ax[0]*pA[0] + ax[1]*pA[1] + ax[2]*pA[2]);
          float aq[3] = {ax[0]*qd, ax[1]*qd, ax[2]*qd};
          v_cross(v, aq, c);
          v_cross(v + 3, aq, c + 3);
        } else {
          h_mul(I.H, ax, U);
          sym_mul(I.M, ax, U + 3);
          d = ax[0]*U[3] + ax[1]*U[4] + ax[2]*U[5];
          u = tau - (ax[0]*pA[3] + ax[1]*pA[4] + ax[2]*pA[5]);
          float aq[3] = {ax[0]*qd, ax[1]*qd, ax[2]*qd};
          c[0] = c[1] = c[2] = 0.f;
          v_cross(v, aq, c + 3);
        }
        d += MF(dof_armature, da) + hdt*damp;
        const float dinv = 1.0f/d;
        /* Ia = I - U U'/d */
        sym_rank1(I.A, U, dinv);
        sym_rank1(I.M, U + 3, dinv);
FB_UNROLL
        for (int i = 0; i < 3; i++)
FB_UNROLL
          for (int j = 0; j < 3; j++) I.H[3*i + j] -= dinv*U[i]*U[3 + j];
        /* pa = pA + Ia c + U u/d */
        float t0[3], t1[3];
        const float ud = u*dinv;
        sym_mul(I.A, c, t0); h_mul(I.H, c + 3, t1);
FB_UNROLL
        for (int k = 0; k < 3; k++) pA[k] += t0[k] + t1[k] + U[k]*ud;
        ht_mul(I.H, c, t0); sym_mul(I.M, c + 3, t1);
FB_UNROLL
        for (int k = 0; k < 3; k++) pA[3 + k] += t0[k] + t1[k] + U[3 + k]*ud;
FB_UNROLL
        for (int k = 0; k < 6; k++) S(m.X.wrench, 6*b + k) = U[k];
        S(m.X.u, b) = u; S(m.X.dinv, b) = dinv; S(m.X.trq, b) = trq;
      }
      if (p == 0) continue;        /* fixed base: nothing above */
      /* move to the parent's anchor and hand over */
      {
        float Rp[9], r[3];
        const Quat qp = quat_of(p);
        q_mat(qp, Rp);
        anchor_offset(b, jtype, jid, qp, Rp, R, dq, r);
        art_shift(I, pA, r);
      }
      if (flags & FT_TO_CARRY) {
        C = I;
FB_UNROLL
        for (int k = 0; k < 6; k++) pc[k] = pA[k];
      } else {
        const int so = m.X.slots + 27*MI(ft_pslot, b);
        if (flags & FT_FIRST_WRITER) {
FB_UNROLL
          for (int k = 0; k < 6; k++) { S(so, k) = I.A[k]; S(so, 15 + k) = I.M[k]; S(so, 21 + k) = pA[k]; }
FB_UNROLL
          for (int k = 0; k < 9; k++) S(so, 6 + k) = I.H[k];
        } else {
FB_UNROLL
          for (int k = 0; k < 6; k++) { S(so, k) += I.A[k]; S(so, 15 + k) += I.M[k]; S(so, 21 + k) += pA[k]; }
FB_UNROLL
          for (int k = 0; k < 9; k++) S(so, 6 + k) += I.H[k];
        }
      }
    }
  }

  /* ---- pass 3: root -> leaves, accelerations, Euler, joints / xfrc rows, drag */
  FB_MEM int pass_accel(const float *aroot, float *row_joints, float *row_xfrc) {
    const int nb = m.nbody;
    const float hdt = m.timestep;
    float ac[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   /* carry: acceleration of body b-1 */
    int bad = 0;
    for (int b = 1; b < nb; b++) {
      const int p = MI(body_parent, b), jid = MI(body_jnt, b), flags = MI(ft_flags, b);
      const int jtype = jid >= 0 ? MI(jnt_type, jid) : -1;
      const Quat q = quat_of(b);
      float R[9], v[6], a[6];
      q_mat(q, R);
FB_UNROLL
      for (int k = 0; k < 6; k++) v[k] = S(m.X.vel, 6*b + k);
      if (jtype == FB_JNT_FREE) {
        const int qa = MI(jnt_qposadr, jid), da = MI(jnt_dofadr, jid);
FB_UNROLL
        for (int k = 0; k < 6; k++) a[k] = aroot[k];
        float cr[3], wl[3], w[3];
        v_cross(v, v + 3, cr);
        m_rot_t(R, a[0], a[1], a[2], wl);
FB_UNROLL
        for (int k = 0; k < 3; k++) {
          float vn = S(m.X.qvel, da + k) + hdt*(a[3 + k] + m.grav[k] + cr[k]);
          S(m.X.qvel, da + k) = vn;
          float pn = S(m.X.qpos, qa + k) + hdt*vn;
          S(m.X.qpos, qa + k) = pn;
          w[k] = S(m.X.qvel, da + 3 + k) + hdt*wl[k];
          S(m.X.qvel, da + 3 + k) = w[k];
          bad |= !(fabsf(pn) < 1e30f);
        }
        float angle = hdt*v_normalize3(w), sn, cs;
        fb_sincos(0.5f*angle, &sn, &cs);
        Quat qr = {cs, w[0]*sn, w[1]*sn, w[2]*sn};
        Quat qn = q_normalize(q_mul(q, qr));
        S(m.X.qpos, qa+3) = qn.w; S(m.X.qpos, qa+4) = qn.x; S(m.X.qpos, qa+5) = qn.y; S(m.X.qpos, qa+6) = qn.z;
        bad |= !(fabsf(qn.w) < 1e30f) | !(fabsf(qn.x) < 1e30f) | !(fabsf(qn.y) < 1e30f) | !(fabsf(qn.z) < 1e30f);
      } else {
        float ap[6] = {0.f, 0.f, 0.f, -m.grav[0], -m.grav[1], -m.grav[2]};
        float Rp[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f}, r[3], cr[3];
        Quat qp = {1.f, 0.f, 0.f, 0.f};
        if (p > 0) {
          qp = quat_of(p);
          q_mat(qp, Rp);
          if (flags & FT_TO_CARRY) {
FB_UNROLL
            for (int k = 0; k < 6; k++) ap[k] = ac[k];
          } else {
            const int so = m.X.slots + 27*MI(ft_pslot, b);
FB_UNROLL
            for (int k = 0; k < 6; k++) ap[k] = S(so, k);
          }
        }
        float dq = 0.f, qj = 0.f, qd = 0.f;
        int qa = 0, da = 0;
        if (jid >= 0) {
          qa = MI(jnt_qposadr, jid); da = MI(jnt_dofadr, jid);
          qj = S(m.X.qpos, qa); qd = S(m.X.qvel, da);
          dq = qj - MF(jnt_qpos0, jid);
        }
        anchor_offset(b, jtype, jid, qp, Rp, R, dq, r);
        v_cross(ap, r, cr);
FB_UNROLL
        for (int k = 0; k < 3; k++) { a[k] = ap[k]; a[3 + k] = ap[3 + k] + cr[k]; }
        if (jid >= 0) {
          float ax[3], U[6];
          m_rot(R, MF(jnt_axis, 3*jid), MF(jnt_axis, 3*jid+1), MF(jnt_axis, 3*jid+2), ax);
FB_UNROLL
          for (int k = 0; k < 6; k++) U[k] = S(m.X.wrench, 6*b + k);
          float aq[3] = {ax[0]*qd, ax[1]*qd, ax[2]*qd}, c[3];
          if (jtype == FB_JNT_HINGE) {
            v_cross(v, aq, c);
            a[0] += c[0]; a[1] += c[1]; a[2] += c[2];
            v_cross(v + 3, aq, c);
            a[3] += c[0]; a[4] += c[1]; a[5] += c[2];
          } else {
            v_cross(v, aq, c);
            a[3] += c[0]; a[4] += c[1]; a[5] += c[2];
          }
          float ua = U[0]*a[0] + U[1]*a[1] + U[2]*a[2] + U[3]*a[3] + U[4]*a[4] + U[5]*a[5];
          const float qdd = (S(m.X.u, b) - ua)*S(m.X.dinv, b);
          const int o3 = jtype == FB_JNT_HINGE ? 0 : 3;
          a[o3] += ax[0]*qdd; a[o3 + 1] += ax[1]*qdd; a[o3 + 2] += ax[2]*qdd;
          const float qdn = qd + hdt*qdd, qn = qj + hdt*qdn;
          S(m.X.qvel, da) = qdn;
          S(m.X.qpos, qa) = qn;
          bad |= !(fabsf(qn) < 1e30f);
          /* joints row: physics.py:481-524 (new position/velocity, forces of the old state) */
          const int fj = MI(ft_fj, b);
          if (fj >= 0) {
            float *row = row_joints + m.joint_cols*fj;
            const float trq = S(m.X.trq, b)*m.inv_torques;
            for (int k = 0; k < m.joint_cols; k += 2) {
              float x0 = 0.f, x1 = 0.f;
              if (k == (m.col_jpos & ~1)) { if (m.col_jpos & 1) x1 = qn; else x0 = qn; }
              if (k == (m.col_jvel & ~1)) { if (m.col_jvel & 1) x1 = qdn*m.inv_angvel; else x0 = qdn*m.inv_angvel; }
              if (k == (m.col_jtrq & ~1)) { if (m.col_jtrq & 1) x1 = trq; else x0 = trq; }
              fb_st2(row + k, x0, x1);
            }
          }
        }
      }
FB_UNROLL
      for (int k = 0; k < 6; k++) ac[k] = a[k];
      if (flags & FT_HAS_SLOT) {
        const int so = m.X.slots + 27*MI(ft_slot, b);
FB_UNROLL
        for (int k = 0; k < 6; k++) S(so, k) = a[k];
      }
      /* xfrc row + the wrench applied during the next step (drag.pyx:152-268, 3.4) */
      const int xr = MI(body_xfrcrow, b);
      if (xr >= 0) {
        float F[3] = {0.f, 0.f, 0.f}, Tq[3] = {0.f, 0.f, 0.f}, wf[3] = {0.f, 0.f, 0.f}, wt[3] = {0.f, 0.f, 0.f};
        const int i = MI(ft_swim, b);
        if (m.water_drag && i >= 0) {
          float h[3], cr[3], lin[3], vl[3], wl[3], uw[3], buoy[3] = {0.f, 0.f, 0.f};
          m_rot(R, MF(ft_hloc, 3*b), MF(ft_hloc, 3*b+1), MF(ft_hloc, 3*b+2), h);
          const float pz = (S(m.X.org, 3*b + 2) + h[2])*m.inv_meters;
          if (!(pz > m.water_surface)) {                 /* drag.pyx:192-194 */
            v_cross(v, h, cr);
FB_UNROLL
            for (int k = 0; k < 3; k++) lin[k] = v[3 + k] + cr[k];
            m_rot_t(R, lin[0]*m.inv_velocity, lin[1]*m.inv_velocity, lin[2]*m.inv_velocity, vl);
            m_rot_t(R, v[0]*m.inv_angvel, v[1]*m.inv_angvel, v[2]*m.inv_angvel, wl);
            const float mass = MF(swim_mass, i);
            if (m.water_buoyancy && mass > 0.f && pz < m.water_surface) {
              float frac = fminf(fmaxf(m.water_surface - pz, 0.f)/MF(swim_height, i), 1.f);
              float lift = -1000.f*mass*(-9.81f)/MF(swim_density, i)*frac;
              m_rot_t(R, 0.f, 0.f, lift, buoy);
            }
            m_rot_t(R, m.water_velocity[0], m.water_velocity[1], m.water_velocity[2], uw);
FB_UNROLL
            for (int k = 0; k < 3; k++) {
              float vv = vl[k] - uw[k], w = wl[k];
              float sv = vv < 0.f ? -vv*vv : vv*vv, sw = w < 0.f ? -w*w : w*w;
              F[k] = sv*m.water_viscosity*MF(swim_coef, 6*i + k) + buoy[k];
              Tq[k] = sw*MF(swim_coef, 6*i + 3 + k);
            }
            m_rot(R, F[0], F[1], F[2], wf);
            m_rot(R, Tq[0], Tq[1], Tq[2], wt);
FB_UNROLL
            for (int k = 0; k < 3; k++) { wf[k] *= m.newtons; wt[k] *= m.torques; }
          }
        }
        float *row = row_xfrc + 6*xr;
        fb_st2(row, F[0], F[1]); fb_st2(row + 2, F[2], Tq[0]); fb_st2(row + 4, Tq[1], Tq[2]);
FB_UNROLL
        for (int k = 0; k < 3; k++) { S(m.X.wrench, 6*b + k) = wf[k]; S(m.X.wrench, 6*b + 3 + k) = wt[k]; }
      } else {
        /* user-applied wrench: persistent, re-read (the slot held U during this step) */
        const float *gx = P.xfrc_applied + ((size_t)env*m.nbody + b)*6;
FB_UNROLL
        for (int k = 0; k < 6; k++) S(m.X.wrench, 6*b + k) = gx[k];
      }
    }
    return bad;
  }

  /* Returns the number of steps taken here (n_steps unless a constraint appeared). */
  FB_MEM int run() {
    load_state();
    const size_t e = (size_t)env;
    const int n = P.n_steps;
    int k = 0;
    for (; k < n; k++) {
      const long long row = (P.it0 + k + 1) % P.ring;
      float *row_links = P.log_links + e*P.links_env_stride + row*(long long)(m.n_links*20);
      float *row_joints = P.log_joints + e*P.joints_env_stride + row*(long long)(m.n_joints*m.joint_cols);
      float *row_contacts = P.log_contacts + e*P.contacts_env_stride + row*(long long)(m.n_contacts*12);
      float *row_xfrc = P.log_xfrc + e*P.xfrc_env_stride + row*(long long)(m.n_xfrc*6);
      const float time = (float)(P.it0 + k)*m.timestep;
      if (pass_poses(row_links)) break;
      float aroot[6] = {0.f, 0.f, 0.f, -m.grav[0], -m.grav[1], -m.grav[2]};
      pass_inertia(time, aroot, k == n - 1 && m.n_wc > 0);   /* ctrl is left as the team path leaves it */
      int bad = pass_accel(aroot, row_joints, row_xfrc);
      /* no contact is active on this path: the contacts rows are zero (sensors.pyx:140-190) */
      for (int i = 0; i < m.n_contacts*3; i++) fb_st4(row_contacts + 4*i, 0.f, 0.f, 0.f, 0.f);
      if (bad) FB_FLAG_OR(P.flags + env, FB_FLAG_NONFINITE);
    }
    store_state(P.it0 + k);
    return k;
  }
};

#endif /* FB_FAST_H_ */
