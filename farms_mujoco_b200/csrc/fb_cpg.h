/*
 * fb_cpg.h -- on-device central pattern generator: a network of coupled phase oscillators that
 * stands where the per-iteration Python of ExperimentTask.step_control stands in the reference
 * (farms_mujoco/simulation/task.py:288-346: controller.step, then ctrl[pos_map] = positions[j],
 * ctrl[trq_map] = torques[j]*units.torques).  SURVEY.md section 8f-1.
 *
 *   theta_i' = 2 pi f_i + sum_j w_ij r_j sin(theta_j - theta_i - phi_ij)
 *   r_i''    = a_i (a_i/4 (R_i - r_i) - r_i')
 *   output_o = offset_o + gain_o (r_a (1 + cos theta_a) - r_b (1 + cos theta_b))      (b >= 0)
 *            = offset_o + gain_o r_a cos theta_a                                      (b < 0)
 *
 * integrated with explicit Euler at the physics time step.  ctrl of iteration k is the output of
 * the state after k Euler steps (the controller's step() advances it between iterations).  One
 * thread = one environment; the kernel runs ahead of the step kernels of a launch and writes the
 * n_steps control vectors into the control sequence they read (FbParams::ctrl_seq), so all three
 * step kernels consume it unchanged.  A torque output is a ctrl index of a motor actuator with
 * gain = units.torques.
 */
#ifndef FB_CPG_H_
#define FB_CPG_H_

#include "fb_device.h"

#define FB_CPG_MAXOSC 64

struct CpgDev {
  int n_osc, n_coupling, n_out;
  const float *freq, *amp, *rate;           /* [n_osc] */
  const int *c_from, *c_to;                 /* [n_coupling] j -> i */
  const float *c_w, *c_phi;                 /* [n_coupling] */
  const int *o_act, *o_a, *o_b;             /* [n_out] */
  const float *o_gain, *o_off;              /* [n_out] */
  float *theta, *r, *rd;                    /* state [n_osc][env_pad] */
  /* spring references (task.py:338-346: model.qpos_spring[joint] = springrefs[joint]): output s
   * goes to row nu + s of the sequence step */
  int n_spring;
  const int *s_a, *s_b;                     /* [n_spring] */
  const float *s_gain, *s_off;              /* [n_spring] */
};

/* n_steps control vectors of one environment into seq[(k*stride + a)*env_pad + env] (stride = nu +
 * n_spring rows per step); actuators the network does not drive keep ctrl[env][a] */
FB_DEV void fb_cpg_env(const CpgDev &c, int env, long long env_pad, int n_steps, int nu, float dt,
                       const float *ctrl, float *seq) {
  const int stride = nu + c.n_spring;
  float th[FB_CPG_MAXOSC], r[FB_CPG_MAXOSC], rd[FB_CPG_MAXOSC], dth[FB_CPG_MAXOSC];
  for (int i = 0; i < c.n_osc; i++) {
    th[i] = c.theta[(long long)i*env_pad + env]; r[i] = c.r[(long long)i*env_pad + env];
    rd[i] = c.rd[(long long)i*env_pad + env];
  }
  for (int k = 0; k < n_steps; k++) {
    float *row = seq + (long long)k*stride*env_pad + env;
    for (int a = 0; a < nu; a++) row[(long long)a*env_pad] = ctrl[(size_t)env*nu + a];
    for (int o = 0; o < c.n_out; o++) {
      const int a = c.o_a[o], b = c.o_b[o];
      const float v = b >= 0 ? r[a]*(1.f + cosf(th[a])) - r[b]*(1.f + cosf(th[b])) : r[a]*cosf(th[a]);
      row[(long long)c.o_act[o]*env_pad] = c.o_off[o] + c.o_gain[o]*v;
    }
    for (int o = 0; o < c.n_spring; o++) {
      const int a = c.s_a[o], b = c.s_b[o];
      const float v = b >= 0 ? r[a]*(1.f + cosf(th[a])) - r[b]*(1.f + cosf(th[b])) : r[a]*cosf(th[a]);
      row[(long long)(nu + o)*env_pad] = c.s_off[o] + c.s_gain[o]*v;
    }
    /* explicit Euler to the next iteration */
    for (int i = 0; i < c.n_osc; i++) dth[i] = 6.283185307179586f*c.freq[i];
    for (int e = 0; e < c.n_coupling; e++) {
      const int i = c.c_to[e], j = c.c_from[e];
      dth[i] += c.c_w[e]*r[j]*sinf(th[j] - th[i] - c.c_phi[e]);
    }
    for (int i = 0; i < c.n_osc; i++) {
      const float a = c.rate[i], rdd = a*(0.25f*a*(c.amp[i] - r[i]) - rd[i]);
      th[i] += dt*dth[i];
      r[i] += dt*rd[i];
      rd[i] += dt*rdd;
    }
    /* keep the phases small: the couplings and outputs only see them modulo 2 pi, and fp32
     * resolves a phase of 60 rad to 4e-6 only.  All of them shift together. */
    if (th[0] > 6.283185307179586f) {
      for (int i = 0; i < c.n_osc; i++) th[i] -= 6.283185307179586f;
    }
  }
  for (int i = 0; i < c.n_osc; i++) {
    c.theta[(long long)i*env_pad + env] = th[i]; c.r[(long long)i*env_pad + env] = r[i];
    c.rd[(long long)i*env_pad + env] = rd[i];
  }
}

#endif /* FB_CPG_H_ */
