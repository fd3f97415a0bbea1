/* farms_mujoco_b200 -- the reference's drag_forces (swimming/drag.pyx:152-268: link_swimming_info
 * :12-63, compute_buoyancy :111-149, compute_force :66-88, compute_torque :91-108) on one link
 * row, float64 like the reference: the body of the stand-alone operator fb_drag_forces
 * (include/farms_b200.h).  The step kernels compute the same forces fused into the step in fp32
 * (fb_fast.h pass_accel, fb_device.h write_log); this file is for callers that hold link rows of
 * their own.  Quaternions are xyzw (the farms links row: CoM position 0..2, CoM orientation 3..6,
 * URDF orientation 10..13, CoM linear velocity 14..16, angular velocity 17..19). */
#ifndef FB_DRAG_H_
#define FB_DRAG_H_

#include "fb_device.h"

struct FbDragArgs {
  const double *links;        /* [n][20] */
  const double *coef;         /* [n][6]: linear x y z, angular x y z */
  const double *mass, *height, *density;   /* [n] */
  double surface, viscosity, gravity, wvel[3];
  int use_buoyancy, n;
  double *xfrc;               /* [n][6], written where applied */
  int *applied;               /* [n] */
};

/* Hamilton product a*b, xyzw */
FB_DEV void fbd_qmul(const double *a, const double *b, double *o) {
  const double x0 = a[0], y0 = a[1], z0 = a[2], w0 = a[3], x1 = b[0], y1 = b[1], z1 = b[2], w1 = b[3];
  o[0] = w0*x1 + x0*w1 + y0*z1 - z0*y1;
  o[1] = w0*y1 - x0*z1 + y0*w1 + z0*x1;
  o[2] = w0*z1 + x0*y1 - y0*x1 + z0*w1;
  o[3] = w0*w1 - x0*x1 - y0*y1 - z0*z1;
}
/* q (v, 0) q*: the reference's quat_rot */
FB_DEV void fbd_qrot(const double *v, const double *q, double *o) {
  const double v4[4] = {v[0], v[1], v[2], 0.0}, qc[4] = {-q[0], -q[1], -q[2], q[3]};
  double t[4], r[4];
  fbd_qmul(q, v4, t);
  fbd_qmul(t, qc, r);
  o[0] = r[0]; o[1] = r[1]; o[2] = r[2];
}

FB_DEV void fb_drag_row(const FbDragArgs &A, int i) {
  const double *row = A.links + 20*(size_t)i;
  const double pos_z = row[2];
  A.applied[i] = 0;
  if (pos_z > A.surface) return;                       /* drag.pyx:192-194: the row stays as it is */
  const double *urdf2global = row + 10, *com2global = row + 3;
  const double global2urdf[4] = {-urdf2global[0], -urdf2global[1], -urdf2global[2], urdf2global[3]};
  double com2urdf[4], lin[3], ang[3], buoy[3] = {0.0, 0.0, 0.0}, wv[3], force[3], torque[3];
  fbd_qmul(global2urdf, com2global, com2urdf);
  const double urdf2com[4] = {-com2urdf[0], -com2urdf[1], -com2urdf[2], com2urdf[3]};
  fbd_qrot(row + 14, global2urdf, lin);
  fbd_qrot(row + 17, global2urdf, ang);
  if (A.use_buoyancy && A.mass[i] > 0.0 && pos_z < A.surface) {
    const double frac = fmin(fmax(A.surface - pos_z, 0.0)/A.height[i], 1.0);
    const double lift[3] = {0.0, 0.0, -1000.0*A.mass[i]*A.gravity/A.density[i]*frac};
    fbd_qrot(lift, global2urdf, buoy);
  }
  fbd_qrot(A.wvel, global2urdf, wv);
  const double *c = A.coef + 6*(size_t)i;
  for (int k = 0; k < 3; k++) {
    const double v = lin[k] - wv[k], w = ang[k];
    force[k] = (v < 0.0 ? -v*v : v*v)*A.viscosity*c[k] + buoy[k];
    torque[k] = (w < 0.0 ? -w*w : w*w)*c[3 + k];
  }
  fbd_qrot(force, urdf2com, A.xfrc + 6*(size_t)i);
  fbd_qrot(torque, urdf2com, A.xfrc + 6*(size_t)i + 3);
  A.applied[i] = 1;
}

#endif  /* FB_DRAG_H_ */
