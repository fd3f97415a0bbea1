/* oracle.h -- data block of the fp64 CPU oracle (test infrastructure only;
 * see the header of mjstep_oracle.c).  All pointers are caller-allocated
 * (NumPy arrays owned by oracle/oracle.py). */
#ifndef FARMS_ORACLE_H_
#define FARMS_ORACLE_H_
#include <stdint.h>

enum { ORC_CNSTR_LIMIT = 3, ORC_CNSTR_CONTACT_PYRAMIDAL = 5 };

typedef struct OrcData {
  /* state / inputs */
  double *qpos, *qvel, *ctrl, *xfrc_applied /* [nbody][6] force,torque */, *qpos_spring;
  double time;
  /* position stage */
  double *xpos, *xquat, *xmat, *xipos, *ximat, *xanchor, *xaxis, *geom_xpos, *geom_xmat;
  double *subtree_com, *cinert, *crb, *cdof, *cdof_dot, *cvel;
  double *qM, *qLD, *qLDiagInv;
  /* contacts [ncand] */
  int32_t ncon;
  int32_t *con_cand, *con_efc_address;
  double *con_dist, *con_pos, *con_frame, *con_force /* [ncand][6] */;
  /* constraints [maxefc] */
  int32_t nefc;
  int32_t *efc_type, *efc_id, *jnt_limit_row;
  double *efc_J, *efc_pos, *efc_margin, *efc_R, *efc_D, *efc_KBIP, *efc_aref, *efc_force;
  /* forces / accelerations [nv] */
  double *qfrc_bias, *qfrc_passive, *qfrc_actuator, *qfrc_smooth, *qacc_smooth, *qacc, *qfrc_constraint;
  double *actuator_force;    /* [nu] */
  /* sensor-stage quantities */
  double *body_linvel, *body_angvel;  /* [nbody][3] framelinvel / frameangvel, objtype=body */
  double *jnt_limit_force;            /* [njnt] */
  int32_t solver_niter;
} OrcData;

#endif
