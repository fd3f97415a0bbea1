"""Experiment behind DESIGN.md section 8, item (3): can the frictionless <pair>s the reference
generates (friction=[0]*5 -> mu = 1e-5, mjcf.py:1029) be solved in fp32?

The constraint problem of one step is taken from the fp64 oracle (M, qacc_smooth, J, D, aref of
SALAMANDER with the feet pressed together) and solved three times in NumPy:

  (a) the primal Newton step of the kernels, all arithmetic in float32;
  (b) the same in float64 (what the oracle does);
  (c) float32 again, but with the stiff rows -- the four pyramidal rows of a frictionless contact,
      D = 1/(2 mu^2 R) ~ 1e10 -- replaced by ONE hard row per contact, n.J.a >= aref with a
      multiplier (the normal force) as unknown: Newton on the soft rows, Schur complement
      S = Jn H^-1 Jn' on the multipliers, active set on lambda >= 0.

(c) drops terms of relative size mu (the 1e-5 n of tangential force a mu = 1e-5 cone can carry and
the O(mu) softness of the normal direction).  Printed: normal forces and qacc against the oracle.

    python tools/experiments/frictionless_pairs_saddle.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import variant_models  # noqa: E402
from farms_mujoco_b200 import mjcf_subset  # noqa: E402
from oracle.oracle import OraclePhysics  # noqa: E402

mjcf_subset.PAIR_MIN_FRICTION = 0.0                     # the oracle (fp64) takes them
spec = variant_models.salamander_foot_pairs(friction=0)
model = mjcf_subset.parse_mjcf(spec.mjcf)
orc = OraclePhysics(model)
orc.reset(keyframe_id=0)
rng = np.random.default_rng(1)
qpos = variant_models.folded_legs_qpos(model, 1.0)
qpos[7:] += rng.uniform(-0.03, 0.03, model.nq - 7)
qpos[7 + 3] = 1.004                                      # a trunk joint past its limit: a soft row in the mix
orc.data.qpos[:] = qpos
orc.data.qvel[:] = rng.uniform(-0.2, 0.2, model.nv)
orc.forward()
nv, nefc = model.nv, orc.nefc
A = orc.arrays
J = A['efc_J'][:nefc*nv].reshape(nefc, nv).copy()
D, aref, force = A['efc_D'][:nefc].copy(), A['efc_aref'][:nefc].copy(), A['efc_force'][:nefc].copy()
a_s, a_ref = A['qacc_smooth'].copy(), A['qacc'].copy()
M = np.zeros((nv, nv))
for i in range(nv):
    adr, k = model.dof_Madr[i], i
    while k >= 0:
        M[i, k] = M[k, i] = A['qM'][adr]
        adr, k = adr + 1, model.dof_parentid[k]
stiff, contacts = np.zeros(nefc, bool), []
for i in range(orc.ncon):
    adr, cand = int(orc._d.con_efc_address[i]), int(A['con_cand'][i]) if 'con_cand' in A else int(orc._d.con_cand[i])
    if adr >= 0 and model.cand_friction[cand] < 1e-3:
        stiff[adr:adr + 4] = True
        contacts.append(np.arange(adr, adr + 4))
print(f'nv {nv}  rows {nefc}  stiff rows {int(stiff.sum())}  D soft ~{np.median(D[~stiff]):.3g}  D stiff ~{np.median(D[stiff]):.3g}')
n_ref = np.array([force[rows].sum() for rows in contacts])
print('oracle normal forces', n_ref)


def primal_newton(dtype, iters=50):
    Mt, Jt, Dt, ar, a = (x.astype(dtype) for x in (M, J, D, aref, a_s))
    a0 = a.copy()
    for _ in range(iters):
        res = Jt @ a - ar
        act = res < 0
        grad = Mt @ (a - a0) + Jt.T @ (Dt*np.where(act, res, 0))
        H = Mt + (Jt[act].T*Dt[act]) @ Jt[act]
        try:
            p = -np.linalg.solve(H.astype(dtype), grad)
        except np.linalg.LinAlgError:
            return a*np.nan, res*np.nan
        # exact line search by bisection on phi'(alpha)
        jp = Jt @ p
        lo, hi = dtype(0), dtype(1)
        dphi = lambda al: p @ (Mt @ (a + al*p - a0)) + np.sum(Dt*jp*np.minimum(0, res + al*jp))
        while dphi(hi) < 0 and hi < 64:
            hi *= 2
        for _ in range(40):
            mid = (lo + hi)/2
            lo, hi = (mid, hi) if dphi(mid) < 0 else (lo, mid)
        a = a + ((lo + hi)/2)*p
    res = Jt @ a - ar
    return a, -Dt*np.minimum(0, res)


def saddle_newton(dtype=np.float32, iters=30):
    soft = ~stiff
    Mt, Js, Ds, ars = (x.astype(dtype) for x in (M, J[soft], D[soft], aref[soft]))
    Jn = np.array([J[rows].mean(axis=0) for rows in contacts]).astype(dtype)       # n.J of each stiff contact
    arn = np.array([aref[rows].mean() for rows in contacts]).astype(dtype)
    a0 = a_s.astype(dtype)
    a, lam = a0.copy(), np.zeros(len(contacts), dtype=dtype)
    on = np.ones(len(contacts), bool)
    for _ in range(iters):
        res = Js @ a - ars
        act = res < 0
        g = Mt @ (a - a0) + Js.T @ (Ds*np.where(act, res, 0)) - Jn[on].T @ lam[on]
        H = Mt + (Js[act].T*Ds[act]) @ Js[act]
        Hi_g = np.linalg.solve(H, g)
        Hi_J = np.linalg.solve(H, Jn[on].T)
        x = Jn[on] @ a - arn[on]                        # violation of the hard rows
        S = Jn[on] @ Hi_J
        dlam = np.linalg.solve(S, -(x - Jn[on] @ Hi_g)) if on.any() else np.zeros(0, dtype)
        p = -Hi_g + Hi_J @ dlam
        a, lam[on] = a + p, lam[on] + dlam
        release = on & (lam < 0)
        lam[release] = 0
        on = on & ~release
        on |= (~on) & (Jn @ a - arn < 0)
    return a, lam


for label, (a, f) in (('(a) primal Newton, float32', primal_newton(np.float32)), ('(b) primal Newton, float64', primal_newton(np.float64))):
    n = np.array([f[rows].sum() for rows in contacts])
    print(f'{label}: normal forces {n}  rel err {np.abs(n - n_ref).max()/np.abs(n_ref).max():.2e}  '
          f'qacc rel err {np.abs(a - a_ref).max()/np.abs(a_ref).max():.2e}')
a, lam = saddle_newton()
print(f'(c) hard normal rows + multipliers, float32: normal forces {lam}  rel err {np.abs(lam - n_ref).max()/np.abs(n_ref).max():.2e}  '
      f'qacc rel err {np.abs(a - a_ref).max()/np.abs(a_ref).max():.2e}')
