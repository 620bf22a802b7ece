"""Analytic known-answer tests of the oracle's contact pipeline (CPU).  The reference's tests pin nothing for contacts and
MuJoCo is not installable here, so these hold the restatement to numbers derived in tests/kat_analytic.py from the scene
parameters and MuJoCo's documented constraint model - not to anything the oracle itself computes.  GPU twins of every case
run the CUDA path against the same numbers (tests/test_scene_gpu.py)."""
import numpy as np
import pytest

import kat_analytic as ka
from oracle.oracle import OracleSim


def _sim():
  o = OracleSim(ka.BLOB, collide=True)
  return o


def _contacts(o):
  return [(c['geom1'], c['geom2'], c['dist'], c['pos'], c['frame'][0]) for c in o.contacts()]


def test_model_constants_match_the_closed_forms():
  o = _sim()
  m = o.meta
  nb = int(m['nbody'][0])
  assert abs(m['body_mass'][nb - 2] - ka.BOX_MASS) < 1e-15 and abs(m['body_mass'][nb - 1] - ka.CAP_MASS) < 1e-15
  # [upstream] mj_setConst: invweight0 of a free body = 1 / m (translation) and mean(1 / I_principal) (rotation)
  iw = m['body_invweight0'].reshape(-1, 2)
  assert abs(iw[nb - 2, 0] * ka.BOX_MASS - 1) < 1e-12
  I = ka.BOX_MASS / 3 * np.array([ka.BOX_HALF[1]**2 + ka.BOX_HALF[2]**2, ka.BOX_HALF[0]**2 + ka.BOX_HALF[2]**2, ka.BOX_HALF[0]**2 + ka.BOX_HALF[1]**2])
  assert abs(iw[nb - 2, 1] / np.mean(1 / I) - 1) < 1e-12
  assert np.allclose(np.sort(m['body_inertia'].reshape(-1, 3)[nb - 2]), np.sort(I), rtol=1e-12)


@pytest.mark.parametrize('name', sorted(ka.geometry_cases()))
def test_contact_geometry_matches_the_closed_form(name):
  """One collision pass at a constructed pose: contact count, pair, dist = -depth, normal = +z, positions under the prop's
  lowest points halfway between the two surfaces (box-box, box-plane, capsule-box, capsule-plane; face, edge and vertex)."""
  case = ka.geometry_cases()[name]
  o = _sim()
  q = ka.scene_state(o.meta['qpos0'], case['box'][0], case['box'][1], case['cap'][0], case['cap'][1])
  o.set_state(q, np.zeros(18)); o.forward()
  curved = name.startswith('capsule_end')   # EPA on the spherical cap terminates by tolerance: normal to ~sqrt(2 tol / r)
  ka.check_contacts(_contacts(o), case, pos_tol=1e-4 if curved else 1e-9, normal_tol=5e-4 if curved else 1e-9, dist_tol=2e-9 if curved else 1e-12)


def _settle(o, q, steps=100):
  o.set_state(q, np.zeros(18))
  for _ in range(steps):
    o.control_step(np.zeros(6))


def test_box_rests_at_the_closed_form_penetration_and_carries_its_weight():
  """Box flat on the table, 2 s: the four corner contacts share m g, so the rest penetration solves
  4 * imp^2 / (1 - imp) * m * K * d = m g (solref mixed to 0.012, solimp default).  Also sum of the normal forces = m g."""
  d = ka.rest_depth(ka.BOX_MASS, 4)
  assert 3e-5 < d < 5e-5
  o = _sim()
  q = ka.scene_state(o.meta['qpos0'], (0.25, 0.0, ka.TABLE_TOP + ka.BOX_HALF[2]))
  _settle(o, q)
  z = o.qpos[8]
  print('box rest penetration: analytic %.6e measured %.6e' % (d, ka.TABLE_TOP + ka.BOX_HALF[2] - z))
  assert abs((ka.TABLE_TOP + ka.BOX_HALF[2] - z) - d) < 1e-9
  assert np.abs(o.qvel[6:12]).max() < 1e-9
  cons = [c for c in o.contacts() if (c['geom1'], c['geom2']) == (ka.TABLE_GEOM, ka.BOX_GEOM)]
  assert len(cons) == 4 and all(c['dim'] == 6 for c in cons)
  f = o.field('efc_force')
  fn = sum(f[c['efc_address']] for c in cons)
  assert abs(fn / (ka.BOX_MASS * ka.G) - 1) < 1e-6, fn
  for c in cons:   # contact parameter mixing: friction = element-wise max (1, 1, 0.01, 0.01, 0.01), solref = mean
    assert np.allclose(c['friction'], [1.0, 1.0, 0.01, 0.01, 0.01]) and np.allclose(c['solref'], ka.SOLREF)


def test_capsule_rests_at_the_closed_form_penetration():
  d = ka.rest_depth(ka.CAP_MASS, 2)
  o = _sim()
  q = ka.scene_state(o.meta['qpos0'], (0.25, -0.25, 0.7), cap_pos=(0.25, 0.0, ka.TABLE_TOP + ka.CAP_R), cap_quat=ka.quat_about((0, 1, 0), np.pi / 2))
  q[6:9] = (0.25, -0.2, ka.TABLE_TOP + ka.BOX_HALF[2])   # the box rests beside it
  _settle(o, q)
  z = o.qpos[15]
  print('capsule rest penetration: analytic %.6e measured %.6e' % (d, ka.TABLE_TOP + ka.CAP_R - z))
  assert abs((ka.TABLE_TOP + ka.CAP_R - z) - d) < 5e-9   # (EPA depth tolerance 1e-9 on the cylindrical side)
  assert np.abs(o.qvel[12:18]).max() < 1e-6


def test_stacked_bodies_rest_at_the_closed_form_penetrations_and_pass_the_load_down():
  """Capsule on box on table (ka.stack_state): contact rows between TWO dynamic bodies (Jacobian blocks of both, invweight sum,
  prop / prop parameter mixing) and the load transfer through the lower body.  Penetrations against the closed forms, normal
  forces: capsule / box contacts carry m_cap g, box / table contacts (m_box + m_cap) g."""
  d_low, d_up = ka.stack_rest_depths()
  o = _sim()
  _settle(o, ka.stack_state(o.meta['qpos0']), steps=150)
  zb, zc = o.qpos[8], o.qpos[15]
  m_low, m_up = ka.TABLE_TOP + ka.BOX_HALF[2] - zb, (zb + ka.BOX_HALF[2]) - (zc - ka.CAP_R)
  print('stack rest penetrations: box/table analytic %.9e measured %.9e; capsule/box analytic %.9e measured %.9e' % (d_low, m_low, d_up, m_up))
  assert abs(m_low - d_low) < 1e-9 and abs(m_up - d_up) < 1e-9
  assert np.abs(o.qvel[6:18]).max() < 1e-9
  f = o.field('efc_force')
  low = [c for c in o.contacts() if (c['geom1'], c['geom2']) == (ka.TABLE_GEOM, ka.BOX_GEOM)]
  up = [c for c in o.contacts() if (c['geom1'], c['geom2']) == (ka.BOX_GEOM, ka.CAPSULE_GEOM)]
  assert len(low) == 4 and len(up) == 2
  assert abs(sum(f[c['efc_address']] for c in low) / ((ka.BOX_MASS + ka.CAP_MASS) * ka.G) - 1) < 1e-6
  assert abs(sum(f[c['efc_address']] for c in up) / (ka.CAP_MASS * ka.G) - 1) < 1e-6
  assert sorted(round(c['pos'][0] - 0.25, 9) for c in up) == [-0.03, 0.03]          # the line contact clipped to the box face
  assert all(np.allclose(c['solref'], (0.004, 1.0)) for c in up)


def test_sliding_box_decelerates_at_mu_g():
  """A box sliding on the table at 0.3 m/s: every contact sits on the friction cone, so the deceleration is mu * g (mu = 1;
  elliptic cone, impratio 10) while it slides, and it stops without reversing."""
  o = _sim()
  q = ka.scene_state(o.meta['qpos0'], (0.1, 0.0, ka.TABLE_TOP + ka.BOX_HALF[2]))
  _settle(o, q, steps=50)
  v = o.qvel.copy(); v[6] = 0.3
  o.set_state(o.qpos.copy(), v)
  for _ in range(3):
    o.substep()
  v0 = o.qvel[6]
  n = 8
  for _ in range(n):
    o.substep()
  a = (o.qvel[6] - v0) / (n * ka.DT)
  print('sliding deceleration: measured %.5f, mu g = %.5f' % (-a, ka.MU * ka.G))
  assert abs(-a / (ka.MU * ka.G) - 1) < 2e-2
  for _ in range(40):
    o.substep()
  assert abs(o.qvel[6]) < 1e-3   # stopped (below the reward's "moving" threshold, success_detector_utils.py:19) ...
  for _ in range(200):
    o.substep()
  assert abs(o.qvel[6]) < 1e-5   # ... and the residual creep of the soft friction rows dies out


def test_free_fall_follows_semi_implicit_euler():
  """No contact: z_n = z_0 - g h^2 n (n + 1) / 2 and v_n = -g h n exactly ([upstream] mj_Euler: velocity first)."""
  o = _sim()
  q = ka.scene_state(o.meta['qpos0'], (0.25, 0.0, 0.8), cap_pos=(0.25, 0.3, 0.9))
  o.set_state(q, np.zeros(18))
  n = 30
  for _ in range(n):
    o.substep()
  assert abs(o.qvel[8] + ka.G * ka.DT * n) < 1e-12 and abs(o.qpos[8] - (0.8 - ka.G * ka.DT**2 * n * (n + 1) / 2)) < 1e-12
  assert abs(o.qpos[15] - (0.9 - ka.G * ka.DT**2 * n * (n + 1) / 2)) < 1e-12


def test_implicitfast_satisfies_its_defining_equation():
  """integrator = implicitfast (named by north_star; the reference itself runs MuJoCo's default Euler): [upstream] mj_implicit
  advances the velocity with x solving (M - h D) x = M qacc, D = d(qfrc_smooth)/d(qvel) without Coriolis terms - here the
  actuators' velocity gain biasprm[2] = +1 (scene_pbr.xml:11) on every actuator whose force is not clamped.  Checked from the
  oracle's own M, qacc and actuator forces of the substep, independently of how integrate() solves it; Euler gives x = qacc."""
  rs = np.random.RandomState(0)
  for integ in ('euler', 'implicitfast'):
    o = OracleSim('so100_arm', collide=False, integrator=integ)
    q = 0.3 * rs.uniform(-1, 1, 6); q[2] = abs(q[2]) + 0.2
    o.set_state(q, rs.uniform(-1, 1, 6)); o.ctrl[:] = q + 0.2 * rs.uniform(-1, 1, 6)
    v0 = o.qvel.copy()
    o.substep()
    M, qacc, fa = o.field('M', 36).reshape(6, 6).copy(), o.field('qacc', 6).copy(), o.field('qfrc_actuator', 6).copy()
    x = (o.qvel - v0) / ka.DT
    D = np.diag((np.abs(fa) < 35.0).astype(float))
    assert D.trace() >= 5
    if integ == 'implicitfast':
      assert np.abs((M - ka.DT * D) @ x - M @ qacc).max() < 1e-11 and np.abs(x - qacc).max() > 1e-2
    else:
      assert np.abs(x - qacc).max() < 1e-11
