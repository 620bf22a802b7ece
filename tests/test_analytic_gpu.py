"""GPU twins of tests/test_oracle_analytic.py: the CUDA contact pipeline (through the C-ABI) against the closed-form known
answers of tests/kat_analytic.py - contact geometry of box / capsule against the table box and the floor plane, rest
penetration from solref / solimp, sliding deceleration mu g, free fall - in float64 and in the float32 product arithmetic.
Nothing here consults the oracle: these pin the contact pipeline to numbers this project did not compute."""
import numpy as np
import pytest
import torch

import kat_analytic as ka

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _kat_env(built, n, precision, control_timestep=0.02):
  from so101_sim_b200.task_suite import BatchedEnvironment, SO100HandOver
  task = SO100HandOver('banana', control_timestep=control_timestep, cameras=())
  task.model_name = ka.BLOB          # the arm scene with the two primitive props (tools/make_kat_blob.py)
  return BatchedEnvironment(task, n, 30.0, 0, DEV, None, precision, 100, 1e-8 if precision == 'f64' else 1e-6)


def _put(env, rows, vel=None):
  q = torch.tensor(np.stack(rows), dtype=torch.float64)
  v = torch.zeros(len(rows), 18, dtype=torch.float64) if vel is None else torch.tensor(np.stack(vel), dtype=torch.float64)
  env.set_initial_state(q, v)
  env.reset()


@pytest.mark.parametrize('precision', ['f64', 'f32'])
def test_contact_geometry_matches_the_closed_form(built, precision):
  cases = ka.geometry_cases()
  names = sorted(cases)
  env = _kat_env(built, len(names), precision, control_timestep=0.002)   # one substep per step: the probe sees the initial pose
  q0 = env.model['qpos0']
  _put(env, [ka.scene_state(q0, cases[n]['box'][0], cases[n]['box'][1], cases[n]['cap'][0], cases[n]['cap'][1]) for n in names])
  env.debug_contacts()
  env.step(torch.zeros(len(names), 6, device=DEV))
  got = env.debug_contacts()
  f32 = precision == 'f32'
  for e, n in enumerate(names):
    curved = n.startswith('capsule_end')   # EPA on the spherical cap ends by tolerance (1e-9 / 1e-6): normal to ~sqrt(2 tol / r)
    # (the contact probe reports float32 values: float64 runs are compared to 1e-7 / 1e-9, not to round-off)
    ka.check_contacts(got[e], cases[n],
                      pos_tol=(3e-3 if f32 else 1e-4) if curved else (2e-6 if f32 else 1e-7),
                      normal_tol=(2e-2 if f32 else 5e-4) if curved else (2e-6 if f32 else 1e-7),
                      dist_tol=(2e-6 if f32 else 2e-9) if curved else (2e-7 if f32 else 1e-9))
  assert env.counters()['contacts_dropped'] == 0
  env.close()


@pytest.mark.parametrize('precision', ['f64', 'f32'])
def test_rest_penetration_matches_the_closed_form(built, precision):
  """Env 0: box flat on the table (4 contacts); env 1: capsule on its side (2 contacts).  After 2 s the penetration solves
  n * imp^2 / (1 - imp) * m * K * d = m g (tests/kat_analytic.py rest_depth)."""
  env = _kat_env(built, 2, precision)
  q0 = env.model['qpos0']
  rows = [ka.scene_state(q0, (0.25, 0.0, ka.TABLE_TOP + ka.BOX_HALF[2])),
          ka.scene_state(q0, (0.25, -0.25, 0.7), cap_pos=(0.25, 0.0, ka.TABLE_TOP + ka.CAP_R), cap_quat=ka.quat_about((0, 1, 0), np.pi / 2))]
  rows[1][6:9] = (0.25, -0.2, ka.TABLE_TOP + ka.BOX_HALF[2])
  _put(env, rows)
  zero = torch.zeros(2, 6, device=DEV)
  for _ in range(100):
    env.step(zero)
  q, v = env.get_state(torch.float64)
  d_box = ka.TABLE_TOP + ka.BOX_HALF[2] - float(q[0, 8]); d_cap = ka.TABLE_TOP + ka.CAP_R - float(q[1, 15])
  print(f'{precision} rest penetration: box analytic {ka.rest_depth(ka.BOX_MASS, 4):.6e} measured {d_box:.6e}; '
        f'capsule analytic {ka.rest_depth(ka.CAP_MASS, 2):.6e} measured {d_cap:.6e}')
  tol = 1e-9 if precision == 'f64' else 2e-7
  assert abs(d_box - ka.rest_depth(ka.BOX_MASS, 4)) < tol
  assert abs(d_cap - ka.rest_depth(ka.CAP_MASS, 2)) < 5 * tol
  assert float(v[0, 6:12].abs().max()) < (1e-9 if precision == 'f64' else 1e-4)
  assert env.counters()['diverged'] == 0
  env.close()


@pytest.mark.parametrize('precision', ['f64', 'f32'])
def test_stacked_bodies_rest_at_the_closed_form_penetrations(built, precision):
  """Capsule on box on table: contact rows between two dynamic bodies and the load passed down (tests/kat_analytic.py
  stack_rest_depths: 7.7409e-05 m into the table under both weights, 1.7322e-05 m between the props)."""
  env = _kat_env(built, 1, precision)
  _put(env, [ka.stack_state(env.model['qpos0'])])
  zero = torch.zeros(1, 6, device=DEV)
  for _ in range(150):
    env.step(zero)
  q, v = env.get_state(torch.float64)
  zb, zc = float(q[0, 8]), float(q[0, 15])
  m_low, m_up = ka.TABLE_TOP + ka.BOX_HALF[2] - zb, (zb + ka.BOX_HALF[2]) - (zc - ka.CAP_R)
  d_low, d_up = ka.stack_rest_depths()
  print(f'{precision} stack rest penetrations: box/table analytic {d_low:.6e} measured {m_low:.6e}; capsule/box analytic {d_up:.6e} measured {m_up:.6e}')
  tol = 1e-9 if precision == 'f64' else 5e-7
  assert abs(m_low - d_low) < tol and abs(m_up - d_up) < tol
  assert float(v[0, 6:18].abs().max()) < (1e-9 if precision == 'f64' else 1e-4)
  assert int(env.debug_read('ncon').flatten()[0]) == 6 and env.counters()['diverged'] == 0
  env.close()


@pytest.mark.parametrize('precision', ['f64', 'f32'])
def test_sliding_box_decelerates_at_mu_g(built, precision):
  env = _kat_env(built, 2, precision, control_timestep=0.002)
  q0 = env.model['qpos0']
  _put(env, [ka.scene_state(q0, (0.1, 0.0, ka.TABLE_TOP + ka.BOX_HALF[2]))] * 2)
  zero = torch.zeros(2, 6, device=DEV)
  for _ in range(500):
    env.step(zero)
  q, v = env.get_state(torch.float64)
  v[:, 6] = 0.3
  env.set_state(q, v)
  for _ in range(3):
    env.step(zero)
  _, v0 = env.get_state(torch.float64)
  for _ in range(8):
    env.step(zero)
  _, v1 = env.get_state(torch.float64)
  a = float(v1[0, 6] - v0[0, 6]) / (8 * ka.DT)
  print(f'{precision} sliding deceleration: measured {-a:.5f}, mu g = {ka.MU * ka.G:.5f}')
  assert abs(-a / (ka.MU * ka.G) - 1) < 2e-2
  for _ in range(240):
    env.step(zero)
  _, v2 = env.get_state(torch.float64)
  assert abs(float(v2[0, 6])) < 1e-4
  env.close()


@pytest.mark.parametrize('precision', ['f64', 'f32'])
def test_free_fall_follows_semi_implicit_euler(built, precision):
  env = _kat_env(built, 2, precision, control_timestep=0.002)
  q0 = env.model['qpos0']
  _put(env, [ka.scene_state(q0, (0.25, 0.0, 0.8), cap_pos=(0.25, 0.3, 0.9))] * 2)
  zero = torch.zeros(2, 6, device=DEV)
  n = 30
  for _ in range(n):
    env.step(zero)
  q, v = env.get_state(torch.float64)
  tol = 1e-12 if precision == 'f64' else 1e-8
  assert abs(float(v[0, 8]) + ka.G * ka.DT * n) < tol * 100 and abs(float(q[0, 8]) - (0.8 - ka.G * ka.DT**2 * n * (n + 1) / 2)) < tol
  assert abs(float(q[0, 15]) - (0.9 - ka.G * ka.DT**2 * n * (n + 1) / 2)) < tol
  env.close()
