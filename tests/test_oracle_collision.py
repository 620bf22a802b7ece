"""CPU checks of the collision pieces the oracle and the CUDA path share: model data (hull vertex graphs) and geometric
invariants of the oracle's contacts (the contact pipeline is unpinned against MuJoCo)."""
import json
import os

import numpy as np
import pytest

from oracle.oracle import OracleSim
from so101_sim_b200.model import read_blob

KAT1 = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'kat1_so101_rl.json')))


@pytest.fixture(scope='module')
def model():
  return read_blob('so100_handover_banana')


def test_hill_climbing_on_the_hull_graphs_finds_the_exhaustive_support(model):
  """Steepest-ascent hill climbing over hull_nbr (what support() does for hulls >= 40 vertices, oracle and CUDA) must reach
  the exhaustive maximum from ANY start vertex: the graphs are connected and the hulls convex."""
  verts = model['hull_vert'].reshape(-1, 3)
  adr, nbr = model['hull_nbradr'], model['hull_nbr']
  rs = np.random.RandomState(0)
  checked = 0
  for g in range(int(np.asarray(model['ngeom']).reshape(-1)[0])):
    if model['geom_type'][g] != 5 or model['geom_vertnum'][g] < 40:
      continue
    a, n = int(model['geom_vertadr'][g]), int(model['geom_vertnum'][g])
    V = verts[a:a + n]
    for _ in range(6):
      d = rs.normal(size=3)
      val = V @ d
      cur = int(rs.randint(n))
      for _guard in range(n):
        nb = nbr[adr[a + cur]:adr[a + cur + 1]]
        assert nb.min() >= 0 and nb.max() < n
        j = nb[np.argmax(val[nb])]
        if val[j] > val[cur]:
          cur = int(j)
        else:
          break
      assert val[cur] >= val.max() - 1e-12 * max(1.0, abs(val.max())), (g, n)
      checked += 1
  assert checked > 300


def test_resting_contacts_are_flat_on_the_table():
  """From the reference notebook's own resting state (KAT-1 pre-step state): every prop-table contact has the table normal,
  sits at the table top (z = 0.42) and penetrates by well under a millimetre; no more than 4 points per geom pair."""
  s = OracleSim('so100_handover_banana', collide=True)
  st = np.array(KAT1['delayed_physics_state'])
  s.set_state(st[:20], st[20:]); s.forward()
  con = s.contacts()
  assert 8 <= len(con) <= 64
  per_pair = {}
  for c in con:
    n = np.asarray(c['frame'][0])
    assert abs(np.linalg.norm(n) - 1) < 1e-12 and c['dist'] < 0
    per_pair[(c['geom1'], c['geom2'])] = per_pair.get((c['geom1'], c['geom2']), 0) + 1
    if abs(c['pos'][2] - 0.42) < 2e-3 and n[2] > 0.99:  # table-top contact
      assert -1e-3 < c['dist'] and np.abs(n - [0, 0, 1]).max() < 1e-9
  assert max(per_pair.values()) <= 4
  # frames are right-handed orthonormal triads
  F = np.asarray(con[0]['frame']).reshape(3, 3)
  np.testing.assert_allclose(F @ F.T, np.eye(3), atol=1e-12)
  assert np.linalg.det(F) > 0.999


def test_props_come_to_rest_near_the_reference_rest_heights():
  """Dropped 3 mm above the table with identity orientation the props settle at the rest heights the reference prints (banana
  z 0.421711, bowl z 0.422622, so101_rl.ipynb:221-223): the bowl within 2.3e-6 m (bar 1e-5), the banana - whose printed rest
  orientation is not the identity - within 2.0e-4 m (bar 3e-4), with small residual velocity."""
  s = OracleSim('so100_handover_banana', collide=True)
  q = s.meta['qpos0'].copy()
  q[:6] = 0
  q[6:13] = [0.25, 0.0, 0.4217 + 0.003, 1, 0, 0, 0]
  q[13:20] = [-0.25, -0.05, 0.4226 + 0.003, 1, 0, 0, 0]
  s.set_state(q, np.zeros(18))
  for _ in range(40):
    s.control_step(np.zeros(6))
  assert abs(s.qpos[8] - 0.421711) < 3e-4 and abs(s.qpos[15] - 0.422622) < 1e-5
  assert np.abs(s.qvel[6:9]).max() < 5e-3 and np.abs(s.qvel[12:15]).max() < 5e-3


def test_pen_in_the_utensil_holder_triggers_the_two_box_reward():
  """SO100HandOverPen (so100_hand_over.py:97-117): the reward needs the pen's box to overlap BOTH container boxes (inside the
  holder and above it, :104-116,263-273) with both props at rest; a pen lying on the table overlaps neither."""
  m = read_blob('so100_handover_pen')
  assert m['reward_box_pos'].size == 6 and m['reward_box_half'].size == 6 and int(m['prop_mass_standin'][0]) == 0
  np.testing.assert_allclose(m['reward_box_pos'].reshape(2, 3), np.array([[0, 0, 0.02666], [0, 0, 0.25]]) * 0.6)
  q = m['qpos0'].copy()
  q[:6] = 0
  q[13:20] = [-0.25, 0.05, 0.4503, 1, 0, 0, 0]
  s = OracleSim('so100_handover_pen', collide=True)
  q[6:13] = [-0.25, 0.05, 0.4503 + 0.105, np.cos(np.pi / 4), np.sin(np.pi / 4), 0, 0]  # long axis (body y) up, inside the holder
  s.set_state(q, np.zeros(18))
  rewards = [s.control_step(np.zeros(6)) for _ in range(12)]
  assert rewards[0] == 0.0 and max(rewards) == 1.0  # moving at first, then at rest inside both boxes
  s = OracleSim('so100_handover_pen', collide=True)
  q[6:13] = [0.25, 0.0, 0.4283, 1, 0, 0, 0]
  s.set_state(q, np.zeros(18))
  assert max(s.control_step(np.zeros(6)) for _ in range(12)) == 0.0
