"""Closed-form known answers for the contact pipeline, shared by the oracle tests (tests/test_oracle_analytic.py, CPU) and their
GPU twins (tests/test_scene_gpu.py).  Nothing here calls the oracle or the CUDA path: the numbers follow from the scene's
parameters and [upstream] MuJoCo's documented constraint model:

  solref (tc, dr) -> K = 1 / (dmax^2 tc^2 dr^2), B = 2 / (dmax tc), tc >= 2 dt
  solimp (d0, dmax, width, mid, power) -> impedance imp(depth): power-law sigmoid of depth / width between d0 and dmax
  aref = -B v - K imp (dist - margin);  R = (1 - imp) / imp * (invweight0 of the two bodies);  force = -D (J qacc - aref), D = 1 / R
  contact parameter mixing: condim = max, friction = element-wise max, solref / solimp = solmix-weighted mean (equal priority)
  contact record: dist = -(penetration), pos = midpoint between the two surfaces, frame x = normal from geom1 to geom2

The scene is tests/golden/kat_primitives.blob (tools/make_kat_blob.py): the reference's scene_pbr.xml plus a free box
(half sizes 0.03 x 0.02 x 0.015) and a free capsule (radius 0.015, half length 0.04), density 200, with the YCB props'
collision class (condim 6, friction 1 0.01 0.01, solref 0.004 1)."""
import os

import numpy as np

BLOB = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'kat_primitives.blob')
G = 9.81
DT = 0.002
TABLE_TOP = 0.42          # table body z 0.4 + half thickness 0.02 (scene_pbr.xml:132-133)
TABLE_GEOM, FLOOR_GEOM, BOX_GEOM, CAPSULE_GEOM = 19, 0, 26, 27
BOX_HALF = np.array([0.03, 0.02, 0.015])
CAP_R, CAP_H = 0.015, 0.04
DENSITY = 200.0
BOX_MASS = DENSITY * 8 * BOX_HALF.prod()
CAP_MASS = DENSITY * (np.pi * CAP_R**2 * 2 * CAP_H + 4.0 / 3.0 * np.pi * CAP_R**3)
MU = 1.0                  # max(prop friction 1.0, table / floor friction 1.0)
# mixed solref of a prop geom (0.004, 1) and a default geom (0.02, 1) with equal solmix; default solimp on both
SOLREF = (0.5 * (0.004 + 0.02), 1.0)
SOLIMP = (0.9, 0.95, 0.001, 0.5, 2.0)


def impedance(depth):
  d0, dmax, width, mid, power = SOLIMP
  x = min(abs(depth) / width, 1.0)
  if x <= mid:
    y = (1.0 / mid) ** (power - 1) * x ** power
  else:
    y = 1.0 - (1.0 / (1.0 - mid)) ** (power - 1) * (1.0 - x) ** power
  return d0 + y * (dmax - d0)


def stiffness():
  tc = max(SOLREF[0], 2 * DT)
  dmax = SOLIMP[1]
  return 1.0 / (dmax**2 * tc**2 * SOLREF[1]**2)


def rest_depth(mass, ncontacts):
  """Penetration at which `ncontacts` equally loaded contacts carry m g:  n * D * K * imp * d = m g with D = imp / ((1 - imp) / m)."""
  K = stiffness()
  f = lambda d: ncontacts * impedance(d)**2 / (1.0 - impedance(d)) * mass * K * d - mass * G
  lo, hi = 0.0, 1e-3
  for _ in range(200):
    mid = 0.5 * (lo + hi)
    lo, hi = (mid, hi) if f(mid) < 0 else (lo, mid)
  return 0.5 * (lo + hi)


def rest_depth_loaded(load_mass, ncontacts, invweight, tc):
  """General form: `ncontacts` equally loaded contacts between bodies whose translational invweight0 sum to `invweight`, mixed
  solref time constant `tc`, carrying load_mass * g:  n * imp^2 / (1 - imp) * K(tc) * d / invweight = load_mass * g."""
  K = 1.0 / (SOLIMP[1]**2 * max(tc, 2 * DT)**2)
  f = lambda d: ncontacts * impedance(d)**2 / (1.0 - impedance(d)) * K * d / invweight - load_mass * G
  lo, hi = 0.0, 1e-3
  for _ in range(200):
    mid = 0.5 * (lo + hi)
    lo, hi = (mid, hi) if f(mid) < 0 else (lo, mid)
  return 0.5 * (lo + hi)


def stack_state(qpos0):
  """The capsule lying along x on top of the box, which lies flat on the table: the capsule (half extent 0.055) overhangs the
  box top (half extent 0.03), so the box / capsule contact is the capsule's bottom line clipped to the face: two points at x =
  +-0.03, each carrying m_cap g / 2; the four box / table contacts carry (m_box + m_cap) g."""
  hz = BOX_HALF[2]
  return scene_state(qpos0, (0.25, 0.0, TABLE_TOP + hz), cap_pos=(0.25, 0.0, TABLE_TOP + 2 * hz + CAP_R), cap_quat=quat_about((0, 1, 0), np.pi / 2))


def stack_rest_depths():
  """(box into table, capsule into box): the lower contacts mix solref (0.004 + 0.02) / 2 and see the box's invweight only (the
  table is static); the upper contacts are prop against prop: solref 0.004 (= 2 dt), invweight 1 / m_box + 1 / m_cap."""
  return (rest_depth_loaded(BOX_MASS + CAP_MASS, 4, 1.0 / BOX_MASS, SOLREF[0]),
          rest_depth_loaded(CAP_MASS, 2, 1.0 / BOX_MASS + 1.0 / CAP_MASS, 0.004))


def quat_about(axis, angle):
  a = np.asarray(axis, dtype=np.float64); a = a / np.linalg.norm(a)
  return np.concatenate([[np.cos(angle / 2)], np.sin(angle / 2) * a])


def quat_to_mat(q):
  w, x, y, z = q / np.linalg.norm(q)
  return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                   [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                   [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def scene_state(qpos0, box_pos, box_quat=(1, 0, 0, 0), cap_pos=(0.25, 0.2, 0.6), cap_quat=(1, 0, 0, 0)):
  """qpos (20) with the arm at 0 and the two props placed; the prop that is not under test floats far above the table."""
  q = np.array(qpos0, dtype=np.float64).copy()
  q[:6] = 0
  q[6:9] = box_pos; q[9:13] = np.asarray(box_quat, dtype=np.float64) / np.linalg.norm(box_quat)
  q[13:16] = cap_pos; q[16:20] = np.asarray(cap_quat, dtype=np.float64) / np.linalg.norm(cap_quat)
  return q


def box_lowest_points(center, quat, tol=1e-9):
  """World corners of the box that are lowest (within tol) - the analytic support feature along -z - and their height."""
  R = quat_to_mat(np.asarray(quat, dtype=np.float64))
  corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]) * BOX_HALF
  w = corners @ R.T + np.asarray(center)
  zmin = w[:, 2].min()
  return w[w[:, 2] <= zmin + tol], zmin


# geometry cases: name -> (box centre, box quaternion, capsule centre, capsule quaternion, expected contacts) where expected =
# list of (geom1, geom2, points [k,3] on the lower surface, depth) for the pairs that must appear, built in geometry_cases()
def geometry_cases():
  cases = {}
  depth = 5e-4
  far_cap, far_box = (0.25, 0.25, 0.7), (0.25, -0.25, 0.7)
  # (a) box flat on the table: 4 corner contacts
  c = np.array([0.25, 0.0, TABLE_TOP + BOX_HALF[2] - depth])
  cases['box_flat_on_table'] = dict(box=(c, (1, 0, 0, 0)), cap=(far_cap, (1, 0, 0, 0)), pair=(TABLE_GEOM, BOX_GEOM), surface=TABLE_TOP, depth=depth,
                                    points=box_lowest_points(c, (1, 0, 0, 0))[0])
  # (b) box on an edge: rotated 45 degrees about x, the lowest edge penetrates by `depth`
  q = quat_about((1, 0, 0), np.pi / 4)
  low = (BOX_HALF[1] + BOX_HALF[2]) * np.sqrt(0.5)
  c = np.array([0.25, 0.0, TABLE_TOP + low - depth])
  cases['box_edge_on_table'] = dict(box=(c, q), cap=(far_cap, (1, 0, 0, 0)), pair=(TABLE_GEOM, BOX_GEOM), surface=TABLE_TOP, depth=depth,
                                    points=box_lowest_points(c, q)[0])
  # (c) box on a vertex: generic rotation, one contact
  q = quat_about((1, 0.3, 0), 0.9)
  pts, zmin = box_lowest_points(np.zeros(3), q)
  c = np.array([0.25, 0.0, TABLE_TOP - zmin - depth])
  cases['box_vertex_on_table'] = dict(box=(c, q), cap=(far_cap, (1, 0, 0, 0)), pair=(TABLE_GEOM, BOX_GEOM), surface=TABLE_TOP, depth=depth,
                                      points=box_lowest_points(c, q)[0])
  # (d) box flat on the floor plane (plane path), away from the table
  c = np.array([1.5, 1.5, BOX_HALF[2] - depth])
  cases['box_flat_on_floor'] = dict(box=(c, (1, 0, 0, 0)), cap=(far_cap, (1, 0, 0, 0)), pair=(FLOOR_GEOM, BOX_GEOM), surface=0.0, depth=depth,
                                    points=box_lowest_points(c, (1, 0, 0, 0))[0])
  # (e) capsule lying on the table (axis along x): two contacts under the segment's end points
  q = quat_about((0, 1, 0), np.pi / 2)
  c = np.array([0.25, 0.0, TABLE_TOP + CAP_R - depth])
  ends = np.array([[c[0] - CAP_H, c[1], c[2] - CAP_R], [c[0] + CAP_H, c[1], c[2] - CAP_R]])
  cases['capsule_side_on_table'] = dict(box=(far_box, (1, 0, 0, 0)), cap=(c, q), pair=(TABLE_GEOM, CAPSULE_GEOM), surface=TABLE_TOP, depth=depth, points=ends)
  # (f) capsule standing on one end (axis along z): one contact under the lower cap
  c = np.array([0.25, 0.0, TABLE_TOP + CAP_R + CAP_H - depth])
  cases['capsule_end_on_table'] = dict(box=(far_box, (1, 0, 0, 0)), cap=(c, (1, 0, 0, 0)), pair=(TABLE_GEOM, CAPSULE_GEOM), surface=TABLE_TOP, depth=depth,
                                       points=np.array([[c[0], c[1], c[2] - CAP_R - CAP_H]]))
  # (g) capsule lying on the floor plane
  q = quat_about((0, 1, 0), np.pi / 2)
  c = np.array([1.5, 1.5, CAP_R - depth])
  ends = np.array([[c[0] - CAP_H, c[1], c[2] - CAP_R], [c[0] + CAP_H, c[1], c[2] - CAP_R]])
  cases['capsule_side_on_floor'] = dict(box=(far_box, (1, 0, 0, 0)), cap=(c, q), pair=(FLOOR_GEOM, CAPSULE_GEOM), surface=0.0, depth=depth, points=ends)
  return cases


def check_contacts(contacts, case, pos_tol, normal_tol, dist_tol):
  """contacts: list of (geom1, geom2, dist, pos[3], normal[3]).  Exactly the expected points, each once, with normal +z (from the
  static lower geom to the prop), dist = -depth and pos halfway between the prop's lowest points and the supporting surface."""
  mine = [c for c in contacts if (c[0], c[1]) == case['pair']]
  others = [c for c in contacts if (c[0], c[1]) != case['pair']]
  assert not others, f'unexpected contacts {[(c[0], c[1]) for c in others]}'
  pts = np.asarray(case['points'])
  assert len(mine) == len(pts), (len(mine), len(pts))
  used = set()
  for g1, g2, dist, pos, normal in mine:
    assert abs(dist + case['depth']) < dist_tol, (dist, case['depth'])
    assert np.abs(np.asarray(normal) - np.array([0, 0, 1.0])).max() < normal_tol, normal
    want_z = case['surface'] - 0.5 * case['depth']
    k = int(np.argmin(np.abs(pts[:, :2] - np.asarray(pos)[:2]).sum(1)))
    assert k not in used
    used.add(k)
    assert np.abs(pts[k, :2] - np.asarray(pos)[:2]).max() < pos_tol and abs(pos[2] - want_z) < pos_tol, (pos, pts[k], want_z)
