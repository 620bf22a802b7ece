#!/usr/bin/env python3
"""CPU-only hunt (float64 oracle) for the episodes that end in divergence under the bench's random actions: finds the seeds,
then prints the substeps around the first anomalous prop acceleration.  Usage: cpu_diverge_hunt.py [n_seeds] [steps] | detail SEED"""
import os, sys, multiprocessing as mp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.oracle import OracleSim

def setup(seed):
  sim = OracleSim('so100_handover_banana', collide=True)
  rs = np.random.RandomState(seed)
  na = 6
  lo = np.array([-np.pi, -3.14158, -3.14158, -3.14158, -3.14158, 0.0]); hi = np.array([np.pi, 3.14158, 3.14158, 3.14158, 3.14158, 0.08])
  q = sim.meta['qpos0'].copy()
  u = rs.uniform(size=5)
  while np.hypot(-0.3 + 0.1 * u[3] + 0.1778, -0.1 + 0.2 * u[4] - 0.1656) < 0.151:
    u[3:5] = rs.uniform(size=2)
  yaw = (2 * u[2] - 1) * 0.1 * np.pi
  q[:na] = 0
  q[na:na + 7] = [0.2 + 0.1 * u[0], -0.1 + 0.2 * u[1], 0.45, np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)]
  q[na + 7:na + 14] = [-0.3 + 0.1 * u[3], -0.1 + 0.2 * u[4], 0.45, 1, 0, 0, 0]
  sim.set_state(q, np.zeros(sim.nv))
  for _ in range(50):
    sim.control_step(np.zeros(na))
  qs, vs = sim.qpos.copy(), sim.qvel.copy()
  qs[:na] = 0; vs[:na] = 0
  sim.set_state(qs, vs)
  acts = rs.uniform(lo, hi, size=(256, na)) * 0.3
  return sim, acts

def hunt(args):
  seed, steps = args
  sim, acts = setup(seed)
  for t in range(steps):
    sim.control_step(acts[t % 256])
    if sim.info('diverged') or np.abs(sim.qvel[6:]).max() > 100.0:
      return seed, t, bool(sim.info('diverged'))
  return seed, -1, False

def detail(seed, upto):
  sim, acts = setup(seed)
  names = sim.meta.get('geom_names')
  hist = []
  for t in range(upto + 1):
    sim.ctrl[:] = acts[t % 256]
    for k in range(10):
      sim.substep()
      cs = sim.contacts()
      dmin = min([c['dist'] for c in cs] + [0.0])
      hist.append((t, k, len(cs), dmin, np.abs(sim.qvel[:6]).max(), np.abs(sim.qvel[6:12]).max(), np.abs(sim.qvel[12:]).max(), np.abs(sim.field('qacc', 18)).max(), sim.info('solver_iter'),
                   [(c['geom1'], c['geom2'], round(c['dist'], 5)) for c in cs if c['dist'] < -0.004]))
  # first substep with a prop beyond 20 m/s
  first = next((i for i, h in enumerate(hist) if max(h[5], h[6]) > 20.0), len(hist) - 1)
  for h in hist[max(0, first - 25):first + 3]:
    print('step %3d.%d ncon %3d dmin %9.5f  |qd| arm %7.2f banana %9.2f bowl %9.2f  |qacc| %10.3g it %3d  deep %s' % h)

if __name__ == '__main__':
  if len(sys.argv) > 1 and sys.argv[1] == 'detail':
    detail(int(sys.argv[2]), int(sys.argv[3]))
  else:
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 800
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 110
    with mp.get_context('fork').Pool(8) as pool:
      res = pool.map(hunt, [(1000 + i, steps) for i in range(n)], chunksize=4)
    bad = [r for r in res if r[1] >= 0]
    print(len(bad), 'of', n, 'episodes launch a prop beyond 100 m/s or diverge within', steps, 'steps:', bad)
