import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.oracle import OracleSim
from so101_sim_b200.task_suite import create_batched_task_env
dev = 'cuda:0'
env = create_batched_task_env('SO100TwoArmHandOverBanana', num_envs=2, time_limit=30.0, seed=0, device=dev, reset_rounds=0, precision='f64', control_timestep=0.002)
rs = np.random.RandomState(3)
q = np.tile(np.asarray(env.model['qpos0'], dtype=np.float64), (2, 1)); q[:, :12] = 0
for e in range(2):
  yaw = rs.uniform(-0.3, 0.3)
  q[e, 12:19] = [rs.uniform(0.2, 0.3), rs.uniform(-0.1, 0.1), 0.4217 + 0.002, np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)]
  q[e, 19:26] = [rs.uniform(-0.3, -0.22), rs.uniform(-0.1, -0.02), 0.4226 + 0.002, 1, 0, 0, 0]
q0 = torch.tensor(q); v0 = torch.zeros(2, 24, dtype=torch.float64)
env.set_initial_state(q0, v0); env.reset()
g = torch.Generator(device=dev); g.manual_seed(4)
spec = env.action_spec()
lo, hi = torch.tensor(spec.minimum, device=dev), torch.tensor(spec.maximum, device=dev)
acts = (lo + torch.rand(25, 2, 12, generator=g, device=dev) * (hi - lo)) * 0.3
sims = []
for e in range(2):
  o = OracleSim('so100_twoarm_banana', collide=True); o.set_state(q[e], np.zeros(24)); sims.append(o)
env.debug_contacts()
done = False
for t in range(250):
  a = acts[t // 10]
  env.step(a)
  got = env.debug_contacts()
  qq, vv = env.get_state(torch.float64); qq = qq.cpu().numpy()
  for e, o in enumerate(sims):
    o.ctrl[:] = a[e].double().cpu().numpy()
    o.forward(); ref = o.contacts()   # contacts at the state the GPU substep saw
    o.substep()
    d = np.abs(qq[e] - o.qpos)
    if d.max() > 1e-7 and not done:
      done = True
      print('FIRST mismatch substep', t, 'env', e, 'arm err', d[:12].max(), 'prop err', d[12:].max(), 'ncon gpu', len(got[e]), 'oracle', len(ref), env.counters())
      gp = [(c[0], c[1]) for c in got[e]]; rp = [(c['geom1'], c['geom2']) for c in ref]
      print(' gpu pairs', gp); print(' ora pairs', rp)
      for cg, cr in zip(got[e], ref):
        if (cg[0], cg[1]) != (cr['geom1'], cr['geom2']) or abs(cg[2] - cr['dist']) > 1e-6: print('  diff', cg[:3], (cr['geom1'], cr['geom2'], cr['dist']))
      print(' gpu dq', np.round(qq[e] - o.qpos, 9))
  if t % 25 == 24: print('substep', t, 'max err', max(np.abs(qq[e] - sims[e].qpos).max() for e in range(2)), 'ncon', [len(x) for x in got])

# ---- isolate: state just before the first mismatching substep, one env, one substep
print('---- isolate')
o = OracleSim('so100_twoarm_banana', collide=True); o.set_state(q[0], np.zeros(24))
for t in range(81):
  o.ctrl[:] = acts[t // 10][0].double().cpu().numpy(); o.substep()
o.ctrl[:] = acts[8][0].double().cpu().numpy()
qs, vs = o.qpos.copy(), o.qvel.copy()
o.forward()
ref = [c for c in o.contacts() if (c['geom1'], c['geom2']) in ((13, 23), (14, 23))]
for c in ref: print(' oracle', c['geom1'], c['geom2'], c['dist'], c['pos'], c['frame'][0])
for prec in ('f64', 'f32'):
  e2 = create_batched_task_env('SO100TwoArmHandOverBanana', num_envs=1, time_limit=30.0, seed=0, device=dev, reset_rounds=0, precision=prec, control_timestep=0.002)
  e2.set_initial_state(torch.tensor(qs)[None], torch.tensor(vs)[None]); e2.reset()
  e2.debug_contacts(); e2.step(acts[8][:1]); got = e2.debug_contacts()[0]
  for c in got:
    if (c[0], c[1]) in ((13, 23), (14, 23)): print(' gpu', prec, c[0], c[1], c[2], c[3], c[4])
  e2.close()
m = env.model
for g in (13, 23): print('geom', g, 'type', m['geom_type'][g], 'nvert', m['geom_vertnum'][g], 'body', m['geom_body'][g])
print('---- batch variants')
o1 = OracleSim('so100_twoarm_banana', collide=True); o1.set_state(q[1], np.zeros(24))
for t in range(81):
  o1.ctrl[:] = acts[t // 10][1].double().cpu().numpy(); o1.substep()
q1s, v1s = o1.qpos.copy(), o1.qvel.copy()
for name, Q, V in (('same x2', [qs, qs], [vs, vs]), ('env0+env1', [qs, q1s], [vs, v1s]), ('env1+env0', [q1s, qs], [v1s, vs])):
  e2 = create_batched_task_env('SO100TwoArmHandOverBanana', num_envs=2, time_limit=30.0, seed=0, device=dev, reset_rounds=0, precision='f64', control_timestep=0.002)
  e2.set_initial_state(torch.tensor(np.stack(Q)), torch.tensor(np.stack(V))); e2.reset()
  e2.debug_contacts(); e2.step(acts[8]); got = e2.debug_contacts()
  for e in range(2):
    print(' ', name, 'env', e, 'ncon', len(got[e]), [(c[0], c[1], round(c[2], 7)) for c in got[e] if (c[0], c[1]) == (13, 23)], e2.counters()['contacts_dropped'])
  e2.close()
