import sys, time, torch
sys.path.insert(0, '/root/repo')
from so101_sim_b200.task_suite import create_batched_task_env
dev = 'cuda:0'
env = create_batched_task_env('SO100HandOverBanana', num_envs=16384, time_limit=2.0, seed=0, device=dev)   # 2 s episodes: many auto-resets
env.randomize_resets(rounds=3, seed=0, settle_steps=25)
g = torch.Generator(device=dev); g.manual_seed(1)
spec = env.action_spec()
lo, hi = torch.tensor(spec.minimum, device=dev), torch.tensor(spec.maximum, device=dev)
nfirst = nlast = 0; rsum = 0.0
t0 = time.time()
for t in range(400):
  a = (lo + torch.rand(16384, 6, generator=g, device=dev) * (hi - lo)) * 0.3
  ts = env.step(a)
  if t % 50 == 49:
    torch.cuda.synchronize()
    st = ts.step_type
    q = ts.observation['physics_state']
    print(t, 'first', int((st == 0).sum()), 'last', int((st == 2).sum()), 'finite', bool(torch.isfinite(q).all()), 'reward sum', float(ts.reward.sum()), env.counters(), flush=True)
torch.cuda.synchronize(); print('wall', time.time() - t0)
env.close()
for prec in ('f64',):
  env = create_batched_task_env('SO100HandOverBanana', num_envs=1024, time_limit=30.0, seed=0, device=dev, precision=prec)
  env.sample_prop_initial_states(seed=0, settle_steps=10)
  t0 = time.time()
  for t in range(20):
    ts = env.step((lo + torch.rand(1024, 6, generator=g, device=dev) * (hi - lo)) * 0.3)
  torch.cuda.synchronize(); print(prec, '20 steps x 1024 envs', time.time() - t0, 's', env.counters(), bool(torch.isfinite(ts.observation['physics_state']).all()))
  env.close()
