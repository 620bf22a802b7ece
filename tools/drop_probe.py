import sys, torch, ctypes
sys.path.insert(0, '/root/repo')
from so101_sim_b200.task_suite import create_batched_task_env
dev='cuda:0'
for N in (4096, 16384):
  env = create_batched_task_env('SO100HandOverBanana', num_envs=N, time_limit=30.0, seed=0, device=dev)
  env.sample_prop_initial_states(seed=0, settle_steps=25)
  g = torch.Generator(device=dev); g.manual_seed(1)
  spec = env.action_spec(); lo, hi = torch.tensor(spec.minimum, device=dev), torch.tensor(spec.maximum, device=dev)
  d0 = env.debug_read('dropcat', 8)[0, :8].tolist() if False else None
  out = torch.empty(8, dtype=torch.float32, device=dev)
  def dc():
    env._check(env._lib.so101_debug_read(env._h, b'dropcat', ctypes.c_void_p(out.data_ptr()), 8, env._stream())); return out.tolist()
  a0 = dc()
  for t in range(200):
    ts = env.step((lo + torch.rand(N, 6, generator=g, device=dev) * (hi - lo)) * 0.3)
    if t in (49, 99, 199):
      a1 = dc(); ncon = env.debug_read('ncon').flatten()
      print(N, t, 'dropcat delta [NOUT,CAND,PAIR,WORKQ,CONBUF,BLOCK,HIT]', [int(x - y) for x, y in zip(a1, a0)][:7], env.counters(), 'ncon mean/max', float(ncon.mean()), float(ncon.max()), flush=True)
  env.close()
