#!/usr/bin/env python3
"""Developer probe: per-stage clock64 breakdown of the scene step kernel (SO101_PROFILE=1).  GPU only.
usage: python tools/profile_stages.py [envs] [steps] [precision]"""
import json, os, sys
os.environ['SO101_PROFILE'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from so101_sim_b200.task_suite import create_batched_task_env

envs = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
prec = sys.argv[3] if len(sys.argv) > 3 else 'f32'
pre = int(sys.argv[4]) if len(sys.argv) > 4 else 0   # random-action control steps before the measured ones
env = create_batched_task_env('SO100HandOverBanana', num_envs=envs, time_limit=30.0, seed=0, device='cuda:0', precision=prec, placement='device', nursery_envs=0)
env.reset()
g = torch.Generator(device='cuda:0'); g.manual_seed(1)
spec = env.action_spec()
lo, hi = torch.tensor(spec.minimum, device='cuda:0'), torch.tensor(spec.maximum, device='cuda:0')
for _ in range(pre):
  env.step((lo + torch.rand(envs, 6, generator=g, device='cuda:0') * (hi - lo)) * 0.3)
acts = (lo + torch.rand(steps, envs, 6, generator=g, device='cuda:0') * (hi - lo)) * 0.3
p0 = env.debug_read('prof', 16)[0, :16].double().cpu() if False else None
names = ['dyn', 'broad', 'gjk_iters', 'gjk', 'epa', 'manifold', 'rows', 'solve', 'epa_iters', 'big_envs', 'npq', 'ncon', 'newton', 'line', 'nepa', 'nsub']
def read():
  out = torch.empty(16, dtype=torch.float32, device='cuda:0')
  import ctypes
  env._check(env._lib.so101_debug_read(env._h, b'prof', ctypes.c_void_p(out.data_ptr()), 16, env._stream()))
  return out.double().cpu()
a = read()
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for t in range(steps):
  env.step(acts[t])
t1.record(); torch.cuda.synchronize()
b = read() - a
nsub = float(b[15])
cyc = {n: float(b[i]) / nsub for i, n in enumerate(names[:10]) if n not in ('gjk_iters', 'epa_iters', 'big_envs')}
tot = sum(cyc.values())
res = dict(envs=envs, steps=steps, precision=prec, ms_per_step=t0.elapsed_time(t1) / steps, cycles_per_substep=tot,
           share={k: round(v / tot, 4) for k, v in cyc.items()}, cycles={k: round(v) for k, v in cyc.items()},
           per_substep=dict(pairs=float(b[10]) / nsub, contacts=float(b[11]) / nsub, newton_iters=float(b[12]) / nsub,
                            line_evals=float(b[13]) / nsub, epa_calls=float(b[14]) / nsub,
                            gjk_iters_per_pair=float(b[2]) / max(float(b[10]), 1), epa_iters_per_call=float(b[8]) / max(float(b[14]), 1), big_tier_env_fraction=float(b[9]) / nsub))
q, v = env.get_state()
ncon = env.debug_read('ncon').flatten(); it = env.debug_read('solver_iter').flatten()
res['ncon_last'] = dict(mean=float(ncon.mean()), max=float(ncon.max())); res['iter_last'] = dict(mean=float(it.mean()), max=float(it.max()))
res['counters'] = env.counters()
eh = torch.empty(8, dtype=torch.float32, device='cuda:0')
import ctypes as _ct
env._check(env._lib.so101_debug_read(env._h, b'epahist', _ct.c_void_p(eh.data_ptr()), 8, env._stream()))
res['epa_iter_hist(<=2,<=5,<=10,<=20,<=40,<cap,cap)'] = [int(x) for x in eh.tolist()[:7]]
npf = torch.empty(16, dtype=torch.float32, device='cuda:0')
env._check(env._lib.so101_debug_read(env._h, b'nprof', _ct.c_void_p(npf.data_ptr()), 16, env._stream()))
npf = npf.tolist()
res['narrow_warp_probe'] = dict(trips=int(npf[2]), mean_epa_cycles=npf[0] / max(npf[2], 1), mean_manifold_cycles=npf[1] / max(npf[2], 1), max_trip_cycles=npf[3], max_trip_geom2=int(npf[6]), max_epa_cycles=npf[4], max_manifold_cycles=npf[5], trip_cycles_hist_100k_200k_400k_800k_1600k_3200k=[int(x) for x in npf[8:15]])
h = torch.histc(ncon.float(), bins=13, min=0, max=104); res['ncon_hist_bins_of_8'] = [int(x) for x in h.tolist()]
print(json.dumps(res))
