"""ctypes binding + env-level restatement for the CPU oracle — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
The env-level class `OracleEnv` restates, for ONE environment, what dm_control's composer.Environment does around
the physics for the SO100 task (reference so100_task.py:266-368, task_suite.py:148-155).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, '_build', 'libso101_oracle.so')
DATA_DIR = os.path.join(_HERE, '..', 'so101_sim_b200', 'data')


def build(force=False):
  src = [os.path.join(_HERE, f) for f in ('so101_oracle.c', 'so101_collide.c', 'so101_oracle.h')]
  if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src):
    subprocess.check_call(['make', '-C', _HERE, '-s'])
  return _LIB_PATH


_lib = None


def lib():
  global _lib
  if _lib is None:
    L = ctypes.CDLL(build())
    L.so_model_load.restype = ctypes.c_void_p; L.so_model_load.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
    L.so_model_free.argtypes = [ctypes.c_void_p]
    L.so_data_new.restype = ctypes.c_void_p; L.so_data_new.argtypes = [ctypes.c_void_p]
    L.so_data_free.argtypes = [ctypes.c_void_p]
    L.so_reset.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.so_forward_position.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.so_substep.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.so_control_step.restype = ctypes.c_double
    L.so_control_step.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    L.so_reward.restype = ctypes.c_double; L.so_reward.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.so_field.restype = ctypes.POINTER(ctypes.c_double)
    L.so_field.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int)]
    L.so_info.restype = ctypes.c_int; L.so_info.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    L.so_set_collide.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.so_set_integrator.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.so_get_contact.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    L.so_overlap_oobb_oobb.restype = ctypes.c_int
    L.so_overlap_oobb_oobb.argtypes = [ctypes.c_void_p] * 6
    _lib = L
  return _lib


def _dp(a):
  return a.ctypes.data_as(ctypes.c_void_p)


def overlap_oobb_oobb(p0, q0, h0, p1, q1, h1) -> bool:
  arrs = [np.ascontiguousarray(x, dtype=np.float64) for x in (p0, q0, h0, p1, q1, h1)]
  return bool(lib().so_overlap_oobb_oobb(*[_dp(a) for a in arrs]))


class OracleSim:
  """One float64 physics instance (model + data)."""

  def __init__(self, model='so100_handover_banana', collide=True, integrator='euler'):
    path = model if os.path.exists(model) else os.path.join(DATA_DIR, model + '.blob')
    with open(path, 'rb') as f:
      self._blob = f.read()
    L = lib()
    self._m = L.so_model_load(self._blob, len(self._blob))
    if not self._m:
      raise RuntimeError('so_model_load failed')
    self._d = L.so_data_new(self._m)
    L.so_set_collide(self._d, int(collide))
    L.so_set_integrator(self._d, int(integrator == 'implicitfast'))
    from so101_sim_b200.model import read_blob  # host-side blob reader (no CUDA involved)
    self.meta = read_blob(path)
    self.nq, self.nv, self.nu, self.nbody = (int(self.meta[k][0]) for k in ('nq', 'nv', 'nu', 'nbody'))

  def __del__(self):
    try:
      lib().so_data_free(self._d); lib().so_model_free(self._m)
    except Exception:
      pass

  def field(self, name, n=None):
    c = ctypes.c_int(0)
    p = lib().so_field(self._d, name.encode(), ctypes.byref(c))
    if not p:
      raise KeyError(name)
    a = np.ctypeslib.as_array(p, shape=(c.value,))
    return a if n is None else a[:n]

  def info(self, name):
    return lib().so_info(self._d, name.encode())

  @property
  def qpos(self): return self.field('qpos', self.nq)
  @property
  def qvel(self): return self.field('qvel', self.nv)
  @property
  def ctrl(self): return self.field('ctrl', self.nu)
  @property
  def time(self): return float(self.field('time')[0])

  def reset(self):
    lib().so_reset(self._m, self._d)

  def set_state(self, qpos, qvel):
    self.qpos[:] = qpos; self.qvel[:] = qvel

  def forward(self):
    lib().so_forward_position(self._m, self._d)

  def substep(self):
    lib().so_substep(self._m, self._d)

  def control_step(self, action, offsets=None, nsub=10) -> float:
    a = np.ascontiguousarray(action, dtype=np.float64)
    o = np.zeros(self.nu) if offsets is None else np.ascontiguousarray(offsets, dtype=np.float64)
    return float(lib().so_control_step(self._m, self._d, _dp(a), _dp(o), nsub))

  def reward(self) -> float:
    return float(lib().so_reward(self._m, self._d))

  def contacts(self):
    out = []
    buf = np.zeros(30)
    for c in range(self.info('ncon')):
      lib().so_get_contact(self._d, c, _dp(buf))
      out.append(dict(dist=buf[0], pos=buf[1:4].copy(), frame=buf[4:13].reshape(3, 3).copy(), dim=int(buf[13]), geom1=int(buf[14]),
                      geom2=int(buf[15]), mu=buf[16], friction=buf[17:22].copy(), solref=buf[22:24].copy(), solimp=buf[24:29].copy(),
                      efc_address=int(buf[29])))
    return out
