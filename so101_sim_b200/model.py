"""Host-side reader for the compiled model blob (format written by tools/compile_model.py).

The blob is the flat, versioned model description of the SO100 scene
(reference: so101_sim/assets/so100/scene_pbr.xml + the two YCB props, so100_hand_over.py:159-206).
"""
from __future__ import annotations

import os
import struct

import numpy as np

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')
BLOB_MAGIC = b'SO1B'
BLOB_VERSION = 2


def blob_path(name: str) -> str:
  return name if os.path.exists(name) else os.path.join(DATA_DIR, name + '.blob')


def read_blob(path: str) -> dict[str, np.ndarray]:
  with open(blob_path(path), 'rb') as f:
    raw = f.read()
  magic, version, n, _ = struct.unpack_from('<4sIII', raw, 0)
  if magic != BLOB_MAGIC or version != BLOB_VERSION:
    raise ValueError(f'{path}: not a so101 model blob (magic={magic!r}, version={version})')
  out = {}
  for i in range(n):
    name, dt, cnt, off = struct.unpack_from('<24sIIQ', raw, 16 + 40 * i)
    out[name.rstrip(b'\0').decode()] = np.frombuffer(raw, dtype=np.float64 if dt == 0 else np.int32, count=cnt, offset=off).copy()
  return out
