#!/usr/bin/env python3
"""Model compiler: MJCF scene (+ free props) -> flat binary "model blob".

Runs ONCE, on the host, in a container that has the reference checkout.  The
blob it writes (so101_sim_b200/data/*.blob) is what travels to the GPU box; no
test, bench or product code reads /root/reference at run time.

What it restates (reference file:line):
  * scene_pbr.xml:1-162            robot + table scene (defaults, bodies, actuators)
  * so100_task.py:151-152          multiccd on, noslip 0
  * so100_hand_over.py:159-206     two free props (banana, bowl x1.5), freejoint
                                   removed, attached with add_free_entity
  * ycb/011_banana/.../model.xml:37-56, ycb/024_bowl/.../model.xml:86-158
[upstream] MuJoCo compile-time behaviour restated here: default-class
resolution, mesh -> convex hull (qhull via scipy), inertia from density,
invweight0 / meaninertia constants (mj_setConst), contact-pair static filtering.

The visual meshes `textured.obj` that define prop mass are absent from the
reference checkout (.MISSING_LARGE_BLOBS:2-3); `coacd_merged.obj` stands in
(same surface, convex-decomposed).  This is recorded in the blob (`prop_mass_standin`).
"""
from __future__ import annotations

import argparse
import os
import struct
import sys
import xml.etree.ElementTree as ET

import numpy as np
from scipy.spatial import ConvexHull

# ----------------------------------------------------------------------------------------------
# geom / joint type codes shared with oracle/so101_oracle.c and csrc/model.h
GEOM_PLANE, GEOM_SPHERE, GEOM_CAPSULE, GEOM_CYLINDER, GEOM_BOX, GEOM_HULL = 0, 1, 2, 3, 4, 5
JNT_FREE, JNT_HINGE = 0, 1

GEOM_DEFAULTS = dict(type='sphere', contype='1', conaffinity='1', condim='3', friction='1 0.005 0.0001',
                     solref='0.02 1', solimp='0.9 0.95 0.001 0.5 2', margin='0', gap='0', solmix='1',
                     priority='0', density='1000', pos='0 0 0', quat='1 0 0 0')
JOINT_DEFAULTS = dict(type='hinge', pos='0 0 0', axis='0 0 1', armature='0', frictionloss='0', damping='0',
                      stiffness='0', solreflimit='0.02 1', solimplimit='0.9 0.95 0.001 0.5 2',
                      solreffriction='0.02 1', solimpfriction='0.9 0.95 0.001 0.5 2', margin='0')
GENERAL_DEFAULTS = dict(gainprm='1 0 0', biasprm='0 0 0', biastype='none', gaintype='fixed', gear='1')


def fvec(s, n=None, fill=None):
  v = np.array([float(x) for x in s.split()], dtype=np.float64)
  if n is not None and len(v) < n:
    v = np.concatenate([v, np.asarray(fill, dtype=np.float64)[len(v):n]])
  return v


# ----------------------------------------------------------------------------------------------
# quaternion helpers (w, x, y, z)
def qmul(a, b):
  return np.array([a[0]*b[0]-a[1]*b[1]-a[2]*b[2]-a[3]*b[3],
                   a[0]*b[1]+a[1]*b[0]+a[2]*b[3]-a[3]*b[2],
                   a[0]*b[2]-a[1]*b[3]+a[2]*b[0]+a[3]*b[1],
                   a[0]*b[3]+a[1]*b[2]-a[2]*b[1]+a[3]*b[0]])


def q2m(q):
  q = q/np.linalg.norm(q)
  w, x, y, z = q
  return np.array([[1-2*(y*y+z*z), 2*(x*y-w*z), 2*(x*z+w*y)],
                   [2*(x*y+w*z), 1-2*(x*x+z*z), 2*(y*z-w*x)],
                   [2*(x*z-w*y), 2*(y*z+w*x), 1-2*(x*x+y*y)]])


def m2q(m):
  # robust matrix -> quaternion
  t = np.trace(m)
  if t > 0:
    s = np.sqrt(t+1.0)*2
    q = np.array([0.25*s, (m[2, 1]-m[1, 2])/s, (m[0, 2]-m[2, 0])/s, (m[1, 0]-m[0, 1])/s])
  elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
    s = np.sqrt(1.0+m[0, 0]-m[1, 1]-m[2, 2])*2
    q = np.array([(m[2, 1]-m[1, 2])/s, 0.25*s, (m[0, 1]+m[1, 0])/s, (m[0, 2]+m[2, 0])/s])
  elif m[1, 1] > m[2, 2]:
    s = np.sqrt(1.0+m[1, 1]-m[0, 0]-m[2, 2])*2
    q = np.array([(m[0, 2]-m[2, 0])/s, (m[0, 1]+m[1, 0])/s, 0.25*s, (m[1, 2]+m[2, 1])/s])
  else:
    s = np.sqrt(1.0+m[2, 2]-m[0, 0]-m[1, 1])*2
    q = np.array([(m[1, 0]-m[0, 1])/s, (m[0, 2]+m[2, 0])/s, (m[1, 2]+m[2, 1])/s, 0.25*s])
  return q/np.linalg.norm(q)


# ----------------------------------------------------------------------------------------------
# mesh IO
def load_stl(path):
  with open(path, 'rb') as f:
    data = f.read()
  n = struct.unpack_from('<I', data, 80)[0]
  assert len(data) == 84+50*n, f'{path}: not a binary STL'
  rec = np.frombuffer(data, dtype=np.dtype([('n', '<f4', 3), ('v', '<f4', (3, 3)), ('a', '<u2')]), count=n, offset=84)
  tri = rec['v'].astype(np.float64)           # [n,3,3]
  verts, inv = np.unique(tri.reshape(-1, 3), axis=0, return_inverse=True)
  return verts, inv.reshape(-1, 3)


def load_obj(path):
  vs, fs = [], []
  with open(path) as f:
    for line in f:
      if line.startswith('v '):
        vs.append([float(x) for x in line.split()[1:4]])
      elif line.startswith('f '):
        idx = [int(t.split('/')[0])-1 for t in line.split()[1:]]
        for k in range(1, len(idx)-1):
          fs.append([idx[0], idx[k], idx[k+1]])
  return np.array(vs, dtype=np.float64), np.array(fs, dtype=np.int64)


def mesh_mass_props(verts, faces):
  """Signed-tetrahedron volume integrals: volume, com, inertia about com (unit density)."""
  a, b, c = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
  vol6 = np.einsum('ij,ij->i', a, np.cross(b, c))
  vol = vol6.sum()/6.0
  com = ((a+b+c)/4.0*(vol6/6.0)[:, None]).sum(0)/vol
  # second moments about origin
  def sm(i, j):
    return (vol6/120.0*(2*a[:, i]*a[:, j]+2*b[:, i]*b[:, j]+2*c[:, i]*c[:, j]
                        + a[:, i]*b[:, j]+a[:, j]*b[:, i]+a[:, i]*c[:, j]+a[:, j]*c[:, i]
                        + b[:, i]*c[:, j]+b[:, j]*c[:, i])).sum()
  P = np.array([[sm(i, j) for j in range(3)] for i in range(3)])
  P = P-vol*np.outer(com, com)                 # about com
  I = np.trace(P)*np.eye(3)-P
  return vol, com, I


def hull_of(verts):
  """Convex hull: vertices, outward-oriented triangles, CSR vertex adjacency."""
  h = ConvexHull(verts)
  vid = np.unique(h.simplices)
  remap = -np.ones(len(verts), dtype=np.int64)
  remap[vid] = np.arange(len(vid))
  hv = verts[vid]
  tris = remap[h.simplices]
  # orient outward using qhull's facet equations
  nrm = np.cross(hv[tris[:, 1]]-hv[tris[:, 0]], hv[tris[:, 2]]-hv[tris[:, 0]])
  flip = np.einsum('ij,ij->i', nrm, h.equations[:, :3]) < 0
  tris[flip] = tris[flip][:, ::-1]
  nbr = [set() for _ in range(len(hv))]
  for t in tris:
    for i in range(3):
      nbr[t[i]].add(int(t[(i+1) % 3])); nbr[t[i]].add(int(t[(i+2) % 3]))
  adr = np.zeros(len(hv)+1, dtype=np.int64)
  lst = []
  for i, s in enumerate(nbr):
    lst.extend(sorted(s)); adr[i+1] = len(lst)
  return hv, tris, adr, np.array(lst, dtype=np.int64)


# ----------------------------------------------------------------------------------------------
# MJCF subset parser with default classes
class Defaults:
  def __init__(self):
    self.cls = {}          # class name -> {tag: {attr: val}}
    self.parent = {}

  def load(self, elem, parent='__none__', top=True):
    name = elem.get('class', 'main' if top else None)
    d = {}
    if parent in self.cls:
      d = {k: dict(v) for k, v in self.cls[parent].items()}
    for ch in elem:
      if ch.tag == 'default':
        continue
      d.setdefault(ch.tag, {}).update(ch.attrib)
    self.cls[name] = d
    for ch in elem:
      if ch.tag == 'default':
        self.load(ch, name, top=False)

  def resolve(self, tag, elem, childclass, base):
    out = dict(base)
    c = elem.get('class', childclass)
    if c is None:
      c = 'main'
    if c in self.cls:
      out.update(self.cls[c].get(tag, {}))
    out.update({k: v for k, v in elem.attrib.items() if k != 'class'})
    return out


class Model:
  """Flat model under construction (lists of dicts)."""

  def __init__(self):
    self.bodies = [dict(name='world', parent=0, pos=np.zeros(3), quat=np.array([1., 0, 0, 0]), joints=[], inertial=None,
                        geoms=[], weld=0)]
    self.joints, self.geoms, self.acts, self.excludes = [], [], [], []
    self.opt = dict(timestep=0.002, gravity=np.array([0, 0, -9.81]), impratio=1.0, cone='pyramidal', tolerance=1e-8,
                    iterations=100, ls_iterations=50, ls_tolerance=0.01)


def parse_mjcf(model: Model, path, prefix='', attach_free=False, mesh_scale=1.0):
  root = ET.parse(path).getroot()
  base = os.path.dirname(path)
  comp = root.find('compiler')
  # (assetdir is the fallback of meshdir: edr/pen and the gso models use it)
  meshdir = os.path.join(base, (comp.get('meshdir') or comp.get('assetdir') or '') if comp is not None else '')
  opt = root.find('option')
  if opt is not None and not attach_free:
    if opt.get('impratio'): model.opt['impratio'] = float(opt.get('impratio'))
    if opt.get('cone'): model.opt['cone'] = opt.get('cone')
    if opt.get('timestep'): model.opt['timestep'] = float(opt.get('timestep'))
  dfl = Defaults()
  d0 = root.find('default')
  if d0 is not None:
    dfl.load(d0)
  meshes = {}
  for m in root.find('asset').findall('mesh'):
    f = m.get('file')
    nm = m.get('name', os.path.splitext(os.path.basename(f))[0])
    meshes[nm] = (os.path.join(meshdir, f), fvec(m.get('scale', '1 1 1'))*mesh_scale)

  def load_mesh(nm):
    p, sc = meshes[nm]
    v, f = (load_stl if p.lower().endswith('.stl') else load_obj)(p)
    return v*sc, f

  def walk(belem, parent_id, childclass):
    for b in belem.findall('body'):
      cc = b.get('childclass', childclass)
      bid = len(model.bodies)
      body = dict(name=prefix+b.get('name', f'body{bid}'), parent=parent_id, pos=fvec(b.get('pos', '0 0 0')),
                  quat=fvec(b.get('quat', '1 0 0 0')), joints=[], inertial=None, geoms=[], massgeoms=[])
      model.bodies.append(body)
      ine = b.find('inertial')
      if ine is not None:
        body['inertial'] = dict(pos=fvec(ine.get('pos')), quat=fvec(ine.get('quat', '1 0 0 0')), mass=float(ine.get('mass')),
                                diag=fvec(ine.get('diaginertia')))
      jl = list(b.findall('joint'))
      if b.find('freejoint') is not None or (attach_free and parent_id == 0 and belem is root.find('worldbody')):
        # add_free_entity: the prop's own freejoint is removed and a free joint is added on the attachment frame
        # (so100_hand_over.py:168,199-206).  Kinematically identical to a free joint on this body.
        model.joints.append(dict(name=body['name']+'/free', type=JNT_FREE, body=bid))
        body['joints'].append(len(model.joints)-1)
      for j in jl:
        a = dfl.resolve('joint', j, cc, JOINT_DEFAULTS)
        assert a['type'] == 'hinge', 'only hinge/free joints are on this path'
        rng = fvec(a['range']) if 'range' in a else None
        model.joints.append(dict(name=prefix+j.get('name'), type=JNT_HINGE, body=bid, pos=fvec(a['pos']),
                                 axis=fvec(a['axis'])/np.linalg.norm(fvec(a['axis'])), armature=float(a['armature']),
                                 frictionloss=float(a['frictionloss']), damping=float(a['damping']), range=rng,
                                 solreflimit=fvec(a['solreflimit'], 2, [0.02, 1]), solimplimit=fvec(a['solimplimit'], 5, [0.9, 0.95, 0.001, 0.5, 2]),
                                 solreffriction=fvec(a['solreffriction'], 2, [0.02, 1]),
                                 solimpfriction=fvec(a['solimpfriction'], 5, [0.9, 0.95, 0.001, 0.5, 2]), margin=float(a['margin'])))
        body['joints'].append(len(model.joints)-1)
      for g in b.findall('geom'):
        add_geom(g, bid, cc, body)
      walk(b, bid, cc)

  def add_geom(g, bid, cc, body):
    a = dfl.resolve('geom', g, cc, GEOM_DEFAULTS)
    collides = int(a['contype']) != 0 or int(a['conaffinity']) != 0
    gd = dict(name=prefix+(g.get('name') or f"geom{len(model.geoms)}"), body=bid, pos=fvec(a['pos']), quat=fvec(a['quat']),
              condim=int(a['condim']), friction=fvec(a['friction'], 3, [1, 0.005, 0.0001]), solref=fvec(a['solref'], 2, [0.02, 1]),
              solimp=fvec(a['solimp'], 5, [0.9, 0.95, 0.001, 0.5, 2]), margin=float(a['margin']), gap=float(a['gap']),
              solmix=float(a['solmix']), priority=int(a['priority']), contype=int(a['contype']), conaffinity=int(a['conaffinity']),
              density=float(a['density']), mass=(float(a['mass']) if 'mass' in a else None), type=a['type'])
    if a['type'] == 'mesh':
      try:
        v, f = load_mesh(a['mesh'])
      except FileNotFoundError:
        gd['missing'] = True
        v = f = None
      gd['mesh'] = (v, f)
    else:
      gd['size'] = fvec(a['size'], 3, [0, 0, 0])
    if collides:
      body['geoms'].append(gd)
    if gd['density'] > 0 or (gd['mass'] or 0) > 0:
      body.setdefault('massgeoms', []).append(gd)

  wb = root.find('worldbody')
  if not attach_free:
    for g in wb.findall('geom'):
      add_geom(g, 0, None, model.bodies[0])
  walk(wb, 0, None)
  con = root.find('contact')
  if con is not None:
    for e in con.findall('exclude'):
      model.excludes.append((prefix+e.get('body1'), prefix+e.get('body2')))
  act = root.find('actuator')
  if act is not None:
    for g in act.findall('general'):
      a = dfl.resolve('general', g, None, GENERAL_DEFAULTS)
      assert a['biastype'] == 'affine' and a['gaintype'] == 'fixed'
      model.acts.append(dict(name=prefix+g.get('name'), joint=prefix+a['joint'], gain=fvec(a['gainprm'])[0],
                             bias=fvec(a['biasprm'], 3, [0, 0, 0]), ctrlrange=fvec(a['ctrlrange']), forcerange=fvec(a['forcerange']),
                             gear=fvec(a['gear'])[0]))


# ----------------------------------------------------------------------------------------------
def finalize(model: Model, standin_meshes):
  """Resolve inertials, hulls, bounding volumes, pair lists, setConst constants; return dict of flat arrays."""
  nb = len(model.bodies)
  A = {}
  body_parent = np.array([b['parent'] for b in model.bodies], dtype=np.int32)
  body_pos = np.array([b['pos'] for b in model.bodies])
  body_quat = np.array([b['quat']/np.linalg.norm(b['quat']) for b in model.bodies])
  body_mass = np.zeros(nb); body_inertia = np.zeros((nb, 3)); body_ipos = np.zeros((nb, 3)); body_iquat = np.tile([1., 0, 0, 0], (nb, 1))
  # joints / dofs
  nq = nv = 0
  jnt_type, jnt_body, jnt_qposadr, jnt_dofadr = [], [], [], []
  for j in model.joints:
    jnt_type.append(j['type']); jnt_body.append(j['body']); jnt_qposadr.append(nq); jnt_dofadr.append(nv)
    nq += 7 if j['type'] == JNT_FREE else 1
    nv += 6 if j['type'] == JNT_FREE else 1
  njnt = len(model.joints)
  body_dofnum = np.zeros(nb, dtype=np.int32)
  for j in model.joints:
    body_dofnum[j['body']] += 6 if j['type'] == JNT_FREE else 1
  # weld id: static bodies share the weld of their parent
  body_weld = np.zeros(nb, dtype=np.int32)
  for i in range(1, nb):
    body_weld[i] = i if body_dofnum[i] > 0 else body_weld[body_parent[i]]

  # inertials
  for i, b in enumerate(model.bodies):
    if i == 0:
      continue
    if b['inertial'] is not None:
      ine = b['inertial']
      body_mass[i], body_inertia[i], body_ipos[i], body_iquat[i] = ine['mass'], ine['diag'], ine['pos'], ine['quat']/np.linalg.norm(ine['quat'])
      continue
    # from geoms: mesh geoms with density (visual mesh; stand-in when missing); primitives with density
    m_tot, c_acc, parts = 0.0, np.zeros(3), []
    for g in b.get('massgeoms', []):
      if g['type'] == 'mesh':
        v, f = g['mesh']
        if v is None:
          v, f = standin_meshes[b['name']]
        vol, com, I = mesh_mass_props(v, f)
        R, p = q2m(g['quat']), g['pos']
        m = g['density']*abs(vol)
        parts.append((m, p+R@com, R@(I*g['density']*np.sign(vol))@R.T))
      elif g['type'] == 'box':
        sx, sy, sz = g['size']; m = g['density']*8*sx*sy*sz
        I = np.diag([m/3*(sy*sy+sz*sz), m/3*(sx*sx+sz*sz), m/3*(sx*sx+sy*sy)]); R = q2m(g['quat'])
        parts.append((m, g['pos'], R@I@R.T))
      elif g['type'] == 'cylinder':
        r, h = g['size'][:2]; m = g['density']*np.pi*r*r*2*h
        I = np.diag([m*(3*r*r+4*h*h)/12, m*(3*r*r+4*h*h)/12, m*r*r/2]); R = q2m(g['quat'])
        parts.append((m, g['pos'], R@I@R.T))
      elif g['type'] == 'capsule':
        r, h = g['size'][:2]
        mc = g['density']*np.pi*r*r*2*h; ms = g['density']*4/3*np.pi*r**3; m = mc+ms
        Iz = mc*r*r/2+ms*2*r*r/5
        Ix = mc*(3*r*r+4*h*h)/12+ms*(2*r*r/5+h*h+3*h*r/4)
        R = q2m(g['quat']); parts.append((m, g['pos'], R@np.diag([Ix, Ix, Iz])@R.T))
      elif g['type'] == 'plane':
        pass
    if not parts:
      continue
    m_tot = sum(p[0] for p in parts)
    com = sum(p[0]*p[1] for p in parts)/m_tot
    I = np.zeros((3, 3))
    for m, c, Ic in parts:
      d = c-com
      I += Ic+m*(d@d*np.eye(3)-np.outer(d, d))
    w, V = np.linalg.eigh(I)
    order = np.argsort(-w)                      # [upstream] principal moments sorted descending
    w, V = w[order], V[:, order]
    if np.linalg.det(V) < 0:
      V[:, 2] = -V[:, 2]
    body_mass[i], body_inertia[i], body_ipos[i], body_iquat[i] = m_tot, w, com, m2q(V)

  # geoms
  G = []
  verts_all, vert_adr, nbr_all, nbr_adr_all = [], [0], [], [0]
  face_all, face_adr = [], [0]
  for bi, b in enumerate(model.bodies):
    for g in b['geoms']:
      t = g['type']
      e = dict(g); e['body'] = bi
      R, p = q2m(g['quat']), g['pos']
      if t == 'mesh':
        v, f = g['mesh']
        hv, tris, adr, lst = hull_of(v)
        hv = hv@R.T+p                           # hull vertices stored in the BODY frame
        e['tcode'] = GEOM_HULL
        e['gpos'], e['gmat'] = np.zeros(3), np.eye(3)
        c = 0.5*(hv.min(0)+hv.max(0))
        e['bcenter'] = c; e['rbound'] = np.linalg.norm(hv-c, axis=1).max(); e['size'] = np.zeros(3)
        e['vadr'], e['vnum'] = len(verts_all), len(hv)
        e['nbr_base'] = len(nbr_all)
        nbr_all.extend(lst.tolist())
        e['nbr_adr'] = (adr+e['nbr_base']).tolist()
        e['fadr'], e['fnum'] = len(face_all), len(tris)
        face_all.extend(tris.tolist())
        verts_all.extend(hv.tolist())
        # [upstream] mesh geoms are re-centred on the mesh's own inertial frame; geom_aabb lives there.
        vol, com, I = mesh_mass_props(hv, tris)
        w, V = np.linalg.eigh(I); order = np.argsort(-w); V = V[:, order]
        if np.linalg.det(V) < 0: V[:, 2] = -V[:, 2]
        loc = (hv-com)@V
        e['mframe'] = (com, V, 0.5*(loc.min(0)+loc.max(0)), 0.5*(loc.max(0)-loc.min(0)))
      else:
        e['tcode'] = dict(plane=GEOM_PLANE, sphere=GEOM_SPHERE, capsule=GEOM_CAPSULE, cylinder=GEOM_CYLINDER, box=GEOM_BOX)[t]
        e['gpos'], e['gmat'] = p, R
        s = g['size']
        e['bcenter'] = p
        e['rbound'] = dict(plane=0.0, sphere=s[0], capsule=s[0]+s[1], cylinder=np.hypot(s[0], s[1]), box=np.linalg.norm(s))[t]
        e['vadr'], e['vnum'], e['fadr'], e['fnum'] = 0, 0, 0, 0
        half = dict(plane=np.zeros(3), sphere=np.full(3, s[0]), capsule=np.array([s[0], s[0], s[0]+s[1]]),
                    cylinder=np.array([s[0], s[0], s[1]]), box=s)[t]
        e['mframe'] = (p, R, np.zeros(3), half)
      G.append(e)
  ng = len(G)
  # non-colliding mesh geoms whose file is present (visual / mass meshes): only their AABB matters, for the body's BVH root box
  vis = {}
  for bi, b in enumerate(model.bodies):
    for g in b.get('massgeoms', []):
      if any(g is c for c in b['geoms']) or g['type'] != 'mesh' or g['mesh'][0] is None:
        continue
      v, f = g['mesh']
      v = v@q2m(g['quat']).T+g['pos']
      vol, com, I = mesh_mass_props(v, f)
      w, V = np.linalg.eigh(I*np.sign(vol)); order = np.argsort(-w); V = V[:, order]
      if np.linalg.det(V) < 0: V[:, 2] = -V[:, 2]
      loc = (v-com)@V
      vis.setdefault(bi, []).append((com, V, 0.5*(loc.min(0)+loc.max(0)), 0.5*(loc.max(0)-loc.min(0))))
  A['_vis'] = vis
  # body geom ranges (colliding geoms are contiguous per body by construction)
  body_geomadr = np.zeros(nb, dtype=np.int32); body_geomnum = np.zeros(nb, dtype=np.int32)
  k = 0
  for bi, b in enumerate(model.bodies):
    body_geomadr[bi] = k; body_geomnum[bi] = len(b['geoms']); k += len(b['geoms'])
  # body bounding spheres over colliding geoms (body frame)
  body_bcenter = np.zeros((nb, 3)); body_rbound = np.zeros(nb)
  for bi in range(nb):
    gs = [G[g] for g in range(body_geomadr[bi], body_geomadr[bi]+body_geomnum[bi]) if G[g]['tcode'] != GEOM_PLANE]
    if not gs:
      continue
    lo = np.min([g['bcenter']-g['rbound'] for g in gs], axis=0); hi = np.max([g['bcenter']+g['rbound'] for g in gs], axis=0)
    c = 0.5*(lo+hi)
    body_bcenter[bi] = c; body_rbound[bi] = max(np.linalg.norm(g['bcenter']-c)+g['rbound'] for g in gs)

  # static body-pair filtering  [upstream] mj_collision filters: same weld, parent-child (unless parent weld is world), exclude
  name2body = {b['name']: i for i, b in enumerate(model.bodies)}
  excl = {(name2body[a], name2body[b]) for a, b in model.excludes}
  excl |= {(b, a) for a, b in excl}
  pairs = []
  for b1 in range(nb):
    for b2 in range(b1+1, nb):
      if body_geomnum[b1] == 0 or body_geomnum[b2] == 0: continue
      w1, w2 = body_weld[b1], body_weld[b2]
      if w1 == w2: continue                      # same weld group (incl. static-static)
      if (b1, b2) in excl: continue
      pw1, pw2 = body_weld[body_parent[w1]] if w1 else 0, body_weld[body_parent[w2]] if w2 else 0
      if (w1 != 0 and w2 != 0) and (pw1 == w2 or pw2 == w1): continue   # parent-child, parent not world-welded
      pairs.append((b1, b2))
  # NB the explicit exclude Base--Rotation_Pitch (scene_pbr.xml:149-151) is needed because Base is world-welded.

  # ---- setConst: M at qpos0, invweight0, meaninertia (numpy, generic tree)
  qpos0 = np.zeros(nq)
  for j, jd in enumerate(model.joints):
    if jd['type'] == JNT_FREE:
      b = jd['body']
      # free body qpos0 = its body pos/quat in the parent (world) frame
      qpos0[jnt_qposadr[j]:jnt_qposadr[j]+3] = body_pos[b]
      qpos0[jnt_qposadr[j]+3:jnt_qposadr[j]+7] = body_quat[b]
  xpos = np.zeros((nb, 3)); xmat = np.tile(np.eye(3), (nb, 1, 1))
  for i in range(1, nb):
    p = body_parent[i]
    xpos[i] = xpos[p]+xmat[p]@body_pos[i]; xmat[i] = xmat[p]@q2m(body_quat[i])
  xipos = np.array([xpos[i]+xmat[i]@body_ipos[i] for i in range(nb)])
  ximat = np.array([xmat[i]@q2m(body_iquat[i]) for i in range(nb)])
  dof_body = np.zeros(nv, dtype=np.int32); dof_jnt = np.zeros(nv, dtype=np.int32)
  armature = np.zeros(nv); frictionloss = np.zeros(nv); damping = np.zeros(nv)
  # world-frame motion axes at qpos0: column d of Jp(b)/Jr(b) for body b
  dof_axis_ang = np.zeros((nv, 3)); dof_axis_lin0 = np.zeros((nv, 3)); dof_anchor = np.zeros((nv, 3)); dof_is_trans = np.zeros(nv, dtype=bool)
  for j, jd in enumerate(model.joints):
    b, d = jd['body'], jnt_dofadr[j]
    if jd['type'] == JNT_HINGE:
      dof_body[d] = b; dof_jnt[d] = j
      dof_axis_ang[d] = xmat[b]@jd['axis']; dof_anchor[d] = xpos[b]+xmat[b]@jd['pos']
      armature[d], frictionloss[d], damping[d] = jd['armature'], jd['frictionloss'], jd['damping']
    else:
      for k in range(6):
        dof_body[d+k] = b; dof_jnt[d+k] = j
      for k in range(3):
        dof_is_trans[d+k] = True; dof_axis_lin0[d+k] = np.eye(3)[k]
        dof_axis_ang[d+3+k] = xmat[b][:, k]; dof_anchor[d+3+k] = xpos[b]

  def ancestors(b):
    out = set()
    while b != 0:
      out.add(b); b = body_parent[b]
    return out
  anc = [ancestors(b) for b in range(nb)]

  def jac_point(b, point):
    Jp = np.zeros((3, nv)); Jr = np.zeros((3, nv))
    for d in range(nv):
      if dof_body[d] in anc[b]:
        if dof_is_trans[d]:
          Jp[:, d] = dof_axis_lin0[d]
        else:
          Jr[:, d] = dof_axis_ang[d]; Jp[:, d] = np.cross(dof_axis_ang[d], point-dof_anchor[d])
    return Jp, Jr
  M = np.diag(armature.copy())
  for b in range(1, nb):
    if body_mass[b] <= 0: continue
    Jp, Jr = jac_point(b, xipos[b])
    Iw = ximat[b]@np.diag(body_inertia[b])@ximat[b].T
    M += body_mass[b]*Jp.T@Jp+Jr.T@Iw@Jr
  Minv = np.linalg.inv(M) if nv else np.zeros((0, 0))
  MINVAL = 1e-15
  body_invweight0 = np.zeros((nb, 2))
  for b in range(1, nb):
    if body_weld[b] == 0: continue
    Jp, Jr = jac_point(b, xipos[b])
    body_invweight0[b, 0] = max(MINVAL, np.trace(Jp@Minv@Jp.T)/3)
    body_invweight0[b, 1] = max(MINVAL, np.trace(Jr@Minv@Jr.T)/3)
  dof_invweight0 = np.zeros(nv)
  for j, jd in enumerate(model.joints):
    d = jnt_dofadr[j]
    if jd['type'] == JNT_HINGE:
      dof_invweight0[d] = Minv[d, d]
    else:
      dof_invweight0[d:d+3] = np.trace(Minv[d:d+3, d:d+3])/3
      dof_invweight0[d+3:d+6] = np.trace(Minv[d+3:d+6, d+3:d+6])/3
  meaninertia = np.trace(M)/max(nv, 1)

  # ---- reward constants: root-BVH box of the banana body in its inertial frame (oobb_utils.py:165-172)
  A['nq'], A['nv'], A['nu'], A['nbody'], A['njnt'], A['ngeom'] = nq, nv, len(model.acts), nb, njnt, ng
  A['opt'] = np.array([model.opt['timestep'], *model.opt['gravity'], model.opt['impratio'], model.opt['tolerance'],
                       model.opt['iterations'], 1.0 if model.opt['cone'] == 'elliptic' else 0.0, meaninertia,
                       model.opt['ls_iterations'], model.opt['ls_tolerance']])
  A['body_parent'] = body_parent; A['body_weld'] = body_weld
  A['body_pos'] = body_pos; A['body_quat'] = body_quat; A['body_ipos'] = body_ipos; A['body_iquat'] = body_iquat
  A['body_mass'] = body_mass; A['body_inertia'] = body_inertia; A['body_invweight0'] = body_invweight0
  A['body_geomadr'] = body_geomadr; A['body_geomnum'] = body_geomnum; A['body_bcenter'] = body_bcenter; A['body_rbound'] = body_rbound
  A['body_jntadr'] = np.array([b['joints'][0] if b['joints'] else -1 for b in model.bodies], dtype=np.int32)
  A['body_jntnum'] = np.array([len(b['joints']) for b in model.bodies], dtype=np.int32)
  A['jnt_type'] = np.array(jnt_type, dtype=np.int32); A['jnt_body'] = np.array(jnt_body, dtype=np.int32)
  A['jnt_qposadr'] = np.array(jnt_qposadr, dtype=np.int32); A['jnt_dofadr'] = np.array(jnt_dofadr, dtype=np.int32)
  A['jnt_pos'] = np.array([j.get('pos', np.zeros(3)) for j in model.joints]); A['jnt_axis'] = np.array([j.get('axis', np.array([0, 0, 1.])) for j in model.joints])
  A['jnt_limited'] = np.array([1 if j.get('range') is not None else 0 for j in model.joints], dtype=np.int32)
  A['jnt_range'] = np.array([j['range'] if j.get('range') is not None else np.zeros(2) for j in model.joints])
  A['jnt_solreflimit'] = np.array([j.get('solreflimit', np.array([0.02, 1])) for j in model.joints])
  A['jnt_solimplimit'] = np.array([j.get('solimplimit', np.array([0.9, 0.95, 0.001, 0.5, 2])) for j in model.joints])
  A['jnt_solreffriction'] = np.array([j.get('solreffriction', np.array([0.02, 1])) for j in model.joints])
  A['jnt_solimpfriction'] = np.array([j.get('solimpfriction', np.array([0.9, 0.95, 0.001, 0.5, 2])) for j in model.joints])
  A['jnt_margin'] = np.array([j.get('margin', 0.0) for j in model.joints])
  A['dof_body'] = dof_body; A['dof_jnt'] = dof_jnt; A['dof_armature'] = armature; A['dof_frictionloss'] = frictionloss
  A['dof_damping'] = damping; A['dof_invweight0'] = dof_invweight0; A['qpos0'] = qpos0
  jname = {j['name']: i for i, j in enumerate(model.joints)}
  A['act_jnt'] = np.array([jname[a['joint']] for a in model.acts], dtype=np.int32)
  A['act_gain'] = np.array([a['gain'] for a in model.acts]); A['act_bias'] = np.array([a['bias'] for a in model.acts]).reshape(-1, 3)
  A['act_ctrlrange'] = np.array([a['ctrlrange'] for a in model.acts]).reshape(-1, 2)
  A['act_forcerange'] = np.array([a['forcerange'] for a in model.acts]).reshape(-1, 2)
  A['act_gear'] = np.array([a['gear'] for a in model.acts])
  A['geom_type'] = np.array([g['tcode'] for g in G], dtype=np.int32); A['geom_body'] = np.array([g['body'] for g in G], dtype=np.int32)
  A['geom_pos'] = np.array([g['gpos'] for g in G]); A['geom_mat'] = np.array([g['gmat'].reshape(9) for g in G])
  A['geom_size'] = np.array([g['size'] for g in G]); A['geom_bcenter'] = np.array([g['bcenter'] for g in G])
  A['geom_rbound'] = np.array([g['rbound'] for g in G]); A['geom_condim'] = np.array([g['condim'] for g in G], dtype=np.int32)
  A['geom_friction'] = np.array([g['friction'] for g in G]); A['geom_solref'] = np.array([g['solref'] for g in G])
  A['geom_solimp'] = np.array([g['solimp'] for g in G]); A['geom_solmix'] = np.array([g['solmix'] for g in G])
  A['geom_margin'] = np.array([g['margin'] for g in G]); A['geom_gap'] = np.array([g['gap'] for g in G])
  A['geom_priority'] = np.array([g['priority'] for g in G], dtype=np.int32)
  A['geom_vertadr'] = np.array([g['vadr'] for g in G], dtype=np.int32); A['geom_vertnum'] = np.array([g['vnum'] for g in G], dtype=np.int32)
  A['geom_faceadr'] = np.array([g['fadr'] for g in G], dtype=np.int32); A['geom_facenum'] = np.array([g['fnum'] for g in G], dtype=np.int32)
  A['hull_vert'] = np.array(verts_all).reshape(-1, 3)
  A['hull_face'] = np.array(face_all, dtype=np.int32).reshape(-1, 3)
  # CSR adjacency (global vertex ids -> local neighbour ids within the geom)
  nadr = np.zeros(len(verts_all)+1, dtype=np.int32)
  for g in G:
    if g['tcode'] == GEOM_HULL:
      nadr[g['vadr']:g['vadr']+g['vnum']+1] = np.array(g['nbr_adr'], dtype=np.int32)
  A['hull_nbradr'] = nadr; A['hull_nbr'] = np.array(nbr_all, dtype=np.int32)
  A['bodypair'] = np.array(pairs, dtype=np.int32).reshape(-1, 2)
  A['_names_body'] = [b['name'] for b in model.bodies]; A['_names_geom'] = [g['name'] for g in G]
  A['_G'] = G
  A['_xipos0'] = xipos
  return A


def body_root_box(A, body):
  """[upstream] root bvh_aabb of a body: union of each geom's AABB (in the geom's own frame) re-boxed in the body inertial
  frame -> (centre3, half3).  Read by oobb_utils.get_oobb (oobb_utils.py:165-172)."""
  Ri = q2m(A['body_iquat'][body]); pi = A['body_ipos'][body]
  lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
  frames = [g['mframe'] for g in A['_G'] if g['body'] == body]+A['_vis'].get(body, [])
  for com, V, c, h in frames:
    for sx in (-1, 1):
      for sy in (-1, 1):
        for sz in (-1, 1):
          pb = com+V@(c+h*np.array([sx, sy, sz]))     # corner in body frame
          pl = Ri.T@(pb-pi)
          lo, hi = np.minimum(lo, pl), np.maximum(hi, pl)
  return np.concatenate([0.5*(lo+hi), 0.5*(hi-lo)])


def write_blob(path, A):
  ent = []
  for k, v in A.items():
    if k.startswith('_'): continue
    if isinstance(v, (int, np.integer)):
      v = np.array([v], dtype=np.int32)
    v = np.ascontiguousarray(v)
    if v.dtype.kind in 'iu' or v.dtype == bool:
      v = v.astype(np.int32); dt = 1
    else:
      v = v.astype(np.float64); dt = 0
    ent.append((k, dt, v))
  hdr = struct.pack('<4sIII', b'SO1B', 2, len(ent), 0)
  off = len(hdr)+40*len(ent)
  off = (off+7)//8*8
  table, payload = b'', b''
  for k, dt, v in ent:
    assert len(k) < 24, k
    raw = v.tobytes()
    table += struct.pack('<24sIIQ', k.encode(), dt, v.size, off+len(payload))
    payload += raw+b'\0'*((-len(raw)) % 8)
  with open(path, 'wb') as f:
    f.write(hdr+table)
    f.write(b'\0'*(off-len(hdr)-len(table)))
    f.write(payload)


# so100_hand_over.py:80-118 — object / container models, container mesh scale and the overlap boxes in the container frame
HANDOVER = {
    'banana': dict(object=('011_banana/', 'ycb/011_banana/google_64k'), container=('024_bowl/', 'ycb/024_bowl/google_64k'), scale=1.5,
                   boxes=[(np.array([-0.017, -0.045, 0.035])*1.5, np.array([0.02, 0.02, 0.01])*1.5)]),
    'pen': dict(object=('pen/', 'edr/pen'), container=('utensil_holder/', 'gso/BIA_Cordon_Bleu_White_Porcelain_Utensil_Holder_900028'),
                scale=0.6, boxes=[(np.array([0.0, 0.0, 0.02666])*0.6, np.array([0.04666, 0.04666, 0.025])*0.6),
                                  (np.array([0.0, 0.0, 0.25])*0.6, np.array([0.1, 0.1, 0.01666])*0.6)]),
}


# BASELINE config 4 ("hand-over task, two arms"): the reference has no two-SO100 scene (its only two-arm hand-over is the ALOHA
# task, so101_sim/tasks/hand_over.py:122), so this is a LABELLED SYNTHETIC scene: scene_pbr.xml with its arm
# moved to (+0.12, 0.3) and a second, identical arm (names prefixed "B_") at (-0.12, 0.3), both reaching over the prop
# placement area along -y; table, static obstacles and the two free props are the reference's.
TWO_ARM_BASES = ((0.12, 0.3, 0.42), (-0.12, 0.3, 0.42))


def two_arm_scene_xml(scene_path):
  """scene_pbr.xml with the Base subtree duplicated: returns the path of a temporary MJCF file (absolute meshdir)."""
  import copy, tempfile
  tree = ET.parse(scene_path)
  root = tree.getroot()
  comp = root.find('compiler')
  comp.set('meshdir', os.path.join(os.path.dirname(scene_path), comp.get('meshdir', '')))
  wb = root.find('worldbody')
  base = [b for b in wb.findall('body') if b.get('name') == 'Base'][0]
  base.set('pos', ' '.join(str(v) for v in TWO_ARM_BASES[0]))
  second = copy.deepcopy(base)
  second.set('pos', ' '.join(str(v) for v in TWO_ARM_BASES[1]))
  for e in second.iter():
    if e.get('name') is not None and e.tag in ('body', 'joint', 'geom'):
      e.set('name', 'B_' + e.get('name'))
  wb.insert(list(wb).index(base) + 1, second)
  con = root.find('contact')
  for e in list(con.findall('exclude')):
    con.append(ET.Element('exclude', body1='B_' + e.get('body1'), body2='B_' + e.get('body2')))
  act = root.find('actuator')
  for g in list(act.findall('general')):
    attrs = dict(g.attrib); attrs['name'] = 'B_' + attrs['name']; attrs['joint'] = 'B_' + attrs['joint']
    act.append(ET.Element('general', attrs))
  f = tempfile.NamedTemporaryFile('w', suffix='.xml', delete=False)
  tree.write(f.name)
  return f.name


def build(ref_root, task=None, two_arms=False):
  assets = os.path.join(ref_root, 'so101_sim', 'assets')
  m = Model()
  scene = os.path.join(assets, 'so100', 'scene_pbr.xml')
  parse_mjcf(m, two_arm_scene_xml(scene) if two_arms else scene)
  standin = {}
  with_props = task is not None
  if with_props:
    cfg = HANDOVER[task]
    # so100_hand_over.py:159-199 — object first, then the container with its meshes scaled
    for (nm, rel), sc in ((cfg['object'], 1.0), (cfg['container'], cfg['scale'])):
      nb0 = len(m.bodies)
      parse_mjcf(m, os.path.join(assets, rel, 'model.xml'), prefix=nm, attach_free=True, mesh_scale=sc)
      v, f = load_obj(os.path.join(assets, rel, 'meshes', 'coacd_merged.obj'))
      standin[m.bodies[nb0]['name']] = (v*sc, f)  # used only when the mass-defining visual mesh is missing from the checkout
      # dm_control names: attachment frame '011_banana/' with the inner (unnamed) body; we keep one body (identity offset)
  A = finalize(m, standin)
  A['nprop'] = 2 if with_props else 0
  if with_props:
    nb = A['nbody']
    A['prop_body'] = np.array([nb-2, nb-1], dtype=np.int32)
    A['reward_obj_box'] = body_root_box(A, nb-2)
    A['reward_box_pos'] = np.concatenate([b[0] for b in cfg['boxes']])
    A['reward_box_half'] = np.concatenate([b[1] for b in cfg['boxes']])
    A['prop_mass_standin'] = int(any(g.get('missing') for bd in m.bodies[-2:] for g in bd.get('massgeoms', [])))
  else:
    A['prop_body'] = np.zeros(0, dtype=np.int32)
    A['reward_obj_box'] = np.zeros(6); A['reward_box_pos'] = np.zeros(3); A['reward_box_half'] = np.zeros(3)
    A['prop_mass_standin'] = 0
  return A


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--ref', default='/root/reference')
  ap.add_argument('--out', default=os.path.join(os.path.dirname(__file__), '..', 'so101_sim_b200', 'data'))
  a = ap.parse_args()
  os.makedirs(a.out, exist_ok=True)
  for name, props in (('so100_arm', None), ('so100_handover_banana', 'banana'), ('so100_handover_pen', 'pen'), ('so100_twoarm_banana', 'banana')):
    A = build(a.ref, props, two_arms=name.startswith('so100_twoarm'))
    p = os.path.join(a.out, name+'.blob')
    write_blob(p, A)
    print(f"{name}: nq={A['nq']} nv={A['nv']} nu={A['nu']} nbody={A['nbody']} ngeom={A['ngeom']} "
          f"nvert={len(A['hull_vert'])} pairs={len(A['bodypair'])} -> {p} ({os.path.getsize(p)} B)")
    for i, n in enumerate(A['_names_body']):
      print(f"   body {i:2d} {n:20s} parent={A['body_parent'][i]} weld={A['body_weld'][i]} mass={A['body_mass'][i]:.5f} "
            f"ngeom={A['body_geomnum'][i]} invw0={A['body_invweight0'][i]}")
    print('   dof_invweight0', A['dof_invweight0'])
    print('   meaninertia', A['opt'][8])
    if props:
      print('   reward_obj_box', A['reward_obj_box'], 'mass stand-in:', A['prop_mass_standin'])


if __name__ == '__main__':
  sys.exit(main())
