// Device-side description of the full contact scene (arm + static table/obstacles + two free props) and the
// host routine that uploads it from the model blob.  Everything here is read-only during stepping and shared by all
// envs; at ~0.5 MB it is L2-resident (126 MB L2 on B200).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "arm_dynamics.cuh"
#include "blob.hpp"

namespace so101 {

enum { G_PLANE = 0, G_SPHERE = 1, G_CAPSULE = 2, G_CYLINDER = 3, G_BOX = 4, G_HULL = 5 };
constexpr int NPROP = 2;
constexpr int MAXREWARDBOX = 2;  // container overlap boxes (banana / bowl: 1, pen / utensil holder: 2)
constexpr int NSLOT = NA + NPROP;   // dynamic bodies whose pose lives in shared memory: 6 links per arm + 2 props
constexpr int NV = NA + 6 * NPROP;  // 18 (one arm) / 24 (two arms)
constexpr int NQ = NA + 7 * NPROP;  // 20 / 26
constexpr int NBLK = NARM + NPROP;  // 6-dof blocks of the mass matrix: one per arm, one per prop

template <typename T>
struct alignas(16) Vec4 {
  T x, y, z, w;
};

// w of a hull_nbrv entry: local vertex id (11 bits) | degree (7 bits) << 11 | adjacency offset within the hull (14 bits) << 18.
// float: the bit pattern (never used in arithmetic); double: the integer value.
constexpr unsigned NBR_ID_BITS = 11, NBR_DEG_BITS = 7, NBR_ADR_BITS = 14;
template <typename T> __host__ __device__ inline T nbr_pack(unsigned v);
template <> __host__ __device__ inline float nbr_pack<float>(unsigned v) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(v);
#else
  float f; std::memcpy(&f, &v, 4); return f;
#endif
}
template <> __host__ __device__ inline double nbr_pack<double>(unsigned v) { return (double)v; }
__device__ __forceinline__ unsigned nbr_bits(float w) { return __float_as_uint(w); }
__device__ __forceinline__ unsigned nbr_bits(double w) { return (unsigned)w; }

template <typename T>
struct SceneModel {
  int ngeom, npair, nbody;
  // geoms
  const int *geom_type, *geom_body, *geom_slot, *geom_vertadr, *geom_vertnum, *geom_condim, *geom_priority;
  const T *geom_pos, *geom_mat, *geom_size, *geom_bcenter, *geom_rbound;  // body frame (static geoms: world frame)
  const T *geom_aabb;  // [ngeom][6] centre, half sizes of the geom's bounding box in the GEOM frame (mid phase)
  const T *geom_friction, *geom_solref, *geom_solimp, *geom_solmix, *geom_margin, *geom_gap;
  const Vec4<T> *hull_vert;  // body-frame hull vertices, 16/32-byte aligned for vector loads
  const int *hull_nbradr, *hull_nbr;  // vertex adjacency: CSR offsets per global vertex id, local neighbour ids (hill-climbing support)
  const Vec4<T> *hull_nbrv;           // per adjacency entry: the neighbour's coordinates; w packs its local id, degree and the offset
                                      //   of ITS adjacency list inside the hull's (nbr_pack): a hill-climb step costs one round trip
  // bodies
  const int *bodypair, *body_slot, *body_geomadr, *body_geomnum;
  const int *pair_start;          // [npair + 1]  range of body pair p in geompair
  const unsigned *geompair;       // static candidate geom pairs (g1 | g2 << 8) in broad-phase order (pair-major, g1-major)
  const T *body_bcenter, *body_rbound, *body_invweight0;  // static bodies: bcenter in the world frame
  // free props
  T prop_mass[NPROP], prop_ipos[NPROP][3], prop_Icom[NPROP][6], prop_Iorg[NPROP][6];  // inertia about COM / body origin, body axes
  T prop_Riq[NPROP][9];  // body_iquat as a matrix (body <- inertial frame), for the reward's ximat
  // task constants (so100_hand_over.py:87-93, oobb_utils.py:165-172)
  T reward_obj_box[6];
  int nreward_box;  // overlap boxes of the container (so100_hand_over.py:87-93,104-116): ALL must be overlapped (:263-273)
  T reward_box_pos[MAXREWARDBOX][3], reward_box_half[MAXREWARDBOX][3];
  T impratio, timestep;
};

// Host: flatten the blob into device arrays.  Static geoms are pre-transformed into the world frame.
template <typename T>
struct SceneModelHost {
  SceneModel<T> dev{};
  std::vector<void *> allocs;

  template <typename U>
  const U *up(const std::vector<U> &v) {
    void *p = nullptr;
    size_t bytes = (v.empty() ? 1 : v.size()) * sizeof(U);
    if (cudaMalloc(&p, bytes) != cudaSuccess) throw std::runtime_error("cudaMalloc(model) failed");
    if (!v.empty() && cudaMemcpy(p, v.data(), v.size() * sizeof(U), cudaMemcpyHostToDevice) != cudaSuccess)
      throw std::runtime_error("cudaMemcpy(model) failed");
    allocs.push_back(p);
    return static_cast<const U *>(p);
  }
  template <typename U>
  std::vector<T> cvt(const std::vector<U> &v) { return std::vector<T>(v.begin(), v.end()); }

  ~SceneModelHost() { for (void *p : allocs) cudaFree(p); }

  static void q2m(const double *q, double *m) {
    double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    double w = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
    m[0] = 1 - 2 * (y * y + z * z); m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
    m[3] = 2 * (x * y + w * z); m[4] = 1 - 2 * (x * x + z * z); m[5] = 2 * (y * z - w * x);
    m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = 1 - 2 * (x * x + y * y);
  }

  void build(const Blob &b) {
    const int nbody = b.scalar("nbody"), ngeom = b.scalar("ngeom");
    if (b.scalar("nq") != NQ || b.scalar("nv") != NV || b.scalar("nprop") != NPROP)
      throw std::runtime_error("this build of the scene kernels expects " + std::to_string(NARM) + " 6-dof arm(s) + 2 free props");
    const auto &bp = b.I("body_parent"), &bw = b.I("body_weld"), &jb = b.I("jnt_body"), &jt = b.I("jnt_type");
    // world poses of all bodies at qpos0 (only the static ones are used)
    std::vector<double> xpos(3 * nbody, 0.0), xmat(9 * nbody, 0.0);
    xmat[0] = xmat[4] = xmat[8] = 1;
    for (int i = 1; i < nbody; i++) {
      const int p = bp[i];
      double R[9];
      q2m(&b.F("body_quat")[4 * i], R);
      for (int r = 0; r < 3; r++) {
        xpos[3 * i + r] = xpos[3 * p + r];
        for (int c = 0; c < 3; c++) xpos[3 * i + r] += xmat[9 * p + 3 * r + c] * b.F("body_pos")[3 * i + c];
        for (int c = 0; c < 3; c++) {
          double s = 0;
          for (int k = 0; k < 3; k++) s += xmat[9 * p + 3 * r + k] * R[3 * k + c];
          xmat[9 * i + 3 * r + c] = s;
        }
      }
    }
    std::vector<int> slot(nbody, -1);
    for (int j = 0; j < NA; j++) slot[jb[j]] = j;
    int np = 0;
    for (size_t j = NA; j < jt.size(); j++) {
      if (jt[j] != 0) throw std::runtime_error("joints after the arms must be free joints");
      if (b.I("jnt_qposadr")[j] != NA + 7 * np || b.I("jnt_dofadr")[j] != NA + 6 * np) throw std::runtime_error("unexpected prop state layout");
      slot[jb[j]] = NA + np++;
    }
    for (int i = 1; i < nbody; i++)
      if (slot[i] < 0 && bw[i] != 0) throw std::runtime_error("dynamic body without a pose slot");
    // geoms: static ones go to the world frame (static hulls, i.e. the arm Base mesh, get their vertices baked instead)
    std::vector<double> gpos = b.F("geom_pos"), gmat = b.F("geom_mat"), gbc = b.F("geom_bcenter"), bbc = b.F("body_bcenter");
    std::vector<double> hv = b.F("hull_vert");
    std::vector<int> gslot(ngeom);
    for (int g = 0; g < ngeom; g++) {
      const int body = b.I("geom_body")[g];
      gslot[g] = slot[body];
      if (slot[body] >= 0) continue;
      const double *X = &xpos[3 * body], *R = &xmat[9 * body];
      auto xf = [&](const double *l, double *w) {
        for (int r = 0; r < 3; r++) w[r] = X[r] + R[3 * r] * l[0] + R[3 * r + 1] * l[1] + R[3 * r + 2] * l[2];
      };
      double c[3];
      xf(&gbc[3 * g], c);
      for (int r = 0; r < 3; r++) gbc[3 * g + r] = c[r];
      if (b.I("geom_type")[g] == G_HULL) {
        const int adr = b.I("geom_vertadr")[g], num = b.I("geom_vertnum")[g];
        for (int v = adr; v < adr + num; v++) {
          double w[3];
          xf(&hv[3 * v], w);
          for (int r = 0; r < 3; r++) hv[3 * v + r] = w[r];
        }
        continue;  // hull geoms carry an identity geom frame
      }
      double p[3], M[9];
      xf(&gpos[3 * g], p);
      for (int r = 0; r < 3; r++)
        for (int k = 0; k < 3; k++) {
          double sm = 0;
          for (int e = 0; e < 3; e++) sm += R[3 * r + e] * gmat[9 * g + 3 * e + k];
          M[3 * r + k] = sm;
        }
      for (int r = 0; r < 3; r++) gpos[3 * g + r] = p[r];
      for (int r = 0; r < 9; r++) gmat[9 * g + r] = M[r];
    }
    for (int i = 0; i < nbody; i++) {
      if (slot[i] >= 0) continue;
      double c[3];
      for (int r = 0; r < 3; r++) { c[r] = xpos[3 * i + r]; for (int k = 0; k < 3; k++) c[r] += xmat[9 * i + 3 * r + k] * bbc[3 * i + k]; }
      for (int r = 0; r < 3; r++) bbc[3 * i + r] = c[r];
    }
    // geom-frame bounding boxes for the oriented-box mid phase (hull geoms carry an identity geom frame: body / world axes)
    std::vector<double> gaabb(6 * ngeom, 0.0);
    for (int g = 0; g < ngeom; g++) {
      const int ty = b.I("geom_type")[g];
      const double *sz = &b.F("geom_size")[3 * g];
      double *o = &gaabb[6 * g];
      if (ty == G_HULL) {
        const int adr = b.I("geom_vertadr")[g], num = b.I("geom_vertnum")[g];
        double lo[3] = {1e30, 1e30, 1e30}, hi[3] = {-1e30, -1e30, -1e30};
        for (int v = adr; v < adr + num; v++)
          for (int r = 0; r < 3; r++) { lo[r] = std::min(lo[r], hv[3 * v + r]); hi[r] = std::max(hi[r], hv[3 * v + r]); }
        for (int r = 0; r < 3; r++) { o[r] = 0.5 * (lo[r] + hi[r]); o[3 + r] = 0.5 * (hi[r] - lo[r]); }
      } else if (ty == G_BOX) { o[3] = sz[0]; o[4] = sz[1]; o[5] = sz[2]; }
      else if (ty == G_CYLINDER) { o[3] = o[4] = sz[0]; o[5] = sz[1]; }
      else if (ty == G_CAPSULE) { o[3] = o[4] = sz[0]; o[5] = sz[0] + sz[1]; }
      else if (ty == G_SPHERE) { o[3] = o[4] = o[5] = sz[0]; }
      else { o[3] = o[4] = o[5] = 1e9; }  // plane: never used as a box
    }
    std::vector<Vec4<T>> verts(hv.size() / 3);
    for (size_t i = 0; i < verts.size(); i++) verts[i] = Vec4<T>{(T)hv[3 * i], (T)hv[3 * i + 1], (T)hv[3 * i + 2], T(0)};
    SceneModel<T> &d = dev;
    d.ngeom = ngeom; d.nbody = nbody; d.npair = (int)b.I("bodypair").size() / 2;
    d.geom_type = up(b.I("geom_type")); d.geom_body = up(b.I("geom_body")); d.geom_slot = up(gslot);
    d.geom_vertadr = up(b.I("geom_vertadr")); d.geom_vertnum = up(b.I("geom_vertnum"));
    d.geom_condim = up(b.I("geom_condim")); d.geom_priority = up(b.I("geom_priority"));
    d.geom_pos = up(cvt(gpos)); d.geom_mat = up(cvt(gmat)); d.geom_size = up(cvt(b.F("geom_size")));
    d.geom_bcenter = up(cvt(gbc)); d.geom_rbound = up(cvt(b.F("geom_rbound"))); d.geom_aabb = up(cvt(gaabb));
    d.geom_friction = up(cvt(b.F("geom_friction"))); d.geom_solref = up(cvt(b.F("geom_solref"))); d.geom_solimp = up(cvt(b.F("geom_solimp")));
    d.geom_solmix = up(cvt(b.F("geom_solmix"))); d.geom_margin = up(cvt(b.F("geom_margin"))); d.geom_gap = up(cvt(b.F("geom_gap")));
    d.hull_vert = up(verts); d.hull_nbradr = up(b.I("hull_nbradr")); d.hull_nbr = up(b.I("hull_nbr"));
    {
      const auto &nadr = b.I("hull_nbradr"), &nbr = b.I("hull_nbr");
      std::vector<Vec4<T>> nv(nbr.size());
      for (int g = 0; g < ngeom; g++) {
        if (b.I("geom_type")[g] != G_HULL) continue;
        const int adr = b.I("geom_vertadr")[g], num = b.I("geom_vertnum")[g];
        for (int v = adr; v < adr + num; v++)
          for (int k = nadr[v]; k < nadr[v + 1]; k++) {
            const int j = nbr[k], deg = nadr[adr + j + 1] - nadr[adr + j], off = nadr[adr + j] - nadr[adr];
            if (j >= (1 << NBR_ID_BITS) || deg >= (1 << NBR_DEG_BITS) || off >= (1 << NBR_ADR_BITS))
              throw std::runtime_error("hull adjacency does not fit the packed neighbour record");
            nv[k] = verts[adr + j];
            nv[k].w = nbr_pack<T>((unsigned)j | ((unsigned)deg << NBR_ID_BITS) | ((unsigned)off << (NBR_ID_BITS + NBR_DEG_BITS)));
          }
      }
      d.hull_nbrv = up(nv);
    }
    {
      std::vector<int> pstart; std::vector<unsigned> gpairs;
      const auto &bpv = b.I("bodypair"), &ga = b.I("body_geomadr"), &gn = b.I("body_geomnum");
      for (size_t p = 0; p < bpv.size() / 2; p++) {
        pstart.push_back((int)gpairs.size());
        const int b1 = bpv[2 * p], b2 = bpv[2 * p + 1];
        for (int g1 = ga[b1]; g1 < ga[b1] + gn[b1]; g1++)
          for (int g2 = ga[b2]; g2 < ga[b2] + gn[b2]; g2++) gpairs.push_back((unsigned)g1 | ((unsigned)g2 << 8));
      }
      pstart.push_back((int)gpairs.size());
      d.pair_start = up(pstart); d.geompair = up(gpairs);
    }
    d.bodypair = up(b.I("bodypair")); d.body_slot = up(slot); d.body_geomadr = up(b.I("body_geomadr")); d.body_geomnum = up(b.I("body_geomnum"));
    d.body_bcenter = up(cvt(bbc)); d.body_rbound = up(cvt(b.F("body_rbound"))); d.body_invweight0 = up(cvt(b.F("body_invweight0")));
    for (int p = 0; p < NPROP; p++) {
      const int body = b.I("prop_body")[p];
      const double m = b.F("body_mass")[body], *ip = &b.F("body_ipos")[3 * body], *in = &b.F("body_inertia")[3 * body];
      double Ri[9], Ic[9];
      q2m(&b.F("body_iquat")[4 * body], Ri);
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) Ic[3 * r + c] = Ri[3 * r] * in[0] * Ri[3 * c] + Ri[3 * r + 1] * in[1] * Ri[3 * c + 1] + Ri[3 * r + 2] * in[2] * Ri[3 * c + 2];
      const double pp = ip[0] * ip[0] + ip[1] * ip[1] + ip[2] * ip[2];
      double Io[9];
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) Io[3 * r + c] = Ic[3 * r + c] + m * ((r == c ? pp : 0.0) - ip[r] * ip[c]);
      d.prop_mass[p] = (T)m;
      for (int c = 0; c < 3; c++) d.prop_ipos[p][c] = (T)ip[c];
      const int idx[6] = {0, 4, 8, 1, 2, 5};
      for (int c = 0; c < 6; c++) { d.prop_Icom[p][c] = (T)Ic[idx[c]]; d.prop_Iorg[p][c] = (T)Io[idx[c]]; }
      for (int c = 0; c < 9; c++) d.prop_Riq[p][c] = (T)Ri[c];
    }
    for (int c = 0; c < 6; c++) d.reward_obj_box[c] = (T)b.F("reward_obj_box")[c];
    d.nreward_box = (int)b.F("reward_box_pos").size() / 3;
    if (d.nreward_box < 1 || d.nreward_box > MAXREWARDBOX || b.F("reward_box_half").size() != b.F("reward_box_pos").size())
      throw std::runtime_error("model blob: unsupported number of reward overlap boxes");
    for (int k = 0; k < d.nreward_box; k++)
      for (int c = 0; c < 3; c++) { d.reward_box_pos[k][c] = (T)b.F("reward_box_pos")[3 * k + c]; d.reward_box_half[k][c] = (T)b.F("reward_box_half")[3 * k + c]; }
    d.impratio = (T)b.F("opt")[4]; d.timestep = (T)b.F("opt")[0];
  }
};

}  // namespace so101
