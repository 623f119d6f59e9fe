#include "mesh.hpp"

#include <cmath>
#include <stdexcept>
#include <unordered_map>

namespace bloch_b200 {

namespace {

struct Key {
  int64_t a, b, c;
  bool operator==(const Key &o) const { return a == o.a && b == o.b && c == o.c; }
};
struct KeyHash {
  size_t operator()(const Key &k) const {
    uint64_t h = (uint64_t)k.a * 0x9E3779B97F4A7C15ull;
    h ^= (uint64_t)k.b + 0x7F4A7C159E3779B9ull + (h << 6) + (h >> 2);
    h ^= (uint64_t)k.c + 0x3779B97F4A7C159Eull + (h << 6) + (h >> 2);
    return (size_t)h;
  }
};

constexpr int64_t QS = int64_t(1) << 30;

// position modulo the lattice -> quantised fractional coordinates.  All entity centres of a
// uniformly subdivided WS cell have rational fractional coordinates with small denominators,
// so the quantisation never sits on a rounding boundary.
inline Key key_of(const double rec[9], const double x[3]) {
  int64_t q[3];
  for (int i = 0; i < 3; i++) {
    double f = rec[3 * i] * x[0] + rec[3 * i + 1] * x[1] + rec[3 * i + 2] * x[2];
    f -= std::floor(f);
    q[i] = (int64_t)std::llround(f * (double)QS) % QS;
  }
  return Key{q[0], q[1], q[2]};
}

// make v lexicographically positive; returns the sign that was applied
inline int canon(double v[3]) {
  double nrm = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  for (int c = 0; c < 3; c++) {
    if (std::fabs(v[c]) > 1e-9 * nrm) {
      if (v[c] < 0) { v[0] = -v[0]; v[1] = -v[1]; v[2] = -v[2]; return -1; }
      return 1;
    }
  }
  return 1;
}
inline bool lex_less(const double a[3], const double b[3]) {
  for (int c = 0; c < 3; c++) {
    double tol = 1e-9 * (std::fabs(a[c]) + std::fabs(b[c]) + 1e-300);
    if (a[c] < b[c] - tol) return true;
    if (a[c] > b[c] + tol) return false;
  }
  return false;
}

struct FaceFrame { int g[2]; int s[2]; int sn; };   // local in-face dir k -> (global axis g, sign s)

inline void other_dirs(int d, int &d1, int &d2) {
  d1 = (d == 0) ? 1 : 0;
  d2 = (d == 2) ? 1 : 2;
}

}  // namespace

void HexMesh::centers(std::vector<double> &c) const {
  c.resize(3 * (size_t)n_elem);
  for (int e = 0; e < n_elem; e++) {
    const double *Jc = &J[9 * cls[e]];
    for (int i = 0; i < 3; i++)
      c[3 * e + i] = x0[3 * e + i] + 0.5 * (Jc[3 * i] + Jc[3 * i + 1] + Jc[3 * i + 2]);
  }
}

void build_mesh(const std::vector<std::array<double, 3>> &vert,
                const std::vector<std::array<int, 8>> &hex, const double rec[9], int n,
                HexMesh &mesh) {
  if (n < 1) throw std::runtime_error("n_sub must be >= 1");
  mesh = HexMesh();
  mesh.n_sub = n;
  mesh.n_class = (int)hex.size();
  mesh.n_elem = mesh.n_class * n * n * n;
  mesh.rec.assign(rec, rec + 9);
  mesh.J.resize(9 * hex.size());
  mesh.x0.reserve(3 * (size_t)mesh.n_elem);
  mesh.cls.reserve(mesh.n_elem);
  static const int ref[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0},
                                {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
  for (size_t c = 0; c < hex.size(); c++) {
    const auto &h = hex[c];
    double Jc[9];
    const auto &v0 = vert[h[0]];
    for (int i = 0; i < 3; i++) {
      Jc[3 * i + 0] = vert[h[1]][i] - v0[i];
      Jc[3 * i + 1] = vert[h[3]][i] - v0[i];
      Jc[3 * i + 2] = vert[h[4]][i] - v0[i];
    }
    for (int k = 0; k < 8; k++)   // is_affine assert
      for (int i = 0; i < 3; i++) {
        double x = v0[i] + Jc[3 * i] * ref[k][0] + Jc[3 * i + 1] * ref[k][1] + Jc[3 * i + 2] * ref[k][2];
        if (std::fabs(x - vert[h[k]][i]) > 1e-12)
          throw std::runtime_error("coarse hex is not a parallelepiped (only affine cells are supported)");
      }
    double det = Jc[0] * (Jc[4] * Jc[8] - Jc[5] * Jc[7]) - Jc[1] * (Jc[3] * Jc[8] - Jc[5] * Jc[6]) +
                 Jc[2] * (Jc[3] * Jc[7] - Jc[4] * Jc[6]);
    if (det < 0) {   // inverted cell: swap the first two local axes (MFEM reorders vertices)
      for (int i = 0; i < 3; i++) std::swap(Jc[3 * i], Jc[3 * i + 1]);
      det = -det;
    }
    if (det < 1e-14) throw std::runtime_error("degenerate coarse hex");
    for (int k = 0; k < 9; k++) mesh.J[9 * c + k] = Jc[k] / n;
    mesh.volume += det;
    for (int kz = 0; kz < n; kz++)
      for (int jy = 0; jy < n; jy++)
        for (int ix = 0; ix < n; ix++) {
          for (int i = 0; i < 3; i++)
            mesh.x0.push_back(v0[i] + (Jc[3 * i] * ix + Jc[3 * i + 1] * jy + Jc[3 * i + 2] * kz) / n);
          mesh.cls.push_back((int)c);
        }
  }
}

void build_ws_mesh(const bravais::BravaisLattice &lat, int n_sub, HexMesh &mesh) {
  std::vector<bravais::Vec3> b;
  lat.GetReciprocalLatticeVectors(b);
  double rec[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) rec[3 * i + j] = b[i][j];
  build_mesh(lat.WignerSeitzVertices(), lat.WignerSeitzHexes(), rec, n_sub, mesh);
}

void build_dofmaps(HexMesh &mesh, int p, DofMaps &M) {
  if (p < 1) throw std::runtime_error("order must be >= 1");
  M = DofMaps();
  M.p = p;
  const int P1 = p + 1;
  M.l_h1 = P1 * P1 * P1;
  M.l_nd = 3 * p * P1 * P1;
  M.l_rt = 3 * p * p * P1;
  const int ne = mesh.n_elem;
  M.h1.assign((size_t)ne * M.l_h1, 0);
  M.nd.assign((size_t)ne * M.l_nd, 0);
  M.rt.assign((size_t)ne * M.l_rt, 0);
  const double *rec = mesh.rec.data();

  // per-class orientation tables (translation invariant)
  std::vector<std::array<int, 3>> edge_sign(mesh.n_class);
  std::vector<std::array<FaceFrame, 3>> frames(mesh.n_class);
  for (int c = 0; c < mesh.n_class; c++) {
    const double *Jc = &mesh.J[9 * c];
    for (int d = 0; d < 3; d++) {
      double v[3] = {Jc[d], Jc[3 + d], Jc[6 + d]};
      edge_sign[c][d] = canon(v);
      int d1, d2;
      other_dirs(d, d1, d2);
      double u[3] = {Jc[d1], Jc[3 + d1], Jc[6 + d1]}, w[3] = {Jc[d2], Jc[3 + d2], Jc[6 + d2]};
      FaceFrame f;
      f.s[0] = canon(u);
      f.s[1] = canon(w);
      if (lex_less(u, w)) { f.g[0] = 0; f.g[1] = 1; } else { f.g[0] = 1; f.g[1] = 0; }
      // physical normal carried by the RT dof functional: adj(J)^T e_d  ||  (J e_{d+1}) x (J e_{d+2})
      int a = (d + 1) % 3, b = (d + 2) % 3;
      double A[3] = {Jc[a], Jc[3 + a], Jc[6 + a]}, B[3] = {Jc[b], Jc[3 + b], Jc[6 + b]};
      double nrm[3] = {A[1] * B[2] - A[2] * B[1], A[2] * B[0] - A[0] * B[2], A[0] * B[1] - A[1] * B[0]};
      f.sn = canon(nrm);
      frames[c][d] = f;
    }
  }

  std::unordered_map<Key, int, KeyHash> vmap, emap, fmap;
  vmap.reserve((size_t)ne * 2);
  emap.reserve((size_t)ne * 4);
  fmap.reserve((size_t)ne * 4);
  std::vector<long> v_h1, e_h1, e_nd, f_h1, f_nd, f_rt;   // DOF base per entity
  long c_h1 = 0, c_nd = 0, c_rt = 0;
  const int eh = p - 1, fh = (p - 1) * (p - 1), ih = (p - 1) * (p - 1) * (p - 1);
  const int en = p, fn = 2 * p * (p - 1), in_ = 3 * p * (p - 1) * (p - 1);
  const int fr = p * p, ir = 3 * p * p * (p - 1);

  for (int e = 0; e < ne; e++) {
    const int c = mesh.cls[e];
    const double *Jc = &mesh.J[9 * c];
    const double *x0 = &mesh.x0[3 * e];
    auto phys = [&](double r0, double r1, double r2, double x[3]) {
      for (int i = 0; i < 3; i++) x[i] = x0[i] + Jc[3 * i] * r0 + Jc[3 * i + 1] * r1 + Jc[3 * i + 2] * r2;
    };
    int vid[2][2][2], eid[3][2][2], fid[3][2];
    double x[3];
    for (int a2 = 0; a2 < 2; a2++)
      for (int a1 = 0; a1 < 2; a1++)
        for (int a0 = 0; a0 < 2; a0++) {
          phys(a0, a1, a2, x);
          auto it = vmap.emplace(key_of(rec, x), (int)vmap.size());
          if (it.second) { v_h1.push_back(c_h1); c_h1 += 1; }
          vid[a0][a1][a2] = it.first->second;
        }
    for (int d = 0; d < 3; d++) {
      int d1, d2;
      other_dirs(d, d1, d2);
      for (int b2 = 0; b2 < 2; b2++)
        for (int b1 = 0; b1 < 2; b1++) {
          double r[3];
          r[d] = 0.5; r[d1] = b1; r[d2] = b2;
          phys(r[0], r[1], r[2], x);
          auto it = emap.emplace(key_of(rec, x), (int)emap.size());
          if (it.second) {
            e_h1.push_back(c_h1); c_h1 += eh;
            e_nd.push_back(c_nd); c_nd += en;
          }
          eid[d][b1][b2] = it.first->second;
        }
    }
    for (int d = 0; d < 3; d++) {
      int d1, d2;
      other_dirs(d, d1, d2);
      for (int a = 0; a < 2; a++) {
        double r[3];
        r[d] = a; r[d1] = 0.5; r[d2] = 0.5;
        phys(r[0], r[1], r[2], x);
        auto it = fmap.emplace(key_of(rec, x), (int)fmap.size());
        if (it.second) {
          f_h1.push_back(c_h1); c_h1 += fh;
          f_nd.push_back(c_nd); c_nd += fn;
          f_rt.push_back(c_rt); c_rt += fr;
        }
        fid[d][a] = it.first->second;
      }
    }
    const long i_h1 = c_h1, i_nd = c_nd, i_rt = c_rt;
    c_h1 += ih; c_nd += in_; c_rt += ir;

    // ---- H1 ----
    {
      int32_t *out = &M.h1[(size_t)e * M.l_h1];
      int icount = 0;
      for (int a2 = 0; a2 <= p; a2++)
        for (int a1 = 0; a1 <= p; a1++)
          for (int a0 = 0; a0 <= p; a0++) {
            int a[3] = {a0, a1, a2};
            bool bd[3];
            int nb = 0;
            for (int d = 0; d < 3; d++) { bd[d] = (a[d] == 0 || a[d] == p); nb += bd[d]; }
            long gid;
            if (nb == 3) {
              gid = v_h1[vid[a0 / p][a1 / p][a2 / p]];
            } else if (nb == 2) {
              int d = !bd[0] ? 0 : (!bd[1] ? 1 : 2);
              int d1, d2;
              other_dirs(d, d1, d2);
              int t = edge_sign[c][d] > 0 ? a[d] : p - a[d];
              gid = e_h1[eid[d][a[d1] / p][a[d2] / p]] + (t - 1);
            } else if (nb == 1) {
              int d = bd[0] ? 0 : (bd[1] ? 1 : 2);
              int d1, d2;
              other_dirs(d, d1, d2);
              const FaceFrame &f = frames[c][d];
              int t[2];
              t[f.g[0]] = f.s[0] > 0 ? a[d1] : p - a[d1];
              t[f.g[1]] = f.s[1] > 0 ? a[d2] : p - a[d2];
              gid = f_h1[fid[d][a[d] / p]] + (t[0] - 1) + (long)(p - 1) * (t[1] - 1);
            } else {
              gid = i_h1 + icount++;
            }
            out[a0 + P1 * (a1 + P1 * a2)] = (int32_t)(gid + 1);
          }
    }
    // ---- ND ----
    {
      int32_t *out = &M.nd[(size_t)e * M.l_nd];
      int icount = 0;
      const int nb_ = p * P1 * P1;
      for (int cc = 0; cc < 3; cc++) {
        int dims[3] = {P1, P1, P1};
        dims[cc] = p;
        int d1, d2;
        other_dirs(cc, d1, d2);
        for (int a2 = 0; a2 < dims[2]; a2++)
          for (int a1 = 0; a1 < dims[1]; a1++)
            for (int a0 = 0; a0 < dims[0]; a0++) {
              int a[3] = {a0, a1, a2};
              bool b1 = (a[d1] == 0 || a[d1] == p), b2 = (a[d2] == 0 || a[d2] == p);
              long gid;
              int sgn = 1;
              if (b1 && b2) {
                sgn = edge_sign[c][cc];
                int t = sgn > 0 ? a[cc] : p - 1 - a[cc];
                gid = e_nd[eid[cc][a[d1] / p][a[d2] / p]] + t;
              } else if (b1 || b2) {
                int dn = b1 ? d1 : d2;     // face normal
                int dw = b1 ? d2 : d1;     // the closed in-face direction
                int f1, f2;
                other_dirs(dn, f1, f2);    // local in-face dirs of that face, increasing
                const FaceFrame &f = frames[c][dn];
                int kc = (cc == f1) ? 0 : 1, kw = 1 - kc;   // positions of cc / dw in (f1,f2)
                (void)f2;
                int gc = f.g[kc], sc = f.s[kc], sw = f.s[kw];
                int to = sc > 0 ? a[cc] : p - 1 - a[cc];
                int tc = sw > 0 ? a[dw] : p - a[dw];
                long idx = (gc == 0) ? (to + (long)p * (tc - 1))
                                     : ((long)p * (p - 1) + (tc - 1) + (long)(p - 1) * to);
                gid = f_nd[fid[dn][a[dn] / p]] + idx;
                sgn = sc;
              } else {
                gid = i_nd + icount++;
              }
              int li = cc * nb_ + a0 + dims[0] * (a1 + dims[1] * a2);
              out[li] = (int32_t)(sgn * (gid + 1));
            }
      }
    }
    // ---- RT ----
    {
      int32_t *out = &M.rt[(size_t)e * M.l_rt];
      int icount = 0;
      const int nb_ = p * p * P1;
      for (int cc = 0; cc < 3; cc++) {
        int dims[3] = {p, p, p};
        dims[cc] = P1;
        int d1, d2;
        other_dirs(cc, d1, d2);
        for (int a2 = 0; a2 < dims[2]; a2++)
          for (int a1 = 0; a1 < dims[1]; a1++)
            for (int a0 = 0; a0 < dims[0]; a0++) {
              int a[3] = {a0, a1, a2};
              long gid;
              int sgn = 1;
              if (a[cc] == 0 || a[cc] == p) {
                const FaceFrame &f = frames[c][cc];
                int t[2];
                t[f.g[0]] = f.s[0] > 0 ? a[d1] : p - 1 - a[d1];
                t[f.g[1]] = f.s[1] > 0 ? a[d2] : p - 1 - a[d2];
                gid = f_rt[fid[cc][a[cc] / p]] + t[0] + (long)p * t[1];
                sgn = f.sn;
              } else {
                gid = i_rt + icount++;
              }
              int li = cc * nb_ + a0 + dims[0] * (a1 + dims[1] * a2);
              out[li] = (int32_t)(sgn * (gid + 1));
            }
      }
    }
  }
  M.n_h1 = c_h1; M.n_nd = c_nd; M.n_rt = c_rt;
  mesh.n_vert = (int)vmap.size();
  mesh.n_edge = (int)emap.size();
  mesh.n_face = (int)fmap.size();
}

}  // namespace bloch_b200
