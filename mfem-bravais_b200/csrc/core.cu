// Eigenproblem object + C ABI (include/bloch_b200.h).  Host logic mirrors
// MaxwellBlochWaveEquation (maxwell/maxwell_bloch.cpp): constructor :34-140, SetKappa :200-210,
// Setup :337-620, GetEigenvalues :1052-1076, GetEigenvectorE/B :1371-1458.
#include "core.hpp"
#include "dense.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "../../include/bloch_b200.h"

using namespace bloch_b200;
using D2 = double2;

static thread_local std::string g_last_error;

#define API_BEGIN try {
#define API_END                                                   \
  }                                                               \
  catch (const CudaError &e) { g_last_error = e.what(); return BLOCH_ERR_CUDA; }      \
  catch (const std::invalid_argument &e) { g_last_error = e.what(); return BLOCH_ERR_ARG; } \
  catch (const std::exception &e) { g_last_error = e.what(); return BLOCH_ERR_INTERNAL; }   \
  catch (...) { g_last_error = "unknown error"; return BLOCH_ERR_INTERNAL; }
#define REQUIRE(cond, msg) \
  do { if (!(cond)) throw std::invalid_argument(msg); } while (0)

// ------------------------------------------------------------------------------------------
// handle internals
// ------------------------------------------------------------------------------------------
static void build_kernel_maps(bloch_handle_s *h) {
  const int p = h->p, Q = p + 1, ne = h->mesh.n_elem;
  const DofMaps &M = h->maps;
  std::vector<int> perm_nd(M.l_nd), perm_rt(M.l_rt), perm_h1(M.l_h1);
  const int nb = p * Q * Q, rb = p * p * Q;
  for (int c = 0; c < 3; c++) {
    for (int o = 0; o < p; o++)
      for (int j1 = 0; j1 < Q; j1++)
        for (int j2 = 0; j2 < Q; j2++) {
          int a[3], dims[3] = {Q, Q, Q};
          dims[c] = p;
          a[c] = o; a[(c + 1) % 3] = j1; a[(c + 2) % 3] = j2;
          perm_nd[c * nb + (o * Q + j1) * Q + j2] = c * nb + a[0] + dims[0] * (a[1] + dims[1] * a[2]);
        }
    for (int j = 0; j < Q; j++)
      for (int o1 = 0; o1 < p; o1++)
        for (int o2 = 0; o2 < p; o2++) {
          int a[3], dims[3] = {p, p, p};
          dims[c] = Q;
          a[c] = j; a[(c + 1) % 3] = o1; a[(c + 2) % 3] = o2;
          perm_rt[c * rb + (j * p + o1) * p + o2] = c * rb + a[0] + dims[0] * (a[1] + dims[1] * a[2]);
        }
  }
  for (int i0 = 0; i0 < Q; i0++)
    for (int i1 = 0; i1 < Q; i1++)
      for (int i2 = 0; i2 < Q; i2++) perm_h1[(i0 * Q + i1) * Q + i2] = i0 + Q * (i1 + Q * i2);
  std::vector<int32_t> knd((size_t)ne * M.l_nd), krt((size_t)ne * M.l_rt), kh1((size_t)ne * M.l_h1);
  for (int e = 0; e < ne; e++) {
    for (int k = 0; k < M.l_nd; k++) knd[(size_t)e * M.l_nd + k] = M.nd[(size_t)e * M.l_nd + perm_nd[k]];
    for (int k = 0; k < M.l_rt; k++) krt[(size_t)e * M.l_rt + k] = M.rt[(size_t)e * M.l_rt + perm_rt[k]];
    for (int k = 0; k < M.l_h1; k++) kh1[(size_t)e * M.l_h1 + k] = M.h1[(size_t)e * M.l_h1 + perm_h1[k]];
  }
  // transpose of the ND map for the atomic-free second pass (counting sort by dof)
  {
    const size_t nnz = (size_t)ne * M.l_nd;
    std::vector<int> ptr(h->maps.n_nd + 1, 0);
    for (size_t t = 0; t < nnz; t++) ptr[std::abs(knd[t])]++;        // |s| = gid + 1
    for (long g = 0; g < h->maps.n_nd; g++) ptr[g + 1] += ptr[g];
    std::vector<int32_t> loc(nnz);
    std::vector<int> fill(ptr.begin(), ptr.end() - 1);
    for (size_t t = 0; t < nnz; t++) {
      const int s = knd[t];
      const int g = std::abs(s) - 1;
      loc[fill[g]++] = (int32_t)(s < 0 ? -(long)(t + 1) : (long)(t + 1));
    }
    h->d_tp_ptr.upload(ptr, h->stream);
    h->d_tp_loc.upload(loc, h->stream);
    // rows the apply kernels REDUCE into (every dof that is not interior to a single element): the only rows that have
    // to be cleared before a fresh apply, the interior ones are written with plain stores (ElemData::partial_clear)
    std::vector<char> interior(h->maps.n_nd, 0);
    for (int e = 0; e < ne; e++)
      for (int k = 0; k < M.l_nd; k++) {
        const int j1 = (k % nb) / Q % Q, j2 = k % Q;
        const int g = std::abs(knd[(size_t)e * M.l_nd + k]) - 1;
        if (j1 > 0 && j1 < p && j2 > 0 && j2 < p && ptr[g + 1] - ptr[g] == 1) interior[g] = 1;
      }
    std::vector<int32_t> shared;
    shared.reserve(h->maps.n_nd);
    for (long g = 0; g < h->maps.n_nd; g++)
      if (!interior[g]) shared.push_back((int32_t)g);
    h->n_shared_rows = (long)shared.size();
    if (!shared.empty()) h->d_shared_rows.upload(shared, h->stream);
    BLOCH_CUDA(cudaStreamSynchronize(h->stream));
  }
  h->knd_host = knd;
  h->d_map_nd.upload(knd, h->stream);
  h->d_map_rt.upload(krt, h->stream);
  h->d_map_h1.upload(kh1, h->stream);
  h->d_cls.upload(h->mesh.cls, h->stream);
  BLOCH_CUDA(cudaStreamSynchronize(h->stream));
}

static void fill_tabs(const Basis1D &B, Tabs &T) {
  std::memset(&T, 0, sizeof(T));
  const int p = B.p, q = p + 1;
  for (int r = 0; r < q; r++)
    for (int j = 0; j < q; j++) { T.TI[r][j] = B.TI[r * q + j]; T.TIinv[r][j] = B.TIinv[r * q + j]; }
  for (int a = 0; a < p; a++)
    for (int r = 0; r < q; r++) T.Dt[a][r] = B.Dt[a * q + r];
  for (int r = 0; r < q; r++) T.om[r] = B.om[r];
}

namespace bloch_b200 {
void class_params(const double *J, const double kappa[3], double out[kClassParDoubles]) {
  double JtJ[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += J[3 * k + i] * J[3 * k + j];
      JtJ[3 * i + j] = s;
    }
  const double det = J[0] * (J[4] * J[8] - J[5] * J[7]) - J[1] * (J[3] * J[8] - J[5] * J[6]) +
                     J[2] * (J[3] * J[7] - J[4] * J[6]);
  for (int d = 0; d < 3; d++) out[d] = J[d] * kappa[0] + J[3 + d] * kappa[1] + J[6 + d] * kappa[2];
  for (int k = 0; k < 9; k++) out[3 + k] = JtJ[k] / det;
  // H = det * (J^T J)^-1 via the adjugate of the symmetric JtJ (det(JtJ) = det^2)
  const double *a = JtJ;
  double adj[9];
  adj[0] = a[4] * a[8] - a[5] * a[7]; adj[1] = a[2] * a[7] - a[1] * a[8]; adj[2] = a[1] * a[5] - a[2] * a[4];
  adj[3] = a[5] * a[6] - a[3] * a[8]; adj[4] = a[0] * a[8] - a[2] * a[6]; adj[5] = a[2] * a[3] - a[0] * a[5];
  adj[6] = a[3] * a[7] - a[4] * a[6]; adj[7] = a[1] * a[6] - a[0] * a[7]; adj[8] = a[0] * a[4] - a[1] * a[3];
  for (int k = 0; k < 9; k++) out[12 + k] = adj[k] / det;
  out[21] = det;
}
}  // namespace bloch_b200

// element-local diagonals of A, M1 and S0 per class, obtained by running the production kernels
// on a probe problem (one private copy of the element per local unit vector)
// largest eigenvalue of diag(X)^-1/2 X diag(X)^-1/2 for a Hermitian PSD local matrix (power iteration)
namespace bloch_b200 {
double local_scaled_lmax(int L, const D2 *X /* column k at X[k*L + l] */) {
  std::vector<double> d(L);
  for (int k = 0; k < L; k++) d[k] = X[(size_t)k * L + k].x > 0 ? 1.0 / std::sqrt(X[(size_t)k * L + k].x) : 0.0;
  std::vector<double> vr(L), vi(L, 0.0), wr(L), wi(L);
  for (int k = 0; k < L; k++) vr[k] = 1.0 + 0.37 * std::sin(1.0 + 2.3 * k);
  double lam = 0;
  for (int it = 0; it < 200; it++) {
    for (int l = 0; l < L; l++) { wr[l] = 0; wi[l] = 0; }
    for (int k = 0; k < L; k++) {
      const double xr = vr[k] * d[k], xi = vi[k] * d[k];
      const D2 *col = X + (size_t)k * L;
      for (int l = 0; l < L; l++) {
        wr[l] += col[l].x * xr - col[l].y * xi;
        wi[l] += col[l].x * xi + col[l].y * xr;
      }
    }
    double nw = 0, nv = 0;
    for (int l = 0; l < L; l++) {
      wr[l] *= d[l]; wi[l] *= d[l];
      nw += wr[l] * wr[l] + wi[l] * wi[l];
      nv += vr[l] * vr[l] + vi[l] * vi[l];
    }
    const double nl = std::sqrt(nw / nv);
    const bool conv = std::fabs(nl - lam) < 1e-4 * nl;
    lam = nl;
    const double sc = 1.0 / std::sqrt(nw);
    for (int l = 0; l < L; l++) { vr[l] = wr[l] * sc; vi[l] = wi[l] * sc; }
    if (conv && it > 20) break;
  }
  return lam;
}
}  // namespace bloch_b200

static void probe_diagonals(bloch_handle_s *h, std::vector<double> &dA, std::vector<double> &dM,
                            std::vector<double> &dS0, std::vector<double> &dM0, double *lmax_bound,
                            double *lmax_bound_h1) {
  const int nc = h->mesh.n_class, Ln = h->L_nd, Lh = h->L_h1;
  cudaStream_t s = h->stream;
  auto run = [&](int L, bool h1space, std::vector<double> *outA, std::vector<double> *outM) {
    const int ne = nc * L, nk = h->nk;
    bloch_handle_s::ProbeWork &pw = h1space ? h->probe_h1 : h->probe_nd;
    if (!pw.built || pw.built_nk != nk) {
      std::vector<int32_t> map((size_t)ne * L);
      std::vector<int> cls(ne);
      std::vector<double> one(ne, 1.0);
      std::vector<D2> x((size_t)ne * L * nk, make_double2(0.0, 0.0));
      for (int c = 0; c < nc; c++)
        for (int k = 0; k < L; k++) {
          const int e = c * L + k;
          cls[e] = c;
          for (int l = 0; l < L; l++) map[(size_t)e * L + l] = (int32_t)((size_t)e * L + l + 1);
          for (int kk = 0; kk < nk; kk++) x[((size_t)e * L + k) * nk + kk].x = 1.0;
        }
      pw.map.upload(map, s); pw.cls.upload(cls, s); pw.one.upload(one, s); pw.x.upload(x, s);
      pw.y.alloc(x.size());
      pw.hy.resize(x.size());
      BLOCH_CUDA(cudaStreamSynchronize(s));
      pw.built = true;
      pw.built_nk = nk;
    }
    DevBuf<int32_t> &dmap = pw.map; DevBuf<int> &dcls = pw.cls; DevBuf<double> &done = pw.one;
    DevBuf<D2> &dx = pw.x, &dy = pw.y;
    const size_t xsize = (size_t)ne * L * nk;
    ElemData E = h->E;
    E.n_elem = ne; E.cls = dcls.p; E.eps = done.p; E.muinv = done.p;
    if (h1space) E.map_h1 = dmap.p; else E.map_nd = dmap.p;
    std::vector<D2> &y = pw.hy;
    auto fetch = [&](std::vector<double> *out) {   // out[(k * nc + c) * L + l]: element-local diagonals per k-point
      BLOCH_CUDA(cudaMemcpyAsync(y.data(), dy.p, sizeof(D2) * xsize, cudaMemcpyDeviceToHost, s));
      BLOCH_CUDA(cudaStreamSynchronize(s));
      out->resize((size_t)nk * nc * L);
      for (int kk = 0; kk < nk; kk++)
        for (int c = 0; c < nc; c++)
          for (int k = 0; k < L; k++) (*out)[((size_t)kk * nc + c) * L + k] = y[(((size_t)(c * L + k)) * L + k) * nk + kk].x;
      double *bound = h1space ? lmax_bound_h1 : lmax_bound;
      if (bound) {  // lambda_max(D^-1 (a A + b M)) <= max_class max(mu_A, mu_M)
        std::vector<D2> loc((size_t)L * L);
        for (int kk = 0; kk < nk; kk++)
          for (int c = 0; c < nc; c++) {
            for (size_t t = 0; t < (size_t)L * L; t++) loc[t] = y[((size_t)c * L * L + t) * nk + kk];
            *bound = std::max(*bound, local_scaled_lmax(L, loc.data()));
          }
      }
    };
    if (h1space) {
      BLOCH_CUDA(cudaMemsetAsync(dy.p, 0, sizeof(D2) * xsize, s));
      BLOCH_CUDA(launch_h1_op(h->p, 3, h->tabs, E, dx.p, nk, dy.p, nk, nk, s, 1.0, 0.0));
      h->count_launch();
      fetch(outA);
      BLOCH_CUDA(cudaMemsetAsync(dy.p, 0, sizeof(D2) * xsize, s));
      BLOCH_CUDA(launch_h1_op(h->p, 3, h->tabs, E, dx.p, nk, dy.p, nk, nk, s, 0.0, 1.0));
      h->count_launch();
      fetch(outM);
    } else {
      BLOCH_CUDA(cudaMemsetAsync(dy.p, 0, sizeof(D2) * xsize, s));
      BLOCH_CUDA(launch_nd_apply(h->p, h->tabs, E, dx.p, nk, dy.p, nk, nk, 1.0, 0.0, s));
      h->count_launch();
      fetch(outA);
      BLOCH_CUDA(cudaMemsetAsync(dy.p, 0, sizeof(D2) * xsize, s));
      BLOCH_CUDA(launch_nd_apply(h->p, h->tabs, E, dx.p, nk, dy.p, nk, nk, 0.0, 1.0, s));
      h->count_launch();
      fetch(outM);
    }
  };
  if (h->p <= 3) run(Ln, false, &dA, &dM);
  run(Lh, true, &dS0, &dM0);
}

void bloch_handle_s::setup() {
  if (device < 0) throw std::invalid_argument("topology-only handle (BLOCH_DEVICE_NONE): no compute available");
  BLOCH_CUDA(cudaSetDevice(device));
  if (dirty_coef) {
    d_eps.upload(eps, stream);
    d_muinv.upload(muinv, stream);
  }
  E.n_elem = mesh.n_elem;
  E.n_class = mesh.n_class;
  E.nk = nk;
  E.cls = d_cls.p;
  E.eps = d_eps.p;
  E.muinv = d_muinv.p;
  E.map_nd = d_map_nd.p;
  E.map_h1 = d_map_h1.p;
  E.map_rt = d_map_rt.p;
  if (dirty_kappa) {
    std::vector<double> cp((size_t)nk * mesh.n_class * kClassParDoubles);
    std::vector<int> gf(nk, 0);
    betas.assign(nk, 0.0);
    any_gamma = false;
    for (int k = 0; k < nk; k++) {
      const double *kp = &kappas[3 * k];
      for (int c = 0; c < mesh.n_class; c++)
        class_params(&mesh.J[9 * c], kp, &cp[((size_t)k * mesh.n_class + c) * kClassParDoubles]);
      betas[k] = std::sqrt(kp[0] * kp[0] + kp[1] * kp[1] + kp[2] * kp[2]);
      gf[k] = betas[k] == 0.0 ? 1 : 0;
      any_gamma = any_gamma || gf[k];
    }
    d_cpar.upload(cp, stream);
    d_gflag.upload(gf, stream);
    BLOCH_CUDA(cudaStreamSynchronize(stream));   // cp / gf are stack temporaries
    for (int d = 0; d < 3; d++) kappa[d] = kappas[d];
    beta = betas[0];
  }
  E.cpar = d_cpar.p;
  if (dirty_kappa || dirty_coef) {
    // Jacobi diagonals of A, M and S0 = G^H M G (element-local diagonals from a probe launch)
    std::vector<double> dA, dM, dS0, dM0;
    double bound = 0.0, bound_h1 = 0.0;
    probe_diagonals(this, dA, dM, dS0, dM0, (lmax_local <= 0.0) ? &bound : nullptr,
                    (lmax_local_h1 <= 0.0) ? &bound_h1 : nullptr);
    if (lmax_local <= 0.0) lmax_local = bound;
    if (lmax_local_h1 <= 0.0) lmax_local_h1 = bound_h1;
    DevBuf<double> &dl = d_dloc;
    // Jacobi diagonals are [n][nk]: one value per (dof, k-point)
    const int nc = mesh.n_class;
    d_diagA.alloc((size_t)N * nk); d_diagM.alloc((size_t)N * nk); d_diagS0.alloc((size_t)N0 * nk); d_diagM0.alloc((size_t)N0 * nk);
    BLOCH_CUDA(cudaMemsetAsync(d_diagA.p, 0, sizeof(double) * N * nk, stream));
    BLOCH_CUDA(cudaMemsetAsync(d_diagM.p, 0, sizeof(double) * N * nk, stream));
    BLOCH_CUDA(cudaMemsetAsync(d_diagS0.p, 0, sizeof(double) * N0 * nk, stream));
    BLOCH_CUDA(cudaMemsetAsync(d_diagM0.p, 0, sizeof(double) * N0 * nk, stream));
    if (p <= 3) {
      dl.upload(dA, stream);
      BLOCH_CUDA(launch_scatter_diag(d_map_nd.p, L_nd, d_cls.p, d_muinv.p, dl.p, mesh.n_elem, d_diagA.p, stream, nk, nc));
      BLOCH_CUDA(cudaStreamSynchronize(stream));
      dl.upload(dM, stream);
      BLOCH_CUDA(launch_scatter_diag(d_map_nd.p, L_nd, d_cls.p, d_eps.p, dl.p, mesh.n_elem, d_diagM.p, stream, nk, nc));
      BLOCH_CUDA(cudaStreamSynchronize(stream));
    }
    dl.upload(dM0, stream);
    BLOCH_CUDA(launch_scatter_diag(d_map_h1.p, L_h1, d_cls.p, d_muinv.p, dl.p, mesh.n_elem, d_diagM0.p, stream, nk, nc));
    BLOCH_CUDA(cudaStreamSynchronize(stream));
    dl.upload(dS0, stream);
    BLOCH_CUDA(launch_scatter_diag(d_map_h1.p, L_h1, d_cls.p, d_eps.p, dl.p, mesh.n_elem, d_diagS0.p, stream, nk, nc));
    BLOCH_CUDA(cudaStreamSynchronize(stream));
    count_launch(3);
  }
  if ((dirty_kappa || dirty_coef) && use_mg) {
    if (!mg) mg = mg_create(this);
    if (mg) mg_setup(mg, this);
    if (mg && use_aux && p <= 3) {
      if (!aux) aux = aux_create(this);
      if (aux) aux_setup(aux, this);
    }
    if (mg && scalar_ready) {   // scalar H1 problem in use on this handle: its preconditioner hierarchy (kind 2)
      const char *e_smg = std::getenv("BLOCH_SCALAR_MG");
      if (!e_smg || std::atoi(e_smg) != 0) {
        if (!mg_scalar) mg_scalar = mg_create(this, 2);
        if (mg_scalar) mg_setup(mg_scalar, this);
      }
    }
  }
  dirty_coef = dirty_kappa = false;
}

// Assembled A = S1 - i beta DKZ (which = 0) or M = M1(eps) (which = 1) in CSR, for the matrix dump of
// maxwell_dispersion.cpp:553-590.  The element matrices of every affine class come from the same probe
// launch Setup() uses for the Jacobi diagonals (the kernels applied to element-local unit vectors); the
// global rows are then merged on the host.  Debug / interchange path, not used by the solver.
void bloch_handle_s::assemble(int which) {
  if (p > 3) throw std::invalid_argument("order not supported");
  if (nk != 1) throw std::invalid_argument("matrix export needs a single-kappa handle (bloch_set_kappa)");
  if (dirty_coef || dirty_kappa) setup();
  cudaStream_t s = stream;
  const int L = L_nd, nc = mesh.n_class;
  ProbeWork &pw = probe_nd;                      // built by setup()
  ElemData Ep = E;
  Ep.n_elem = nc * L; Ep.cls = pw.cls.p; Ep.eps = pw.one.p; Ep.muinv = pw.one.p; Ep.map_nd = pw.map.p;
  const size_t xsize = (size_t)nc * L * L;
  BLOCH_CUDA(cudaMemsetAsync(pw.y.p, 0, sizeof(D2) * xsize, s));
  BLOCH_CUDA(launch_nd_apply(p, tabs, Ep, pw.x.p, 1, pw.y.p, 1, 1, which == 0 ? 1.0 : 0.0, which == 0 ? 0.0 : 1.0, s));
  count_launch();
  std::vector<D2> loc(xsize);                    // loc[(c*L + k)*L + l] = (X_c)[l][k]
  BLOCH_CUDA(cudaMemcpyAsync(loc.data(), pw.y.p, sizeof(D2) * xsize, cudaMemcpyDeviceToHost, s));
  BLOCH_CUDA(cudaStreamSynchronize(s));
  const std::vector<double> &coef = which == 0 ? muinv : eps;
  // dof -> its local copies (counting sort)
  const size_t nloc = (size_t)mesh.n_elem * L;
  std::vector<int64_t> ptr(N + 1, 0);
  for (size_t t = 0; t < nloc; t++) ptr[std::abs(knd_host[t])]++;
  for (long g = 0; g < N; g++) ptr[g + 1] += ptr[g];
  std::vector<int64_t> where(nloc), fill(ptr.begin(), ptr.end() - 1);
  for (size_t t = 0; t < nloc; t++) where[fill[std::abs(knd_host[t]) - 1]++] = (int64_t)t;
  csr_ptr.assign(N + 1, 0);
  csr_col.clear(); csr_re.clear(); csr_im.clear();
  struct Ent { int32_t col; double re, im; };
  std::vector<Ent> row;
  for (long g = 0; g < N; g++) {
    row.clear();
    for (int64_t q = ptr[g]; q < ptr[g + 1]; q++) {
      const int64_t t = where[q];
      const int e = (int)(t / L), l = (int)(t - (int64_t)e * L);
      const int c = mesh.cls[e];
      const double sl = knd_host[t] < 0 ? -coef[e] : coef[e];
      for (int k = 0; k < L; k++) {
        const int32_t sk = knd_host[(size_t)e * L + k];
        const D2 v = loc[((size_t)c * L + k) * L + l];
        if (v.x == 0.0 && v.y == 0.0) continue;
        const double sg = sk < 0 ? -sl : sl;
        row.push_back({std::abs(sk) - 1, sg * v.x, sg * v.y});
      }
    }
    std::sort(row.begin(), row.end(), [](const Ent &a, const Ent &b) { return a.col < b.col; });
    for (size_t i = 0; i < row.size();) {
      size_t j = i;
      double re = 0, im = 0;
      while (j < row.size() && row[j].col == row[i].col) { re += row[j].re; im += row[j].im; j++; }
      csr_col.push_back(row[i].col); csr_re.push_back(re); csr_im.push_back(im);
      i = j;
    }
    csr_ptr[g + 1] = (int64_t)csr_col.size();
  }
}

void bloch_handle_s::field_averages(int i, double out24[24]) {
  cudaStream_t s = stream;
  if (!avg_ready) {
    const int q = p + 1;
    Basis1D B = make_basis(p);
    std::vector<double> xq, wq, v, dv;
    detail::gauss_legendre01(q, xq, wq);
    std::memset(&avg_tabs, 0, sizeof(avg_tabs));
    for (int a = 0; a < q; a++) {
      avg_tabs.xq[a] = xq[a];
      avg_tabs.wq[a] = wq[a];
      detail::lagrange(B.g, xq[a], v, dv);
      for (int o = 0; o < p; o++) avg_tabs.bo[a][o] = v[o];
      detail::lagrange(B.l, xq[a], v, dv);
      for (int j = 0; j < q; j++) avg_tabs.bc[a][j] = v[j];
    }
    std::vector<double> geom((size_t)mesh.n_class * 18);
    for (int c = 0; c < mesh.n_class; c++) {
      const double *J = &mesh.J[9 * c];
      std::vector<double> Jm(J, J + 9), Ji;
      detail::invert(3, Jm, Ji);
      for (int k = 0; k < 9; k++) { geom[18 * c + k] = J[k]; geom[18 * c + 9 + k] = Ji[k]; }
    }
    d_geom.upload(geom, s);
    d_x0.upload(mesh.x0, s);
    BLOCH_CUDA(cudaStreamSynchronize(s));
    avg_ready = true;
  }
  if (dirty_coef || dirty_kappa) setup();
  d_fa_e.alloc(N); d_fa_b.alloc(Nrt);
  d_fa_part.alloc((size_t)mesh.n_elem * 12); d_fa_out.alloc(12);
  BLOCH_CUDA(cudaMemcpy2DAsync(d_fa_e.p, sizeof(D2), d_X.p + (size_t)sel * block + i, sizeof(D2) * nk * block, sizeof(D2), N,
                               cudaMemcpyDeviceToDevice, s));
  const ElemData Ek = elem_of_k(sel);
  BLOCH_CUDA(launch_curl(p, tabs, Ek, d_fa_e.p, 1, d_fa_b.p, 1, 1, s));
  BLOCH_CUDA(launch_field_avg(p, avg_tabs, Ek, d_x0.p, d_geom.p, &kappas[3 * sel], d_fa_e.p, 1, d_fa_b.p, 1, 1, d_fa_part.p, d_fa_out.p, s));
  count_launch(2);
  D2 o[12];
  BLOCH_CUDA(cudaMemcpyAsync(o, d_fa_out.p, sizeof(o), cudaMemcpyDeviceToHost, s));
  BLOCH_CUDA(cudaStreamSynchronize(s));
  // B = i C E / sqrt|lambda| (Bi = Re(C E), Br = -Im(C E), maxwell_bloch.cpp:1432-1457); output order of
  // the reference's argument list: Er, Ei, Br, Bi, Dr, Di, Hr, Hi
  const double lam = std::fabs(eigenvalues[(size_t)sel * nbands + i]);
  const double sc = lam > 0 ? 1.0 / std::sqrt(lam) : 1.0;
  for (int k = 0; k < 3; k++) {
    out24[k] = o[k].x;            out24[3 + k] = o[k].y;
    out24[6 + k] = -sc * o[3 + k].y;  out24[9 + k] = sc * o[3 + k].x;
    out24[12 + k] = o[6 + k].x;   out24[15 + k] = o[6 + k].y;
    out24[18 + k] = -sc * o[9 + k].y; out24[21 + k] = sc * o[9 + k].x;
  }
}

void bloch_handle_s::set_kappas(int n, const double *k3) {
  if (n != nk) {   // the block layout [N][nk * block] changes: no warm start across a change of the batch size
    have_vectors = 0;
    have_hist = 0;
    kappas_prev.clear();
    kappas_prev2.clear();
    have_vectors_s = 0;
    eigenvalues.clear();
    eigenvalues_s.clear();
  }
  nk = n;
  kappas.assign(k3, k3 + 3 * (size_t)n);
  for (int d = 0; d < 3; d++) kappa[d] = kappas[d];
  sel = 0;
  dirty_kappa = true;
}

void bloch_handle_s::apply_nd(const D2 *x, D2 *y, int nvec, double ca, double cm) {
  apply_nd_ld(x, nvec, y, nvec, nvec, ca, cm);
}
// y[:, :nvec] (pitch ldy) = ca A x + cm M x.  BLOCH_TWO_PASS=1: atomic-free two-pass (element kernel ->
// E-vector, then one owner-computes reduction per dof; deterministic, bitwise reproducible).  Default
// is the single-pass variant with fp64 atomics, which measured ~15% faster on B200.
void bloch_handle_s::apply_nd_ld(const D2 *x, int ldx, D2 *y, int ldy, int nvec, double ca, double cm) {
  if (p > 3) throw std::invalid_argument("the Nedelec operators support orders 1..3 (order 4: scalar H1 problem only)");
  if (nvec % nk != 0) throw std::invalid_argument("the number of vectors must be a multiple of the k-point batch size");
  if (two_pass) {
    d_evec.alloc((size_t)mesh.n_elem * L_nd * nvec);
    BLOCH_CUDA(launch_nd_apply(p, tabs, E, x, ldx, y, ldy, nvec, ca, cm, stream, d_evec.p));
    BLOCH_CUDA(launch_nd_reduce(d_tp_ptr.p, d_tp_loc.p, d_evec.p, y, N, nvec, ldy, stream));
    count_launch(2);
  } else {
    const int cat = prof_in_precond ? 7 : 1;
    prof_begin(cat);
    ElemData Ef = E;
    static const bool fresh = [] { const char *e = std::getenv("BLOCH_ND_FRESH_Y"); return !e || std::atoi(e) != 0; }();
    static const bool partial = [] { const char *e = std::getenv("BLOCH_ND_PARTIAL_CLEAR"); return !e || std::atoi(e) != 0; }();
    Ef.fresh_y = fresh ? 1 : 0;     // y is cleared right here and only this launch writes it
    Ef.n_rows_y = N;
    if (fresh && partial && p >= 3 && n_shared_rows > 0) {
      // only the rows the kernels reduce into: the 3p(p-1)^2 element-interior dofs per element (44 % of the rows at
      // order 3) are written with plain stores.  configs[2]: 50.4 -> 51.1 GDOF/s at 10 RHS, 53.2 -> 54.5 at 30; at order 2
      // (25 % interior rows) the contiguous memset is faster than the row-wise clear (49.9 vs 48.5), so it stays
      BLOCH_CUDA(launch_clear_rows(y, ldy, nvec, d_shared_rows.p, n_shared_rows, stream));
      Ef.partial_clear = 1;
      count_launch();
    } else {
      BLOCH_CUDA(cudaMemset2DAsync(y, sizeof(D2) * ldy, 0, sizeof(D2) * nvec, N, stream));
    }
    BLOCH_CUDA(launch_nd_apply(p, tabs, Ef, x, ldx, y, ldy, nvec, ca, cm, stream));
    count_launch();
    prof_end(cat);
  }
  if (ca != 0.0) stats.applies_A += nvec;
}
// scalar variant: y[:, :nvec] = ca G^H M1(k) G x + cm M0(m) x on H1 block vectors
void bloch_handle_s::apply_scalar_ld(const D2 *x, int ldx, D2 *y, int ldy, int nvec, double ca, double cm) {
  BLOCH_CUDA(cudaMemset2DAsync(y, sizeof(D2) * ldy, 0, sizeof(D2) * nvec, N0, stream));
  BLOCH_CUDA(launch_h1_op(p, 3, tabs, E, x, ldx, y, ldy, nvec, stream, ca, cm));
  count_launch();
  if (ca != 0.0) stats.applies_A += nvec;
}
void bloch_handle_s::apply_h1(int mode, const D2 *x, D2 *y, int nvec) {
  if (p > 3) throw std::invalid_argument("the Nedelec / projector operators support orders 1..3 (order 4: scalar H1 problem only)");
  if (mode != 1) BLOCH_CUDA(cudaMemsetAsync(y, 0, sizeof(D2) * (size_t)N0 * nvec, stream));
  else BLOCH_CUDA(cudaMemsetAsync(y, 0, sizeof(D2) * (size_t)N * nvec, stream));
  BLOCH_CUDA(launch_h1_op(p, mode, tabs, E, x, nvec, y, nvec, nvec, stream));
  count_launch();
}
void bloch_handle_s::apply_curl(const D2 *x, D2 *y, int nvec) {
  BLOCH_CUDA(launch_curl(p, tabs, E, x, nvec, y, nvec, nvec, stream));
  count_launch();
}

static bloch_handle_s *make_handle(const std::vector<std::array<double, 3>> &vert,
                                   const std::vector<std::array<int, 8>> &hex, const double rec[9],
                                   int n_sub, int order, int device) {
  REQUIRE(order >= 1 && order <= 4, "order must be 1..3 (Maxwell) or 1..4 (scalar H1 problem)");
  REQUIRE(n_sub >= 1, "n_sub must be >= 1");
  REQUIRE((int)hex.size() <= kMaxClasses, "too many coarse hexes (element classes)");
  const bool host_only = (device == BLOCH_DEVICE_NONE);
  if (!host_only) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
      throw CudaError("no CUDA device available (this library has no CPU fallback)");
    if (device < 0) BLOCH_CUDA(cudaGetDevice(&device));
    REQUIRE(device < ndev, "device index out of range");
    BLOCH_CUDA(cudaSetDevice(device));
  }
  bloch_handle_s *h = new bloch_handle_s();
  try {
    h->device = device;
    if (!host_only) {
      BLOCH_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
      h->stream = h->own_stream;
    }
    h->p = order;
    h->coarse_vert = vert;
    h->coarse_hex = hex;
    build_mesh(vert, hex, rec, n_sub, h->mesh);
    build_dofmaps(h->mesh, order, h->maps);
    h->N = h->maps.n_nd; h->N0 = h->maps.n_h1; h->Nrt = h->maps.n_rt;
    h->L_nd = h->maps.l_nd; h->L_h1 = h->maps.l_h1; h->L_rt = h->maps.l_rt;
    h->basis = make_basis(order);
    fill_tabs(h->basis, h->tabs);
    h->eps.assign(h->mesh.n_elem, 1.0);
    h->muinv.assign(h->mesh.n_elem, 1.0);
    if (const char *e = std::getenv("BLOCH_TWO_PASS")) h->two_pass = std::atoi(e);
    if (const char *e = std::getenv("BLOCH_MG")) h->use_mg = std::atoi(e);
    if (const char *e = std::getenv("BLOCH_PRECOND")) h->use_aux = std::string(e) != "cheb";
    if (!host_only) build_kernel_maps(h);
  } catch (...) {
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    throw;
  }
  return h;
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

const char *bloch_last_error(void) { return g_last_error.c_str(); }
int bloch_version(void) { return 100; }
int bloch_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// ---- lattice ----
int bloch_lattice_create(bloch_lattice *out, int type, double a, double b, double c, double alpha,
                         double beta, double gamma) {
  API_BEGIN
  REQUIRE(out, "null output");
  bravais::BravaisLattice *L = bravais::BravaisLatticeFactory((bravais::BRAVAIS_LATTICE_TYPE)type, a, b, c, alpha, beta, gamma);
  REQUIRE(L, "lattice type not available (CUB=7, FCC=8, BCC=9, HEX=16)");
  *out = new bloch_lattice_s{L};
  return BLOCH_OK;
  API_END
}
int bloch_lattice_destroy(bloch_lattice lat) {
  if (lat) { delete lat->lat; delete lat; }
  return BLOCH_OK;
}
static int copy_label(const std::string &s, char *buf, int buflen) {
  if (buf && buflen > 0) {
    std::strncpy(buf, s.c_str(), buflen - 1);
    buf[buflen - 1] = 0;
  }
  return (int)s.size();
}
int bloch_lattice_label(bloch_lattice lat, char *buf, int buflen) {
  API_BEGIN
  REQUIRE(lat, "null lattice");
  copy_label(lat->lat->GetLatticeTypeLabel(), buf, buflen);
  return BLOCH_OK;
  API_END
}
double bloch_lattice_volume(bloch_lattice lat) { return lat ? lat->lat->GetUnitCellVolume() : 0.0; }
int bloch_lattice_vectors(bloch_lattice lat, double lat9[9], double rec9[9]) {
  API_BEGIN
  REQUIRE(lat, "null lattice");
  std::vector<bravais::Vec3> a, b;
  lat->lat->GetLatticeVectors(a);
  lat->lat->GetReciprocalLatticeVectors(b);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      if (lat9) lat9[3 * i + j] = a[i][j];
      if (rec9) rec9[3 * i + j] = b[i][j];
    }
  return BLOCH_OK;
  API_END
}
int bloch_lattice_num_translations(bloch_lattice lat) {
  if (!lat) return BLOCH_ERR_ARG;
  std::vector<bravais::Vec3> t;
  lat->lat->GetTranslationVectors(t);
  return (int)t.size();
}
int bloch_lattice_translations(bloch_lattice lat, double *trn, double *radii) {
  API_BEGIN
  REQUIRE(lat, "null lattice");
  std::vector<bravais::Vec3> t;
  std::vector<double> r;
  lat->lat->GetTranslationVectors(t);
  lat->lat->GetFaceRadii(r);
  for (size_t i = 0; i < t.size(); i++) {
    if (trn) for (int j = 0; j < 3; j++) trn[3 * i + j] = t[i][j];
    if (radii) radii[i] = r[i];
  }
  return BLOCH_OK;
  API_END
}
int bloch_lattice_num_symmetry_points(bloch_lattice lat) { return lat ? (int)lat->lat->GetNumberSymmetryPoints() : BLOCH_ERR_ARG; }
int bloch_lattice_symmetry_point(bloch_lattice lat, int i, double kappa[3], char *label, int buflen) {
  API_BEGIN
  REQUIRE(lat && i >= 0 && i < (int)lat->lat->GetNumberSymmetryPoints(), "bad symmetry point index");
  bravais::Vec3 k;
  lat->lat->GetSymmetryPoint(i, k);
  if (kappa) for (int j = 0; j < 3; j++) kappa[j] = k[j];
  copy_label(lat->lat->GetSymmetryPointLabel(i), label, buflen);
  return BLOCH_OK;
  API_END
}
int bloch_lattice_symmetry_point_index(bloch_lattice lat, const char *label) {
  if (!lat || !label) return -1;
  return lat->lat->GetSymmetryPointIndex(label);
}
int bloch_lattice_num_paths(bloch_lattice lat) { return lat ? (int)lat->lat->GetNumberPaths() : BLOCH_ERR_ARG; }
int bloch_lattice_num_path_segments(bloch_lattice lat, int p) {
  if (!lat || p < 0 || p >= (int)lat->lat->GetNumberPaths()) return BLOCH_ERR_ARG;
  return (int)lat->lat->GetNumberPathSegments(p);
}
int bloch_lattice_path_segment(bloch_lattice lat, int p, int s, int *e0, int *e1) {
  API_BEGIN
  REQUIRE(lat && p >= 0 && p < (int)lat->lat->GetNumberPaths(), "bad path index");
  REQUIRE(s >= 0 && s < (int)lat->lat->GetNumberPathSegments(p), "bad segment index");
  int a, b;
  lat->lat->GetPathSegmentEndPointIndices(p, s, a, b);
  if (e0) *e0 = a;
  if (e1) *e1 = b;
  return BLOCH_OK;
  API_END
}
int bloch_lattice_intermediate_point(bloch_lattice lat, int p, int s, double kappa[3], char *label, int buflen) {
  API_BEGIN
  REQUIRE(lat && p >= 0 && p < (int)lat->lat->GetNumberPaths(), "bad path index");
  REQUIRE(s >= 0 && s < (int)lat->lat->GetNumberPathSegments(p), "bad segment index");
  bravais::Vec3 k;
  lat->lat->GetIntermediatePoint(p, s, k);
  if (kappa) for (int j = 0; j < 3; j++) kappa[j] = k[j];
  copy_label(lat->lat->GetIntermediatePointLabel(p, s), label, buflen);
  return BLOCH_OK;
  API_END
}
int bloch_lattice_map_to_primitive_cell(bloch_lattice lat, const double pt[3], double ipt[3]) {
  if (!lat || !pt || !ipt) return BLOCH_ERR_ARG;
  bravais::Vec3 a{pt[0], pt[1], pt[2]}, b;
  bool m = lat->lat->MapToPrimitiveCell(a, b);
  for (int j = 0; j < 3; j++) ipt[j] = b[j];
  return m ? 1 : 0;
}

// ---- eigenproblem object ----
int bloch_create(bloch_handle *out, bloch_lattice lat, int n_sub, int order, int device) {
  API_BEGIN
  REQUIRE(out && lat, "null argument");
  std::vector<bravais::Vec3> b;
  lat->lat->GetReciprocalLatticeVectors(b);
  double rec[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) rec[3 * i + j] = b[i][j];
  *out = make_handle(lat->lat->WignerSeitzVertices(), lat->lat->WignerSeitzHexes(), rec, n_sub, order, device);
  return BLOCH_OK;
  API_END
}
int bloch_create_from_hexes(bloch_handle *out, int n_vert, const double *xyz, int n_hex, const int *hex,
                            const double rec9[9], int n_sub, int order, int device) {
  API_BEGIN
  REQUIRE(out && xyz && hex && rec9 && n_vert > 0 && n_hex > 0, "bad mesh description");
  std::vector<std::array<double, 3>> v(n_vert);
  std::vector<std::array<int, 8>> hx(n_hex);
  for (int i = 0; i < n_vert; i++) v[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
  for (int i = 0; i < n_hex; i++)
    for (int k = 0; k < 8; k++) {
      REQUIRE(hex[8 * i + k] >= 0 && hex[8 * i + k] < n_vert, "hex vertex index out of range");
      hx[i][k] = hex[8 * i + k];
    }
  *out = make_handle(v, hx, rec9, n_sub, order, device);
  return BLOCH_OK;
  API_END
}
int bloch_destroy(bloch_handle h) {
  if (!h) return BLOCH_OK;
  if (h->device < 0) { delete h; return BLOCH_OK; }
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  cudaStream_t own = h->own_stream;
  for (cudaEvent_t e : h->prof_pool) cudaEventDestroy(e);
  if (h->mg) mg_destroy(h->mg);
  if (h->aux) aux_destroy(h->aux);
  if (h->mg_scalar) mg_destroy(h->mg_scalar);
  delete h;
  if (own) cudaStreamDestroy(own);
  return BLOCH_OK;
}
int bloch_set_stream(bloch_handle h, void *s) {
  if (!h || h->device < 0) return BLOCH_ERR_ARG;
  h->stream = s ? (cudaStream_t)s : h->own_stream;
  return BLOCH_OK;
}
int bloch_num_elements(bloch_handle h, int64_t *n_elem, int *n_class) {
  if (!h) return BLOCH_ERR_ARG;
  if (n_elem) *n_elem = h->mesh.n_elem;
  if (n_class) *n_class = h->mesh.n_class;
  return BLOCH_OK;
}
int bloch_num_dofs(bloch_handle h, int64_t *n_nd, int64_t *n_rt, int64_t *n_h1) {
  if (!h) return BLOCH_ERR_ARG;
  if (n_nd) *n_nd = h->N;
  if (n_rt) *n_rt = h->Nrt;
  if (n_h1) *n_h1 = h->N0;
  return BLOCH_OK;
}
int bloch_mesh_counts(bloch_handle h, int64_t *nv, int64_t *ne, int64_t *nf, double *vol) {
  if (!h) return BLOCH_ERR_ARG;
  if (nv) *nv = h->mesh.n_vert;
  if (ne) *ne = h->mesh.n_edge;
  if (nf) *nf = h->mesh.n_face;
  if (vol) *vol = h->mesh.volume;
  return BLOCH_OK;
}
int bloch_element_centers(bloch_handle h, double *xyz) {
  if (!h || !xyz) return BLOCH_ERR_ARG;
  std::vector<double> c;
  h->mesh.centers(c);
  std::memcpy(xyz, c.data(), sizeof(double) * c.size());
  return BLOCH_OK;
}
int bloch_element_geometry(bloch_handle h, double *x0, int *cls, double *J) {
  if (!h) return BLOCH_ERR_ARG;
  if (x0) std::memcpy(x0, h->mesh.x0.data(), sizeof(double) * h->mesh.x0.size());
  if (cls) std::memcpy(cls, h->mesh.cls.data(), sizeof(int) * h->mesh.cls.size());
  if (J) std::memcpy(J, h->mesh.J.data(), sizeof(double) * h->mesh.J.size());
  return BLOCH_OK;
}
int bloch_local_size(bloch_handle h, int space) {
  if (!h) return BLOCH_ERR_ARG;
  return space == 0 ? h->L_h1 : (space == 1 ? h->L_nd : (space == 2 ? h->L_rt : BLOCH_ERR_ARG));
}
int bloch_get_dofmap(bloch_handle h, int space, int32_t *out) {
  if (!h || !out || space < 0 || space > 2) return BLOCH_ERR_ARG;
  const std::vector<int32_t> &m = space == 0 ? h->maps.h1 : (space == 1 ? h->maps.nd : h->maps.rt);
  std::memcpy(out, m.data(), sizeof(int32_t) * m.size());
  return BLOCH_OK;
}
int bloch_set_eps(bloch_handle h, const double *eps) {
  API_BEGIN
  REQUIRE(h && eps, "null argument");
  for (int e = 0; e < h->mesh.n_elem; e++) REQUIRE(eps[e] > 0.0, "eps must be positive");
  h->eps.assign(eps, eps + h->mesh.n_elem);
  h->dirty_coef = true;
  return BLOCH_OK;
  API_END
}
int bloch_set_muinv(bloch_handle h, const double *mu) {
  API_BEGIN
  REQUIRE(h && mu, "null argument");
  for (int e = 0; e < h->mesh.n_elem; e++) REQUIRE(mu[e] > 0.0, "1/mu must be positive");
  h->muinv.assign(mu, mu + h->mesh.n_elem);
  h->dirty_coef = true;
  return BLOCH_OK;
  API_END
}
int bloch_set_kappa(bloch_handle h, const double kappa[3]) {
  if (!h || !kappa) return BLOCH_ERR_ARG;
  h->set_kappas(1, kappa);
  return BLOCH_OK;
}
int bloch_set_kappa_batch(bloch_handle h, int nk, const double *kappa) {
  API_BEGIN
  REQUIRE(h && kappa, "null argument");
  REQUIRE(nk >= 1 && nk <= BLOCH_MAX_BATCH, "batch size must be in [1, BLOCH_MAX_BATCH]");
  h->set_kappas(nk, kappa);
  return BLOCH_OK;
  API_END
}
int bloch_batch_size(bloch_handle h) { return h ? h->nk : BLOCH_ERR_ARG; }
int bloch_select_kpoint(bloch_handle h, int k) {
  API_BEGIN
  REQUIRE(h && k >= 0 && k < h->nk, "k-point index outside the batch");
  h->sel = k;
  return BLOCH_OK;
  API_END
}
int bloch_set_profile(bloch_handle h, int on) {
  if (!h) return BLOCH_ERR_ARG;
  h->profile = on != 0;
  return BLOCH_OK;
}
int bloch_get_profile(bloch_handle h, double *ms, int n) {
  if (!h || !ms || n < 1) return BLOCH_ERR_ARG;
  for (int i = 0; i < n; i++) ms[i] = i < 8 ? h->stats.prof_ms[i] : 0.0;
  return BLOCH_OK;
}
int bloch_set_num_bands(bloch_handle h, int n) {
  API_BEGIN
  REQUIRE(h && n >= 1 && n <= 20, "number of bands must be in [1,20]");
  if (n != h->nbands) h->have_vectors = 0;
  h->nbands = n;
  return BLOCH_OK;
  API_END
}
int bloch_set_tol(bloch_handle h, double tol, int max_iter) {
  API_BEGIN
  REQUIRE(h && tol > 0.0 && max_iter > 0, "bad tolerance / iteration limit");
  h->tol = tol;
  h->max_iter = max_iter;
  return BLOCH_OK;
  API_END
}
int bloch_setup(bloch_handle h) {
  API_BEGIN
  REQUIRE(h, "null handle");
  h->setup();
  return BLOCH_OK;
  API_END
}
int bloch_set_initial_vectors(bloch_handle h, int m, const double *vecs) {
  API_BEGIN
  REQUIRE(h, "null handle");
  if (m <= 0 || !vecs) { h->n_init = 0; h->init_vecs.clear(); return BLOCH_OK; }
  h->n_init = m;
  h->init_vecs.assign(vecs, vecs + (size_t)m * 2 * h->N);
  return BLOCH_OK;
  API_END
}
int bloch_solve(bloch_handle h) {
  API_BEGIN
  REQUIRE(h, "null handle");
  h->setup();
  h->solve();
  h->sel = 0;
  return h->stats.converged >= h->nbands ? BLOCH_OK : BLOCH_ERR_NOCONV;
  API_END
}
int bloch_get_eigenvalues(bloch_handle h, double *lambda, int n) {
  API_BEGIN
  REQUIRE(h && lambda, "null argument");
  REQUIRE((int)h->eigenvalues.size() == h->nk * h->nbands, "no eigenvalues (call bloch_solve first)");
  REQUIRE(n >= 0 && n <= h->nbands, "more eigenvalues requested than computed");
  std::memcpy(lambda, h->eigenvalues.data() + (size_t)h->sel * h->nbands, sizeof(double) * n);
  return BLOCH_OK;
  API_END
}
int bloch_get_stats(bloch_handle h, bloch_stats *st) {
  if (!h || !st) return BLOCH_ERR_ARG;
  const bool perk = (int)h->stats.k_iterations.size() == h->nk && h->nk > 1;
  st->iterations = perk ? h->stats.k_iterations[h->sel] : h->stats.iterations;
  st->converged_bands = perk ? h->stats.k_converged[h->sel] : h->stats.converged;
  st->inner_iterations = h->stats.inner_iterations;
  st->solve_seconds = h->stats.seconds;
  st->max_residual = perk ? h->stats.k_max_residual[h->sel] : h->stats.max_residual;
  st->applies_A = h->stats.applies_A;
  st->kernel_launches = h->stats.launches;
  return BLOCH_OK;
}

// host-pointer operator applications: H2D, pack, kernel, unpack, D2H
enum { OP_A, OP_M, OP_PROJ, OP_C };
static int apply_host(bloch_handle h, int op, const double *x, double *y, int nvec) {
  API_BEGIN
  REQUIRE(h && x && y && nvec >= 1, "bad argument");
  h->setup();
  const long N = h->N, Nout = (op == OP_C) ? h->Nrt : h->N;
  cudaStream_t s = h->stream;
  h->d_io_a.alloc((size_t)2 * N * nvec);
  h->d_io_b.alloc((size_t)2 * Nout * nvec);
  h->d_blk_a.alloc((size_t)N * nvec);
  h->d_blk_b.alloc((size_t)Nout * nvec);
  BLOCH_CUDA(cudaMemcpyAsync(h->d_io_a.p, x, sizeof(double) * 2 * N * nvec, cudaMemcpyHostToDevice, s));
  BLOCH_CUDA(launch_pack(h->d_io_a.p, h->d_blk_a.p, N, nvec, s));
  h->count_launch();
  switch (op) {
    case OP_A: h->apply_nd(h->d_blk_a.p, h->d_blk_b.p, nvec, 1.0, 0.0); break;
    case OP_M: h->apply_nd(h->d_blk_a.p, h->d_blk_b.p, nvec, 0.0, 1.0); break;
    case OP_C: h->apply_curl(h->d_blk_a.p, h->d_blk_b.p, nvec); break;
    case OP_PROJ:
      BLOCH_CUDA(cudaMemcpyAsync(h->d_blk_b.p, h->d_blk_a.p, sizeof(D2) * N * nvec, cudaMemcpyDeviceToDevice, s));
      h->project(h->d_blk_b.p, nvec, 1e-13, nullptr);
      break;
  }
  BLOCH_CUDA(launch_unpack(h->d_blk_b.p, h->d_io_b.p, Nout, nvec, s));
  h->count_launch();
  BLOCH_CUDA(cudaMemcpyAsync(y, h->d_io_b.p, sizeof(double) * 2 * Nout * nvec, cudaMemcpyDeviceToHost, s));
  BLOCH_CUDA(cudaStreamSynchronize(s));
  return BLOCH_OK;
  API_END
}
int bloch_apply_A(bloch_handle h, const double *x, double *y, int nvec) { return apply_host(h, OP_A, x, y, nvec); }
int bloch_apply_M(bloch_handle h, const double *x, double *y, int nvec) { return apply_host(h, OP_M, x, y, nvec); }
int bloch_apply_projector(bloch_handle h, const double *x, double *y, int nvec) { return apply_host(h, OP_PROJ, x, y, nvec); }
int bloch_apply_C(bloch_handle h, const double *x, double *y, int nvec) { return apply_host(h, OP_C, x, y, nvec); }

int bloch_apply_A_device(bloch_handle h, const double *d_x, double *d_y, int nvec) {
  API_BEGIN
  REQUIRE(h && d_x && d_y && nvec >= 1, "bad argument");
  if (h->device < 0 || h->dirty_coef || h->dirty_kappa) h->setup();
  BLOCH_CUDA(cudaSetDevice(h->device));
  h->apply_nd((const D2 *)d_x, (D2 *)d_y, nvec, 1.0, 0.0);
  return BLOCH_OK;
  API_END
}
int bloch_apply_M_device(bloch_handle h, const double *d_x, double *d_y, int nvec) {
  API_BEGIN
  REQUIRE(h && d_x && d_y && nvec >= 1, "bad argument");
  if (h->device < 0 || h->dirty_coef || h->dirty_kappa) h->setup();
  BLOCH_CUDA(cudaSetDevice(h->device));
  h->apply_nd((const D2 *)d_x, (D2 *)d_y, nvec, 0.0, 1.0);
  return BLOCH_OK;
  API_END
}
int bloch_pack_device(bloch_handle h, const double *d_reim, double *d_block, int nvec) {
  API_BEGIN
  REQUIRE(h && d_reim && d_block && nvec >= 1 && h->device >= 0, "bad argument");
  BLOCH_CUDA(cudaSetDevice(h->device));
  BLOCH_CUDA(launch_pack(d_reim, (D2 *)d_block, h->N, nvec, h->stream));
  h->count_launch();
  return BLOCH_OK;
  API_END
}
int bloch_unpack_device(bloch_handle h, const double *d_block, double *d_reim, int nvec) {
  API_BEGIN
  REQUIRE(h && d_reim && d_block && nvec >= 1 && h->device >= 0, "bad argument");
  BLOCH_CUDA(cudaSetDevice(h->device));
  BLOCH_CUDA(launch_unpack((const D2 *)d_block, d_reim, h->N, nvec, h->stream));
  h->count_launch();
  return BLOCH_OK;
  API_END
}

// test hook: H1 <-> ND operators of the projector on host vectors.
// mode 0: y(2 N0) = S0 x(2 N0); mode 1: y(2 N) = G x(2 N0); mode 2: y(2 N0) = G^H M x(2 N)
int bloch_debug_apply_h1op(bloch_handle h, int mode, const double *x, double *y, int nvec) {
  API_BEGIN
  REQUIRE(h && x && y && nvec >= 1 && mode >= 0 && mode <= 2, "bad argument");
  h->setup();
  const long nin = (mode == 2) ? h->N : h->N0, nout = (mode == 1) ? h->N : h->N0;
  cudaStream_t s = h->stream;
  DevBuf<double> ia, ib;
  DevBuf<D2> ba, bb;
  ia.alloc((size_t)2 * nin * nvec); ib.alloc((size_t)2 * nout * nvec);
  ba.alloc((size_t)nin * nvec); bb.alloc((size_t)nout * nvec);
  BLOCH_CUDA(cudaMemcpyAsync(ia.p, x, sizeof(double) * 2 * nin * nvec, cudaMemcpyHostToDevice, s));
  BLOCH_CUDA(launch_pack(ia.p, ba.p, nin, nvec, s));
  h->apply_h1(mode, ba.p, bb.p, nvec);
  BLOCH_CUDA(launch_unpack(bb.p, ib.p, nout, nvec, s));
  BLOCH_CUDA(cudaMemcpyAsync(y, ib.p, sizeof(double) * 2 * nout * nvec, cudaMemcpyDeviceToHost, s));
  BLOCH_CUDA(cudaStreamSynchronize(s));
  return BLOCH_OK;
  API_END
}

// test hook: pieces of the auxiliary-space preconditioner (aux.cu) on host vectors; (H1)^3 vectors hold the three
// Cartesian components as blocks of N0 entries: [re(3 N0); im(3 N0)].
// mode 0: y(2 N) = Pi u(2 * 3 N0); mode 1: y(2 * 3 N0) = Pi^T x(2 N); mode 2: y(2 * 3 N0) = B u, one V-cycle per component
int bloch_debug_apply_aux(bloch_handle h, int mode, const double *x, double *y, int nvec) {
  API_BEGIN
  REQUIRE(h && x && y && nvec >= 1 && mode >= 0 && mode <= 2, "bad argument");
  h->setup();
  REQUIRE(h->aux, "no auxiliary space on this handle (odd n_sub, order > 3 or BLOCH_PRECOND=cheb)");
  REQUIRE(mode != 2 || nvec % h->nk == 0, "block width must be a multiple of the k-point batch size");
  const long n3 = 3 * h->N0, nin = (mode == 1) ? h->N : n3, nout = (mode == 0) ? h->N : n3;
  cudaStream_t s = h->stream;
  DevBuf<double> ia, ib;
  DevBuf<D2> ba, bb;
  ia.alloc((size_t)2 * nin * nvec); ib.alloc((size_t)2 * nout * nvec);
  ba.alloc((size_t)nin * nvec); bb.alloc((size_t)nout * nvec);
  BLOCH_CUDA(cudaMemcpyAsync(ia.p, x, sizeof(double) * 2 * nin * nvec, cudaMemcpyHostToDevice, s));
  BLOCH_CUDA(launch_pack(ia.p, ba.p, nin, nvec, s));
  if (mode == 0) aux_apply_pi(h->aux, h, ba.p, bb.p, nvec, false);
  else if (mode == 1) aux_apply_pit(h->aux, h, ba.p, bb.p, nvec);
  else aux_vcycles(h->aux, h, ba.p, bb.p, nvec);
  BLOCH_CUDA(launch_unpack(bb.p, ib.p, nout, nvec, s));
  BLOCH_CUDA(cudaMemcpyAsync(y, ib.p, sizeof(double) * 2 * nout * nvec, cudaMemcpyDeviceToHost, s));
  BLOCH_CUDA(cudaStreamSynchronize(s));
  return BLOCH_OK;
  API_END
}

// test hook (host only, also on topology-only handles): the nodal interpolation Pi of the auxiliary-space preconditioner
// in CSR form.  rowptr == NULL: only *nnz is returned.
int bloch_debug_pi_matrix(bloch_handle h, int64_t *nnz, int64_t *rowptr, int32_t *col, double *val) {
  API_BEGIN
  REQUIRE(h && nnz, "null argument");
  REQUIRE(h->p <= 3, "the Nedelec space supports orders 1..3");
  std::vector<int> ptr;
  std::vector<int32_t> c;
  std::vector<double> v;
  aux_build_pi(h, ptr, c, v);
  *nnz = (int64_t)c.size();
  if (!rowptr) return BLOCH_OK;
  REQUIRE(col && val, "null argument");
  for (size_t i = 0; i < ptr.size(); i++) rowptr[i] = ptr[i];
  std::memcpy(col, c.data(), sizeof(int32_t) * c.size());
  std::memcpy(val, v.data(), sizeof(double) * v.size());
  return BLOCH_OK;
  API_END
}

// test hook: nested-mesh transfer of the H1 multigrid between the handle's mesh and the next coarser one in its three
// implementations (variant 0 CSR, 1 sum-factorised, 2 element-wise); dir 0: y(2 N0) = P x(2 N0c), dir 1: y(2 N0c) = P^T x(2 N0)
int bloch_debug_mg_transfer(bloch_handle h, int variant, int dir, const double *x, double *y, int nvec, int64_t *n_coarse) {
  API_BEGIN
  REQUIRE(h && variant >= 0 && variant <= 2 && dir >= 0 && dir <= 1, "bad argument");
  h->setup();
  REQUIRE(h->mg && mg_level_size(h->mg, 1) > 0, "no nested coarser mesh (odd n_sub)");
  const long nc = mg_level_size(h->mg, 1), nf = h->N0;
  if (n_coarse) *n_coarse = nc;
  if (!x || !y) return BLOCH_OK;          // size query
  REQUIRE(nvec >= 1, "bad argument");
  const long nin = dir == 0 ? nc : nf, nout = dir == 0 ? nf : nc;
  cudaStream_t s = h->stream;
  DevBuf<double> ia, ib;
  DevBuf<D2> ba, bb;
  ia.alloc((size_t)2 * nin * nvec); ib.alloc((size_t)2 * nout * nvec);
  ba.alloc((size_t)nin * nvec); bb.alloc((size_t)nout * nvec);
  BLOCH_CUDA(cudaMemcpyAsync(ia.p, x, sizeof(double) * 2 * nin * nvec, cudaMemcpyHostToDevice, s));
  BLOCH_CUDA(launch_pack(ia.p, ba.p, nin, nvec, s));
  mg_debug_transfer(h->mg, h, variant, dir, ba.p, bb.p, nvec);
  BLOCH_CUDA(launch_unpack(bb.p, ib.p, nout, nvec, s));
  BLOCH_CUDA(cudaMemcpyAsync(y, ib.p, sizeof(double) * 2 * nout * nvec, cudaMemcpyDeviceToHost, s));
  BLOCH_CUDA(cudaStreamSynchronize(s));
  return BLOCH_OK;
  API_END
}

// ---- host-side dense Rayleigh-Ritz solver, exposed for the CPU tests (no GPU involved) ----
int bloch_debug_hegv(int n, int m, const double *ga_reim, const double *gm_reim, double *lambda, double *c_reim,
                     int values_only) {
  API_BEGIN
  REQUIRE(n >= 1 && m >= 1 && m <= n && ga_reim && gm_reim && lambda, "bad argument");
  dense::Mat GA((size_t)n * n), GM((size_t)n * n), Cm;
  for (size_t i = 0; i < (size_t)n * n; i++) {
    GA[i] = dense::cplx(ga_reim[2 * i], ga_reim[2 * i + 1]);
    GM[i] = dense::cplx(gm_reim[2 * i], gm_reim[2 * i + 1]);
  }
  std::vector<double> lam;
  if (values_only) {
    if (!dense::hegv_lowest_values(n, m, GA, GM, lam)) return BLOCH_ERR_NOCONV;
  } else {
    REQUIRE(c_reim, "null eigenvector output");
    if (!dense::hegv_lowest(n, m, GA, GM, lam, Cm)) return BLOCH_ERR_NOCONV;
    for (size_t i = 0; i < (size_t)n * m; i++) { c_reim[2 * i] = Cm[i].real(); c_reim[2 * i + 1] = Cm[i].imag(); }
  }
  for (int i = 0; i < m; i++) lambda[i] = lam[i];
  return BLOCH_OK;
  API_END
}

// ---- device-side dense Rayleigh-Ritz solver (rr_device.cu), exposed for the parity tests against dense.hpp ----
int bloch_debug_hegv_device(int n, int m, int nk, const double *ga_reim, const double *gm_reim, const unsigned char *act,
                            int use_p, double *lambda, double *c_reim, int *info) {
  API_BEGIN
  REQUIRE(n >= 1 && n <= 63 && m >= 1 && m <= n && m <= 32 && nk >= 1 && ga_reim && gm_reim && lambda && c_reim && info,
          "bad argument");
  int ndev = 0;
  REQUIRE(cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0, "no CUDA device");
  DevBuf<D2> dGA, dGM, dC;
  DevBuf<double> dlam;
  DevBuf<unsigned char> dact, dusep;
  DevBuf<int> dinfo;
  const size_t nn = (size_t)nk * n * n;
  dGA.alloc(nn); dGM.alloc(nn); dC.alloc((size_t)nk * n * m); dlam.alloc((size_t)nk * m);
  dact.alloc((size_t)nk * m); dusep.alloc(nk); dinfo.alloc(nk);
  std::vector<unsigned char> a((size_t)nk * m, 1), up(nk, use_p ? 1 : 0);
  if (act) a.assign(act, act + (size_t)nk * m);
  BLOCH_CUDA(cudaMemcpy(dGA.p, ga_reim, sizeof(D2) * nn, cudaMemcpyHostToDevice));
  BLOCH_CUDA(cudaMemcpy(dGM.p, gm_reim, sizeof(D2) * nn, cudaMemcpyHostToDevice));
  BLOCH_CUDA(cudaMemcpy(dact.p, a.data(), a.size(), cudaMemcpyHostToDevice));
  BLOCH_CUDA(cudaMemcpy(dusep.p, up.data(), up.size(), cudaMemcpyHostToDevice));
  BLOCH_CUDA(launch_rr_solve(dGA.p, dGM.p, n, m, nk, dact.p, dusep.p, dC.p, dlam.p, dinfo.p, nullptr));
  BLOCH_CUDA(cudaDeviceSynchronize());
  BLOCH_CUDA(cudaMemcpy(lambda, dlam.p, sizeof(double) * nk * m, cudaMemcpyDeviceToHost));
  BLOCH_CUDA(cudaMemcpy(c_reim, dC.p, sizeof(D2) * nk * n * m, cudaMemcpyDeviceToHost));
  BLOCH_CUDA(cudaMemcpy(info, dinfo.p, sizeof(int) * nk, cudaMemcpyDeviceToHost));
  return BLOCH_OK;
  API_END
}

// ---- assembled operators (the -wm dump of maxwell_dispersion.cpp:553-590) ----
int bloch_assemble_matrix(bloch_handle h, int which, int64_t *nnz) {
  API_BEGIN
  REQUIRE(h && nnz, "null argument");
  REQUIRE(h->device >= 0, "topology-only handle");
  REQUIRE(which == 0 || which == 1, "which must be 0 (A) or 1 (M)");
  BLOCH_CUDA(cudaSetDevice(h->device));
  h->assemble(which);
  *nnz = (int64_t)h->csr_col.size();
  return BLOCH_OK;
  API_END
}
int bloch_get_matrix(bloch_handle h, int64_t *rowptr, int32_t *col, double *re, double *im) {
  API_BEGIN
  REQUIRE(h && rowptr && col && re, "null argument");
  REQUIRE((long)h->csr_ptr.size() == h->N + 1, "no assembled matrix (call bloch_assemble_matrix first)");
  std::copy(h->csr_ptr.begin(), h->csr_ptr.end(), rowptr);
  std::copy(h->csr_col.begin(), h->csr_col.end(), col);
  std::copy(h->csr_re.begin(), h->csr_re.end(), re);
  if (im) std::copy(h->csr_im.begin(), h->csr_im.end(), im);
  return BLOCH_OK;
  API_END
}

// ---- multilevel warm start: refinement of eigenvectors (meta_material_solver.cpp:2829-2853) ----
int bloch_prolong_eigenvectors(bloch_handle coarse, bloch_handle fine) {
  API_BEGIN
  REQUIRE(coarse && fine && coarse != fine, "bad handles");
  REQUIRE(coarse->device >= 0 && coarse->device == fine->device, "handles must live on the same device");
  REQUIRE(coarse->p == fine->p && coarse->p <= 3, "orders differ / not supported");
  REQUIRE(fine->mesh.n_sub == 2 * coarse->mesh.n_sub && fine->mesh.n_elem == 8 * coarse->mesh.n_elem &&
              coarse->coarse_hex == fine->coarse_hex && coarse->coarse_vert == fine->coarse_vert,
          "fine mesh must be the uniform refinement of the coarse one");
  REQUIRE(coarse->have_vectors > 0, "no coarse eigenvectors (call bloch_solve on the coarse handle first)");
  REQUIRE(coarse->nk == 1 && fine->nk == 1, "multilevel warm start works on single-kappa handles");
  BLOCH_CUDA(cudaSetDevice(fine->device));
  const int p = fine->p, q = p + 1, mb = coarse->block;
  NdTransfer1D T;
  std::memset(&T, 0, sizeof(T));
  std::vector<double> v, dv;
  for (int a = 0; a < 2; a++) {
    for (int i = 0; i < q; i++) {
      detail::lagrange(fine->basis.l, 0.5 * (a + fine->basis.l[i]), v, dv);
      for (int j = 0; j < q; j++) T.Pc[a][i][j] = std::fabs(v[j]) < 1e-15 ? 0.0 : v[j];
    }
    for (int o = 0; o < p; o++) {
      detail::lagrange(fine->basis.g, 0.5 * (a + fine->basis.g[o]), v, dv);
      for (int k = 0; k < p; k++) T.Po[a][o][k] = 0.5 * v[k];
    }
  }
  // coarse stream finished -> prolong on the fine stream -> the fine solve warm-starts from d_X
  BLOCH_CUDA(cudaStreamSynchronize(coarse->stream));
  fine->d_X.alloc((size_t)fine->N * mb);
  BLOCH_CUDA(launch_nd_prolong(p, T, fine->d_map_nd.p, coarse->d_map_nd.p, fine->mesh.n_elem, fine->mesh.n_sub,
                               coarse->d_X.p, mb, fine->d_X.p, mb, mb, fine->stream));
  fine->count_launch();
  BLOCH_CUDA(cudaStreamSynchronize(fine->stream));
  fine->block = mb;
  fine->have_vectors = coarse->have_vectors;
  fine->eigenvalues = coarse->eigenvalues;
  fine->n_init = 0;
  return BLOCH_OK;
  API_END
}

// ---- field averages: GetFieldAverages (maxwell/maxwell_bloch.cpp:1550-1632) ----
int bloch_get_field_averages(bloch_handle h, int i, double out24[24]) {
  API_BEGIN
  REQUIRE(h && out24, "null argument");
  REQUIRE(h->device >= 0, "topology-only handle");
  REQUIRE(i >= 0 && i < h->have_vectors && i < h->nbands && (int)h->eigenvalues.size() == h->nk * h->nbands,
          "eigenvector index out of range (only the bands of the last bloch_solve are addressable)");
  REQUIRE(h->p <= 3, "order not supported");
  BLOCH_CUDA(cudaSetDevice(h->device));
  h->field_averages(i, out24);
  return BLOCH_OK;
  API_END
}

// ---- reduced-basis sweep: MaxwellDispersion (meta-material/meta_material_solver.cpp:3132-3410) ----
int bloch_rb_clear(bloch_handle h) {
  if (!h) return BLOCH_ERR_ARG;
  h->rb_size = 0;
  return BLOCH_OK;
}
int bloch_rb_append(bloch_handle h) {
  API_BEGIN
  REQUIRE(h && h->device >= 0, "bad handle");
  BLOCH_CUDA(cudaSetDevice(h->device));
  h->rb_append();
  return BLOCH_OK;
  API_END
}
int bloch_rb_size(bloch_handle h) { return h ? h->rb_size : BLOCH_ERR_ARG; }
int bloch_rb_approx(bloch_handle h, const double kappa[3], double *lambda, int n) {
  API_BEGIN
  REQUIRE(h && kappa && lambda && n >= 1, "bad argument");
  h->set_kappas(1, kappa);
  auto t0 = std::chrono::steady_clock::now();
  h->setup();
  if (std::getenv("BLOCH_VERBOSE"))
    std::printf("[rb] setup %.2f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  h->rb_approx(lambda, n);
  return BLOCH_OK;
  API_END
}

// ---- scalar H1 Bloch Helmholtz variant: ScalarFloquetWaveEquation (misc/scalar3d.cpp:662-818) ----
int bloch_scalar_set_coefs(bloch_handle h, const double *stiffness_k, const double *mass_m) {
  API_BEGIN
  REQUIRE(h && stiffness_k && mass_m, "null argument");
  for (int e = 0; e < h->mesh.n_elem; e++) REQUIRE(stiffness_k[e] > 0.0 && mass_m[e] > 0.0, "coefficients must be positive");
  h->eps.assign(stiffness_k, stiffness_k + h->mesh.n_elem);     // M1(k) weight of the gradient form
  h->muinv.assign(mass_m, mass_m + h->mesh.n_elem);             // M0(m)
  h->scalar_ready = true;
  h->dirty_coef = true;
  return BLOCH_OK;
  API_END
}
int bloch_scalar_set_num_modes(bloch_handle h, int n) {
  API_BEGIN
  REQUIRE(h && n >= 1 && n <= 20, "number of modes must be in [1,20]");
  if (n != h->nbands_s) h->have_vectors_s = 0;
  h->nbands_s = n;
  return BLOCH_OK;
  API_END
}
int bloch_scalar_solve(bloch_handle h) {
  API_BEGIN
  REQUIRE(h, "null handle");
  h->setup();
  h->solve_scalar();
  return h->stats.converged >= h->nbands_s ? BLOCH_OK : BLOCH_ERR_NOCONV;
  API_END
}
int bloch_scalar_get_eigenvalues(bloch_handle h, double *lambda, int n) {
  API_BEGIN
  REQUIRE(h && lambda && n >= 0 && n <= h->nbands_s && (int)h->eigenvalues_s.size() == h->nk * h->nbands_s,
          "more eigenvalues requested than computed");
  std::memcpy(lambda, h->eigenvalues_s.data() + (size_t)h->sel * h->nbands_s, sizeof(double) * n);
  return BLOCH_OK;
  API_END
}
int bloch_scalar_apply(bloch_handle h, int which, const double *x, double *y, int nvec) {
  API_BEGIN
  REQUIRE(h && x && y && nvec >= 1 && (which == 0 || which == 1), "bad argument");
  h->setup();
  const long n = h->N0;
  cudaStream_t s = h->stream;
  DevBuf<double> ia, ib;
  DevBuf<D2> ba, bb;
  ia.alloc((size_t)2 * n * nvec); ib.alloc((size_t)2 * n * nvec); ba.alloc((size_t)n * nvec); bb.alloc((size_t)n * nvec);
  BLOCH_CUDA(cudaMemcpyAsync(ia.p, x, sizeof(double) * 2 * n * nvec, cudaMemcpyHostToDevice, s));
  BLOCH_CUDA(launch_pack(ia.p, ba.p, n, nvec, s));
  h->apply_scalar_ld(ba.p, nvec, bb.p, nvec, nvec, which == 0 ? 1.0 : 0.0, which == 0 ? 0.0 : 1.0);
  BLOCH_CUDA(launch_unpack(bb.p, ib.p, n, nvec, s));
  BLOCH_CUDA(cudaMemcpyAsync(y, ib.p, sizeof(double) * 2 * n * nvec, cudaMemcpyDeviceToHost, s));
  BLOCH_CUDA(cudaStreamSynchronize(s));
  return BLOCH_OK;
  API_END
}

int bloch_debug_fp64_peak(bloch_handle h, double *tflops) {
  API_BEGIN
  REQUIRE(h && tflops && h->device >= 0, "bad argument");
  BLOCH_CUDA(cudaSetDevice(h->device));
  BLOCH_CUDA(measure_fp64_peak(tflops, h->stream));
  return BLOCH_OK;
  API_END
}

int bloch_get_eigenvector_E(bloch_handle h, int i, double *re, double *im) {
  API_BEGIN
  REQUIRE(h && re && im, "null argument");
  REQUIRE(i >= 0 && i < h->have_vectors && i < h->nbands && (int)h->eigenvalues.size() == h->nk * h->nbands,
          "eigenvector index out of range (only the bands of the last bloch_solve are addressable)");
  BLOCH_CUDA(cudaSetDevice(h->device));
  std::vector<D2> col(h->N);
  BLOCH_CUDA(cudaMemcpy2DAsync(col.data(), sizeof(D2), h->d_X.p + (size_t)h->sel * h->block + i,
                               sizeof(D2) * h->nk * h->block, sizeof(D2), h->N, cudaMemcpyDeviceToHost, h->stream));
  BLOCH_CUDA(cudaStreamSynchronize(h->stream));
  for (long k = 0; k < h->N; k++) { re[k] = col[k].x; im[k] = col[k].y; }
  return BLOCH_OK;
  API_END
}
int bloch_get_eigenvector_B(bloch_handle h, int i, double *re, double *im) {
  API_BEGIN
  REQUIRE(h && re && im, "null argument");
  REQUIRE(i >= 0 && i < h->have_vectors && i < h->nbands && (int)h->eigenvalues.size() == h->nk * h->nbands,
          "eigenvector index out of range (only the bands of the last bloch_solve are addressable)");
  BLOCH_CUDA(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  DevBuf<D2> e, b;
  e.alloc(h->N); b.alloc(h->Nrt);
  BLOCH_CUDA(cudaMemcpy2DAsync(e.p, sizeof(D2), h->d_X.p + (size_t)h->sel * h->block + i, sizeof(D2) * h->nk * h->block,
                               sizeof(D2), h->N, cudaMemcpyDeviceToDevice, s));
  BLOCH_CUDA(launch_curl(h->p, h->tabs, h->elem_of_k(h->sel), e.p, 1, b.p, 1, 1, s));
  h->count_launch();
  std::vector<D2> col(h->Nrt);
  BLOCH_CUDA(cudaMemcpyAsync(col.data(), b.p, sizeof(D2) * h->Nrt, cudaMemcpyDeviceToHost, s));
  BLOCH_CUDA(cudaStreamSynchronize(s));
  // B = C E / sqrt|lambda| (maxwell_bloch.cpp:1432-1457).  In the reference's real block form
  // C E = [Cr; Ci] and it returns Bi = block 0, Br = -block 1.
  const double lam = std::fabs(h->eigenvalues[(size_t)h->sel * h->nbands + i]);
  const double sc = lam > 0 ? 1.0 / std::sqrt(lam) : 1.0;
  for (long k = 0; k < h->Nrt; k++) { im[k] = sc * col[k].x; re[k] = -sc * col[k].y; }
  return BLOCH_OK;
  API_END
}

}  // extern "C"
