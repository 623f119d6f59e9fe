// Internal definition of the eigenproblem object behind the C ABI (include/bloch_b200.h).
// Mirrors the state of mfem::bloch::MaxwellBlochWaveEquation (maxwell/maxwell_bloch.hpp:236-390):
// spaces, coefficients, kappa = beta*zeta, dirty flags, solver controls and results.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <functional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "basis.hpp"
#include "bravais.hpp"
#include "kernels.hpp"
#include "mesh.hpp"
#include "mg.hpp"
#include "aux.hpp"

struct bloch_lattice_s {
  bloch_b200::bravais::BravaisLattice *lat = nullptr;
};

namespace bloch_b200 {

struct CudaError : std::runtime_error {
  explicit CudaError(const std::string &m) : std::runtime_error(m) {}
};
#define BLOCH_CUDA(call)                                                                     \
  do {                                                                                       \
    cudaError_t err__ = (call);                                                              \
    if (err__ != cudaSuccess)                                                                \
      throw ::bloch_b200::CudaError(std::string(#call) + ": " + cudaGetErrorString(err__)); \
  } while (0)

// Host wait for the handle's stream.  BLOCH_BLOCKING_SYNC=1: wait on a blocking event so the thread
// sleeps instead of spinning - for runs with more solver threads than host cores (several
// concurrent k-point solves per GPU x several ranks per box).
inline void h_sync(cudaStream_t s) {
  static const bool blocking = [] { const char *e = std::getenv("BLOCH_BLOCKING_SYNC"); return e && std::atoi(e) != 0; }();
  if (!blocking) {
    BLOCH_CUDA(cudaStreamSynchronize(s));
    return;
  }
  static thread_local cudaEvent_t evs[kMaxDevices] = {};   // an event belongs to the device it was created on
  cudaEvent_t &ev = evs[current_device_slot()];
  if (!ev) BLOCH_CUDA(cudaEventCreateWithFlags(&ev, cudaEventBlockingSync | cudaEventDisableTiming));
  BLOCH_CUDA(cudaEventRecord(ev, s));
  BLOCH_CUDA(cudaEventSynchronize(ev));
}

template <class T>
struct DevBuf {   // owning device buffer
  T *p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  void alloc(size_t count) {
    if (count <= n && p) return;
    release();
    BLOCH_CUDA(cudaMalloc(&p, sizeof(T) * (count ? count : 1)));
    n = count;
  }
  void upload(const std::vector<T> &h, cudaStream_t s) {
    alloc(h.size());
    BLOCH_CUDA(cudaMemcpyAsync(p, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, s));
  }
};

// one Hermitian pencil for the block eigensolver: y = ca A x + cm M x, Jacobi diagonals, optional
// divergence constraint (ND problem) - lets the ND Maxwell and the scalar H1 problems share lobpcg()
struct EigProblem {
  long n = 0;
  std::function<void(const double2 *, int, double2 *, int, int, double, double)> apply;
  const double *diagA = nullptr, *diagM = nullptr;
  double lmax_local = 0;
  bool constrained = false, use_init = false;
  H1Multigrid *mg_precond = nullptr;   // scalar H1 problem: T = one V-cycle of this hierarchy (kind 2)
  DevBuf<double2> *X = nullptr;
  std::vector<double> *evals = nullptr;
  int *have = nullptr, *blk = nullptr;
  int nbands = 0;
};

// shared host helpers (core.cu)
void class_params(const double *J, const double kappa[3], double out[kClassParDoubles]);
double local_scaled_lmax(int L, const double2 *X);

struct SolverStats {
  int iterations = 0, converged = 0, inner_iterations = 0;
  double seconds = 0, max_residual = 0;
  int64_t applies_A = 0, launches = 0;
  // per k-point of a batched solve (size nk): outer iterations until that k-point converged, converged bands,
  // largest residual of its wanted bands
  std::vector<int> k_iterations, k_converged;
  std::vector<double> k_max_residual;
  // phase times of the last solve in ms (only with profiling on, bloch_set_profile): whole solve, operator-apply
  // kernels (ND A/M applies incl. the clearing of y), projector (inner S0 solves + G, G^H M applies),
  // Chebyshev vector updates, Gram + rotation kernels, host Rayleigh-Ritz (incl. the copies around it)
  double prof_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

}  // namespace bloch_b200

struct bloch_handle_s {
  using D2 = double2;
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  int p = 1;
  bloch_b200::HexMesh mesh;
  std::vector<std::array<double, 3>> coarse_vert;   // coarse WS cell, kept for the multigrid levels
  std::vector<std::array<int, 8>> coarse_hex;
  bloch_b200::H1Multigrid *mg = nullptr;            // h-multigrid for the projector's S0 solves
  int use_mg = 1;
  bloch_b200::AuxSpace *aux = nullptr;              // auxiliary nodal space of the ND preconditioner (aux.cu)
  int use_aux = 1;                                  // 0: Chebyshev-Jacobi polynomial only (BLOCH_PRECOND=cheb)
  bloch_b200::H1Multigrid *mg_scalar = nullptr;     // kind-2 hierarchy: preconditioner of the scalar H1 problem
  bloch_b200::DofMaps maps;
  bloch_b200::Basis1D basis;
  bloch_b200::Tabs tabs;
  long N = 0, N0 = 0, Nrt = 0;
  int L_nd = 0, L_h1 = 0, L_rt = 0;

  std::vector<double> eps, muinv;
  // k-point batch: nk Bloch vectors solved together (independent eigenproblems sharing mesh, maps and
  // coefficients, maxwell_dispersion.cpp:475-531); kappa / beta alias k-point 0 for the single-kappa entry points
  int nk = 1;
  std::vector<double> kappas = {0, 0, 0};   // [nk][3]
  std::vector<double> betas = {0};          // [nk]
  double kappa[3] = {0, 0, 0};
  double beta = 0;
  int sel = 0;                              // k-point the getters refer to (bloch_select_kpoint)
  bloch_b200::DevBuf<int> d_gflag;          // [nk] 1 where kappa == 0 (S0 singular on constants)
  bool any_gamma = false;
  bool dirty_coef = true, dirty_kappa = true;
  bool profile = false;

  bloch_b200::DevBuf<int> d_cls;
  bloch_b200::DevBuf<double> d_eps, d_muinv, d_cpar;
  bloch_b200::DevBuf<int32_t> d_map_nd, d_map_h1, d_map_rt;
  bloch_b200::DevBuf<int32_t> d_shared_rows; // ND dofs that are NOT interior to an element (rows the apply reduces into)
  long n_shared_rows = 0;
  bloch_b200::DevBuf<int> d_tp_ptr;          // transpose of map_nd: dof -> its local copies
  bloch_b200::DevBuf<int32_t> d_tp_loc;      // signed 1-based positions e*L_nd + j
  bloch_b200::DevBuf<D2> d_evec;             // E-vector of the atomic-free apply
  int two_pass = 0;            // 1: atomic-free E-vector + owner reduction (deterministic, ~15% slower)
  bloch_b200::DevBuf<double> d_diagA, d_diagM, d_diagS0;   // Jacobi diagonals (real)
  bloch_b200::DevBuf<double> d_jac, d_jac0;                // 1/(diagA + sigma diagM), 1/diagS0
  bloch_b200::ElemData E{};

  // solver controls / results
  int nbands = 10, block = 0, max_iter = 2000;
  double tol = 1e-6;
  double sigma = 0;                 // shift of the preconditioner A + sigma M
  int cheb_degree = 8;
  double lmaxA = 0;                 // bound of lambda_max(D^-1 (A + sigma M)) used by Chebyshev
  double lmax_local = 0;            // max over classes of lambda_max(diag(X_e)^-1 X_e), X = A, M
  std::vector<double> eigenvalues;  // ascending, [nk][nbands]
  bloch_b200::DevBuf<D2> d_X;       // eigenvectors, block layout [N][nk * block] (column k * block + j)
  int have_vectors = 0;             // number of valid columns per k-point in d_X
  // sweep history: the eigenvectors of the solve BEFORE the last one and the kappas of the last two solves.  When a
  // k-point continues a walk along a line (kappa_prev2 -> kappa_prev -> kappa in the same direction, similar steps),
  // span{X(kappa_prev), X(kappa_prev2)} contains the linear extrapolation of the bands: X(kappa_prev2) enters the
  // first Rayleigh-Ritz as the P block (solver.cu)
  bloch_b200::DevBuf<D2> d_Xhist;
  int have_hist = 0;                // columns per k-point valid in d_Xhist (0: none)
  std::vector<double> kappas_prev, kappas_prev2;   // [nk][3] of the last / last but one solve
  std::vector<double> init_vecs;    // user supplied [m][2N]
  int n_init = 0;
  bloch_b200::SolverStats stats;

  // assembled operators for the -wm dump (maxwell_dispersion.cpp:553-590)
  std::vector<int32_t> knd_host;                 // kernel-order ND map (host copy)
  std::vector<int64_t> csr_ptr;
  std::vector<int32_t> csr_col;
  std::vector<double> csr_re, csr_im;
  void assemble(int which);

  // field averages (maxwell_bloch.cpp:1550-1632)
  bloch_b200::DevBuf<double> d_x0, d_geom;
  bloch_b200::DevBuf<D2> d_fa_part, d_fa_out, d_fa_e, d_fa_b;
  bloch_b200::AvgTabs avg_tabs;
  bool avg_ready = false;
  void field_averages(int i, double out24[24]);

  // reduced-basis sweep (meta-material/meta_material_solver.cpp:3132-3305)
  bloch_b200::DevBuf<D2> d_rb, d_rb_p, d_rb_ap, d_rb_mp, d_rb_tmp;   // raw basis [N][rb_cap] and projected work arrays
  int rb_size = 0, rb_cap = 0;
  void rb_append();
  void rb_approx(double *lambda, int n);

  // scalar H1 variant (misc/scalar3d.cpp): stiffness coefficient k -> eps slot, mass coefficient m -> muinv slot
  bloch_b200::DevBuf<double> d_diagM0;
  double lmax_local_h1 = 0;
  int nbands_s = 5, block_s = 0, have_vectors_s = 0;
  std::vector<double> eigenvalues_s;
  bloch_b200::DevBuf<D2> d_Xs;
  bool scalar_ready = false;

  // probe problems (one private element copy per local unit vector), built once per space;
  // persistent so that Setup() never calls cudaMalloc/cudaFree (device-wide syncs)
  struct ProbeWork {
    bloch_b200::DevBuf<int32_t> map;
    bloch_b200::DevBuf<int> cls;
    bloch_b200::DevBuf<double> one;
    bloch_b200::DevBuf<D2> x, y;     // [n_class * L * L][nk]: the unit vectors replicated per k-point
    std::vector<D2> hy;
    bool built = false;
    int built_nk = 0;
  } probe_nd, probe_h1;
  bloch_b200::DevBuf<double> d_dloc;

  struct LobpcgWork {
    bloch_b200::DevBuf<D2> S, AS, MS, R, Wc, Dd, Tq, Qb, dC, dGA, dGM, Lu, Lphi, Lg;
    bloch_b200::DevBuf<double> dlam, drn, dtau;
    bloch_b200::DevBuf<unsigned char> dact, dusep;   // device Rayleigh-Ritz: active-column mask [nk][m], P-block flags [nk]
    bloch_b200::DevBuf<int> dinfo;
  } lw;

  // workspace of the divergence projector (inner S0 solve); owned by the handle so that it is released by
  // bloch_destroy and always lives on the handle's device
  struct ProjWork {
    bloch_b200::DevBuf<D2> rhs, phi, z, p, q, g;
    bloch_b200::DevBuf<double> scal;
  } pw;

  // scratch for host-pointer entry points
  bloch_b200::DevBuf<double> d_io_a, d_io_b;
  bloch_b200::DevBuf<D2> d_blk_a, d_blk_b, d_blk_c;

  void setup();                                 // Setup()
  void apply_nd(const D2 *x, D2 *y, int nvec, double ca, double cm);   // y = ca A x + cm M x
  void apply_nd_ld(const D2 *x, int ldx, D2 *y, int ldy, int nvec, double ca, double cm);
  void apply_h1(int mode, const D2 *x, D2 *y, int nvec);               // zeroes y for modes 0, 2
  void apply_curl(const D2 *x, D2 *y, int nvec);
  void project(D2 *x, int nvec, double rel_tol, int *iters);           // in place x <- P x
  bloch_b200::ElemData elem_of_k(int k) const {                         // single-kappa view of k-point k's class table
    bloch_b200::ElemData K = E;
    K.nk = 1;
    K.cpar = E.cpar + (size_t)k * E.n_class * bloch_b200::kClassParDoubles;
    return K;
  }
  void set_kappas(int n, const double *k3);
  void solve();
  void solve_scalar();
  void lobpcg(bloch_b200::EigProblem &prob);
  void apply_scalar_ld(const D2 *x, int ldx, D2 *y, int ldy, int nvec, double ca, double cm);
  void setup_scalar();
  void count_launch(int n = 1) { stats.launches += n; }

  // phase timing (bloch_set_profile / bloch_get_profile); categories: see SolverStats::prof_ms
  std::vector<cudaEvent_t> prof_pool;
  std::pair<cudaEvent_t, cudaEvent_t> prof_open[8] = {};
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_done[8];
  bool prof_in_precond = false;
  void prof_begin(int cat);
  void prof_end(int cat);
  void prof_collect();
};
