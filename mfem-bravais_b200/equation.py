"""Host-side mirror of the reference interface for this path, over the C ABI:

  * BravaisLattice            <-> bravais::BravaisLattice           (lib/bravais.hpp:64-181)
  * MaxwellBlochWaveEquation  <-> mfem::bloch::MaxwellBlochWaveEquation (maxwell/maxwell_bloch.hpp:140-234)

Same method names and argument meaning as the reference; MFEM types are flattened to numpy
arrays.  Vectors of length 2N are stored [re(N); im(N)] like the reference's block vectors.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import BlochError, check, dptr


class BravaisLattice:
    """BravaisLatticeFactory(type, a, b, c, alpha, beta, gamma) (lib/bravais.cpp:8598-8846)."""

    def __init__(self, lattice_type, a=1.0, b=1.0, c=1.0, alpha=0.0, beta=0.0, gamma=0.0):
        if isinstance(lattice_type, str):
            lattice_type = capi.LATTICE_TYPES[lattice_type.upper()]
        self._L = capi.lib()
        self._h = C.c_void_p()
        check(self._L.bloch_lattice_create(C.byref(self._h), int(lattice_type), a, b, c, alpha, beta, gamma),
              "bloch_lattice_create")

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.bloch_lattice_destroy(self._h)
            self._h = C.c_void_p()

    def _label(self, fn, *args):
        buf = C.create_string_buffer(64)
        check(fn(self._h, *args, buf, 64))
        return buf.value.decode()

    def GetLatticeTypeLabel(self):
        return self._label(self._L.bloch_lattice_label)

    def GetUnitCellVolume(self):
        return self._L.bloch_lattice_volume(self._h)

    def GetLatticeVectors(self):
        a = np.zeros((3, 3))
        check(self._L.bloch_lattice_vectors(self._h, dptr(a), None))
        return a

    def GetReciprocalLatticeVectors(self):
        b = np.zeros((3, 3))
        check(self._L.bloch_lattice_vectors(self._h, None, dptr(b)))
        return b

    def GetTranslationVectors(self):
        n = self._L.bloch_lattice_num_translations(self._h)
        t = np.zeros((n, 3))
        check(self._L.bloch_lattice_translations(self._h, dptr(t), None))
        return t

    def GetFaceRadii(self):
        n = self._L.bloch_lattice_num_translations(self._h)
        r = np.zeros(n)
        check(self._L.bloch_lattice_translations(self._h, None, dptr(r)))
        return r

    def GetNumberSymmetryPoints(self):
        return self._L.bloch_lattice_num_symmetry_points(self._h)

    def GetSymmetryPoint(self, i):
        k = np.zeros(3)
        check(self._L.bloch_lattice_symmetry_point(self._h, i, dptr(k), None, 0))
        return k

    def GetSymmetryPointLabel(self, i):
        buf = C.create_string_buffer(64)
        check(self._L.bloch_lattice_symmetry_point(self._h, i, None, buf, 64))
        return buf.value.decode()

    def GetSymmetryPointIndex(self, label):
        return self._L.bloch_lattice_symmetry_point_index(self._h, label.encode())

    def GetNumberPaths(self):
        return self._L.bloch_lattice_num_paths(self._h)

    def GetNumberPathSegments(self, p):
        return check(self._L.bloch_lattice_num_path_segments(self._h, p))

    def GetPathSegmentEndPointIndices(self, p, s):
        e0, e1 = C.c_int(), C.c_int()
        check(self._L.bloch_lattice_path_segment(self._h, p, s, C.byref(e0), C.byref(e1)))
        return e0.value, e1.value

    def GetIntermediatePoint(self, p, s):
        k = np.zeros(3)
        check(self._L.bloch_lattice_intermediate_point(self._h, p, s, dptr(k), None, 0))
        return k

    def GetIntermediatePointLabel(self, p, s):
        buf = C.create_string_buffer(64)
        check(self._L.bloch_lattice_intermediate_point(self._h, p, s, None, buf, 64))
        return buf.value.decode()

    def MapToPrimitiveCell(self, pt):
        pt = np.ascontiguousarray(pt, float)
        out = np.zeros(3)
        mapped = self._L.bloch_lattice_map_to_primitive_cell(self._h, dptr(pt), dptr(out))
        return bool(mapped), out


class MaxwellBlochWaveEquation:
    """MaxwellBlochWaveEquation(pmesh, order): here the periodic Wigner-Seitz hex mesh is
    described by (lattice, n_sub) - n_sub = 2^r reproduces r uniform refinements."""

    def __init__(self, lattice, n_sub, order, device=-1):
        self._L = capi.lib()
        self._h = C.c_void_p()
        self.lattice = lattice
        check(self._L.bloch_create(C.byref(self._h), lattice._h, int(n_sub), int(order), int(device)),
              "bloch_create")
        ne, nc = C.c_int64(), C.c_int()
        check(self._L.bloch_num_elements(self._h, C.byref(ne), C.byref(nc)))
        self.n_elem, self.n_class = ne.value, nc.value
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        check(self._L.bloch_num_dofs(self._h, C.byref(a), C.byref(b), C.byref(c)))
        self.N, self.N_rt, self.N_h1 = a.value, b.value, c.value
        self.order = int(order)
        self.nev = 20

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.bloch_destroy(self._h)
            self._h = C.c_void_p()

    # ---- sizes / geometry ----
    def GetHCurlTrueVSize(self):
        return self.N

    def mesh_counts(self):
        v, e, f, vol = C.c_int64(), C.c_int64(), C.c_int64(), C.c_double()
        check(self._L.bloch_mesh_counts(self._h, C.byref(v), C.byref(e), C.byref(f), C.byref(vol)))
        return v.value, e.value, f.value, vol.value

    def element_centers(self):
        c = np.zeros((self.n_elem, 3))
        check(self._L.bloch_element_centers(self._h, dptr(c)))
        return c

    def element_geometry(self):
        x0 = np.zeros((self.n_elem, 3))
        cls = np.zeros(self.n_elem, dtype=np.int32)
        J = np.zeros((self.n_class, 3, 3))
        check(self._L.bloch_element_geometry(self._h, dptr(x0), cls.ctypes.data_as(C.POINTER(C.c_int)), dptr(J)))
        return x0, cls, J

    def dofmap(self, space):
        """space: 'h1' | 'nd' | 'rt' -> (gid[n_elem, L], sign[n_elem, L])"""
        sp = {"h1": 0, "nd": 1, "rt": 2}[space]
        L = self._L.bloch_local_size(self._h, sp)
        m = np.zeros((self.n_elem, L), dtype=np.int32)
        check(self._L.bloch_get_dofmap(self._h, sp, m.ctypes.data_as(C.POINTER(C.c_int32))))
        return np.abs(m).astype(np.int64) - 1, np.sign(m).astype(float)

    # ---- reference interface ----
    def SetMassCoef(self, eps_per_elem):
        e = np.ascontiguousarray(eps_per_elem, float)
        assert e.shape == (self.n_elem,)
        check(self._L.bloch_set_eps(self._h, dptr(e)), "bloch_set_eps")

    def SetStiffnessCoef(self, muinv_per_elem):
        e = np.ascontiguousarray(muinv_per_elem, float)
        assert e.shape == (self.n_elem,)
        check(self._L.bloch_set_muinv(self._h, dptr(e)), "bloch_set_muinv")

    def SetKappa(self, kappa):
        k = np.ascontiguousarray(kappa, float)
        self._kappa_set = k.copy()
        check(self._L.bloch_set_kappa(self._h, dptr(k)), "bloch_set_kappa")

    # ---- k-point batch: several Bloch vectors iterated together (independent eigenproblems, one set of kernel
    # launches; the k-loop of maxwell_dispersion.cpp:475-531 taken nk points at a time) ----
    def SetKappaBatch(self, kappas):
        k = np.ascontiguousarray(kappas, float).reshape(-1, 3)
        check(self._L.bloch_set_kappa_batch(self._h, k.shape[0], dptr(k)), "bloch_set_kappa_batch")

    def BatchSize(self):
        return int(self._L.bloch_batch_size(self._h))

    def SelectKPoint(self, k):
        """the getters (band_eigenvalues, GetEigenvector*, GetSolverStats, GetFieldAverages) refer to k-point k"""
        check(self._L.bloch_select_kpoint(self._h, int(k)), "bloch_select_kpoint")

    def SolveBatch(self, kappas):
        """SetKappaBatch + Setup + Solve; returns (eigenvalues[nk, n_bands], stats per k-point)"""
        self.SetKappaBatch(kappas)
        self.Setup()
        self.Solve()
        lam, stats = [], []
        for k in range(self.BatchSize()):
            self.SelectKPoint(k)
            lam.append(self.band_eigenvalues())
            stats.append(self.GetSolverStats())
        self.SelectKPoint(0)
        return np.array(lam), stats

    def SetProfile(self, on):
        check(self._L.bloch_set_profile(self._h, 1 if on else 0))

    def GetProfile(self):
        """phase times (ms) of the last Solve() with profiling on, see bloch_get_profile in include/bloch_b200.h"""
        ms = np.zeros(8)
        check(self._L.bloch_get_profile(self._h, dptr(ms), 8))
        names = ["solve", "nd_apply_outside_precond", "projector", "precond", "gram_rotation", "host_rr", "lift_extra",
                 "nd_apply_in_precond"]
        return dict(zip(names, ms.tolist()))

    def SetNumEigs(self, nev):
        """nev counts REAL modes like the reference (2 per complex band)."""
        self.nev = int(nev)
        check(self._L.bloch_set_num_bands(self._h, (self.nev + 1) // 2), "bloch_set_num_bands")

    def SetAbsoluteTolerance(self, atol, max_iter=2000):
        check(self._L.bloch_set_tol(self._h, float(atol), int(max_iter)), "bloch_set_tol")

    def Setup(self):
        check(self._L.bloch_setup(self._h), "bloch_setup")

    def SetInitialVectors(self, vecs):
        if vecs is None or len(vecs) == 0:
            check(self._L.bloch_set_initial_vectors(self._h, 0, None))
            return
        v = np.ascontiguousarray(vecs, float)
        assert v.ndim == 2 and v.shape[1] == 2 * self.N
        check(self._L.bloch_set_initial_vectors(self._h, v.shape[0], dptr(v)), "bloch_set_initial_vectors")

    def Solve(self):
        check(self._L.bloch_solve(self._h), "bloch_solve")

    def GetEigenvalues(self, nev=None, kappa=None, init_vecs=None):
        """GetEigenvalues(eigenvalues) or the 4-argument convenience form
        (SetNumEigs + SetKappa + Setup + SetInitialVectors + Solve + get, maxwell_bloch.cpp:1078-1097).
        Returns nev values: each complex band twice, like the reference's real block form."""
        if kappa is not None:
            self.SetNumEigs(nev)
            self.SetKappa(kappa)
            self.Setup()
            if init_vecs is not None:
                self.SetInitialVectors(init_vecs)
            self.Solve()
        nb = (self.nev + 1) // 2
        lam = np.zeros(nb)
        check(self._L.bloch_get_eigenvalues(self._h, dptr(lam), nb), "bloch_get_eigenvalues")
        return np.repeat(lam, 2)[: self.nev]

    def band_eigenvalues(self):
        nb = (self.nev + 1) // 2
        lam = np.zeros(nb)
        check(self._L.bloch_get_eigenvalues(self._h, dptr(lam), nb), "bloch_get_eigenvalues")
        return lam

    def AssembleMatrix(self, which):
        """'A' (complex Hermitian S1 - i beta DKZ) or 'M' (real M1(eps)) as scipy CSR - the operators the
        reference dumps with -wm (maxwell_dispersion.cpp:553-590)."""
        import scipy.sparse as sp
        w = {"A": 0, "M": 1}[which]
        nnz = C.c_int64()
        check(self._L.bloch_assemble_matrix(self._h, w, C.byref(nnz)), "bloch_assemble_matrix")
        ptr = np.zeros(self.N + 1, np.int64)
        col = np.zeros(nnz.value, np.int32)
        re, im = np.zeros(nnz.value), np.zeros(nnz.value)
        check(self._L.bloch_get_matrix(self._h, ptr.ctypes.data_as(C.POINTER(C.c_int64)),
                                       col.ctypes.data_as(C.POINTER(C.c_int32)), dptr(re), dptr(im)), "bloch_get_matrix")
        data = re + 1j * im if w == 0 else re
        return sp.csr_matrix((data, col, ptr), shape=(self.N, self.N))

    def ProlongEigenvectorsTo(self, fine):
        """Interpolates this (coarse) equation's eigenvectors onto `fine` (its uniform refinement) and
        installs them as the starting block of fine.Solve() (meta_material_solver.cpp:2829-2853)."""
        check(self._L.bloch_prolong_eigenvectors(self._h, fine._h), "bloch_prolong_eigenvectors")

    # reduced-basis sweep pieces (MaxwellDispersion, meta_material_solver.cpp:3132-3305)
    def ReducedBasisClear(self):
        check(self._L.bloch_rb_clear(self._h), "bloch_rb_clear")

    def ReducedBasisAppend(self):
        """Appends the bands of the last Solve() to the device-resident raw basis."""
        check(self._L.bloch_rb_append(self._h), "bloch_rb_append")

    def ReducedBasisSize(self):
        return int(self._L.bloch_rb_size(self._h))

    def ApproxEigenvalues(self, kappa, n_bands):
        """approxEigenfrequencies (meta_material_solver.cpp:3213-3305) at kappa, as eigenvalues."""
        k = np.ascontiguousarray(kappa, float)
        lam = np.zeros(int(n_bands))
        check(self._L.bloch_rb_approx(self._h, dptr(k), dptr(lam), int(n_bands)), "bloch_rb_approx")
        return lam

    def GetEigenvectorE(self, i):
        re, im = np.zeros(self.N), np.zeros(self.N)
        check(self._L.bloch_get_eigenvector_E(self._h, i, dptr(re), dptr(im)), "bloch_get_eigenvector_E")
        return re, im

    def GetEigenvectorB(self, i):
        re, im = np.zeros(self.N_rt), np.zeros(self.N_rt)
        check(self._L.bloch_get_eigenvector_B(self._h, i, dptr(re), dptr(im)), "bloch_get_eigenvector_B")
        return re, im

    def GetEigenvector(self, i):
        """GetEigenvector(i, Er, Ei, Br, Bi) (maxwell_bloch.hpp:187-190, .cpp:1371-1458): all four parts of mode i"""
        er, ei = self.GetEigenvectorE(i)
        br, bi = self.GetEigenvectorB(i)
        return er, ei, br, bi

    def IdentifyDegeneracies(self, zero_tol, rel_tol):
        """IdentifyDegeneracies (maxwell_bloch.cpp:1493-1548): groups of (real-mode) eigenvalue indices that are
        zero (< zero_tol) or agree to rel_tol relative to their mean; list of sets in ascending order."""
        ev = self.GetEigenvalues()
        if len(ev) == 0:
            return []
        degen = [{0}]
        zeroes = ev[0] < zero_tol
        for i in range(1, len(ev)):
            if zeroes:
                if ev[i] > zero_tol:
                    degen.append(set())
                    zeroes = False
            elif abs(ev[i] - ev[i - 1]) > 0.5 * (ev[i] + ev[i - 1]) * rel_tol:
                degen.append(set())
            degen[-1].add(i)
        return degen

    def DetermineBasis(self, v1):
        """DetermineBasis (maxwell_bloch.cpp:1700-1727): right-handed orthonormal frame (e0, e1, e2) with e2 = kappa /
        |kappa|, e1 = the part of v1 orthogonal to kappa, e0 = e1 x e2; the Cartesian frame when |kappa| < 1e-4.
        (The reference's body subtracts (e2.v1) v1 instead of (e2.v1) e2 and assigns e0[0] twice, leaving e0[1]
        unset - INTEGRATION.md section 4; this returns the frame those lines are evidently meant to build.)"""
        kappa = self._kappa0()
        kn = np.linalg.norm(kappa)
        if kn < 1e-4:
            return [np.eye(3)[i] for i in range(3)]
        e2 = kappa / kn
        v1 = np.asarray(v1, float)
        e1 = v1 - (e2 @ v1) * e2
        e1 /= np.linalg.norm(e1)
        return [np.cross(e1, e2), e1, e2]

    def ComputeHomogenizedCoefs(self):
        """empty in the reference too (maxwell_bloch.cpp:1694-1698); the effective-medium inputs are GetFieldAverages"""
        return None

    def _kappa0(self):
        return np.array(getattr(self, "_kappa_set", np.zeros(3)), float)

    def WriteVisitFields(self, prefix, label, eps=None, muinv=None):
        """WriteVisitFields (maxwell_bloch.cpp:1730-1822): E_r, E_i, B_r, B_i of every mode (cycle = mode number,
        time = omega) plus the coefficient fields.  The reference writes an MFEM VisIt data collection; here one
        legacy-VTK file per mode, `<prefix>/<label>_<cycle:06d>.vtk`, with the fields evaluated at the element corners
        (VisIt and ParaView read these directly) and `<prefix>/<label>.visit` listing them with their times."""
        import os
        from .dispersion import evaluate_fields, omega_of_lambda, write_vtk_fields
        os.makedirs(prefix, exist_ok=True)
        lam = self.GetEigenvalues()
        om = omega_of_lambda(lam)
        names = []
        cells = {}
        if eps is not None:
            cells["epsilon"] = eps
        if muinv is not None:
            cells["muInv"] = muinv
        for i in range(self.nev):
            band = i // 2
            er, ei, br, bi = self.GetEigenvector(band)
            if i & 1:      # the reference's second real mode of a complex band is i times the first
                er, ei, br, bi = -ei, er, -bi, br
            _, E, B = evaluate_fields(self, np.concatenate([er, ei]), np.concatenate([br, bi]))
            fn = "%s_%06d.vtk" % (label, i + 1)
            write_vtk_fields(self, os.path.join(prefix, fn), {"E_r": E.real, "E_i": E.imag, "B_r": B.real, "B_i": B.imag}, cells)
            names.append((fn, om[i]))
        with open(os.path.join(prefix, label + ".visit"), "w") as f:
            f.write("!NBLOCKS 1\n")
            for fn, t in names:
                f.write("%s\n" % fn)
        with open(os.path.join(prefix, label + ".times"), "w") as f:
            for k, (fn, t) in enumerate(names):
                f.write("%d %s %.12g\n" % (k + 1, fn, t))
        return [n for n, _ in names]

    def GetFieldAverages(self, i):
        """GetFieldAverages (maxwell_bloch.cpp:1550-1632) of band i: dict of complex 3-vectors E, B, D, H."""
        o = np.zeros(24)
        check(self._L.bloch_get_field_averages(self._h, int(i), dptr(o)), "bloch_get_field_averages")
        return {"E": o[0:3] + 1j * o[3:6], "B": o[6:9] + 1j * o[9:12],
                "D": o[12:15] + 1j * o[15:18], "H": o[18:21] + 1j * o[21:24]}

    def GetSolverStats(self):
        st = capi.bloch_stats()
        check(self._L.bloch_get_stats(self._h, C.byref(st)))
        return {k: getattr(st, k) for k, _ in st._fields_}

    def _apply(self, fn, x, nout):
        x = np.ascontiguousarray(x, float)
        single = x.ndim == 1
        x2 = x.reshape(1, -1) if single else x
        y = np.zeros((x2.shape[0], 2 * nout))
        check(fn(self._h, dptr(x2), dptr(y), x2.shape[0]))
        return y[0] if single else y

    def MultA(self, x):      # GetAOperator()->Mult
        return self._apply(self._L.bloch_apply_A, x, self.N)

    def MultM(self, x):      # GetMOperator()->Mult
        return self._apply(self._L.bloch_apply_M, x, self.N)

    def MultProjector(self, x):   # GetSubSpaceProjector()->Mult
        return self._apply(self._L.bloch_apply_projector, x, self.N)

    def MultC(self, x):      # C_ block operator (curl + kappa cross)
        return self._apply(self._L.bloch_apply_C, x, self.N_rt)

    def debug_h1op(self, mode, x):
        x = np.ascontiguousarray(x, float)
        x2 = x.reshape(1, -1) if x.ndim == 1 else x
        nout = self.N if mode == 1 else self.N_h1
        y = np.zeros((x2.shape[0], 2 * nout))
        check(self._L.bloch_debug_apply_h1op(self._h, mode, dptr(x2), dptr(y), x2.shape[0]))
        return y[0] if x.ndim == 1 else y

    def debug_aux(self, mode, x):
        """Pieces of the auxiliary-space preconditioner (bloch_debug_apply_aux): 0 = Pi, 1 = Pi^T, 2 = V-cycles."""
        x = np.ascontiguousarray(x, float)
        x2 = x.reshape(1, -1) if x.ndim == 1 else x
        nout = self.N if mode == 0 else 3 * self.N_h1
        y = np.zeros((x2.shape[0], 2 * nout))
        check(self._L.bloch_debug_apply_aux(self._h, mode, dptr(x2), dptr(y), x2.shape[0]))
        return y[0] if x.ndim == 1 else y

    def pi_matrix(self):
        """Nodal interpolation Pi: (H1)^3 -> ND of the auxiliary-space preconditioner as scipy CSR [N, 3 N_h1]
        (bloch_debug_pi_matrix; host only, works on topology-only handles)."""
        import scipy.sparse as sp
        nnz = C.c_int64()
        check(self._L.bloch_debug_pi_matrix(self._h, C.byref(nnz), None, None, None), "bloch_debug_pi_matrix")
        ptr = np.zeros(self.N + 1, np.int64)
        col = np.zeros(nnz.value, np.int32)
        val = np.zeros(nnz.value)
        check(self._L.bloch_debug_pi_matrix(self._h, C.byref(nnz), ptr.ctypes.data_as(C.POINTER(C.c_int64)),
                                            col.ctypes.data_as(C.POINTER(C.c_int32)), dptr(val)), "bloch_debug_pi_matrix")
        return sp.csr_matrix((val, col, ptr), shape=(self.N, 3 * self.N_h1))

    def debug_mg_transfer(self, variant, direction, x):
        """Level 0 <-> 1 transfer of the H1 multigrid (bloch_debug_mg_transfer); x: [nvec, 2 n_in] in [re; im] layout."""
        nc = C.c_int64()
        check(self._L.bloch_debug_mg_transfer(self._h, variant, direction, None, None, 0, C.byref(nc)))
        x = np.ascontiguousarray(x, float)
        nout = self.N_h1 if direction == 0 else nc.value
        y = np.zeros((x.shape[0], 2 * nout))
        check(self._L.bloch_debug_mg_transfer(self._h, variant, direction, dptr(x), dptr(y), x.shape[0], C.byref(nc)))
        return y

    def mg_coarse_size(self):
        nc = C.c_int64()
        check(self._L.bloch_debug_mg_transfer(self._h, 0, 0, None, None, 0, C.byref(nc)))
        return int(nc.value)

    def fp64_peak_tflops(self):
        v = C.c_double()
        check(self._L.bloch_debug_fp64_peak(self._h, C.byref(v)))
        return v.value

    # device-pointer entry points (integers = CUDA device addresses, e.g. torch tensor.data_ptr())
    def set_stream(self, stream_ptr):
        check(self._L.bloch_set_stream(self._h, C.c_void_p(stream_ptr)))

    def apply_A_device(self, x_ptr, y_ptr, nvec):
        check(self._L.bloch_apply_A_device(self._h, C.c_void_p(x_ptr), C.c_void_p(y_ptr), nvec))

    def apply_M_device(self, x_ptr, y_ptr, nvec):
        check(self._L.bloch_apply_M_device(self._h, C.c_void_p(x_ptr), C.c_void_p(y_ptr), nvec))

    def pack_device(self, reim_ptr, blk_ptr, nvec):
        check(self._L.bloch_pack_device(self._h, C.c_void_p(reim_ptr), C.c_void_p(blk_ptr), nvec))

    def unpack_device(self, blk_ptr, reim_ptr, nvec):
        check(self._L.bloch_unpack_device(self._h, C.c_void_p(blk_ptr), C.c_void_p(reim_ptr), nvec))


class ScalarFloquetWaveEquation:
    """ScalarFloquetWaveEquation(pmesh, order) of misc/scalar3d.cpp:592-898 - the scalar H1 Bloch
    Helmholtz variant  (G - i Z_kappa)^T M1(k) (G - i Z_kappa) u = lambda M0(m) u.  The reference takes
    the phase shift beta in degrees (SetBeta) and a unit direction zeta (SetZeta)."""

    def __init__(self, lattice, n_sub, order, device=-1):
        self._eq = MaxwellBlochWaveEquation(lattice, n_sub, order, device)
        self._L, self._h = self._eq._L, self._eq._h
        self.n_elem, self.N = self._eq.n_elem, self._eq.N_h1
        self._beta_deg, self._zeta = 0.0, np.array([1.0, 0.0, 0.0])
        self.nev = 10

    def element_centers(self):
        return self._eq.element_centers()

    def dofmap(self):
        return self._eq.dofmap("h1")[0]

    def element_geometry(self):
        return self._eq.element_geometry()

    def SetMassCoef(self, m_per_elem):
        self._m = np.ascontiguousarray(m_per_elem, float)
        self._push()

    def SetStiffnessCoef(self, k_per_elem):
        self._k = np.ascontiguousarray(k_per_elem, float)
        self._push()

    def _push(self):
        if hasattr(self, "_m") and hasattr(self, "_k"):
            check(self._L.bloch_scalar_set_coefs(self._h, dptr(self._k), dptr(self._m)), "bloch_scalar_set_coefs")

    def SetBeta(self, beta_degrees):
        self._beta_deg = float(beta_degrees)
        self._push_kappa()

    def SetZeta(self, zeta):
        self._zeta = np.asarray(zeta, float)
        self._push_kappa()

    def SetAzimuth(self, alpha_a_degrees):
        """direction of the phase shift from azimuth / inclination in degrees (scalar3d.cpp:109-110, 665-669)"""
        self._alpha_a = float(alpha_a_degrees)
        self._push_angles()

    def SetInclination(self, alpha_i_degrees):
        self._alpha_i = float(alpha_i_degrees)
        self._push_angles()

    def _push_angles(self):
        d = np.pi / 180.0
        a, i = getattr(self, "_alpha_a", 0.0), getattr(self, "_alpha_i", 90.0)      # class defaults, scalar3d.cpp:594-595
        self.SetZeta([np.cos(i * d) * np.cos(a * d), np.cos(i * d) * np.sin(a * d), np.sin(i * d)])

    def SetKappa(self, kappa):
        self._eq.SetKappa(kappa)

    def _push_kappa(self):
        self._eq.SetKappa(self._beta_deg * np.pi / 180.0 * self._zeta)     # scalar3d.cpp:733,784-785

    def SetNumEigs(self, nev):
        """real-mode count like the reference (2 per complex mode)"""
        self.nev = int(nev)
        check(self._L.bloch_scalar_set_num_modes(self._h, (self.nev + 1) // 2), "bloch_scalar_set_num_modes")

    def SetAbsoluteTolerance(self, atol, max_iter=1000):
        self._eq.SetAbsoluteTolerance(atol, max_iter)

    def Setup(self):
        self._eq.Setup()

    def Solve(self):
        check(self._L.bloch_scalar_solve(self._h), "bloch_scalar_solve")

    def mode_eigenvalues(self):
        nb = (self.nev + 1) // 2
        lam = np.zeros(nb)
        check(self._L.bloch_scalar_get_eigenvalues(self._h, dptr(lam), nb), "bloch_scalar_get_eigenvalues")
        return lam

    def GetEigenvalues(self):
        return np.repeat(self.mode_eigenvalues(), 2)[: self.nev]

    def GetSolverStats(self):
        return self._eq.GetSolverStats()

    def _apply(self, which, x):
        x = np.ascontiguousarray(x, float)
        x2 = x.reshape(1, -1) if x.ndim == 1 else x
        y = np.zeros_like(x2)
        check(self._L.bloch_scalar_apply(self._h, which, dptr(x2), dptr(y), x2.shape[0]), "bloch_scalar_apply")
        return y[0] if x.ndim == 1 else y

    def MultA(self, x):
        return self._apply(0, x)

    def MultM(self, x):
        return self._apply(1, x)
