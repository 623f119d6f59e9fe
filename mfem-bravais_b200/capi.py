"""ctypes binding of include/bloch_b200.h (the same stub a C++/MFEM maintainer would write with
`extern "C"` declarations)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
lib_path = os.path.join(_HERE, "lib", "libbloch_b200.so")

LATTICE_TYPES = {"CUB": 7, "FCC": 8, "BCC": 9, "HEX": 16}


class BlochError(RuntimeError):
    pass


class bloch_stats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("converged_bands", C.c_int), ("inner_iterations", C.c_int),
                ("solve_seconds", C.c_double), ("max_residual", C.c_double),
                ("applies_A", C.c_int64), ("kernel_launches", C.c_int64)]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol declared in include/bloch_b200.h
SIGNATURES = {
    "bloch_last_error": (C.c_char_p, []),
    "bloch_version": (C.c_int, []),
    "bloch_device_count": (C.c_int, []),
    "bloch_lattice_create": (C.c_int, [C.POINTER(_vp), C.c_int] + [C.c_double] * 6),
    "bloch_lattice_destroy": (C.c_int, [_vp]),
    "bloch_lattice_label": (C.c_int, [_vp, C.c_char_p, C.c_int]),
    "bloch_lattice_volume": (C.c_double, [_vp]),
    "bloch_lattice_vectors": (C.c_int, [_vp, _dp, _dp]),
    "bloch_lattice_num_translations": (C.c_int, [_vp]),
    "bloch_lattice_translations": (C.c_int, [_vp, _dp, _dp]),
    "bloch_lattice_num_symmetry_points": (C.c_int, [_vp]),
    "bloch_lattice_symmetry_point": (C.c_int, [_vp, C.c_int, _dp, C.c_char_p, C.c_int]),
    "bloch_lattice_symmetry_point_index": (C.c_int, [_vp, C.c_char_p]),
    "bloch_lattice_num_paths": (C.c_int, [_vp]),
    "bloch_lattice_num_path_segments": (C.c_int, [_vp, C.c_int]),
    "bloch_lattice_path_segment": (C.c_int, [_vp, C.c_int, C.c_int, _ip, _ip]),
    "bloch_lattice_intermediate_point": (C.c_int, [_vp, C.c_int, C.c_int, _dp, C.c_char_p, C.c_int]),
    "bloch_lattice_map_to_primitive_cell": (C.c_int, [_vp, _dp, _dp]),
    "bloch_create": (C.c_int, [C.POINTER(_vp), _vp, C.c_int, C.c_int, C.c_int]),
    "bloch_create_from_hexes": (C.c_int, [C.POINTER(_vp), C.c_int, _dp, C.c_int, _ip, _dp, C.c_int, C.c_int, C.c_int]),
    "bloch_destroy": (C.c_int, [_vp]),
    "bloch_set_stream": (C.c_int, [_vp, _vp]),
    "bloch_num_elements": (C.c_int, [_vp, _i64p, _ip]),
    "bloch_num_dofs": (C.c_int, [_vp, _i64p, _i64p, _i64p]),
    "bloch_mesh_counts": (C.c_int, [_vp, _i64p, _i64p, _i64p, _dp]),
    "bloch_element_centers": (C.c_int, [_vp, _dp]),
    "bloch_element_geometry": (C.c_int, [_vp, _dp, _ip, _dp]),
    "bloch_local_size": (C.c_int, [_vp, C.c_int]),
    "bloch_get_dofmap": (C.c_int, [_vp, C.c_int, _i32p]),
    "bloch_set_eps": (C.c_int, [_vp, _dp]),
    "bloch_set_muinv": (C.c_int, [_vp, _dp]),
    "bloch_set_kappa": (C.c_int, [_vp, _dp]),
    "bloch_set_kappa_batch": (C.c_int, [_vp, C.c_int, _dp]),
    "bloch_batch_size": (C.c_int, [_vp]),
    "bloch_select_kpoint": (C.c_int, [_vp, C.c_int]),
    "bloch_set_profile": (C.c_int, [_vp, C.c_int]),
    "bloch_get_profile": (C.c_int, [_vp, _dp, C.c_int]),
    "bloch_set_num_bands": (C.c_int, [_vp, C.c_int]),
    "bloch_set_tol": (C.c_int, [_vp, C.c_double, C.c_int]),
    "bloch_setup": (C.c_int, [_vp]),
    "bloch_set_initial_vectors": (C.c_int, [_vp, C.c_int, _dp]),
    "bloch_solve": (C.c_int, [_vp]),
    "bloch_get_eigenvalues": (C.c_int, [_vp, _dp, C.c_int]),
    "bloch_get_eigenvector_E": (C.c_int, [_vp, C.c_int, _dp, _dp]),
    "bloch_get_eigenvector_B": (C.c_int, [_vp, C.c_int, _dp, _dp]),
    "bloch_get_stats": (C.c_int, [_vp, C.POINTER(bloch_stats)]),
    "bloch_apply_A": (C.c_int, [_vp, _dp, _dp, C.c_int]),
    "bloch_apply_M": (C.c_int, [_vp, _dp, _dp, C.c_int]),
    "bloch_apply_projector": (C.c_int, [_vp, _dp, _dp, C.c_int]),
    "bloch_apply_C": (C.c_int, [_vp, _dp, _dp, C.c_int]),
    "bloch_apply_A_device": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "bloch_apply_M_device": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "bloch_pack_device": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "bloch_unpack_device": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "bloch_debug_hegv": (C.c_int, [C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_int]),
    "bloch_debug_hegv_device": (C.c_int, [C.c_int, C.c_int, C.c_int, _dp, _dp, C.POINTER(C.c_ubyte), C.c_int, _dp, _dp, _ip]),
    "bloch_assemble_matrix": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int64)]),
    "bloch_get_matrix": (C.c_int, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int32), _dp, _dp]),
    "bloch_prolong_eigenvectors": (C.c_int, [_vp, _vp]),
    "bloch_get_field_averages": (C.c_int, [_vp, C.c_int, _dp]),
    "bloch_rb_clear": (C.c_int, [_vp]),
    "bloch_rb_append": (C.c_int, [_vp]),
    "bloch_rb_size": (C.c_int, [_vp]),
    "bloch_rb_approx": (C.c_int, [_vp, _dp, _dp, C.c_int]),
    "bloch_scalar_set_coefs": (C.c_int, [_vp, _dp, _dp]),
    "bloch_scalar_set_num_modes": (C.c_int, [_vp, C.c_int]),
    "bloch_scalar_solve": (C.c_int, [_vp]),
    "bloch_scalar_get_eigenvalues": (C.c_int, [_vp, _dp, C.c_int]),
    "bloch_scalar_apply": (C.c_int, [_vp, C.c_int, _dp, _dp, C.c_int]),
    "bloch_debug_apply_h1op": (C.c_int, [_vp, C.c_int, _dp, _dp, C.c_int]),
    "bloch_debug_apply_aux": (C.c_int, [_vp, C.c_int, _dp, _dp, C.c_int]),
    "bloch_debug_pi_matrix": (C.c_int, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32), _dp]),
    "bloch_debug_mg_transfer": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, C.c_int, C.POINTER(C.c_int64)]),
    "bloch_debug_fp64_peak": (C.c_int, [_vp, _dp]),
}

_lib = None


def lib():
    """Loads libbloch_b200.so (built by __graft_entry__.build() / csrc/Makefile).  Fails loudly
    if it is missing: there is no fallback implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(lib_path):
            raise BlochError("%s not found - run `python -c 'import __graft_entry__ as g; g.build()'`" % lib_path)
        L = C.CDLL(lib_path)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(rc, what=""):
    if rc < 0:
        msg = lib().bloch_last_error()
        raise BlochError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))
    return rc


def dptr(a):
    return a.ctypes.data_as(_dp)
