"""Shared helpers for the parity tests: build the oracle operators on the PRODUCT's dof maps so
that vectors can be compared entry by entry."""
import numpy as np

from oracle.bloch_oracle import BlochOperators, Lattice, Mesh, Spaces


def oracle_on_product_maps(eq, name, n, p, eps, muinv=None):
    lat = Lattice(name)
    mesh = Mesh(lat, n)
    x0, cls, J = eq.element_geometry()
    assert np.allclose(x0, mesh.x0, atol=1e-13) and np.allclose(J, mesh.J, atol=1e-13)
    assert (cls == mesh.cls).all()
    ng, ns = eq.dofmap("nd")
    hg, _ = eq.dofmap("h1")
    rg, rs = eq.dofmap("rt")
    sp = Spaces(mesh, p, dofmaps=dict(nd_gid=ng, nd_sign=ns, h1_gid=hg, rt_gid=rg, rt_sign=rs))
    return BlochOperators(sp, eps, muinv), mesh


def rel_err(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
