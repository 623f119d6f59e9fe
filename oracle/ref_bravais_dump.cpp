// Runs the reference's OWN lib/bravais.cpp (compiled unmodified against oracle/ref_shim/mfem.hpp, see oracle/Makefile
// target `_ref`) and dumps, as JSON, the lattice tables that the dispersion path depends on: lattice / reciprocal /
// translation vectors, cell volumes, face radii, symmetry points and labels, k-paths with their intermediate points,
// the coarse Wigner-Seitz hex cell (vertices + elements) and the vertex identification that the reference's
// MakePeriodicMesh derives from the translation vectors.  tests/golden/ref_bravais.json is this program's output;
// tests/test_ref_bravais.py compares the oracle and the product with it.  Test infrastructure only.
#include <cstdio>
#include <string>
#include <vector>

#include "bravais.hpp"

using namespace mfem;
using namespace mfem::bravais;

static void vec(const Vector &v) {
  std::printf("[");
  for (int i = 0; i < v.Size(); i++) std::printf("%s%.17g", i ? ", " : "", v[i]);
  std::printf("]");
}
static void vecs(const char *name, const std::vector<Vector> &a) {
  std::printf("  \"%s\": [", name);
  for (size_t i = 0; i < a.size(); i++) { if (i) std::printf(", "); vec(a[i]); }
  std::printf("],\n");
}
static void mesh_json(const char *name, Mesh *m) {
  std::printf("  \"%s\": {\"n_refine\": %d, \"vertices\": [", name, m->n_refine);
  for (int i = 0; i < m->GetNV(); i++) {
    const double *x = m->GetVertex(i);
    std::printf("%s[%.17g, %.17g, %.17g]", i ? ", " : "", x[0], x[1], x[2]);
  }
  std::printf("], \"elements\": [");
  for (int e = 0; e < m->GetNE(); e++) {
    const Element *el = m->GetElement(e);
    std::printf("%s{\"geom\": %d, \"attr\": %d, \"v\": [", e ? ", " : "", (int)el->GetGeometryType(), el->GetAttribute());
    for (int k = 0; k < el->GetNVertices(); k++) std::printf("%s%d", k ? ", " : "", el->GetVertices()[k]);
    std::printf("]}");
  }
  std::printf("]},\n");
}

int main() {
  struct Case { const char *tag; BRAVAIS_LATTICE_TYPE type; double a, b, c; } cases[] = {
      {"CUB", PRIMITIVE_CUBIC, 1.0, 1.0, 1.0},
      {"FCC", FACE_CENTERED_CUBIC, 1.0, 1.0, 1.0},
      {"BCC", BODY_CENTERED_CUBIC, 1.0, 1.0, 1.0},
      {"HEX", PRIMITIVE_HEXAGONAL_PRISM, 1.0, 1.0, 1.0},
      {"FCC_a2", FACE_CENTERED_CUBIC, 2.0, 2.0, 2.0},                 // scaled cells
      {"BCC_a0.5", BODY_CENTERED_CUBIC, 0.5, 0.5, 0.5},
      {"HEX_c1.5", PRIMITIVE_HEXAGONAL_PRISM, 1.0, 1.0, 1.5},
  };
  BravaisLatticeFactory fact;
  std::printf("{\n");
  for (size_t ci = 0; ci < sizeof(cases) / sizeof(cases[0]); ci++) {
    const Case &c = cases[ci];
    BravaisLattice *L = fact.GetLattice(c.type, c.a, c.b, c.c, 0.0, 0.0, 0.0);
    if (!L) { std::fprintf(stderr, "no lattice for %s\n", c.tag); return 1; }
    std::printf(" \"%s\": {\n", c.tag);
    std::printf("  \"type\": %d, \"label\": \"%s\", \"dim\": %u,\n", (int)L->GetLatticeType(), L->GetLatticeTypeLabel().c_str(), L->GetDim());
    std::printf("  \"cell_volume\": %.17g, \"bz_volume\": %.17g,\n", L->GetUnitCellVolume(), L->GetBrillouinZoneVolume());
    std::vector<Vector> a, b, t;
    std::vector<double> r;
    L->GetLatticeVectors(a); L->GetReciprocalLatticeVectors(b); L->GetTranslationVectors(t); L->GetFaceRadii(r);
    vecs("lattice_vectors", a); vecs("reciprocal_vectors", b); vecs("translation_vectors", t);
    std::printf("  \"face_radii\": [");
    for (size_t i = 0; i < r.size(); i++) std::printf("%s%.17g", i ? ", " : "", r[i]);
    std::printf("],\n  \"symmetry_points\": [");
    for (unsigned i = 0; i < L->GetNumberSymmetryPoints(); i++) {
      Vector p;
      L->GetSymmetryPoint(i, p);
      std::printf("%s{\"label\": \"%s\", \"kappa\": ", i ? ", " : "", L->GetSymmetryPointLabel(i).c_str());
      vec(p);
      std::printf(", \"index_of_label\": %d}", L->GetSymmetryPointIndex(L->GetSymmetryPointLabel(i)));
    }
    std::printf("],\n  \"paths\": [");
    for (unsigned p = 0; p < L->GetNumberPaths(); p++) {
      std::printf("%s[", p ? ", " : "");
      for (unsigned s = 0; s < L->GetNumberPathSegments(p); s++) {
        int e0, e1;
        L->GetPathSegmentEndPointIndices(p, s, e0, e1);
        Vector ip;
        L->GetIntermediatePoint(p, s, ip);
        std::printf("%s{\"e0\": %d, \"e1\": %d, \"mid_label\": \"%s\", \"mid\": ", s ? ", " : "", e0, e1,
                    L->GetIntermediatePointLabel(p, s).c_str());
        vec(ip);
        std::printf("}");
      }
      std::printf("]");
    }
    std::printf("],\n");
    // MapToPrimitiveCell on a fixed pseudo-random point set (linear congruential sequence in [-1.5, 1.5]^3)
    std::printf("  \"map_to_primitive_cell\": [");
    unsigned long long lcg = 12345;
    for (int k = 0; k < 24; k++) {
      Vector pt(3), ipt(3);
      for (int d = 0; d < 3; d++) {
        lcg = (lcg * 6364136223846793005ULL + 1442695040888963407ULL);
        pt[d] = 3.0 * ((double)(lcg >> 11) / 9007199254740992.0) - 1.5;
      }
      const bool moved = L->MapToPrimitiveCell(pt, ipt);
      std::printf("%s{\"pt\": ", k ? ", " : "");
      vec(pt);
      std::printf(", \"ipt\": ");
      vec(ipt);
      std::printf(", \"moved\": %d}", moved ? 1 : 0);
    }
    std::printf("],\n");
    // plane-wave phases of CreateInitialVectors: Real/ImagModeCoefficient::Eval (lib/bravais.cpp:8953-8976) at a few
    // points (the shim's ElementTransformation maps an integration point to itself)
    {
      RealModeCoefficient cr;
      ImagModeCoefficient ci;
      cr.SetReciprocalLatticeVectors(b); ci.SetReciprocalLatticeVectors(b);
      const int modes[3][3] = {{1, 0, 0}, {1, -1, 1}, {0, 2, -1}};
      ElementTransformation T;
      std::printf("  \"mode_coefficient\": [");
      bool first = true;
      for (auto &n : modes) {
        cr.SetModeIndices(n[0], n[1], n[2]); ci.SetModeIndices(n[0], n[1], n[2]);
        for (int k = 0; k < 4; k++) {
          IntegrationPoint ip;
          ip.Set3(0.11 + 0.17 * k, -0.23 + 0.05 * k * k, 0.31 - 0.09 * k);
          std::printf("%s{\"n\": [%d, %d, %d], \"x\": [%.17g, %.17g, %.17g], \"re\": %.17g, \"im\": %.17g}", first ? "" : ", ",
                      n[0], n[1], n[2], ip.x, ip.y, ip.z, cr.Eval(T, ip), ci.Eval(T, ip));
          first = false;
        }
      }
      std::printf("],\n");
    }
    // LatticeCoefficient (rods along the translation vectors, lib/bravais.cpp:9834-9862) on a pseudo-random point set
    {
      LatticeCoefficient lc(*L, 0.5, 1.0, 10.0);
      ElementTransformation T;
      std::printf("  \"lattice_coefficient\": {\"frac\": 0.5, \"val0\": 1.0, \"val1\": 10.0, \"samples\": [");
      unsigned long long lcg2 = 777;
      for (int k = 0; k < 40; k++) {
        double x[3];
        for (int d = 0; d < 3; d++) {
          lcg2 = (lcg2 * 6364136223846793005ULL + 1442695040888963407ULL);
          x[d] = ((double)(lcg2 >> 11) / 9007199254740992.0) - 0.5;
        }
        IntegrationPoint ip;
        ip.Set3(x[0], x[1], x[2]);
        std::printf("%s[%.17g, %.17g, %.17g, %.17g]", k ? ", " : "", x[0], x[1], x[2], lc.Eval(T, ip));
      }
      std::printf("]},\n");
    }
    Mesh *ws = L->GetWignerSeitzMesh(false);
    mesh_json("ws_mesh", ws);
    Mesh *per = MakePeriodicMesh(ws, t);
    mesh_json("ws_mesh_periodic", per);
    delete per;
    delete ws;
    std::printf("  \"n_transformations\": %u\n }%s\n", L->GetNumberTransformations(), ci + 1 < sizeof(cases) / sizeof(cases[0]) ? "," : "");
    delete L;
  }
  std::printf("}\n");
  return 0;
}
