// Minimal stand-in for the parts of MFEM that the reference's lib/bravais.cpp touches, written for ONE purpose:
// compile that file UNMODIFIED, from where it lies under /root/reference, so that its lattice tables (lattice /
// reciprocal / translation vectors, symmetry points, labels, k-paths, Wigner-Seitz cell vertex and element tables)
// can be executed here and pin the oracle's and the product's restatement of them (oracle/Makefile target `_ref`).
// This is test infrastructure: no MFEM code is reproduced, only the interface subset (names, argument meaning) the
// reference calls.  Vector / DenseMatrix carry real arithmetic; Mesh records vertices and elements and implements
// just what the table builders need (no refinement: UniformRefinement is a no-op, so only COARSE cells are dumped).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#define MFEM_ASSERT(c, msg) do { if (!(c)) { std::cerr << "MFEM_ASSERT failed: " << msg << std::endl; std::abort(); } } while (0)
#define MFEM_VERIFY(c, msg) MFEM_ASSERT(c, msg)
#define MFEM_ABORT(msg) do { std::cerr << "MFEM_ABORT: " << msg << std::endl; std::abort(); } while (0)

namespace mfem {

inline void mfem_error(const char *m = "") { std::cerr << "mfem_error: " << m << std::endl; std::abort(); }

template <class T>
class Array {
  std::vector<T> d_;
public:
  Array() {}
  explicit Array(int n) : d_(n) {}
  Array(T *p, int n) : d_(p, p + n) {}
  int Size() const { return (int)d_.size(); }
  void SetSize(int n) { d_.resize(n); }
  void SetSize(int n, const T &v) { d_.resize(n, v); }
  T &operator[](int i) { return d_[i]; }
  const T &operator[](int i) const { return d_[i]; }
  void Append(const T &v) { d_.push_back(v); }
  T *GetData() { return d_.data(); }
  const T *GetData() const { return d_.data(); }
  operator T *() { return d_.data(); }
  operator const T *() const { return d_.data(); }
  Array &operator=(const T &v) { for (auto &x : d_) x = v; return *this; }
  T Max() const { T m = d_[0]; for (auto &x : d_) if (x > m) m = x; return m; }
  void Sort() { std::sort(d_.begin(), d_.end()); }
  void Print(std::ostream &os = std::cout, int w = 4) const { for (int i = 0; i < Size(); i++) os << d_[i] << ((i + 1) % w ? " " : "\n"); os << "\n"; }
};

class Vector {
  std::vector<double> own_;
  double *p_ = nullptr;      // own_.data() or external memory (SetData)
  int n_ = 0;
  void own(int n) { own_.assign(n, 0.0); p_ = own_.data(); n_ = n; }
public:
  Vector() {}
  explicit Vector(int n) { own(n); }
  Vector(double *p, int n) : p_(p), n_(n) {}                 // aliases, like MFEM
  Vector(const Vector &o) { own(o.n_); for (int i = 0; i < n_; i++) p_[i] = o.p_[i]; }
  Vector &operator=(const Vector &o) {
    if (this != &o) { if (n_ != o.n_) own(o.n_); for (int i = 0; i < n_; i++) p_[i] = o.p_[i]; }
    return *this;
  }
  int Size() const { return n_; }
  void SetSize(int n) {
    if (n == n_) return;
    std::vector<double> keep(p_, p_ + (n < n_ ? n : n_));
    own(n);
    for (size_t i = 0; i < keep.size(); i++) p_[i] = keep[i];
  }
  void SetData(double *p) { p_ = p; }
  void SetDataAndSize(double *p, int n) { p_ = p; n_ = n; }
  double &operator[](int i) { return p_[i]; }
  const double &operator[](int i) const { return p_[i]; }
  double &operator()(int i) { return p_[i]; }
  const double &operator()(int i) const { return p_[i]; }
  double *GetData() { return p_; }
  const double *GetData() const { return p_; }
  operator double *() { return p_; }
  operator const double *() const { return p_; }
  Vector &operator=(double v) { for (int i = 0; i < n_; i++) p_[i] = v; return *this; }
  Vector &operator*=(double v) { for (int i = 0; i < n_; i++) p_[i] *= v; return *this; }
  Vector &operator/=(double v) { for (int i = 0; i < n_; i++) p_[i] /= v; return *this; }
  Vector &operator+=(const Vector &o) { for (int i = 0; i < n_; i++) p_[i] += o[i]; return *this; }
  Vector &operator-=(const Vector &o) { for (int i = 0; i < n_; i++) p_[i] -= o[i]; return *this; }
  Vector &operator-=(double v) { for (int i = 0; i < n_; i++) p_[i] -= v; return *this; }
  Vector &operator+=(double v) { for (int i = 0; i < n_; i++) p_[i] += v; return *this; }
  double operator*(const Vector &o) const { double s = 0; for (int i = 0; i < n_; i++) s += p_[i] * o[i]; return s; }
  double operator*(const double *o) const { double s = 0; for (int i = 0; i < n_; i++) s += p_[i] * o[i]; return s; }
  Vector &Set(double a, const Vector &x) { SetSize(x.Size()); for (int i = 0; i < n_; i++) p_[i] = a * x[i]; return *this; }
  Vector &Add(double a, const Vector &x) { for (int i = 0; i < n_; i++) p_[i] += a * x[i]; return *this; }
  void Neg() { for (int i = 0; i < n_; i++) p_[i] = -p_[i]; }
  double Norml2() const { return std::sqrt((*this) * (*this)); }
  double Normlinf() const { double m = 0; for (int i = 0; i < n_; i++) m = std::fmax(m, std::fabs(p_[i])); return m; }
  double Sum() const { double s = 0; for (int i = 0; i < n_; i++) s += p_[i]; return s; }
  double Min() const { double m = p_[0]; for (int i = 0; i < n_; i++) m = std::fmin(m, p_[i]); return m; }
  double Max() const { double m = p_[0]; for (int i = 0; i < n_; i++) m = std::fmax(m, p_[i]); return m; }
  double DistanceTo(const double *q) const { double s = 0; for (int i = 0; i < n_; i++) s += (p_[i] - q[i]) * (p_[i] - q[i]); return std::sqrt(s); }
  void Print(std::ostream &os = std::cout, int w = 8) const { for (int i = 0; i < n_; i++) os << p_[i] << ((i + 1) % w && i + 1 < n_ ? " " : "\n"); }
};

inline void add(const Vector &a, const Vector &b, Vector &c) { c.SetSize(a.Size()); for (int i = 0; i < a.Size(); i++) c[i] = a[i] + b[i]; }
inline void add(const Vector &a, double s, const Vector &b, Vector &c) { c.SetSize(a.Size()); for (int i = 0; i < a.Size(); i++) c[i] = a[i] + s * b[i]; }
inline void add(double a, const Vector &x, double b, const Vector &y, Vector &z) { z.SetSize(x.Size()); for (int i = 0; i < x.Size(); i++) z[i] = a * x[i] + b * y[i]; }
inline void subtract(const Vector &a, const Vector &b, Vector &c) { c.SetSize(a.Size()); for (int i = 0; i < a.Size(); i++) c[i] = a[i] - b[i]; }

class DenseMatrix {
  int h_ = 0, w_ = 0;
  std::vector<double> d_;   // column major like MFEM
public:
  DenseMatrix() {}
  explicit DenseMatrix(int n) : h_(n), w_(n), d_((size_t)n * n, 0.0) {}
  DenseMatrix(int h, int w) : h_(h), w_(w), d_((size_t)h * w, 0.0) {}
  int Height() const { return h_; }
  int Width() const { return w_; }
  int Size() const { return w_; }
  void SetSize(int n) { SetSize(n, n); }
  void SetSize(int h, int w) { h_ = h; w_ = w; d_.assign((size_t)h * w, 0.0); }
  double &operator()(int i, int j) { return d_[(size_t)j * h_ + i]; }
  const double &operator()(int i, int j) const { return d_[(size_t)j * h_ + i]; }
  double &Elem(int i, int j) { return (*this)(i, j); }
  double *GetData() { return d_.data(); }
  const double *GetData() const { return d_.data(); }
  DenseMatrix &operator=(double v) { for (auto &x : d_) x = v; return *this; }
  DenseMatrix &operator*=(double v) { for (auto &x : d_) x *= v; return *this; }
  void Mult(const Vector &x, Vector &y) const {
    y.SetSize(h_);
    for (int i = 0; i < h_; i++) { double s = 0; for (int j = 0; j < w_; j++) s += (*this)(i, j) * x[j]; y[i] = s; }
  }
  void Mult(const double *x, double *y) const {
    for (int i = 0; i < h_; i++) { double s = 0; for (int j = 0; j < w_; j++) s += (*this)(i, j) * x[j]; y[i] = s; }
  }
  void MultTranspose(const Vector &x, Vector &y) const {
    y.SetSize(w_);
    for (int j = 0; j < w_; j++) { double s = 0; for (int i = 0; i < h_; i++) s += (*this)(i, j) * x[i]; y[j] = s; }
  }
  double Det() const {
    if (h_ == 1) return d_[0];
    if (h_ == 2) return (*this)(0, 0) * (*this)(1, 1) - (*this)(0, 1) * (*this)(1, 0);
    const DenseMatrix &A = *this;
    return A(0, 0) * (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)) - A(0, 1) * (A(1, 0) * A(2, 2) - A(1, 2) * A(2, 0)) +
           A(0, 2) * (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0));
  }
  void Transpose() { DenseMatrix t(w_, h_); for (int i = 0; i < h_; i++) for (int j = 0; j < w_; j++) t(j, i) = (*this)(i, j); *this = t; }
  void Print(std::ostream &os = std::cout, int w = 4) const { for (int i = 0; i < h_; i++) { for (int j = 0; j < w_; j++) os << (*this)(i, j) << " "; os << "\n"; } }
};

class IntegrationPoint {
public:
  double x = 0, y = 0, z = 0, weight = 0;
  void Set(const double *p, int dim) { x = p[0]; if (dim > 1) y = p[1]; if (dim > 2) z = p[2]; }
  void Set3(double a, double b, double c) { x = a; y = b; z = c; }
};

class ElementTransformation {
public:
  int Attribute = 1, ElementNo = 0;
  virtual ~ElementTransformation() {}
  virtual void Transform(const IntegrationPoint &ip, Vector &x) { x.SetSize(3); x[0] = ip.x; x[1] = ip.y; x[2] = ip.z; }
  void SetIntPoint(const IntegrationPoint *) {}
  int GetSpaceDim() const { return 3; }
};

class Coefficient {
public:
  virtual ~Coefficient() {}
  virtual double Eval(ElementTransformation &T, const IntegrationPoint &ip) = 0;
};

class VectorCoefficient {
protected:
  int vdim;
public:
  explicit VectorCoefficient(int vd) : vdim(vd) {}
  virtual ~VectorCoefficient() {}
  int GetVDim() const { return vdim; }
  virtual void Eval(Vector &V, ElementTransformation &T, const IntegrationPoint &ip) = 0;
};

class Geometry {
public:
  enum Type { POINT, SEGMENT, TRIANGLE, SQUARE, TETRAHEDRON, CUBE, PRISM, PYRAMID };
  static int NumVerts(int t) { static const int nv[] = {1, 2, 3, 4, 4, 8, 6, 5}; return nv[t]; }
};

class Element {
public:
  enum Type { POINT, SEGMENT, TRIANGLE, QUADRILATERAL, TETRAHEDRON, HEXAHEDRON, WEDGE, PYRAMID };
  Geometry::Type geom;
  int attr = 1;
  std::vector<int> v;
  Element(Geometry::Type g, const int *ind, int a) : geom(g), attr(a), v(ind, ind + Geometry::NumVerts(g)) {}
  int GetNVertices() const { return (int)v.size(); }
  const int *GetVertices() const { return v.data(); }
  int *GetVertices() { return v.data(); }
  void GetVertices(Array<int> &a) const { a.SetSize((int)v.size()); for (size_t i = 0; i < v.size(); i++) a[(int)i] = v[i]; }
  void SetVertices(const int *ind) { for (size_t i = 0; i < v.size(); i++) v[i] = ind[i]; }
  int GetAttribute() const { return attr; }
  void SetAttribute(int a) { attr = a; }
  Geometry::Type GetGeometryType() const { return geom; }
  int GetType() const { static const int t[] = {0, 1, 2, 3, 4, 5, 6, 7}; return t[geom]; }
};

// Records what the table builders put in; geometry / topology queries beyond that are not provided.
class Mesh {
public:
  int dim_ = 3, sdim_ = 3;
  std::vector<std::vector<double>> verts;
  std::vector<Element *> elems, bdr;
  int n_refine = 0;
  Mesh() {}
  Mesh(int dim, int nv, int ne, int nbe = 0, int sdim = -1) : dim_(dim), sdim_(sdim < 0 ? dim : sdim) { (void)nv; (void)ne; (void)nbe; }
  Mesh(double *vertices, int num_vertices, int *element_indices, Geometry::Type element_type, int *element_attributes,
       int num_elements, int *boundary_indices, Geometry::Type boundary_type, int *boundary_attributes,
       int num_boundary_elements, int dimension, int space_dimension = -1)
      : dim_(dimension), sdim_(space_dimension < 0 ? dimension : space_dimension) {
    for (int i = 0; i < num_vertices; i++) verts.emplace_back(vertices + (size_t)i * sdim_, vertices + (size_t)(i + 1) * sdim_);
    const int nve = Geometry::NumVerts(element_type), nvb = Geometry::NumVerts(boundary_type);
    for (int e = 0; e < num_elements; e++) elems.push_back(new Element(element_type, element_indices + (size_t)e * nve, element_attributes[e]));
    for (int b = 0; b < num_boundary_elements; b++) bdr.push_back(new Element(boundary_type, boundary_indices + (size_t)b * nvb, boundary_attributes[b]));
    if (bdr.empty()) GenerateBoundaryElements();
  }
  Mesh(const Mesh &o, bool = true) : dim_(o.dim_), sdim_(o.sdim_), verts(o.verts), n_refine(o.n_refine) {
    for (auto *e : o.elems) elems.push_back(new Element(*e));
    for (auto *e : o.bdr) bdr.push_back(new Element(*e));
  }
  virtual ~Mesh() { for (auto *e : elems) delete e; for (auto *e : bdr) delete e; }
  int Dimension() const { return dim_; }
  int SpaceDimension() const { return sdim_; }
  int GetNV() const { return (int)verts.size(); }
  int GetNE() const { return (int)elems.size(); }
  int GetNBE() const { return (int)bdr.size(); }
  double *GetVertex(int i) { return verts[i].data(); }
  const double *GetVertex(int i) const { return verts[i].data(); }
  Element *GetElement(int i) { return elems[i]; }
  const Element *GetElement(int i) const { return elems[i]; }
  Element *GetBdrElement(int i) { return bdr[i]; }
  const Element *GetBdrElement(int i) const { return bdr[i]; }
  void GetElementVertices(int i, Array<int> &v) const { elems[i]->GetVertices(v); }
  void GetBdrElementVertices(int i, Array<int> &v) const { bdr[i]->GetVertices(v); }
  void SetCurvature(int, bool = false, int = -1, int = 1) {}
  int GetAttribute(int i) const { return elems[i]->GetAttribute(); }
  void SetAttribute(int i, int a) { elems[i]->SetAttribute(a); }
  Element *NewElement(int geom_or_type) {
    static const int zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    return new Element((Geometry::Type)geom_or_type, zero, 1);
  }
  void AddElement(Element *e) { elems.push_back(e); }
  void AddBdrElement(Element *e) { bdr.push_back(e); }
  void AddVertex(const double *x) { verts.emplace_back(x, x + sdim_); }
  void AddVertex(double x, double y = 0.0, double z = 0.0) { double p[3] = {x, y, z}; verts.emplace_back(p, p + sdim_); }
  void add_(std::vector<Element *> &to, Geometry::Type g, const int *vi, int a) { to.push_back(new Element(g, vi, a)); }
  void AddSegment(const int *vi, int a = 1) { add_(elems, Geometry::SEGMENT, vi, a); }
  void AddSegment(int v0, int v1, int a = 1) { int vi[2] = {v0, v1}; add_(elems, Geometry::SEGMENT, vi, a); }
  void AddTri(const int *vi, int a = 1) { add_(elems, Geometry::TRIANGLE, vi, a); }
  void AddTriangle(const int *vi, int a = 1) { add_(elems, Geometry::TRIANGLE, vi, a); }
  void AddQuad(const int *vi, int a = 1) { add_(elems, Geometry::SQUARE, vi, a); }
  void AddTet(const int *vi, int a = 1) { add_(elems, Geometry::TETRAHEDRON, vi, a); }
  void AddHex(const int *vi, int a = 1) { add_(elems, Geometry::CUBE, vi, a); }
  void AddWedge(const int *vi, int a = 1) { add_(elems, Geometry::PRISM, vi, a); }
  void AddPyramid(const int *vi, int a = 1) { add_(elems, Geometry::PYRAMID, vi, a); }
  void AddBdrPoint(int v, int a = 1) { add_(bdr, Geometry::POINT, &v, a); }
  void AddBdrSegment(const int *vi, int a = 1) { add_(bdr, Geometry::SEGMENT, vi, a); }
  void AddBdrTriangle(const int *vi, int a = 1) { add_(bdr, Geometry::TRIANGLE, vi, a); }
  void AddBdrQuad(const int *vi, int a = 1) { add_(bdr, Geometry::SQUARE, vi, a); }
  void FinalizeTriMesh(int = 0, int = 0, bool = false) { if (bdr.empty()) GenerateBoundaryElements(); }
  void FinalizeQuadMesh(int = 0, int = 0, bool = false) { if (bdr.empty()) GenerateBoundaryElements(); }
  void FinalizeTetMesh(int = 0, int = 0, bool = false) { if (bdr.empty()) GenerateBoundaryElements(); }
  void FinalizeHexMesh(int = 0, int = 0, bool = false) { if (bdr.empty()) GenerateBoundaryElements(); }
  void FinalizeWedgeMesh(int = 0, int = 0, bool = false) { if (bdr.empty()) GenerateBoundaryElements(); }
  void FinalizeMesh(int = 0, bool = false) { if (bdr.empty()) GenerateBoundaryElements(); }
  void FinalizeTopology(bool = true) {}
  void Finalize(bool = false, bool = false) { if (bdr.empty()) GenerateBoundaryElements(); }
  void UniformRefinement(int = 0) { n_refine++; }
  void RemoveUnusedVertices() {}
  void RemoveInternalBoundaries() {}
  // faces that belong to exactly one element (what MFEM does when no boundary is given)
  void GenerateBoundaryElements() {
    static const int hexf[6][4] = {{3, 2, 1, 0}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}};
    static const int tetf[4][3] = {{1, 2, 3}, {0, 3, 2}, {0, 1, 3}, {0, 2, 1}};
    static const int prif[5][4] = {{0, 2, 1, -1}, {3, 4, 5, -1}, {0, 1, 4, 3}, {1, 2, 5, 4}, {2, 0, 3, 5}};
    static const int quae[4][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}};
    static const int trie[3][2] = {{0, 1}, {1, 2}, {2, 0}};
    std::map<std::vector<int>, std::pair<int, std::vector<int>>> faces;
    auto add_face = [&](const Element *e, const int *loc, int n) {
      std::vector<int> f;
      for (int k = 0; k < n; k++) if (loc[k] >= 0) f.push_back(e->v[loc[k]]);
      std::vector<int> key = f;
      std::sort(key.begin(), key.end());
      auto &slot = faces[key];
      slot.first++;
      slot.second = f;
    };
    for (const Element *e : elems) {
      switch (e->geom) {
        case Geometry::CUBE: for (auto &f : hexf) add_face(e, f, 4); break;
        case Geometry::TETRAHEDRON: for (auto &f : tetf) add_face(e, f, 3); break;
        case Geometry::PRISM: for (auto &f : prif) add_face(e, f, 4); break;
        case Geometry::SQUARE: for (auto &f : quae) add_face(e, f, 2); break;
        case Geometry::TRIANGLE: for (auto &f : trie) add_face(e, f, 2); break;
        default: break;
      }
    }
    for (auto *b : bdr) delete b;
    bdr.clear();
    for (auto &kv : faces)
      if (kv.second.first == 1) {
        const std::vector<int> &f = kv.second.second;
        const Geometry::Type g = f.size() == 4 ? Geometry::SQUARE : (f.size() == 3 ? Geometry::TRIANGLE : Geometry::SEGMENT);
        bdr.push_back(new Element(g, f.data(), 1));
      }
  }
  void CheckElementOrientation(bool = true) {}
  void CheckBdrElementOrientation(bool = true) {}
  void SetAttributes() {}
  void EnsureNodes() {}
  int EulerNumber() const { return 0; }
  int EulerNumber2D() const { return 0; }
  void GetBoundingBox(Vector &mn, Vector &mx, int = 2) {
    mn.SetSize(sdim_); mx.SetSize(sdim_);
    for (int d = 0; d < sdim_; d++) { mn[d] = 1e300; mx[d] = -1e300; }
    for (auto &v : verts) for (int d = 0; d < sdim_; d++) { mn[d] = std::fmin(mn[d], v[d]); mx[d] = std::fmax(mx[d], v[d]); }
  }
  virtual void Print(std::ostream & = std::cout) const {}
};

// ---- opaque stand-ins for the parallel FE types named by the (off-path) Fourier-series classes of lib/bravais ----
class ParFiniteElementSpace { public: virtual ~ParFiniteElementSpace() {} };
class LinearFormIntegrator { public: virtual ~LinearFormIntegrator() {} };
class DomainLFIntegrator : public LinearFormIntegrator { public: explicit DomainLFIntegrator(Coefficient &) {} };
class VectorFunctionCoefficient : public VectorCoefficient {
  void (*f_)(const Vector &, Vector &) = nullptr;
public:
  VectorFunctionCoefficient(int vd, void (*f)(const Vector &, Vector &), Coefficient * = nullptr) : VectorCoefficient(vd), f_(f) {}
  template <class F> VectorFunctionCoefficient(int vd, F, Coefficient * = nullptr) : VectorCoefficient(vd) {}
  void Eval(Vector &V, ElementTransformation &T, const IntegrationPoint &ip) override { Vector x; T.Transform(ip, x); V.SetSize(vdim); if (f_) f_(x, V); }
};
class VectorConstantCoefficient : public VectorCoefficient {
  Vector v_;
public:
  explicit VectorConstantCoefficient(const Vector &v) : VectorCoefficient(v.Size()), v_(v) {}
  void Eval(Vector &V, ElementTransformation &, const IntegrationPoint &) override { V = v_; }
};
class VectorFEDomainLFIntegrator : public LinearFormIntegrator { public: explicit VectorFEDomainLFIntegrator(VectorCoefficient &) {} };
class HypreParVector : public Vector { public: HypreParVector() {} explicit HypreParVector(ParFiniteElementSpace *) {} };
inline double InnerProduct(const HypreParVector &a, const HypreParVector &b) { return a.Size() == b.Size() ? a * b : 0.0; }
class ParLinearForm : public Vector {
public:
  explicit ParLinearForm(ParFiniteElementSpace *) {}
  ~ParLinearForm() {}
  void AddDomainIntegrator(LinearFormIntegrator *i) { delete i; }
  void Assemble() {}
  void ParallelAssemble(HypreParVector &) {}
  HypreParVector *ParallelAssemble() { return new HypreParVector; }
  ParLinearForm &operator=(double) { return *this; }
};

}  // namespace mfem
