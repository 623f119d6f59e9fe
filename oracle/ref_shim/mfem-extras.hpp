// Stand-in for the reference's absent "mfem-extras.hpp" (see mfem.hpp in this directory): only the four FE-space type
// names that the off-path Fourier-series classes of lib/bravais mention.
#pragma once
#include "mfem.hpp"
namespace mfem { namespace miniapps {
class H1_ParFESpace : public ParFiniteElementSpace {};
class L2_ParFESpace : public ParFiniteElementSpace {};
class ND_ParFESpace : public ParFiniteElementSpace {};
class RT_ParFESpace : public ParFiniteElementSpace {};
} }
