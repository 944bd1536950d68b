// Face normals used by the sweep DAG and the system matrices, and the text dumps the reference's driver writes after a run.
// Same signatures as the reference's include/Utils.hpp:16-46.  The normal arithmetic (and its order) is on the parity path:
// the sign of normal . direction decides the upwind DAG bit-for-bit.  The dumps are post-processing output.
#pragma once

#include "AngularQuadrature.hpp"

#include "mfem.hpp"

#include <ostream>
#include <string>
#include <vector>

namespace pbte
{
struct ElementIntegralData;
}

namespace utils
{
void WriteVector(std::ostream &os, const mfem::Vector &v, const std::string &label);
void WriteMatrix(std::ostream &os, const mfem::DenseMatrix &m, const std::string &label);

/// output/log/integrals_all.txt (reference: src/Utils.cpp:48-164).
bool DumpElementIntegrals(const std::vector<pbte::ElementIntegralData> &element_data, bool is_root);
/// output/log/coeff_all.txt: every block coeff[dir][branch][spec] as text (reference: src/Utils.cpp:166-226).
/// The text of a benchmark-size run would be tens to hundreds of GB (33 GB for C5, 344 GB for C4): a dump whose estimate
/// exceeds PBTE_DUMP_LIMIT_MB (default 4096) is refused with a message and `false`.
bool DumpCoefficients(const std::vector<std::vector<std::vector<mfem::DenseMatrix>>> &coeff, const pbte::AngleQuadrature &quad, bool is_root);
/// output/log/Tc_all.txt (reference: src/Utils.cpp:228-260).
bool DumpTemperature(const mfem::DenseMatrix &Tc, bool is_root);

mfem::Vector ComputeFaceNormal(const mfem::Mesh &mesh, int face_id);
mfem::Vector ComputeOutwardFaceNormal(const mfem::Mesh &mesh, int face_id, int elem_id);
}  // namespace utils
