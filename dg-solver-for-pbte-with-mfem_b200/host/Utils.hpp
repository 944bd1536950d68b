// Face normals used by the sweep DAG and the system matrices.  Same signatures as the reference's
// include/Utils.hpp:41-46.  The arithmetic (and its order) is on the parity path: the sign of
// normal . direction decides the upwind DAG bit-for-bit.
#pragma once

#include "mfem.hpp"

namespace utils
{
mfem::Vector ComputeFaceNormal(const mfem::Mesh &mesh, int face_id);
mfem::Vector ComputeOutwardFaceNormal(const mfem::Mesh &mesh, int face_id, int elem_id);
}  // namespace utils
