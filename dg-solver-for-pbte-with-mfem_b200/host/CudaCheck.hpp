// Maps C-ABI status codes back to the exception types the reference's classes raise.
#pragma once

#include "pbte_cuda.h"

#include "mfem.hpp"

#include <stdexcept>
#include <string>

namespace pbte
{
inline void ThrowOnStatus(int status, const char *what)
{
    if (status == PBTE_OK) return;
    const std::string msg = std::string(what) + ": " + pbte_cuda_last_error();
    switch (status)
    {
        case PBTE_ERR_SWEEP_STALL: throw std::runtime_error("Angular sweep ordering stalled; check mesh connectivity.");
        case PBTE_ERR_MISSING_BC: throw std::out_of_range(msg);
        case PBTE_ERR_INVALID:
        case PBTE_ERR_ZERO_NORM:
        case PBTE_ERR_STATE: mfem::mfem_error(msg.c_str());
        default: throw std::runtime_error(msg);
    }
}
}  // namespace pbte
