// oracle/oracle_multistage.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// placeholder: filled in by the multistage milestone.
#pragma once
#include "oracle_sparse.hpp"
namespace oracle {
inline std::unique_ptr<KKTBackend> make_multistage_backend(const SparseMatrices&) { return nullptr; }
}
