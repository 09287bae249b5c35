#pragma once
#include "piqp/typedefs.hpp"
namespace piqp { namespace sparse {
template<typename T, typename I> struct Data { long n = 0, p = 0, m = 0; SparseMat<T, I> P_utri, AT, GT; };
}}
