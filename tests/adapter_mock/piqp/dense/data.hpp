#pragma once
#include "piqp/typedefs.hpp"
namespace piqp { namespace dense {
template<typename T> struct Data { long n = 0, p = 0, m = 0; Mat<T> P_utri, AT, GT; };
}}
