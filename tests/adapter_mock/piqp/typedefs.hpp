// tests/adapter_mock -- a MINIMAL MOCK of the four reference types include/piqp_b200_adapter.hpp touches, so that the adapter can be
// compiled and exercised in an image without Eigen.  Test infrastructure only: it mirrors the accessor API the adapter relies on
// (Vec<T>::data(), SparseMat::outerIndexPtr() / innerIndexPtr() / valuePtr(), Data::n/p/m/P_utri/AT/GT) and nothing else of
// include/piqp/typedefs.hpp, dense/data.hpp:23-51, sparse/data.hpp:26-54, kkt_solver_base.hpp:21-44.
#pragma once
#include <cstdio>
#include <vector>
#define piqp_eprint(...) std::fprintf(stderr, __VA_ARGS__)
namespace piqp {
enum { PIQP_DENSE = 0, PIQP_SPARSE = 1 };
template<typename T> struct Vec {
    std::vector<T> v;
    Vec() = default;
    explicit Vec(size_t n) : v(n) {}
    T* data() { return v.data(); }
    const T* data() const { return v.data(); }
    size_t size() const { return v.size(); }
    T& operator()(size_t i) { return v[i]; }
    const T& operator()(size_t i) const { return v[i]; }
};
template<typename T> struct Mat {      // column-major dense
    int r = 0, c = 0; std::vector<T> v;
    Mat() = default;
    Mat(int r_, int c_) : r(r_), c(c_), v((size_t)r_ * c_) {}
    T* data() { return v.data(); }
    const T* data() const { return v.data(); }
    T& operator()(int i, int j) { return v[(size_t)j * r + i]; }
    const T& operator()(int i, int j) const { return v[(size_t)j * r + i]; }
};
template<typename T, typename I> struct SparseMat {      // CSC
    int r = 0, c = 0; std::vector<I> outer, inner; std::vector<T> val;
    const I* outerIndexPtr() const { return outer.data(); }
    const I* innerIndexPtr() const { return inner.data(); }
    const T* valuePtr() const { return val.data(); }
};
}  // namespace piqp
