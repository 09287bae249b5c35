// integration/piqp_python_batched.cpp -- REFERENCE-SIDE Python binding of the batched path: what a maintainer adds to
// interfaces/python/src/piqp_python.cpp (:130-188 bind piqp::DenseSolver / SparseSolver) so that `import piqp` also offers
//     piqp.DenseSolverBatched().setup(P, c, A, b, G, h_l, h_u, x_l, x_u)  /  .update(...)  /  .solve()  /  .result
// with a leading batch axis on every array.  pybind11 + numpy only (no Eigen): the arrays go straight to the C-ABI of
// include/piqp_b200.h.  Builds as its own module here (tests/test_adapter_header.py: `piqp_batched`), or is pasted into the
// reference's PYBIND11_MODULE body.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "piqp_b200.h"

namespace py = pybind11;
using arr = py::array_t<double, py::array::c_style | py::array::forcecast>;
using opt_arr = py::object;      // None or array

namespace {
const double* ptr(const opt_arr& o, std::vector<arr>& keep, py::ssize_t expect, const char* name) {
    if (o.is_none()) return nullptr;
    arr a = arr::ensure(o);
    if (!a || a.size() != expect) throw std::invalid_argument(std::string(name) + " has wrong dimensions");
    keep.push_back(a);
    return a.data();
}
void check(int rc, const char* what) { if (rc < 0) throw std::runtime_error(std::string(what) + ": " + b200_last_error()); }

struct DenseSolverBatched {
    b200qp_handle* h = nullptr;
    b200qp_settings settings;
    int batch = 0, n = 0, p = 0, m = 0;
    DenseSolverBatched() { b200qp_set_default_settings_dense(&settings); }
    ~DenseSolverBatched() { if (h) b200qp_cleanup(h); }
    void setup(const arr& P, const arr& c, opt_arr A, opt_arr b, opt_arr G, opt_arr h_l, opt_arr h_u, opt_arr x_l, opt_arr x_u) {
        if (P.ndim() != 3 || P.shape(1) != P.shape(2)) throw std::invalid_argument("P must be (batch, n, n)");
        batch = (int)P.shape(0); n = (int)P.shape(1);
        p = A.is_none() ? 0 : (int)arr::ensure(A).shape(1);
        m = G.is_none() ? 0 : (int)arr::ensure(G).shape(1);
        std::vector<arr> keep;
        const py::ssize_t B = batch;
        if (h) { b200qp_cleanup(h); h = nullptr; }
        check(b200qp_setup_dense(&h, batch, n, p, m, P.data(), ptr(c, keep, B * n, "c"), ptr(A, keep, B * p * n, "A"), ptr(b, keep, B * p, "b"),
                                 ptr(G, keep, B * m * n, "G"), ptr(h_l, keep, B * m, "h_l"), ptr(h_u, keep, B * m, "h_u"), ptr(x_l, keep, B * n, "x_l"),
                                 ptr(x_u, keep, B * n, "x_u"), &settings, 0, 0), "b200qp_setup_dense");
    }
    void update(opt_arr P, opt_arr c, opt_arr A, opt_arr b, opt_arr G, opt_arr h_l, opt_arr h_u, opt_arr x_l, opt_arr x_u) {
        if (!h) throw std::runtime_error("Solver not setup yet");
        std::vector<arr> keep;
        const py::ssize_t B = batch;
        check(b200qp_update_settings(h, &settings), "b200qp_update_settings");
        check(b200qp_update_dense(h, ptr(P, keep, B * n * n, "P"), ptr(c, keep, B * n, "c"), ptr(A, keep, B * p * n, "A"), ptr(b, keep, B * p, "b"),
                                  ptr(G, keep, B * m * n, "G"), ptr(h_l, keep, B * m, "h_l"), ptr(h_u, keep, B * m, "h_u"), ptr(x_l, keep, B * n, "x_l"),
                                  ptr(x_u, keep, B * n, "x_u"), 0), "b200qp_update_dense");
    }
    std::vector<int> solve() {
        if (!h) throw std::runtime_error("Solver not setup yet");
        { py::gil_scoped_release nogil; check(b200qp_solve(h), "b200qp_solve"); }
        std::vector<b200qp_info> infos(batch);
        check(b200qp_get_info(h, infos.data()), "b200qp_get_info");
        std::vector<int> st(batch);
        for (int k = 0; k < batch; k++) st[k] = infos[k].status;
        return st;
    }
    py::dict result() {
        if (!h) throw std::runtime_error("Solver not setup yet");
        const py::ssize_t B = batch;
        arr x({B, (py::ssize_t)n}), y({B, (py::ssize_t)p}), z_l({B, (py::ssize_t)m}), z_u({B, (py::ssize_t)m}), z_bl({B, (py::ssize_t)n}), z_bu({B, (py::ssize_t)n}),
            s_l({B, (py::ssize_t)m}), s_u({B, (py::ssize_t)m}), s_bl({B, (py::ssize_t)n}), s_bu({B, (py::ssize_t)n});
        check(b200qp_get_result(h, x.mutable_data(), y.mutable_data(), z_l.mutable_data(), z_u.mutable_data(), z_bl.mutable_data(), z_bu.mutable_data(),
                                s_l.mutable_data(), s_u.mutable_data(), s_bl.mutable_data(), s_bu.mutable_data(), 0), "b200qp_get_result");
        std::vector<b200qp_info> infos(batch);
        check(b200qp_get_info(h, infos.data()), "b200qp_get_info");
        py::list iters, objs;
        for (const auto& i : infos) { iters.append(i.iter); objs.append(i.primal_obj); }
        py::dict d;
        d["x"] = x; d["y"] = y; d["z_l"] = z_l; d["z_u"] = z_u; d["z_bl"] = z_bl; d["z_bu"] = z_bu; d["s_l"] = s_l; d["s_u"] = s_u; d["s_bl"] = s_bl; d["s_bu"] = s_bu;
        d["iter"] = iters; d["primal_obj"] = objs;
        return d;
    }
};
}  // namespace

PYBIND11_MODULE(piqp_batched, mod) {
    mod.doc() = "batched twin of piqp.DenseSolver over libpiqp_b200 (device-resident interior-point loop)";
    py::class_<b200qp_settings>(mod, "Settings")      // the fields piqp_python.cpp:100-128 exposes, same names
        .def_readwrite("rho_init", &b200qp_settings::rho_init).def_readwrite("delta_init", &b200qp_settings::delta_init)
        .def_readwrite("eps_abs", &b200qp_settings::eps_abs).def_readwrite("eps_rel", &b200qp_settings::eps_rel)
        .def_readwrite("check_duality_gap", &b200qp_settings::check_duality_gap)
        .def_readwrite("eps_duality_gap_abs", &b200qp_settings::eps_duality_gap_abs).def_readwrite("eps_duality_gap_rel", &b200qp_settings::eps_duality_gap_rel)
        .def_readwrite("infeasibility_threshold", &b200qp_settings::infeasibility_threshold)
        .def_readwrite("reg_lower_limit", &b200qp_settings::reg_lower_limit).def_readwrite("reg_finetune_lower_limit", &b200qp_settings::reg_finetune_lower_limit)
        .def_readwrite("max_iter", &b200qp_settings::max_iter).def_readwrite("max_factor_retires", &b200qp_settings::max_factor_retires)
        .def_readwrite("preconditioner_scale_cost", &b200qp_settings::preconditioner_scale_cost)
        .def_readwrite("preconditioner_reuse_on_update", &b200qp_settings::preconditioner_reuse_on_update)
        .def_readwrite("preconditioner_iter", &b200qp_settings::preconditioner_iter).def_readwrite("tau", &b200qp_settings::tau)
        .def_readwrite("iterative_refinement_always_enabled", &b200qp_settings::iterative_refinement_always_enabled)
        .def_readwrite("verbose", &b200qp_settings::verbose).def_readwrite("compute_timings", &b200qp_settings::compute_timings);
    py::class_<DenseSolverBatched>(mod, "DenseSolverBatched")
        .def(py::init<>())
        .def_readwrite("settings", &DenseSolverBatched::settings)
        .def("setup", &DenseSolverBatched::setup, py::arg("P"), py::arg("c"), py::arg("A") = py::none(), py::arg("b") = py::none(), py::arg("G") = py::none(),
             py::arg("h_l") = py::none(), py::arg("h_u") = py::none(), py::arg("x_l") = py::none(), py::arg("x_u") = py::none())
        .def("update", &DenseSolverBatched::update, py::arg("P") = py::none(), py::arg("c") = py::none(), py::arg("A") = py::none(), py::arg("b") = py::none(),
             py::arg("G") = py::none(), py::arg("h_l") = py::none(), py::arg("h_u") = py::none(), py::arg("x_l") = py::none(), py::arg("x_u") = py::none())
        .def("solve", &DenseSolverBatched::solve)
        .def_property_readonly("result", &DenseSolverBatched::result);
}
