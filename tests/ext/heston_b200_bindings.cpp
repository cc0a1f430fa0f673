// tests/ext/heston_b200_bindings.cpp -- the pybind11 binding of INTEGRATION.md section 2, compiled as a
// TEST-ONLY extension (tests/ext/build_ext.py) so that the snippet a maintainer of the reference would add next to
// src/cpp/bindings/heston_bindings.cpp is known to compile against include/heston_b200.h and to work through
// the C ABI.  bind_heston_b200() below is that snippet verbatim; only the PYBIND11_MODULE wrapper at the end is
// test scaffolding (in the reference it is one call from quant_cpp.cpp:135).
#include <pybind11/pybind11.h>
#include <pybind11/numpy.h>

#include <stdexcept>
#include <string>

#include "heston_b200.h"
namespace py = pybind11;
using arr = py::array_t<double, py::array::c_style | py::array::forcecast>;

struct B200Plan {
    hb_plan* p = nullptr;
    B200Plan(const std::string& mode, int n_grid, double eta, double alpha, int device) {
        if (hb_plan_create(mode == "fft" ? HB_MODE_FFT : HB_MODE_REFGRID, n_grid, eta, alpha, device, &p))
            throw std::runtime_error(hb_last_error());
    }
    ~B200Plan() { hb_plan_destroy(p); }
};
static void chk(int rc) {
    if (rc == HB_ERR_INVALID_ARGUMENT || rc == HB_ERR_INVALID_PARAMETER) throw std::invalid_argument(hb_last_error());
    if (rc) throw std::runtime_error(hb_last_error());
}
void bind_heston_b200(py::module_& m) {
    py::class_<B200Plan>(m, "B200Plan")
        .def(py::init<const std::string&, int, double, double, int>(), py::arg("mode") = "refgrid",
             py::arg("n_grid") = 4096, py::arg("eta") = 0.25, py::arg("alpha") = 0.75, py::arg("device") = 0)
        .def("set_surface", [](B200Plan& s, arr K, arr T, py::array_t<uint8_t> c, arr mkt, double S0, double r, double q) {
            chk(hb_surface_set(s.p, (int)K.size(), K.data(), T.data(), c.data(), mkt.size() ? mkt.data() : nullptr, S0, r, q)); })
        .def("objective", [](B200Plan& s, arr X) {          // X: [P,5] -> loss [P]
            arr out(X.shape(0));
            { py::gil_scoped_release g; chk(hb_objective_host(s.p, X.data(), (int)X.shape(0), out.mutable_data())); }
            return out; })
        .def("prices", [](B200Plan& s, arr X) {             // X: [P,5] -> prices [P,n]
            arr out({X.shape(0), (py::ssize_t)hb_plan_n_options(s.p)});
            { py::gil_scoped_release g; chk(hb_price_host(s.p, X.data(), (int)X.shape(0), out.mutable_data())); }
            return out; })
        .def("normal_equations", [](B200Plan& s, arr X) {   // X: [P,5] -> [P,22]
            arr out({X.shape(0), (py::ssize_t)HB_NEQ_WIDTH});
            { py::gil_scoped_release g; chk(hb_normal_eq_host(s.p, X.data(), (int)X.shape(0), out.mutable_data())); }
            return out; })
        .def("implied_vols", [](B200Plan& s, arr X) {       // loop of implied_volatility, heston.cpp:311-349
            arr out({X.shape(0), (py::ssize_t)hb_plan_n_options(s.p)});
            { py::gil_scoped_release g; chk(hb_implied_vol_host(s.p, X.data(), (int)X.shape(0), out.mutable_data())); }
            return out; })
        .def("greeks", [](B200Plan& s, arr X) {             // loop of price_option_with_greeks, heston.cpp:168-217
            arr out({X.shape(0), (py::ssize_t)hb_plan_n_options(s.p), (py::ssize_t)5});  // delta gamma vega theta rho
            { py::gil_scoped_release g; chk(hb_greeks_host(s.p, X.data(), (int)X.shape(0), out.mutable_data())); }
            return out; });
}

PYBIND11_MODULE(quant_cpp_b200, m) {
    m.doc() = "test build of INTEGRATION.md section 2: quant_cpp.heston.B200Plan over libheston_b200.so";
    py::module_ heston = m.def_submodule("heston");
    bind_heston_b200(heston);
}
