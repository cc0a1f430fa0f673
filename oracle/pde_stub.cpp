// pde_stub.cpp -- stands in for the reference's bindings/pde_bindings.cpp, which
// needs Eigen (absent in this image).  TEST INFRASTRUCTURE ONLY: lets
// oracle/Makefile's `pyref` target link the reference's own pybind11 module so
// tests/golden/make_golden.py can drive the reference's Python calibrator.
#include <pybind11/pybind11.h>
void init_pde_bindings(pybind11::module_&) {}
