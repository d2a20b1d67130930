#include <pybind11/pybind11.h>
namespace py = pybind11;
void bind_mlp(py::module& m) { (void)m; }
