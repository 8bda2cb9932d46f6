#include <torch/extension.h>
void register_comm(pybind11::module_& m) {}
