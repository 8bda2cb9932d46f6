#include <torch/extension.h>
void register_conv(pybind11::module_& m) {}
