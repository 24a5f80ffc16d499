// Binding used ONLY by make_torch_impl_golden.py: exposes the reference's own torch oracle
// (/root/reference/tests/torch_impl.cpp, compiled from where it lies, never copied) to Python so
// golden vectors for the SH and tile-intersection rows can be generated in the build container.
#include <torch/library.h>

#include "torch_impl.hpp"

namespace {
at::Tensor sh(int64_t degree, const at::Tensor &dirs, const at::Tensor &coeffs) {
    return reference::spherical_harmonics((int)degree, dirs, coeffs);
}
std::tuple<at::Tensor, at::Tensor, at::Tensor> isect(const at::Tensor &means2d, const at::Tensor &radii,
                                                     const at::Tensor &depths, int64_t tile_size, int64_t tile_width,
                                                     int64_t tile_height, bool sort) {
    return reference::isect_tiles(means2d, radii, depths, (int)tile_size, (int)tile_width, (int)tile_height, sort);
}
at::Tensor q2r(const at::Tensor &quats) { return reference::quat_to_rotmat(quats); }
} // namespace

TORCH_LIBRARY(ref_torch_impl, m) {
    m.def("spherical_harmonics", &sh);
    m.def("isect_tiles", &isect);
    m.def("quat_to_rotmat", &q2r);
}
