// oracle/ref_shim/torch/extension.h -- TEST INFRASTRUCTURE ONLY.
// A minimal stand-in for <torch/extension.h> so that the reference's own
// lidargen/ops/roiaware_pool3d/src/roiaware_pool3d.cpp compiles UNMODIFIED, from where it lies under
// /root/reference, with plain g++ (its torch idioms -- tensor.data<T>() -- were removed from torch 2.11, so the
// real headers no longer build it).  Only what that file touches is provided: at::Tensor::size()/data<T>() over a
// caller-owned buffer and a no-op PYBIND11_MODULE.
#pragma once
#include <math.h>
#include <cmath>
#include <stdint.h>

namespace at {
struct Tensor {
    void* ptr = nullptr;
    int64_t dims[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int64_t size(int i) const { return dims[i]; }
    template <typename T> T* data() const { return static_cast<T*>(ptr); }
};
}  // namespace at

namespace ref_shim {
struct Module {
    template <typename Fn> void def(const char*, Fn, const char*) {}
};
}  // namespace ref_shim

#define TORCH_EXTENSION_NAME ref_shim_module
#define PYBIND11_MODULE(name, m) void ref_shim_register_##name(ref_shim::Module& m)
