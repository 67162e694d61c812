// Test infrastructure only.  The reference's voxel_pooling extension (seg3d/ops/voxel_pooling/src/voxel_pooling.cpp)
// binds its CPU functions and its CUDA functions in one pybind module; oracle/build_ref.py compiles that .cpp where it
// lies under /root/reference for its CPU functions only, and these stubs satisfy the two CUDA symbols it references
// (voxel_pooling.h:17-22) without compiling voxel_pooling_cuda.cu (GPU-only, never called by the oracle checks).
#include <stdexcept>

#include <torch/torch.h>

at::Tensor voxel_pooling_forward_cuda(const at::Tensor, const at::Tensor, const at::Tensor) {
  throw std::runtime_error("oracle/_ref: the reference's CUDA pooling kernel is not part of the CPU reference build");
}

at::Tensor voxel_pooling_backward_cuda(const at::Tensor, const at::Tensor, const at::Tensor, const int) {
  throw std::runtime_error("oracle/_ref: the reference's CUDA pooling kernel is not part of the CPU reference build");
}
