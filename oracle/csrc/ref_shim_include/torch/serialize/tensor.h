// TEST INFRASTRUCTURE.  Stand-in for <torch/serialize/tensor.h> when the reference's stand-alone CUDA sources
// (pcdet/ops/pointnet2/pointnet2_stack/src/*_gpu.cu) are compiled into oracle/_ref by oracle/ref_build.py: their headers
// only DECLARE wrapper functions taking at::Tensor by value, so an incomplete type is enough and libtorch is not needed.
#pragma once
namespace at {
class Tensor;
}
