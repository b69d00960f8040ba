// Stand-in for <torch/serialize/tensor.h>: the reference's *_gpu.h headers only need the
// name at::Tensor to DECLARE their (unused here) wrapper functions.  Keeps the legacy build
// free of libtorch and takes 1 s instead of 25 s per file.
#pragma once
namespace at { class Tensor {}; }
