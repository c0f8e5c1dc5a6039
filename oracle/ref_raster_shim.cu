// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// O-gpu oracle: compiles the reference's OWN rasteriser kernels, unmodified and
// in place, by #including the reference translation unit
//   third_party/neural_renderer/neural_renderer/cuda/rasterize_cuda_kernel.cu
// (path injected at build time with -DJAF_REF_KERNEL_FILE=..., see
// oracle/build_ref.sh; nothing is copied into this repository) and exposes the
// two forward kernels
//   forward_face_index_map_cuda_kernel_1  (rasterize_cuda_kernel.cu:24-67)
//   forward_face_index_map_cuda_kernel_2  (rasterize_cuda_kernel.cu:70-169)
// behind a raw-pointer C entry point that launches them exactly as
// forward_face_index_map_cuda does (rasterize_cuda_kernel.cu:596-651: 512
// threads, ceil-div grids, legacy default stream).
//
// The reference file targets the torch-1.2 ATen API.  Instead of editing it we
// neutralise the one macro that no longer compiles: AT_DISPATCH_FLOATING_TYPES
// is redefined to instantiate the float path only (the hot path is float32,
// SURVEY.md §8b "Dtypes"), so `x.type()` is never evaluated.
//
// Output goes to oracle/_ref/libjaf_ref_raster.so (git-ignored, travels with
// gpurun).  Only tests/, __graft_entry__.smoke() and bench.py may load it.
#include <ATen/ATen.h>
#include <ATen/Dispatch.h>
#include <cuda_runtime.h>
#include <cstdint>

#undef AT_DISPATCH_FLOATING_TYPES
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...) \
  [&] { using scalar_t = float; return __VA_ARGS__(); }()

#ifndef JAF_REF_KERNEL_FILE
#error "build with -DJAF_REF_KERNEL_FILE=\"/root/reference/.../rasterize_cuda_kernel.cu\""
#endif
#include JAF_REF_KERNEL_FILE

// All pointers are device pointers.  fim/wim/depth must be pre-filled by the
// caller the way rasterize.py:50-52 does (-1 / 0 / far); faces_inv must be
// zero-filled (rasterize.py:164).  face_inv_map may be a 1-element dummy when
// return_depth == 0 (rasterize.py:66-69).
extern "C" int jaf_ref_forward_face_index_map(const float* faces, float* faces_inv,
                                              int32_t* fim, float* wim, float* depth,
                                              float* face_inv_map, int batch_size,
                                              int num_faces, int image_size, float near_,
                                              float far_, int return_depth) {
  const int threads = 512;
  const dim3 blocks_1((batch_size * num_faces - 1) / threads + 1);
  forward_face_index_map_cuda_kernel_1<float><<<blocks_1, threads>>>(
      faces, faces_inv, batch_size, num_faces, image_size);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return -(int)err;
  const dim3 blocks_2((batch_size * image_size * image_size - 1) / threads + 1);
  forward_face_index_map_cuda_kernel_2<float><<<blocks_2, threads>>>(
      faces, faces_inv, fim, wim, depth, face_inv_map, batch_size, num_faces, image_size,
      near_, far_, 0, 0, return_depth);
  err = cudaGetLastError();
  if (err != cudaSuccess) return -(int)err;
  err = cudaDeviceSynchronize();
  return err == cudaSuccess ? 0 : -(int)err;
}
