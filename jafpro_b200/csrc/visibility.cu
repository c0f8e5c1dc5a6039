// Per-reference visibility for row F: the reference's rule is SMPLRenderer.get_vis_f2pts
// (src/nmr.py:507-546) — faces that do not appear in the SOURCE pose's face-index map get the sentinel -2, i.e. a
// target pixel can only be filled from reference k if the face it shows is visible in reference k:
//     vis_k[p] = fim_tgt[p] in unique(fim_src_k).
// Two streaming passes: mark the faces present in each source fim, then look every target pixel's face up.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
k_mark_faces(const int* __restrict__ fim_src, long n, int HW, int F, uint8_t* __restrict__ seen) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int f = i < n ? ld_stream_s32(fim_src + i) : -1;
  const long img = i < n ? i / HW : -1;
  // face-index maps are piecewise constant along x: only the first pixel of a run inside the warp stores
  const int fl = __shfl_up_sync(0xffffffffu, f, 1);
  const long il = __shfl_up_sync(0xffffffffu, img, 1);
  const bool first = (threadIdx.x & 31) == 0 || fl != f || il != img;
  if (first && f >= 0 && f < F) seen[img * F + f] = 1;  // racing stores all write 1
}

__global__ void __launch_bounds__(256)
k_vis_lookup(const int* __restrict__ fim_tgt, const uint8_t* __restrict__ seen, int K, int HW, int F, long n,
             float* __restrict__ vis) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long bk = i / HW;
  const int p = (int)(i - bk * HW);
  const int f = __ldg(fim_tgt + (bk / K) * HW + p);
  const bool v = f >= 0 && f < F && seen[bk * F + f] != 0;
  st_stream_f32(vis + i, v ? 1.f : 0.f);
}

}  // namespace

extern "C" int jaf_face_visibility(const int32_t* fim_src, const int32_t* fim_tgt, int B, int K, int HW, int F,
                                   uint8_t* seen, float* vis, void* stream) {
  JAF_REQUIRE(fim_src && seen, "null pointer");
  JAF_REQUIRE((fim_tgt == nullptr) == (vis == nullptr), "fim_tgt and vis go together");
  JAF_REQUIRE(B > 0 && K > 0 && HW > 0 && F > 0, "bad sizes");
  cudaStream_t st = jaf::as_stream(stream);
  const long n = (long)B * K * HW;
  JAF_CUDA(cudaMemsetAsync(seen, 0, (size_t)B * K * F, st));
  k_mark_faces<<<jaf::ceil_div(n, 256), 256, 0, st>>>(fim_src, n, HW, F, seen);
  int launches = 1;
  if (vis) {
    k_vis_lookup<<<jaf::ceil_div(n, 256), 256, 0, st>>>(fim_tgt, seen, K, HW, F, n, vis);
    ++launches;
  }
  return jaf::finish_launch("k_face_visibility", launches);
}
