// k8_nethook.cu -- SURVEY 8(f)-4: the two elementwise passes either side of the
// segmentation network, so that an external RangeNet++ engine can be plugged in between the
// range image of stage a1 and the {0, 1, 255} mask stage a2 consumes.
//
// Replaces Segmentation::_makeTensor (sloam/src/segmentation/inference.cpp:167-198) and
// Segmentation::_mask (inference.cpp:275-300).  The network itself stays out of scope.
//
// HBM-bound streaming kernels: 4 B in, 5 B out per pixel (tensor); 12 B + 1 B in, 1 B out
// per pixel (mask).
#include "common.cuh"

namespace sb {

// _makeTensor, one channel (_img_d = 1, the range): a pixel is invalid when its value
// converts to the int 0 (the reference's all_of lambda takes an `int`, :183), i.e. |v| < 1;
// valid pixels are normalised (v - mean) / std in float, invalid ones are passed through.
// The reference collects invalid pixel indices; a per-pixel flag carries the same information
// and keeps the pass order-free (n_invalid = its population count).
__global__ void make_tensor_kernel(long long total, int N, const float *__restrict__ range_image, float mean,
                                   float stdv, float *__restrict__ tensor, uint8_t *__restrict__ invalid,
                                   int32_t *__restrict__ n_invalid) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool inv = false;
  if (g < total) {
    const float v = range_image[g];
    // float -> int conversion of the lambda argument: truncation; NaN and out-of-range values
    // are undefined in C++ (x86 gives INT_MIN), treated as "not zero" here
    inv = (v > -1.0f) && (v < 1.0f);
    tensor[g] = inv ? v : (v - mean) / stdv;
    invalid[g] = inv ? 1 : 0;
  }
  if (n_invalid) {
    // a warp may straddle two keyframes when N is not a multiple of 32: count per lane group
    const int k = g < total ? (int)(g / N) : -1;
    const unsigned same = __match_any_sync(0xFFFFFFFFu, k);
    const unsigned votes = __ballot_sync(0xFFFFFFFFu, inv) & same;
    if (k >= 0 && votes && (threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(&n_invalid[k], __popc(votes));
  }
}

// _mask: argmax over the 3 class scores with a strict '<' (the first maximum wins), class 2
// -> 255, invalid pixels -> 0.  logits are channel-major per keyframe: [K][3][N].
__global__ void mask_from_logits_kernel(long long total, int N, const float *__restrict__ logits,
                                        const uint8_t *__restrict__ invalid, uint8_t *__restrict__ mask) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  const long long k = g / N, i = g - k * N;
  const float *o = logits + k * 3 * (long long)N + i;
  float best = o[0];
  unsigned char out = 0;
  const float c1 = o[N], c2 = o[2 * (long long)N];
  if (best < c1) { best = c1; out = 1; }
  if (best < c2) { out = 255; }
  if (invalid && invalid[g]) out = 0;
  mask[g] = out;
}

}  // namespace sb

using namespace sb;

extern "C" {

int sloam_b200_make_tensor_dev(sloam_ctx *c, int K, const float *range_image, float mean, float stdv,
                               float *tensor, uint8_t *invalid, int32_t *n_invalid) {
  if (!c || K <= 0 || K > c->max_k || !range_image || !tensor || !invalid || !(stdv != 0.0f))
    return set_err(c, SLOAM_E_INVALID, "make_tensor: bad arguments");
  const long long total = (long long)K * c->hp.N;
  if (n_invalid) SB_CUDA(c, cudaMemsetAsync(n_invalid, 0, sizeof(int32_t) * (size_t)K, c->stream));
  make_tensor_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(total, c->hp.N, range_image, mean, stdv,
                                                                             tensor, invalid, n_invalid);
  SB_LAUNCH_CHECK(c);
  return SLOAM_OK;
}

int sloam_b200_mask_from_logits_dev(sloam_ctx *c, int K, const float *logits, const uint8_t *invalid,
                                    uint8_t *mask) {
  if (!c || K <= 0 || K > c->max_k || !logits || !mask)
    return set_err(c, SLOAM_E_INVALID, "mask_from_logits: bad arguments");
  const long long total = (long long)K * c->hp.N;
  mask_from_logits_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(total, c->hp.N, logits, invalid, mask);
  SB_LAUNCH_CHECK(c);
  return SLOAM_OK;
}

}  // extern "C"
