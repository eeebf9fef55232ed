/* oracle/orc_nethook.cpp -- SURVEY 8(f)-4 (test infrastructure).
 * Restates Segmentation::_makeTensor (sloam/src/segmentation/inference.cpp:167-198) for the
 * one-channel (range) configuration and Segmentation::_mask (inference.cpp:275-300). */
#include <cstdint>
#include <vector>

extern "C" {

/* returns the number of invalid pixels; invalid_idx (may be null) receives them in order */
int orc_make_tensor(const float *range_image, int n, float mean, float stdv, float *tensor,
                    uint8_t *invalid, int32_t *invalid_idx) {
  int n_inv = 0;
  for (int pixel_id = 0; pixel_id < n; ++pixel_id) {
    const float v = range_image[pixel_id];
    /* the lambda of :183 takes an int: (i == 0.0f) || isnan(i) on the truncated value */
    const bool all_zeros = (v > -1.0f) && (v < 1.0f);
    if (all_zeros) {
      if (invalid_idx) invalid_idx[n_inv] = pixel_id;
      ++n_inv;
    }
    invalid[pixel_id] = all_zeros ? 1 : 0;
    tensor[pixel_id] = all_zeros ? v : (v - mean) / stdv; /* :190 */
  }
  return n_inv;
}

void orc_mask_from_logits(const float *output, int n, const uint8_t *invalid, uint8_t *mask) {
  const size_t channel_offset = (size_t)n;
  for (int pixel_id = 0; pixel_id < n; ++pixel_id) {
    size_t max_idx = (size_t)pixel_id;
    unsigned char out_idx = 0;
    for (unsigned char i = 1; i < 3; ++i) { /* _n_classes = 3 */
      const size_t buffer_idx = channel_offset * i + pixel_id;
      if (output[max_idx] < output[buffer_idx]) { max_idx = buffer_idx; out_idx = i; }
    }
    if (out_idx == 2) out_idx = 255;
    mask[pixel_id] = out_idx;
  }
  if (invalid)
    for (int i = 0; i < n; ++i)
      if (invalid[i]) mask[i] = 0;
}

}
