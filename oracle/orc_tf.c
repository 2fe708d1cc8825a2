/*
 * orc_tf.c -- ORACLE (test infrastructure): transfer-function tables.
 *
 * Follows (reference file:line):
 *   TransferFunction1D::SetStdFunction       IO/TransferFunction1D.cpp:74-113
 *   TransferFunction1D::GetByteArray         IO/TransferFunction1D.cpp:311-330
 *   TransferFunction1D::ComputeNonZeroLimits IO/TransferFunction1D.cpp:362-371
 *   TransferFunction2D::ComputeNonZeroLimits IO/TransferFunction2D.cpp:378-397
 */
#include "orc.h"

static float smoothstep_(float x) { return 3 * x * x - 2 * x * x * x; }

void orc_tf1d_std(float* rgba, uint32_t n, float center, float inv_gradient) {
  float c = center < 0 ? 0 : center > 1 ? 1 : center;
  float g = inv_gradient < 0 ? 0 : inv_gradient > 1 ? 1 : inv_gradient;
  size_t ic = (size_t)((n - 1) * c);
  size_t ig = (size_t)((n - 1) * g);
  size_t start = (ig / 2 > ic) ? 0 : ic - ig / 2;
  size_t end = (ig / 2 + ic > n) ? n : ic + ig / 2;
  for (int comp = 0; comp < 4; comp++) {
    for (size_t i = 0; i < start; i++) rgba[4 * i + comp] = 0;
    for (size_t i = start; i < end; i++)
      rgba[4 * i + comp] = smoothstep_((float)(i - ic + ig / 2) / (float)ig);
    for (size_t i = end; i < n; i++) rgba[4 * i + comp] = 1;
  }
}

void orc_tf1d_bytes(const float* rgba, uint32_t n, uint8_t* out) {
  for (size_t i = 0; i < (size_t)n * 4; i++) {
    float v = rgba[i];
    v = v < 1.0f ? v : 1.0f;
    v = v > 0.0f ? v : 0.0f;
    out[i] = (uint8_t)(v * 255);
  }
}

void orc_tf1d_nonzero(const float* rgba, uint32_t n, uint64_t* lo, uint64_t* hi) {
  *lo = n;
  *hi = 0;
  for (uint32_t i = 0; i < n; i++)
    if (rgba[4 * i + 3] != 0) {
      if (i < *lo) *lo = i;
      *hi = i;
    }
}

void orc_tf2d_nonzero(const uint8_t* rgba, uint32_t w, uint32_t h, uint64_t out[4]) {
  /* (xmin, xmax, ymin, ymax) over texels with alpha != 0; initialised (w, 0, h, 0) */
  out[0] = w; out[1] = 0; out[2] = h; out[3] = 0;
  for (uint32_t y = 0; y < h; y++)
    for (uint32_t x = 0; x < w; x++)
      if (rgba[4 * ((size_t)y * w + x) + 3] != 0) {
        if (x < out[0]) out[0] = x;
        if (x > out[1]) out[1] = x;
        if (y < out[2]) out[2] = y;
        if (y > out[3]) out[3] = y;
      }
}
