// CPU statement of the engine's sample-rate converter (sw_resample_f32). TEST INFRASTRUCTURE ONLY.
//
// PARITY UNPINNED against the reference: /root/reference/src/stt_engine.cpp:87-115 calls libsamplerate's
// src_simple(SRC_SINC_FASTEST), a third-party library that is neither in the reference tree nor in this
// image, and whose coefficient table cannot be restated from its published description. What is built
// instead is the same published method (band-limited interpolation: a windowed sinc sampled on a fine grid,
// linear interpolation between grid points, cut-off scaled by the ratio when downsampling - J. O. Smith,
// "Digital Audio Resampling", the method libsamplerate's sinc converters implement) with its own window.
// This file is the bit-exact definition the CUDA kernel is tested against, and the quality checks
// (tests/test_resample.py: tone SNR, agreement with scipy.signal.resample_poly) are what stands in for a
// reference comparison.
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include <vector>

extern "C" {

enum { RS_ZEROS = 16, RS_GRID = 256 };  // zero crossings per side, table points per zero crossing

// half of the symmetric filter: h[i] = sinc(i / GRID) * kaiser(i / (ZEROS * GRID)), i = 0 .. ZEROS*GRID (+1 guard)
void ora_resample_table(float* table) {
  const double beta = 9.0;
  auto i0 = [](double x) {  // modified Bessel function of the first kind, order 0
    double s = 1.0, t = 1.0;
    for (int k = 1; k < 60; ++k) {
      t *= (x / (2.0 * k)) * (x / (2.0 * k));
      s += t;
    }
    return s;
  };
  const int n = RS_ZEROS * RS_GRID;
  for (int i = 0; i <= n + 1; ++i) {
    const double t = (double)i / RS_GRID;
    const double r = (double)i / n;
    const double w = r >= 1.0 ? 0.0 : i0(beta * sqrt(1.0 - r * r)) / i0(beta);
    const double s = i == 0 ? 1.0 : sin(M_PI * t) / (M_PI * t);
    table[i] = (float)(s * w);
  }
}

int64_t ora_resample_out_len(int64_t n_in, int sr_in, int sr_out) {
  return (int64_t)(((__int128)n_in * sr_out) / sr_in);  // floor(n_in * sr_out / sr_in)
}

// out[n] = scale * sum_k in[k] * h(|k - x| * scale),  x = n * sr_in / sr_out,  scale = min(1, sr_out / sr_in)
// Single precision, no FMA contraction, taps walked left to right: what the CUDA kernel does.
void ora_resample_f32(const float* in, int64_t n_in, int sr_in, int sr_out, float* out) {
  std::vector<float> table(RS_ZEROS * RS_GRID + 2);
  ora_resample_table(table.data());
  const int64_t n_out = ora_resample_out_len(n_in, sr_in, sr_out);
  const float scale = sr_out < sr_in ? (float)sr_out / (float)sr_in : 1.0f;
  const float gscale = scale * (float)RS_GRID;        // table steps per input sample
  const int half = (int)ceilf((float)RS_ZEROS / scale);  // taps per side
  for (int64_t n = 0; n < n_out; ++n) {
    const __int128 num = (__int128)n * sr_in;
    const int64_t xi = (int64_t)(num / sr_out);
    const float frac = (float)((double)(int64_t)(num % sr_out) / (double)sr_out);
    float acc = 0.0f;
    for (int k = -half + 1; k <= half; ++k) {
      const int64_t idx = xi + k;
      if (idx < 0 || idx >= n_in) continue;
      const float dist = fabsf((float)k - frac) * gscale;  // position in the table
      const int ti = (int)dist;
      if (ti >= RS_ZEROS * RS_GRID) continue;
      const float tf = dist - (float)ti;
      const float w = table[ti] + tf * (table[ti + 1] - table[ti]);
      acc = acc + in[idx] * w;
    }
    out[n] = acc * scale;
  }
}

}  // extern "C"
