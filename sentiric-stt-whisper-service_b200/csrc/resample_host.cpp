// C ABI of the sample-rate converter (include/sw_whisper.h: sw_resample_*): see resample.cu.
#include <math.h>

#include <algorithm>
#include <vector>

#include "sequencer.h"

namespace sw {

struct ResampleState {
  std::mutex mu;
  cudaStream_t stream = nullptr;
  DevBuf<float> d_table, d_in, d_out;
  size_t in_cap = 0, out_cap = 0;
  ~ResampleState() {
    if (stream) cudaStreamDestroy(stream);
  }
};
void resample_state_free(void* p) { delete static_cast<ResampleState*>(p); }

namespace {

// half of the symmetric filter on a grid of RS_GRID points per zero crossing: sinc x Kaiser(beta = 9)
void build_table(std::vector<float>& table) {
  const double beta = 9.0;
  auto i0 = [](double x) {
    double s = 1.0, t = 1.0;
    for (int k = 1; k < 60; ++k) {
      t *= (x / (2.0 * k)) * (x / (2.0 * k));
      s += t;
    }
    return s;
  };
  const int n = RS_ZEROS * RS_GRID;
  table.resize(n + 2);
  for (int i = 0; i <= n + 1; ++i) {
    const double t = (double)i / RS_GRID;
    const double r = (double)i / n;
    const double w = r >= 1.0 ? 0.0 : i0(beta * sqrt(1.0 - r * r)) / i0(beta);
    const double s = i == 0 ? 1.0 : sin(M_PI * t) / (M_PI * t);
    table[i] = (float)(s * w);
  }
}

}  // namespace
}  // namespace sw

extern "C" {

int64_t sw_resample_out_len(int64_t n_in, int sr_in, int sr_out) {
  if (n_in <= 0 || sr_in <= 0 || sr_out <= 0) return 0;
  return (int64_t)(((__int128)n_in * sr_out) / sr_in);
}

int sw_resample_f32(sw_ctx* ctx, const float* in, int64_t n_in, int sr_in, int sr_out, float* out) {
  using namespace sw;
  SW_CHECK(ctx && ctx->e && in && out && n_in > 0, "bad arguments");
  SW_CHECK(sr_in >= 1000 && sr_in <= 768000 && sr_out >= 1000 && sr_out <= 768000, "resample: rates %d -> %d", sr_in,
           sr_out);
  SW_CUDA_CHECK(cudaSetDevice(ctx->e->device));
  if (!ctx->resample.load(std::memory_order_acquire)) {
    std::lock_guard<std::mutex> lk(ctx->aux_mu);
    if (!ctx->resample.load(std::memory_order_relaxed)) {
      ResampleState* st = new ResampleState();
      std::vector<float> table;
      build_table(table);
      if (cudaStreamCreateWithFlags(&st->stream, cudaStreamNonBlocking) != cudaSuccess || st->d_table.alloc(table.size()) ||
          cudaMemcpy(st->d_table.p, table.data(), table.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
        delete st;
        set_last_error("resample: cannot set up the device state");
        return -1;
      }
      ctx->resample.store(st, std::memory_order_release);
    }
  }
  ResampleState& st = *static_cast<ResampleState*>(ctx->resample.load(std::memory_order_acquire));
  std::lock_guard<std::mutex> lk(st.mu);
  const int64_t n_out = sw_resample_out_len(n_in, sr_in, sr_out);
  if (n_out <= 0) return 0;
  if ((size_t)n_in > st.in_cap) {
    st.d_in.release();
    if (st.d_in.alloc((size_t)n_in * 5 / 4 + 1024)) return -1;
    st.in_cap = st.d_in.n;
  }
  if ((size_t)n_out > st.out_cap) {
    st.d_out.release();
    if (st.d_out.alloc((size_t)n_out * 5 / 4 + 1024)) return -1;
    st.out_cap = st.d_out.n;
  }
  const float scale = sr_out < sr_in ? (float)sr_out / (float)sr_in : 1.0f;
  const float gscale = scale * (float)RS_GRID;
  const int half = (int)ceilf((float)RS_ZEROS / scale);
  SW_CUDA_CHECK(cudaMemcpyAsync(st.d_in.p, in, (size_t)n_in * 4, cudaMemcpyDefault, st.stream));
  if (resample_f32(st.d_in.p, n_in, sr_in, sr_out, st.d_table.p, scale, gscale, half, n_out, st.d_out.p, st.stream))
    return -1;
  SW_CUDA_CHECK(cudaMemcpyAsync(out, st.d_out.p, (size_t)n_out * 4, cudaMemcpyDefault, st.stream));
  SW_CUDA_CHECK(cudaStreamSynchronize(st.stream));
  return 0;
}

}  // extern "C"
