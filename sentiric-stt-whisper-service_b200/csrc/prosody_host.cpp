// C ABI of the batched prosody path (include/sw_whisper.h: sw_prosody_segments_*): uploads the utterance,
// runs prosody.cu over all its segments, and finishes each segment with the reference's scalar heuristics
// (/root/reference/src/prosody_extractor.cpp:128-221: octave corrections, gender / emotion proxies,
// 8-D speaker vector) on the host - a few dozen float operations per segment, in the reference's order.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "sequencer.h"

namespace sw {

struct ProsodyState {  // per context: own stream and buffers, so that it never waits for a transcription batch
  std::mutex mu;
  cudaStream_t stream = nullptr;
  DevBuf<uint8_t> d_pcm;
  DevBuf<ProsodySeg> d_segs;
  DevBuf<ProsodyFrame> d_frames;
  DevBuf<ProsodyRaw> d_raw;
  size_t pcm_cap = 0, seg_cap = 0, frame_cap = 0;
  ~ProsodyState() {
    if (stream) cudaStreamDestroy(stream);
  }
};

void prosody_state_free(void* p) { delete static_cast<ProsodyState*>(p); }

namespace {

float soft_norm(float v, float lo, float hi) {  // prosody_extractor.cpp:25-28
  const float t = (v - lo) / (hi - lo);
  return std::max(0.0f, std::min(1.0f, t));
}

void finish_segment(const ProsodyRaw& r, int64_t n_samples, int sample_rate, const sw_prosody_opts& o, sw_prosody* out) {
  out->pitch_mean = r.pitch_median;
  out->pitch_std = r.pitch_std;
  out->energy_mean = r.energy_mean;
  out->energy_std = r.energy_std;
  out->spectral_centroid = r.sc_mean;
  out->zero_crossing_rate = r.zcr_mean;
  // :138-146 octave corrections
  const bool high = out->pitch_mean > o.gender_threshold, low_zcr = out->zero_crossing_rate < 0.024f;
  if (high && low_zcr) out->pitch_mean *= 0.5f;
  else if (out->energy_mean > 0.12f && out->pitch_mean < 240.0f && out->spectral_centroid < 90.0f)
    out->pitch_mean *= 0.5f;
  const float dur_s = (float)n_samples / sample_rate;
  const float rate = dur_s > 0 ? (float)r.peaks / dur_s : 0.0f;
  // :153-162
  if (out->pitch_mean == 0.0f || out->energy_mean < 0.018f) out->gender = '?';
  else if (out->zero_crossing_rate < 0.030f) out->gender = 'M';
  else out->gender = out->pitch_mean > o.gender_threshold ? 'F' : 'M';
  // :165-185
  const float np = out->gender == 'M' ? soft_norm(out->pitch_mean, 60.0f, 180.0f) : soft_norm(out->pitch_mean, 160.0f, 350.0f);
  const float nb = soft_norm(out->spectral_centroid, 40.0f, 150.0f);
  out->valence = ((np * 0.4f) + (nb * 0.6f)) * 2.0f - 1.0f;
  out->valence += 0.35f;
  const float ne = soft_norm(out->energy_mean, 0.02f, 0.20f), nr = soft_norm(rate, 2.0f, 9.0f);
  out->arousal = (ne * 0.7f) + (nr * 0.3f);
  if (out->arousal > 0.65f) out->emotion = out->valence > 0.1f ? SW_EMOTION_EXCITED : SW_EMOTION_ANGRY;
  else if (out->arousal < 0.30f) out->emotion = out->valence < -0.4f ? SW_EMOTION_SAD : SW_EMOTION_NEUTRAL;
  else out->emotion = SW_EMOTION_NEUTRAL;
  // :190-221
  float base;
  if (out->gender == 'M') base = soft_norm(out->pitch_mean, 60.0f, 200.0f) * 0.4f;
  else if (out->gender == 'F') base = 0.6f + (soft_norm(out->pitch_mean, 160.0f, 350.0f) * 0.4f);
  else base = 0.5f;
  float* s = out->speaker_vec;
  s[0] = base;
  s[1] = soft_norm(out->spectral_centroid, 40.0f, 250.0f);
  s[4] = soft_norm(out->zero_crossing_rate, 0.0f, 0.5f) * 0.8f;
  s[2] = soft_norm(out->pitch_std, 5.0f, 100.0f) * 0.1f;
  s[3] = soft_norm(out->energy_mean, 0.0f, 0.3f) * 0.1f;
  s[5] = soft_norm(rate, 1.0f, 12.0f) * 0.1f;
  s[6] = out->arousal * 0.05f;
  s[7] = ((out->valence + 1.0f) / 2.0f) * 0.05f;
}

int run(sw_ctx* ctx, const void* pcm, int is_f32, int64_t n_samples, int sample_rate, const int64_t* seg_begin,
        const int64_t* seg_end, int n_segs, const sw_prosody_opts* opts, sw_prosody* out) {
  SW_CHECK(ctx && ctx->e && out && n_segs >= 0 && (n_segs == 0 || (seg_begin && seg_end)), "bad arguments");
  if (n_segs == 0) return 0;
  SW_CHECK(sample_rate >= 100, "prosody: sample rate %d too low", sample_rate);
  sw_prosody_opts o = opts ? *opts : sw_prosody_default_opts();
  SW_CHECK(o.lpf_alpha > 0.0f && o.lpf_alpha <= 1.0f, "prosody: lpf_alpha %g outside (0, 1]", o.lpf_alpha);
  SW_CUDA_CHECK(cudaSetDevice(ctx->e->device));
  if (!ctx->prosody.load(std::memory_order_acquire)) {
    std::lock_guard<std::mutex> lk(ctx->aux_mu);
    if (!ctx->prosody.load(std::memory_order_relaxed)) {
      ProsodyState* st = new ProsodyState();
      if (cudaStreamCreateWithFlags(&st->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete st;
        set_last_error("prosody: cannot create a stream");
        return -1;
      }
      ctx->prosody.store(st, std::memory_order_release);
    }
  }
  ProsodyState& st = *static_cast<ProsodyState*>(ctx->prosody.load(std::memory_order_acquire));
  std::lock_guard<std::mutex> lk(st.mu);
  const int shift = sample_rate / 100;
  std::vector<ProsodySeg> segs;
  std::vector<int> seg_of;  // analysed segment -> caller's index
  int frame_off = 0, max_frames = 0;
  for (int i = 0; i < n_segs; ++i) {
    memset(&out[i], 0, sizeof(out[i]));
    out[i].gender = '?';
    out[i].emotion = SW_EMOTION_NEUTRAL;
    const int64_t b = seg_begin[i], e = seg_end[i];
    SW_CHECK(b >= 0 && e >= b && e <= n_samples, "prosody: segment %d [%lld, %lld) outside the %lld samples", i,
             (long long)b, (long long)e, (long long)n_samples);
    if (e - b < 160 || !pcm) continue;  // prosody_extractor.cpp:35-47: the neutral record above
    ProsodySeg s;
    s.begin = b;
    s.n_frames = (int)((e - b) / shift);  // i + shift <= n
    s.frame_off = frame_off;
    if (s.n_frames <= 0) continue;
    frame_off += s.n_frames;
    max_frames = std::max(max_frames, s.n_frames);
    segs.push_back(s);
    seg_of.push_back(i);
  }
  if (segs.empty()) return 0;
  const size_t pcm_bytes = (size_t)n_samples * (is_f32 ? 4 : 2);
  if (pcm_bytes > st.pcm_cap) {
    st.d_pcm.release();
    if (st.d_pcm.alloc(pcm_bytes + pcm_bytes / 4)) return -1;
    st.pcm_cap = st.d_pcm.n;
  }
  if (segs.size() > st.seg_cap) {
    st.d_segs.release();
    st.d_raw.release();
    if (st.d_segs.alloc(segs.size() * 2) || st.d_raw.alloc(segs.size() * 2)) return -1;
    st.seg_cap = st.d_segs.n;
  }
  if ((size_t)frame_off > st.frame_cap) {
    st.d_frames.release();
    if (st.d_frames.alloc((size_t)frame_off * 2)) return -1;
    st.frame_cap = st.d_frames.n;
  }
  // the low-pass state forgets its start as (1 - alpha)^n: 48 / alpha samples leave < 1e-20 of it
  const int warm = (int)std::min<double>(1 << 20, ceil(48.0 / o.lpf_alpha));
  cudaStream_t s = st.stream;
  SW_CUDA_CHECK(cudaMemcpyAsync(st.d_pcm.p, pcm, pcm_bytes, cudaMemcpyDefault, s));
  SW_CUDA_CHECK(cudaMemcpyAsync(st.d_segs.p, segs.data(), segs.size() * sizeof(ProsodySeg), cudaMemcpyHostToDevice, s));
  if (prosody_frames(st.d_pcm.p, is_f32, st.d_segs.p, (int)segs.size(), max_frames, shift, warm, o.lpf_alpha,
                     st.d_frames.p, s))
    return -1;
  if (prosody_reduce(st.d_segs.p, st.d_frames.p, (int)segs.size(), shift, sample_rate, o.min_pitch, o.max_pitch,
                     st.d_raw.p, s))
    return -1;
  std::vector<ProsodyRaw> raw(segs.size());
  SW_CUDA_CHECK(cudaMemcpyAsync(raw.data(), st.d_raw.p, raw.size() * sizeof(ProsodyRaw), cudaMemcpyDeviceToHost, s));
  SW_CUDA_CHECK(cudaStreamSynchronize(s));
  for (size_t k = 0; k < segs.size(); ++k) {
    const int i = seg_of[k];
    finish_segment(raw[k], seg_end[i] - seg_begin[i], sample_rate, o, &out[i]);
  }
  return 0;
}

}  // namespace
}  // namespace sw

extern "C" {

sw_prosody_opts sw_prosody_default_opts(void) {  // ProsodyOptions, prosody_extractor.h:20-26
  sw_prosody_opts o;
  o.lpf_alpha = 0.07f;
  o.gender_threshold = 170.0f;
  o.min_pitch = 60.0f;
  o.max_pitch = 500.0f;
  return o;
}

int sw_prosody_segments_f32(sw_ctx* ctx, const float* pcm, int64_t n_samples, int sample_rate,
                            const int64_t* seg_begin, const int64_t* seg_end, int n_segs,
                            const sw_prosody_opts* opts, sw_prosody* out) {
  try {
    return sw::run(ctx, pcm, 1, n_samples, sample_rate, seg_begin, seg_end, n_segs, opts, out);
  } catch (const std::exception& ex) {
    sw::set_last_error("internal error: %s", ex.what());
    return -1;
  }
}

int sw_prosody_segments_pcm16(sw_ctx* ctx, const int16_t* pcm, int64_t n_samples, int sample_rate,
                              const int64_t* seg_begin, const int64_t* seg_end, int n_segs,
                              const sw_prosody_opts* opts, sw_prosody* out) {
  try {
    return sw::run(ctx, pcm, 0, n_samples, sample_rate, seg_begin, seg_end, n_segs, opts, out);
  } catch (const std::exception& ex) {
    sw::set_last_error("internal error: %s", ex.what());
    return -1;
  }
}

}  // extern "C"
