/*
 * samplerate.h - source-compatibility shim for the one libsamplerate call the reference makes
 * (SttEngine::resample_audio, /root/reference/src/stt_engine.cpp:87-106: src_simple with SRC_SINC_FASTEST,
 * one channel). libsamplerate is not in the reference tree; libwhisper_compat.so implements src_simple on
 * the engine's CUDA windowed-sinc resampler (sw_resample_f32). Same published method, its own window: not
 * bit-compatible with libsamplerate.
 */
#ifndef SAMPLERATE_H
#define SAMPLERATE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  const float* data_in;
  float* data_out;
  long input_frames, output_frames;
  long input_frames_used, output_frames_gen;
  int end_of_input;
  double src_ratio;
} SRC_DATA;

enum {
  SRC_SINC_BEST_QUALITY = 0,
  SRC_SINC_MEDIUM_QUALITY = 1,
  SRC_SINC_FASTEST = 2,
  SRC_ZERO_ORDER_HOLD = 3,
  SRC_LINEAR = 4,
};

__attribute__((visibility("default"))) int src_simple(SRC_DATA* data, int converter_type, int channels);
__attribute__((visibility("default"))) const char* src_strerror(int error);

#ifdef __cplusplus
}
#endif
#endif
