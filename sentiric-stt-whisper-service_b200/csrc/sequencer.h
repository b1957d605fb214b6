// Result containers behind the C ABI and the batched whisper_full_with_state entry point.
#pragma once
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "engine.h"

struct sw_ctx {
  sw::Engine* e = nullptr;            // lane 0: owns the model; the stage hooks run here
  std::vector<sw::Engine*> lanes;     // further lanes: own buffers and stream, weights shared with e
  // side states, created on first use under aux_mu (never under Engine::mu, which a running transcription
  // batch holds for its whole duration): sw::ProsodyState (prosody_host.cpp), sw::ResampleState (resample_host.cpp)
  std::mutex aux_mu;
  std::atomic<void*> prosody{nullptr};
  std::atomic<void*> resample{nullptr};
};
struct sw_segment {
  int64_t t0 = 0, t1 = 0;
  std::string text;
  std::vector<sw_token_data> tokens;
  bool speaker_turn_next = false;
};
struct sw_result {
  std::vector<sw_segment> segs;
  int lang_id = -1;
  int n_decode_steps = 0;
  int n_windows = 0;
};

namespace sw {
// pcm[i]: n_samples[i] host samples (int16 or f32). out[i] receives a new sw_result.
// Returns 0, or non-zero on failure/abort (all out[i] are null then).
// langs (optional): language of every utterance ("en", ..., "auto" / null = detect); null = params->language for all
int run_full_batch(Engine* e, const sw_full_params* params, const void* const* pcm, const int* n_samples,
                   int n, bool is_f32, sw_result** out, const char* const* langs = nullptr);
// The same over all lanes of a context: utterances are dealt to the lanes (they are independent units,
// SURVEY.md §8e), one host thread per lane drives its engine, results land in the caller's order.
int run_full_batch_lanes(sw_ctx* ctx, const sw_full_params* params, const void* const* pcm, const int* n_samples,
                         int n, bool is_f32, sw_result** out, const char* const* langs = nullptr);
const char* last_error_string();
void prosody_state_free(void* p);
void resample_state_free(void* p);
}  // namespace sw
