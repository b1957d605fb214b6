// SttEngine: the C++ class the service's HTTP / gRPC handlers call (reference:
// /root/reference/src/stt_engine.h:59-116). Public API, value types, units (centiseconds) and error
// behaviour are the reference's; underneath, whisper.cpp's context + state pool are replaced by one
// sw_ctx (include/sw_whisper.h) and a dispatcher thread that BATCHES concurrent callers into a
// single device pass - the reference runs one whisper_state per request and never batches
// (stt_engine.cpp:36-42, 245).
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sw_whisper.h"
#include "config.h"
#include "prosody.h"
#include "speaker_cluster.h"

struct TokenData {
  std::string text;
  float p;
  int64_t t0;
  int64_t t1;
};

struct RequestOptions {
  std::string language;
  std::string prompt;
  bool translate = false;
  bool enable_diarization = false;
  float temperature = -1.0f;
  int beam_size = -1;
  int best_of = -1;
  ProsodyOptions prosody_opts;
  std::function<bool()> should_abort = nullptr;
};

struct TranscriptionResult {
  std::string text;
  std::string language;
  float prob;
  int64_t t0;
  int64_t t1;
  bool speaker_turn_next;
  std::vector<TokenData> tokens;
  int token_count = 0;
  std::string gender_proxy;
  std::string emotion_proxy;
  float arousal = 0.0f;
  float valence = 0.0f;
  AffectiveTags affective;
  std::string speaker_id;
};

class EngineBusyException : public std::runtime_error {
 public:
  EngineBusyException(const std::string& msg) : std::runtime_error(msg) {}
};

class SttEngine {
 public:
  explicit SttEngine(const Settings& settings);
  ~SttEngine();
  bool is_ready() const;
  const Settings& get_settings() const { return settings_; }

  struct PerformanceMetrics {
    double queue_time_ms;
    double processing_time_ms;
    int token_count;
  };

  std::vector<TranscriptionResult> transcribe(const std::vector<float>& pcmf32, int input_sample_rate,
                                              const RequestOptions& options,
                                              PerformanceMetrics* out_metrics = nullptr);
  std::vector<TranscriptionResult> transcribe_pcm16(const std::vector<int16_t>& pcm16, int input_sample_rate,
                                                    const RequestOptions& options,
                                                    PerformanceMetrics* out_metrics = nullptr);

  // B200 additions
  void set_prosody_fn(ProsodyFn fn) { prosody_fn_ = std::move(fn); }
  // Speech pre-gate of stt_engine.cpp:108-115,169-194. The reference evaluates a Silero model through
  // whisper.cpp's CPU VAD (whisper_vad_detect_speech); neither that code nor the model file is part of this
  // build, so the gate is a hook: return false for "no speech" and the request gets the reference's
  // placeholder result (:173-192) without touching the GPU. Without a hook the gate is open, which is what
  // the reference does when its VAD context failed to load (vad_ctx_ == nullptr -> true, :109); the
  // constructor says so loudly when Settings.enable_vad asks for a gate that is not there.
  using VadFn = std::function<bool(const float* pcm, size_t n_samples)>;
  void set_vad_fn(VadFn fn) { vad_fn_ = std::move(fn); }
  bool vad_gate_active() const { return settings_.enable_vad && (bool)vad_fn_; }
  long batches_run() const { return batches_run_; }
  long requests_batched() const { return requests_batched_; }

 private:
  struct Request;
  std::vector<TranscriptionResult> run_request(const float* pcm, size_t n, const int16_t* pcm16,
                                               const RequestOptions& options, PerformanceMetrics* out_metrics,
                                               std::chrono::high_resolution_clock::time_point t_start);
  void acquire_slot();
  void release_slot();
  void dispatcher_loop();

  Settings settings_;
  sw_ctx* ctx_ = nullptr;
  ProsodyFn prosody_fn_;  // empty: the batched GPU path (sw_prosody_segments_*)
  VadFn vad_fn_;          // empty: no speech pre-gate (see set_vad_fn)
  std::mutex vad_mutex_;  // the reference serialises VAD calls (stt_engine.cpp:110)

  // admission: at most parallel_requests requests in the engine (the reference's whisper_state pool)
  std::mutex pool_mutex_;
  std::condition_variable pool_cv_;
  int free_slots_ = 0;
  int capacity_ = 0;             // admission slots in total
  std::atomic<int> inside_{0};   // callers currently admitted

  // batching dispatcher
  std::mutex q_mutex_;
  std::condition_variable q_cv_;
  std::deque<Request*> queue_;
  bool stopping_ = false;
  std::thread dispatcher_;
  long batches_run_ = 0, requests_batched_ = 0;
  int lanes_ = 1;  // lanes of the context: a full device pass is max_batch * lanes_ requests
};
