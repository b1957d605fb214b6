// See stt_engine.h. Line references are to /root/reference/src/stt_engine.cpp, whose observable
// behaviour each block reproduces.
#include "stt_engine.h"

#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <cmath>

#include "model_manager.h"
#include "text_filters.h"

using Clock = std::chrono::high_resolution_clock;

struct SttEngine::Request {
  const float* pcm = nullptr;      // exactly one of pcm / pcm16 is set
  const int16_t* pcm16 = nullptr;
  int n = 0;
  sw_full_params params;
  std::string language, prompt;   // own the strings params points at
  std::function<bool()> abort_fn;
  sw_result* result = nullptr;
  int rc = 0;
  std::string error;
  bool done = false;
  std::mutex m;
  std::condition_variable cv;
  // requests run in one device pass only if every decode setting is identical
  bool compatible(const Request& o) const {
    const sw_full_params &a = params, &b = o.params;
    return !abort_fn && !o.abort_fn && (pcm16 != nullptr) == (o.pcm16 != nullptr) && a.strategy == b.strategy &&
           a.beam_size == b.beam_size && a.best_of == b.best_of && a.temperature == b.temperature &&
           a.translate == b.translate && a.tdrz_enable == b.tdrz_enable && prompt == o.prompt;
    // (the language may differ: it only selects a prompt token, sw_full_batch_*_lang takes one per utterance)
  }
};

static int abort_trampoline(void* user) {  // stt_engine.cpp:17-23
  auto* fn = static_cast<std::function<bool()>*>(user);
  return (fn && *fn && (*fn)()) ? 1 : 0;
}

SttEngine::SttEngine(const Settings& settings) : settings_(settings) {
  // :26-34 - load the model or throw
  const std::string model_path = settings_.model_dir + "/" + settings_.model_filename;
  // row budget of the context: max_batch windows x the configured decoders per window. A request may still ask
  // for up to 8 decoders (RequestOptions.beam_size / best_of; whisper.cpp's WHISPER_MAX_DECODERS) - such a pass
  // then carries fewer windows (sw_ctx_params.max_beams)
  const int max_beams = std::max(1, std::min(8, std::max(settings_.beam_size, settings_.best_of)));
  ctx_ = ModelManager::load_to_device(settings_, model_path, max_beams);
  if (!ctx_) {
    fprintf(stderr, "[stt_engine] %s\n", sw_last_error());
    throw std::runtime_error("Whisper model initialization failed");
  }
  // :36-42 - the state pool becomes an admission counter. The reference sizes it with parallel_requests
  // (default 2: one whisper_state each, decoded one by one); here a device pass decodes max_batch windows
  // together, so admitting only 2 callers would starve it. admission_slots (0 = auto) decouples the two:
  // auto admits max(parallel_requests, 2 * max_batch) callers; the bounded wait and EngineBusyException stay.
  free_slots_ = settings_.admission_slots > 0
                    ? settings_.admission_slots
                    : std::max(std::max(1, settings_.parallel_requests), 2 * std::max(1, settings_.max_batch));
  capacity_ = free_slots_;
  // :44-52 - the reference loads a Silero model into whisper.cpp's CPU VAD. Not part of this build (see
  // set_vad_fn in the header): say so instead of silently running without the gate.
  if (settings_.enable_vad) {
    const std::string vad_path = settings_.model_dir + "/" + settings_.vad_model_filename;
    FILE* f = fopen(vad_path.c_str(), "rb");
    if (f) fclose(f);
    fprintf(stderr,
            "[stt_engine] WARNING: enable_vad is set but this engine has no Silero evaluator (%s %s). The speech "
            "pre-gate stays OPEN, as in the reference when whisper_vad_init_from_file_with_params fails, until "
            "SttEngine::set_vad_fn() installs one; silence is decoded by Whisper and filtered by no_speech_thold.\n",
            vad_path.c_str(), f ? "is present but cannot be evaluated" : "is missing");
  }
  {
    sw_stats st;
    if (sw_ctx_get_stats(ctx_, &st, 0) == 0 && st.n_lanes > 1) lanes_ = (int)st.n_lanes;
  }
  dispatcher_ = std::thread([this] { dispatcher_loop(); });
}

SttEngine::~SttEngine() {
  {
    std::lock_guard<std::mutex> lk(q_mutex_);
    stopping_ = true;
  }
  q_cv_.notify_all();
  if (dispatcher_.joinable()) dispatcher_.join();
  if (ctx_) sw_ctx_destroy(ctx_);
}

bool SttEngine::is_ready() const { return ctx_ != nullptr; }

void SttEngine::acquire_slot() {  // :63-79
  std::unique_lock<std::mutex> lock(pool_mutex_);
  const bool ok = pool_cv_.wait_for(lock, std::chrono::milliseconds(settings_.request_queue_timeout_ms),
                                    [this] { return free_slots_ > 0; });
  if (!ok) throw EngineBusyException("Server is busy (Queue timeout)");
  --free_slots_;
  inside_.fetch_add(1, std::memory_order_relaxed);
}

void SttEngine::release_slot() {  // :81-85
  {
    std::lock_guard<std::mutex> lock(pool_mutex_);
    ++free_slots_;
    inside_.fetch_sub(1, std::memory_order_relaxed);
  }
  pool_cv_.notify_one();
}

void SttEngine::dispatcher_loop() {
  while (true) {
    std::vector<Request*> batch;
    {
      std::unique_lock<std::mutex> lk(q_mutex_);
      q_cv_.wait(lk, [this] { return stopping_ || !queue_.empty(); });
      if (stopping_ && queue_.empty()) return;
      // a full device pass is max_batch windows on EVERY lane of the context (two lanes: 2 x max_batch)
      const int full_pass = settings_.max_batch * lanes_;
      if (settings_.batch_window_us > 0 && (int)queue_.size() < full_pass) {
        // give concurrent callers a moment to join this device pass - unless nobody else CAN join: when every
        // admission slot is taken and all of their holders already wait in this queue, no caller can arrive before
        // a pass has run, and the rest of the window would be a pure loss
        q_cv_.wait_for(lk, std::chrono::microseconds(settings_.batch_window_us), [this, full_pass] {
          return stopping_ || (int)queue_.size() >= full_pass ||
                 ((int)queue_.size() >= capacity_ && inside_.load(std::memory_order_relaxed) >= capacity_);
        });
      }
      Request* head = queue_.front();
      queue_.pop_front();
      batch.push_back(head);
      for (auto it = queue_.begin(); it != queue_.end() && (int)batch.size() < 4 * settings_.max_batch;) {
        if (head->compatible(**it)) {
          batch.push_back(*it);
          it = queue_.erase(it);
        } else {
          ++it;
        }
      }
    }
    const int n = (int)batch.size();
    std::vector<int> lens(n);
    std::vector<sw_result*> res(n, nullptr);
    for (int i = 0; i < n; ++i) lens[i] = batch[i]->n;
    int rc;
    std::vector<const char*> langs(n);
    for (int i = 0; i < n; ++i) langs[i] = batch[i]->language.c_str();
    if (batch[0]->pcm16) {
      std::vector<const int16_t*> ptrs(n);
      for (int i = 0; i < n; ++i) ptrs[i] = batch[i]->pcm16;
      rc = sw_full_batch_pcm16_lang(ctx_, &batch[0]->params, ptrs.data(), lens.data(), n, langs.data(), res.data());
    } else {
      std::vector<const float*> ptrs(n);
      for (int i = 0; i < n; ++i) ptrs[i] = batch[i]->pcm;
      rc = sw_full_batch_f32_lang(ctx_, &batch[0]->params, ptrs.data(), lens.data(), n, langs.data(), res.data());
    }
    const std::string err = rc ? sw_last_error() : "";
    ++batches_run_;
    requests_batched_ += n;
    for (int i = 0; i < n; ++i) {
      Request* r = batch[i];
      {
        std::lock_guard<std::mutex> lk(r->m);
        r->rc = rc;
        r->error = err;
        r->result = res[i];
        r->done = true;
        // notify while holding r->m: the Request lives on the waiter's stack, and a waiter that saw done == true
        // after an unlock-then-notify could return and destroy the condition variable under notify_one()
        r->cv.notify_one();
      }
    }
  }
}

std::vector<TranscriptionResult> SttEngine::transcribe_pcm16(const std::vector<int16_t>& pcm16, int input_sample_rate,
                                                             const RequestOptions& options,
                                                             PerformanceMetrics* out_metrics) {
  // :117-125 converts to float on the host; here the int16 samples go to the device as they are
  // and the /32768 happens in the front-end kernel's load.
  const auto t_start = Clock::now();
  if (input_sample_rate != 16000) {
    std::vector<float> f(pcm16.size());
    for (size_t i = 0; i < pcm16.size(); ++i) f[i] = static_cast<float>(pcm16[i]) / 32768.0f;
    return transcribe(f, input_sample_rate, options, out_metrics);
  }
  return run_request(nullptr, pcm16.size(), pcm16.data(), options, out_metrics, t_start);
}

std::vector<TranscriptionResult> SttEngine::transcribe(const std::vector<float>& pcmf32, int input_sample_rate,
                                                       const RequestOptions& options,
                                                       PerformanceMetrics* out_metrics) {
  const auto t_start = Clock::now();
  if (input_sample_rate != 16000 && ctx_ && !pcmf32.empty() && input_sample_rate > 0) {
    // :138-145 resamples with libsamplerate (SRC_SINC_FASTEST); here the engine's own windowed-sinc
    // converter on the GPU (sw_resample_f32: same method, own window). Like the
    // reference, a failed conversion falls through to the unconverted samples (:141-144).
    std::vector<float> resampled((size_t)sw_resample_out_len((int64_t)pcmf32.size(), input_sample_rate, 16000));
    if (!resampled.empty() && sw_resample_f32(ctx_, pcmf32.data(), (int64_t)pcmf32.size(), input_sample_rate, 16000,
                                              resampled.data()) == 0)
      return run_request(resampled.data(), resampled.size(), nullptr, options, out_metrics, t_start);
    fprintf(stderr, "[stt_engine] resampling %d -> 16000 Hz failed: %s\n", input_sample_rate, sw_last_error());
  }
  return run_request(pcmf32.data(), pcmf32.size(), nullptr, options, out_metrics, t_start);
}

std::vector<TranscriptionResult> SttEngine::run_request(const float* pcm, size_t pcm_size, const int16_t* pcm16,
                                                        const RequestOptions& options,
                                                        PerformanceMetrics* out_metrics, Clock::time_point t_start) {
  if (!ctx_) return {};                                          // :132
  if (options.should_abort && options.should_abort()) return {};  // :133

  // :153-167 - too short to process
  const size_t min_samples = static_cast<size_t>((settings_.vad_ms_min_duration * 16000) / 1000);
  if (pcm_size < min_samples) {
    if (out_metrics) {
      out_metrics->queue_time_ms = 0;
      out_metrics->processing_time_ms = 0;
      out_metrics->token_count = 0;
    }
    return {};
  }
  // :169-194 - speech pre-gate (hook, see set_vad_fn): no speech -> the reference's placeholder result
  if (settings_.enable_vad && vad_fn_) {
    bool speech;
    {
      std::vector<float> tmp;
      const float* f = pcm;
      if (!f) {  // int16 caller: the gate sees what the reference's /32768 loop (:117-125) would hand it
        tmp.resize(pcm_size);
        for (size_t i = 0; i < pcm_size; ++i) tmp[i] = static_cast<float>(pcm16[i]) / 32768.0f;
        f = tmp.data();
      }
      std::lock_guard<std::mutex> lk(vad_mutex_);  // :110
      speech = vad_fn_(f, pcm_size);
    }
    if (!speech) {
      TranscriptionResult empty_res{};
      empty_res.text = "";
      empty_res.language = "unknown";
      empty_res.prob = 0.0f;
      empty_res.t0 = 0;
      empty_res.t1 = static_cast<int64_t>(pcm_size / 16.0);
      empty_res.speaker_turn_next = false;
      empty_res.token_count = 0;
      empty_res.affective = prosody_fn_ ? prosody_fn_(nullptr, 0, 16000, options.prosody_opts)
                                        : neutral_prosody(nullptr, 0, 16000, options.prosody_opts);
      empty_res.speaker_id = "unknown";
      if (out_metrics) {
        out_metrics->queue_time_ms = 0;
        out_metrics->processing_time_ms = std::chrono::duration<double, std::milli>(Clock::now() - t_start).count();
        out_metrics->token_count = 0;
      }
      return {empty_res};
    }
  }

  acquire_slot();  // throws EngineBusyException after request_queue_timeout_ms (:197-198)
  struct SlotGuard {
    SttEngine& e;
    ~SlotGuard() { e.release_slot(); }
  } guard{*this};
  const auto t_acquired = Clock::now();

  SpeakerClusterer clusterer(settings_.cluster_threshold);  // :202

  // :204-243 - parameter mapping
  const int active_beam = options.beam_size >= 0 ? options.beam_size : settings_.beam_size;
  const float active_temp = options.temperature >= 0.0f ? options.temperature : settings_.temperature;
  const int active_best_of = options.best_of >= 0 ? options.best_of : settings_.best_of;
  const int strategy = active_beam > 1 ? 1 : 0;
  Request req;
  req.pcm = pcm;
  req.pcm16 = pcm16;
  req.n = static_cast<int>(pcm_size);
  req.params = sw_full_default_params(strategy);
  req.abort_fn = options.should_abort;
  if (req.abort_fn) {
    req.params.abort_callback = abort_trampoline;
    req.params.abort_callback_user_data = &req.abort_fn;
  }
  req.params.token_timestamps = 1;
  req.params.suppress_nst = settings_.suppress_nst ? 1 : 0;
  req.params.no_speech_thold = settings_.no_speech_threshold;
  req.params.translate = options.translate ? 1 : 0;
  req.params.tdrz_enable = options.enable_diarization ? 1 : 0;
  req.language = options.language.empty() ? settings_.language : options.language;
  req.params.language = req.language.c_str();
  req.prompt = options.prompt;
  req.params.initial_prompt = req.prompt.empty() ? nullptr : req.prompt.c_str();
  req.params.temperature = active_temp;
  if (strategy == 1) req.params.beam_size = active_beam;
  else req.params.best_of = active_best_of;
  req.params.entropy_thold = 2.40f;
  req.params.logprob_thold = settings_.logprob_threshold;
  req.params.n_threads = settings_.n_threads;

  // :245-246 - whisper_full_with_state, here: join the next device pass
  {
    std::lock_guard<std::mutex> lk(q_mutex_);
    queue_.push_back(&req);
  }
  q_cv_.notify_all();
  {
    std::unique_lock<std::mutex> lk(req.m);
    req.cv.wait(lk, [&] { return req.done; });
  }
  const auto t_end = Clock::now();
  if (out_metrics) {  // :250-256
    out_metrics->queue_time_ms = std::chrono::duration<double, std::milli>(t_acquired - t_start).count();
    out_metrics->processing_time_ms = std::chrono::duration<double, std::milli>(t_end - t_acquired).count();
    out_metrics->token_count = 0;
  }

  std::vector<TranscriptionResult> results;
  if (req.rc != 0 || !req.result) {  // :341-346 - log and return nothing
    if (!(options.should_abort && options.should_abort()))
      fprintf(stderr, "[stt_engine] Whisper processing failed: %d (%s)\n", req.rc, req.error.c_str());
    return results;
  }
  struct ResultGuard {
    sw_result* r;
    ~ResultGuard() { sw_result_free(r); }
  } rguard{req.result};

  const float kMinAvgTokenProb = 0.40f;  // :264
  const int eot = sw_token_eot(ctx_);
  const int n_segments = sw_result_n_segments(req.result);  // :261
  // Pass 1 (:261-311): filters; the kept segments and their PCM slices (:313-320).
  struct Kept {
    std::string text;
    int64_t t0, t1, s0, s1;
    bool turn;
    std::vector<TokenData> tokens;
    int valid;
    float avg_prob;
  };
  std::vector<Kept> kept;
  for (int i = 0; i < n_segments; ++i) {
    const char* text_c = sw_result_segment_text(req.result, i);
    const std::string text = text_c ? text_c : "";
    if (sentiric::utils::is_hallucination(text)) continue;  // :272-278
    const int64_t t0 = sw_result_segment_t0(req.result, i), t1 = sw_result_segment_t1(req.result, i);
    const bool turn = sw_result_segment_speaker_turn_next(req.result, i) != 0;

    std::vector<TokenData> tokens;  // :285-296
    double total_prob = 0.0;
    int valid = 0;
    const int n_tokens = sw_result_n_tokens(req.result, i);
    for (int j = 0; j < n_tokens; ++j) {
      const sw_token_data d = sw_result_token_data(req.result, i, j);
      if (d.id >= eot) continue;
      tokens.push_back({std::string(sw_token_to_str(ctx_, d.id)), d.p, d.t0, d.t1});
      total_prob += d.p;
      ++valid;
    }
    if (out_metrics) out_metrics->token_count += valid;
    const float avg_prob = valid > 0 ? static_cast<float>(total_prob / valid) : 0.0f;
    if (avg_prob < kMinAvgTokenProb && valid > 0) continue;  // :300-311

    int64_t s0 = static_cast<int64_t>((static_cast<double>(t0) / 100.0) * 16000.0);
    int64_t s1 = static_cast<int64_t>((static_cast<double>(t1) / 100.0) * 16000.0);
    s0 = std::max<int64_t>(0, std::min<int64_t>(s0, (int64_t)pcm_size));
    s1 = std::max<int64_t>(s0, std::min<int64_t>(s1, (int64_t)pcm_size));
    kept.push_back({text, t0, t1, s0, s1, turn, std::move(tokens), valid, avg_prob});
  }

  // Pass 2 (:322-334): prosody of every kept segment. Default: all segments of the utterance in one
  // batched GPU call (sw_prosody_segments_*: bit-identical to the reference's extract_prosody); a
  // caller-supplied ProsodyFn (set_prosody_fn) is evaluated per segment on the host instead.
  std::vector<AffectiveTags> tags(kept.size());
  if (!prosody_fn_) {
    std::vector<int64_t> b(kept.size()), e(kept.size());
    for (size_t k = 0; k < kept.size(); ++k) b[k] = kept[k].s0, e[k] = kept[k].s1;
    std::vector<sw_prosody> out(kept.size());
    sw_prosody_opts po;
    po.lpf_alpha = options.prosody_opts.lpf_alpha;
    po.gender_threshold = options.prosody_opts.gender_threshold;
    po.min_pitch = options.prosody_opts.min_pitch;
    po.max_pitch = options.prosody_opts.max_pitch;
    const int rc = kept.empty() ? 0
                   : pcm ? sw_prosody_segments_f32(ctx_, pcm, (int64_t)pcm_size, 16000, b.data(), e.data(),
                                                   (int)kept.size(), &po, out.data())
                         : sw_prosody_segments_pcm16(ctx_, pcm16, (int64_t)pcm_size, 16000, b.data(), e.data(),
                                                     (int)kept.size(), &po, out.data());
    if (rc != 0) fprintf(stderr, "[stt_engine] prosody failed: %s\n", sw_last_error());
    static const char* kEmotion[] = {"neutral", "excited", "sad", "angry"};
    for (size_t k = 0; k < kept.size(); ++k) {
      AffectiveTags& t = tags[k];
      if (rc != 0) {
        t = neutral_prosody(nullptr, 0, 16000, options.prosody_opts);
        continue;
      }
      const sw_prosody& p = out[k];
      t.gender_proxy = std::string(1, p.gender);
      t.emotion_proxy = kEmotion[p.emotion & 3];
      t.arousal = p.arousal; t.valence = p.valence; t.pitch_mean = p.pitch_mean; t.pitch_std = p.pitch_std;
      t.energy_mean = p.energy_mean; t.energy_std = p.energy_std; t.spectral_centroid = p.spectral_centroid;
      t.zero_crossing_rate = p.zero_crossing_rate;
      t.speaker_vec.assign(p.speaker_vec, p.speaker_vec + 8);
    }
  } else {
    std::vector<float> slice;
    for (size_t k = 0; k < kept.size(); ++k) {
      const size_t seg = static_cast<size_t>(kept[k].s1 - kept[k].s0);
      if (seg < 160) {
        tags[k] = prosody_fn_(nullptr, 0, 16000, options.prosody_opts);
        continue;
      }
      const float* p = nullptr;
      if (pcm) {
        p = pcm + kept[k].s0;
      } else {  // prosody works on float samples: convert only the slices that are analysed
        slice.resize(seg);
        for (size_t q = 0; q < seg; ++q) slice[q] = static_cast<float>(pcm16[kept[k].s0 + q]) / 32768.0f;
        p = slice.data();
      }
      tags[k] = prosody_fn_(p, seg, 16000, options.prosody_opts);
    }
  }
  for (size_t k = 0; k < kept.size(); ++k) {
    Kept& s = kept[k];
    std::string spk = "?";  // :323, :330-332: segments of >= 160 samples are clustered
    if (s.s1 - s.s0 >= 160 && !tags[k].speaker_vec.empty()) spk = clusterer.assign_or_add(tags[k].speaker_vec);
    const AffectiveTags& pros = tags[k];
    results.push_back({s.text, req.language, s.avg_prob, s.t0, s.t1, s.turn, s.tokens, s.valid, pros.gender_proxy,
                       pros.emotion_proxy, pros.arousal, pros.valence, pros, spk});  // :336-339
  }
  return results;
}
