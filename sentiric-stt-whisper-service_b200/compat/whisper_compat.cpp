// libwhisper_compat.so: the whisper.cpp C API subset of include/compat/whisper.h (+ libsamplerate's
// src_simple, include/compat/samplerate.h) implemented on the B200 engine's C ABI (include/sw_whisper.h).
// With it the reference's own src/stt_engine.cpp compiles and runs UNMODIFIED (INTEGRATION.md route B):
// every function below names the reference line that calls it.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "../../include/compat/samplerate.h"
#include "../../include/compat/whisper.h"
#include "../../include/sw_whisper.h"

struct whisper_context {
  sw_ctx* sw = nullptr;
};
struct whisper_state {  // the reference keeps a pool of these, one per request in flight (stt_engine.cpp:36-42)
  whisper_context* ctx = nullptr;
  sw_result* res = nullptr;
};
struct whisper_vad_context {
  int unused;
};

namespace {
std::atomic<sw_ctx*> g_last_ctx{nullptr};  // src_simple has no context argument: it uses the latest one

ggml_log_callback g_log_cb = nullptr;
void* g_log_user = nullptr;
void log_bridge(int level, const char* text, void*) {
  if (g_log_cb) g_log_cb(static_cast<ggml_log_level>(level), text, g_log_user);
}
int abort_bridge(void* user) {
  auto* p = static_cast<const whisper_full_params*>(user);
  return p->abort_callback && p->abort_callback(p->abort_callback_user_data) ? 1 : 0;
}
int env_int(const char* name, int def) {
  const char* v = getenv(name);
  return v ? atoi(v) : def;
}
}  // namespace

extern "C" {

whisper_context_params whisper_context_default_params(void) {
  whisper_context_params p;
  memset(&p, 0, sizeof(p));
  p.use_gpu = true;
  p.dtw_n_top = -1;
  p.dtw_mem_size = 1024 * 1024 * 128;
  return p;
}

whisper_context* whisper_init_from_file_with_params(const char* path_model, whisper_context_params params) {
  sw_ctx_params cp = sw_ctx_default_params();
  cp.device = params.gpu_device;
  cp.flash_attn = params.flash_attn ? 1 : 0;
  // one window per whisper_state and request: small batches (SW_COMPAT_MAX_BATCH raises it for direct users)
  cp.max_batch = env_int("SW_COMPAT_MAX_BATCH", 4);
  cp.max_beams = 8;
  cp.n_lanes = 1;
  sw_ctx* sw = sw_ctx_create(path_model, &cp);
  if (!sw) return nullptr;  // upstream convention: NULL on failure (stt_engine.cpp:34 then throws)
  whisper_context* c = new whisper_context();
  c->sw = sw;
  g_last_ctx.store(sw);
  return c;
}

whisper_state* whisper_init_state(whisper_context* ctx) {
  if (!ctx) return nullptr;
  whisper_state* s = new whisper_state();
  s->ctx = ctx;
  return s;
}

void whisper_free_state(whisper_state* state) {
  if (!state) return;
  if (state->res) sw_result_free(state->res);
  delete state;
}

void whisper_free(whisper_context* ctx) {
  if (!ctx) return;
  sw_ctx* expect = ctx->sw;
  g_last_ctx.compare_exchange_strong(expect, nullptr);
  sw_ctx_destroy(ctx->sw);
  delete ctx;
}

const char* whisper_token_to_str(whisper_context* ctx, whisper_token token) {
  return ctx ? sw_token_to_str(ctx->sw, token) : "";
}
whisper_token whisper_token_eot(whisper_context* ctx) { return ctx ? sw_token_eot(ctx->sw) : 0; }
int whisper_lang_id(const char* lang) { return sw_lang_id(lang); }

whisper_full_params whisper_full_default_params(whisper_sampling_strategy strategy) {
  whisper_full_params p;
  memset(&p, 0, sizeof(p));
  const sw_full_params d = sw_full_default_params(strategy == WHISPER_SAMPLING_BEAM_SEARCH ? 1 : 0);
  p.strategy = strategy;
  p.n_threads = d.n_threads;
  p.n_max_text_ctx = 16384;
  p.no_context = d.no_context != 0;
  p.print_progress = true;   // upstream defaults; the reference switches them off (stt_engine.cpp:220-223)
  p.print_timestamps = true;
  p.thold_pt = 0.01f;
  p.thold_ptsum = 0.01f;
  p.language = "en";
  p.suppress_blank = d.suppress_blank != 0;
  p.temperature = d.temperature;
  p.max_initial_ts = d.max_initial_ts;
  p.length_penalty = d.length_penalty;
  p.temperature_inc = d.temperature_inc;
  p.entropy_thold = d.entropy_thold;
  p.logprob_thold = d.logprob_thold;
  p.no_speech_thold = d.no_speech_thold;
  p.greedy.best_of = strategy == WHISPER_SAMPLING_GREEDY ? d.best_of : -1;
  p.beam_search.beam_size = strategy == WHISPER_SAMPLING_BEAM_SEARCH ? d.beam_size : -1;
  p.beam_search.patience = -1.0f;
  p.grammar_penalty = 100.0f;
  return p;
}

int whisper_full_with_state(whisper_context* ctx, whisper_state* state, whisper_full_params params,
                            const float* samples, int n_samples) {
  if (!ctx || !state) return -1;
  sw_full_params p = sw_full_default_params(params.strategy == WHISPER_SAMPLING_BEAM_SEARCH ? 1 : 0);
  p.beam_size = params.beam_search.beam_size;
  p.best_of = params.greedy.best_of;
  p.temperature = params.temperature;
  p.temperature_inc = params.temperature_inc;
  p.entropy_thold = params.entropy_thold;
  p.logprob_thold = params.logprob_thold;
  p.no_speech_thold = params.no_speech_thold;
  p.translate = params.translate;
  p.tdrz_enable = params.tdrz_enable;
  p.suppress_nst = params.suppress_nst;
  p.suppress_blank = params.suppress_blank;
  p.token_timestamps = params.token_timestamps;
  p.no_timestamps = params.no_timestamps;
  p.single_segment = params.single_segment;
  p.no_context = params.no_context;
  p.max_initial_ts = params.max_initial_ts;
  p.length_penalty = params.length_penalty;
  p.language = params.detect_language ? "auto" : params.language;
  p.initial_prompt = params.initial_prompt;
  p.prompt_tokens = params.prompt_tokens;
  p.prompt_n_tokens = params.prompt_n_tokens;
  p.n_threads = params.n_threads;
  if (params.abort_callback) {  // bool(void*) upstream, int(void*) in the C ABI
    p.abort_callback = abort_bridge;
    p.abort_callback_user_data = &params;
  }
  if (state->res) {
    sw_result_free(state->res);
    state->res = nullptr;
  }
  return sw_full(ctx->sw, &p, samples, n_samples, &state->res);
}

int whisper_full_n_segments_from_state(whisper_state* state) {
  return state && state->res ? sw_result_n_segments(state->res) : 0;
}
const char* whisper_full_get_segment_text_from_state(whisper_state* state, int i) {
  return state && state->res ? sw_result_segment_text(state->res, i) : nullptr;
}
int64_t whisper_full_get_segment_t0_from_state(whisper_state* state, int i) {
  return state && state->res ? sw_result_segment_t0(state->res, i) : 0;
}
int64_t whisper_full_get_segment_t1_from_state(whisper_state* state, int i) {
  return state && state->res ? sw_result_segment_t1(state->res, i) : 0;
}
bool whisper_full_get_segment_speaker_turn_next_from_state(whisper_state* state, int i) {
  return state && state->res && sw_result_segment_speaker_turn_next(state->res, i) != 0;
}
int whisper_full_n_tokens_from_state(whisper_state* state, int i) {
  return state && state->res ? sw_result_n_tokens(state->res, i) : 0;
}
whisper_token_data whisper_full_get_token_data_from_state(whisper_state* state, int i, int j) {
  whisper_token_data t;
  memset(&t, 0, sizeof(t));
  if (!state || !state->res) return t;
  const sw_token_data s = sw_result_token_data(state->res, i, j);
  t.id = s.id;
  t.tid = s.tid;
  t.p = s.p;
  t.plog = s.plog;
  t.pt = s.pt;
  t.ptsum = s.ptsum;
  t.t0 = s.t0;
  t.t1 = s.t1;
  t.t_dtw = s.t_dtw;
  t.vlen = s.vlen;
  return t;
}
int whisper_full_lang_id_from_state(whisper_state* state) {
  return state && state->res ? sw_result_lang_id(state->res) : -1;
}

whisper_vad_context_params whisper_vad_default_context_params(void) {
  whisper_vad_context_params p;
  p.n_threads = 4;
  p.use_gpu = false;
  p.gpu_device = 0;
  return p;
}
whisper_vad_context* whisper_vad_init_from_file_with_params(const char*, whisper_vad_context_params) {
  return nullptr;  // no Silero model / evaluator in this build: the reference then leaves its gate open (:109)
}
bool whisper_vad_detect_speech(whisper_vad_context*, const float*, int) { return true; }
void whisper_vad_free(whisper_vad_context*) {}

void whisper_log_set(ggml_log_callback cb, void* user) {
  g_log_cb = cb;
  g_log_user = user;
  sw_log_set(cb ? log_bridge : nullptr, nullptr);
}

// ---- libsamplerate's src_simple (stt_engine.cpp:87-106) on sw_resample_f32
int src_simple(SRC_DATA* d, int /*converter_type*/, int channels) {
  if (!d || channels != 1 || !d->data_in || !d->data_out || d->src_ratio <= 0.0) return 1;
  sw_ctx* ctx = g_last_ctx.load();
  if (!ctx) return 2;
  // the caller passes only the ratio target / source: recover integer rates (the reference always
  // converts TO 16 kHz, so try that first)
  int sr_out = 16000, sr_in = (int)llround(16000.0 / d->src_ratio);
  if (sr_in <= 0 || fabs((double)sr_out / sr_in - d->src_ratio) > 1e-9 * d->src_ratio) {
    sr_in = 1000000;
    sr_out = (int)llround(1000000.0 * d->src_ratio);
  }
  const int64_t n_out = sw_resample_out_len(d->input_frames, sr_in, sr_out);
  if (n_out > d->output_frames) return 3;
  if (n_out > 0 && sw_resample_f32(ctx, d->data_in, d->input_frames, sr_in, sr_out, d->data_out)) return 4;
  d->input_frames_used = d->input_frames;
  d->output_frames_gen = (long)n_out;
  return 0;
}
const char* src_strerror(int error) { return error ? "sw_whisper resampler error" : "no error"; }

}  // extern "C"
