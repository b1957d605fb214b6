/*
 * whisper.h - source-compatibility shim (INTEGRATION.md route B).
 *
 * The reference reaches its whole inference path through whisper.cpp's C API from one file,
 * /root/reference/src/stt_engine.cpp (+ whisper_log_set in main.cpp:71). whisper.cpp v1.8.2 itself is NOT
 * part of the reference tree (it is cloned at Docker build time, Dockerfile:24-27). This header restates the
 * part of that published API the reference binds - the 21 functions, 3 parameter structs, the token record
 * and the enums listed in SURVEY.md §8(b) - so that the reference's stt_engine.cpp compiles UNMODIFIED and
 * links against libwhisper_compat.so, which forwards every call to the B200 engine's C ABI
 * (include/sw_whisper.h). Members the reference never touches are declared for source compatibility and
 * ignored; members it sets are honoured (list at whisper_full_params below).
 *
 * This is a compatibility route, not the fast one: each whisper_state runs its own device pass, as the
 * reference's design prescribes. The batching facade is host/stt_engine.cpp (route A).
 */
#ifndef WHISPER_H
#define WHISPER_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WHISPER_API __attribute__((visibility("default")))
#define WHISPER_SAMPLE_RATE 16000
#define WHISPER_N_FFT 400
#define WHISPER_HOP_LENGTH 160
#define WHISPER_CHUNK_SIZE 30

struct whisper_context;
struct whisper_state;
struct whisper_vad_context;

typedef int32_t whisper_pos;
typedef int32_t whisper_token;
typedef int32_t whisper_seq_id;

/* ggml.h: the log levels main.cpp:42-48 switches on, and the callback types */
enum ggml_log_level {
  GGML_LOG_LEVEL_NONE = 0,
  GGML_LOG_LEVEL_DEBUG = 1,
  GGML_LOG_LEVEL_INFO = 2,
  GGML_LOG_LEVEL_WARN = 3,
  GGML_LOG_LEVEL_ERROR = 4,
  GGML_LOG_LEVEL_CONT = 5,
};
typedef void (*ggml_log_callback)(enum ggml_log_level level, const char* text, void* user_data);
typedef bool (*ggml_abort_callback)(void* data);

enum whisper_alignment_heads_preset { WHISPER_AHEADS_NONE = 0 };
typedef struct whisper_ahead {
  int n_text_layer;
  int n_head;
} whisper_ahead;
typedef struct whisper_aheads {
  size_t n_heads;
  const whisper_ahead* heads;
} whisper_aheads;

struct whisper_context_params { /* stt_engine.cpp:28-31 sets use_gpu, flash_attn */
  bool use_gpu;
  bool flash_attn;
  int gpu_device;
  bool dtw_token_timestamps;
  enum whisper_alignment_heads_preset dtw_aheads_preset;
  int dtw_n_top;
  struct whisper_aheads dtw_aheads;
  size_t dtw_mem_size;
};

typedef struct whisper_token_data { /* stt_engine.cpp:290-293 reads id, p, t0, t1 */
  whisper_token id;
  whisper_token tid;
  float p;
  float plog;
  float pt;
  float ptsum;
  int64_t t0;
  int64_t t1;
  int64_t t_dtw;
  float vlen;
} whisper_token_data;

WHISPER_API struct whisper_context_params whisper_context_default_params(void);            /* :28 */
WHISPER_API struct whisper_context* whisper_init_from_file_with_params(const char* path_model,
                                                                       struct whisper_context_params params); /* :33 */
WHISPER_API struct whisper_state* whisper_init_state(struct whisper_context* ctx);         /* :39 */
WHISPER_API void whisper_free_state(struct whisper_state* state);                           /* :56 */
WHISPER_API void whisper_free(struct whisper_context* ctx);                                 /* :57 */

WHISPER_API const char* whisper_token_to_str(struct whisper_context* ctx, whisper_token token); /* :291 */
WHISPER_API whisper_token whisper_token_eot(struct whisper_context* ctx);                    /* :292 */
WHISPER_API int whisper_lang_id(const char* lang);

enum whisper_sampling_strategy {
  WHISPER_SAMPLING_GREEDY,      /* stt_engine.cpp:212 */
  WHISPER_SAMPLING_BEAM_SEARCH, /* stt_engine.cpp:211 */
};

typedef void (*whisper_new_segment_callback)(struct whisper_context* ctx, struct whisper_state* state, int n_new,
                                             void* user_data);
typedef void (*whisper_progress_callback)(struct whisper_context* ctx, struct whisper_state* state, int progress,
                                          void* user_data);
typedef bool (*whisper_encoder_begin_callback)(struct whisper_context* ctx, struct whisper_state* state,
                                               void* user_data);
typedef void (*whisper_logits_filter_callback)(struct whisper_context* ctx, struct whisper_state* state,
                                               const whisper_token_data* tokens, int n_tokens, float* logits,
                                               void* user_data);
typedef struct whisper_grammar_element {
  int type;
  uint32_t value;
} whisper_grammar_element;

typedef struct whisper_vad_params {
  float threshold;
  int min_speech_duration_ms;
  int min_silence_duration_ms;
  float max_speech_duration_s;
  int speech_pad_ms;
  float samples_overlap;
} whisper_vad_params;

/* Honoured (everything stt_engine.cpp:217-243 sets): strategy, n_threads, translate, token_timestamps,
 * tdrz_enable, initial_prompt, prompt_tokens / prompt_n_tokens, language, suppress_blank, suppress_nst,
 * temperature, max_initial_ts, length_penalty, temperature_inc, entropy_thold, logprob_thold,
 * no_speech_thold, greedy.best_of, beam_search.beam_size, no_context, no_timestamps, single_segment,
 * abort_callback / abort_callback_user_data. The print_* flags are accepted (the engine never prints). */
struct whisper_full_params {
  enum whisper_sampling_strategy strategy;
  int n_threads;
  int n_max_text_ctx;
  int offset_ms;
  int duration_ms;
  bool translate;
  bool no_context;
  bool no_timestamps;
  bool single_segment;
  bool print_special;
  bool print_progress;
  bool print_realtime;
  bool print_timestamps;
  bool token_timestamps;
  float thold_pt;
  float thold_ptsum;
  int max_len;
  bool split_on_word;
  int max_tokens;
  bool debug_mode;
  int audio_ctx;
  bool tdrz_enable;
  const char* suppress_regex;
  const char* initial_prompt;
  const whisper_token* prompt_tokens;
  int prompt_n_tokens;
  const char* language;
  bool detect_language;
  bool suppress_blank;
  bool suppress_nst;
  float temperature;
  float max_initial_ts;
  float length_penalty;
  float temperature_inc;
  float entropy_thold;
  float logprob_thold;
  float no_speech_thold;
  struct {
    int best_of;
  } greedy;
  struct {
    int beam_size;
    float patience;
  } beam_search;
  whisper_new_segment_callback new_segment_callback;
  void* new_segment_callback_user_data;
  whisper_progress_callback progress_callback;
  void* progress_callback_user_data;
  whisper_encoder_begin_callback encoder_begin_callback;
  void* encoder_begin_callback_user_data;
  ggml_abort_callback abort_callback;
  void* abort_callback_user_data;
  whisper_logits_filter_callback logits_filter_callback;
  void* logits_filter_callback_user_data;
  const whisper_grammar_element** grammar_rules;
  size_t n_grammar_rules;
  size_t i_start_rule;
  float grammar_penalty;
  bool vad;
  const char* vad_model_path;
  whisper_vad_params vad_params;
};

WHISPER_API struct whisper_full_params whisper_full_default_params(enum whisper_sampling_strategy strategy); /* :214 */
WHISPER_API int whisper_full_with_state(struct whisper_context* ctx, struct whisper_state* state,
                                        struct whisper_full_params params, const float* samples,
                                        int n_samples);                                               /* :245 */

WHISPER_API int whisper_full_n_segments_from_state(struct whisper_state* state);                       /* :261 */
WHISPER_API const char* whisper_full_get_segment_text_from_state(struct whisper_state* state, int i_segment); /* :267 */
WHISPER_API int64_t whisper_full_get_segment_t0_from_state(struct whisper_state* state, int i_segment); /* :280 */
WHISPER_API int64_t whisper_full_get_segment_t1_from_state(struct whisper_state* state, int i_segment); /* :281 */
WHISPER_API bool whisper_full_get_segment_speaker_turn_next_from_state(struct whisper_state* state,
                                                                       int i_segment);                  /* :283 */
WHISPER_API int whisper_full_n_tokens_from_state(struct whisper_state* state, int i_segment);           /* :286 */
WHISPER_API whisper_token_data whisper_full_get_token_data_from_state(struct whisper_state* state, int i_segment,
                                                                      int i_token);                     /* :290 */
WHISPER_API int whisper_full_lang_id_from_state(struct whisper_state* state);

/* Silero VAD through whisper.cpp (stt_engine.cpp:44-52,108-115). Neither that code nor a model file is part
 * of this build: init returns NULL, which the reference treats as "gate open" (stt_engine.cpp:109). */
struct whisper_vad_context_params {
  int n_threads;
  bool use_gpu;
  int gpu_device;
};
WHISPER_API struct whisper_vad_context_params whisper_vad_default_context_params(void);                /* :48 */
WHISPER_API struct whisper_vad_context* whisper_vad_init_from_file_with_params(
    const char* path_model, struct whisper_vad_context_params params);                                /* :51 */
WHISPER_API bool whisper_vad_detect_speech(struct whisper_vad_context* vctx, const float* samples,
                                           int n_samples);                                             /* :114 */
WHISPER_API void whisper_vad_free(struct whisper_vad_context* ctx);                                    /* :58 */

WHISPER_API void whisper_log_set(ggml_log_callback log_callback, void* user_data);                    /* main.cpp:71 */

#ifdef __cplusplus
}
#endif
#endif /* WHISPER_H */
