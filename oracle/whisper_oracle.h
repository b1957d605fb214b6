/*
 * whisper_oracle.h - CPU ORACLE for the Whisper inference hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ may be imported, linked or
 * executed by the product path (sentiric-stt-whisper-service_b200/); only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, and only as the checker / the CPU arm.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in third-party
 * ggerganov/whisper.cpp @ v1.8.2 (pinned by /root/reference/Dockerfile:24-27,
 * Dockerfile.gpu:24-27, consumed by CMakeLists.txt:49), which is NOT vendored
 * under /root/reference, is nowhere on disk and cannot be fetched. The
 * reference holds no golden vectors, tests or fixtures for this path
 * (SURVEY.md §0.3). This file restates the published algorithm of that
 * dependency anchored on the reference's own call sites
 * in src/stt_engine.cpp. It remains UNPINNED AGAINST UPSTREAM ITSELF; what pins
 * it is the independent HuggingFace Whisper implementation:
 *   - log-mel, encoder output, teacher-forced logits: tests/golden/make_golden.py
 *     -> golden/micro_hf.npz (tests/test_oracle_golden.py);
 *   - the logit rules (process_logits: suppression sets, timestamp grammar,
 *     max_initial_ts, timestamp-mass rule) and greedy sequencing: HF's three
 *     logits processors and generate(), tests/golden/make_rules_golden.py ->
 *     golden/rules_hf.npz (tests/test_oracle_vs_hf_rules.py), with the two
 *     intentional whisper.cpp-vs-OpenAI differences asserted by name.
 * Recall-only (no independent counterpart here): seek / segment logic, beam
 * search draws, sequence scoring, fallback thresholds, token-level timestamps.
 */
#ifndef WHISPER_ORACLE_H
#define WHISPER_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ora_model ora_model;
typedef struct ora_result ora_result;

/* weight_round: 0 = weights as stored in the file (f16/f32), 1 = additionally rounded to bf16
 * (what the B200 engine holds in HBM). */
ora_model* ora_load(const char* ggml_path, int weight_round);
void ora_free(ora_model* m);
const char* ora_last_error(void);

/* numerics switches (every low-confidence upstream choice is a named switch) */
enum { ORA_ACT_F32 = 0, ORA_ACT_F16 = 1, ORA_ACT_BF16 = 2 };
void ora_set_act_round(ora_model* m, int mode);    /* rounding of matmul activations */
void ora_set_gelu_erf(ora_model* m, int use_erf);  /* 0: tanh GELU (whisper.cpp), 1: erf (HF) */
void ora_set_threads(ora_model* m, int n_threads);

typedef struct ora_hparams {
  int n_vocab, n_audio_ctx, n_audio_state, n_audio_head, n_audio_layer;
  int n_text_ctx, n_text_state, n_text_head, n_text_layer, n_mels, ftype;
  int token_eot, token_sot, token_translate, token_transcribe, token_solm;
  int token_prev, token_nosp, token_not, token_beg;
  int is_multilingual;
} ora_hparams;
void ora_get_hparams(const ora_model* m, ora_hparams* out);
const char* ora_token_to_str(const ora_model* m, int id);
int ora_lang_id(const char* lang);
/* whisper_tokenize restatement (greedy longest match over the vocabulary) */
int ora_tokenize(const ora_model* m, const char* text, int32_t* out, int max_tokens);

/* --- stages ------------------------------------------------------------ */
/* log_mel_spectrogram (A.3). out: [n_mel][n_len] f32 or NULL to query sizes. */
int ora_mel(const ora_model* m, const float* pcm, int n_samples, float* out, int* n_len,
            int* n_len_org);
/* whisper_encode_internal (A.4) on one window mel[n_mel][3000]; out [1500][d]; also fills the
 * model's cross-KV cache used by ora_decode. */
int ora_encode(ora_model* m, const float* mel_window, float* enc_out);
/* intermediate taps of the last ora_encode (for kernel-level parity): 0 = conv stem + pos
 * [1500][d], 1.. = output of encoder block i-1. Returns number of floats or <0. */
int ora_encode_tap(const ora_model* m, int which, float* out);
/* whisper_decode_internal (A.5): feeds n_tokens tokens at positions n_past.. of decoder slot
 * `slot` (slots have private self-KV). logits_out [n_tokens][n_vocab] or NULL. */
int ora_decode(ora_model* m, int slot, const int32_t* tokens, int n_tokens, int n_past,
               float* logits_out);

/* --- whole path: whisper_full_with_state (A.6), call site stt_engine.cpp:245 --- */
typedef struct ora_full_params {
  int strategy; /* 0 greedy, 1 beam */
  int beam_size, best_of;
  float temperature, temperature_inc;
  float entropy_thold, logprob_thold, no_speech_thold;
  int translate, tdrz_enable, suppress_nst, suppress_blank;
  int token_timestamps, no_timestamps, single_segment, no_context;
  float max_initial_ts, length_penalty;
  const char* language;
  const char* initial_prompt;
  const int32_t* prompt_tokens;
  int prompt_n_tokens;
  int max_tokens;
} ora_full_params;
ora_full_params ora_full_default_params(int strategy);

typedef struct ora_token_data {
  int32_t id, tid;
  float p, plog, pt, ptsum;
  int64_t t0, t1, t_dtw;
  float vlen;
} ora_token_data;

int ora_full(ora_model* m, const ora_full_params* p, const float* pcm, int n_samples,
             ora_result** out);
int ora_result_n_segments(const ora_result* r);
const char* ora_result_segment_text(const ora_result* r, int i);
int64_t ora_result_segment_t0(const ora_result* r, int i);
int64_t ora_result_segment_t1(const ora_result* r, int i);
int ora_result_segment_speaker_turn_next(const ora_result* r, int i);
int ora_result_n_tokens(const ora_result* r, int i);
ora_token_data ora_result_token_data(const ora_result* r, int i, int j);
int ora_result_lang_id(const ora_result* r);
int ora_result_n_decode_steps(const ora_result* r);
int ora_result_n_windows(const ora_result* r);
double ora_result_ms_mel(const ora_result* r);
double ora_result_ms_encode(const ora_result* r);
double ora_result_ms_decode(const ora_result* r);
void ora_result_free(ora_result* r);

/* logit rules on one row (whisper_process_logits), exposed for unit parity of the device kernel:
 * tokens_cur: tokens sampled so far in this window; returns logprobs/probs. */
void ora_process_logits(const ora_model* m, const ora_full_params* p, const int32_t* tokens_cur,
                        int n_cur, int has_ts, int seek_delta, float temperature,
                        const float* logits_in, float* logits_out, float* logprobs, float* probs);

#ifdef __cplusplus
}
#endif
#endif
