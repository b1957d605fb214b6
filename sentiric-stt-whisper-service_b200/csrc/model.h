// Whisper model in HBM: parsed from the legacy ggml .bin the reference loads through
// whisper_init_from_file_with_params (stt_engine.cpp:33; format: SURVEY.md A.2). Matrix weights
// are converted to bf16 on the device; biases, LayerNorm affine and positional embeddings stay f32.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include <string>
#include <unordered_map>
#include <vector>

namespace sw {

struct HParams {
  int n_vocab = 0, n_audio_ctx = 0, n_audio_state = 0, n_audio_head = 0, n_audio_layer = 0;
  int n_text_ctx = 0, n_text_state = 0, n_text_head = 0, n_text_layer = 0, n_mels = 0, ftype = 0;
};

struct Vocab {
  std::vector<std::string> id_to_token;
  std::unordered_map<std::string, int> token_to_id;
  int n_langs = 0;
  bool multilingual = false;
  int eot = 50256, sot = 50257, translate = 50357, transcribe = 50358, solm = 50359, prev = 50360,
      nosp = 50361, not_ = 50362, beg = 50363;
  int space = -1;                // id of " "
  std::vector<int> nst_ids;      // suppress_nst token ids
};

struct LayerNormW {
  float* g = nullptr;
  float* b = nullptr;
};
struct EncLayerW {
  LayerNormW ln1, ln2;
  __nv_bfloat16 *wqkv = nullptr, *wo = nullptr, *w1 = nullptr, *w2 = nullptr;
  float *bqkv = nullptr, *bo = nullptr, *b1 = nullptr, *b2 = nullptr;
};
struct DecLayerW {
  LayerNormW ln1, lnx, ln2;
  __nv_bfloat16 *wqkv = nullptr, *wo = nullptr, *wxq = nullptr, *wxkv = nullptr, *wxo = nullptr,
                *w1 = nullptr, *w2 = nullptr;
  float *bqkv = nullptr, *bo = nullptr, *bxq = nullptr, *bxkv = nullptr, *bxo = nullptr, *b1 = nullptr,
        *b2 = nullptr;
};

struct Model {
  HParams hp;
  Vocab vocab;
  // device
  float* filters = nullptr;                 // [n_mel][201]
  int2* filter_span = nullptr;              // [n_mel] first / one-past-last bin (multiples of 4) with a non-zero weight
  __nv_bfloat16* conv1_w = nullptr;         // [d][3][n_mel]  (K-major for the implicit GEMM)
  __nv_bfloat16* conv2_w = nullptr;         // [d][3][d]
  float *conv1_b = nullptr, *conv2_b = nullptr;
  float* enc_pos = nullptr;                 // [1500][d]
  std::vector<EncLayerW> enc;
  LayerNormW ln_post;
  float* dec_pos = nullptr;                 // [448][d]
  __nv_bfloat16* tok_emb = nullptr;         // [n_vocab][d]
  std::vector<DecLayerW> dec;
  LayerNormW dec_ln;
  // arena
  uint8_t* arena = nullptr;
  size_t arena_bytes = 0, arena_used = 0;
  size_t weight_bytes_decoder = 0;          // bytes streamed by one decode step (roofline accounting)

  ~Model();
};

// Returns nullptr on failure (sw_last_error()). Current CUDA device must be set.
Model* load_model(const char* path);
int lang_id(const char* lang);
const char* lang_code(int id);

}  // namespace sw
