/*
 * whisper_oracle.cpp - CPU ORACLE for the Whisper inference hot path (see whisper_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY - never linked or loaded by the product path.
 * PARITY UNPINNED (whisper.cpp v1.8.2 is not vendored under /root/reference; see header).
 *
 * Every function names the upstream routine it restates (whisper.cpp v1.8.2, src/whisper.cpp,
 * summarised in SURVEY.md Appendix A) and the reference call site that reaches it
 * (/root/reference/src/stt_engine.cpp).
 */
#include "whisper_oracle.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

thread_local char g_err[512] = "";
void set_err(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---------------------------------------------------------------------------------------------
// scalar format helpers
// ---------------------------------------------------------------------------------------------
inline float f16_to_f32(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000) << 16;
  uint32_t exp = (h >> 10) & 0x1f;
  uint32_t man = h & 0x3ff;
  uint32_t out;
  if (exp == 0) {
    if (man == 0) {
      out = sign;
    } else {
      int e = -1;
      do {
        man <<= 1;
        ++e;
      } while (!(man & 0x400));
      out = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ff) << 13);
    }
  } else if (exp == 31) {
    out = sign | 0x7f800000u | (man << 13);
  } else {
    out = sign | ((exp + 112) << 23) | (man << 13);
  }
  float f;
  memcpy(&f, &out, 4);
  return f;
}
inline float round_f16(float f) {  // f32 -> f16 (RNE) -> f32
  _Float16 h = (_Float16)f;
  return (float)h;
}
inline float round_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return f;
  u += 0x7fffu + ((u >> 16) & 1u);
  u &= 0xffff0000u;
  memcpy(&f, &u, 4);
  return f;
}
inline float round_mode(float f, int mode) {
  return mode == ORA_ACT_F16 ? round_f16(f) : mode == ORA_ACT_BF16 ? round_bf16(f) : f;
}
void round_buf(const float* in, float* out, size_t n, int mode) {
  if (mode == ORA_ACT_F32) {
    if (in != out) memcpy(out, in, n * sizeof(float));
    return;
  }
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)n; ++i) out[i] = round_mode(in[i], mode);
}

// ---------------------------------------------------------------------------------------------
// C[M][N] = A[M][K] * W[N][K]^T (+ bias[N]); both operands K-major, f32 accumulate.
// Restates ggml mul_mat as whisper.cpp uses it (weights f16 x activations, f32 out).
// ---------------------------------------------------------------------------------------------
typedef float v8f __attribute__((vector_size(32), aligned(4)));
inline float hsum(v8f v) {
  float s = 0;
  for (int i = 0; i < 8; ++i) s += v[i];
  return s;
}
template <int MR>
inline void dot_block(const float* A, int lda, const float* W, int ldw, int K, float* acc /*MR*4*/) {
  v8f c[MR][4];
  for (int i = 0; i < MR; ++i)
    for (int j = 0; j < 4; ++j) c[i][j] = v8f{0, 0, 0, 0, 0, 0, 0, 0};
  int k = 0;
  for (; k + 8 <= K; k += 8) {
    v8f w0 = *(const v8f*)(W + k), w1 = *(const v8f*)(W + ldw + k);
    v8f w2 = *(const v8f*)(W + 2 * ldw + k), w3 = *(const v8f*)(W + 3 * ldw + k);
    for (int i = 0; i < MR; ++i) {
      v8f a = *(const v8f*)(A + (size_t)i * lda + k);
      c[i][0] += a * w0;
      c[i][1] += a * w1;
      c[i][2] += a * w2;
      c[i][3] += a * w3;
    }
  }
  for (int i = 0; i < MR; ++i)
    for (int j = 0; j < 4; ++j) {
      float s = hsum(c[i][j]);
      for (int kk = k; kk < K; ++kk) s += A[(size_t)i * lda + kk] * W[(size_t)j * ldw + kk];
      acc[i * 4 + j] = s;
    }
}
void matmul_nt(const float* A, int M, int K, int lda, const float* W, int N, const float* bias,
               float* C, int ldc) {
  const int NB = 4;
  const int n_blocks = (N + NB - 1) / NB;
  const int m_blocks = (M + 2) / 3;
  if (M >= 12) {
    // tile so that a panel of W (64 rows) stays in L2 while M streams past
    const int NP = 64;
    const int n_panels = (N + NP - 1) / NP;
    const int MP = 96;
    const int m_panels = (M + MP - 1) / MP;
#pragma omp parallel for collapse(2) schedule(dynamic)
    for (int mp = 0; mp < m_panels; ++mp)
      for (int np = 0; np < n_panels; ++np) {
        const int m0 = mp * MP, m1 = std::min(M, m0 + MP);
        const int n0 = np * NP, n1 = std::min(N, n0 + NP);
        for (int n = n0; n < n1; n += 4) {
          float wtmp_acc[12];
          const int nn = std::min(4, n1 - n);
          const float* Wp = W + (size_t)n * K;
          std::vector<float> wpad;
          if (nn < 4) {
            wpad.assign((size_t)4 * K, 0.f);
            memcpy(wpad.data(), Wp, (size_t)nn * K * sizeof(float));
            Wp = wpad.data();
          }
          for (int m = m0; m < m1; m += 3) {
            const int mm = std::min(3, m1 - m);
            if (mm == 3)
              dot_block<3>(A + (size_t)m * lda, lda, Wp, K, K, wtmp_acc);
            else if (mm == 2)
              dot_block<2>(A + (size_t)m * lda, lda, Wp, K, K, wtmp_acc);
            else
              dot_block<1>(A + (size_t)m * lda, lda, Wp, K, K, wtmp_acc);
            for (int i = 0; i < mm; ++i)
              for (int j = 0; j < nn; ++j)
                C[(size_t)(m + i) * ldc + n + j] = wtmp_acc[i * 4 + j] + (bias ? bias[n + j] : 0.f);
          }
        }
      }
    (void)m_blocks;
    (void)n_blocks;
    return;
  }
#pragma omp parallel for schedule(static)
  for (int nb = 0; nb < n_blocks; ++nb) {
    const int n = nb * NB;
    const int nn = std::min(4, N - n);
    const float* Wp = W + (size_t)n * K;
    std::vector<float> wpad;
    if (nn < 4) {
      wpad.assign((size_t)4 * K, 0.f);
      memcpy(wpad.data(), Wp, (size_t)nn * K * sizeof(float));
      Wp = wpad.data();
    }
    float acc[12];
    for (int m = 0; m < M; m += 3) {
      const int mm = std::min(3, M - m);
      if (mm == 3)
        dot_block<3>(A + (size_t)m * lda, lda, Wp, K, K, acc);
      else if (mm == 2)
        dot_block<2>(A + (size_t)m * lda, lda, Wp, K, K, acc);
      else
        dot_block<1>(A + (size_t)m * lda, lda, Wp, K, K, acc);
      for (int i = 0; i < mm; ++i)
        for (int j = 0; j < nn; ++j)
          C[(size_t)(m + i) * ldc + n + j] = acc[i * 4 + j] + (bias ? bias[n + j] : 0.f);
    }
  }
}

inline float gelu_tanh(float x) {  // ggml_gelu_f32 (tanh approximation)
  return 0.5f * x * (1.0f + tanhf(0.79788456080286535587989211986876f * x * (1.0f + 0.044715f * x * x)));
}
inline float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// ggml_norm + affine (whisper.cpp: ggml_norm(eps 1e-5), mul gamma, add beta)
void layer_norm(const float* x, int rows, int d, const float* g, const float* b, float* out) {
#pragma omp parallel for schedule(static)
  for (int r = 0; r < rows; ++r) {
    const float* xr = x + (size_t)r * d;
    float* o = out + (size_t)r * d;
    double sum = 0.0;
    for (int i = 0; i < d; ++i) sum += xr[i];
    const float mean = (float)(sum / d);
    double sum2 = 0.0;
    for (int i = 0; i < d; ++i) {
      const float v = xr[i] - mean;
      sum2 += (double)v * v;
    }
    const float scale = 1.0f / sqrtf((float)(sum2 / d) + 1e-5f);
    for (int i = 0; i < d; ++i) o[i] = (xr[i] - mean) * scale * g[i] + b[i];
  }
}

struct Tensor {
  std::vector<float> d;
  int ne[4] = {1, 1, 1, 1};  // ggml order (ne[0] contiguous)
  int nd = 0;
  int ttype = 0;
};

struct AttnW {
  const float *ln_g, *ln_b, *qw, *qb, *kw, *vw, *vb, *ow, *ob;
};
struct MlpW {
  const float *ln_g, *ln_b, *w0, *b0, *w2, *b2;
};
struct EncLayer {
  AttnW attn;
  MlpW mlp;
};
struct DecLayer {
  AttnW self, cross;
  MlpW mlp;
};

const char* const k_langs[] = {
    "en", "zh", "de", "es", "ru", "ko", "fr", "ja", "pt", "tr", "pl", "ca", "nl", "ar", "sv", "it",
    "id", "hi", "fi", "vi", "he", "uk", "el", "ms", "cs", "ro", "da", "hu", "ta", "no", "th", "ur",
    "hr", "bg", "lt", "la", "mi", "ml", "cy", "sk", "te", "fa", "lv", "bn", "sr", "az", "sl", "kn",
    "et", "mk", "br", "eu", "is", "hy", "ne", "mn", "bs", "kk", "sq", "sw", "gl", "mr", "pa", "si",
    "km", "sn", "yo", "so", "af", "oc", "ka", "be", "tg", "sd", "gu", "am", "yi", "lo", "uz", "fo",
    "ht", "ps", "tk", "nn", "mt", "sa", "lb", "my", "bo", "tl", "mg", "as", "tt", "haw", "ln", "ha",
    "ba", "jw", "su", "yue"};
const int k_n_langs = 100;

// the strings whisper_process_logits matches against the vocabulary when suppress_nst is set
const char* const k_non_speech[] = {
    "\"", "#", "(", ")", "*", "+", "/", ":", ";", "<", "=", ">", "@", "[", "\\", "]", "^", "_", "`",
    "{", "|", "}", "~", "「", "」", "『", "』", "<<", ">>", "<<<", ">>>", "--", "---", "-(", "-[",
    "('", "(\"", "((", "))", "(((", ")))", "[[", "]]", "{{", "}}", "♪♪", "♪♪♪", "♩", "♪", "♫", "♬",
    "♭", "♮", "♯"};

constexpr int MAX_SLOTS = 16;  // decoder slots (beam/best_of decoders + scratch copies)

}  // namespace

struct ora_model {
  ora_hparams hp;
  int n_langs = 0;
  std::vector<float> filters;  // [n_mel][201]
  std::vector<std::string> id_to_token;
  std::unordered_map<std::string, int> token_to_id;
  std::map<std::string, Tensor> tensors;
  std::vector<EncLayer> enc;
  std::vector<DecLayer> dec;
  const float *enc_pos = nullptr, *conv1_w = nullptr, *conv1_b = nullptr, *conv2_w = nullptr,
              *conv2_b = nullptr, *ln_post_g = nullptr, *ln_post_b = nullptr;
  const float *dec_pos = nullptr, *tok_emb = nullptr, *dec_ln_g = nullptr, *dec_ln_b = nullptr;
  std::vector<float> conv1_wr, conv2_wr;  // conv weights re-laid-out [d_out][k][c_in]
  std::vector<int> nst_ids;               // suppress_nst token ids
  int tok_space = -1;                     // id of " "

  int act_round = ORA_ACT_F16;
  int gelu_erf_flag = 0;
  int n_threads = 0;

  // state of the last ora_encode
  std::vector<std::vector<float>> cross_k, cross_v;  // per decoder layer [1500][d]
  std::vector<std::vector<float>> taps;              // [0] stem, [1+i] block i
  // decoder self-KV: [slot][layer] -> [n_text_ctx][d]
  std::vector<std::vector<std::vector<float>>> self_k, self_v;

  // token-level timestamp state (whisper_state::t_beg/t_last/tid_last/energy)
  std::vector<float> energy;
  int64_t t_beg = 0, t_last = 0;
  int tid_last = 0;

  const float* T(const std::string& name) const {
    auto it = tensors.find(name);
    if (it == tensors.end()) return nullptr;
    return it->second.d.data();
  }
};

struct ora_segment {
  int64_t t0, t1;
  std::string text;
  std::vector<ora_token_data> tokens;
  bool speaker_turn_next;
};
struct ora_result {
  std::vector<ora_segment> segs;
  int lang_id = -1;
  int n_decode_steps = 0;
  int n_windows = 0;
  double ms_mel = 0, ms_encode = 0, ms_decode = 0;
};

namespace {

double now_ms() {
  return std::chrono::duration<double, std::milli>(
             std::chrono::steady_clock::now().time_since_epoch())
      .count();
}

struct ThreadScope {
  int prev = 0;
  explicit ThreadScope(int n) {
#ifdef _OPENMP
    prev = omp_get_max_threads();
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
  }
  ~ThreadScope() {
#ifdef _OPENMP
    omp_set_num_threads(prev);
#endif
  }
};

// ---------------------------------------------------------------------------------------------
// whisper_model_load (legacy ggml .bin). Reached from
// whisper_init_from_file_with_params, stt_engine.cpp:33.
// ---------------------------------------------------------------------------------------------
bool read_exact(FILE* f, void* p, size_t n) { return fread(p, 1, n, f) == n; }

bool load_model(ora_model* m, const char* path, int weight_round) {
  FILE* f = fopen(path, "rb");
  if (!f) {
    set_err("cannot open %s", path);
    return false;
  }
  uint32_t magic = 0;
  if (!read_exact(f, &magic, 4) || magic != 0x67676d6c) {
    set_err("bad magic in %s", path);
    fclose(f);
    return false;
  }
  int32_t h[11];
  if (!read_exact(f, h, sizeof(h))) {
    set_err("truncated hparams");
    fclose(f);
    return false;
  }
  ora_hparams& hp = m->hp;
  hp.n_vocab = h[0];
  hp.n_audio_ctx = h[1];
  hp.n_audio_state = h[2];
  hp.n_audio_head = h[3];
  hp.n_audio_layer = h[4];
  hp.n_text_ctx = h[5];
  hp.n_text_state = h[6];
  hp.n_text_head = h[7];
  hp.n_text_layer = h[8];
  hp.n_mels = h[9];
  hp.ftype = h[10];
  int32_t fm[2];
  if (!read_exact(f, fm, 8) || fm[0] != hp.n_mels || fm[1] != 201) {
    set_err("bad mel filter header");
    fclose(f);
    return false;
  }
  m->filters.resize((size_t)fm[0] * fm[1]);
  if (!read_exact(f, m->filters.data(), m->filters.size() * 4)) {
    set_err("truncated filters");
    fclose(f);
    return false;
  }
  int32_t n_vocab_file = 0;
  read_exact(f, &n_vocab_file, 4);
  m->id_to_token.resize(hp.n_vocab);
  for (int i = 0; i < n_vocab_file; ++i) {
    uint32_t len = 0;
    if (!read_exact(f, &len, 4)) {
      set_err("truncated vocab");
      fclose(f);
      return false;
    }
    std::string s(len, '\0');
    if (len && !read_exact(f, &s[0], len)) {
      set_err("truncated vocab");
      fclose(f);
      return false;
    }
    if (i < hp.n_vocab) {
      m->id_to_token[i] = s;
      m->token_to_id[s] = i;
    }
  }
  // special-token ids (whisper_vocab; multilingual shift dt)
  hp.is_multilingual = hp.n_vocab >= 51865;
  hp.token_eot = 50256;
  hp.token_sot = 50257;
  hp.token_translate = 50357;
  hp.token_transcribe = 50358;
  hp.token_solm = 50359;
  hp.token_prev = 50360;
  hp.token_nosp = 50361;
  hp.token_not = 50362;
  hp.token_beg = 50363;
  m->n_langs = 0;
  if (hp.is_multilingual) {
    m->n_langs = hp.n_vocab - 51765 - 1;
    const int dt = m->n_langs - 98;
    hp.token_eot++;
    hp.token_sot++;
    hp.token_translate += dt;
    hp.token_transcribe += dt;
    hp.token_solm += dt;
    hp.token_prev += dt;
    hp.token_nosp += dt;
    hp.token_not += dt;
    hp.token_beg += dt;
  }
  // names for the tokens the file does not list (upstream synthesises them the same way)
  for (int i = n_vocab_file; i < hp.n_vocab; ++i) {
    std::string w;
    char buf[64];
    if (i > hp.token_beg) {
      snprintf(buf, sizeof(buf), "[_TT_%d]", i - hp.token_beg);
      w = buf;
    } else if (i == hp.token_eot)
      w = "[_EOT_]";
    else if (i == hp.token_sot)
      w = "[_SOT_]";
    else if (i == hp.token_translate)
      w = "[_TRANSLATE_]";
    else if (i == hp.token_transcribe)
      w = "[_TRANSCRIBE_]";
    else if (i == hp.token_solm)
      w = "[_SOLM_]";
    else if (i == hp.token_prev)
      w = "[_PREV_]";
    else if (i == hp.token_nosp)
      w = "[_NOSP_]";
    else if (i == hp.token_not)
      w = "[_NOT_]";
    else if (i == hp.token_beg)
      w = "[_BEG_]";
    else if (i > hp.token_sot && i <= hp.token_sot + m->n_langs) {
      snprintf(buf, sizeof(buf), "[_LANG_%s]", k_langs[i - hp.token_sot - 1]);
      w = buf;
    } else {
      snprintf(buf, sizeof(buf), "[_extra_token_%d]", i);
      w = buf;
    }
    m->id_to_token[i] = w;
    m->token_to_id[w] = i;
  }
  // tensors
  while (true) {
    int32_t hd[3];
    if (fread(hd, 1, 12, f) != 12) break;
    const int n_dims = hd[0], name_len = hd[1], ttype = hd[2];
    if (n_dims < 1 || n_dims > 4 || name_len <= 0 || name_len > 256) {
      set_err("corrupt tensor header");
      fclose(f);
      return false;
    }
    Tensor t;
    t.nd = n_dims;
    t.ttype = ttype;
    size_t nel = 1;
    for (int i = 0; i < n_dims; ++i) {
      int32_t v;
      read_exact(f, &v, 4);
      t.ne[i] = v;
      nel *= (size_t)v;
    }
    std::string name(name_len, '\0');
    read_exact(f, &name[0], name_len);
    t.d.resize(nel);
    if (ttype == 0) {
      if (!read_exact(f, t.d.data(), nel * 4)) {
        set_err("truncated tensor %s", name.c_str());
        fclose(f);
        return false;
      }
    } else if (ttype == 1) {
      std::vector<uint16_t> tmp(nel);
      if (!read_exact(f, tmp.data(), nel * 2)) {
        set_err("truncated tensor %s", name.c_str());
        fclose(f);
        return false;
      }
      for (size_t i = 0; i < nel; ++i) t.d[i] = f16_to_f32(tmp[i]);
    } else {
      set_err("tensor %s: quantised ggml type %d not supported by the oracle", name.c_str(), ttype);
      fclose(f);
      return false;
    }
    // the B200 engine keeps every >=2-D matrix weight as bf16 in HBM
    const bool is_matrix = n_dims >= 2 && name.find("positional_embedding") == std::string::npos &&
                           name.find("bias") == std::string::npos;
    if (weight_round && is_matrix)
      for (size_t i = 0; i < nel; ++i) t.d[i] = round_bf16(t.d[i]);
    m->tensors[name] = std::move(t);
  }
  fclose(f);

  auto need = [&](const std::string& n) -> const float* {
    const float* p = m->T(n);
    if (!p) set_err("missing tensor %s", n.c_str());
    return p;
  };
  bool ok = true;
  auto get = [&](const std::string& n) {
    const float* p = need(n);
    if (!p) ok = false;
    return p;
  };
  auto load_attn = [&](const std::string& p, const std::string& ln) {
    AttnW a;
    a.ln_g = get(ln + ".weight");
    a.ln_b = get(ln + ".bias");
    a.qw = get(p + ".query.weight");
    a.qb = get(p + ".query.bias");
    a.kw = get(p + ".key.weight");
    a.vw = get(p + ".value.weight");
    a.vb = get(p + ".value.bias");
    a.ow = get(p + ".out.weight");
    a.ob = get(p + ".out.bias");
    return a;
  };
  auto load_mlp = [&](const std::string& p) {
    MlpW w;
    w.ln_g = get(p + ".mlp_ln.weight");
    w.ln_b = get(p + ".mlp_ln.bias");
    w.w0 = get(p + ".mlp.0.weight");
    w.b0 = get(p + ".mlp.0.bias");
    w.w2 = get(p + ".mlp.2.weight");
    w.b2 = get(p + ".mlp.2.bias");
    return w;
  };
  m->enc_pos = get("encoder.positional_embedding");
  m->conv1_w = get("encoder.conv1.weight");
  m->conv1_b = get("encoder.conv1.bias");
  m->conv2_w = get("encoder.conv2.weight");
  m->conv2_b = get("encoder.conv2.bias");
  m->ln_post_g = get("encoder.ln_post.weight");
  m->ln_post_b = get("encoder.ln_post.bias");
  for (int i = 0; i < hp.n_audio_layer && ok; ++i) {
    const std::string p = "encoder.blocks." + std::to_string(i);
    EncLayer L;
    L.attn = load_attn(p + ".attn", p + ".attn_ln");
    L.mlp = load_mlp(p);
    m->enc.push_back(L);
  }
  m->dec_pos = get("decoder.positional_embedding");
  m->tok_emb = get("decoder.token_embedding.weight");
  m->dec_ln_g = get("decoder.ln.weight");
  m->dec_ln_b = get("decoder.ln.bias");
  for (int i = 0; i < hp.n_text_layer && ok; ++i) {
    const std::string p = "decoder.blocks." + std::to_string(i);
    DecLayer L;
    L.self = load_attn(p + ".attn", p + ".attn_ln");
    L.cross = load_attn(p + ".cross_attn", p + ".cross_attn_ln");
    L.mlp = load_mlp(p);
    m->dec.push_back(L);
  }
  if (!ok) return false;
  // conv weights [d_out][c_in][3] -> [d_out][3][c_in] so that a conv is a K-major dot
  const int d = hp.n_audio_state;
  auto relayout = [&](const float* w, int cin, std::vector<float>& out) {
    out.resize((size_t)d * 3 * cin);
    for (int o = 0; o < d; ++o)
      for (int c = 0; c < cin; ++c)
        for (int k = 0; k < 3; ++k) out[((size_t)o * 3 + k) * cin + c] = w[((size_t)o * cin + c) * 3 + k];
  };
  relayout(m->conv1_w, hp.n_mels, m->conv1_wr);
  relayout(m->conv2_w, d, m->conv2_wr);

  // suppress_nst id list: token and " "+token, plus " -" and " '"
  for (const char* s : k_non_speech) {
    for (const std::string& cand : {std::string(s), std::string(" ") + s}) {
      auto it = m->token_to_id.find(cand);
      if (it != m->token_to_id.end()) m->nst_ids.push_back(it->second);
    }
  }
  for (const char* s : {" -", " '"}) {
    auto it = m->token_to_id.find(s);
    if (it != m->token_to_id.end()) m->nst_ids.push_back(it->second);
  }
  std::sort(m->nst_ids.begin(), m->nst_ids.end());
  m->nst_ids.erase(std::unique(m->nst_ids.begin(), m->nst_ids.end()), m->nst_ids.end());
  {
    auto it = m->token_to_id.find(" ");
    m->tok_space = it == m->token_to_id.end() ? -1 : it->second;
  }

  m->cross_k.assign(hp.n_text_layer, {});
  m->cross_v.assign(hp.n_text_layer, {});
  m->self_k.assign(MAX_SLOTS, std::vector<std::vector<float>>(hp.n_text_layer));
  m->self_v.assign(MAX_SLOTS, std::vector<std::vector<float>>(hp.n_text_layer));
  return true;
}

// ---------------------------------------------------------------------------------------------
// log_mel_spectrogram. fft()/dft() restate upstream's table-driven radix-2
// recursion that bottoms out in a naive DFT at odd sizes (400 -> 200 -> 100 -> 50 -> 25).
// ---------------------------------------------------------------------------------------------
constexpr int N_FFT = 400, HOP = 160, SR = 16000;
struct MelTables {
  float sin_t[N_FFT], cos_t[N_FFT], hann[N_FFT];
  MelTables() {
    for (int i = 0; i < N_FFT; ++i) {
      const double th = (2.0 * M_PI * i) / N_FFT;
      sin_t[i] = sinf(th);
      cos_t[i] = cosf(th);
      hann[i] = 0.5 * (1.0 - cosf((2.0 * M_PI * i) / N_FFT));  // periodic Hann
    }
  }
};
const MelTables& mel_tables() {
  static MelTables t;
  return t;
}
void dft(const float* in, int N, float* out) {
  const MelTables& t = mel_tables();
  const int step = N_FFT / N;
  for (int k = 0; k < N; ++k) {
    float re = 0, im = 0;
    for (int n = 0; n < N; ++n) {
      const int idx = (k * n * step) % N_FFT;
      re += in[n] * t.cos_t[idx];
      im -= in[n] * t.sin_t[idx];
    }
    out[2 * k] = re;
    out[2 * k + 1] = im;
  }
}
void fft(float* in, int N, float* out) {
  if (N == 1) {
    out[0] = in[0];
    out[1] = 0;
    return;
  }
  const int half = N / 2;
  if (N - half * 2 == 1) {
    dft(in, N, out);
    return;
  }
  const MelTables& t = mel_tables();
  float* even = in + N;
  for (int i = 0; i < half; ++i) even[i] = in[2 * i];
  float* even_fft = out + 2 * N;
  fft(even, half, even_fft);
  float* odd = even;
  for (int i = 0; i < half; ++i) odd[i] = in[2 * i + 1];
  float* odd_fft = even_fft + N;
  fft(odd, half, odd_fft);
  const int step = N_FFT / N;
  for (int k = 0; k < half; ++k) {
    const int idx = k * step;
    const float re = t.cos_t[idx], im = -t.sin_t[idx];
    const float ro = odd_fft[2 * k], io = odd_fft[2 * k + 1];
    out[2 * k] = even_fft[2 * k] + re * ro - im * io;
    out[2 * k + 1] = even_fft[2 * k + 1] + re * io + im * ro;
    out[2 * (k + half)] = even_fft[2 * k] - re * ro + im * io;
    out[2 * (k + half) + 1] = even_fft[2 * k + 1] - re * io - im * ro;
  }
}

int mel_impl(const ora_model* m, const float* pcm, int n_samples, float* out, int* n_len_out,
             int* n_len_org_out) {
  const int n_mel = m->hp.n_mels;
  const int64_t pad1 = (int64_t)SR * 30, pad2 = N_FFT / 2;
  const int64_t n_padded = n_samples + pad1 + 2 * pad2;
  const int n_len = (int)((n_padded - N_FFT) / HOP);
  const int n_len_org = 1 + (int)((n_samples + pad2 - N_FFT) / HOP);
  if (n_len_out) *n_len_out = n_len;
  if (n_len_org_out) *n_len_org_out = n_len_org;
  if (!out) return 0;
  std::vector<float> sp((size_t)n_padded, 0.f);
  if (n_samples > 0) memcpy(sp.data() + pad2, pcm, (size_t)n_samples * 4);
  // reflective pad at the beginning: reverse_copy(samples+1, samples+1+200)
  for (int i = 0; i < pad2; ++i) {
    const int src = (int)pad2 - i;  // samples[200], samples[199], ... samples[1]
    sp[i] = src < n_samples ? pcm[src] : 0.f;
  }
  const MelTables& tb = mel_tables();
  const int n_eff = n_samples + (int)pad2;  // the worker's n_samples
  const int n_active = std::min(n_eff / HOP + 1, n_len);
  const float* filt = m->filters.data();
#pragma omp parallel
  {
    std::vector<float> fin(N_FFT * 2, 0.f), fout(N_FFT * 8, 0.f);
#pragma omp for schedule(static)
    for (int i = 0; i < n_len; ++i) {
      if (i >= n_active) {
        const float v = (float)log10(1e-10);
        for (int j = 0; j < n_mel; ++j) out[(size_t)j * n_len + i] = v;
        continue;
      }
      const int off = i * HOP;
      const int lim = std::min(N_FFT, n_eff - off);
      for (int j = 0; j < lim; ++j) fin[j] = tb.hann[j] * sp[off + j];
      for (int j = std::max(lim, 0); j < N_FFT; ++j) fin[j] = 0.f;
      fft(fin.data(), N_FFT, fout.data());
      for (int j = 0; j < 201; ++j) fout[j] = fout[2 * j] * fout[2 * j] + fout[2 * j + 1] * fout[2 * j + 1];
      for (int j = 0; j < n_mel; ++j) {
        double sum = 0.0;
        const float* fr = filt + (size_t)j * 201;
        int k = 0;
        for (; k < 201 - 3; k += 4)
          sum += fout[k] * fr[k] + fout[k + 1] * fr[k + 1] + fout[k + 2] * fr[k + 2] +
                 fout[k + 3] * fr[k + 3];
        for (; k < 201; ++k) sum += fout[k] * fr[k];
        sum = log10(std::max(sum, 1e-10));
        out[(size_t)j * n_len + i] = (float)sum;
      }
    }
  }
  // clamp + normalise over the WHOLE buffer (not per 30 s window)
  double mmax = -1e20;
  const size_t tot = (size_t)n_mel * n_len;
  for (size_t i = 0; i < tot; ++i)
    if (out[i] > mmax) mmax = out[i];
  mmax -= 8.0;
  for (size_t i = 0; i < tot; ++i) {
    if (out[i] < mmax) out[i] = (float)mmax;
    out[i] = (float)((out[i] + 4.0) / 4.0);
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// attention over [Tq][d] queries and [Tk][d] keys/values, heads of 64
// (ggml_flash_attn_ext path: scale 1/sqrt(64), f32 softmax).
// ---------------------------------------------------------------------------------------------
void attention(const float* q, int Tq, const float* k, const float* v, int Tk, int d, int n_head,
               bool causal, int q_pos0, int p_round, float* out) {
  const int dh = d / n_head;
  const float scale = 1.0f / sqrtf((float)dh);
#pragma omp parallel
  {
    std::vector<float> s(Tk);
#pragma omp for collapse(2) schedule(static)
    for (int h = 0; h < n_head; ++h)
      for (int i = 0; i < Tq; ++i) {
        const float* qi = q + (size_t)i * d + h * dh;
        const int lim = causal ? std::min(Tk, q_pos0 + i + 1) : Tk;
        float mx = -INFINITY;
        for (int j = 0; j < lim; ++j) {
          const float* kj = k + (size_t)j * d + h * dh;
          float acc = 0;
          for (int c = 0; c < dh; ++c) acc += qi[c] * kj[c];
          s[j] = acc * scale;
          mx = std::max(mx, s[j]);
        }
        float sum = 0;
        for (int j = 0; j < lim; ++j) {
          s[j] = expf(s[j] - mx);
          sum += s[j];
        }
        float o[256];
        for (int c = 0; c < dh; ++c) o[c] = 0;
        for (int j = 0; j < lim; ++j) {
          const float p = round_mode(s[j], p_round);
          const float* vj = v + (size_t)j * d + h * dh;
          for (int c = 0; c < dh; ++c) o[c] += p * vj[c];
        }
        const float inv = 1.0f / sum;
        float* oi = out + (size_t)i * d + h * dh;
        for (int c = 0; c < dh; ++c) oi[c] = o[c] * inv;
      }
  }
}

// rounding helper: returns pointer to rounded copy (or the input in f32 mode)
const float* rounded(const ora_model* m, const float* x, size_t n, std::vector<float>& tmp) {
  if (m->act_round == ORA_ACT_F32) return x;
  tmp.resize(n);
  round_buf(x, tmp.data(), n, m->act_round);
  return tmp.data();
}

inline float gelu(const ora_model* m, float x) { return m->gelu_erf_flag ? gelu_erf(x) : gelu_tanh(x); }

// ---------------------------------------------------------------------------------------------
// whisper_encode_internal: conv stem -> encoder blocks -> ln_post -> cross K/V
// ---------------------------------------------------------------------------------------------
int encode_impl(ora_model* m, const float* mel /*[n_mel][3000]*/, float* enc_out) {
  const ora_hparams& hp = m->hp;
  const int d = hp.n_audio_state, n_mel = hp.n_mels, T2 = 2 * hp.n_audio_ctx, T = hp.n_audio_ctx;
  const int pr = m->act_round == ORA_ACT_BF16 ? ORA_ACT_BF16 : ORA_ACT_F32;
  std::vector<float> tmp;
  // conv1 as im2col: rows t, cols [k][c]
  std::vector<float> col1((size_t)T2 * 3 * n_mel);
#pragma omp parallel for schedule(static)
  for (int t = 0; t < T2; ++t)
    for (int k = 0; k < 3; ++k) {
      const int ts = t + k - 1;
      for (int c = 0; c < n_mel; ++c)
        col1[((size_t)t * 3 + k) * n_mel + c] =
            (ts >= 0 && ts < T2) ? round_mode(mel[(size_t)c * T2 + ts], m->act_round) : 0.f;
    }
  std::vector<float> h1((size_t)T2 * d);
  matmul_nt(col1.data(), T2, 3 * n_mel, 3 * n_mel, m->conv1_wr.data(), d, m->conv1_b, h1.data(), d);
  for (size_t i = 0; i < h1.size(); ++i) h1[i] = gelu(m, h1[i]);
  std::vector<float>().swap(col1);
  std::vector<float> col2((size_t)T * 3 * d);
#pragma omp parallel for schedule(static)
  for (int t = 0; t < T; ++t)
    for (int k = 0; k < 3; ++k) {
      const int ts = 2 * t + k - 1;
      float* dst = &col2[((size_t)t * 3 + k) * d];
      if (ts >= 0 && ts < T2)
        for (int c = 0; c < d; ++c) dst[c] = round_mode(h1[(size_t)ts * d + c], m->act_round);
      else
        for (int c = 0; c < d; ++c) dst[c] = 0.f;
    }
  std::vector<float> x((size_t)T * d);
  matmul_nt(col2.data(), T, 3 * d, 3 * d, m->conv2_wr.data(), d, m->conv2_b, x.data(), d);
  for (size_t i = 0; i < x.size(); ++i) x[i] = gelu(m, x[i]) + m->enc_pos[i];
  std::vector<float>().swap(col2);
  std::vector<float>().swap(h1);
  m->taps.assign(1 + hp.n_audio_layer, {});
  m->taps[0] = x;

  std::vector<float> hbuf((size_t)T * d), q((size_t)T * d), k((size_t)T * d), v((size_t)T * d),
      att((size_t)T * d), proj((size_t)T * d), ff((size_t)T * 4 * d);
  for (int l = 0; l < hp.n_audio_layer; ++l) {
    const EncLayer& L = m->enc[l];
    layer_norm(x.data(), T, d, L.attn.ln_g, L.attn.ln_b, hbuf.data());
    const float* hr = rounded(m, hbuf.data(), hbuf.size(), tmp);
    matmul_nt(hr, T, d, d, L.attn.qw, d, L.attn.qb, q.data(), d);
    matmul_nt(hr, T, d, d, L.attn.kw, d, nullptr, k.data(), d);
    matmul_nt(hr, T, d, d, L.attn.vw, d, L.attn.vb, v.data(), d);
    round_buf(q.data(), q.data(), q.size(), m->act_round);
    round_buf(k.data(), k.data(), k.size(), m->act_round);
    round_buf(v.data(), v.data(), v.size(), m->act_round);
    attention(q.data(), T, k.data(), v.data(), T, d, hp.n_audio_head, false, 0, pr, att.data());
    const float* ar = rounded(m, att.data(), att.size(), tmp);
    matmul_nt(ar, T, d, d, L.attn.ow, d, L.attn.ob, proj.data(), d);
    for (size_t i = 0; i < x.size(); ++i) x[i] += proj[i];
    layer_norm(x.data(), T, d, L.mlp.ln_g, L.mlp.ln_b, hbuf.data());
    hr = rounded(m, hbuf.data(), hbuf.size(), tmp);
    matmul_nt(hr, T, d, d, L.mlp.w0, 4 * d, L.mlp.b0, ff.data(), 4 * d);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)ff.size(); ++i) ff[i] = round_mode(gelu(m, ff[i]), m->act_round);
    matmul_nt(ff.data(), T, 4 * d, 4 * d, L.mlp.w2, d, L.mlp.b2, proj.data(), d);
    for (size_t i = 0; i < x.size(); ++i) x[i] += proj[i];
    m->taps[1 + l] = x;
  }
  std::vector<float> enc((size_t)T * d);
  layer_norm(x.data(), T, d, m->ln_post_g, m->ln_post_b, enc.data());
  if (enc_out) memcpy(enc_out, enc.data(), enc.size() * 4);
  // cross K/V per decoder layer, stored rounded (whisper.cpp: f16 kv_cross)
  const float* er = rounded(m, enc.data(), enc.size(), tmp);
  for (int l = 0; l < hp.n_text_layer; ++l) {
    const AttnW& c = m->dec[l].cross;
    m->cross_k[l].resize((size_t)T * d);
    m->cross_v[l].resize((size_t)T * d);
    matmul_nt(er, T, d, d, c.kw, d, nullptr, m->cross_k[l].data(), d);
    matmul_nt(er, T, d, d, c.vw, d, c.vb, m->cross_v[l].data(), d);
    round_buf(m->cross_k[l].data(), m->cross_k[l].data(), (size_t)T * d, m->act_round);
    round_buf(m->cross_v[l].data(), m->cross_v[l].data(), (size_t)T * d, m->act_round);
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// whisper_decode_internal
// ---------------------------------------------------------------------------------------------
int decode_impl(ora_model* m, int slot, const int32_t* tokens, int n, int n_past, float* logits_out,
                bool last_only) {
  const ora_hparams& hp = m->hp;
  const int d = hp.n_text_state, T = hp.n_audio_ctx, nctx = hp.n_text_ctx;
  if (slot < 0 || slot >= MAX_SLOTS) {
    set_err("bad decoder slot %d", slot);
    return -1;
  }
  if (n_past + n > nctx) {
    set_err("decoder context overflow: %d + %d > %d", n_past, n, nctx);
    return -1;
  }
  if (m->cross_k.empty() || m->cross_k[0].empty()) {
    set_err("ora_decode before ora_encode");
    return -1;
  }
  const int pr = m->act_round == ORA_ACT_BF16 ? ORA_ACT_BF16 : ORA_ACT_F32;
  std::vector<float> tmp;
  std::vector<float> x((size_t)n * d), h((size_t)n * d), q((size_t)n * d), kk((size_t)n * d),
      vv((size_t)n * d), att((size_t)n * d), proj((size_t)n * d), ff((size_t)n * 4 * d);
  for (int i = 0; i < n; ++i) {
    const int id = tokens[i];
    if (id < 0 || id >= hp.n_vocab) {
      set_err("token id %d out of range", id);
      return -1;
    }
    for (int c = 0; c < d; ++c)
      x[(size_t)i * d + c] = m->tok_emb[(size_t)id * d + c] + m->dec_pos[(size_t)(n_past + i) * d + c];
  }
  for (int l = 0; l < hp.n_text_layer; ++l) {
    const DecLayer& L = m->dec[l];
    auto& sk = m->self_k[slot][l];
    auto& sv = m->self_v[slot][l];
    if (sk.empty()) {
      sk.assign((size_t)nctx * d, 0.f);
      sv.assign((size_t)nctx * d, 0.f);
    }
    // self attention
    layer_norm(x.data(), n, d, L.self.ln_g, L.self.ln_b, h.data());
    const float* hr = rounded(m, h.data(), h.size(), tmp);
    matmul_nt(hr, n, d, d, L.self.qw, d, L.self.qb, q.data(), d);
    matmul_nt(hr, n, d, d, L.self.kw, d, nullptr, kk.data(), d);
    matmul_nt(hr, n, d, d, L.self.vw, d, L.self.vb, vv.data(), d);
    round_buf(q.data(), q.data(), q.size(), m->act_round);
    round_buf(kk.data(), sk.data() + (size_t)n_past * d, (size_t)n * d, m->act_round);
    round_buf(vv.data(), sv.data() + (size_t)n_past * d, (size_t)n * d, m->act_round);
    attention(q.data(), n, sk.data(), sv.data(), n_past + n, d, hp.n_text_head, true, n_past, pr,
              att.data());
    const float* ar = rounded(m, att.data(), att.size(), tmp);
    matmul_nt(ar, n, d, d, L.self.ow, d, L.self.ob, proj.data(), d);
    for (size_t i = 0; i < x.size(); ++i) x[i] += proj[i];
    // cross attention
    layer_norm(x.data(), n, d, L.cross.ln_g, L.cross.ln_b, h.data());
    hr = rounded(m, h.data(), h.size(), tmp);
    matmul_nt(hr, n, d, d, L.cross.qw, d, L.cross.qb, q.data(), d);
    round_buf(q.data(), q.data(), q.size(), m->act_round);
    attention(q.data(), n, m->cross_k[l].data(), m->cross_v[l].data(), T, d, hp.n_text_head, false, 0,
              pr, att.data());
    ar = rounded(m, att.data(), att.size(), tmp);
    matmul_nt(ar, n, d, d, L.cross.ow, d, L.cross.ob, proj.data(), d);
    for (size_t i = 0; i < x.size(); ++i) x[i] += proj[i];
    // mlp
    layer_norm(x.data(), n, d, L.mlp.ln_g, L.mlp.ln_b, h.data());
    hr = rounded(m, h.data(), h.size(), tmp);
    matmul_nt(hr, n, d, d, L.mlp.w0, 4 * d, L.mlp.b0, ff.data(), 4 * d);
    for (size_t i = 0; i < ff.size(); ++i) ff[i] = round_mode(gelu(m, ff[i]), m->act_round);
    matmul_nt(ff.data(), n, 4 * d, 4 * d, L.mlp.w2, d, L.mlp.b2, proj.data(), d);
    for (size_t i = 0; i < x.size(); ++i) x[i] += proj[i];
  }
  if (logits_out) {
    layer_norm(x.data(), n, d, m->dec_ln_g, m->dec_ln_b, h.data());
    const float* hr = rounded(m, h.data(), h.size(), tmp);
    if (last_only)
      matmul_nt(hr + (size_t)(n - 1) * d, 1, d, d, m->tok_emb, hp.n_vocab, nullptr, logits_out,
                hp.n_vocab);
    else
      matmul_nt(hr, n, d, d, m->tok_emb, hp.n_vocab, nullptr, logits_out, hp.n_vocab);
  }
  return 0;
}

void kv_copy(ora_model* m, int dst, int src, int n_tok) {
  if (dst == src) return;
  const int d = m->hp.n_text_state;
  for (int l = 0; l < m->hp.n_text_layer; ++l) {
    if (m->self_k[src][l].empty()) continue;
    if (m->self_k[dst][l].empty()) {
      m->self_k[dst][l].assign((size_t)m->hp.n_text_ctx * d, 0.f);
      m->self_v[dst][l].assign((size_t)m->hp.n_text_ctx * d, 0.f);
    }
    memcpy(m->self_k[dst][l].data(), m->self_k[src][l].data(), (size_t)n_tok * d * 4);
    memcpy(m->self_v[dst][l].data(), m->self_v[src][l].data(), (size_t)n_tok * d * 4);
  }
}

// ---------------------------------------------------------------------------------------------
// whisper_process_logits (SURVEY.md A.6 "Logit rules")
// ---------------------------------------------------------------------------------------------
void process_logits_impl(const ora_model* m, const ora_full_params* p, const int32_t* cur, int n_cur,
                         int has_ts, int seek_delta, float temperature, float* logits,
                         float* logprobs, float* probs) {
  const ora_hparams& hp = m->hp;
  const int n = hp.n_vocab;
  const bool is_initial = n_cur == 0;
  if (temperature > 0.0f)
    for (int i = 0; i < n; ++i) logits[i] /= temperature;
  if (p->suppress_blank && is_initial) {
    logits[hp.token_eot] = -INFINITY;
    if (m->tok_space >= 0) logits[m->tok_space] = -INFINITY;
  }
  logits[hp.token_not] = -INFINITY;
  if (p->no_timestamps)
    for (int i = hp.token_beg; i < n; ++i) logits[i] = -INFINITY;
  logits[hp.token_sot] = -INFINITY;
  logits[hp.token_nosp] = -INFINITY;
  if (!p->tdrz_enable) logits[hp.token_solm] = -INFINITY;
  logits[hp.token_translate] = -INFINITY;
  logits[hp.token_transcribe] = -INFINITY;
  logits[hp.token_prev] = -INFINITY;
  for (int i = 0; i < m->n_langs; ++i) logits[hp.token_sot + 1 + i] = -INFINITY;
  if (p->suppress_nst)
    for (int id : m->nst_ids) logits[id] = -INFINITY;
  // timestamps come in pairs, except directly before EOT
  {
    const bool last_ts = n_cur > 0 && cur[n_cur - 1] >= hp.token_beg;
    const bool penult_ts = n_cur < 2 || cur[n_cur - 2] >= hp.token_beg;
    if (last_ts) {
      if (penult_ts)
        for (int i = hp.token_beg; i < n; ++i) logits[i] = -INFINITY;
      else
        for (int i = 0; i < hp.token_eot; ++i) logits[i] = -INFINITY;
    }
  }
  if (is_initial && p->max_initial_ts > 0.0f) {
    const float precision = 30.0f / hp.n_audio_ctx;
    const int tid0 = (int)roundf(p->max_initial_ts / precision);
    for (int i = hp.token_beg + tid0 + 1; i < n; ++i) logits[i] = -INFINITY;
  }
  if (has_ts) {
    const int tid0 = seek_delta / 2;
    for (int i = hp.token_beg; i < hp.token_beg + tid0 && i < n; ++i) logits[i] = -INFINITY;
  }
  // log-softmax
  {
    float mx = -INFINITY;
    for (int i = 0; i < n; ++i) mx = std::max(mx, logits[i]);
    float lse = 0.0f;
    for (int i = 0; i < n; ++i)
      if (logits[i] > -INFINITY) lse += expf(logits[i] - mx);
    lse = logf(lse) + mx;
    for (int i = 0; i < n; ++i) logprobs[i] = logits[i] > -INFINITY ? logits[i] - lse : -INFINITY;
  }
  // if the timestamp mass beats every text token, force a timestamp
  {
    float ts_logprob = -INFINITY;
    {
      float mx = -INFINITY;
      for (int i = hp.token_beg; i < n; ++i) mx = std::max(mx, logprobs[i]);
      float s = 0.0f;
      for (int i = hp.token_beg; i < n; ++i)
        if (logprobs[i] > -INFINITY) s += expf(logprobs[i] - mx);
      if (s > 0.0f) ts_logprob = logf(s) + mx;
    }
    float max_text = -INFINITY;
    for (int i = 0; i < hp.token_beg; ++i) max_text = std::max(max_text, logprobs[i]);
    if (ts_logprob > max_text)
      for (int i = 0; i < hp.token_beg; ++i) {
        logits[i] = -INFINITY;
        logprobs[i] = -INFINITY;
      }
  }
  for (int i = 0; i < n; ++i) probs[i] = logits[i] == -INFINITY ? 0.0f : expf(logprobs[i]);
}

// whisper_sample_token (best / sampled) incl. timestamp statistics
ora_token_data sample_token(const ora_model* m, const float* probs, const float* logprobs, bool best,
                            std::mt19937& rng) {
  const ora_hparams& hp = m->hp;
  const int n = hp.n_vocab;
  ora_token_data r = {0, 0, 0.f, 0.f, 0.f, 0.f, -1, -1, -1, 0.f};
  {
    double sum_ts = 0.0, max_ts = 0.0;
    for (int i = hp.token_beg; i < n; ++i) {
      sum_ts += probs[i];
      if (max_ts < probs[i]) {
        max_ts = probs[i];
        r.tid = i;
      }
    }
    r.pt = (float)(max_ts / (sum_ts + 1e-10));
    r.ptsum = (float)sum_ts;
  }
  if (best) {
    for (int i = 0; i < n; ++i)
      if (r.p < probs[i]) {
        r.id = i;
        r.p = probs[i];
        r.plog = logprobs[i];
      }
  } else {
    std::discrete_distribution<> dist(probs, probs + n);
    r.id = dist(rng);
    r.p = probs[r.id];
    r.plog = logprobs[r.id];
  }
  if (r.id >= hp.token_beg) {
    r.tid = r.id;
    r.pt = r.p;
  }
  return r;
}

struct Sequence {
  std::vector<ora_token_data> tokens;
  int result_len = 0;
  double sum_logprobs_all = 0, sum_logprobs = 0, avg_logprobs = 0, entropy = 0, score = 0;
};
struct Decoder {
  Sequence seq;
  int seek_delta = 0;
  bool failed = false, completed = false, has_ts = false;
  std::vector<float> logits, logprobs, probs;
  std::mt19937 rng;
};

// whisper_sequence_score
void sequence_score(const ora_full_params* p, Sequence& s) {
  if (s.result_len == 0) return;
  double r = 0.0;
  for (int i = 0; i < s.result_len; ++i) r += s.tokens[i].plog;
  s.sum_logprobs = r;
  s.avg_logprobs = r / s.result_len;
  double penalty = s.result_len;
  if (p->length_penalty > 0.0f) penalty = pow((5.0 + penalty) / 6.0, p->length_penalty);
  s.score = r / penalty;
  std::map<int, int> cnt;
  int c = 0;
  for (int i = std::max(0, s.result_len - 32); i < s.result_len; ++i) {
    cnt[s.tokens[i].id]++;
    c++;
  }
  double e = 0.0;
  for (auto& kv : cnt) {
    const double q = kv.second / (double)c;
    e -= q * log(q);
  }
  s.entropy = e;
}

// ---- token-level timestamps (whisper_exp_compute_token_level_timestamps) ----
float voice_length(const std::string& text) {
  float r = 0.f;
  for (char c : text) {
    if (c == ' ')
      r += 0.01f;
    else if (c == ',')
      r += 2.0f;
    else if (c == '.' || c == '!' || c == '?')
      r += 3.0f;
    else if (c >= '0' && c <= '9')
      r += 3.0f;
    else
      r += 1.0f;
  }
  return r;
}
inline int ts_to_sample(int64_t t, int n) {
  return std::max(0, std::min(n - 1, (int)((t * SR) / 100)));
}
inline int64_t sample_to_ts(int i) { return (100ll * i) / SR; }

void signal_energy(const float* s, int n, int hw, std::vector<float>& out) {
  out.resize(n);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    float sum = 0;
    for (int j = -hw; j <= hw; ++j)
      if (i + j >= 0 && i + j < n) sum += fabsf(s[i + j]);
    out[i] = sum / (2 * hw + 1);
  }
}

void token_level_timestamps(ora_model* m, ora_segment& seg, float thold_pt, float thold_ptsum) {
  auto& tk = seg.tokens;
  const int n_samples = (int)m->energy.size();
  if (n_samples == 0) return;
  const int64_t t0 = seg.t0, t1 = seg.t1;
  const int n = (int)tk.size();
  if (n == 0) return;
  if (n == 1) {
    tk[0].t0 = t0;
    tk[0].t1 = t1;
    return;
  }
  const int beg = m->hp.token_beg;
  for (int j = 0; j < n; ++j) {
    if (j == 0) {
      if (tk[j].id == beg) {
        tk[j].t0 = t0;
        tk[j].t1 = t0;
        tk[j + 1].t0 = t0;
        m->t_beg = t0;
        m->t_last = t0;
        m->tid_last = beg;
      } else {
        tk[j].t0 = m->t_last;
      }
    }
    const int64_t tt = m->t_beg + 2 * (tk[j].tid - beg);
    tk[j].vlen = voice_length(m->id_to_token[tk[j].id]);
    if (tk[j].pt > thold_pt && tk[j].ptsum > thold_ptsum && tk[j].tid > m->tid_last && tt <= t1) {
      if (j > 0) tk[j - 1].t1 = tt;
      tk[j].t0 = tt;
      m->tid_last = tk[j].tid;
    }
  }
  tk[n - 2].t1 = t1;
  tk[n - 1].t0 = t1;
  tk[n - 1].t1 = t1;
  m->t_last = t1;
  {
    int p0 = 0, p1 = 0;
    while (true) {
      while (p1 < n && tk[p1].t1 < 0) p1++;
      if (p1 >= n) p1--;
      if (p1 > p0) {
        double psum = 0.0;
        for (int j = p0; j <= p1; ++j) psum += tk[j].vlen;
        const double dt = (double)(tk[p1].t1 - tk[p0].t0);
        for (int j = p0 + 1; j <= p1; ++j) {
          const double ct = tk[j - 1].t0 + dt * tk[j - 1].vlen / psum;
          tk[j - 1].t1 = (int64_t)ct;
          tk[j].t0 = (int64_t)ct;
        }
      }
      p1++;
      p0 = p1;
      if (p1 >= n) break;
    }
  }
  for (int j = 0; j < n - 1; ++j) {
    if (tk[j].t1 < 0) tk[j + 1].t0 = tk[j].t1;
    if (j > 0 && tk[j - 1].t1 > tk[j].t0) {
      tk[j].t0 = tk[j - 1].t1;
      tk[j].t1 = std::max(tk[j].t0, tk[j].t1);
    }
  }
  // expand or contract tokens based on voice activity
  const std::vector<float>& en = m->energy;
  const int hw = SR / 8;
  for (int j = 0; j < n; ++j) {
    if (tk[j].id >= m->hp.token_eot) continue;
    int s0 = ts_to_sample(tk[j].t0, n_samples);
    int s1 = ts_to_sample(tk[j].t1, n_samples);
    const int ss0 = std::max(s0 - hw, 0), ss1 = std::min(s1 + hw, n_samples);
    const int ns = ss1 - ss0;
    float sum = 0.f;
    for (int k = ss0; k < ss1; ++k) sum += en[k];
    const float thold = 0.5f * sum / ns;
    {
      int k = s0;
      if (en[k] > thold && j > 0) {
        while (k > 0 && en[k] > thold) k--;
        tk[j].t0 = sample_to_ts(k);
        if (tk[j].t0 < tk[j - 1].t1)
          tk[j].t0 = tk[j - 1].t1;
        else
          s0 = k;
      } else {
        while (en[k] < thold && k < s1) k++;
        s0 = k;
        tk[j].t0 = sample_to_ts(k);
      }
    }
    {
      int k = s1;
      if (en[k] > thold) {
        while (k < n_samples - 1 && en[k] > thold) k++;
        tk[j].t1 = sample_to_ts(k);
        if (j < n - 1 && tk[j].t1 > tk[j + 1].t0)
          tk[j].t1 = tk[j + 1].t0;
        else
          s1 = k;
      } else {
        while (en[k] < thold && k > s0) k--;
        s1 = k;
        tk[j].t1 = sample_to_ts(k);
      }
    }
  }
}

}  // namespace

// =============================================================================================
// C API
// =============================================================================================
extern "C" {

const char* ora_last_error(void) { return g_err; }

ora_model* ora_load(const char* path, int weight_round) {
  ora_model* m = new ora_model();
  if (!load_model(m, path, weight_round)) {
    delete m;
    return nullptr;
  }
  return m;
}
void ora_free(ora_model* m) { delete m; }
void ora_set_act_round(ora_model* m, int mode) { m->act_round = mode; }
void ora_set_gelu_erf(ora_model* m, int e) { m->gelu_erf_flag = e; }
void ora_set_threads(ora_model* m, int n) { m->n_threads = n; }
void ora_get_hparams(const ora_model* m, ora_hparams* out) { *out = m->hp; }
const char* ora_token_to_str(const ora_model* m, int id) {
  if (id < 0 || id >= (int)m->id_to_token.size()) return "";
  return m->id_to_token[id].c_str();
}
int ora_lang_id(const char* lang) {
  if (!lang) return -1;
  for (int i = 0; i < k_n_langs; ++i)
    if (strcmp(lang, k_langs[i]) == 0) return i;
  return -1;
}

/* whisper_tokenize restatement: upstream splits the text into words with a GPT-2 style regex and
 * then matches greedily the longest vocabulary entry from each position; here the longest-match is
 * applied to the whole string (identical for prompts whose words are vocabulary entries). */
int ora_tokenize(const ora_model* m, const char* text, int32_t* out, int max_tokens) {
  const std::string s(text ? text : "");
  int n = 0;
  size_t i = 0;
  while (i < s.size()) {
    int found = -1;
    size_t flen = 0;
    for (size_t len = std::min<size_t>(s.size() - i, 32); len >= 1; --len) {
      auto it = m->token_to_id.find(s.substr(i, len));
      if (it != m->token_to_id.end() && it->second < m->hp.token_eot) {
        found = it->second;
        flen = len;
        break;
      }
    }
    if (found < 0) {
      ++i;
      continue;
    }
    if (n < max_tokens && out) out[n] = found;
    ++n;
    i += flen;
  }
  return n;
}

int ora_mel(const ora_model* m, const float* pcm, int n_samples, float* out, int* n_len,
            int* n_len_org) {
  ThreadScope ts(m->n_threads);
  return mel_impl(m, pcm, n_samples, out, n_len, n_len_org);
}
int ora_encode(ora_model* m, const float* mel_window, float* enc_out) {
  ThreadScope ts(m->n_threads);
  return encode_impl(m, mel_window, enc_out);
}
int ora_encode_tap(const ora_model* m, int which, float* out) {
  if (which < 0 || which >= (int)m->taps.size()) return -1;
  if (out) memcpy(out, m->taps[which].data(), m->taps[which].size() * 4);
  return (int)m->taps[which].size();
}
int ora_decode(ora_model* m, int slot, const int32_t* tokens, int n_tokens, int n_past,
               float* logits_out) {
  ThreadScope ts(m->n_threads);
  return decode_impl(m, slot, tokens, n_tokens, n_past, logits_out, false);
}
void ora_process_logits(const ora_model* m, const ora_full_params* p, const int32_t* tokens_cur,
                        int n_cur, int has_ts, int seek_delta, float temperature,
                        const float* logits_in, float* logits_out, float* logprobs, float* probs) {
  memcpy(logits_out, logits_in, (size_t)m->hp.n_vocab * 4);
  process_logits_impl(m, p, tokens_cur, n_cur, has_ts, seek_delta, temperature, logits_out, logprobs,
                      probs);
}

ora_full_params ora_full_default_params(int strategy) {  // whisper_full_default_params
  ora_full_params p;
  memset(&p, 0, sizeof(p));
  p.strategy = strategy;
  p.beam_size = strategy == 1 ? 5 : -1;
  p.best_of = strategy == 0 ? 5 : -1;
  p.temperature = 0.0f;
  p.temperature_inc = 0.2f;
  p.entropy_thold = 2.4f;
  p.logprob_thold = -1.0f;
  p.no_speech_thold = 0.6f;
  p.suppress_blank = 1;
  p.no_context = 1;
  p.max_initial_ts = 1.0f;
  p.length_penalty = -1.0f;
  p.language = "en";
  return p;
}

/* whisper_full_with_state; reference call site stt_engine.cpp:245-246. */
int ora_full(ora_model* m, const ora_full_params* pp, const float* pcm, int n_samples,
             ora_result** out) {
  ThreadScope tsc(m->n_threads);
  const ora_full_params& p = *pp;
  const ora_hparams& hp = m->hp;
  ora_result* res = new ora_result();
  *out = res;
  const int n_mel = hp.n_mels;

  double t0 = now_ms();
  int n_len = 0, n_len_org = 0;
  mel_impl(m, pcm, n_samples, nullptr, &n_len, &n_len_org);
  std::vector<float> mel((size_t)n_mel * n_len);
  mel_impl(m, pcm, n_samples, mel.data(), &n_len, &n_len_org);
  if (p.token_timestamps) {
    m->t_beg = 0;
    m->t_last = 0;
    m->tid_last = 0;
    if (n_samples > 0) signal_energy(pcm, n_samples, 32, m->energy);
  } else {
    m->energy.clear();
  }
  res->ms_mel += now_ms() - t0;

  const int T2 = 2 * hp.n_audio_ctx;
  auto window = [&](int seek, std::vector<float>& w) {
    w.assign((size_t)n_mel * T2, 0.f);
    const int i0 = std::min(seek, n_len), i1 = std::min(seek + T2, n_len);
    for (int j = 0; j < n_mel; ++j)
      for (int i = i0; i < i1; ++i) w[(size_t)j * T2 + (i - i0)] = mel[(size_t)j * n_len + i];
  };

  // language
  int lang_id = -1;
  std::vector<float> win;
  std::vector<float> logits_buf((size_t)hp.n_vocab);
  if (hp.is_multilingual) {
    if (!p.language || !p.language[0] || strcmp(p.language, "auto") == 0) {
      // whisper_lang_auto_detect_with_state: encode window 0, decode [sot], argmax over languages
      window(0, win);
      double te = now_ms();
      encode_impl(m, win.data(), nullptr);
      res->ms_encode += now_ms() - te;
      const int32_t sot = hp.token_sot;
      if (decode_impl(m, 0, &sot, 1, 0, logits_buf.data(), true)) return -1;
      float best = -INFINITY;
      for (int i = 0; i < m->n_langs; ++i) {
        const float v = logits_buf[hp.token_sot + 1 + i];
        if (v > best) {
          best = v;
          lang_id = i;
        }
      }
    } else {
      lang_id = ora_lang_id(p.language);
      if (lang_id < 0 || lang_id >= m->n_langs) {
        set_err("unknown language '%s'", p.language);
        return -1;
      }
    }
  }
  res->lang_id = lang_id;

  int seek = 0;
  const int seek_end = n_len_org;
  if (seek_end < seek + 10) return 0;  // "input is too short"

  std::vector<int32_t> prompt_past;
  if (p.prompt_tokens && p.prompt_n_tokens > 0)
    prompt_past.assign(p.prompt_tokens, p.prompt_tokens + p.prompt_n_tokens);
  else if (p.initial_prompt && p.initial_prompt[0]) {
    std::vector<int32_t> tk(1024);
    int n = ora_tokenize(m, p.initial_prompt, tk.data(), 1024);
    tk.resize(std::min(n, 1024));
    prompt_past = tk;
  }

  int n_decoders = 1;
  if (p.strategy == 0)
    n_decoders = p.best_of;
  else
    n_decoders = std::max(p.best_of, p.beam_size);
  n_decoders = std::max(1, std::min(n_decoders, 8));
  std::vector<Decoder> dec(n_decoders);
  for (int j = 0; j < n_decoders; ++j) {
    dec[j].rng = std::mt19937(j);
    dec[j].logits.resize(hp.n_vocab);
    dec[j].logprobs.resize(hp.n_vocab);
    dec[j].probs.resize(hp.n_vocab);
  }
  std::vector<float> temps;
  if (p.temperature_inc > 0.0f)
    for (float t = p.temperature; t < 1.0f + 1e-6f; t += p.temperature_inc) temps.push_back(t);
  else
    temps.push_back(p.temperature);

  std::vector<int32_t> prompt_init;
  prompt_init.push_back(hp.token_sot);
  if (hp.is_multilingual) {
    prompt_init.push_back(hp.token_sot + 1 + lang_id);
    prompt_init.push_back(p.translate ? hp.token_translate : hp.token_transcribe);
  }
  if (p.no_timestamps) prompt_init.push_back(hp.token_not);

  const int n_max = hp.n_text_ctx / 2 - 4;
  const int delta_min = 10;
  struct Cand {
    int decoder_idx, seek_delta;
    bool has_ts;
    Sequence seq;
  };

  while (true) {
    if (seek + 100 >= seek_end) break;
    res->n_windows++;
    window(seek, win);
    double te = now_ms();
    encode_impl(m, win.data(), nullptr);
    res->ms_encode += now_ms() - te;
    double td = now_ms();

    std::vector<int32_t> prompt;
    int best_id = 0;
    float no_speech_prob = 0.f;
    for (size_t it = 0; it < temps.size(); ++it) {
      const float t_cur = temps[it];
      int n_cur = 1;
      if (p.strategy == 0) {
        if (t_cur > 0.0f) n_cur = p.best_of;
      } else {
        n_cur = t_cur > 0.0f ? p.best_of : p.beam_size;
      }
      n_cur = std::max(1, std::min(n_cur, n_decoders));
      // prompt: [prev, tail of prompt_past] only on the t < 0.5 passes, then sot/lang/task
      prompt.clear();
      if (!prompt_past.empty() && t_cur < 0.5f) {
        const int n_take = std::min(std::min(16384, hp.n_text_ctx / 2), (int)prompt_past.size());
        prompt.push_back(hp.token_prev);
        prompt.insert(prompt.end(), prompt_past.end() - n_take, prompt_past.end());
      }
      prompt.insert(prompt.end(), prompt_init.begin(), prompt_init.end());
      for (int j = 0; j < n_cur; ++j) {
        Decoder& d = dec[j];
        d.seq = Sequence();
        d.seek_delta = 100 * 30;
        d.failed = d.completed = d.has_ts = false;
      }
      if (decode_impl(m, 0, prompt.data(), (int)prompt.size(), 0, logits_buf.data(), true)) return -1;
      res->n_decode_steps++;
      {  // no-speech probability from the unfiltered logits of the last prompt token
        float mx = -INFINITY;
        for (int i = 0; i < hp.n_vocab; ++i) mx = std::max(mx, logits_buf[i]);
        float s = 0;
        for (int i = 0; i < hp.n_vocab; ++i) s += expf(logits_buf[i] - mx);
        const float lse = logf(s) + mx;
        no_speech_prob = expf(logits_buf[hp.token_nosp] - lse);
      }
      memcpy(dec[0].logits.data(), logits_buf.data(), (size_t)hp.n_vocab * 4);
      process_logits_impl(m, &p, nullptr, 0, 0, dec[0].seek_delta, t_cur, dec[0].logits.data(),
                          dec[0].logprobs.data(), dec[0].probs.data());
      for (int j = 1; j < n_cur; ++j) {
        kv_copy(m, j, 0, (int)prompt.size());
        dec[j].logits = dec[0].logits;
        dec[j].logprobs = dec[0].logprobs;
        dec[j].probs = dec[0].probs;
      }

      for (int i = 0; i < n_max; ++i) {
        std::vector<Cand> cands;
        if (p.strategy == 0) {
          for (int j = 0; j < n_cur; ++j) {
            Decoder& d = dec[j];
            if (d.completed || d.failed) continue;
            ora_token_data tk = sample_token(m, d.probs.data(), d.logprobs.data(), t_cur < 1e-6f, d.rng);
            d.seq.tokens.push_back(tk);
            d.seq.sum_logprobs_all += tk.plog;
          }
        } else {
          // whisper_sample_token_topk: k draws from the decoder's distribution
          for (int j = 0; j < n_cur; ++j) {
            Decoder& d = dec[j];
            if (d.completed || d.failed) continue;
            std::discrete_distribution<> dist(d.probs.begin(), d.probs.end());
            // pt/ptsum/tid are those of the distribution, before the id override
            double sum_ts = 0.0, max_ts = 0.0;
            int tid = hp.token_beg;
            for (int k = hp.token_beg; k < hp.n_vocab; ++k) {
              sum_ts += d.probs[k];
              if (max_ts < d.probs[k]) {
                max_ts = d.probs[k];
                tid = k;
              }
            }
            for (int k = 0; k < p.beam_size; ++k) {
              const int id = dist(d.rng);
              ora_token_data tk = {id, tid, d.probs[id], d.logprobs[id], (float)(max_ts / (sum_ts + 1e-10)),
                                   (float)sum_ts, -1, -1, -1, 0.f};
              if (id >= hp.token_beg) {
                tk.tid = id;
                tk.pt = tk.p;
              }
              Cand c{j, d.seek_delta, d.has_ts, d.seq};
              c.seq.tokens.push_back(tk);
              c.seq.sum_logprobs_all += tk.plog;
              cands.push_back(std::move(c));
            }
          }
          std::stable_sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) {
            if (a.seq.sum_logprobs_all != b.seq.sum_logprobs_all)
              return a.seq.sum_logprobs_all > b.seq.sum_logprobs_all;
            return a.decoder_idx < b.decoder_idx;
          });
          auto same = [](const Sequence& a, const Sequence& b) {
            if (a.tokens.size() != b.tokens.size()) return false;
            for (size_t k = 0; k < a.tokens.size(); ++k)
              if (a.tokens[k].id != b.tokens[k].id) return false;
            return true;
          };
          size_t cc = 0;
          std::vector<int> src(n_cur, -1);
          for (int j = 0; j < n_cur; ++j) {
            Decoder& d = dec[j];
            if (d.completed || d.failed) continue;
            if (cc >= cands.size()) cc = 0;
            Cand& cur = cands[cc++];
            while (cands.size() > cc && same(cands[cc].seq, cur.seq) && i > 0) ++cc;
            d.seek_delta = cur.seek_delta;
            d.has_ts = cur.has_ts;
            d.seq = cur.seq;
            src[j] = cur.decoder_idx;
          }
          const int n_kv = (int)prompt.size() + i;
          for (int j = 0; j < n_cur; ++j)
            if (src[j] >= 0) kv_copy(m, 8 + j, src[j], n_kv);
          for (int j = 0; j < n_cur; ++j)
            if (src[j] >= 0) kv_copy(m, j, 8 + j, n_kv);
        }
        // update decoder state
        for (int j = 0; j < n_cur; ++j) {
          Decoder& d = dec[j];
          if (d.completed || d.failed) continue;
          const ora_token_data& tk = d.seq.tokens.back();
          if (tk.id > hp.token_beg) {
            const int sd_new = 2 * (tk.id - hp.token_beg);
            if (d.has_ts && d.seek_delta > sd_new && d.seq.result_len < i) {
              d.failed = true;
              continue;
            }
            d.seek_delta = sd_new;
            d.seq.result_len = i + 1;
            d.has_ts = true;
          }
          if (tk.id == hp.token_eot || (p.max_tokens > 0 && i >= p.max_tokens) ||
              (d.has_ts && seek + d.seek_delta + delta_min >= seek_end)) {
            if (d.seq.result_len == 0 && !p.no_timestamps) {
              if (seek + d.seek_delta + delta_min >= seek_end)
                d.seq.result_len = i + 1;
              else {
                d.failed = true;
                continue;
              }
            }
            if (p.single_segment || p.no_timestamps) {
              d.seq.result_len = i + 1;
              d.seek_delta = 100 * 30;
            }
            d.completed = true;
            continue;
          }
          if (i == n_max - 1 && (d.seq.result_len == 0 || d.seek_delta < 100 * 30 / 2)) {
            d.failed = true;
            continue;
          }
        }
        bool all_done = true;
        for (int j = 0; j < n_cur; ++j)
          if (!dec[j].completed && !dec[j].failed) all_done = false;
        if (all_done) break;
        // next logits
        for (int j = 0; j < n_cur; ++j) {
          Decoder& d = dec[j];
          if (d.completed || d.failed) continue;
          const int32_t id = d.seq.tokens.back().id;
          if (decode_impl(m, j, &id, 1, (int)prompt.size() + i, d.logits.data(), true)) return -1;
          std::vector<int32_t> ids(d.seq.tokens.size());
          for (size_t k = 0; k < ids.size(); ++k) ids[k] = d.seq.tokens[k].id;
          process_logits_impl(m, &p, ids.data(), (int)ids.size(), d.has_ts, d.seek_delta, t_cur,
                              d.logits.data(), d.logprobs.data(), d.probs.data());
        }
        res->n_decode_steps++;
      }
      // rank
      double best_score = -INFINITY;
      best_id = 0;
      for (int j = 0; j < n_cur; ++j) {
        Decoder& d = dec[j];
        if (d.failed) continue;
        d.seq.tokens.resize(d.seq.result_len);
        sequence_score(&p, d.seq);
        if (d.seq.result_len > 32 && d.seq.entropy < p.entropy_thold) {
          d.failed = true;
          continue;
        }
        if (best_score < d.seq.score) {
          best_score = d.seq.score;
          best_id = j;
        }
      }
      bool success = true;
      const Decoder& bd = dec[best_id];
      if (bd.failed || (bd.seq.avg_logprobs < p.logprob_thold && no_speech_prob < p.no_speech_thold))
        success = false;
      if (success) break;
    }
    res->ms_decode += now_ms() - td;

    // segments
    {
      Decoder& bd = dec[best_id];
      int seek_delta = bd.seek_delta;
      const int result_len = bd.seq.result_len;
      auto& tc = bd.seq.tokens;
      if ((int)tc.size() > result_len) tc.resize(result_len);
      const bool is_no_speech = no_speech_prob > p.no_speech_thold && bd.seq.avg_logprobs < p.logprob_thold;
      prompt_past.clear();
      if (!prompt.empty() && prompt.front() == hp.token_prev)
        prompt_past.insert(prompt_past.end(), prompt.begin() + 1, prompt.end() - prompt_init.size());
      for (int i = 0; i < result_len && !is_no_speech; ++i) prompt_past.push_back(tc[i].id);
      if (!tc.empty() && !is_no_speech) {
        int i0 = 0;
        int64_t t0s = seek + 2 * (tc.front().tid - hp.token_beg);
        std::string text;
        bool turn = false;
        auto emit = [&](int64_t a, int64_t b, int j0, int j1) {
          ora_segment sg;
          sg.t0 = a;
          sg.t1 = b;
          sg.text = text;
          sg.speaker_turn_next = turn;
          for (int j = j0; j <= j1; ++j) sg.tokens.push_back(tc[j]);
          if (p.token_timestamps) token_level_timestamps(m, sg, 0.01f, 0.01f);
          res->segs.push_back(std::move(sg));
        };
        for (int i = 0; i < (int)tc.size(); ++i) {
          if (tc[i].id < hp.token_eot) text += m->id_to_token[tc[i].id];
          if (p.tdrz_enable && tc[i].id == hp.token_solm) turn = true;
          if (tc[i].id > hp.token_beg && !p.single_segment) {
            const int64_t t1s = seek + 2 * (tc[i].tid - hp.token_beg);
            if (!text.empty()) emit(t0s, t1s, i0, i);
            text.clear();
            while (i < (int)tc.size() && tc[i].id > hp.token_beg) i++;
            i--;
            t0s = t1s;
            i0 = i + 1;
            turn = false;
          }
        }
        if (!text.empty()) emit(t0s, seek + seek_delta, i0, (int)tc.size() - 1);
      }
      const bool single_ts_end = tc.size() > 1 && tc[tc.size() - 2].id < hp.token_beg &&
                                 tc[tc.size() - 1].id > hp.token_beg;
      if (single_ts_end) seek_delta = std::min(seek_end - seek, 30 * 100);
      seek += seek_delta;
    }
  }
  return 0;
}

int ora_result_n_segments(const ora_result* r) { return (int)r->segs.size(); }
const char* ora_result_segment_text(const ora_result* r, int i) { return r->segs[i].text.c_str(); }
int64_t ora_result_segment_t0(const ora_result* r, int i) { return r->segs[i].t0; }
int64_t ora_result_segment_t1(const ora_result* r, int i) { return r->segs[i].t1; }
int ora_result_segment_speaker_turn_next(const ora_result* r, int i) {
  return r->segs[i].speaker_turn_next;
}
int ora_result_n_tokens(const ora_result* r, int i) { return (int)r->segs[i].tokens.size(); }
ora_token_data ora_result_token_data(const ora_result* r, int i, int j) { return r->segs[i].tokens[j]; }
int ora_result_lang_id(const ora_result* r) { return r->lang_id; }
int ora_result_n_decode_steps(const ora_result* r) { return r->n_decode_steps; }
int ora_result_n_windows(const ora_result* r) { return r->n_windows; }
double ora_result_ms_mel(const ora_result* r) { return r->ms_mel; }
double ora_result_ms_encode(const ora_result* r) { return r->ms_encode; }
double ora_result_ms_decode(const ora_result* r) { return r->ms_decode; }
void ora_result_free(ora_result* r) { delete r; }

}  // extern "C"
