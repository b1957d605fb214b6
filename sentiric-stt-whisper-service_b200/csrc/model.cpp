// ggml .bin -> HBM loader (see model.h). Two passes over the file: index the tensors, size one
// device arena, then stream each tensor through a pinned staging buffer and convert f16/f32 -> bf16
// on the device. Q/K/V (and cross K/V) weights are concatenated so each projection group is one GEMM.
#include "model.h"

#include <stdio.h>
#include <string.h>

#include <map>

#include "host_common.h"
#include "kernels.cuh"

namespace sw {

namespace {

const char* const k_langs[] = {
    "en", "zh", "de", "es", "ru", "ko", "fr", "ja", "pt", "tr", "pl", "ca", "nl", "ar", "sv", "it",
    "id", "hi", "fi", "vi", "he", "uk", "el", "ms", "cs", "ro", "da", "hu", "ta", "no", "th", "ur",
    "hr", "bg", "lt", "la", "mi", "ml", "cy", "sk", "te", "fa", "lv", "bn", "sr", "az", "sl", "kn",
    "et", "mk", "br", "eu", "is", "hy", "ne", "mn", "bs", "kk", "sq", "sw", "gl", "mr", "pa", "si",
    "km", "sn", "yo", "so", "af", "oc", "ka", "be", "tg", "sd", "gu", "am", "yi", "lo", "uz", "fo",
    "ht", "ps", "tk", "nn", "mt", "sa", "lb", "my", "bo", "tl", "mg", "as", "tt", "haw", "ln", "ha",
    "ba", "jw", "su", "yue"};
constexpr int k_n_langs = sizeof(k_langs) / sizeof(k_langs[0]);

// symbols whisper.cpp matches against the vocabulary for suppress_nst
const char* const k_non_speech[] = {
    "\"", "#", "(", ")", "*", "+", "/", ":", ";", "<", "=", ">", "@", "[", "\\", "]", "^", "_", "`",
    "{", "|", "}", "~", "「", "」", "『", "』", "<<", ">>", "<<<", ">>>", "--", "---", "-(", "-[",
    "('", "(\"", "((", "))", "(((", ")))", "[[", "]]", "{{", "}}", "♪♪", "♪♪♪", "♩", "♪", "♫", "♬",
    "♭", "♮", "♯"};

// ---- ggml tensor types this loader reads (ggml.h enum ggml_type): f32, f16 and the five "legacy" block
// quantisations whisper.cpp's quantize tool writes (q4_0, q4_1, q5_0, q5_1, q8_0). Blocks of 32 values along
// ne[0]; layouts restated from ggml's block_q* structs and dequantize_row_q* and checked against the independent
// `gguf` Python package (tests/test_quantized_models.py). Quantised matrices are dequantised to f32 on the host
// at load and then take the same f32 -> bf16 path as an f32 file: HBM always holds bf16.
constexpr int QK = 32;
inline bool type_supported(int t) { return t == 0 || t == 1 || t == 2 || t == 3 || t == 6 || t == 7 || t == 8; }
inline bool type_quantised(int t) { return t >= 2; }
inline size_t block_bytes(int t) {
  switch (t) {
    case 2: return 2 + 16;       // q4_0: f16 d, 16 bytes of nibbles
    case 3: return 2 + 2 + 16;   // q4_1: f16 d, f16 m, nibbles
    case 6: return 2 + 4 + 16;   // q5_0: f16 d, 32 high bits, nibbles
    case 7: return 2 + 2 + 4 + 16;  // q5_1: f16 d, f16 m, high bits, nibbles
    case 8: return 2 + 32;       // q8_0: f16 d, 32 int8
    default: return 0;
  }
}
inline size_t tensor_bytes(int t, int64_t nel) {
  if (t == 0) return (size_t)nel * 4;
  if (t == 1) return (size_t)nel * 2;
  return (size_t)(nel / QK) * block_bytes(t);
}
inline float half_to_float(const uint8_t* p) {
  _Float16 h;
  memcpy(&h, p, 2);
  return (float)h;
}
// n values (a multiple of 32) of type t at src -> f32
void dequantise(int t, const uint8_t* src, float* dst, int64_t n) {
  const size_t bb = block_bytes(t);
  for (int64_t b = 0; b < n / QK; ++b, src += bb, dst += QK) {
    const float d = half_to_float(src);
    if (t == 8) {
      const int8_t* q = reinterpret_cast<const int8_t*>(src + 2);
      for (int j = 0; j < QK; ++j) dst[j] = q[j] * d;
      continue;
    }
    const bool has_m = t == 3 || t == 7, has_h = t == 6 || t == 7;
    const float m = has_m ? half_to_float(src + 2) : 0.f;
    const uint8_t* p = src + 2 + (has_m ? 2 : 0);
    uint32_t qh = 0;
    if (has_h) {
      memcpy(&qh, p, 4);
      p += 4;
    }
    for (int j = 0; j < QK / 2; ++j) {
      int x0 = p[j] & 0x0F, x1 = p[j] >> 4;
      if (has_h) {
        x0 |= ((qh >> (j + 0)) << 4) & 0x10;
        x1 |= (qh >> (j + 12)) & 0x10;
      }
      if (has_m) {  // q4_1 / q5_1: unsigned code * d + m
        dst[j] = x0 * d + m;
        dst[j + QK / 2] = x1 * d + m;
      } else {      // q4_0 / q5_0: signed around the middle code
        const int off = has_h ? 16 : 8;
        dst[j] = (x0 - off) * d;
        dst[j + QK / 2] = (x1 - off) * d;
      }
    }
  }
}

struct TInfo {
  int n_dims = 0, ttype = 0;
  int64_t ne[4] = {1, 1, 1, 1};
  int64_t nel = 0;
  long offset = 0;
};

struct Loader {
  FILE* f = nullptr;
  std::map<std::string, TInfo> idx;
  uint8_t* h_stage = nullptr;  // pinned
  uint8_t* d_stage = nullptr;
  size_t stage_bytes = 0;
  cudaStream_t stream = nullptr;
  std::vector<uint8_t> raw;  // a quantised tensor as stored
  ~Loader() {
    if (f) fclose(f);
    if (h_stage) cudaFreeHost(h_stage);
    if (d_stage) cudaFree(d_stage);
  }
  const TInfo* find(const std::string& n) {
    auto it = idx.find(n);
    if (it == idx.end()) {
      set_last_error("model file lacks tensor '%s'", n.c_str());
      return nullptr;
    }
    return &it->second;
  }
  int read_host(const TInfo& t, void* dst) {
    const size_t bytes = tensor_bytes(t.ttype, t.nel);
    SW_CHECK(fseek(f, t.offset, SEEK_SET) == 0, "seek failed");
    SW_CHECK(fread(dst, 1, bytes, f) == bytes, "model file truncated");
    return 0;
  }
  // matrix tensor -> bf16 at dst (device)
  int matrix(const std::string& name, __nv_bfloat16* dst, int64_t expect) {
    const TInfo* t = find(name);
    if (!t) return -1;
    SW_CHECK(t->nel == expect, "tensor '%s' has %lld elements, expected %lld", name.c_str(),
             (long long)t->nel, (long long)expect);
    size_t bytes = tensor_bytes(t->ttype, t->nel);
    SW_CHECK(bytes <= stage_bytes && (!type_quantised(t->ttype) || (size_t)t->nel * 4 <= stage_bytes),
             "staging buffer too small for '%s'", name.c_str());
    if (type_quantised(t->ttype)) {  // blocks -> f32 on the host, then the f32 path
      raw.resize(bytes);
      if (read_host(*t, raw.data())) return -1;
      dequantise(t->ttype, raw.data(), reinterpret_cast<float*>(h_stage), t->nel);
      bytes = (size_t)t->nel * 4;
    } else if (read_host(*t, h_stage)) {
      return -1;
    }
    SW_CUDA_CHECK(cudaMemcpyAsync(d_stage, h_stage, bytes, cudaMemcpyHostToDevice, stream));
    if (t->ttype == 1) {
      if (convert_f16_to_bf16(reinterpret_cast<const uint16_t*>(d_stage), dst, t->nel, stream)) return -1;
    } else {
      if (convert_f32_to_bf16(reinterpret_cast<const float*>(d_stage), dst, t->nel, stream)) return -1;
    }
    SW_CUDA_CHECK(cudaStreamSynchronize(stream));
    return 0;
  }
  // any tensor -> f32 host vector
  int host_f32(const std::string& name, std::vector<float>& out, int64_t expect) {
    const TInfo* t = find(name);
    if (!t) return -1;
    SW_CHECK(t->nel == expect, "tensor '%s' has %lld elements, expected %lld", name.c_str(),
             (long long)t->nel, (long long)expect);
    out.resize(t->nel);
    if (t->ttype == 0) return read_host(*t, out.data());
    if (type_quantised(t->ttype)) {
      raw.resize(tensor_bytes(t->ttype, t->nel));
      if (read_host(*t, raw.data())) return -1;
      dequantise(t->ttype, raw.data(), out.data(), t->nel);
      return 0;
    }
    std::vector<uint16_t> tmp(t->nel);
    if (read_host(*t, tmp.data())) return -1;
    for (int64_t i = 0; i < t->nel; ++i) {
      _Float16 h;
      memcpy(&h, &tmp[i], 2);
      out[i] = (float)h;
    }
    return 0;
  }
  int vector_f32(const std::string& name, float* dst, int64_t expect) {
    std::vector<float> v;
    if (host_f32(name, v, expect)) return -1;
    SW_CUDA_CHECK(cudaMemcpy(dst, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
    return 0;
  }
};

template <typename T>
T* arena_alloc(Model* m, int64_t n) {
  size_t off = (m->arena_used + 255) & ~size_t(255);
  const size_t bytes = (size_t)n * sizeof(T);
  if (off + bytes > m->arena_bytes) return nullptr;
  m->arena_used = off + bytes;
  return reinterpret_cast<T*>(m->arena + off);
}

int build_vocab(Model* m, FILE* f) {
  const HParams& hp = m->hp;
  Vocab& v = m->vocab;
  int32_t n_file = 0;
  SW_CHECK(fread(&n_file, 4, 1, f) == 1 && n_file >= 0 && n_file <= hp.n_vocab + 100000, "bad vocab size");
  v.id_to_token.assign(hp.n_vocab, std::string());
  for (int i = 0; i < n_file; ++i) {
    uint32_t len = 0;
    SW_CHECK(fread(&len, 4, 1, f) == 1 && len < 4096, "bad vocab entry");
    std::string s(len, '\0');
    if (len) SW_CHECK(fread(&s[0], 1, len, f) == len, "vocab truncated");
    if (i < hp.n_vocab) {
      v.id_to_token[i] = s;
      v.token_to_id[s] = i;
    }
  }
  v.multilingual = hp.n_vocab >= 51865;
  if (v.multilingual) {
    v.n_langs = hp.n_vocab - 51765 - 1;
    const int dt = v.n_langs - 98;
    v.eot += 1;
    v.sot += 1;
    v.translate += dt;
    v.transcribe += dt;
    v.solm += dt;
    v.prev += dt;
    v.nosp += dt;
    v.not_ += dt;
    v.beg += dt;
  }
  SW_CHECK(v.n_langs <= k_n_langs, "vocabulary implies %d languages", v.n_langs);
  for (int i = n_file; i < hp.n_vocab; ++i) {
    char buf[64];
    if (i > v.beg) snprintf(buf, sizeof(buf), "[_TT_%d]", i - v.beg);
    else if (i == v.eot) snprintf(buf, sizeof(buf), "[_EOT_]");
    else if (i == v.sot) snprintf(buf, sizeof(buf), "[_SOT_]");
    else if (i == v.translate) snprintf(buf, sizeof(buf), "[_TRANSLATE_]");
    else if (i == v.transcribe) snprintf(buf, sizeof(buf), "[_TRANSCRIBE_]");
    else if (i == v.solm) snprintf(buf, sizeof(buf), "[_SOLM_]");
    else if (i == v.prev) snprintf(buf, sizeof(buf), "[_PREV_]");
    else if (i == v.nosp) snprintf(buf, sizeof(buf), "[_NOSP_]");
    else if (i == v.not_) snprintf(buf, sizeof(buf), "[_NOT_]");
    else if (i == v.beg) snprintf(buf, sizeof(buf), "[_BEG_]");
    else if (i > v.sot && i <= v.sot + v.n_langs) snprintf(buf, sizeof(buf), "[_LANG_%s]", k_langs[i - v.sot - 1]);
    else snprintf(buf, sizeof(buf), "[_extra_token_%d]", i);
    v.id_to_token[i] = buf;
    v.token_to_id[buf] = i;
  }
  std::vector<char> seen(hp.n_vocab, 0);
  auto add = [&](const std::string& s) {
    auto it = v.token_to_id.find(s);
    if (it != v.token_to_id.end() && !seen[it->second]) {
      seen[it->second] = 1;
      v.nst_ids.push_back(it->second);
    }
  };
  for (const char* s : k_non_speech) {
    add(s);
    add(std::string(" ") + s);
  }
  add(" -");
  add(" '");
  auto sp = v.token_to_id.find(" ");
  v.space = sp == v.token_to_id.end() ? -1 : sp->second;
  return 0;
}

int load_impl(Model* m, const char* path) {
  Loader L;
  L.f = fopen(path, "rb");
  SW_CHECK(L.f != nullptr, "cannot open model file '%s'", path);
  uint32_t magic = 0;
  SW_CHECK(fread(&magic, 4, 1, L.f) == 1 && magic == 0x67676d6c, "'%s' is not a ggml model (bad magic)", path);
  int32_t h[11];
  SW_CHECK(fread(h, 4, 11, L.f) == 11, "model header truncated");
  HParams& hp = m->hp;
  hp.n_vocab = h[0]; hp.n_audio_ctx = h[1]; hp.n_audio_state = h[2]; hp.n_audio_head = h[3];
  hp.n_audio_layer = h[4]; hp.n_text_ctx = h[5]; hp.n_text_state = h[6]; hp.n_text_head = h[7];
  hp.n_text_layer = h[8]; hp.n_mels = h[9]; hp.ftype = h[10];
  SW_CHECK(hp.n_audio_state == hp.n_text_state, "encoder/decoder widths differ");
  SW_CHECK(hp.n_audio_state % 128 == 0 && hp.n_audio_state == 64 * hp.n_audio_head &&
               hp.n_text_state == 64 * hp.n_text_head,
           "unsupported model width %d / heads %d", hp.n_audio_state, hp.n_audio_head);
  SW_CHECK(hp.n_audio_ctx == 1500 && hp.n_text_ctx == 448, "unsupported context sizes %d/%d",
           hp.n_audio_ctx, hp.n_text_ctx);
  SW_CHECK(hp.n_mels == 80 || hp.n_mels == 128, "unsupported n_mels %d", hp.n_mels);
  SW_CHECK(hp.n_vocab >= 51864 && hp.n_vocab <= 52000, "unsupported n_vocab %d", hp.n_vocab);
  {
    // ggml_ftype, possibly with the quantisation version in the thousands (GGML_QNT_VERSION_FACTOR): all f32 (0),
    // mostly f16 (1), mostly q4_0 (2), q4_1 (3), q8_0 (7), q5_0 (8), q5_1 (9); k-quants (10+) are not read
    const int ft = hp.ftype % 1000;
    SW_CHECK(ft == 0 || ft == 1 || ft == 2 || ft == 3 || ft == 7 || ft == 8 || ft == 9,
             "ggml ftype %d is not supported (f32, f16, q4_0, q4_1, q5_0, q5_1, q8_0 are)", hp.ftype);
  }
  int32_t fm[2];
  SW_CHECK(fread(fm, 4, 2, L.f) == 2 && fm[0] == hp.n_mels && fm[1] == 201, "bad mel filterbank header");
  std::vector<float> filters((size_t)fm[0] * fm[1]);
  SW_CHECK(fread(filters.data(), 4, filters.size(), L.f) == filters.size(), "filterbank truncated");
  if (build_vocab(m, L.f)) return -1;

  // pass 1: index
  size_t max_bytes = 0;
  while (true) {
    int32_t hd[3];
    if (fread(hd, 4, 3, L.f) != 3) break;
    TInfo t;
    t.n_dims = hd[0];
    t.ttype = hd[2];
    SW_CHECK(t.n_dims >= 1 && t.n_dims <= 4 && hd[1] > 0 && hd[1] < 256, "corrupt tensor header");
    SW_CHECK(type_supported(t.ttype), "ggml tensor type %d is not supported (f32, f16, q4_0, q4_1, q5_0, q5_1, q8_0 are)",
             t.ttype);
    t.nel = 1;
    for (int i = 0; i < t.n_dims; ++i) {
      int32_t v;
      SW_CHECK(fread(&v, 4, 1, L.f) == 1 && v > 0, "corrupt tensor dims");
      t.ne[i] = v;
      t.nel *= v;
    }
    std::string name(hd[1], '\0');
    SW_CHECK(fread(&name[0], 1, hd[1], L.f) == (size_t)hd[1], "corrupt tensor name");
    t.offset = ftell(L.f);
    SW_CHECK(!type_quantised(t.ttype) || t.ne[0] % QK == 0, "quantised tensor '%s': row length %lld is not a multiple of 32",
             name.c_str(), (long long)t.ne[0]);
    const size_t bytes = tensor_bytes(t.ttype, t.nel);
    SW_CHECK(fseek(L.f, (long)bytes, SEEK_CUR) == 0, "seek failed");
    max_bytes = bytes > max_bytes ? bytes : max_bytes;
    if (type_quantised(t.ttype)) max_bytes = std::max(max_bytes, (size_t)t.nel * 4);  // staged as f32
    L.idx[name] = t;
  }
  {
    // the last tensor must be fully present
    fseek(L.f, 0, SEEK_END);
    const long fsz = ftell(L.f);
    for (auto& kv : L.idx) {
      const size_t bytes = tensor_bytes(kv.second.ttype, kv.second.nel);
      SW_CHECK(kv.second.offset + (long)bytes <= fsz, "model file truncated in tensor '%s'", kv.first.c_str());
    }
  }

  const int64_t d = hp.n_audio_state, nm = hp.n_mels, V = hp.n_vocab;
  const int Le = hp.n_audio_layer, Ld = hp.n_text_layer;
  // arena size: bf16 matrices + f32 vectors, generous alignment slack
  size_t need = 0;
  need += (size_t)(d * 3 * nm + d * 3 * d) * 2 + (size_t)(1500 * d + 448 * d + hp.n_mels * 201 + 8 * d) * 4;
  need += (size_t)Le * ((size_t)12 * d * d * 2 + (size_t)16 * d * 4);
  need += (size_t)Ld * ((size_t)16 * d * d * 2 + (size_t)24 * d * 4);
  need += (size_t)V * d * 2;
  need += (size_t)(Le + Ld) * 40 * 256 + (1 << 20);
  SW_CUDA_CHECK(cudaMalloc(&m->arena, need));
  m->arena_bytes = need;
  SW_CUDA_CHECK(cudaMemset(m->arena, 0, need));
  L.stage_bytes = max_bytes;
  SW_CUDA_CHECK(cudaMallocHost(&L.h_stage, max_bytes));
  SW_CUDA_CHECK(cudaMalloc(&L.d_stage, max_bytes));
  SW_CUDA_CHECK(cudaStreamCreate(&L.stream));

#define ALLOC(ptr, T, n)                                       \
  do {                                                         \
    ptr = arena_alloc<T>(m, (n));                              \
    SW_CHECK(ptr != nullptr, "weight arena exhausted");        \
  } while (0)
  auto ln = [&](LayerNormW& w, const std::string& p) -> int {
    ALLOC(w.g, float, d);
    ALLOC(w.b, float, d);
    if (L.vector_f32(p + ".weight", w.g, d)) return -1;
    return L.vector_f32(p + ".bias", w.b, d);
  };

  ALLOC(m->filters, float, nm * 201);
  SW_CUDA_CHECK(cudaMemcpy(m->filters, filters.data(), filters.size() * 4, cudaMemcpyHostToDevice));
  {  // the triangles are a few bins wide: the front end only walks the groups of four bins that hold a weight
    std::vector<int2> span(nm);
    for (size_t i = 0; i < (size_t)nm; ++i) {
      int lo = 201, hi = 0;
      for (int k = 0; k < 201; ++k)
        if (filters[i * 201 + k] != 0.0f) {
          lo = std::min(lo, k);
          hi = std::max(hi, k + 1);
        }
      if (lo >= hi) lo = hi = 0;
      span[i].x = lo / 4 * 4;
      span[i].y = hi;
    }
    ALLOC(m->filter_span, int2, nm);
    SW_CUDA_CHECK(cudaMemcpy(m->filter_span, span.data(), span.size() * sizeof(int2), cudaMemcpyHostToDevice));
  }
  // conv weights: [d_out][c_in][3] -> [d_out][3][c_in]
  auto conv = [&](const std::string& name, int64_t cin, __nv_bfloat16*& dst) -> int {
    std::vector<float> w, r((size_t)d * 3 * cin);
    if (L.host_f32(name, w, d * cin * 3)) return -1;
    for (int64_t o = 0; o < d; ++o)
      for (int64_t c = 0; c < cin; ++c)
        for (int k = 0; k < 3; ++k) r[(o * 3 + k) * cin + c] = w[(o * cin + c) * 3 + k];
    ALLOC(dst, __nv_bfloat16, d * 3 * cin);
    SW_CHECK(r.size() * 4 <= L.stage_bytes, "staging buffer too small for conv weights");
    SW_CUDA_CHECK(cudaMemcpy(L.d_stage, r.data(), r.size() * 4, cudaMemcpyHostToDevice));
    if (convert_f32_to_bf16(reinterpret_cast<const float*>(L.d_stage), dst, (int64_t)r.size(), L.stream)) return -1;
    SW_CUDA_CHECK(cudaStreamSynchronize(L.stream));
    return 0;
  };
  if (conv("encoder.conv1.weight", nm, m->conv1_w)) return -1;
  if (conv("encoder.conv2.weight", d, m->conv2_w)) return -1;
  ALLOC(m->conv1_b, float, d);
  ALLOC(m->conv2_b, float, d);
  if (L.vector_f32("encoder.conv1.bias", m->conv1_b, d)) return -1;
  if (L.vector_f32("encoder.conv2.bias", m->conv2_b, d)) return -1;
  ALLOC(m->enc_pos, float, 1500 * d);
  if (L.vector_f32("encoder.positional_embedding", m->enc_pos, 1500 * d)) return -1;

  m->enc.resize(Le);
  for (int i = 0; i < Le; ++i) {
    EncLayerW& w = m->enc[i];
    const std::string p = "encoder.blocks." + std::to_string(i);
    if (ln(w.ln1, p + ".attn_ln")) return -1;
    ALLOC(w.wqkv, __nv_bfloat16, 3 * d * d);
    ALLOC(w.bqkv, float, 3 * d);
    if (L.matrix(p + ".attn.query.weight", w.wqkv, d * d)) return -1;
    if (L.matrix(p + ".attn.key.weight", w.wqkv + d * d, d * d)) return -1;
    if (L.matrix(p + ".attn.value.weight", w.wqkv + 2 * d * d, d * d)) return -1;
    if (L.vector_f32(p + ".attn.query.bias", w.bqkv, d)) return -1;
    if (L.vector_f32(p + ".attn.value.bias", w.bqkv + 2 * d, d)) return -1;
    ALLOC(w.wo, __nv_bfloat16, d * d);
    ALLOC(w.bo, float, d);
    if (L.matrix(p + ".attn.out.weight", w.wo, d * d)) return -1;
    if (L.vector_f32(p + ".attn.out.bias", w.bo, d)) return -1;
    if (ln(w.ln2, p + ".mlp_ln")) return -1;
    ALLOC(w.w1, __nv_bfloat16, 4 * d * d);
    ALLOC(w.b1, float, 4 * d);
    ALLOC(w.w2, __nv_bfloat16, 4 * d * d);
    ALLOC(w.b2, float, d);
    if (L.matrix(p + ".mlp.0.weight", w.w1, 4 * d * d)) return -1;
    if (L.vector_f32(p + ".mlp.0.bias", w.b1, 4 * d)) return -1;
    if (L.matrix(p + ".mlp.2.weight", w.w2, 4 * d * d)) return -1;
    if (L.vector_f32(p + ".mlp.2.bias", w.b2, d)) return -1;
  }
  if (ln(m->ln_post, "encoder.ln_post")) return -1;

  ALLOC(m->dec_pos, float, 448 * d);
  if (L.vector_f32("decoder.positional_embedding", m->dec_pos, 448 * d)) return -1;
  ALLOC(m->tok_emb, __nv_bfloat16, V * d);
  if (L.matrix("decoder.token_embedding.weight", m->tok_emb, V * d)) return -1;
  m->dec.resize(Ld);
  for (int i = 0; i < Ld; ++i) {
    DecLayerW& w = m->dec[i];
    const std::string p = "decoder.blocks." + std::to_string(i);
    if (ln(w.ln1, p + ".attn_ln")) return -1;
    ALLOC(w.wqkv, __nv_bfloat16, 3 * d * d);
    ALLOC(w.bqkv, float, 3 * d);
    if (L.matrix(p + ".attn.query.weight", w.wqkv, d * d)) return -1;
    if (L.matrix(p + ".attn.key.weight", w.wqkv + d * d, d * d)) return -1;
    if (L.matrix(p + ".attn.value.weight", w.wqkv + 2 * d * d, d * d)) return -1;
    if (L.vector_f32(p + ".attn.query.bias", w.bqkv, d)) return -1;
    if (L.vector_f32(p + ".attn.value.bias", w.bqkv + 2 * d, d)) return -1;
    ALLOC(w.wo, __nv_bfloat16, d * d);
    ALLOC(w.bo, float, d);
    if (L.matrix(p + ".attn.out.weight", w.wo, d * d)) return -1;
    if (L.vector_f32(p + ".attn.out.bias", w.bo, d)) return -1;
    if (ln(w.lnx, p + ".cross_attn_ln")) return -1;
    ALLOC(w.wxq, __nv_bfloat16, d * d);
    ALLOC(w.bxq, float, d);
    if (L.matrix(p + ".cross_attn.query.weight", w.wxq, d * d)) return -1;
    if (L.vector_f32(p + ".cross_attn.query.bias", w.bxq, d)) return -1;
    ALLOC(w.wxkv, __nv_bfloat16, 2 * d * d);
    ALLOC(w.bxkv, float, 2 * d);
    if (L.matrix(p + ".cross_attn.key.weight", w.wxkv, d * d)) return -1;
    if (L.matrix(p + ".cross_attn.value.weight", w.wxkv + d * d, d * d)) return -1;
    if (L.vector_f32(p + ".cross_attn.value.bias", w.bxkv + d, d)) return -1;
    ALLOC(w.wxo, __nv_bfloat16, d * d);
    ALLOC(w.bxo, float, d);
    if (L.matrix(p + ".cross_attn.out.weight", w.wxo, d * d)) return -1;
    if (L.vector_f32(p + ".cross_attn.out.bias", w.bxo, d)) return -1;
    if (ln(w.ln2, p + ".mlp_ln")) return -1;
    ALLOC(w.w1, __nv_bfloat16, 4 * d * d);
    ALLOC(w.b1, float, 4 * d);
    ALLOC(w.w2, __nv_bfloat16, 4 * d * d);
    ALLOC(w.b2, float, d);
    if (L.matrix(p + ".mlp.0.weight", w.w1, 4 * d * d)) return -1;
    if (L.vector_f32(p + ".mlp.0.bias", w.b1, 4 * d)) return -1;
    if (L.matrix(p + ".mlp.2.weight", w.w2, 4 * d * d)) return -1;
    if (L.vector_f32(p + ".mlp.2.bias", w.b2, d)) return -1;
  }
  if (ln(m->dec_ln, "decoder.ln")) return -1;
#undef ALLOC
  // per-step streamed weights: self qkv+o (4d^2), cross q+o (2d^2), mlp (8d^2), logits matrix
  m->weight_bytes_decoder = ((size_t)Ld * 14 * d * d + (size_t)V * d) * 2;
  cudaStreamDestroy(L.stream);
  L.stream = nullptr;
  return 0;
}

}  // namespace

Model::~Model() {
  if (arena) cudaFree(arena);
}

Model* load_model(const char* path) {
  Model* m = new Model();
  if (load_impl(m, path)) {
    delete m;
    return nullptr;
  }
  return m;
}

int lang_id(const char* lang) {
  if (!lang) return -1;
  for (int i = 0; i < k_n_langs; ++i)
    if (strcmp(lang, k_langs[i]) == 0) return i;
  return -1;
}
const char* lang_code(int id) { return (id >= 0 && id < k_n_langs) ? k_langs[id] : ""; }

}  // namespace sw
