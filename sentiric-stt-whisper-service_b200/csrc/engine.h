// Engine context behind the C ABI: model in HBM, activation/KV buffers sized for max_batch
// windows, and the device pipelines (front end, encoder, one batched decoder step).
// The host sequencer (sequencer.cpp) drives them; capi.cpp exposes them.
#pragma once
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/sw_whisper.h"
#include "host_common.h"
#include "kernels.cuh"
#include "model.h"

namespace sw {

struct StageTimes {  // accumulated device time per stage (CUDA events), for the benchmark
  double ms_mel = 0, ms_encode = 0, ms_decode = 0;
  long n_windows = 0, n_steps = 0, n_launches = 0;
  double h2d_bytes = 0, d2h_bytes = 0;  // PCM in, picks / energy out
  double ms_xattn = 0;        // device time inside cross_attention launches (kernel timing on)
  long n_xattn = 0;
  double xattn_bytes = 0;     // algorithmic bytes of those launches (cross-KV of the active windows)
  double decode_bytes = 0;  // algorithmic bytes streamed by the decode steps (weights + cross-KV + self-KV)
};

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  int alloc(size_t count) {
    n = count;
    SW_CUDA_CHECK(cudaMalloc(&p, count * sizeof(T)));
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
};
template <typename T>
struct PinBuf {
  T* p = nullptr;
  size_t n = 0;
  int alloc(size_t count) {
    n = count;
    SW_CUDA_CHECK(cudaMallocHost(&p, count * sizeof(T)));
    return 0;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = 0;
  }
  ~PinBuf() { release(); }
};

struct Engine {
  int device = 0;
  int max_batch = 64, max_beams = 5, max_rows = 320;
  Model* model = nullptr;
  bool owns_model = true;   // false for the further lanes of a context (they share lane 0's weights)
  int xa_max_ctas = 0;      // cap on the persistent cross-attention grid (0 = every SM); contexts with
                            // several lanes leave SMs to the other lane's latency-bound kernels
  size_t buffer_bytes = 0;  // device memory of this lane's buffers (without the weights)
  cudaStream_t stream = nullptr;
  std::mutex mu;  // one batch in flight per context

  // ---- front end
  DevBuf<uint8_t> d_pcm;        // batch PCM (int16 or f32)
  size_t pcm_capacity = 0;
  DevBuf<float> d_log;          // log-mel before normalisation, all utterances of the batch
  size_t log_capacity = 0;
  DevBuf<MelUtt> d_utts;
  DevBuf<unsigned> d_max_enc;
  DevBuf<int> d_win_utt, d_win_seek;
  DevBuf<float> d_energy;       // token-timestamp signal energy, same offsets as the PCM; never leaves HBM
  DevBuf<float> d_eblk;         // per 256-sample block min | max of the energy (two halves)
  size_t eblk_capacity = 0;
  size_t energy_capacity = 0;
  // token-time refinement of a finished batch (token_times.cu), on its own stream: it runs from the host
  // thread that post-processes batch i while the main stream already decodes batch i + 1
  cudaStream_t post_stream = nullptr;
  // interleaved halves of a decoder step (engine.cu::enqueue_decode_step): second stream of the step graph and
  // the events that order the two halves' cross attentions
  bool interleave = false;
  int interleave_min_groups = 8;  // per half
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::vector<cudaEvent_t> ev_half;
  DevBuf<TtSeg> d_tt_seg;
  PinBuf<TtSeg> h_tt_seg;
  DevBuf<long long> d_tt_t;     // t0 | t1 (two halves of tt_tok_capacity)
  PinBuf<long long> h_tt_t;
  DevBuf<uint8_t> d_tt_flag;
  PinBuf<uint8_t> h_tt_flag;
  DevBuf<float> d_tt_thold;
  size_t tt_seg_capacity = 0, tt_tok_capacity = 0;
  int utt_capacity = 0;
  // ---- encoder activations (max_batch windows)
  DevBuf<bf16> conv_in, h1, hb, qkv, ff;
  DevBuf<float> x;
  DevBuf<bf16> cross_kv;        // [n_text_layer][max_batch*1500][2d]
  // ---- decoder
  DevBuf<float> dx, logits, xa_ws, dpart;  // dpart: split-K partial sums [split][R][d]
  DevBuf<bf16> dh, dqkv, datt, dq, dff;
  DevBuf<bf16> kv_pool;
  int n_pages = 0;
  DevBuf<int> d_page_table, d_tok, d_pos, d_grp_win, d_grp_start, d_grp_count;
  DevBuf<DecRow> d_rows;
  DevBuf<LogitRow> d_lrows;
  DevBuf<PickOut> d_picks;
  DevBuf<uint8_t> d_suppress;
  DevBuf<int> d_copy_pairs;
  PinBuf<int> h_page_table, h_tok, h_pos, h_grp;
  PinBuf<DecRow> h_rows;
  PinBuf<LogitRow> h_lrows;
  PinBuf<PickOut> h_picks;
  int64_t logits_ld = 0;

  StageTimes times;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool legacy_attention = false;
  bool use_graphs = true;       // replay each decoder step from a CUDA graph captured per step shape
  void* step_graphs = nullptr;  // StepGraphCache (engine.cu)
  bool kernel_timing = false;   // bracket each cross_attention launch with events (bench roofline)
  std::vector<cudaEvent_t> xa_ev;

  ~Engine();
};

Engine* engine_create(const char* model_path, const sw_ctx_params* params);
// a further lane next to `primary`: same model (shared, not owned), same limits, own buffers and stream
Engine* engine_create_lane(const Engine* primary);

// conv stem + encoder stack + ln_post + cross-KV for n_win windows whose conv input is in conv_in.
// enc_out_f32 (device, [n_win*1500][d]) may be null.
int engine_encode(Engine* e, int n_win, float* enc_out_f32);

// one decoder step over R rows (h_rows/h_tok/h_pos filled; rows grouped by window in h_grp).
// want_logits: final LN + logits GEMM for all rows (e->logits). If n_lrows > 0, h_lrows[0..n_lrows) are
// processed into h_picks. Synchronises the stream before returning so the host can read them.
int engine_decode_step(Engine* e, int R, int n_groups, int max_count, bool want_logits, int n_lrows,
                       const LogitCfg& cfg, bool upload_page_table);
// device-side copies of whole KV pages (beam reshuffle): pairs [src0, dst0, src1, dst1, ...]
int engine_copy_pages(Engine* e, const std::vector<int>& pairs);

}  // namespace sw
