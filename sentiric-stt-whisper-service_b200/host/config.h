// Settings consumed by SttEngine / ModelManager. Field names and defaults follow the reference's
// `struct Settings` (/root/reference/src/config.h:10-82) so that code constructing it keeps
// compiling; the env-var parser (config.h:84-172) is control-plane code and stays with the service.
#pragma once
#include <algorithm>
#include <string>
#include <thread>

struct Settings {
  std::string host = "0.0.0.0";
  int http_port = 15030;
  int grpc_port = 15031;
  int metrics_port = 15032;

  std::string model_dir = "/models";
  std::string model_filename = "ggml-medium.bin";
  std::string model_url_template =
      "https://huggingface.co/ggerganov/whisper.cpp/resolve/main/ggml-{model_name}.bin";
  int model_load_timeout = 600;

  std::string vad_model_filename = "ggml-silero-vad.bin";
  std::string vad_model_url = "https://huggingface.co/ggml-org/whisper-vad/resolve/main/ggml-silero-v6.2.0.bin";
  bool enable_vad = true;
  float vad_threshold = 0.75f;
  int vad_ms_min_duration = 500;

  int n_threads = std::min(4, (int)std::thread::hardware_concurrency());
  int parallel_requests = 2;
  int request_queue_timeout_ms = 5000;

  std::string device = "auto";
  std::string compute_type = "int8";

  std::string language = "auto";
  bool translate = false;
  bool no_timestamps = false;

  int beam_size = 5;
  float temperature = 0.0f;
  int best_of = 5;
  float logprob_threshold = -0.7f;
  float no_speech_threshold = 0.85f;

  bool flash_attn = true;
  bool suppress_nst = true;

  bool enable_diarization = false;
  float cluster_threshold = 0.88f;

  int sample_rate = 16000;
  int stream_buffer_samples = 8000;

  std::string log_level = "info";
  std::string grpc_ca_path = "";
  std::string grpc_cert_path = "";
  std::string grpc_key_path = "";

  // ---- additions of the B200 engine (ignored by reference code) ----
  int gpu_device = 0;        // CUDA ordinal of this replica (one engine per GPU)
  int max_batch = 64;        // 30 s windows decoded together
  int batch_window_us = 300; // how long the dispatcher waits for more concurrent callers
  int admission_slots = 0;   // callers admitted at once (the reference's state pool size); 0 = auto:
                             // max(parallel_requests, 2 * max_batch), so that the defaults can fill a device pass
};
