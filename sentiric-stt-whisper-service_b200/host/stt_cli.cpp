// Command-line driver of the SttEngine facade (tests + manual smoke; the reference's only E2E tool
// is the gRPC `stt_cli`, /root/reference/src/cli/). Usage:
//   stt_cli <model_dir> <model_file> <raw 16 kHz mono int16 file> [n_concurrent=1] [beam=1]
// Prints one JSON object per request: segments with t0/t1 (centiseconds), text, per-token p / t0 / t1.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <math.h>

#include <atomic>
#include <chrono>
#include <fstream>
#include <map>
#include <thread>

#include "stream_session.h"
#include "stt_engine.h"

static std::string json_escape(const std::string& s) {
  std::string o;
  char buf[8];
  for (unsigned char c : s) {
    if (c == '"' || c == '\\') {
      o += '\\';
      o += (char)c;
    } else if (c < 0x20 || c >= 0x7f) {
      snprintf(buf, sizeof(buf), "\\u%04x", c);
      o += buf;
    } else {
      o += (char)c;
    }
  }
  return o;
}

int main(int argc, char** argv) {
  if (argc < 4) {
    fprintf(stderr, "usage: %s <model_dir> <model_file> <pcm16.raw> [n_concurrent] [beam] [stream|batch] [sample_rate]\n", argv[0]);
    return 2;
  }
  // trailing key=value arguments (any position after the raw file): split them off the positional ones
  std::map<std::string, std::string> kv;
  {
    int w = 1;
    for (int i = 1; i < argc; ++i) {
      const char* eq = i >= 4 ? strchr(argv[i], '=') : nullptr;
      if (eq) kv[std::string(argv[i], eq - argv[i])] = eq + 1;
      else argv[w++] = argv[i];
    }
    argc = w;
  }
  auto opt = [&](const char* k, int def) { return kv.count(k) ? atoi(kv[k].c_str()) : def; };
  Settings s;
  s.model_dir = argv[1];
  s.model_filename = argv[2];
  const int n_conc = argc > 4 ? atoi(argv[4]) : 1;
  s.beam_size = argc > 5 ? atoi(argv[5]) : 1;
  s.language = kv.count("language") ? kv["language"] : "en";
  s.parallel_requests = std::max(1, n_conc);
  s.max_batch = opt("max_batch", 8);
  s.batch_window_us = opt("batch_window_us", 20000);
  s.admission_slots = opt("admission", 0);
  s.request_queue_timeout_ms = opt("timeout_ms", s.request_queue_timeout_ms);
  s.n_threads = opt("n_threads", s.n_threads);
  s.gpu_device = opt("device", 0);
  s.enable_vad = opt("enable_vad", 0) != 0;  // the tests run without the gate unless they ask for it
  RequestOptions ropt;
  ropt.beam_size = opt("req_beam", -1);
  ropt.best_of = opt("req_best_of", -1);
  if (kv.count("req_temperature")) ropt.temperature = (float)atof(kv["req_temperature"].c_str());
  // vad=energy: a stand-in gate for the hook (mean |x| above 1e-3 = speech); the reference's Silero model is not here
  auto install_vad = [&](SttEngine& e) {
    if (kv.count("vad") && kv["vad"] == "energy")
      e.set_vad_fn([](const float* x, size_t n) {
        double a = 0;
        for (size_t i = 0; i < n; ++i) a += fabs(x[i]);
        return n > 0 && a / n > 1e-3;
      });
  };
  std::ifstream f(argv[3], std::ios::binary);
  std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  std::vector<int16_t> pcm(raw.size() / 2);
  memcpy(pcm.data(), raw.data(), pcm.size() * 2);
  const bool stream_mode = argc > 6 && std::string(argv[6]) == "stream";
  const int sample_rate = argc > 7 ? atoi(argv[7]) : 16000;
  if (stream_mode) {
    // n_conc concurrent streams (grpc_server.cpp:129-305 policy through StreamSession), each fed the clip in
    // 0.5 s chunks and closed with an empty chunk: their re-transcriptions meet in the dispatcher
    try {
      SttEngine engine(s);
      install_vad(engine);
      std::vector<std::string> finals(n_conc);
      std::vector<int> partials(n_conc, 0), calls(n_conc, 0);
      std::vector<std::thread> th;
      for (int i = 0; i < n_conc; ++i)
        th.emplace_back([&, i] {
          StreamSession ss(
              [&, i](const std::vector<int16_t>& b) {
                ++calls[i];
                return engine.transcribe_pcm16(b, 16000, RequestOptions());
              },
              (size_t)engine.get_settings().stream_buffer_samples);
          const size_t step = 8000;
          for (size_t off = 0; off < pcm.size(); off += step) {
            const size_t n = std::min(step, pcm.size() - off);
            for (auto& e : ss.feed(std::string(reinterpret_cast<const char*>(pcm.data() + off), n * 2)))
              partials[i] += e.is_final ? 0 : 1;
          }
          for (auto& e : ss.feed(""))
            if (e.is_final) finals[i] += e.transcription + "|";
        });
      for (auto& t : th) t.join();
      int total_calls = 0;
      for (int c : calls) total_calls += c;
      printf("{\"streams\": %d, \"transcribe_calls\": %d, \"batches_run\": %ld, \"partials\": %d, \"finals\": [", n_conc,
             total_calls, engine.batches_run(), partials[0]);
      for (int i = 0; i < n_conc; ++i) printf("%s\"%s\"", i ? ", " : "", json_escape(finals[i]).c_str());
      printf("]}\n");
      return 0;
    } catch (const std::exception& e) {
      fprintf(stderr, "error: %s\n", e.what());
      return 1;
    }
  }
  const std::string mode = argc > 6 ? argv[6] : "batch";
  if (mode == "bench") {
    // The reference-facing front door under load: n_conc caller threads, each handing its own 30 s clip to
    // SttEngine::transcribe_pcm16 as a pageable std::vector (what the HTTP / gRPC handlers do), `steps` timed
    // rounds after `warmup` untimed ones. The raw file holds the clips back to back (clip=<samples per clip>).
    try {
      const size_t clip = (size_t)opt("clip", 480000);
      const int n_clips = std::max<int>(1, (int)(pcm.size() / clip));
      const int steps = opt("steps", 3), warmup = opt("warmup", 1);
      SttEngine engine(s);
      install_vad(engine);
      std::vector<std::vector<int16_t>> clips(n_conc);
      for (int i = 0; i < n_conc; ++i) {
        const size_t off = (size_t)(i % n_clips) * clip;
        clips[i].assign(pcm.begin() + off, pcm.begin() + std::min(pcm.size(), off + clip));
      }
      std::atomic<long> tokens{0}, segments{0}, failures{0};
      auto round = [&] {
        std::vector<std::thread> th;
        for (int i = 0; i < n_conc; ++i)
          th.emplace_back([&, i] {
            SttEngine::PerformanceMetrics m{};
            try {
              auto r = engine.transcribe_pcm16(clips[i], sample_rate, ropt, &m);
              tokens += m.token_count;
              segments += (long)r.size();
            } catch (const std::exception&) {
              ++failures;
            }
          });
        for (auto& t : th) t.join();
      };
      for (int k = 0; k < warmup; ++k) round();
      tokens = 0; segments = 0;
      const long b0 = engine.batches_run();
      const auto t0 = std::chrono::steady_clock::now();
      for (int k = 0; k < steps; ++k) round();
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      double audio_s = 0;
      for (auto& c : clips) audio_s += (double)c.size() / sample_rate;
      printf("{\"mode\": \"bench\", \"callers\": %d, \"steps\": %d, \"seconds\": %.6f, \"audio_s_per_s\": %.3f, "
             "\"ms_per_step\": %.3f, \"tokens\": %ld, \"segments\": %ld, \"failures\": %ld, \"device_passes\": %ld, "
             "\"beam\": %d, \"language\": \"%s\"}\n",
             n_conc, steps, dt, audio_s * steps / dt, 1e3 * dt / steps, tokens.load(), segments.load(), failures.load(),
             engine.batches_run() - b0, ropt.beam_size >= 0 ? ropt.beam_size : s.beam_size, s.language.c_str());
      return 0;
    } catch (const std::exception& e) {
      fprintf(stderr, "error: %s\n", e.what());
      return 1;
    }
  }
  try {
    SttEngine engine(s);
    install_vad(engine);
    std::vector<std::vector<TranscriptionResult>> out(n_conc);
    std::vector<SttEngine::PerformanceMetrics> met(n_conc);
    std::vector<int> busy(n_conc, 0);
    std::vector<std::thread> th;
    for (int i = 0; i < n_conc; ++i)
      th.emplace_back([&, i] {
        try {
          out[i] = engine.transcribe_pcm16(pcm, sample_rate, ropt, &met[i]);
        } catch (const EngineBusyException& ex) {  // stt_engine.cpp:70-74
          busy[i] = std::string(ex.what()) == "Server is busy (Queue timeout)" ? 1 : -1;
        }
      });
    for (auto& t : th) t.join();
    for (int i = 0; i < n_conc; ++i) {
      printf("{\"request\": %d, \"busy\": %d, \"token_count\": %d, \"batches_run\": %ld, \"segments\": [", i, busy[i],
             met[i].token_count, engine.batches_run());
      for (size_t k = 0; k < out[i].size(); ++k) {
        const TranscriptionResult& r = out[i][k];
        printf("%s{\"t0\": %lld, \"t1\": %lld, \"prob\": %.6f, \"language\": \"%s\", \"speaker\": \"%s\", \"text\": \"%s\", ",
               k ? ", " : "", (long long)r.t0, (long long)r.t1, r.prob, r.language.c_str(), r.speaker_id.c_str(),
               json_escape(r.text).c_str());
        printf("\"gender\": \"%s\", \"emotion\": \"%s\", \"prosody\": [%.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g], \"tokens\": [",
               r.gender_proxy.c_str(), r.emotion_proxy.c_str(), r.affective.arousal, r.affective.valence, r.affective.pitch_mean,
               r.affective.pitch_std, r.affective.energy_mean, r.affective.energy_std, r.affective.spectral_centroid,
               r.affective.zero_crossing_rate);
        for (size_t j = 0; j < r.tokens.size(); ++j)
          printf("%s[\"%s\", %.6f, %lld, %lld]", j ? ", " : "", json_escape(r.tokens[j].text).c_str(), r.tokens[j].p,
                 (long long)r.tokens[j].t0, (long long)r.tokens[j].t1);
        printf("]}");
      }
      printf("]}\n");
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
