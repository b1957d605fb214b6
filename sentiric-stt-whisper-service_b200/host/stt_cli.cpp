// Command-line driver of the SttEngine facade (tests + manual smoke; the reference's only E2E tool
// is the gRPC `stt_cli`, /root/reference/src/cli/). Usage:
//   stt_cli <model_dir> <model_file> <raw 16 kHz mono int16 file> [n_concurrent=1] [beam=1]
// Prints one JSON object per request: segments with t0/t1 (centiseconds), text, per-token p / t0 / t1.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fstream>
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
  Settings s;
  s.model_dir = argv[1];
  s.model_filename = argv[2];
  const int n_conc = argc > 4 ? atoi(argv[4]) : 1;
  s.beam_size = argc > 5 ? atoi(argv[5]) : 1;
  s.language = "en";
  s.parallel_requests = std::max(1, n_conc);
  s.max_batch = 8;
  s.batch_window_us = 20000;
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
  try {
    SttEngine engine(s);
    std::vector<std::vector<TranscriptionResult>> out(n_conc);
    std::vector<SttEngine::PerformanceMetrics> met(n_conc);
    std::vector<std::thread> th;
    for (int i = 0; i < n_conc; ++i)
      th.emplace_back([&, i] { out[i] = engine.transcribe_pcm16(pcm, sample_rate, RequestOptions(), &met[i]); });
    for (auto& t : th) t.join();
    for (int i = 0; i < n_conc; ++i) {
      printf("{\"request\": %d, \"token_count\": %d, \"batches_run\": %ld, \"segments\": [", i, met[i].token_count,
             engine.batches_run());
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
