// Route B driver (test infrastructure): built by this directory's Makefile (target routeb) together with the reference's own, unmodified
// stt_engine.cpp / prosody_extractor.cpp / speaker_cluster.cpp. It includes the REFERENCE's stt_engine.h and
// uses nothing but the reference's public API; the JSON it prints has the layout of host/stt_cli.cpp's batch
// mode, so a test can hold the two facades side by side on the same clip.
//   route_b_cli <model_dir> <model_file> <pcm16.raw> [beam=1] [sample_rate=16000] [language=en]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fstream>

#include "stt_engine.h"  // /root/reference/src/stt_engine.h

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
  if (argc < 4) return 2;
  Settings s;
  s.model_dir = argv[1];
  s.model_filename = argv[2];
  s.beam_size = argc > 4 ? atoi(argv[4]) : 1;
  const int sample_rate = argc > 5 ? atoi(argv[5]) : 16000;
  s.language = argc > 6 ? argv[6] : "en";
  s.parallel_requests = 1;
  s.enable_vad = false;
  std::ifstream f(argv[3], std::ios::binary);
  std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  std::vector<int16_t> pcm(raw.size() / 2);
  memcpy(pcm.data(), raw.data(), pcm.size() * 2);
  try {
    SttEngine engine(s);
    SttEngine::PerformanceMetrics met{};
    std::vector<TranscriptionResult> out = engine.transcribe_pcm16(pcm, sample_rate, RequestOptions(), &met);
    printf("{\"request\": 0, \"busy\": 0, \"token_count\": %d, \"segments\": [", met.token_count);
    for (size_t k = 0; k < out.size(); ++k) {
      const TranscriptionResult& r = out[k];
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
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
