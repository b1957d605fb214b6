// C entry point of the WAV front door for the tests (host/wav.h); part of libstt_engine.so.
#include <string.h>

#include "wav.h"

extern "C" __attribute__((visibility("default"))) long stt_parse_wav(const char* bytes, size_t n, short* out,
                                                                     size_t cap, int* sample_rate, int* channels) {
  try {
    const sentiric::utils::DecodedAudio a = sentiric::utils::parse_wav_robust(std::string(bytes, n));
    if (!a.is_valid) return -1;
    *sample_rate = a.sample_rate;
    *channels = a.channels;
    const size_t m = a.pcm_data.size() < cap ? a.pcm_data.size() : cap;
    if (m) memcpy(out, a.pcm_data.data(), m * 2);
    return (long)a.pcm_data.size();
  } catch (const std::exception&) {
    return -1;
  }
}

#include "text_filters.h"
extern "C" __attribute__((visibility("default"))) int stt_is_hallucination(const char* text) {
  return sentiric::utils::is_hallucination(std::string(text)) ? 1 : 0;
}
