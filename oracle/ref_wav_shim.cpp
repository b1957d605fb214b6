// C entry point around the REFERENCE's own parse_wav_robust (/root/reference/src/utils.h:107-202), compiled
// where it lies (with oracle/ref_stubs/spdlog for its logging include) into oracle/_ref/libref_wav.so.
// Test infrastructure: pins host/wav.h.
#include <string.h>

#include <string>

#include "utils.h"  // -I/root/reference/src

// returns the number of samples (<= cap written to out), or -1 when the reference throws; rate / channels out
extern "C" long ref_parse_wav(const char* bytes, size_t n, short* out, size_t cap, int* sample_rate, int* channels) {
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

// the segment post-filter of stt_engine.cpp:272-278 (utils.h:214-306), for tests/test_text_filters.py
extern "C" int ref_is_hallucination(const char* text) { return sentiric::utils::is_hallucination(std::string(text)) ? 1 : 0; }
