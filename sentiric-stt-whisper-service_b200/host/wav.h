// Codec front door of the service (reference: /root/reference/src/utils.h:101-202, has_wav_header and
// parse_wav_robust, called from http_server.cpp:161 and grpc_server.cpp:48):
// RIFF/WAVE chunk walk, 16-bit PCM only, stereo mixed to mono as (l + r) / 2, more channels -> channel 0.
// The reference tries ffmpeg on input without a RIFF header and otherwise assumes raw 16 kHz mono int16;
// ffmpeg is not part of this tree, so the fallback is the raw-PCM assumption directly.
// Pinned against the reference's own code: tests/test_wav.py compares with a build of the reference's utils.h.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace sentiric::utils {

struct DecodedAudio {
  std::vector<int16_t> pcm_data;
  int sample_rate = 16000;
  int channels = 1;
  bool is_valid = false;
};

inline bool has_wav_header(const std::string& bytes) {  // utils.h:101-105
  return bytes.size() >= 12 && memcmp(bytes.data(), "RIFF", 4) == 0 && memcmp(bytes.data() + 8, "WAVE", 4) == 0;
}

inline DecodedAudio parse_wav_robust(const std::string& bytes) {
  DecodedAudio out;
  if (!has_wav_header(bytes)) {  // utils.h:111-137 (without the ffmpeg attempt): raw 16 kHz mono int16
    out.pcm_data.resize(bytes.size() / 2);
    if (!out.pcm_data.empty()) memcpy(out.pcm_data.data(), bytes.data(), out.pcm_data.size() * 2);
    out.is_valid = true;
    return out;
  }
  const uint8_t* data = reinterpret_cast<const uint8_t*>(bytes.data());
  const size_t n = bytes.size();
  size_t pos = 12, pcm_bytes = 0;
  const uint8_t* pcm = nullptr;
  int16_t bits = 0;
  bool have_fmt = false;
  while (pos + 8 < n) {  // utils.h:145-175 chunk walk
    const uint8_t* id = data + pos;
    uint32_t size;
    memcpy(&size, data + pos + 4, 4);
    pos += 8;
    if (pos + size > n) break;
    if (memcmp(id, "fmt ", 4) == 0) {
      if (size < 16) throw std::runtime_error("Invalid fmt");
      uint16_t tag;
      memcpy(&tag, data + pos, 2);
      if (tag != 1 && tag != 0xFFFE) throw std::runtime_error("Unsupported WAV tag");
      // the reference copies 2 bytes into the low half of an int initialised to 1 (little endian)
      memcpy(&out.channels, data + pos + 2, 2);
      memcpy(&out.sample_rate, data + pos + 4, 4);
      memcpy(&bits, data + pos + 14, 2);
      // not in the reference: a header with 0 channels reaches `n_samples / channels` below (SIGFPE on an
      // untrusted upload, utils.h:194), and a 0 Hz rate only fails much later; both are malformed files
      if (out.channels < 1) throw std::runtime_error("Invalid channel count");
      if (out.sample_rate <= 0) throw std::runtime_error("Invalid sample rate");
      have_fmt = true;
      pos += size;
    } else if (memcmp(id, "data", 4) == 0) {
      if (!have_fmt) throw std::runtime_error("No fmt chunk");
      pcm = data + pos;
      pcm_bytes = size;
      break;
    } else {
      pos += size;
    }
    if (size % 2 != 0 && pos < n) ++pos;  // chunks are word aligned
  }
  if (!pcm || pcm_bytes == 0) throw std::runtime_error("No data chunk");
  if (bits != 16) throw std::runtime_error("Unsupported bit depth");
  const size_t left = n - (size_t)(pcm - data);
  if (pcm_bytes > left) pcm_bytes = left;
  const size_t n_samples = pcm_bytes / 2;
  std::vector<int16_t> raw(n_samples);  // the payload need not be 2-byte aligned inside the container
  if (n_samples) memcpy(raw.data(), pcm, n_samples * 2);
  if (out.channels == 1) {
    out.pcm_data = std::move(raw);
  } else if (out.channels == 2) {  // utils.h:186-193
    out.pcm_data.resize(n_samples / 2);
    for (size_t i = 0; i < out.pcm_data.size(); ++i)
      out.pcm_data[i] = static_cast<int16_t>(((int32_t)raw[2 * i] + (int32_t)raw[2 * i + 1]) / 2);
  } else {  // utils.h:194-199 first channel only
    const size_t frames = n_samples / out.channels;
    out.pcm_data.resize(frames);
    for (size_t i = 0; i < frames; ++i) out.pcm_data[i] = raw[i * out.channels];
  }
  out.is_valid = true;
  return out;
}

}  // namespace sentiric::utils
