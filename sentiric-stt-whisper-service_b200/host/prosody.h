// Affective / prosody tags attached to every segment (reference:
// /root/reference/src/prosody_extractor.h). The DSP itself (prosody_extractor.cpp:31-224, SURVEY.md
// §8(f) rank 3) runs on the GPU for all segments of an utterance at once (csrc/prosody.cu behind
// sw_prosody_segments_*, bit-identical to the reference's host code); SttEngine::set_prosody_fn can
// replace it with a host function. neutral_prosody is what the reference returns for a segment too
// short to analyse (prosody_extractor.cpp:35-47).
#pragma once
#include <cstddef>
#include <functional>
#include <string>
#include <vector>

struct AffectiveTags {
  std::string gender_proxy;
  std::string emotion_proxy;
  float arousal = 0.0f;
  float valence = 0.0f;
  float pitch_mean = 0.0f;
  float pitch_std = 0.0f;
  float energy_mean = 0.0f;
  float energy_std = 0.0f;
  float spectral_centroid = 0.0f;
  float zero_crossing_rate = 0.0f;
  std::vector<float> speaker_vec;
};

struct ProsodyOptions {
  float lpf_alpha = 0.07f;
  float gender_threshold = 170.0f;
  float min_pitch = 60.0f;
  float max_pitch = 500.0f;
};

using ProsodyFn = std::function<AffectiveTags(const float*, size_t, int, const ProsodyOptions&)>;

inline AffectiveTags neutral_prosody(const float*, size_t, int, const ProsodyOptions&) {
  AffectiveTags t;
  t.gender_proxy = "?";
  t.emotion_proxy = "neutral";
  t.speaker_vec.assign(8, 0.0f);
  return t;
}
