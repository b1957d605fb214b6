// C entry points around the REFERENCE's own extract_prosody / SpeakerClusterer, compiled together with
// /root/reference/src/prosody_extractor.cpp and speaker_cluster.cpp (where they lie; nothing is copied)
// into oracle/_ref/libref_prosody.so. Test infrastructure: pins oracle/prosody_oracle.cpp.
#include <string.h>

#include <string>
#include <vector>

#include "prosody_extractor.h"   // -I/root/reference/src
#include "speaker_cluster.h"
#include "prosody_oracle.h"

extern "C" void ref_prosody_extract(const float* pcm, size_t n, int sr, const ora_prosody_opts* o, ora_prosody* out) {
  ProsodyOptions po;
  po.lpf_alpha = o->lpf_alpha;
  po.gender_threshold = o->gender_threshold;
  po.min_pitch = o->min_pitch;
  po.max_pitch = o->max_pitch;
  const AffectiveTags t = extract_prosody(pcm, n, sr, po);
  memset(out, 0, sizeof(*out));
  out->gender = t.gender_proxy.empty() ? '?' : t.gender_proxy[0];
  out->emotion = t.emotion_proxy == "excited" ? 1 : t.emotion_proxy == "sad" ? 2 : t.emotion_proxy == "angry" ? 3 : 0;
  out->arousal = t.arousal; out->valence = t.valence; out->pitch_mean = t.pitch_mean; out->pitch_std = t.pitch_std;
  out->energy_mean = t.energy_mean; out->energy_std = t.energy_std; out->spectral_centroid = t.spectral_centroid;
  out->zero_crossing_rate = t.zero_crossing_rate;
  for (size_t i = 0; i < 8 && i < t.speaker_vec.size(); ++i) out->speaker_vec[i] = t.speaker_vec[i];
}

extern "C" void ref_speaker_cluster(const float* vecs, int n, float threshold, int* ids) {
  SpeakerClusterer c(threshold);
  for (int i = 0; i < n; ++i) {
    const std::string id = c.assign_or_add(std::vector<float>(vecs + 8 * i, vecs + 8 * i + 8));
    ids[i] = atoi(id.c_str() + 4);  // "spk_<k>"
  }
}
