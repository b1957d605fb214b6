// See prosody_oracle.h. Each block cites the reference lines it restates.
#include "prosody_oracle.h"

#include <math.h>

#include <algorithm>
#include <vector>

namespace {

// prosody_extractor.cpp:9-28
float mean_of(const std::vector<float>& v) {
  if (v.empty()) return 0.0f;
  float s = 0.0f;
  for (float x : v) s += x;  // std::accumulate with a float zero: sequential single-precision sum
  return s / v.size();
}
float stdev_of(const std::vector<float>& v, float mean) {
  if (v.empty()) return 0.0f;
  float acc = 0.0f;
  for (float x : v) acc += (x - mean) * (x - mean);
  return sqrtf(acc / v.size());
}
float median_of(std::vector<float> v) {  // upper median: element size/2 of the sorted order
  if (v.empty()) return 0.0f;
  const size_t n = v.size() / 2;
  std::nth_element(v.begin(), v.begin() + n, v.end());
  return v[n];
}
float soft_norm(float val, float lo, float hi) {
  const float t = (val - lo) / (hi - lo);
  return std::max(0.0f, std::min(1.0f, t));
}

}  // namespace

extern "C" void ora_prosody_extract(const float* pcm, size_t n, int sr, const ora_prosody_opts* o, ora_prosody* out) {
  *out = ora_prosody();
  out->gender = '?';
  out->emotion = 0;
  if (n < 160 || !pcm) return;  // :35-47: every number stays 0, speaker_vec = 8 zeros

  const int shift = sr / 100;  // :49 10 ms frames
  std::vector<float> f0s, rmses, zcrs, scs;
  int peaks = 0;
  float last_rms = 0.0f, lpf = 0.0f;  // the one-pole low-pass state runs through the whole segment (:58-59)
  const int fs = std::min(shift, 1600);
  std::vector<float> filt(1600);
  for (size_t i = 0; i + shift <= n; i += shift) {  // :62
    float r0 = 0.0f;
    for (int k = 0; k < fs; ++k) {  // :68-75
      const float x = pcm[i + k];
      r0 += x * x;
      lpf += o->lpf_alpha * (x - lpf);
      filt[k] = lpf;
    }
    const float rms = sqrtf(r0 / fs);  // :76
    rmses.push_back(rms);
    if (rms > 0.05f && last_rms <= 0.05f) ++peaks;  // :79-82 syllable-like onsets
    last_rms = rms;

    const float clip = std::max(0.002f, rms * 0.15f);  // :84
    int cycles = 0, zc = 0;
    bool positive = false, started = false;
    for (int k = 1; k < fs; ++k) {  // :90-108 zero crossings and hysteresis cycle count on the filtered frame
      const float v = filt[k];
      if ((v >= 0) != (filt[k - 1] >= 0)) ++zc;
      if (!started) {
        if (v > clip) positive = true, started = true;
        else if (v < -clip) positive = false, started = true;
      } else if (positive && v < -clip) {
        positive = false;
        ++cycles;
      } else if (!positive && v > clip) {
        positive = true;
      }
    }
    zcrs.push_back((float)zc / fs);  // :109-110

    if (rms > 0.015f && cycles > 0) {  // :112-117
      const float dur = (float)shift / sr;
      const float f0 = cycles / dur;
      if (f0 >= o->min_pitch && f0 <= o->max_pitch) f0s.push_back(f0);
    }
    float power = 0.0f, weighted = 0.0f;  // :119-125 first-difference "centroid"
    for (int k = 1; k < fs; ++k) {
      const float d = fabsf(pcm[i + k] - pcm[i + k - 1]);
      weighted += d * k;
      power += d;
    }
    scs.push_back(power > 0 ? weighted / power : 0.0f);
  }

  // :128-133
  out->pitch_mean = median_of(f0s);
  out->pitch_std = f0s.empty() ? 0.0f : stdev_of(f0s, mean_of(f0s));
  out->energy_mean = rmses.empty() ? 0.01f : mean_of(rmses);
  out->energy_std = rmses.empty() ? 0.0f : stdev_of(rmses, out->energy_mean);
  out->spectral_centroid = scs.empty() ? 50.0f : mean_of(scs);
  out->zero_crossing_rate = zcrs.empty() ? 0.1f : mean_of(zcrs);

  // :138-146 octave corrections
  const bool high = out->pitch_mean > o->gender_threshold, low_zcr = out->zero_crossing_rate < 0.024f;
  if (high && low_zcr) out->pitch_mean *= 0.5f;
  else if (out->energy_mean > 0.12f && out->pitch_mean < 240.0f && out->spectral_centroid < 90.0f)
    out->pitch_mean *= 0.5f;

  const float dur_s = (float)n / sr;  // :148-150
  const float rate = dur_s > 0 ? (float)peaks / dur_s : 0.0f;

  // :153-162 gender proxy
  if (out->pitch_mean == 0.0f || out->energy_mean < 0.018f) out->gender = '?';
  else if (out->zero_crossing_rate < 0.030f) out->gender = 'M';
  else out->gender = out->pitch_mean > o->gender_threshold ? 'F' : 'M';

  // :165-185 valence / arousal / emotion
  const float np = out->gender == 'M' ? soft_norm(out->pitch_mean, 60.0f, 180.0f) : soft_norm(out->pitch_mean, 160.0f, 350.0f);
  const float nb = soft_norm(out->spectral_centroid, 40.0f, 150.0f);
  out->valence = ((np * 0.4f) + (nb * 0.6f)) * 2.0f - 1.0f;
  out->valence += 0.35f;
  const float ne = soft_norm(out->energy_mean, 0.02f, 0.20f), nr = soft_norm(rate, 2.0f, 9.0f);
  out->arousal = (ne * 0.7f) + (nr * 0.3f);
  if (out->arousal > 0.65f) out->emotion = out->valence > 0.1f ? 1 : 3;
  else if (out->arousal < 0.30f) out->emotion = out->valence < -0.4f ? 2 : 0;
  else out->emotion = 0;

  // :190-221 speaker vector
  float base;
  if (out->gender == 'M') base = soft_norm(out->pitch_mean, 60.0f, 200.0f) * 0.4f;
  else if (out->gender == 'F') base = 0.6f + (soft_norm(out->pitch_mean, 160.0f, 350.0f) * 0.4f);
  else base = 0.5f;
  float* s = out->speaker_vec;
  s[0] = base;
  s[1] = soft_norm(out->spectral_centroid, 40.0f, 250.0f);
  s[4] = soft_norm(out->zero_crossing_rate, 0.0f, 0.5f) * 0.8f;
  s[2] = soft_norm(out->pitch_std, 5.0f, 100.0f) * 0.1f;
  s[3] = soft_norm(out->energy_mean, 0.0f, 0.3f) * 0.1f;
  s[5] = soft_norm(rate, 1.0f, 12.0f) * 0.1f;
  s[6] = out->arousal * 0.05f;
  s[7] = ((out->valence + 1.0f) / 2.0f) * 0.05f;
}

// speaker_cluster.cpp:5-38. The reference walks an unordered_map, so exact ties between two
// clusters are broken in an unspecified order there; here the earlier cluster wins.
extern "C" void ora_speaker_cluster(const float* vecs, int n, float threshold, int* ids) {
  struct Cl {
    float c[8];
    size_t count;
  };
  std::vector<Cl> cls;
  for (int i = 0; i < n; ++i) {
    const float* v = vecs + 8 * i;
    int best = -1;
    float best_sim = 0.0f;
    for (size_t k = 0; k < cls.size(); ++k) {
      float dot = 0, na = 0, nb = 0;
      for (int j = 0; j < 8; ++j) {
        dot += v[j] * cls[k].c[j];
        na += v[j] * v[j];
        nb += cls[k].c[j] * cls[k].c[j];
      }
      const float sim = (na == 0 || nb == 0) ? 0.0f : dot / (sqrtf(na) * sqrtf(nb));
      if (sim > best_sim) best_sim = sim, best = (int)k;
    }
    if (best >= 0 && best_sim >= threshold) {
      Cl& c = cls[best];
      for (int j = 0; j < 8; ++j) c.c[j] = (c.c[j] * c.count + v[j]) / (c.count + 1);
      ++c.count;
      ids[i] = best;
    } else {
      Cl c;
      for (int j = 0; j < 8; ++j) c.c[j] = v[j];
      c.count = 1;
      cls.push_back(c);
      ids[i] = (int)cls.size() - 1;
    }
  }
}
