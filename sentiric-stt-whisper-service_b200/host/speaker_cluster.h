// Per-request online speaker clustering of the 8-D prosody vectors (reference:
// /root/reference/src/speaker_cluster.{h,cpp}: cosine similarity against running-mean centroids,
// new cluster below the threshold). Host-only and trivially cheap; kept so TranscriptionResult
// carries the same speaker ids.
#pragma once
#include <cmath>
#include <string>
#include <vector>

struct SpeakerCluster {
  std::string id;
  std::vector<float> centroid;
  size_t count = 0;
};

class SpeakerClusterer {
 public:
  explicit SpeakerClusterer(float threshold = 0.88f) : threshold_(threshold) {}
  std::string assign_or_add(const std::vector<float>& vec) {
    int best = -1;
    float best_sim = 0.0f;
    for (size_t c = 0; c < clusters_.size(); ++c) {
      const float s = cosine(vec, clusters_[c].centroid);
      if (s > best_sim) {
        best_sim = s;
        best = (int)c;
      }
    }
    if (best >= 0 && best_sim >= threshold_) {
      SpeakerCluster& k = clusters_[best];
      for (size_t i = 0; i < k.centroid.size() && i < vec.size(); ++i)
        k.centroid[i] = (k.centroid[i] * k.count + vec[i]) / (k.count + 1);
      ++k.count;
      return k.id;
    }
    clusters_.push_back({"spk_" + std::to_string(clusters_.size()), vec, 1});
    return clusters_.back().id;
  }
  const std::vector<SpeakerCluster>& clusters() const { return clusters_; }

 private:
  static float cosine(const std::vector<float>& a, const std::vector<float>& b) {
    float dot = 0, na = 0, nb = 0;
    for (size_t i = 0; i < a.size() && i < b.size(); ++i) {
      dot += a[i] * b[i];
      na += a[i] * a[i];
      nb += b[i] * b[i];
    }
    if (na == 0 || nb == 0) return 0;
    return dot / (std::sqrt(na) * std::sqrt(nb));
  }
  float threshold_;
  std::vector<SpeakerCluster> clusters_;
};
