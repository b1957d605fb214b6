// ModelManager: resolves and validates the ggml model file the engine loads (reference:
// /root/reference/src/model_manager.{h,cpp}). ensure_model keeps the reference's contract - path =
// model_dir/model_filename, a file of <= 1 MiB counts as corrupt and is removed - but never
// downloads (the reference forks `curl`; there is no network here and fetching is control-plane
// work). The parse + bf16 upload into HBM that north_star moves "under model_manager" is
// load_to_device(), a thin wrapper over sw_ctx_create.
#pragma once
#include <string>

#include "../../include/sw_whisper.h"
#include "config.h"

class ModelManager {
 public:
  // returns the validated path, or "" when the file is missing / too small (and logs why)
  static std::string ensure_model(const Settings& settings);
  static std::string ensure_vad_model(const Settings& settings);
  // ggml .bin -> bf16 weights resident in the HBM of settings.gpu_device; nullptr on failure
  static sw_ctx* load_to_device(const Settings& settings, const std::string& path, int max_beams);

 private:
  static std::string ensure_file(const std::string& dir, const std::string& filename, size_t min_size_bytes);
};
