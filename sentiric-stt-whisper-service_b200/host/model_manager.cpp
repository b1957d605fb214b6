#include "model_manager.h"

#include <stdio.h>

#include <filesystem>

namespace fs = std::filesystem;

std::string ModelManager::ensure_model(const Settings& settings) {
  return ensure_file(settings.model_dir, settings.model_filename, 1024 * 1024);
}

std::string ModelManager::ensure_vad_model(const Settings& settings) {
  if (!settings.enable_vad) return "";
  return ensure_file(settings.model_dir, settings.vad_model_filename, 100 * 1024);
}

std::string ModelManager::ensure_file(const std::string& dir, const std::string& filename, size_t min_size_bytes) {
  std::error_code ec;
  const fs::path p = fs::path(dir) / filename;
  if (!fs::exists(dir, ec)) fs::create_directories(dir, ec);
  if (fs::exists(p, ec)) {
    const auto size = fs::file_size(p, ec);
    if (!ec && size > min_size_bytes) return p.string();
    fprintf(stderr, "[model_manager] %s is corrupt or too small (%llu bytes): removed\n", p.c_str(),
            (unsigned long long)size);
    fs::remove(p, ec);
  }
  fprintf(stderr, "[model_manager] %s is missing; provisioning (download) is outside the engine\n", p.c_str());
  return "";
}

sw_ctx* ModelManager::load_to_device(const Settings& settings, const std::string& path, int max_beams) {
  sw_ctx_params cp = sw_ctx_default_params();
  cp.device = settings.gpu_device;
  cp.max_batch = settings.max_batch;
  cp.max_beams = max_beams;
  cp.flash_attn = settings.flash_attn ? 1 : 0;
  return sw_ctx_create(path.c_str(), &cp);
}
