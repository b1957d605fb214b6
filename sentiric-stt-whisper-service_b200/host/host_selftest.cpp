// CPU self-test of the host-side pieces that need no GPU: segment filters, speaker clustering,
// model file validation, and the constructor's failure contract.
#include <assert.h>
#include <stdio.h>

#include <filesystem>
#include <fstream>

#include "model_manager.h"
#include "stt_engine.h"
#include "text_filters.h"

int main(int argc, char** argv) {
  using sentiric::utils::is_hallucination;
  // utils.h:214-306 decisions
  assert(is_hallucination(""));
  assert(is_hallucination("  .  "));
  assert(is_hallucination("a"));
  assert(is_hallucination("[music]"));
  assert(is_hallucination("(applause)"));
  assert(is_hallucination("Thanks for watching"));
  assert(is_hallucination("  thank you!  "));
  assert(is_hallucination("Altyazı M.K."));
  assert(!is_hallucination("visit www.example.org"));  // 4-byte phrases only match a whole segment
  assert(is_hallucination("Okay."));
  assert(is_hallucination("Hmm"));
  assert(is_hallucination("ehem..."));
  assert(!is_hallucination("Okay, let us start the meeting."));
  assert(!is_hallucination("hello world"));
  assert(!is_hallucination("The senkr report"));
  // speaker_cluster.cpp:19-39
  SpeakerClusterer c(0.88f);
  const std::vector<float> a = {1, 0, 0, 0, 0, 0, 0, 0}, b = {0.98f, 0.05f, 0, 0, 0, 0, 0, 0}, d = {0, 1, 0, 0, 0, 0, 0, 0};
  assert(c.assign_or_add(a) == "spk_0");
  assert(c.assign_or_add(b) == "spk_0");
  assert(c.assign_or_add(d) == "spk_1");
  assert(c.clusters().size() == 2 && c.clusters()[0].count == 2);
  // model_manager.cpp:39-80: a <= 1 MiB file is corrupt and removed; a missing one is reported
  const std::string dir = argc > 1 ? argv[1] : "/tmp/sw_host_selftest";
  Settings s;
  s.model_dir = dir;
  s.model_filename = "ggml-small-junk.bin";
  {
    std::filesystem::create_directories(dir);
    std::ofstream f(dir + "/" + s.model_filename, std::ios::binary);
    f << "tiny";
  }
  assert(ModelManager::ensure_model(s).empty());
  assert(!std::filesystem::exists(dir + "/" + s.model_filename));
  s.enable_vad = false;
  assert(ModelManager::ensure_vad_model(s).empty());
  // stt_engine.cpp:34: the constructor throws std::runtime_error when the model cannot be loaded
  bool threw = false;
  try {
    SttEngine e(s);
  } catch (const std::runtime_error& ex) {
    threw = std::string(ex.what()) == "Whisper model initialization failed";
  }
  assert(threw);
  printf("HOST_SELFTEST_OK\n");
  return 0;
}
