// CPU self-test of the host-side pieces that need no GPU: segment filters, speaker clustering,
// model file validation, and the constructor's failure contract.
#include <assert.h>
#include <stdio.h>
#include <string.h>

#include <filesystem>
#include <fstream>

#include "model_manager.h"
#include "stream_session.h"
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
  // grpc_server.cpp:129-305 streaming policy, with a fake engine that reports the buffer length
  {
    std::vector<size_t> passes;
    StreamSession ss(
        [&](const std::vector<int16_t>& b) {
          passes.push_back(b.size());
          TranscriptionResult r{};
          r.text = "n=" + std::to_string(b.size());
          return std::vector<TranscriptionResult>{r};
        },
        8000);
    std::string wav(44 + 2 * 3000, '\0');
    memcpy(&wav[0], "RIFF", 4);
    memcpy(&wav[8], "WAVE", 4);
    assert(ss.feed(wav).empty() && ss.buffered_samples() == 3000);   // header skipped, below the 0.5 s step
    const std::string pcm(2 * 5000, '\1');
    auto ev = ss.feed(pcm);                                          // 8000 samples: one partial over ALL of them
    assert(ev.size() == 1 && !ev[0].is_final && ev[0].transcription == "n=8000 " && ss.buffered_samples() == 8000);
    assert(ss.feed(std::string(2 * 100, '\1')).empty());            // 100 new samples: no pass
    ev = ss.feed("");                                                // end of sentence: final, buffer reset
    assert(ev.size() == 1 && ev[0].is_final && ev[0].transcription == "n=8100" && ss.buffered_samples() == 0);
    assert(ss.feed("").empty());
    for (int i = 0; i < 60; ++i) ev = ss.feed(std::string(2 * 8000, '\1'));  // 30 s without a pause
    assert(ev.size() == 1 && !ev[0].is_final && ss.buffered_samples() == 480000);
    ev = ss.feed(std::string(2 * 8000, '\1'));                      // past 30 s: partial + forced final, reset
    assert(ev.size() == 2 && !ev[0].is_final && ev[1].is_final && ss.buffered_samples() == 0);
    assert(passes.size() == 1 + 1 + 61);
  }
  printf("HOST_SELFTEST_OK\n");
  return 0;
}
