// Transport-free restatement of the service's streaming policy (reference:
// /root/reference/src/grpc_server.cpp:129-305, WhisperTranscribeStream).
// The reference keeps one growing int16 buffer per stream and re-transcribes ALL of it every
// `stream_buffer_samples` new samples (partial result), finalises on an empty chunk (end of
// sentence) and force-finalises past 30 s. Whisper's encoder is not causal, so nothing of a previous
// pass can be reused; what the B200 engine adds is that the re-transcriptions of concurrent streams
// meet in SttEngine's dispatcher and share device passes (one batch of up to max_batch windows, two
// lanes), instead of queueing for a state each. The gRPC plumbing stays in the service; it feeds
// chunks in and writes the events out.
#pragma once
#include <cstdint>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "stt_engine.h"

struct StreamEvent {
  bool is_final = false;                    // set_is_final (:152, :258, :281)
  std::string transcription;                // partials: the segments' texts joined with ' ' (:239-241)
  TranscriptionResult last;                 // the (last) segment the affective fields / words come from
  std::vector<TranscriptionResult> words_of;  // finals carry per-token words of their own segment (:168-174)
};

class StreamSession {
 public:
  using TranscribeFn = std::function<std::vector<TranscriptionResult>(const std::vector<int16_t>&)>;
  // transcribe: engine->transcribe_pcm16(buffer, 16000, RequestOptions()) in the service
  StreamSession(TranscribeFn transcribe, size_t stream_buffer_samples)
      : transcribe_(std::move(transcribe)), step_(stream_buffer_samples) {}
  explicit StreamSession(SttEngine* engine)
      : StreamSession([engine](const std::vector<int16_t>& b) { return engine->transcribe_pcm16(b, 16000, RequestOptions()); },
                      (size_t)engine->get_settings().stream_buffer_samples) {}

  // One request of the stream (request.audio_chunk()). Returns the responses the reference would write.
  std::vector<StreamEvent> feed(const std::string& chunk) {
    std::vector<StreamEvent> out;
    if (chunk.empty()) {  // :140-184 end-of-sentence signal
      if (!buffer_.empty()) {
        for (auto& r : transcribe_(buffer_))
          if (!r.text.empty()) out.push_back(final_event(r));
        buffer_.clear();
        last_processed_ = 0;
      }
      return out;
    }
    const uint8_t* data = reinterpret_cast<const uint8_t*>(chunk.data());
    size_t len = chunk.size();
    if (first_) {  // :189-195 a RIFF/WAVE container: skip its 44-byte header
      if (chunk.size() >= 12 && memcmp(chunk.data(), "RIFF", 4) == 0 && memcmp(chunk.data() + 8, "WAVE", 4) == 0) {
        wav_ = true;
        if (chunk.size() > 44) skip_ = 44;
      }
      first_ = false;
    }
    if (wav_ && skip_ > 0) {  // :197-206
      if (len >= skip_) {
        data += skip_;
        len -= skip_;
        skip_ = 0;
      } else {
        skip_ -= len;
        len = 0;
      }
    }
    if (len > 0) {  // :208-213
      const size_t n = len / 2, cur = buffer_.size();
      buffer_.resize(cur + n);
      memcpy(buffer_.data() + cur, data, n * 2);
    }
    if (buffer_.size() - last_processed_ >= step_) {  // :216 partial pass over the WHOLE buffer
      std::vector<TranscriptionResult> results;
      try {
        results = transcribe_(buffer_);
      } catch (const std::exception&) {  // :300-303 logged, the stream goes on
        return out;
      }
      last_processed_ = buffer_.size();
      StreamEvent partial;
      bool any = false;
      for (auto& r : results)
        if (!r.text.empty()) {
          partial.transcription += r.text + " ";
          partial.last = r;
          any = true;
        }
      if (any) out.push_back(partial);  // :262-268 is_final = false
      if (buffer_.size() > kMaxBuffer) {  // :273-298 30 s without a pause: finalise everything
        for (auto& r : results)
          if (!r.text.empty()) {
            StreamEvent e;
            e.is_final = true;
            e.transcription = r.text;
            e.last = r;
            out.push_back(e);
          }
        buffer_.clear();
        last_processed_ = 0;
      }
    }
    return out;
  }
  size_t buffered_samples() const { return buffer_.size(); }

 private:
  static constexpr size_t kMaxBuffer = 16000 * 30;  // :137
  static StreamEvent final_event(const TranscriptionResult& r) {
    StreamEvent e;
    e.is_final = true;
    e.transcription = r.text;
    e.last = r;
    e.words_of.push_back(r);
    return e;
  }
  TranscribeFn transcribe_;
  size_t step_;
  std::vector<int16_t> buffer_;
  size_t last_processed_ = 0, skip_ = 0;
  bool first_ = true, wav_ = false;
};
