// Post-filter the reference applies to every segment inside SttEngine::transcribe
// (stt_engine.cpp:272; rules: /root/reference/src/utils.h:214-306): a segment is dropped when it
// is empty / pure punctuation / bracketed, contains one of the known subtitle-credit phrases, or
// is nothing but a short filler. Same decisions, table-driven.
#pragma once
#include <algorithm>
#include <cctype>
#include <string>

namespace sentiric {
namespace utils {

inline std::string trim_ws(const std::string& s) {
  const char* ws = " \t\n\r\f\v";
  const size_t a = s.find_first_not_of(ws);
  if (a == std::string::npos) return "";
  return s.substr(a, s.find_last_not_of(ws) - a + 1);
}

inline bool is_hallucination(const std::string& raw_text) {
  static const char* const kPhrases[] = {
      "altyazı", "Altyazı", "ALTYAZI", "sesli betimleme", "Sesli betimleme", "senkron", "Senkron", "www.", ".com",
      "izlediğiniz için", "İzlediğiniz için", "İZLEDİĞİNİZ İÇİN", "teşekkürler", "Teşekkürler", "TEŞEKKÜRLER",
      "teşekkür ederim", "Teşekkür ederim", "TEŞEKKÜR EDERİM", "thank you", "Thank you", "Thanks for watching",
      "abone ol", "Abone ol", "videoyu beğen", "bir sonraki videoda", "devam edecek", "Devam edecek",
      "transcription:", "subtitle:", "2分", "ご視聴", "I'm going to go", "Okay.", "Bye.", "Ahem.", "Ahem", "Umarım",
      "umarım"};
  static const char* const kFillers[] = {"Hıhı", "hıhı", "Pffft", "pffft", "Ehem", "ehem", "Hmm", "hmm",
                                         "Aa",   "aa",   "Ah",    "ah",    "Oh",   "oh",   "Eh",  "eh"};
  const std::string text = trim_ws(raw_text);
  if (text.size() < 2) return true;
  if (text.find_first_not_of(" \t\n\v\f\r.,?!") == std::string::npos) return true;
  if ((text.front() == '[' && text.back() == ']') || (text.front() == '(' && text.back() == ')')) return true;

  auto lower_of = [](std::string s) {
    std::transform(s.begin(), s.end(), s.begin(), ::tolower);
    return s;
  };
  auto strip_punct = [](std::string s) {
    while (!s.empty() && ispunct((unsigned char)s.back())) s.pop_back();
    size_t i = 0;
    while (i < s.size() && ispunct((unsigned char)s[i])) ++i;
    return s.substr(i);
  };
  const std::string lower = lower_of(text);
  const std::string core_lower = strip_punct(lower), core = strip_punct(text);
  for (const char* p : kPhrases) {
    const std::string phrase(p);
    if (phrase.size() > 4) {  // long phrases: substring match in either casing
      if (lower.find(phrase) != std::string::npos || text.find(phrase) != std::string::npos) return true;
    }
    if (phrase.size() <= 6) {  // short phrases: the whole (punctuation-stripped) segment
      if (core_lower == lower_of(phrase) || core == phrase) return true;
    }
  }
  for (const char* f : kFillers)
    if (core_lower == f || core == f) return true;
  return false;
}

}  // namespace utils
}  // namespace sentiric
