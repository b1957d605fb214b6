// Thread-local error string + log sink behind the C ABI (include/sw_whisper.h).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <mutex>

#include "../../include/sw_whisper.h"

namespace sw {
static thread_local char g_err[1024] = "";
static sw_log_callback g_log_cb = nullptr;
static void* g_log_user = nullptr;
static std::mutex g_log_mu;

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const char* last_error_string() { return g_err; }

void log_msg(int level, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  std::lock_guard<std::mutex> lk(g_log_mu);
  if (g_log_cb)
    g_log_cb(level, buf, g_log_user);
  else if (level >= 3)
    fprintf(stderr, "[sw_whisper] %s\n", buf);
}
}  // namespace sw

extern "C" {
const char* sw_last_error(void) { return sw::g_err; }
const char* sw_version(void) { return "sw_whisper 0.1 (sm_100a)"; }
void sw_log_set(sw_log_callback cb, void* user) {
  std::lock_guard<std::mutex> lk(sw::g_log_mu);
  sw::g_log_cb = cb;
  sw::g_log_user = user;
}
}
