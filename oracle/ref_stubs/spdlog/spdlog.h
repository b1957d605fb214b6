// Stand-in for the spdlog header the reference's utils.h includes, so that oracle/Makefile can compile
// /root/reference/src/utils.h where it lies (logging calls become no-ops). Test infrastructure only.
#pragma once
namespace spdlog {
template <typename... A> inline void info(A&&...) {}
template <typename... A> inline void warn(A&&...) {}
template <typename... A> inline void error(A&&...) {}
template <typename... A> inline void debug(A&&...) {}
}  // namespace spdlog
