// Stand-in for spdlog (absent from this image): logging becomes a no-op, fmt::format/join return empty strings.
// Written for oracle/ref_host (TEST INFRASTRUCTURE): lets the reference's own headers be compiled where they lie.
#pragma once
#include <string>
#define SPDLOG_LEVEL_TRACE 0
#define SPDLOG_LEVEL_DEBUG 1
#define SPDLOG_LEVEL_INFO 2
#define SPDLOG_LEVEL_WARN 3
#define SPDLOG_LEVEL_ERROR 4
#define SPDLOG_LEVEL_CRITICAL 5
#define SPDLOG_LEVEL_OFF 6
#ifndef SPDLOG_ACTIVE_LEVEL
#define SPDLOG_ACTIVE_LEVEL SPDLOG_LEVEL_OFF
#endif
namespace spdlog {
namespace level { enum level_enum { trace, debug, info, warn, err, critical, off }; }
inline void set_pattern(const char*) {}
inline void set_level(level::level_enum) {}
template <class... A> inline void info(A&&...) {}
template <class... A> inline void warn(A&&...) {}
template <class... A> inline void error(A&&...) {}
template <class... A> inline void critical(A&&...) {}
template <class... A> inline void debug(A&&...) {}
template <class... A> inline void trace(A&&...) {}
}  // namespace spdlog
namespace fmt {
template <class... A> inline std::string format(A&&...) { return std::string(); }
template <class... A> inline std::string join(A&&...) { return std::string(); }
}  // namespace fmt
namespace m3stub { template <class... A> inline void sink(A&&...) {} }
#define SPDLOG_TRACE(...) ::m3stub::sink(__VA_ARGS__)
#define SPDLOG_DEBUG(...) ::m3stub::sink(__VA_ARGS__)
#define SPDLOG_INFO(...) ::m3stub::sink(__VA_ARGS__)
#define SPDLOG_WARN(...) ::m3stub::sink(__VA_ARGS__)
#define SPDLOG_ERROR(...) ::m3stub::sink(__VA_ARGS__)
#define SPDLOG_CRITICAL(...) ::m3stub::sink(__VA_ARGS__)
