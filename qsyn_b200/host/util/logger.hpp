// logger.hpp -- the subset of spdlog the tensor-verification path uses, with qsyn's level
// prefixes ("[info]     ", pattern "%L%v", CMakeLists.txt:152-156, qsyn_helper.cpp:128).
// Lets the host layer print byte-identical log lines without the spdlog dependency.  When the
// headers are dropped into the real qsyn tree, include <spdlog/spdlog.h> instead.
#pragma once

#include <cstdio>
#include <format>
#include <string>
#include <string_view>

namespace spdlog {

namespace level {
enum level_enum : int { trace = 0, debug = 1, info = 2, warn = 3, err = 4, critical = 5, off = 6 };
}

namespace detail {
inline level::level_enum& current_level() {
    static level::level_enum lvl = level::warn;  // qsyn default (qsyn_helper.cpp:129)
    return lvl;
}
inline void emit(level::level_enum lvl, std::string const& msg) {
    static char const* const prefix[] = {"[trace]    ", "[debug]    ", "[info]     ", "[warn]     ", "[error]    ", "[critical] "};
    if (lvl < current_level()) return;
    std::fputs(prefix[lvl], stdout);
    std::fputs(msg.c_str(), stdout);
    std::fputc('\n', stdout);
    std::fflush(stdout);
}
}  // namespace detail

inline void set_level(level::level_enum lvl) { detail::current_level() = lvl; }
inline level::level_enum get_level() { return detail::current_level(); }

template <class... Args>
void trace(std::format_string<Args...> f, Args&&... a) {
    if (level::trace < detail::current_level()) return;  // (like spdlog: a disabled level formats nothing)
    detail::emit(level::trace, std::format(f, std::forward<Args>(a)...));
}
template <class... Args>
void debug(std::format_string<Args...> f, Args&&... a) {
    if (level::debug < detail::current_level()) return;  // (like spdlog: a disabled level formats nothing)
    detail::emit(level::debug, std::format(f, std::forward<Args>(a)...));
}
template <class... Args>
void info(std::format_string<Args...> f, Args&&... a) {
    if (level::info < detail::current_level()) return;  // (like spdlog: a disabled level formats nothing)
    detail::emit(level::info, std::format(f, std::forward<Args>(a)...));
}
template <class... Args>
void warn(std::format_string<Args...> f, Args&&... a) {
    if (level::warn < detail::current_level()) return;  // (like spdlog: a disabled level formats nothing)
    detail::emit(level::warn, std::format(f, std::forward<Args>(a)...));
}
template <class... Args>
void error(std::format_string<Args...> f, Args&&... a) {
    if (level::err < detail::current_level()) return;  // (like spdlog: a disabled level formats nothing)
    detail::emit(level::err, std::format(f, std::forward<Args>(a)...));
}
template <class... Args>
void critical(std::format_string<Args...> f, Args&&... a) {
    if (level::critical < detail::current_level()) return;  // (like spdlog: a disabled level formats nothing)
    detail::emit(level::critical, std::format(f, std::forward<Args>(a)...));
}

}  // namespace spdlog
