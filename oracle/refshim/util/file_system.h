// refshim: MVE util/file_system.h stand-in -- path helpers only (see ../README.md)
#pragma once
#include <string>
#include <sys/stat.h>

namespace util { namespace fs {

inline bool exists(char const* p) { struct stat s; return ::stat(p, &s) == 0; }
inline bool dir_exists(char const* p) { struct stat s; return ::stat(p, &s) == 0 && S_ISDIR(s.st_mode); }
inline bool file_exists(char const* p) { struct stat s; return ::stat(p, &s) == 0 && S_ISREG(s.st_mode); }
inline std::string join_path(std::string const& a, std::string const& b) { return a.empty() ? b : (a + "/" + b); }
inline std::string dirname(std::string const& p) { std::size_t k = p.find_last_of('/'); return k == std::string::npos ? "." : p.substr(0, k); }
inline std::string basename(std::string const& p) { std::size_t k = p.find_last_of('/'); return k == std::string::npos ? p : p.substr(k + 1); }
inline std::string replace_extension(std::string const& p, std::string const& e) { std::size_t k = p.find_last_of('.'); return (k == std::string::npos ? p : p.substr(0, k)) + "." + e; }
inline std::string sanitize_path(std::string const& p) { return p; }
inline std::string abspath(std::string const& p) { return p; }
inline bool mkdir(char const* p) { return ::mkdir(p, 0755) == 0; }

} }  // namespace util::fs
