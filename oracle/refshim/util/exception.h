// refshim: MVE util/exception.h stand-in (see ../README.md)
#pragma once
#include <stdexcept>
#include <string>

namespace util {

class Exception : public std::exception, public std::string {
public:
    Exception() {}
    Exception(std::string const& msg) : std::string(msg) {}
    Exception(std::string const& msg, char const* msg2) : std::string(msg) { append(msg2); }
    Exception(std::string const& msg, std::string const& msg2) : std::string(msg) { append(msg2); }
    virtual ~Exception() throw() {}
    virtual const char* what() const throw() { return c_str(); }
};

class FileException : public Exception {
public:
    std::string filename;
    FileException(std::string const& fn, std::string const& msg) : Exception(msg), filename(fn) {}
    FileException(std::string const& fn, char const* msg) : Exception(msg), filename(fn) {}
    virtual ~FileException() throw() {}
};

}  // namespace util
