// ref_shim/Error.h -- TEST INFRASTRUCTURE ONLY.  Stand-in for PSRCHIVE's Error.h: same constructor forms
// (code, function, printf-style message), operator += for context, stream insertion.
#ifndef REF_SHIM_ERROR_H
#define REF_SHIM_ERROR_H
#include <cstdarg>
#include <cstdio>
#include <sstream>
#include <string>
enum ErrorCode { Undefined, BadAllocation, BadPointer, InvalidParam, InvalidState, InvalidRange, FileNotFound,
                 FailedCall, FailedSys, EndOfFile };
class Error {
 public:
  Error(ErrorCode c, const std::string& func) : code(c), function(func) {}
  Error(ErrorCode c, const std::string& func, const char* fmt, ...) : code(c), function(func) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    message = buf;
  }
  Error(ErrorCode c, const std::string& func, const std::string& msg) : code(c), function(func), message(msg) {}
  const Error& operator+=(const char* ctx) { function += std::string(" <- ") + ctx; return *this; }
  const Error& operator+=(const std::string& ctx) { function += " <- " + ctx; return *this; }
  template <class T> Error& operator<<(const T& t) { std::ostringstream s; s << t; message += s.str(); return *this; }
  const std::string& get_message() const { return message; }
  ErrorCode get_code() const { return code; }
  ErrorCode code;
  std::string function, message;
};
#endif
