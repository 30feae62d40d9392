// awm_util.hh -- logging, Error value type, hex / bit-string helpers of the host side.
// Mirrors the names of the reference's src/utils.hh:76-130 so host code reads like the original
// (Error is truthy on failure, .message() gives the text; error/warning/info/debug log to stderr).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

#if defined(__GNUC__)
#define AWM_PRINTF(f, a) __attribute__ ((__format__ (__printf__, f, a)))
#else
#define AWM_PRINTF(f, a)
#endif

enum class Log { DEBUG = 0, INFO = 1, WARNING = 2, ERROR = 3 };
void set_log_level (Log level);
void error (const char *format, ...) AWM_PRINTF (1, 2);
void warning (const char *format, ...) AWM_PRINTF (1, 2);
void info (const char *format, ...) AWM_PRINTF (1, 2);
void debug (const char *format, ...) AWM_PRINTF (1, 2);
std::string string_printf (const char *format, ...) AWM_PRINTF (1, 2);

class Error
{
public:
  enum class Code { NONE, STR };
  Error (Code code = Code::NONE) : m_code (code), m_message (code == Code::NONE ? "OK" : "Unknown error") {}
  explicit Error (const std::string& message) : m_code (Code::STR), m_message (message) {}
  Code        code() const    { return m_code; }
  const char *message() const { return m_message.c_str(); }
  operator bool() const       { return m_code != Code::NONE; }
private:
  Code        m_code;
  std::string m_message;
};

std::vector<int>           bit_str_to_vec (const std::string& bits);          // src/utils.cc:95-111
std::string                bit_vec_to_str (const std::vector<int>& bit_vec);  // src/utils.cc:113-133
std::vector<unsigned char> hex_str_to_vec (const std::string& str);
std::string                vec_to_hex_str (const std::vector<unsigned char>& vec);
double                     get_time();
