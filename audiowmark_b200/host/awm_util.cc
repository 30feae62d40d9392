#include "awm_util.hh"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <sys/time.h>

static Log g_log_level = Log::INFO;

void
set_log_level (Log level)
{
  g_log_level = level;
}

static std::string
vformat (const char *format, va_list ap)
{
  va_list ap2;
  va_copy (ap2, ap);
  const int n = vsnprintf (nullptr, 0, format, ap2);
  va_end (ap2);
  if (n < 0)
    return format;
  std::string s (size_t (n) + 1, '\0');
  vsnprintf (&s[0], s.size(), format, ap);
  s.resize (n);
  return s;
}

static void
log_msg (Log level, const char *format, va_list ap)
{
  if (level < g_log_level)
    return;
  const std::string s = vformat (format, ap);
  fputs (s.c_str(), stderr);
  fflush (stderr);
}

#define AWM_LOG_FN(name, level) \
  void name (const char *format, ...) { va_list ap; va_start (ap, format); log_msg (level, format, ap); va_end (ap); }
AWM_LOG_FN (error, Log::ERROR)
AWM_LOG_FN (warning, Log::WARNING)
AWM_LOG_FN (info, Log::INFO)
AWM_LOG_FN (debug, Log::DEBUG)

std::string
string_printf (const char *format, ...)
{
  va_list ap;
  va_start (ap, format);
  std::string s = vformat (format, ap);
  va_end (ap);
  return s;
}

static int
hex_nibble (char c)
{
  if (c >= '0' && c <= '9') return c - '0';
  if (c >= 'a' && c <= 'f') return c - 'a' + 10;
  if (c >= 'A' && c <= 'F') return c - 'A' + 10;
  return -1;
}

std::vector<int>
bit_str_to_vec (const std::string& bits)
{
  std::vector<int> v;
  v.reserve (bits.size() * 4);
  for (char ch : bits)
    {
      const int n = hex_nibble (ch);
      if (n < 0)
        return {};
      for (int b = 3; b >= 0; b--)      // most significant bit of each nibble first
        v.push_back ((n >> b) & 1);
    }
  return v;
}

std::string
bit_vec_to_str (const std::vector<int>& bit_vec)
{
  std::string s;
  for (size_t pos = 0; pos + 3 < bit_vec.size(); pos += 4)   // whole nibbles only
    {
      int n = 0;
      for (int j = 0; j < 4; j++)
        n = (n << 1) | (bit_vec[pos + j] ? 1 : 0);
      s += "0123456789abcdef"[n];
    }
  return s;
}

std::vector<unsigned char>
hex_str_to_vec (const std::string& str)
{
  if (str.size() % 2)
    return {};
  std::vector<unsigned char> out;
  for (size_t i = 0; i < str.size(); i += 2)
    {
      const int h = hex_nibble (str[i]), l = hex_nibble (str[i + 1]);
      if (h < 0 || l < 0)
        return {};
      out.push_back ((h << 4) | l);
    }
  return out;
}

std::string
vec_to_hex_str (const std::vector<unsigned char>& vec)
{
  std::string s;
  for (unsigned char b : vec)
    s += string_printf ("%02x", b);
  return s;
}

double
get_time()
{
  timeval tv;
  gettimeofday (&tv, nullptr);
  return tv.tv_sec + tv.tv_usec / 1e6;
}
