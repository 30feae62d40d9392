#include "awm_random.hh"
#include "awm_util.hh"

#include <inttypes.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- AES-128 (FIPS-197), encryption only */

namespace {
struct SBox
{
  uint8_t s[256];
  SBox()
  {
    /* multiplicative inverse in GF(2^8) followed by the affine map; generated, not tabulated */
    uint8_t p = 1, q = 1;
    do
      {
        p = p ^ (p << 1) ^ ((p & 0x80) ? 0x1B : 0);            // p *= 3
        q ^= q << 1; q ^= q << 2; q ^= q << 4;                 // q /= 3
        if (q & 0x80) q ^= 0x09;
        const uint8_t x = q ^ rotl (q, 1) ^ rotl (q, 2) ^ rotl (q, 3) ^ rotl (q, 4);
        s[p] = x ^ 0x63;
      }
    while (p != 1);
    s[0] = 0x63;
  }
  static uint8_t rotl (uint8_t v, int n) { return uint8_t ((v << n) | (v >> (8 - n))); }
};
const SBox g_sbox;
inline uint8_t xtime (uint8_t x) { return uint8_t ((x << 1) ^ ((x & 0x80) ? 0x1B : 0)); }
}

AES128::AES128 (const unsigned char *key)
{
  memcpy (m_rk, key, 16);
  uint8_t rcon = 1;
  for (int i = 16; i < 176; i += 4)
    {
      uint8_t t[4] = { m_rk[i - 4], m_rk[i - 3], m_rk[i - 2], m_rk[i - 1] };
      if (i % 16 == 0)
        {
          const uint8_t t0 = t[0];
          t[0] = g_sbox.s[t[1]] ^ rcon;
          t[1] = g_sbox.s[t[2]];
          t[2] = g_sbox.s[t[3]];
          t[3] = g_sbox.s[t0];
          rcon = xtime (rcon);
        }
      for (int j = 0; j < 4; j++)
        m_rk[i + j] = m_rk[i - 16 + j] ^ t[j];
    }
}

namespace {
/* T-table for SubBytes + MixColumns of one state column (generated from the S-box) */
struct TTab
{
  uint32_t t[4][256];
  TTab()
  {
    for (int x = 0; x < 256; x++)
      {
        const uint8_t s = g_sbox.s[x], s2 = xtime (s), s3 = s2 ^ s;
        const uint32_t w = uint32_t (s2) | (uint32_t (s) << 8) | (uint32_t (s) << 16) | (uint32_t (s3) << 24);   // bytes (2s, s, s, 3s), little endian
        for (int r = 0; r < 4; r++)
          t[r][x] = (w << (8 * r)) | (w >> ((32 - 8 * r) & 31));
      }
  }
};
const TTab g_ttab;
inline uint32_t load_le32 (const uint8_t *p) { uint32_t v; memcpy (&v, p, 4); return v; }
}

#if defined(__x86_64__)
#include <immintrin.h>
__attribute__ ((target ("aes,sse2"))) static void
encrypt_block_aesni (const uint8_t *rk, const uint8_t in[16], uint8_t out[16])
{
  __m128i s = _mm_xor_si128 (_mm_loadu_si128 ((const __m128i *) in), _mm_loadu_si128 ((const __m128i *) rk));
  for (int r = 1; r < 10; r++)
    s = _mm_aesenc_si128 (s, _mm_loadu_si128 ((const __m128i *) (rk + 16 * r)));
  s = _mm_aesenclast_si128 (s, _mm_loadu_si128 ((const __m128i *) (rk + 160)));
  _mm_storeu_si128 ((__m128i *) out, s);
}
static const bool g_have_aesni = __builtin_cpu_supports ("aes") && !getenv ("AWM_NO_AESNI");   // env switch lets the tests cover the table path
#else
static const bool g_have_aesni = false;
static void encrypt_block_aesni (const uint8_t *, const uint8_t *, uint8_t *) {}
#endif

void
AES128::encrypt_block (const uint8_t in[16], uint8_t out[16]) const
{
  if (g_have_aesni)
    {
      encrypt_block_aesni (m_rk, in, out);
      return;
    }
  /* column c of the state as a little-endian word: byte (r, c) in bits 8r..8r+7 */
  uint32_t c0 = load_le32 (in) ^ load_le32 (m_rk), c1 = load_le32 (in + 4) ^ load_le32 (m_rk + 4);
  uint32_t c2 = load_le32 (in + 8) ^ load_le32 (m_rk + 8), c3 = load_le32 (in + 12) ^ load_le32 (m_rk + 12);
  const uint32_t (*T)[256] = g_ttab.t;
  for (int round = 1; round < 10; round++)
    {
      const uint8_t *rk = m_rk + 16 * round;
      const uint32_t n0 = T[0][c0 & 0xff] ^ T[1][(c1 >> 8) & 0xff] ^ T[2][(c2 >> 16) & 0xff] ^ T[3][c3 >> 24] ^ load_le32 (rk);
      const uint32_t n1 = T[0][c1 & 0xff] ^ T[1][(c2 >> 8) & 0xff] ^ T[2][(c3 >> 16) & 0xff] ^ T[3][c0 >> 24] ^ load_le32 (rk + 4);
      const uint32_t n2 = T[0][c2 & 0xff] ^ T[1][(c3 >> 8) & 0xff] ^ T[2][(c0 >> 16) & 0xff] ^ T[3][c1 >> 24] ^ load_le32 (rk + 8);
      const uint32_t n3 = T[0][c3 & 0xff] ^ T[1][(c0 >> 8) & 0xff] ^ T[2][(c1 >> 16) & 0xff] ^ T[3][c2 >> 24] ^ load_le32 (rk + 12);
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    }
  const uint32_t col[4] = { c0, c1, c2, c3 };
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++)
      out[4 * c + r] = g_sbox.s[(col[(c + r) & 3] >> (8 * r)) & 0xff] ^ m_rk[160 + 4 * c + r];
}

/* ---------------------------------------------------------------- Random */

Random::Random (const Key& key, uint64_t start_seed, Stream stream) :
  m_aes (key.aes_key())
{
  seed (start_seed, stream);
}

void
Random::seed (uint64_t seed, Stream stream)
{
  uint8_t plain[16] = { 0 };
  for (int i = 0; i < 8; i++)
    plain[i] = uint8_t (seed >> (56 - 8 * i));      // big endian, endian independent
  plain[8] = uint8_t (stream);
  m_aes.encrypt_block (plain, m_ctr);
  m_pos = 32;
}

void
Random::refill()
{
  /* 256 bytes of CTR keystream; the counter is one 128-bit big-endian integer */
  for (int blk = 0; blk < 16; blk++)
    {
      uint8_t ks[16];
      m_aes.encrypt_block (m_ctr, ks);
      for (int i = 15; i >= 0; i--)
        if (++m_ctr[i] != 0)
          break;
      for (int w = 0; w < 2; w++)
        {
          uint64_t v = 0;
          for (int b = 0; b < 8; b++)
            v = (v << 8) | ks[8 * w + b];
          m_buffer[2 * blk + w] = v;
        }
    }
  m_pos = 0;
}

double
Random::random_double()
{
  /* generate_canonical<double,53> with a 64-bit source: one draw, (double) r / 2^64, kept below 1 */
  const double r = double ((*this)()) / 18446744073709551616.0;
  return r >= 1.0 ? nextafter (1.0, 0.0) : r;
}

std::string
Random::gen_key()
{
  std::vector<unsigned char> key (16);
  FILE *f = fopen ("/dev/urandom", "rb");
  if (!f || fread (key.data(), 1, 16, f) != 16)
    {
      error ("audiowmark: unable to read random key material from /dev/urandom\n");
      exit (1);
    }
  fclose (f);
  return vec_to_hex_str (key);
}

/* ---------------------------------------------------------------- Key */

void
Key::set_test_key (uint64_t key)
{
  for (int i = 0; i < 8; i++)
    m_aes_key[i] = uint8_t (key >> (56 - 8 * i));
  for (int i = 8; i < 16; i++)
    m_aes_key[i] = 0;
  m_name = string_printf ("test-key-%" PRId64, int64_t (key));
}

void
Key::set_key (const unsigned char *bytes16, const std::string& name)
{
  m_aes_key.assign (bytes16, bytes16 + SIZE);
  m_name = name;
}

/* key file grammar (src/random.cc:209-360): lines of tokens; `key <32 hex>` exactly once, optional
 * `name <string>`; tokens are bare words [A-Za-z0-9.:=/_-]+ or "quoted strings" with \ escapes; # starts a comment */
static bool
split_tokens (const std::string& line_in, std::vector<std::string>& tokens)
{
  auto word_char = [] (char c) {
    return (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || (c >= '0' && c <= '9') || strchr (".:=/-_", c) != nullptr;
  };
  auto space = [] (char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r'; };
  const std::string line = line_in + '\n';
  tokens.clear();
  size_t i = 0;
  while (i < line.size())
    {
      const char c = line[i];
      if (space (c))
        i++;
      else if (c == '#')
        return true;
      else if (c == '"')
        {
          std::string s;
          for (i++; ; i++)
            {
              if (i >= line.size())
                return false;                 // unterminated string
              if (line[i] == '\\')
                {
                  if (++i >= line.size())
                    return false;
                  s += line[i];
                }
              else if (line[i] == '"')
                break;
              else
                s += line[i];
            }
          i++;
          tokens.push_back (s);
        }
      else if (word_char (c))
        {
          std::string s;
          while (i < line.size() && word_char (line[i]))
            s += line[i++];
          if (i < line.size() && !space (line[i]))
            return false;                     // a word must be followed by white space
          tokens.push_back (s);
        }
      else
        return false;
    }
  return true;
}

void
Key::load_key (const std::string& key_file)
{
  FILE *f = fopen (key_file.c_str(), "r");
  if (!f)
    {
      error ("audiowmark: error opening key file: '%s'\n", key_file.c_str());
      exit (1);
    }
  m_name = key_file;
  const size_t sep = m_name.find_last_of ("\\/");
  if (sep != std::string::npos)
    m_name = m_name.substr (sep + 1);

  char buffer[1024];
  int line = 1, keys = 0;
  while (fgets (buffer, sizeof (buffer), f))
    {
      std::vector<std::string> tokens;
      bool ok = false;
      if (split_tokens (buffer, tokens))
        {
          if (tokens.size() == 2 && tokens[0] == "key")
            {
              std::vector<unsigned char> key = hex_str_to_vec (tokens[1]);
              if (key.size() != SIZE)
                {
                  error ("audiowmark: wrong key length in key file '%s', line %d\n => required key length is %zd bits\n", key_file.c_str(), line, SIZE * 8);
                  exit (1);
                }
              m_aes_key = key;
              keys++;
              ok = true;
            }
          if (tokens.size() == 2 && tokens[0] == "name")
            {
              m_name = tokens[1];
              ok = true;
            }
          if (tokens.empty())
            ok = true;
        }
      if (!ok)
        {
          error ("audiowmark: parse error in key file '%s', line %d\n", key_file.c_str(), line);
          exit (1);
        }
      line++;
    }
  fclose (f);
  if (keys > 1)
    {
      error ("audiowmark: key file '%s' contains more than one key\n", key_file.c_str());
      exit (1);
    }
  if (keys == 0)
    {
      error ("audiowmark: key file '%s' contains no key\n", key_file.c_str());
      exit (1);
    }
}
