// awm_random.hh -- Key + keyed PRNG of the host side.
// Same generator as the reference's Random (src/random.cc:97-161, src/random.hh:53-113):
// AES-128; the counter block is ECB(key, be64(seed) || stream || 0^7) and the output is the
// CTR keystream read as big-endian u64 words.  AES is implemented here (FIPS-197), the
// reference links libgcrypt for it.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

class Key
{
  std::vector<unsigned char> m_aes_key;
  std::string                m_name;
public:
  static constexpr size_t SIZE = 16;
  Key() : m_aes_key (SIZE) {}
  bool operator== (const Key& o) const { return m_aes_key == o.m_aes_key && m_name == o.m_name; }
  void set_test_key (uint64_t key);                 // src/random.cc:202-207
  void set_key (const unsigned char *bytes16, const std::string& name);
  void load_key (const std::string& filename);      // src/random.cc:295-360 (exits on error like the reference)
  const unsigned char *aes_key() const { return m_aes_key.data(); }
  const std::string&   name() const    { return m_name; }
};

class AES128
{
  uint8_t m_rk[176];
public:
  explicit AES128 (const unsigned char *key16);
  void encrypt_block (const uint8_t in[16], uint8_t out[16]) const;
};

class Random
{
public:
  enum class Stream { data_up_down = 1, sync_up_down = 2, speed_clip = 3, mix = 4, bit_order = 5, frame_position = 6 };
  typedef uint64_t result_type;
private:
  AES128   m_aes;
  uint8_t  m_ctr[16];
  uint64_t m_buffer[32];
  size_t   m_pos = 32;
  void     refill();
public:
  Random (const Key& key, uint64_t seed, Stream stream);
  void seed (uint64_t seed, Stream stream);
  result_type operator()() { if (m_pos == 32) refill(); return m_buffer[m_pos++]; }
  static constexpr result_type min() { return 0; }
  static constexpr result_type max() { return UINT64_MAX; }
  double random_double();          // [0,1): what libstdc++'s uniform_real_distribution<double> yields for one 64-bit draw

  template<class T> void
  shuffle (std::vector<T>& v)      // Fisher-Yates, j = i + r % (n - i)   (src/random.hh:102-113)
  {
    for (size_t i = 0; i < v.size(); i++)
      {
        const uint64_t r = (*this)();
        const size_t j = i + r % (v.size() - i);
        std::swap (v[i], v[j]);
      }
  }
  static std::string gen_key();    // 16 random bytes as hex (reads /dev/urandom)
};
