/* standins.cc -- bodies for the reference classes whose real implementation
 * needs an absent third-party library (libsndfile, mpg123, ffmpeg).
 *
 * TEST INFRASTRUCTURE ONLY; linked into oracle/_ref/audiowmark together with
 * the unmodified reference sources (see oracle/Makefile.ref).  The class
 * declarations come from the reference headers; only behaviour needed to run
 * `add`, `get`, `cmp` and the `test-*` helpers on WAV files is provided:
 *
 *  - SFInputStream  : RIFF/WAVE reader delegating sample conversion to the
 *                     reference's own WavPipeInputStream (int samples are
 *                     left-justified to 32 bit and scaled by 2^-31, the same
 *                     normalisation src/sfinputstream.cc:189-210 applies).
 *  - SFOutputStream : canonical 44-byte-header WAV writer; integer samples are
 *                     float_to_int_clip<32>() then the top <bit_depth> bits
 *                     are kept (what sf_writef_int does, src/sfoutputstream.cc:148-155).
 *  - MP3InputStream, hls_add, hls_prepare : report "not available".
 */
#include "sfinputstream.hh"
#include "sfoutputstream.hh"
#include "mp3inputstream.hh"
#include "wavpipeinputstream.hh"
#include "rawconverter.hh"
#include "wavdata.hh"
#include "random.hh"
#include "hls.hh"

#include <stdio.h>
#include <string.h>
#include <errno.h>

using std::string;
using std::vector;

/* ---------------- SFInputStream ---------------- */

SFVirtualData::SFVirtualData()
{
  memset (&io, 0, sizeof (io));
}

namespace {
struct InImpl
{
  WavPipeInputStream wav;
};
InImpl *in_impl (SNDFILE *s) { return reinterpret_cast<InImpl *> (s); }

/* scan RIFF chunks for the size of the data chunk; returns false if unknown */
bool
wav_data_bytes (const string& filename, uint64_t& data_bytes)
{
  FILE *f = fopen (filename.c_str(), "rb");
  if (!f)
    return false;
  unsigned char hdr[12];
  bool ok = false;
  if (fread (hdr, 12, 1, f) == 1 && !memcmp (hdr, "RIFF", 4) && !memcmp (hdr + 8, "WAVE", 4))
    {
      unsigned char ch[8];
      while (fread (ch, 8, 1, f) == 1)
        {
          uint32_t sz = ch[4] | (ch[5] << 8) | (ch[6] << 16) | (uint32_t (ch[7]) << 24);
          if (!memcmp (ch, "data", 4))
            {
              long pos = ftell (f);
              fseek (f, 0, SEEK_END);
              long end = ftell (f);
              uint64_t avail = end - pos;
              data_bytes = (sz == 0xffffffffu || sz > avail) ? avail : sz;
              ok = true;
              break;
            }
          if (fseek (f, sz + (sz & 1), SEEK_CUR))
            break;
        }
    }
  fclose (f);
  return ok;
}
}

SFInputStream::~SFInputStream()
{
  close();
}

Error
SFInputStream::open (const string& filename)
{
  InImpl *impl = new InImpl();
  Error err = impl->wav.open (filename);
  if (err)
    {
      delete impl;
      return err;
    }
  m_sndfile     = reinterpret_cast<SNDFILE *> (impl);
  m_n_channels  = impl->wav.n_channels();
  m_sample_rate = impl->wav.sample_rate();
  m_bit_depth   = impl->wav.bit_depth();
  m_encoding    = impl->wav.encoding();
  if (m_encoding == Encoding::UNSIGNED) /* 8 bit wav */
    m_encoding = Encoding::SIGNED;
  m_is_stdin    = (filename == "-");
  m_n_frames    = AudioInputStream::N_FRAMES_UNKNOWN;

  uint64_t data_bytes = 0;
  if (!m_is_stdin && wav_data_bytes (filename, data_bytes))
    m_n_frames = data_bytes / (m_n_channels * ((m_bit_depth + 7) / 8));

  m_state = State::OPEN;
  return Error::Code::NONE;
}

Error
SFInputStream::open (const vector<unsigned char> *)
{
  return Error ("reference stand-in: in-memory sndfile input not available");
}

Error
SFInputStream::read_frames (vector<float>& samples, size_t count)
{
  return in_impl (m_sndfile)->wav.read_frames (samples, count);
}

void
SFInputStream::close()
{
  if (m_state == State::OPEN)
    {
      delete in_impl (m_sndfile);
      m_sndfile = nullptr;
      m_state = State::CLOSED;
    }
}

int      SFInputStream::sample_rate() const { return m_sample_rate; }
int      SFInputStream::bit_depth() const   { return m_bit_depth; }
Encoding SFInputStream::encoding() const    { return m_encoding; }

/* ---------------- SFOutputStream ---------------- */

namespace {
struct OutImpl
{
  FILE    *file = nullptr;
  uint64_t data_bytes = 0;
};
OutImpl *out_impl (SNDFILE *s) { return reinterpret_cast<OutImpl *> (s); }

void
put_u32 (unsigned char *p, uint32_t u)
{
  p[0] = u; p[1] = u >> 8; p[2] = u >> 16; p[3] = u >> 24;
}
void
put_u16 (unsigned char *p, uint16_t u)
{
  p[0] = u; p[1] = u >> 8;
}
void
write_header (OutImpl *impl, int n_channels, int sample_rate, int bit_depth, bool is_float)
{
  unsigned char h[44];
  memcpy (h, "RIFF", 4);
  put_u32 (h + 4, uint32_t (36 + impl->data_bytes + (impl->data_bytes & 1)));
  memcpy (h + 8, "WAVEfmt ", 8);
  put_u32 (h + 16, 16);
  put_u16 (h + 20, is_float ? 3 : 1);
  put_u16 (h + 22, n_channels);
  put_u32 (h + 24, sample_rate);
  put_u32 (h + 28, sample_rate * n_channels * bit_depth / 8);
  put_u16 (h + 32, n_channels * bit_depth / 8);
  put_u16 (h + 34, bit_depth);
  memcpy (h + 36, "data", 4);
  put_u32 (h + 40, uint32_t (impl->data_bytes));
  fseek (impl->file, 0, SEEK_SET);
  fwrite (h, 1, 44, impl->file);
}
}

SFOutputStream::~SFOutputStream()
{
  close();
}

Error
SFOutputStream::open (const string& filename, int n_channels, int sample_rate, int bit_depth, Encoding encoding, OutFormat)
{
  m_write_float_data = (encoding == Encoding::FLOAT);
  if (m_write_float_data)
    {
      if (bit_depth != 32 && bit_depth != 64)
        return Error ("reference stand-in: unsupported float bit depth");
    }
  else if (bit_depth != 16 && bit_depth != 24 && bit_depth != 32)
    return Error ("reference stand-in: unsupported bit depth");

  OutImpl *impl = new OutImpl();
  impl->file = fopen (filename.c_str(), "wb");
  if (!impl->file)
    {
      delete impl;
      return Error (strerror (errno));
    }
  m_sndfile     = reinterpret_cast<SNDFILE *> (impl);
  m_n_channels  = n_channels;
  m_sample_rate = sample_rate;
  m_bit_depth   = bit_depth;
  write_header (impl, n_channels, sample_rate, bit_depth, m_write_float_data);
  m_state = State::OPEN;
  return Error::Code::NONE;
}

Error
SFOutputStream::open (vector<unsigned char> *, int, int, int, Encoding, OutFormat)
{
  return Error ("reference stand-in: in-memory sndfile output not available");
}

Error
SFOutputStream::write_frames (const vector<float>& samples)
{
  OutImpl *impl = out_impl (m_sndfile);
  const int width = m_bit_depth / 8;
  vector<unsigned char> bytes (samples.size() * width);
  unsigned char *p = bytes.data();
  if (m_write_float_data)
    {
      for (float s : samples)
        {
          if (m_bit_depth == 32)
            {
              float f = float_clip (s);
              memcpy (p, &f, 4);
            }
          else
            {
              double d = float_clip (s);
              memcpy (p, &d, 8);
            }
          p += width;
        }
    }
  else
    {
      for (float s : samples)
        {
          const int v = float_to_int_clip<32> (s) >> (32 - m_bit_depth); /* keep most significant bits */
          for (int b = 0; b < width; b++)
            p[b] = (unsigned (v) >> (8 * b)) & 0xff;
          p += width;
        }
    }
  if (fwrite (bytes.data(), 1, bytes.size(), impl->file) != bytes.size())
    return Error ("writing sample data failed: short write");
  impl->data_bytes += bytes.size();
  return Error::Code::NONE;
}

Error
SFOutputStream::close()
{
  if (m_state == State::OPEN)
    {
      OutImpl *impl = out_impl (m_sndfile);
      if (impl->data_bytes & 1)
        fputc (0, impl->file);
      write_header (impl, m_n_channels, m_sample_rate, m_bit_depth, m_write_float_data);
      const bool err = fclose (impl->file) != 0;
      delete impl;
      m_sndfile = nullptr;
      m_state = State::CLOSED;
      if (err)
        return Error ("sf_close returned an error");
    }
  return Error::Code::NONE;
}

int SFOutputStream::bit_depth() const   { return m_bit_depth; }
int SFOutputStream::sample_rate() const { return m_sample_rate; }
int SFOutputStream::n_channels() const  { return m_n_channels; }

/* ---------------- MP3InputStream ---------------- */

MP3InputStream::~MP3InputStream() {}
Error    MP3InputStream::open (const string&) { return Error ("reference stand-in: mp3 input not available"); }
Error    MP3InputStream::read_frames (vector<float>&, size_t) { return Error ("reference stand-in: mp3 input not available"); }
void     MP3InputStream::close() {}
int      MP3InputStream::bit_depth() const   { return 24; }
int      MP3InputStream::sample_rate() const { return m_sample_rate; }
int      MP3InputStream::n_channels() const  { return m_n_channels; }
size_t   MP3InputStream::n_frames() const    { return N_FRAMES_UNKNOWN; }
Encoding MP3InputStream::encoding() const    { return Encoding::SIGNED; }
bool     MP3InputStream::detect (const string&) { return false; }

/* ---------------- HLS ---------------- */

int
hls_add (const Key&, const string&, const string&, const string&)
{
  error ("audiowmark: reference stand-in: hls-add not available\n");
  return 1;
}

int
hls_prepare (const string&, const string&, const string&, const string&)
{
  error ("audiowmark: reference stand-in: hls-prepare not available\n");
  return 1;
}
