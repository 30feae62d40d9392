#include "awm_code.hh"
#include "awm_params.hh"

#include <assert.h>

// rate 1/12 mother code, constraint order 15 (src/convcode.cc:42-49); A = even, B = odd generators
static const unsigned k_generators[12] = { 066561, 075211, 071545, 054435, 063635, 052475,
                                           063543, 075307, 052547, 045627, 067657, 051757 };
static const unsigned k_order = 15;

static std::vector<unsigned>
generators_for (ConvBlockType t)
{
  std::vector<unsigned> g;
  for (unsigned i = 0; i < 12; i++)
    if (t == ConvBlockType::ab || (i % 2) == (t == ConvBlockType::b ? 1u : 0u))
      g.push_back (k_generators[i]);
  return g;
}

size_t
conv_code_size (ConvBlockType block_type, size_t msg_size)
{
  const size_t rate = block_type == ConvBlockType::ab ? 12 : 6;
  return (msg_size + k_order) * rate;
}

std::vector<int>
conv_encode (ConvBlockType block_type, const std::vector<int>& in_bits)
{
  const std::vector<unsigned> gens = generators_for (block_type);
  std::vector<int> out;
  out.reserve ((in_bits.size() + k_order) * gens.size());
  unsigned reg = 0;
  for (size_t i = 0; i < in_bits.size() + k_order; i++)      // k_order zero bits flush the register
    {
      reg = (reg << 1) | (i < in_bits.size() ? (in_bits[i] & 1) : 0);
      for (unsigned poly : gens)
        out.push_back (__builtin_parity (reg & poly));
    }
  return out;
}

/* ---- short payload mode (src/shortcode.cc): a binary linear block code [n, k] in front of the convolutional code.
 * The generator matrices are the best known linear codes [65,20,20], [61,16,21], [56,12,22] over GF(2) the reference
 * uses (interoperability constants; src/shortcode.cc:26-83 cites codetables.de); row i is stored as a bit mask,
 * bit j = column j. */
struct ShortRow { uint64_t lo, hi; };
static const ShortRow short_56_12[12] = {
  { 0x00feb8b646cb1001ull, 0x0ull },
  { 0x0005d0daf7f1b002ull, 0x0ull },
  { 0x0068aec1274e8804ull, 0x0ull },
  { 0x0073c692698c2808ull, 0x0ull },
  { 0x00da51f4b6048810ull, 0x0ull },
  { 0x0057617a230f1020ull, 0x0ull },
  { 0x00b9eda54a308040ull, 0x0ull },
  { 0x003f9dfcd0163080ull, 0x0ull },
  { 0x00d4b8e8ef2d2900ull, 0x0ull },
  { 0x006b339794612200ull, 0x0ull },
  { 0x008acc5794991c00ull, 0x0ull },
  { 0x009ff7fc1fffc000ull, 0x0ull },
};
static const ShortRow short_61_16[16] = {
  { 0x0498284fd74f0001ull, 0x0ull },
  { 0x0930509fae9e0002ull, 0x0ull },
  { 0x1260a13f5d3c0004ull, 0x0ull },
  { 0x139f97d14b610008ull, 0x0ull },
  { 0x1061fa0d67db0010ull, 0x0ull },
  { 0x179d21b53eaf0020ull, 0x0ull },
  { 0x186496c58c470040ull, 0x0ull },
  { 0x0797f824e9970080ull, 0x0ull },
  { 0x0f2ff049d32e0100ull, 0x0ull },
  { 0x1e5fe093a65c0200ull, 0x0ull },
  { 0x0be11488bda10400ull, 0x0ull },
  { 0x17c229117b420800ull, 0x0ull },
  { 0x18da878d079d1000ull, 0x0ull },
  { 0x06ebdab5fe232000ull, 0x0ull },
  { 0x0dd7b56bfc464000ull, 0x0ull },
  { 0x1baf6ad7f88c8000ull, 0x0ull },
};
static const ShortRow short_65_20[20] = {
  { 0xdcfaff02fec40001ull, 0x1ull },
  { 0xfb826f058a840002ull, 0x1ull },
  { 0xb5734f0b62040004ull, 0x1ull },
  { 0x28910f16b3040008ull, 0x1ull },
  { 0x13558f2d11040010ull, 0x0ull },
  { 0xab9a385b11e00020ull, 0x1ull },
  { 0x448828b599e00040ull, 0x1ull },
  { 0x9aac096889e00080ull, 0x0ull },
  { 0xe9a2fdd3ed040100ull, 0x0ull },
  { 0x5e74dda6e9e00200ull, 0x0ull },
  { 0x6013544f2d040400ull, 0x1ull },
  { 0x8251299e2d440800ull, 0x0ull },
  { 0x8993753d69601000ull, 0x0ull },
  { 0xcfdc15782c442000ull, 0x0ull },
  { 0x12891cf16b204000ull, 0x0ull },
  { 0xf9e8d6e028848000ull, 0x1ull },
  { 0xb1a62cc026450000ull, 0x1ull },
  { 0x213bc8803b860000ull, 0x1ull },
  { 0x8d31360133a80000ull, 0x1ull },
  { 0x09de2401dd300000ull, 0x1ull },
};

static const ShortRow *g_short_rows = nullptr;
static size_t g_short_in = 0, g_short_out = 0;

size_t
short_code_init (size_t k)
{
  if (k == 12)      { g_short_rows = short_56_12; g_short_out = 56; }
  else if (k == 16) { g_short_rows = short_61_16; g_short_out = 61; }
  else if (k == 20) { g_short_rows = short_65_20; g_short_out = 65; }
  else
    return 0;        /* unsupported k */
  g_short_in = k;
  return g_short_out;
}

/* number of message bits the convolutional code carries: the payload, or the block code word in short mode */
size_t
code_message_bits()
{
  return Params::payload_short ? g_short_out : Params::payload_size;
}

static ShortRow
short_codeword (uint32_t msg)
{
  ShortRow w { 0, 0 };
  for (size_t bit = 0; bit < g_short_in; bit++)
    if (msg & (1u << bit))
      {
        w.lo ^= g_short_rows[bit].lo;
        w.hi ^= g_short_rows[bit].hi;
      }
  return w;
}

std::vector<int>
short_encode_blk (const std::vector<int>& in_bits)        /* src/shortcode.cc:136-157 */
{
  assert (in_bits.size() == g_short_in);
  uint32_t msg = 0;
  for (size_t bit = 0; bit < g_short_in; bit++)
    if (in_bits[bit])
      msg |= 1u << bit;
  const ShortRow w = short_codeword (msg);
  std::vector<int> out_bits (g_short_out);
  for (size_t j = 0; j < g_short_out; j++)
    out_bits[j] = j < 64 ? (w.lo >> j) & 1 : (w.hi >> (j - 64)) & 1;
  return out_bits;
}

/* exhaustive search for the message whose code word equals the received bits; empty if there is none
 * (src/shortcode.cc:171-213: the first match in ascending message order -- code words are distinct, so "the" match) */
std::vector<int>
short_decode_blk (const std::vector<int>& coded_bits)
{
  assert (coded_bits.size() == g_short_out);
  ShortRow r { 0, 0 };
  for (size_t j = 0; j < g_short_out; j++)
    if (coded_bits[j])
      (j < 64 ? r.lo : r.hi) |= 1ull << (j & 63);
  std::vector<int> out_bits;
  for (uint32_t c = 0; c < (1u << g_short_in); c++)
    {
      const ShortRow w = short_codeword (c);
      if (w.lo == r.lo && w.hi == r.hi)
        {
          for (size_t bit = 0; bit < g_short_in; bit++)
            out_bits.push_back ((c >> bit) & 1);
          return out_bits;
        }
    }
  return out_bits;
}

size_t
code_size (ConvBlockType block_type, size_t msg_size)
{
  if (Params::payload_short)
    {
      assert (msg_size == g_short_in);
      return conv_code_size (block_type, g_short_out);
    }
  return conv_code_size (block_type, msg_size);
}

std::vector<int>
code_encode (ConvBlockType block_type, const std::vector<int>& in_bits)
{
  return Params::payload_short ? conv_encode (block_type, short_encode_blk (in_bits)) : conv_encode (block_type, in_bits);
}
