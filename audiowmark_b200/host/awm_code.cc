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

size_t
code_size (ConvBlockType block_type, size_t msg_size)
{
  assert (!Params::payload_short);
  return conv_code_size (block_type, msg_size);
}

std::vector<int>
code_encode (ConvBlockType block_type, const std::vector<int>& in_bits)
{
  assert (!Params::payload_short);
  return conv_encode (block_type, in_bits);
}

size_t
short_code_init (size_t)
{
  return 0;
}
