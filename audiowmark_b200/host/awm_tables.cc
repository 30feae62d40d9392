#include "awm_tables.hh"
#include "awm_params.hh"
#include "awm_util.hh"

#include <algorithm>
#include <assert.h>

size_t
mark_data_frame_count()
{
  return code_size (ConvBlockType::a, Params::payload_size) * Params::frames_per_bit;
}

size_t
mark_sync_frame_count()
{
  return Params::sync_bits * Params::sync_frames_per_bit;
}

void
UpDownGen::get (int f, UpDownArray& up, UpDownArray& down)
{
  std::vector<int> bands (Params::max_band - Params::min_band + 1);
  for (size_t i = 0; i < bands.size(); i++)
    bands[i] = Params::min_band + i;
  m_random.seed (f, m_stream);              // one reseed per frame
  m_random.shuffle (bands);
  for (size_t i = 0; i < Params::bands_per_frame; i++)
    {
      up[i] = bands[i];
      down[i] = bands[Params::bands_per_frame + i];
    }
}

BitPosGen::BitPosGen (const Key& key)
{
  m_pos.resize (frames_per_block());
  for (size_t i = 0; i < m_pos.size(); i++)
    m_pos[i] = i;
  Random random (key, 0, Random::Stream::frame_position);
  random.shuffle (m_pos);
}

std::vector<MixEntry>
gen_mix_entries (const Key& key)
{
  const int frame_count = mark_data_frame_count();
  std::vector<MixEntry> entries;
  entries.reserve (frame_count * Params::bands_per_frame);
  UpDownGen up_down_gen (key, Random::Stream::data_up_down);
  BitPosGen bit_pos_gen (key);
  for (int f = 0; f < frame_count; f++)
    {
      UpDownArray up, down;
      up_down_gen.get (f, up, down);
      for (size_t i = 0; i < up.size(); i++)
        entries.push_back ({ bit_pos_gen.data_frame (f), up[i], down[i] });
    }
  /* --linear (src/wmget.cc:110-152): the same (frame, up, down) triples in generation order, no mixing */
  if (Params::mix)
    {
      Random random (key, 0, Random::Stream::mix);
      random.shuffle (entries);
    }
  return entries;
}

std::vector<unsigned>
bit_order (const Key& key, size_t n)
{
  std::vector<unsigned> order (n);
  for (size_t i = 0; i < n; i++)
    order[i] = i;
  Random random (key, 0, Random::Stream::bit_order);
  random.shuffle (order);
  return order;
}

SyncTable
gen_sync_table (const Key& key, int mode)
{
  SyncTable tab;
  const int first_block_end = frames_per_block();
  const int block_count = mode == AWM_MODE_CLIP ? 2 : 1;    // a "long" block repeats the sync pattern with up/down swapped
  UpDownGen up_down_gen (key, Random::Stream::sync_up_down);
  BitPosGen bit_pos_gen (key);
  tab.bit_offsets.push_back (0);
  for (int bit = 0; bit < Params::sync_bits; bit++)
    {
      std::vector<awm_sync_entry> bit_entries;
      for (int f = 0; f < Params::sync_frames_per_bit; f++)
        {
          const int sf = f + bit * Params::sync_frames_per_bit;
          UpDownArray up, down;
          up_down_gen.get (sf, up, down);
          for (int block = 0; block < block_count; block++)
            {
              awm_sync_entry e;
              e.frame = bit_pos_gen.sync_frame (sf) + block * first_block_end;
              const UpDownArray& u = block == 0 ? up : down;
              const UpDownArray& d = block == 0 ? down : up;
              for (size_t i = 0; i < u.size(); i++)
                {
                  e.up[i] = u[i] - Params::min_band;
                  e.down[i] = d[i] - Params::min_band;
                }
              std::sort (e.up, e.up + AWM_BANDS_PER_FRAME);
              std::sort (e.down, e.down + AWM_BANDS_PER_FRAME);
              bit_entries.push_back (e);
            }
        }
      std::sort (bit_entries.begin(), bit_entries.end(), [] (const awm_sync_entry& a, const awm_sync_entry& b) { return a.frame < b.frame; });
      tab.entries.insert (tab.entries.end(), bit_entries.begin(), bit_entries.end());
      tab.bit_offsets.push_back (tab.entries.size());
    }
  return tab;
}

std::vector<uint8_t>
gen_frame_mod_ab (const Key& key, const std::vector<int>& bitvec)
{
  enum { KEEP = 0, UP = 1, DOWN = 2 };
  const size_t fpb = frames_per_block(), n_bins = Params::max_band + 1;
  std::vector<uint8_t> tab (2 * fpb * n_bins, KEEP);
  const std::vector<MixEntry> mix_entries = Params::mix ? gen_mix_entries (key) : std::vector<MixEntry>();
  BitPosGen bit_pos_gen (key);
  for (int ab = 0; ab < 2; ab++)
    {
      uint8_t *fm = &tab[ab * fpb * n_bins];
      auto mark = [&] (int frame, const UpDownArray& up, const UpDownArray& down, int data_bit)
        {
          for (int u : up)   fm[frame * n_bins + u] = data_bit ? UP : DOWN;
          for (int d : down) fm[frame * n_bins + d] = data_bit ? DOWN : UP;
        };
      const std::vector<int> fec = randomize_bit_order (key, code_encode (ab ? ConvBlockType::b : ConvBlockType::a, bitvec), /* encode */ true);
      /* sync frames: 010101 for A, 101010 for B, written in linear order */
      UpDownGen sync_gen (key, Random::Stream::sync_up_down);
      for (int f = 0; f < int (mark_sync_frame_count()); f++)
        {
          UpDownArray up, down;
          sync_gen.get (f, up, down);
          mark (bit_pos_gen.sync_frame (f), up, down, (f / Params::sync_frames_per_bit + ab) & 1);
        }
      /* data frames */
      const int data_frames = mark_data_frame_count();
      if (Params::mix)
        {
          for (int f = 0; f < data_frames; f++)
            for (size_t fb = 0; fb < Params::bands_per_frame; fb++)
              {
                const MixEntry& me = mix_entries[f * Params::bands_per_frame + fb];
                const int data_bit = fec[f / Params::frames_per_bit];
                fm[me.frame * n_bins + me.up]   = data_bit ? UP : DOWN;
                fm[me.frame * n_bins + me.down] = data_bit ? DOWN : UP;
              }
        }
      else
        {
          UpDownGen data_gen (key, Random::Stream::data_up_down);
          for (int f = 0; f < data_frames; f++)
            {
              UpDownArray up, down;
              data_gen.get (f, up, down);
              mark (bit_pos_gen.data_frame (f), up, down, fec[f / Params::frames_per_bit]);
            }
        }
    }
  return tab;
}

std::vector<int>
parse_payload (const std::string& bits)
{
  std::vector<int> bitvec = bit_str_to_vec (bits);
  if (bitvec.empty())
    {
      error ("audiowmark: cannot parse bits '%s'\n", bits.c_str());
      return {};
    }
  if ((Params::payload_short || Params::strict) && bitvec.size() != Params::payload_size)
    {
      error ("audiowmark: number of message bits must match payload size (%zd bits)\n", Params::payload_size);
      return {};
    }
  if (bitvec.size() > Params::payload_size)
    {
      error ("audiowmark: number of bits in message '%s' larger than payload size\n", bits.c_str());
      return {};
    }
  if (bitvec.size() < Params::payload_size)
    {
      /* short messages are repeated cyclically (disabled by --strict) */
      std::vector<int> expanded (Params::payload_size);
      for (size_t i = 0; i < expanded.size(); i++)
        expanded[i] = bitvec[i % bitvec.size()];
      bitvec = expanded;
    }
  return bitvec;
}
