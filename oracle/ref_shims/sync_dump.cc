/* sync_dump.cc -- TEST INFRASTRUCTURE: prints what SyncFinder::search of the UNMODIFIED reference returns.
 *
 * The reference binary never prints sync positions (patterns carry mm:ss only), but BASELINE.json's north_star asks for
 * bit-exact sync positions.  This driver links the reference's own object files (compiled where they lie by
 * oracle/Makefile.ref, everything except audiowmark.o) and calls the reference's public classes in the order
 * get_watermark / decode do (src/wmget.cc:886-1013): per chunk of WavChunkLoader one BLOCK mode search; for the first chunk
 * of a short file additionally the two zero padded CLIP mode searches of ClipDecoder::run_block (src/wmget.cc:823-866).
 * All arithmetic is the reference's; only this print loop is ours.
 *
 *   sync_dump in.wav        ->  "search <BLOCK|CLIP> <n_frames>" then one "score <index> <quality %.9g> <A|B>" line per score
 */
#include <stdio.h>
#include <algorithm>
#include <vector>

#include "wavdata.hh"
#include "wmcommon.hh"
#include "syncfinder.hh"
#include "wavchunkloader.hh"

using std::vector;

static void
dump (const vector<SyncFinder::KeyResult>& key_results, const char *mode, const WavData& wav_data)
{
  printf ("search %s %zd\n", mode, wav_data.n_values() / wav_data.n_channels());
  for (const auto& kr : key_results)
    for (const auto& s : kr.sync_scores)
      printf ("score %zd %.9g %s\n", s.index, s.quality, s.block_type == ConvBlockType::a ? "A" : "B");
}

static void
clip_search (const vector<Key>& key_list, const WavData& wav_data, bool at_end)
{
  /* the padding rule of ClipDecoder::run_block, restated */
  const size_t per_block = mark_sync_frame_count() + mark_data_frame_count();
  const size_t n = (per_block + 5) * Params::frame_size * wav_data.n_channels();
  size_t first = 0, last = std::min (n, wav_data.n_values()), pad_start = n, pad_end = n;
  if (!at_end)
    {
      if (last < n)
        pad_start += n - last;
    }
  else
    {
      if (wav_data.n_values() <= n)
        return;
      first = wav_data.n_values() - n;
      last = wav_data.n_values();
    }
  vector<float> ext (pad_start, 0.f);
  ext.insert (ext.end(), wav_data.samples().begin() + first, wav_data.samples().begin() + last);
  ext.insert (ext.end(), pad_end, 0.f);
  WavData padded (ext, wav_data.n_channels(), wav_data.sample_rate(), wav_data.bit_depth());
  SyncFinder sync_finder;
  dump (sync_finder.search (key_list, padded, SyncFinder::Mode::CLIP), "CLIP", padded);
}

int
main (int argc, char **argv)
{
  if (argc != 2)
    {
      fprintf (stderr, "usage: sync_dump in.wav\n");
      return 2;
    }
  vector<Key> key_list (1);                      /* the zero key, as audiowmark without --key */
  WavChunkLoader loader (argv[1]);
  bool first_chunk = true;
  while (!loader.done())
    {
      Error err = loader.load_next_chunk();
      if (err)
        {
          fprintf (stderr, "sync_dump: %s\n", err.message());
          return 1;
        }
      if (loader.done())
        break;
      const WavData& wav_data = loader.wav_data();
      printf ("chunk %.6f\n", loader.time_offset());
      SyncFinder sync_finder;
      dump (sync_finder.search (key_list, wav_data, SyncFinder::Mode::BLOCK), "BLOCK", wav_data);
      const size_t per_block = mark_sync_frame_count() + mark_data_frame_count();
      const int wav_frames = wav_data.n_values() / (Params::frame_size * wav_data.n_channels());
      if (first_chunk && wav_frames < per_block * 3.1)
        {
          clip_search (key_list, wav_data, false);
          clip_search (key_list, wav_data, true);
        }
      first_chunk = false;
    }
  return 0;
}
