// awm_sync.cc -- SyncFinder::search (reference src/syncfinder.cc:487-558): the scoring sweeps run on
// the GPU (awm_sync_approx = search_approx, awm_sync_refine = search_refine); candidate selection
// (local maxima, false-positive masking, threshold / n-best) stays on the host and follows the
// reference's comparison operators exactly (src/syncfinder.cc:258-391).
#include "awm_wm.hh"
#include "awm_engine.hh"
#include "awm_util.hh"

#include <algorithm>
#include <math.h>

using std::vector;

namespace {

typedef awm_search_score SearchScore;
inline double abs_quality (const SearchScore& s) { return fabs (s.raw_quality - s.local_mean); }
constexpr int local_mean_distance = 20;

void
select_local_maxima (vector<SearchScore>& scores)
{
  vector<SearchScore> selected;
  for (size_t i = 0; i < scores.size(); i++)
    {
      const double q = abs_quality (scores[i]);
      const double q_last = i > 0 ? abs_quality (scores[i - 1]) : 0;
      const double q_next = i + 1 < scores.size() ? abs_quality (scores[i + 1]) : 0;
      if (q >= q_last && q >= q_next)
        {
          selected.push_back (scores[i]);
          i++;       // the next score cannot be a local maximum
        }
    }
  scores.swap (selected);
}

/* subtracting the local mean biases the neighbourhood of a strong peak in the opposite direction:
 * drop peaks that have a 3x stronger peak of opposite sign within 23 search steps */
void
mask_avg_false_positives (vector<SearchScore>& scores)
{
  constexpr int    mask_distance = local_mean_distance + 3;
  constexpr double mask_factor   = 3;
  auto sign = [] (const SearchScore& s) { return (s.raw_quality - s.local_mean < 0) ? -1 : 1; };
  vector<SearchScore> out;
  for (int i = 0; i < int (scores.size()); i++)
    {
      bool mask = false;
      for (int d = -mask_distance; d <= mask_distance && !mask; d++)
        {
          const int j = i + d;
          if (j == i || j < 0 || j >= int (scores.size()))
            continue;
          const int distance = std::abs (int (scores[i].index) - int (scores[j].index)) / Params::sync_search_step;
          if (distance <= mask_distance && abs_quality (scores[j]) > abs_quality (scores[i]) * mask_factor && sign (scores[j]) != sign (scores[i]))
            mask = true;
        }
      if (!mask)
        out.push_back (scores[i]);
    }
  scores.swap (out);
}

void
select_threshold_and_n_best (vector<SearchScore>& scores, double threshold)
{
  std::sort (scores.begin(), scores.end(), [] (const SearchScore& a, const SearchScore& b) { return abs_quality (a) > abs_quality (b); });
  int i = 0;
  while (i < int (scores.size()) && abs_quality (scores[i]) > threshold)
    i++;
  if (i >= Params::get_n_best)
    scores.resize (i);                        // all matches above the threshold
  else if (int (scores.size()) > Params::get_n_best)
    scores.resize (Params::get_n_best);       // otherwise the n best
}

void
select_truncate_n (vector<SearchScore>& scores, size_t n)
{
  std::sort (scores.begin(), scores.end(), [] (const SearchScore& a, const SearchScore& b) { return abs_quality (a) > abs_quality (b); });
  if (scores.size() > n)
    scores.resize (n);
}

} // namespace

double
SyncFinder::normalize_sync_quality (double raw_quality)
{
  return raw_quality / std::min (Params::water_delta, 0.080) / 2.9;
}

vector<SyncFinder::KeyResult>
SyncFinder::search (const vector<Key>& key_list, size_t n_frames, int n_channels, Mode mode, size_t wav_first, size_t wav_last)
{
  vector<KeyResult> key_results;
  awm_ctx *ctx = Engine::ctx();
  const int amode = mode == Mode::CLIP ? AWM_MODE_CLIP : AWM_MODE_BLOCK;

  if (Params::test_no_sync)               // fake_sync (src/syncfinder.cc:460-485): the positions a clean file has
    {
      vector<Score> result_scores;
      if (mode == Mode::BLOCK)
        {
          const size_t expect0 = Params::frames_pad_start * Params::frame_size;
          const size_t expect_step = (mark_sync_frame_count() + mark_data_frame_count()) * Params::frame_size;
          const size_t expect_end = (n_frames / Params::frame_size) * Params::frame_size;
          int ab = 0;
          for (size_t expect_index = expect0; expect_index + expect_step < expect_end; expect_index += expect_step)
            result_scores.push_back (Score { expect_index, 1.0, (ab++ & 1) ? ConvBlockType::b : ConvBlockType::a });
        }
      for (const auto& key : key_list)
        key_results.push_back (KeyResult { key, result_scores });
      return key_results;
    }
  if (mode == Mode::BLOCK)                 // no special treatment of silence in block mode
    {
      wav_first = 0;
      wav_last = n_frames * n_channels;
    }
  for (const auto& key : key_list)
    {
      KeyResult key_result;
      key_result.key = key;
      const int slot = ctx ? Engine::key_slot (key) : -1;
      vector<SearchScore> scores;
      bool ok = slot >= 0;
      if (ok)
        {
          size_t n_scores = 0;
          ok = awm_sync_approx (ctx, slot, amode, wav_first, wav_last, Params::water_delta, nullptr, 0, &n_scores) == 0;
          if (ok && n_scores)
            {
              scores.resize (n_scores);
              ok = awm_sync_approx (ctx, slot, amode, wav_first, wav_last, Params::water_delta, scores.data(), scores.size(), &n_scores) == 0;
            }
        }
      if (ok)
        {
          select_local_maxima (scores);
          mask_avg_false_positives (scores);
          select_threshold_and_n_best (scores, Params::sync_threshold2 * 0.75);
          if (mode == Mode::CLIP)               // ClipDecoder: at most n_best matches, but at least 5
            select_truncate_n (scores, std::max (Params::get_n_best, 5));
          ok = awm_sync_refine (ctx, slot, amode, wav_first, wav_last, Params::water_delta, scores.data(), scores.size()) == 0;
        }
      if (!ok)
        {
          error ("audiowmark: sync search failed: %s\n", Engine::last_error().c_str());
          key_results.push_back (key_result);
          continue;
        }
      std::sort (scores.begin(), scores.end(), [] (const SearchScore& a, const SearchScore& b) { return a.index < b.index; });
      select_threshold_and_n_best (scores, Params::sync_threshold2);
      std::sort (scores.begin(), scores.end(), [] (const SearchScore& a, const SearchScore& b) { return a.index < b.index; });
      for (const auto& s : scores)
        {
          const double q = s.raw_quality - s.local_mean;
          key_result.sync_scores.push_back (Score { size_t (s.index), fabs (q), q > 0 ? ConvBlockType::a : ConvBlockType::b });
        }
      key_results.push_back (key_result);
    }
  return key_results;
}
