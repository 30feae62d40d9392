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

/* select_local_maxima + mask_avg_false_positives + select_threshold_and_n_best in one pass with the same result:
 * only the few best peaks can survive the threshold / n-best rule, so the (expensive) false-positive mask is
 * evaluated lazily in descending quality order instead of for every one of the ~n/3 local maxima of a chunk. */
void
select_candidates (vector<SearchScore>& scores, double threshold)
{
  select_local_maxima (scores);
  const size_t n = scores.size();
  constexpr int    mask_distance = local_mean_distance + 3;
  constexpr double mask_factor   = 3;
  vector<double> aq (n);
  vector<uint32_t> order (n);
  for (size_t i = 0; i < n; i++)
    {
      aq[i] = abs_quality (scores[i]);
      order[i] = i;
    }
  auto sign = [&] (size_t i) { return (scores[i].raw_quality - scores[i].local_mean < 0) ? -1 : 1; };
  auto masked = [&] (int i)
    {
      for (int d = -mask_distance; d <= mask_distance; d++)
        {
          const int j = i + d;
          if (j == i || j < 0 || j >= int (n))
            continue;
          const int distance = std::abs (int (scores[i].index) - int (scores[j].index)) / Params::sync_search_step;
          if (distance <= mask_distance && aq[j] > aq[i] * mask_factor && sign (j) != sign (i))
            return true;
        }
      return false;
    };
  vector<SearchScore> out;
  size_t sorted = 0, batch = 64;
  bool done = n == 0;
  while (!done)
    {
      const size_t end = std::min (n, sorted + batch);
      std::partial_sort (order.begin() + sorted, order.begin() + end, order.end(), [&] (uint32_t a, uint32_t b) { return aq[a] > aq[b]; });
      for (size_t k = sorted; k < end && !done; k++)
        {
          const uint32_t i = order[k];
          if (aq[i] <= threshold && int (out.size()) >= Params::get_n_best)
            done = true;                         // everything above the threshold is in, and at least n_best matches
          else if (!masked (i))
            out.push_back (scores[i]);
        }
      sorted = end;
      batch *= 4;
      if (sorted == n)
        done = true;
    }
  scores.swap (out);
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
          select_candidates (scores, Params::sync_threshold2 * 0.75);
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
