// awm_results.cc -- ResultSet: what `audiowmark get` found, rated, ordered, merged across chunks and printed.
// Behaviour pinned by the reference (src/wmget.cc:163-474): the text and --json formats are the public output of the tool, the
// order of the patterns and the merge rule decide what a user sees.  The implementation below is organised around two small
// ideas: a pattern's sort position is a tuple (compared lexicographically), and both printers render one pattern to a string.
#include "awm_results.hh"
#include "awm_util.hh"

#include <algorithm>
#include <map>
#include <unordered_map>
#include <stdint.h>
#include <math.h>
#include <stdlib.h>
#include <tuple>

using std::string;
using std::vector;
typedef ResultSet::Pattern Pattern;

namespace {

const char *
block_letters (ConvBlockType t)
{
  return t == ConvBlockType::a ? "A" : t == ConvBlockType::b ? "B" : "AB";
}

/* A < B < AB inside one position */
int
block_order (ConvBlockType t)
{
  return t == ConvBlockType::a ? 0 : t == ConvBlockType::b ? 1 : t == ConvBlockType::ab ? 2 : 99;
}

/* "A" / "B" / "AB" (JSON: "ALL" for combined patterns), prefixed by CLIP-, suffixed by -SPEED */
string
type_label (const Pattern& p, bool for_json)
{
  string label = (for_json && p.type == ResultSet::Type::ALL) ? "ALL" : block_letters (p.sync_score.block_type);
  if (p.type == ResultSet::Type::CLIP)
    label.insert (0, "CLIP-");
  if (p.speed != 1)
    label += "-SPEED";
  return label;
}

string
minutes_seconds (double time, const char *fmt)
{
  const int seconds = int (time);
  return string_printf (fmt, seconds / 60, seconds % 60);
}

string
json_string (const string& raw)
{
  string out;
  for (unsigned char c : raw)
    switch (c)
      {
        case '"': case '\\': out += '\\'; out += char (c); break;
        default:          out += c < 32 ? string_printf ("\\u%04x", c) : string (1, char (c));
      }
  return out;
}

string
json_record (const Pattern& p)
{
  return string_printf ("    { \"key\": \"%s\", \"pos\": \"%s\", \"bits\": \"%s\", \"quality\": %.5f, \"error\": %.6f, \"rating\": %.5f, \"type\": \"%s\", \"speed\": %.6f }",
                        json_string (p.key.name()).c_str(), minutes_seconds (p.time, "%d:%02d").c_str(), bit_vec_to_str (p.bit_vec).c_str(),
                        p.sync_score.quality, p.decode_error, p.rating, type_label (p, true).c_str(), p.speed);
}

string
text_record (const Pattern& p)
{
  const string bits = bit_vec_to_str (p.bit_vec);
  if (p.type == ResultSet::Type::ALL)
    return string_printf ("pattern   all %s %.3f %.3f%s\n", bits.c_str(), p.sync_score.quality, p.decode_error, p.speed != 1 ? " SPEED" : "");
  return string_printf ("pattern %s %s %.3f %.3f %s\n", minutes_seconds (p.time, "%2d:%02d").c_str(), bits.c_str(), p.sync_score.quality, p.decode_error,
                        type_label (p, false).c_str());
}

}

bool
Pattern::approx_match (const Pattern& other) const
{
  /* the same detection seen from two overlapping chunks: same key / bits / kind, positions less than one frame apart (combined
   * patterns have no position), speeds within a percent */
  const double one_frame = Params::frame_size / double (Params::mark_sample_rate);
  if (type != other.type || sync_score.block_type != other.sync_score.block_type || !(fabs (speed - other.speed) < 0.01))
    return false;
  if (type != Type::ALL && !(fabs (time - other.time) < one_frame))          // cheap tests first: almost every pair fails here
    return false;
  return key == other.key && bit_vec == other.bit_vec;
}

void
ResultSet::add_pattern (const Key& key, double time, SyncFinder::Score sync_score, const vector<int>& bit_vec, float decode_error, Type pattern_type, double speed)
{
  patterns.emplace_back();
  Pattern& p = patterns.back();
  p.key = key;
  p.time = time;
  p.bit_vec = bit_vec;
  p.decode_error = decode_error;
  p.sync_score = sync_score;
  p.type = pattern_type;
  p.speed = speed;
}

void
ResultSet::apply_time_offset (double time_offset)
{
  std::for_each (patterns.begin(), patterns.end(), [time_offset] (Pattern& p) { p.time += time_offset; });
  forget_index();
}

/* rating of a pattern = summed sync quality of every pattern of the key that decoded to the same bits; combined ("all") patterns
 * weigh double */
void
ResultSet::rate_patterns (const Key& key)
{
  std::map<vector<int>, float> by_bits;
  for (const Pattern& p : patterns)
    if (p.key == key)
      by_bits[p.bit_vec] += p.sync_score.quality * (p.type == Type::ALL ? 2.f : 1.f);
  for (Pattern& p : patterns)
    if (p.key == key)
      p.rating = by_bits[p.bit_vec];
}

void
ResultSet::sort (const vector<Key>& key_list)
{
  for (const Key& key : key_list)
    rate_patterns (key);
  /* key name; best rated payload first; combined patterns after the blocks they were built from; stream position; A, B, AB; bits.
   * The positions are built once (the bit string is the last tie breaker and costly to make), then the patterns are permuted. */
  typedef std::tuple<string, double, bool, double, int, string> Position;
  vector<std::pair<Position, size_t>> order;
  order.reserve (patterns.size());
  for (size_t i = 0; i < patterns.size(); i++)
    {
      const Pattern& p = patterns[i];
      order.emplace_back (Position (p.key.name(), -p.rating, p.type == Type::ALL, p.time, block_order (p.sync_score.block_type), bit_vec_to_str (p.bit_vec)), i);
    }
  std::sort (order.begin(), order.end(), [] (const std::pair<Position, size_t>& x, const std::pair<Position, size_t>& y) { return x.first < y.first; });
  vector<Pattern> sorted;
  sorted.reserve (patterns.size());
  for (const auto& o : order)
    sorted.push_back (std::move (patterns[o.second]));
  patterns.swap (sorted);
  forget_index();
}

/* add the patterns of the next chunk (in time order) unless this set already holds the same detection (approx_match).  Only a
 * pattern with the same bits, and -- unless it is a combined one -- a position less than a frame away can be that detection, so the
 * candidates come from an index (payload -> positions) instead of a scan over everything found so far: the merged document of an
 * 8 h stream has ~900 patterns, and the merge runs on one rank while the others wait. */
void
ResultSet::merge (ResultSet& other)
{
  const double one_frame = Params::frame_size / double (Params::mark_sample_rate);
  auto bits_hash = [] (const vector<int>& bits)      /* FNV-1a over the payload bits: equal payloads are confirmed by approx_match */
    {
      uint64_t h = 1469598103934665603ull;
      for (int b : bits)
        h = (h ^ uint64_t (b & 1)) * 1099511628211ull;
      return h;
    };
  vector<Pattern> incoming;
  incoming.swap (other.patterns);                    /* the chunk result is consumed */
  std::stable_sort (incoming.begin(), incoming.end(), [] (const Pattern& x, const Pattern& y) { return x.time < y.time; });
  auto remember = [&] (size_t i)
    {
      Known& k = by_bits[bits_hash (patterns[i].bit_vec)];
      if (patterns[i].type == Type::ALL)
        k.combined.push_back (i);
      else
        k.positioned.emplace (patterns[i].time, i);
    };
  for (; indexed < patterns.size(); indexed++)      /* what earlier merges / add_pattern calls appended since the index was last complete */
    remember (indexed);
  patterns.reserve (patterns.size() + incoming.size());
  for (Pattern& p : incoming)
    {
      const Known& k = by_bits[bits_hash (p.bit_vec)];
      bool have = false;
      if (p.type == Type::ALL)
        have = std::any_of (k.combined.begin(), k.combined.end(), [&] (size_t i) { return patterns[i].approx_match (p); });
      else
        for (auto it = k.positioned.lower_bound (p.time - one_frame); it != k.positioned.end() && it->first <= p.time + one_frame && !have; ++it)
          have = patterns[it->second].approx_match (p);
      if (!have)
        {
          patterns.push_back (std::move (p));
          remember (indexed++);
        }
    }
  if (debug_sync.empty())
    debug_sync = other.debug_sync;
}

void
ResultSet::print_json (FILE *outfile, size_t time_length)
{
  string doc = string_printf ("{ \"length\": \"%ld:%02ld\",\n  \"matches\": [\n", long (time_length / 60), long (time_length % 60));
  for (size_t i = 0; i < patterns.size(); i++)
    doc += (i ? ",\n" : "") + json_record (patterns[i]);
  doc += " ]\n}\n";
  fputs (doc.c_str(), outfile);
}

void
ResultSet::print_json (size_t time_length, const string& json_file)
{
  FILE *outfile = fopen (json_file == "-" ? "/dev/stdout" : json_file.c_str(), "w");
  if (!outfile)
    {
      perror (("audiowmark: failed to open \"" + json_file + "\":").c_str());
      exit (127);
    }
  print_json (outfile, time_length);
  fclose (outfile);
}

/* text output: a "key" line whenever the key changes, followed by one "speed" line if any pattern of that key was found at a
 * speed other than 1, then the patterns */
void
ResultSet::print (FILE *out)
{
  string shown;                  /* key name of the block being printed; "" up front, so an unnamed key gets no "key" line */
  bool speed_pending = true;
  for (const Pattern& p : patterns)
    {
      if (p.key.name() != shown)
        {
          fprintf (out, "key %s\n", p.key.name().c_str());
          shown = p.key.name();
          speed_pending = true;
        }
      if (speed_pending)
        {
          auto stretched = std::find_if (patterns.begin(), patterns.end(), [&] (const Pattern& q) { return q.key == p.key && q.speed != 1; });
          if (stretched != patterns.end())
            fprintf (out, "speed %.6f\n", stretched->speed);
          speed_pending = false;
        }
      fputs (text_record (p).c_str(), out);
    }
}

int
ResultSet::match_count (const vector<int>& orig_bits) const
{
  return int (std::count_if (patterns.begin(), patterns.end(), [&] (const Pattern& p) { return p.bit_vec == orig_bits; }));
}

int
ResultSet::print_match_count (const vector<int>& orig_bits)
{
  const int n = match_count (orig_bits);
  printf ("match_count %d %zd\n", n, patterns.size());
  return n;
}
