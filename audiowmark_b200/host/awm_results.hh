// awm_results.hh -- ResultSet of `audiowmark get` (reference src/wmget.cc:163-474): pattern list,
// rating, ordering, chunk merge and the text / JSON printers whose format is the public output API.
#pragma once
#include <map>
#include <stdint.h>
#include <string>
#include <unordered_map>
#include <vector>
#include "awm_wm.hh"

class ResultSet
{
public:
  enum class Type { BLOCK, CLIP, ALL };
  struct Pattern
  {
    Key               key;
    double            time = 0;
    std::vector<int>  bit_vec;
    float             decode_error = 0;
    SyncFinder::Score sync_score;
    Type              type;
    double            speed = 0;
    double            rating = 0;
    bool approx_match (const Pattern& p) const;
  };
  void add_pattern (const Key& key, double time, SyncFinder::Score sync_score, const std::vector<int>& bit_vec,
                    float decode_error, Type pattern_type, double speed);
  void apply_time_offset (double time_offset);
  void sort (const std::vector<Key>& key_list);
  void merge (ResultSet& other);
  void print_json (FILE *outfile, size_t time_length);
  void print_json (size_t time_length, const std::string& json_file);
  void print (FILE *out = stdout);
  int  print_match_count (const std::vector<int>& orig_bits);
  int  match_count (const std::vector<int>& orig_bits) const;
  void set_debug_sync (const std::string& ds) { debug_sync = ds; }
  void print_debug_sync()                     { printf ("%s", debug_sync.c_str()); }
  const std::vector<Pattern>& all() const     { return patterns; }
private:
  std::vector<Pattern> patterns;
  std::string          debug_sync;
  void rate_patterns (const Key& key);
  /* merge index: patterns [0, indexed) by payload hash -> stream positions / combined patterns; rebuilt after anything that moves
   * or re-times patterns (sort, apply_time_offset) */
  struct Known { std::multimap<double, size_t> positioned; std::vector<size_t> combined; };
  std::unordered_map<uint64_t, Known> by_bits;
  size_t indexed = 0;
  void forget_index() { by_bits.clear(); indexed = 0; }
};
