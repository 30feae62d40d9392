// awm_cli_tools.hh -- the WAV helper commands of the CLI (what the reference's tests/*.sh call besides add / get / cmp):
// test-gen-noise, cut-start, test-snr, test-info, test-clip, test-subtract, gentest, test-speed, test-change-speed,
// test-resample, gen-key (reference: src/audiowmark.cc:201-538).  Every function returns the process exit code.
#pragma once
#include <string>
#include "awm_random.hh"

namespace cli_tools {

int gen_key (const std::string& key_file, const std::string& key_name);
int gen_noise (const Key& key, const std::string& out_file, double seconds, int rate, int bits);
int gentest (const std::string& in_file, const std::string& out_file);
int cut_start (const std::string& in_file, const std::string& out_file, size_t frames);
int subtract (const std::string& file1, const std::string& file2, const std::string& out_file);
int snr (const std::string& orig_file, const std::string& wm_file);
int clip (const Key& key, const std::string& in_file, const std::string& out_file, int seed, int seconds);
int info (const std::string& in_file, const std::string& property);
int speed (const Key& key, int seed);
int change_speed (const std::string& in_file, const std::string& out_file, double speed);
int resample (const std::string& in_file, const std::string& out_file, int new_rate);

}
