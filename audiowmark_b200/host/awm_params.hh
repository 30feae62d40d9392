// awm_params.hh -- configuration of the host side.
// Same fields and defaults as the reference's Params (src/wmcommon.hh:33-89, src/wmcommon.cc:27-58);
// kept as static members because the CLI / test scripts set them exactly like the reference does.
#pragma once
#include <string>
#include "awm_streams.hh"

enum class Format { AUTO = 1, RAW, RF64, WAV_PIPE };

class Params
{
public:
  static constexpr size_t frame_size      = 1024;
  static           int    frames_per_bit;
  static constexpr size_t bands_per_frame = 30;
  static constexpr int    max_band        = 100;
  static constexpr int    min_band        = 20;

  static double      water_delta;
  static std::string json_output;
  static bool        strict;
  static bool        mix;
  static bool        hard;
  static bool        snr;
  static bool        detect_speed;
  static bool        detect_speed_patient;
  static double      try_speed;
  static double      test_speed;
  static size_t      payload_size;
  static bool        payload_short;

  static constexpr int sync_bits           = 6;
  static constexpr int sync_frames_per_bit = 85;
  static constexpr int sync_search_step    = 256;
  static constexpr int sync_search_fine    = 8;
  static double        sync_threshold2;
  static int           get_n_best;

  static constexpr size_t frames_pad_start = 250;
  static constexpr int    mark_sample_rate = 44100;
  static constexpr double limiter_block_size_ms = 1000;
  static constexpr double limiter_ceiling       = 0.99;

  static double get_chunk_size;
  static int    test_cut;
  static bool   test_no_sync;
  static bool   test_no_limiter;
  static int    test_truncate;
  static int    expect_matches;

  static Format    input_format;
  static Format    output_format;
  static RawFormat raw_input_format;
  static RawFormat raw_output_format;

  static std::string input_label;
  static std::string output_label;

  static int gpu_device;     // CUDA device the context is created on (new: --gpu-device)
};
