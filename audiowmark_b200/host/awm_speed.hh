// awm_speed.hh -- speed detection (reference src/wmspeed.hh:24-30) and resampling (src/resample.hh) entry points.
#pragma once
#include <vector>
#include "awm_params.hh"
#include "awm_random.hh"

struct DetectSpeedResult
{
  Key    key;
  double speed = 0;
};

/* detect_speed (src/wmspeed.cc:622-781): samples = one chunk at Params::mark_sample_rate, host or device memory */
struct DetectSpeedInfo { bool valid = false; double speed = 0, quality = 0; };   /* what the reference prints for the last key */
std::vector<DetectSpeedResult> detect_speed (const std::vector<Key>& key_list, const float *samples, size_t n_frames, int n_channels,
                                             int sample_rate, bool print_results, DetectSpeedInfo *info = nullptr);

/* resample_ratio (src/resample.cc:127-131): out gets lrint (n_frames * ratio) frames */
bool   resample_ratio (const float *in, size_t n_frames, int n_channels, double ratio, std::vector<float>& out);
/* number of frames the streaming resampler of WavChunkLoader / WatermarkResampler delivers for n_in input frames */
size_t resample_stream_frames (size_t n_in, double ratio);
/* outputs available from a streaming resampler after `fed` frames were written (no trailing frames yet) */
size_t resample_stream_available (size_t fed, double ratio);
