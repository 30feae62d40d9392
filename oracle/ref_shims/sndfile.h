/* Minimal stand-in for <sndfile.h>: only the types the reference headers
 * mention (src/sfinputstream.hh:27-33, src/sfoutputstream.hh).  libsndfile is
 * absent; sf_standin.cc provides SFInputStream/SFOutputStream bodies that use
 * the reference's own dependency-free WAV reader / RawConverter instead. */
#ifndef AWM_REF_SHIM_SNDFILE_H
#define AWM_REF_SHIM_SNDFILE_H
#include <stdint.h>
typedef int64_t sf_count_t;
typedef struct SNDFILE_tag SNDFILE;
typedef struct { sf_count_t frames; int samplerate, channels, format, sections, seekable; } SF_INFO;
typedef struct { void *get_filelen, *seek, *read, *write, *tell; } SF_VIRTUAL_IO;
#endif
