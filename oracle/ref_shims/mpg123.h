/* Minimal stand-in for <mpg123.h> (MP3 input is out of scope; mp3_standin.cc
 * makes MP3InputStream report "not available"). */
#ifndef AWM_REF_SHIM_MPG123_H
#define AWM_REF_SHIM_MPG123_H
typedef struct mpg123_handle_struct mpg123_handle;
#endif
