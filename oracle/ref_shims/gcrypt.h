/* Minimal stand-in for <gcrypt.h>: prototypes + ABI constants for the calls
 * the reference makes (src/random.cc:38-48,101-111,130-135,151,180,188).
 * The real runtime library libgcrypt.so.20 is linked, so the keyed PRNG of the
 * reference binary is the real reference arithmetic.  Test infrastructure only. */
#ifndef AWM_REF_SHIM_GCRYPT_H
#define AWM_REF_SHIM_GCRYPT_H
#include <stddef.h>
#include <stdlib.h>
#include <stdarg.h>
#include <string.h>   /* the real header pulls these in; src/random.cc relies on it for memset */
#include <sys/types.h>
#ifdef __cplusplus
extern "C" {
#endif
#define GCRYPT_VERSION "1.8.0"   /* minimum accepted by gcry_check_version */
typedef unsigned int gcry_error_t;
struct gcry_cipher_handle;
typedef struct gcry_cipher_handle *gcry_cipher_hd_t;
enum gcry_ctl_cmds { GCRYCTL_DISABLE_SECMEM = 37, GCRYCTL_INITIALIZATION_FINISHED = 38 };
enum gcry_cipher_algos { GCRY_CIPHER_AES128 = 7 };
enum gcry_cipher_modes { GCRY_CIPHER_MODE_ECB = 1, GCRY_CIPHER_MODE_CTR = 6 };
enum gcry_md_algos { GCRY_MD_SHA1 = 2 };
enum gcry_random_level { GCRY_WEAK_RANDOM = 0, GCRY_STRONG_RANDOM = 1, GCRY_VERY_STRONG_RANDOM = 2 };
const char  *gcry_check_version (const char *req_version);
gcry_error_t gcry_control (enum gcry_ctl_cmds cmd, ...);
gcry_error_t gcry_cipher_open (gcry_cipher_hd_t *handle, int algo, int mode, unsigned int flags);
void         gcry_cipher_close (gcry_cipher_hd_t h);
gcry_error_t gcry_cipher_setkey (gcry_cipher_hd_t hd, const void *key, size_t keylen);
gcry_error_t gcry_cipher_setctr (gcry_cipher_hd_t hd, const void *ctr, size_t ctrlen);
gcry_error_t gcry_cipher_encrypt (gcry_cipher_hd_t h, void *out, size_t outsize, const void *in, size_t inlen);
const char  *gcry_strsource (gcry_error_t err);
const char  *gcry_strerror (gcry_error_t err);
void         gcry_randomize (void *buffer, size_t length, enum gcry_random_level level);
void         gcry_md_hash_buffer (int algo, void *digest, const void *buffer, size_t length);
#ifdef __cplusplus
}
#endif
#endif
