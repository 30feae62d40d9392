// awm_code.hh -- channel code sizes + encoder (the decoder is the GPU Viterbi, awm_viterbi).
// Reference: src/convcode.cc:42-125, src/shortcode.cc:117-133.
#pragma once
#include <stddef.h>
#include <vector>

enum class ConvBlockType { a, b, ab };

size_t           conv_code_size (ConvBlockType block_type, size_t msg_size);
std::vector<int> conv_encode (ConvBlockType block_type, const std::vector<int>& in_bits);

/* code_* dispatch between the plain convolutional code and the (deprecated) short payload mode */
size_t           code_size (ConvBlockType block_type, size_t msg_size);
std::vector<int> code_encode (ConvBlockType block_type, const std::vector<int>& in_bits);
/* --short <bits> (12, 16 or 20): selects the block code, returns its length n (0 = unsupported size) */
size_t           short_code_init (size_t k);
std::vector<int> short_encode_blk (const std::vector<int>& in_bits);
std::vector<int> short_decode_blk (const std::vector<int>& coded_bits);   /* empty: no code word matches */
/* message bits of the convolutional code: the payload, or the block code word in short mode */
size_t           code_message_bits();
