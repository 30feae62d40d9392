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
/* --short <bits> needs the block-code generator matrices of src/shortcode.cc:28-83 (tabulated data,
 * not restated here): reports 0 = unsupported */
size_t           short_code_init (size_t k);
