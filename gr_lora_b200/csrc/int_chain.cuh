// int_chain.cuh -- the integer half of the LoRa decode chain as host/device functions:
// Gray, reduced-rate fold, deinterleave, deshuffle, dewhiten, Hamming, header parse.
// Bit-exact counterparts of lib/decoder_impl.cc:506-528 (demodulate tail), :535-565
// (deinterleave), :611-637 (deshuffle), :639-652 (dewhiten), :654-706 (hamming_decode,
// extract_data_only) and :833-847 (header parse).  Used by the K8 kernel and by the stream
// state machine (which must parse the header on the device to know the payload length).
#pragma once
#include "lora_common.cuh"
#include "lora_whitening.h"

namespace lb {

constexpr int LB_MAX_CW = 1024;        // code words buffered per frame (reference vectors are unbounded)
constexpr int LB_MAX_FRAME = 18 + 544; // loratap + phy + payload bytes

LB_HD uint32_t rotl_bits(uint32_t bits, uint32_t count, uint32_t size) {   // include/lora/utilities.h:96-103
    const uint32_t mask = (1u << size) - 1u;
    count %= size;
    bits &= mask;
    return count ? (((bits << count) & mask) | (bits >> (size - count))) : bits;
}

LB_HD uint32_t gray_encode(uint32_t bin) { return bin ^ (bin >> 1); }     // decoder_impl.cc:512

// std::lround(bin / 4.0f) % N_hdr (:508): bin/4 has at most 2 fractional bits, so the
// round-half-away-from-zero of lround is (bin + 2) >> 2 exactly.
LB_HD uint32_t reduce_bin(uint32_t bin, uint32_t n_bins_hdr) { return ((bin + 2u) >> 2) % n_bins_hdr; }

// one interleaver block: n_words Gray words -> ppm code words of n_words bits (:547-553)
LB_HD void deinterleave_block(const uint32_t *words, uint32_t n_words, uint32_t ppm, uint8_t *out) {
    for (uint32_t x = 0; x < ppm; x++) out[x] = 0;
    for (uint32_t i = 0; i < n_words; i++) {
        const uint32_t w = rotl_bits(words[i], i, ppm);
        for (uint32_t x = 0; x < ppm; x++) out[x] |= (uint8_t)(((w >> x) & 1u) << i);
    }
}

LB_HD uint8_t deshuffle_byte(uint8_t v) {      // pattern {5,0,1,2,4,3,6,7}, :568,:616-624
    return (uint8_t)(((v >> 5) & 1u) | (((v >> 0) & 1u) << 1) | (((v >> 1) & 1u) << 2) | (((v >> 2) & 1u) << 3) |
                     (((v >> 4) & 1u) << 4) | (((v >> 3) & 1u) << 5) | (v & 0xC0u));
}

LB_HD uint8_t hamming84_encode(uint8_t v) {    // hamming_encode_soft, include/lora/utilities.h:257-264
    const uint32_t d0 = v & 1u, d1 = (v >> 1) & 1u, d2 = (v >> 2) & 1u, d3 = (v >> 3) & 1u;
    const uint32_t p1 = d1 ^ d2 ^ d3, p2 = d0 ^ d1 ^ d2, p3 = d0 ^ d1 ^ d3, p4 = d0 ^ d2 ^ d3;
    return (uint8_t)(p1 | (d0 << 1) | (d1 << 2) | (d2 << 3) | (p2 << 4) | (d3 << 5) | (p3 << 6) | (p4 << 7));
}

// Hamming(8,4) decode of one received byte: stand-in for liquid-dsp's fec_decode table
// (decoder_impl.cc:661): nearest code word, lowest symbol on ties.
LB_HD uint8_t hamming84_decode(uint8_t cw) {
    int best = 0, best_d = 9;
    for (int s = 0; s < 16; s++) {
        uint32_t x = (uint32_t)(cw ^ hamming84_encode((uint8_t)s));
        int dd = 0;
        for (; x; x &= x - 1) dd++;
        if (dd < best_d) { best_d = dd; best = s; }
    }
    return (uint8_t)best;
}

LB_HD uint8_t extract_data(uint8_t v) {        // select_bits {1,2,3,5}, :694
    return (uint8_t)(((v >> 1) & 1u) | (((v >> 2) & 1u) << 1) | (((v >> 3) & 1u) << 2) | (((v >> 5) & 1u) << 3));
}

LB_HD uint8_t whitening_byte(int is_header, uint32_t cr, uint32_t i) {    // table choice :579-580
#ifdef __CUDA_ARCH__
    if (is_header) return i < LB_PRNG_HEADER_LEN ? lb_dev_prng_header[i] : 0;
    if (cr <= 2) return i < LB_PRNG_PAYLOAD_CR56_LEN ? lb_dev_prng_payload_cr56[i] : 0;
    return i < LB_PRNG_PAYLOAD_CR78_LEN ? lb_dev_prng_payload_cr78[i] : 0;
#else
    if (is_header) return i < LB_PRNG_HEADER_LEN ? lb_prng_header[i] : 0;
    if (cr <= 2) return i < LB_PRNG_PAYLOAD_CR56_LEN ? lb_prng_payload_cr56[i] : 0;
    return i < LB_PRNG_PAYLOAD_CR78_LEN ? lb_prng_payload_cr78[i] : 0;
#endif
}

// length (in code words) of the dewhitened vector and number of decoded bytes of decode()
LB_HD uint32_t decode_len_words(uint32_t n_cw, int is_header) { return is_header ? 6u : n_cw; }   // 5 + pad, :612,:633
LB_HD uint32_t decode_len_bytes(uint32_t len_words, uint32_t cr) {
    if (cr == 3 || cr == 4) {
        // ceil(len * 4.0f / (4.0f + cr)), :658 -- exact in integers for these sizes
        return (len_words * 4u + (3u + cr)) / (4u + cr);
    }
    if (cr == 1 || cr == 2) return (len_words + 1u) / 2u;
    return 0;
}

// decoded byte `i` of decode(is_header) given the deinterleaved code words (B2-B4)
LB_HD uint8_t decode_byte(const uint8_t *cw, uint32_t n_cw, int is_header, uint32_t cr, uint32_t i) {
    const uint32_t len = decode_len_words(n_cw, is_header);
    uint8_t w[2];
    for (uint32_t k = 0; k < 2; k++) {
        const uint32_t idx = 2u * i + k;
        uint8_t v = 0;
        if (idx < len) {
            const bool pad = is_header && idx == 5u;
            const uint8_t raw = (pad || idx >= n_cw) ? 0 : cw[idx];
            v = pad ? 0 : deshuffle_byte(raw);
            v ^= whitening_byte(is_header, cr, idx);
        }
        w[k] = v;
    }
    if (cr == 3 || cr == 4) {
        const uint8_t s0 = hamming84_decode(w[0]);                 // missing bytes read as 0 (oracle D3)
        const uint8_t s1 = hamming84_decode(w[1]);
        const uint8_t b = (uint8_t)((s0 << 4) | s1);
        return is_header ? b : (uint8_t)(((b & 0x0f) << 4) | ((b & 0xf0) >> 4));   // swap_nibbles :663
    }
    const uint8_t d1 = extract_data(w[0]);
    const uint8_t d2 = (2u * i + 1u < len) ? extract_data(w[1]) : 0;
    return is_header ? (uint8_t)((d1 << 4) | d2) : (uint8_t)((d2 << 4) | d1);      // :701-704
}

// payload symbol count, :842-847, evaluated with the reference's fp32 expressions
LB_HD int32_t payload_symbols(uint32_t payload_len, uint32_t cr, uint32_t sf, int reduced_rate) {
    const int spb = (int)cr + 4;
    const float bits_needed = (float)payload_len * 8.0f;
    const float symbols_needed = bits_needed * ((float)spb / 4.0f) / (float)(sf - (reduced_rate ? 2u : 0u));
    const int blocks = (int)ceilf(symbols_needed / (float)spb);
    return blocks * spb;
}

}  // namespace lb
