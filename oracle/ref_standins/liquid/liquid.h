// stand-in for <liquid/liquid.h>: the FFT plan and Hamming(8,4) FEC calls of lib/decoder_impl.cc
// (:112-113,116-117,136-138,443,459,661).  liquid-dsp is not in the container; this follows its published
// behaviour: power-of-two plans run an fp32 radix-2 decimation-in-time FFT (bit-reversed copy, then log2 n
// passes with twiddles cexpf(-+ j 2 pi k / n)), other sizes a direct DFT; fec_decode(HAMMING84) maps every
// received byte to the nearest of the 16 code words (the same code book as include/lora/utilities.h:257-264)
// and packs two decoded nibbles per output byte, first nibble high.  Ties between equally near code words
// (only possible with >= 2 bit errors) resolve to the lowest symbol: that choice is liquid's table and is
// NOT pinned here.
#pragma once
#include <cmath>
#include <complex>
#include <cstdlib>
#include <vector>

typedef std::complex<float> liquid_float_complex;

#define LIQUID_FFT_FORWARD  (+1)
#define LIQUID_FFT_BACKWARD (-1)

struct fftplan_s {
    unsigned int nfft;
    liquid_float_complex *x, *y;
    int direction;
    bool radix2;
    unsigned int m;
    std::vector<unsigned int> index_rev;
    std::vector<liquid_float_complex> twiddle;
};
typedef fftplan_s *fftplan;

static inline fftplan fft_create_plan(unsigned int nfft, liquid_float_complex *x, liquid_float_complex *y, int dir, int /*flags*/) {
    fftplan q = new fftplan_s;
    q->nfft = nfft; q->x = x; q->y = y; q->direction = dir;
    q->radix2 = nfft >= 2 && (nfft & (nfft - 1)) == 0;
    q->m = 0;
    while ((1u << q->m) < nfft) q->m++;
    const float d = dir == LIQUID_FFT_FORWARD ? -1.0f : 1.0f;
    q->twiddle.resize(nfft);
    for (unsigned int i = 0; i < nfft; i++)
        q->twiddle[i] = std::exp(liquid_float_complex(0.0f, (float)(d * 2.0 * M_PI * (double)i / (double)nfft)));
    if (q->radix2) {
        q->index_rev.resize(nfft);
        for (unsigned int i = 0; i < nfft; i++) {
            unsigned int r = 0;
            for (unsigned int b = 0; b < q->m; b++) r |= ((i >> b) & 1u) << (q->m - 1 - b);
            q->index_rev[i] = r;
        }
    }
    return q;
}
static inline void fft_destroy_plan(fftplan q) { delete q; }
static inline void fft_execute(fftplan q) {
    const unsigned int n = q->nfft;
    liquid_float_complex *y = q->y;
    if (!q->radix2) {
        for (unsigned int k = 0; k < n; k++) {
            liquid_float_complex acc(0.0f, 0.0f);
            for (unsigned int i = 0; i < n; i++) acc += q->x[i] * q->twiddle[(unsigned int)(((unsigned long long)i * k) % n)];
            y[k] = acc;
        }
        return;
    }
    for (unsigned int i = 0; i < n; i++) y[i] = q->x[q->index_rev[i]];
    unsigned int n1 = 0, n2 = 1, stride = n;
    for (unsigned int s = 0; s < q->m; s++) {
        n1 = n2; n2 *= 2; stride >>= 1;
        unsigned int tw = 0;
        for (unsigned int j = 0; j < n1; j++) {
            const liquid_float_complex t = q->twiddle[tw];
            tw = (tw + stride) % n;
            for (unsigned int k = j; k < n; k += n2) {
                const liquid_float_complex yp = y[k + n1] * t;
                y[k + n1] = y[k] - yp;
                y[k] += yp;
            }
        }
    }
}

typedef enum { LIQUID_FEC_UNKNOWN = 0, LIQUID_FEC_NONE, LIQUID_FEC_REP3, LIQUID_FEC_REP5, LIQUID_FEC_HAMMING74,
               LIQUID_FEC_HAMMING84, LIQUID_FEC_HAMMING128 } fec_scheme;
struct fec_s {
    fec_scheme scheme;
    unsigned char dec[256];
};
typedef fec_s *fec;

static inline fec fec_create(fec_scheme scheme, void * /*opts*/) {
    static const unsigned char enc[16] = {0x00, 0xd2, 0x55, 0x87, 0x99, 0x4b, 0xcc, 0x1e, 0xe1, 0x33, 0xb4, 0x66, 0x78, 0xaa, 0x2d, 0xff};
    if (scheme != LIQUID_FEC_HAMMING84) abort();      // the reference only ever asks for HAMMING84 (:116)
    fec q = new fec_s;
    q->scheme = scheme;
    for (unsigned int r = 0; r < 256; r++) {
        int best = 9;
        unsigned char sym = 0;
        for (unsigned int s = 0; s < 16; s++) {
            const int dist = __builtin_popcount(r ^ enc[s]);
            if (dist < best) { best = dist; sym = (unsigned char)s; }
        }
        q->dec[r] = sym;
    }
    return q;
}
static inline void fec_destroy(fec q) { delete q; }
static inline int fec_decode(fec q, unsigned int dec_msg_len, unsigned char *msg_enc, unsigned char *msg_dec) {
    for (unsigned int i = 0; i < dec_msg_len; i++)
        msg_dec[i] = (unsigned char)((q->dec[msg_enc[2 * i]] << 4) | q->dec[msg_enc[2 * i + 1]]);
    return 0;
}
