/*
 * lora_oracle.c -- CPU ORACLE (test infrastructure, see lora_oracle.h for scope and pinning).
 *
 * Every function cites the reference lines it restates (paths relative to the reference
 * root, rpp0/gr-lora @ 90343d45).  Arithmetic types follow the reference expression by
 * expression (float vs double promotions included) because thresholds and argmax
 * decisions downstream depend on them.
 */
#include "lora_oracle.h"
#include "lo_tables.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define LO_MAC_CRC_SIZE 2u        /* include/lora/utilities.h:29 */
#define LO_PWR_QUEUE 4            /* include/lora/utilities.h:30 */

/* ---- tiny growable byte / word vectors (std::vector stand-ins) ------------------- */
typedef struct { uint8_t *p; size_t n, cap; } vec8;
typedef struct { uint32_t *p; size_t n, cap; } vec32;

static void v8_push(vec8 *v, uint8_t b) {
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 64; v->p = (uint8_t *)realloc(v->p, v->cap); }
    v->p[v->n++] = b;
}
static void v8_erase_front(vec8 *v, size_t k) {
    if (k > v->n) k = v->n;
    memmove(v->p, v->p + k, v->n - k);
    v->n -= k;
}
static void v32_push(vec32 *v, uint32_t w) {
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 16; v->p = (uint32_t *)realloc(v->p, v->cap * sizeof(uint32_t)); }
    v->p[v->n++] = w;
}

typedef struct { uint8_t *data; size_t len; } lo_frame;

struct lo_decoder {
    /* members of decoder_impl, lib/decoder_impl.h:70-123 */
    int       state;
    lo_cf    *downchirp, *upchirp;
    float    *downchirp_ifreq, *upchirp_ifreq, *upchirp_ifreq_v;
    int       implicit, reduced_rate;
    uint8_t   sf;
    uint32_t  bw;
    uint8_t   phdr[3];                /* loraphy_header_t, include/lora/loraphy.h:25-32 */
    double    bits_per_second, symbols_per_second, bits_per_symbol, period, dt;
    uint32_t  delay_after_sync, samples_per_second, sps, n_bins, n_bins_hdr, decim;
    int32_t   payload_symbols;
    uint32_t  payload_length, corr_fails;
    float     energy_threshold, snr;
    float     pwr_queue[LO_PWR_QUEUE]; int pwr_n, pwr_head;   /* boost::circular_buffer<float>(4) */
    vec32     words;
    vec8      demodulated, decoded;
    int       enable_fine_sync;
    int32_t   fine_sync;
    /* oracle-only */
    int       demod_method;
    lo_cf    *fft_tw;                 /* twiddles for the sps-point FFT */
    lo_cf    *fft_buf;
    uint32_t *fft_rev;
    float    *scratch_f;              /* 2*sps floats */
    lo_frame *frames; size_t n_frames, cap_frames;
    char     *out; size_t out_n, out_cap;
    int32_t   last_bin;
};

/* ---- phy header bitfield access (include/lora/loraphy.h:25-32, LSB-first bitfields) - */
static uint8_t hdr_cr(const lo_decoder *d) { return (uint8_t)(d->phdr[1] >> 5); }
static uint8_t hdr_has_crc(const lo_decoder *d) { return (uint8_t)((d->phdr[1] >> 4) & 1u); }
static void hdr_set_cr(lo_decoder *d, uint8_t cr) { d->phdr[1] = (uint8_t)((d->phdr[1] & 0x1f) | ((cr & 7u) << 5)); }
static void hdr_set_crc(lo_decoder *d, int crc) { d->phdr[1] = (uint8_t)((d->phdr[1] & 0xef) | ((crc ? 1u : 0u) << 4)); }

static void out_printf(lo_decoder *d, const char *fmt, ...) {
    char tmp[512];
    va_list ap;
    va_start(ap, fmt);
    int k = vsnprintf(tmp, sizeof tmp, fmt, ap);
    va_end(ap);
    if (k < 0) return;
    if (d->out_n + (size_t)k + 1 > d->out_cap) {
        d->out_cap = (d->out_n + (size_t)k + 1) * 2;
        d->out = (char *)realloc(d->out, d->out_cap);
    }
    memcpy(d->out + d->out_n, tmp, (size_t)k + 1);
    d->out_n += (size_t)k;
}

/* print_vector_hex, include/lora/utilities.h:351-368 */
static void out_hex(lo_decoder *d, const uint8_t *v, size_t have, uint32_t size, int endline, int ascii) {
    char asc[1024]; size_t an = 0;
    for (uint32_t i = 0; i < size; i++) {
        uint8_t b = i < have ? v[i] : 0;                 /* D5 */
        out_printf(d, " %02x", b);
        if (b >= ' ' && b <= '~' && an + 1 < sizeof asc) asc[an++] = (char)b;
    }
    asc[an] = 0;
    if (ascii) out_printf(d, " (%s)", asc);
    if (endline) out_printf(d, "\n");
}

/* ==================================================================================== */
/* A3  instantaneous_frequency, lib/decoder_impl.cc:224-244                              */
/* ==================================================================================== */
void lo_instantaneous_frequency(const lo_cf *in, float *out, uint32_t window) {
    if (window < 2u) return;                                   /* :225-228 */
    for (uint32_t i = 1u; i < window; i++) {
        const float iphase_1 = atan2f(in[i - 1].im, in[i - 1].re);   /* std::arg, :232 */
        float iphase_2 = atan2f(in[i].im, in[i].re);                 /* :233 */
        /* :236-237 -- float difference compared against the double M_PI, update in double */
        while ((iphase_2 - iphase_1) > M_PI) iphase_2 = (float)(iphase_2 - 2.0f * M_PI);
        while ((iphase_2 - iphase_1) < -M_PI) iphase_2 = (float)(iphase_2 + 2.0f * M_PI);
        out[i - 1] = iphase_2 - iphase_1;                            /* :239 */
    }
    out[window - 1] = out[window - 2];                               /* :243 */
}

/* ==================================================================================== */
/* A2  build_ideal_chirps, lib/decoder_impl.cc:141-175                                   */
/* ==================================================================================== */
static lo_cf expj_times_1p1j(float phase) {
    /* gr_complex(1,1) * gr_expj(phase); gr_expj takes a FLOAT phase and uses sincosf (:159) */
    const float c = cosf(phase), s = sinf(phase);
    lo_cf r; r.re = 1.0f * c - 1.0f * s; r.im = 1.0f * s + 1.0f * c;
    return r;
}

static void build_ideal_chirps(lo_decoder *d) {
    const uint32_t sps = d->sps;
    const double T = -0.5 * d->bw * d->symbols_per_second;          /* :149 */
    const double f0 = (d->bw / 2.0);                                /* :150 */
    const double pre_dir = 2.0 * M_PI;                              /* :151 */
    lo_cf *tmp = (lo_cf *)malloc(sizeof(lo_cf) * sps * 3);
    for (uint32_t i = 0u; i < sps; i++) {
        const double t = d->dt * i;                                 /* :158 */
        d->downchirp[i] = expj_times_1p1j((float)(pre_dir * t * (f0 + T * t)));            /* :159 */
        d->upchirp[i] = expj_times_1p1j((float)(pre_dir * t * (f0 + T * t) * -1.0f));      /* :160 */
    }
    lo_instantaneous_frequency(d->downchirp, d->downchirp_ifreq, sps);  /* :164 */
    lo_instantaneous_frequency(d->upchirp, d->upchirp_ifreq, sps);      /* :165 */
    for (int k = 0; k < 3; k++) memcpy(tmp + (size_t)k * sps, d->upchirp, sizeof(lo_cf) * sps);   /* :171-173 */
    lo_instantaneous_frequency(tmp, d->upchirp_ifreq_v, sps * 3);       /* :174 */
    free(tmp);
}

/* ==================================================================================== */
/* lifecycle, A1: lib/decoder_impl.cc:41-122                                             */
/* ==================================================================================== */
static void fft_init(lo_decoder *d) {
    const uint32_t n = d->sps;
    d->fft_tw = (lo_cf *)malloc(sizeof(lo_cf) * (n / 2 + 1));
    d->fft_buf = (lo_cf *)malloc(sizeof(lo_cf) * n);
    d->fft_rev = (uint32_t *)malloc(sizeof(uint32_t) * n);
    for (uint32_t k = 0; k < n / 2; k++) {
        const double a = -2.0 * M_PI * (double)k / (double)n;       /* forward DFT: e^{-j2pi kn/N} */
        d->fft_tw[k].re = (float)cos(a);
        d->fft_tw[k].im = (float)sin(a);
    }
    uint32_t bits = 0; while ((1u << bits) < n) bits++;
    for (uint32_t i = 0; i < n; i++) {
        uint32_t r = 0;
        for (uint32_t b = 0; b < bits; b++) if (i & (1u << b)) r |= 1u << (bits - 1 - b);
        d->fft_rev[i] = r;
    }
}

lo_decoder *lo_create(float samp_rate, uint32_t bandwidth, uint8_t sf, int implicit, uint8_t cr,
                      int crc, int reduced_rate, int disable_drift_correction) {
    if (sf < 6 || sf > 13) return NULL;                             /* :57-61 (reference exit(1)) */
    lo_decoder *d = (lo_decoder *)calloc(1, sizeof *d);
    d->state = LO_DETECT;                                           /* :55 */
    d->bw = bandwidth;                                              /* :69 */
    d->implicit = implicit != 0;
    d->reduced_rate = reduced_rate != 0;
    hdr_set_cr(d, cr);                                              /* :72, 3-bit field */
    hdr_set_crc(d, crc);                                            /* :73 */
    d->samples_per_second = (uint32_t)samp_rate;                    /* :74, float -> uint32_t member */
    d->payload_symbols = 0;
    d->dt = 1.0f / d->samples_per_second;                           /* :77, float divide stored in double */
    d->sf = sf;
    d->bits_per_second = (double)d->sf * (double)(4.0 / (4.0 + hdr_cr(d))) / (1u << d->sf) * d->bw;   /* :79 */
    d->symbols_per_second = (double)d->bw / (1u << d->sf);          /* :80 */
    d->period = 1.0f / (double)d->symbols_per_second;               /* :81 */
    d->bits_per_symbol = (double)(d->bits_per_second / d->symbols_per_second);   /* :82 */
    d->sps = (uint32_t)(d->samples_per_second / d->symbols_per_second);          /* :83 */
    d->delay_after_sync = d->sps / 4u;                              /* :84 */
    d->n_bins = (uint32_t)(1u << d->sf);                            /* :85 */
    d->n_bins_hdr = (uint32_t)(1u << (d->sf - 2));                  /* :86 */
    d->decim = d->sps / d->n_bins;                                  /* :87 */
    d->energy_threshold = 0.0f;
    d->fine_sync = 0;
    d->enable_fine_sync = !disable_drift_correction;                /* :90 */
    d->snr = 1.0f;                                                  /* D4 */
    d->demod_method = LO_DEMOD_GRADIENT;
    d->last_bin = -1;

    /* banner, :93-103 (std::cout default double format == %g) */
    out_printf(d, "Bits (nominal) per symbol: \t%g\n", d->bits_per_symbol);
    out_printf(d, "Bins per symbol: \t%u\n", d->n_bins);
    out_printf(d, "Samples per symbol: \t%u\n", d->sps);
    out_printf(d, "Decimation: \t\t%u\n", d->decim);
    if (!d->enable_fine_sync) out_printf(d, "Warning: clock drift correction disabled\n");
    if (d->implicit) {
        out_printf(d, "CR: \t\t%d\n", (int)hdr_cr(d));
        out_printf(d, "CRC: \t\t%d\n", (int)hdr_has_crc(d));
    }

    const uint32_t sps = d->sps;
    d->downchirp = (lo_cf *)malloc(sizeof(lo_cf) * sps);
    d->upchirp = (lo_cf *)malloc(sizeof(lo_cf) * sps);
    d->downchirp_ifreq = (float *)malloc(sizeof(float) * sps);
    d->upchirp_ifreq = (float *)malloc(sizeof(float) * sps);
    d->upchirp_ifreq_v = (float *)malloc(sizeof(float) * sps * 3);
    d->scratch_f = (float *)malloc(sizeof(float) * sps * 2);
    build_ideal_chirps(d);                                          /* :106 */
    fft_init(d);                                                    /* :109-113 (liquid plan stand-in) */
    return d;
}

void lo_destroy(lo_decoder *d) {
    if (!d) return;
    free(d->downchirp); free(d->upchirp); free(d->downchirp_ifreq); free(d->upchirp_ifreq);
    free(d->upchirp_ifreq_v); free(d->scratch_f); free(d->fft_tw); free(d->fft_buf); free(d->fft_rev);
    free(d->words.p); free(d->demodulated.p); free(d->decoded.p);
    lo_frames_clear(d); free(d->frames); free(d->out);
    free(d);
}

void lo_set_demod(lo_decoder *d, int method) { d->demod_method = method; }
uint32_t lo_sps(const lo_decoder *d) { return d->sps; }
uint32_t lo_bins(const lo_decoder *d) { return d->n_bins; }
uint32_t lo_decim(const lo_decoder *d) { return d->decim; }
double lo_bits_per_symbol(const lo_decoder *d) { return d->bits_per_symbol; }
const lo_cf *lo_downchirp(const lo_decoder *d) { return d->downchirp; }
const lo_cf *lo_upchirp(const lo_decoder *d) { return d->upchirp; }
const float *lo_downchirp_ifreq(const lo_decoder *d) { return d->downchirp_ifreq; }
const float *lo_upchirp_ifreq(const lo_decoder *d) { return d->upchirp_ifreq; }
const float *lo_upchirp_ifreq_v(const lo_decoder *d) { return d->upchirp_ifreq_v; }
int lo_state(const lo_decoder *d) { return d->state; }
const char *lo_stdout(const lo_decoder *d) { return d->out ? d->out : ""; }

/* ==================================================================================== */
/* A4  get_shift_fft, lib/decoder_impl.cc:430-464 (dormant in the reference, :500)       */
/* liquid-dsp's fft_execute (unpinned third-party) is restated as an unnormalised       */
/* forward radix-2 DIT FFT in fp32 with double-derived twiddles.                         */
/* ==================================================================================== */
static void fft_forward(lo_decoder *d, lo_cf *x) {
    const uint32_t n = d->sps;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t r = d->fft_rev[i];
        if (r > i) { lo_cf t = x[i]; x[i] = x[r]; x[r] = t; }
    }
    for (uint32_t half = 1; half < n; half <<= 1) {
        const uint32_t step = n / (2 * half);
        for (uint32_t base = 0; base < n; base += 2 * half) {
            for (uint32_t k = 0; k < half; k++) {
                const lo_cf w = d->fft_tw[k * step];
                lo_cf *a = &x[base + k], *b = &x[base + k + half];
                const float tr = b->re * w.re - b->im * w.im;
                const float ti = b->re * w.im + b->im * w.re;
                b->re = a->re - tr; b->im = a->im - ti;
                a->re = a->re + tr; a->im = a->im + ti;
            }
        }
    }
}

uint32_t lo_get_shift_fft(lo_decoder *d, const lo_cf *samples, float *mag_out) {
    const uint32_t sps = d->sps, N = d->n_bins;
    lo_cf *m = d->fft_buf;
    for (uint32_t i = 0u; i < sps; i++) {                            /* :436-438, plain (not conj) product */
        const lo_cf a = samples[i], b = d->downchirp[i];
        m[i].re = a.re * b.re - a.im * b.im;
        m[i].im = a.re * b.im + a.im * b.re;
    }
    fft_forward(d, m);                                               /* :443 */
    /* :447-450 decimate: tmp[0,N/2) = F[0,N/2); tmp[N/2,N) = F[sps-N/2, sps); tmp[N/2] += F[N/2] */
    uint32_t best = 0; float best_mag = -1.0f;
    for (uint32_t i = 0u; i < N; i++) {
        lo_cf t = (i < (N + 1u) / 2u) ? m[i] : m[sps - N / 2u + (i - (N + 1u) / 2u)];
        if (i == N / 2u) { t.re += m[N / 2u].re; t.im += m[N / 2u].im; }
        const float mag = hypotf(t.re, t.im);                        /* std::abs, :454 */
        if (mag > best_mag) { best_mag = mag; best = i; }            /* std::max_element: first max, :463 */
    }
    if (mag_out) *mag_out = best_mag;
    return best;
}

/* ==================================================================================== */
/* A5  max_frequency_gradient_idx, lib/decoder_impl.cc:466-491 (live demodulator)        */
/* ==================================================================================== */
uint32_t lo_max_frequency_gradient_idx(lo_decoder *d, const lo_cf *samples) {
    const uint32_t N = d->n_bins, decim = d->decim;
    float *ifreq = d->scratch_f;
    float *avg = (float *)malloc(sizeof(float) * N);
    lo_instantaneous_frequency(samples, ifreq, d->sps);              /* :472 */
    for (uint32_t i = 0; i < N; i++) {                               /* :474-477 */
        float acc = 0.0f;                                            /* volk_32f_accumulator_s32f */
        for (uint32_t k = 0; k < decim; k++) acc += ifreq[i * decim + k];
        avg[i] = acc / decim;
    }
    float max_gradient = 0.1f;                                       /* :479 */
    uint32_t max_index = 0;
    for (uint32_t i = 1u; i < N; i++) {                              /* :482-488 */
        const float gradient = avg[i - 1] - avg[i];
        if (gradient > max_gradient) { max_gradient = gradient; max_index = i + 1; }
    }
    free(avg);
    return (N - max_index) % N;                                      /* :490 */
}

/* ==================================================================================== */
/* A6  fine_sync + cross_correlate_ifreq_fast, lib/decoder_impl.cc:300-338, :259-263     */
/* ==================================================================================== */
static float dot_f(const float *a, const float *b, uint32_t n) {     /* volk_32f_x2_dot_prod_32f */
    float r = 0.0f;
    for (uint32_t i = 0; i < n; i++) r += a[i] * b[i];
    return r;
}

int32_t lo_fine_sync(lo_decoder *d, const lo_cf *samples, int32_t bin_idx, int32_t search_space) {
    const uint32_t sps = d->sps;
    const int32_t shift_ref = (bin_idx + 1) * (int32_t)d->decim;     /* :301 */
    float *ifreq = d->scratch_f;
    float max_correlation = 0.0f;
    int32_t lag = 0;
    lo_instantaneous_frequency(samples, ifreq, sps);                 /* :306 */
    for (int32_t i = -search_space + 1; i < search_space; i++) {     /* :308 */
        const int64_t start = (int64_t)shift_ref + i + (int64_t)sps; /* :310 */
        float c = 0.0f;
        for (uint32_t k = 0; k < sps; k++) {
            int64_t idx = start + k;
            if (idx < 0) idx = 0;
            if (idx >= (int64_t)3 * sps) idx = (int64_t)3 * sps - 1; /* D1 */
            c += ifreq[k] * d->upchirp_ifreq_v[idx];
        }
        if (c > max_correlation) { max_correlation = c; lag = i; }   /* :311-314 */
    }
    d->fine_sync = -lag;                                             /* :321 */
    return d->fine_sync;
}

/* ==================================================================================== */
/* A8  detect_preamble_autocorr, lib/decoder_impl.cc:340-366                             */
/* ==================================================================================== */
static void pwr_push(lo_decoder *d, float v) {                       /* circular_buffer::push_back */
    if (d->pwr_n < LO_PWR_QUEUE) {
        d->pwr_queue[(d->pwr_head + d->pwr_n) % LO_PWR_QUEUE] = v; d->pwr_n++;
    } else {
        d->pwr_queue[d->pwr_head] = v; d->pwr_head = (d->pwr_head + 1) % LO_PWR_QUEUE;
    }
}

float lo_detect_preamble_autocorr(lo_decoder *d, const lo_cf *samples) {
    const uint32_t window = d->sps;
    const lo_cf *c1 = samples, *c2 = samples + d->sps;               /* :341-342 */
    float dr = 0.0f, di = 0.0f, e1 = 0.0f, e2 = 0.0f;
    for (uint32_t i = 0; i < window; i++) {                          /* :350 a * conj(b) */
        dr += c1[i].re * c2[i].re + c1[i].im * c2[i].im;
        di += c1[i].im * c2[i].re - c1[i].re * c2[i].im;
    }
    for (uint32_t i = 0; i < window; i++) e1 += c1[i].re * c1[i].re + c1[i].im * c1[i].im;   /* :351,353 */
    for (uint32_t i = 0; i < window; i++) e2 += c2[i].re * c2[i].re + c2[i].im * c2[i].im;   /* :352,354 */
    d->energy_threshold = e2 / 2.0f;                                 /* :357 */
    pwr_push(d, e1 / d->sps);                                        /* :360 */
    const float s = sqrtf(e1 * e2);                                  /* :363 */
    return hypotf(dr / s, di / s);                                   /* |dot / (s + 0j)| */
}

/* A11 determine_energy :368-375, determine_snr :377-383 */
float lo_determine_energy(lo_decoder *d, const lo_cf *samples) {
    float e = 0.0f;
    for (uint32_t i = 0; i < d->sps; i++) e += samples[i].re * samples[i].re + samples[i].im * samples[i].im;
    return e;
}
static void determine_snr(lo_decoder *d) {
    if (d->pwr_n >= 2) {
        const float pwr_noise = d->pwr_queue[d->pwr_head];
        const float pwr_signal = d->pwr_queue[(d->pwr_head + d->pwr_n - 1) % LO_PWR_QUEUE];
        d->snr = pwr_signal / pwr_noise;
    }
}

/* ==================================================================================== */
/* A9  detect_upchirp / sliding_norm_cross_correlate_upchirp, :392-413                    */
/* ==================================================================================== */
float lo_detect_upchirp(lo_decoder *d, const lo_cf *samples, int32_t *index) {
    const uint32_t window = d->sps;
    float *ifreq = d->scratch_f;
    lo_instantaneous_frequency(samples, ifreq, window * 2);          /* :394 */
    float max_correlation = 0;
    for (uint32_t i = 0; i < window; i++) {                          /* :403-410 */
        const float c = dot_f(ifreq + i, d->upchirp_ifreq, window - 1u);
        if (c > max_correlation) { *index = (int32_t)i; max_correlation = c; }
    }
    return max_correlation;
}

/* ==================================================================================== */
/* A10 detect_downchirp / cross_correlate_ifreq / stddev, :385-390, :283-298, :415-425    */
/* ==================================================================================== */
static float stddev_f(const float *v, uint32_t len, float mean) {
    float variance = 0.0f;
    for (uint32_t i = 0u; i < len; i++) { const float t = v[i] - mean; variance += t * t; }
    variance /= (float)len;
    return sqrtf(variance);
}

float lo_detect_downchirp(lo_decoder *d, const lo_cf *samples) {
    const uint32_t window = d->sps, to_idx = window - 1u;
    float *ifreq = d->scratch_f;
    const float *ideal = d->downchirp_ifreq;
    lo_instantaneous_frequency(samples, ifreq, window);              /* :387 */
    float acc = 0.0f;
    for (uint32_t i = 0; i < to_idx; i++) acc += ifreq[i];           /* std::accumulate, :286 */
    const float average = acc / (float)to_idx;
    acc = 0.0f;
    for (uint32_t i = 0; i < to_idx; i++) acc += ideal[i];           /* :287 */
    const float chirp_avg = acc / (float)to_idx;
    const float sd = stddev_f(ifreq, to_idx, average) * stddev_f(ideal, to_idx, chirp_avg);   /* :288-289 */
    float result = 0.0f;
    for (uint32_t i = 0u; i < to_idx; i++) result += (ifreq[i] - average) * (ideal[i] - chirp_avg) / sd;   /* :291-293 */
    result /= (float)to_idx;                                         /* :295 */
    return result;
}

/* ==================================================================================== */
/* integer stage B1-B5                                                                   */
/* ==================================================================================== */
uint32_t lo_rotl(uint32_t bits, uint32_t count, uint32_t size) {     /* include/lora/utilities.h:96-103 */
    const uint32_t len_mask = (1u << size) - 1u;
    count %= size;
    bits &= len_mask;
    if (count == 0) return bits;                 /* (bits >> size) would be UB-free here anyway; same value */
    return ((bits << count) & len_mask) | (bits >> (size - count));
}

uint32_t lo_gray(uint32_t bin) { return bin ^ (bin >> 1u); }         /* :512 */

uint32_t lo_reduce_bin(uint32_t bin, uint32_t n_bins_hdr) {          /* :508 */
    return (uint32_t)(lroundf(bin / 4.0f) % (long)n_bins_hdr);
}

void lo_deinterleave_words(const uint32_t *words, uint32_t n_words, uint32_t ppm, uint8_t *out) {   /* :535-553 */
    const uint32_t offset_start = ppm - 1u;
    memset(out, 0, ppm);
    for (uint32_t i = 0u; i < n_words; i++) {
        const uint32_t word = lo_rotl(words[i], i, ppm);             /* :548 */
        for (uint32_t j = (1u << offset_start), x = offset_start; j; j >>= 1u, x--)   /* :550-552 */
            out[x] |= (uint8_t)((!!(word & j)) << i);
    }
}

uint8_t lo_deshuffle_byte(uint8_t v) {                               /* :568, :616-624 */
    static const uint8_t pattern[8] = {5, 0, 1, 2, 4, 3, 6, 7};
    uint8_t result = 0u;
    for (uint32_t j = 0u; j < 8u; j++) result |= (uint8_t)((!!(v & (1u << pattern[j]))) << j);
    return result;
}

static uint8_t bit8(uint8_t v, uint8_t i) { return (uint8_t)((v >> i) & 1u); }

uint8_t lo_hamming84_encode(uint8_t v) {                             /* hamming_encode_soft, utilities.h:257-264 */
    const uint8_t p1 = bit8(v, 1) ^ bit8(v, 2) ^ bit8(v, 3);
    const uint8_t p2 = bit8(v, 0) ^ bit8(v, 1) ^ bit8(v, 2);
    const uint8_t p3 = bit8(v, 0) ^ bit8(v, 1) ^ bit8(v, 3);
    const uint8_t p4 = bit8(v, 0) ^ bit8(v, 2) ^ bit8(v, 3);
    return (uint8_t)(p1 | (bit8(v, 0) << 1) | (bit8(v, 1) << 2) | (bit8(v, 2) << 3) | (p2 << 4) |
                     (bit8(v, 3) << 5) | (p3 << 6) | (p4 << 7));
}

uint8_t lo_hamming_decode_soft_byte(uint8_t v) {                     /* utilities.h:288-339 (deprecated there) */
    static const uint8_t H[16] = {0x0, 0x0, 0x4, 0x0, 0x6, 0x0, 0x0, 0x2, 0x7, 0x0, 0x0, 0x3, 0x0, 0x5, 0x1, 0x0};
    const uint8_t p1 = bit8(v, 0), p2 = bit8(v, 4), p3 = bit8(v, 6), p4 = bit8(v, 7);
    const uint8_t p1c = bit8(v, 2) ^ bit8(v, 3) ^ bit8(v, 5);
    const uint8_t p2c = bit8(v, 1) ^ bit8(v, 2) ^ bit8(v, 3);
    const uint8_t p3c = bit8(v, 1) ^ bit8(v, 2) ^ bit8(v, 5);
    const uint8_t p4c = bit8(v, 1) ^ bit8(v, 3) ^ bit8(v, 5);
    const uint8_t syndrome = (uint8_t)((p1 != p1c) | ((p2 != p2c) << 1) | ((p3 != p3c) << 2) | ((p4 != p4c) << 3));
    if (syndrome) v ^= (uint8_t)(1u << H[syndrome]);
    return (uint8_t)(bit8(v, 1) | (bit8(v, 2) << 1) | (bit8(v, 3) << 2) | (bit8(v, 5) << 3));
}

/* liquid-dsp fec_decode(LIQUID_FEC_HAMMING84) per-byte stand-in (decoder_impl.cc:661).
 * liquid decodes with a 256-entry table; the code book equals lo_hamming84_encode(0..15)
 * (SURVEY 8a B4).  Restated as nearest-codeword, lowest symbol wins ties: pinned for <=1 bit
 * error (unique nearest codeword, d_min = 4), UNPINNED for 2-bit errors. */
uint8_t lo_hamming84_decode(uint8_t cw) {
    int best = 0, best_d = 9;
    for (int s = 0; s < 16; s++) {
        const int dd = __builtin_popcount((unsigned)(cw ^ lo_hamming84_encode((uint8_t)s)));
        if (dd < best_d) { best_d = dd; best = s; }
    }
    return (uint8_t)best;
}

static uint32_t select_bits(uint32_t data, const uint8_t *idx, uint8_t n) {   /* utilities.h:209-216 */
    uint32_t r = 0u;
    for (uint8_t i = 0u; i < n; ++i) r |= (data & (1u << idx[i])) ? (1u << i) : 0u;
    return r;
}

/* decode(): deshuffle :611-637, dewhiten :639-652, hamming_decode :654-675, extract_data_only :693-706 */
size_t lo_decode_codewords(const uint8_t *dem, size_t n, int is_header, uint8_t cr,
                           uint8_t *out, size_t out_cap, size_t *consumed) {
    uint8_t *w = (uint8_t *)malloc(n + 16);
    size_t len = 0;
    const size_t to_decode = is_header ? 5u : n;                     /* :612 */
    for (size_t i = 0; i < to_decode; i++) w[len++] = lo_deshuffle_byte(i < n ? dem[i] : 0);
    if (is_header) w[len++] = 0;                                     /* :633 pad */
    if (consumed) *consumed = is_header ? (n < 5 ? n : 5) : n;       /* :632 / :635 */

    const uint8_t *prng; size_t prng_len;                            /* :579-580 */
    if (is_header) { prng = lo_prng_header; prng_len = LO_PRNG_HEADER_LEN; }
    else if (cr <= 2) { prng = lo_prng_payload_cr56; prng_len = LO_PRNG_PAYLOAD_CR56_LEN; }
    else { prng = lo_prng_payload_cr78; prng_len = LO_PRNG_PAYLOAD_CR78_LEN; }
    for (size_t i = 0; i < len; i++) w[i] ^= (i < prng_len ? prng[i] : 0);   /* :642-645, D2 */

    size_t n_out = 0;
    switch (cr) {                                                    /* :655 */
    case 4: case 3: {
        const uint32_t nn = (uint32_t)ceilf(len * 4.0f / (4.0f + cr));   /* :658 */
        for (uint32_t i = 0; i < nn; i++) {                          /* fec_decode, :661 */
            const uint8_t s0 = lo_hamming84_decode((2u * i) < len ? w[2u * i] : 0);          /* D3 */
            const uint8_t s1 = lo_hamming84_decode((2u * i + 1u) < len ? w[2u * i + 1u] : 0);
            uint8_t b = (uint8_t)((s0 << 4) | s1);
            if (!is_header) b = (uint8_t)(((b & 0x0f) << 4) | ((b & 0xf0) >> 4));            /* swap_nibbles :663 */
            if (n_out < out_cap) out[n_out] = b;
            n_out++;
        }
        break;
    }
    case 2: case 1: {
        static const uint8_t data_indices[4] = {1, 2, 3, 5};         /* :694 */
        for (size_t i = 0u; i < len; i += 2u) {                      /* :697-705 */
            const uint8_t d2 = (i + 1u < len) ? (uint8_t)(select_bits(w[i + 1u], data_indices, 4u) & 0xFF) : 0u;
            const uint8_t d1 = (uint8_t)(select_bits(w[i], data_indices, 4u) & 0xFF);
            const uint8_t b = is_header ? (uint8_t)((d1 << 4u) | d2) : (uint8_t)((d2 << 4u) | d1);
            if (n_out < out_cap) out[n_out] = b;
            n_out++;
        }
        break;
    }
    default: break;                                                  /* no case: d_decoded untouched */
    }
    free(w);
    return n_out < out_cap ? n_out : out_cap;
}

int32_t lo_payload_symbols(uint32_t payload_len, uint8_t cr, uint8_t sf, int reduced_rate) {   /* :842-847 */
    const uint8_t redundancy = (uint8_t)(reduced_rate ? 2 : 0);
    const int symbols_per_block = cr + 4u;
    const float bits_needed = (float)payload_len * 8.0f;
    const float symbols_needed = bits_needed * (symbols_per_block / 4.0f) / (float)(sf - redundancy);
    const int blocks_needed = (int)ceilf(symbols_needed / symbols_per_block);
    return blocks_needed * symbols_per_block;
}

/* ==================================================================================== */
/* A7  demodulate :493-529, decode :567-586, msg_lora_frame :588-609                      */
/* ==================================================================================== */
static void deinterleave(lo_decoder *d, uint32_t ppm) {              /* :535-565 */
    uint8_t out[32];
    lo_deinterleave_words(d->words.p, (uint32_t)d->words.n, ppm, out);
    for (uint32_t i = 0; i < ppm; i++) v8_push(&d->demodulated, out[i]);   /* :561 */
    d->words.n = 0;                                                  /* :564 */
}

static int demodulate(lo_decoder *d, const lo_cf *samples, int is_first) {
    const int reduced_rate = is_first || d->reduced_rate;            /* :495 */
    uint32_t bin_idx;
    if (d->demod_method == LO_DEMOD_FFT) {
        /* north-star variant of :499-500: FFT bin mapped onto the gradient index convention */
        bin_idx = (lo_get_shift_fft(d, samples, NULL) + d->n_bins - 1u) % d->n_bins;
    } else {
        bin_idx = lo_max_frequency_gradient_idx(d, samples);         /* :499 */
    }
    d->last_bin = (int32_t)bin_idx;
    if (d->enable_fine_sync) {                                       /* :501-502 */
        uint32_t s = d->decim / 4u; if (s < 2u) s = 2u;
        lo_fine_sync(d, samples, (int32_t)bin_idx, (int32_t)s);
    }
    if (reduced_rate) bin_idx = lo_reduce_bin(bin_idx, d->n_bins_hdr);   /* :507-509 */
    const uint32_t word = lo_gray(bin_idx);                          /* :512 */
    v32_push(&d->words, word);                                       /* :517 */
    if (d->words.n == (4u + (is_first ? 4u : hdr_cr(d)))) {          /* :521 */
        deinterleave(d, reduced_rate ? d->sf - 2u : d->sf);          /* :523 */
        return 1;
    }
    return 0;
}

static void decode(lo_decoder *d, int is_header) {                   /* :567-586 */
    uint8_t out[1024];
    size_t consumed = 0;
    const size_t n = lo_decode_codewords(d->demodulated.p, d->demodulated.n, is_header, hdr_cr(d), out, sizeof out, &consumed);
    if (is_header) v8_erase_front(&d->demodulated, consumed); else d->demodulated.n = 0;
    const uint8_t cr = hdr_cr(d);
    if (cr >= 1 && cr <= 4) {
        if (cr >= 3) d->decoded.n = 0;                               /* d_decoded.assign, :664 */
        for (size_t i = 0; i < n; i++) v8_push(&d->decoded, out[i]); /* push_back for cr 1,2 :702-704 */
    }
}

static void msg_lora_frame(lo_decoder *d) {                          /* :588-609 */
    const uint32_t len = 15u + 3u + d->payload_length;
    uint8_t *buf = (uint8_t *)calloc(1, len ? len : 1);
    /* loratap_header_t (include/lora/loratap.h:48-55) zeroed except rssi.snr at byte 13 (:597) */
    const double snr_db = 10.0f * log10f(d->snr) + 0.5;
    buf[13] = (uint8_t)(int32_t)snr_db;                              /* D6 */
    memcpy(buf + 15, d->phdr, 3);                                    /* :600 */
    for (uint32_t i = 0; i < d->payload_length; i++)                 /* :601, D5 */
        buf[18 + i] = i < d->decoded.n ? d->decoded.p[i] : 0;
    if (d->n_frames == d->cap_frames) {
        d->cap_frames = d->cap_frames ? d->cap_frames * 2 : 8;
        d->frames = (lo_frame *)realloc(d->frames, d->cap_frames * sizeof(lo_frame));
    }
    d->frames[d->n_frames].data = buf;
    d->frames[d->n_frames].len = len;
    d->n_frames++;
}

size_t lo_frame_count(const lo_decoder *d) { return d->n_frames; }
size_t lo_frame_len(const lo_decoder *d, size_t i) { return i < d->n_frames ? d->frames[i].len : 0; }
const uint8_t *lo_frame_data(const lo_decoder *d, size_t i) { return i < d->n_frames ? d->frames[i].data : NULL; }
void lo_frames_clear(lo_decoder *d) {
    for (size_t i = 0; i < d->n_frames; i++) free(d->frames[i].data);
    d->n_frames = 0;
}

/* ==================================================================================== */
/* A12 work(), lib/decoder_impl.cc:740-903                                               */
/* ==================================================================================== */
int lo_work(lo_decoder *d, const lo_cf *input, lo_step *tr) {
    int consumed = 0;
    lo_step t; t.state = d->state; t.bin = -1; t.metric = 0.0f; t.consumed = 0; t.fine_sync = 0;
    d->fine_sync = 0;                                                /* :749 */
    d->last_bin = -1;
    switch (d->state) {
    case LO_DETECT: {                                                /* :752-768 */
        const float correlation = lo_detect_preamble_autocorr(d, input);
        t.metric = correlation;
        if (correlation >= 0.90f) {
            determine_snr(d);
            d->corr_fails = 0u;
            d->state = LO_SYNC;
            break;
        }
        consumed = (int)d->sps;
        break;
    }
    case LO_SYNC: {                                                  /* :770-783 */
        int32_t i = 0;
        t.metric = lo_detect_upchirp(d, input, &i);
        consumed = i;
        d->state = LO_FIND_SFD;
        break;
    }
    case LO_FIND_SFD: {                                              /* :785-818 */
        const float c = lo_detect_downchirp(d, input);
        t.metric = c;
        if (c > 0.96f) {
            d->state = LO_PAUSE;
        } else {
            if (c < -0.97f) lo_fine_sync(d, input, -1, (int32_t)d->decim * 4);   /* :803 */
            else d->corr_fails++;
            if (d->corr_fails > 4u) d->state = LO_DETECT;            /* :808-813 */
        }
        consumed = (int32_t)d->sps + d->fine_sync;                   /* :816 */
        break;
    }
    case LO_PAUSE: {                                                 /* :820-824 */
        d->state = LO_DECODE_HEADER;
        consumed = (int)(d->sps + d->delay_after_sync);
        break;
    }
    case LO_DECODE_HEADER: {                                         /* :826-858 */
        if (demodulate(d, input, 1)) {
            if (d->implicit) {
                d->payload_symbols = 1;                              /* :829 */
            } else {
                decode(d, 1);                                        /* :831 */
                out_hex(d, d->decoded.p, d->decoded.n, (uint32_t)d->decoded.n, 0, 0);   /* :832 */
                for (int k = 0; k < 3; k++) d->phdr[k] = k < (int)d->decoded.n ? d->decoded.p[k] : 0;   /* :833 */
                if (hdr_cr(d) > 4) hdr_set_cr(d, 4);                 /* :834-835 */
                d->decoded.n = 0;                                    /* :836 */
                d->payload_length = d->phdr[0] + LO_MAC_CRC_SIZE * hdr_has_crc(d);   /* :838 */
                d->payload_symbols = lo_payload_symbols(d->payload_length, hdr_cr(d), d->sf, d->reduced_rate);
            }
            d->state = LO_DECODE_PAYLOAD;                            /* :853 */
        }
        t.bin = d->last_bin;
        consumed = (int32_t)d->sps + d->fine_sync;                   /* :856 */
        break;
    }
    case LO_DECODE_PAYLOAD: {                                        /* :860-886 */
        if (d->implicit && lo_determine_energy(d, input) < d->energy_threshold) {
            d->payload_symbols = 0;
            d->payload_length = (uint32_t)(int32_t)(d->demodulated.n / 2);   /* :864 */
        } else if (demodulate(d, input, 0)) {
            if (!d->implicit) d->payload_symbols -= (int32_t)(4u + hdr_cr(d));   /* :866-867 */
        }
        t.bin = d->last_bin;
        if (d->payload_symbols <= 0) {                               /* :870 */
            decode(d, 0);
            out_hex(d, d->decoded.p, d->decoded.n, d->payload_length, 1, 1);     /* :872 */
            msg_lora_frame(d);
            d->state = LO_DETECT;
            d->decoded.n = 0; d->words.n = 0; d->demodulated.n = 0;  /* :876-880 */
        }
        consumed = (int32_t)d->sps + d->fine_sync;                   /* :883 */
        break;
    }
    case LO_STOP: consumed = (int)d->sps; break;                     /* :888-891 */
    default: break;
    }
    t.consumed = consumed;
    t.fine_sync = d->fine_sync;
    if (tr) *tr = t;
    return consumed;
}

size_t lo_run(lo_decoder *d, const lo_cf *in, size_t n_items, lo_step *steps, size_t max_steps, size_t *n_steps) {
    size_t pos = 0, k = 0;
    const size_t need = 2u * (size_t)d->sps;                         /* set_output_multiple, :91 */
    while (pos + need <= n_items) {
        lo_step t;
        const int c = lo_work(d, in + pos, &t);
        if (steps && k < max_steps) steps[k] = t;
        k++;
        pos += (size_t)(c < 0 ? 0 : c);
    }
    if (n_steps) *n_steps = k;
    return pos;
}

/* ==================================================================================== */
/* batch helpers (bench cpu_baseline / parity at scale)                                  */
/* ==================================================================================== */
void lo_demod_fft_batch(lo_decoder *d, const lo_cf *iq, size_t n_symbols, uint32_t *bins, float *mags) {
    for (size_t s = 0; s < n_symbols; s++) {
        float m;
        bins[s] = lo_get_shift_fft(d, iq + s * (size_t)d->sps, &m);
        if (mags) mags[s] = m;
    }
}

void lo_demod_grad_batch(lo_decoder *d, const lo_cf *iq, size_t n_symbols, uint32_t *bins) {
    for (size_t s = 0; s < n_symbols; s++) bins[s] = lo_max_frequency_gradient_idx(d, iq + s * (size_t)d->sps);
}
