// ref_wrap.cc -- C API around the REFERENCE'S OWN decoder (oracle/_ref/liblora_ref.so).  TEST INFRASTRUCTURE ONLY.
//
// This translation unit #includes the reference's unmodified lib/decoder_impl.cc from where it lies under
// /root/reference (the Makefile passes -I/root/reference/lib -I/root/reference/include; nothing is copied into this
// repository) and compiles it against the stand-in headers of oracle/ref_standins/ (GNU Radio, VOLK, liquid-dsp and
// Boost are not installed here; see oracle/ref_standins/README.md for what that leaves unpinned).  `private` is
// redefined to `public` for the reference's class so that every stage function of SURVEY.md 8(a) -- not only work() --
// can be driven from tests; standard headers are included first so the redefinition touches the reference's code
// only.  Access specifiers do not change layout or code generation with GCC.
//
// It is the pin of oracle/lora_oracle.c: tests/test_ref_pins_oracle.py requires restatement == reference for the
// chirp tables, instantaneous frequency, both demodulators, fine sync, the three detectors, the integer chain and
// the whole work() state machine (per-step state, consume amount, fine sync; frames; stdout) on the 13 golden frame
// cases.  It exists only in the build container (the GPU box has no /root/reference; the built .so travels there).
//
// Members the reference never initialises before first use -- d_snr when the power ring holds < 2 entries
// (lib/decoder_impl.cc:377-383,597), d_corr_fails, d_payload_length, d_mac_crc, and in implicit-header mode the length /
// crc nibbles of d_phdr that go out in every frame (:72-73,600) -- would be heap garbage: the object is therefore
// constructed in zeroed storage and d_snr set to the 1.0f the restatement documents (D4 in oracle/lora_oracle.h), so
// that runs are reproducible.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <numeric>
#include <sstream>
#include <string>
#include <vector>

#include <gnuradio/sync_block.h>
#include <gnuradio/expj.h>
#include <volk/volk.h>
#include <liquid/liquid.h>
#include <boost/circular_buffer.hpp>

#define private public
#define protected public
#include "decoder_impl.cc"          // /root/reference/lib/decoder_impl.cc, unmodified
#undef private
#undef protected

using gr::lora::decoder_impl;
using gr::lora::DecoderState;

namespace {

// Everything the reference writes to std::cout goes to the log of the decoder that is running on the calling thread.
// std::cout's buffer is replaced ONCE by a stream buffer that appends to a thread-local target (bench.py runs one
// reference decoder per host thread; swapping rdbuf per call would race), and forwards to the original buffer when no
// decoder call is active on the thread.
thread_local std::string *tls_cout_target = nullptr;

class TlsCoutBuf : public std::streambuf {
public:
    explicit TlsCoutBuf(std::streambuf *orig) : d_orig(orig) {}
protected:
    int overflow(int c) override {
        if (c == traits_type::eof()) return c;
        if (tls_cout_target) { tls_cout_target->push_back((char)c); return c; }
        return d_orig ? d_orig->sputc((char)c) : c;
    }
    std::streamsize xsputn(const char *s, std::streamsize n) override {
        if (tls_cout_target) { tls_cout_target->append(s, (size_t)n); return n; }
        return d_orig ? d_orig->sputn(s, n) : n;
    }
    int sync() override { return d_orig && !tls_cout_target ? d_orig->pubsync() : 0; }
private:
    std::streambuf *d_orig;
};

void install_cout_buffer() {
    static TlsCoutBuf *buf = [] {
        TlsCoutBuf *b = new TlsCoutBuf(std::cout.rdbuf());
        std::cout.rdbuf(b);
        return b;
    }();
    (void)buf;
}

struct CoutCapture {
    std::string *prev;
    explicit CoutCapture(std::string &to) : prev(tls_cout_target) {
        install_cout_buffer();
        tls_cout_target = &to;
        // print_vector_hex leaves std::hex / setfill('0') set on std::cout (include/lora/utilities.h:356); in the
        // reference that is process-wide state (a second decoder's banner would print "Bins per symbol: 80").  Each
        // captured call starts from default formatting so that one decoder's output does not depend on another's.
        std::cout.flags(std::ios::dec | std::ios::skipws);
        std::cout.fill(' ');
        std::cout.precision(6);
        std::cout.width(0);
    }
    ~CoutCapture() { tls_cout_target = prev; }
};

}  // namespace

struct lr_decoder {
    std::string out;
    void *storage = nullptr;
    decoder_impl *impl = nullptr;
    std::vector<std::vector<uint8_t>> frames;
};

typedef struct { float re, im; } lr_cf;
typedef struct { int32_t state, consumed, bin, fine_sync; float metric; } lr_step;

static inline const gr_complex *cx(const lr_cf *p) { return reinterpret_cast<const gr_complex *>(p); }

static void drain_frames(lr_decoder *d) {
    auto &port = d->impl->standin_ports["frames"];
    for (auto &m : port) d->frames.push_back(m->blob);
    port.clear();
}

extern "C" {

lr_decoder *lr_create(float samp_rate, uint32_t bandwidth, uint8_t sf, int implicit, uint8_t cr, int crc, int reduced_rate,
                      int disable_drift_correction) {
    if (sf < 6 || sf > 13) return nullptr;             // the reference prints an error and exit(1)s (:57-61)
    lr_decoder *d = new lr_decoder;
    CoutCapture cap(d->out);
    d->storage = calloc(1, sizeof(decoder_impl));
    d->impl = new (d->storage) decoder_impl(samp_rate, bandwidth, sf, implicit != 0, cr, crc != 0, reduced_rate != 0, disable_drift_correction != 0);
    d->impl->d_snr = 1.0f;
    return d;
}
void lr_destroy(lr_decoder *d) {
    if (!d) return;
    d->impl->~decoder_impl();
    free(d->storage);
    delete d;
}

uint32_t lr_sps(const lr_decoder *d) { return d->impl->d_samples_per_symbol; }
uint32_t lr_bins(const lr_decoder *d) { return d->impl->d_number_of_bins; }
uint32_t lr_bins_hdr(const lr_decoder *d) { return d->impl->d_number_of_bins_hdr; }
uint32_t lr_decim(const lr_decoder *d) { return d->impl->d_decim_factor; }
uint32_t lr_delay_after_sync(const lr_decoder *d) { return d->impl->d_delay_after_sync; }
int lr_output_multiple(const lr_decoder *d) { return d->impl->standin_output_multiple; }
double lr_bits_per_symbol(const lr_decoder *d) { return d->impl->d_bits_per_symbol; }
double lr_dt(const lr_decoder *d) { return d->impl->d_dt; }

const lr_cf *lr_downchirp(const lr_decoder *d) { return reinterpret_cast<const lr_cf *>(d->impl->d_downchirp.data()); }
const lr_cf *lr_upchirp(const lr_decoder *d) { return reinterpret_cast<const lr_cf *>(d->impl->d_upchirp.data()); }
const float *lr_downchirp_ifreq(const lr_decoder *d) { return d->impl->d_downchirp_ifreq.data(); }
const float *lr_upchirp_ifreq(const lr_decoder *d) { return d->impl->d_upchirp_ifreq.data(); }
const float *lr_upchirp_ifreq_v(const lr_decoder *d) { return d->impl->d_upchirp_ifreq_v.data(); }

void lr_instantaneous_frequency(lr_decoder *d, const lr_cf *in, float *out, uint32_t window) {
    d->impl->instantaneous_frequency(cx(in), out, window);
}
uint32_t lr_get_shift_fft(lr_decoder *d, const lr_cf *samples, float *mag_out) {       // :430-464
    const uint32_t b = d->impl->get_shift_fft(cx(samples));
    // fft_execute(d_qr) "for debugging" (:459) leaves d_tmp untouched (it writes d_mult_hf): the kept bins are still there
    if (mag_out) *mag_out = std::abs(d->impl->d_tmp[b]);
    return b;
}
void lr_get_shift_fft_spectrum(lr_decoder *d, const lr_cf *samples, lr_cf *bins_out) {  // the N kept bins of :447-450
    d->impl->get_shift_fft(cx(samples));
    memcpy(bins_out, d->impl->d_tmp.data(), sizeof(lr_cf) * d->impl->d_number_of_bins);
}
uint32_t lr_max_frequency_gradient_idx(lr_decoder *d, const lr_cf *samples) { return d->impl->max_frequency_gradient_idx(cx(samples)); }
int32_t lr_fine_sync(lr_decoder *d, const lr_cf *samples, int32_t bin_idx, int32_t search_space) {
    d->impl->fine_sync(cx(samples), bin_idx, search_space);
    return d->impl->d_fine_sync;
}
float lr_detect_preamble_autocorr(lr_decoder *d, const lr_cf *samples) {
    return d->impl->detect_preamble_autocorr(cx(samples), d->impl->d_samples_per_symbol);
}
float lr_energy_threshold(const lr_decoder *d) { return d->impl->d_energy_threshold; }
float lr_detect_upchirp(lr_decoder *d, const lr_cf *samples, int32_t *index) {
    int32_t i = 0;
    const float c = d->impl->detect_upchirp(cx(samples), d->impl->d_samples_per_symbol, &i);
    if (index) *index = i;
    return c;
}
float lr_detect_downchirp(lr_decoder *d, const lr_cf *samples) { return d->impl->detect_downchirp(cx(samples), d->impl->d_samples_per_symbol); }
float lr_experimental_determine_cfo(lr_decoder *d, const lr_cf *samples) {       // :730-738 (its call site :774 is commented out)
    return d->impl->experimental_determine_cfo(cx(samples), d->impl->d_samples_per_symbol);
}
float lr_determine_energy(lr_decoder *d, const lr_cf *samples) { return d->impl->determine_energy(cx(samples)); }

void lr_demod_fft_batch(lr_decoder *d, const lr_cf *iq, size_t n_symbols, uint32_t *bins, float *mags) {
    const size_t sps = d->impl->d_samples_per_symbol;
    for (size_t s = 0; s < n_symbols; s++) bins[s] = lr_get_shift_fft(d, iq + s * sps, mags ? mags + s : nullptr);
}
void lr_demod_grad_batch(lr_decoder *d, const lr_cf *iq, size_t n_symbols, uint32_t *bins) {
    const size_t sps = d->impl->d_samples_per_symbol;
    for (size_t s = 0; s < n_symbols; s++) bins[s] = d->impl->max_frequency_gradient_idx(cx(iq + s * sps));
}

int lr_state(const lr_decoder *d) { return (int)d->impl->d_state; }

// one work() call (:740-903).  `in` must hold at least 2*sps items (the block's output multiple, :91).
int lr_work(lr_decoder *d, const lr_cf *in, lr_step *trace) {
    decoder_impl *p = d->impl;
    const gr_complex *x = cx(in);
    if (trace) {
        trace->state = (int32_t)p->d_state;
        trace->bin = -1;
        trace->metric = 0.0f;
        // side-effect-free previews of what this call will look at (the reference keeps them in locals)
        if (p->d_state == DecoderState::DECODE_HEADER ||
            (p->d_state == DecoderState::DECODE_PAYLOAD && !(p->d_implicit && p->determine_energy(x) < p->d_energy_threshold)))
            trace->bin = (int32_t)p->max_frequency_gradient_idx(x);
        if (p->d_state == DecoderState::FIND_SFD) trace->metric = p->detect_downchirp(x, p->d_samples_per_symbol);
        if (p->d_state == DecoderState::SYNC) { int32_t i = 0; trace->metric = p->detect_upchirp(x, p->d_samples_per_symbol, &i); }
        if (p->d_state == DecoderState::DETECT) {
            const float thr = p->d_energy_threshold;
            const boost::circular_buffer<float> q = p->d_pwr_queue;
            trace->metric = p->detect_preamble_autocorr(x, p->d_samples_per_symbol);
            p->d_energy_threshold = thr;
            p->d_pwr_queue = q;
        }
    }
    gr_vector_const_void_star ins(1, in);
    gr_vector_void_star outs;
    p->standin_consumed = 0;
    {
        CoutCapture cap(d->out);
        p->work(2 * (int)p->d_samples_per_symbol, ins, outs);
    }
    drain_frames(d);
    if (trace) {
        trace->consumed = (int32_t)p->standin_consumed;
        trace->fine_sync = p->d_fine_sync;
    }
    return (int)p->standin_consumed;
}

// fake scheduler, same contract as lo_run: work() is called while at least 2*sps unconsumed items remain
size_t lr_run(lr_decoder *d, const lr_cf *in, size_t n_items, lr_step *steps, size_t max_steps, size_t *n_steps) {
    const size_t need = 2 * (size_t)d->impl->d_samples_per_symbol;
    size_t pos = 0, n = 0;
    while (n_items - pos >= need) {
        lr_step st;
        const int c = lr_work(d, in + pos, &st);
        if (steps && n < max_steps) steps[n] = st;
        n++;
        if (c < 0) break;
        pos += (size_t)c;
    }
    if (n_steps) *n_steps = n;
    return pos;
}

size_t lr_frame_count(const lr_decoder *d) { return d->frames.size(); }
size_t lr_frame_len(const lr_decoder *d, size_t i) { return d->frames[i].size(); }
const uint8_t *lr_frame_data(const lr_decoder *d, size_t i) { return d->frames[i].data(); }
void lr_frames_clear(lr_decoder *d) { d->frames.clear(); }
const char *lr_stdout(lr_decoder *d) { return d->out.c_str(); }

// ---- integer stage, driven through the reference's member functions -------------------------------------------
uint32_t lr_rotl(uint32_t bits, uint32_t count, uint32_t size) { return gr::lora::rotl(bits, count, size); }
uint8_t lr_hamming_encode_soft(uint8_t nibble) { return gr::lora::hamming_encode_soft(nibble); }
uint8_t lr_hamming_decode_soft_byte(uint8_t v) { return gr::lora::hamming_decode_soft_byte(v); }

// B1: d_words -> deinterleave(ppm) -> the ppm code words appended to d_demodulated (:535-565)
void lr_deinterleave_words(lr_decoder *d, const uint32_t *words, uint32_t n_words, uint32_t ppm, uint8_t *out) {
    decoder_impl *p = d->impl;
    p->d_words.assign(words, words + n_words);
    p->d_demodulated.clear();
    p->deinterleave(ppm);
    for (uint32_t i = 0; i < ppm; i++) out[i] = p->d_demodulated[i];
    p->d_demodulated.clear();
}

// B2-B4: decode(is_header) (:567-586) on a code-word vector with d_phdr.cr = cr.  Returns d_decoded.size(); *consumed =
// code words removed from d_demodulated (5 for a header, all for a payload).
size_t lr_decode_codewords(lr_decoder *d, const uint8_t *demodulated, size_t n, int is_header, uint8_t cr, uint8_t *out,
                           size_t out_cap, size_t *consumed) {
    decoder_impl *p = d->impl;
    const uint8_t cr_saved = p->d_phdr.cr;
    p->d_phdr.cr = cr;
    p->d_demodulated.assign(demodulated, demodulated + n);
    // fec_decode reads 2*ceil(len*4/(4+cr)) bytes of d_words_dewhitened, sometimes past its size (:658-661, D3 in
    // oracle/lora_oracle.h); reserve zeroed room so the over-read is defined and equal to the restatement's
    p->d_words_dewhitened.clear();
    p->d_words_dewhitened.reserve(2 * n + 16);
    memset(p->d_words_dewhitened.data(), 0, p->d_words_dewhitened.capacity());
    p->d_decoded.clear();
    p->decode(is_header != 0);
    if (consumed) *consumed = n - p->d_demodulated.size();
    const size_t m = std::min(out_cap, p->d_decoded.size());
    memcpy(out, p->d_decoded.data(), m);
    const size_t total = p->d_decoded.size();
    p->d_decoded.clear();
    p->d_demodulated.clear();
    p->d_phdr.cr = cr_saved;
    return total;
}

}  // extern "C"
