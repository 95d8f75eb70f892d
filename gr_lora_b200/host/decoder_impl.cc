// decoder_impl.cc -- the drop-in body of gr::lora::decoder_impl over liblora_b200.so.
// This is the file INTEGRATION.md describes: it keeps decoder::make's signature, the sync_block shape
// (lib/decoder_impl.cc:50-52), set_output_multiple(2*sps) (:91), the banner and hex lines on std::cout
// (:93-103,:832,:872), the "frames"/"control" message ports (:120-121) and exit(1) for a bad SF
// (:57-61).  Compiled here against gr_stub/ (GNU Radio is absent); against the real GNU Radio the same
// source builds with the usual gr-lora CMake, linking lora_b200 instead of liquid.
#include <gnuradio/io_signature.h>
#include <cstdlib>
#include <iostream>
#include "decoder_impl.h"

namespace gr {
namespace lora {

decoder::sptr decoder::make(float samp_rate, uint32_t bandwidth, uint8_t sf, bool implicit, uint8_t cr, bool crc,
                            bool reduced_rate, bool disable_drift_correction) {
    return gnuradio::get_initial_sptr(new decoder_impl(samp_rate, bandwidth, sf, implicit, cr, crc, reduced_rate,
                                                       disable_drift_correction));
}

decoder_impl::decoder_impl(float samp_rate, uint32_t bandwidth, uint8_t sf, bool implicit, uint8_t cr, bool crc,
                           bool reduced_rate, bool disable_drift_correction)
    : gr::sync_block("decoder", gr::io_signature::make(1, -1, sizeof(gr_complex)), gr::io_signature::make(0, 0, 0)) {
    lora_b200_config cfg = {};
    cfg.samp_rate = samp_rate;
    cfg.bandwidth = bandwidth;
    cfg.sf = sf;
    cfg.implicit = implicit;
    cfg.cr = cr;
    cfg.crc = crc;
    cfg.reduced_rate = reduced_rate;
    cfg.disable_drift_correction = disable_drift_correction;
    const char *demod = std::getenv("LORA_B200_DEMOD");           // "fft" selects the north-star demodulator
    cfg.demod = (demod && std::string(demod) == "fft") ? LORA_B200_DEMOD_FFT : LORA_B200_DEMOD_GRADIENT;
    cfg.n_streams = 1;
    cfg.device = -1;
    cfg.max_items_per_call = 1u << 20;
    d_gpu = lora_b200_create(&cfg);
    if (!d_gpu) {
        std::cerr << lora_b200_last_error() << std::endl;
        exit(1);
    }
    char banner[512];
    lora_b200_banner(d_gpu, banner, sizeof banner);
    std::cout << banner;
    set_output_multiple(2 * (int)lora_b200_samples_per_symbol(d_gpu));
    set_min_noutput_items(64 * (int)lora_b200_samples_per_symbol(d_gpu));
    message_port_register_out(pmt::mp("frames"));
    message_port_register_out(pmt::mp("control"));
}

decoder_impl::~decoder_impl() { lora_b200_destroy(d_gpu); }

static void on_frame(void *user, uint32_t, const uint8_t *frame, size_t len) {
    static_cast<decoder_impl *>(user)->message_port_pub(pmt::mp("frames"), pmt::make_blob(frame, len));
}

int decoder_impl::work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &) {
    size_t consumed = 0;
    if (lora_b200_work(d_gpu, 0, input_items[0], (size_t)noutput_items, &consumed, on_frame, this) != 0) {
        std::cerr << lora_b200_last_error() << std::endl;
        return WORK_DONE;
    }
    char lines[1 << 14];
    if (lora_b200_stdout_last(d_gpu, 0, lines, sizeof lines) > 0) std::cout << lines << std::flush;
    consume_each((int)consumed);
    return 0;
}

void decoder_impl::set_sf(uint8_t sf) {
    lora_b200_set_sf(d_gpu, sf);
    std::cerr << lora_b200_last_error() << std::endl;
}

void decoder_impl::set_samp_rate(float samp_rate) {
    lora_b200_set_samp_rate(d_gpu, samp_rate);
    std::cerr << lora_b200_last_error() << std::endl;
}

}  // namespace lora
}  // namespace gr
