// channelizer_impl.cc -- the drop-in body of gr::lora::channelizer_impl over liblora_b200.so.
// Keeps channelizer::make's signature (include/lora/channelizer.h:49), the hier-block shape and wiring of
// lib/channelizer_impl.cc:40-60 (1 input, channel_list.size() outputs, "control" message input) and apply_cfo (:68-71).
// GNU Radio's freq_xlating_fir_filter_ccf(decimation, firdes::low_pass(...), channel_list[0] - center_freq, samp_rate) is
// replaced by xlating_fir_b200, whose work() hands the scheduler's buffer to the GPU FIR bank; unlike the reference, which
// wires channel_list[0] only, every listed channel gets its output.  Compiled here against gr_stub/.
#include <gnuradio/io_signature.h>
#include <cstdlib>
#include <iostream>
#include "channelizer_impl.h"

namespace gr {
namespace lora {

xlating_fir_b200::xlating_fir_b200(lora_b200_channelizer *c, uint32_t n_channels, uint32_t decimation)
    : gr::sync_block("xlating_fir_b200", gr::io_signature::make(1, 1, sizeof(gr_complex)),
                     gr::io_signature::make((int)n_channels, (int)n_channels, sizeof(gr_complex))),
      d_c(c), d_n_channels(n_channels), d_decimation(decimation) {}

int xlating_fir_b200::work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items) {
    size_t n_out = 0;
    if (lora_b200_channelizer_work_host(d_c, input_items[0], (size_t)noutput_items * d_decimation, &n_out) != 0) {
        std::cerr << lora_b200_channelizer_last_error() << std::endl;
        return WORK_DONE;
    }
    for (uint32_t ch = 0; ch < d_n_channels && ch < output_items.size(); ch++)
        if (lora_b200_channelizer_read_output(d_c, ch, output_items[ch], n_out) != 0) {
            std::cerr << lora_b200_channelizer_last_error() << std::endl;
            return WORK_DONE;
        }
    return (int)n_out;
}

channelizer::sptr channelizer::make(float samp_rate, float center_freq, std::vector<float> channel_list, uint32_t bandwidth,
                                    uint32_t decimation) {
    return gnuradio::get_initial_sptr(new channelizer_impl(samp_rate, center_freq, channel_list, bandwidth, decimation));
}

channelizer_impl::channelizer_impl(float samp_rate, float center_freq, std::vector<float> channel_list, uint32_t bandwidth,
                                   uint32_t decimation)
    : gr::hier_block2("channelizer", gr::io_signature::make(1, 1, sizeof(gr_complex)),
                      gr::io_signature::make((int)channel_list.size(), (int)channel_list.size(), sizeof(gr_complex))) {
    d_gpu = lora_b200_channelizer_create(samp_rate, center_freq, channel_list.data(), (uint32_t)channel_list.size(), bandwidth,
                                         decimation, -1);
    if (!d_gpu) {
        std::cerr << lora_b200_channelizer_last_error() << std::endl;
        exit(1);
    }
    d_xlating_fir_filter = std::make_shared<xlating_fir_b200>(d_gpu, (uint32_t)channel_list.size(), decimation);
    message_port_register_hier_in(pmt::intern("control"));          // :53 (the controller block that turns ("cfo" . x) into
                                                                    // apply_cfo(x), lib/controller_impl.cc:52-57, is host plumbing)
    connect(self(), 0, d_xlating_fir_filter, 0);                     // :55
    connect(d_xlating_fir_filter, 0, self(), 0);                     // :56
}

channelizer_impl::~channelizer_impl() { lora_b200_channelizer_destroy(d_gpu); }

void channelizer_impl::apply_cfo(float cfo) { lora_b200_channelizer_apply_cfo(d_gpu, 0, cfo); }

}  // namespace lora
}  // namespace gr
