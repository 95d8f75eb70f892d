// channelizer_impl: the reference's hier block (lib/channelizer_impl.h, lib/channelizer_impl.cc:40-71) with the GNU Radio
// freq_xlating_fir_filter_ccf inside replaced by one block whose work() runs the FIR bank on the GPU.
#pragma once
#include <lora/channelizer.h>
#include <lora_b200.h>

namespace gr {
namespace lora {

// the inner block: decimating sync block, 1 input, one output per channel of channel_list
class xlating_fir_b200 : public gr::sync_block {
public:
    xlating_fir_b200(lora_b200_channelizer *c, uint32_t n_channels, uint32_t decimation);
    // noutput_items outputs per channel from noutput_items * decimation inputs (gr::sync_decimator's contract)
    int work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items) override;
    uint32_t decimation() const { return d_decimation; }

private:
    lora_b200_channelizer *d_c;
    uint32_t d_n_channels, d_decimation;
};

class channelizer_impl : public channelizer {
public:
    channelizer_impl(float samp_rate, float center_freq, std::vector<float> channel_list, uint32_t bandwidth, uint32_t decimation);
    ~channelizer_impl() override;
    void apply_cfo(float cfo);                                   // lib/channelizer_impl.cc:68-71
    std::shared_ptr<xlating_fir_b200> filter() { return d_xlating_fir_filter; }
    lora_b200_channelizer *handle() { return d_gpu; }

private:
    lora_b200_channelizer *d_gpu;
    std::shared_ptr<xlating_fir_b200> d_xlating_fir_filter;
};

}  // namespace lora
}  // namespace gr
