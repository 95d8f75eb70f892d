// decoder_impl: the reference's private members (lib/decoder_impl.h:70-123) collapse into one handle.
#pragma once
#include <lora/decoder.h>
#include <lora_b200.h>

namespace gr {
namespace lora {

class decoder_impl : public decoder {
public:
    decoder_impl(float samp_rate, uint32_t bandwidth, uint8_t sf, bool implicit, uint8_t cr, bool crc, bool reduced_rate,
                 bool disable_drift_correction);
    ~decoder_impl() override;
    int work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items) override;
    void set_sf(uint8_t sf) override;
    void set_samp_rate(float samp_rate) override;
    lora_b200_decoder *handle() { return d_gpu; }

private:
    lora_b200_decoder *d_gpu;
};

}  // namespace lora
}  // namespace gr
