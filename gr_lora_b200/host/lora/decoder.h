// Public block class: same name, namespace and make() signature as the reference's
// include/lora/decoder.h:693-709 (written from its documented interface, not copied).
#pragma once
#include <gnuradio/sync_block.h>
#include <cstdint>
#include <memory>

namespace gr {
namespace lora {

class decoder : virtual public gr::sync_block {
public:
    typedef std::shared_ptr<decoder> sptr;
    static sptr make(float samp_rate, uint32_t bandwidth, uint8_t sf, bool implicit, uint8_t cr, bool crc,
                     bool reduced_rate, bool disable_drift_correction);
    virtual void set_sf(uint8_t sf) = 0;
    virtual void set_samp_rate(float samp_rate) = 0;

protected:
    decoder() : gr::sync_block("decoder", nullptr, nullptr) {}
};

}  // namespace lora
}  // namespace gr
