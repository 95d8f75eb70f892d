// Public block class: same name, namespace and make() signature as the reference's include/lora/channelizer.h:36-50
// (written from its documented interface, not copied).
#pragma once
#include <gnuradio/hier_block2.h>
#include <cstdint>
#include <memory>
#include <vector>

namespace gr {
namespace lora {

class channelizer : virtual public gr::hier_block2 {
public:
    typedef std::shared_ptr<channelizer> sptr;
    static sptr make(float samp_rate, float center_freq, std::vector<float> channel_list, uint32_t bandwidth, uint32_t decimation);
};

}  // namespace lora
}  // namespace gr
