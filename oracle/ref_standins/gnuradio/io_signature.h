// stand-in for <gnuradio/io_signature.h> (lib/decoder_impl.cc:26,50-52)
#pragma once
#include <memory>
namespace gr {
class io_signature {
public:
    typedef std::shared_ptr<io_signature> sptr;
    static sptr make(int min_streams, int max_streams, int sizeof_stream_item) {
        return sptr(new io_signature{min_streams, max_streams, sizeof_stream_item});
    }
    int d_min, d_max, d_item;
};
}  // namespace gr
