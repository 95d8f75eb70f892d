// stand-in for <gnuradio/sync_block.h> (include/lora/decoder.h:683): the members decoder_impl uses.
// consume_each() adds to a counter the test driver reads after every work() call; message_port_pub()
// appends the blob to a per-port list.  No scheduler: oracle/ref_wrap.cc calls work() itself.
#pragma once
#include <gnuradio/gr_complex.h>
#include <gnuradio/io_signature.h>
#include <pmt/pmt.h>
#include <map>
#include <string>
#include <vector>

typedef std::vector<const void *> gr_vector_const_void_star;
typedef std::vector<void *> gr_vector_void_star;

namespace gr {
class sync_block {
public:
    sync_block() {}     // virtual base of lora::decoder: default-constructible like gr::basic_block's protected ctor
    sync_block(const std::string &name, io_signature::sptr in, io_signature::sptr out) : d_name(name), d_insig(in), d_outsig(out) {}
    virtual ~sync_block() {}
    virtual int work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items) = 0;
    void set_output_multiple(int m) { standin_output_multiple = m; }
    void consume_each(int n) { standin_consumed += n; }
    void message_port_register_out(pmt::pmt_t port) { standin_ports[port->symbol]; }
    void message_port_pub(pmt::pmt_t port, pmt::pmt_t msg) { standin_ports[port->symbol].push_back(msg); }

    int standin_output_multiple = 1;
    long standin_consumed = 0;
    std::map<std::string, std::vector<pmt::pmt_t>> standin_ports;

private:
    std::string d_name;
    io_signature::sptr d_insig, d_outsig;
};
}  // namespace gr

namespace gnuradio {
template <class T>
std::shared_ptr<T> get_initial_sptr(T *p) { return std::shared_ptr<T>(p); }
}  // namespace gnuradio
