// Stand-in for gr::hier_block2 (GNU Radio is absent from the build container): just enough for the channelizer shim
// (gr_lora_b200/host/channelizer_impl.cc) to keep the reference's class shape -- name, signatures, the "control"
// message port and the wiring calls -- while the test driver calls the inner filter block's work() itself.
#pragma once
#include <gnuradio/sync_block.h>

namespace gr {

typedef std::shared_ptr<sync_block> basic_block_sptr;

class hier_block2 {
public:
    hier_block2() {}
    hier_block2(const std::string &name, io_signature::sptr in, io_signature::sptr out) : d_name(name), d_in(in), d_out(out) {}
    virtual ~hier_block2() {}
    hier_block2 *self() { return this; }
    void connect(hier_block2 *, int, basic_block_sptr dst, int) { d_first = dst; }      // self -> first inner block
    void connect(basic_block_sptr src, int, hier_block2 *, int) { d_last = src; }       // last inner block -> self
    void message_port_register_hier_in(pmt::pmt_t port) { d_msg_in.push_back(port->sym); }
    const std::string &name() const { return d_name; }
    basic_block_sptr first_block() const { return d_first; }
    io_signature::sptr output_signature() const { return d_out; }

private:
    std::string d_name;
    io_signature::sptr d_in, d_out;
    basic_block_sptr d_first, d_last;
    std::vector<std::string> d_msg_in;
};

}  // namespace gr
