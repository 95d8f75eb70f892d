// Minimal stand-in for the slice of GNU Radio's public API that lib/decoder_impl.{h,cc} of gr-lora
// touches (gr::sync_block, gr::io_signature, gnuradio::get_initial_sptr, message ports, pmt blobs).
// GNU Radio is not installed in the build container; this stub exists ONLY so that the drop-in shim
// gr_lora_b200/host/decoder_impl.cc -- the file a gr-lora maintainer would compile against the real
// GNU Radio -- is compiled and exercised by tests/test_gpu_cpp_shim.py.  It implements a one-block fake
// scheduler: run() calls work() with at least output_multiple items and honours consume_each().
#pragma once
#include <complex>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

typedef std::complex<float> gr_complex;
typedef std::vector<const void *> gr_vector_const_void_star;
typedef std::vector<void *> gr_vector_void_star;

namespace pmt {
struct pmt_base {
    std::string sym;
    std::vector<uint8_t> blob;
};
typedef std::shared_ptr<pmt_base> pmt_t;
inline pmt_t mp(const std::string &s) { auto p = std::make_shared<pmt_base>(); p->sym = s; return p; }
inline pmt_t intern(const std::string &s) { return mp(s); }
inline pmt_t make_blob(const void *buf, size_t len) {
    auto p = std::make_shared<pmt_base>();
    p->blob.assign((const uint8_t *)buf, (const uint8_t *)buf + len);
    return p;
}
inline const void *blob_data(const pmt_t &p) { return p->blob.data(); }
inline size_t blob_length(const pmt_t &p) { return p->blob.size(); }
}  // namespace pmt

namespace gr {
class io_signature {
public:
    typedef std::shared_ptr<io_signature> sptr;
    static sptr make(int min_streams, int max_streams, int sizeof_item) {
        auto s = std::make_shared<io_signature>();
        s->min = min_streams; s->max = max_streams; s->item = sizeof_item;
        return s;
    }
    int min = 0, max = 0, item = 0;
};

class sync_block {
public:
    enum { WORK_DONE = -1 };
    sync_block(const std::string &name, io_signature::sptr in, io_signature::sptr out) : d_name(name), d_in(in), d_out(out) {}
    virtual ~sync_block() {}
    virtual int work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items) = 0;
    void set_output_multiple(int m) { d_output_multiple = m; }
    int output_multiple() const { return d_output_multiple; }
    void set_min_noutput_items(int m) { d_min_noutput = m; }
    void consume_each(int n) { d_consumed += n; }
    void message_port_register_out(pmt::pmt_t port) { d_ports[port->sym]; }
    void message_port_pub(pmt::pmt_t port, pmt::pmt_t msg) {
        for (auto &h : d_ports[port->sym]) h(msg);
    }
    void message_port_subscribe(const std::string &port, std::function<void(pmt::pmt_t)> h) { d_ports[port].push_back(h); }
    const std::string &name() const { return d_name; }

    // fake scheduler: present the unconsumed tail again, in chunks of at most max_items
    size_t run(const gr_complex *samples, size_t n_items, size_t max_items) {
        size_t pos = 0;
        while (n_items - pos >= (size_t)d_output_multiple) {
            size_t n = n_items - pos;
            if (n > max_items) n = max_items;
            n -= n % (size_t)d_output_multiple;
            gr_vector_const_void_star in(1, samples + pos);
            gr_vector_void_star out;
            d_consumed = 0;
            if (work((int)n, in, out) == WORK_DONE) break;
            if (d_consumed <= 0) break;
            pos += (size_t)d_consumed;
        }
        return pos;
    }

private:
    std::string d_name;
    io_signature::sptr d_in, d_out;
    int d_output_multiple = 1, d_min_noutput = 1;
    long d_consumed = 0;
    std::map<std::string, std::vector<std::function<void(pmt::pmt_t)>>> d_ports;
};
}  // namespace gr

namespace gnuradio {
template <class T>
std::shared_ptr<T> get_initial_sptr(T *p) { return std::shared_ptr<T>(p); }
}  // namespace gnuradio
