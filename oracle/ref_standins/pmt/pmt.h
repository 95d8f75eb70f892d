// stand-in for <pmt/pmt.h>: the three calls of lib/decoder_impl.cc:120-121,607-608
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
namespace pmt {
struct pmt_base {
    std::string symbol;
    std::vector<uint8_t> blob;
};
typedef std::shared_ptr<pmt_base> pmt_t;
inline pmt_t mp(const std::string &s) { pmt_t p(new pmt_base); p->symbol = s; return p; }
inline pmt_t intern(const std::string &s) { return mp(s); }
inline pmt_t make_blob(const void *buf, size_t len) {
    pmt_t p(new pmt_base);
    p->blob.assign((const uint8_t *)buf, (const uint8_t *)buf + len);
    return p;
}
}  // namespace pmt
