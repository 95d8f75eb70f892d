// chan_shim_main.cc -- file in, file out through lora::channelizer::make(...) under a hand-driven scheduler: the inner
// block's work() is called with GNU-Radio-sized buffers, every channel's output is appended to <out>.<ch>.cf32.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "channelizer_impl.h"

int main(int argc, char **argv) {
    if (argc < 8) {
        std::fprintf(stderr, "usage: %s in.cf32 out_prefix samp_rate center_freq bandwidth decimation chunk_out ch0 [ch1 ...]\n", argv[0]);
        return 2;
    }
    std::FILE *f = std::fopen(argv[1], "rb");
    if (!f) { std::perror("open"); return 2; }
    std::fseek(f, 0, SEEK_END);
    const long bytes = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<gr_complex> iq((size_t)bytes / sizeof(gr_complex));
    if (std::fread(iq.data(), sizeof(gr_complex), iq.size(), f) != iq.size()) { std::perror("read"); return 2; }
    std::fclose(f);
    const uint32_t decim = (uint32_t)std::atoi(argv[6]);
    const size_t chunk = (size_t)std::atol(argv[7]);
    std::vector<float> chans;
    for (int i = 8; i < argc; i++) chans.push_back((float)std::atof(argv[i]));
    auto blk = gr::lora::channelizer::make((float)std::atof(argv[3]), (float)std::atof(argv[4]), chans, (uint32_t)std::atoi(argv[5]), decim);
    auto impl = std::dynamic_pointer_cast<gr::lora::channelizer_impl>(blk);
    std::vector<std::FILE *> outs;
    for (size_t c = 0; c < chans.size(); c++) outs.push_back(std::fopen((std::string(argv[2]) + "." + std::to_string(c) + ".cf32").c_str(), "wb"));
    std::vector<std::vector<gr_complex>> bufs(chans.size(), std::vector<gr_complex>(chunk));
    size_t pos = 0, total = 0;
    while (iq.size() - pos >= decim) {
        size_t n_out = std::min(chunk, (iq.size() - pos) / decim);
        gr_vector_const_void_star in(1, iq.data() + pos);
        gr_vector_void_star out;
        for (auto &b : bufs) out.push_back(b.data());
        const int got = impl->filter()->work((int)n_out, in, out);
        if (got <= 0) break;
        for (size_t c = 0; c < chans.size(); c++) std::fwrite(bufs[c].data(), sizeof(gr_complex), (size_t)got, outs[c]);
        pos += (size_t)got * decim;
        total += (size_t)got;
        if (total == chunk && argc > 8 && std::getenv("CHAN_SHIM_CFO")) impl->apply_cfo((float)std::atof(std::getenv("CHAN_SHIM_CFO")));
    }
    for (auto *o : outs) std::fclose(o);
    std::fprintf(stderr, "PRODUCED %zu\n", total);
    return 0;
}
