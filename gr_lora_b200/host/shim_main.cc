// shim_main.cc -- what apps/lora_receive_file_nogui.py does, in C++ against the shim: read a cf32 file,
// run it through lora::decoder::make(...) under the fake scheduler, print every published frame as hex
// on stderr ("FRAME <hex>") while std::cout carries exactly what the reference block prints.
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>
#include <lora/decoder.h>

int main(int argc, char **argv) {
    if (argc < 8) {
        std::fprintf(stderr, "usage: %s file.cf32 samp_rate bandwidth sf implicit cr crc [reduced_rate] [chunk_items]\n", argv[0]);
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
    auto blk = gr::lora::decoder::make((float)std::atof(argv[2]), (uint32_t)std::atoi(argv[3]), (uint8_t)std::atoi(argv[4]),
                                       std::atoi(argv[5]) != 0, (uint8_t)std::atoi(argv[6]), std::atoi(argv[7]) != 0,
                                       argc > 8 ? std::atoi(argv[8]) != 0 : false, false);
    blk->message_port_subscribe("frames", [](pmt::pmt_t msg) {
        std::fprintf(stderr, "FRAME ");
        const uint8_t *p = (const uint8_t *)pmt::blob_data(msg);
        for (size_t i = 0; i < pmt::blob_length(msg); i++) std::fprintf(stderr, "%02x", p[i]);
        std::fprintf(stderr, "\n");
    });
    const size_t chunk = argc > 9 ? (size_t)std::atol(argv[9]) : (size_t)1 << 20;
    const size_t consumed = blk->run(iq.data(), iq.size(), chunk);
    std::fprintf(stderr, "CONSUMED %zu\n", consumed);
    return 0;
}
