// stand-in for <gnuradio/expj.h> (lib/decoder_impl.cc:27,159-160): e^{j phase} from a FLOAT phase,
// cosine and sine in single precision like GNU Radio's gr::sincosf
#pragma once
#include <cmath>
#include <gnuradio/gr_complex.h>
static inline gr_complex gr_expj(float phase) {
    float s, c;
    ::sincosf(phase, &s, &c);
    return gr_complex(c, s);
}
