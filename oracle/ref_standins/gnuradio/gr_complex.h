// stand-in for <gnuradio/gr_complex.h> (include/lora/debugger.h:24)
#pragma once
#include <complex>
typedef std::complex<float> gr_complex;
typedef std::complex<double> gr_complexd;
