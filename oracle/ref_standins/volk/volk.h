// stand-in for <volk/volk.h>: the six kernels lib/decoder_impl.cc calls (:261,267,350-354,371-372,475,734),
// as VOLK's generic protokernels compute them: scalar, in order, single precision.
#pragma once
#include <complex>
typedef std::complex<float> lv_32fc_t;

static inline void volk_32f_x2_dot_prod_32f(float *result, const float *input, const float *taps, unsigned int num_points) {
    float acc = 0.0f;
    for (unsigned int i = 0; i < num_points; i++) acc += input[i] * taps[i];
    *result = acc;
}
static inline void volk_32fc_x2_conjugate_dot_prod_32fc(lv_32fc_t *result, const lv_32fc_t *input, const lv_32fc_t *taps, unsigned int num_points) {
    lv_32fc_t acc(0.0f, 0.0f);
    for (unsigned int i = 0; i < num_points; i++) acc += input[i] * std::conj(taps[i]);
    *result = acc;
}
static inline void volk_32fc_magnitude_squared_32f(float *magnitude, const lv_32fc_t *in, unsigned int num_points) {
    for (unsigned int i = 0; i < num_points; i++) {
        const float re = in[i].real(), im = in[i].imag();
        magnitude[i] = re * re + im * im;
    }
}
static inline void volk_32f_accumulator_s32f(float *result, const float *input, unsigned int num_points) {
    float acc = 0.0f;
    for (unsigned int i = 0; i < num_points; i++) acc += input[i];
    *result = acc;
}
static inline void volk_32fc_x2_multiply_32fc(lv_32fc_t *out, const lv_32fc_t *a, const lv_32fc_t *b, unsigned int num_points) {
    for (unsigned int i = 0; i < num_points; i++) out[i] = a[i] * b[i];
}
