// oracle/ref_csfd_wrap.cpp — TEST INFRASTRUCTURE ONLY.
//
// C-ABI shell around the UNMODIFIED reference host bicomplex type (DeviceArray/include/DoubleComplex.h,
// DeviceArray/src/DoubleComplex.cpp), compiled from /root/reference into oracle/_ref/libref_csfd.so.
// It is (a) the pin for oracle/csfd_oracle.cpp and for the product's DoubleComplex / SoA kernels and
// (b) the "reference CPU path" timed by bench.py (the DeviceArray CSFD math exercised by
// Experiments/test_CSFD/main.cpp:194-219).
//
// A bicomplex value is 4 floats [re.re, re.im, im.re, im.im] = [v, h d1, h d2, h^2 d12].
#include "DoubleComplex.h"

#include <chrono>
#include <cstring>

namespace {
inline DoubleComplex ld(const float *p) { return DoubleComplex(p[0], p[1], p[2], p[3]); }
inline void st(float *p, const DoubleComplex &x) {
    p[0] = x.real().real();
    p[1] = x.real().imag();
    p[2] = x.imag().real();
    p[3] = x.imag().imag();
}
// Experiments/test_CSFD/main.cpp:9-12
inline DoubleComplex f1(DoubleComplex x, DoubleComplex y) { return (x + y) * (x + y); }
}  // namespace

extern "C" {

enum { OP_ADD = 0, OP_SUB, OP_MUL, OP_DIV, OP_SQRT, OP_EXP, OP_LOG, OP_SIN, OP_COS, OP_ATAN2, OP_POW, OP_ATAN };

// n elementwise operations; b may be NULL for unary ops; pow uses the float exponent p.
int ref_dc_apply(int op, const float *a, const float *b, float p, float *out, long n) {
    for (long i = 0; i < n; ++i) {
        DoubleComplex x = ld(a + 4 * i), y = b ? ld(b + 4 * i) : DoubleComplex(0.f), r;
        switch (op) {
            case OP_ADD: r = x + y; break;
            case OP_SUB: r = x - y; break;
            case OP_MUL: r = x * y; break;
            case OP_DIV: r = x / y; break;
            case OP_SQRT: r = sqrt(x); break;
            case OP_EXP: r = exp(x); break;
            case OP_LOG: r = log(x); break;
            case OP_SIN: r = sin(x); break;
            case OP_COS: r = cos(x); break;
            case OP_ATAN2: r = atan2(x, y); break;
            case OP_POW: r = pow(x, p); break;
            case OP_ATAN: r = atan(x); break;
            default: return -1;
        }
        st(out + 4 * i, r);
    }
    return 0;
}

// The DCSFD self-check of test_CSFD (main.cpp:194-205) evaluated at n points t[i]:
// loss = f1(t*t, sin t), t seeded as DoubleComplex((t,h),(h,0)).  out: n x 4 floats.
int ref_dc_chain(const float *t, float h, float *out, long n) {
    for (long i = 0; i < n; ++i) {
        DoubleComplex tt(SingleComplex(t[i], h), SingleComplex(h, 0));
        DoubleComplex x = tt * tt;
        DoubleComplex y = sin(tt);
        st(out + 4 * i, f1(x, y));
    }
    return 0;
}

// Same chain, timed over `reps` passes; returns evaluations per second (1 thread) and a checksum that
// consumes every result (the stock test_CSFD loops discard theirs and are dead code at -O2).
double ref_dc_chain_bench(const float *t, float h, long n, int reps, double *checksum) {
    double acc = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r)
        for (long i = 0; i < n; ++i) {
            DoubleComplex tt(SingleComplex(t[i], h), SingleComplex(h, 0));
            DoubleComplex l = f1(tt * tt, sin(tt));
            acc += l.real().imag() + l.imag().imag();
        }
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (checksum) *checksum = acc;
    return double(n) * reps / s;
}

}  // extern "C"
