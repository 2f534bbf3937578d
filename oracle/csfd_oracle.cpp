// oracle/csfd_oracle.cpp — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the reference's host second-order-complex type (DeviceArray/src/DoubleComplex.cpp) and of
// the DCSFD self-check of Experiments/test_CSFD/main.cpp:194-219, on plain arrays of 4 numbers
// [re.re, re.im, im.re, im.im] = [v, h d1, h d2, h^2 d12].  T = float is the reference's precision
// (std::complex<float>, DoubleComplex.h:5,14); T = double is the FP64 restatement.
// Pinned against (a) the known answers the reference prints (gradient 2.73911, second order 9.26892 at
// t = 0.5) and (b) oracle/_ref/libref_csfd.so = the reference's own DoubleComplex.cpp.
#include <cmath>
#include <complex>

namespace {

template <class T> struct B2 {
    std::complex<T> re, im;
};
template <class T> B2<T> ld(const T *p) { return {{p[0], p[1]}, {p[2], p[3]}}; }
template <class T> void st(T *p, const B2<T> &x) {
    p[0] = x.re.real(), p[1] = x.re.imag(), p[2] = x.im.real(), p[3] = x.im.imag();
}
template <class T> B2<T> add(B2<T> a, B2<T> b) { return {a.re + b.re, a.im + b.im}; }  // .cpp:143-148
template <class T> B2<T> sub(B2<T> a, B2<T> b) { return {a.re - b.re, a.im - b.im}; }  // .cpp:151-156
template <class T> B2<T> mul(B2<T> a, B2<T> b) {                                         // .cpp:159-166
    return {a.re * b.re - a.im * b.im, a.im * b.re + a.re * b.im};
}
template <class T> std::complex<T> nrm(B2<T> x) { return x.re * x.re + x.im * x.im; }  // .cpp:320-323
template <class T> B2<T> dvd(B2<T> a, B2<T> b) {                                          // .cpp:168-175
    const std::complex<T> r = a.re * b.re + a.im * b.im, n = nrm(b);
    return {r / n, (a.im * b.re - a.re * b.im) / n};
}
template <class T> std::complex<T> absb(B2<T> x) { return std::sqrt(x.re * x.re + x.im * x.im); }  // .cpp:303-307
template <class T> B2<T> sqrtb(B2<T> x) {                                                            // .cpp:332-349
    B2<T> result = x;
    const std::complex<T> r = absb(x), sqrt_r = std::sqrt(r);
    result.re += r;
    const std::complex<T> zr = absb(result);
    if (std::fabs(zr.real()) < T(1e-20) && std::fabs(zr.imag()) < T(1e-20)) return {result.re * sqrt_r, result.im * sqrt_r};
    const std::complex<T> scale = sqrt_r / zr;
    return {result.re * scale, result.im * scale};
}
template <class T> B2<T> expb(B2<T> x) { return {std::exp(x.re) * std::cos(x.im), std::exp(x.re) * std::sin(x.im)}; }  // .cpp:351-356
template <class T> std::complex<T> atan2c(std::complex<T> y, std::complex<T> x) {                                          // .cpp:386-401
    std::complex<T> r = std::sqrt(x * x + y * y);
    if (r.real() > T(0)) {
        r += x;
        r = y / r;
    } else {
        r -= x;
        r = r / y;
    }
    r = std::atan(r);
    r *= T(2);
    return r;
}
template <class T> B2<T> logb_(B2<T> x) { return {std::log(absb(x)), atan2c(x.im, x.re)}; }  // .cpp:358-366
template <class T> B2<T> sinb(B2<T> x) {                                                      // .cpp:421-426
    return {std::cosh(-x.im) * std::sin(x.re), -std::sinh(-x.im) * std::cos(x.re)};
}
template <class T> B2<T> cosb(B2<T> x) {  // .cpp:428-433
    return {std::cosh(-x.im) * std::cos(x.re), std::sinh(-x.im) * std::sin(x.re)};
}
template <class T> B2<T> powb(B2<T> x, T y) {  // .cpp:435-440 with polar (.cpp:325-330)
    const B2<T> r = logb_(x);
    const std::complex<T> rho = std::exp(y * r.re), theta = y * r.im;
    return {rho * std::cos(theta), rho * std::sin(theta)};
}

template <class T> int apply(int op, const T *a, const T *b, T p, T *out, long n) {
    for (long i = 0; i < n; ++i) {
        const B2<T> x = ld(a + 4 * i), y = b ? ld(b + 4 * i) : x;
        B2<T> r;
        switch (op) {
            case 0: r = add(x, y); break;
            case 1: r = sub(x, y); break;
            case 2: r = mul(x, y); break;
            case 3: r = dvd(x, y); break;
            case 4: r = sqrtb(x); break;
            case 5: r = expb(x); break;
            case 6: r = logb_(x); break;
            case 7: r = sinb(x); break;
            case 8: r = cosb(x); break;
            case 10: r = powb(x, p); break;
            default: return -1;
        }
        st(out + 4 * i, r);
    }
    return 0;
}
// main.cpp:194-205
template <class T> void chain(const T *t, T h, T *out, long n) {
    for (long i = 0; i < n; ++i) {
        const B2<T> tt = {{t[i], h}, {h, 0}};
        const B2<T> x = mul(tt, tt), y = sinb(tt), s = add(x, y);
        st(out + 4 * i, mul(s, s));
    }
}

}  // namespace

extern "C" {
int oracle_dc_apply_f32(int op, const float *a, const float *b, float p, float *out, long n) { return apply<float>(op, a, b, p, out, n); }
int oracle_dc_apply_f64(int op, const double *a, const double *b, double p, double *out, long n) { return apply<double>(op, a, b, p, out, n); }
int oracle_dc_chain_f32(const float *t, float h, float *out, long n) { return chain<float>(t, h, out, n), 0; }
int oracle_dc_chain_f64(const double *t, double h, double *out, long n) { return chain<double>(t, h, out, n), 0; }
}
