// xslam_dcomplex.hpp — header-only host number type with the API surface of the reference's DoubleComplex
// (DeviceArray/include/DoubleComplex.h:15-95, src/DoubleComplex.cpp) and the accessors of its device twin d_complex<T>
// (DeviceArray/include/cuda_double_complex.hpp:47-55: value(), grad(), hessian()), plus an owning RAII view of the packed-SoA
// device arrays of libxslam_b200 (xs_dc_array_*, the DeviceArray<T> semantics of device_array.hpp:25-134).
//
// A second-order complex (bicomplex) number is z = a + j b with a, b complex<float> and i, j two commuting imaginary units:
//     z = v + i e1 + j e2 + ij e12        a = (v, e1),  b = (e2, e12)
// Seeding a variable t as t + i h + j h (addPerturbation, DoubleComplex.h:33 / .cpp:61-66, h = 1e-6) makes
// f(t).grad() / h = f'(t) and f(t).hessian() / h^2 = f''(t): the DCSFD rule of Experiments/test_CSFD/main.cpp:194-205.
// The object is four floats in the order (v, e1, e2, e12) - the memory layout of the reference's class - so an array of
// them is what xs_dc_array_upload / DeviceArray4::upload take.
//
// Every operation is the tower formula over complex<float> (z1 z2 = (a1 a2 - b1 b2) + j (b1 a2 + a1 b2), ...) evaluated with
// std::complex<float> like the reference's host type, so results agree with it to FP32 rounding (tests/test_dcomplex_host.py
// holds them against the reference's own DoubleComplex.cpp).  Stated deviation, shared with the device kernels
// (x-slam_b200/csrc/dc_array.cu): the reference's atanh evaluates log(a - a) (DoubleComplex.cpp:372-377), which makes its
// atan / atan2 non-finite; here atanh(x) = (log(1 + x) - log(1 - x)) / 2.
#pragma once
#include "xslam_b200.h"

#include <cmath>
#include <complex>
#include <ostream>
#include <vector>

namespace xslam_b200 {

typedef float MyFloat;
typedef std::complex<MyFloat> SingleComplex;

class DoubleComplex {
    SingleComplex a_, b_;  // z = a_ + j b_

  public:
    DoubleComplex() : a_(0), b_(0) {}
    DoubleComplex(MyFloat real) : a_(real), b_(0) {}
    DoubleComplex(SingleComplex real) : a_(real), b_(0) {}
    DoubleComplex(SingleComplex real, SingleComplex imag) : a_(real), b_(imag) {}
    DoubleComplex(MyFloat real_real, MyFloat real_imag, MyFloat imag_real, MyFloat imag_imag)
        : a_(real_real, real_imag), b_(imag_real, imag_imag) {}

    SingleComplex real() const { return a_; }
    SingleComplex imag() const { return b_; }
    void real(SingleComplex r) { a_ = r; }
    void imag(SingleComplex i) { b_ = i; }
    // d_complex accessors (cuda_double_complex.hpp:47-55)
    MyFloat value() const { return a_.real(); }
    MyFloat grad() const { return a_.imag(); }
    MyFloat hessian() const { return b_.imag(); }

    // DoubleComplex.cpp:61-72: h = 1e-6 on both first-order slots
    void addPerturbation() {
        const MyFloat h = 1e-6f;
        a_ = SingleComplex(a_.real(), h);
        b_ = SingleComplex(h, 0);
    }
    void clearPerturbation() {
        a_ = SingleComplex(a_.real(), 0);
        b_ = SingleComplex(0, 0);
    }

    DoubleComplex operator-() const { return DoubleComplex(-a_, -b_); }
    DoubleComplex &operator=(const SingleComplex &o) { return a_ = o, b_ = 0, *this; }
    DoubleComplex &operator=(const MyFloat &o) { return a_ = o, b_ = 0, *this; }

    DoubleComplex &operator+=(const MyFloat &o) { return a_ += o, *this; }
    DoubleComplex &operator-=(const MyFloat &o) { return a_ -= o, *this; }
    DoubleComplex &operator*=(const MyFloat &o) { return a_ *= o, b_ *= o, *this; }
    DoubleComplex &operator/=(const MyFloat &o) { return a_ /= o, b_ /= o, *this; }
    DoubleComplex &operator+=(const SingleComplex &o) { return a_ += o, *this; }
    DoubleComplex &operator-=(const SingleComplex &o) { return a_ -= o, *this; }
    DoubleComplex &operator*=(const SingleComplex &o) { return a_ *= o, b_ *= o, *this; }
    DoubleComplex &operator/=(const SingleComplex &o) { return a_ /= o, b_ /= o, *this; }
    DoubleComplex &operator+=(const DoubleComplex &o) { return a_ += o.a_, b_ += o.b_, *this; }
    DoubleComplex &operator-=(const DoubleComplex &o) { return a_ -= o.a_, b_ -= o.b_, *this; }
    // (a1 + j b1)(a2 + j b2) = (a1 a2 - b1 b2) + j (b1 a2 + a1 b2)
    DoubleComplex &operator*=(const DoubleComplex &o) {
        const SingleComplex a = a_ * o.a_ - b_ * o.b_, b = b_ * o.a_ + a_ * o.b_;
        return a_ = a, b_ = b, *this;
    }
    // z1 / z2 = z1 conj_j(z2) / (a2^2 + b2^2): the j-norm of z2 is a plain complex number
    DoubleComplex &operator/=(const DoubleComplex &o) {
        const SingleComplex n = o.a_ * o.a_ + o.b_ * o.b_;
        const SingleComplex a = (a_ * o.a_ + b_ * o.b_) / n, b = (b_ * o.a_ - a_ * o.b_) / n;
        return a_ = a, b_ = b, *this;
    }
};

inline std::ostream &operator<<(std::ostream &os, const DoubleComplex &x) { return os << '(' << x.real() << ',' << x.imag() << ')'; }

// comparisons look at the value only (DoubleComplex.cpp:248-276)
inline bool operator>(const DoubleComplex &l, const DoubleComplex &r) { return l.value() > r.value(); }
inline bool operator>(const DoubleComplex &l, const SingleComplex &r) { return l.value() > r.real(); }
inline bool operator>(const DoubleComplex &l, const MyFloat &r) { return l.value() > r; }
inline bool operator<(const DoubleComplex &l, const DoubleComplex &r) { return l.value() < r.value(); }
inline bool operator<(const DoubleComplex &l, const SingleComplex &r) { return l.value() < r.real(); }
inline bool operator<(const DoubleComplex &l, const MyFloat &r) { return l.value() < r; }

inline DoubleComplex operator+(DoubleComplex l, const MyFloat &r) { return l += r; }
inline DoubleComplex operator-(DoubleComplex l, const MyFloat &r) { return l -= r; }
inline DoubleComplex operator*(DoubleComplex l, const MyFloat &r) { return l *= r; }
inline DoubleComplex operator/(DoubleComplex l, const MyFloat &r) { return l /= r; }
inline DoubleComplex operator+(DoubleComplex l, const DoubleComplex &r) { return l += r; }
inline DoubleComplex operator-(DoubleComplex l, const DoubleComplex &r) { return l -= r; }
inline DoubleComplex operator*(DoubleComplex l, const DoubleComplex &r) { return l *= r; }
inline DoubleComplex operator/(DoubleComplex l, const DoubleComplex &r) { return l /= r; }

inline SingleComplex real(const DoubleComplex &x) { return x.real(); }
inline SingleComplex imag(const DoubleComplex &x) { return x.imag(); }
inline MyFloat fabs(const DoubleComplex &x) { return std::fabs(x.value()); }
inline SingleComplex norm(const DoubleComplex &x) { return x.real() * x.real() + x.imag() * x.imag(); }  // a^2 + b^2
inline SingleComplex abs(const DoubleComplex &x) { return std::sqrt(norm(x)); }                            // r = sqrt(a^2 + b^2)
inline DoubleComplex abs2(const DoubleComplex &x) { return x * x; }
inline DoubleComplex conj(const DoubleComplex &x) { return DoubleComplex(x.real(), -x.imag()); }
inline DoubleComplex polar(const SingleComplex &rho, const SingleComplex &theta) {
    return DoubleComplex(rho * std::cos(theta), rho * std::sin(theta));
}

// atan2 of two complex numbers through the half-angle identity 2 atan(y / (r + x)), r = sqrt(x^2 + y^2); the branch is
// decided on real parts like every comparison of this type (DoubleComplex.cpp:386-401)
inline SingleComplex atan2(const SingleComplex &y, const SingleComplex &x) {
    const SingleComplex r = std::sqrt(x * x + y * y);
    const SingleComplex q = r.real() > 0.f ? y / (r + x) : (r - x) / y;
    return std::atan(q) * 2.0f;
}
inline SingleComplex arg(const DoubleComplex &x) { return atan2(x.imag(), x.real()); }

// sqrt(z) = sqrt(r) (z + r) / |z + r| with r = |z| (both moduli are j-norms, i.e. complex numbers), DoubleComplex.cpp:332-349
inline DoubleComplex sqrt(const DoubleComplex &x) {
    const SingleComplex r = abs(x), root_r = std::sqrt(r);
    DoubleComplex w = x;
    w += r;
    const SingleComplex m = abs(w);
    if (std::fabs(m.real()) < 1e-20f && std::fabs(m.imag()) < 1e-20f) return w *= root_r, w;
    return w *= root_r / m, w;
}
inline DoubleComplex abs_d(const DoubleComplex &x) { return sqrt(x * x); }
// exp(a + j b) = e^a (cos b + j sin b)
inline DoubleComplex exp(const DoubleComplex &x) {
    const SingleComplex e = std::exp(x.real());
    return DoubleComplex(e * std::cos(x.imag()), e * std::sin(x.imag()));
}
// log z = log r + j atan2(b, a)
inline DoubleComplex log(const DoubleComplex &x) { return DoubleComplex(std::log(abs(x)), atan2(x.imag(), x.real())); }
// sin(a + j b) = sin a cosh b + j cos a sinh b,  cos(a + j b) = cos a cosh b - j sin a sinh b
inline DoubleComplex sin(const DoubleComplex &x) {
    return DoubleComplex(std::cosh(-x.imag()) * std::sin(x.real()), -std::sinh(-x.imag()) * std::cos(x.real()));
}
inline DoubleComplex cos(const DoubleComplex &x) {
    return DoubleComplex(std::cosh(-x.imag()) * std::cos(x.real()), std::sinh(-x.imag()) * std::sin(x.real()));
}
// x^y = exp(y log x) written in polar form
inline DoubleComplex pow(const DoubleComplex &x, const MyFloat y) {
    const DoubleComplex l = log(x);
    return polar(std::exp(y * l.real()), y * l.imag());
}
inline DoubleComplex atanh(const DoubleComplex &x) {  // see the header comment for the deviation from the reference
    const DoubleComplex one(SingleComplex(1, 0), SingleComplex(0, 0));
    return (log(one + x) - log(one - x)) * 0.5f;
}
// atan z = -j atanh(j z)
inline DoubleComplex atan(const DoubleComplex &x) {
    const DoubleComplex r = atanh(DoubleComplex(-x.imag(), x.real()));
    return DoubleComplex(r.imag(), -r.real());
}
inline DoubleComplex atan2(const DoubleComplex &y, const DoubleComplex &x) {
    DoubleComplex r = sqrt(x * x + y * y);
    if (r > 0.0f) {
        r += x;
        r = y / r;
    } else {
        r -= x;
        r = r / y;
    }
    r = atan(r);
    return r *= 2.0f, r;
}

static_assert(sizeof(DoubleComplex) == 4 * sizeof(float), "DoubleComplex is four packed floats (v, e1, e2, e12)");

// Owning packed-SoA device array of bicomplex numbers: the method names and semantics of DeviceArray<T>
// (device_array.hpp:25-134): create is a no-op when the size is unchanged, upload / download block, copyTo is a deep copy.
class DeviceArray4 {
    xs_dc_array *h_ = nullptr;

    void ensure() {
        if (!h_) h_ = xs_dc_array_create(0);
    }

  public:
    DeviceArray4() = default;
    explicit DeviceArray4(size_t n) { create(n); }
    DeviceArray4(const DoubleComplex *host, size_t n) { upload(host, n); }
    DeviceArray4(const DeviceArray4 &) = delete;
    DeviceArray4 &operator=(const DeviceArray4 &) = delete;
    DeviceArray4(DeviceArray4 &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    ~DeviceArray4() { release(); }

    bool create(size_t n) { return ensure(), h_ && xs_dc_array_resize(h_, (long) n) == XS_OK; }
    void release() {
        xs_dc_array_release(h_);
        h_ = nullptr;
    }
    bool upload(const DoubleComplex *host, size_t n) {
        return ensure(), h_ && xs_dc_array_upload(h_, reinterpret_cast<const float *>(host), (long) n) == XS_OK;
    }
    bool upload(const std::vector<DoubleComplex> &v) { return upload(v.data(), v.size()); }
    bool download(DoubleComplex *host) const { return h_ && xs_dc_array_download(h_, reinterpret_cast<float *>(host)) == XS_OK; }
    bool download(std::vector<DoubleComplex> &v) const {
        v.resize(size());
        return download(v.data());
    }
    bool copyTo(DeviceArray4 &other) const { return other.ensure(), h_ && other.h_ && xs_dc_array_copy(h_, other.h_) == XS_OK; }
    size_t size() const { return (size_t) xs_dc_array_size(h_); }
    bool empty() const { return size() == 0; }
    float *ptr() { return xs_dc_array_ptr(h_); }               // float[4][n] device planes
    const float *ptr() const { return xs_dc_array_ptr(h_); }
    // element-wise out = op(*this, b) on the device (xs_dc_apply); b may be null for unary operations
    bool apply(xs_dc_op op, const DeviceArray4 *b, float p, DeviceArray4 &out) const {
        if (!out.create(size())) return false;
        return xs_dc_apply(op, ptr(), b ? b->ptr() : nullptr, p, out.ptr(), (long) size(), nullptr) == XS_OK;
    }
};

}  // namespace xslam_b200
