// oracle/oracle_types.h — TEST INFRASTRUCTURE ONLY.
//
// Complex arithmetic of the reference's device code, restated for the CPU oracle:
// cuda::std::complex<float> as shipped with CUDA 12.9 / CCCL 2.8.2 (the dependency behind `devComplex`,
// XKinectFusion/include/Internal.h:24; it lives in the CUDA toolkit, not under /root/reference):
//   operator*   four products, x = ac - bd, y = ad + bc         (libcxx/include/complex:503-548)
//   operator/   logb/scalbn pre-scaling, (ac+bd)/(cc+dd)        (:633-701)
//   sqrt        polar(sqrt(abs(z)), arg(z)/2)                   (:1039-1055)
//   complex op scalar: component-wise                           (:393-416)
// T = float follows the reference's precision; T = double is the "FP64 restatement" the derivative
// tolerances of BASELINE.json are stated against.
#pragma once
#include <cmath>
#include <cstddef>

namespace xo {

// ---- second order in FP64: T = Dual makes every Cx<T> a complex number whose real and imaginary parts are dual
// numbers (value, d/d theta_j).  With the imaginary unit carrying h d/d theta_i as in the reference, im.d / h is the mixed second
// derivative d2 / (d theta_i d theta_j) - the quantity a bicomplex run would return in its eps1 eps2 part - evaluated without any
// step-size error in j.  Integer decisions still follow the real VALUE (re.v), exactly as with the other types.
struct Dual {
    double v, d;
    Dual() : v(0), d(0) {}
    Dual(double a, double b = 0) : v(a), d(b) {}
    explicit operator double() const { return v; }
    explicit operator float() const { return (float) v; }
};
inline Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
inline Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
inline Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
inline Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, a.d * b.v + a.v * b.d); }
inline Dual operator/(Dual a, Dual b) {
    const double q = a.v / b.v;
    return Dual(q, (a.d - q * b.d) / b.v);
}
inline Dual &operator+=(Dual &a, Dual b) { return a = a + b; }
inline bool operator<(Dual a, Dual b) { return a.v < b.v; }
inline bool operator>(Dual a, Dual b) { return a.v > b.v; }
inline bool operator<=(Dual a, Dual b) { return a.v <= b.v; }
inline bool operator>=(Dual a, Dual b) { return a.v >= b.v; }
inline bool operator==(Dual a, Dual b) { return a.v == b.v; }
inline bool operator!=(Dual a, Dual b) { return a.v != b.v; }
inline Dual fabs(Dual a) { return a.v < 0 ? -a : a; }
inline Dual fmax(Dual a, Dual b) { return a.v >= b.v ? a : b; }
inline double logb(Dual a) { return std::logb(a.v); }
inline Dual scalbn(Dual a, int n) { return Dual(std::scalbn(a.v, n), std::scalbn(a.d, n)); }
inline bool isnan(Dual a) { return std::isnan(a.v); }
inline Dual copysign(Dual a, Dual s) { return std::signbit(a.v) == std::signbit(s.v) ? a : -a; }
inline Dual sqrt(Dual a) {
    const double r = std::sqrt(a.v);
    return Dual(r, r != 0 ? a.d / (2 * r) : 0.0);
}
inline Dual hypot(Dual a, Dual b) {
    const double r = std::hypot(a.v, b.v);
    return Dual(r, r != 0 ? (a.v * a.d + b.v * b.d) / r : 0.0);
}
inline Dual atan2(Dual y, Dual x) {
    const double q = x.v * x.v + y.v * y.v;
    return Dual(std::atan2(y.v, x.v), q != 0 ? (x.v * y.d - y.v * x.d) / q : 0.0);
}
inline Dual cos(Dual a) { return Dual(std::cos(a.v), -std::sin(a.v) * a.d); }
inline Dual sin(Dual a) { return Dual(std::sin(a.v), std::cos(a.v) * a.d); }
using std::atan2;
using std::copysign;
using std::cos;
using std::fabs;
using std::fmax;
using std::hypot;
using std::isnan;
using std::logb;
using std::scalbn;
using std::sin;
using std::sqrt;

template <class T> struct Cx {
    T re, im;
    Cx() : re(0), im(0) {}
    Cx(T r, T i = 0) : re(r), im(i) {}
};
template <class T> inline Cx<T> operator+(Cx<T> a, Cx<T> b) { return Cx<T>(a.re + b.re, a.im + b.im); }
template <class T> inline Cx<T> operator-(Cx<T> a, Cx<T> b) { return Cx<T>(a.re - b.re, a.im - b.im); }
template <class T> inline Cx<T> operator-(Cx<T> a) { return Cx<T>(-a.re, -a.im); }
template <class T> inline Cx<T> operator+(Cx<T> a, T s) { return Cx<T>(a.re + s, a.im); }
template <class T> inline Cx<T> operator-(Cx<T> a, T s) { return Cx<T>(a.re - s, a.im); }
template <class T> inline Cx<T> operator-(T s, Cx<T> a) { return Cx<T>(s - a.re, -a.im); }
template <class T> inline Cx<T> operator*(Cx<T> a, T s) { return Cx<T>(a.re * s, a.im * s); }
template <class T> inline Cx<T> operator*(T s, Cx<T> a) { return Cx<T>(a.re * s, a.im * s); }
template <class T> inline Cx<T> operator/(Cx<T> a, T s) { return Cx<T>(a.re / s, a.im / s); }
template <class T> inline Cx<T> operator*(Cx<T> z, Cx<T> w) {
    const T ac = z.re * w.re, bd = z.im * w.im, ad = z.re * w.im, bc = z.im * w.re;
    return Cx<T>(ac - bd, ad + bc);
}
template <class T> inline Cx<T> operator/(Cx<T> z, Cx<T> w) {
    int il = 0;
    T c = w.re, d = w.im;
    const double lb = (double) logb(fmax(fabs(c), fabs(d)));
    if (std::isfinite(lb)) {
        il = (int) lb;
        c = scalbn(c, -il);
        d = scalbn(d, -il);
    }
    const T denom = c * c + d * d;
    T x = scalbn((z.re * c + z.im * d) / denom, -il);
    T y = scalbn((z.im * c - z.re * d) / denom, -il);
    if (isnan(x) && isnan(y) && denom == T(0) && (!isnan(z.re) || !isnan(z.im))) {
        x = copysign(T(INFINITY), w.re) * z.re;
        y = copysign(T(INFINITY), w.re) * z.im;
    }
    return Cx<T>(x, y);
}
template <class T> inline Cx<T> operator/(T s, Cx<T> w) { return Cx<T>(s, 0) / w; }
template <class T> inline Cx<T> csqrt(Cx<T> x) {
    const T rho = sqrt(hypot(x.re, x.im)), theta = atan2(x.im, x.re) / T(2);
    return Cx<T>(rho * cos(theta), rho * sin(theta));
}

// devComplex3 / MatS33 and their operators, Internal.h:63-154
template <class T> struct V3 {
    Cx<T> x, y, z;
};
template <class T> inline V3<T> operator+(V3<T> a, V3<T> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class T> inline V3<T> operator-(V3<T> a, V3<T> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class T> inline V3<T> operator*(V3<T> a, T s) { return {a.x * s, a.y * s, a.z * s}; }
template <class T> inline V3<T> operator*(V3<T> a, Cx<T> s) { return {a.x * s, a.y * s, a.z * s}; }
template <class T> inline Cx<T> dot(V3<T> a, V3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> inline Cx<T> norm(V3<T> a) { return csqrt(dot(a, a)); }
template <class T> inline V3<T> normalized(V3<T> a) { return {a.x / norm(a), a.y / norm(a), a.z / norm(a)}; }
template <class T> inline V3<T> cross(V3<T> a, V3<T> b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <class T> struct M33 {
    V3<T> r[3];
};
template <class T> inline V3<T> operator*(const M33<T> &m, V3<T> v) { return {dot(m.r[0], v), dot(m.r[1], v), dot(m.r[2], v)}; }

// load an interleaved float pose (9 or 3 complex numbers as re,im pairs)
template <class T> inline M33<T> load_mat(const float *p) {
    M33<T> M;
    for (int r = 0; r < 3; ++r) {
        M.r[r].x = Cx<T>(p[(r * 3 + 0) * 2], p[(r * 3 + 0) * 2 + 1]);
        M.r[r].y = Cx<T>(p[(r * 3 + 1) * 2], p[(r * 3 + 1) * 2 + 1]);
        M.r[r].z = Cx<T>(p[(r * 3 + 2) * 2], p[(r * 3 + 2) * 2 + 1]);
    }
    return M;
}
template <class T> inline V3<T> load_vec(const float *p) { return {Cx<T>(p[0], p[1]), Cx<T>(p[2], p[3]), Cx<T>(p[4], p[5])}; }

// ---- storage.  Arrays are float (re, im) pairs.  With T = Dual every array is followed, `sh` floats later, by a shadow array of
// the same layout that holds the dual parts (sh is ignored by the scalar types).
template <class T> struct is_dual {
    static const bool value = false;
};
template <> struct is_dual<Dual> {
    static const bool value = true;
};
template <class T> inline T ldt(const float *p, size_t) { return (T) p[0]; }
template <> inline Dual ldt<Dual>(const float *p, size_t sh) { return Dual(p[0], p[sh]); }
template <class T> inline Cx<T> ldc(const float *p, size_t) { return Cx<T>(p[0], p[1]); }
template <> inline Cx<Dual> ldc<Dual>(const float *p, size_t sh) { return Cx<Dual>(Dual(p[0], p[sh]), Dual(p[1], p[sh + 1])); }
template <class T> inline void stc(float *p, size_t, const Cx<T> &v) { p[0] = (float) v.re, p[1] = (float) v.im; }
template <> inline void stc<Dual>(float *p, size_t sh, const Cx<Dual> &v) {
    p[0] = (float) v.re.v, p[1] = (float) v.im.v, p[sh] = (float) v.re.d, p[sh + 1] = (float) v.im.d;
}
template <class T> inline M33<T> load_mat_sh(const float *p) {  // 9 complex numbers (+ 18 floats of dual parts)
    M33<T> M;
    for (int r = 0; r < 3; ++r) {
        M.r[r].x = ldc<T>(p + (r * 3 + 0) * 2, 18);
        M.r[r].y = ldc<T>(p + (r * 3 + 1) * 2, 18);
        M.r[r].z = ldc<T>(p + (r * 3 + 2) * 2, 18);
    }
    return M;
}
template <class T> inline V3<T> load_vec_sh(const float *p) { return {ldc<T>(p, 6), ldc<T>(p + 2, 6), ldc<T>(p + 4, 6)}; }

}  // namespace xo
