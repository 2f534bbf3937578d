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

namespace xo {

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
    const T lb = std::logb(std::fmax(std::fabs(c), std::fabs(d)));
    if (std::isfinite(lb)) {
        il = (int) lb;
        c = std::scalbn(c, -il);
        d = std::scalbn(d, -il);
    }
    const T denom = c * c + d * d;
    T x = std::scalbn((z.re * c + z.im * d) / denom, -il);
    T y = std::scalbn((z.im * c - z.re * d) / denom, -il);
    if (std::isnan(x) && std::isnan(y) && denom == T(0) && (!std::isnan(z.re) || !std::isnan(z.im))) {
        x = std::copysign(T(INFINITY), w.re) * z.re;
        y = std::copysign(T(INFINITY), w.re) * z.im;
    }
    return Cx<T>(x, y);
}
template <class T> inline Cx<T> operator/(T s, Cx<T> w) { return Cx<T>(s, 0) / w; }
template <class T> inline Cx<T> csqrt(Cx<T> x) {
    const T rho = std::sqrt(std::hypot(x.re, x.im)), theta = std::atan2(x.im, x.re) / T(2);
    return Cx<T>(rho * std::cos(theta), rho * std::sin(theta));
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

}  // namespace xo
