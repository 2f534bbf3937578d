// xs_jet.cuh — batched complex-step numbers for device code.
//
// A Jet<C,K> carries one real value and K perturbation directions of C derivative components:
//   C = 1  CSFD   : value + eps                (cuda::std::complex<float>,  Internal.h:24)
//   C = 3  DCSFD  : value + eps1 + eps2 + eps1eps2   (d_complex<float>, cuda_double_complex.hpp:16-134,
//                   value()=re.re, grad()=re.im, imag().real()=im.re, hessian()=im.im)
// Components are h-scaled exactly like the reference's imaginary parts (H_ = 1e-7, Internal.h:33), so
// eps^2 terms (h^2 ~ 1e-14 relative) are below FP32 resolution and the complex / bicomplex product
// reduces to the truncated algebra implemented here.  The REAL part of every operation reproduces the
// rounding sequence the reference's libcu++ complex arithmetic performs when imaginary parts are zero
// (see DESIGN.md "rounding contract"):
//   complex * complex  -> one rounded product (no FMA contraction: libcu++ forms ac - bd)      jmul
//   complex * float    -> rounded product, contracted with a following add where nvcc does     jmulf / jfmaf
//   complex / complex  -> logb/scalbn pre-scaled quotient  (a*c')/(c'*c')                         jdiv
//   complex / float    -> IEEE division                                                           jdivf
//   sqrt(complex)      -> IEEE sqrt of the real part (polar(sqrt|z|, arg/2) with arg == 0)        jsqrt
// so that integer pixel / voxel indices and validity masks derived from real parts are bit-exact
// against a zero-seed reference run.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define XS_DEV __device__ __forceinline__

namespace xs {

template <int C, int K> struct Jet {
    static constexpr int N = C * K;
    float v;
    float d[N > 0 ? N : 1];
};

// ---- reference-faithful real-part helpers -------------------------------------------------
// Real part of complex<float>(a,0) / complex<float>(c,0) as libcu++ evaluates it
// (cuda/std/detail/libcxx/include/complex operator/: logb/scalbn scaling, (ac+bd)/(cc+dd)).
XS_DEV float ref_cdiv_re(float a, float c) {
    float lb = logbf(fabsf(c));
    int il = 0;
    float cs = c;
    if (isfinite(lb)) {
        il = (int) lb;
        cs = scalbnf(c, -il);
    }
    float denom = __fmul_rn(cs, cs);
    float q = __fdiv_rn(__fmul_rn(a, cs), denom);
    float x = scalbnf(q, -il);
    if (isnan(x) && denom == 0.f && !isnan(a)) x = copysignf(INFINITY, c) * a;
    return x;
}

template <int C, int K> XS_DEV Jet<C, K> jconst(float v) {
    Jet<C, K> r;
    r.v = v;
#pragma unroll
    for (int i = 0; i < Jet<C, K>::N; ++i) r.d[i] = 0.f;
    return r;
}

template <int C, int K> XS_DEV Jet<C, K> operator+(const Jet<C, K> &a, const Jet<C, K> &b) {
    Jet<C, K> r;
    r.v = __fadd_rn(a.v, b.v);
#pragma unroll
    for (int i = 0; i < Jet<C, K>::N; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
}
template <int C, int K> XS_DEV Jet<C, K> operator-(const Jet<C, K> &a, const Jet<C, K> &b) {
    Jet<C, K> r;
    r.v = __fsub_rn(a.v, b.v);
#pragma unroll
    for (int i = 0; i < Jet<C, K>::N; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
}
template <int C, int K> XS_DEV Jet<C, K> operator-(const Jet<C, K> &a) {
    Jet<C, K> r;
    r.v = -a.v;
#pragma unroll
    for (int i = 0; i < Jet<C, K>::N; ++i) r.d[i] = -a.d[i];
    return r;
}
// complex + float / complex - float / float - complex
template <int C, int K> XS_DEV Jet<C, K> jaddf(const Jet<C, K> &a, float s) {
    Jet<C, K> r = a;
    r.v = __fadd_rn(a.v, s);
    return r;
}
template <int C, int K> XS_DEV Jet<C, K> jsubf(const Jet<C, K> &a, float s) {
    Jet<C, K> r = a;
    r.v = __fsub_rn(a.v, s);
    return r;
}
template <int C, int K> XS_DEV Jet<C, K> jrsubf(float s, const Jet<C, K> &a) {  // s - a
    Jet<C, K> r;
    r.v = __fsub_rn(s, a.v);
#pragma unroll
    for (int i = 0; i < Jet<C, K>::N; ++i) r.d[i] = -a.d[i];
    return r;
}
// complex * float (rounded product)
template <int C, int K> XS_DEV Jet<C, K> jmulf(const Jet<C, K> &a, float s) {
    Jet<C, K> r;
    r.v = __fmul_rn(a.v, s);
#pragma unroll
    for (int i = 0; i < Jet<C, K>::N; ++i) r.d[i] = a.d[i] * s;
    return r;
}
// a * s + b with the real part contracted into one FMA (RayCaster.cu:91,227,238: origin + dir * time)
template <int C, int K> XS_DEV Jet<C, K> jfmaf(const Jet<C, K> &a, float s, const Jet<C, K> &b) {
    Jet<C, K> r;
    r.v = __fmaf_rn(a.v, s, b.v);
#pragma unroll
    for (int i = 0; i < Jet<C, K>::N; ++i) r.d[i] = fmaf(a.d[i], s, b.d[i]);
    return r;
}
// complex / float
template <int C, int K> XS_DEV Jet<C, K> jdivf(const Jet<C, K> &a, float s) {
    Jet<C, K> r;
    r.v = __fdiv_rn(a.v, s);
    float inv = __fdiv_rn(1.f, s);
#pragma unroll
    for (int i = 0; i < Jet<C, K>::N; ++i) r.d[i] = a.d[i] * inv;
    return r;
}

// complex * complex
template <int K> XS_DEV Jet<1, K> operator*(const Jet<1, K> &a, const Jet<1, K> &b) {
    Jet<1, K> r;
    r.v = __fmul_rn(a.v, b.v);
#pragma unroll
    for (int k = 0; k < K; ++k) r.d[k] = fmaf(a.v, b.d[k], a.d[k] * b.v);
    return r;
}
template <int K> XS_DEV Jet<3, K> operator*(const Jet<3, K> &a, const Jet<3, K> &b) {
    Jet<3, K> r;
    r.v = __fmul_rn(a.v, b.v);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float a1 = a.d[3 * k], a2 = a.d[3 * k + 1], a12 = a.d[3 * k + 2];
        const float b1 = b.d[3 * k], b2 = b.d[3 * k + 1], b12 = b.d[3 * k + 2];
        r.d[3 * k] = fmaf(a.v, b1, a1 * b.v);
        r.d[3 * k + 1] = fmaf(a.v, b2, a2 * b.v);
        r.d[3 * k + 2] = fmaf(a.v, b12, fmaf(a12, b.v, fmaf(a1, b2, a2 * b1)));
    }
    return r;
}
// float * complex is complex * float in the reference (operator*(const T&, const complex<T>&))
template <int C, int K> XS_DEV Jet<C, K> operator*(float s, const Jet<C, K> &a) { return jmulf(a, s); }

// complex / complex
template <int K> XS_DEV Jet<1, K> operator/(const Jet<1, K> &a, const Jet<1, K> &b) {
    Jet<1, K> r;
    r.v = ref_cdiv_re(a.v, b.v);
    const float inv = __fdiv_rn(1.f, b.v);
#pragma unroll
    for (int k = 0; k < K; ++k) r.d[k] = (a.d[k] - r.v * b.d[k]) * inv;
    return r;
}
template <int K> XS_DEV Jet<3, K> operator/(const Jet<3, K> &a, const Jet<3, K> &b) {
    Jet<3, K> r;
    r.v = ref_cdiv_re(a.v, b.v);
    const float inv = __fdiv_rn(1.f, b.v);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float a1 = a.d[3 * k], a2 = a.d[3 * k + 1], a12 = a.d[3 * k + 2];
        const float b1 = b.d[3 * k], b2 = b.d[3 * k + 1], b12 = b.d[3 * k + 2];
        const float q1 = (a1 - r.v * b1) * inv;
        const float q2 = (a2 - r.v * b2) * inv;
        r.d[3 * k] = q1;
        r.d[3 * k + 1] = q2;
        r.d[3 * k + 2] = (a12 - r.v * b12 - q1 * b2 - q2 * b1) * inv;
    }
    return r;
}

// sqrt
template <int K> XS_DEV Jet<1, K> jsqrt(const Jet<1, K> &a) {
    Jet<1, K> r;
    r.v = __fsqrt_rn(a.v);
    const float h = __fdiv_rn(0.5f, r.v);
#pragma unroll
    for (int k = 0; k < K; ++k) r.d[k] = a.d[k] * h;
    return r;
}
template <int K> XS_DEV Jet<3, K> jsqrt(const Jet<3, K> &a) {
    Jet<3, K> r;
    r.v = __fsqrt_rn(a.v);
    const float h = __fdiv_rn(0.5f, r.v);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float s1 = a.d[3 * k] * h, s2 = a.d[3 * k + 1] * h;
        r.d[3 * k] = s1;
        r.d[3 * k + 1] = s2;
        r.d[3 * k + 2] = (a.d[3 * k + 2] - 2.f * s1 * s2) * h;
    }
    return r;
}

// ---- 3-vectors (devComplex3 / devDComplex3, Internal.h:63-142,159-189) ---------------------
template <int C, int K> struct Jet3 {
    Jet<C, K> x, y, z;
};
template <int C, int K> XS_DEV Jet3<C, K> operator+(const Jet3<C, K> &a, const Jet3<C, K> &b) {
    return {a.x + b.x, a.y + b.y, a.z + b.z};
}
template <int C, int K> XS_DEV Jet3<C, K> operator-(const Jet3<C, K> &a, const Jet3<C, K> &b) {
    return {a.x - b.x, a.y - b.y, a.z - b.z};
}
// dot, Internal.h:75-79: ((x*x' + y*y') + z*z')
template <int C, int K> XS_DEV Jet<C, K> jdot(const Jet3<C, K> &a, const Jet3<C, K> &b) {
    return (a.x * b.x + a.y * b.y) + a.z * b.z;
}
// cross, Internal.h:139-142
template <int C, int K> XS_DEV Jet3<C, K> jcross(const Jet3<C, K> &a, const Jet3<C, K> &b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// norm, Internal.h:124-127
template <int C, int K> XS_DEV Jet<C, K> jnorm(const Jet3<C, K> &a) { return jsqrt(jdot(a, a)); }
// normalized, Internal.h:134-137 (the reference re-evaluates norm(v) per component; same value)
template <int C, int K> XS_DEV Jet3<C, K> jnormalized(const Jet3<C, K> &a) {
    const Jet<C, K> n = jnorm(a);
    return {a.x / n, a.y / n, a.z / n};
}

// ---- "fast-real" variants ------------------------------------------------------------------
// Used where a kernel re-derives derivative components per direction from a real path that was already evaluated
// bit-faithfully elsewhere: the real part here is only a coefficient of the derivative formulas (1e-7 relative is
// plenty), so the reference-faithful logb/scalbn quotient, IEEE division and IEEE sqrt are replaced by MUFU-based ones.
template <int K> XS_DEV Jet<1, K> jdiv_fast(const Jet<1, K> &a, const Jet<1, K> &b) {
    Jet<1, K> r;
    const float inv = __fdividef(1.f, b.v);
    r.v = a.v * inv;
#pragma unroll
    for (int k = 0; k < K; ++k) r.d[k] = (a.d[k] - r.v * b.d[k]) * inv;
    return r;
}
template <int K> XS_DEV Jet<3, K> jdiv_fast(const Jet<3, K> &a, const Jet<3, K> &b) {
    Jet<3, K> r;
    const float inv = __fdividef(1.f, b.v);
    r.v = a.v * inv;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float a1 = a.d[3 * k], a2 = a.d[3 * k + 1], a12 = a.d[3 * k + 2];
        const float b1 = b.d[3 * k], b2 = b.d[3 * k + 1], b12 = b.d[3 * k + 2];
        const float q1 = (a1 - r.v * b1) * inv;
        const float q2 = (a2 - r.v * b2) * inv;
        r.d[3 * k] = q1;
        r.d[3 * k + 1] = q2;
        r.d[3 * k + 2] = (a12 - r.v * b12 - q1 * b2 - q2 * b1) * inv;
    }
    return r;
}
template <int K> XS_DEV Jet<1, K> jsqrt_fast(const Jet<1, K> &a) {
    Jet<1, K> r;
    const float rs = rsqrtf(a.v);
    r.v = a.v * rs;
    const float h = 0.5f * rs;
#pragma unroll
    for (int k = 0; k < K; ++k) r.d[k] = a.d[k] * h;
    return r;
}
template <int K> XS_DEV Jet<3, K> jsqrt_fast(const Jet<3, K> &a) {
    Jet<3, K> r;
    const float rs = rsqrtf(a.v);
    r.v = a.v * rs;
    const float h = 0.5f * rs;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float s1 = a.d[3 * k] * h, s2 = a.d[3 * k + 1] * h;
        r.d[3 * k] = s1;
        r.d[3 * k + 1] = s2;
        r.d[3 * k + 2] = (a.d[3 * k + 2] - 2.f * s1 * s2) * h;
    }
    return r;
}
template <int C, int K> XS_DEV Jet3<C, K> jnormalized_fast(const Jet3<C, K> &a) {
    const Jet<C, K> n = jsqrt_fast(jdot(a, a));
    return {jdiv_fast(a.x, n), jdiv_fast(a.y, n), jdiv_fast(a.z, n)};
}

// ---- rigid transforms with derivative components -------------------------------------------
struct DevPose {  // real part, passed by value
    float R[9];
    float t[3];
};

// Loads the derivative components of direction tile [k0, k0+K) of entry e (0..11: R row-major, then t)
// from dpose[ncomp][12]; directions beyond `dirs` are zero.
template <int C, int K>
XS_DEV Jet<C, K> pose_entry(const DevPose &P, const float *__restrict__ dpose, int e, int k0, int dirs) {
    Jet<C, K> r;
    r.v = (e < 9) ? P.R[e] : P.t[e - 9];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int c = 0; c < C; ++c) r.d[k * C + c] = (k0 + k < dirs) ? dpose[((k0 + k) * C + c) * 12 + e] : 0.f;
    return r;
}

template <int C, int K> struct JetPose {
    Jet3<C, K> r0, r1, r2, t;
};
template <int C, int K>
XS_DEV JetPose<C, K> load_pose(const DevPose &P, const float *__restrict__ dpose, int k0, int dirs) {
    JetPose<C, K> J;
    J.r0 = {pose_entry<C, K>(P, dpose, 0, k0, dirs), pose_entry<C, K>(P, dpose, 1, k0, dirs),
            pose_entry<C, K>(P, dpose, 2, k0, dirs)};
    J.r1 = {pose_entry<C, K>(P, dpose, 3, k0, dirs), pose_entry<C, K>(P, dpose, 4, k0, dirs),
            pose_entry<C, K>(P, dpose, 5, k0, dirs)};
    J.r2 = {pose_entry<C, K>(P, dpose, 6, k0, dirs), pose_entry<C, K>(P, dpose, 7, k0, dirs),
            pose_entry<C, K>(P, dpose, 8, k0, dirs)};
    J.t = {pose_entry<C, K>(P, dpose, 9, k0, dirs), pose_entry<C, K>(P, dpose, 10, k0, dirs),
           pose_entry<C, K>(P, dpose, 11, k0, dirs)};
    return J;
}
// One direction (K = 1), all 12 entries of its C derivative components with three 128-bit loads per component
// (dpose rows are 48 bytes, 16-byte aligned).
template <int C> XS_DEV JetPose<C, 1> load_pose_vec(const DevPose &P, const float *__restrict__ dpose, int q) {
    float e[C][12];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float4 *p4 = reinterpret_cast<const float4 *>(dpose + (size_t) (q * C + c) * 12);
        const float4 a = __ldg(p4), b = __ldg(p4 + 1), d = __ldg(p4 + 2);
        e[c][0] = a.x, e[c][1] = a.y, e[c][2] = a.z, e[c][3] = a.w;
        e[c][4] = b.x, e[c][5] = b.y, e[c][6] = b.z, e[c][7] = b.w;
        e[c][8] = d.x, e[c][9] = d.y, e[c][10] = d.z, e[c][11] = d.w;
    }
    JetPose<C, 1> J;
    Jet<C, 1> *out[12] = {&J.r0.x, &J.r0.y, &J.r0.z, &J.r1.x, &J.r1.y, &J.r1.z, &J.r2.x, &J.r2.y, &J.r2.z, &J.t.x, &J.t.y, &J.t.z};
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        out[i]->v = (i < 9) ? P.R[i] : P.t[i - 9];
#pragma unroll
        for (int c = 0; c < C; ++c) out[i]->d[c] = e[c][i];
    }
    return J;
}
// The same for an explicit list of component indices (Hessian batches: (F_i, F_j, S_ij) are not consecutive planes).
template <int C> XS_DEV JetPose<C, 1> load_pose_comps(const DevPose &P, const float *__restrict__ dpose, const int (&comp)[C]) {
    float e[C][12];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float4 *p4 = reinterpret_cast<const float4 *>(dpose + (size_t) comp[c] * 12);
        const float4 a = __ldg(p4), b = __ldg(p4 + 1), d = __ldg(p4 + 2);
        e[c][0] = a.x, e[c][1] = a.y, e[c][2] = a.z, e[c][3] = a.w;
        e[c][4] = b.x, e[c][5] = b.y, e[c][6] = b.z, e[c][7] = b.w;
        e[c][8] = d.x, e[c][9] = d.y, e[c][10] = d.z, e[c][11] = d.w;
    }
    JetPose<C, 1> J;
    Jet<C, 1> *out[12] = {&J.r0.x, &J.r0.y, &J.r0.z, &J.r1.x, &J.r1.y, &J.r1.z, &J.r2.x, &J.r2.y, &J.r2.z, &J.t.x, &J.t.y, &J.t.z};
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        out[i]->v = (i < 9) ? P.R[i] : P.t[i - 9];
#pragma unroll
        for (int c = 0; c < C; ++c) out[i]->d[c] = e[c][i];
    }
    return J;
}
// MatS33 * devComplex3, Internal.h:150-154
template <int C, int K> XS_DEV Jet3<C, K> jrot(const JetPose<C, K> &P, const Jet3<C, K> &v) {
    return {jdot(P.r0, v), jdot(P.r1, v), jdot(P.r2, v)};
}

}  // namespace xs
