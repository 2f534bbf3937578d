// dc_array.cu — elementwise second-order-complex (bicomplex) arithmetic on packed-SoA device arrays.
//
// Device counterpart of the host DoubleComplex type (DeviceArray/include/DoubleComplex.h:15-95,
// DeviceArray/src/DoubleComplex.cpp) and of Experiments/test_CSFD's DCSFD chain (main.cpp:194-205).
// An array of n bicomplex numbers is stored as four planes float[4][n] = value | eps1 | eps2 | eps1eps2
// (re.re | re.im | im.re | im.im); each thread handles four consecutive elements with one 128-bit load per
// plane, so a warp moves 4 x 512 contiguous bytes per instruction group.  The operations restate
// DoubleComplex.cpp line by line on complex<float> pairs.  Deviation, documented in DESIGN.md: the
// reference's atanh evaluates log(a - a) (DoubleComplex.cpp:372-377), which makes atan / atan2 return
// non-finite values; here atanh uses log(a - x).
#include "xs_common.cuh"

#include <cuda/std/complex>

namespace xs {

typedef cuda::std::complex<float> cfl;
struct B2 {
    cfl re, im;
};

XS_DEV B2 b2(cfl re, cfl im) { return {re, im}; }
XS_DEV B2 operator+(B2 a, B2 b) { return {a.re + b.re, a.im + b.im}; }
XS_DEV B2 operator-(B2 a, B2 b) { return {a.re - b.re, a.im - b.im}; }
XS_DEV B2 operator*(B2 a, B2 b) { return {a.re * b.re - a.im * b.im, a.im * b.re + a.re * b.im}; }  // .cpp:159-166
XS_DEV B2 operator*(B2 a, float s) { return {a.re * s, a.im * s}; }
XS_DEV cfl b2_norm(B2 x) { return x.re * x.re + x.im * x.im; }                                       // .cpp:320-323
XS_DEV B2 operator/(B2 a, B2 b) {                                                                    // .cpp:168-175
    const cfl r = a.re * b.re + a.im * b.im, n = b2_norm(b);
    return {r / n, (a.im * b.re - a.re * b.im) / n};
}
XS_DEV cfl b2_abs(B2 x) { return cuda::std::sqrt(x.re * x.re + x.im * x.im); }  // .cpp:303-307
XS_DEV B2 b2_sqrt(B2 x) {                                                        // .cpp:332-349
    B2 result = x;
    const cfl r = b2_abs(x), sqrt_r = cuda::std::sqrt(r);
    result.re += r;
    const cfl zrnorm = b2_abs(result);
    if (fabsf(zrnorm.real()) < 1e-20f && fabsf(zrnorm.imag()) < 1e-20f) return {result.re * sqrt_r, result.im * sqrt_r};
    const cfl scale = sqrt_r / zrnorm;
    return {result.re * scale, result.im * scale};
}
XS_DEV B2 b2_exp(B2 x) {  // .cpp:351-356
    const cfl e = cuda::std::exp(x.re);
    return {e * cuda::std::cos(x.im), e * cuda::std::sin(x.im)};
}
XS_DEV cfl c_atan2(cfl y, cfl x) {  // .cpp:386-401 (the comparison of a complex with 0 uses operator> on real parts)
    cfl r = cuda::std::sqrt(x * x + y * y);
    if (r.real() > 0.0f) {
        r += x;
        r = y / r;
    } else {
        r -= x;
        r = r / y;
    }
    r = cuda::std::atan(r);
    r *= 2.0f;
    return r;
}
XS_DEV B2 b2_log(B2 x) { return {cuda::std::log(b2_abs(x)), c_atan2(x.im, x.re)}; }  // .cpp:358-366
XS_DEV B2 b2_sin(B2 x) {                                                              // .cpp:421-426
    return {cuda::std::cosh(-x.im) * cuda::std::sin(x.re), -cuda::std::sinh(-x.im) * cuda::std::cos(x.re)};
}
XS_DEV B2 b2_cos(B2 x) {  // .cpp:428-433
    return {cuda::std::cosh(-x.im) * cuda::std::cos(x.re), cuda::std::sinh(-x.im) * cuda::std::sin(x.re)};
}
XS_DEV B2 b2_polar(cfl rho, cfl theta) { return {rho * cuda::std::cos(theta), rho * cuda::std::sin(theta)}; }  // .cpp:325-330
XS_DEV B2 b2_pow(B2 x, float y) {                                                                                // .cpp:435-440
    const B2 r = b2_log(x);
    return b2_polar(cuda::std::exp(y * r.re), y * r.im);
}
XS_DEV B2 b2_atanh(B2 x) {  // .cpp:372-377 with the a - x fix
    const B2 a = {cfl(1.f, 0.f), cfl(0.f, 0.f)};
    return (b2_log(a + x) - b2_log(a - x)) * 0.5f;
}
XS_DEV B2 b2_atan(B2 x) {  // .cpp:379-384
    B2 r = {-x.im, x.re};
    r = b2_atanh(r);
    return {r.im, -r.re};
}
XS_DEV B2 b2_atan2(B2 y, B2 x) {  // .cpp:403-419
    B2 r = b2_sqrt(x * x + y * y);
    if (r.re.real() > 0.0f) {
        r = r + x;
        r = y / r;
    } else {
        r = r - x;
        r = r / y;
    }
    r = b2_atan(r);
    return r * 2.0f;
}

XS_DEV B2 apply_op(int op, B2 x, B2 y, float p) {
    switch (op) {
        case XS_DC_ADD: return x + y;
        case XS_DC_SUB: return x - y;
        case XS_DC_MUL: return x * y;
        case XS_DC_DIV: return x / y;
        case XS_DC_SQRT: return b2_sqrt(x);
        case XS_DC_EXP: return b2_exp(x);
        case XS_DC_LOG: return b2_log(x);
        case XS_DC_SIN: return b2_sin(x);
        case XS_DC_COS: return b2_cos(x);
        case XS_DC_ATAN2: return b2_atan2(x, y);
        case XS_DC_POW: return b2_pow(x, p);
        default: return b2_atan(x);
    }
}

struct Quad {
    float4 p[4];  // plane-major: p[c] = component c of 4 consecutive elements
};
XS_DEV Quad load_quad(const float *__restrict__ a, long n, long i4) {
    Quad q;
#pragma unroll
    for (int c = 0; c < 4; ++c) q.p[c] = __ldg(reinterpret_cast<const float4 *>(a + (size_t) c * n) + i4);
    return q;
}
XS_DEV B2 quad_get(const Quad &q, int j) {
    const float *v = reinterpret_cast<const float *>(q.p);
    return {cfl(v[j], v[4 + j]), cfl(v[8 + j], v[12 + j])};
}
XS_DEV void quad_set(Quad &q, int j, B2 x) {
    float *v = reinterpret_cast<float *>(q.p);
    v[j] = x.re.real();
    v[4 + j] = x.re.imag();
    v[8 + j] = x.im.real();
    v[12 + j] = x.im.imag();
}

// n must be a multiple of 4 and the planes 16-byte aligned (checked by the launcher); the tail is handled
// by the scalar kernel below.
__global__ void __launch_bounds__(256) dc_apply_vec4(int op, const float *__restrict__ a, const float *__restrict__ b, float p,
                                                     float *__restrict__ out, long n, long n4) {
    for (long i4 = (long) blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += (long) gridDim.x * blockDim.x) {
        const Quad qa = load_quad(a, n, i4);
        Quad qb = qa, qo;
        if (b) qb = load_quad(b, n, i4);
#pragma unroll
        for (int j = 0; j < 4; ++j) quad_set(qo, j, apply_op(op, quad_get(qa, j), quad_get(qb, j), p));
#pragma unroll
        for (int c = 0; c < 4; ++c) reinterpret_cast<float4 *>(out + (size_t) c * n)[i4] = qo.p[c];
    }
}
__global__ void dc_apply_scalar(int op, const float *__restrict__ a, const float *__restrict__ b, float p,
                                float *__restrict__ out, long n, long begin) {
    const long i = begin + (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const B2 x = {cfl(a[i], a[n + i]), cfl(a[2 * n + i], a[3 * n + i])};
    const B2 y = b ? B2{cfl(b[i], b[n + i]), cfl(b[2 * n + i], b[3 * n + i])} : x;
    const B2 r = apply_op(op, x, y, p);
    out[i] = r.re.real();
    out[n + i] = r.re.imag();
    out[2 * n + i] = r.im.real();
    out[3 * n + i] = r.im.imag();
}

// test_CSFD main.cpp:194-205: t = ((t,h),(h,0)); x = t*t; y = sin(t); loss = (x+y)*(x+y)
__global__ void __launch_bounds__(256) dc_chain_kernel(const float *__restrict__ t, float h, float *__restrict__ out, long n) {
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
        const B2 tt = {cfl(t[i], h), cfl(h, 0.f)};
        const B2 x = tt * tt, y = b2_sin(tt);
        const B2 l = (x + y) * (x + y);
        out[i] = l.re.real();
        out[n + i] = l.re.imag();
        out[2 * n + i] = l.im.real();
        out[3 * n + i] = l.im.imag();
    }
}

// AoS (the host DoubleComplex object: re.re, re.im, im.re, im.im - DoubleComplex.h:15-19) <-> packed-SoA planes
__global__ void dc_aos_to_soa_kernel(const float4 *__restrict__ aos, float *__restrict__ soa, long n) {
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
        const float4 v = aos[i];
        soa[i] = v.x, soa[n + i] = v.y, soa[2 * n + i] = v.z, soa[3 * n + i] = v.w;
    }
}
__global__ void dc_soa_to_aos_kernel(const float *__restrict__ soa, float4 *__restrict__ aos, long n) {
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x)
        aos[i] = make_float4(soa[i], soa[n + i], soa[2 * n + i], soa[3 * n + i]);
}

}  // namespace xs

// Owning packed-SoA bicomplex array with the semantics of the reference's DeviceArray<T> (device_array.hpp:25-134,
// device_memory.cpp:72-178): create(n) is a no-op when the size is unchanged, upload / download block and synchronise,
// copyTo is a deep copy that (re)creates the destination, release frees.
struct xs_dc_array {
    float *d = nullptr;      // float[4][n]: value | eps1 | eps2 | eps1eps2
    float4 *stage = nullptr;  // device staging for the AoS <-> SoA conversion
    long n = 0;
};

using namespace xs;

extern "C" {

xs_dc_array *xs_dc_array_create(long n) {
    int ndev = 0;
    if (n < 0 || cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("xs_dc_array_create: no CUDA device (libxslam_b200 has no CPU fallback) or negative size");
        return nullptr;
    }
    xs_dc_array *a = new xs_dc_array();
    if (xs_dc_array_resize(a, n) != XS_OK) {
        delete a;
        return nullptr;
    }
    return a;
}
void xs_dc_array_release(xs_dc_array *a) {
    if (!a) return;
    cudaFree(a->d);
    cudaFree(a->stage);
    delete a;
}
int xs_dc_array_resize(xs_dc_array *a, long n) {  // DeviceArray::create
    if (!a || n < 0) return XS_ERR_ARG;
    if (n == a->n) return XS_OK;
    cudaFree(a->d);
    cudaFree(a->stage);
    a->d = nullptr, a->stage = nullptr, a->n = 0;
    if (n > 0) {
        XS_CUDA(cudaMalloc(&a->d, (size_t) n * 4 * sizeof(float)));
        XS_CUDA(cudaMalloc(&a->stage, (size_t) n * sizeof(float4)));
    }
    a->n = n;
    return XS_OK;
}
long xs_dc_array_size(const xs_dc_array *a) { return a ? a->n : 0; }
float *xs_dc_array_ptr(xs_dc_array *a) { return a ? a->d : nullptr; }
int xs_dc_array_upload(xs_dc_array *a, const float *host_aos, long n) {  // DeviceArray::upload (create + blocking copy)
    if (!a || (n > 0 && !host_aos)) return XS_ERR_ARG;
    const int rc = xs_dc_array_resize(a, n);
    if (rc != XS_OK || n == 0) return rc;
    XS_CUDA(cudaMemcpy(a->stage, host_aos, (size_t) n * sizeof(float4), cudaMemcpyHostToDevice));
    const int grid = (int) (n / 256 + 1 < sm_count() * 8 ? n / 256 + 1 : sm_count() * 8);
    dc_aos_to_soa_kernel<<<grid, 256>>>(a->stage, a->d, n);
    XS_LAUNCH_CHECK();
    XS_CUDA(cudaDeviceSynchronize());  // device_memory.cpp:141-146 synchronises after the copy
    return XS_OK;
}
int xs_dc_array_download(const xs_dc_array *a, float *host_aos) {  // DeviceArray::download
    if (!a || (a->n > 0 && !host_aos)) return XS_ERR_ARG;
    if (a->n == 0) return XS_OK;
    const int grid = (int) (a->n / 256 + 1 < sm_count() * 8 ? a->n / 256 + 1 : sm_count() * 8);
    dc_soa_to_aos_kernel<<<grid, 256>>>(a->d, a->stage, a->n);
    XS_LAUNCH_CHECK();
    XS_CUDA(cudaMemcpy(host_aos, a->stage, (size_t) a->n * sizeof(float4), cudaMemcpyDeviceToHost));
    XS_CUDA(cudaDeviceSynchronize());
    return XS_OK;
}
int xs_dc_array_copy(const xs_dc_array *src, xs_dc_array *dst) {  // DeviceArray::copyTo
    if (!src || !dst) return XS_ERR_ARG;
    const int rc = xs_dc_array_resize(dst, src->n);
    if (rc != XS_OK || src->n == 0) return rc;
    XS_CUDA(cudaMemcpy(dst->d, src->d, (size_t) src->n * 4 * sizeof(float), cudaMemcpyDeviceToDevice));
    XS_CUDA(cudaDeviceSynchronize());
    return XS_OK;
}

int xs_dc_apply(int op, const float *d_a, const float *d_b, float p, float *d_out, long n, void *stream) {
    if (!d_a || !d_out || n <= 0 || op < 0 || op > XS_DC_ATAN) return XS_ERR_ARG;
    const bool binary = (op <= XS_DC_DIV) || op == XS_DC_ATAN2;
    if (binary && !d_b) return XS_ERR_ARG;
    cudaStream_t s = (cudaStream_t) stream;
    const bool aligned = (n % 4 == 0) && ((size_t) d_a % 16 == 0) && ((size_t) d_out % 16 == 0) && (!d_b || (size_t) d_b % 16 == 0);
    long done = 0;
    if (aligned) {
        const long n4 = n / 4;
        const int grid = (int) (n4 / 256 + 1 < sm_count() * 8 ? n4 / 256 + 1 : sm_count() * 8);
        dc_apply_vec4<<<grid, 256, 0, s>>>(op, d_a, binary ? d_b : nullptr, p, d_out, n, n4);
        XS_LAUNCH_CHECK();
        done = n;
    }
    if (done < n) {
        dc_apply_scalar<<<(unsigned) ((n - done + 255) / 256), 256, 0, s>>>(op, d_a, binary ? d_b : nullptr, p, d_out, n, done);
        XS_LAUNCH_CHECK();
    }
    return XS_OK;
}

int xs_dc_chain(const float *d_t, float h, float *d_out, long n, void *stream) {
    if (!d_t || !d_out || n <= 0) return XS_ERR_ARG;
    const int grid = (int) (n / 256 + 1 < sm_count() * 8 ? n / 256 + 1 : sm_count() * 8);
    dc_chain_kernel<<<grid, 256, 0, (cudaStream_t) stream>>>(d_t, h, d_out, n);
    XS_LAUNCH_CHECK();
    return XS_OK;
}

}  // extern "C"
