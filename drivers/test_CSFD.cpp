// test_CSFD — the bicomplex demo of Experiments/test_CSFD/main.cpp:88-219 on the packed-SoA device arrays.
//   1. complex-step arithmetic on arrays of n = 10^6 bicomplex numbers (value | eps1 | eps2 | eps1eps2 planes):
//      * / exp sin pow at a = (0.5, h), b = (-1.5, h) — the five value pairs the reference prints (main.cpp:90-191) —
//      timed per launch (the reference times 10^6 scalar host calls);
//   2. the DCSFD chain-rule self check (main.cpp:194-219): t = ((0.5, h), (h, 0)), x = t*t, y = sin t,
//      loss = (x + y)^2; gradient = loss.eps1 / h, second order = loss.eps1eps2 / h^2, next to the analytic chain rule.
// Host C++ over the C-ABI (xs_dc_apply, xs_dc_chain); fails loudly without a CUDA device.
#include "../include/xslam_b200.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <iostream>
#include <vector>

int xs_driver_device_alloc(float **p, size_t floats);
int xs_driver_device_download(float *dst, const float *src, size_t floats);
int xs_driver_device_upload(float *dst, const float *src, size_t floats);
void xs_driver_device_free(float *p);

int main() {
    const long n = 1000000;  // launch_number, main.cpp:94
    const float h = 1e-6f;   // main.cpp:93
    float *d_a, *d_b, *d_o;
    if (xs_driver_device_alloc(&d_a, 4 * n) || xs_driver_device_alloc(&d_b, 4 * n) || xs_driver_device_alloc(&d_o, 4 * n)) {
        std::cerr << "test_CSFD needs a CUDA device (libxslam_b200 has no CPU fallback)\n";
        return -1;
    }
    std::vector<float> a(4 * n, 0.f), b(4 * n, 0.f), o(4 * n);
    for (long i = 0; i < n; ++i) {
        a[i] = 0.5f, a[n + i] = h;   // a = (0.5, h): value plane, eps1 plane
        b[i] = -1.5f, b[n + i] = h;  // b = (-1.5, h)
    }
    xs_driver_device_upload(d_a, a.data(), 4 * n);
    xs_driver_device_upload(d_b, b.data(), 4 * n);
    std::cout << "1. simple test for complex acceleration (" << n << " bicomplex elements per launch)" << std::endl;
    // c = a + b = (-1, 2h): the argument the reference feeds to exp and sin (main.cpp:130-171)
    float *d_c;
    if (xs_driver_device_alloc(&d_c, 4 * n) || xs_dc_apply(XS_DC_ADD, d_a, d_b, 0.f, d_c, n, nullptr) != XS_OK) {
        std::cerr << "xs_dc_apply failed: " << xs_last_error() << "\n";
        return -1;
    }
    struct Op {
        const char *name;
        int op;
        float p;
        const float *x;
    } ops[] = {{"multiplication a*b", XS_DC_MUL, 0, d_a}, {"division a/b", XS_DC_DIV, 0, d_a}, {"exp(a+b)", XS_DC_EXP, 0, d_c},
               {"sin(a+b)", XS_DC_SIN, 0, d_c}, {"pow(a,3)", XS_DC_POW, 3, d_a}};
    for (const Op &op : ops) {
        if (xs_dc_apply(op.op, op.x, d_b, op.p, d_o, n, nullptr) != XS_OK) {  // warm-up
            std::cerr << "xs_dc_apply failed: " << xs_last_error() << "\n";
            return -1;
        }
        xs_driver_device_download(o.data(), d_o, 4 * n);
        const auto t0 = std::chrono::steady_clock::now();
        const int reps = 20;
        for (int r = 0; r < reps; ++r) xs_dc_apply(op.op, op.x, d_b, op.p, d_o, n, nullptr);
        xs_driver_device_download(o.data(), d_o, 4);  // synchronises
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / reps;
        std::cout << "run our " << op.name << std::endl;
        printf("mean compute time = %.3f ms per %ld elements\n", ms, n);
        std::cout << "result = (" << o[0] << "," << o[n] << ")" << std::endl;
    }
    std::cout << "2. test for DCSFD" << std::endl;
    std::vector<float> t(n, 0.5f);
    xs_driver_device_upload(d_a, t.data(), n);
    if (xs_dc_chain(d_a, h, d_o, n, nullptr) != XS_OK) {
        std::cerr << "xs_dc_chain failed: " << xs_last_error() << "\n";
        return -1;
    }
    xs_driver_device_download(o.data(), d_o, 4 * n);
    std::cout << "DCSFD result: " << std::endl;
    std::cout << "gradient = " << o[n] / h << std::endl;
    std::cout << "second order differentiation = " << o[3 * n] / h / h << std::endl;
    // chain rule: f = (x + y)^2, x = t^2, y = sin t
    const double tt = 0.5, x = tt * tt, y = std::sin(tt), dx = 2 * tt, dy = std::cos(tt);
    std::cout << "chain rule result: " << std::endl;
    std::cout << "gradient = " << 2 * (x + y) * (dx + dy) << std::endl;
    std::cout << "second order differentiation = " << 2 * (dx + dy) * (dx + dy) + 2 * (x + y) * (2 - std::sin(tt)) << std::endl;
    xs_driver_device_free(d_a), xs_driver_device_free(d_b), xs_driver_device_free(d_o), xs_driver_device_free(d_c);
    return 0;
}
