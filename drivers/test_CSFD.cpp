// test_CSFD — the DCSFD demo of Experiments/test_CSFD/main.cpp:88-219, written against this repo's number types the way the
// reference writes it against its own:
//   1. "simple test for complex acceleration" (main.cpp:90-191): the five complex-step operations * / exp sin pow at
//      a = (0.5, h), b = (-1.5, h) - printed as value pairs ("ours", i.e. the truncated first-order form, beside std::complex)
//      - and, as the B200 counterpart of the reference's 10^6-iteration host loops, the same operations on packed-SoA device
//      arrays of 10^6 bicomplex elements (xslam_b200::DeviceArray4 over xs_dc_apply), timed per launch;
//   2. "test high order chain rule" (main.cpp:193-219): t = ((0.5, h), (h, 0)), x = t*t, y = sin t, loss = f1(x, y) with the
//      host scalar xslam_b200::DoubleComplex: gradient = loss.real().imag() / h, second order = loss.imag().imag() / h^2,
//      next to the chain rule assembled from partial derivatives exactly as the reference does; the same chain evaluated on
//      the device (xs_dc_chain) must print the same two numbers.
// Host C++ over the C-ABI; the device part fails loudly without a CUDA device (no CPU fallback for device arrays).
#include "../include/xslam_dcomplex.hpp"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <iostream>
#include <vector>

using xslam_b200::DeviceArray4;
using xslam_b200::DoubleComplex;
using xslam_b200::MyFloat;
using xslam_b200::SingleComplex;

static DoubleComplex f1(DoubleComplex x, DoubleComplex y) { return (x + y) * (x + y); }  // main.cpp:9-12

// first-order ("ours") forms of main.cpp:17-86: the real part ignores h^2 terms
static SingleComplex mul_first_order(SingleComplex a, SingleComplex b) { return {a.real() * b.real(), a.imag() * b.real() + a.real() * b.imag()}; }
static SingleComplex div_first_order(SingleComplex a, SingleComplex b) {
    return {a.real() / b.real(), (a.imag() * b.real() - a.real() * b.imag()) / (b.real() * b.real() + b.imag() * b.imag())};
}
static SingleComplex exp_first_order(SingleComplex a) { return {std::exp(a.real()), std::exp(a.real()) * std::sin(a.imag())}; }
static SingleComplex sin_first_order(SingleComplex a) { return {std::sin(a.real()), -std::sinh(-a.imag()) * std::cos(a.real())}; }
static SingleComplex pow_first_order(SingleComplex a, int n) { return {(float) std::pow(a.real(), n), (float) (std::pow(std::norm(a), n) * std::sin(n * std::arg(a)))}; }

int main() {
    const float h = 1e-6f;       // main.cpp:93
    const long n = 1000000;      // launch_number, main.cpp:94
    std::cout << "1. simple test for complex acceleration" << std::endl;
    const SingleComplex a(0.5f, h), b(-1.5f, h);
    std::cout << "multiplication value: " << mul_first_order(a, b) << "\t" << a * b << std::endl;
    std::cout << "division value: " << div_first_order(a, b) << "\t" << a / b << std::endl;
    std::cout << "exp value: " << exp_first_order(a + b) << "\t" << std::exp(a + b) << std::endl;
    std::cout << "sin value: " << sin_first_order(a + b) << "\t" << std::sin(a + b) << std::endl;
    std::cout << "pow value: " << pow_first_order(a + b, 3) << "\t" << std::pow(a + b, 3) << std::endl;

    // the same operations on device arrays of n bicomplex elements (value | eps1 | eps2 | eps1eps2 planes)
    std::vector<DoubleComplex> ha(n, DoubleComplex(0.5f, h, 0.f, 0.f)), hb(n, DoubleComplex(-1.5f, h, 0.f, 0.f)), ho;
    DeviceArray4 da, db, dc, dout;
    if (!da.upload(ha) || !db.upload(hb) || !da.apply(XS_DC_ADD, &db, 0.f, dc)) {
        std::cerr << "test_CSFD needs a CUDA device (libxslam_b200 has no CPU fallback): " << xs_last_error() << "\n";
        return -1;
    }
    struct Op {
        const char *name;
        xs_dc_op op;
        float p;
        const DeviceArray4 *x;
    } ops[] = {{"multiplication", XS_DC_MUL, 0, &da}, {"division", XS_DC_DIV, 0, &da}, {"exp", XS_DC_EXP, 0, &dc}, {"sin", XS_DC_SIN, 0, &dc},
               {"pow", XS_DC_POW, 3, &da}};  // pow on a = (0.5, h): the bicomplex log of a negative base is NaN in the reference too (atan2, DoubleComplex.cpp:386-401)
    for (const Op &op : ops) {
        if (!op.x->apply(op.op, &db, op.p, dout) || !dout.download(ho)) {  // warm-up launch + result
            std::cerr << "xs_dc_apply failed: " << xs_last_error() << "\n";
            return -1;
        }
        const int reps = 20;
        const auto t0 = std::chrono::steady_clock::now();
        for (int r = 0; r < reps; ++r) op.x->apply(op.op, &db, op.p, dout);
        DeviceArray4 probe;
        dout.copyTo(probe);  // blocking: the launches above have completed
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / reps;
        std::cout << "run our " << op.name << " on the device" << std::endl;
        printf("mean compute time = %.3f ms per %ld elements\n", ms, n);
        std::cout << "value: " << ho[0].real() << std::endl;
    }

    // main.cpp:193-219
    std::cout << "2. test high order chain rule by compute f1(x,y)=(x+y)^2, x=t*t,  y=sin(t)" << std::endl;
    DoubleComplex t(0.5f);
    t.addPerturbation();  // t + i h + j h, h = 1e-6 (DoubleComplex.cpp:61-66)
    const DoubleComplex x = t * t, y = sin(t), loss = f1(x, y);
    std::cout << "a. compute gradient and second order differentiation by DCSFD" << std::endl;
    std::cout << "gradient = " << loss.grad() / h << std::endl;
    std::cout << "second order differentiation = " << loss.hessian() / h / h << std::endl;
    // b. the chain rule from partial derivatives, each obtained by seeding one argument of f1 (main.cpp:206-219)
    const float xv = x.value(), yv = y.value();
    const DoubleComplex both_x(xv, h, h, 0), both_y(yv, h, h, 0), plain_x(xv), plain_y(yv);
    const float fx = f1(both_x, plain_y).grad() / h, fy = f1(plain_x, both_y).grad() / h;
    const float fxx = f1(both_x, plain_y).hessian() / h / h, fyy = f1(plain_x, both_y).hessian() / h / h;
    const float fxy = f1(DoubleComplex(xv, h, 0, 0), DoubleComplex(yv, 0, h, 0)).hessian() / h / h;
    const float xt = x.grad() / h, xtt = x.hessian() / h / h, yt = y.grad() / h, ytt = y.hessian() / h / h;
    std::cout << "b. compute gradient and second order differentiation by chain rule" << std::endl;
    std::cout << "gradient = " << fx * xt + fy * yt << std::endl;
    std::cout << "second order differentiation = " << fx * xtt + fy * ytt + xt * xt * fxx + yt * yt * fyy + 2 * xt * yt * fxy << std::endl;
    // c. the same chain on the device arrays (xs_dc_chain takes the real parts of t and seeds them itself)
    DeviceArray4 dt_, dl;
    std::vector<DoubleComplex> ht(n, DoubleComplex(0.5f)), hl;
    if (!dt_.upload(ht) || !dl.create(n) || xs_dc_chain(dt_.ptr(), h, dl.ptr(), n, nullptr) != XS_OK || !dl.download(hl)) {
        std::cerr << "xs_dc_chain failed: " << xs_last_error() << "\n";
        return -1;
    }
    std::cout << "c. the same chain on " << n << " device elements (packed-SoA DeviceArray4)" << std::endl;
    std::cout << "gradient = " << hl[n / 2].grad() / h << std::endl;
    std::cout << "second order differentiation = " << hl[n / 2].hessian() / h / h << std::endl;
    return 0;
}
