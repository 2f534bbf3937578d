// Test harness (CPU only): the pose chain of the frame loop (KinectFusionReconstruction.cpp:167-173,231,248-258,305-320)
//   c2w = inverse(world2camera); Rprev_inv = inverse(rotation(c2w)); c2v = world2volume * c2w; v2c = inverse(c2v)
// evaluated twice on the same inputs: with the product's host jets (x-slam_b200/csrc/host_jet.h, one CSFD component) and with
// the oracle's restatement of the reference's Eigen / std::complex<float> semantics (oracle/host_algebra.h, the perturbation
// in the imaginary part).  Prints "real_jet real_oracle deriv_jet imag_oracle" per entry.  Driven by tests/test_host_jet.py.
#include "../../oracle/host_algebra.h"
#include "../../x-slam_b200/csrc/host_jet.h"

#include <cstdio>

using namespace xs;

int main() {
    hj_ctx().comps = 1;
    hj_ctx().dirs = 1;
    float w2c[2][16], w2v[16];
    for (int q = 0; q < 2; ++q)
        for (int e = 0; e < 16; ++e)
            if (scanf("%f", &w2c[q][e]) != 1) return 1;
    for (int e = 0; e < 16; ++e)
        if (scanf("%f", &w2v[e]) != 1) return 1;
    HMat4 J, JV = HMat4::identity();
    xo::Mat4c C, CV;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            J.m[i][j].v = w2c[0][i * 4 + j];
            J.m[i][j].d[0] = w2c[1][i * 4 + j];
            C.m[i][j] = xo::cf(w2c[0][i * 4 + j], w2c[1][i * 4 + j]);
            JV.m[i][j] = HJet(w2v[i * 4 + j]);
            CV.m[i][j] = xo::cf(w2v[i * 4 + j], 0.f);
        }
    const HMat4 j_c2w = hinverse(J), j_c2v = hmul(JV, j_c2w), j_v2c = hinverse(j_c2v);
    const HMat3 j_rinv = hinverse(hrotation(j_c2w));
    const xo::Mat4c c_c2w = xo::inverse(C), c_c2v = xo::mul(CV, c_c2w), c_v2c = xo::inverse(c_c2v);
    const xo::Mat3c c_rinv = xo::inverse(xo::rotation_of(c_c2w));
    const HMat4 *jm[3] = {&j_c2w, &j_c2v, &j_v2c};
    const xo::Mat4c *cm[3] = {&c_c2w, &c_c2v, &c_v2c};
    for (int m = 0; m < 3; ++m)
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j)
                printf("%.9g %.9g %.9g %.9g\n", jm[m]->m[i][j].v, cm[m]->m[i][j].real(), jm[m]->m[i][j].d[0], cm[m]->m[i][j].imag());
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            printf("%.9g %.9g %.9g %.9g\n", j_rinv.m[i][j].v, c_rinv.m[i][j].real(), j_rinv.m[i][j].d[0], c_rinv.m[i][j].imag());
    return 0;
}
