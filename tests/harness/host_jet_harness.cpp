// Test harness (CPU only): evaluates the host jet algebra of x-slam_b200/csrc/host_jet.h - the pose algebra the frame loop
// runs between ICP and integration (KinectFusionReconstruction.cpp:167-173,231,248-258,305-320) - on inputs read from stdin
// and prints every component.  Built and driven by tests/test_host_jet.py.
//   stdin : comps dirs, then for A and B: (1 + ncomp) x 16 floats (component 0 = real 4x4, row-major), then one angle jet
//           (1 + ncomp floats)
//   stdout: inverse(A), A*B, inverse(rotation(A)), Rz(angle)*Ry(angle)*Rx(angle): each (1 + ncomp) x (16 | 9) floats
#include "../../x-slam_b200/csrc/host_jet.h"

#include <cstdio>
#include <vector>

using namespace xs;

static HMat4 read4(int ncomp) {
    std::vector<float> v((size_t) (1 + ncomp) * 16);
    for (float &x : v)
        if (scanf("%f", &x) != 1) x = 0.f;
    HMat4 M;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            M.m[i][j].v = v[i * 4 + j];
            for (int q = 0; q < ncomp; ++q) M.m[i][j].d[q] = v[(size_t) (1 + q) * 16 + i * 4 + j];
        }
    return M;
}
static void print4(const HMat4 &M, int ncomp) {
    for (int q = 0; q <= ncomp; ++q)
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) printf("%.9g\n", q == 0 ? M.m[i][j].v : M.m[i][j].d[q - 1]);
}
static void print3(const HMat3 &M, int ncomp) {
    for (int q = 0; q <= ncomp; ++q)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) printf("%.9g\n", q == 0 ? M.m[i][j].v : M.m[i][j].d[q - 1]);
}

int main() {
    int comps = 1, dirs = 0;
    if (scanf("%d %d", &comps, &dirs) != 2) return 1;
    hj_ctx().comps = comps;
    hj_ctx().dirs = dirs;
    const int ncomp = comps * dirs;
    const HMat4 A = read4(ncomp), B = read4(ncomp);
    HJet angle;
    if (scanf("%f", &angle.v) != 1) return 1;
    for (int q = 0; q < ncomp; ++q)
        if (scanf("%f", &angle.d[q]) != 1) return 1;
    print4(hinverse(A), ncomp);
    print4(hmul(A, B), ncomp);
    print3(hinverse(hrotation(A)), ncomp);
    print3(hmul(hmul(haxis_rotation(angle, 2), haxis_rotation(angle, 1)), haxis_rotation(angle, 0)), ncomp);
    return 0;
}
