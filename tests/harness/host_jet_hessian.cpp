// Test harness (CPU only): the host jet algebra of x-slam_b200/csrc/host_jet.h evaluated on the same seeded 4x4 pose chain
// (inverse, product, inverse of the rotation block, Rz*Ry*Rx of batched angles, a quotient) twice: as a Hessian batch
// (comps = 2: n first-order + pair second-order components) and as the DCSFD list of the same pairs (comps = 3).
// Prints per pair the largest |F_i - eps1|, |F_j - eps2|, |S_ij - eps1eps2| over all outputs, relative to the component scale.
// Driven by tests/test_host_jet.py.
#include "../../x-slam_b200/csrc/host_jet.h"

#include <cstdio>
#include <random>
#include <vector>

using namespace xs;

static std::vector<HJet> chain(const std::vector<std::vector<float>> &seed_comps /* [16 + 3][ncomp] */) {
    HMat4 W = HMat4::identity();
    const float base[16] = {0.98f, -0.05f, 0.17f, 0.3f, 0.06f, 0.99f, -0.04f, -0.2f, -0.16f, 0.05f, 0.98f, 0.5f, 0, 0, 0, 1};
    const int nc = hj_ctx().ncomp();
    for (int e = 0; e < 16; ++e) {
        W.m[e / 4][e % 4] = HJet(base[e]);
        for (int q = 0; q < nc; ++q) W.m[e / 4][e % 4].d[q] = seed_comps[e][q];
    }
    HJet ang[3] = {HJet(0.02f), HJet(-0.01f), HJet(0.03f)};
    for (int a = 0; a < 3; ++a)
        for (int q = 0; q < nc; ++q) ang[a].d[q] = seed_comps[16 + a][q];
    HMat4 Wi = hinverse(W);
    HMat4 P = hmul(Wi, hmul(W, Wi));
    HMat3 R = hmul(hmul(haxis_rotation(ang[2], 2), haxis_rotation(ang[1], 1)), haxis_rotation(ang[0], 0));
    HMat3 Q = hmul(hinverse(hrotation(P)), R);
    std::vector<HJet> out;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) out.push_back(P.m[i][j]);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) out.push_back(Q.m[i][j]);
    out.push_back(Q.m[0][0] / (P.m[1][1] + ang[0]));
    return out;
}

int main() {
    const int n = 4;
    std::vector<HPair> pairs;
    for (int i = 0; i < n; ++i)
        for (int j = i; j < n; ++j) pairs.push_back(HPair{i, j});
    pairs.erase(pairs.begin() + 3);  // a subset, as a rank of a multi-GPU run holds it
    const int m = (int) pairs.size();
    std::mt19937 rng(3);
    std::normal_distribution<float> nd(0.f, 1.f);
    const float h = 1e-7f;
    std::vector<std::vector<float>> first(19, std::vector<float>(n)), second(19, std::vector<float>(m));
    for (auto &v : first)
        for (auto &x : v) x = h * nd(rng);
    for (auto &v : second)
        for (auto &x : v) x = h * h * nd(rng);
    // Hessian batch
    hj_ctx().comps = 2, hj_ctx().dirs = n, hj_ctx().npairs = m, hj_ctx().pairs = pairs.data();
    std::vector<std::vector<float>> sh(19, std::vector<float>(n + m));
    for (int e = 0; e < 19; ++e) {
        for (int i = 0; i < n; ++i) sh[e][i] = first[e][i];
        for (int k = 0; k < m; ++k) sh[e][n + k] = second[e][k];
    }
    const std::vector<HJet> H = chain(sh);
    std::vector<std::vector<float>> hv;
    for (const HJet &j : H) hv.push_back(std::vector<float>(j.d, j.d + n + m));
    // DCSFD list of the same pairs
    hj_ctx().comps = 3, hj_ctx().dirs = m, hj_ctx().npairs = 0, hj_ctx().pairs = nullptr;
    std::vector<std::vector<float>> sl(19, std::vector<float>(3 * m));
    for (int e = 0; e < 19; ++e)
        for (int k = 0; k < m; ++k) sl[e][3 * k] = first[e][pairs[k].i], sl[e][3 * k + 1] = first[e][pairs[k].j], sl[e][3 * k + 2] = second[e][k];
    const std::vector<HJet> L = chain(sl);
    for (int k = 0; k < m; ++k) {
        double e1 = 0, e2 = 0, e12 = 0, s1 = 0, s12 = 0;
        for (size_t o = 0; o < H.size(); ++o) {
            e1 = std::max(e1, (double) std::fabs(hv[o][pairs[k].i] - L[o].d[3 * k]));
            e2 = std::max(e2, (double) std::fabs(hv[o][pairs[k].j] - L[o].d[3 * k + 1]));
            e12 = std::max(e12, (double) std::fabs(hv[o][n + k] - L[o].d[3 * k + 2]));
            s1 = std::max(s1, (double) std::fabs(L[o].d[3 * k]));
            s12 = std::max(s12, (double) std::fabs(L[o].d[3 * k + 2]));
        }
        std::printf("%d %d %.3e %.3e %.3e\n", pairs[k].i, pairs[k].j, e1 / s1, e2 / s1, e12 / s12);
    }
    return 0;
}
