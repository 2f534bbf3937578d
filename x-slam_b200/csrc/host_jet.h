// host_jet.h — host-side batched complex-step numbers and the small pose algebra of the orchestrator.
//
// HJet is the host counterpart of xs::Jet: one float real part plus ncomp h-scaled derivative components
// (comps = 1: eps per direction; comps = 3: eps1, eps2, eps1eps2 per direction; comps = 2: Hessian batch - dirs first-order
// components F_i followed by one second-order component S_k per listed parameter pair, xs_common.cuh).  It replaces the
// Eigen::Matrix4cf / Matrix3frm / Vector3cf algebra of the reference orchestrator
// (XKinectFusion/src/KinectFusionReconstruction.cpp:161-332, std::complex<float> scalars) for a whole
// batch of perturbation directions at once.  Real parts are evaluated with plain float operations, which
// is what std::complex<float> arithmetic reduces to when imaginary parts vanish.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

namespace xs {

constexpr int HJ_MAX = 256;  // max derivative components per number

struct HPair {
    int i, j;
};
struct HJetCtx {
    int comps = 1, dirs = 0;
    int npairs = 0;                 // comps == 2 only
    const HPair *pairs = nullptr;   // comps == 2 only: [npairs], component dirs + k is S(pairs[k].i, pairs[k].j)
    int ncomp() const { return comps == 2 ? dirs + npairs : comps * dirs; }
};
inline HJetCtx &hj_ctx() {
    static thread_local HJetCtx c;
    return c;
}

struct HJetNoInit {};
struct HJet {
    float v = 0.f;
    float d[HJ_MAX];
    HJet() { std::memset(d, 0, sizeof(float) * hj_ctx().ncomp()); }
    explicit HJet(HJetNoInit) {}  // every live component is about to be overwritten
    HJet(float x) : v(x) { std::memset(d, 0, sizeof(float) * hj_ctx().ncomp()); }
    // copies touch only the live components
    HJet(const HJet &o) : v(o.v) { std::memcpy(d, o.d, sizeof(float) * hj_ctx().ncomp()); }
    HJet &operator=(const HJet &o) {
        v = o.v;
        std::memcpy(d, o.d, sizeof(float) * hj_ctx().ncomp());
        return *this;
    }
};

inline HJet operator+(const HJet &a, const HJet &b) {
    HJet r{HJetNoInit()};
    r.v = a.v + b.v;
    const int n = hj_ctx().ncomp();
    for (int i = 0; i < n; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
}
inline HJet operator-(const HJet &a, const HJet &b) {
    HJet r{HJetNoInit()};
    r.v = a.v - b.v;
    const int n = hj_ctx().ncomp();
    for (int i = 0; i < n; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
}
inline HJet operator-(const HJet &a) {
    HJet r{HJetNoInit()};
    r.v = -a.v;
    const int n = hj_ctx().ncomp();
    for (int i = 0; i < n; ++i) r.d[i] = -a.d[i];
    return r;
}
inline HJet operator*(const HJet &a, const HJet &b) {
    HJet r{HJetNoInit()};
    r.v = a.v * b.v;
    const HJetCtx &c = hj_ctx();
    const int n = c.ncomp();
    const float av = a.v, bv = b.v;
    // the product rule term of every component as one flat (vectorisable) loop; the eps1eps2 cross terms are added after, in
    // the order of the expression  a.v*b12 + a12*b.v + a1*b2 + a2*b1
    for (int i = 0; i < n; ++i) r.d[i] = av * b.d[i] + a.d[i] * bv;
    if (c.comps == 3) {
        for (int k = 0; k < c.dirs; ++k) {
            float t = r.d[3 * k + 2];
            t = t + a.d[3 * k] * b.d[3 * k + 1];
            t = t + a.d[3 * k + 1] * b.d[3 * k];
            r.d[3 * k + 2] = t;
        }
    } else if (c.comps == 2) {
        for (int k = 0; k < c.npairs; ++k) {
            const int i = c.pairs[k].i, j = c.pairs[k].j;
            float t = r.d[c.dirs + k];
            t = t + a.d[i] * b.d[j];
            t = t + a.d[j] * b.d[i];
            r.d[c.dirs + k] = t;
        }
    }
    return r;
}
inline HJet operator/(const HJet &a, const HJet &b) {
    HJet r{HJetNoInit()};
    r.v = a.v / b.v;
    const float inv = 1.f / b.v;
    const HJetCtx &c = hj_ctx();
    if (c.comps == 1) {
        for (int k = 0; k < c.dirs; ++k) r.d[k] = (a.d[k] - r.v * b.d[k]) * inv;
    } else if (c.comps == 2) {
        for (int i = 0; i < c.dirs; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
        for (int k = 0; k < c.npairs; ++k) {
            const int i = c.pairs[k].i, j = c.pairs[k].j, s = c.dirs + k;
            r.d[s] = (a.d[s] - r.v * b.d[s] - r.d[i] * b.d[j] - r.d[j] * b.d[i]) * inv;
        }
    } else {
        for (int k = 0; k < c.dirs; ++k) {
            const float a1 = a.d[3 * k], a2 = a.d[3 * k + 1], a12 = a.d[3 * k + 2];
            const float b1 = b.d[3 * k], b2 = b.d[3 * k + 1], b12 = b.d[3 * k + 2];
            const float q1 = (a1 - r.v * b1) * inv, q2 = (a2 - r.v * b2) * inv;
            r.d[3 * k] = q1;
            r.d[3 * k + 1] = q2;
            r.d[3 * k + 2] = (a12 - r.v * b12 - q1 * b2 - q2 * b1) * inv;
        }
    }
    return r;
}
// f(a) with first and second derivative f1, f2 at a.v
inline HJet hj_apply(const HJet &a, float f0, float f1, float f2) {
    HJet r;
    r.v = f0;
    const HJetCtx &c = hj_ctx();
    if (c.comps == 1) {
        for (int k = 0; k < c.dirs; ++k) r.d[k] = f1 * a.d[k];
    } else if (c.comps == 2) {
        for (int i = 0; i < c.dirs; ++i) r.d[i] = f1 * a.d[i];
        for (int k = 0; k < c.npairs; ++k) r.d[c.dirs + k] = f1 * a.d[c.dirs + k] + f2 * a.d[c.pairs[k].i] * a.d[c.pairs[k].j];
    } else {
        for (int k = 0; k < c.dirs; ++k) {
            r.d[3 * k] = f1 * a.d[3 * k];
            r.d[3 * k + 1] = f1 * a.d[3 * k + 1];
            r.d[3 * k + 2] = f1 * a.d[3 * k + 2] + f2 * a.d[3 * k] * a.d[3 * k + 1];
        }
    }
    return r;
}
inline HJet hj_sqrt(const HJet &a) {
    const float s = std::sqrt(a.v);
    return hj_apply(a, s, 0.5f / s, -0.25f / (s * a.v));
}
inline HJet hj_sin(const HJet &a) { return hj_apply(a, std::sin(a.v), std::cos(a.v), -std::sin(a.v)); }
inline HJet hj_cos(const HJet &a) { return hj_apply(a, std::cos(a.v), -std::sin(a.v), -std::cos(a.v)); }

struct HMat3 {
    HJet m[3][3];
};
struct HVec3 {
    HJet v[3];
};
struct HMat4 {
    HJet m[4][4];
    static HMat4 identity() {
        HMat4 r;
        for (int i = 0; i < 4; ++i) r.m[i][i] = HJet(1.f);
        return r;
    }
};

inline HMat4 hmul(const HMat4 &a, const HMat4 &b) {
    HMat4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            HJet s = a.m[i][0] * b.m[0][j];
            for (int k = 1; k < 4; ++k) s = s + a.m[i][k] * b.m[k][j];
            r.m[i][j] = s;
        }
    return r;
}
inline HMat3 hmul(const HMat3 &a, const HMat3 &b) {
    HMat3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            HJet s = a.m[i][0] * b.m[0][j];
            for (int k = 1; k < 3; ++k) s = s + a.m[i][k] * b.m[k][j];
            r.m[i][j] = s;
        }
    return r;
}
inline HVec3 hmul(const HMat3 &a, const HVec3 &b) {
    HVec3 r;
    for (int i = 0; i < 3; ++i) {
        HJet s = a.m[i][0] * b.v[0];
        for (int k = 1; k < 3; ++k) s = s + a.m[i][k] * b.v[k];
        r.v[i] = s;
    }
    return r;
}

// fixed-size inverses by cofactors (what Eigen does for 3x3 / 4x4), no conjugation
inline HMat3 hinverse(const HMat3 &A) {
    const HJet(&m)[3][3] = A.m;
    HJet c00 = m[1][1] * m[2][2] - m[1][2] * m[2][1];
    HJet c01 = m[1][2] * m[2][0] - m[1][0] * m[2][2];
    HJet c02 = m[1][0] * m[2][1] - m[1][1] * m[2][0];
    HJet det = m[0][0] * c00 + m[0][1] * c01 + m[0][2] * c02;
    HJet inv = HJet(1.f) / det;
    HMat3 r;
    r.m[0][0] = c00 * inv;
    r.m[1][0] = c01 * inv;
    r.m[2][0] = c02 * inv;
    r.m[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * inv;
    r.m[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * inv;
    r.m[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * inv;
    r.m[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * inv;
    r.m[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * inv;
    r.m[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * inv;
    return r;
}
inline HJet hdet3(const HJet &a, const HJet &b, const HJet &c, const HJet &d, const HJet &e, const HJet &f,
                  const HJet &g, const HJet &h, const HJet &i) {
    return a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
}
inline HMat4 hinverse(const HMat4 &A) {
    HMat4 cof;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            int r[3], c[3];
            for (int k = 0, n = 0; k < 4; ++k)
                if (k != i) r[n++] = k;
            for (int k = 0, n = 0; k < 4; ++k)
                if (k != j) c[n++] = k;
            HJet d = hdet3(A.m[r[0]][c[0]], A.m[r[0]][c[1]], A.m[r[0]][c[2]], A.m[r[1]][c[0]], A.m[r[1]][c[1]],
                           A.m[r[1]][c[2]], A.m[r[2]][c[0]], A.m[r[2]][c[1]], A.m[r[2]][c[2]]);
            cof.m[i][j] = ((i + j) & 1) ? -d : d;
        }
    HJet det = A.m[0][0] * cof.m[0][0] + A.m[0][1] * cof.m[0][1] + A.m[0][2] * cof.m[0][2] + A.m[0][3] * cof.m[0][3];
    HMat4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r.m[i][j] = cof.m[j][i] / det;
    return r;
}
inline HMat3 hrotation(const HMat4 &T) {
    HMat3 R;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R.m[i][j] = T.m[i][j];
    return R;
}
inline HVec3 htranslation(const HMat4 &T) {
    HVec3 t;
    for (int i = 0; i < 3; ++i) t.v[i] = T.m[i][3];
    return t;
}
// rotation about a coordinate axis (0=X,1=Y,2=Z) by a batched angle: Eigen AngleAxis::toRotationMatrix()
// specialised to a unit axis (KinectFusionReconstruction.cpp:215-218)
inline HMat3 haxis_rotation(const HJet &angle, int axis) {
    const HJet s = hj_sin(angle), c = hj_cos(angle);
    HMat3 R;
    for (int i = 0; i < 3; ++i) R.m[i][i] = (i == axis) ? HJet(1.f) - c + c : c;
    const int a = (axis + 1) % 3, b = (axis + 2) % 3;
    R.m[a][b] = -s;
    R.m[b][a] = s;
    return R;
}

}  // namespace xs
