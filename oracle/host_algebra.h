// oracle/host_algebra.h — TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Host-side pose algebra of the reference orchestrator, restated without Eigen
// (Eigen is an un-vendored, unpinned dependency of the reference:
// XKinectFusion/CMakeLists.txt:4, "find_package(Eigen3 REQUIRED)"; Ubuntu 22.04 per
// README.md:31 => Eigen 3.4.0).  PARITY UNPINNED at this boundary: the reference ships
// no test for it and Eigen is not available offline; what is restated here is Eigen
// 3.4's published algorithm for each call site:
//
//   Matrix4cf::inverse()        KinectFusionReconstruction.cpp:167,182,231,248,250,305,307
//                               fixed-size 4x4 => cofactors / determinant, no conjugation
//   Matrix3frm::inverse()       KinectFusionReconstruction.cpp:170   (3x3 cofactors)
//   A.real().determinant()      KinectFusionReconstruction.cpp:203   (value only guards a threshold)
//   A.llt().solve(b)            KinectFusionReconstruction.cpp:211   unblocked in-place Cholesky on
//                               the LOWER triangle treating A as Hermitian: x = real(A_kk) - |L_k,:|^2,
//                               A21 -= A20 * conj(A10)^T, A21 /= sqrt(x); then L y = b, L^H x = y.
//   (Matrix3c)AngleAxisc(g,Z) * AngleAxisc(b,Y) * AngleAxisc(a,X)
//                               KinectFusionReconstruction.cpp:215-218   the C cast binds to the first
//                               factor only, so this is three AngleAxis::toRotationMatrix() products.
//
// All arithmetic is std::complex<float> / std::complex<double> exactly as in the reference.
#pragma once
#include <cmath>
#include <complex>

namespace xo {

typedef std::complex<float> cf;
typedef std::complex<double> cd;

struct Mat3c {
    cf m[3][3];
};
struct Vec3c {
    cf v[3];
};
struct Mat4c {
    cf m[4][4];
    static Mat4c identity() {
        Mat4c r;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) r.m[i][j] = cf(i == j ? 1.f : 0.f, 0.f);
        return r;
    }
};

inline Mat4c mul(const Mat4c &a, const Mat4c &b) {
    Mat4c r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            cf s = a.m[i][0] * b.m[0][j];
            for (int k = 1; k < 4; ++k) s += a.m[i][k] * b.m[k][j];
            r.m[i][j] = s;
        }
    return r;
}
inline Mat3c mul(const Mat3c &a, const Mat3c &b) {
    Mat3c r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            cf s = a.m[i][0] * b.m[0][j];
            for (int k = 1; k < 3; ++k) s += a.m[i][k] * b.m[k][j];
            r.m[i][j] = s;
        }
    return r;
}
inline Vec3c mul(const Mat3c &a, const Vec3c &b) {
    Vec3c r;
    for (int i = 0; i < 3; ++i) {
        cf s = a.m[i][0] * b.v[0];
        for (int k = 1; k < 3; ++k) s += a.m[i][k] * b.v[k];
        r.v[i] = s;
    }
    return r;
}

// 3x3 determinant helper on arbitrary entries
template <class S> inline S det3(S a, S b, S c, S d, S e, S f, S g, S h, S i) {
    return a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
}

// Cofactor inverse (Eigen 3.4 fixed-size path).  No conjugation anywhere.
inline Mat3c inverse(const Mat3c &A) {
    const cf(&m)[3][3] = A.m;
    cf c00 = m[1][1] * m[2][2] - m[1][2] * m[2][1];
    cf c01 = m[1][2] * m[2][0] - m[1][0] * m[2][2];
    cf c02 = m[1][0] * m[2][1] - m[1][1] * m[2][0];
    cf det = m[0][0] * c00 + m[0][1] * c01 + m[0][2] * c02;
    cf inv = cf(1.f, 0.f) / det;
    Mat3c r;
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

inline Mat4c inverse(const Mat4c &A) {
    Mat4c cof;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            int r[3], c[3];
            for (int k = 0, n = 0; k < 4; ++k)
                if (k != i) r[n++] = k;
            for (int k = 0, n = 0; k < 4; ++k)
                if (k != j) c[n++] = k;
            cf d = det3<cf>(A.m[r[0]][c[0]], A.m[r[0]][c[1]], A.m[r[0]][c[2]], A.m[r[1]][c[0]], A.m[r[1]][c[1]],
                            A.m[r[1]][c[2]], A.m[r[2]][c[0]], A.m[r[2]][c[1]], A.m[r[2]][c[2]]);
            cof.m[i][j] = ((i + j) & 1) ? -d : d;
        }
    cf det = A.m[0][0] * cof.m[0][0] + A.m[0][1] * cof.m[0][1] + A.m[0][2] * cof.m[0][2] + A.m[0][3] * cof.m[0][3];
    Mat4c r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r.m[i][j] = cof.m[j][i] / det;
    return r;
}

inline Mat3c rotation_of(const Mat4c &T) {
    Mat3c R;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R.m[i][j] = T.m[i][j];
    return R;
}
inline Vec3c translation_of(const Mat4c &T) {
    Vec3c t;
    for (int i = 0; i < 3; ++i) t.v[i] = T.m[i][3];
    return t;
}

// AngleAxis<complex<float>>::toRotationMatrix() for a unit coordinate axis (0=X,1=Y,2=Z).
inline Mat3c angle_axis_matrix(cf angle, int axis) {
    cf ax[3] = {cf(0.f, 0.f), cf(0.f, 0.f), cf(0.f, 0.f)};
    ax[axis] = cf(1.f, 0.f);
    cf s = std::sin(angle), c = std::cos(angle);
    cf sin_axis[3], cos1_axis[3];
    for (int i = 0; i < 3; ++i) {
        sin_axis[i] = s * ax[i];
        cos1_axis[i] = (cf(1.f, 0.f) - c) * ax[i];
    }
    Mat3c R;
    cf tmp;
    tmp = cos1_axis[0] * ax[1];
    R.m[0][1] = tmp - sin_axis[2];
    R.m[1][0] = tmp + sin_axis[2];
    tmp = cos1_axis[0] * ax[2];
    R.m[0][2] = tmp + sin_axis[1];
    R.m[2][0] = tmp - sin_axis[1];
    tmp = cos1_axis[1] * ax[2];
    R.m[1][2] = tmp - sin_axis[0];
    R.m[2][1] = tmp + sin_axis[0];
    for (int i = 0; i < 3; ++i) R.m[i][i] = cos1_axis[i] * ax[i] + c;
    return R;
}

// |det(Re A)| guard of KinectFusionReconstruction.cpp:203-210 (6x6, Gaussian elimination, partial pivoting).
inline double det6_real(const cd *A /* column-major 6x6 */) {
    double M[6][6];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) M[i][j] = A[j * 6 + i].real();
    double det = 1.0;
    for (int k = 0; k < 6; ++k) {
        int p = k;
        for (int i = k + 1; i < 6; ++i)
            if (std::fabs(M[i][k]) > std::fabs(M[p][k])) p = i;
        if (M[p][k] == 0.0) return 0.0;
        if (p != k) {
            for (int j = 0; j < 6; ++j) std::swap(M[p][j], M[k][j]);
            det = -det;
        }
        det *= M[k][k];
        for (int i = k + 1; i < 6; ++i) {
            double f = M[i][k] / M[k][k];
            for (int j = k; j < 6; ++j) M[i][j] -= f * M[k][j];
        }
    }
    return det;
}

// Eigen 3.4 LLT<Matrix<complex<double>,6,6>, Lower>::compute + solve (unblocked path, n < 32).
// Returns false when the factorisation meets a non-positive pivot (Eigen reports NumericalIssue
// but solve() still runs on the partial factor; the reference never checks info()).
inline bool llt_solve6(const cd *A_colmajor, const cd *b, cd *x) {
    cd L[6][6];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) L[i][j] = A_colmajor[j * 6 + i];
    bool ok = true;
    for (int k = 0; k < 6; ++k) {
        double xk = L[k][k].real();
        for (int j = 0; j < k; ++j) xk -= std::norm(L[k][j]);
        if (xk <= 0.0) {
            ok = false;
            break;
        }
        xk = std::sqrt(xk);
        L[k][k] = cd(xk, 0.0);
        for (int i = k + 1; i < 6; ++i) {
            cd s = L[i][k];
            for (int j = 0; j < k; ++j) s -= L[i][j] * std::conj(L[k][j]);
            L[i][k] = s / xk;
        }
    }
    cd y[6];
    for (int i = 0; i < 6; ++i) {  // L y = b
        cd s = b[i];
        for (int j = 0; j < i; ++j) s -= L[i][j] * y[j];
        y[i] = s / L[i][i];
    }
    for (int i = 5; i >= 0; --i) {  // L^H x = y
        cd s = y[i];
        for (int j = i + 1; j < 6; ++j) s -= std::conj(L[j][i]) * x[j];
        x[i] = s / std::conj(L[i][i]);
    }
    return ok;
}

// One Gauss-Newton pose update, KinectFusionReconstruction.cpp:211-224.
inline void pose_update(const cd *x, Mat3c &Rcurr, Vec3c &tcurr) {
    cf r[6];
    for (int i = 0; i < 6; ++i) r[i] = cf((float) x[i].real(), (float) x[i].imag());
    Mat3c Rinc = mul(mul(angle_axis_matrix(r[2], 2), angle_axis_matrix(r[1], 1)), angle_axis_matrix(r[0], 0));
    Vec3c t = mul(Rinc, tcurr);
    for (int i = 0; i < 3; ++i) tcurr.v[i] = t.v[i] + r[3 + i];
    Rcurr = mul(Rinc, Rcurr);
}

// se3Exp, KinectFusionReconstruction.h:176-219 (xi = [v; omega]).  Used to seed pose perturbations.
inline Mat4c se3_exp(const cf xi[6]) {
    cf v[3] = {xi[0], xi[1], xi[2]}, w[3] = {xi[3], xi[4], xi[5]};
    Mat3c W;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) W.m[i][j] = cf(0.f, 0.f);
    W.m[0][1] = -w[2];
    W.m[0][2] = w[1];
    W.m[1][2] = -w[0];
    W.m[1][0] = w[2];
    W.m[2][0] = -w[1];
    W.m[2][1] = w[0];
    Mat3c R, V;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R.m[i][j] = V.m[i][j] = cf(i == j ? 1.f : 0.f, 0.f);
    // Eigen's norm() of a complex vector is sqrt(sum |w_i|^2)
    float nrm = std::sqrt(std::norm(w[0]) + std::norm(w[1]) + std::norm(w[2]));
    if (nrm < 1e-6) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                R.m[i][j] += W.m[i][j];
                V.m[i][j] += W.m[i][j];
            }
    } else {
        cf sum = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
        cf theta = std::sqrt(sum);
        cf s = std::sin(theta), c = std::cos(theta);
        Mat3c W2 = mul(W, W);
        cf A = s / theta;
        cf B = (1.0f - c) / std::pow(theta, 2.0f);
        cf C = (theta - s) / std::pow(theta, 3.0f);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                R.m[i][j] = R.m[i][j] + A * W.m[i][j] + B * W2.m[i][j];
                V.m[i][j] = V.m[i][j] + B * W.m[i][j] + C * W2.m[i][j];
            }
    }
    Mat4c T;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) T.m[i][j] = cf(0.f, 0.f);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) T.m[i][j] = R.m[i][j];
        T.m[i][3] = V.m[i][0] * v[0] + V.m[i][1] * v[1] + V.m[i][2] * v[2];
    }
    T.m[3][3] = cf(1.f, 0.f);
    return T;
}

}  // namespace xo
