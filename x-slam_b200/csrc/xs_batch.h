// xs_batch.h — the meaning of the derivative components a pipeline carries (host + device, plain C++: included by the
// CUDA sources through xs_common.cuh and by the host orchestrator kinfu.cpp).
#pragma once
#include <vector_types.h>

namespace xs {

// ---- derivative batches ------------------------------------------------------------------
// What the ncomp derivative components carried beside every real value mean ("comps" of the C-ABI):
//   kind 1  CSFD list      n independent first-order directions: component q = eps of direction q
//   kind 3  DCSFD list     n independent bicomplex directions: components 3q, 3q+1, 3q+2 = eps1, eps2, eps1eps2
//   kind 2  Hessian batch  n parameters: components 0..n-1 = first-order F_i (h d/dtheta_i), then one second-order component
//                          S_k = h^2 d2/dtheta_i dtheta_j per listed pair k = (i, j), i <= j, at index n + k.  Truncated product
//                          rule: (ab)_F_i = a F_i(b) + F_i(a) b;  (ab)_S_ij = a S(b) + S(a) b + F_i(a) F_j(b) + F_j(a) F_i(b),
//                          i.e. the bicomplex product of cuda_double_complex.hpp:119-133 with (eps1, eps2, eps1eps2) =
//                          (F_i, F_j, S_ij) - every first-order plane is stored once instead of once per pair it occurs in.
//
// Intrinsic parameters (kind 2 only; BASELINE.json configs[3]: "Hessian w.r.t. pose + intrinsics"): a parameter may also move
// the camera intrinsics, dintr[p] = h d(fx, fy, cx, cy) / d theta_p.  The reference's Intr is plain floats (Internal.h:49-59),
// so this is new behaviour: the current-frame vertex / normal maps then carry derivative components - but only for the
// parameters that move an intrinsic and for the pairs of two such parameters (a map of the current frame does not depend on
// the pose).  cslot[c] is the slot of batch component c in those maps (-1: the component is identically zero), ncurr their number.
struct BatchView {  // passed by value to kernels
    int kind;
    int n;  // directions (kind 1, 3) or parameters (kind 2)
    int m;  // kind 2: listed pairs
    int ncomp;
    const int2 *pairs;   // kind 2: device [m], sorted by i
    const float *dintr;  // kind 2: device [n][4] first-order intrinsic seeds (level 0), or null
    const int *cslot;    // device [ncomp] slot in the current-frame derivative maps, or null
    int ncurr;           // derivative components of the current-frame maps
    float gx0, gy0;      // 1 / fx, 1 / fy at pyramid level 0 (with dintr)
};
struct Batch {
    BatchView v = {1, 0, 0, 0, nullptr, nullptr, nullptr, 0, 0.f, 0.f};
    int2 *h_pairs = nullptr;   // host copy [m]
    float *h_dintr = nullptr;  // host copy [n][4]
    int *h_cslot = nullptr;    // host copy [ncomp]
};
// kind 2: attaches first-order intrinsic seeds dintr[n][4] (null / all zero: none) and builds the slot table
int batch_set_intrinsics(Batch &b, const float *dintr, float fx0, float fy0);
// comps: 1, 3 (dirs = directions) or 2 (dirs = parameters; pairs = null -> all n(n+1)/2 pairs).  Uploads the pair table.
int batch_init(Batch &b, int comps, int dirs, int npairs, const int *pairs);
void batch_free(Batch &b);
inline int batch_ncomp(int comps, int dirs, int npairs) {
    return comps == 2 ? dirs + (npairs >= 0 ? npairs : dirs * (dirs + 1) / 2) : comps * dirs;
}

}  // namespace xs
