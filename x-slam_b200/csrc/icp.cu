// icp.cu — direction-batched projective point-to-plane ICP normal equations.
//
// Replaces Combined::{search_newton, operator()} / combinedKernel (XKinectFusion/src/ICP.cu:166-281,357),
// TranformReduction / TransformEstimatorKernel (ICP.cu:120-164) and estimateCombined (ICP.cu:365-429).
//
// Two launches per Gauss-Newton iteration and NO host round trip (the reference: 2 launches per direction and
// iteration, then sync + download + host Eigen solve, ICP.cu:414-417 / KinectFusionReconstruction.cpp:196-224):
//   1. icp_assoc_kernel   data association (projection, bounds, NaN, distance and angle gates) ONCE per pixel on real
//                         parts, the real 7-vector row [cross(s,n), n, n.(d-s)], its 27 upper-triangular products
//                         widened to double exactly as the reference does (product in float, sum in double,
//                         ICP.cu:273-274), and a per-pixel association record (matched index + the 16 real numbers
//                         every direction needs) that stays L2-resident (68 B/pixel).
//   2. icp_deriv_kernel   persistent: the (direction group, pixel unit) items are cut into equal contiguous ranges for
//                         296 CTAs.  A thread reads the record, gathers its direction's derivative planes of the previous
//                         maps at the matched pixel (cp.async, one pixel ahead) and accumulates the derivative components
//                         of the 27 products (linearised row algebra, no recomputation of the real path) for at most 32
//                         pixels in FP32; then a fixed FP32 butterfly over the lanes, double totals per thread, the 8 warps
//                         and the CTAs of a group in fixed order.  The LAST CTA of a direction group to arrive adds the
//                         group's partials and runs the host Gauss-Newton step of
//                         KinectFusionReconstruction.cpp:203-224 for its direction on the device (icp_solve_direction:
//                         det guard, 6x6 LLT solve in double, Rinc = Rz*Ry*Rx, pose update - one thread per direction,
//                         every thread redoing the tiny real part).  The current pose therefore lives in device memory
//                         and the 12 iterations of a frame are queued back to back.
// Split chains (frame loop with derivative components): the real part of an iteration - association, real sums, real step
// - depends on no derivative component, so the real chain of a frame (icp_assoc_kernel + a one-thread icp_solve_kernel per
// iteration) runs ahead on a second stream, leaving record, sums and real pose of every iteration in a slot of its own; the
// derivative kernels follow back to back on the pipeline's stream and their tail updates derivative components only
// (SolveParams::deriv_only).  The derivative kernel fills every SM, so the real chain advances in the tails of the derivative
// launches (where the SMs would idle behind the one-thread solves): about a third of its latency is hidden.
// Instead of the reference's 27 sequential 256-thread shared-memory tree reductions per direction, the summation
// order is fixed by construction, so results are deterministic run to run.
#include "xs_common.cuh"

#include <algorithm>
#include <utility>
#include <vector>

namespace xs {

// upper-triangular product order of ICP.cu:267-279: e -> (i, j), i = 0..5, j = i..6 (j == 6 is b).
// constexpr so that the fully unrolled product loops index the row registers statically.
__host__ __device__ constexpr int tri_i(int e) {
    int i = 0, n = 7;
    while (e >= n) {
        e -= n;
        --n;
        ++i;
    }
    return i;
}
__host__ __device__ constexpr int tri_j(int e) {
    int i = 0, n = 7;
    while (e >= n) {
        e -= n;
        --n;
        ++i;
    }
    return i + e;
}

// Sums v[e] over the 32 lanes of a warp for e = 0..31; lane L returns the total of element L.
XS_DEV double warp_transpose_reduce(double (&v)[32]) {
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const double keep = upper ? v[i + half] : v[i];
            const double send = upper ? v[i] : v[i + half];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}

// real products row_i * row_j for e = 0..26 with compile-time row indices (fold over E)
template <int... E> XS_DEV void add_real(double (&v)[32], const float (&r)[7], std::integer_sequence<int, E...>) {
    ((v[E] += (double) __fmul_rn(r[tri_i(E)], r[tri_j(E)])), ...);
}

// per-pixel association record, SoA planes of npix elements each
constexpr int REC_F = 16;  // vc(3) s(3) n(3) e=d-s(3) cr=cross(s,n)(3) r6
struct IcpParams {
    DevPose prev;                         // prev.R = Rprev_inv, prev.t = tprev
    const float *pose_curr;               // [(1+ncomp)][12] device: Rcurr row-major + tcurr; component 0 = real part
    const float *vmap_curr, *nmap_curr;   // [3][rows][cols]
    const float *vmap_prev, *nmap_prev;   // [(1+ncomp)][3][rows][cols]
    xs_intr intr;
    int rows, cols, dirs, ncomp;
    float dist_thres, angle_thres;
    int *rec_idx;      // [npix] matched linear index in the previous maps, -1 = no correspondence
    float4 *rec_f;     // [npix][4]: (vc.xyz, s.x) (s.yz, n.xy) (n.z, e.xyz) (cr.xyz, r6) - 64 contiguous bytes per pixel
    double *partials;  // real: [gridDim.x][27]
    double *sums;      // [27*(1+ncomp)]
    unsigned int *ticket;
    int tiles_x, tiles_y;
    // derivative pass
    double *dpartials;  // [groups][max_writers][81]
    int chunks, groups, ppt;  // chunks = pixel units of 256 * ppt pixels
    unsigned int *group_ticket;  // [groups] arrival counters of the derivative pass (self-resetting)
    int max_writers;             // upper bound of the CTAs that write a partial for one group
    BatchView batch;             // Hessian batch (kind 2): a group is one task = parameter i (component i) or pair k (component n + k)
    unsigned int *done_ticket;   // kind 2: tasks whose sums are complete (self-resetting); the CTA that completes the last one solves
    const struct HTask *htasks;  // kind 2: [groups] task table (device)
    const int *tile_wp;          // kind 2, tile form: [warps][pairs per warp] pair index of every warp slot (-1: empty)
};

// the Gauss-Newton step that closes an iteration (icp_solve_direction below)
struct SolveParams {
    const double *sums;    // [27*(1+ncomp)]
    const float *pose_in;  // [(1+ncomp)][12] current pose
    float *pose_out;       // [(1+ncomp)][12] updated pose (a different buffer: blocks do not synchronise); null = no solve
    int *status;           // [2]: 0 = ok; 1 = |det(Re A)| < 1e-15; 2 = NaN det.  [0] sticky (later iterations are skipped)
    double *log;           // optional [27*(1+ncomp)] copy of the sums of this iteration
    int dirs, ncomp, solve_mode;
    int deriv_only;        // the real part of the step is owned by another launch (split chains): only the derivative
                           // components of pose_out are written, the status flags are read but not set
    const int *status_in;  // Hessian tails: the two status flags staged in shared memory (null: read P.status)
    int staged;            // Hessian tails: sums / real_cache / pose_in point to shared-memory copies (plain loads)
    double *real_cache;    // split chains, [REAL_CACHE]: Cholesky factor (36), reciprocal pivots (6) and real solution (6) of
                           // this iteration, stored by the real step and loaded by the derivative tails (which then skip the
                           // determinant guard - the real step has already set the status - the factorisation and the real solve)
};
constexpr int REAL_CACHE = 48;
template <int C> __device__ __noinline__ void icp_solve_direction(const SolveParams &P, int q, const double *real_sums,
                                                                  const double *comp_sums);
template <int NT> __device__ void icp_hessian_tail(const IcpParams &P, const SolveParams &S, unsigned char *s_raw, bool reduced, int tid);

// search_newton (ICP.cu:196-244) on real parts with the reference's rounding sequence: projection of the current
// vertex into the previous frame, bounds / NaN / distance / angle gates.  Outputs vcurr, vcurr_g and the matched pixel.
XS_DEV bool search_newton_real(const IcpParams &P, const float *s_curr, int x, int y, size_t plane, float &vcx, float &vcy,
                               float &vcz, float &gx, float &gy, float &gz, int &ux, int &uy, float (&nprev)[3], float (&vprev)[3]) {
    const size_t pix = (size_t) y * P.cols + x;
    // the six current-frame values in one round trip (the reference tests the NaN first; the loads are valid either way)
    const float ncx = P.nmap_curr[pix], ncy = P.nmap_curr[pix + plane], ncz = P.nmap_curr[pix + 2 * plane];
    vcx = P.vmap_curr[pix];
    vcy = P.vmap_curr[pix + plane];
    vcz = P.vmap_curr[pix + 2 * plane];
    if (isnan(ncx)) return false;
    const float *R = s_curr, *t = s_curr + 9, *Q = P.prev.R, *tp = P.prev.t;
    // vcurr_g = Rcurr * vcurr + tcurr
    gx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[0], vcx), __fmul_rn(R[1], vcy)), __fmul_rn(R[2], vcz)), t[0]);
    gy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[3], vcx), __fmul_rn(R[4], vcy)), __fmul_rn(R[5], vcz)), t[1]);
    gz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[6], vcx), __fmul_rn(R[7], vcy)), __fmul_rn(R[8], vcz)), t[2]);
    // vcurr_cp = Rprev_inv * (vcurr_g - tprev)
    const float ex = __fsub_rn(gx, tp[0]), ey = __fsub_rn(gy, tp[1]), ez = __fsub_rn(gz, tp[2]);
    const float px = __fadd_rn(__fadd_rn(__fmul_rn(Q[0], ex), __fmul_rn(Q[1], ey)), __fmul_rn(Q[2], ez));
    const float py = __fadd_rn(__fadd_rn(__fmul_rn(Q[3], ex), __fmul_rn(Q[4], ey)), __fmul_rn(Q[5], ez));
    const float pz = __fadd_rn(__fadd_rn(__fmul_rn(Q[6], ex), __fmul_rn(Q[7], ey)), __fmul_rn(Q[8], ez));
    ux = __float2int_rn(__fadd_rn(__fdiv_rn(__fmul_rn(px, P.intr.fx), pz), P.intr.cx));
    uy = __float2int_rn(__fadd_rn(__fdiv_rn(__fmul_rn(py, P.intr.fy), pz), P.intr.cy));
    if (ux < 0 || uy < 0 || ux >= P.cols || uy >= P.rows || pz < 0) return false;
    const size_t q = (size_t) uy * P.cols + ux;
    // the six matched values in one round trip
    const float npx = P.nmap_prev[q], npy = P.nmap_prev[q + plane], npz = P.nmap_prev[q + 2 * plane];
    const float vpx = P.vmap_prev[q], vpy = P.vmap_prev[q + plane], vpz = P.vmap_prev[q + 2 * plane];
    if (isnan(npx)) return false;
    nprev[0] = npx, nprev[1] = npy, nprev[2] = npz;
    vprev[0] = vpx, vprev[1] = vpy, vprev[2] = vpz;
    // dist = norm(vprev_g - vcurr_g)
    const float ddx = __fsub_rn(vpx, gx), ddy = __fsub_rn(vpy, gy), ddz = __fsub_rn(vpz, gz);
    const float dist = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)), __fmul_rn(ddz, ddz)));
    if (dist > P.dist_thres) return false;
    // sine = norm(cross(Rcurr * ncurr, nprev_g))
    const float mx = __fadd_rn(__fadd_rn(__fmul_rn(R[0], ncx), __fmul_rn(R[1], ncy)), __fmul_rn(R[2], ncz));
    const float my = __fadd_rn(__fadd_rn(__fmul_rn(R[3], ncx), __fmul_rn(R[4], ncy)), __fmul_rn(R[5], ncz));
    const float mz = __fadd_rn(__fadd_rn(__fmul_rn(R[6], ncx), __fmul_rn(R[7], ncy)), __fmul_rn(R[8], ncz));
    const float kx = __fsub_rn(__fmul_rn(my, npz), __fmul_rn(mz, npy));
    const float ky = __fsub_rn(__fmul_rn(mz, npx), __fmul_rn(mx, npz));
    const float kz = __fsub_rn(__fmul_rn(mx, npy), __fmul_rn(my, npx));
    const float sine = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(kx, kx), __fmul_rn(ky, ky)), __fmul_rn(kz, kz)));
    return !(sine >= P.angle_thres);
}

// ---------------------------------------------------------------------------------------------------------------
// pass 1: association + real normal equations
// SR.pose_out != null: the last block also runs the real Gauss-Newton step of the iteration (one thread, from shared memory)
// instead of a separate one-thread launch - one launch and one kernel-to-kernel dependency less per iteration.
__global__ void __launch_bounds__(256) icp_assoc_kernel(const IcpParams P, const SolveParams SR) {
    __shared__ double s_acc[27];
    __shared__ double s_stage[8][32];
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    __shared__ float s_curr[12];
    if (tid < 27) s_acc[tid] = 0.0;
    if (tid < 12) s_curr[tid] = P.pose_curr[tid];
    __syncthreads();
    const size_t plane = (size_t) P.rows * P.cols;
    const int ntiles = P.tiles_x * P.tiles_y;
    // per-thread double sums over the CTA's tiles (a thread sees at most ceil(ntiles / gridDim.x) pixels), reduced once
    double v[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = 0.0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int x = (tile % P.tiles_x) * 32 + threadIdx.x;
        const int y = (tile / P.tiles_x) * 8 + threadIdx.y;
        // ---------------- search_newton, ICP.cu:196-244 (real parts)
        float vcx = 0, vcy = 0, vcz = 0;  // vcurr (camera frame)
        float gx = 0, gy = 0, gz = 0;     // vcurr_g
        int ux = 0, uy = 0;
        const bool inside = x < P.cols && y < P.rows;
        float np_[3], vp_[3];
        const bool found = inside && search_newton_real(P, s_curr, x, y, plane, vcx, vcy, vcz, gx, gy, gz, ux, uy, np_, vp_);
        // ---------------- the real row, ICP.cu:254-260: s = vcurr_g, n = nprev_g, d = vprev_g
        float row[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (found) {
            const float nx = np_[0], ny = np_[1], nz = np_[2];
            const float ex = __fsub_rn(vp_[0], gx), ey = __fsub_rn(vp_[1], gy), ez = __fsub_rn(vp_[2], gz);
            row[0] = __fsub_rn(__fmul_rn(gy, nz), __fmul_rn(gz, ny));
            row[1] = __fsub_rn(__fmul_rn(gz, nx), __fmul_rn(gx, nz));
            row[2] = __fsub_rn(__fmul_rn(gx, ny), __fmul_rn(gy, nx));
            row[3] = nx;
            row[4] = ny;
            row[5] = nz;
            row[6] = __fadd_rn(__fadd_rn(__fmul_rn(nx, ex), __fmul_rn(ny, ey)), __fmul_rn(nz, ez));
            if (P.ncomp > 0) {
                float4 *f = P.rec_f + ((size_t) y * P.cols + x) * 4;
                f[0] = make_float4(vcx, vcy, vcz, gx);
                f[1] = make_float4(gy, gz, nx, ny);
                f[2] = make_float4(nz, ex, ey, ez);
                f[3] = make_float4(row[0], row[1], row[2], row[6]);
            }
        }
        if (inside && P.ncomp > 0) P.rec_idx[(size_t) y * P.cols + x] = found ? (int) ((size_t) uy * P.cols + ux) : -1;
        // 27 products, widened to double (ICP.cu:273-274)
        if (found) add_real(v, row, std::make_integer_sequence<int, 27>());
    }
    // warp transpose-reduce, then the 8 warps in fixed order
    s_stage[warp][lane] = warp_transpose_reduce(v);
    __syncthreads();
    if (tid < 27) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += s_stage[w][tid];
        s_acc[tid] = sum;
    }
    __syncthreads();
    // ---------------- block partials, then the last block reduces over blocks in block order
    if (tid < 27) P.partials[(size_t) blockIdx.x * 27 + tid] = s_acc[tid];
    __threadfence();
    __shared__ bool s_last;
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(P.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // fixed-order sum over the blocks: warp w adds its contiguous eighth of the per-block partials in order (lane = product,
    // independent L2 loads in flight), then the eight segment sums are added in warp order
    {
        const int nb = (int) gridDim.x, seg = (nb + 7) / 8, b0 = warp * seg, b1 = min(nb, b0 + seg);
        double sum = 0.0;
        if (lane < 27) {
#pragma unroll 8
            for (int b = b0; b < b1; ++b) sum += __ldcg(P.partials + (size_t) b * 27 + lane);
        }
        s_stage[warp][lane] = sum;
    }
    __syncthreads();
    if (tid < 27) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += s_stage[w][tid];
        P.sums[tid] = sum;
        s_acc[tid] = sum;
    }
    if (tid == 0) *P.ticket = 0u;
    if (SR.pose_out == nullptr) return;
    __syncthreads();
    if (tid == 0) icp_solve_direction<1>(SR, 0, s_acc, s_acc);
}

// ---------------------------------------------------------------------------------------------------------------
// pass 2: derivative components of the 27 products for one direction group.
//   C == 3 : the group is one bicomplex direction, slots (0,1,2) = (eps1, eps2, eps1eps2)
//   C == 1 : the group is three independent first-order components
// acc[a][e] accumulates  d_a (row_i * row_j)  for e <-> (i, j).
template <int C, int... E>
XS_DEV void accumulate_products(float (&acc)[3][27], const float (&r)[7], const float (&d0)[7], const float (&d1)[7],
                                const float (&d2)[7], std::integer_sequence<int, E...>) {
    ((acc[0][E] = fmaf(r[tri_i(E)], d0[tri_j(E)], fmaf(d0[tri_i(E)], r[tri_j(E)], acc[0][E]))), ...);
    ((acc[1][E] = fmaf(r[tri_i(E)], d1[tri_j(E)], fmaf(d1[tri_i(E)], r[tri_j(E)], acc[1][E]))), ...);
    if (C == 3) {
        ((acc[2][E] = fmaf(r[tri_i(E)], d2[tri_j(E)],
                           fmaf(d2[tri_i(E)], r[tri_j(E)],
                                fmaf(d0[tri_i(E)], d1[tri_j(E)], fmaf(d1[tri_i(E)], d0[tri_j(E)], acc[2][E]))))),
         ...);
    } else {
        ((acc[2][E] = fmaf(r[tri_i(E)], d2[tri_j(E)], fmaf(d2[tri_i(E)], r[tri_j(E)], acc[2][E]))), ...);
    }
}

XS_DEV void cross3(const float *a, const float *b, float *o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
XS_DEV void cross3_add(const float *a, const float *b, float *o) {
    o[0] += a[1] * b[2] - a[2] * b[1];
    o[1] += a[2] * b[0] - a[0] * b[2];
    o[2] += a[0] * b[1] - a[1] * b[0];
}
XS_DEV float dot3(const float *a, const float *b) { return fmaf(a[0], b[0], fmaf(a[1], b[1], a[2] * b[2])); }

// asynchronous 4-byte global -> shared copies (LDGSTS): the prefetch of the next pixel costs no registers
XS_DEV void cp_async4(float *smem_dst, const float *gsrc) {
    const unsigned sdst = (unsigned) __cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sdst), "l"(gsrc) : "memory");
}
// streamed-once variant: the derivative planes are read once per launch, so their lines are marked evict-first in L2 and
// leave the (re-read) association record resident
XS_DEV void cp_async4_stream(float *smem_dst, const float *gsrc, unsigned long long policy) {
    const unsigned sdst = (unsigned) __cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;" ::"r"(sdst), "l"(gsrc), "l"(policy) : "memory");
}
XS_DEV unsigned long long l2_policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
XS_DEV void cp_async4_keep(float *smem_dst, const float *gsrc, unsigned long long policy) {
    const unsigned sdst = (unsigned) __cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;" ::"r"(sdst), "l"(gsrc), "l"(policy) : "memory");
}
XS_DEV unsigned long long l2_policy_evict_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// pulls the line of a streamed plane into L2 several pixels ahead of its cp.async (no register, no shared memory)
XS_DEV void prefetch_l2(const float *g) { asm volatile("prefetch.global.L2 [%0];" ::"l"(g)); }
XS_DEV void cp_async16(float4 *smem_dst, const float4 *gsrc) {
    const unsigned sdst = (unsigned) __cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst), "l"(gsrc) : "memory");
}
XS_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> XS_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// staged per-pixel inputs: 16 record floats + 3 components x (dn[3], dd[3]) gathered at the matched pixel
constexpr int DERIV_IN = REC_F + 18;
constexpr int DERIV_MAX_STAGES = 3;
// staging [stage][DERIV_IN][256] floats + per-thread double totals [3][256]
constexpr size_t deriv_smem(int stages) { return (size_t) stages * DERIV_IN * 256 * sizeof(float) + 3 * 256 * sizeof(double); }

// Sums v[e] over the 32 lanes of a warp for e = 0..31 in FP32 (fixed butterfly order); lane L returns the total of element L.
XS_DEV float warp_transpose_reduce_f(float (&v)[32]) {
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float keep = upper ? v[i + half] : v[i];
            const float send = upper ? v[i] : v[i + half];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}

// Work decomposition of the derivative pass.  The (direction group, pixel unit) pairs - a unit is 256 * upt consecutive
// pixels - are linearised group-major into U = groups * units items and cut into gridDim.x equal contiguous ranges, one per
// persistent CTA (grid = min(296, U): every CTA slot of the 148 SMs gets the same amount of work whatever the number of
// directions, e.g. the 7 directions a rank of an 8-GPU run holds).  A CTA therefore walks at most a few group segments;
// per segment it writes one 81-value partial and takes a ticket of that group.
//   owner(u) = the CTA whose range [b U / B, (b + 1) U / B) holds item u
XS_DEV int deriv_owner(long long u, long long U, long long B) { return (int) (((u + 1) * B - 1) / U); }

// ST = depth of the cp.async pipeline: the inputs of pixel j + ST - 1 are in flight while pixel j is processed.
template <int C, int ST> __global__ void __launch_bounds__(256, 2) icp_deriv_kernel(const IcpParams P, const SolveParams S) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ float s_pose[3][12];
    __shared__ bool s_last;
    float *s_in = reinterpret_cast<float *>(s_raw);  // [stage][DERIV_IN][256]
    double *s_tot = reinterpret_cast<double *>(s_raw + (size_t) ST * DERIV_IN * 256 * sizeof(float));  // [3][256]
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const size_t plane = (size_t) P.rows * P.cols;
    const int npix = P.rows * P.cols;
    const unsigned long long stream_policy = l2_policy_evict_first();
    const long long U = (long long) P.groups * P.chunks, B = gridDim.x;
    const long long u_begin = blockIdx.x * U / B, u_end = (blockIdx.x + 1) * U / B;
    for (long long u = u_begin; u < u_end;) {
        const int group = (int) (u / P.chunks), c_begin = (int) (u % P.chunks);
        const int c_end = (int) min((long long) P.chunks, c_begin + (u_end - u));
        u += c_end - c_begin;
        const int comp0 = group * 3;
        __syncthreads();  // s_pose / s_tot of the previous segment are no longer read
        if (tid < 36) {
            const int a = tid / 12, e = tid % 12;
            s_pose[a][e] = (comp0 + a < P.ncomp) ? P.pose_curr[(size_t) (1 + comp0 + a) * 12 + e] : 0.f;
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) s_tot[a * 256 + tid] = 0.0;
        __syncthreads();
        float acc[3][27];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int e = 0; e < 27; ++e) acc[a][e] = 0.f;
        // A thread sums at most 32 pixels in FP32; then the 32 lanes are combined by a fixed FP32 butterfly and lane e adds
        // the warp total of product e to its double total in shared memory (no barrier, the pipeline keeps running).
        auto flush = [&]() {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                float v[32];
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] = e < 27 ? acc[a][e] : 0.f;
                s_tot[a * 256 + tid] += (double) warp_transpose_reduce_f(v);
#pragma unroll
                for (int e = 0; e < 27; ++e) acc[a][e] = 0.f;
            }
        };
        const int base = c_begin * 256 * P.ppt;
        const int nitems = (c_end - c_begin) * P.ppt;
        // Software pipeline without registers: the inputs of the pixels ahead (record + the gathers that depend on the
        // matched index) are copied global -> shared asynchronously while pixel j is processed; matched indices run ST
        // pixels ahead in registers.  Every thread reads back only what it copied itself, so no barrier is needed.
        auto issue = [&](int stage, int p, int q) {
            if (q >= 0) {
                float *dst = s_in + (size_t) stage * DERIV_IN * 256 + tid;
                float4 *dst4 = reinterpret_cast<float4 *>(s_in + (size_t) stage * DERIV_IN * 256) + tid;  // [4][256] float4
                const float4 *f = P.rec_f + (size_t) p * 4;
#pragma unroll
                for (int i = 0; i < 4; ++i) cp_async16(dst4 + i * 256, f + i);
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const int comp = min(comp0 + a, P.ncomp - 1);  // out-of-range slots re-read a valid plane; their sums are dropped
                    // 32-bit element offsets (a map set is < 2^32 floats: checked by the host wrapper): one IMAD.WIDE per copy
                    const unsigned uplane = (unsigned) plane;
                    const unsigned o = (unsigned) q + (unsigned) (1 + comp) * 3u * uplane;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        cp_async4_stream(dst + (REC_F + a * 6 + c) * 256, P.nmap_prev + (o + c * uplane), stream_policy);
                        cp_async4_stream(dst + (REC_F + a * 6 + 3 + c) * 256, P.vmap_prev + (o + c * uplane), stream_policy);
                    }
                }
            }
            cp_async_commit();
        };
        auto idx_at = [&](int j) {
            const int p = base + j * 256 + tid;
            return (j < nitems && p < npix) ? P.rec_idx[p] : -1;
        };
        int qs[ST];  // qs[i] = matched index of pixel j + i
#pragma unroll
        for (int i = 0; i < ST; ++i) qs[i] = idx_at(i);
#pragma unroll
        for (int i = 0; i < ST - 1; ++i) issue(i, base + i * 256 + tid, qs[i]);
        int st = 0;  // stage that holds pixel j
        for (int j = 0; j < nitems; ++j) {
            const int p = base + j * 256 + tid;
            int st_in = st + ST - 1;
            if (st_in >= ST) st_in -= ST;
            issue(st_in, p + (ST - 1) * 256, qs[ST - 1]);
            const int q = qs[0];
#pragma unroll
            for (int i = 0; i < ST - 1; ++i) qs[i] = qs[i + 1];
            qs[ST - 1] = idx_at(j + ST);
            const int cur = st;
            st = (st + 1 == ST) ? 0 : st + 1;
            cp_async_wait<ST - 1>();  // the copies of pixel j have landed
            if (q >= 0) {
                const float *in = s_in + (size_t) cur * DERIV_IN * 256 + tid;
                const float4 *in4 = reinterpret_cast<const float4 *>(s_in + (size_t) cur * DERIV_IN * 256) + tid;
                const float4 f0 = in4[0], f1 = in4[256], f2 = in4[512], f3 = in4[768];
                const float vc[3] = {f0.x, f0.y, f0.z};
                const float s[3] = {f0.w, f1.x, f1.y};
                const float n[3] = {f1.z, f1.w, f2.x};
                const float e[3] = {f2.y, f2.z, f2.w};
                const float r[7] = {f3.x, f3.y, f3.z, n[0], n[1], n[2], f3.w};
                float ds[3][3], dn[3][3], de[3][3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const float *m = s_pose[a];
                    ds[a][0] = fmaf(m[0], vc[0], fmaf(m[1], vc[1], fmaf(m[2], vc[2], m[9])));
                    ds[a][1] = fmaf(m[3], vc[0], fmaf(m[4], vc[1], fmaf(m[5], vc[2], m[10])));
                    ds[a][2] = fmaf(m[6], vc[0], fmaf(m[7], vc[1], fmaf(m[8], vc[2], m[11])));
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        dn[a][c] = in[(REC_F + a * 6 + c) * 256];
                        de[a][c] = in[(REC_F + a * 6 + 3 + c) * 256] - ds[a][c];
                    }
                }
                float d[3][7];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    cross3(ds[a], n, d[a]);        // d(s x n) = ds x n + s x dn
                    cross3_add(s, dn[a], d[a]);
                    d[a][3] = dn[a][0];
                    d[a][4] = dn[a][1];
                    d[a][5] = dn[a][2];
                    d[a][6] = dot3(dn[a], e) + dot3(n, de[a]);  // d(n . (d - s))
                }
                if (C == 3) {  // eps1eps2 cross terms
                    cross3_add(ds[0], dn[1], d[2]);
                    cross3_add(ds[1], dn[0], d[2]);
                    d[2][6] += dot3(dn[0], de[1]) + dot3(dn[1], de[0]);
                }
                accumulate_products<C>(acc, r, d[0], d[1], d[2], std::make_integer_sequence<int, 27>());
            }
            if ((j & 31) == 31) flush();  // warp-uniform: j is
        }
        if (nitems & 31) flush();
        cp_async_wait<0>();
        __syncthreads();
        // ---------------- the 8 warps in order -> one 81-value partial of this CTA for this group
        const int first_b = deriv_owner((long long) group * P.chunks, U, B);
        const int writers = deriv_owner((long long) (group + 1) * P.chunks - 1, U, B) - first_b + 1;
        double *part = P.dpartials + ((size_t) group * P.max_writers + (blockIdx.x - first_b)) * 81;
        if (tid < 81) {
            const int a = tid / 27, e = tid % 27;
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) sum += s_tot[a * 256 + w * 32 + e];
            part[tid] = sum;
        }
        // ---------------- tail: the last CTA of a direction group to arrive adds the partials of its writers in CTA order
        // and runs the Gauss-Newton step of its direction(s) from shared memory: no separate finish / solve launches.
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = atomicAdd(P.group_ticket + group, 1u) == (unsigned) writers - 1u;
        __syncthreads();
        if (!s_last) continue;
        __threadfence();
        // warp w adds its contiguous eighth of the writers' partials in order for products lane, lane + 32, lane + 64
        // (coalesced 648-byte rows, independent L2 loads in flight), then the eight segment sums are added in warp order
        double(*s_seg)[81] = reinterpret_cast<double(*)[81]>(s_raw);   // [8][81]; the staging area is idle here
        double *s_sums = reinterpret_cast<double *>(s_raw) + 8 * 81;   // [27 real][81 of this group]
        {
            const int seg = (writers + 7) / 8, w0 = warp * seg, w1 = min(writers, w0 + seg);
            double sum[3] = {0.0, 0.0, 0.0};
            const double *src = P.dpartials + (size_t) group * P.max_writers * 81 + lane;
#pragma unroll 4
            for (int w = w0; w < w1; ++w) {
                const double *row = src + (size_t) w * 81;
                sum[0] += __ldcg(row);
                sum[1] += __ldcg(row + 32);
                if (lane < 17) sum[2] += __ldcg(row + 64);
            }
            s_seg[warp][lane] = sum[0];
            s_seg[warp][lane + 32] = sum[1];
            if (lane < 17) s_seg[warp][lane + 64] = sum[2];
        }
        __syncthreads();
        if (tid < 81) {
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) sum += s_seg[w][tid];
            s_sums[27 + tid] = sum;
            const int comp = comp0 + tid / 27;
            if (comp < P.ncomp) P.sums[(size_t) (1 + comp) * 27 + tid % 27] = sum;
        } else if (tid >= 96 && tid < 96 + 27) {
            s_sums[tid - 96] = __ldcg(P.sums + (tid - 96));  // real sums of icp_assoc_kernel (previous launch)
        }
        if (tid == 0) P.group_ticket[group] = 0u;
        __syncthreads();
        if (S.pose_out && tid < (C == 3 ? 1 : 3)) {
            const int q = C == 3 ? group : group * 3 + tid;
            if (q < S.dirs) icp_solve_direction<C>(S, q, s_sums, s_sums + 27 + (C == 3 ? 0 : tid * 27));
        }
        __syncthreads();  // s_sums aliases the staging area the next segment copies into
    }
}

// ---------------------------------------------------------------------------------------------------------------
// pass 4: the Gauss-Newton step on the device (KinectFusionReconstruction.cpp:203-224).
struct cplx {
    double re, im;
};
XS_DEV cplx cmul(cplx a, cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
XS_DEV cplx cconj(cplx a) { return {a.re, -a.im}; }
XS_DEV cplx csub(cplx a, cplx b) { return {a.re - b.re, a.im - b.im}; }
XS_DEV cplx cdiv(cplx a, cplx b) {
    const double d = b.re * b.re + b.im * b.im;
    return {(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
}

// unpack the upper-triangular product order of ICP.cu:419-428 into a symmetric 6x6 and b
XS_DEV void unpack_sums(const double *v, double (&A)[6][6], double (&b)[6]) {
    int shift = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = i; j < 7; ++j) {
            const double val = v[shift++];
            if (j == 6)
                b[i] = val;
            else
                A[i][j] = A[j][i] = val;
        }
}

// A.real().determinant() (KinectFusionReconstruction.cpp:203): Gaussian elimination with partial pivoting.  All loops
// are fully unrolled with static row indices (the pivot is bubbled into row k by predicated swaps) so the matrix
// lives in registers.
XS_DEV double det6_dev(const double (&A)[6][6]) {
    double M[6][6];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) M[i][j] = A[i][j];
    double det = 1.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
#pragma unroll
        for (int i = k + 1; i < 6; ++i) {
            if (fabs(M[i][k]) > fabs(M[k][k])) {
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const double t = M[i][j];
                    M[i][j] = M[k][j];
                    M[k][j] = t;
                }
                det = -det;
            }
        }
        if (M[k][k] == 0.0) return 0.0;
        det *= M[k][k];
        const double inv = 1.0 / M[k][k];
#pragma unroll
        for (int i = k + 1; i < 6; ++i) {
            const double f = M[i][k] * inv;
#pragma unroll
            for (int j = k + 1; j < 6; ++j) M[i][j] -= f * M[k][j];
        }
    }
    return det;
}

// Real Cholesky factor of the lower triangle of A as Eigen's unblocked LLT computes it (KinectFusionReconstruction.cpp:
// 211; with zero imaginary parts the Hermitian factorisation of SURVEY.md 0.6 is the real one).  L holds the factor,
// Linv the reciprocals of its diagonal.  A non-positive pivot stops the factorisation like Eigen does.
struct Chol6 {
    double L[6][6];
    double inv[6];
};
XS_DEV bool chol6_factor(const double (&A)[6][6], Chol6 &F) {  // false: a pivot was not positive (the factor is partial, as Eigen's)
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) F.L[i][j] = A[i][j];
#pragma unroll
    for (int k = 0; k < 6; ++k) F.inv[k] = 1.0 / A[k][k];
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double xk = F.L[k][k];
#pragma unroll
        for (int j = 0; j < k; ++j) xk -= F.L[k][j] * F.L[k][j];
        ok = ok && xk > 0.0;
        if (ok) {
            const double r = rsqrt(xk);
            F.L[k][k] = xk * r;
            F.inv[k] = r;
#pragma unroll
            for (int i = k + 1; i < 6; ++i) {
                double sacc = F.L[i][k];
#pragma unroll
                for (int j = 0; j < k; ++j) sacc -= F.L[i][j] * F.L[k][j];
                F.L[i][k] = sacc * r;
            }
        }
    }
    return ok;
}
XS_DEV void chol6_solve(const Chol6 &F, const double (&b)[6], double (&x)[6]) {
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double sacc = b[i];
#pragma unroll
        for (int j = 0; j < i; ++j) sacc -= F.L[i][j] * y[j];
        y[i] = sacc * F.inv[i];
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        double sacc = y[i];
#pragma unroll
        for (int j = i + 1; j < 6; ++j) sacc -= F.L[j][i] * x[j];
        x[i] = sacc * F.inv[i];
    }
}

// Eigen 3.4 LLT<Matrix<complex<double>,6,6>,Lower>::solve (unblocked, n < 32) with a non-zero imaginary part: the
// factor is built from the LOWER triangle with real(A_kk) on the diagonal and conj() in the updates, i.e. the
// complex-symmetric A of ICP.cu:427 is treated as Hermitian (KinectFusionReconstruction.cpp:211).
XS_DEV void llt_hermitian_solve6_dev(const double (&Ar)[6][6], const double (&Ai)[6][6], const double (&br)[6],
                                     const double (&bi)[6], cplx (&x)[6]) {
    cplx L[6][6];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) L[i][j] = {Ar[i][j], Ai[i][j]};
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double xk = L[k][k].re;
#pragma unroll
        for (int j = 0; j < k; ++j) xk -= L[k][j].re * L[k][j].re + L[k][j].im * L[k][j].im;
        ok = ok && xk > 0.0;
        if (ok) {
            const double r = rsqrt(xk);
            L[k][k] = {xk * r, 0.0};
#pragma unroll
            for (int i = k + 1; i < 6; ++i) {
                cplx sacc = L[i][k];
#pragma unroll
                for (int j = 0; j < k; ++j) sacc = csub(sacc, cmul(L[i][j], cconj(L[k][j])));
                L[i][k] = {sacc.re * r, sacc.im * r};
            }
        }
    }
    cplx y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        cplx sacc = {br[i], bi[i]};
#pragma unroll
        for (int j = 0; j < i; ++j) sacc = csub(sacc, cmul(L[i][j], y[j]));
        y[i] = cdiv(sacc, L[i][i]);
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        cplx sacc = y[i];
#pragma unroll
        for (int j = i + 1; j < 6; ++j) sacc = csub(sacc, cmul(cconj(L[j][i]), x[j]));
        x[i] = cdiv(sacc, cconj(L[i][i]));
    }
}

XS_DEV void matvec6_dev(const double (&A)[6][6], const double (&x)[6], double (&y)[6]) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double sacc = 0;
#pragma unroll
        for (int j = 0; j < 6; ++j) sacc += A[i][j] * x[j];
        y[i] = sacc;
    }
}

// sin / cos of a batched angle; the real part is evaluated in double and rounded (matches the host's correctly
// rounded sinf / cosf), derivative components by the chain rule
// coef_only: the real parts are only coefficients of the derivative formulas (a derivative tail under split chains never
// writes a real part), so single-precision sincosf replaces the two double-precision evaluations
template <int C> XS_DEV void jsincos(const Jet<C, 1> &a, Jet<C, 1> &sn, Jet<C, 1> &cs, bool coef_only = false) {
    float s0, c0;
    if (coef_only) {
        sincosf(a.v, &s0, &c0);
    } else {
        s0 = (float) sin((double) a.v);
        c0 = (float) cos((double) a.v);
    }
    sn.v = s0;
    cs.v = c0;
    sn.d[0] = c0 * a.d[0];
    cs.d[0] = -s0 * a.d[0];
    if (C == 3) {
        sn.d[1 % C] = c0 * a.d[1 % C];
        cs.d[1 % C] = -s0 * a.d[1 % C];
        sn.d[2 % C] = c0 * a.d[2 % C] - s0 * a.d[0] * a.d[1 % C];
        cs.d[2 % C] = -s0 * a.d[2 % C] - c0 * a.d[0] * a.d[1 % C];
    }
}
template <int C> struct JMat3 {
    Jet<C, 1> m[3][3];
};
// rotation about a coordinate axis: Eigen AngleAxis::toRotationMatrix() specialised to a unit axis (host_jet.h)
template <int C> XS_DEV JMat3<C> jaxis_rotation(const Jet<C, 1> &angle, int axis, bool coef_only = false) {
    Jet<C, 1> sn, cs;
    jsincos<C>(angle, sn, cs, coef_only);
    JMat3<C> R;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R.m[i][j] = jconst<C, 1>(0.f);
    for (int i = 0; i < 3; ++i) R.m[i][i] = (i == axis) ? (jconst<C, 1>(1.f) - cs) + cs : cs;
    const int a = (axis + 1) % 3, b = (axis + 2) % 3;
    R.m[a][b] = -sn;
    R.m[b][a] = sn;
    return R;
}
template <int C> XS_DEV JMat3<C> jmatmul(const JMat3<C> &a, const JMat3<C> &b) {
    JMat3<C> r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            Jet<C, 1> sacc = a.m[i][0] * b.m[0][j];
            for (int k = 1; k < 3; ++k) sacc = sacc + a.m[i][k] * b.m[k][j];
            r.m[i][j] = sacc;
        }
    return r;
}

// One direction q of the Gauss-Newton step (direction 0 also owns the real part).  Called by the last CTA of a direction
// group in icp_deriv_kernel, or by icp_solve_kernel when there are no derivative components.
// real_sums: the 27 real sums; comp_sums: the C x 27 sums of this direction's components (shared memory in both callers).
template <int C>
__device__ __noinline__ void icp_solve_direction(const SolveParams &P, int q, const double *real_sums, const double *comp_sums) {
    typedef Jet<C, 1> J;
    const bool has_dir = q < P.dirs;
    if (P.log) {
        if (q == 0)
            for (int i = 0; i < 27; ++i) P.log[i] = real_sums[i];
        if (has_dir)
            for (int i = 0; i < 27 * C; ++i) P.log[27 * (1 + q * C) + i] = comp_sums[i];
    }
    const int st0 = __ldcg(P.status), st1 = __ldcg(P.status + 1);
    const bool owns_real = q == 0 && !P.deriv_only;
    if (st0 != 0 || st1 != 0) {  // [0] sticky flag of earlier iterations, [1] written by an earlier launch
        if (owns_real) P.status[0] = st0 != 0 ? st0 : st1;
        return;
    }
    double A[6][6], b[6];
    unpack_sums(real_sums, A, b);
    Chol6 F;
    double xr[6];
    if (P.deriv_only && P.real_cache) {
        // the real step of this iteration has completed (stream order): a degenerate system was caught by the status test
        // above; its factor and solution are loaded instead of being recomputed
        const double *rc = P.real_cache;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
            for (int j = 0; j < 6; ++j) F.L[i][j] = __ldcg(rc + i * 6 + j);
            F.inv[i] = __ldcg(rc + 36 + i);
            xr[i] = __ldcg(rc + 42 + i);
        }
    } else {
        // A.real().determinant() guard (KinectFusionReconstruction.cpp:203-210).  For a positive definite A the determinant is
        // the squared product of the Cholesky pivots, which the solve needs anyway; only when the factorisation meets a
        // non-positive pivot (or NaN) is the determinant taken by elimination.
        const bool spd = chol6_factor(A, F);
        double det = 1.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) det *= F.L[i][i] * F.L[i][i];
        if (!spd || !isfinite(det)) det = det6_dev(A);
        if (fabs(det) < 1e-15 || isnan(det)) {
            // every thread of every block computes the same det and takes this branch; the flag is only read at kernel
            // entry by later launches
            if (owns_real) P.status[1] = isnan(det) ? 2 : 1;
            return;
        }
        chol6_solve(F, b, xr);  // zero-seed solve: the canonical real part
        if (owns_real && P.real_cache) {
#pragma unroll
            for (int i = 0; i < 6; ++i) {
#pragma unroll
                for (int j = 0; j < 6; ++j) P.real_cache[i * 6 + j] = F.L[i][j];
                P.real_cache[36 + i] = F.inv[i];
                P.real_cache[42 + i] = xr[i];
            }
        }
    }
    J x[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] = jconst<C, 1>((float) xr[i]);
    if (has_dir) {
        if (C == 1 && P.solve_mode == XS_SOLVE_EIGEN_LLT) {
            // one Hermitian-LLT solve with this direction's imaginary part, as the reference's one-direction run
            double Ai[6][6], bi[6];
            unpack_sums(comp_sums, Ai, bi);
            cplx xq[6];
            llt_hermitian_solve6_dev(A, Ai, b, bi, xq);
            for (int i = 0; i < 6; ++i) x[i].d[0] = (float) xq[i].im;
        } else {
            // analytic: x_a = A^-1 (b_a - A_a x);  x_12 = A^-1 (b_12 - A_12 x - A_1 x_2 - A_2 x_1)
            double xa[3][6];
#pragma unroll
            for (int a = 0; a < (C == 3 ? 2 : 1); ++a) {
                double Aa[6][6], ba[6], t[6], rhs[6];
                unpack_sums(comp_sums + a * 27, Aa, ba);
                if (C == 3 && P.solve_mode == XS_SOLVE_EIGEN_LLT) {
                    // bicomplex direction in LLT mode: each first-order component is what a one-direction complex run of
                    // the reference yields for that imaginary part (Hermitian LLT quirk); the eps1eps2 component below is
                    // the truncated-algebra second derivative built on them
                    cplx xq[6];
                    llt_hermitian_solve6_dev(A, Aa, b, ba, xq);
                    for (int i = 0; i < 6; ++i) xa[a][i] = xq[i].im;
                    continue;
                }
                matvec6_dev(Aa, xr, t);
                for (int i = 0; i < 6; ++i) rhs[i] = ba[i] - t[i];
                chol6_solve(F, rhs, xa[a]);
            }
            if (C == 3) {
                double A1[6][6], A2[6][6], A12[6][6], b1[6], b2[6], b12[6], t[6], rhs[6];
                unpack_sums(comp_sums, A1, b1);
                unpack_sums(comp_sums + 27, A2, b2);
                unpack_sums(comp_sums + 54, A12, b12);
                for (int i = 0; i < 6; ++i) rhs[i] = b12[i];
                matvec6_dev(A12, xr, t);
                for (int i = 0; i < 6; ++i) rhs[i] -= t[i];
                matvec6_dev(A1, xa[1], t);
                for (int i = 0; i < 6; ++i) rhs[i] -= t[i];
                matvec6_dev(A2, xa[0], t);
                for (int i = 0; i < 6; ++i) rhs[i] -= t[i];
                chol6_solve(F, rhs, xa[2]);
            }
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int a = 0; a < C; ++a) x[i].d[a] = (float) xa[a][i];
        }
    }
    // ---- pose update, KinectFusionReconstruction.cpp:212-224
    JMat3<C> Rcurr;
    J tcurr[3];
    const float *pr = P.pose_in;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            Rcurr.m[i][j].v = pr[i * 3 + j];
            for (int a = 0; a < C; ++a) Rcurr.m[i][j].d[a] = has_dir ? pr[(size_t) (1 + q * C + a) * 12 + i * 3 + j] : 0.f;
        }
        tcurr[i].v = pr[9 + i];
        for (int a = 0; a < C; ++a) tcurr[i].d[a] = has_dir ? pr[(size_t) (1 + q * C + a) * 12 + 9 + i] : 0.f;
    }
    // Rinc = Rz(gamma) * Ry(beta) * Rx(alpha)
    const bool co = P.deriv_only != 0;
    const JMat3<C> Rinc = jmatmul(jmatmul(jaxis_rotation<C>(x[2], 2, co), jaxis_rotation<C>(x[1], 1, co)), jaxis_rotation<C>(x[0], 0, co));
    J tn[3];
    for (int i = 0; i < 3; ++i) {
        J sacc = Rinc.m[i][0] * tcurr[0];
        for (int k = 1; k < 3; ++k) sacc = sacc + Rinc.m[i][k] * tcurr[k];
        tn[i] = sacc + x[3 + i];
    }
    const JMat3<C> Rn = jmatmul(Rinc, Rcurr);
    float *pw = P.pose_out;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            if (owns_real) pw[i * 3 + j] = Rn.m[i][j].v;
            if (has_dir)
                for (int a = 0; a < C; ++a) pw[(size_t) (1 + q * C + a) * 12 + i * 3 + j] = Rn.m[i][j].d[a];
        }
        if (owns_real) pw[9 + i] = tn[i].v;
        if (has_dir)
            for (int a = 0; a < C; ++a) pw[(size_t) (1 + q * C + a) * 12 + 9 + i] = tn[i].d[a];
    }
}

// the step without derivative components (dirs == 0): one thread
template <int C> __global__ void __launch_bounds__(32) icp_solve_kernel(const SolveParams P) {
    __shared__ double s_real[27];
    if (threadIdx.x < 27) s_real[threadIdx.x] = P.sums[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) icp_solve_direction<C>(P, 0, s_real, s_real);
}

// ---------------------------------------------------------------------------------------------------------------
// pass 2 for a Hessian batch (kind 2).  A task is one parameter i (first-order component i: gathers F_i of the previous
// maps) or one listed pair k = (i, j) (second-order component n + k: gathers F_i, F_j and S_k); either way a thread
// accumulates ONE set of 27 sums - those of the task's own component - so 27 instead of 81 FP32 accumulators per thread
// (three CTAs per SM instead of two) and every first-order plane contributes its sums once instead of once per pair.
//   first order :  d(r_a r_b) = r_a F(r_b) + F(r_a) r_b
//   pair        :  d(r_a r_b) = r_a S(r_b) + S(r_a) r_b + F_i(r_a) F_j(r_b) + F_j(r_a) F_i(r_b)        (xs_batch.h)
// Work decomposition, staging, flush and per-task partial reduction are those of icp_deriv_kernel; the CTA that completes the
// LAST task's sums (done_ticket) runs the Gauss-Newton step of every component: first-order solves first (their solutions
// are needed by the pairs), then the pairs.
// A task of the Hessian-batch derivative pass: either one parameter (first-order sums of component i) or up to HPMAX pairs
// that share their first parameter i, (i, j_0) .. (i, j_np-1) with second-order components s_0 .. s_np-1.  The pairs of a task
// share the association record and the gathers of F_i - the derivative pass is bound by L2 -> SM traffic and instruction
// issue, not by HBM (profiles/r02_ab_table.md).
//
// Two forms of the pair sums:
//   FULL     the 27 second-order sums A_ij, b_ij of every pair (what the ICP log and the seam-level estimateCombined return);
//   REDUCED  under split chains the real solution x of the iteration is known before the derivative pass (the real step ran
//            ahead), and the Gauss-Newton step needs A_ij, b_ij only through g_ij = b_ij - A_ij x.  With rho = row_6 - sum_j
//            row_j x_j (the linearised residual after the step) and its first / second-order parts,
//              g_ij = sum_pixels [ S(row) rho + row S(rho) + F_i(row) F_j(rho) + F_j(row) F_i(rho) ]   (6 values)
//            i.e. 6 instead of 27 accumulators and 52 instead of 108 FMAs per pixel and pair, and three pairs per task.
constexpr int HPMAX = 3;
struct HTask {
    int i;          // parameter: first-order component i
    int np;         // 0: first-order task (sums of component i); 1..HPMAX: pair task
    int j[HPMAX];   // second parameters
    int s[HPMAX];   // second-order component indices (n + pair index)
};
__host__ __device__ constexpr int deriv_h_in(int hp) { return 12 + 6 + hp * 12; }  // staged floats per pixel: record (3 float4) | F_i (dn, dv) | per pair F_j, S
constexpr size_t deriv_h_smem(int stages, int hp, bool reduced) {
    return (size_t) stages * deriv_h_in(hp) * 256 * sizeof(float) + (reduced ? 1 : hp) * 256 * sizeof(double);
}

template <int... E>
XS_DEV void accumulate_first(float *acc, const float (&r)[7], const float (&d0)[7], std::integer_sequence<int, E...>) {
    ((acc[E] = fmaf(r[tri_i(E)], d0[tri_j(E)], fmaf(d0[tri_i(E)], r[tri_j(E)], acc[E]))), ...);
}
template <int... E>
XS_DEV void accumulate_pair(float *acc, const float (&r)[7], const float (&d0)[7], const float (&d1)[7], const float (&d2)[7],
                            std::integer_sequence<int, E...>) {
    ((acc[E] = fmaf(r[tri_i(E)], d2[tri_j(E)],
                    fmaf(d2[tri_i(E)], r[tri_j(E)], fmaf(d0[tri_i(E)], d1[tri_j(E)], fmaf(d1[tri_i(E)], d0[tri_j(E)], acc[E]))))),
     ...);
}
// rho-type contraction of a row (or of a row's derivative) with the solution: v_6 - sum_j v_j x_j
// rotation part of a pose component (row-major in three float4: R (9), t (3)) times a 3-vector, added to o
XS_DEV void rot_add(const float4 *mm4, const float (&v)[3], float (&o)[3]) {
    const float4 m0 = mm4[0], m1 = mm4[1], m2 = mm4[2];
    o[0] += fmaf(m0.x, v[0], fmaf(m0.y, v[1], m0.z * v[2]));
    o[1] += fmaf(m0.w, v[0], fmaf(m1.x, v[1], m1.y * v[2]));
    o[2] += fmaf(m1.z, v[0], fmaf(m1.w, v[1], m2.x * v[2]));
}
XS_DEV float rho_of(const float (&v)[7], const float (&x)[6]) {
    return v[6] - fmaf(v[0], x[0], fmaf(v[1], x[1], fmaf(v[2], x[2], fmaf(v[3], x[3], fmaf(v[4], x[4], v[5] * x[5])))));
}
// derivative of the row [s x n, n, n . e] for one component: ds = dpose * vc + dt, (dn, dv) gathered at the matched pixel
// (ex: what the derivative of the current-frame vertex adds to ds when a parameter moves the intrinsics: R dvc + cross terms)
XS_DEV void row_first(const float4 *mm4, const float (&vc)[3], const float (&sv)[3], const float (&nv)[3], const float (&ev)[3],
                      const float (&dn)[3], const float (&dv)[3], const float (&ex)[3], float (&ds)[3], float (&de)[3], float (&d)[7]) {
    const float4 m0 = mm4[0], m1 = mm4[1], m2 = mm4[2];  // R row-major (9), t (3): broadcast 128-bit shared loads
    ds[0] = fmaf(m0.x, vc[0], fmaf(m0.y, vc[1], fmaf(m0.z, vc[2], m2.y))) + ex[0];
    ds[1] = fmaf(m0.w, vc[0], fmaf(m1.x, vc[1], fmaf(m1.y, vc[2], m2.z))) + ex[1];
    ds[2] = fmaf(m1.z, vc[0], fmaf(m1.w, vc[1], fmaf(m2.x, vc[2], m2.w))) + ex[2];
#pragma unroll
    for (int c = 0; c < 3; ++c) de[c] = dv[c] - ds[c];
    cross3(ds, nv, d);  // d(s x n) = ds x n + s x dn
    cross3_add(sv, dn, d);
    d[3] = dn[0];
    d[4] = dn[1];
    d[5] = dn[2];
    d[6] = dot3(dn, ev) + dot3(nv, de);  // d(n . (d - s))
}

// CURR: some parameter moves the intrinsics, so the current-frame vertex has derivative components (xs_batch.h):
// s = R vc + t then has d s = dR vc + dt + R dvc, and for a pair also dR_i dvc_j + dR_j dvc_i.  The components of vc are not
// read from memory: with vx = z (u - cx) / fx (Map.cu:19-21) they are linear in the real vertex,
//   F_p(vx) = -(dcx_p z + dfx_p vx) / fx,   S_ij(vx) = ((dcx_i dfx_j + dcx_j dfx_i) z + 2 dfx_i dfx_j vx) / fx^2,   d vz = 0
// (vy likewise), with coefficients that are the same at every pyramid level (seed and focal length scale together) and are
// formed once per task (s_cur).
template <int ST, int MINB, int HPK, bool REDUCED, bool CURR>
__global__ void __launch_bounds__(256, MINB) icp_deriv_h_kernel(const IcpParams P, const SolveParams S) {
    constexpr int IN = deriv_h_in(HPK);
    constexpr int NACC = REDUCED ? (HPK * 6 > 27 ? HPK * 6 : 27) : HPK * 27;  // accumulators per thread
    constexpr int NPART = REDUCED ? 27 : HPK * 27;                             // values of a (task, writer) partial
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ float4 s_pose[1 + 2 * HPK][3];  // F_i | F_j0, S_0 | F_j1, S_1 | ...
    __shared__ float4 s_real_pose[3];          // CURR: the real current pose
    __shared__ float4 s_cur[1 + 2 * HPK];      // CURR: d vc = (x z + y vx, z z + w vy, 0) per gather slot
    __shared__ float s_x[6];
    __shared__ bool s_last, s_final;
    float *s_in = reinterpret_cast<float *>(s_raw);  // [stage][IN][256]
    double *s_tot = reinterpret_cast<double *>(s_raw + (size_t) ST * IN * 256 * sizeof(float));  // [REDUCED ? 1 : HPK][256]
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const size_t plane = (size_t) P.rows * P.cols;
    const unsigned uplane = (unsigned) plane;
    const int npix = P.rows * P.cols;
    const int ntasks = P.groups;
    const unsigned long long stream_policy = l2_policy_evict_first(), keep_policy = l2_policy_evict_last();
    constexpr int R = ST + 2;  // matched indices kept ahead in registers (they come from L2)
    const long long U = (long long) P.groups * P.chunks, B = gridDim.x;
    const long long u_begin = blockIdx.x * U / B, u_end = (blockIdx.x + 1) * U / B;
    if (REDUCED && tid < 6) s_x[tid] = (float) __ldcg(S.real_cache + 42 + tid);  // real solution of this iteration (the real step has completed)
    if (CURR && tid >= 32 && tid < 44) reinterpret_cast<float *>(&s_real_pose[0])[tid - 32] = P.pose_curr[tid - 32];
    bool final_cta = false;
    for (long long u = u_begin; u < u_end;) {
        const int task = (int) (u / P.chunks), c_begin = (int) (u % P.chunks);
        const int c_end = (int) min((long long) P.chunks, c_begin + (u_end - u));
        u += c_end - c_begin;
        const HTask T = P.htasks[task];
        const int np = T.np, ci = T.i;
        // component of every gather slot as scalars (an indexed array would live in local memory)
        int cj[HPK], cs[HPK];
#pragma unroll
        for (int h = 0; h < HPK; ++h) cj[h] = T.j[h], cs[h] = T.s[h];
        // CURR: slots of these components in the current-frame derivative maps (-1: identically zero)
        int ki = -1, kj[HPK], ks[HPK];
#pragma unroll
        for (int h = 0; h < HPK; ++h) kj[h] = ks[h] = -1;
        if (CURR) {
            ki = __ldg(P.batch.cslot + ci);
#pragma unroll
            for (int h = 0; h < HPK; ++h) kj[h] = __ldg(P.batch.cslot + cj[h]), ks[h] = __ldg(P.batch.cslot + cs[h]);
        }
        __syncthreads();  // s_pose / s_tot of the previous segment are no longer read
        if (tid < 12 * (1 + 2 * HPK)) {
            const int a = tid / 12, e = tid % 12;
            int ca = ci;
#pragma unroll
            for (int h = 0; h < HPK; ++h) {
                if (a == 1 + 2 * h) ca = cj[h];
                if (a == 2 + 2 * h) ca = cs[h];
            }
            reinterpret_cast<float *>(&s_pose[a][0])[e] = P.pose_curr[(size_t) (1 + ca) * 12 + e];
        }
        if (CURR && tid >= 128 && tid < 128 + 1 + 2 * HPK) {
            const int a = tid - 128;
            const float4 *din = reinterpret_cast<const float4 *>(P.batch.dintr);  // (dfx, dfy, dcx, dcy) at level 0
            const float gx = P.batch.gx0, gy = P.batch.gy0;
            const float4 di = __ldg(din + ci);
            float4 o = make_float4(-di.z * gx, -di.x * gx, -di.w * gy, -di.y * gy);
#pragma unroll
            for (int h = 0; h < HPK; ++h) {
                if (h < np) {
                    const float4 dj = __ldg(din + cj[h]);
                    if (a == 1 + 2 * h) o = make_float4(-dj.z * gx, -dj.x * gx, -dj.w * gy, -dj.y * gy);
                    if (a == 2 + 2 * h)
                        o = make_float4((di.z * dj.x + dj.z * di.x) * gx * gx, 2.f * di.x * dj.x * gx * gx, (di.w * dj.y + dj.w * di.y) * gy * gy,
                                        2.f * di.y * dj.y * gy * gy);
                }
            }
            s_cur[a] = o;
        }
#pragma unroll
        for (int h = 0; h < (REDUCED ? 1 : HPK); ++h) s_tot[h * 256 + tid] = 0.0;
        __syncthreads();
        float x[6];
#pragma unroll
        for (int e = 0; e < 6; ++e) x[e] = REDUCED ? s_x[e] : 0.f;
        float acc[NACC];
#pragma unroll
        for (int e = 0; e < NACC; ++e) acc[e] = 0.f;
        auto flush = [&]() {
            if (REDUCED) {  // first-order task: 27 sums; pair task: HPK x 6 values, one transpose-reduce either way
                float v[32];
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] = e < NACC ? acc[e] : 0.f;
                s_tot[tid] += (double) warp_transpose_reduce_f(v);
#pragma unroll
                for (int e = 0; e < NACC; ++e) acc[e] = 0.f;
            } else {
#pragma unroll
                for (int h = 0; h < HPK; ++h) {
                    if (h == 0 || np > h) {  // block-uniform
                        float v[32];
#pragma unroll
                        for (int e = 0; e < 32; ++e) v[e] = e < 27 ? acc[h * 27 + e] : 0.f;
                        s_tot[h * 256 + tid] += (double) warp_transpose_reduce_f(v);
#pragma unroll
                        for (int e = 0; e < 27; ++e) acc[h * 27 + e] = 0.f;
                    }
                }
            }
        };
        const int base = c_begin * 256 * P.ppt;
        const int nitems = (c_end - c_begin) * P.ppt;
        auto gather = [&](float *dst, int slot, int comp, unsigned q, bool stream) {
            const unsigned o = q + (unsigned) (1 + comp) * 3u * uplane;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (stream) {  // the task's own second-order planes are read once: evict-first keeps the re-read planes in L2
                    cp_async4_stream(dst + (12 + slot * 6 + c) * 256, P.nmap_prev + (o + c * uplane), stream_policy);
                    cp_async4_stream(dst + (12 + slot * 6 + 3 + c) * 256, P.vmap_prev + (o + c * uplane), stream_policy);
                } else {
                    cp_async4_keep(dst + (12 + slot * 6 + c) * 256, P.nmap_prev + (o + c * uplane), keep_policy);
                    cp_async4_keep(dst + (12 + slot * 6 + 3 + c) * 256, P.vmap_prev + (o + c * uplane), keep_policy);
                }
            }
        };
        auto issue = [&](int stage, int p, int q) {
            if (q >= 0) {
                float *dst = s_in + (size_t) stage * IN * 256 + tid;
                float4 *dst4 = reinterpret_cast<float4 *>(s_in + (size_t) stage * IN * 256) + tid;  // [3][256] float4
                const float4 *f = P.rec_f + (size_t) p * 4;  // (vc, s.x) (s.yz, n.xy) (n.z, e): the fourth (s x n, n . e) is recomputed
#pragma unroll
                for (int i = 0; i < 3; ++i) cp_async16(dst4 + i * 256, f + i);
                gather(dst, 0, ci, (unsigned) q, np == 0);
#pragma unroll
                for (int h = 0; h < HPK; ++h) {
                    if (h < np) {  // block-uniform
                        gather(dst, 1 + 2 * h, cj[h], (unsigned) q, false);
                        gather(dst, 2 + 2 * h, cs[h], (unsigned) q, true);
                    }
                }
            }
            cp_async_commit();
        };
        auto idx_at = [&](int j) {
            const int p = base + j * 256 + tid;
            return (j < nitems && p < npix) ? P.rec_idx[p] : -1;
        };
        int qs[R];
#pragma unroll
        for (int i = 0; i < R; ++i) qs[i] = idx_at(i);
#pragma unroll
        for (int i = 0; i < ST - 1; ++i) issue(i, base + i * 256 + tid, qs[i]);
        int st = 0;
        for (int j = 0; j < nitems; ++j) {
            const int p = base + j * 256 + tid;
            int st_in = st + ST - 1;
            if (st_in >= ST) st_in -= ST;
            issue(st_in, p + (ST - 1) * 256, qs[ST - 1]);
            const int q = qs[0];
#pragma unroll
            for (int i = 0; i < R - 1; ++i) qs[i] = qs[i + 1];
            qs[R - 1] = idx_at(j + R);
            const int cur = st;
            st = (st + 1 == ST) ? 0 : st + 1;
            cp_async_wait<ST - 1>();
            if (q >= 0) {
                const float *in = s_in + (size_t) cur * IN * 256 + tid;
                const float4 *in4 = reinterpret_cast<const float4 *>(s_in + (size_t) cur * IN * 256) + tid;
                const float4 f0 = in4[0], f1 = in4[256], f2 = in4[512];
                const float vc[3] = {f0.x, f0.y, f0.z};
                const float sv[3] = {f0.w, f1.x, f1.y};
                const float nv[3] = {f1.z, f1.w, f2.x};
                const float ev[3] = {f2.y, f2.z, f2.w};
                float r[7];
                cross3(sv, nv, r);
                r[3] = nv[0], r[4] = nv[1], r[5] = nv[2];
                r[6] = dot3(nv, ev);
                auto load6 = [&](int slot, float (&dn)[3], float (&dv)[3]) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        dn[c] = in[(12 + slot * 6 + c) * 256];
                        dv[c] = in[(12 + slot * 6 + 3 + c) * 256];
                    }
                };
                auto load_dvc = [&](int a, float (&o)[3]) {  // derivative of the current-frame vertex for gather slot a
                    const float4 k = s_cur[a];
                    o[0] = fmaf(k.x, vc[2], k.y * vc[0]);
                    o[1] = fmaf(k.z, vc[2], k.w * vc[1]);
                    o[2] = 0.f;
                };
                float dn0[3], dv0[3], ds0[3], de0[3], d0[7], ex0[3] = {0.f, 0.f, 0.f}, dvc0[3] = {0.f, 0.f, 0.f};
                load6(0, dn0, dv0);
                if (CURR && ki >= 0) {  // block-uniform
                    load_dvc(0, dvc0);
                    rot_add(s_real_pose, dvc0, ex0);
                }
                row_first(s_pose[0], vc, sv, nv, ev, dn0, dv0, ex0, ds0, de0, d0);
                if (np == 0) {
                    accumulate_first(acc, r, d0, std::make_integer_sequence<int, 27>());
                } else {
                    const float rho = REDUCED ? rho_of(r, x) : 0.f, rho0 = REDUCED ? rho_of(d0, x) : 0.f;
#pragma unroll
                    for (int h = 0; h < HPK; ++h) {
                        if (h < np) {  // block-uniform
                            float dn1[3], dv1[3], ds1[3], de1[3], d1[7], dn2[3], dv2[3], ds2[3], de2[3], d2[7];
                            float ex1[3] = {0.f, 0.f, 0.f}, ex2[3] = {0.f, 0.f, 0.f};
                            if (CURR) {
                                float dvc1[3] = {0.f, 0.f, 0.f}, dvc2[3];
                                if (kj[h] >= 0) {  // block-uniform
                                    load_dvc(1 + 2 * h, dvc1);
                                    rot_add(s_real_pose, dvc1, ex1);
                                    rot_add(s_pose[0], dvc1, ex2);  // dR_i dvc_j
                                }
                                if (ki >= 0) rot_add(s_pose[1 + 2 * h], dvc0, ex2);  // dR_j dvc_i
                                if (ks[h] >= 0) {
                                    load_dvc(2 + 2 * h, dvc2);
                                    rot_add(s_real_pose, dvc2, ex2);
                                }
                            }
                            load6(1 + 2 * h, dn1, dv1);
                            row_first(s_pose[1 + 2 * h], vc, sv, nv, ev, dn1, dv1, ex1, ds1, de1, d1);
                            load6(2 + 2 * h, dn2, dv2);
                            row_first(s_pose[2 + 2 * h], vc, sv, nv, ev, dn2, dv2, ex2, ds2, de2, d2);
                            cross3_add(ds0, dn1, d2);  // second-order cross terms of the row
                            cross3_add(ds1, dn0, d2);
                            d2[6] += dot3(dn0, de1) + dot3(dn1, de0);
                            if (REDUCED) {
                                const float rho1 = rho_of(d1, x), rho2 = rho_of(d2, x);
#pragma unroll
                                for (int c = 0; c < 6; ++c)
                                    acc[h * 6 + c] = fmaf(d2[c], rho, fmaf(r[c], rho2, fmaf(d0[c], rho1, fmaf(d1[c], rho0, acc[h * 6 + c]))));
                            } else {
                                accumulate_pair(acc + h * 27, r, d0, d1, d2, std::make_integer_sequence<int, 27>());
                            }
                        }
                    }
                }
            }
            if ((j & 31) == 31) flush();
        }
        if (nitems & 31) flush();
        cp_async_wait<0>();
        __syncthreads();
        // ---------------- the 8 warps in order -> one NPART-value partial of this CTA for this task
        const int first_b = deriv_owner((long long) task * P.chunks, U, B);
        const int writers = deriv_owner((long long) (task + 1) * P.chunks - 1, U, B) - first_b + 1;
        double *part = P.dpartials + ((size_t) task * P.max_writers + (blockIdx.x - first_b)) * NPART;
        if (tid < NPART) {
            const int h = REDUCED ? 0 : tid / 27, e = REDUCED ? tid : tid % 27;
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) sum += s_tot[h * 256 + w * 32 + e];
            part[tid] = sum;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = atomicAdd(P.group_ticket + task, 1u) == (unsigned) writers - 1u;
        __syncthreads();
        if (!s_last) continue;
        __threadfence();
        // the last writer of the task adds the partials in CTA order and files them under the components they belong to:
        // 27 sums per component (first-order tasks, FULL pairs), or the 6 values of g_ij in the first 6 slots (REDUCED pairs)
        if (tid < NPART) {
            double sum = 0.0;
            const double *src = P.dpartials + (size_t) task * P.max_writers * NPART + tid;
            for (int w = 0; w < writers; ++w) sum += __ldcg(src + (size_t) w * NPART);
            if (np == 0) {
                if (tid < 27) P.sums[(size_t) (1 + ci) * 27 + tid] = sum;
            } else if (REDUCED) {
                const int h = tid / 6;
                if (h < np) {
                    int comp = cs[0];
#pragma unroll
                    for (int hh = 1; hh < HPK; ++hh)
                        if (h == hh) comp = cs[hh];
                    P.sums[(size_t) (1 + comp) * 27 + tid % 6] = sum;
                }
            } else {
                const int h = tid / 27;
                if (h < np) {
                    int comp = cs[0];
#pragma unroll
                    for (int hh = 1; hh < HPK; ++hh)
                        if (h == hh) comp = cs[hh];
                    P.sums[(size_t) (1 + comp) * 27 + tid % 27] = sum;
                }
            }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            P.group_ticket[task] = 0u;
            s_final = atomicAdd(P.done_ticket, 1u) == (unsigned) ntasks - 1u;
        }
        __syncthreads();
        final_cta = final_cta || s_final;
        __syncthreads();
    }
    if (!final_cta) return;  // block-uniform
    // ---------------- tail: every task's sums are complete -> the Gauss-Newton step of every component
    __threadfence();
    if (tid == 0) *P.done_ticket = 0u;
    if (!S.pose_out) return;
    icp_hessian_tail<256>(P, S, s_raw, REDUCED, tid);
}

// ---- tile form of the REDUCED Hessian derivative pass -------------------------------------------------------------------------
// The task form above reads the association record once per task and the first-order planes of parameter i once per pair run:
// ~3.4x the algorithmic bytes cross L2 -> SM at 10 parameters / 55 pairs, and that link, not HBM, bounds it (profiles/
// r02_ab_table.md).  Here a CTA owns pixels instead of tasks: a tile of 32 pixels (lane = pixel) is staged ONCE for all
// components - the record, the n first-order and the m second-order gathers at the matched pixels, cp.async, two tiles ahead -
// so every byte crosses L2 -> SM once per launch.  The NW warps of the CTA then split the components of the tile:
//   phase B: warp p < n forms the first-order row of parameter p (d_p, ds_p, de_p, rho_p = d_p6 - d_p . x: 14 floats per pixel,
//            left in shared memory) and accumulates its 27 first-order sums;
//   phase C: warp w owns a contiguous run of <= PPW pairs (the pair list is sorted by i, so the row of i stays in registers
//            along the run), reads row j, its own gathers and accumulates the 6 values of g_ij = b_ij - A_ij x.
// The rows are double buffered: between two barriers a warp runs phase B of tile j + 1 and phase C of tile j (one barrier per
// tile); the staging runs two iterations ahead of its reader (4 buffers for the record + first-order part, 3 for the
// second-order part, which only its own warp reads).
// FP32 accumulators are flushed every 32 tiles through the lane transpose into per-lane doubles (as in the task form); a CTA
// writes one partial [n x 27 + m x 6], the last CTA (ticket) adds the partials in CTA order and runs the tail.
constexpr int TILE_ROW = 14;
// staging buffers for a look-ahead of `depth` tiles: record + first-order gathers (read by phases B and C) | second-order gathers
__host__ __device__ constexpr int tile_small(int depth) { return depth + 2; }
__host__ __device__ constexpr int tile_big(int depth) { return depth + 1; }
__host__ __device__ constexpr int tile_groups(int ppw) { return (27 + 6 * ppw + 31) / 32; }
constexpr size_t tile_smem(int n, int m, int nw, int ppw, int depth, bool pipe) {
    return (size_t) (tile_small(depth) * (12 + 6 * n) + tile_big(depth) * 6 * m + (pipe ? 2 : 1) * n * TILE_ROW) * 32 * sizeof(float) + (size_t) (n + m) * 4 * sizeof(float4) +
           (size_t) nw * tile_groups(ppw) * 32 * sizeof(double) + (size_t) m * sizeof(int2);
}

template <int NW, int PPW, bool CURR, bool PIPE, int DEPTH>
__global__ void __launch_bounds__(NW * 32, 1) icp_deriv_tile_kernel(const IcpParams P, const SolveParams S) {
    constexpr int G = tile_groups(PPW);
    constexpr int NS = tile_small(DEPTH), NB = tile_big(DEPTH);
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ float4 s_real_pose[3];
    __shared__ __align__(16) float s_x[6];
    __shared__ bool s_last;
    __shared__ int s_tb;  // first tile of this CTA (read where needed: held in a register it is spilled, and the staging traffic keeps evicting L1)
    const int n = P.batch.n, m = P.batch.m;
    const int INS = 12 + 6 * n, INB = 6 * m;
    float *s_small = reinterpret_cast<float *>(s_raw);                    // [NS][INS][32]: record, first-order gathers
    float *s_big = s_small + (size_t) NS * INS * 32;                      // [NB][INB][32]: second-order gathers
    float *s_row = s_big + (size_t) NB * INB * 32;                        // [PIPE ? 2 : 1][n][TILE_ROW][32]
    float4 *s_pose = reinterpret_cast<float4 *>(s_row + (PIPE ? 2 : 1) * n * TILE_ROW * 32);  // [n + m][3]
    float4 *s_cur = s_pose + 3 * (n + m);                                 // [n + m]: d vc = (x z + y vx, z z + w vy, 0)
    double *s_tot = reinterpret_cast<double *>(s_cur + (n + m));          // [NW][G][32]
    int2 *s_pair = reinterpret_cast<int2 *>(s_tot + NW * G * 32);         // [m]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned uplane = (unsigned) (P.rows * P.cols);
    const int npix = P.rows * P.cols;
    const int ntile = (npix + 31) >> 5;
    const unsigned long long stream_policy = l2_policy_evict_first();
    // ---- per-CTA tables
    if (tid < 6) s_x[tid] = (float) __ldcg(S.real_cache + 42 + tid);
    if (tid == 6) s_tb = (int) ((long long) blockIdx.x * ntile / gridDim.x);
    if (tid >= 32 && tid < 44) reinterpret_cast<float *>(&s_real_pose[0])[tid - 32] = P.pose_curr[tid - 32];
    for (int e = tid; e < 12 * (n + m); e += NW * 32) reinterpret_cast<float *>(s_pose)[e] = P.pose_curr[12 + e];
    if (CURR) {
        const float4 *din = reinterpret_cast<const float4 *>(P.batch.dintr);  // (dfx, dfy, dcx, dcy) at level 0
        const float gx = P.batch.gx0, gy = P.batch.gy0;
        for (int c = tid; c < n + m; c += NW * 32) {
            float4 o;
            if (c < n) {
                const float4 d = __ldg(din + c);
                o = make_float4(-d.z * gx, -d.x * gx, -d.w * gy, -d.y * gy);
            } else {
                const int2 pr = __ldg(P.batch.pairs + (c - n));
                const float4 di = __ldg(din + pr.x), dj = __ldg(din + pr.y);
                o = make_float4((di.z * dj.x + dj.z * di.x) * gx * gx, 2.f * di.x * dj.x * gx * gx, (di.w * dj.y + dj.w * di.y) * gy * gy,
                                2.f * di.y * dj.y * gy * gy);
            }
            s_cur[c] = o;
        }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) s_tot[(warp * G + g) * 32 + lane] = 0.0;
    // ---- this warp's components
    const int fp = warp < n ? warp : -1;  // first-order parameter (host: n <= NW)
    // this warp's pairs (P.tile_wp: [NW][PPW] pair indices, -1 = none; dealt by the host so that the warps carry equal work)
    // (kept in shared memory: as registers they are spilled, and the local-memory reloads miss an L1 the staging keeps evicting)
    __shared__ int s_wk[NW * PPW];
    for (int e = tid; e < NW * PPW; e += NW * 32) s_wk[e] = __ldg(P.tile_wp + e);
    const int *wk = s_wk + warp * PPW;
    for (int k = tid; k < m; k += NW * 32) s_pair[k] = __ldg(P.batch.pairs + k);  // read per pair (broadcast) rather than held in registers
    __syncthreads();
    const float (&x)[6] = s_x;  // broadcast shared loads where used
    float accf[27], accp[PPW][6];
#pragma unroll
    for (int e = 0; e < 27; ++e) accf[e] = 0.f;
#pragma unroll
    for (int h = 0; h < PPW; ++h)
#pragma unroll
        for (int c = 0; c < 6; ++c) accp[h][c] = 0.f;
    // FP32 accumulators -> per-lane double totals: accumulator a is summed over the lanes by a fixed butterfly and added by lane
    // a % 32 to its slot of group a / 32 (one accumulator at a time: the lane transpose of the task form would need 32 more
    // registers here, and what it spills is reloaded from an L1 that the staging traffic keeps evicting)
    auto flush = [&]() {
        double tot[G];
#pragma unroll
        for (int g = 0; g < G; ++g) tot[g] = 0.0;
#pragma unroll
        for (int a = 0; a < 27 + 6 * PPW; ++a) {
            float v = a < 27 ? accf[a % 27] : accp[((a < 27 ? 27 : a) - 27) / 6][((a < 27 ? 27 : a) - 27) % 6];
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == (a & 31)) tot[a / 32] += (double) v;
        }
#pragma unroll
        for (int g = 0; g < G; ++g) s_tot[(warp * G + g) * 32 + lane] += tot[g];
#pragma unroll
        for (int e = 0; e < 27; ++e) accf[e] = 0.f;
#pragma unroll
        for (int h = 0; h < PPW; ++h)
#pragma unroll
            for (int c = 0; c < 6; ++c) accp[h][c] = 0.f;
    };
    const int nt = (int) ((long long) (blockIdx.x + 1) * ntile / gridDim.x) - (int) ((long long) blockIdx.x * ntile / gridDim.x);
    auto idx_at = [&](int j) {
        const int p = (s_tb + j) * 32 + lane;
        return (j < nt && p < npix) ? P.rec_idx[p] : -1;
    };
    // staging of tile j.  Small part: the record (3 x 16 bytes per pixel) by the last three warps, the six gather rows (normal and
    // vertex planes at the matched pixel) of parameter p by warp p.  Big part: every warp stages the gathers of its own pairs.
    auto gather6 = [&](float *d, unsigned o) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            cp_async4_stream(d + c * 32, P.nmap_prev + (o + c * uplane), stream_policy);
            cp_async4_stream(d + (3 + c) * 32, P.vmap_prev + (o + c * uplane), stream_policy);
        }
    };
    auto issue_small = [&](int buf, int j, int q) {
        if (q < 0) return;
        float *dst = s_small + (size_t) buf * INS * 32;
        if (warp >= NW - 3)
            cp_async16(reinterpret_cast<float4 *>(dst) + (warp - (NW - 3)) * 32 + lane, P.rec_f + (size_t) ((s_tb + j) * 32 + lane) * 4 + (warp - (NW - 3)));
        if (fp >= 0) gather6(dst + (12 + fp * 6) * 32 + lane, (unsigned) q + (unsigned) (1 + fp) * 3u * uplane);
    };
    auto issue_big = [&](int buf, int q) {
        if (q < 0) return;
        float *d = s_big + (size_t) buf * INB * 32 + lane;
        const unsigned o = (unsigned) q + (unsigned) (1 + n) * 3u * uplane;
#pragma unroll
        for (int h = 0; h < PPW; ++h)
            if (wk[h] >= 0) gather6(d + wk[h] * (6 * 32), o + (unsigned) wk[h] * 3u * uplane);
    };
    // per-pixel real quantities of a staged tile
    struct Px {
        float vc[3], sv[3], nv[3], ev[3], r[7], rho;
    };
    auto pixel = [&](int buf, Px &p) {
        const float4 *in4 = reinterpret_cast<const float4 *>(s_small + (size_t) buf * INS * 32) + lane;
        const float4 f0 = in4[0], f1 = in4[32], f2 = in4[64];
        p.vc[0] = f0.x, p.vc[1] = f0.y, p.vc[2] = f0.z;
        p.sv[0] = f0.w, p.sv[1] = f1.x, p.sv[2] = f1.y;
        p.nv[0] = f1.z, p.nv[1] = f1.w, p.nv[2] = f2.x;
        p.ev[0] = f2.y, p.ev[1] = f2.z, p.ev[2] = f2.w;
        cross3(p.sv, p.nv, p.r);
        p.r[3] = p.nv[0], p.r[4] = p.nv[1], p.r[5] = p.nv[2];
        p.r[6] = dot3(p.nv, p.ev);
        p.rho = rho_of(p.r, x);
    };
    auto dvc_of = [&](int c, const Px &p, float (&o)[3]) {
        const float4 k = s_cur[c];
        o[0] = fmaf(k.x, p.vc[2], k.y * p.vc[0]);
        o[1] = fmaf(k.z, p.vc[2], k.w * p.vc[1]);
        o[2] = 0.f;
    };
    // phase B of a staged tile: the first-order row of this warp's parameter -> rows[rb], and its 27 sums
    auto phase_b = [&](int buf, int rb, int q) {
        if (fp < 0 || q < 0) return;  // fp: warp-uniform
        Px p;
        pixel(buf, p);
        const float *in = s_small + (size_t) buf * INS * 32 + lane;
        float dn[3], dv[3], ds[3], de[3], d[7], ex[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; ++c) dn[c] = in[(12 + fp * 6 + c) * 32], dv[c] = in[(12 + fp * 6 + 3 + c) * 32];
        if (CURR) {
            float dvc[3];
            dvc_of(fp, p, dvc);
            rot_add(s_real_pose, dvc, ex);
        }
        row_first(s_pose + 3 * fp, p.vc, p.sv, p.nv, p.ev, dn, dv, ex, ds, de, d);
        accumulate_first(accf, p.r, d, std::make_integer_sequence<int, 27>());
        float *o = s_row + ((size_t) rb * n + fp) * TILE_ROW * 32 + lane;
        o[0] = d[0], o[32] = d[1], o[64] = d[2];
        o[96] = dn[0], o[128] = dn[1], o[160] = dn[2];
        o[192] = d[6];
        o[224] = ds[0], o[256] = ds[1], o[288] = ds[2];
        o[320] = de[0], o[352] = de[1], o[384] = de[2];
        o[416] = rho_of(d, x);
    };
    // CURR: which of this warp's pairs involve a parameter that moves the intrinsics (bit h: parameter i, bit 8 + h: parameter j)
    unsigned curmask = 0u;
    if (CURR) {
#pragma unroll
        for (int h = 0; h < PPW; ++h) {
            if (wk[h] >= 0) {
                const float4 ki = s_cur[s_pair[wk[h]].x], kj = s_cur[s_pair[wk[h]].y];
                if ((ki.x != 0.f) | (ki.y != 0.f) | (ki.z != 0.f) | (ki.w != 0.f)) curmask |= 1u << h;
                if ((kj.x != 0.f) | (kj.y != 0.f) | (kj.z != 0.f) | (kj.w != 0.f)) curmask |= 1u << (8 + h);
            }
        }
    }
    // Pipeline.  Between two barriers a warp runs phase B of tile j + 1 and phase C of tile j; the small part of tile j + 3 and
    // the big part of tile j + 2 are issued at the top of iteration j (one commit group), two iterations before they are read.
    constexpr int R = DEPTH + 3;
    int qs[R];  // matched indices of tiles j .. j + DEPTH + 2
#pragma unroll
    for (int i = 0; i < R; ++i) qs[i] = idx_at(i);
    issue_small(0, 0, qs[0]);
#pragma unroll
    for (int g = 0; g < DEPTH; ++g) {  // commit group g: small part of tile g + 1, big part of tile g
        issue_small(g + 1, g + 1, qs[g + 1]);
        issue_big(g, qs[g]);
        cp_async_commit();
    }
    if (PIPE) {
        cp_async_wait<DEPTH - 1>();
        __syncthreads();
        phase_b(0, 0, qs[0]);
    }
    int sb = 0, bb = 0;  // buffers of tile j: small part (of NS), big part (of NB)
    for (int j = 0; j < nt; ++j) {
        cp_async_wait<DEPTH - 1>();  // this thread's copies up to (small j + 1, big j) have landed
        __syncthreads();             // ... everyone's have; rows of tile j are complete; phase C of tile j - 1 is over
        issue_small((sb + DEPTH + 1) % NS, j + DEPTH + 1, qs[DEPTH + 1]);
        issue_big((bb + DEPTH) % NB, qs[DEPTH]);
        cp_async_commit();
        if (PIPE) {
            phase_b((sb + 1) % NS, (j + 1) & 1, qs[1]);
        } else {  // A/B form: rows of tile j, a second barrier, then its pairs
            phase_b(sb, 0, qs[0]);
            __syncthreads();
        }
        const int q = qs[0];
#pragma unroll
        for (int i = 0; i < R - 1; ++i) qs[i] = qs[i + 1];
        qs[R - 1] = idx_at(j + R);
        // ---- phase C: this warp's pairs on tile j
        if (q >= 0) {
            Px p;
            pixel(sb, p);
            const float *in = s_big + (size_t) bb * INB * 32 + lane;
            const float *rows = s_row + (size_t) (PIPE ? (j & 1) : 0) * n * TILE_ROW * 32 + lane;
            int have_i = -1;
            float di[7], dsi[3], dei[3], rhoi = 0.f;
            auto load_row = [&](int p_, float (&d)[7], float (&ds)[3], float (&de)[3], float &rh) {
                const float *o = rows + (size_t) p_ * TILE_ROW * 32;
#pragma unroll
                for (int e = 0; e < 7; ++e) d[e] = o[e * 32];
#pragma unroll
                for (int e = 0; e < 3; ++e) ds[e] = o[(7 + e) * 32], de[e] = o[(10 + e) * 32];
                rh = o[13 * 32];
            };
#pragma unroll
            for (int h = 0; h < PPW; ++h) {
                if (wk[h] >= 0) {  // warp-uniform
                    const int k = wk[h];
                    const int2 pr = s_pair[k];
                    const int i = pr.x, jj = pr.y;
                    if (i != have_i) {
                        load_row(i, di, dsi, dei, rhoi);
                        have_i = i;
                    }
                    float dj[7], dsj[3], dej[3], rhoj;
                    load_row(jj, dj, dsj, dej, rhoj);
                    float dn2[3], dv2[3], ds2[3], de2[3], d2[7], ex2[3] = {0.f, 0.f, 0.f};
#pragma unroll
                    for (int c = 0; c < 3; ++c) dn2[c] = in[(k * 6 + c) * 32], dv2[c] = in[(k * 6 + 3 + c) * 32];
                    if (CURR) {
                        const bool ci = (curmask >> h) & 1u, cj = (curmask >> (8 + h)) & 1u;
                        if (ci | cj) {  // warp-uniform
                            float a[3];
                            if (cj) {
                                dvc_of(jj, p, a);
                                rot_add(s_pose + 3 * i, a, ex2);  // dR_i dvc_j
                            }
                            if (ci) {
                                dvc_of(i, p, a);
                                rot_add(s_pose + 3 * jj, a, ex2);  // dR_j dvc_i
                            }
                            if (ci & cj) {
                                dvc_of(n + k, p, a);
                                rot_add(s_real_pose, a, ex2);  // R dvc_ij
                            }
                        }
                    }
                    row_first(s_pose + 3 * (n + k), p.vc, p.sv, p.nv, p.ev, dn2, dv2, ex2, ds2, de2, d2);
                    const float *dni = di + 3, *dnj = dj + 3;
                    cross3_add(dsi, dnj, d2);  // second-order cross terms of the row
                    cross3_add(dsj, dni, d2);
                    d2[6] += dot3(dni, dej) + dot3(dnj, dei);
                    const float rho2 = rho_of(d2, x);
#pragma unroll
                    for (int c = 0; c < 6; ++c)
                        accp[h][c] = fmaf(d2[c], p.rho, fmaf(p.r[c], rho2, fmaf(di[c], rhoj, fmaf(dj[c], rhoi, accp[h][c]))));
                }
            }
        }
        sb = sb + 1 == NS ? 0 : sb + 1;
        bb = bb + 1 == NB ? 0 : bb + 1;
        if ((j & 31) == 31) flush();
    }
    if (nt & 31) flush();
    cp_async_wait<0>();
    // ---- one partial per CTA: [n][27] first-order sums, [m][6] reduced pair values
    const int NV = n * 27 + m * 6;
    double *part = P.dpartials + (size_t) blockIdx.x * NV;
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const int a = g * 32 + lane;
        const double v = s_tot[(warp * G + g) * 32 + lane];
        if (a < 27) {
            if (fp >= 0) part[fp * 27 + a] = v;
        } else if (a - 27 < 6 * PPW) {
            int k = -1;
#pragma unroll
            for (int h = 0; h < PPW; ++h)
                if ((a - 27) / 6 == h) k = wk[h];
            if (k >= 0) part[n * 27 + k * 6 + (a - 27) % 6] = v;
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(P.done_ticket, 1u) == gridDim.x - 1u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int v = tid; v < NV; v += NW * 32) {
        double sum = 0.0;
        const double *src = P.dpartials + v;
#pragma unroll 8
        for (unsigned b = 0; b < gridDim.x; ++b) sum += __ldcg(src + (size_t) b * NV);
        if (v < n * 27) {
            P.sums[27 + v] = sum;  // component 1 + p, element e: 27 (1 + p) + e
        } else {
            const int k = (v - n * 27) / 6, c = (v - n * 27) - k * 6;
            P.sums[(size_t) (1 + n + k) * 27 + c] = sum;
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) *P.done_ticket = 0u;
    if (!S.pose_out) return;
    icp_hessian_tail<NW * 32>(P, S, s_raw, true, tid);  // the Gauss-Newton step of every component
}

// ---- Gauss-Newton step of a Hessian batch (kind 2), one thread per component.
// Real prelude shared by both task kinds: status flags, real A / b, Cholesky factor and real solution (loaded from the real
// step's cache under split chains, computed otherwise).  Returns false when the iteration is skipped (degenerate system).
XS_DEV double ld_sum(const SolveParams &P, const double *p) { return P.staged ? *p : __ldcg(p); }
XS_DEV bool gn_real_prelude(const SolveParams &P, const double *real_sums, bool owns_real, double (&A)[6][6], double (&b)[6], Chol6 &F,
                            double (&xr)[6]) {
    const int st0 = P.status_in ? P.status_in[0] : __ldcg(P.status), st1 = P.status_in ? P.status_in[1] : __ldcg(P.status + 1);
    if (st0 != 0 || st1 != 0) {
        if (owns_real) P.status[0] = st0 != 0 ? st0 : st1;
        return false;
    }
    unpack_sums(real_sums, A, b);
    if (P.deriv_only && P.real_cache) {
        const double *rc = P.real_cache;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
            for (int j = 0; j < 6; ++j) F.L[i][j] = ld_sum(P, rc + i * 6 + j);
            F.inv[i] = ld_sum(P, rc + 36 + i);
            xr[i] = ld_sum(P, rc + 42 + i);
        }
        return true;
    }
    const double det = det6_dev(A);
    if (fabs(det) < 1e-15 || isnan(det)) {
        if (owns_real) P.status[1] = isnan(det) ? 2 : 1;
        return false;
    }
    chol6_factor(A, F);
    chol6_solve(F, b, xr);
    return true;
}

// pose update of KinectFusionReconstruction.cpp:212-224 for the C components comp[0..C-1] of the pose tables; components
// c >= first_store are written, the real part only by its owner
template <int C>
XS_DEV void gn_pose_update(const SolveParams &P, const int (&comp)[C], const Jet<C, 1> (&x)[6], int first_store, bool owns_real) {
    typedef Jet<C, 1> J;
    JMat3<C> Rcurr;
    J tcurr[3];
    const float *pr = P.pose_in;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            Rcurr.m[i][j].v = pr[i * 3 + j];
            for (int a = 0; a < C; ++a) Rcurr.m[i][j].d[a] = pr[(size_t) (1 + comp[a]) * 12 + i * 3 + j];
        }
        tcurr[i].v = pr[9 + i];
        for (int a = 0; a < C; ++a) tcurr[i].d[a] = pr[(size_t) (1 + comp[a]) * 12 + 9 + i];
    }
    const bool co = P.deriv_only != 0;
    const JMat3<C> Rinc = jmatmul(jmatmul(jaxis_rotation<C>(x[2], 2, co), jaxis_rotation<C>(x[1], 1, co)), jaxis_rotation<C>(x[0], 0, co));
    J tn[3];
    for (int i = 0; i < 3; ++i) {
        J sacc = Rinc.m[i][0] * tcurr[0];
        for (int k = 1; k < 3; ++k) sacc = sacc + Rinc.m[i][k] * tcurr[k];
        tn[i] = sacc + x[3 + i];
    }
    const JMat3<C> Rn = jmatmul(Rinc, Rcurr);
    float *pw = P.pose_out;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            if (owns_real) pw[i * 3 + j] = Rn.m[i][j].v;
            for (int a = first_store; a < C; ++a) pw[(size_t) (1 + comp[a]) * 12 + i * 3 + j] = Rn.m[i][j].d[a];
        }
        if (owns_real) pw[9 + i] = tn[i].v;
        for (int a = first_store; a < C; ++a) pw[(size_t) (1 + comp[a]) * 12 + 9 + i] = tn[i].d[a];
    }
}

// first-order solution of parameter p: x_p = A^-1 (b_p - A_p x) (analytic) or the imaginary part of the Hermitian-LLT solve
// (LLT mode, what a one-direction complex run of the reference yields); A_p is returned for the pairs
XS_DEV void first_order_solution(const SolveParams &P, int p, const double (&A)[6][6], const double (&b)[6], const Chol6 &F,
                                 const double (&xr)[6], double (&Ap)[6][6], double (&xp)[6]) {
    double bp[6], sums[27];
    for (int e = 0; e < 27; ++e) sums[e] = ld_sum(P, P.sums + (size_t) (1 + p) * 27 + e);
    unpack_sums(sums, Ap, bp);
    if (P.solve_mode == XS_SOLVE_EIGEN_LLT) {
        cplx xq[6];
        llt_hermitian_solve6_dev(A, Ap, b, bp, xq);
        for (int e = 0; e < 6; ++e) xp[e] = xq[e].im;
    } else {
        double t[6], rhs[6];
        matvec6_dev(Ap, xr, t);
        for (int e = 0; e < 6; ++e) rhs[e] = bp[e] - t[e];
        chol6_solve(F, rhs, xp);
    }
}

// first-order component i of a Hessian batch: solution + pose update
__device__ __noinline__ void icp_solve_hessian_first(const SolveParams &P, int i, const double *real_sums) {
    const bool owns_real = i == 0 && !P.deriv_only;
    double A[6][6], b[6], xr[6], Ai[6][6], xi[6];
    Chol6 F;
    if (!gn_real_prelude(P, real_sums, owns_real, A, b, F, xr)) return;
    first_order_solution(P, i, A, b, F, xr, Ai, xi);
    Jet<1, 1> x[6];
#pragma unroll
    for (int e = 0; e < 6; ++e) {
        x[e].v = (float) xr[e];
        x[e].d[0] = (float) xi[e];
    }
    const int comp[1] = {i};
    gn_pose_update<1>(P, comp, x, 0, owns_real);
}

// pair k = (i, j): x_ij = A^-1 (b_ij - A_ij x - A_i x_j - A_j x_i); the pose update runs in the bicomplex algebra on
// (F_i, F_j, S_ij) and stores S_ij only.  The thread forms x_i and x_j itself (the same arithmetic as the first-order threads,
// so the same values) rather than waiting for them: one phase, no barrier.
// (reduced: the derivative pass has already formed g_ij = b_ij - A_ij x, the first 6 values filed under the component)
__device__ __noinline__ void icp_solve_hessian_pair(const SolveParams &P, int n, int k, int2 pr, const double *real_sums, bool reduced) {
    double A[6][6], b[6], xr[6];
    Chol6 F;
    if (!gn_real_prelude(P, real_sums, false, A, b, F, xr)) return;
    double M[6][6], v[6], sums[27], rhs[6], t[6], xs[6];
    if (reduced) {
        for (int e = 0; e < 6; ++e) rhs[e] = ld_sum(P, P.sums + (size_t) (1 + n + k) * 27 + e);
    } else {
        for (int e = 0; e < 27; ++e) sums[e] = ld_sum(P, P.sums + (size_t) (1 + n + k) * 27 + e);
        unpack_sums(sums, M, v);  // A_ij, b_ij
        matvec6_dev(M, xr, t);
        for (int e = 0; e < 6; ++e) rhs[e] = v[e] - t[e];
    }
    double xi[6], xj[6];
    first_order_solution(P, pr.x, A, b, F, xr, M, xi);  // M = A_i
    if (pr.y == pr.x) {
        for (int e = 0; e < 6; ++e) xj[e] = xi[e];
        matvec6_dev(M, xi, t);
        for (int e = 0; e < 6; ++e) rhs[e] -= 2.0 * t[e];
    } else {
        double Mj[6][6];
        first_order_solution(P, pr.y, A, b, F, xr, Mj, xj);
        matvec6_dev(M, xj, t);  // A_i x_j
        for (int e = 0; e < 6; ++e) rhs[e] -= t[e];
        matvec6_dev(Mj, xi, t);  // A_j x_i
        for (int e = 0; e < 6; ++e) rhs[e] -= t[e];
    }
    chol6_solve(F, rhs, xs);
    Jet<3, 1> x[6];
#pragma unroll
    for (int e = 0; e < 6; ++e) {
        x[e].v = (float) xr[e];
        x[e].d[0] = (float) xi[e];
        x[e].d[1] = (float) xj[e];
        x[e].d[2] = (float) xs[e];
    }
    const int comp[3] = {pr.x, pr.y, n + k};
    gn_pose_update<3>(P, comp, x, 2, false);
}

// ---- register-resident form of the two solves for the staged, analytic, split-chain case (the frame loop's): the generic
// functions above unpack every 6x6 system into local arrays (five of them per pair: 1.4 KB of stack) and the single thread of a
// component then waits on local-memory round trips.  Here the packed sums and the real step's factor are read from their
// shared-memory copies at the point of use with compile-time indices; the arithmetic and its order are those of
// unpack_sums / matvec6_dev / chol6_solve.
XS_DEV constexpr int sum_idx(int i, int j) { return i <= j ? i * 7 - i * (i - 1) / 2 + (j - i) : j * 7 - j * (j - 1) / 2 + (i - j); }
XS_DEV void packed_matvec(const double *s27, const double (&x)[6], double (&y)[6]) {  // y = A x, A from the 27 packed sums
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double sacc = 0;
#pragma unroll
        for (int j = 0; j < 6; ++j) sacc += s27[sum_idx(i, j)] * x[j];
        y[i] = sacc;
    }
}
XS_DEV void staged_chol_solve(const double *rc, const double (&b)[6], double (&x)[6]) {  // rc: L[6][6], inv[6] (REAL_CACHE layout)
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double sacc = b[i];
#pragma unroll
        for (int j = 0; j < i; ++j) sacc -= rc[i * 6 + j] * y[j];
        y[i] = sacc * rc[36 + i];
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        double sacc = y[i];
#pragma unroll
        for (int j = i + 1; j < 6; ++j) sacc -= rc[j * 6 + i] * x[j];
        x[i] = sacc * rc[36 + i];
    }
}
XS_DEV void staged_first_order(const SolveParams &P, int p, const double (&xr)[6], double (&xp)[6]) {  // x_p = A^-1 (b_p - A_p x)
    const double *sp = P.sums + (size_t) (1 + p) * 27;
    double t[6], rhs[6];
    packed_matvec(sp, xr, t);
#pragma unroll
    for (int e = 0; e < 6; ++e) rhs[e] = sp[sum_idx(e, 6)] - t[e];
    staged_chol_solve(P.real_cache, rhs, xp);
}
XS_DEV bool staged_fast_path(const SolveParams &P) { return P.staged && P.deriv_only && P.real_cache && P.solve_mode != XS_SOLVE_EIGEN_LLT; }
__device__ __noinline__ void staged_solve_first(const SolveParams &P, int i) {
    if (P.status_in[0] != 0 || P.status_in[1] != 0) return;
    double xr[6], xi[6];
#pragma unroll
    for (int e = 0; e < 6; ++e) xr[e] = P.real_cache[42 + e];
    staged_first_order(P, i, xr, xi);
    Jet<1, 1> x[6];
#pragma unroll
    for (int e = 0; e < 6; ++e) {
        x[e].v = (float) xr[e];
        x[e].d[0] = (float) xi[e];
    }
    const int comp[1] = {i};
    gn_pose_update<1>(P, comp, x, 0, false);
}
__device__ __noinline__ void staged_solve_pair(const SolveParams &P, int n, int k, int2 pr, bool reduced) {
    if (P.status_in[0] != 0 || P.status_in[1] != 0) return;
    double xr[6], xi[6], xj[6], rhs[6], t[6], xs[6];
#pragma unroll
    for (int e = 0; e < 6; ++e) xr[e] = P.real_cache[42 + e];
    const double *ss = P.sums + (size_t) (1 + n + k) * 27, *si = P.sums + (size_t) (1 + pr.x) * 27, *sj = P.sums + (size_t) (1 + pr.y) * 27;
    if (reduced) {
#pragma unroll
        for (int e = 0; e < 6; ++e) rhs[e] = ss[e];
    } else {
        packed_matvec(ss, xr, t);
#pragma unroll
        for (int e = 0; e < 6; ++e) rhs[e] = ss[sum_idx(e, 6)] - t[e];
    }
    staged_first_order(P, pr.x, xr, xi);
    if (pr.y == pr.x) {
        packed_matvec(si, xi, t);
#pragma unroll
        for (int e = 0; e < 6; ++e) xj[e] = xi[e], rhs[e] -= 2.0 * t[e];
    } else {
        staged_first_order(P, pr.y, xr, xj);
        packed_matvec(si, xj, t);  // A_i x_j
#pragma unroll
        for (int e = 0; e < 6; ++e) rhs[e] -= t[e];
        packed_matvec(sj, xi, t);  // A_j x_i
#pragma unroll
        for (int e = 0; e < 6; ++e) rhs[e] -= t[e];
    }
    staged_chol_solve(P.real_cache, rhs, xs);
    Jet<3, 1> x[6];
#pragma unroll
    for (int e = 0; e < 6; ++e) {
        x[e].v = (float) xr[e];
        x[e].d[0] = (float) xi[e];
        x[e].d[1] = (float) xj[e];
        x[e].d[2] = (float) xs[e];
    }
    const int comp[3] = {pr.x, pr.y, n + k};
    gn_pose_update<3>(P, comp, x, 2, false);
}

// Tail of a Hessian derivative pass (run by the CTA that completed the last sums): the Gauss-Newton step of every component.
// Everything the solves read - the sums of all components, the real step's factor, the pose tables, the status flags - is
// staged into shared memory by the whole CTA first (one round of global latency instead of one per dependent load of a
// single thread), then first-order components and pairs run side by side on different warps.
template <int NT>
XS_DEV void icp_hessian_tail(const IcpParams &P, const SolveParams &S, unsigned char *s_raw, bool reduced, int tid) {
    const int n = P.batch.n, m = P.batch.m, nv = 27 * (1 + n + m), npose = 12 * (1 + n + m);
    double *s_sums = reinterpret_cast<double *>(s_raw);  // the staging area of the derivative pass is idle here
    double *s_rc = s_sums + nv;
    float *s_pose = reinterpret_cast<float *>(s_rc + REAL_CACHE);
    int *s_st = reinterpret_cast<int *>(s_pose + npose);
    const bool cached_real = S.deriv_only && S.real_cache;
    const int first_threads = ((n + 31) & ~31) % NT;  // pairs start on the warp after the first-order components
    unsigned smem_avail;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(smem_avail));
    if ((size_t) nv * sizeof(double) + REAL_CACHE * sizeof(double) + (size_t) npose * sizeof(float) + 2 * sizeof(int) > smem_avail) {
        // a batch too large for the staging area (several hundred components): every thread reads what it needs itself
        if (S.log)
            for (int i = tid; i < nv; i += NT) S.log[i] = __ldcg(P.sums + i);
        double *s_real = reinterpret_cast<double *>(s_raw);
        if (tid < 27) s_real[tid] = __ldcg(P.sums + tid);
        __syncthreads();
        for (int i = tid; i < n; i += NT) icp_solve_hessian_first(S, i, s_real);
        for (int k = (tid + NT - first_threads) % NT; k < m; k += NT) icp_solve_hessian_pair(S, n, k, __ldg(P.batch.pairs + k), s_real, reduced);
        return;
    }
    for (int i = tid; i < nv; i += NT) s_sums[i] = __ldcg(P.sums + i);
    if (cached_real)
        for (int i = tid; i < REAL_CACHE; i += NT) s_rc[i] = __ldcg(S.real_cache + i);
    for (int i = tid; i < npose; i += NT) s_pose[i] = __ldcg(S.pose_in + i);
    if (tid < 2) s_st[tid] = __ldcg(S.status + tid);
    __syncthreads();
    if (S.log)
        for (int i = tid; i < nv; i += NT) S.log[i] = s_sums[i];
    SolveParams L = S;
    L.sums = s_sums;
    L.pose_in = s_pose;
    L.status_in = s_st;
    L.staged = 1;
    if (cached_real) L.real_cache = s_rc;
    if (staged_fast_path(L)) {
        for (int i = tid; i < n; i += NT) staged_solve_first(L, i);
        for (int k = (tid + NT - first_threads) % NT; k < m; k += NT) staged_solve_pair(L, n, k, __ldg(P.batch.pairs + k), reduced);
        return;
    }
    for (int i = tid; i < n; i += NT) icp_solve_hessian_first(L, i, s_sums);
    for (int k = (tid + NT - first_threads) % NT; k < m; k += NT) icp_solve_hessian_pair(L, n, k, __ldg(P.batch.pairs + k), s_sums, reduced);
}

// persistent scratch of the ICP operator (gbuf / mbuf of the reference, ICP.cu:400-403)
struct IcpScratch {
    double *d_partials = nullptr, *d_sums = nullptr, *h_sums = nullptr, *d_dpartials = nullptr;
    unsigned int *d_ticket = nullptr, *d_group_ticket = nullptr, *d_done_ticket = nullptr;
    int *d_tile_wp = nullptr;          // Hessian batch, tile form: pair slots of the warps (built with the task table)
    int tile_wp_nw = 0;
    bool tile_wp_cur = false;
    const void *tile_wp_key = nullptr;
    struct HTask *d_htasks = nullptr;  // Hessian batch: task table of the derivative pass, built once per pair list
    const void *htasks_key = nullptr;
    int n_htasks = 0, htasks_hp = 0;
    int cap_groups = 0;
    float *d_pose = nullptr, *h_pose = nullptr;  // seam-level entry point only: [(1+ncomp)][12]
    int *d_rec_idx = nullptr;
    float4 *d_rec_f = nullptr;
    int cap_vals = 0, cap_comp = -1, cap_pix = 0, cap_slots = 0;
    // split chains (icp_iteration_async with a second stream): per-iteration slots of the sums and of the association record,
    // and the event that marks the real chain's iteration as complete
    static constexpr int MAX_SLOTS = 16;
    cudaEvent_t ev_real[MAX_SLOTS] = {};
    double *d_real_cache = nullptr;  // [MAX_SLOTS][REAL_CACHE]
    size_t cap_dpart = 0;
    int max_blocks = 296;  // set from the device's SM count at creation (two CTAs per SM)
    int device = -1;       // the device every buffer and event of this scratch lives on
    // CUDA-event brackets of the derivative kernel launches since the last reset (roofline timing, bench.py)
    static constexpr int MAX_TIMED = 16;
    cudaEvent_t ev0[MAX_TIMED] = {}, ev1[MAX_TIMED] = {};
    int timed_npix[MAX_TIMED] = {};
    int n_timed = 0;
};
// One scratch per pipeline object (xs_kinfu owns one; two pipelines in one process - e.g. one per device - never share
// tickets, slots or events).  The seam-level entry points (xs_estimate_combined, xs_compute_optimize_matrix), which have
// no handle, share one lazily created scratch per device and are therefore not re-entrant, like the reference's
// estimateCombined with its caller-owned gbuf / mbuf (ICP.cu:400-403).
IcpScratch *icp_scratch_create() {
    IcpScratch *sc = new IcpScratch();
    cudaGetDevice(&sc->device);
    sc->max_blocks = 2 * sm_count();
    return sc;
}
void icp_scratch_destroy(IcpScratch *sc) {
    if (!sc) return;
    cudaFree(sc->d_partials);
    cudaFree(sc->d_sums);
    cudaFreeHost(sc->h_sums);
    cudaFree(sc->d_dpartials);
    cudaFree(sc->d_ticket);
    cudaFree(sc->d_group_ticket);
    cudaFree(sc->d_done_ticket);
    cudaFree(sc->d_htasks);
    cudaFree(sc->d_tile_wp);
    cudaFree(sc->d_pose);
    cudaFreeHost(sc->h_pose);
    cudaFree(sc->d_rec_idx);
    cudaFree(sc->d_rec_f);
    cudaFree(sc->d_real_cache);
    for (int i = 0; i < IcpScratch::MAX_SLOTS; ++i)
        if (sc->ev_real[i]) cudaEventDestroy(sc->ev_real[i]);
    for (int i = 0; i < IcpScratch::MAX_TIMED; ++i) {
        if (sc->ev0[i]) cudaEventDestroy(sc->ev0[i]);
        if (sc->ev1[i]) cudaEventDestroy(sc->ev1[i]);
    }
    delete sc;
}
static IcpScratch *g_last_timed = nullptr;  // scratch whose derivative launches xs_icp_deriv_times reports
static IcpScratch *seam_scratch() {
    static IcpScratch *per_device[64] = {nullptr};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!per_device[dev]) per_device[dev] = icp_scratch_create();
    return per_device[dev];
}

static int icp_reserve(IcpScratch *scp, int ncomp, int npix, size_t dpart, int groups, int slots = 1) {
    IcpScratch &g_icp = *scp;
    const int nvals = 27 * (1 + ncomp);
    if (slots > g_icp.cap_slots) {  // the slot count multiplies the sums and the record: start over with the larger one
        g_icp.cap_vals = 0;
        g_icp.cap_pix = 0;
        g_icp.cap_slots = slots;
    }
    slots = g_icp.cap_slots;
    if (nvals > g_icp.cap_vals) {
        cudaFree(g_icp.d_sums);
        cudaFreeHost(g_icp.h_sums);
        g_icp.d_sums = nullptr;
        g_icp.h_sums = nullptr;
        XS_CUDA(cudaMalloc(&g_icp.d_sums, (size_t) nvals * slots * sizeof(double)));
        XS_CUDA(cudaMallocHost(&g_icp.h_sums, (size_t) nvals * sizeof(double)));
        g_icp.cap_vals = nvals;
    }
    if (!g_icp.d_partials) XS_CUDA(cudaMalloc(&g_icp.d_partials, (size_t) g_icp.max_blocks * 27 * sizeof(double)));
    if (!g_icp.d_ticket) {
        XS_CUDA(cudaMalloc(&g_icp.d_ticket, sizeof(unsigned int)));
        XS_CUDA(cudaMemset(g_icp.d_ticket, 0, sizeof(unsigned int)));
        XS_CUDA(cudaMalloc(&g_icp.d_done_ticket, sizeof(unsigned int)));
        XS_CUDA(cudaMemset(g_icp.d_done_ticket, 0, sizeof(unsigned int)));
    }
    if (ncomp > g_icp.cap_comp) {
        cudaFree(g_icp.d_pose);
        cudaFreeHost(g_icp.h_pose);
        const size_t n = (size_t) (1 + ncomp) * 12;
        XS_CUDA(cudaMalloc(&g_icp.d_pose, n * sizeof(float)));
        XS_CUDA(cudaMallocHost(&g_icp.h_pose, n * sizeof(float)));
        g_icp.cap_comp = ncomp;
    }
    if (ncomp > 0 && npix > g_icp.cap_pix) {
        cudaFree(g_icp.d_rec_idx);
        cudaFree(g_icp.d_rec_f);
        g_icp.d_rec_idx = nullptr;
        g_icp.d_rec_f = nullptr;
        XS_CUDA(cudaMalloc(&g_icp.d_rec_idx, (size_t) npix * slots * sizeof(int)));
        XS_CUDA(cudaMalloc(&g_icp.d_rec_f, (size_t) npix * slots * 4 * sizeof(float4)));
        g_icp.cap_pix = npix;
    }
    if (groups > g_icp.cap_groups) {
        cudaFree(g_icp.d_group_ticket);
        XS_CUDA(cudaMalloc(&g_icp.d_group_ticket, (size_t) groups * sizeof(unsigned int)));
        XS_CUDA(cudaMemset(g_icp.d_group_ticket, 0, (size_t) groups * sizeof(unsigned int)));
        g_icp.cap_groups = groups;
    }
    if (dpart > g_icp.cap_dpart) {
        cudaFree(g_icp.d_dpartials);
        XS_CUDA(cudaMalloc(&g_icp.d_dpartials, dpart * sizeof(double)));
        g_icp.cap_dpart = dpart;
    }
    return XS_OK;
}

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

template <int C, int ST> static int launch_deriv(const IcpParams &P, const SolveParams &S, int grid, cudaStream_t s) {
    static bool smem_set = false;
    if (!smem_set) {
        XS_CUDA(cudaFuncSetAttribute(icp_deriv_kernel<C, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) deriv_smem(ST)));
        smem_set = true;
    }
    icp_deriv_kernel<C, ST><<<grid, 256, deriv_smem(ST), s>>>(P, S);
    XS_LAUNCH_CHECK();
    return XS_OK;
}

constexpr int TILE_NW = 11, TILE_PPW = 5;  // 10 parameters + 55 pairs: one row and five pairs per warp
constexpr int TILE_NW_DEFAULT = 12;  // 0.302 vs 0.316 ms per level-0 launch at 10 parameters / 55 pairs (profiles/r02_ab_table.md)
template <int NW, int PPW, bool CURR, bool PIPE, int DEPTH> static int launch_deriv_tile(const IcpParams &P, const SolveParams &S, int grid, cudaStream_t s) {
    static size_t smem_set = 0;
    static const int pad_kb = env_int("XS_ICP_TILE_PAD_KB", 0);  // experiment: unused shared memory (shrinks L1)
    const size_t smem = std::min<size_t>(tile_smem(P.batch.n, P.batch.m, NW, PPW, DEPTH, PIPE) + (size_t) pad_kb * 1024, 227 * 1024);
    if (smem > smem_set) {
        XS_CUDA(cudaFuncSetAttribute(icp_deriv_tile_kernel<NW, PPW, CURR, PIPE, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        smem_set = smem;
    }
    icp_deriv_tile_kernel<NW, PPW, CURR, PIPE, DEPTH><<<grid, NW * 32, smem, s>>>(P, S);
    XS_LAUNCH_CHECK();
    return XS_OK;
}

template <int ST, int MINB, int HPK, bool REDUCED, bool CURR> static int launch_deriv_h(const IcpParams &P, const SolveParams &S, int grid, cudaStream_t s) {
    static bool smem_set = false;
    if (!smem_set) {
        XS_CUDA(cudaFuncSetAttribute(icp_deriv_h_kernel<ST, MINB, HPK, REDUCED, CURR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int) deriv_h_smem(ST, HPK, REDUCED)));
        smem_set = true;
    }
    icp_deriv_h_kernel<ST, MINB, HPK, REDUCED, CURR><<<grid, 256, deriv_h_smem(ST, HPK, REDUCED), s>>>(P, S);
    XS_LAUNCH_CHECK();
    return XS_OK;
}

// Queues one Gauss-Newton iteration: association + real sums, then (ncomp > 0) the derivative pass whose tail sums the
// partials and - when d_pose_out is given - runs the Gauss-Newton step per direction; with ncomp == 0 the step is a
// one-thread kernel.  d_pose_out == nullptr: accumulate only (the sums land in g_icp.d_sums).
int icp_iteration_async(IcpScratch *scp, const float *d_pose_curr, const float *d_vmap_curr, const float *d_nmap_curr, const xs_pose *prev,
                        xs_intr intr, const float *d_vmap_g_prev, const float *d_nmap_g_prev, int rows, int cols, const Batch &batch_full,
                        float dist_thres, float angle_thres, float *d_pose_out, int solve_mode, int *d_status,
                        double *d_log, cudaStream_t s, cudaStream_t s_real, int slot) {
    IcpScratch &g_icp = *scp;
    g_last_timed = scp;
    const BatchView &batch = batch_full.v;
    const int comps = batch.kind, dirs = batch.n;
    const int ncomp = batch.ncomp;
    const bool hessian = batch.kind == 2;
    // Split chains: the real part of an iteration (association, real sums, real Gauss-Newton step) does not depend on any
    // derivative component, so with a second stream the real chain of a frame - association + one-thread solve per
    // iteration - runs ahead on s_real, each iteration leaving its record, sums and real pose in a slot of its own, and
    // the derivative kernels follow back to back on s (their tail then only updates derivative components).
    const bool split = s_real != nullptr && ncomp > 0 && d_pose_out != nullptr;
    if (slot < 0 || slot >= IcpScratch::MAX_SLOTS || !split) slot = 0;
    const int npix = rows * cols;
    if ((double) (1 + ncomp) * 3.0 * npix >= 4294967296.0) {
        set_error("icp: a map set must hold fewer than 2^32 floats");
        return XS_ERR_ARG;
    }
    IcpParams P;
    // derivative pass decomposition (icp_deriv_kernel): pixel units of 256 * ppt pixels, (group, unit) items cut into equal
    // contiguous ranges for min(296, items) persistent CTAs; ppt shrinks until there are at least as many items as CTA slots
    // REDUCED pair sums (g_ij = b_ij - A_ij x, see icp_deriv_h_kernel) need the real solution before the derivative pass: split
    // chains with a solve; the ICP log and accumulate-only calls want the full A_ij, b_ij
    static const int h_full = env_int("XS_ICP_H_FULL", 0);  // A/B knob
    const bool h_reduced = hessian && split && d_log == nullptr && !h_full;
    static const int red_hp = env_int("XS_ICP_H_RED_HP", 2);  // pairs per task of the reduced form: 2 measured faster than 3 (profiles/r02_ab_table.md)
    const int hp = h_reduced ? ((red_hp == 2 || batch.ncurr > 0) ? 2 : 3) : 2;
    if (hessian && (g_icp.htasks_key != (const void *) batch.pairs || !g_icp.d_htasks || g_icp.htasks_hp != hp)) {
        // task table: one task per parameter (first-order sums), then the pairs in runs of up to hp that share their first parameter
        std::vector<HTask> tasks;
        for (int i = 0; i < batch.n; ++i) tasks.push_back(HTask{i, 0, {i, i, i}, {i, i, i}});
        for (int k = 0; k < batch.m;) {
            HTask t = {batch_full.h_pairs[k].x, 0, {0, 0, 0}, {0, 0, 0}};
            while (k < batch.m && t.np < hp && batch_full.h_pairs[k].x == t.i) {
                t.j[t.np] = batch_full.h_pairs[k].y;
                t.s[t.np] = batch.n + k;
                ++t.np;
                ++k;
            }
            for (int h = t.np; h < HPMAX; ++h) t.j[h] = t.j[0], t.s[h] = t.s[0];
            tasks.push_back(t);
        }
        g_icp.htasks_hp = hp;
        XS_CUDA(cudaStreamSynchronize(s));  // a table still in use by queued iterations (another pair list) must not be freed under them
        cudaFree(g_icp.d_htasks);
        g_icp.d_htasks = nullptr;
        XS_CUDA(cudaMalloc(&g_icp.d_htasks, tasks.size() * sizeof(HTask)));
        XS_CUDA(cudaMemcpy(g_icp.d_htasks, tasks.data(), tasks.size() * sizeof(HTask), cudaMemcpyHostToDevice));
        g_icp.n_htasks = (int) tasks.size();
        g_icp.htasks_key = (const void *) batch.pairs;
    }
    P.groups = hessian ? g_icp.n_htasks : (ncomp + 2) / 3;  // Hessian batch: one group per task
    P.batch = batch;
    P.htasks = g_icp.d_htasks;
    P.done_ticket = nullptr;
    static const int ppt_env = env_int("XS_ICP_PPT", 0), stages_env = env_int("XS_ICP_STAGES", 0);
    P.ppt = 4;
    const int cta_slots = g_icp.max_blocks;  // two persistent CTAs per SM
    while (P.ppt > 1 && (long long) div_up(npix, 256 * P.ppt) * P.groups < cta_slots) P.ppt >>= 1;
    if (ppt_env > 0) P.ppt = ppt_env < 32 ? ppt_env : 32;
    P.chunks = div_up(npix, 256 * P.ppt);
    const long long items = (long long) P.groups * P.chunks;
    const int deriv_grid = (int) (items < cta_slots ? items : cta_slots);
    P.max_writers = ncomp > 0 ? (int) (P.chunks / (items / deriv_grid)) + 2 : 0;
    // tile form of the reduced pass (icp_deriv_tile_kernel): one CTA per SM over 32-pixel tiles, all components of a tile at once
    static const int h_tile = env_int("XS_ICP_H_TILE", 1);
    const size_t smem_cap = 227 * 1024;
    // warps per CTA: 12 (default: the register file is allotted per four warps, so the twelfth costs nothing and takes pairs off
    // the others) or 11 (one first-order row + five pairs each at 10 parameters / 55 pairs); XS_ICP_TILE_NW selects.  Measured and
    // rejected: 16 warps x 4 pairs at 128 registers (spills; 0.48 vs 0.39 ms, profiles/r02_ab_table.md)
    static const int tile_nw_env = env_int("XS_ICP_TILE_NW", TILE_NW_DEFAULT);
    const int tile_nw = tile_nw_env == 12 ? 12 : 11;
    // (a small share of the pairs - one rank of an 8-GPU job holds 7 - leaves the tiles too little work per barrier: the task
    // form is faster there, 0.104 vs 0.120 ms per level-0 launch at 4 parameters + 7 pairs)
    const bool tile = h_reduced && h_tile && batch.m >= 24 && batch.n <= TILE_NW && batch.m <= TILE_NW * TILE_PPW &&
                      tile_smem(batch.n, batch.m, TILE_NW, TILE_PPW, 1, true) <= smem_cap;
    const int tile_grid = std::min((npix + 31) / 32, sm_count());
    P.tile_wp = nullptr;
    if (tile) {
        const int nw = tile_nw, ppw = TILE_PPW;
        if (g_icp.tile_wp_key != (const void *) batch.pairs || g_icp.tile_wp_nw != nw || g_icp.tile_wp_cur != (batch.ncurr > 0)) {
            // pairs -> warp slots: greedy on the estimated cost of a pair (more with parameters that move the intrinsics, and when
            // the row of its first parameter is not already held by the warp), warps that also form a first-order row start loaded
            std::vector<char> moves(batch.n, 0);
            if (batch.ncurr > 0 && batch_full.h_dintr)
                for (int i = 0; i < batch.n; ++i)
                    for (int c = 0; c < 4; ++c) moves[i] |= batch_full.h_dintr[4 * i + c] != 0.f;
            std::vector<double> cost(batch.m), load(nw, 0.0);
            std::vector<int> order(batch.m);
            for (int k = 0; k < batch.m; ++k) {
                const bool ci = moves[batch_full.h_pairs[k].x], cj = moves[batch_full.h_pairs[k].y];
                cost[k] = 1.0 + 0.12 * ci + 0.12 * cj + 0.12 * (ci && cj);
                order[k] = k;
            }
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
            for (int w = 0; w < nw && w < batch.n; ++w) load[w] = 0.5;  // a first-order row + its 27 sums: about half a pair
            std::vector<std::vector<int>> slots(nw);
            for (int k : order) {
                int best = -1;
                double best_load = 0.0;
                for (int w = 0; w < nw; ++w) {
                    if ((int) slots[w].size() >= ppw) continue;
                    bool has_i = false;
                    for (int o : slots[w]) has_i |= batch_full.h_pairs[o].x == batch_full.h_pairs[k].x;
                    const double l = load[w] + cost[k] + (has_i ? 0.0 : 0.13);
                    if (best < 0 || l < best_load) best = w, best_load = l;
                }
                slots[best].push_back(k);
                load[best] = best_load;
            }
            std::vector<int> wp((size_t) 16 * 5, -1);
            for (int w = 0; w < nw; ++w) {
                std::sort(slots[w].begin(), slots[w].end());
                for (size_t h = 0; h < slots[w].size(); ++h) wp[(size_t) w * ppw + h] = slots[w][h];
            }
            XS_CUDA(cudaStreamSynchronize(s));  // queued iterations may still read the previous table
            if (!g_icp.d_tile_wp) XS_CUDA(cudaMalloc(&g_icp.d_tile_wp, wp.size() * sizeof(int)));
            XS_CUDA(cudaMemcpy(g_icp.d_tile_wp, wp.data(), wp.size() * sizeof(int), cudaMemcpyHostToDevice));
            g_icp.tile_wp_key = (const void *) batch.pairs;
            g_icp.tile_wp_nw = nw;
            g_icp.tile_wp_cur = batch.ncurr > 0;
        }
        P.tile_wp = g_icp.d_tile_wp;
    }
    size_t dpart_need = (size_t) P.groups * P.max_writers * 81;
    if (tile) dpart_need = std::max(dpart_need, (size_t) tile_grid * (batch.n * 27 + batch.m * 6));
    int rc = icp_reserve(scp, ncomp, npix, dpart_need, P.groups, split ? IcpScratch::MAX_SLOTS : 1);
    if (rc != XS_OK) return rc;
    P.done_ticket = g_icp.d_done_ticket;
    const int nvals = 27 * (1 + ncomp);
    // only the current pose's derivative components enter the rows (s = Rcurr*v + tcurr); the previous pose is used
    // for the real projection only (ICP.cu:206-217 takes real parts)
    for (int i = 0; i < 9; ++i) P.prev.R[i] = prev->R[i];
    for (int i = 0; i < 3; ++i) P.prev.t[i] = prev->t[i];
    P.pose_curr = d_pose_curr;
    P.vmap_curr = d_vmap_curr;
    P.nmap_curr = d_nmap_curr;
    P.vmap_prev = d_vmap_g_prev;
    P.nmap_prev = d_nmap_g_prev;
    P.intr = intr;
    P.rows = rows;
    P.cols = cols;
    P.dirs = dirs;
    P.ncomp = ncomp;
    P.dist_thres = dist_thres;
    P.angle_thres = angle_thres;
    P.rec_idx = g_icp.d_rec_idx ? g_icp.d_rec_idx + (size_t) slot * g_icp.cap_pix : nullptr;
    P.rec_f = g_icp.d_rec_f ? g_icp.d_rec_f + (size_t) slot * g_icp.cap_pix * 4 : nullptr;
    P.partials = g_icp.d_partials;
    P.dpartials = g_icp.d_dpartials;
    P.sums = g_icp.d_sums + (size_t) slot * nvals;
    P.ticket = g_icp.d_ticket;
    P.group_ticket = g_icp.d_group_ticket;
    P.tiles_x = div_up(cols, 32);
    P.tiles_y = div_up(rows, 8);
    SolveParams S;
    S.status_in = nullptr;
    S.staged = 0;
    S.sums = P.sums;
    S.deriv_only = split ? 1 : 0;
    S.real_cache = nullptr;
    if (split) {
        if (!g_icp.d_real_cache) XS_CUDA(cudaMalloc(&g_icp.d_real_cache, (size_t) IcpScratch::MAX_SLOTS * REAL_CACHE * sizeof(double)));
        S.real_cache = g_icp.d_real_cache + (size_t) slot * REAL_CACHE;
    }
    S.pose_in = d_pose_curr;
    S.pose_out = d_pose_out;
    static const int no_tail = env_int("XS_ICP_NO_TAIL", 0);  // timing experiment only: the derivative tails skip the solve (derivative poses are then stale)
    S.status = d_status;
    S.log = d_log;
    S.dirs = dirs;
    S.ncomp = ncomp;
    S.solve_mode = solve_mode;
    const int ntiles = P.tiles_x * P.tiles_y;
    const int grid = ntiles < g_icp.max_blocks ? ntiles : g_icp.max_blocks;
    // the real step rides in the tail of the association kernel when it depends on no derivative component: under split
    // chains (the derivative kernels then update derivative components only) and when there are none at all
    SolveParams SR = S;  // one thread, no derivative components, owns the status flags
    SR.dirs = 0;
    SR.ncomp = 0;
    SR.deriv_only = 0;
    if (split) SR.log = nullptr;
    if (!(split || (ncomp == 0 && d_pose_out))) SR.pose_out = nullptr;
    icp_assoc_kernel<<<grid, dim3(32, 8), 0, split ? s_real : s>>>(P, SR);
    XS_LAUNCH_CHECK();
    if (split) {
        if (!g_icp.ev_real[slot]) XS_CUDA(cudaEventCreateWithFlags(&g_icp.ev_real[slot], cudaEventDisableTiming));
        XS_CUDA(cudaEventRecord(g_icp.ev_real[slot], s_real));
        XS_CUDA(cudaStreamWaitEvent(s, g_icp.ev_real[slot], 0));
    }
    if (ncomp > 0) {
        const int tslot = g_icp.n_timed < IcpScratch::MAX_TIMED ? g_icp.n_timed : -1;
        if (tslot >= 0) {
            if (!g_icp.ev0[tslot]) {
                XS_CUDA(cudaEventCreate(&g_icp.ev0[tslot]));
                XS_CUDA(cudaEventCreate(&g_icp.ev1[tslot]));
            }
            XS_CUDA(cudaEventRecord(g_icp.ev0[tslot], s));
        }
        // pipeline depth: 2 stages (one pixel ahead) measured faster than 3 on B200 (0.475 vs 0.542 ms per level-0 launch at
        // 55 directions: the deeper pipeline costs L1 capacity and does not raise issue utilisation); XS_ICP_STAGES=3 selects it
        const int stages = stages_env == 3 ? 3 : 2;
        if (no_tail && split) S.pose_out = nullptr;
        if (hessian && tile) {
            // look-ahead of the staging: as deep as fits beside ~96 KB of L1 (the L1 left over holds the lines of the in-flight
            // gathers: profiles/r02_ab_table.md) - one tile at 10 parameters / 55 pairs, four for an 8-rank share
            static const int tile_pipe = env_int("XS_ICP_TILE_PIPE", 1), tile_depth_env = env_int("XS_ICP_TILE_DEPTH", 0);
            const bool cur = batch.ncurr > 0;
            int tile_depth = 1;
            for (int d : {2, 4})
                if (tile_smem(batch.n, batch.m, TILE_NW, TILE_PPW, d, true) <= 160 * 1024) tile_depth = d;
            if (tile_depth_env == 1 || tile_depth_env == 2 || tile_depth_env == 4) tile_depth = tile_depth_env;
            if (tile_smem(batch.n, batch.m, TILE_NW, TILE_PPW, tile_depth, true) > smem_cap) tile_depth = 1;
#define XS_TILE(NW_, PPW_, PIPE_, DEPTH_) \
    (cur ? launch_deriv_tile<NW_, PPW_, true, PIPE_, DEPTH_>(P, S, tile_grid, s) : launch_deriv_tile<NW_, PPW_, false, PIPE_, DEPTH_>(P, S, tile_grid, s))
            if (!tile_pipe)
                rc = XS_TILE(TILE_NW, TILE_PPW, false, 1);
            else if (tile_nw == 12)
                rc = tile_depth == 4 ? XS_TILE(12, TILE_PPW, true, 4) : tile_depth == 2 ? XS_TILE(12, TILE_PPW, true, 2) : XS_TILE(12, TILE_PPW, true, 1);
            else if (tile_depth == 4)
                rc = XS_TILE(TILE_NW, TILE_PPW, true, 4);
            else if (tile_depth == 2)
                rc = XS_TILE(TILE_NW, TILE_PPW, true, 2);
            else
                rc = XS_TILE(TILE_NW, TILE_PPW, true, 1);
#undef XS_TILE
        } else if (hessian) {
            if (batch.ncurr > 0)  // a parameter moves the intrinsics: the current-frame maps carry derivative components
                rc = h_reduced ? launch_deriv_h<2, 2, 2, true, true>(P, S, deriv_grid, s) : launch_deriv_h<2, 2, 2, false, true>(P, S, deriv_grid, s);
            else
                rc = h_reduced ? (hp == 2 ? launch_deriv_h<2, 2, 2, true, false>(P, S, deriv_grid, s) : launch_deriv_h<2, 2, 3, true, false>(P, S, deriv_grid, s))
                               : launch_deriv_h<2, 2, 2, false, false>(P, S, deriv_grid, s);
        }
        else if (comps == 1)
            rc = stages == 2 ? launch_deriv<1, 2>(P, S, deriv_grid, s) : launch_deriv<1, DERIV_MAX_STAGES>(P, S, deriv_grid, s);
        else
            rc = stages == 2 ? launch_deriv<3, 2>(P, S, deriv_grid, s) : launch_deriv<3, DERIV_MAX_STAGES>(P, S, deriv_grid, s);
        if (rc != XS_OK) return rc;
        if (tslot >= 0) {
            XS_CUDA(cudaEventRecord(g_icp.ev1[tslot], s));
            g_icp.timed_npix[tslot] = npix;
            ++g_icp.n_timed;
        }
    }
    return XS_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// Combined::computeOptimizeMatrix (ICP.cu:283-355): per-correspondence analytic Jacobian (3x4) and Gauss-Newton Hessian
// (12x12, upper triangle = 78 sums) of the point-to-plane energy with respect to the 12 entries of [R | t], real parts.
// The reference runs 12 + 78 sequential 256-thread shared-memory tree reductions per block and 90 thrust::reduce calls;
// here a thread accumulates its pixels in double, the block reduces with shuffles and the last block adds the partials.
constexpr int OPT_VALS = 12 + 78;
__global__ void __launch_bounds__(256) icp_optimize_matrix_kernel(const IcpParams P, double *partials, double *out,
                                                                   unsigned long long *count_out) {
    __shared__ float s_curr[12];
    __shared__ double s_red[8][OPT_VALS];
    __shared__ unsigned int s_cnt[8];
    __shared__ bool s_last;
    const int tid = threadIdx.y * 32 + threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < 12) s_curr[tid] = P.pose_curr[tid];
    __syncthreads();
    const size_t plane = (size_t) P.rows * P.cols;
    const int ntiles = P.tiles_x * P.tiles_y;
    double acc[OPT_VALS];
#pragma unroll
    for (int e = 0; e < OPT_VALS; ++e) acc[e] = 0.0;
    unsigned int cnt = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int x = (tile % P.tiles_x) * 32 + threadIdx.x;
        const int y = (tile / P.tiles_x) * 8 + threadIdx.y;
        float vcx, vcy, vcz, gx, gy, gz;
        int ux, uy;
        float n1[3], p1[3];
        if (!(x < P.cols && y < P.rows && search_newton_real(P, s_curr, x, y, plane, vcx, vcy, vcz, gx, gy, gz, ux, uy, n1, p1))) continue;
        ++cnt;
        const float p0[4] = {vcx, vcy, vcz, 1.f};
        // proj_norm = (p0_trans - p1) . n1, ICP.cu:312-314
        const float proj = (gx - p1[0]) * n1[0] + (gy - p1[1]) * n1[1] + (gz - p1[2]) * n1[2];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i * 4 + j] += (double) (2 * n1[i] * proj * p0[j]);  // :315
        int e = 12;
#pragma unroll
        for (int a = 0; a < 12; ++a)
#pragma unroll
            for (int b = a; b < 12; ++b, ++e) acc[e] += (double) (2 * (p0[a % 4] * (n1[a / 4] * n1[b / 4] * p0[b % 4])));  // :334-342
    }
#pragma unroll
    for (int e = 0; e < OPT_VALS; ++e) {
        double v = acc[e];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) s_red[warp][e] = v;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    if (tid < OPT_VALS) {
        double v = 0;
        for (int w = 0; w < 8; ++w) v += s_red[w][tid];
        partials[(size_t) blockIdx.x * OPT_VALS + tid] = v;
    }
    if (tid == 0) {
        unsigned int c = 0;
        for (int w = 0; w < 8; ++w) c += s_cnt[w];
        atomicAdd(count_out, (unsigned long long) c);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(P.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (tid < OPT_VALS) {
        double v = 0;
        for (unsigned b = 0; b < gridDim.x; ++b) v += __ldcg(partials + (size_t) b * OPT_VALS + tid);
        out[tid] = v;
    }
    if (tid == 0) *P.ticket = 0u;
}

void icp_timing_reset(IcpScratch *sc) {
    if (sc) sc->n_timed = 0;
}

}  // namespace xs

using namespace xs;

// Device durations (CUDA events on the launching stream) of the icp_deriv_kernel launches since the last
// xs_kinfu_pose_estimate / xs_estimate_combined began; call after the stream has been synchronised.
extern "C" int xs_icp_deriv_times(float *ms, int *npix, int max_n) {
    int n = 0;
    if (!g_last_timed) return 0;
    IcpScratch &g_icp = *g_last_timed;
    for (; n < g_icp.n_timed && n < max_n; ++n) {
        if (cudaEventElapsedTime(&ms[n], g_icp.ev0[n], g_icp.ev1[n]) != cudaSuccess) break;
        npix[n] = g_icp.timed_npix[n];
    }
    return n;
}

extern "C" int xs_estimate_combined(const xs_pose *curr, const float *d_vmap_curr, const float *d_nmap_curr,
                                    const xs_pose *prev, xs_intr intr, const float *d_vmap_g_prev,
                                    const float *d_nmap_g_prev, int rows, int cols, int comps, int dirs,
                                    float dist_thres, float angle_thres, double *A_host, double *b_host,
                                    void *stream) {
    if (!curr || !prev || !d_vmap_curr || !d_nmap_curr || !d_vmap_g_prev || !d_nmap_g_prev || !A_host || !b_host ||
        rows <= 0 || cols <= 0 || (comps != 1 && comps != 2 && comps != 3) || dirs < 0)
        return XS_ERR_ARG;
    const int ncomp = batch_ncomp(comps, dirs, -1);  // comps == 2: dirs parameters and all their pairs
    Batch batch;
    if (batch_init(batch, comps, dirs, -1, nullptr) != XS_OK) return XS_ERR_ARG;
    struct BatchGuard {
        Batch &b;
        ~BatchGuard() { batch_free(b); }
    } batch_guard{batch};
    if (curr->ncomp != ncomp || prev->ncomp != ncomp) {
        set_error("xs_estimate_combined: pose derivative component count mismatch");
        return XS_ERR_ARG;
    }
    cudaStream_t s = (cudaStream_t) stream;
    IcpScratch *scp = seam_scratch();
    IcpScratch &g_icp = *scp;
    int rc = icp_reserve(scp, ncomp, rows * cols, 0, 0);
    if (rc != XS_OK) return rc;
    icp_timing_reset(scp);
    XS_CUDA(cudaStreamSynchronize(s));  // pinned staging reuse
    float *h = g_icp.h_pose;
    for (int e = 0; e < 9; ++e) h[e] = curr->R[e];
    for (int e = 0; e < 3; ++e) h[9 + e] = curr->t[e];
    for (int q = 0; q < ncomp; ++q) {
        float *hc = h + (size_t) (1 + q) * 12;
        for (int e = 0; e < 9; ++e) hc[e] = curr->dR[q * 9 + e];
        for (int e = 0; e < 3; ++e) hc[9 + e] = curr->dt[q * 3 + e];
    }
    XS_CUDA(cudaMemcpyAsync(g_icp.d_pose, h, (size_t) (1 + ncomp) * 12 * sizeof(float), cudaMemcpyHostToDevice, s));
    rc = icp_iteration_async(scp, g_icp.d_pose, d_vmap_curr, d_nmap_curr, prev, intr, d_vmap_g_prev, d_nmap_g_prev, rows, cols,
                             batch, dist_thres, angle_thres, nullptr, 0, nullptr, nullptr, s, nullptr, 0);
    if (rc != XS_OK) return rc;
    const int nvals = 27 * (1 + ncomp);
    XS_CUDA(cudaMemcpyAsync(g_icp.h_sums, g_icp.d_sums, (size_t) nvals * sizeof(double), cudaMemcpyDeviceToHost, s));
    XS_CUDA(cudaStreamSynchronize(s));  // estimateCombined syncs and downloads, ICP.cu:414-417
    // unpack upper-triangular order into column-major symmetric A and b, ICP.cu:419-428
    for (int c = 0; c <= ncomp; ++c) {
        const double *v = g_icp.h_sums + (size_t) c * 27;
        double *A = A_host + (size_t) c * 36, *b = b_host + (size_t) c * 6;
        int shift = 0;
        for (int i = 0; i < 6; ++i)
            for (int j = i; j < 7; ++j) {
                const double val = v[shift++];
                if (j == 6)
                    b[i] = val;
                else
                    A[j * 6 + i] = A[i * 6 + j] = val;
            }
    }
    return XS_OK;
}

extern "C" long xs_compute_optimize_matrix(const xs_pose *curr, const float *d_vmap_curr, const float *d_nmap_curr,
                                           const xs_pose *prev, xs_intr intr, const float *d_vmap_g_prev,
                                           const float *d_nmap_g_prev, int rows, int cols, float dist_thres, float angle_thres,
                                           double *jacobi_host, double *hessian_host, void *stream) {
    if (!curr || !prev || !d_vmap_curr || !d_nmap_curr || !d_vmap_g_prev || !d_nmap_g_prev || !jacobi_host || !hessian_host ||
        rows <= 0 || cols <= 0) {
        set_error("xs_compute_optimize_matrix: null argument");
        return -1;
    }
    cudaStream_t s = (cudaStream_t) stream;
    IcpScratch *scp = seam_scratch();
    IcpScratch &g_icp = *scp;
    if (icp_reserve(scp, 0, rows * cols, 0, 0) != XS_OK) return -1;
    const int grid = 2 * sm_count();
    double *d_buf = nullptr;
    if (cudaMalloc(&d_buf, ((size_t) (grid + 1) * OPT_VALS + 1) * sizeof(double) + 12 * sizeof(float)) != cudaSuccess) {
        set_error("xs_compute_optimize_matrix: cudaMalloc failed");
        return -1;
    }
    double *d_out = d_buf + (size_t) grid * OPT_VALS;
    unsigned long long *d_count = reinterpret_cast<unsigned long long *>(d_out + OPT_VALS);
    float *d_pose = reinterpret_cast<float *>(d_count + 1);
    float h_pose[12];
    for (int e = 0; e < 9; ++e) h_pose[e] = curr->R[e];
    for (int e = 0; e < 3; ++e) h_pose[9 + e] = curr->t[e];
    cudaMemcpyAsync(d_pose, h_pose, sizeof(h_pose), cudaMemcpyHostToDevice, s);
    cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), s);
    IcpParams P = {};
    for (int i = 0; i < 9; ++i) P.prev.R[i] = prev->R[i];
    for (int i = 0; i < 3; ++i) P.prev.t[i] = prev->t[i];
    P.pose_curr = d_pose;
    P.vmap_curr = d_vmap_curr;
    P.nmap_curr = d_nmap_curr;
    P.vmap_prev = d_vmap_g_prev;
    P.nmap_prev = d_nmap_g_prev;
    P.intr = intr;
    P.rows = rows;
    P.cols = cols;
    P.dist_thres = dist_thres;
    P.angle_thres = angle_thres;
    P.ticket = g_icp.d_ticket;
    P.tiles_x = div_up(cols, 32);
    P.tiles_y = div_up(rows, 8);
    const int ntiles = P.tiles_x * P.tiles_y;
    icp_optimize_matrix_kernel<<<ntiles < grid ? ntiles : grid, dim3(32, 8), 0, s>>>(P, d_buf, d_out, d_count);
    long rc = -1;
    double h_out[OPT_VALS];
    unsigned long long h_count = 0;
    if (cudaGetLastError() == cudaSuccess && cudaMemcpyAsync(h_out, d_out, sizeof(h_out), cudaMemcpyDeviceToHost, s) == cudaSuccess &&
        cudaMemcpyAsync(&h_count, d_count, sizeof(h_count), cudaMemcpyDeviceToHost, s) == cudaSuccess &&
        cudaStreamSynchronize(s) == cudaSuccess) {
        ++xs::g_launches;
        // jacobi_host(i, j), i < 3, j < 4 -> row-major [3][4]; hessian_host[i1][j1](i2, j2) -> [12][12], index i * 4 + j (ICP.cu:471-488)
        for (int e = 0; e < 12; ++e) jacobi_host[e] = h_out[e];
        int e = 12;
        for (int a = 0; a < 12; ++a)
            for (int b = a; b < 12; ++b, ++e) hessian_host[a * 12 + b] = hessian_host[b * 12 + a] = h_out[e];
        rc = (long) h_count;
    } else {
        set_error("xs_compute_optimize_matrix: CUDA failure");
    }
    cudaFree(d_buf);
    return rc;
}
