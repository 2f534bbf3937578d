// icp.cu — direction-batched projective point-to-plane ICP normal equations.
//
// Replaces Combined::{search_newton, operator()} / combinedKernel (XKinectFusion/src/ICP.cu:166-281,357),
// TranformReduction / TransformEstimatorKernel (ICP.cu:120-164) and estimateCombined (ICP.cu:365-429).
//
// Data association (projection, bounds, NaN, distance and angle gates) is evaluated once per pixel on
// real parts; the 7-vector row [cross(s,n), n, n.(d-s)] is then formed per direction tile as Jet<C,K>
// numbers and its 27 upper-triangular products are widened to double exactly as the reference does
// (product in float, sum in double, ICP.cu:273-274).  Instead of the reference's 27 sequential
// 256-thread shared-memory tree reductions plus a second kernel, each warp reduces a 32-wide vector of
// sums with a 31-shuffle transpose reduction, warps are combined in a fixed order, per-block partials go
// to global memory and the last block to finish (ticket) adds them in block order: one launch per
// iteration, deterministic summation order.
#include "xs_common.cuh"

#include <utility>

namespace xs {

struct IcpParams {
    DevPose curr, prev;  // prev.R = Rprev_inv, prev.t = tprev
    const float *dpose_curr, *dpose_prev;
    const float *vmap_curr, *nmap_curr;  // [3][rows][cols]
    const float *vmap_prev, *nmap_prev;  // [(1+ncomp)][3][rows][cols]
    xs_intr intr;
    int rows, cols, dirs, ncomp;
    float dist_thres, angle_thres;
    double *partials;  // [gridDim.x][27*(1+ncomp)]
    double *sums;      // [27*(1+ncomp)]
    unsigned int *ticket;
    int tiles_x, tiles_y;
};

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

template <int C, int K> struct Row7 {
    Jet<C, K> r[7];
};

// products row_i * row_j for e = 0..26 with compile-time row indices (fold over E)
template <int C, int K, int... E>
XS_DEV void fill_real(double (&v)[32], const Row7<C, K> &row, std::integer_sequence<int, E...>) {
    ((v[E] = (double) __fmul_rn(row.r[tri_i(E)].v, row.r[tri_j(E)].v)), ...);
}
template <int C, int K> XS_DEV float prod_deriv(const Jet<C, K> &a, const Jet<C, K> &b, int i) {
    if (C == 1 || (i % 3) != 2) return fmaf(a.v, b.d[i], a.d[i] * b.v);
    // eps1eps2 of direction i/3
    return fmaf(a.v, b.d[i], fmaf(a.d[i], b.v, fmaf(a.d[i - 2], b.d[i - 1], a.d[i - 1] * b.d[i - 2])));
}
template <int C, int K, int... E>
XS_DEV void fill_deriv(double (&v)[32], const Row7<C, K> &row, int i, std::integer_sequence<int, E...>) {
    ((v[E] = (double) prod_deriv<C, K>(row.r[tri_i(E)], row.r[tri_j(E)], i)), ...);
}

template <int C, int K> __global__ void __launch_bounds__(256) icp_kernel(const IcpParams P) {
    constexpr int N = C * K;
    extern __shared__ double s_mem[];
    const int nvals = 27 * (1 + P.ncomp);
    double *s_acc = s_mem;            // [nvals]
    double *s_stage = s_mem + nvals;  // [8 warps][1+N][32]
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < nvals; i += 256) s_acc[i] = 0.0;
    __syncthreads();

    const size_t plane = (size_t) P.rows * P.cols;
    const int ntiles = P.tiles_x * P.tiles_y;
    const int dtiles = P.dirs > 0 ? (P.dirs + K - 1) / K : 1;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int x = (tile % P.tiles_x) * 32 + threadIdx.x;
        const int y = (tile / P.tiles_x) * 8 + threadIdx.y;
        // ---------------- search_newton, ICP.cu:196-244 (real parts)
        bool found = false;
        float vcx = 0, vcy = 0, vcz = 0;  // vcurr (camera frame)
        int ux = 0, uy = 0;
        if (x < P.cols && y < P.rows) {
            const size_t pix = (size_t) y * P.cols + x;
            const float ncx = P.nmap_curr[pix];
            if (!isnan(ncx)) {
                const float ncy = P.nmap_curr[pix + plane], ncz = P.nmap_curr[pix + 2 * plane];
                vcx = P.vmap_curr[pix];
                vcy = P.vmap_curr[pix + plane];
                vcz = P.vmap_curr[pix + 2 * plane];
                const float *R = P.curr.R, *t = P.curr.t, *Q = P.prev.R, *tp = P.prev.t;
                // vcurr_g = Rcurr * vcurr + tcurr
                const float gx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[0], vcx), __fmul_rn(R[1], vcy)), __fmul_rn(R[2], vcz)), t[0]);
                const float gy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[3], vcx), __fmul_rn(R[4], vcy)), __fmul_rn(R[5], vcz)), t[1]);
                const float gz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[6], vcx), __fmul_rn(R[7], vcy)), __fmul_rn(R[8], vcz)), t[2]);
                // vcurr_cp = Rprev_inv * (vcurr_g - tprev)
                const float ex = __fsub_rn(gx, tp[0]), ey = __fsub_rn(gy, tp[1]), ez = __fsub_rn(gz, tp[2]);
                const float px = __fadd_rn(__fadd_rn(__fmul_rn(Q[0], ex), __fmul_rn(Q[1], ey)), __fmul_rn(Q[2], ez));
                const float py = __fadd_rn(__fadd_rn(__fmul_rn(Q[3], ex), __fmul_rn(Q[4], ey)), __fmul_rn(Q[5], ez));
                const float pz = __fadd_rn(__fadd_rn(__fmul_rn(Q[6], ex), __fmul_rn(Q[7], ey)), __fmul_rn(Q[8], ez));
                ux = __float2int_rn(__fadd_rn(__fdiv_rn(__fmul_rn(px, P.intr.fx), pz), P.intr.cx));
                uy = __float2int_rn(__fadd_rn(__fdiv_rn(__fmul_rn(py, P.intr.fy), pz), P.intr.cy));
                if (!(ux < 0 || uy < 0 || ux >= P.cols || uy >= P.rows || pz < 0)) {
                    const size_t q = (size_t) uy * P.cols + ux;
                    const float npx = P.nmap_prev[q];
                    if (!isnan(npx)) {
                        const float npy = P.nmap_prev[q + plane], npz = P.nmap_prev[q + 2 * plane];
                        const float vpx = P.vmap_prev[q], vpy = P.vmap_prev[q + plane], vpz = P.vmap_prev[q + 2 * plane];
                        // dist = norm(vprev_g - vcurr_g)
                        const float ddx = __fsub_rn(vpx, gx), ddy = __fsub_rn(vpy, gy), ddz = __fsub_rn(vpz, gz);
                        const float dist = __fsqrt_rn(
                            __fadd_rn(__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)), __fmul_rn(ddz, ddz)));
                        if (!(dist > P.dist_thres)) {
                            // sine = norm(cross(Rcurr * ncurr, nprev_g))
                            const float mx = __fadd_rn(__fadd_rn(__fmul_rn(R[0], ncx), __fmul_rn(R[1], ncy)), __fmul_rn(R[2], ncz));
                            const float my = __fadd_rn(__fadd_rn(__fmul_rn(R[3], ncx), __fmul_rn(R[4], ncy)), __fmul_rn(R[5], ncz));
                            const float mz = __fadd_rn(__fadd_rn(__fmul_rn(R[6], ncx), __fmul_rn(R[7], ncy)), __fmul_rn(R[8], ncz));
                            const float kx = __fsub_rn(__fmul_rn(my, npz), __fmul_rn(mz, npy));
                            const float ky = __fsub_rn(__fmul_rn(mz, npx), __fmul_rn(mx, npz));
                            const float kz = __fsub_rn(__fmul_rn(mx, npy), __fmul_rn(my, npx));
                            const float sine = __fsqrt_rn(
                                __fadd_rn(__fadd_rn(__fmul_rn(kx, kx), __fmul_rn(ky, ky)), __fmul_rn(kz, kz)));
                            found = !(sine >= P.angle_thres);
                        }
                    }
                }
            }
        }
        // ---------------- rows and products per direction tile, ICP.cu:254-279
        for (int dt = 0; dt < dtiles; ++dt) {
            const int k0 = dt * K;
            Row7<C, K> row;
            if (found) {
                const JetPose<C, K> cur = load_pose<C, K>(P.curr, P.dpose_curr, k0, P.dirs);
                const Jet3<C, K> vc = {jconst<C, K>(vcx), jconst<C, K>(vcy), jconst<C, K>(vcz)};
                const Jet3<C, K> s = jrot(cur, vc) + cur.t;
                const size_t q = (size_t) uy * P.cols + ux;
                Jet3<C, K> n, d;
                n.x.v = P.nmap_prev[q];
                n.y.v = P.nmap_prev[q + plane];
                n.z.v = P.nmap_prev[q + 2 * plane];
                d.x.v = P.vmap_prev[q];
                d.y.v = P.vmap_prev[q + plane];
                d.z.v = P.vmap_prev[q + 2 * plane];
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const int comp = k0 * C + i;
                    if (comp < P.ncomp) {
                        const size_t o = q + (size_t) (1 + comp) * 3 * plane;
                        n.x.d[i] = P.nmap_prev[o];
                        n.y.d[i] = P.nmap_prev[o + plane];
                        n.z.d[i] = P.nmap_prev[o + 2 * plane];
                        d.x.d[i] = P.vmap_prev[o];
                        d.y.d[i] = P.vmap_prev[o + plane];
                        d.z.d[i] = P.vmap_prev[o + 2 * plane];
                    } else {
                        n.x.d[i] = n.y.d[i] = n.z.d[i] = 0.f;
                        d.x.d[i] = d.y.d[i] = d.z.d[i] = 0.f;
                    }
                }
                const Jet3<C, K> cr = jcross(s, n);
                row.r[0] = cr.x;
                row.r[1] = cr.y;
                row.r[2] = cr.z;
                row.r[3] = n.x;
                row.r[4] = n.y;
                row.r[5] = n.z;
                row.r[6] = jdot(n, d - s);
            } else {
#pragma unroll
                for (int i = 0; i < 7; ++i) row.r[i] = jconst<C, K>(0.f);
            }
            // one 32-wide group per component: 27 products, widened to double (ICP.cu:273-274)
            double *stage = s_stage + (size_t) warp * (1 + N) * 32;
            if (k0 == 0) {
                double v[32];
#pragma unroll
                for (int e = 27; e < 32; ++e) v[e] = 0.0;
                fill_real<C, K>(v, row, std::make_integer_sequence<int, 27>());
                stage[lane] = warp_transpose_reduce(v);
            }
#pragma unroll
            for (int i = 0; i < N; ++i) {
                double v[32];
#pragma unroll
                for (int e = 27; e < 32; ++e) v[e] = 0.0;
                fill_deriv<C, K>(v, row, i, std::make_integer_sequence<int, 27>());
                stage[(1 + i) * 32 + lane] = warp_transpose_reduce(v);
            }
            __syncthreads();
            // combine the 8 warps in fixed order; thread (g, e) owns accumulator (component, product e)
            for (int idx = tid; idx < (1 + N) * 32; idx += 256) {
                const int g = idx >> 5, e = idx & 31;
                if (e >= 27) continue;
                if (g == 0 && k0 != 0) continue;
                const int comp = (g == 0) ? 0 : 1 + k0 * C + (g - 1);
                if (comp > P.ncomp) continue;
                double sum = 0.0;
#pragma unroll
                for (int w = 0; w < 8; ++w) sum += s_stage[(size_t) w * (1 + N) * 32 + idx];
                s_acc[comp * 27 + e] += sum;
            }
            __syncthreads();
        }
    }
    // ---------------- block partials, then the last block reduces over blocks in block order
    double *mine = P.partials + (size_t) blockIdx.x * nvals;
    for (int i = tid; i < nvals; i += 256) mine[i] = s_acc[i];
    __threadfence();
    __shared__ bool s_last;
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(P.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int i = tid; i < nvals; i += 256) {
        double sum = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) sum += __ldcg(P.partials + (size_t) b * nvals + i);
        P.sums[i] = sum;
    }
    if (tid == 0) *P.ticket = 0u;
}

// persistent scratch of the ICP operator (gbuf / mbuf of the reference, ICP.cu:400-403)
struct IcpScratch {
    double *d_partials = nullptr, *d_sums = nullptr, *h_sums = nullptr;
    unsigned int *d_ticket = nullptr;
    float *d_dpose = nullptr, *h_dpose = nullptr;
    int cap_vals = 0, cap_comp = -1;
    int max_blocks = 296;
};
static IcpScratch g_icp;

static int icp_reserve(int ncomp) {
    const int nvals = 27 * (1 + ncomp);
    if (nvals > g_icp.cap_vals) {
        cudaFree(g_icp.d_partials);
        cudaFree(g_icp.d_sums);
        cudaFreeHost(g_icp.h_sums);
        XS_CUDA(cudaMalloc(&g_icp.d_partials, (size_t) g_icp.max_blocks * nvals * sizeof(double)));
        XS_CUDA(cudaMalloc(&g_icp.d_sums, (size_t) nvals * sizeof(double)));
        XS_CUDA(cudaMallocHost(&g_icp.h_sums, (size_t) nvals * sizeof(double)));
        g_icp.cap_vals = nvals;
    }
    if (!g_icp.d_ticket) {
        XS_CUDA(cudaMalloc(&g_icp.d_ticket, sizeof(unsigned int)));
        XS_CUDA(cudaMemset(g_icp.d_ticket, 0, sizeof(unsigned int)));
    }
    if (ncomp > g_icp.cap_comp) {
        cudaFree(g_icp.d_dpose);
        cudaFreeHost(g_icp.h_dpose);
        const size_t n = (size_t) (ncomp > 0 ? ncomp : 1) * 24;
        XS_CUDA(cudaMalloc(&g_icp.d_dpose, n * sizeof(float)));
        XS_CUDA(cudaMallocHost(&g_icp.h_dpose, n * sizeof(float)));
        g_icp.cap_comp = ncomp;
    }
    return XS_OK;
}

}  // namespace xs

using namespace xs;

extern "C" int xs_estimate_combined(const xs_pose *curr, const float *d_vmap_curr, const float *d_nmap_curr,
                                    const xs_pose *prev, xs_intr intr, const float *d_vmap_g_prev,
                                    const float *d_nmap_g_prev, int rows, int cols, int comps, int dirs,
                                    float dist_thres, float angle_thres, double *A_host, double *b_host,
                                    void *stream) {
    if (!curr || !prev || !d_vmap_curr || !d_nmap_curr || !d_vmap_g_prev || !d_nmap_g_prev || !A_host || !b_host ||
        rows <= 0 || cols <= 0 || (comps != 1 && comps != 3) || dirs < 0)
        return XS_ERR_ARG;
    const int ncomp = comps * dirs;
    if (curr->ncomp != ncomp || prev->ncomp != ncomp) {
        set_error("xs_estimate_combined: pose derivative component count mismatch");
        return XS_ERR_ARG;
    }
    cudaStream_t s = (cudaStream_t) stream;
    int rc = icp_reserve(ncomp);
    if (rc != XS_OK) return rc;
    for (int q = 0; q < ncomp; ++q) {
        float *hc = g_icp.h_dpose + q * 12, *hp = g_icp.h_dpose + (size_t) ncomp * 12 + q * 12;
        for (int e = 0; e < 9; ++e) {
            hc[e] = curr->dR[q * 9 + e];
            hp[e] = prev->dR[q * 9 + e];
        }
        for (int e = 0; e < 3; ++e) {
            hc[9 + e] = curr->dt[q * 3 + e];
            hp[9 + e] = prev->dt[q * 3 + e];
        }
    }
    if (ncomp)
        XS_CUDA(cudaMemcpyAsync(g_icp.d_dpose, g_icp.h_dpose, (size_t) ncomp * 24 * sizeof(float), cudaMemcpyHostToDevice, s));
    IcpParams P;
    for (int i = 0; i < 9; ++i) {
        P.curr.R[i] = curr->R[i];
        P.prev.R[i] = prev->R[i];
    }
    for (int i = 0; i < 3; ++i) {
        P.curr.t[i] = curr->t[i];
        P.prev.t[i] = prev->t[i];
    }
    P.dpose_curr = g_icp.d_dpose;
    P.dpose_prev = g_icp.d_dpose + (size_t) ncomp * 12;
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
    P.partials = g_icp.d_partials;
    P.sums = g_icp.d_sums;
    P.ticket = g_icp.d_ticket;
    P.tiles_x = div_up(cols, 32);
    P.tiles_y = div_up(rows, 8);
    const int ntiles = P.tiles_x * P.tiles_y;
    const int grid = ntiles < g_icp.max_blocks ? ntiles : g_icp.max_blocks;
    const int nvals = 27 * (1 + ncomp);
    const int N = (comps == 1) ? 6 : 6;
    const size_t smem = ((size_t) nvals + (size_t) 8 * (1 + N) * 32) * sizeof(double);
    dim3 blk(32, 8);
    if (comps == 1) {
        XS_CUDA(cudaFuncSetAttribute(icp_kernel<1, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        icp_kernel<1, 6><<<grid, blk, smem, s>>>(P);
    } else {
        XS_CUDA(cudaFuncSetAttribute(icp_kernel<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        icp_kernel<3, 2><<<grid, blk, smem, s>>>(P);
    }
    XS_LAUNCH_CHECK();
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
