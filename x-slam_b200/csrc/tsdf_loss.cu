// tsdf_loss.cu — DCSFD volume loss: per-voxel squared TSDF residual against a ground-truth volume in
// bicomplex arithmetic, reduced to {sum loss, sum grad, sum hessian, count}.
//
// Replaces ComputeLocalTsdfHessianKernel / ComputeLocalTsdf_hessian (XKinectFusion/src/TsdfFusion.cu:204-331).
// The bicomplex number follows d_complex<float> (DeviceArray/include/cuda_double_complex.hpp:16-134,242-260)
// over the reference's own complex<T> (DeviceArray/include/cuda_complex.hpp: naive 4-product *, logb/scalbn
// scaled /, sqrt = polar(sqrt|z|, arg/2)); it is evaluated in full (not truncated) so that value(), the
// masks derived from it and the accumulated components follow the reference operation by operation.
// The reference writes four dense N^3 temporaries and runs four thrust::reduce passes over them; here each
// block reduces in double with warp shuffles and the last block combines the partials in block order.
#include "xs_common.cuh"

#include <vector>

namespace xs {

struct Cx {
    float re, im;
};
XS_DEV Cx cx(float r, float i = 0.f) { return {r, i}; }
XS_DEV Cx operator+(Cx a, Cx b) { return {a.re + b.re, a.im + b.im}; }
XS_DEV Cx operator-(Cx a, Cx b) { return {a.re - b.re, a.im - b.im}; }
XS_DEV Cx operator*(Cx a, Cx b) {  // cuda_complex.hpp:168-228
    const float ac = a.re * b.re, bd = a.im * b.im, ad = a.re * b.im, bc = a.im * b.re;
    return {ac - bd, ad + bc};
}
XS_DEV Cx operator*(Cx a, float s) { return {a.re * s, a.im * s}; }
XS_DEV Cx operator/(Cx a, float s) { return {a.re / s, a.im / s}; }
XS_DEV Cx operator/(Cx z, Cx w) {  // cuda_complex.hpp:286-300 (float specialisation, recovery code commented out)
    int il = 0;
    float c = w.re, d = w.im;
    const float lb = logbf(fmaxf(fabsf(c), fabsf(d)));
    if (isfinite(lb)) {
        il = (int) lb;
        c = scalbnf(c, -il);
        d = scalbnf(d, -il);
    }
    const float denom = c * c + d * d;
    return {scalbnf((z.re * c + z.im * d) / denom, -il), scalbnf((z.im * c - z.re * d) / denom, -il)};
}
XS_DEV Cx cx_sqrt(Cx x) {  // cuda_complex.hpp:581-593 (finite arguments)
    const float rho = sqrtf(hypotf(x.re, x.im)), theta = atan2f(x.im, x.re) / 2.f;
    float s, c;
    sincosf(theta, &s, &c);
    return {rho * c, rho * s};
}

struct BiC {  // d_complex<float>: re_ = (value, grad), im_ = (., hessian)
    Cx re, im;
};
XS_DEV BiC bic(float v) { return {cx(v), cx(0.f)}; }
XS_DEV BiC operator+(BiC a, BiC b) { return {a.re + b.re, a.im + b.im}; }
XS_DEV BiC operator-(BiC a, BiC b) { return {a.re - b.re, a.im - b.im}; }
XS_DEV BiC operator*(BiC a, BiC b) {  // cuda_double_complex.hpp:119-125
    return {a.re * b.re - a.im * b.im, a.im * b.re + a.re * b.im};
}
XS_DEV BiC operator*(BiC a, float s) { return {a.re * s, a.im * s}; }          // :79-83
XS_DEV BiC operator/(BiC a, float s) { return {a.re / s, a.im / s}; }          // :84-88
XS_DEV BiC operator+(BiC a, float s) { return {{a.re.re + s, a.re.im}, a.im}; }  // :71-74
XS_DEV BiC operator-(BiC a, float s) { return {{a.re.re - s, a.re.im}, a.im}; }
XS_DEV BiC operator/(BiC a, BiC b) {  // :126-133
    const Cx r = a.re * b.re + a.im * b.im;
    const Cx n = b.re * b.re + b.im * b.im;
    return {r / n, (a.im * b.re - a.re * b.im) / n};
}
XS_DEV Cx bic_abs(BiC x) { return cx_sqrt(x.re * x.re + x.im * x.im); }  // :233-239
XS_DEV BiC bic_sqrt(BiC x) {                                               // :242-260
    BiC result = x;
    const Cx r = bic_abs(x);
    const Cx sqrt_r = cx_sqrt(r);
    result.re = result.re + r;
    const Cx zrnorm = bic_abs(result);
    if (fabsf(zrnorm.re) < 1e-20f && fabsf(zrnorm.im) < 1e-20f) return {result.re * sqrt_r, result.im * sqrt_r};
    const Cx scale = sqrt_r / zrnorm;
    return {result.re * scale, result.im * scale};
}
struct BiC3 {
    BiC x, y, z;
};
XS_DEV BiC bdot(const BiC3 &a, const BiC3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // Internal.h:171-174
XS_DEV BiC bnorm(const BiC3 &a) { return bic_sqrt(bdot(a, a)); }                              // Internal.h:186-189

// block sums of N per-thread doubles (warp shuffles, then the 8 warps in order), per-block partials, and the last block
// to arrive adds the partials in block order: deterministic run to run
template <int N> XS_DEV void block_reduce_to_out(const double (&acc)[N], double *partials, double *out, unsigned int *ticket) {
    __shared__ double s_part[8][N];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < N; ++c) {
        double v = acc[c];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) s_part[warp][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < N) {
        double v = 0;
        for (int w = 0; w < 8; ++w) v += s_part[w][threadIdx.x];
        partials[(size_t) blockIdx.x * N + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // fixed order: warp w adds its contiguous eighth of the blocks (lane = value, independent L2 loads), then the 8 warps in order
    static_assert(N <= 32, "one lane per value");
    __shared__ double s_seg[8][N];
    {
        const int nb = (int) gridDim.x, seg = (nb + 7) / 8, b0 = warp * seg, b1 = min(nb, b0 + seg);
        double v = 0;
        if (lane < N) {
#pragma unroll 8
            for (int b = b0; b < b1; ++b) v += __ldcg(partials + (size_t) b * N + lane);
            s_seg[warp][lane] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x < N) {
        double v = 0;
        for (int w = 0; w < 8; ++w) v += s_seg[w][threadIdx.x];
        out[threadIdx.x] = v;
    }
    if (threadIdx.x == 0) *ticket = 0u;
}

struct LossParams {
    const float *depth;  // metres
    int rows, cols;
    int rx, ry, rz;
    float voxel, trunc;
    xs_intr intr;
    BiC R[9], t[3];
    const float *gt;
    double *partials;  // [grid][4]
    double *out;       // [4]
    unsigned int *ticket;
};

// One voxel of ComputeLocalTsdfHessianKernel (TsdfFusion.cu:214-281) for the bicomplex pose (R, t); false where the
// reference `continue`s.
XS_DEV bool hessian_voxel(const LossParams &P, const BiC *R, const BiC *t, float gt, size_t index, float trunc_inv, BiC &loss) {
    const int x = (int) (index % P.rx), y = (int) ((index / P.rx) % P.ry), z = (int) (index / ((size_t) P.rx * P.ry));
    const BiC3 vg = {bic((float(x) + 0.5f) * P.voxel), bic((float(y) + 0.5f) * P.voxel), bic((float(z) + 0.5f) * P.voxel)};
    const BiC3 r0 = {R[0], R[1], R[2]}, r1 = {R[3], R[4], R[5]}, r2 = {R[6], R[7], R[8]};
    const BiC3 vc = {bdot(r0, vg) + t[0], bdot(r1, vg) + t[1], bdot(r2, vg) + t[2]};
    const BiC inv_z = bic(1.0f) / vc.z;
    if (inv_z.re.re < 0) return false;
    const BiC image_x = vc.x * inv_z * P.intr.fx + P.intr.cx;  // :232-233
    const BiC image_y = vc.y * inv_z * P.intr.fy + P.intr.cy;
    const int coox = __float2int_rd(image_x.re.re - 0.5f), cooy = __float2int_rd(image_y.re.re - 0.5f);
    if (!(coox > 1 && cooy > 1 && coox < P.cols - 1 && cooy < P.rows - 1)) return false;
    const int nx = __float2int_rn(image_x.re.re), ny = __float2int_rn(image_y.re.re);
    const float d00 = __ldg(P.depth + (size_t) cooy * P.cols + coox), d10 = __ldg(P.depth + (size_t) cooy * P.cols + coox + 1);
    const float d01 = __ldg(P.depth + (size_t) (cooy + 1) * P.cols + coox),
                d11 = __ldg(P.depth + (size_t) (cooy + 1) * P.cols + coox + 1);
    BiC Dp;
    if (d00 != 0.0f && d01 != 0.0f && d10 != 0.0f && d11 != 0.0f) {  // :248-256 (no threshold test)
        const BiC one = bic(1.0f);
        const BiC a = image_x - bic(float(coox) + 0.5f), b = image_y - bic(float(cooy) + 0.5f);
        Dp = bic(d00) * (one - a) * (one - b) + bic(d10) * a * (one - b) + bic(d01) * (one - a) * b + bic(d11) * a * b;
    } else {
        Dp = bic(__ldg(P.depth + (size_t) ny * P.cols + nx));
    }
    if (Dp.re.re > 5 || Dp.re.re < 0.2) return false;  // :260 (compared in double, as the literals are)
    const BiC xl = (image_x - P.intr.cx) / P.intr.fx, yl = (image_y - P.intr.cy) / P.intr.fy;
    const BiC3 v1 = {Dp * xl, Dp * yl, Dp};
    const BiC distance = bnorm(v1) - bnorm(vc);
    const BiC gt_distance = bic(gt) * P.trunc;
    const BiC error = (distance - gt_distance) * trunc_inv;
    if (fabsf(error.re.re) > 1) return false;
    loss = error * error;
    return true;
}

__global__ void __launch_bounds__(256) tsdf_hessian_kernel(const LossParams P) {
    const size_t nvox = (size_t) P.rx * P.ry * P.rz;
    const float trunc_inv = 1.0f / P.trunc;
    double acc[4] = {0, 0, 0, 0};
    for (size_t index = (size_t) blockIdx.x * blockDim.x + threadIdx.x; index < nvox; index += (size_t) gridDim.x * blockDim.x) {
        const float gt = __ldg(P.gt + index);
        if (gt == 0 || fabsf(gt) > 0.95f) continue;  // TsdfFusion.cu:222
        BiC loss;
        if (!hessian_voxel(P, P.R, P.t, gt, index, trunc_inv, loss)) continue;
        acc[0] += (double) loss.re.re;  // value()
        acc[1] += (double) loss.re.im;  // grad()
        acc[2] += (double) loss.im.im;  // hessian()
        acc[3] += 1.0;
    }
    block_reduce_to_out<4>(acc, P.partials, P.out, P.ticket);
}

// DB bicomplex directions per sweep of the ground-truth volume: the volume (4 GiB at 1024^3) is streamed once for all of
// them instead of once per direction - the kernel is bound by that stream, since only the few per cent of voxels inside
// the ground truth's truncation band reach the bicomplex arithmetic.  Every direction runs the single-direction code on
// the same voxels in the same order, so its sums are bit-identical to a tsdf_hessian_kernel launch of its own.
template <int DB> __global__ void __launch_bounds__(256) tsdf_hessian_batch_kernel(const LossParams P, const BiC *__restrict__ poses /* [DB][12] */) {
    __shared__ BiC s_pose[DB][12];
    for (int i = threadIdx.x; i < DB * 12; i += blockDim.x) s_pose[i / 12][i % 12] = poses[i];
    __syncthreads();
    const size_t nvox = (size_t) P.rx * P.ry * P.rz;
    const float trunc_inv = 1.0f / P.trunc;
    double acc[DB * 4];
#pragma unroll
    for (int i = 0; i < DB * 4; ++i) acc[i] = 0.0;
    for (size_t index = (size_t) blockIdx.x * blockDim.x + threadIdx.x; index < nvox; index += (size_t) gridDim.x * blockDim.x) {
        const float gt = __ldg(P.gt + index);
        if (gt == 0 || fabsf(gt) > 0.95f) continue;
#pragma unroll
        for (int q = 0; q < DB; ++q) {
            BiC loss;
            if (!hessian_voxel(P, s_pose[q], s_pose[q] + 9, gt, index, trunc_inv, loss)) continue;
            acc[q * 4 + 0] += (double) loss.re.re;
            acc[q * 4 + 1] += (double) loss.re.im;
            acc[q * 4 + 2] += (double) loss.im.im;
            acc[q * 4 + 3] += 1.0;
        }
    }
    block_reduce_to_out<DB * 4>(acc, P.partials, P.out, P.ticket);
}

// ComputeLocalTsdfLossKernel (TsdfFusion.cu:335-406): the real-only twin of the kernel above - plain FP32 with the
// reference's expression shapes (so nvcc contracts the same FMAs), reduced to {sum loss, count}.
struct RealLossParams {
    const float *depth;  // metres
    int rows, cols;
    int rx, ry, rz;
    float voxel, trunc;
    xs_intr intr;
    float R[9], t[3];
    const float *gt;
    double *partials;  // [grid][2]
    double *out;       // [2]
    unsigned int *ticket;
};
XS_DEV float fdot3(float ax, float ay, float az, float bx, float by, float bz) { return ax * bx + ay * by + az * bz; }  // Internal.h:215-218

__global__ void __launch_bounds__(256) tsdf_loss_kernel(const RealLossParams P) {
    const size_t nvox = (size_t) P.rx * P.ry * P.rz;
    const float tranc_dist_inv = 1.0f / P.trunc;
    double acc[2] = {0, 0};
    for (size_t index = (size_t) blockIdx.x * blockDim.x + threadIdx.x; index < nvox; index += (size_t) gridDim.x * blockDim.x) {
        const float gt_tsdf = __ldg(P.gt + index);
        if (gt_tsdf == 0 || fabs(gt_tsdf) > 0.95) continue;  // :352 (compared in double, as the literal is)
        const int x = (int) (index % P.rx), y = (int) ((index / P.rx) % P.ry), z = (int) (index / ((size_t) P.rx * P.ry));
        const float v_g_x = (float(x) + 0.5f) * P.voxel, v_g_y = (float(y) + 0.5f) * P.voxel, v_g_z = (float(z) + 0.5f) * P.voxel;
        const float v_c_x = fdot3(P.R[0], P.R[1], P.R[2], v_g_x, v_g_y, v_g_z) + P.t[0];  // Rv2c * v_g + tv2c, :358
        const float v_c_y = fdot3(P.R[3], P.R[4], P.R[5], v_g_x, v_g_y, v_g_z) + P.t[1];
        const float v_c_z = fdot3(P.R[6], P.R[7], P.R[8], v_g_x, v_g_y, v_g_z) + P.t[2];
        const float inv_z = 1.0f / v_c_z;
        if (inv_z < 0) continue;
        const float image_x = v_c_x * inv_z * P.intr.fx + P.intr.cx;
        const float image_y = v_c_y * inv_z * P.intr.fy + P.intr.cy;
        const int coox = __float2int_rd(image_x - 0.5f), cooy = __float2int_rd(image_y - 0.5f);
        if (!(coox > 1 && cooy > 1 && coox < P.cols - 1 && cooy < P.rows - 1)) continue;
        const int nx = __float2int_rn(image_x), ny = __float2int_rn(image_y);
        const float d00 = __ldg(P.depth + (size_t) cooy * P.cols + coox), d10 = __ldg(P.depth + (size_t) cooy * P.cols + coox + 1);
        const float d01 = __ldg(P.depth + (size_t) (cooy + 1) * P.cols + coox),
                    d11 = __ldg(P.depth + (size_t) (cooy + 1) * P.cols + coox + 1);
        float Dp;
        if (d00 != 0.0f && d01 != 0.0f && d10 != 0.0f && d11 != 0.0f) {  // :378-386
            const float one = 1.0f;
            const float a = image_x - (float(coox) + 0.5f);
            const float b = image_y - (float(cooy) + 0.5f);
            Dp = d00 * (one - a) * (one - b) + d10 * a * (one - b) + d01 * (one - a) * b + d11 * a * b;
        } else {
            Dp = __ldg(P.depth + (size_t) ny * P.cols + nx);
        }
        if (Dp > 5 || Dp < 0.2) continue;  // :390
        const float xl = (image_x - P.intr.cx) / P.intr.fx;
        const float yl = (image_y - P.intr.cy) / P.intr.fy;
        const float v1x = Dp * xl, v1y = Dp * yl, v1z = Dp;
        const float distance = sqrt(fdot3(v1x, v1y, v1z, v1x, v1y, v1z)) - sqrt(fdot3(v_c_x, v_c_y, v_c_z, v_c_x, v_c_y, v_c_z));
        const float gt_distance = gt_tsdf * P.trunc;
        const float error = (distance - gt_distance) * tranc_dist_inv;
        if (fabs(error) > 1) continue;
        const float loss = error * error;
        acc[0] += (double) loss;
        acc[1] += 1.0;
    }
    block_reduce_to_out<2>(acc, P.partials, P.out, P.ticket);
}

__global__ void scale_depth_kernel2(const uint16_t *__restrict__ depth, size_t step, int rows, int cols, float *__restrict__ out) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    int Dp = *((const uint16_t *) ((const char *) depth + (size_t) y * step) + x);
    out[(size_t) y * cols + x] = (Dp > 5000 || Dp < 200) ? 0.f : __fdiv_rn(float(Dp), 1000.f);
}

}  // namespace xs

using namespace xs;

extern "C" int xs_tsdf_hessian(const uint16_t *d_depth, size_t depth_step_bytes, int rows, int cols, xs_intr intr,
                               const int res[3], float voxel_size, const xs_pose *v2c, float trunc, const float *d_gt,
                               double *out4_host, void *stream) {
    if (!d_depth || !res || !v2c || !d_gt || !out4_host || v2c->ncomp != 3 || !v2c->dR || !v2c->dt) {
        set_error("xs_tsdf_hessian: needs one bicomplex direction (ncomp == 3: eps1, eps2, eps1eps2)");
        return XS_ERR_ARG;
    }
    cudaStream_t s = (cudaStream_t) stream;
    const int grid = sm_count() * 8;
    float *d_scaled = nullptr;
    double *d_part = nullptr;
    unsigned int *d_ticket = nullptr;
    XS_CUDA(cudaMalloc(&d_scaled, (size_t) rows * cols * sizeof(float)));
    XS_CUDA(cudaMalloc(&d_part, (size_t) (grid + 1) * 4 * sizeof(double)));
    XS_CUDA(cudaMalloc(&d_ticket, sizeof(unsigned int)));
    XS_CUDA(cudaMemsetAsync(d_ticket, 0, sizeof(unsigned int), s));
    dim3 blk(32, 8), grd(div_up(cols, 32), div_up(rows, 8));
    scale_depth_kernel2<<<grd, blk, 0, s>>>(d_depth, depth_step_bytes, rows, cols, d_scaled);
    XS_LAUNCH_CHECK();
    LossParams P;
    P.depth = d_scaled;
    P.rows = rows;
    P.cols = cols;
    P.rx = res[0];
    P.ry = res[1];
    P.rz = res[2];
    P.voxel = voxel_size;
    P.trunc = trunc;
    P.intr = intr;
    for (int i = 0; i < 9; ++i) P.R[i] = {{v2c->R[i], v2c->dR[i]}, {v2c->dR[9 + i], v2c->dR[18 + i]}};
    for (int i = 0; i < 3; ++i) P.t[i] = {{v2c->t[i], v2c->dt[i]}, {v2c->dt[3 + i], v2c->dt[6 + i]}};
    P.gt = d_gt;
    P.partials = d_part;
    P.out = d_part + (size_t) grid * 4;
    P.ticket = d_ticket;
    tsdf_hessian_kernel<<<grid, 256, 0, s>>>(P);
    XS_LAUNCH_CHECK();
    XS_CUDA(cudaMemcpyAsync(out4_host, P.out, 4 * sizeof(double), cudaMemcpyDeviceToHost, s));
    XS_CUDA(cudaStreamSynchronize(s));
    cudaFree(d_scaled);
    cudaFree(d_part);
    cudaFree(d_ticket);
    return XS_OK;
}

extern "C" int xs_tsdf_loss(const uint16_t *d_depth, size_t depth_step_bytes, int rows, int cols, xs_intr intr, const int res[3],
                            float voxel_size, const float Rv2c[9], const float tv2c[3], float trunc, const float *d_gt,
                            double *out2_host, void *stream) {
    if (!d_depth || !res || !Rv2c || !tv2c || !d_gt || !out2_host || rows <= 0 || cols <= 0) {
        set_error("xs_tsdf_loss: null argument");
        return XS_ERR_ARG;
    }
    cudaStream_t s = (cudaStream_t) stream;
    const int grid = sm_count() * 8;
    float *d_scaled = nullptr;
    double *d_part = nullptr;
    unsigned int *d_ticket = nullptr;
    XS_CUDA(cudaMalloc(&d_scaled, (size_t) rows * cols * sizeof(float)));
    XS_CUDA(cudaMalloc(&d_part, (size_t) (grid + 1) * 2 * sizeof(double)));
    XS_CUDA(cudaMalloc(&d_ticket, sizeof(unsigned int)));
    XS_CUDA(cudaMemsetAsync(d_ticket, 0, sizeof(unsigned int), s));
    dim3 blk(32, 8), grd(div_up(cols, 32), div_up(rows, 8));
    scale_depth_kernel2<<<grd, blk, 0, s>>>(d_depth, depth_step_bytes, rows, cols, d_scaled);
    XS_LAUNCH_CHECK();
    RealLossParams P;
    P.depth = d_scaled;
    P.rows = rows;
    P.cols = cols;
    P.rx = res[0];
    P.ry = res[1];
    P.rz = res[2];
    P.voxel = voxel_size;
    P.trunc = trunc;
    P.intr = intr;
    for (int i = 0; i < 9; ++i) P.R[i] = Rv2c[i];
    for (int i = 0; i < 3; ++i) P.t[i] = tv2c[i];
    P.gt = d_gt;
    P.partials = d_part;
    P.out = d_part + (size_t) grid * 2;
    P.ticket = d_ticket;
    tsdf_loss_kernel<<<grid, 256, 0, s>>>(P);
    XS_LAUNCH_CHECK();
    XS_CUDA(cudaMemcpyAsync(out2_host, P.out, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    XS_CUDA(cudaStreamSynchronize(s));
    cudaFree(d_scaled);
    cudaFree(d_part);
    cudaFree(d_ticket);
    return XS_OK;
}

// Batched ComputeLocalTsdf_hessian: v2c carries ncomp = 3 * dirs components (eps1, eps2, eps1eps2 per direction); out_host is
// [dirs][4] = {sum loss, sum grad, sum hessian, count} per direction.  Directions are processed 8 (then 4, 2, 1) per sweep of the
// ground-truth volume.
extern "C" int xs_tsdf_hessian_batch(const uint16_t *d_depth, size_t depth_step_bytes, int rows, int cols, xs_intr intr,
                                     const int res[3], float voxel_size, const xs_pose *v2c, float trunc, const float *d_gt,
                                     double *out_host, void *stream) {
    if (!d_depth || !res || !v2c || !d_gt || !out_host || v2c->ncomp <= 0 || v2c->ncomp % 3 || !v2c->dR || !v2c->dt) {
        set_error("xs_tsdf_hessian_batch: needs ncomp = 3 * dirs components (eps1, eps2, eps1eps2 per direction)");
        return XS_ERR_ARG;
    }
    const int dirs = v2c->ncomp / 3;
    cudaStream_t s = (cudaStream_t) stream;
    const int grid = sm_count() * 8;
    float *d_scaled = nullptr;
    double *d_part = nullptr;
    unsigned int *d_ticket = nullptr;
    BiC *d_poses = nullptr;
    std::vector<BiC> h_poses((size_t) dirs * 12);
    for (int q = 0; q < dirs; ++q) {
        for (int i = 0; i < 9; ++i)
            h_poses[(size_t) q * 12 + i] = {{v2c->R[i], v2c->dR[(3 * q) * 9 + i]}, {v2c->dR[(3 * q + 1) * 9 + i], v2c->dR[(3 * q + 2) * 9 + i]}};
        for (int i = 0; i < 3; ++i)
            h_poses[(size_t) q * 12 + 9 + i] = {{v2c->t[i], v2c->dt[(3 * q) * 3 + i]}, {v2c->dt[(3 * q + 1) * 3 + i], v2c->dt[(3 * q + 2) * 3 + i]}};
    }
    XS_CUDA(cudaMalloc(&d_scaled, (size_t) rows * cols * sizeof(float)));
    XS_CUDA(cudaMalloc(&d_part, ((size_t) grid * 32 + (size_t) dirs * 4) * sizeof(double)));
    XS_CUDA(cudaMalloc(&d_ticket, sizeof(unsigned int)));
    XS_CUDA(cudaMalloc(&d_poses, h_poses.size() * sizeof(BiC)));
    XS_CUDA(cudaMemsetAsync(d_ticket, 0, sizeof(unsigned int), s));
    XS_CUDA(cudaMemcpyAsync(d_poses, h_poses.data(), h_poses.size() * sizeof(BiC), cudaMemcpyHostToDevice, s));
    dim3 blk(32, 8), grd(div_up(cols, 32), div_up(rows, 8));
    scale_depth_kernel2<<<grd, blk, 0, s>>>(d_depth, depth_step_bytes, rows, cols, d_scaled);
    XS_LAUNCH_CHECK();
    LossParams P;
    P.depth = d_scaled;
    P.rows = rows;
    P.cols = cols;
    P.rx = res[0];
    P.ry = res[1];
    P.rz = res[2];
    P.voxel = voxel_size;
    P.trunc = trunc;
    P.intr = intr;
    P.gt = d_gt;
    P.partials = d_part;
    P.ticket = d_ticket;
    double *d_out = d_part + (size_t) grid * 32;
    for (int q = 0; q < dirs;) {  // stream-ordered launches share the partial buffer and the (self-resetting) ticket
        P.out = d_out + (size_t) q * 4;
        const BiC *pq = d_poses + (size_t) q * 12;
        if (dirs - q >= 8) {
            tsdf_hessian_batch_kernel<8><<<grid, 256, 0, s>>>(P, pq);
            q += 8;
        } else if (dirs - q >= 4) {
            tsdf_hessian_batch_kernel<4><<<grid, 256, 0, s>>>(P, pq);
            q += 4;
        } else if (dirs - q >= 2) {
            tsdf_hessian_batch_kernel<2><<<grid, 256, 0, s>>>(P, pq);
            q += 2;
        } else {
            tsdf_hessian_batch_kernel<1><<<grid, 256, 0, s>>>(P, pq);
            q += 1;
        }
        XS_LAUNCH_CHECK();
    }
    XS_CUDA(cudaMemcpyAsync(out_host, d_out, (size_t) dirs * 4 * sizeof(double), cudaMemcpyDeviceToHost, s));
    XS_CUDA(cudaStreamSynchronize(s));
    cudaFree(d_scaled);
    cudaFree(d_part);
    cudaFree(d_ticket);
    cudaFree(d_poses);
    return XS_OK;
}
