// surface.cu — surface measurement: bilateral filter, depth pyramid, vertex and normal maps.
//
// Replaces bilateralKernel / pyrDownKernel / computeVmapKernel / computeNmapKernel and their wrappers
// (XKinectFusion/src/Map.cu:8-102,155-230,262-283).  With real depth and real intrinsics the reference
// writes devComplex(x, 0) everywhere on this stage (Map.cu:196-198,228-229), so these maps carry the
// real plane only: float[rows][cols] depth levels and float[3][rows][cols] vertex / normal maps.
// Integer rounding (bilateral -> int mm, pyrDown integer mean) follows the reference operation by
// operation, including the exclusive, clamped window bounds (Map.cu:172-173,213-214).
#include "xs_common.cuh"

namespace xs {

__global__ void __launch_bounds__(256)
bilateral_kernel(const uint16_t *__restrict__ src, size_t step, int rows, int cols, float *__restrict__ dst,
                 float sigma_space2_inv_half, float sigma_color2_inv_half) {
    const int x = threadIdx.x + blockIdx.x * blockDim.x;
    const int y = threadIdx.y + blockIdx.y * blockDim.y;
    if (x >= cols || y >= rows) return;
    auto at = [&](int yy, int xx) -> int { return *((const uint16_t *) ((const char *) src + (size_t) yy * step) + xx); };
    const int value = at(y, x);
    const int R = 6, D = R * 2 + 1;
    const int tx = min(x - D / 2 + D, cols - 1);
    const int ty = min(y - D / 2 + D, rows - 1);
    float sum1 = 0, sum2 = 0;
    for (int cy = max(y - D / 2, 0); cy < ty; ++cy) {
        for (int cx = max(x - D / 2, 0); cx < tx; ++cx) {
            const int tmp = at(cy, cx);
            const float space2 = (float) (unsigned) ((x - cx) * (x - cx) + (y - cy) * (y - cy));
            const float color2 = (float) (unsigned) ((value - tmp) * (value - tmp));
            // Map.cu:185-189: fma(space2, s, color2*c) -> __expf -> fma accumulate
            const float weight = __expf(-__fmaf_rn(space2, sigma_space2_inv_half, __fmul_rn(color2, sigma_color2_inv_half)));
            sum1 = __fmaf_rn(weight, (float) tmp, sum1);
            sum2 = __fadd_rn(sum2, weight);
        }
    }
    int round = __float2int_rn(__fdiv_rn(sum1, sum2));
    if (round > 5000 || round < 200) round = 0;
    round = max(0, min(round, 32767));
    dst[(size_t) y * cols + x] = __int2float_rd(round);
}

__global__ void pyr_down_kernel(const float *__restrict__ src, int srows, int scols, float *__restrict__ dst, int drows,
                                int dcols, float sigma_color) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dcols || y >= drows) return;
    const int D = 5;
    const int center = __float2int_rn(src[(size_t) (2 * y) * scols + 2 * x]);
    const int tx = min(2 * x - D / 2 + D, scols - 1);
    const int ty = min(2 * y - D / 2 + D, srows - 1);
    int sum = 0, count = 0;
    for (int cy = max(0, 2 * y - D / 2); cy < ty; ++cy)
        for (int cx = max(0, 2 * x - D / 2); cx < tx; ++cx) {
            const int val = __float2int_rn(src[(size_t) cy * scols + cx]);
            if (abs(val - center) < 3 * sigma_color) {
                sum += val;
                ++count;
            }
        }
    dst[(size_t) y * dcols + x] = __int2float_rd(sum / count);
}

__global__ void vmap_kernel(const float *__restrict__ depth, int rows, int cols, float *__restrict__ vmap, float fx_inv,
                            float fy_inv, float cx, float cy) {
    const int u = threadIdx.x + blockIdx.x * blockDim.x;
    const int v = threadIdx.y + blockIdx.y * blockDim.y;
    if (u >= cols || v >= rows) return;
    const size_t plane = (size_t) rows * cols, pix = (size_t) v * cols + u;
    const float z = __fdiv_rn(depth[pix], 1000.f);  // Map.cu:16
    if (z != 0) {
        vmap[pix] = __fmul_rn(__fmul_rn(z, __fsub_rn(float(u), cx)), fx_inv);
        vmap[pix + plane] = __fmul_rn(__fmul_rn(z, __fsub_rn(float(v), cy)), fy_inv);
        vmap[pix + 2 * plane] = z;
    } else {
        vmap[pix] = __int_as_float(0x7fffffff);  // NaN in the x plane only (Map.cu:27); y,z made deterministic
        vmap[pix + plane] = 0.f;
        vmap[pix + 2 * plane] = 0.f;
    }
}

__global__ void nmap_kernel(int rows, int cols, const float *__restrict__ vmap, float *__restrict__ nmap) {
    const int u = threadIdx.x + blockIdx.x * blockDim.x;
    const int v = threadIdx.y + blockIdx.y * blockDim.y;
    if (u >= cols || v >= rows) return;
    const size_t plane = (size_t) rows * cols, pix = (size_t) v * cols + u;
    const float qnan = __int_as_float(0x7fffffff);
    bool ok = !(u == cols - 1 || v == rows - 1);
    typedef Jet<1, 0> J;
    Jet3<1, 0> v00, v01, v10;
    if (ok) {
        v00.x.v = vmap[pix];
        v01.x.v = vmap[pix + 1];
        v10.x.v = vmap[pix + cols];
        ok = !isnan(v00.x.v) && !isnan(v01.x.v) && !isnan(v10.x.v);
    }
    if (!ok) {
        nmap[pix] = qnan;
        nmap[pix + plane] = 0.f;
        nmap[pix + 2 * plane] = 0.f;
        return;
    }
    v00.y.v = vmap[pix + plane];
    v01.y.v = vmap[pix + plane + 1];
    v10.y.v = vmap[pix + plane + cols];
    v00.z.v = vmap[pix + 2 * plane];
    v01.z.v = vmap[pix + 2 * plane + 1];
    v10.z.v = vmap[pix + 2 * plane + cols];
    // Map.cu:60: normalized(cross(v01 - v00, v10 - v00)) in complex arithmetic with zero imaginary parts
    const Jet3<1, 0> r = jnormalized(jcross(v01 - v00, v10 - v00));
    nmap[pix] = r.x.v;
    nmap[pix + plane] = r.y.v;
    nmap[pix + 2 * plane] = r.z.v;
}

// ---- derivative components of the current-frame maps (intrinsic parameters of a Hessian batch, xs_batch.h) ---------------
// createVMap with a perturbed intrinsic: vx = z (u - cx) g, g = 1 / fx (Map.cu:19-21, 84).  For parameter p with seeds
// (dfx, dcx):  F_p(vx) = z [ -dcx g - (u - cx) g^2 dfx ];  for a pair (i, j) of such parameters
// S_ij(vx) = z [ (dcx_i dfx_j + dcx_j dfx_i) g^2 + 2 (u - cx) g^3 dfx_i dfx_j ];  vy likewise with (fy, cy); vz = z is constant.
// vmap: [(1 + ncurr)][3][rows][cols]; thread = pixel, loop over the slots.  `scale` = 1 / 2^level (Intr::operator(), Internal.h:55-58).
__global__ void vmap_deriv_kernel(const float *__restrict__ depth, int rows, int cols, float *__restrict__ vmap, xs_intr k, float scale,
                                  BatchView B) {
    const int u = threadIdx.x + blockIdx.x * blockDim.x;
    const int v = threadIdx.y + blockIdx.y * blockDim.y;
    if (u >= cols || v >= rows) return;
    const size_t plane = (size_t) rows * cols, pix = (size_t) v * cols + u;
    const float z = __fdiv_rn(depth[pix], 1000.f);
    const float gx = 1.f / k.fx, gy = 1.f / k.fy, ax = float(u) - k.cx, ay = float(v) - k.cy;
    for (int c = 0; c < B.ncomp; ++c) {
        const int slot = __ldg(B.cslot + c);
        if (slot < 0) continue;
        float dx = 0.f, dy = 0.f;
        if (z != 0.f) {
            if (c < B.n) {
                const float4 d = __ldg(reinterpret_cast<const float4 *>(B.dintr) + c);  // (dfx, dfy, dcx, dcy) at level 0
                dx = z * (-d.z * scale * gx - ax * gx * gx * d.x * scale);
                dy = z * (-d.w * scale * gy - ay * gy * gy * d.y * scale);
            } else {
                const int2 pr = __ldg(B.pairs + (c - B.n));
                const float4 di = __ldg(reinterpret_cast<const float4 *>(B.dintr) + pr.x), dj = __ldg(reinterpret_cast<const float4 *>(B.dintr) + pr.y);
                const float s2 = scale * scale;
                dx = z * s2 * ((di.z * dj.x + dj.z * di.x) * gx * gx + 2.f * ax * gx * gx * gx * di.x * dj.x);
                dy = z * s2 * ((di.w * dj.y + dj.w * di.y) * gy * gy + 2.f * ay * gy * gy * gy * di.y * dj.y);
            }
        }
        float *o = vmap + (size_t) (1 + slot) * 3 * plane + pix;
        o[0] = dx;
        o[plane] = dy;
        o[2 * plane] = 0.f;
    }
}

// createNMap on jets: n = normalized(cross(v01 - v00, v10 - v00)) (Map.cu:32-70) for the derivative slots; first-order
// slots in the dual algebra, pair slots in the bicomplex one on (F_i, F_j, S_ij).
template <int C>
XS_DEV void nmap_deriv_task(const float *__restrict__ vmap, float *__restrict__ nmap, size_t plane, size_t pix, int cols, bool ok,
                            const int (&slot)[C], int out_slot) {
    float *o = nmap + (size_t) (1 + out_slot) * 3 * plane + pix;
    if (!ok) {
        o[0] = o[plane] = o[2 * plane] = 0.f;
        return;
    }
    Jet3<C, 1> p[3];  // v00, v01, v10
    const size_t off[3] = {0, 1, (size_t) cols};
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        Jet<C, 1> *cmp[3] = {&p[t].x, &p[t].y, &p[t].z};
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) {
            cmp[pl]->v = vmap[pix + off[t] + pl * plane];
#pragma unroll
            for (int c = 0; c < C; ++c)
                cmp[pl]->d[c] = slot[c] >= 0 ? vmap[(size_t) (1 + slot[c]) * 3 * plane + pix + off[t] + pl * plane] : 0.f;
        }
    }
    const Jet3<C, 1> r = jnormalized_fast(jcross(p[1] - p[0], p[2] - p[0]));
    o[0] = r.x.d[C - 1];
    o[plane] = r.y.d[C - 1];
    o[2 * plane] = r.z.d[C - 1];
}
__global__ void nmap_deriv_kernel(int rows, int cols, const float *__restrict__ vmap, float *__restrict__ nmap, BatchView B) {
    const int u = threadIdx.x + blockIdx.x * blockDim.x;
    const int v = threadIdx.y + blockIdx.y * blockDim.y;
    if (u >= cols || v >= rows) return;
    const size_t plane = (size_t) rows * cols, pix = (size_t) v * cols + u;
    bool ok = !(u == cols - 1 || v == rows - 1);
    if (ok) ok = !isnan(vmap[pix]) && !isnan(vmap[pix + 1]) && !isnan(vmap[pix + cols]);
    for (int c = 0; c < B.ncomp; ++c) {
        const int slot = __ldg(B.cslot + c);
        if (slot < 0) continue;
        if (c < B.n) {
            const int s1[1] = {slot};
            nmap_deriv_task<1>(vmap, nmap, plane, pix, cols, ok, s1, slot);
        } else {
            const int2 pr = __ldg(B.pairs + (c - B.n));
            const int s3[3] = {__ldg(B.cslot + pr.x), __ldg(B.cslot + pr.y), slot};
            nmap_deriv_task<3>(vmap, nmap, plane, pix, cols, ok, s3, slot);
        }
    }
}

// derivative components of the vertex / normal maps of one pyramid level (after the real maps): [(1 + ncurr)][3][rows][cols]
int surface_derivs(const float *d_depth, int rows, int cols, int level, xs_intr intr_level0, const BatchView &B, float *d_vmap, float *d_nmap,
                   cudaStream_t s) {
    if (B.ncurr == 0) return XS_OK;
    const float scale = 1.f / float(1 << level);
    const xs_intr k = {intr_level0.fx * scale, intr_level0.fy * scale, intr_level0.cx * scale, intr_level0.cy * scale};
    dim3 blk(32, 8), grd(div_up(cols, 32), div_up(rows, 8));
    vmap_deriv_kernel<<<grd, blk, 0, s>>>(d_depth, rows, cols, d_vmap, k, scale, B);
    XS_LAUNCH_CHECK();
    nmap_deriv_kernel<<<grd, blk, 0, s>>>(rows, cols, d_vmap, d_nmap, B);
    XS_LAUNCH_CHECK();
    return XS_OK;
}

// Seam layout <-> packed SoA.  The reference's maps are pitched arrays of interleaved complex<float> with the planes
// stacked by rows (MapArr = DeviceArray2D<devComplex>, Internal.h:31; `nplanes*rows` x cols).  Component 0 of the SoA
// map is the real part; the imaginary part maps to derivative component `comp` (comp < 0: dropped / written as zero).
__global__ void complex_to_soa_kernel(const char *__restrict__ src, size_t step, int prow, int cols, float *__restrict__ real,
                                      float *__restrict__ imag) {
    const int x = threadIdx.x + blockIdx.x * blockDim.x, y = threadIdx.y + blockIdx.y * blockDim.y;
    if (x >= cols || y >= prow) return;
    const float2 v = reinterpret_cast<const float2 *>(src + (size_t) y * step)[x];
    real[(size_t) y * cols + x] = v.x;
    if (imag) imag[(size_t) y * cols + x] = v.y;
}
__global__ void soa_to_complex_kernel(const float *__restrict__ real, const float *__restrict__ imag, int prow, int cols,
                                      char *__restrict__ dst, size_t step) {
    const int x = threadIdx.x + blockIdx.x * blockDim.x, y = threadIdx.y + blockIdx.y * blockDim.y;
    if (x >= cols || y >= prow) return;
    reinterpret_cast<float2 *>(dst + (size_t) y * step)[x] =
        make_float2(real[(size_t) y * cols + x], imag ? imag[(size_t) y * cols + x] : 0.f);
}

}  // namespace xs

using namespace xs;

extern "C" {

int xs_map_complex_to_soa(const void *d_src, size_t step_bytes, int nplanes, int rows, int cols, float *d_soa, int ncomp,
                          int comp, void *stream) {
    if (!d_src || !d_soa || nplanes <= 0 || rows <= 0 || cols <= 0 || comp >= ncomp) return XS_ERR_ARG;
    const size_t csize = (size_t) nplanes * rows * cols;
    dim3 blk(32, 8), grd(div_up(cols, 32), div_up(nplanes * rows, 8));
    complex_to_soa_kernel<<<grd, blk, 0, (cudaStream_t) stream>>>((const char *) d_src, step_bytes, nplanes * rows, cols, d_soa,
                                                                  comp >= 0 ? d_soa + (size_t) (1 + comp) * csize : nullptr);
    XS_LAUNCH_CHECK();
    return XS_OK;
}

int xs_map_soa_to_complex(const float *d_soa, int ncomp, int comp, int nplanes, int rows, int cols, void *d_dst,
                          size_t step_bytes, void *stream) {
    if (!d_dst || !d_soa || nplanes <= 0 || rows <= 0 || cols <= 0 || comp >= ncomp) return XS_ERR_ARG;
    const size_t csize = (size_t) nplanes * rows * cols;
    dim3 blk(32, 8), grd(div_up(cols, 32), div_up(nplanes * rows, 8));
    soa_to_complex_kernel<<<grd, blk, 0, (cudaStream_t) stream>>>(d_soa, comp >= 0 ? d_soa + (size_t) (1 + comp) * csize : nullptr,
                                                                  nplanes * rows, cols, (char *) d_dst, step_bytes);
    XS_LAUNCH_CHECK();
    return XS_OK;
}

int xs_bilateral_filter(const uint16_t *d_depth, size_t depth_step_bytes, int rows, int cols, float *d_out, void *stream) {
    if (!d_depth || !d_out || rows <= 0 || cols <= 0) return XS_ERR_ARG;
    const float sigma_color = 30, sigma_space = 4.5;  // Map.cu:4-5
    dim3 blk(32, 8), grd(div_up(cols, 32), div_up(rows, 8));
    bilateral_kernel<<<grd, blk, 0, (cudaStream_t) stream>>>(d_depth, depth_step_bytes, rows, cols, d_out,
                                                             0.5f / (sigma_space * sigma_space),
                                                             0.5f / (sigma_color * sigma_color));
    XS_LAUNCH_CHECK();
    return XS_OK;
}

int xs_pyr_down(const float *d_src, int rows, int cols, float *d_dst, void *stream) {
    if (!d_src || !d_dst || rows < 2 || cols < 2) return XS_ERR_ARG;
    const float sigma_color = 30;
    const int drows = rows / 2, dcols = cols / 2;
    dim3 blk(32, 8), grd(div_up(dcols, 32), div_up(drows, 8));
    pyr_down_kernel<<<grd, blk, 0, (cudaStream_t) stream>>>(d_src, rows, cols, d_dst, drows, dcols, sigma_color);
    XS_LAUNCH_CHECK();
    return XS_OK;
}

int xs_create_vmap(xs_intr intr, const float *d_depth, int rows, int cols, float *d_vmap, void *stream) {
    if (!d_depth || !d_vmap || rows <= 0 || cols <= 0) return XS_ERR_ARG;
    dim3 blk(32, 8), grd(div_up(cols, 32), div_up(rows, 8));
    vmap_kernel<<<grd, blk, 0, (cudaStream_t) stream>>>(d_depth, rows, cols, d_vmap, 1.f / intr.fx, 1.f / intr.fy, intr.cx,
                                                        intr.cy);
    XS_LAUNCH_CHECK();
    return XS_OK;
}

int xs_create_nmap(const float *d_vmap, int rows, int cols, float *d_nmap, void *stream) {
    if (!d_vmap || !d_nmap || rows <= 0 || cols <= 0) return XS_ERR_ARG;
    dim3 blk(32, 8), grd(div_up(cols, 32), div_up(rows, 8));
    nmap_kernel<<<grd, blk, 0, (cudaStream_t) stream>>>(rows, cols, d_vmap, d_nmap);
    XS_LAUNCH_CHECK();
    return XS_OK;
}

}  // extern "C"
