// integrate.cu — TSDF volume container and brick-tiled, direction-batched TSDF integration.
//
// Replaces TsdfVolume (XKinectFusion/src/TsdfVolume.cpp:11-77), initVolume (TsdfFusion.cu:4-43),
// scaleDepthKernal (TsdfFusion.cu:68-82), tsdfFusionKernal / integrateTsdfVolume (TsdfFusion.cu:85-201)
// and pack/unpack_tsdf (TsdfFusion.h:7-26).
//
// One CTA pass = half of an 8x8x8 brick (256 threads, thread = voxel, brick-local index = thread index, so every
// warp access to value / weight / deriv[comp] is one fully coalesced 128-byte line).  The real path of a
// voxel (projection, depth lookup, sdf, all integer decisions) is evaluated ONCE together with the
// Jacobian (and for DCSFD the Hessian) of sdf with respect to the camera-frame position v_c; every stored
// derivative component q is then updated from its own pose-derivative (dR_q, dt_q):
//      d v_c,q = dR_q * v_g + dt_q,        d sdf_q = J . d v_c,q   (+ d1 v_c^T H d2 v_c for eps1eps2)
// so one pass over the brick serves all k perturbation directions (the reference re-runs the whole
// kernel once per direction).  Bricks that cannot project into the image are culled before any load.
#include "xs_common.cuh"

#include <cstring>

namespace xs {

__global__ void scale_depth_kernel(const uint16_t *__restrict__ depth, size_t step, int rows, int cols,
                                   float *__restrict__ out) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    int Dp = *((const uint16_t *) ((const char *) depth + (size_t) y * step) + x);
    // TsdfFusion.cu:76-81
    out[(size_t) y * cols + x] = (Dp > 5000 || Dp < 200) ? 0.f : __fdiv_rn(float(Dp), 1000.f);
}

__global__ void reset_volume_kernel(float *value, int *weight, float *deriv, size_t nvox, size_t nderiv) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t) gridDim.x * blockDim.x;
    for (size_t j = i; j < nvox; j += stride) {
        value[j] = 0.f;
        weight[j] = 0;
    }
    for (size_t j = i; j < nderiv; j += stride) deriv[j] = 0.f;
}

struct IntegrateParams {
    VolumeView V;
    DevPose v2c;
    const float *dpose;  // [ncomp][12]
    const float *depth;  // metres, [rows][cols]
    int rows, cols;
    xs_intr intr;
    int max_weight;
    float threshold;
    float trunc_inv;
    unsigned long long *stats;  // [0] updated voxels, [1] bricks in the list, [2] voxels whose derivative planes were read+written
    int nbricks;
    const int2 *brick_list;     // bricks that survive the cull (written by cull_bricks_kernel): (index, packed x|y<<10|z<<20)
    unsigned int *list_count;
    unsigned char *live;        // [nbricks] 0 = every derivative plane of the brick is still exactly zero
    const float *tile_max;      // [tiles_y][tiles_x] largest valid depth (metres) of each 16 x 16 pixel tile, 0 = none
    int tiles_x, tiles_y;
    BatchView batch;            // meaning of the ncomp derivative planes (kind 2: first-order planes, then one plane per pair)
};

// Largest depth of every 16 x 16 pixel tile (invalid pixels are 0), for the depth-aware part of the brick cull.
constexpr int CULL_TILE = 16;
__global__ void __launch_bounds__(256) depth_tile_max_kernel(const float *__restrict__ depth, int rows, int cols, float *__restrict__ tile_max,
                                                             int tiles_x) {
    const int x = blockIdx.x * CULL_TILE + (threadIdx.x & 15), y = blockIdx.y * CULL_TILE + (threadIdx.x >> 4);
    float v = (x < cols && y < rows) ? depth[(size_t) y * cols + x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    __shared__ float s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) v = fmaxf(v, s[w]);
        tile_max[blockIdx.y * tiles_x + blockIdx.x] = v;
    }
}

// Conservative brick cull: bounding sphere vs. camera half-space / image planes, then against the depth image: a brick
// whose nearest point lies farther than the largest depth under its (padded) footprint plus the truncation distance holds
// only voxels the update rule skips (sdf < -trunc, TsdfFusion.cu:150) - in a room that is every brick behind the walls.  One thread per brick; survivors are
// appended to the brick list (warp-aggregated), so the integration kernel never touches a brick outside the frustum.
__global__ void __launch_bounds__(256) cull_bricks_kernel(const IntegrateParams P, int2 *__restrict__ list) {
    const int b = blockIdx.x * 256 + threadIdx.x;
    bool keep = false;
    int bx = 0, by = 0, bz = 0;
    if (b < P.nbricks) {
        const float vs = P.V.voxel;
        const float *R = P.v2c.R, *t = P.v2c.t;
        bx = b % P.V.bx, by = (b / P.V.bx) % P.V.by, bz = b / (P.V.bx * P.V.by);
        const float cxw = (bx * 8 + 4) * vs, cyw = (by * 8 + 4) * vs, czw = (bz * 8 + 4) * vs;
        const float ccx = R[0] * cxw + R[1] * cyw + R[2] * czw + t[0];
        const float ccy = R[3] * cxw + R[4] * cyw + R[5] * czw + t[1];
        const float ccz = R[6] * cxw + R[7] * cyw + R[8] * czw + t[2];
        const float rad = 6.4f * vs + 1e-4f;  // > half diagonal 3.5*sqrt(3) = 6.06 voxels
        bool cull = false;
        if (ccz + rad < 0.f) {
            cull = true;  // every voxel has z < 0  =>  Re(1/z) < 0
        } else if (ccz - rad > 1e-3f) {
            const float fx = P.intr.fx, fy = P.intr.fy, pcx = P.intr.cx, pcy = P.intr.cy;
            // ix >= lo  <=>  fx*X - (lo-cx)*Z >= 0 ; a voxel needs 2 <= ix < cols and 2 <= iy < rows
            float nz, nn;
            nz = -(1.5f - pcx);
            nn = sqrtf(fx * fx + nz * nz);
            if (fx * ccx + nz * ccz + rad * nn < 0.f) cull = true;
            nz = (P.cols + 0.5f - pcx);
            nn = sqrtf(fx * fx + nz * nz);
            if (-fx * ccx + nz * ccz + rad * nn < 0.f) cull = true;
            nz = -(1.5f - pcy);
            nn = sqrtf(fy * fy + nz * nz);
            if (fy * ccy + nz * ccz + rad * nn < 0.f) cull = true;
            nz = (P.rows + 0.5f - pcy);
            nn = sqrtf(fy * fy + nz * nz);
            if (-fy * ccy + nz * ccz + rad * nn < 0.f) cull = true;
            if (!cull) {
                // footprint of the bounding sphere: |u - u_c| <= |f| rad (Z + |X|) / (Z (Z - rad)), padded by the 2 pixels the
                // bilinear / nearest depth look-up can reach; sdf = (Dp - z) * |ray| with |ray| >= 1, so z_min > max Dp + trunc
                // implies sdf < -trunc for every voxel (and max Dp = 0, no valid depth, implies Dp <= 0 for every voxel)
                const float zmin = ccz - rad, inv = 1.f / (ccz * zmin);
                const float ru = fabsf(fx) * rad * (ccz + fabsf(ccx)) * inv + 2.5f, rv = fabsf(fy) * rad * (ccz + fabsf(ccy)) * inv + 2.5f;
                const float uc = fx * ccx / ccz + pcx, vc = fy * ccy / ccz + pcy;
                const int tx0 = max(0, (int) floorf((uc - ru) * (1.f / CULL_TILE))), tx1 = min(P.tiles_x - 1, (int) floorf((uc + ru) * (1.f / CULL_TILE)));
                const int ty0 = max(0, (int) floorf((vc - rv) * (1.f / CULL_TILE))), ty1 = min(P.tiles_y - 1, (int) floorf((vc + rv) * (1.f / CULL_TILE)));
                if (tx1 >= tx0 && ty1 >= ty0 && (tx1 - tx0 + 1) * (ty1 - ty0 + 1) <= 64) {
                    float dmax = 0.f;
                    for (int ty = ty0; ty <= ty1; ++ty)
                        for (int tx = tx0; tx <= tx1; ++tx) dmax = fmaxf(dmax, __ldg(P.tile_max + ty * P.tiles_x + tx));
                    if (zmin > dmax + P.V.trunc + 2e-3f) cull = true;
                }
            }
        }
        keep = !cull;
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    unsigned base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(P.list_count, (unsigned) __popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    // entry = (brick index, packed brick coordinates) so the integration kernel does no integer division
    if (keep) list[base + __popc(m & ((1u << lane) - 1u))] = make_int2(b, bx | (by << 10) | (bz << 20));
}

// Real path + derivative of sdf w.r.t. v_c for one voxel.  K = 3 (C=1: gradient) or 6 (C=3: pairs
// (0,0),(0,1),(0,2),(1,1),(1,2),(2,2) -> gradient + Hessian).  Returns false when the voxel is skipped.
template <int C, int K>
XS_DEV bool eval_voxel(const IntegrateParams &P, float vcx, float vcy, float vcz, Jet<C, K> &sdf) {
    typedef Jet<C, K> J;
    J X = jconst<C, K>(vcx), Y = jconst<C, K>(vcy), Z = jconst<C, K>(vcz);
    if (K == 0) {
        // real-only evaluation (decisions and the real sdf): no seeds
    } else if (C == 1) {
        X.d[0] = 1.f;
        Y.d[1] = 1.f;
        Z.d[2] = 1.f;
    } else {
        // direction p=(i,j): eps1 along e_i, eps2 along e_j
        const int pi[6] = {0, 0, 0, 1, 1, 2}, pj[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
        for (int p = 0; p < 6; ++p) {
            if (pi[p] == 0) X.d[3 * p] = 1.f;
            if (pi[p] == 1) Y.d[3 * p] = 1.f;
            if (pi[p] == 2) Z.d[3 * p] = 1.f;
            if (pj[p] == 0) X.d[3 * p + 1] = 1.f;
            if (pj[p] == 1) Y.d[3 * p + 1] = 1.f;
            if (pj[p] == 2) Z.d[3 * p + 1] = 1.f;
        }
    }
    // TsdfFusion.cu:115-119
    J inv_z = jconst<C, K>(1.0f) / Z;
    if (inv_z.v < 0) return false;
    J image_x = jaddf(jmulf(X, P.intr.fx) * inv_z, P.intr.cx);
    J image_y = jaddf(jmulf(Y, P.intr.fy) * inv_z, P.intr.cy);
    // :120-124
    int coox = __float2int_rd(__fsub_rn(image_x.v, 0.5f));
    int cooy = __float2int_rd(__fsub_rn(image_y.v, 0.5f));
    if (!(coox > 1 && cooy > 1 && coox < P.cols - 1 && cooy < P.rows - 1)) return false;
    // :125-143
    int nx = __float2int_rn(image_x.v), ny = __float2int_rn(image_y.v);
    const float *d = P.depth;
    float d00 = __ldg(d + (size_t) cooy * P.cols + coox);
    float d10 = __ldg(d + (size_t) cooy * P.cols + coox + 1);
    float d01 = __ldg(d + (size_t) (cooy + 1) * P.cols + coox);
    float d11 = __ldg(d + (size_t) (cooy + 1) * P.cols + coox + 1);
    float gmax = fmaxf(d00, fmaxf(d01, fmaxf(d10, d11)));
    float gmin = fminf(d00, fminf(d01, fminf(d10, d11)));
    J Dp;
    if (__fsub_rn(gmax, gmin) < P.threshold && ((d00 != 0.0f) & (d01 != 0.0f) & (d10 != 0.0f) & (d11 != 0.0f))) {
        J a = jsubf(image_x, __fadd_rn(float(coox), 0.5f));
        J b = jsubf(image_y, __fadd_rn(float(cooy), 0.5f));
        J one_a = jrsubf(1.0f, a), one_b = jrsubf(1.0f, b);
        Dp = ((jmulf(one_a, d00) * one_b + jmulf(a, d10) * one_b) + jmulf(one_a, d01) * b) + jmulf(a, d11) * b;
    } else {
        Dp = jconst<C, K>(__ldg(d + (size_t) ny * P.cols + nx));
    }
    // :144-150
    J xl = jdivf(jsubf(image_x, P.intr.cx), P.intr.fx);
    J yl = jdivf(jsubf(image_y, P.intr.cy), P.intr.fy);
    Jet3<C, K> v1 = {Dp * xl, Dp * yl, Dp};
    Jet3<C, K> vc = {X, Y, Z};
    sdf = jnorm(v1) - jnorm(vc);
    return Dp.v > 0 && sdf.v >= -P.V.trunc;
}

// One CTA pass = half a brick (256 threads, 4 z-slices): 3 CTAs per SM at <= 85 registers instead of one 512-thread CTA,
// i.e. 24 resident warps and three independent streams of derivative-plane loads per SM.
constexpr int INT_THREADS = 256;
// KIND = batch kind (xs_batch.h): 1 = first-order list, 3 = bicomplex list, 2 = Hessian batch (needs gradient AND Hessian of
// sdf like kind 3, but updates every first-order plane once and one second-order plane per listed pair).
// ---- cp.async.bulk staging (experiment, XS_INT_BULK=1): north_star (2) asks for shared-memory staging of the brick spans with
// TMA bulk copies "where the brick layout allows".  The layout allows it - the planes of a half brick are ncomp contiguous 1 KB
// spans - so the variant below stages every plane of an already-live half brick with cp.async.bulk (UBLKCP) behind an mbarrier,
// updates it in shared memory and writes it back with bulk stores; bricks that are not live yet keep the predicated LDG / STG
// path.  Measured against the default path in profiles/r02_ab_table.md.
XS_DEV unsigned smem_u32(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }
XS_DEV void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
XS_DEV void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
XS_DEV bool mbar_try_wait(unsigned long long *bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
XS_DEV void bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
XS_DEV void bulk_s2g(void *gdst, const void *smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
XS_DEV void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
XS_DEV void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
XS_DEV void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
XS_DEV void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int KIND, bool BULK = false> __global__ void __launch_bounds__(INT_THREADS, 3) integrate_kernel(const IntegrateParams P) {
    constexpr int C = (KIND == 1) ? 1 : 3;
    constexpr int K = (C == 1) ? 3 : 6;
    extern __shared__ float4 s_dpose4[];  // [ncomp][3] float4 = [ncomp][12] floats, then (kind 2) the pair table int2[m]
    float *s_dpose = reinterpret_cast<float *>(s_dpose4);
    const int ncomp = P.V.ncomp;
    for (int i = threadIdx.x; i < ncomp * 12; i += INT_THREADS) s_dpose[i] = P.dpose[i];
    int2 *s_pairs = reinterpret_cast<int2 *>(s_dpose + (size_t) ncomp * 12);
    if (KIND == 2)
        for (int i = threadIdx.x; i < P.batch.m; i += INT_THREADS) s_pairs[i] = P.batch.pairs[i];
    // BULK: [ncomp][256] floats of the half brick in flight, 128-byte aligned behind the tables
    float *s_stage = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(s_pairs + (KIND == 2 ? P.batch.m : 0)) + 127) & ~uintptr_t(127));
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ bool s_live;
    unsigned bar_phase = 0;
    if (BULK && threadIdx.x == 0) mbar_init(&s_bar, 1);
    __syncthreads();

    const float vs = P.V.voxel;
    const float *R = P.v2c.R, *t = P.v2c.t;
    unsigned long long n_upd = 0, n_der = 0;
    const int nlist = (int) *P.list_count;

    for (int hi = blockIdx.x; hi < 2 * nlist; hi += gridDim.x) {
        const int li = hi >> 1;
        const int tid = ((hi & 1) << 8) | threadIdx.x;  // brick-local voxel index
        const int lx = tid & 7, ly = (tid >> 3) & 7, lz = tid >> 6;
        const int2 entry = P.brick_list[li];
        const int b = entry.x;
        const int bx = entry.y & 1023, by = (entry.y >> 10) & 1023, bz = entry.y >> 20;
        // derivative planes of a brick that never held a truncation-band voxel are exactly zero: scaling them by
        // w/(w+1) is the identity, so free-space bricks move no derivative bytes at all
        bool live = ncomp > 0 && P.live[b] != 0;
        if (BULK) {  // the flag may be raised by the CTA that holds the other half of the brick while this one reads it: one reader
            __syncthreads();
            if (threadIdx.x == 0) s_live = live;
            __syncthreads();
            live = s_live;
        }
        const bool staged = BULK && live;  // CTA-uniform
        if (staged) {
            if (threadIdx.x == 0) bulk_wait_read0();  // the stores of the previous staged half brick have left the buffer
            __syncthreads();
            if (threadIdx.x == 0) {
                mbar_expect_tx(&s_bar, (unsigned) ncomp * 1024u);
                const float *src = P.V.deriv + (size_t) b * ncomp * BRICK_VOX + ((hi & 1) << 8);
                for (int q = 0; q < ncomp; ++q) bulk_g2s(s_stage + q * 256, src + (size_t) q * BRICK_VOX, 1024u, &s_bar);
            }
        }
        [&]() {
        // ---- per-voxel real path (TsdfFusion.cu:110-114)
        const int x = bx * 8 + lx, y = by * 8 + ly, z = bz * 8 + lz;
        const float vgx = __fmul_rn(__fadd_rn(float(x), 0.5f), vs);
        const float vgy = __fmul_rn(__fadd_rn(float(y), 0.5f), vs);
        const float vgz = __fmul_rn(__fadd_rn(float(z), 0.5f), vs);
        const float vcx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[0], vgx), __fmul_rn(R[1], vgy)), __fmul_rn(R[2], vgz)), t[0]);
        const float vcy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[3], vgx), __fmul_rn(R[4], vgy)), __fmul_rn(R[5], vgz)), t[1]);
        const float vcz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[6], vgx), __fmul_rn(R[7], vgy)), __fmul_rn(R[8], vgz)), t[2]);
        // real pass first: most voxels in view are skipped (behind the surface) or saturated (free space, tsdf = 1,
        // zero derivative); only voxels inside the truncation band pay for the Jacobian / Hessian jets below.
        Jet<1, 0> sdf;
        const bool upd = eval_voxel<1, 0>(P, vcx, vcy, vcz, sdf);
        if (!upd) return;  // no barrier inside the per-voxel work: threads are independent
        ++n_upd;
        // ---- TsdfFusion.cu:152-167
        const bool saturated = sdf.v > P.V.trunc;
        const float tsdf = saturated ? 1.0f : __fmul_rn(sdf.v, P.trunc_inv);
        const size_t vi = (size_t) b * BRICK_VOX + tid;
        const int w_prev = P.V.weight[vi];
        const float wf = __int2float_rn(w_prev), wf1 = __int2float_rn(w_prev + 1);
        P.V.value[vi] = __fdiv_rn(__fmaf_rn(P.V.value[vi], wf, tsdf), wf1);
        P.V.weight[vi] = min(w_prev + 1, P.max_weight);
        if (ncomp == 0) return;
        // ---- derivative components
        const float inv_w1 = __fdiv_rn(1.f, wf1);
        const float a_keep = wf * inv_w1;
        // derivative planes of this voxel: in HBM (plane stride BRICK_VOX) or, staged, in shared memory (plane stride 256)
        float *dp = staged ? s_stage + threadIdx.x : P.V.deriv + (size_t) b * ncomp * BRICK_VOX + tid;
        const size_t DS = staged ? 256 : BRICK_VOX;
        if (staged)
            while (!mbar_try_wait(&s_bar, bar_phase)) {
            }
        if (saturated) {  // tsdf = (1, 0): F_q <- F_q * w / (w + 1)
            if (!live) return;
            ++n_der;
#pragma unroll 8
            for (int q = 0; q < ncomp; ++q) dp[(size_t) q * DS] *= a_keep;
            return;
        }
        ++n_der;
        if (!live) P.live[b] = 1;  // benign race: every writer stores 1; other voxels of the brick hold zeros
        const float sc = P.trunc_inv * inv_w1;
        Jet<C, K> sdfj;
        eval_voxel<C, K>(P, vcx, vcy, vcz, sdfj);
        if (KIND == 2) {
            // Hessian batch: gradient J and Hessian H of sdf w.r.t. v_c once per voxel (as for kind 3), then
            //   F_i  += J . d_i v_c                                   for the n first-order planes,
            //   S_ij += J . d_ij v_c + (d_j v_c)^T H (d_i v_c)        for the listed pairs (sorted by i: d_i v_c and H d_i v_c
            //                                                          are formed once per run of pairs with the same i)
            const float J0 = sdfj.d[0] * sc, J1 = sdfj.d[9] * sc, J2 = sdfj.d[15] * sc;
            const float H00 = sdfj.d[2] * sc, H01 = sdfj.d[5] * sc, H02 = sdfj.d[8] * sc, H11 = sdfj.d[11] * sc,
                        H12 = sdfj.d[14] * sc, H22 = sdfj.d[17] * sc;
            const int n = P.batch.n, m = P.batch.m;
#pragma unroll 4
            for (int q = 0; q < n; ++q) {
                const float o = dp[(size_t) q * DS];
                const float4 *mq = s_dpose4 + 3 * q;
                const float4 a0 = mq[0], a1 = mq[1], a2 = mq[2];
                const float dx = fmaf(a0.x, vgx, fmaf(a0.y, vgy, fmaf(a0.z, vgz, a2.y)));
                const float dy = fmaf(a0.w, vgx, fmaf(a1.x, vgy, fmaf(a1.y, vgz, a2.z)));
                const float dz = fmaf(a1.z, vgx, fmaf(a1.w, vgy, fmaf(a2.x, vgz, a2.w)));
                dp[(size_t) q * DS] = fmaf(o, a_keep, fmaf(J0, dx, fmaf(J1, dy, J2 * dz)));
            }
            int cur_i = -1;
            float hx = 0.f, hy = 0.f, hz = 0.f;  // H d_i v_c
            float *ps = dp + (size_t) n * DS;
#pragma unroll 4
            for (int k = 0; k < m; ++k) {
                const float o = ps[(size_t) k * DS];  // the load first: it bounds this loop
                const int2 pr = s_pairs[k];
                if (pr.x != cur_i) {  // block-uniform
                    cur_i = pr.x;
                    const float4 *mi = s_dpose4 + 3 * cur_i;
                    const float4 a0 = mi[0], a1 = mi[1], a2 = mi[2];
                    const float ax = fmaf(a0.x, vgx, fmaf(a0.y, vgy, fmaf(a0.z, vgz, a2.y)));
                    const float ay = fmaf(a0.w, vgx, fmaf(a1.x, vgy, fmaf(a1.y, vgz, a2.z)));
                    const float az = fmaf(a1.z, vgx, fmaf(a1.w, vgy, fmaf(a2.x, vgz, a2.w)));
                    hx = fmaf(H00, ax, fmaf(H01, ay, H02 * az));
                    hy = fmaf(H01, ax, fmaf(H11, ay, H12 * az));
                    hz = fmaf(H02, ax, fmaf(H12, ay, H22 * az));
                }
                const float4 *mj = s_dpose4 + 3 * pr.y, *ms = s_dpose4 + 3 * (n + k);
                const float4 b0 = mj[0], b1 = mj[1], b2 = mj[2], c0 = ms[0], c1 = ms[1], c2 = ms[2];
                const float bx = fmaf(b0.x, vgx, fmaf(b0.y, vgy, fmaf(b0.z, vgz, b2.y)));
                const float by_ = fmaf(b0.w, vgx, fmaf(b1.x, vgy, fmaf(b1.y, vgz, b2.z)));
                const float bz = fmaf(b1.z, vgx, fmaf(b1.w, vgy, fmaf(b2.x, vgz, b2.w)));
                const float cx2 = fmaf(c0.x, vgx, fmaf(c0.y, vgy, fmaf(c0.z, vgz, c2.y)));
                const float cy2 = fmaf(c0.w, vgx, fmaf(c1.x, vgy, fmaf(c1.y, vgz, c2.z)));
                const float cz2 = fmaf(c1.z, vgx, fmaf(c1.w, vgy, fmaf(c2.x, vgz, c2.w)));
                const float T = fmaf(J0, cx2, fmaf(J1, cy2, J2 * cz2)) + fmaf(bx, hx, fmaf(by_, hy, bz * hz));
                ps[(size_t) k * DS] = fmaf(o, a_keep, T);
            }
        } else if (C == 1) {
            const float J0 = sdfj.d[0] * sc, J1 = sdfj.d[1] * sc, J2 = sdfj.d[2] * sc;
#pragma unroll 4
            for (int q = 0; q < ncomp; ++q) {
                const float *m = s_dpose + q * 12;
                const float dx = fmaf(m[0], vgx, fmaf(m[1], vgy, fmaf(m[2], vgz, m[9])));
                const float dy = fmaf(m[3], vgx, fmaf(m[4], vgy, fmaf(m[5], vgz, m[10])));
                const float dz = fmaf(m[6], vgx, fmaf(m[7], vgy, fmaf(m[8], vgz, m[11])));
                const float T = fmaf(J0, dx, fmaf(J1, dy, J2 * dz));
                dp[(size_t) q * DS] = fmaf(dp[(size_t) q * DS], a_keep, T);
            }
        } else {
            // gradient from the diagonal pairs, Hessian from eps1eps2 of each pair
            const float J0 = sdfj.d[0] * sc, J1 = sdfj.d[9] * sc, J2 = sdfj.d[15] * sc;
            const float H00 = sdfj.d[2] * sc, H01 = sdfj.d[5] * sc, H02 = sdfj.d[8] * sc, H11 = sdfj.d[11] * sc,
                        H12 = sdfj.d[14] * sc, H22 = sdfj.d[17] * sc;
            const int dirs = ncomp / 3;
#pragma unroll 2
            for (int k = 0; k < dirs; ++k) {
                float *p = dp + (size_t) (3 * k) * DS;
                const float o1 = p[0], o2 = p[DS], o12 = p[2 * DS];  // loads first: they bound this loop
                const float4 *m = s_dpose4 + 9 * k;  // rows of (dR | dt) for eps1, eps2, eps1eps2
                const float4 a0 = m[0], a1 = m[1], a2 = m[2], b0 = m[3], b1 = m[4], b2 = m[5], c0 = m[6], c1 = m[7], c2 = m[8];
                const float ax = fmaf(a0.x, vgx, fmaf(a0.y, vgy, fmaf(a0.z, vgz, a2.y)));
                const float ay = fmaf(a0.w, vgx, fmaf(a1.x, vgy, fmaf(a1.y, vgz, a2.z)));
                const float az = fmaf(a1.z, vgx, fmaf(a1.w, vgy, fmaf(a2.x, vgz, a2.w)));
                const float bxx = fmaf(b0.x, vgx, fmaf(b0.y, vgy, fmaf(b0.z, vgz, b2.y)));
                const float byy = fmaf(b0.w, vgx, fmaf(b1.x, vgy, fmaf(b1.y, vgz, b2.z)));
                const float bzz = fmaf(b1.z, vgx, fmaf(b1.w, vgy, fmaf(b2.x, vgz, b2.w)));
                const float cx2 = fmaf(c0.x, vgx, fmaf(c0.y, vgy, fmaf(c0.z, vgz, c2.y)));
                const float cy2 = fmaf(c0.w, vgx, fmaf(c1.x, vgy, fmaf(c1.y, vgz, c2.z)));
                const float cz2 = fmaf(c1.z, vgx, fmaf(c1.w, vgy, fmaf(c2.x, vgz, c2.w)));
                const float T1 = fmaf(J0, ax, fmaf(J1, ay, J2 * az));
                const float T2 = fmaf(J0, bxx, fmaf(J1, byy, J2 * bzz));
                const float hx = fmaf(H00, bxx, fmaf(H01, byy, H02 * bzz));
                const float hy = fmaf(H01, bxx, fmaf(H11, byy, H12 * bzz));
                const float hz = fmaf(H02, bxx, fmaf(H12, byy, H22 * bzz));
                const float T12 = fmaf(J0, cx2, fmaf(J1, cy2, J2 * cz2)) + fmaf(ax, hx, fmaf(ay, hy, az * hz));
                p[0] = fmaf(o1, a_keep, T1);
                p[DS] = fmaf(o2, a_keep, T2);
                p[2 * DS] = fmaf(o12, a_keep, T12);
            }
        }
        }();
        if (staged) {  // the updated planes go back as bulk stores
            while (!mbar_try_wait(&s_bar, bar_phase)) {  // every thread: the loads have landed (also when no voxel of the half brick
            }                                            // was updated), so the barrier's phase is over before it is armed again
            fence_async_smem();
            __syncthreads();
            if (threadIdx.x == 0) {
                float *dst = P.V.deriv + (size_t) b * ncomp * BRICK_VOX + ((hi & 1) << 8);
                for (int q = 0; q < ncomp; ++q) bulk_s2g(dst + (size_t) q * BRICK_VOX, s_stage + q * 256, 1024u);
                bulk_commit();
            }
            bar_phase ^= 1u;
        }
    }
    if (BULK && threadIdx.x == 0) bulk_wait0();
    // updated-voxel count (drives the algorithmic-bytes model)
    if (P.stats) {
        for (int o = 16; o > 0; o >>= 1) {
            n_upd += __shfl_down_sync(0xffffffffu, n_upd, o);
            n_der += __shfl_down_sync(0xffffffffu, n_der, o);
        }
        if ((threadIdx.x & 31) == 0 && n_upd) atomicAdd(P.stats, n_upd);
        if ((threadIdx.x & 31) == 0 && n_der) atomicAdd(P.stats + 2, n_der);
        if (threadIdx.x == 0 && blockIdx.x == 0) P.stats[1] = (unsigned long long) nlist;
    }
}

// dense [z][y][x] planes <-> brick layout (seam views of TsdfVolume::value/weight/grad)
template <bool TO_DENSE>
__global__ void convert_planes_kernel(VolumeView V, int comp, float *value, int *weight, float *grad) {
    size_t n = (size_t) V.rx * V.ry * V.rz;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        int x = (int) (i % V.rx), y = (int) ((i / V.rx) % V.ry), z = (int) (i / ((size_t) V.rx * V.ry));
        size_t vi = value_index(V, x, y, z);
        if (TO_DENSE) {
            if (value) value[i] = V.value[vi];
            if (weight) weight[i] = V.weight[vi];
            if (grad) grad[i] = V.deriv[deriv_index(V, x, y, z, comp)];
        } else {
            if (value) V.value[vi] = value[i];
            if (weight) V.weight[vi] = weight[i];
            if (grad) V.deriv[deriv_index(V, x, y, z, comp)] = grad[i];
        }
    }
}

int upload_pose_derivs(const xs_volume *v, const xs_pose *p, int slot, cudaStream_t s);

}  // namespace xs

using namespace xs;

extern "C" {

xs_volume *xs_volume_create(const int res[3], float voxel_size, float thres_range, int comps, int dirs) {
    return xs_volume_create_hessian(res, voxel_size, thres_range, comps == 2 ? dirs : -comps * 1000 - dirs, -1, nullptr);
}

// comps = 2 (Hessian batch): nparams first-order planes + one second-order plane per listed pair (pairs == NULL: all pairs).
// Internally also the creation path of the list kinds, encoded as nparams = -(comps * 1000 + dirs) by xs_volume_create.
xs_volume *xs_volume_create_hessian(const int res[3], float voxel_size, float thres_range, int nparams, int npairs, const int *pairs) {
    int comps = 2, dirs = nparams;
    if (nparams < 0) {
        comps = (-nparams) / 1000;
        dirs = (-nparams) % 1000;
    }
    if (!res || res[0] <= 0 || (res[0] % 8) || (res[1] % 8) || (res[2] % 8) || (comps != 1 && comps != 2 && comps != 3) || dirs < 0) {
        set_error("xs_volume_create: resolution must be a positive multiple of 8 and comps in {1,2,3}");
        return nullptr;
    }
    xs_volume *v = new xs_volume();
    if (batch_init(v->batch, comps, dirs, npairs, pairs) != XS_OK) {
        delete v;
        return nullptr;
    }
    VolumeView &V = v->view;
    V.rx = res[0];
    V.ry = res[1];
    V.rz = res[2];
    V.bx = res[0] / 8;
    V.by = res[1] / 8;
    V.bz = res[2] / 8;
    V.ncomp = v->batch.v.ncomp;
    V.voxel = voxel_size;
    V.trunc = fmaxf(voxel_size * thres_range, 2.1f * voxel_size);  // TsdfVolume.cpp:25,37
    v->comps = comps;
    v->dirs = dirs;
    size_t nvox = (size_t) res[0] * res[1] * res[2];
    v->bytes = nvox * 8 + nvox * 4 * V.ncomp;
    cudaError_t e = cudaMalloc(&V.value, nvox * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&V.weight, nvox * sizeof(int));
    V.deriv = nullptr;
    if (e == cudaSuccess && V.ncomp) e = cudaMalloc(&V.deriv, nvox * sizeof(float) * V.ncomp);
    size_t pose_floats = (size_t) (V.ncomp > 0 ? V.ncomp : 1) * 12 * 3;
    v->pipelined = false;
    v->d_tile_max = nullptr;
    v->tile_capacity = 0;
    if (e == cudaSuccess) e = cudaMalloc(&v->d_dpose, pose_floats * sizeof(float));
    if (e == cudaSuccess) e = cudaMallocHost(&v->h_dpose, pose_floats * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&v->d_stats, 8 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMallocHost(&v->h_stats, 8 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(v->d_stats, 0, 8 * sizeof(unsigned long long));
    if (e == cudaSuccess) std::memset(v->h_stats, 0, 8 * sizeof(unsigned long long));
    const size_t nbricks = nvox / BRICK_VOX;
    v->d_brick_list = nullptr;
    v->d_list_count = nullptr;
    v->d_live = nullptr;
    if (e == cudaSuccess) e = cudaMalloc(&v->d_brick_list, nbricks * sizeof(int2));
    if (e == cudaSuccess) e = cudaMalloc(&v->d_list_count, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc(&v->d_live, nbricks);
    v->ev_k0 = v->ev_k1 = nullptr;
    v->last_kernel_ms = 0.f;
    if (e == cudaSuccess) e = cudaEventCreate(&v->ev_k0);
    if (e == cudaSuccess) e = cudaEventCreate(&v->ev_k1);
    if (e == cudaSuccess) e = cudaEventCreate(&v->ev_h0);
    if (e == cudaSuccess) e = cudaEventCreate(&v->ev_h1);
    v->d_depth_m = nullptr;
    v->depth_capacity = 0;
    v->d_hit_time = nullptr;
    v->hit_capacity = 0;
    if (e != cudaSuccess) {
        set_error(std::string("xs_volume_create: ") + cudaGetErrorString(e));
        xs_volume_destroy(v);
        return nullptr;
    }
    if (xs_volume_reset(v, nullptr) != XS_OK) {
        xs_volume_destroy(v);
        return nullptr;
    }
    return v;
}

void xs_volume_destroy(xs_volume *v) {
    if (!v) return;
    cudaFree(v->view.value);
    cudaFree(v->view.weight);
    cudaFree(v->view.deriv);
    cudaFree(v->d_dpose);
    cudaFree(v->d_tile_max);
    cudaFreeHost(v->h_dpose);
    cudaFree(v->d_depth_m);
    cudaFree(v->d_hit_time);
    cudaFree(v->d_stats);
    cudaFreeHost(v->h_stats);
    cudaFree(v->d_brick_list);
    cudaFree(v->d_list_count);
    cudaFree(v->d_live);
    if (v->ev_k0) cudaEventDestroy(v->ev_k0);
    if (v->ev_k1) cudaEventDestroy(v->ev_k1);
    if (v->ev_h0) cudaEventDestroy(v->ev_h0);
    if (v->ev_h1) cudaEventDestroy(v->ev_h1);
    batch_free(v->batch);
    delete v;
}

int xs_volume_reset(xs_volume *v, void *stream) {
    if (!v) return XS_ERR_ARG;
    size_t nvox = (size_t) v->view.rx * v->view.ry * v->view.rz;
    reset_volume_kernel<<<sm_count() * 8, 256, 0, (cudaStream_t) stream>>>(v->view.value, v->view.weight, v->view.deriv, nvox,
                                                                     nvox * v->view.ncomp);
    XS_LAUNCH_CHECK();
    XS_CUDA(cudaMemsetAsync(v->d_live, 0, nvox / BRICK_VOX, (cudaStream_t) stream));
    XS_CUDA(cudaStreamSynchronize((cudaStream_t) stream));  // initVolume syncs, TsdfFusion.cu:42
    return XS_OK;
}

// intrinsic parameters of a Hessian batch (xs_batch.h): the raycast then differentiates the pixel ray as well
int xs_volume_set_intrinsic_seeds(xs_volume *v, const float *dintr) {
    if (!v) return XS_ERR_ARG;
    return batch_set_intrinsics(v->batch, dintr, 1.f, 1.f);  // the raycast takes 1 / fx, 1 / fy from its own intrinsics argument
}

float xs_volume_trunc_dist(const xs_volume *v) { return v ? v->view.trunc : 0.f; }
size_t xs_volume_bytes(const xs_volume *v) { return v ? v->bytes : 0; }
float xs_volume_last_integrate_ms(const xs_volume *v) { return v ? v->last_kernel_ms : 0.f; }

int xs_volume_export_planes(const xs_volume *v, int comp, float *d_value, int *d_weight, float *d_grad, void *stream) {
    if (!v || (d_grad && (comp < 0 || comp >= v->view.ncomp))) return XS_ERR_ARG;
    convert_planes_kernel<true><<<sm_count() * 8, 256, 0, (cudaStream_t) stream>>>(v->view, comp, d_value, d_weight, d_grad);
    XS_LAUNCH_CHECK();
    return XS_OK;
}
int xs_volume_import_planes(xs_volume *v, int comp, const float *d_value, const int *d_weight, const float *d_grad,
                            void *stream) {
    if (!v || (d_grad && (comp < 0 || comp >= v->view.ncomp))) return XS_ERR_ARG;
    convert_planes_kernel<false><<<sm_count() * 8, 256, 0, (cudaStream_t) stream>>>(
        v->view, comp, const_cast<float *>(d_value), const_cast<int *>(d_weight), const_cast<float *>(d_grad));
    XS_LAUNCH_CHECK();
    if (d_grad)  // imported derivative planes may be non-zero anywhere
        XS_CUDA(cudaMemsetAsync(v->d_live, 1, (size_t) v->view.bx * v->view.by * v->view.bz, (cudaStream_t) stream));
    return XS_OK;
}

}  // extern "C"

namespace xs {
int batch_init(Batch &b, int comps, int dirs, int npairs, const int *pairs) {
    batch_free(b);
    if ((comps != 1 && comps != 2 && comps != 3) || dirs < 0) {
        set_error("batch: comps must be 1 (CSFD list), 3 (DCSFD list) or 2 (Hessian batch)");
        return XS_ERR_ARG;
    }
    b.v.kind = comps;
    b.v.n = dirs;
    b.v.m = 0;
    b.v.pairs = nullptr;
    b.v.dintr = nullptr;
    b.v.cslot = nullptr;
    b.v.ncurr = 0;
    if (comps == 2) {
        const int m = pairs ? npairs : dirs * (dirs + 1) / 2;
        if (m < 0 || (pairs == nullptr && npairs > 0 && npairs != m)) {
            set_error("batch: bad pair list");
            return XS_ERR_ARG;
        }
        b.v.m = m;
        if (m > 0) {
            b.h_pairs = new int2[m];
            int k = 0;
            if (pairs) {
                for (; k < m; ++k) b.h_pairs[k] = make_int2(pairs[2 * k], pairs[2 * k + 1]);
            } else {
                for (int i = 0; i < dirs; ++i)
                    for (int j = i; j < dirs; ++j) b.h_pairs[k++] = make_int2(i, j);
            }
            for (k = 0; k < m; ++k) {
                const int2 p = b.h_pairs[k];
                if (p.x < 0 || p.y < p.x || p.y >= dirs || (k > 0 && b.h_pairs[k - 1].x > p.x)) {
                    set_error("batch: pairs must satisfy 0 <= i <= j < nparams and be sorted by i");
                    batch_free(b);
                    return XS_ERR_ARG;
                }
            }
            int2 *d = nullptr;
            if (cudaMalloc(&d, (size_t) m * sizeof(int2)) != cudaSuccess ||
                cudaMemcpy(d, b.h_pairs, (size_t) m * sizeof(int2), cudaMemcpyHostToDevice) != cudaSuccess) {
                set_error("batch: cannot upload the pair table (no CUDA device?)");
                cudaFree(d);
                batch_free(b);
                return XS_ERR_CUDA;
            }
            b.v.pairs = d;
        }
    }
    b.v.ncomp = comps == 2 ? dirs + b.v.m : comps * dirs;
    return XS_OK;
}
void batch_free(Batch &b) {
    cudaFree(const_cast<int2 *>(b.v.pairs));
    cudaFree(const_cast<float *>(b.v.dintr));
    cudaFree(const_cast<int *>(b.v.cslot));
    delete[] b.h_pairs;
    delete[] b.h_dintr;
    delete[] b.h_cslot;
    b.h_pairs = nullptr;
    b.h_dintr = nullptr;
    b.h_cslot = nullptr;
    b.v = BatchView{1, 0, 0, 0, nullptr, nullptr, nullptr, 0, 0.f, 0.f};
}
int batch_set_intrinsics(Batch &b, const float *dintr, float fx0, float fy0) {
    if (b.v.kind != 2) {
        set_error("intrinsic parameters need a Hessian batch (comps = 2)");
        return XS_ERR_ARG;
    }
    cudaFree(const_cast<float *>(b.v.dintr));
    cudaFree(const_cast<int *>(b.v.cslot));
    delete[] b.h_dintr;
    delete[] b.h_cslot;
    b.h_dintr = nullptr, b.h_cslot = nullptr, b.v.dintr = nullptr, b.v.cslot = nullptr, b.v.ncurr = 0;
    bool any = false;
    for (int i = 0; dintr && i < 4 * b.v.n; ++i) any = any || dintr[i] != 0.f;
    if (!any) return XS_OK;
    const int n = b.v.n, m = b.v.m;
    b.h_dintr = new float[4 * n];
    b.h_cslot = new int[n + m];
    std::memcpy(b.h_dintr, dintr, sizeof(float) * 4 * n);
    int slots = 0;
    auto moves = [&](int p) { return dintr[4 * p] != 0.f || dintr[4 * p + 1] != 0.f || dintr[4 * p + 2] != 0.f || dintr[4 * p + 3] != 0.f; };
    for (int p = 0; p < n; ++p) b.h_cslot[p] = moves(p) ? slots++ : -1;
    for (int k = 0; k < m; ++k) b.h_cslot[n + k] = (moves(b.h_pairs[k].x) && moves(b.h_pairs[k].y)) ? slots++ : -1;
    float *dd = nullptr;
    int *dc = nullptr;
    if (cudaMalloc(&dd, sizeof(float) * 4 * n) != cudaSuccess || cudaMalloc(&dc, sizeof(int) * (n + m)) != cudaSuccess ||
        cudaMemcpy(dd, b.h_dintr, sizeof(float) * 4 * n, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(dc, b.h_cslot, sizeof(int) * (n + m), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(dd);
        cudaFree(dc);
        set_error("batch: cannot upload the intrinsic seeds");
        return XS_ERR_CUDA;
    }
    b.v.dintr = dd;
    b.v.cslot = dc;
    b.v.ncurr = slots;
    b.v.gx0 = 1.f / fx0;
    b.v.gy0 = 1.f / fy0;
    return XS_OK;
}

// Copies the derivative components of a pose into staging slot `slot` (0 or 1) of the volume.
int upload_pose_derivs(const xs_volume *v, const xs_pose *p, int slot, cudaStream_t s) {
    const int ncomp = v->view.ncomp;
    if (p->ncomp != ncomp) {
        set_error("pose carries " + std::to_string(p->ncomp) + " derivative components, volume expects " + std::to_string(ncomp));
        return XS_ERR_ARG;
    }
    if (ncomp == 0) return XS_OK;
    float *h = v->h_dpose + (size_t) slot * ncomp * 12;
    for (int q = 0; q < ncomp; ++q) {
        for (int e = 0; e < 9; ++e) h[q * 12 + e] = p->dR[q * 9 + e];
        for (int e = 0; e < 3; ++e) h[q * 12 + 9 + e] = p->dt[q * 3 + e];
    }
    XS_CUDA(cudaMemcpyAsync(v->d_dpose + (size_t) slot * ncomp * 12, h, (size_t) ncomp * 12 * sizeof(float),
                            cudaMemcpyHostToDevice, s));
    return XS_OK;
}
}  // namespace xs

extern "C" int xs_volume_finish_frame(xs_volume *v, unsigned long long *stats_host);

namespace xs {
// The pose-independent head of an integration: metric depth, the per-tile depth maxima of the brick cull, cleared counters.
// The frame loop queues it behind the download of the ICP result so that it runs while the host does the pose algebra;
// xs_integrate then finds it done (same frame, same stream) and starts with the cull.
int integrate_prepare(xs_volume *v, const uint16_t *d_depth, size_t depth_step_bytes, int rows, int cols, cudaStream_t s) {
    if (v->depth_capacity < rows * cols) {
        cudaFree(v->d_depth_m);
        XS_CUDA(cudaMalloc(&v->d_depth_m, (size_t) rows * cols * sizeof(float)));
        v->depth_capacity = rows * cols;
    }
    const int tiles_x = div_up(cols, CULL_TILE), tiles_y = div_up(rows, CULL_TILE);
    if (v->tile_capacity < tiles_x * tiles_y) {
        cudaFree(v->d_tile_max);
        v->d_tile_max = nullptr;
        XS_CUDA(cudaMalloc(&v->d_tile_max, (size_t) tiles_x * tiles_y * sizeof(float)));
        v->tile_capacity = tiles_x * tiles_y;
    }
    dim3 blk(32, 8), grd(div_up(cols, 32), div_up(rows, 8));
    scale_depth_kernel<<<grd, blk, 0, s>>>(d_depth, depth_step_bytes, rows, cols, v->d_depth_m);
    XS_LAUNCH_CHECK();
    depth_tile_max_kernel<<<dim3(tiles_x, tiles_y), 256, 0, s>>>(v->d_depth_m, rows, cols, v->d_tile_max, tiles_x);
    XS_LAUNCH_CHECK();
    XS_CUDA(cudaMemsetAsync(v->d_stats, 0, 3 * sizeof(unsigned long long), s));  // [3] = extraction counter, [4..5] = raycast
    XS_CUDA(cudaMemsetAsync(v->d_list_count, 0, sizeof(unsigned int), s));
    v->prepared_depth = d_depth;
    return XS_OK;
}
}  // namespace xs

extern "C" int xs_integrate(xs_volume *v, const uint16_t *d_depth, size_t depth_step_bytes, int rows, int cols,
                            xs_intr intr, int max_weight, const xs_pose *v2c, float bilinear_threshold,
                            unsigned long long *stats_host, void *stream) {
    if (!v || !d_depth || !v2c || rows <= 0 || cols <= 0) return XS_ERR_ARG;
    cudaStream_t s = (cudaStream_t) stream;
    if (v->batch.v.dintr != nullptr && bilinear_threshold > 0.f) {
        // with the nearest-neighbour depth look-up (biInterpolate_threshold = 0, the reference's default) sdf does not depend on
        // the intrinsics at all: xl = (image_x - cx) / fx = X / Z.  The bilinear branch does (through the sub-pixel weights).
        set_error("xs_integrate: intrinsic parameters with the bilinear depth look-up (biInterpolate_threshold > 0) are not implemented");
        return XS_ERR_ARG;
    }
    // the staging buffer is reused by the next call: make sure the previous consumer is done.  In the frame loop (pipelined)
    // the caller synchronises once per frame and integration has a staging slot of its own, so nothing waits here.
    if (!v->pipelined) XS_CUDA(cudaStreamSynchronize(s));
    int rc = upload_pose_derivs(v, v2c, 2, s);
    if (rc != XS_OK) return rc;
    const bool prepared = v->pipelined && v->prepared_depth == d_depth;
    v->prepared_depth = nullptr;
    if (!prepared) {
        rc = integrate_prepare(v, d_depth, depth_step_bytes, rows, cols, s);
        v->prepared_depth = nullptr;
        if (rc != XS_OK) return rc;
    }

    IntegrateParams P;
    P.V = v->view;
    for (int i = 0; i < 9; ++i) P.v2c.R[i] = v2c->R[i];
    for (int i = 0; i < 3; ++i) P.v2c.t[i] = v2c->t[i];
    P.dpose = v->d_dpose + (size_t) 2 * v->view.ncomp * 12;
    P.depth = v->d_depth_m;
    P.rows = rows;
    P.cols = cols;
    P.intr = intr;
    P.max_weight = max_weight;
    P.threshold = bilinear_threshold;
    P.trunc_inv = 1.0f / v->view.trunc;  // TsdfFusion.cu:99
    P.stats = v->d_stats;
    P.nbricks = v->view.bx * v->view.by * v->view.bz;
    P.brick_list = v->d_brick_list;
    P.list_count = v->d_list_count;
    P.live = v->d_live;
    P.tiles_x = div_up(cols, CULL_TILE);
    P.tiles_y = div_up(rows, CULL_TILE);
    P.tile_max = v->d_tile_max;
    cull_bricks_kernel<<<div_up(P.nbricks, 256), 256, 0, s>>>(P, v->d_brick_list);
    XS_LAUNCH_CHECK();
    int grid = 2 * P.nbricks < sm_count() * 24 ? 2 * P.nbricks : sm_count() * 24;
    size_t smem = (size_t) (v->view.ncomp > 0 ? v->view.ncomp : 1) * 12 * sizeof(float) + (size_t) v->batch.v.m * sizeof(int2);
    P.batch = v->batch.v;
    XS_CUDA(cudaEventRecord(v->ev_k0, s));
    if (v->comps == 1)
        integrate_kernel<1><<<grid, INT_THREADS, smem, s>>>(P);
    else if (v->comps == 2) {
        static const char *bulk_env = getenv("XS_INT_BULK");  // experiment: cp.async.bulk staging of live half bricks
        if (bulk_env && *bulk_env == '1' && v->view.ncomp > 0) {
            const size_t smem_bulk = smem + 128 + (size_t) v->view.ncomp * 1024;
            static size_t smem_set = 0;
            if (smem_bulk > smem_set) {
                XS_CUDA(cudaFuncSetAttribute(integrate_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_bulk));
                smem_set = smem_bulk;
            }
            integrate_kernel<2, true><<<grid, INT_THREADS, smem_bulk, s>>>(P);
        } else {
            integrate_kernel<2><<<grid, INT_THREADS, smem, s>>>(P);
        }
    }
    else
        integrate_kernel<3><<<grid, INT_THREADS, smem, s>>>(P);
    XS_LAUNCH_CHECK();
    XS_CUDA(cudaEventRecord(v->ev_k1, s));
    XS_CUDA(cudaMemcpyAsync(v->h_stats, v->d_stats, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    if (v->pipelined) return XS_OK;     // the frame loop queues the raycast behind this without a host round trip
    XS_CUDA(cudaStreamSynchronize(s));  // integrateTsdfVolume syncs, TsdfFusion.cu:200
    return xs_volume_finish_frame(v, stats_host);
}

// After the stream has been synchronised: statistics and kernel time of the last integration.
extern "C" int xs_volume_finish_frame(xs_volume *v, unsigned long long *stats_host) {
    if (!v) return XS_ERR_ARG;
    if (stats_host)
        for (int i = 0; i < 4; ++i) stats_host[i] = v->h_stats[i];
    cudaEventElapsedTime(&v->last_kernel_ms, v->ev_k0, v->ev_k1);
    // the raycast of the collected frame has completed too: its hit kernel's duration (the events are re-recorded by the next
    // frame, so the value is taken here and kept)
    if (cudaEventElapsedTime(&v->last_hit_ms, v->ev_h0, v->ev_h1) != cudaSuccess) {
        v->last_hit_ms = 0.f;
        cudaGetLastError();
    }
    v->hit_stats[0] = v->h_stats[4];
    v->hit_stats[1] = v->h_stats[5];
    return XS_OK;
}

extern "C" int xs_volume_set_pipelined(xs_volume *v, int on) {
    if (!v) return XS_ERR_ARG;
    v->pipelined = on != 0;
    if (!v->pipelined) v->prepared_depth = nullptr;
    return XS_OK;
}
