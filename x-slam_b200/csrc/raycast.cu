// raycast.cu — direction-batched raycasting of the brick-tiled TSDF and map pyramid resizing.
//
// Replaces rayCastKernel / RayCaster::{operator(), readTsdf, getVoxel, interpolateTrilineary}
// (XKinectFusion/src/RayCaster.cu:69-141,197-310), raycast (RayCaster.cu:327-368) and
// resizeMapKernel / resizeVMap / resizeNMap (XKinectFusion/src/Map.cu:105-152,233-259).
//
// The march (RayCaster.cu:236-247) runs on the real value plane only and is shared by all perturbation
// directions.  The hit is then evaluated in a second kernel: the real path once per pixel (reference-faithful
// arithmetic), then one pass per direction that propagates derivative components only, reading derivative planes
// only at the 8 trilinear samples of the hit.
#include "xs_common.cuh"

namespace xs {

struct RaycastParams {
    VolumeView V;
    DevPose c2v, v2w;
    const float *dpose_c2v;  // [ncomp][12]
    const float *dpose_v2w;  // [ncomp][12]
    xs_intr intr;
    int rows, cols;
    int dirs;
    float time_step;
    float *vmap, *nmap;  // [(1+ncomp)][3][rows][cols]
    BatchView batch;     // meaning of the derivative components
    unsigned long long *stats;  // [4] pixels with a valid vertex, [5] with a valid normal (algorithmic-bytes model)
};

XS_DEV float read_value(const VolumeView &V, int x, int y, int z) {
    return __fadd_rn(__ldg(V.value + value_index(V, x, y, z)), 1e-5f);  // RayCaster.cu:74-76
}

// ray origin and direction in volume coordinates, RayCaster.cu:56-62,207-213
template <int C, int K>
XS_DEV void ray_setup(const RaycastParams &P, int x, int y, int k0, Jet3<C, K> &start, Jet3<C, K> &dir) {
    const JetPose<C, K> c2v = load_pose<C, K>(P.c2v, P.dpose_c2v, k0, P.dirs);
    Jet3<C, K> next;
    next.x = jconst<C, K>(__fdiv_rn(__fsub_rn(float(x), P.intr.cx), P.intr.fx));
    next.y = jconst<C, K>(__fdiv_rn(__fsub_rn(float(y), P.intr.cy), P.intr.fy));
    next.z = jconst<C, K>(1.f);
    start = c2v.t;
    const Jet3<C, K> ray_next = jrot(c2v, next) + c2v.t;
    dir = jnormalized(ray_next - start);
    // degenerate-direction patch (:211-213), decided on the real part
    if (dir.x.v == 0.f) dir.x = jconst<C, K>(1e-15f);
    if (dir.y.v == 0.f) dir.y = jconst<C, K>(1e-15f);
    if (dir.z.v == 0.f) dir.z = jconst<C, K>(1e-15f);
}

XS_DEV void store3(float *map, int comp, int rows, int cols, int y, int x, float a, float b, float c) {
    const size_t plane = (size_t) rows * cols;
    float *p = map + (size_t) comp * 3 * plane + (size_t) y * cols + x;
    p[0] = a;
    p[plane] = b;
    p[2 * plane] = c;
}

// floor(RN(p / vs)) — the voxel index of RayCaster.cu:80-86 — without an IEEE division on the common path: the product
// with the rounded reciprocal is within a few ulp of the correctly rounded quotient, so its floor can only differ when
// the quotient is that close to an integer; those cases (a fraction below 1e-3 of the samples) take the exact division.
XS_DEV int voxel_floor(float p, float vs, float inv_vs) {
    const float q = p * inv_vs;
    const float f = floorf(q);
    const float frac = q - f;
    if (frac < 1e-3f || frac > 0.999f || !(fabsf(q) < 4096.f)) return __float2int_rd(__fdiv_rn(p, vs));
    return (int) f;
}

// ---- pass 1: the real march, RayCaster.cu:222-247.  One thread per pixel on the value plane only; the result
// (time of the sample before the + -> - crossing, or a negative number) is shared by every direction.
// The loop reads LOOKAHEAD samples ahead of the exit tests (sample positions do not depend on loaded values), which
// turns a chain of dependent L2/HBM round trips into batches; the exit tests are still applied in order.
constexpr int MARCH_LOOKAHEAD = 4;
__global__ void __launch_bounds__(256) raycast_march_kernel(const RaycastParams P, float *__restrict__ hit_time) {
    const int x = threadIdx.x + blockIdx.x * 32;
    const int y = threadIdx.y + blockIdx.y * 8;
    if (x >= P.cols || y >= P.rows) return;
    const VolumeView &V = P.V;
    Jet3<1, 0> s0, d0;
    ray_setup<1, 0>(P, x, y, 0, s0, d0);
    const float sx = s0.x.v, sy = s0.y.v, sz = s0.z.v, dx = d0.x.v, dy = d0.y.v, dz = d0.z.v;
    const float vs = V.voxel;
    const float inv_vs = __fdiv_rn(1.f, vs);
    float time_curr = 0.2f;
    const float max_time = 5.0f;
    int gx = __float2int_rd(__fdiv_rn(__fmaf_rn(dx, time_curr, sx), vs));
    int gy = __float2int_rd(__fdiv_rn(__fmaf_rn(dy, time_curr, sy), vs));
    int gz = __float2int_rd(__fdiv_rn(__fmaf_rn(dz, time_curr, sz), vs));
    gx = max(0, min(gx, V.rx - 1));
    gy = max(0, min(gy, V.ry - 1));
    gz = max(0, min(gz, V.rz - 1));
    float tsdf = read_value(V, gx, gy, gz);
    float result = -1.f;
    bool done = false;
    while (!done && time_curr < max_time) {
        float tc[MARCH_LOOKAHEAD], val[MARCH_LOOKAHEAD];
        bool inside[MARCH_LOOKAHEAD];
        float t = time_curr;
#pragma unroll
        for (int i = 0; i < MARCH_LOOKAHEAD; ++i) {
            tc[i] = t;
            const float tt = __fadd_rn(t, P.time_step);
            const int ix = voxel_floor(__fmaf_rn(dx, tt, sx), vs, inv_vs);
            const int iy = voxel_floor(__fmaf_rn(dy, tt, sy), vs, inv_vs);
            const int iz = voxel_floor(__fmaf_rn(dz, tt, sz), vs, inv_vs);
            inside[i] = (ix >= 0 && iy >= 0 && iz >= 0 && ix < V.rx && iy < V.ry && iz < V.rz);
            val[i] = inside[i] ? read_value(V, ix, iy, iz) : 0.f;
            t = tt;  // time_curr += time_step (:236)
        }
#pragma unroll
        for (int i = 0; i < MARCH_LOOKAHEAD; ++i) {
            if (done) break;
            if (!(tc[i] < max_time) || !inside[i]) {
                done = true;
                break;
            }
            const float tsdf_prev = tsdf;
            tsdf = val[i];
            if (tsdf_prev < 0.f && tsdf > 0.f) {
                done = true;
                break;
            }
            if (tsdf_prev > 0.f && tsdf < 0.f) {
                result = tc[i];
                done = true;
                break;
            }
        }
        time_curr = t;
    }
    hit_time[(size_t) y * P.cols + x] = result;
}

// ---- pass 2: hit evaluation with the real path evaluated ONCE per pixel.
//
// A CTA owns HIT_PG groups of 32 consecutive pixels of one image row.  One warp per group evaluates the real hit (RayCaster.cu:249-305) with the
// reference-faithful arithmetic above and leaves, per pixel and per trilinear sample (2 for the crossing, 6 for the
// normal), a small context in shared memory: packed per-axis corner offsets, the weights (a, b, c), the gradient G and
// mixed second partials H of the trilinear interpolant with respect to (a, b, c), and its value.  Then the 8 warps
// loop over the perturbation directions (warp w takes directions w, w+8, ...; thread = pixel) and propagate ONLY
// derivative components:  d tri = <dF>_w + G.d(abc)  and, for the eps1eps2 component,
//   d12 tri = <F12>_w + G.d12(abc) + grad<F2>.d1(abc) + grad<F1>.d2(abc) + d1(abc)^T H d2(abc),
// where <.>_w is the trilinear contraction of that component's 8 corner values (7 lerps).  Real parts inside this loop
// are only coefficients, so they use MUFU-based division / rsqrt instead of the reference-faithful sequences.
constexpr int HIT_PX = 32, HIT_WARPS = 8;
// per-sample context fields
// S_OFF: 64-bit element offset (low / high word) of corner (0, 0, 0) into the derivative planes (component 0 of direction
// 0); S_SX / S_SY / S_SZ: signed 32-bit offset steps to the +x / +y / +z neighbour.  The brick-tiled index is separable,
// offset(x, y, z) = fx(x) + fy(y) + fz(z), so corner (i, j, k) sits at base + i*sx + j*sy + k*sz: 5 words instead of 16,
// which takes 11 shared-memory loads per sample and direction off the L1 pipe that bounds this kernel.
enum { S_OFF = 0, S_SX = 2, S_SY, S_SZ, S_A, S_B, S_C, S_GA, S_GB, S_GC, S_HAB, S_HAC, S_HBC, S_VAL, S_FIELDS };
// per-pixel context fields
enum { X_FLAGS = 0, X_T0, X_VX, X_VY, X_VZ, X_OK, X_FIELDS };
constexpr int HIT_CTX_WORDS = 8 * S_FIELDS + X_FIELDS;

// value, gradient and (optionally) mixed second partials of the trilinear interpolant of f[i*4 + j*2 + k]
// (i, j, k = x, y, z corner bits) with respect to the weights (a, b, c)
template <bool HESS>
XS_DEV void contract(const float (&f)[8], float a, float b, float c, float &val, float &ga, float &gb, float &gc, float &hab,
                     float &hac, float &hbc) {
    const float d00 = f[1] - f[0], v00 = fmaf(c, d00, f[0]);
    const float d01 = f[3] - f[2], v01 = fmaf(c, d01, f[2]);
    const float d10 = f[5] - f[4], v10 = fmaf(c, d10, f[4]);
    const float d11 = f[7] - f[6], v11 = fmaf(c, d11, f[6]);
    const float e0 = v01 - v00, w0 = fmaf(b, e0, v00), k0 = d01 - d00, gc0 = fmaf(b, k0, d00);
    const float e1 = v11 - v10, w1 = fmaf(b, e1, v10), k1 = d11 - d10, gc1 = fmaf(b, k1, d10);
    ga = w1 - w0;
    val = fmaf(a, ga, w0);
    gb = fmaf(a, e1 - e0, e0);
    gc = fmaf(a, gc1 - gc0, gc0);
    if (HESS) {
        hab = e1 - e0;
        hac = gc1 - gc0;
        hbc = fmaf(a, k1 - k0, k0);
    }
}
XS_DEV float contract_value(const float (&f)[8], float a, float b, float c) {
    const float v00 = fmaf(c, f[1] - f[0], f[0]), v01 = fmaf(c, f[3] - f[2], f[2]);
    const float v10 = fmaf(c, f[5] - f[4], f[4]), v11 = fmaf(c, f[7] - f[6], f[6]);
    const float w0 = fmaf(b, v01 - v00, v00), w1 = fmaf(b, v11 - v10, v10);
    return fmaf(a, w1 - w0, w0);
}

// Real trilinear sample, interpolateTrilineary (RayCaster.cu:99-141) in the reference's operation order, that also
// records the sample context.  Returns false where the reference returns NaN.
XS_DEV bool trilinear_real(const VolumeView &V, float px, float py, float pz, float *ctx /* [S_FIELDS][HIT_PX] + lane */,
                           float &out) {
    const float vs = V.voxel;
    int gx = __float2int_rd(__fdiv_rn(px, vs));
    int gy = __float2int_rd(__fdiv_rn(py, vs));
    int gz = __float2int_rd(__fdiv_rn(pz, vs));
    if (gx <= 0 || gx >= V.rx - 1) return false;
    if (gy <= 0 || gy >= V.ry - 1) return false;
    if (gz <= 0 || gz >= V.rz - 1) return false;
    if (__fmul_rn(__fadd_rn(float(gx), 0.5f), vs) >= px) gx -= 1;
    if (__fmul_rn(__fadd_rn(float(gy), 0.5f), vs) >= py) gy -= 1;
    if (__fmul_rn(__fadd_rn(float(gz), 0.5f), vs) >= pz) gz -= 1;
    const float a0 = __fdiv_rn(__fmaf_rn(-__fadd_rn(float(gx), 0.5f), vs, px), vs);
    const float b0 = __fdiv_rn(__fmaf_rn(-__fadd_rn(float(gy), 0.5f), vs, py), vs);
    const float c0 = __fdiv_rn(__fmaf_rn(-__fadd_rn(float(gz), 0.5f), vs, pz), vs);
    const float a1 = __fsub_rn(1.0f, a0), b1 = __fsub_rn(1.0f, b0), c1 = __fsub_rn(1.0f, c0);
    float f[8];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 2; ++k) f[i * 4 + j * 2 + k] = read_value(V, gx + i, gy + j, gz + k);
    float r = __fmul_rn(__fmul_rn(__fmul_rn(f[0], a1), b1), c1);
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(__fmul_rn(f[1], a1), b1), c0));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(__fmul_rn(f[2], a1), b0), c1));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(__fmul_rn(f[3], a1), b0), c0));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(__fmul_rn(f[4], a0), b1), c1));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(__fmul_rn(f[5], a0), b1), c0));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(__fmul_rn(f[6], a0), b0), c1));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(__fmul_rn(f[7], a0), b0), c0));
    out = r;
    float val, ga, gb, gc, hab, hac, hbc;
    contract<true>(f, a0, b0, c0, val, ga, gb, gc, hab, hac, hbc);
    {
        const long long off = (long long) deriv_index(V, gx, gy, gz, 0);
        ctx[S_OFF * HIT_PX] = __uint_as_float((unsigned) off);
        ctx[(S_OFF + 1) * HIT_PX] = __uint_as_float((unsigned) ((unsigned long long) off >> 32));
        ctx[S_SX * HIT_PX] = __int_as_float((int) ((long long) deriv_index(V, gx + 1, gy, gz, 0) - off));
        ctx[S_SY * HIT_PX] = __int_as_float((int) ((long long) deriv_index(V, gx, gy + 1, gz, 0) - off));
        ctx[S_SZ * HIT_PX] = __int_as_float((int) ((long long) deriv_index(V, gx, gy, gz + 1, 0) - off));
    }
    ctx[S_A * HIT_PX] = a0;
    ctx[S_B * HIT_PX] = b0;
    ctx[S_C * HIT_PX] = c0;
    ctx[S_GA * HIT_PX] = ga;
    ctx[S_GB * HIT_PX] = gb;
    ctx[S_GC * HIT_PX] = gc;
    ctx[S_HAB * HIT_PX] = hab;
    ctx[S_HAC * HIT_PX] = hac;
    ctx[S_HBC * HIT_PX] = hbc;
    ctx[S_VAL * HIT_PX] = r;
    return true;
}

// Real hit evaluation, RayCaster.cu:249-305, in three phases so that the six normal samples run on six warps.
// flags: bit 0 = vertex valid, bit 1 = normal valid, bit 2 = the normal samples are needed (vertex well inside).
// phase A (warp 0): crossing samples, hit time, vertex -> world vertex; leaves the volume-frame vertex in the context
XS_DEV unsigned hit_phase_a(const RaycastParams &P, int x, int y, float time_curr, float *ctx, float *xctx, float (&vw)[3]) {
    const VolumeView &V = P.V;
    Jet3<1, 0> start, dir;
    ray_setup<1, 0>(P, x, y, 0, start, dir);
    const float t1 = __fadd_rn(time_curr, P.time_step);
    float Ftdt, Ft;
    if (!trilinear_real(V, __fmaf_rn(dir.x.v, t1, start.x.v), __fmaf_rn(dir.y.v, t1, start.y.v), __fmaf_rn(dir.z.v, t1, start.z.v),
                        ctx + 0 * S_FIELDS * HIT_PX, Ftdt))
        return 0u;
    if (!trilinear_real(V, __fmaf_rn(dir.x.v, time_curr, start.x.v), __fmaf_rn(dir.y.v, time_curr, start.y.v),
                        __fmaf_rn(dir.z.v, time_curr, start.z.v), ctx + 1 * S_FIELDS * HIT_PX, Ft))
        return 0u;
    if (isnan(Ftdt) || isnan(Ft)) return 0u;
    const float coef = ref_cdiv_re(Ft, __fsub_rn(Ftdt, Ft));
    if (Ft < 0.0f || Ftdt > 0.0f) return 0u;
    const float Ts = __fmaf_rn(-coef, P.time_step, time_curr);
    Jet3<1, 0> vertex;
    vertex.x.v = __fadd_rn(start.x.v, __fmul_rn(dir.x.v, Ts));
    vertex.y.v = __fadd_rn(start.y.v, __fmul_rn(dir.y.v, Ts));
    vertex.z.v = __fadd_rn(start.z.v, __fmul_rn(dir.z.v, Ts));
    const JetPose<1, 0> v2w = load_pose<1, 0>(P.v2w, P.dpose_v2w, 0, P.dirs);
    const Jet3<1, 0> w = jrot(v2w, vertex) + v2w.t;
    vw[0] = w.x.v, vw[1] = w.y.v, vw[2] = w.z.v;
    xctx[X_VX * HIT_PX] = vertex.x.v;
    xctx[X_VY * HIT_PX] = vertex.y.v;
    xctx[X_VZ * HIT_PX] = vertex.z.v;
    const float vs = V.voxel;
    const int gx = __float2int_rd(__fdiv_rn(vertex.x.v, vs));
    const int gy = __float2int_rd(__fdiv_rn(vertex.y.v, vs));
    const int gz = __float2int_rd(__fdiv_rn(vertex.z.v, vs));
    if (!(gx > 1 && gy > 1 && gz > 1 && gx < V.rx - 2 && gy < V.ry - 2 && gz < V.rz - 2)) return 1u;
    return 1u | 4u;
}
// phase B (warps 2..7, sample = warp index): one normal sample at vertex +- half a voxel along axis (sample-2)/2
XS_DEV void hit_phase_b(const RaycastParams &P, int sample, float *ctx, float *xctx) {
    const float hv = __fmul_rn(P.V.voxel, 0.5f);
    float p[3] = {xctx[X_VX * HIT_PX], xctx[X_VY * HIT_PX], xctx[X_VZ * HIT_PX]};
    const int axis = (sample - 2) >> 1;
    p[axis] = (sample & 1) ? __fsub_rn(p[axis], hv) : __fadd_rn(p[axis], hv);
    float F;
    const bool ok = trilinear_real(P.V, p[0], p[1], p[2], ctx + sample * S_FIELDS * HIT_PX, F);
    if (!ok) xctx[X_OK * HIT_PX] = 0.f;  // cannot happen for g in (1, N-2); the reference would propagate NaN
}
// phase C (warp 0): the world-frame normal
XS_DEV bool hit_phase_c(const RaycastParams &P, const float *ctx, float (&ng)[3]) {
    Jet3<1, 0> n;
    n.x.v = __fsub_rn(ctx[(2 * S_FIELDS + S_VAL) * HIT_PX], ctx[(3 * S_FIELDS + S_VAL) * HIT_PX]);
    n.y.v = __fsub_rn(ctx[(4 * S_FIELDS + S_VAL) * HIT_PX], ctx[(5 * S_FIELDS + S_VAL) * HIT_PX]);
    n.z.v = __fsub_rn(ctx[(6 * S_FIELDS + S_VAL) * HIT_PX], ctx[(7 * S_FIELDS + S_VAL) * HIT_PX]);
    if (jdot(n, n).v == 0.f) return false;
    const JetPose<1, 0> v2w = load_pose<1, 0>(P.v2w, P.dpose_v2w, 0, P.dirs);
    const Jet3<1, 0> g = jrot(v2w, jnormalized(n));
    ng[0] = g.x.v, ng[1] = g.y.v, ng[2] = g.z.v;
    return true;
}

// derivative components of one trilinear sample for one direction; dpos = derivative of the sample position
// dq[cc] = first element of the plane of component cc inside brick 0 (deriv + comp * BRICK_VOX): the C components of a
// direction are consecutive planes for the list kinds, arbitrary ones (F_i, F_j, S_ij) for a Hessian batch.
template <int C>
XS_DEV Jet<C, 1> sample_deriv(const VolumeView &V, const float *const (&dq)[C], const float *ctx, const Jet3<C, 1> &pos, float inv_vs) {
    const float a = ctx[S_A * HIT_PX], b = ctx[S_B * HIT_PX], c = ctx[S_C * HIT_PX];
    float f[C][8];
    const long long o000 = (long long) ((unsigned long long) __float_as_uint(ctx[S_OFF * HIT_PX]) |
                                        ((unsigned long long) __float_as_uint(ctx[(S_OFF + 1) * HIT_PX]) << 32));
    const long long sx = __float_as_int(ctx[S_SX * HIT_PX]), sy = __float_as_int(ctx[S_SY * HIT_PX]), sz = __float_as_int(ctx[S_SZ * HIT_PX]);
#pragma unroll
    for (int cidx = 0; cidx < 8; ++cidx) {
        const long long o = o000 + ((cidx & 4) ? sx : 0) + ((cidx & 2) ? sy : 0) + ((cidx & 1) ? sz : 0);
#pragma unroll
        for (int cc = 0; cc < C; ++cc) f[cc][cidx] = __ldg(dq[cc] + o);
    }
    Jet<C, 1> r;
    r.v = ctx[S_VAL * HIT_PX];
    const float ga = ctx[S_GA * HIT_PX], gb = ctx[S_GB * HIT_PX], gc = ctx[S_GC * HIT_PX];
    float unused;
    if (C == 1) {
        const float v1 = contract_value(f[0], a, b, c);
        r.d[0] = fmaf(ga, pos.x.d[0] * inv_vs, fmaf(gb, pos.y.d[0] * inv_vs, fmaf(gc, pos.z.d[0] * inv_vs, v1)));
    } else {
        const float a1 = pos.x.d[0] * inv_vs, b1 = pos.y.d[0] * inv_vs, c1 = pos.z.d[0] * inv_vs;
        const float a2 = pos.x.d[1] * inv_vs, b2 = pos.y.d[1] * inv_vs, c2 = pos.z.d[1] * inv_vs;
        const float a12 = pos.x.d[2] * inv_vs, b12 = pos.y.d[2] * inv_vs, c12 = pos.z.d[2] * inv_vs;
        float v1, g1a, g1b, g1c, v2, g2a, g2b, g2c;
        contract<false>(f[0], a, b, c, v1, g1a, g1b, g1c, unused, unused, unused);
        contract<false>(f[1 % C], a, b, c, v2, g2a, g2b, g2c, unused, unused, unused);
        const float v12 = contract_value(f[2 % C], a, b, c);
        const float hab = ctx[S_HAB * HIT_PX], hac = ctx[S_HAC * HIT_PX], hbc = ctx[S_HBC * HIT_PX];
        r.d[0] = fmaf(ga, a1, fmaf(gb, b1, fmaf(gc, c1, v1)));
        r.d[1 % C] = fmaf(ga, a2, fmaf(gb, b2, fmaf(gc, c2, v2)));
        float t = fmaf(ga, a12, fmaf(gb, b12, fmaf(gc, c12, v12)));
        t = fmaf(g2a, a1, fmaf(g2b, b1, fmaf(g2c, c1, t)));
        t = fmaf(g1a, a2, fmaf(g1b, b2, fmaf(g1c, c2, t)));
        t = fmaf(hab, fmaf(a1, b2, b1 * a2), fmaf(hac, fmaf(a1, c2, c1 * a2), fmaf(hbc, fmaf(b1, c2, c1 * b2), t)));
        r.d[2 % C] = t;
    }
    return r;
}

// One perturbation direction of one pixel group: propagates the C derivative components comp[0..C-1] (indices into the
// pose derivative tables, the derivative planes and the output maps) through the hit evaluation; components
// c >= first_store are written (a pair of a Hessian batch recomputes F_i, F_j as coefficients and stores only S_ij).
template <int C>
XS_DEV void hit_direction(const RaycastParams &P, const float *ctx, int x, int y, float ny, float inv_vs, const int (&comp)[C],
                          int first_store) {
    typedef Jet<C, 1> J;
    const VolumeView &V = P.V;
    const float *xctx = ctx + 8 * S_FIELDS * HIT_PX;
    const unsigned flags = __float_as_uint(xctx[X_FLAGS * HIT_PX]);
    const float t0 = xctx[X_T0 * HIT_PX];
    const float nx = (float(x) - P.intr.cx) * __fdividef(1.f, P.intr.fx);
    Jet3<C, 1> vw, ng;
    if (flags & 1u) {
        // ray: start = t, dir = normalized(R * next)  (RayCaster.cu:56-62,207-213)
        const JetPose<C, 1> c2v = load_pose_comps<C>(P.c2v, P.dpose_c2v, comp);
        Jet3<C, 1> next = {jconst<C, 1>(nx), jconst<C, 1>(ny), jconst<C, 1>(1.f)};
        const Jet3<C, 1> start = c2v.t;
        Jet3<C, 1> dir = jnormalized_fast(jrot(c2v, next));
        // the reference patches exactly-zero direction components with a constant (:211-213)
        if (dir.x.v == 0.f) dir.x = jconst<C, 1>(1e-15f);
        if (dir.y.v == 0.f) dir.y = jconst<C, 1>(1e-15f);
        if (dir.z.v == 0.f) dir.z = jconst<C, 1>(1e-15f);
        const float t1 = t0 + P.time_step;
        const Jet3<C, 1> p1 = {jfmaf(dir.x, t1, start.x), jfmaf(dir.y, t1, start.y), jfmaf(dir.z, t1, start.z)};
        const Jet3<C, 1> p0 = {jfmaf(dir.x, t0, start.x), jfmaf(dir.y, t0, start.y), jfmaf(dir.z, t0, start.z)};
        const float *dq[C];
#pragma unroll
        for (int c = 0; c < C; ++c) dq[c] = V.deriv + (size_t) comp[c] * BRICK_VOX;
        const J Ftdt = sample_deriv<C>(V, dq, ctx + 0 * S_FIELDS * HIT_PX, p1, inv_vs);
        const J Ft = sample_deriv<C>(V, dq, ctx + 1 * S_FIELDS * HIT_PX, p0, inv_vs);
        const J coef = jdiv_fast(Ft, Ftdt - Ft);
        J Ts;
        Ts.v = fmaf(-coef.v, P.time_step, t0);
#pragma unroll
        for (int i = 0; i < C; ++i) Ts.d[i] = -P.time_step * coef.d[i];
        const Jet3<C, 1> vertex = {start.x + dir.x * Ts, start.y + dir.y * Ts, start.z + dir.z * Ts};
        const JetPose<C, 1> v2w = load_pose_comps<C>(P.v2w, P.dpose_v2w, comp);
        vw = jrot(v2w, vertex) + v2w.t;
        if (flags & 2u) {
            // the six normal samples sit at vertex +- half a voxel along one axis: same position derivative
            Jet3<C, 1> n;
            n.x = sample_deriv<C>(V, dq, ctx + 2 * S_FIELDS * HIT_PX, vertex, inv_vs) -
                  sample_deriv<C>(V, dq, ctx + 3 * S_FIELDS * HIT_PX, vertex, inv_vs);
            n.y = sample_deriv<C>(V, dq, ctx + 4 * S_FIELDS * HIT_PX, vertex, inv_vs) -
                  sample_deriv<C>(V, dq, ctx + 5 * S_FIELDS * HIT_PX, vertex, inv_vs);
            n.z = sample_deriv<C>(V, dq, ctx + 6 * S_FIELDS * HIT_PX, vertex, inv_vs) -
                  sample_deriv<C>(V, dq, ctx + 7 * S_FIELDS * HIT_PX, vertex, inv_vs);
            ng = jrot(v2w, jnormalized_fast(n));
        }
    }
#pragma unroll
    for (int i = 0; i < C; ++i) {
        if (i < first_store) continue;
        const int out = 1 + comp[i];
        if (flags & 1u)
            store3(P.vmap, out, P.rows, P.cols, y, x, vw.x.d[i], vw.y.d[i], vw.z.d[i]);
        else
            store3(P.vmap, out, P.rows, P.cols, y, x, 0.f, 0.f, 0.f);
        if (flags & 2u)
            store3(P.nmap, out, P.rows, P.cols, y, x, ng.x.d[i], ng.y.d[i], ng.z.d[i]);
        else
            store3(P.nmap, out, P.rows, P.cols, y, x, 0.f, 0.f, 0.f);
    }
}

// ---- Hessian batches: first-order results cached per (pixel, parameter) --------------------------------------------------
// The second-order component of a pair (i, j) needs, at every trilinear sample, the VALUE and the GRADIENT (w.r.t. the
// weights a, b, c) of the trilinear contraction of the first-order planes F_i and F_j - which do not depend on the pair.  The
// parameter tasks therefore leave them in shared memory, and a pair task gathers only its own plane S_ij (8 corners per sample
// instead of 24) and contracts it for the value only:
//   tri_ij = <S>_w + G.abc_ij + g_j.abc_i + g_i.abc_j + abc_i^T H abc_j          (G, H: real interpolant, ctx; g: cached)
// The six normal samples share one position derivative (the vertex's), so per axis only the DIFFERENCES (+ minus -) of value
// and gradient are cached.  Per (pixel, parameter): 2 crossing samples x (v, g[3]) + 3 axes x (dv, dg[3]) = 20 floats.
constexpr int HC_FIELDS = 20;
enum { HC_S0 = 0, HC_S1 = 4, HC_AX = 8 };  // + 4 * axis

// Pixel ray ((x - cx) / fx, (y - cy) / fy, 1) (RayCaster.cu:56-62) with the derivative components of parameters that move the
// intrinsics (xs_batch.h): with g = 1 / fx, a = x - cx:  F_p = -dcx_p g - a g^2 dfx_p,
// S_ij = (dcx_i dfx_j + dcx_j dfx_i) g^2 + 2 a g^3 dfx_i dfx_j; the y component likewise with (fy, cy).
XS_DEV void next_first(const RaycastParams &P, int x, int y, int p, float &dx, float &dy) {
    dx = dy = 0.f;
    if (P.batch.dintr == nullptr) return;
    const float4 d = __ldg(reinterpret_cast<const float4 *>(P.batch.dintr) + p);  // (dfx, dfy, dcx, dcy)
    const float gx = __fdividef(1.f, P.intr.fx), gy = __fdividef(1.f, P.intr.fy);
    dx = -d.z * gx - (float(x) - P.intr.cx) * gx * gx * d.x;
    dy = -d.w * gy - (float(y) - P.intr.cy) * gy * gy * d.y;
}
XS_DEV void next_second(const RaycastParams &P, int x, int y, int i, int j, float &dx, float &dy) {
    dx = dy = 0.f;
    if (P.batch.dintr == nullptr) return;
    const float4 a = __ldg(reinterpret_cast<const float4 *>(P.batch.dintr) + i), b = __ldg(reinterpret_cast<const float4 *>(P.batch.dintr) + j);
    const float gx = __fdividef(1.f, P.intr.fx), gy = __fdividef(1.f, P.intr.fy);
    dx = (a.z * b.x + b.z * a.x) * gx * gx + 2.f * (float(x) - P.intr.cx) * gx * gx * gx * a.x * b.x;
    dy = (a.w * b.y + b.w * a.y) * gy * gy + 2.f * (float(y) - P.intr.cy) * gy * gy * gy * a.y * b.y;
}

// value + gradient of the contraction of one first-order plane at one sample (8 corner gathers)
XS_DEV void sample_first(const float *__restrict__ plane, const float *ctx, float &v, float &ga, float &gb, float &gc) {
    const float a = ctx[S_A * HIT_PX], b = ctx[S_B * HIT_PX], c = ctx[S_C * HIT_PX];
    const long long o000 = (long long) ((unsigned long long) __float_as_uint(ctx[S_OFF * HIT_PX]) |
                                        ((unsigned long long) __float_as_uint(ctx[(S_OFF + 1) * HIT_PX]) << 32));
    const long long sx = __float_as_int(ctx[S_SX * HIT_PX]), sy = __float_as_int(ctx[S_SY * HIT_PX]), sz = __float_as_int(ctx[S_SZ * HIT_PX]);
    float f[8];
#pragma unroll
    for (int cidx = 0; cidx < 8; ++cidx)
        f[cidx] = __ldg(plane + o000 + ((cidx & 4) ? sx : 0) + ((cidx & 2) ? sy : 0) + ((cidx & 1) ? sz : 0));
    float unused;
    contract<false>(f, a, b, c, v, ga, gb, gc, unused, unused, unused);
}
XS_DEV float sample_value(const float *__restrict__ plane, const float *ctx) {
    const float a = ctx[S_A * HIT_PX], b = ctx[S_B * HIT_PX], c = ctx[S_C * HIT_PX];
    const long long o000 = (long long) ((unsigned long long) __float_as_uint(ctx[S_OFF * HIT_PX]) |
                                        ((unsigned long long) __float_as_uint(ctx[(S_OFF + 1) * HIT_PX]) << 32));
    const long long sx = __float_as_int(ctx[S_SX * HIT_PX]), sy = __float_as_int(ctx[S_SY * HIT_PX]), sz = __float_as_int(ctx[S_SZ * HIT_PX]);
    float f[8];
#pragma unroll
    for (int cidx = 0; cidx < 8; ++cidx)
        f[cidx] = __ldg(plane + o000 + ((cidx & 4) ? sx : 0) + ((cidx & 2) ? sy : 0) + ((cidx & 1) ? sz : 0));
    return contract_value(f, a, b, c);
}

// Parameter task (first order) of one pixel group: the C = 1 propagation of hit_direction, leaving the per-sample value /
// gradient of the F_i contraction in the cache hc[HC_FIELDS][HIT_PX] (+ lane)
XS_DEV void hit_first_cached(const RaycastParams &P, const float *ctx, float *hc, int x, int y, float ny, float inv_vs, int comp) {
    typedef Jet<1, 1> J;
    const VolumeView &V = P.V;
    const float *xctx = ctx + 8 * S_FIELDS * HIT_PX;
    const unsigned flags = __float_as_uint(xctx[X_FLAGS * HIT_PX]);
    const float t0 = xctx[X_T0 * HIT_PX];
    const float nx = (float(x) - P.intr.cx) * __fdividef(1.f, P.intr.fx);
    Jet3<1, 1> vw, ng;
    if (flags & 1u) {
        const int cc[1] = {comp};
        const JetPose<1, 1> c2v = load_pose_comps<1>(P.c2v, P.dpose_c2v, cc);
        Jet3<1, 1> next = {jconst<1, 1>(nx), jconst<1, 1>(ny), jconst<1, 1>(1.f)};
        next_first(P, x, y, comp, next.x.d[0], next.y.d[0]);
        const Jet3<1, 1> start = c2v.t;
        Jet3<1, 1> dir = jnormalized_fast(jrot(c2v, next));
        if (dir.x.v == 0.f) dir.x = jconst<1, 1>(1e-15f);
        if (dir.y.v == 0.f) dir.y = jconst<1, 1>(1e-15f);
        if (dir.z.v == 0.f) dir.z = jconst<1, 1>(1e-15f);
        const float t1 = t0 + P.time_step;
        const float *plane = V.deriv + (size_t) comp * BRICK_VOX;
        // crossing samples: sample 0 at t1, sample 1 at t0 (their position derivatives differ)
        J Fs[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const float *cs = ctx + s * S_FIELDS * HIT_PX;
            const float ts = s == 0 ? t1 : t0;
            float v, ga, gb, gc;
            sample_first(plane, cs, v, ga, gb, gc);
            hc[(HC_S0 + 4 * s + 0) * HIT_PX] = v;
            hc[(HC_S0 + 4 * s + 1) * HIT_PX] = ga;
            hc[(HC_S0 + 4 * s + 2) * HIT_PX] = gb;
            hc[(HC_S0 + 4 * s + 3) * HIT_PX] = gc;
            const float px = fmaf(dir.x.d[0], ts, start.x.d[0]) * inv_vs, py = fmaf(dir.y.d[0], ts, start.y.d[0]) * inv_vs,
                        pz = fmaf(dir.z.d[0], ts, start.z.d[0]) * inv_vs;
            Fs[s].v = cs[S_VAL * HIT_PX];
            Fs[s].d[0] = fmaf(cs[S_GA * HIT_PX], px, fmaf(cs[S_GB * HIT_PX], py, fmaf(cs[S_GC * HIT_PX], pz, v)));
        }
        const J coef = jdiv_fast(Fs[1], Fs[0] - Fs[1]);
        J Ts;
        Ts.v = fmaf(-coef.v, P.time_step, t0);
        Ts.d[0] = -P.time_step * coef.d[0];
        const Jet3<1, 1> vertex = {start.x + dir.x * Ts, start.y + dir.y * Ts, start.z + dir.z * Ts};
        const JetPose<1, 1> v2w = load_pose_comps<1>(P.v2w, P.dpose_v2w, cc);
        vw = jrot(v2w, vertex) + v2w.t;
        if (flags & 2u) {
            const float ax = vertex.x.d[0] * inv_vs, ay = vertex.y.d[0] * inv_vs, az = vertex.z.d[0] * inv_vs;
            Jet<1, 1> nn[3];
#pragma unroll
            for (int axis = 0; axis < 3; ++axis) {
                const float *cp = ctx + (2 + 2 * axis) * S_FIELDS * HIT_PX, *cm = ctx + (3 + 2 * axis) * S_FIELDS * HIT_PX;
                float vp, gpa, gpb, gpc, vm, gma, gmb, gmc;
                sample_first(plane, cp, vp, gpa, gpb, gpc);
                sample_first(plane, cm, vm, gma, gmb, gmc);
                const float dv = vp - vm;
                hc[(HC_AX + 4 * axis + 0) * HIT_PX] = dv;
                hc[(HC_AX + 4 * axis + 1) * HIT_PX] = gpa - gma;
                hc[(HC_AX + 4 * axis + 2) * HIT_PX] = gpb - gmb;
                hc[(HC_AX + 4 * axis + 3) * HIT_PX] = gpc - gmc;
                const float dGa = cp[S_GA * HIT_PX] - cm[S_GA * HIT_PX], dGb = cp[S_GB * HIT_PX] - cm[S_GB * HIT_PX],
                            dGc = cp[S_GC * HIT_PX] - cm[S_GC * HIT_PX];
                nn[axis].v = cp[S_VAL * HIT_PX] - cm[S_VAL * HIT_PX];
                nn[axis].d[0] = fmaf(dGa, ax, fmaf(dGb, ay, fmaf(dGc, az, dv)));
            }
            const Jet3<1, 1> n = {nn[0], nn[1], nn[2]};
            ng = jrot(v2w, jnormalized_fast(n));
        }
    }
    const int out = 1 + comp;
    if (flags & 1u)
        store3(P.vmap, out, P.rows, P.cols, y, x, vw.x.d[0], vw.y.d[0], vw.z.d[0]);
    else
        store3(P.vmap, out, P.rows, P.cols, y, x, 0.f, 0.f, 0.f);
    if (flags & 2u)
        store3(P.nmap, out, P.rows, P.cols, y, x, ng.x.d[0], ng.y.d[0], ng.z.d[0]);
    else
        store3(P.nmap, out, P.rows, P.cols, y, x, 0.f, 0.f, 0.f);
}

// Pair task (second order) of one pixel group from the cached first-order results of parameters i and j (hci, hcj): the
// bicomplex propagation of hit_direction<3> in which every sample gathers the pair's own plane only.
XS_DEV void hit_pair_cached(const RaycastParams &P, const float *ctx, const float *hci, const float *hcj, int x, int y, float ny,
                            float inv_vs, int ci, int cj, int cs_) {
    typedef Jet<3, 1> J;
    const VolumeView &V = P.V;
    const float *xctx = ctx + 8 * S_FIELDS * HIT_PX;
    const unsigned flags = __float_as_uint(xctx[X_FLAGS * HIT_PX]);
    const float t0 = xctx[X_T0 * HIT_PX];
    const float nx = (float(x) - P.intr.cx) * __fdividef(1.f, P.intr.fx);
    Jet3<3, 1> vw, ng;
    if (flags & 1u) {
        const int comp[3] = {ci, cj, cs_};
        const JetPose<3, 1> c2v = load_pose_comps<3>(P.c2v, P.dpose_c2v, comp);
        Jet3<3, 1> next = {jconst<3, 1>(nx), jconst<3, 1>(ny), jconst<3, 1>(1.f)};
        next_first(P, x, y, ci, next.x.d[0], next.y.d[0]);
        next_first(P, x, y, cj, next.x.d[1], next.y.d[1]);
        next_second(P, x, y, ci, cj, next.x.d[2], next.y.d[2]);
        const Jet3<3, 1> start = c2v.t;
        Jet3<3, 1> dir = jnormalized_fast(jrot(c2v, next));
        if (dir.x.v == 0.f) dir.x = jconst<3, 1>(1e-15f);
        if (dir.y.v == 0.f) dir.y = jconst<3, 1>(1e-15f);
        if (dir.z.v == 0.f) dir.z = jconst<3, 1>(1e-15f);
        const float t1 = t0 + P.time_step;
        const float *plane = V.deriv + (size_t) cs_ * BRICK_VOX;
        J Fs[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const float *cs = ctx + s * S_FIELDS * HIT_PX;
            const float ts = s == 0 ? t1 : t0;
            // position derivatives of the sample in voxel units: (a, b, c) components 1, 2, 12
            float pa[3], pb[3], pc[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                pa[q] = fmaf(dir.x.d[q], ts, start.x.d[q]) * inv_vs;
                pb[q] = fmaf(dir.y.d[q], ts, start.y.d[q]) * inv_vs;
                pc[q] = fmaf(dir.z.d[q], ts, start.z.d[q]) * inv_vs;
            }
            const float ga = cs[S_GA * HIT_PX], gb = cs[S_GB * HIT_PX], gc = cs[S_GC * HIT_PX];
            const float hab = cs[S_HAB * HIT_PX], hac = cs[S_HAC * HIT_PX], hbc = cs[S_HBC * HIT_PX];
            const float vi = hci[(HC_S0 + 4 * s) * HIT_PX], gia = hci[(HC_S0 + 4 * s + 1) * HIT_PX], gib = hci[(HC_S0 + 4 * s + 2) * HIT_PX],
                        gic = hci[(HC_S0 + 4 * s + 3) * HIT_PX];
            const float vj = hcj[(HC_S0 + 4 * s) * HIT_PX], gja = hcj[(HC_S0 + 4 * s + 1) * HIT_PX], gjb = hcj[(HC_S0 + 4 * s + 2) * HIT_PX],
                        gjc = hcj[(HC_S0 + 4 * s + 3) * HIT_PX];
            const float vs12 = sample_value(plane, cs);
            Fs[s].v = cs[S_VAL * HIT_PX];
            Fs[s].d[0] = fmaf(ga, pa[0], fmaf(gb, pb[0], fmaf(gc, pc[0], vi)));
            Fs[s].d[1] = fmaf(ga, pa[1], fmaf(gb, pb[1], fmaf(gc, pc[1], vj)));
            float t = fmaf(ga, pa[2], fmaf(gb, pb[2], fmaf(gc, pc[2], vs12)));
            t = fmaf(gja, pa[0], fmaf(gjb, pb[0], fmaf(gjc, pc[0], t)));
            t = fmaf(gia, pa[1], fmaf(gib, pb[1], fmaf(gic, pc[1], t)));
            t = fmaf(hab, fmaf(pa[0], pb[1], pb[0] * pa[1]), fmaf(hac, fmaf(pa[0], pc[1], pc[0] * pa[1]), fmaf(hbc, fmaf(pb[0], pc[1], pc[0] * pb[1]), t)));
            Fs[s].d[2] = t;
        }
        const J coef = jdiv_fast(Fs[1], Fs[0] - Fs[1]);
        J Ts;
        Ts.v = fmaf(-coef.v, P.time_step, t0);
#pragma unroll
        for (int q = 0; q < 3; ++q) Ts.d[q] = -P.time_step * coef.d[q];
        const Jet3<3, 1> vertex = {start.x + dir.x * Ts, start.y + dir.y * Ts, start.z + dir.z * Ts};
        const JetPose<3, 1> v2w = load_pose_comps<3>(P.v2w, P.dpose_v2w, comp);
        vw = jrot(v2w, vertex) + v2w.t;
        if (flags & 2u) {
            float pa[3], pb[3], pc[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) pa[q] = vertex.x.d[q] * inv_vs, pb[q] = vertex.y.d[q] * inv_vs, pc[q] = vertex.z.d[q] * inv_vs;
            J nn[3];
#pragma unroll
            for (int axis = 0; axis < 3; ++axis) {
                const float *cp = ctx + (2 + 2 * axis) * S_FIELDS * HIT_PX, *cm = ctx + (3 + 2 * axis) * S_FIELDS * HIT_PX;
                const float ga = cp[S_GA * HIT_PX] - cm[S_GA * HIT_PX], gb = cp[S_GB * HIT_PX] - cm[S_GB * HIT_PX], gc = cp[S_GC * HIT_PX] - cm[S_GC * HIT_PX];
                const float hab = cp[S_HAB * HIT_PX] - cm[S_HAB * HIT_PX], hac = cp[S_HAC * HIT_PX] - cm[S_HAC * HIT_PX],
                            hbc = cp[S_HBC * HIT_PX] - cm[S_HBC * HIT_PX];
                const float vi = hci[(HC_AX + 4 * axis) * HIT_PX], gia = hci[(HC_AX + 4 * axis + 1) * HIT_PX], gib = hci[(HC_AX + 4 * axis + 2) * HIT_PX],
                            gic = hci[(HC_AX + 4 * axis + 3) * HIT_PX];
                const float vj = hcj[(HC_AX + 4 * axis) * HIT_PX], gja = hcj[(HC_AX + 4 * axis + 1) * HIT_PX], gjb = hcj[(HC_AX + 4 * axis + 2) * HIT_PX],
                            gjc = hcj[(HC_AX + 4 * axis + 3) * HIT_PX];
                const float vs12 = sample_value(plane, cp) - sample_value(plane, cm);
                nn[axis].v = cp[S_VAL * HIT_PX] - cm[S_VAL * HIT_PX];
                nn[axis].d[0] = fmaf(ga, pa[0], fmaf(gb, pb[0], fmaf(gc, pc[0], vi)));
                nn[axis].d[1] = fmaf(ga, pa[1], fmaf(gb, pb[1], fmaf(gc, pc[1], vj)));
                float t = fmaf(ga, pa[2], fmaf(gb, pb[2], fmaf(gc, pc[2], vs12)));
                t = fmaf(gja, pa[0], fmaf(gjb, pb[0], fmaf(gjc, pc[0], t)));
                t = fmaf(gia, pa[1], fmaf(gib, pb[1], fmaf(gic, pc[1], t)));
                t = fmaf(hab, fmaf(pa[0], pb[1], pb[0] * pa[1]), fmaf(hac, fmaf(pa[0], pc[1], pc[0] * pa[1]), fmaf(hbc, fmaf(pb[0], pc[1], pc[0] * pb[1]), t)));
                nn[axis].d[2] = t;
            }
            const Jet3<3, 1> n = {nn[0], nn[1], nn[2]};
            ng = jrot(v2w, jnormalized_fast(n));
        }
    }
    const int out = 1 + cs_;
    if (flags & 1u)
        store3(P.vmap, out, P.rows, P.cols, y, x, vw.x.d[2], vw.y.d[2], vw.z.d[2]);
    else
        store3(P.vmap, out, P.rows, P.cols, y, x, 0.f, 0.f, 0.f);
    if (flags & 2u)
        store3(P.nmap, out, P.rows, P.cols, y, x, ng.x.d[2], ng.y.d[2], ng.z.d[2]);
    else
        store3(P.nmap, out, P.rows, P.cols, y, x, 0.f, 0.f, 0.f);
}

// HIT_PG pixel groups of 32 pixels per CTA: the real phases (A: one warp per group, B: one warp per (group, normal
// sample), C: one warp per group) fill the 8 warps four times better than with a single group, and the derivative loop
// runs over (direction, group) tasks.
// CACHED (Hessian batches with few enough parameters): the parameter tasks leave their per-sample contractions in shared
// memory ([HIT_PG][n][HC_FIELDS][HIT_PX] floats behind the context) and the pair tasks gather their own plane only.
constexpr int HIT_PG_LIST = 4, HIT_PG_CACHED = 2;
constexpr size_t hit_smem_bytes(int pg, int n_cached) { return (size_t) pg * (HIT_CTX_WORDS + (size_t) n_cached * HC_FIELDS) * HIT_PX * sizeof(float); }
template <int KIND, int HIT_PG, bool CACHED>
__global__ void __launch_bounds__(HIT_PX *HIT_WARPS, 2) raycast_hit_kernel(const RaycastParams P, const float *__restrict__ hit_time) {
    extern __shared__ float s_ctx[];  // [HIT_PG][HIT_CTX_WORDS][HIT_PX] (+ the first-order cache)
    const int lane = threadIdx.x, warp = threadIdx.y;
    const int y = blockIdx.y;
    const float qnan = __int_as_float(0x7fffffff);
    auto ctx_of = [&](int pg) { return s_ctx + (size_t) pg * HIT_CTX_WORDS * HIT_PX + lane; };
    auto x_of = [&](int pg) { return lane + (blockIdx.x * HIT_PG + pg) * HIT_PX; };
    unsigned flags_a = 0u;  // of the group this warp owns in phases A and C (warp < HIT_PG)
    if (warp < HIT_PG) {
        float *ctx = ctx_of(warp), *xctx = ctx + 8 * S_FIELDS * HIT_PX;
        const int x = x_of(warp);
        float t0 = -1.f;
        xctx[X_OK * HIT_PX] = 1.f;
        if (x < P.cols) {
            t0 = hit_time[(size_t) y * P.cols + x];
            float vw[3];
            if (t0 >= 0.f) flags_a = hit_phase_a(P, x, y, t0, ctx, xctx, vw);
            if (flags_a & 1u)
                store3(P.vmap, 0, P.rows, P.cols, y, x, vw[0], vw[1], vw[2]);
            else
                store3(P.vmap, 0, P.rows, P.cols, y, x, qnan, 0.f, 0.f);
        }
        xctx[X_FLAGS * HIT_PX] = __uint_as_float(flags_a);
        xctx[X_T0 * HIT_PX] = t0;
    }
    __syncthreads();
    for (int task = warp; task < 6 * HIT_PG; task += HIT_WARPS) {
        const int pg = task / 6, sample = 2 + task % 6;
        float *ctx = ctx_of(pg), *xctx = ctx + 8 * S_FIELDS * HIT_PX;
        if (__float_as_uint(xctx[X_FLAGS * HIT_PX]) & 4u) hit_phase_b(P, sample, ctx, xctx);
    }
    __syncthreads();
    if (warp < HIT_PG) {
        float *ctx = ctx_of(warp), *xctx = ctx + 8 * S_FIELDS * HIT_PX;
        const int x = x_of(warp);
        if (x < P.cols) {
            float ng[3];
            bool n_ok = false;
            if ((flags_a & 4u) && xctx[X_OK * HIT_PX] != 0.f) n_ok = hit_phase_c(P, ctx, ng);
            if (n_ok)
                store3(P.nmap, 0, P.rows, P.cols, y, x, ng[0], ng[1], ng[2]);
            else
                store3(P.nmap, 0, P.rows, P.cols, y, x, qnan, 0.f, 0.f);
            xctx[X_FLAGS * HIT_PX] = __uint_as_float((flags_a & 1u) | (n_ok ? 2u : 0u));
            flags_a = (flags_a & 1u) | (n_ok ? 2u : 0u);
        } else {
            flags_a = 0u;
        }
        if (P.stats) {  // valid-vertex / valid-normal pixel counts of the frame
            const unsigned mv = __ballot_sync(0xffffffffu, (flags_a & 1u) != 0), mn = __ballot_sync(0xffffffffu, (flags_a & 2u) != 0);
            if (lane == 0 && mv) atomicAdd(P.stats + 4, (unsigned long long) __popc(mv));
            if (lane == 0 && mn) atomicAdd(P.stats + 5, (unsigned long long) __popc(mn));
        }
    }
    __syncthreads();
    if (P.batch.ncomp == 0) return;
    const float inv_vs = __fdividef(1.f, P.V.voxel);
    const float ny = (float(y) - P.intr.cy) * __fdividef(1.f, P.intr.fy);
    // tasks: (direction, pixel group) for the list kinds; for a Hessian batch first the n parameters (first-order algebra,
    // component i) and then the m pairs (bicomplex algebra on (F_i, F_j, S_ij), of which only S_ij is stored)
    const int ntasks = (KIND == 2 ? (CACHED ? 0 : P.batch.n + P.batch.m) : P.dirs) * HIT_PG;
    for (int task = warp; task < ntasks; task += HIT_WARPS) {
        const int q = task / HIT_PG, pg = task % HIT_PG;
        const int x = x_of(pg);
        if (x >= P.cols) continue;
        const float *ctx = ctx_of(pg);
        if (KIND == 1) {
            const int comp[1] = {q};
            hit_direction<1>(P, ctx, x, y, ny, inv_vs, comp, 0);
        } else if (KIND == 3) {
            const int comp[3] = {3 * q, 3 * q + 1, 3 * q + 2};
            hit_direction<3>(P, ctx, x, y, ny, inv_vs, comp, 0);
        } else if (!CACHED) {
            if (q < P.batch.n) {
                const int comp[1] = {q};
                hit_direction<1>(P, ctx, x, y, ny, inv_vs, comp, 0);
            } else {
                const int2 pr = __ldg(P.batch.pairs + (q - P.batch.n));
                const int comp[3] = {pr.x, pr.y, q};
                hit_direction<3>(P, ctx, x, y, ny, inv_vs, comp, 2);
            }
        }
    }
    if (KIND == 2 && CACHED) {
        const int n = P.batch.n;
        float *s_hc = s_ctx + (size_t) HIT_PG * HIT_CTX_WORDS * HIT_PX;  // [HIT_PG][n][HC_FIELDS][HIT_PX]
        for (int task = warp; task < n * HIT_PG; task += HIT_WARPS) {
            const int q = task / HIT_PG, pg = task % HIT_PG;
            const int x = x_of(pg);
            if (x >= P.cols) continue;
            hit_first_cached(P, ctx_of(pg), s_hc + ((size_t) (pg * n + q) * HC_FIELDS) * HIT_PX + lane, x, y, ny, inv_vs, q);
        }
        __syncthreads();  // a pair reads the caches two parameter tasks (other warps) wrote
        for (int task = warp; task < P.batch.m * HIT_PG; task += HIT_WARPS) {
            const int k = task / HIT_PG, pg = task % HIT_PG;
            const int x = x_of(pg);
            if (x >= P.cols) continue;
            const int2 pr = __ldg(P.batch.pairs + k);
            hit_pair_cached(P, ctx_of(pg), s_hc + ((size_t) (pg * n + pr.x) * HC_FIELDS) * HIT_PX + lane,
                            s_hc + ((size_t) (pg * n + pr.y) * HC_FIELDS) * HIT_PX + lane, x, y, ny, inv_vs, pr.x, pr.y, n + k);
        }
    }
}

// resizeMapKernel, Map.cu:105-152, for packed-SoA maps with derivative components.
// Thread = (output pixel, slot): slot 0 writes the real part with the reference's arithmetic, slot s >= 1 writes the
// derivative components of task s-1 (its real coefficients come from the same four real texels, MUFU-normalised).  A task
// is a direction for the list kinds; for a Hessian batch the n parameters come first (first-order algebra), then the m
// pairs (bicomplex algebra on (F_i, F_j, S_ij), only S_ij stored).  Without normalisation (vertex maps) the operation is
// linear, so every component is an independent first-order task whatever the kind.
template <int C>
XS_DEV void resize_task(const float *__restrict__ p00, size_t splane, int scols, bool normalize, bool invalid, const int (&comp)[C],
                        int first_store, float *__restrict__ out, int drows, int dcols, int y, int x) {
    if (invalid) {
#pragma unroll
        for (int i = 0; i < C; ++i)
            if (i >= first_store) store3(out, 1 + comp[i], drows, dcols, y, x, 0.f, 0.f, 0.f);
        return;
    }
    Jet<C, 1> c[3];
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
        const float *p = p00 + pl * splane;
        const float2 a = *reinterpret_cast<const float2 *>(p), b = *reinterpret_cast<const float2 *>(p + scols);
        c[pl].v = (a.x + a.y + b.x + b.y) * 0.25f;
#pragma unroll
        for (int i = 0; i < C; ++i) {
            const float *pd = p + (size_t) (1 + comp[i]) * 3 * splane;
            const float2 da = *reinterpret_cast<const float2 *>(pd), db = *reinterpret_cast<const float2 *>(pd + scols);
            c[pl].d[i] = (da.x + da.y + db.x + db.y) * 0.25f;
        }
    }
    Jet3<C, 1> n = {c[0], c[1], c[2]};
    if (normalize) n = jnormalized_fast(n);
#pragma unroll
    for (int i = 0; i < C; ++i)
        if (i >= first_store) store3(out, 1 + comp[i], drows, dcols, y, x, n.x.d[i], n.y.d[i], n.z.d[i]);
}

template <int KIND, bool NORMALIZE>
__global__ void __launch_bounds__(256) resize_map_kernel(int drows, int dcols, int srows, int scols, BatchView B,
                                                         const float *__restrict__ in, float *__restrict__ out) {
    const int x = threadIdx.x + blockIdx.x * 32;
    const int y = blockIdx.y;
    const int slot = threadIdx.y + blockIdx.z * 8;
    const int ntasks = KIND == 2 ? B.n + B.m : B.n;
    if (x >= dcols || slot > ntasks) return;
    const size_t splane = (size_t) srows * scols;
    const float *p00 = in + (size_t) (y * 2) * scols + x * 2;
    const float qnan = __int_as_float(0x7fffffff);
    const float2 t0 = *reinterpret_cast<const float2 *>(p00), t1 = *reinterpret_cast<const float2 *>(p00 + scols);
    const bool invalid = isnan(t0.x) || isnan(t0.y) || isnan(t1.x) || isnan(t1.y);
    if (slot == 0) {
        if (invalid) {
            store3(out, 0, drows, dcols, y, x, qnan, 0.f, 0.f);
            return;
        }
        Jet3<1, 0> n;
        Jet<1, 0> *c[3] = {&n.x, &n.y, &n.z};
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) {
            const float *p = p00 + pl * splane;
            // (x00 + x01 + x10 + x11) / 4.0f
            c[pl]->v = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(p[0], p[1]), p[scols]), p[scols + 1]), 4.0f);
        }
        if (NORMALIZE) n = jnormalized(n);
        store3(out, 0, drows, dcols, y, x, n.x.v, n.y.v, n.z.v);
        return;
    }
    const int q = slot - 1;
    if (KIND == 1) {
        const int comp[1] = {q};
        resize_task<1>(p00, splane, scols, NORMALIZE, invalid, comp, 0, out, drows, dcols, y, x);
    } else if (KIND == 3) {
        const int comp[3] = {3 * q, 3 * q + 1, 3 * q + 2};
        resize_task<3>(p00, splane, scols, NORMALIZE, invalid, comp, 0, out, drows, dcols, y, x);
    } else if (q < B.n) {
        const int comp[1] = {q};
        resize_task<1>(p00, splane, scols, NORMALIZE, invalid, comp, 0, out, drows, dcols, y, x);
    } else {
        const int2 pr = __ldg(B.pairs + (q - B.n));
        const int comp[3] = {pr.x, pr.y, q};
        resize_task<3>(p00, splane, scols, NORMALIZE, invalid, comp, 2, out, drows, dcols, y, x);
    }
}

int upload_pose_derivs(const xs_volume *v, const xs_pose *p, int slot, cudaStream_t s);

// resizeVMap / resizeNMap for a batch description (the frame loop's entry; the public comps / dirs forms wrap it)
int resize_map_batch(bool normalize, const float *d_in, int rows, int cols, const BatchView &B, float *d_out, cudaStream_t s) {
    if (!d_in || !d_out || rows < 2 || cols < 2) return XS_ERR_ARG;
    if ((cols & 1) || (rows & 1)) return XS_ERR_ARG;  // 64-bit texel pairs
    const int drows = rows / 2, dcols = cols / 2;
    BatchView L = B;
    if (!normalize && B.kind != 1) {  // linear operation: every component is an independent first-order task
        L.kind = 1;
        L.n = B.ncomp;
        L.m = 0;
    }
    const int ntasks = L.kind == 2 ? L.n + L.m : L.n;
    dim3 blk(32, 8), grd(div_up(dcols, 32), drows, div_up(ntasks + 1, 8));
    if (L.kind == 1) {
        if (normalize)
            resize_map_kernel<1, true><<<grd, blk, 0, s>>>(drows, dcols, rows, cols, L, d_in, d_out);
        else
            resize_map_kernel<1, false><<<grd, blk, 0, s>>>(drows, dcols, rows, cols, L, d_in, d_out);
    } else if (L.kind == 3) {
        resize_map_kernel<3, true><<<grd, blk, 0, s>>>(drows, dcols, rows, cols, L, d_in, d_out);
    } else {
        resize_map_kernel<2, true><<<grd, blk, 0, s>>>(drows, dcols, rows, cols, L, d_in, d_out);
    }
    XS_LAUNCH_CHECK();
    return XS_OK;
}

// public form: comps = 1 / 3 with dirs directions, or comps = 2 with dirs parameters and all their pairs
template <bool NORMALIZE>
static int resize_map(const float *d_in, int rows, int cols, int comps, int dirs, float *d_out, void *stream) {
    if ((comps != 1 && comps != 2 && comps != 3) || dirs < 0) return XS_ERR_ARG;
    if (comps != 2 || !NORMALIZE) {
        const BatchView B = {comps == 2 ? 1 : comps, comps == 2 ? batch_ncomp(2, dirs, -1) : dirs, 0, batch_ncomp(comps, dirs, -1), nullptr, nullptr, nullptr, 0, 0.f, 0.f};
        return resize_map_batch(NORMALIZE, d_in, rows, cols, B, d_out, (cudaStream_t) stream);
    }
    Batch b;  // normal maps of a Hessian batch need the pair table on the device
    int rc = batch_init(b, 2, dirs, -1, nullptr);
    if (rc != XS_OK) return rc;
    rc = resize_map_batch(true, d_in, rows, cols, b.v, d_out, (cudaStream_t) stream);
    if (rc == XS_OK && cudaStreamSynchronize((cudaStream_t) stream) != cudaSuccess) rc = XS_ERR_CUDA;
    batch_free(b);
    return rc;
}

}  // namespace xs

using namespace xs;

extern "C" {

int xs_raycast(const xs_volume *v, xs_intr intr, const xs_pose *c2v, const xs_pose *v2w, int rows, int cols,
               float *d_vmap, float *d_nmap, void *stream) {
    if (!v || !c2v || !v2w || !d_vmap || !d_nmap || rows <= 0 || cols <= 0) return XS_ERR_ARG;
    // the hit kernel keeps the +x / +y / +z steps between derivative-plane corners as signed 32-bit element offsets
    if ((double) v->view.bx * v->view.by * (double) (v->view.ncomp > 0 ? v->view.ncomp : 1) * BRICK_VOX >= 2147483648.0) {
        set_error("xs_raycast: a z-step between bricks of the derivative planes must stay below 2^31 elements");
        return XS_ERR_ARG;
    }
    cudaStream_t s = (cudaStream_t) stream;
    if (!v->pipelined) XS_CUDA(cudaStreamSynchronize(s));  // staging buffer reuse (the frame loop synchronises once per frame)
    int rc = upload_pose_derivs(v, c2v, 0, s);
    if (rc != XS_OK) return rc;
    rc = upload_pose_derivs(v, v2w, 1, s);
    if (rc != XS_OK) return rc;
    RaycastParams P;
    P.V = v->view;
    for (int i = 0; i < 9; ++i) {
        P.c2v.R[i] = c2v->R[i];
        P.v2w.R[i] = v2w->R[i];
    }
    for (int i = 0; i < 3; ++i) {
        P.c2v.t[i] = c2v->t[i];
        P.v2w.t[i] = v2w->t[i];
    }
    P.dpose_c2v = v->d_dpose;
    P.dpose_v2w = v->d_dpose + (size_t) v->view.ncomp * 12;
    P.intr = intr;
    P.rows = rows;
    P.cols = cols;
    P.dirs = v->dirs;
    P.batch = v->batch.v;
    P.stats = v->d_stats;
    P.time_step = v->view.trunc * 0.8f;  // RayCaster.cu:350
    P.vmap = d_vmap;
    P.nmap = d_nmap;
    if (v->hit_capacity < rows * cols) {
        cudaFree(v->d_hit_time);
        xs_volume *vm = const_cast<xs_volume *>(v);
        vm->d_hit_time = nullptr;
        XS_CUDA(cudaMalloc(&vm->d_hit_time, (size_t) rows * cols * sizeof(float)));
        vm->hit_capacity = rows * cols;
    }
    dim3 blk(32, 8), grd(div_up(cols, 32), div_up(rows, 8));
    raycast_march_kernel<<<grd, blk, 0, s>>>(P, v->d_hit_time);
    XS_LAUNCH_CHECK();
    // Hessian batches cache the first-order contractions in shared memory when two CTAs of two pixel groups still fit an SM
    static const bool no_cache = getenv("XS_HIT_NO_CACHE") != nullptr;  // A/B knob
    const bool cached = v->comps == 2 && !no_cache && 2 * (hit_smem_bytes(HIT_PG_CACHED, v->batch.v.n) + 1024) <= 227 * 1024;
    if (v->batch.v.dintr != nullptr && !cached) {
        set_error("xs_raycast: intrinsic parameters are implemented on the cached Hessian-batch path (too many parameters for shared memory)");
        return XS_ERR_ARG;
    }
    const int pg = cached ? HIT_PG_CACHED : HIT_PG_LIST;
    dim3 g2(div_up(cols, HIT_PX * pg), rows), b2(HIT_PX, HIT_WARPS);
    const size_t hit_smem = hit_smem_bytes(pg, cached ? v->batch.v.n : 0);
    static bool hit_smem_set = false;
    if (!hit_smem_set) {
        XS_CUDA(cudaFuncSetAttribute(raycast_hit_kernel<1, HIT_PG_LIST, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) hit_smem_bytes(HIT_PG_LIST, 0)));
        XS_CUDA(cudaFuncSetAttribute(raycast_hit_kernel<3, HIT_PG_LIST, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) hit_smem_bytes(HIT_PG_LIST, 0)));
        XS_CUDA(cudaFuncSetAttribute(raycast_hit_kernel<2, HIT_PG_LIST, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) hit_smem_bytes(HIT_PG_LIST, 0)));
        XS_CUDA(cudaFuncSetAttribute(raycast_hit_kernel<2, HIT_PG_CACHED, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
        hit_smem_set = true;
    }
    XS_CUDA(cudaMemsetAsync(v->d_stats + 4, 0, 2 * sizeof(unsigned long long), s));
    XS_CUDA(cudaEventRecord(v->ev_h0, s));
    if (v->comps == 1)
        raycast_hit_kernel<1, HIT_PG_LIST, false><<<g2, b2, hit_smem, s>>>(P, v->d_hit_time);
    else if (cached)
        raycast_hit_kernel<2, HIT_PG_CACHED, true><<<g2, b2, hit_smem, s>>>(P, v->d_hit_time);
    else if (v->comps == 2)
        raycast_hit_kernel<2, HIT_PG_LIST, false><<<g2, b2, hit_smem, s>>>(P, v->d_hit_time);
    else
        raycast_hit_kernel<3, HIT_PG_LIST, false><<<g2, b2, hit_smem, s>>>(P, v->d_hit_time);
    XS_LAUNCH_CHECK();
    XS_CUDA(cudaEventRecord(v->ev_h1, s));
    XS_CUDA(cudaMemcpyAsync(v->h_stats + 4, v->d_stats + 4, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    return XS_OK;  // raycast does not sync, RayCaster.cu:367
}

// device duration of the last raycast hit kernel and its valid-vertex / valid-normal pixel counts; call after the stream has
// been synchronised (the frame loop: after the frame has been collected)
// (values of the last COLLECTED frame: xs_volume_finish_frame takes them once the frame's work has completed)
float xs_volume_last_raycast_hit_ms(const xs_volume *v) { return v ? v->last_hit_ms : 0.f; }
int xs_volume_raycast_stats(const xs_volume *v, unsigned long long *out2) {
    if (!v || !out2) return XS_ERR_ARG;
    out2[0] = v->hit_stats[0];
    out2[1] = v->hit_stats[1];
    return XS_OK;
}

int xs_resize_vmap(const float *d_in, int rows, int cols, int comps, int dirs, float *d_out, void *stream) {
    return resize_map<false>(d_in, rows, cols, comps, dirs, d_out, stream);
}
int xs_resize_nmap(const float *d_in, int rows, int cols, int comps, int dirs, float *d_out, void *stream) {
    return resize_map<true>(d_in, rows, cols, comps, dirs, d_out, stream);
}

}  // extern "C"
