// raycast.cu — direction-batched raycasting of the brick-tiled TSDF and map pyramid resizing.
//
// Replaces rayCastKernel / RayCaster::{operator(), readTsdf, getVoxel, interpolateTrilineary}
// (XKinectFusion/src/RayCaster.cu:69-141,197-310), raycast (RayCaster.cu:327-368) and
// resizeMapKernel / resizeVMap / resizeNMap (XKinectFusion/src/Map.cu:105-152,233-259).
//
// The march (RayCaster.cu:236-247) runs on the real value plane only and is shared by all perturbation
// directions.  Threads leave the march loop before the (expensive) hit evaluation so that a warp does the
// trilinear / normal work convergently; the hit is then evaluated once per direction tile with Jet<C,K>
// numbers, reading derivative planes only at the 8 trilinear samples of the hit.
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
};

XS_DEV float read_value(const VolumeView &V, int x, int y, int z) {
    return __fadd_rn(__ldg(V.value + value_index(V, x, y, z)), 1e-5f);  // RayCaster.cu:74-76
}

template <int C, int K> XS_DEV Jet<C, K> read_tsdf(const VolumeView &V, int x, int y, int z, int k0, int dirs) {
    Jet<C, K> r;
    const size_t base = (size_t) brick_of(V, x, y, z) * V.ncomp * BRICK_VOX + local_of(x, y, z);
    r.v = __fadd_rn(__ldg(V.value + value_index(V, x, y, z)), 1e-5f);
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int c = 0; c < C; ++c)
            r.d[k * C + c] = (k0 + k < dirs) ? __ldg(V.deriv + base + (size_t) ((k0 + k) * C + c) * BRICK_VOX) : 0.f;
    return r;
}

// interpolateTrilineary, RayCaster.cu:99-141.  Returns false where the reference returns NaN.
template <int C, int K>
XS_DEV bool trilinear(const VolumeView &V, const Jet3<C, K> &p, int k0, int dirs, Jet<C, K> &out) {
    const float vs = V.voxel;
    int gx = __float2int_rd(__fdiv_rn(p.x.v, vs));
    int gy = __float2int_rd(__fdiv_rn(p.y.v, vs));
    int gz = __float2int_rd(__fdiv_rn(p.z.v, vs));
    if (gx <= 0 || gx >= V.rx - 1) return false;
    if (gy <= 0 || gy >= V.ry - 1) return false;
    if (gz <= 0 || gz >= V.rz - 1) return false;
    // g -= (v >= p): the reference's sign trick (:117-122) also steps down on equality
    if (__fmul_rn(__fadd_rn(float(gx), 0.5f), vs) >= p.x.v) gx -= 1;
    if (__fmul_rn(__fadd_rn(float(gy), 0.5f), vs) >= p.y.v) gy -= 1;
    if (__fmul_rn(__fadd_rn(float(gz), 0.5f), vs) >= p.z.v) gz -= 1;
    const float inv_vs = __fdiv_rn(1.f, vs);
    Jet<C, K> a0, b0, c0;
    a0.v = __fdiv_rn(__fmaf_rn(-__fadd_rn(float(gx), 0.5f), vs, p.x.v), vs);
    b0.v = __fdiv_rn(__fmaf_rn(-__fadd_rn(float(gy), 0.5f), vs, p.y.v), vs);
    c0.v = __fdiv_rn(__fmaf_rn(-__fadd_rn(float(gz), 0.5f), vs, p.z.v), vs);
#pragma unroll
    for (int i = 0; i < Jet<C, K>::N; ++i) {
        a0.d[i] = p.x.d[i] * inv_vs;
        b0.d[i] = p.y.d[i] * inv_vs;
        c0.d[i] = p.z.d[i] * inv_vs;
    }
    const Jet<C, K> a1 = jrsubf(1.0f, a0), b1 = jrsubf(1.0f, b0), c1 = jrsubf(1.0f, c0);
    Jet<C, K> r = ((read_tsdf<C, K>(V, gx, gy, gz, k0, dirs) * a1) * b1) * c1;
    r = r + ((read_tsdf<C, K>(V, gx, gy, gz + 1, k0, dirs) * a1) * b1) * c0;
    r = r + ((read_tsdf<C, K>(V, gx, gy + 1, gz, k0, dirs) * a1) * b0) * c1;
    r = r + ((read_tsdf<C, K>(V, gx, gy + 1, gz + 1, k0, dirs) * a1) * b0) * c0;
    r = r + ((read_tsdf<C, K>(V, gx + 1, gy, gz, k0, dirs) * a0) * b1) * c1;
    r = r + ((read_tsdf<C, K>(V, gx + 1, gy, gz + 1, k0, dirs) * a0) * b1) * c0;
    r = r + ((read_tsdf<C, K>(V, gx + 1, gy + 1, gz, k0, dirs) * a0) * b0) * c1;
    r = r + ((read_tsdf<C, K>(V, gx + 1, gy + 1, gz + 1, k0, dirs) * a0) * b0) * c0;
    out = r;
    return true;
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

// Result of one hit evaluation for direction tile [k0,k0+K): world-frame vertex / normal with their derivative
// components, and which of the two the reference would have written (RayCaster.cu:268-271,297-303).
template <int C, int K> struct HitOut {
    Jet3<C, K> vw, ng;
    bool v_ok, n_ok;
};

// Hit evaluation for direction tile [k0,k0+K), RayCaster.cu:249-305.
template <int C, int K>
XS_DEV void eval_hit(const RaycastParams &P, int x, int y, float time_curr, int k0, HitOut<C, K> &out) {
    typedef Jet<C, K> J;
    const VolumeView &V = P.V;
    out.v_ok = out.n_ok = false;
    Jet3<C, K> start, dir;
    ray_setup<C, K>(P, x, y, k0, start, dir);
    const float t1 = __fadd_rn(time_curr, P.time_step);
    Jet3<C, K> p1 = {jfmaf(dir.x, t1, start.x), jfmaf(dir.y, t1, start.y), jfmaf(dir.z, t1, start.z)};
    J Ftdt, Ft;
    if (!trilinear<C, K>(V, p1, k0, P.dirs, Ftdt)) return;
    Jet3<C, K> p0 = {jfmaf(dir.x, time_curr, start.x), jfmaf(dir.y, time_curr, start.y), jfmaf(dir.z, time_curr, start.z)};
    if (!trilinear<C, K>(V, p0, k0, P.dirs, Ft)) return;
    if (isnan(Ftdt.v) || isnan(Ft.v)) return;
    const J coef = Ft / (Ftdt - Ft);
    if (Ft.v < 0.0f || Ftdt.v > 0.0f) return;
    // Ts = time_curr - time_step * coef
    J Ts;
    Ts.v = __fmaf_rn(-coef.v, P.time_step, time_curr);
#pragma unroll
    for (int i = 0; i < J::N; ++i) Ts.d[i] = -P.time_step * coef.d[i];
    const Jet3<C, K> vertex = {start.x + dir.x * Ts, start.y + dir.y * Ts, start.z + dir.z * Ts};
    const JetPose<C, K> v2w = load_pose<C, K>(P.v2w, P.dpose_v2w, k0, P.dirs);
    out.vw = jrot(v2w, vertex) + v2w.t;
    out.v_ok = true;

    const float vs = V.voxel;
    const int gx = __float2int_rd(__fdiv_rn(vertex.x.v, vs));
    const int gy = __float2int_rd(__fdiv_rn(vertex.y.v, vs));
    const int gz = __float2int_rd(__fdiv_rn(vertex.z.v, vs));
    if (!(gx > 1 && gy > 1 && gz > 1 && gx < V.rx - 2 && gy < V.ry - 2 && gz < V.rz - 2)) return;
    const float hv = __fmul_rn(vs, 0.5f);
    Jet3<C, K> t, n;
    J F1, F2;
    bool ok = true;
    t = vertex;
    t.x = jaddf(vertex.x, hv);
    ok &= trilinear<C, K>(V, t, k0, P.dirs, F1);
    t.x = jsubf(vertex.x, hv);
    ok &= trilinear<C, K>(V, t, k0, P.dirs, F2);
    n.x = F1 - F2;
    t = vertex;
    t.y = jaddf(vertex.y, hv);
    ok &= trilinear<C, K>(V, t, k0, P.dirs, F1);
    t.y = jsubf(vertex.y, hv);
    ok &= trilinear<C, K>(V, t, k0, P.dirs, F2);
    n.y = F1 - F2;
    t = vertex;
    t.z = jaddf(vertex.z, hv);
    ok &= trilinear<C, K>(V, t, k0, P.dirs, F1);
    t.z = jsubf(vertex.z, hv);
    ok &= trilinear<C, K>(V, t, k0, P.dirs, F2);
    n.z = F1 - F2;
    if (!ok) return;  // cannot happen for g in (1, N-2); the reference would propagate NaN
    if (jdot(n, n).v == 0.f) return;
    out.ng = jrot(v2w, jnormalized(n));
    out.n_ok = true;
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
            const int ix = __float2int_rd(__fdiv_rn(__fmaf_rn(dx, tt, sx), vs));
            const int iy = __float2int_rd(__fdiv_rn(__fmaf_rn(dy, tt, sy), vs));
            const int iz = __float2int_rd(__fdiv_rn(__fmaf_rn(dz, tt, sz), vs));
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

// ---- pass 2: hit evaluation, one thread per (pixel, direction tile).  A CTA is 32 consecutive pixels of a row x
// 8 direction tiles, so the 8 warps share the real value samples of the same pixels through L1 while each reads its
// own derivative planes.  Every output element is written exactly once (values, or the NaN / 0 fill of :204-205).
template <int C, int K> __global__ void __launch_bounds__(256) raycast_hit_kernel(const RaycastParams P, const float *__restrict__ hit_time) {
    const int x = threadIdx.x + blockIdx.x * 32;
    const int y = blockIdx.y;
    const int tile = threadIdx.y + blockIdx.z * 8;
    const int tiles = P.dirs > 0 ? (P.dirs + K - 1) / K : 1;
    if (x >= P.cols || tile >= tiles) return;
    const int k0 = tile * K;
    HitOut<C, K> o;
    o.v_ok = o.n_ok = false;
    const float time_curr = hit_time[(size_t) y * P.cols + x];
    if (time_curr >= 0.f) eval_hit<C, K>(P, x, y, time_curr, k0, o);
    const float qnan = __int_as_float(0x7fffffff);
    if (k0 == 0) {
        if (o.v_ok)
            store3(P.vmap, 0, P.rows, P.cols, y, x, o.vw.x.v, o.vw.y.v, o.vw.z.v);
        else
            store3(P.vmap, 0, P.rows, P.cols, y, x, qnan, 0.f, 0.f);
        if (o.n_ok)
            store3(P.nmap, 0, P.rows, P.cols, y, x, o.ng.x.v, o.ng.y.v, o.ng.z.v);
        else
            store3(P.nmap, 0, P.rows, P.cols, y, x, qnan, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < Jet<C, K>::N; ++i) {
        if (k0 * C + i < P.V.ncomp) {
            if (o.v_ok)
                store3(P.vmap, 1 + k0 * C + i, P.rows, P.cols, y, x, o.vw.x.d[i], o.vw.y.d[i], o.vw.z.d[i]);
            else
                store3(P.vmap, 1 + k0 * C + i, P.rows, P.cols, y, x, 0.f, 0.f, 0.f);
            if (o.n_ok)
                store3(P.nmap, 1 + k0 * C + i, P.rows, P.cols, y, x, o.ng.x.d[i], o.ng.y.d[i], o.ng.z.d[i]);
            else
                store3(P.nmap, 1 + k0 * C + i, P.rows, P.cols, y, x, 0.f, 0.f, 0.f);
        }
    }
}

// resizeMapKernel, Map.cu:105-152, for packed-SoA maps with derivative components.
template <int C, int K, bool NORMALIZE>
__global__ void resize_map_kernel(int drows, int dcols, int srows, int scols, int dirs, const float *__restrict__ in,
                                  float *__restrict__ out) {
    const int x = threadIdx.x + blockIdx.x * blockDim.x;
    const int y = threadIdx.y + blockIdx.y * blockDim.y;
    if (x >= dcols || y >= drows) return;
    const int ncomp = C * dirs;
    const size_t splane = (size_t) srows * scols;
    const int xs_ = x * 2, ys_ = y * 2;
    const float *p00 = in + (size_t) ys_ * scols + xs_;
    const float qnan = __int_as_float(0x7fffffff);
    const float x00 = p00[0], x01 = p00[1], x10 = p00[scols], x11 = p00[scols + 1];
    if (isnan(x00) || isnan(x01) || isnan(x10) || isnan(x11)) {
        store3(out, 0, drows, dcols, y, x, qnan, 0.f, 0.f);
        for (int q = 0; q < ncomp; ++q) store3(out, 1 + q, drows, dcols, y, x, 0.f, 0.f, 0.f);
        return;
    }
    const int tiles = dirs > 0 ? (dirs + K - 1) / K : 1;
    for (int tile = 0; tile < tiles; ++tile) {
        const int k0 = tile * K;
        Jet<C, K> c[3];
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) {
            const float *p = p00 + pl * splane;
            // (x00 + x01 + x10 + x11) / 4.0f
            c[pl].v = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(p[0], p[1]), p[scols]), p[scols + 1]), 4.0f);
#pragma unroll
            for (int i = 0; i < Jet<C, K>::N; ++i) {
                const int q = k0 * C + i;
                if (q < ncomp) {
                    const float *pd = p + (size_t) (1 + q) * 3 * splane;
                    c[pl].d[i] = (pd[0] + pd[1] + pd[scols] + pd[scols + 1]) * 0.25f;
                } else
                    c[pl].d[i] = 0.f;
            }
        }
        Jet3<C, K> n = {c[0], c[1], c[2]};
        if (NORMALIZE) n = jnormalized(n);
        if (k0 == 0) store3(out, 0, drows, dcols, y, x, n.x.v, n.y.v, n.z.v);
#pragma unroll
        for (int i = 0; i < Jet<C, K>::N; ++i)
            if (k0 * C + i < ncomp) store3(out, 1 + k0 * C + i, drows, dcols, y, x, n.x.d[i], n.y.d[i], n.z.d[i]);
    }
}

int upload_pose_derivs(const xs_volume *v, const xs_pose *p, int slot, cudaStream_t s);

template <bool NORMALIZE>
static int resize_map(const float *d_in, int rows, int cols, int comps, int dirs, float *d_out, void *stream) {
    if (!d_in || !d_out || rows < 2 || cols < 2 || (comps != 1 && comps != 3) || dirs < 0) return XS_ERR_ARG;
    const int drows = rows / 2, dcols = cols / 2;
    dim3 blk(32, 8), grd(div_up(dcols, 32), div_up(drows, 8));
    cudaStream_t s = (cudaStream_t) stream;
    if (comps == 1)
        resize_map_kernel<1, 6, NORMALIZE><<<grd, blk, 0, s>>>(drows, dcols, rows, cols, dirs, d_in, d_out);
    else
        resize_map_kernel<3, 2, NORMALIZE><<<grd, blk, 0, s>>>(drows, dcols, rows, cols, dirs, d_in, d_out);
    XS_LAUNCH_CHECK();
    return XS_OK;
}

}  // namespace xs

using namespace xs;

extern "C" {

int xs_raycast(const xs_volume *v, xs_intr intr, const xs_pose *c2v, const xs_pose *v2w, int rows, int cols,
               float *d_vmap, float *d_nmap, void *stream) {
    if (!v || !c2v || !v2w || !d_vmap || !d_nmap || rows <= 0 || cols <= 0) return XS_ERR_ARG;
    cudaStream_t s = (cudaStream_t) stream;
    XS_CUDA(cudaStreamSynchronize(s));  // staging buffer reuse
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
    if (v->comps == 1) {
        constexpr int K = 3;
        const int tiles = v->dirs > 0 ? div_up(v->dirs, K) : 1;
        dim3 g2(div_up(cols, 32), rows, div_up(tiles, 8));
        raycast_hit_kernel<1, K><<<g2, blk, 0, s>>>(P, v->d_hit_time);
    } else {
        constexpr int K = 1;
        const int tiles = v->dirs > 0 ? div_up(v->dirs, K) : 1;
        dim3 g2(div_up(cols, 32), rows, div_up(tiles, 8));
        raycast_hit_kernel<3, K><<<g2, blk, 0, s>>>(P, v->d_hit_time);
    }
    XS_LAUNCH_CHECK();
    return XS_OK;  // raycast does not sync, RayCaster.cu:367
}

int xs_resize_vmap(const float *d_in, int rows, int cols, int comps, int dirs, float *d_out, void *stream) {
    return resize_map<false>(d_in, rows, cols, comps, dirs, d_out, stream);
}
int xs_resize_nmap(const float *d_in, int rows, int cols, int comps, int dirs, float *d_out, void *stream) {
    return resize_map<true>(d_in, rows, cols, comps, dirs, d_out, stream);
}

}  // extern "C"
