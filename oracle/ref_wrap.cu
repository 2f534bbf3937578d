// oracle/ref_wrap.cu — TEST INFRASTRUCTURE ONLY (checker + reported GPU baseline; never shipped,
// never linked into libxslam_b200.so).
//
// A thin C-ABI shell around the UNMODIFIED reference CUDA operators, compiled from where they lie
// under /root/reference (see oracle/Makefile; outputs go to oracle/_ref/ only).  It exposes
//   (1) stateless per-operator entry points on dense host arrays, used by the stage-wise parity
//       tests (upload -> reference operator -> download), and
//   (2) a restated orchestrator ("ref_kinfu") that drives the reference kernels exactly as
//       XKinectFusion/src/KinectFusionReconstruction.cpp:147-332 does.  The original orchestrator
//       cannot be compiled here (real Eigen, yaml-cpp, OpenCV, Sophus are absent), so its host
//       algebra is restated in oracle/host_algebra.h.  One perturbation direction per instance,
//       as in the reference: the seed is the imaginary part of world2camera
//       (the commented line KinectFusionReconstruction.cpp:22).
#include "CudaFunctions.h"  // reference: Map.h, TsdfFusion.h, RayCaster.h, ICP.h, ExtractPointCloud.h
#include "host_algebra.h"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

using xo::cd;
using xo::cf;

extern "C" {

struct xo_config {
    int res[3];
    float voxel_size;
    int max_weight;
    float thres_range;
    float init_xyz[3];
    float r_deg[3];
    int width, height;
    float fx, fy, cx, cy;
    int num_levels;
    float dist_thres;
    float angle_thres_deg;
    float bi_threshold;
    float trunc_k;
};

}  // extern "C"

namespace {

typedef DeviceArray2D<devComplex> Map;

void up(Map &m, const float *host /* interleaved re,im */, int rows, int cols) {
    m.upload(host, cols * sizeof(devComplex), rows, cols);
}
void down(const Map &m, float *host) { m.download(host, m.cols() * sizeof(devComplex)); }

MatS33 to_mat(const float *p /* 9 x (re,im) row-major */) {
    MatS33 M;
    for (int r = 0; r < 3; ++r) {
        M.data[r].x = devComplex(p[(r * 3 + 0) * 2], p[(r * 3 + 0) * 2 + 1]);
        M.data[r].y = devComplex(p[(r * 3 + 1) * 2], p[(r * 3 + 1) * 2 + 1]);
        M.data[r].z = devComplex(p[(r * 3 + 2) * 2], p[(r * 3 + 2) * 2 + 1]);
    }
    return M;
}
devComplex3 to_vec(const float *p /* 3 x (re,im) */) {
    devComplex3 v;
    v.x = devComplex(p[0], p[1]);
    v.y = devComplex(p[2], p[3]);
    v.z = devComplex(p[4], p[5]);
    return v;
}
MatS33 to_mat(const xo::Mat3c &R) {
    MatS33 M;
    for (int r = 0; r < 3; ++r) {
        M.data[r].x = devComplex(R.m[r][0].real(), R.m[r][0].imag());
        M.data[r].y = devComplex(R.m[r][1].real(), R.m[r][1].imag());
        M.data[r].z = devComplex(R.m[r][2].real(), R.m[r][2].imag());
    }
    return M;
}
devComplex3 to_vec(const xo::Vec3c &t) {
    devComplex3 v;
    v.x = devComplex(t.v[0].real(), t.v[0].imag());
    v.y = devComplex(t.v[1].real(), t.v[1].imag());
    v.z = devComplex(t.v[2].real(), t.v[2].imag());
    return v;
}

struct Timer {
    cudaEvent_t a, b;
    Timer() {
        cudaEventCreate(&a);
        cudaEventCreate(&b);
    }
    ~Timer() {
        cudaEventDestroy(a);
        cudaEventDestroy(b);
    }
    void start() { cudaEventRecord(a); }
    float stop() {
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        return ms;
    }
};

}  // namespace

extern "C" {

// ------------------------------------------------------------------ stateless operators
// Map.cu:262 bilateralFilter
int ref_bilateral(const uint16_t *depth, int rows, int cols, float *out /* rows*cols*(re,im) */) {
    DeviceArray2D<ushort> src;
    src.upload(depth, cols * sizeof(ushort), rows, cols);
    Map dst;
    dst.create(rows, cols);
    bilateralFilter(src, dst);
    cudaDeviceSynchronize();
    down(dst, out);
    return 0;
}
// Map.cu:274 pyrDown
int ref_pyrdown(const float *src_h, int rows, int cols, float *out) {
    Map src, dst;
    up(src, src_h, rows, cols);
    pyrDown(src, dst);
    cudaDeviceSynchronize();
    down(dst, out);
    return 0;
}
// Map.cu:73,89 createVMap / createNMap.  The reference leaves the y/z planes of invalid pixels
// unwritten (Map.cu:27,42,69); the wrapper zero-fills the allocation first so that outputs are
// deterministic.
int ref_vmap_nmap(const float *depth_h, int rows, int cols, float fx, float fy, float cx, float cy, float *vmap_out,
                  float *nmap_out) {
    Map depth, vmap, nmap;
    up(depth, depth_h, rows, cols);
    vmap.create(rows * 3, cols);
    nmap.create(rows * 3, cols);
    cudaMemset2D(vmap.ptr(), vmap.step(), 0, cols * sizeof(devComplex), rows * 3);
    cudaMemset2D(nmap.ptr(), nmap.step(), 0, cols * sizeof(devComplex), rows * 3);
    createVMap(Intr(fx, fy, cx, cy), depth, vmap);
    createNMap(vmap, nmap);
    cudaDeviceSynchronize();
    down(vmap, vmap_out);
    down(nmap, nmap_out);
    return 0;
}
// Map.cu:252,257 resizeVMap / resizeNMap
int ref_resize_map(const float *in_h, int rows /* of one plane */, int cols, int normalize, float *out) {
    Map in, outm;
    up(in, in_h, rows * 3, cols);
    outm.create(rows / 2 * 3, cols / 2);
    cudaMemset2D(outm.ptr(), outm.step(), 0, (cols / 2) * sizeof(devComplex), rows / 2 * 3);
    if (normalize)
        resizeNMap(in, outm);
    else
        resizeVMap(in, outm);
    down(outm, out);
    return 0;
}

// TsdfFusion.cu:173 integrateTsdfVolume on a host-provided volume state (dense x-fastest planes).
int ref_integrate(const uint16_t *depth, int rows, int cols, float fx, float fy, float cx, float cy, int max_weight,
                  const int *res, float voxel, const float *Rv2c /*18*/, const float *tv2c /*6*/, float trunc,
                  float *value, int *weight, float *grad, float threshold, float *ms_out) {
    DeviceArray2D<ushort> d;
    d.upload(depth, cols * sizeof(ushort), rows, cols);
    DeviceArray2D<float> dv, dg, scaled;
    DeviceArray2D<int> dw;
    int R = res[1] * res[2], C = res[0];
    dv.upload(value, C * sizeof(float), R, C);
    dg.upload(grad, C * sizeof(float), R, C);
    dw.upload(weight, C * sizeof(int), R, C);
    int3 r3 = make_int3(res[0], res[1], res[2]);
    devComplex3 tc2v = to_vec(tv2c);  // unused by the reference kernel (TsdfFusion.cu:173-179)
    Timer t;
    t.start();
    integrateTsdfVolume(d, Intr(fx, fy, cx, cy), max_weight, r3, voxel, to_mat(Rv2c), to_vec(tv2c), tc2v, trunc, dv,
                        dw, dg, scaled, 0, threshold, 0.f);
    float ms = t.stop();
    if (ms_out) *ms_out = ms;
    dv.download(value, C * sizeof(float));
    dg.download(grad, C * sizeof(float));
    dw.download(weight, C * sizeof(int));
    return 0;
}

// ExtractPointCloud.cu:181-210 extractPoints + :342-362 extractNormals on a host-provided volume, exactly as
// KinectFusionReconstruction::ExportPointCloud calls them (KinectFusionReconstruction.cpp:334-346): normals are computed
// for the whole buffer (points.size = max_points), only the first `count` entries are meaningful.  Returns the count.
long ref_extract(const int *res, float voxel, const float *value, const int *weight, const float *grad, long max_points,
                 float *points_out /* [max_points][3] */, float *normals_out /* [max_points][3] */) {
    DeviceArray2D<float> dv, dg;
    DeviceArray2D<int> dw;
    int R = res[1] * res[2], C = res[0];
    dv.upload(value, C * sizeof(float), R, C);
    dg.upload(grad, C * sizeof(float), R, C);
    dw.upload(weight, C * sizeof(int), R, C);
    DeviceArray<float3> cloud, normals;
    cloud.create(max_points);
    normals.create(max_points);
    cudaMemset(cloud.ptr(), 0, max_points * sizeof(float3));
    int3 r3 = make_int3(res[0], res[1], res[2]);
    const size_t n = extractPoints(dv, dw, dg, r3, voxel, cloud);
    extractNormals(dv, dw, dg, r3, voxel, cloud, normals);
    cudaMemcpy(points_out, cloud.ptr(), n * sizeof(float3), cudaMemcpyDeviceToHost);
    cudaMemcpy(normals_out, normals.ptr(), n * sizeof(float3), cudaMemcpyDeviceToHost);
    return (long) n;
}

// RayCaster.cu:327 raycast on a host-provided volume.
int ref_raycast(float fx, float fy, float cx, float cy, const float *Rc2v, const float *tc2v, const float *Rv2w,
                const float *tv2w, float trunc, const int *res, float voxel, const float *value, const float *grad,
                int rows, int cols, float *vmap_out, float *nmap_out, float *ms_out) {
    DeviceArray2D<float> dv, dg;
    int R = res[1] * res[2], C = res[0];
    dv.upload(value, C * sizeof(float), R, C);
    dg.upload(grad, C * sizeof(float), R, C);
    Map vmap, nmap;
    vmap.create(rows * 3, cols);
    nmap.create(rows * 3, cols);
    cudaMemset2D(vmap.ptr(), vmap.step(), 0, cols * sizeof(devComplex), rows * 3);
    cudaMemset2D(nmap.ptr(), nmap.step(), 0, cols * sizeof(devComplex), rows * 3);
    int3 r3 = make_int3(res[0], res[1], res[2]);
    Timer t;
    t.start();
    raycast(Intr(fx, fy, cx, cy), to_mat(Rc2v), to_vec(tc2v), to_mat(Rv2w), to_vec(tv2w), trunc, r3, voxel, dv, dg,
            vmap, nmap);
    float ms = t.stop();
    if (ms_out) *ms_out = ms;
    down(vmap, vmap_out);
    down(nmap, nmap_out);
    return 0;
}

// ICP.cu:365 estimateCombined.  A: column-major 6x6 (re,im) doubles; b: 6 (re,im) doubles.
int ref_estimate_combined(const float *Rcurr, const float *tcurr, const float *vmap_curr, const float *nmap_curr,
                          const float *Rprev_inv, const float *tprev, float fx, float fy, float cx, float cy,
                          const float *vmap_g_prev, const float *nmap_g_prev, int rows, int cols, float dist_thres,
                          float angle_thres, double *A_out /*72*/, double *b_out /*12*/, float *ms_out) {
    Map vc, nc, vp, np;
    up(vc, vmap_curr, rows * 3, cols);
    up(nc, nmap_curr, rows * 3, cols);
    up(vp, vmap_g_prev, rows * 3, cols);
    up(np, nmap_g_prev, rows * 3, cols);
    DeviceArray2D<devComplexICP> gbuf;
    DeviceArray<devComplexICP> mbuf;
    hostComplexICP A[36], b[6];
    Timer t;
    t.start();
    estimateCombined(to_mat(Rcurr), to_vec(tcurr), vc, nc, to_mat(Rprev_inv), to_vec(tprev), Intr(fx, fy, cx, cy), vp,
                     np, dist_thres, angle_thres, gbuf, mbuf, A, b);
    float ms = t.stop();
    if (ms_out) *ms_out = ms;
    memcpy(A_out, A, sizeof(A));
    memcpy(b_out, b, sizeof(b));
    return 0;
}

// ICP.cu:431 computeOptimizeMatrix (dormant in the reference's frame loop).  J: [3][4] row-major; H: [12][12].
int ref_compute_optimize_matrix(const float *Rcurr, const float *tcurr, const float *vmap_curr, const float *nmap_curr,
                                const float *Rprev_inv, const float *tprev, float fx, float fy, float cx, float cy,
                                const float *vmap_g_prev, const float *nmap_g_prev, int rows, int cols, float dist_thres,
                                float angle_thres, float *J_out /*12*/, float *H_out /*144*/) {
    Map vc, nc, vp, np;
    up(vc, vmap_curr, rows * 3, cols);
    up(nc, nmap_curr, rows * 3, cols);
    up(vp, vmap_g_prev, rows * 3, cols);
    up(np, nmap_g_prev, rows * 3, cols);
    DeviceArray2D<float> jacobi_buf, hessian_buf[12];
    Eigen::Matrix4f jacobi_host, hessian_store[3][4], *hessian_rows[3] = {hessian_store[0], hessian_store[1], hessian_store[2]};
    jacobi_host.setZero();
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) hessian_store[i][j].setZero();
    computeOptimizeMatrix(vc, nc, vp, np, to_mat(Rcurr), to_vec(tcurr), to_mat(Rprev_inv), to_vec(tprev), Intr(fx, fy, cx, cy),
                          dist_thres, angle_thres, jacobi_buf, jacobi_host, hessian_buf, hessian_rows);
    cudaDeviceSynchronize();
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) J_out[i * 4 + j] = jacobi_host(i, j);
    for (int a = 0; a < 12; ++a)
        for (int b = 0; b < 12; ++b) H_out[a * 12 + b] = hessian_store[a / 4][a % 4](b / 4, b % 4);
    return 0;
}

// TsdfFusion.cu:286 ComputeLocalTsdf_hessian.  Rv2c: 9 x (re.re, re.im, im.re, im.im); tv2c: 3 x 4.
int ref_tsdf_hessian(const uint16_t *depth, int rows, int cols, float fx, float fy, float cx, float cy, const int *res,
                     float voxel, const float *Rv2c /*36*/, const float *tv2c /*12*/, float trunc, const float *gt,
                     float *out4, float *ms_out) {
    DeviceArray2D<ushort> d;
    d.upload(depth, cols * sizeof(ushort), rows, cols);
    DeviceArray2D<float> scaled;
    size_t n = (size_t) res[0] * res[1] * res[2];
    thrustDvec<float> gt_vec(gt, gt + n), real_vec, grad_vec, hess_vec;
    thrustDvec<int> count_vec;
    MatD33 R;
    devDComplex3 t;
    devDComplex *Rp = &R.data[0].x;
    for (int i = 0; i < 9; ++i) Rp[i] = devDComplex(Rv2c[4 * i], Rv2c[4 * i + 1], Rv2c[4 * i + 2], Rv2c[4 * i + 3]);
    devDComplex *tp = &t.x;
    for (int i = 0; i < 3; ++i) tp[i] = devDComplex(tv2c[4 * i], tv2c[4 * i + 1], tv2c[4 * i + 2], tv2c[4 * i + 3]);
    int3 r3 = make_int3(res[0], res[1], res[2]);
    Timer tm;
    tm.start();
    float4 r = ComputeLocalTsdf_hessian(d, Intr(fx, fy, cx, cy), scaled, r3, voxel, R, t, trunc, 0.f, 0.f, gt_vec,
                                        real_vec, grad_vec, hess_vec, count_vec);
    float ms = tm.stop();
    if (ms_out) *ms_out = ms;
    out4[0] = r.x;
    out4[1] = r.y;
    out4[2] = r.z;
    out4[3] = r.w;
    return 0;
}

// TsdfFusion.cu:409 ComputeLocalTsdf_loss (real-only).  Rv2c: 9 floats row-major; tv2c: 3 floats.
int ref_tsdf_loss(const uint16_t *depth, int rows, int cols, float fx, float fy, float cx, float cy, const int *res, float voxel,
                  const float *Rv2c, const float *tv2c, float trunc, const float *gt, float *out2, float *ms_out) {
    DeviceArray2D<ushort> d;
    d.upload(depth, cols * sizeof(ushort), rows, cols);
    DeviceArray2D<float> scaled;
    size_t n = (size_t) res[0] * res[1] * res[2];
    thrustDvec<float> gt_vec(gt, gt + n), real_vec;
    thrustDvec<int> count_vec;
    Mat33 R;
    for (int i = 0; i < 3; ++i) R.data[i] = make_float3(Rv2c[3 * i], Rv2c[3 * i + 1], Rv2c[3 * i + 2]);
    float3 t = make_float3(tv2c[0], tv2c[1], tv2c[2]);
    int3 r3 = make_int3(res[0], res[1], res[2]);
    Timer tm;
    tm.start();
    float2 r = ComputeLocalTsdf_loss(d, Intr(fx, fy, cx, cy), scaled, r3, voxel, R, t, trunc, 0.f, 0.f, gt_vec, real_vec, count_vec);
    float ms = tm.stop();
    if (ms_out) *ms_out = ms;
    out2[0] = r.x;
    out2[1] = r.y;
    return 0;
}

// ------------------------------------------------------------------ restated orchestrator
struct ref_kinfu {
    xo_config cfg;
    Intr intr;
    xo::Mat4c world2camera, world2volume;
    std::vector<xo::Mat4c> record;
    int frame_id = 0;
    int icp_iterations[3] = {5, 4, 3};
    float angle_thres;
    TsdfVolume *volume = nullptr;
    std::vector<Map> depths, vmaps_curr, nmaps_curr, vmaps_prev, nmaps_prev;
    DeviceArray2D<devComplexICP> g_buf;
    DeviceArray<devComplexICP> sum_buf;
    DeviceArray2D<float> depth_scaled;
    DeviceArray2D<ushort> depth_d;
    std::vector<double> icp_log;  // per iteration: A (72 doubles) + b (12 doubles)
    float ms[5] = {0, 0, 0, 0, 0};  // surface, icp, integrate, raycast(+resize), total
};

ref_kinfu *ref_kinfu_create(const xo_config *cfg, const float *seed_w2c_imag /* 16 floats or NULL */) {
    ref_kinfu *k = new ref_kinfu;
    k->cfg = *cfg;
    k->intr = Intr(cfg->fx, cfg->fy, cfg->cx, cfg->cy);
    // KinectFusionReconstruction.cpp:21-38
    k->world2camera = xo::Mat4c::identity();
    if (seed_w2c_imag)
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) k->world2camera.m[i][j] = cf(k->world2camera.m[i][j].real(), seed_w2c_imag[i * 4 + j]);
    k->record.push_back(k->world2camera);
    k->world2volume = xo::Mat4c::identity();
    {
        // (Rx * Ry * Rz).matrix() in real float; identity for the reference config (r_* = 0).
        float a[3];
        for (int i = 0; i < 3; ++i) a[i] = cfg->r_deg[i] / 180.0f * float(M_PI);
        xo::Mat3c Rx = xo::angle_axis_matrix(cf(a[0], 0.f), 0), Ry = xo::angle_axis_matrix(cf(a[1], 0.f), 1),
                  Rz = xo::angle_axis_matrix(cf(a[2], 0.f), 2);
        xo::Mat3c R = xo::mul(xo::mul(Rx, Ry), Rz);
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) k->world2volume.m[i][j] = cf(R.m[i][j].real(), 0.f);
            k->world2volume.m[i][3] = cf(cfg->init_xyz[i], 0.f);
        }
    }
    k->angle_thres = float(sin(cfg->angle_thres_deg / 180.f * M_PI));  // :58
    int L = cfg->num_levels;
    k->depths.resize(L);
    k->vmaps_curr.resize(L);
    k->nmaps_curr.resize(L);
    k->vmaps_prev.resize(L);
    k->nmaps_prev.resize(L);
    for (int i = 0; i < L; ++i) {  // :84-92
        int r = cfg->height >> i, c = cfg->width >> i;
        k->depths[i].create(r, c);
        k->vmaps_curr[i].create(r * 3, c);
        k->nmaps_curr[i].create(r * 3, c);
        k->vmaps_prev[i].create(r * 3, c);
        k->nmaps_prev[i].create(r * 3, c);
        // deterministic contents for never-written planes (see ref_vmap_nmap)
        cudaMemset2D(k->vmaps_curr[i].ptr(), k->vmaps_curr[i].step(), 0, c * sizeof(devComplex), r * 3);
        cudaMemset2D(k->nmaps_curr[i].ptr(), k->nmaps_curr[i].step(), 0, c * sizeof(devComplex), r * 3);
        cudaMemset2D(k->vmaps_prev[i].ptr(), k->vmaps_prev[i].step(), 0, c * sizeof(devComplex), r * 3);
        cudaMemset2D(k->nmaps_prev[i].ptr(), k->nmaps_prev[i].step(), 0, c * sizeof(devComplex), r * 3);
    }
    k->g_buf.create(27, 20 * 60);
    k->sum_buf.create(27);
    Eigen::Vector3i res;
    res(0) = cfg->res[0];
    res(1) = cfg->res[1];
    res(2) = cfg->res[2];
    k->volume = new TsdfVolume(res, cfg->voxel_size, cfg->thres_range);
    return k;
}

void ref_kinfu_destroy(ref_kinfu *k) {
    if (!k) return;
    delete k->volume;
    delete k;
}

static int ref_pose_estimate(ref_kinfu *k) {  // KinectFusionReconstruction.cpp:161-235
    if (k->frame_id == 0) return 0;
    xo::Mat4c c2w_prev = xo::inverse(k->record.back());
    xo::Mat3c Rprev = xo::rotation_of(c2w_prev);
    xo::Vec3c tprev = xo::translation_of(c2w_prev);
    xo::Mat3c Rprev_inv = xo::inverse(Rprev);
    xo::Mat3c Rcurr = Rprev;
    xo::Vec3c tcurr = tprev;
    xo::Mat4c c2w_curr = c2w_prev;
    for (int level = k->cfg.num_levels - 1; level >= 0; --level) {
        for (int iter = 0; iter < k->icp_iterations[level]; ++iter) {
            hostComplexICP A[36], b[6];
            estimateCombined(to_mat(Rcurr), to_vec(tcurr), k->vmaps_curr[level], k->nmaps_curr[level], to_mat(Rprev_inv),
                             to_vec(tprev), k->intr(level), k->vmaps_prev[level], k->nmaps_prev[level],
                             k->cfg.dist_thres, k->angle_thres, k->g_buf, k->sum_buf, A, b);
            const double *Ap = reinterpret_cast<const double *>(A), *bp = reinterpret_cast<const double *>(b);
            k->icp_log.insert(k->icp_log.end(), Ap, Ap + 72);
            k->icp_log.insert(k->icp_log.end(), bp, bp + 12);
            double det = xo::det6_real(A);
            if (fabs(det) < 1e-15 || std::isnan(det)) return 0;
            cd x[6];
            xo::llt_solve6(A, b, x);
            xo::pose_update(x, Rcurr, tcurr);
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) c2w_curr.m[i][j] = Rcurr.m[i][j];
                c2w_curr.m[i][3] = tcurr.v[i];
            }
            c2w_curr.m[3][3] = cf(1.f, 0.f);
        }
    }
    k->world2camera = xo::inverse(c2w_curr);
    k->record.push_back(k->world2camera);
    return 1;
}

// ProcessFrame, KinectFusionReconstruction.cpp:147-159.  depth: host uint16 mm, dense.
int ref_kinfu_process_frame(ref_kinfu *k, const uint16_t *depth) {
    const xo_config &c = k->cfg;
    k->depth_d.upload(depth, c.width * sizeof(ushort), c.height, c.width);
    Timer total, t;
    total.start();
    // SurfaceMeasure :280-299
    t.start();
    bilateralFilter(k->depth_d, k->depths[0]);
    for (int i = 1; i < c.num_levels; ++i) pyrDown(k->depths[i - 1], k->depths[i]);
    for (int i = 0; i < c.num_levels; ++i) {
        createVMap(k->intr(i), k->depths[i], k->vmaps_curr[i]);
        createNMap(k->vmaps_curr[i], k->nmaps_curr[i]);
    }
    k->ms[0] = t.stop();
    t.start();
    int ok = ref_pose_estimate(k);
    k->ms[1] = t.stop();
    if (k->frame_id > 0 && !ok) return 0;
    // IntegrateFrame :237-278
    xo::Mat4c c2w = xo::inverse(k->record.back());
    xo::Mat4c c2v = xo::mul(k->world2volume, c2w);
    xo::Mat4c v2c = xo::inverse(c2v);
    int3 res = make_int3(c.res[0], c.res[1], c.res[2]);
    t.start();
    integrateTsdfVolume(k->depth_d, k->intr, c.max_weight, res, c.voxel_size, to_mat(xo::rotation_of(v2c)),
                        to_vec(xo::translation_of(v2c)), to_vec(xo::translation_of(c2v)), k->volume->getTsdfTruncDist(),
                        k->volume->value(), k->volume->weight(), k->volume->grad(), k->depth_scaled, k->frame_id,
                        c.bi_threshold, c.trunc_k);
    k->ms[2] = t.stop();
    // CalculatePointCloud :302-332 (uses world2camera, which equals record.back())
    t.start();
    xo::Mat4c c2w2 = xo::inverse(k->world2camera);
    xo::Mat4c c2v2 = xo::mul(k->world2volume, c2w2);
    xo::Mat4c v2w = xo::inverse(k->world2volume);
    raycast(k->intr, to_mat(xo::rotation_of(c2v2)), to_vec(xo::translation_of(c2v2)), to_mat(xo::rotation_of(v2w)),
            to_vec(xo::translation_of(v2w)), k->volume->getTsdfTruncDist(), res, c.voxel_size, k->volume->value(),
            k->volume->grad(), k->vmaps_prev[0], k->nmaps_prev[0]);
    for (int i = 1; i < c.num_levels; ++i) {
        resizeVMap(k->vmaps_prev[i - 1], k->vmaps_prev[i]);
        resizeNMap(k->nmaps_prev[i - 1], k->nmaps_prev[i]);
    }
    k->ms[3] = t.stop();
    k->ms[4] = total.stop();
    k->frame_id += 1;
    return 1;
}

void ref_kinfu_get_pose(const ref_kinfu *k, float *w2c_out /* 16 x (re,im) row-major */) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            w2c_out[(i * 4 + j) * 2] = k->world2camera.m[i][j].real();
            w2c_out[(i * 4 + j) * 2 + 1] = k->world2camera.m[i][j].imag();
        }
}
void ref_kinfu_get_times(const ref_kinfu *k, float *ms5) { memcpy(ms5, k->ms, sizeof(k->ms)); }
float ref_kinfu_trunc(const ref_kinfu *k) { return k->volume->getTsdfTruncDist(); }

// which: 0 depth pyramid level L (rows*cols complex), 1 vmap_curr, 2 nmap_curr, 3 vmap_g_prev, 4 nmap_g_prev
int ref_kinfu_get_map(const ref_kinfu *k, int which, int level, float *out) {
    const Map *m = nullptr;
    switch (which) {
        case 0: m = &k->depths[level]; break;
        case 1: m = &k->vmaps_curr[level]; break;
        case 2: m = &k->nmaps_curr[level]; break;
        case 3: m = &k->vmaps_prev[level]; break;
        case 4: m = &k->nmaps_prev[level]; break;
        default: return -1;
    }
    down(*m, out);
    return 0;
}
int ref_kinfu_get_volume(const ref_kinfu *k, float *value, int *weight, float *grad) {
    int C = k->cfg.res[0];
    if (value) k->volume->value().download(value, C * sizeof(float));
    if (weight) k->volume->weight().download(weight, C * sizeof(int));
    if (grad) k->volume->grad().download(grad, C * sizeof(float));
    return 0;
}
// ICP log: n iterations x 84 doubles (A column-major 6x6 complex, then b); returns count, clears on read.
int ref_kinfu_take_icp_log(ref_kinfu *k, double *out, int max_iters) {
    int n = (int) (k->icp_log.size() / 84);
    if (n > max_iters) n = max_iters;
    if (out) memcpy(out, k->icp_log.data(), (size_t) n * 84 * sizeof(double));
    k->icp_log.clear();
    return n;
}

}  // extern "C"
