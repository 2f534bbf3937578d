// xslam_b200.hpp — C++ wrappers with the REFERENCE's operator signatures over the C-ABI of libxslam_b200.so.
//
// The reference's seam is the set of free functions declared in XKinectFusion/include/{Map,TsdfFusion,RayCaster,ICP}.h
// (aggregated by CudaFunctions.h:4-8) and called only from KinectFusionReconstruction.cpp.  A maintainer switches the
// frame loop to this library by including this header instead of CudaFunctions.h and adding
// `using namespace xslam_b200::seam;` — the calls in KinectFusionReconstruction.cpp then compile unchanged
// (INTEGRATION.md shows the diff).  The wrappers are templates over the reference's own container / POD types
// (DeviceArray2D<T>, PtrStep<T>, PtrStepSz<T>, Intr, MatS33, devComplex3 — Internal.h, device_array.hpp,
// kernel_containers.hpp), so this header does not include any reference header; oracle/hpp_check.cu instantiates
// every wrapper with the real reference types as a compile check, and oracle/seam_run.cu links and RUNS a frame through
// them on the reference's own containers beside the same calls into the reference's kernels (tests/test_seam_run.py).
//
// Layout at the seam: maps are pitched interleaved complex<float>, three planes stacked by rows;
// volumes are three pitched (Y*Z) x X planes.  Behind the seam everything is packed SoA / brick-tiled, so each
// wrapper converts at the boundary (xs_map_complex_to_soa / xs_map_soa_to_complex, xs_volume_import/export_planes).
// One perturbation direction is carried — the reference's imaginary part — i.e. comps = 1, dirs = 1.  Code that wants
// k directions per pass uses the C-ABI (or xs_kinfu) directly.  Errors follow the reference: print and exit(-1)
// (cudaSafeCall, Common/include/cx.h:124-130).
#pragma once
#include "xslam_b200.h"

#include <cuda_runtime_api.h>

#include <complex>
#include <cstdio>
#include <cstdlib>
#include <map>

namespace xslam_b200 {

inline void check(int rc, const char *what) {
    if (rc < 0) {
        std::fprintf(stderr, "CUDA error(%s): %s\n", what, xs_last_error());
        std::exit(-1);
    }
}

// scratch SoA buffers, grown on demand and reused (the caller owns every seam buffer; these are library-side temporaries)
struct Scratch {
    float *p = nullptr;
    size_t cap = 0;
    float *get(size_t floats) {
        if (floats > cap) {
            cudaFree(p);
            if (cudaMalloc((void **) &p, floats * sizeof(float)) != cudaSuccess) check(XS_ERR_CUDA, "scratch");
            cap = floats;
        }
        return p;
    }
};
inline Scratch &scratch(int slot) {
    static Scratch s[8];
    return s[slot];
}

template <class IntrT> inline xs_intr to_intr(const IntrT &k) { return xs_intr{k.fx, k.fy, k.cx, k.cy}; }

// The reference's call sites pass either the owning containers (DeviceArray2D<T>: ptr(), step(), rows(), cols(); DeviceArray<T>:
// ptr(), size()) or their kernel views (PtrStep / PtrStepSz / PtrSz: data, step, rows, cols, size members) - the containers
// convert implicitly to the views in the reference's non-template signatures (device_array.hpp:96-98,225-231).  These
// adapters accept both, so that the template wrappers below bind to unchanged call sites.
struct Plane {
    void *data;
    size_t step;
    int rows, cols;
};
template <class T> inline auto as_plane(const T &p, int) -> decltype(p.step(), Plane()) { return Plane{(void *) p.ptr(), p.step(), p.rows(), p.cols()}; }
template <class T> inline auto as_plane_dims(const T &p, int) -> decltype(p.rows + 0, Plane()) { return Plane{(void *) p.data, p.step, p.rows, p.cols}; }
template <class T> inline Plane as_plane_dims(const T &p, long) { return Plane{(void *) p.data, p.step, 0, 0}; }
template <class T> inline auto as_plane(const T &p, long) -> decltype(p.step + 0, Plane()) { return as_plane_dims(p, 0); }
struct Linear {
    void *data;
    size_t size;
};
template <class T> inline auto as_linear(const T &p, int) -> decltype(p.size(), Linear()) { return Linear{(void *) p.ptr(), p.size()}; }
template <class T> inline auto as_linear(const T &p, long) -> decltype(p.size + 0, Linear()) { return Linear{(void *) p.data, p.size}; }

// MatS33 (Internal.h:146-148: data[3] rows of devComplex3) + devComplex3 -> xs_pose with one derivative component
struct PoseHolder {
    xs_pose p;
    float dR[9], dt[3];
    template <class Mat, class Vec> PoseHolder(const Mat &R, const Vec &t) {
        const auto *rows = R.data;
        for (int i = 0; i < 3; ++i) {
            const std::complex<float> e[3] = {{rows[i].x.real(), rows[i].x.imag()}, {rows[i].y.real(), rows[i].y.imag()}, {rows[i].z.real(), rows[i].z.imag()}};
            for (int j = 0; j < 3; ++j) {
                p.R[i * 3 + j] = e[j].real();
                dR[i * 3 + j] = e[j].imag();
            }
        }
        p.t[0] = t.x.real(), p.t[1] = t.y.real(), p.t[2] = t.z.real();
        dt[0] = t.x.imag(), dt[1] = t.y.imag(), dt[2] = t.z.imag();
        p.ncomp = 1;
        p.dR = dR;
        p.dt = dt;
    }
};

// import a seam map ([nplanes*rows] x cols interleaved complex) into scratch slot `slot` as SoA [(1+ncomp)][nplanes][rows][cols]
template <class MapT> inline float *import_map(const MapT &m, int nplanes, int ncomp, int slot) {
    const int rows = m.rows() / nplanes, cols = m.cols();
    float *soa = scratch(slot).get((size_t) (1 + ncomp) * nplanes * rows * cols);
    check(xs_map_complex_to_soa(m.ptr(), m.step(), nplanes, rows, cols, soa, ncomp, ncomp ? 0 : -1, nullptr), "import_map");
    return soa;
}
template <class MapT> inline void export_map(const float *soa, int ncomp, int nplanes, int rows, int cols, MapT &m) {
    m.create(nplanes * rows, cols);
    check(xs_map_soa_to_complex(soa, ncomp, ncomp ? 0 : -1, nplanes, rows, cols, m.ptr(), m.step(), nullptr), "export_map");
}

namespace seam {

// Map.h:16 / Map.cu:262 (no sync, like the reference)
template <class DepthT, class MapT> void bilateralFilter(const DepthT &src, MapT &dst) {
    float *out = scratch(0).get((size_t) src.rows() * src.cols());
    check(xs_bilateral_filter(src.ptr(), src.step(), src.rows(), src.cols(), out, nullptr), "bilateralFilter");
    export_map(out, 0, 1, src.rows(), src.cols(), dst);
}
// Map.h:22 / Map.cu:274
template <class MapT> void pyrDown(const MapT &src, MapT &dst) {
    float *in = import_map(src, 1, 0, 0);
    float *out = scratch(1).get((size_t) (src.rows() / 2) * (src.cols() / 2));
    check(xs_pyr_down(in, src.rows(), src.cols(), out, nullptr), "pyrDown");
    export_map(out, 0, 1, src.rows() / 2, src.cols() / 2, dst);
}
// Map.h:29 / Map.cu:73
template <class IntrT, class MapT> void createVMap(const IntrT &intr, const MapT &depth, MapT &vmap) {
    float *in = import_map(depth, 1, 0, 0);
    float *out = scratch(1).get((size_t) 3 * depth.rows() * depth.cols());
    check(xs_create_vmap(to_intr(intr), in, depth.rows(), depth.cols(), out, nullptr), "createVMap");
    export_map(out, 0, 3, depth.rows(), depth.cols(), vmap);
}
// Map.h:35 / Map.cu:89
template <class MapT> void createNMap(const MapT &vmap, MapT &nmap) {
    const int rows = vmap.rows() / 3, cols = vmap.cols();
    float *in = import_map(vmap, 3, 0, 0);
    float *out = scratch(1).get((size_t) 3 * rows * cols);
    check(xs_create_nmap(in, rows, cols, out, nullptr), "createNMap");
    export_map(out, 0, 3, rows, cols, nmap);
}
// Map.h:47,54 / Map.cu:252,257 (sync, like the reference)
template <class MapT> void resizeVMap(const MapT &input, MapT &output) {
    const int rows = input.rows() / 3, cols = input.cols();
    float *in = import_map(input, 3, 1, 0);
    float *out = scratch(1).get((size_t) 2 * 3 * (rows / 2) * (cols / 2));
    check(xs_resize_vmap(in, rows, cols, 1, 1, out, nullptr), "resizeVMap");
    export_map(out, 1, 3, rows / 2, cols / 2, output);
    cudaDeviceSynchronize();
}
template <class MapT> void resizeNMap(const MapT &input, MapT &output) {
    const int rows = input.rows() / 3, cols = input.cols();
    float *in = import_map(input, 3, 1, 0);
    float *out = scratch(1).get((size_t) 2 * 3 * (rows / 2) * (cols / 2));
    check(xs_resize_nmap(in, rows, cols, 1, 1, out, nullptr), "resizeNMap");
    export_map(out, 1, 3, rows / 2, cols / 2, output);
    cudaDeviceSynchronize();
}

// The brick-tiled volume that shadows one reference TsdfVolume (keyed by the address of its value plane).  A maintainer
// would instead hold the xs_volume inside class TsdfVolume; this cache keeps the call sites unchanged.
struct VolumeShadow {
    xs_volume *v = nullptr;
    float *dense = nullptr;  // 3 dense [z][y][x] planes: value, weight, grad
    int res[3] = {0, 0, 0};
};
inline std::map<const void *, VolumeShadow> &shadow_cache() {
    static std::map<const void *, VolumeShadow> cache;
    return cache;
}
inline VolumeShadow &shadow_of(const void *value_plane, const int res[3], float voxel, float trunc) {
    VolumeShadow &s = shadow_cache()[value_plane];
    if (!s.v) {
        // trunc = max(voxel * thres_range, 2.1 voxel) (TsdfVolume.cpp:25,37): recover thres_range from the caller's trunc
        s.v = xs_volume_create(res, voxel, trunc / voxel, 1, 1);
        if (!s.v) check(XS_ERR_CUDA, "xs_volume_create");
        const size_t n = (size_t) res[0] * res[1] * res[2];
        if (cudaMalloc((void **) &s.dense, 3 * n * sizeof(float)) != cudaSuccess) check(XS_ERR_CUDA, "volume shadow");
        for (int i = 0; i < 3; ++i) s.res[i] = res[i];
    }
    return s;
}
// pitched (Y*Z) x X seam planes <-> dense planes <-> bricks
inline void volume_pull(VolumeShadow &s, const Plane &value, const Plane &weight, const Plane &grad) {
    const size_t n = (size_t) s.res[0] * s.res[1] * s.res[2], row = (size_t) s.res[0] * sizeof(float);
    const size_t h = (size_t) s.res[1] * s.res[2];
    cudaMemcpy2D(s.dense, row, value.data, value.step, row, h, cudaMemcpyDeviceToDevice);
    cudaMemcpy2D(s.dense + n, row, weight.data, weight.step, row, h, cudaMemcpyDeviceToDevice);
    cudaMemcpy2D(s.dense + 2 * n, row, grad.data, grad.step, row, h, cudaMemcpyDeviceToDevice);
    check(xs_volume_import_planes(s.v, 0, s.dense, (const int *) (s.dense + n), s.dense + 2 * n, nullptr), "volume import");
}
inline void volume_push(VolumeShadow &s, const Plane &value, const Plane &weight, const Plane &grad) {
    const size_t n = (size_t) s.res[0] * s.res[1] * s.res[2], row = (size_t) s.res[0] * sizeof(float);
    const size_t h = (size_t) s.res[1] * s.res[2];
    check(xs_volume_export_planes(s.v, 0, s.dense, (int *) (s.dense + n), s.dense + 2 * n, nullptr), "volume export");
    cudaMemcpy2D((void *) value.data, value.step, s.dense, row, row, h, cudaMemcpyDeviceToDevice);
    cudaMemcpy2D((void *) weight.data, weight.step, s.dense + n, row, row, h, cudaMemcpyDeviceToDevice);
    cudaMemcpy2D((void *) grad.data, grad.step, s.dense + 2 * n, row, row, h, cudaMemcpyDeviceToDevice);
}

// TsdfVolume.h:16 / TsdfFusion.cu:34-43 (sync): value, weight and grad planes are zeroed (pack_tsdf(0, 0)); like the reference's
// kernel, the packed short2 `volume` plane is not touched.  The brick-tiled shadow of this volume, if one exists, is reset.
template <class PS, class PV, class PW, class Int3>
void initVolume(const PS & /*volume*/, const PV &value_volume, const PW &weight_volume, const PV &grad_volume,
                const Int3 &volume_resolution) {
    const size_t h = (size_t) volume_resolution.y * volume_resolution.z, row = (size_t) volume_resolution.x * sizeof(float);
    const Plane pv = as_plane(value_volume, 0), pw = as_plane(weight_volume, 0), pg = as_plane(grad_volume, 0);
    if (cudaMemset2D(pv.data, pv.step, 0, row, h) != cudaSuccess || cudaMemset2D(pw.data, pw.step, 0, row, h) != cudaSuccess ||
        cudaMemset2D(pg.data, pg.step, 0, row, h) != cudaSuccess)
        check(XS_ERR_CUDA, "initVolume");
    auto it = shadow_cache().find((const void *) pv.data);
    if (it != shadow_cache().end() && it->second.v) check(xs_volume_reset(it->second.v, nullptr), "initVolume");
    cudaDeviceSynchronize();
}

// TsdfFusion.h:40-45 / TsdfFusion.cu:173 (sync).  tc2v, depthScaled, frame_id and k are unused by the reference too.
template <class DepthT, class IntrT, class Int3, class Mat, class Vec, class PV, class PW, class ScaledT>
void integrateTsdfVolume(const DepthT &depth, const IntrT &intr, int max_weight, const Int3 &volume_size, float voxel_size,
                         const Mat &Rv2c, const Vec &tv2c, const Vec & /*tc2v*/, float trunc_dist, const PV &value_volume,
                         const PW &weight_volume, const PV &grad_volume, ScaledT & /*depthScaled*/, int /*frame_id*/, float threshold,
                         float /*k*/) {
    const int res[3] = {volume_size.x, volume_size.y, volume_size.z};
    const Plane pv = as_plane(value_volume, 0), pw = as_plane(weight_volume, 0), pg = as_plane(grad_volume, 0), d = as_plane(depth, 0);
    VolumeShadow &s = shadow_of(pv.data, res, voxel_size, trunc_dist);
    volume_pull(s, pv, pw, pg);
    PoseHolder v2c(Rv2c, tv2c);
    check(xs_integrate(s.v, (const uint16_t *) d.data, d.step, d.rows, d.cols, to_intr(intr), max_weight, &v2c.p, threshold,
                       nullptr, nullptr),
          "integrateTsdfVolume");
    volume_push(s, pv, pw, pg);
    cudaDeviceSynchronize();
}

// RayCaster.h:21-25 / RayCaster.cu:327 (no sync)
template <class IntrT, class Mat, class Vec, class Int3, class PV, class MapT>
void raycast(const IntrT &intr, const Mat &Rc2v, const Vec &tc2v, const Mat &Rv2w, const Vec &tv2w, float trunc_dist,
             const Int3 &volume_size, float voxel_size, const PV &value_volume, const PV &grad_volume, MapT &vmap, MapT &nmap) {
    const int res[3] = {volume_size.x, volume_size.y, volume_size.z};
    const Plane pv = as_plane(value_volume, 0), pg = as_plane(grad_volume, 0);
    VolumeShadow &s = shadow_of(pv.data, res, voxel_size, trunc_dist);
    {   // the planes may have been written by the caller since the last integration
        const size_t n = (size_t) res[0] * res[1] * res[2], row = (size_t) res[0] * sizeof(float), h = (size_t) res[1] * res[2];
        cudaMemcpy2D(s.dense, row, pv.data, pv.step, row, h, cudaMemcpyDeviceToDevice);
        cudaMemcpy2D(s.dense + 2 * n, row, pg.data, pg.step, row, h, cudaMemcpyDeviceToDevice);
        check(xs_volume_import_planes(s.v, 0, s.dense, nullptr, s.dense + 2 * n, nullptr), "volume import");
    }
    const int rows = vmap.rows() / 3, cols = vmap.cols();
    PoseHolder c2v(Rc2v, tc2v), v2w(Rv2w, tv2w);
    float *v = scratch(2).get((size_t) 2 * 3 * rows * cols), *n = scratch(3).get((size_t) 2 * 3 * rows * cols);
    check(xs_raycast(s.v, to_intr(intr), &c2v.p, &v2w.p, rows, cols, v, n, nullptr), "raycast");
    export_map(v, 1, 3, rows, cols, vmap);
    export_map(n, 1, 3, rows, cols, nmap);
}

// ExtractPointCloud.h:19-20 / ExtractPointCloud.cu:181-210 (sync).  output: PtrSz<float3>; returns min(size, points found).
// The order of the points is scheduling dependent on both sides (atomic appends); the point SET is the reference's.
template <class PV, class PW, class Int3, class Out>
size_t extractPoints(const PV &value_volume, const PW &weight_volume, const PV &grad_volume, const Int3 &volume_resolution,
                     float voxel_size, const Out &output) {
    const int res[3] = {volume_resolution.x, volume_resolution.y, volume_resolution.z};
    const Plane pv = as_plane(value_volume, 0), pw = as_plane(weight_volume, 0), pg = as_plane(grad_volume, 0);
    const Linear out = as_linear(output, 0);
    VolumeShadow &s = shadow_of(pv.data, res, voxel_size, 3.f * voxel_size);  // the truncation distance plays no role here
    volume_pull(s, pv, pw, pg);
    const long n = xs_extract_points(s.v, reinterpret_cast<float *>(out.data), nullptr, (long) out.size, nullptr);
    if (n < 0) check((int) n, "extractPoints");
    return (size_t) n;
}
// ExtractPointCloud.h:22-23 / ExtractPointCloud.cu:342-362 (sync): one normal per entry of `points` (points.size entries, as
// the reference: ExportPointCloud passes the whole buffer), divided by the squared norm (:305-306).
template <class PV, class PW, class Int3, class Pts>
void extractNormals(const PV &value_volume, const PW &weight_volume, const PV &grad_volume, const Int3 &volume_resolution,
                    float voxel_size, const Pts &points, const Pts &normal) {
    const int res[3] = {volume_resolution.x, volume_resolution.y, volume_resolution.z};
    const Plane pv = as_plane(value_volume, 0), pw = as_plane(weight_volume, 0), pg = as_plane(grad_volume, 0);
    const Linear pts = as_linear(points, 0), nrm = as_linear(normal, 0);
    VolumeShadow &s = shadow_of(pv.data, res, voxel_size, 3.f * voxel_size);
    volume_pull(s, pv, pw, pg);
    check(xs_extract_normals(s.v, reinterpret_cast<const float *>(pts.data), reinterpret_cast<float *>(nrm.data), (long) pts.size,
                             nullptr),
          "extractNormals");
}

// ICP.h:24-31 / ICP.cu:365 (sync + download).  gbuf / mbuf are the reference's scratch; unused here.
template <class Mat, class Vec, class MapT, class IntrT, class GBuf, class MBuf, class HostC>
void estimateCombined(const Mat &Rcurr, const Vec &tcurr, const MapT &vmap_curr, const MapT &nmap_curr, const Mat &Rprev_inv,
                      const Vec &tprev, const IntrT &intr, const MapT &vmap_g_prev, const MapT &nmap_g_prev, float distThres,
                      float angleThres, GBuf & /*gbuf*/, MBuf & /*mbuf*/, HostC *matrixA_host, HostC *vectorB_host) {
    const int rows = vmap_curr.rows() / 3, cols = vmap_curr.cols();
    // current-frame maps are real in the reference pipeline (Map.cu:196,229); their imaginary parts are dropped
    float *vc = import_map(vmap_curr, 3, 0, 4), *nc = import_map(nmap_curr, 3, 0, 5);
    float *vp = import_map(vmap_g_prev, 3, 1, 6), *np_ = import_map(nmap_g_prev, 3, 1, 7);
    PoseHolder curr(Rcurr, tcurr), prev(Rprev_inv, tprev);
    double A[2][36], b[2][6];
    check(xs_estimate_combined(&curr.p, vc, nc, &prev.p, to_intr(intr), vp, np_, rows, cols, 1, 1, distThres, angleThres,
                               &A[0][0], &b[0][0], nullptr),
          "estimateCombined");
    for (int i = 0; i < 36; ++i) matrixA_host[i] = HostC(A[0][i], A[1][i]);  // column-major 6x6, ICP.cu:419-428
    for (int i = 0; i < 6; ++i) vectorB_host[i] = HostC(b[0][i], b[1][i]);
}

// TsdfFusion.h:55-60 / TsdfFusion.cu:286 (sync).  MatD33 / devDComplex3 hold d_complex<float> = (value, eps1 | eps2, eps1eps2);
// the reference's per-voxel temporaries real_vec / grad_vec / hessian_vec / count_vec are not materialised (the sums are
// formed in one pass); depthScaled, threshold and k are unused by the reference kernel too.
template <class DepthT, class IntrT, class ScaledT, class Int3, class MatD, class VecD, class FVec, class IVec>
float4 ComputeLocalTsdf_hessian(const DepthT &depth, const IntrT &intr, ScaledT & /*depthScaled*/, const Int3 &volume_resolution,
                                float voxel_size, const MatD &Rv2c, const VecD &tv2c, float tranc_dist, float /*threshold*/,
                                float /*k*/, FVec &gt_vec, FVec & /*real_vec*/, FVec & /*grad_vec*/, FVec & /*hessian_vec*/,
                                IVec & /*count_vec*/) {
    const int res[3] = {volume_resolution.x, volume_resolution.y, volume_resolution.z};
    xs_pose p;
    float dR[27], dt[9];
    const auto *rows = Rv2c.data;
    for (int i = 0; i < 3; ++i) {
        const auto *e = &rows[i].x;  // x, y, z are consecutive members (Internal.h devDComplex3)
        for (int j = 0; j < 3; ++j) {
            p.R[i * 3 + j] = e[j].real().real();
            dR[i * 3 + j] = e[j].real().imag();
            dR[9 + i * 3 + j] = e[j].imag().real();
            dR[18 + i * 3 + j] = e[j].imag().imag();
        }
    }
    const auto *tv = &tv2c.x;
    for (int i = 0; i < 3; ++i) {
        p.t[i] = tv[i].real().real();
        dt[i] = tv[i].real().imag();
        dt[3 + i] = tv[i].imag().real();
        dt[6 + i] = tv[i].imag().imag();
    }
    p.ncomp = 3;
    p.dR = dR;
    p.dt = dt;
    double out[4];
    const Plane d = as_plane(depth, 0);
    check(xs_tsdf_hessian((const uint16_t *) d.data, d.step, d.rows, d.cols, to_intr(intr), res, voxel_size, &p, tranc_dist,
                          gt_vec.data().get(), out, nullptr),
          "ComputeLocalTsdf_hessian");
    return make_float4((float) out[0], (float) out[1], (float) out[2], (float) out[3]);
}

// TsdfFusion.h:48-52 / TsdfFusion.cu:409 (sync).  Mat33 = float3 data[3] (Internal.h:229-231).
template <class DepthT, class IntrT, class ScaledT, class Int3, class Mat, class FVec, class IVec>
float2 ComputeLocalTsdf_loss(const DepthT &depth, const IntrT &intr, ScaledT & /*depthScaled*/, const Int3 &volume_resolution,
                             float voxel_size, const Mat &Rv2c, const float3 &tv2c, float tranc_dist, float /*threshold*/,
                             float /*k*/, FVec &gt_vec, FVec & /*real_vec*/, IVec & /*count_vec*/) {
    const int res[3] = {volume_resolution.x, volume_resolution.y, volume_resolution.z};
    const float R[9] = {Rv2c.data[0].x, Rv2c.data[0].y, Rv2c.data[0].z, Rv2c.data[1].x, Rv2c.data[1].y,
                        Rv2c.data[1].z, Rv2c.data[2].x, Rv2c.data[2].y, Rv2c.data[2].z};
    const float t[3] = {tv2c.x, tv2c.y, tv2c.z};
    double out[2];
    const Plane d = as_plane(depth, 0);
    check(xs_tsdf_loss((const uint16_t *) d.data, d.step, d.rows, d.cols, to_intr(intr), res, voxel_size, R, t, tranc_dist,
                       gt_vec.data().get(), out, nullptr),
          "ComputeLocalTsdf_loss");
    return make_float2((float) out[0], (float) out[1]);
}

// ICP.h:34-40 / ICP.cu:431 (sync).  jacobi_buf / hessian_buf are the reference's scratch; unused here.  jacobi_host is an
// Eigen::Matrix4f of which rows 0..2 are written; hessian_host[i1][j1](i2, j2) as the reference.
template <class MapT, class Mat, class Vec, class IntrT, class JBuf, class M4>
void computeOptimizeMatrix(const MapT &vmap_curr, const MapT &nmap_curr, const MapT &vmap_g_prev, const MapT &nmap_g_prev,
                           const Mat &Rcurr, const Vec &tcurr, const Mat &Rprev_inv, const Vec &tprev, const IntrT &intr,
                           float distThres, float angleThres, JBuf & /*jacobi_buf*/, M4 &jacobi_host, JBuf * /*hessian_buf*/,
                           M4 **hessian_host) {
    const int rows = vmap_curr.rows() / 3, cols = vmap_curr.cols();
    float *vc = import_map(vmap_curr, 3, 0, 4), *nc = import_map(nmap_curr, 3, 0, 5);
    float *vp = import_map(vmap_g_prev, 3, 0, 6), *np_ = import_map(nmap_g_prev, 3, 0, 7);
    PoseHolder curr(Rcurr, tcurr), prev(Rprev_inv, tprev);
    double J[12], H[144];
    if (xs_compute_optimize_matrix(&curr.p, vc, nc, &prev.p, to_intr(intr), vp, np_, rows, cols, distThres, angleThres, J, H,
                                   nullptr) < 0)
        check(XS_ERR_CUDA, "computeOptimizeMatrix");
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) jacobi_host(i, j) = (float) J[i * 4 + j];
    for (int a = 0; a < 12; ++a)
        for (int b = 0; b < 12; ++b) hessian_host[a / 4][a % 4](b / 4, b % 4) = (float) H[a * 12 + b];
}

}  // namespace seam
}  // namespace xslam_b200
