// oracle/seam_run.cu — TEST INFRASTRUCTURE ONLY: run-time proof of the drop-in boundary ("Binding A").
//
// One program links BOTH sides of the seam:
//   * the reference's own operators (global namespace; unmodified Map.cu / TsdfFusion.cu / RayCaster.cu / ICP.cu /
//     ExtractPointCloud.cu compiled into oracle/_ref/libxslam_ref.so), and
//   * the wrappers with the reference's signatures in include/xslam_b200.hpp (namespace xslam_b200::seam), which forward
//     to the C-ABI of libxslam_b200.so,
// and drives them with the REFERENCE's container / POD types (DeviceArray2D, DeviceArray, MatS33, devComplex3, Intr) through
// the call sequence of KinectFusionReconstruction.cpp: initVolume (TsdfVolume.cpp:50) -> SurfaceMeasure (:280-299) ->
// integrateTsdfVolume (:264) -> raycast (:327) -> resizeVMap / resizeNMap (:274-275) -> [next frame] SurfaceMeasure ->
// estimateCombined (:198) -> extractPoints / extractNormals (:343-346), with a complex pose perturbation (imaginary seed on
// the volume-to-camera pose).  Every output of the wrapper side is compared with the reference side; the program prints
// one JSON object and exits 0 only if all comparisons are within the tolerances of tests/test_gpu_stages.py.
// Built by `make -C oracle ref` into oracle/_ref/seam_run (needs /root/reference; the binary travels to the GPU box);
// run by tests/test_seam_run.py.
#include "CudaFunctions.h"  // reference: Map.h, TsdfFusion.h, RayCaster.h, ICP.h, ExtractPointCloud.h
#include "TsdfVolume.h"

#include "../include/xslam_b200.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace S = xslam_b200::seam;

namespace {

struct Cmp {
    long n = 0, nan_mismatch = 0, real_ulp_gt0 = 0;
    double max_real_abs = 0, max_imag_abs = 0, imag_scale = 0;
};
long ulp(float a, float b) {
    int ia, ib;
    std::memcpy(&ia, &a, 4);
    std::memcpy(&ib, &b, 4);
    if (ia < 0) ia = -(ia & 0x7fffffff);
    if (ib < 0) ib = -(ib & 0x7fffffff);
    return std::labs((long) ia - (long) ib);
}
// complex maps: rows x cols interleaved (re, im); NaN patterns must agree on the real part
Cmp compare(const MapArr &mine, const MapArr &ref, int nplanes) {
    Cmp c;
    const int rows = ref.rows(), cols = ref.cols();
    if (mine.rows() != rows || mine.cols() != cols) {
        c.nan_mismatch = -1;
        return c;
    }
    std::vector<float> a((size_t) rows * cols * 2), b(a.size());
    mine.download(a.data(), cols * sizeof(devComplex));
    ref.download(b.data(), cols * sizeof(devComplex));
    // x-plane NaN marks an invalid pixel; the y / z planes of invalid pixels are unspecified in the reference (Map.cu:27)
    const int prow = rows / nplanes;
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            const size_t i = ((size_t) y * cols + x) * 2, ix = ((size_t) (y % prow) * cols + x) * 2;
            const bool inv_a = std::isnan(a[ix]), inv_b = std::isnan(b[ix]);
            if (inv_a != inv_b) {
                if (y < prow) ++c.nan_mismatch;
                continue;
            }
            if (inv_b) continue;
            ++c.n;
            if (ulp(a[i], b[i]) > 0) ++c.real_ulp_gt0;
            c.max_real_abs = std::max(c.max_real_abs, (double) std::fabs(a[i] - b[i]));
            c.max_imag_abs = std::max(c.max_imag_abs, (double) std::fabs(a[i + 1] - b[i + 1]));
            c.imag_scale = std::max(c.imag_scale, (double) std::fabs(b[i + 1]));
        }
    return c;
}
void print_cmp(const char *name, const Cmp &c, bool last = false) {
    std::printf("  \"%s\": {\"n\": %ld, \"nan_mismatch\": %ld, \"real_ulp_gt0\": %ld, \"max_real_abs\": %.3g, \"imag_rel\": %.3g}%s\n", name, c.n,
                c.nan_mismatch, c.real_ulp_gt0, c.max_real_abs, c.imag_scale > 0 ? c.max_imag_abs / c.imag_scale : 0.0, last ? "" : ",");
}

struct Side {  // the buffers one implementation of the frame loop owns (KinectFusionReconstruction.h:84-111)
    MapArr depth[3], vmap_curr[3], nmap_curr[3], vmap_prev[3], nmap_prev[3];
    DeviceArray2D<int> volume;
    DeviceArray2D<float> value, grad, depth_scaled;
    DeviceArray2D<int> weight;
    DeviceArray<float3> cloud, normals;
    size_t npoints = 0;
};

MatS33 mat(const float R[9], const float dR[9]) {
    MatS33 M;
    for (int r = 0; r < 3; ++r) {
        M.data[r].x = devComplex(R[r * 3 + 0], dR[r * 3 + 0]);
        M.data[r].y = devComplex(R[r * 3 + 1], dR[r * 3 + 1]);
        M.data[r].z = devComplex(R[r * 3 + 2], dR[r * 3 + 2]);
    }
    return M;
}
devComplex3 vec(const float t[3], const float dt[3]) {
    devComplex3 v;
    v.x = devComplex(t[0], dt[0]);
    v.y = devComplex(t[1], dt[1]);
    v.z = devComplex(t[2], dt[2]);
    return v;
}

}  // namespace

// localises a CUDA failure to the stage that produced it (the reference's own cudaSafeCall only names the next checker)
static void step(const char *name) {
    const cudaError_t e1 = cudaDeviceSynchronize(), e2 = cudaGetLastError();
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        std::fprintf(stderr, "seam_run: CUDA error after %s: %s / %s\n", name, cudaGetErrorString(e1), cudaGetErrorString(e2));
        std::exit(2);
    }
}

int main() {
    const int W = 320, H = 240, RES = 128;
    const float voxel = 0.06f, thres_range = 3.f, trunc = std::max(voxel * thres_range, 2.1f * voxel);
    const Intr intr(481.20f / 2, -480.00f / 2, 319.50f / 2, 239.50f / 2);
    const xs_intr xi = {intr.fx, intr.fy, intr.cx, intr.cy};
    const int3 res = make_int3(RES, RES, RES);
    // two frames of the synthetic stream (host code of the product library)
    std::vector<ushort> d0((size_t) W * H), d1((size_t) W * H);
    float c2w0[16], c2w1[16];
    xs_synth_pose(0, c2w0);
    xs_synth_pose(4, c2w1);
    xs_synth_depth(c2w0, xi, H, W, d0.data());
    xs_synth_depth(c2w1, xi, H, W, d1.data());
    DeviceArray2D<ushort> depth0, depth1;
    depth0.upload(d0.data(), W * sizeof(ushort), H, W);
    depth1.upload(d1.data(), W * sizeof(ushort), H, W);
    // frame 0 is the identity camera: camera-to-volume = translation by init_xyz, with an imaginary seed on every entry
    const float I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, init[3] = {3.2f, 3.2f, 3.2f}, ninit[3] = {-3.2f, -3.2f, -3.2f}, zero3[3] = {0, 0, 0}, zero9[9] = {0};
    float dR[9], dt[3];
    for (int i = 0; i < 9; ++i) dR[i] = 1e-7f * (0.3f + 0.1f * i) * (i % 2 ? -1.f : 1.f);
    for (int i = 0; i < 3; ++i) dt[i] = 1e-7f * (0.5f - 0.4f * i);
    const MatS33 Rv2c = mat(I3, dR), Rc2v = mat(I3, dR), Rv2w = mat(I3, zero9);
    const devComplex3 tv2c = vec(ninit, dt), tc2v = vec(init, dt), tv2w = vec(ninit, zero3);
    Side ref, mine;
    for (Side *s : {&ref, &mine}) {
        s->volume.create(RES * RES, RES);
        s->value.create(RES * RES, RES);
        s->weight.create(RES * RES, RES);
        s->grad.create(RES * RES, RES);
        s->cloud.create(1000000);
        s->normals.create(1000000);
        for (int l = 0; l < 3; ++l) {  // AllocateBuffers, KinectFusionReconstruction.cpp:84-92 (bilateralFilter does not create its output)
            s->depth[l].create(H >> l, W >> l);
            s->vmap_prev[l].create(3 * (H >> l), W >> l);
            s->nmap_prev[l].create(3 * (H >> l), W >> l);
        }
    }
    // ---- TsdfVolume::reset (TsdfVolume.cpp:44-51): poison first so that the zeroing is observable
    for (Side *s : {&ref, &mine}) {
        cudaMemset2D(s->value.ptr(), s->value.step(), 0x3f, RES * sizeof(float), RES * RES);
        cudaMemset2D(s->weight.ptr(), s->weight.step(), 0x01, RES * sizeof(int), RES * RES);
        cudaMemset2D(s->grad.ptr(), s->grad.step(), 0x3f, RES * sizeof(float), RES * RES);
    }
    step("setup");
    ::initVolume(ref.volume, ref.value, ref.weight, ref.grad, res);
    step("reference initVolume");
    S::initVolume(mine.volume, mine.value, mine.weight, mine.grad, res);
    step("wrapper initVolume");
    // ---- SurfaceMeasure, frame 0 (KinectFusionReconstruction.cpp:280-299)
    auto surface_ref = [&](Side &s, const DeviceArray2D<ushort> &d) {
        ::bilateralFilter(d, s.depth[0]);
        for (int i = 1; i < 3; ++i) ::pyrDown(s.depth[i - 1], s.depth[i]);
        for (int i = 0; i < 3; ++i) {
            ::createVMap(intr(i), s.depth[i], s.vmap_curr[i]);
            ::createNMap(s.vmap_curr[i], s.nmap_curr[i]);
        }
    };
    auto surface_mine = [&](Side &s, const DeviceArray2D<ushort> &d) {
        S::bilateralFilter(d, s.depth[0]);
        for (int i = 1; i < 3; ++i) S::pyrDown(s.depth[i - 1], s.depth[i]);
        for (int i = 0; i < 3; ++i) {
            S::createVMap(intr(i), s.depth[i], s.vmap_curr[i]);
            S::createNMap(s.vmap_curr[i], s.nmap_curr[i]);
        }
    };
    surface_ref(ref, depth0);
    step("reference SurfaceMeasure");
    surface_mine(mine, depth0);
    step("wrapper SurfaceMeasure");
    std::printf("{\n");
    int bad = 0;
    auto gate = [&](const char *name, const Cmp &c, double real_abs, double imag_rel) {
        print_cmp(name, c);
        const double ir = c.imag_scale > 0 ? c.max_imag_abs / c.imag_scale : 0.0;
        if (c.n <= 0 || c.nan_mismatch != 0 || c.max_real_abs > real_abs || ir > imag_rel) {
            std::fprintf(stderr, "seam_run: %s out of tolerance\n", name);
            ++bad;
        }
    };
    gate("depth_l2", compare(mine.depth[2], ref.depth[2], 1), 0.0, 0.0);
    gate("vmap_curr_l0", compare(mine.vmap_curr[0], ref.vmap_curr[0], 3), 0.0, 0.0);
    gate("nmap_curr_l1", compare(mine.nmap_curr[1], ref.nmap_curr[1], 3), 0.0, 0.0);
    // ---- IntegrateFrame (KinectFusionReconstruction.cpp:237-278): integrate, raycast, pyramid
    ::integrateTsdfVolume(depth0, intr, 100, res, voxel, Rv2c, tv2c, tc2v, trunc, ref.value, ref.weight, ref.grad, ref.depth_scaled, 0, 0.f, 0.f);
    step("reference integrateTsdfVolume");
    S::integrateTsdfVolume(depth0, intr, 100, res, voxel, Rv2c, tv2c, tc2v, trunc, mine.value, mine.weight, mine.grad, mine.depth_scaled, 0, 0.f, 0.f);
    step("wrapper integrateTsdfVolume");
    {
        std::vector<float> va((size_t) RES * RES * RES), vb(va.size()), ga(va.size()), gb(va.size());
        std::vector<int> wa(va.size()), wb(va.size());
        mine.value.download(va.data(), RES * sizeof(float));
        ref.value.download(vb.data(), RES * sizeof(float));
        mine.grad.download(ga.data(), RES * sizeof(float));
        ref.grad.download(gb.data(), RES * sizeof(float));
        mine.weight.download(wa.data(), RES * sizeof(int));
        ref.weight.download(wb.data(), RES * sizeof(int));
        long wm = 0, vm = 0, upd = 0;
        double gmax = 0, gsc = 0;
        for (size_t i = 0; i < va.size(); ++i) {
            wm += wa[i] != wb[i];
            vm += ulp(va[i], vb[i]) > 0;
            upd += wb[i] > 0;
            gmax = std::max(gmax, (double) std::fabs(ga[i] - gb[i]));
            gsc = std::max(gsc, (double) std::fabs(gb[i]));
        }
        std::printf("  \"volume\": {\"updated\": %ld, \"weight_mismatch\": %ld, \"value_ulp_gt0\": %ld, \"grad_max_rel\": %.3g},\n", upd, wm, vm, gsc > 0 ? gmax / gsc : 0.0);
        if (upd < 10000 || wm != 0 || vm != 0 || gmax > 2e-3 * gsc) {
            std::fprintf(stderr, "seam_run: volume out of tolerance\n");
            ++bad;
        }
    }
    ::raycast(intr, Rc2v, tc2v, Rv2w, tv2w, trunc, res, voxel, ref.value, ref.grad, ref.vmap_prev[0], ref.nmap_prev[0]);
    step("reference raycast");
    S::raycast(intr, Rc2v, tc2v, Rv2w, tv2w, trunc, res, voxel, mine.value, mine.grad, mine.vmap_prev[0], mine.nmap_prev[0]);
    step("wrapper raycast");
    for (int i = 1; i < 3; ++i) {
        ::resizeVMap(ref.vmap_prev[i - 1], ref.vmap_prev[i]);
        ::resizeNMap(ref.nmap_prev[i - 1], ref.nmap_prev[i]);
        S::resizeVMap(mine.vmap_prev[i - 1], mine.vmap_prev[i]);
        S::resizeNMap(mine.nmap_prev[i - 1], mine.nmap_prev[i]);
    }
    step("pyramid");
    // raycast maps.  The stage tests show 0 ulp against a ZERO-seed reference pass; here the reference runs with the seeded
    // complex pose, and its own real part then moves by a few ulp (ac - bd, c^2 + d^2 and the polar sqrt
    // couple h^2-sized terms into the real part), amplified in the normals, which are differences of nearly equal samples:
    // gates 2e-6 m on vertices, 2e-5 on unit normals, with the ulp counts reported
    gate("vmap_g_prev_l0", compare(mine.vmap_prev[0], ref.vmap_prev[0], 3), 2e-6, 5e-3);
    gate("nmap_g_prev_l0", compare(mine.nmap_prev[0], ref.nmap_prev[0], 3), 2e-5, 5e-3);
    gate("vmap_g_prev_l2", compare(mine.vmap_prev[2], ref.vmap_prev[2], 3), 2e-6, 5e-3);
    gate("nmap_g_prev_l2", compare(mine.nmap_prev[2], ref.nmap_prev[2], 3), 2e-5, 5e-3);
    // ---- next frame: SurfaceMeasure + one estimateCombined per level (KinectFusionReconstruction.cpp:186-202)
    surface_ref(ref, depth1);
    surface_mine(mine, depth1);
    step("SurfaceMeasure, next frame");
    const float angle_thres = std::sin(15.f / 180.f * 3.14159265f);
    const MatS33 Rcurr = mat(I3, dR), Rprev_inv = mat(I3, zero9);
    const devComplex3 tcurr = vec(zero3, dt), tprev = vec(zero3, zero3);
    for (int level = 2; level >= 0; --level) {
        DeviceArray2D<devComplexICP> gbuf_r, gbuf_m;
        DeviceArray<devComplexICP> mbuf_r, mbuf_m;
        hostComplexICP A_r[36], b_r[6], A_m[36], b_m[6];
        ::estimateCombined(Rcurr, tcurr, ref.vmap_curr[level], ref.nmap_curr[level], Rprev_inv, tprev, intr(level), ref.vmap_prev[level],
                           ref.nmap_prev[level], 0.10f, angle_thres, gbuf_r, mbuf_r, A_r, b_r);
        S::estimateCombined(Rcurr, tcurr, mine.vmap_curr[level], mine.nmap_curr[level], Rprev_inv, tprev, intr(level), mine.vmap_prev[level],
                            mine.nmap_prev[level], 0.10f, angle_thres, gbuf_m, mbuf_m, A_m, b_m);
        step("estimateCombined");
        double re = 0, im = 0, rs = 0, is = 0;
        for (int i = 0; i < 36; ++i) {
            re = std::max(re, std::fabs(A_m[i].real() - A_r[i].real()));
            im = std::max(im, std::fabs(A_m[i].imag() - A_r[i].imag()));
            rs = std::max(rs, std::fabs(A_r[i].real()));
            is = std::max(is, std::fabs(A_r[i].imag()));
        }
        for (int i = 0; i < 6; ++i) {
            re = std::max(re, std::fabs(b_m[i].real() - b_r[i].real()));
            im = std::max(im, std::fabs(b_m[i].imag() - b_r[i].imag()));
        }
        std::printf("  \"icp_l%d\": {\"A00\": %.6g, \"real_rel\": %.3g, \"imag_rel\": %.3g},\n", level, A_r[0].real(), re / rs, is > 0 ? im / is : 0.0);
        if (!(rs > 0) || re > 1e-6 * rs || im > 2e-4 * is) {
            std::fprintf(stderr, "seam_run: icp level %d out of tolerance\n", level);
            ++bad;
        }
    }
    // ---- ExportPointCloud (KinectFusionReconstruction.cpp:334-346)
    ref.npoints = ::extractPoints(ref.value, ref.weight, ref.grad, res, voxel, ref.cloud);
    ::extractNormals(ref.value, ref.weight, ref.grad, res, voxel, ref.cloud, ref.normals);
    mine.npoints = S::extractPoints(mine.value, mine.weight, mine.grad, res, voxel, mine.cloud);
    S::extractNormals(mine.value, mine.weight, mine.grad, res, voxel, mine.cloud, mine.normals);
    step("ExportPointCloud");
    {
        struct PN {
            float p[3], n[3];
            bool operator<(const PN &o) const { return std::lexicographical_compare(p, p + 3, o.p, o.p + 3); }
        };
        auto fetch = [](const Side &s) {
            std::vector<float3> p(s.npoints), n(s.npoints);
            cudaMemcpy(p.data(), s.cloud.ptr(), s.npoints * sizeof(float3), cudaMemcpyDeviceToHost);
            cudaMemcpy(n.data(), s.normals.ptr(), s.npoints * sizeof(float3), cudaMemcpyDeviceToHost);
            std::vector<PN> v(s.npoints);
            for (size_t i = 0; i < s.npoints; ++i) v[i] = PN{{p[i].x, p[i].y, p[i].z}, {n[i].x, n[i].y, n[i].z}};
            std::sort(v.begin(), v.end());
            return v;
        };
        const std::vector<PN> a = fetch(mine), b = fetch(ref);
        long pdiff = a.size() != b.size(), nulp = 0;
        for (size_t i = 0; i < std::min(a.size(), b.size()); ++i)
            for (int c = 0; c < 3; ++c) {
                pdiff += a[i].p[c] != b[i].p[c];
                if (std::isfinite(a[i].n[c]) && std::isfinite(b[i].n[c])) nulp = std::max(nulp, ulp(a[i].n[c], b[i].n[c]));
            }
        std::printf("  \"extract\": {\"points_ref\": %zu, \"points\": %zu, \"point_mismatch\": %ld, \"normal_max_ulp\": %ld}\n", b.size(), a.size(), pdiff, nulp);
        if (b.size() < 200 || pdiff != 0 || nulp > 4) {
            std::fprintf(stderr, "seam_run: point cloud out of tolerance\n");
            ++bad;
        }
    }
    std::printf("}\n");
    std::fprintf(stderr, bad ? "seam_run: %d comparison(s) FAILED\n" : "seam_run: all comparisons within tolerance\n", bad);
    return bad ? 1 : 0;
}
