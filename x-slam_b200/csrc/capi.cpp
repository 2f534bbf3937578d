// capi.cpp — process-wide state of the C-ABI, output writers and the synthetic depth source.
//
// Output formats follow the reference byte for byte: saveTxtMatrix (XKinectFusion/src/IOHelper.cpp:21-32)
// and CPointCloud::exportPly (Visualization/src/CPointCloud.cpp:42-67).  The synthetic depth source
// replaces the dataset readers (XKinectFusion/src/Dataset.cpp), which need files that do not exist
// offline; its scene, intrinsics and depth encoding are the ones fixed by SURVEY.md §8(d).
#include "../../include/xslam_b200.h"
#include "../../include/xslam_dcomplex.hpp"

#include <cuda_runtime_api.h>

#include <cmath>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <string>

namespace xs {
static thread_local std::string g_error;
long long g_launches = 0;
void set_error(const std::string &msg) { g_error = msg; }
// cudaDevAttrMultiProcessorCount of the current device, cached per device ordinal (148 on B200)
int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        cached[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
    }
    return cached[dev];
}
}  // namespace xs

namespace {

struct V3 {
    double x, y, z;
};
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline double len(V3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
inline double sd_box(V3 p, V3 c, V3 h) {
    V3 q = {std::fabs(p.x - c.x) - h.x, std::fabs(p.y - c.y) - h.y, std::fabs(p.z - c.z) - h.z};
    V3 m = {std::max(q.x, 0.0), std::max(q.y, 0.0), std::max(q.z, 0.0)};
    return len(m) + std::min(std::max(q.x, std::max(q.y, q.z)), 0.0);
}
inline double sd_sphere(V3 p, V3 c, double r) { return len(p - c) - r; }
// Box room 5.0 x 2.8 x 5.0 m around the first camera with two boxes and two spheres inside.
inline double scene_sdf(V3 p) {
    double d = -sd_box(p, {0.0, 0.0, 0.3}, {2.5, 1.4, 2.5});
    d = std::min(d, sd_sphere(p, {0.6, -0.9, 1.6}, 0.5));
    d = std::min(d, sd_box(p, {-0.9, -1.0, 1.9}, {0.4, 0.4, 0.4}));
    d = std::min(d, sd_box(p, {1.5, 0.0, 2.2}, {0.2, 1.4, 0.2}));
    d = std::min(d, sd_sphere(p, {-0.3, 0.3, 2.4}, 0.3));
    return d;
}

}  // namespace

extern "C" {

const char *xs_last_error(void) { return xs::g_error.c_str(); }
int xs_version(void) { return 100; }
long long xs_launch_count(void) { return xs::g_launches; }

// The host number type of include/xslam_dcomplex.hpp evaluated element-wise over host AoS arrays (n x 4 floats).  The
// reference's DoubleComplex is a HOST type (DeviceArray/src/DoubleComplex.cpp), so this is the host side of the API surface,
// not a fallback for any device path; tests hold it against the reference's own DoubleComplex.cpp.
int xs_dc_host_apply(int op, const float *a_aos, const float *b_aos, float p, float *out_aos, long n) {
    using xslam_b200::DoubleComplex;
    if (!a_aos || !out_aos || n < 0 || op < 0 || op > XS_DC_ATAN) return XS_ERR_ARG;
    const bool binary = op <= XS_DC_DIV || op == XS_DC_ATAN2;
    if (binary && !b_aos) return XS_ERR_ARG;
    for (long i = 0; i < n; ++i) {
        const DoubleComplex a(a_aos[4 * i], a_aos[4 * i + 1], a_aos[4 * i + 2], a_aos[4 * i + 3]);
        const DoubleComplex b = binary ? DoubleComplex(b_aos[4 * i], b_aos[4 * i + 1], b_aos[4 * i + 2], b_aos[4 * i + 3]) : DoubleComplex();
        DoubleComplex r;
        switch (op) {
        case XS_DC_ADD: r = a + b; break;
        case XS_DC_SUB: r = a - b; break;
        case XS_DC_MUL: r = a * b; break;
        case XS_DC_DIV: r = a / b; break;
        case XS_DC_SQRT: r = sqrt(a); break;
        case XS_DC_EXP: r = exp(a); break;
        case XS_DC_LOG: r = log(a); break;
        case XS_DC_SIN: r = sin(a); break;
        case XS_DC_COS: r = cos(a); break;
        case XS_DC_ATAN2: r = atan2(a, b); break;
        case XS_DC_POW: r = pow(a, p); break;
        default: r = atan(a); break;
        }
        out_aos[4 * i] = r.real().real(), out_aos[4 * i + 1] = r.real().imag();
        out_aos[4 * i + 2] = r.imag().real(), out_aos[4 * i + 3] = r.imag().imag();
    }
    return XS_OK;
}

int xs_save_pose_txt(const char *path, const float *m16) {
    if (!path || !m16) return XS_ERR_ARG;
    std::ofstream out(path);
    if (!out) return XS_ERR_ARG;
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 4; ++j) out << std::setprecision(7) << std::fixed << m16[i * 4 + j] << " ";
        out << "\n";
    }
    return XS_OK;
}

int xs_export_ply(const char *path, const float *points_xyz, const float *normals_xyz, long n) {
    if (!path || (n > 0 && (!points_xyz || !normals_xyz))) return XS_ERR_ARG;
    std::ofstream out(path);
    if (!out) return XS_ERR_ARG;
    out << "ply\nformat ascii 1.0\ncomment Created by myself\n";
    out << "element vertex " << n << "\n";
    out << "property float x\nproperty float y\nproperty float z\n";
    out << "property float nx\nproperty float ny\nproperty float nz\n";
    out << "end_header\n";
    for (long i = 0; i < n; ++i)
        out << points_xyz[3 * i] << " " << points_xyz[3 * i + 1] << " " << points_xyz[3 * i + 2] << " " << normals_xyz[3 * i]
            << " " << normals_xyz[3 * i + 1] << " " << normals_xyz[3 * i + 2] << "\n";
    return XS_OK;
}

int xs_synth_pose(int frame, float *c2w) {
    if (!c2w) return XS_ERR_ARG;
    const double phi = 2.0 * M_PI * frame / 300.0;
    const double yaw = 0.15 * std::sin(phi), pitch = 0.05 * std::sin(2 * phi);
    const double cy = std::cos(yaw), sy = std::sin(yaw), cp = std::cos(pitch), sp = std::sin(pitch);
    // R = Ry(yaw) * Rx(pitch)
    const double R[9] = {cy, sy * sp, sy * cp, 0, cp, -sp, -sy, cy * sp, cy * cp};
    const double t[3] = {0.4 * std::sin(phi), 0.1 * std::sin(2 * phi), 0.3 * (1 - std::cos(phi))};
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) c2w[i * 4 + j] = (float) R[i * 3 + j];
        c2w[i * 4 + 3] = (float) t[i];
    }
    c2w[12] = c2w[13] = c2w[14] = 0.f;
    c2w[15] = 1.f;
    return XS_OK;
}

int xs_synth_depth(const float *c2w, xs_intr intr, int rows, int cols, uint16_t *out) {
    if (!c2w || !out || rows <= 0 || cols <= 0) return XS_ERR_ARG;
#pragma omp parallel for schedule(dynamic, 4)
    for (int v = 0; v < rows; ++v)
        for (int u = 0; u < cols; ++u) {
            const double dx = (u - intr.cx) / intr.fx, dy = (v - intr.cy) / intr.fy, dz = 1.0;
            const double nrm = std::sqrt(dx * dx + dy * dy + dz * dz);
            V3 dir = {(c2w[0] * dx + c2w[1] * dy + c2w[2] * dz) / nrm, (c2w[4] * dx + c2w[5] * dy + c2w[6] * dz) / nrm,
                      (c2w[8] * dx + c2w[9] * dy + c2w[10] * dz) / nrm};
            V3 o = {c2w[3], c2w[7], c2w[11]};
            double t = 0.0;
            bool hit = false;
            for (int it = 0; it < 256; ++it) {
                V3 p = {o.x + dir.x * t, o.y + dir.y * t, o.z + dir.z * t};
                const double d = scene_sdf(p);
                if (d < 1e-5) {
                    hit = true;
                    break;
                }
                t += d;
                if (t > 20.0) break;
            }
            const double z_mm = hit ? std::round(t / nrm * 1000.0) : 0.0;
            out[(size_t) v * cols + u] = (z_mm < 200.0 || z_mm > 5000.0) ? 0 : (uint16_t) z_mm;
        }
    return XS_OK;
}

}  // extern "C"
