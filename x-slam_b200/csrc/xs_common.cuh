// xs_common.cuh — shared host/device declarations of libxslam_b200 (not part of the public ABI).
#pragma once
#include "../../include/xslam_b200.h"
#include "xs_batch.h"
#include "xs_jet.cuh"

#include <cstdio>
#include <string>

namespace xs {

// ---- error handling (the reference prints and exit(-1)s, cx.h:124-130; the C-ABI returns a status) ----
void set_error(const std::string &msg);
extern long long g_launches;  // kernels launched by this library (xs_launch_count)

#define XS_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            xs::set_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + \
                          std::to_string(__LINE__));                                               \
            return XS_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

#define XS_LAUNCH_CHECK()            \
    do {                             \
        ++xs::g_launches;            \
        XS_CUDA(cudaGetLastError()); \
    } while (0)

inline int div_up(int a, int b) { return (a + b - 1) / b; }  // cx::divUp, cx.h:131

// Streaming multiprocessors of the current device (148 on B200): persistent grids are sized in multiples of it.
int sm_count();

// ---- brick-tiled volume ------------------------------------------------------------------
// 8x8x8 bricks, x fastest inside a brick and across bricks:
//   value [brick][512] float, weight [brick][512] int32, deriv [brick][ncomp][512] float
// Replaces the reference's three pitched (Y*Z) x X planes (TsdfVolume.h:27-36).
constexpr int BRICK = 8;
constexpr int BRICK_VOX = 512;

struct VolumeView {
    float *value;
    int *weight;
    float *deriv;
    int rx, ry, rz;  // resolution in voxels
    int bx, by, bz;  // resolution in bricks
    int ncomp;
    float voxel;
    float trunc;
};

XS_DEV int brick_of(const VolumeView &V, int x, int y, int z) {
    return ((z >> 3) * V.by + (y >> 3)) * V.bx + (x >> 3);
}
XS_DEV int local_of(int x, int y, int z) { return ((z & 7) << 6) | ((y & 7) << 3) | (x & 7); }
XS_DEV size_t value_index(const VolumeView &V, int x, int y, int z) {
    return (size_t) brick_of(V, x, y, z) * BRICK_VOX + local_of(x, y, z);
}
XS_DEV size_t deriv_index(const VolumeView &V, int x, int y, int z, int comp) {
    return ((size_t) brick_of(V, x, y, z) * V.ncomp + comp) * BRICK_VOX + local_of(x, y, z);
}

}  // namespace xs

struct xs_volume {
    xs::VolumeView view;
    int comps, dirs;
    xs::Batch batch;     // meaning of the view.ncomp derivative planes
    float *d_dpose;      // staging for pose derivative components [3][ncomp][12]: slots 0, 1 = c2v, v2w (raycast), 2 = v2c (integration)
    bool pipelined;      // frame-loop mode: xs_integrate / xs_raycast do not synchronise (xs_volume_finish_frame collects stats)
    float *h_dpose;      // pinned host mirror
    float *d_depth_m;    // scaled depth (metres), TsdfFusion.cu:68-82
    int depth_capacity;  // pixels
    const uint16_t *prepared_depth = nullptr;  // depth frame whose pose-independent head (integrate_prepare) is already queued
    float *d_tile_max;   // largest depth per 16 x 16 pixel tile (depth-aware brick cull)
    int tile_capacity;   // tiles
    float *d_hit_time;   // raycast pass 1 -> pass 2: time of the sample before the crossing, < 0 = no hit
    int hit_capacity;    // pixels
    unsigned long long *d_stats;
    unsigned long long *h_stats;
    int2 *d_brick_list;           // [nbricks] bricks surviving the frustum cull of the current integration
    unsigned int *d_list_count;
    unsigned char *d_live;        // [nbricks] brick has (possibly) non-zero derivative planes
    size_t bytes;
    cudaEvent_t ev_k0, ev_k1;  // bracket the integration kernel (roofline timing)
    float last_kernel_ms;
    cudaEvent_t ev_h0 = nullptr, ev_h1 = nullptr;  // bracket the raycast hit kernel
    float last_hit_ms = 0.f;
    unsigned long long hit_stats[2] = {0, 0};  // pixels with a valid vertex / normal of the last collected raycast
};
