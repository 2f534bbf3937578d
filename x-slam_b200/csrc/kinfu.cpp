// kinfu.cpp — the frame-loop orchestrator behind the C-ABI (host C++).
//
// Mirrors class KinectFusionReconstruction (XKinectFusion/include/KinectFusionReconstruction.h:19-220,
// XKinectFusion/src/KinectFusionReconstruction.cpp:9-332): SetYamlParameters / AllocateBuffers,
// ProcessFrame = AlignDepthToReconstruction (SurfaceMeasure + PoseEstimate) + IntegrateFrame
// (integrateTsdfVolume + raycast + pyramid resize), with every pose-dependent quantity carried as a batch of
// k perturbation directions (HJet) instead of one std::complex<float> imaginary part.
//
// Differences that are deliberate: all device buffers are allocated once (the reference
// mallocs/frees depthScaled every frame, TsdfFusion.cu:180,198, and keeps two dead N^3 arrays,
// KinectFusionReconstruction.cpp:81, TsdfVolume.cpp:17); one stream; no cudaDeviceSynchronize between
// stages except where the host needs a result (the 6x6 solve).
#include "../../include/xslam_b200.h"
#include "host_jet.h"
#include "xs_batch.h"

#include <cuda_runtime_api.h>

#include <chrono>
#include <complex>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace xs {
void set_error(const std::string &msg);
extern long long g_launches;
// icp.cu: one Gauss-Newton iteration queued on the stream, no host round trip
struct IcpScratch;
IcpScratch *icp_scratch_create();
void icp_scratch_destroy(IcpScratch *sc);
int icp_iteration_async(IcpScratch *sc, const float *d_pose_curr, const float *d_vmap_curr, const float *d_nmap_curr, const xs_pose *prev,
                        xs_intr intr, const float *d_vmap_g_prev, const float *d_nmap_g_prev, int rows, int cols, const Batch &batch,
                        float dist_thres, float angle_thres, float *d_pose_out, int solve_mode, int *d_status,
                        double *d_log, cudaStream_t s, cudaStream_t s_real, int slot);
// surface.cu: derivative components of the current-frame maps (parameters that move the intrinsics)
int surface_derivs(const float *d_depth, int rows, int cols, int level, xs_intr intr_level0, const BatchView &B, float *d_vmap, float *d_nmap,
                   cudaStream_t s);
// raycast.cu: resizeVMap / resizeNMap for a batch description
int resize_map_batch(bool normalize, const float *d_in, int rows, int cols, const BatchView &B, float *d_out, cudaStream_t s);
void icp_timing_reset(IcpScratch *sc);
// integrate.cu: pose-independent head of the integration (metric depth, tile maxima, cleared counters)
int integrate_prepare(xs_volume *v, const uint16_t *d_depth, size_t depth_step_bytes, int rows, int cols, cudaStream_t s);
}  // namespace xs
using namespace xs;

#define KCUDA(expr)                                                                          \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            set_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at kinfu.cpp:" + std::to_string(__LINE__)); \
            return XS_ERR_CUDA;                                                              \
        }                                                                                    \
    } while (0)

// xs_kinfu_pose_estimate returns 1 / 0 like the reference's AlignDepthToReconstruction: a CUDA failure is "not aligned" (0)
// with the error text set, never a negative status that a caller's `if (!aligned)` would read as success.
#define KCUDA0(expr)                                                                         \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            set_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at kinfu.cpp:" + std::to_string(__LINE__)); \
            return 0;                                                                        \
        }                                                                                    \
    } while (0)

struct xs_kinfu {
    xs_config cfg;
    int comps, dirs, ncomp, solve_mode;
    Batch batch;                 // what the ncomp derivative components mean (xs_batch.h)
    std::vector<HPair> hpairs;   // host copy of the pair table for the host jets (comps == 2)
    xs_intr intr;
    HMat4 world2camera, world2volume;
    std::vector<HMat4> record;  // world2camera_record
    int frame_id = 0;
    int frame_step = 1;  // KinectFusionReconstruction.cpp:72: frame_id += frame_step after every processed frame
    int icp_iterations[3] = {5, 4, 3};  // KinectFusionReconstruction.cpp:54
    float angle_thres;
    xs_volume *volume = nullptr;
    IcpScratch *icp = nullptr;  // ICP scratch of this pipeline (records, sums, tickets, events): nothing process-wide
    cudaStream_t stream = nullptr;
    uint16_t *d_depth2[2] = {nullptr, nullptr};  // uploaded host frames, alternating (the previous frame may still be integrating)
    int depth_idx = 0;
    uint16_t *h_depth = nullptr;  // pinned staging
    std::vector<float *> depths, vmaps_curr, nmaps_curr, vmaps_prev, nmaps_prev;
    float *d_record = nullptr, *h_record = nullptr;  // [(1+ncomp)][16]
    static constexpr int POSE_SLOTS = 13;            // 12 Gauss-Newton iterations + the initial estimate
    float *d_pose_all = nullptr;                     // ICP pose after each iteration: [POSE_SLOTS][(1+ncomp)][12] (R row-major, t)
    float *d_pose_slot(int it) const { return d_pose_all + (size_t) (it % POSE_SLOTS) * (1 + ncomp) * 12; }
    cudaStream_t stream_real = nullptr;              // the real chain of the ICP runs ahead of the derivative kernels (icp.cu)
    cudaEvent_t ev_icp_start = nullptr;              // inputs of the ICP (maps, initial pose, cleared status) are ready
    float *h_pose = nullptr;                         // pinned staging of the same
    std::vector<float> gt_poses;  // [n][16] camera-to-world, KinectFusionReconstruction.h:36
    bool use_gt_pose = false;     // flag_use_gtPose, KinectFusionReconstruction.cpp:69
    int *d_status = nullptr, *h_status = nullptr;    // [2] ICP degeneracy flag (icp.cu SolveParams::status)
    bool log_icp = false;
    double *d_icp_log = nullptr, *h_icp_log = nullptr;  // [max 16 iterations][27*(1+ncomp)] sums per iteration
    std::vector<double> icp_log;
    cudaEvent_t ev[2][6];  // stage brackets ([5] = the pipeline's stream takes over after the head), two sets: a deferred frame is collected after the next frame's head is queued
    int ev_cur = 0;        // set of the frame being queued / last queued
    cudaEvent_t ev_head = nullptr;  // the early head (upload + surface measurement on stream_real) is complete
    float ms[5] = {0, 0, 0, 0, 0};
    long long launches[5] = {0, 0, 0, 0, 0};
    unsigned long long stats[4] = {0, 0, 0, 0};
    int icp_iters_done = 0;
    std::vector<float> dR, dt, dR2, dt2;  // scratch for xs_pose
    const uint16_t *next_depth = nullptr;  // ProcessFrame: frame whose integration head is queued behind the ICP download
    cudaEvent_t ev_icp = nullptr;          // the ICP result has reached the host buffers
    bool h_depth_free = true;  // no upload from the pinned staging frame is in flight
    // multi-GPU (comm.cpp): the record of every rank is all-gathered after each frame on a stream of its own
    xs_comm *comm = nullptr;
    int record_floats = 0;                 // floats per rank in the gather (>= (1 + ncomp) * 16, the same on every rank)
    float *d_gathered[2] = {nullptr, nullptr};  // [world][record_floats], alternating per frame: the host may read frame f - 1's
                                                // records while frame f's gather is in flight (xs_kinfu_get_gathered_records_lagged)
    int gather_slot = 0;                        // buffer of the last queued gather
    cudaStream_t stream_comm = nullptr;
    cudaEvent_t ev_record = nullptr, ev_gather[2] = {nullptr, nullptr};  // record uploaded / gather has read it and written its buffer
    bool gather_in_flight = false;
    bool keep_curr_derivs = false;  // xs_kinfu_keep_current_map_derivatives
    bool deferred = false;  // xs_kinfu_set_deferred: ProcessFrame returns once integration + raycast are queued
    bool pending = false;   // a frame's integration / raycast may still be running; its statistics are not collected yet
};

namespace {

xs_intr level_intr(const xs_intr &k, int level) {  // Intr::operator(), Internal.h:55-58
    const int div = 1 << level;
    return xs_intr{k.fx / div, k.fy / div, k.cx / div, k.cy / div};
}

// HMat3/HVec3 -> xs_pose (real + derivative components)
void to_pose(const HMat3 &R, const HVec3 &t, int ncomp, std::vector<float> &dR, std::vector<float> &dt, xs_pose &p) {
    dR.resize((size_t) (ncomp > 0 ? ncomp : 1) * 9);
    dt.resize((size_t) (ncomp > 0 ? ncomp : 1) * 3);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            p.R[i * 3 + j] = R.m[i][j].v;
            for (int q = 0; q < ncomp; ++q) dR[q * 9 + i * 3 + j] = R.m[i][j].d[q];
        }
        p.t[i] = t.v[i].v;
        for (int q = 0; q < ncomp; ++q) dt[q * 3 + i] = t.v[i].d[q];
    }
    p.ncomp = ncomp;
    p.dR = dR.data();
    p.dt = dt.data();
}

size_t map_floats(const xs_kinfu *k, int level, bool jets) {
    const size_t r = k->cfg.height >> level, c = k->cfg.width >> level;
    return r * c * 3 * (jets ? (1 + k->ncomp) : 1);
}

void set_ctx(const xs_kinfu *k) {
    hj_ctx().comps = k->comps;
    hj_ctx().dirs = k->dirs;
    hj_ctx().npairs = (int) k->hpairs.size();
    hj_ctx().pairs = k->hpairs.empty() ? nullptr : k->hpairs.data();
}

// After the frame's work has completed on the stream: integration statistics and per-stage device times.
void collect_frame(xs_kinfu *k) {
    xs_volume_finish_frame(k->volume, k->stats);
    cudaEvent_t *ev = k->ev[k->ev_cur];
    // surface: head (in deferred mode it may have run beside the previous frame's raycast, on the second stream); icp,
    // integrate, raycast: on the pipeline's stream from the point where it takes over; total = their sum
    cudaEventElapsedTime(&k->ms[0], ev[0], ev[1]);
    cudaEventElapsedTime(&k->ms[1], ev[5], ev[2]);
    cudaEventElapsedTime(&k->ms[2], ev[2], ev[3]);
    cudaEventElapsedTime(&k->ms[3], ev[3], ev[4]);
    k->ms[4] = k->ms[0] + k->ms[1] + k->ms[2] + k->ms[3];
    k->launches[4] = k->launches[0] + k->launches[1] + k->launches[2] + k->launches[3];
}

// Deferred mode: waits for the frame whose integration / raycast were left running and collects its statistics.
int finish_pending(xs_kinfu *k) {
    if (!k->pending) return XS_OK;
    k->pending = false;
    if (cudaStreamSynchronize(k->stream) != cudaSuccess) {
        set_error("xs_kinfu: the deferred frame failed on the device");
        return XS_ERR_CUDA;
    }
    k->h_depth_free = true;
    collect_frame(k);
    return XS_OK;
}

}  // namespace

extern "C" {

static xs_kinfu *kinfu_create(const xs_config *cfg, int comps, int dirs, int npairs, const int *pairs, const float *seeds, int solve_mode);

xs_kinfu *xs_kinfu_create(const xs_config *cfg, int comps, int dirs, const float *seeds, int solve_mode) {
    return kinfu_create(cfg, comps, dirs, -1, nullptr, seeds, solve_mode);
}
// Hessian batch (comps = 2): nparams first-order components and one second-order component per listed pair (i <= j, sorted by
// i; pairs == NULL: all nparams (nparams + 1) / 2 pairs).  seeds: [(nparams + npairs)][16].
xs_kinfu *xs_kinfu_create_hessian(const xs_config *cfg, int nparams, int npairs, const int *pairs, const float *seeds, int solve_mode) {
    return kinfu_create(cfg, 2, nparams, pairs ? npairs : -1, pairs, seeds, solve_mode);
}

static xs_kinfu *kinfu_create(const xs_config *cfg, int comps, int dirs, int npairs, const int *pairs, const float *seeds, int solve_mode) {
    if (!cfg || (comps != 1 && comps != 2 && comps != 3) || dirs < 0 || batch_ncomp(comps, dirs, npairs) > HJ_MAX || cfg->num_levels < 1 ||
        cfg->num_levels > 3) {
        set_error("xs_kinfu_create: bad arguments (comps in {1,2,3}, at most 256 derivative components, 1 <= num_levels <= 3)");
        return nullptr;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("xs_kinfu_create: no CUDA device (libxslam_b200 has no CPU fallback)");
        return nullptr;
    }
    // the host jets size their component arrays from the thread's context: set it before any HJet is constructed
    hj_ctx().comps = comps;
    hj_ctx().dirs = dirs;
    hj_ctx().npairs = comps == 2 ? batch_ncomp(2, dirs, npairs) - dirs : 0;
    hj_ctx().pairs = nullptr;
    xs_kinfu *k = new xs_kinfu();
    k->cfg = *cfg;
    k->comps = comps;
    k->dirs = dirs;
    if (batch_init(k->batch, comps, dirs, npairs, pairs) != XS_OK) {
        delete k;
        return nullptr;
    }
    for (int i = 0; i < k->batch.v.m; ++i) k->hpairs.push_back(HPair{k->batch.h_pairs[i].x, k->batch.h_pairs[i].y});
    k->ncomp = k->batch.v.ncomp;
    set_ctx(k);
    k->solve_mode = solve_mode;
    k->frame_step = cfg->frame_step > 0 ? cfg->frame_step : 1;
    k->intr = xs_intr{cfg->fx, cfg->fy, cfg->cx, cfg->cy};
    // KinectFusionReconstruction.cpp:21-38
    k->world2camera = HMat4::identity();
    if (seeds)
        for (int q = 0; q < k->ncomp; ++q)
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j) k->world2camera.m[i][j].d[q] = seeds[(size_t) q * 16 + i * 4 + j];
    k->record.push_back(k->world2camera);
    k->world2volume = HMat4::identity();
    {
        const float ax = cfg->r_deg[0] / 180.0f * float(M_PI), ay = cfg->r_deg[1] / 180.0f * float(M_PI),
                    az = cfg->r_deg[2] / 180.0f * float(M_PI);
        HMat3 R = hmul(hmul(haxis_rotation(HJet(ax), 0), haxis_rotation(HJet(ay), 1)), haxis_rotation(HJet(az), 2));
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) k->world2volume.m[i][j] = HJet(R.m[i][j].v);
            k->world2volume.m[i][3] = HJet(cfg->init_xyz[i]);
        }
    }
    k->angle_thres = float(sin(cfg->angle_thres_deg / 180.f * M_PI));  // :58
    cudaError_t e = cudaStreamCreate(&k->stream);
    const int L = cfg->num_levels;
    k->depths.assign(L, nullptr);
    k->vmaps_curr.assign(L, nullptr);
    k->nmaps_curr.assign(L, nullptr);
    k->vmaps_prev.assign(L, nullptr);
    k->nmaps_prev.assign(L, nullptr);
    const size_t px = (size_t) cfg->width * cfg->height;
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaMalloc((void **) &k->d_depth2[i], px * sizeof(uint16_t));
    if (e == cudaSuccess) e = cudaMallocHost((void **) &k->h_depth, px * sizeof(uint16_t));
    for (int i = 0; i < L && e == cudaSuccess; ++i) {  // AllocateBuffers, :84-92
        const size_t r = cfg->height >> i, c = cfg->width >> i;
        e = cudaMalloc((void **) &k->depths[i], r * c * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc((void **) &k->vmaps_curr[i], map_floats(k, i, false) * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc((void **) &k->nmaps_curr[i], map_floats(k, i, false) * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc((void **) &k->vmaps_prev[i], map_floats(k, i, true) * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc((void **) &k->nmaps_prev[i], map_floats(k, i, true) * sizeof(float));
        if (e == cudaSuccess) e = cudaMemset(k->vmaps_prev[i], 0, map_floats(k, i, true) * sizeof(float));
        if (e == cudaSuccess) e = cudaMemset(k->nmaps_prev[i], 0, map_floats(k, i, true) * sizeof(float));
    }
    const size_t rec = (size_t) (1 + k->ncomp) * 16;
    if (e == cudaSuccess) e = cudaMalloc((void **) &k->d_record, rec * sizeof(float));
    if (e == cudaSuccess) e = cudaMallocHost((void **) &k->h_record, rec * sizeof(float));
    for (int i = 0; i < 12 && e == cudaSuccess; ++i) e = cudaEventCreate(&k->ev[i / 6][i % 6]);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&k->ev_head, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&k->ev_icp, cudaEventDisableTiming);
    const size_t pose_floats = (size_t) (1 + k->ncomp) * 12;
    if (e == cudaSuccess) e = cudaMalloc((void **) &k->d_pose_all, xs_kinfu::POSE_SLOTS * pose_floats * sizeof(float));
    if (e == cudaSuccess) e = cudaStreamCreate(&k->stream_real);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&k->ev_icp_start, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMallocHost((void **) &k->h_pose, pose_floats * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void **) &k->d_status, 2 * sizeof(int));
    if (e == cudaSuccess) e = cudaMallocHost((void **) &k->h_status, 2 * sizeof(int));
    if (e != cudaSuccess) {
        set_error(std::string("xs_kinfu_create: ") + cudaGetErrorString(e));
        xs_kinfu_destroy(k);
        return nullptr;
    }
    k->icp = icp_scratch_create();
    k->volume = comps == 2 ? xs_volume_create_hessian(cfg->res, cfg->voxel_size, cfg->thres_range, dirs, k->batch.v.m, (const int *) k->batch.h_pairs)
                           : xs_volume_create(cfg->res, cfg->voxel_size, cfg->thres_range, comps, dirs);  // :66-67
    if (!k->volume) {
        xs_kinfu_destroy(k);
        return nullptr;
    }
    return k;
}

void xs_kinfu_destroy(xs_kinfu *k) {
    if (!k) return;
    if (k->stream) cudaStreamSynchronize(k->stream);
    xs_volume_destroy(k->volume);
    if (k->stream_real) cudaStreamSynchronize(k->stream_real);
    icp_scratch_destroy(k->icp);
    batch_free(k->batch);
    cudaFree(k->d_depth2[0]);
    cudaFree(k->d_depth2[1]);
    cudaFreeHost(k->h_depth);
    for (size_t i = 0; i < k->depths.size(); ++i) {
        cudaFree(k->depths[i]);
        cudaFree(k->vmaps_curr[i]);
        cudaFree(k->nmaps_curr[i]);
        cudaFree(k->vmaps_prev[i]);
        cudaFree(k->nmaps_prev[i]);
    }
    if (k->stream_comm) {
        cudaStreamSynchronize(k->stream_comm);
        cudaStreamDestroy(k->stream_comm);
        cudaEventDestroy(k->ev_record);
        cudaEventDestroy(k->ev_gather[0]);
        cudaEventDestroy(k->ev_gather[1]);
    }
    cudaFree(k->d_gathered[0]);
    cudaFree(k->d_gathered[1]);
    cudaFree(k->d_record);
    cudaFreeHost(k->h_record);
    if (k->stream_real) cudaStreamSynchronize(k->stream_real);
    cudaFree(k->d_pose_all);
    if (k->ev_icp_start) cudaEventDestroy(k->ev_icp_start);
    if (k->stream_real) cudaStreamDestroy(k->stream_real);
    cudaFreeHost(k->h_pose);
    cudaFree(k->d_status);
    cudaFreeHost(k->h_status);
    cudaFree(k->d_icp_log);
    cudaFreeHost(k->h_icp_log);
    for (int i = 0; i < 12; ++i)
        if (k->ev[i / 6][i % 6]) cudaEventDestroy(k->ev[i / 6][i % 6]);
    if (k->ev_head) cudaEventDestroy(k->ev_head);
    if (k->ev_icp) cudaEventDestroy(k->ev_icp);
    if (k->stream) cudaStreamDestroy(k->stream);
    delete k;
}

// SurfaceMeasure, KinectFusionReconstruction.cpp:280-299
static int surface_measure_on(xs_kinfu *k, const uint16_t *d_depth, cudaStream_t stream);
int xs_kinfu_surface_measure(xs_kinfu *k, const uint16_t *d_depth) { return k ? surface_measure_on(k, d_depth, k->stream) : XS_ERR_ARG; }
static int surface_measure_on(xs_kinfu *k, const uint16_t *d_depth, cudaStream_t stream) {
    const xs_config &c = k->cfg;
    if (c.width <= 0 || c.height <= 0) {
        set_error("error::KinectFusionReconstruction, not created yet");
        return XS_ERR_ARG;
    }
    int rc = xs_bilateral_filter(d_depth, c.width * sizeof(uint16_t), c.height, c.width, k->depths[0], stream);
    for (int i = 1; i < c.num_levels && rc == XS_OK; ++i)
        rc = xs_pyr_down(k->depths[i - 1], c.height >> (i - 1), c.width >> (i - 1), k->depths[i], stream);
    for (int i = 0; i < c.num_levels && rc == XS_OK; ++i) {
        rc = xs_create_vmap(level_intr(k->intr, i), k->depths[i], c.height >> i, c.width >> i, k->vmaps_curr[i], stream);
        if (rc == XS_OK) rc = xs_create_nmap(k->vmaps_curr[i], c.height >> i, c.width >> i, k->nmaps_curr[i], stream);
        // parameters that move the intrinsics: the maps of the current frame carry derivative components.  The frame loop does
        // not need them stored (the ICP forms d vcurr analytically from the real vertex), so they are written only on request
        if (rc == XS_OK && k->keep_curr_derivs)
            rc = surface_derivs(k->depths[i], c.height >> i, c.width >> i, i, k->intr, k->batch.v, k->vmaps_curr[i], k->nmaps_curr[i], stream);
    }
    return rc;
}

// AlignDepthToReconstruction (after SurfaceMeasure) + PoseEstimate, KinectFusionReconstruction.cpp:161-235.
// Returns 1 when a pose was estimated, 0 on frame 0 or when the normal equations are degenerate.
// The Gauss-Newton loop (levels 2..0 with 3, 4, 5 iterations, :186-192) is queued on the stream without any host
// round trip: two kernels per iteration (association, derivative pass + device-side solve / pose update in its tail, icp.cu), one download of the
// final pose (all derivative components) and the degeneracy flag at the end.
int xs_kinfu_pose_estimate(xs_kinfu *k) {
    set_ctx(k);
    k->icp_iters_done = 0;
    k->icp_log.clear();
    icp_timing_reset(k->icp);
    if (k->use_gt_pose) return 1;  // mapping with known poses: AlignDepthToReconstruction returns before ICP, :164-166
    if (k->frame_id == 0) return 0;
    const xs_config &c = k->cfg;
    const int ncomp = k->ncomp;
    HMat4 c2w_prev = hinverse(k->record.back());
    HMat3 Rprev = hrotation(c2w_prev);
    HVec3 tprev = htranslation(c2w_prev);
    HMat3 Rprev_inv = hinverse(Rprev);
    xs_pose prev_pose;
    to_pose(Rprev_inv, tprev, ncomp, k->dR2, k->dt2, prev_pose);
    // initial estimate = previous pose (:182-185)
    const size_t pose_floats = (size_t) (1 + ncomp) * 12;
    for (int q = 0; q <= ncomp; ++q) {
        float *h = k->h_pose + (size_t) q * 12;
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) h[i * 3 + j] = q == 0 ? Rprev.m[i][j].v : Rprev.m[i][j].d[q - 1];
            h[9 + i] = q == 0 ? tprev.v[i].v : tprev.v[i].d[q - 1];
        }
    }
    KCUDA0(cudaMemcpyAsync(k->d_pose_slot(0), k->h_pose, pose_floats * sizeof(float), cudaMemcpyHostToDevice, k->stream));
    KCUDA0(cudaMemsetAsync(k->d_status, 0, 2 * sizeof(int), k->stream));
    // with derivative components the real chain (association + real step per iteration) runs ahead on its own stream
    const bool split = ncomp > 0 && !getenv("XS_ICP_NO_SPLIT");
    if (split) {
        KCUDA0(cudaEventRecord(k->ev_icp_start, k->stream));
        KCUDA0(cudaStreamWaitEvent(k->stream_real, k->ev_icp_start, 0));
    }
    const size_t log_stride = (size_t) 27 * (1 + ncomp);
    if (k->log_icp && !k->d_icp_log) {
        KCUDA0(cudaMalloc((void **) &k->d_icp_log, 16 * log_stride * sizeof(double)));
        KCUDA0(cudaMallocHost((void **) &k->h_icp_log, 16 * log_stride * sizeof(double)));
    }
    int it = 0;
    for (int level = c.num_levels - 1; level >= 0; --level) {
        const int rows = c.height >> level, cols = c.width >> level;
        for (int iter = 0; iter < k->icp_iterations[level]; ++iter, ++it) {
            const int rc = icp_iteration_async(k->icp, k->d_pose_slot(it), k->vmaps_curr[level], k->nmaps_curr[level], &prev_pose,
                                               level_intr(k->intr, level), k->vmaps_prev[level], k->nmaps_prev[level], rows,
                                               cols, k->batch, c.dist_thres, k->angle_thres, k->d_pose_slot(it + 1),
                                               k->solve_mode, k->d_status,
                                               k->log_icp && it < 16 ? k->d_icp_log + (size_t) it * log_stride : nullptr,
                                               k->stream, split ? k->stream_real : nullptr, it);
            if (rc != XS_OK) {
                cudaStreamSynchronize(k->stream_real);
                return 0;
            }
        }
    }
    KCUDA0(cudaMemcpyAsync(k->h_pose, k->d_pose_slot(it), pose_floats * sizeof(float), cudaMemcpyDeviceToHost, k->stream));
    KCUDA0(cudaMemcpyAsync(k->h_status, k->d_status, 2 * sizeof(int), cudaMemcpyDeviceToHost, k->stream));
    if (k->log_icp)
        KCUDA0(cudaMemcpyAsync(k->h_icp_log, k->d_icp_log, (size_t) (it < 16 ? it : 16) * log_stride * sizeof(double),
                              cudaMemcpyDeviceToHost, k->stream));
    if (k->next_depth) {
        // ProcessFrame: the pose-independent head of the integration runs on the device while the host turns the ICP
        // result into the volume-to-camera pose; the host waits for the download only
        KCUDA0(cudaEventRecord(k->ev_icp, k->stream));
        xs_volume_set_pipelined(k->volume, 1);
        if (integrate_prepare(k->volume, k->next_depth, c.width * sizeof(uint16_t), c.height, c.width, k->stream) != XS_OK)
            return 0;
        KCUDA0(cudaEventSynchronize(k->ev_icp));
    } else {
        KCUDA0(cudaStreamSynchronize(k->stream));
    }
    k->h_depth_free = true;  // everything queued before the ICP download has completed
    k->icp_iters_done = it;
    if (k->log_icp) {  // repack the 27 sums per component into A (36, column-major) + b (6), ICP.cu:419-428
        const int n = it < 16 ? it : 16;
        k->icp_log.assign((size_t) n * (1 + ncomp) * 42, 0.0);
        for (int i = 0; i < n; ++i)
            for (int q = 0; q <= ncomp; ++q) {
                const double *v = k->h_icp_log + (size_t) i * log_stride + (size_t) q * 27;
                double *A = &k->icp_log[((size_t) i * (1 + ncomp) + q) * 42], *b = A + 36;
                int shift = 0;
                for (int r = 0; r < 6; ++r)
                    for (int cc = r; cc < 7; ++cc) {
                        const double val = v[shift++];
                        if (cc == 6)
                            b[r] = val;
                        else
                            A[cc * 6 + r] = A[r * 6 + cc] = val;
                    }
            }
    }
    const int status = k->h_status[0] != 0 ? k->h_status[0] : k->h_status[1];
    if (status != 0) {  // A.real().determinant() guard, :203-210
        set_error(status == 2 ? "qnan det" : "eps det");
        return 0;
    }
    HMat4 c2w_curr = HMat4::identity();
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            c2w_curr.m[i][j].v = k->h_pose[i * 3 + j];
            for (int q = 0; q < ncomp; ++q) c2w_curr.m[i][j].d[q] = k->h_pose[(size_t) (1 + q) * 12 + i * 3 + j];
        }
        c2w_curr.m[i][3].v = k->h_pose[9 + i];
        for (int q = 0; q < ncomp; ++q) c2w_curr.m[i][3].d[q] = k->h_pose[(size_t) (1 + q) * 12 + 9 + i];
    }
    k->world2camera = hinverse(c2w_curr);  // :231
    k->record.push_back(k->world2camera);
    return 1;
}

// CalculatePointCloud, KinectFusionReconstruction.cpp:302-332, plus the pyramid of :272-276
int xs_kinfu_calculate_point_cloud(xs_kinfu *k) {
    set_ctx(k);
    const xs_config &c = k->cfg;
    HMat4 c2w = hinverse(k->world2camera);
    HMat4 c2v = hmul(k->world2volume, c2w);
    HMat4 v2w = hinverse(k->world2volume);
    xs_pose p_c2v, p_v2w;
    to_pose(hrotation(c2v), htranslation(c2v), k->ncomp, k->dR, k->dt, p_c2v);
    to_pose(hrotation(v2w), htranslation(v2w), k->ncomp, k->dR2, k->dt2, p_v2w);
    int rc = xs_raycast(k->volume, k->intr, &p_c2v, &p_v2w, c.height, c.width, k->vmaps_prev[0], k->nmaps_prev[0], k->stream);
    for (int i = 1; i < c.num_levels && rc == XS_OK; ++i) {
        rc = resize_map_batch(false, k->vmaps_prev[i - 1], c.height >> (i - 1), c.width >> (i - 1), k->batch.v, k->vmaps_prev[i], k->stream);
        if (rc == XS_OK)
            rc = resize_map_batch(true, k->nmaps_prev[i - 1], c.height >> (i - 1), c.width >> (i - 1), k->batch.v, k->nmaps_prev[i], k->stream);
    }
    return rc;
}

// IntegrateFrame, KinectFusionReconstruction.cpp:237-278
int xs_kinfu_integrate_frame(xs_kinfu *k, const uint16_t *d_depth) {
    set_ctx(k);
    const xs_config &c = k->cfg;
    if (k->use_gt_pose) {  // :239-247: world2camera = inverse(gt c2w) with zero imaginary part; the record's last entry is replaced
        if ((size_t) k->frame_id >= k->gt_poses.size() / 16) {
            set_error("IntegrateFrame: flag_use_gtPose is set but no ground-truth pose was given for this frame");
            return XS_ERR_ARG;
        }
        HMat4 gt = HMat4::identity();
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) gt.m[i][j] = HJet(k->gt_poses[(size_t) k->frame_id * 16 + i * 4 + j]);
        k->world2camera = hinverse(gt);
        k->record.back() = k->world2camera;
    }
    HMat4 c2w = hinverse(k->record.back());
    HMat4 c2v = hmul(k->world2volume, c2w);
    HMat4 v2c = hinverse(c2v);
    xs_pose p_v2c;
    to_pose(hrotation(v2c), htranslation(v2c), k->ncomp, k->dR, k->dt, p_v2c);
    return xs_integrate(k->volume, d_depth, c.width * sizeof(uint16_t), c.height, c.width, k->intr, c.max_weight, &p_v2c,
                        c.bi_threshold, k->stats, k->stream);
}

// ProcessFrame, KinectFusionReconstruction.cpp:147-159
int xs_kinfu_process_frame(xs_kinfu *k, const uint16_t *depth, int depth_on_device) {
    if (!k || !depth) return 0;
    const xs_config &c = k->cfg;
    const size_t bytes = (size_t) c.width * c.height * sizeof(uint16_t);
    const uint16_t *d_depth = depth;
    // Deferred mode with a frame in flight: the head of this frame (upload + surface measurement: it reads the new depth
    // frame and writes the current-frame maps, which the integration / raycast still running do not touch) is queued on the
    // second stream BEFORE waiting for that frame, so that it runs beside its raycast.
    const bool early = k->pending;
    cudaStream_t hs = early ? k->stream_real : k->stream;
    cudaEvent_t *ev = k->ev[k->ev_cur ^ 1];
    // the pinned staging frame is free once its previous upload is known to be complete (ICP result read, or a full wait)
    const bool early_copy = !depth_on_device && k->h_depth_free;
    if (early_copy) std::memcpy(k->h_depth, depth, bytes);
    if (!early || (!depth_on_device && !early_copy)) {
        if (finish_pending(k) != XS_OK) return 0;
        hs = k->stream;
    }
    if (!depth_on_device) {
        // the upload is outside the reference's timed region (main.cpp:51-57) but inside bench.py's e2e region
        if (!early_copy) std::memcpy(k->h_depth, depth, bytes);
        k->h_depth_free = false;
        k->depth_idx ^= 1;  // the frame before this one may still be integrating from the other buffer
        uint16_t *dd = k->d_depth2[k->depth_idx];
        if (cudaMemcpyAsync(dd, k->h_depth, bytes, cudaMemcpyHostToDevice, hs) != cudaSuccess) return 0;
        d_depth = dd;
    }
    long long l0 = g_launches;
    cudaEventRecord(ev[0], hs);
    if (surface_measure_on(k, d_depth, hs) != XS_OK) return 0;
    cudaEventRecord(ev[1], hs);
    if (hs != k->stream) {
        cudaEventRecord(k->ev_head, hs);
        if (finish_pending(k) != XS_OK) return 0;  // collects the previous frame from its own event set
        cudaStreamWaitEvent(k->stream, k->ev_head, 0);
    }
    cudaEventRecord(ev[5], k->stream);
    k->ev_cur ^= 1;
    k->launches[0] = g_launches - l0;
    l0 = g_launches;
    k->next_depth = d_depth;
    const int aligned = xs_kinfu_pose_estimate(k);
    k->next_depth = nullptr;
    const auto dbg_t0 = std::chrono::steady_clock::now();
    cudaEventRecord(ev[2], k->stream);
    k->launches[1] = g_launches - l0;
    l0 = g_launches;
    if (aligned < 0 || (k->frame_id > 0 && !aligned)) {
        fprintf(stderr, "Frame align failed!\n");
        xs_volume_set_pipelined(k->volume, 0);  // drops the queued integration head: the volume is not touched
        cudaStreamSynchronize(k->stream);
        return 0;
    }
    // integration and raycast are queued back to back (no host round trip between them): their only host input is the pose
    xs_volume_set_pipelined(k->volume, 1);
    const int rc_int = xs_kinfu_integrate_frame(k, d_depth);
    const auto dbg_t1 = std::chrono::steady_clock::now();
    cudaEventRecord(ev[3], k->stream);
    k->launches[2] = g_launches - l0;
    l0 = g_launches;
    const int rc_ray = rc_int == XS_OK ? xs_kinfu_calculate_point_cloud(k) : rc_int;
    xs_volume_set_pipelined(k->volume, 0);
    if (getenv("XS_TIMING")) {
        const auto dbg_t2 = std::chrono::steady_clock::now();
        fprintf(stderr, "[xs timing] host: integrate queued after %.1f us, raycast queued after %.1f us more\n",
                std::chrono::duration<double, std::micro>(dbg_t1 - dbg_t0).count(),
                std::chrono::duration<double, std::micro>(dbg_t2 - dbg_t1).count());
    }
    if (rc_int != XS_OK || rc_ray != XS_OK) {
        cudaStreamSynchronize(k->stream);
        return 0;
    }
    // per-frame derivative record for the multi-GPU gather
    for (int q = 0; q <= k->ncomp; ++q)
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j)
                k->h_record[(size_t) q * 16 + i * 4 + j] = q == 0 ? k->world2camera.m[i][j].v : k->world2camera.m[i][j].d[q - 1];
    if (k->comm && k->gather_in_flight) cudaStreamWaitEvent(k->stream, k->ev_gather[k->gather_slot], 0);  // the previous gather has read d_record
    cudaMemcpyAsync(k->d_record, k->h_record, (size_t) (1 + k->ncomp) * 16 * sizeof(float), cudaMemcpyHostToDevice, k->stream);
    if (k->comm) {  // derivatives gathered by NCCL all-gather over NVLink, beside the next frame's kernels
        cudaEventRecord(k->ev_record, k->stream);
        cudaStreamWaitEvent(k->stream_comm, k->ev_record, 0);
        k->gather_slot ^= 1;
        if (xs_comm_all_gather(k->comm, k->d_record, k->d_gathered[k->gather_slot], k->record_floats, k->stream_comm) != XS_OK) return 0;
        cudaEventRecord(k->ev_gather[k->gather_slot], k->stream_comm);
        k->gather_in_flight = true;
    }
    cudaEventRecord(ev[4], k->stream);
    k->launches[3] = g_launches - l0;
    if (k->deferred) {  // pose and status are final (the ICP result was read on the host); the volume and the maps follow
        k->pending = true;
        k->frame_id += k->frame_step;
        return 1;
    }
    if (cudaStreamSynchronize(k->stream) != cudaSuccess) {
        set_error("xs_kinfu_process_frame: stream synchronize failed");
        return 0;
    }
    k->h_depth_free = true;
    collect_frame(k);
    k->frame_id += k->frame_step;  // KinectFusionReconstruction.cpp:157
    return 1;
}

// Deferred mode (off by default).  ProcessFrame has two host round trips in the reference-shaped loop: the ICP result
// (needed on the host for the pose algebra) and the end of the frame.  With deferred != 0 the second one moves to the start
// of the next xs_kinfu_process_frame (or to xs_kinfu_sync): the call returns as soon as integration, raycast and the pyramid
// are queued, so whatever the caller does between frames (logging, the multi-GPU gather, the next frame's upload) overlaps
// the device work.  The pose / status it returns are final; statistics and stage times (xs_kinfu_get_times / _stats /
// _algorithmic_bytes, xs_volume_last_integrate_ms) describe the last COLLECTED frame until the next call or xs_kinfu_sync.
// A device-resident depth frame must stay valid until then.
int xs_kinfu_set_deferred(xs_kinfu *k, int on) {
    if (!k) return XS_ERR_ARG;
    const int rc = finish_pending(k);
    k->deferred = on != 0;
    return rc;
}

// Waits for everything queued on the pipeline's stream and collects the statistics of a deferred frame.
int xs_kinfu_sync(xs_kinfu *k) {
    if (!k) return XS_ERR_ARG;
    if (k->stream_comm && cudaStreamSynchronize(k->stream_comm) != cudaSuccess) return XS_ERR_CUDA;
    if (k->pending) return finish_pending(k);
    return cudaStreamSynchronize(k->stream) == cudaSuccess ? XS_OK : XS_ERR_CUDA;
}

int xs_kinfu_frame_id(const xs_kinfu *k) { return k ? k->frame_id : -1; }

int xs_kinfu_get_world2camera(const xs_kinfu *k, float *out) {
    if (!k || !out) return XS_ERR_ARG;
    for (int q = 0; q <= k->ncomp; ++q)
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j)
                out[(size_t) q * 16 + i * 4 + j] = q == 0 ? k->world2camera.m[i][j].v : k->world2camera.m[i][j].d[q - 1];
    return XS_OK;
}

// Replaces world2camera (real part and every derivative component) before the first frame: the general form of the
// reference's commented seeding line (KinectFusionReconstruction.cpp:22), e.g. to start from a relocalised pose.
int xs_kinfu_set_world2camera(xs_kinfu *k, const float *in) {
    if (!k || !in || k->frame_id != 0) {
        set_error("xs_kinfu_set_world2camera: only before the first frame");
        return XS_ERR_ARG;
    }
    set_ctx(k);
    for (int q = 0; q <= k->ncomp; ++q)
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                const float v = in[(size_t) q * 16 + i * 4 + j];
                if (q == 0)
                    k->world2camera.m[i][j].v = v;
                else
                    k->world2camera.m[i][j].d[q - 1] = v;
            }
    k->record.clear();
    k->record.push_back(k->world2camera);
    return XS_OK;
}

// Intrinsic parameters of a Hessian batch (BASELINE.json configs[3]: "Hessian w.r.t. pose + intrinsics"): dintr[nparams][4] =
// h d(fx, fy, cx, cy) / d theta_p.  New behaviour - the reference's Intr is plain floats (Internal.h:49-59): the current-frame
// vertex / normal maps then carry derivative components (Map.cu:8-70 on jets), the ICP row takes d s = dR vc + dt + R dvc, and
// the raycast ray ((x - cx) / fx, (y - cy) / fy, 1) its intrinsic derivatives (RayCaster.cu:56-62).  TSDF integration does not
// depend on the intrinsics with the nearest-neighbour depth look-up (xl = (image_x - cx) / fx = X / Z, TsdfFusion.cu:118-146);
// the bilinear branch is rejected.  Only before the first frame.
int xs_kinfu_set_intrinsic_seeds(xs_kinfu *k, const float *dintr) {
    if (!k || k->frame_id != 0 || k->comps != 2) {
        set_error("xs_kinfu_set_intrinsic_seeds: needs a Hessian batch (comps = 2), before the first frame");
        return XS_ERR_ARG;
    }
    if (k->cfg.bi_threshold > 0.f && dintr) {
        set_error("xs_kinfu_set_intrinsic_seeds: not implemented with biInterpolate_threshold > 0 (bilinear depth look-up)");
        return XS_ERR_ARG;
    }
    KCUDA(cudaStreamSynchronize(k->stream));
    int rc = batch_set_intrinsics(k->batch, dintr, k->intr.fx, k->intr.fy);
    if (rc == XS_OK) rc = xs_volume_set_intrinsic_seeds(k->volume, dintr);
    if (rc != XS_OK) return rc;
    return XS_OK;
}

// With intrinsic parameters the current-frame vertex / normal maps have derivative components (xs_kinfu_map then reports
// [(1 + ncurr)][3][rows][cols]).  The frame loop itself does not read them, so they are stored only when asked for.
int xs_kinfu_keep_current_map_derivatives(xs_kinfu *k, int on) {
    if (!k || k->frame_id != 0) {
        set_error("xs_kinfu_keep_current_map_derivatives: only before the first frame");
        return XS_ERR_ARG;
    }
    KCUDA(cudaStreamSynchronize(k->stream));
    k->keep_curr_derivs = on != 0 && k->batch.v.ncurr > 0;
    const int ncurr = k->keep_curr_derivs ? k->batch.v.ncurr : 0;
    for (int i = 0; i < k->cfg.num_levels; ++i) {  // current-frame maps: [(1 + ncurr)][3][rows][cols]
        const size_t floats = map_floats(k, i, false) * (size_t) (1 + ncurr);
        cudaFree(k->vmaps_curr[i]);
        cudaFree(k->nmaps_curr[i]);
        k->vmaps_curr[i] = k->nmaps_curr[i] = nullptr;
        KCUDA(cudaMalloc((void **) &k->vmaps_curr[i], floats * sizeof(float)));
        KCUDA(cudaMalloc((void **) &k->nmaps_curr[i], floats * sizeof(float)));
    }
    return XS_OK;
}

// gt_poses (camera-to-world, row-major 4x4 per frame) + flag_use_gtPose (KinectFusionReconstruction.h:36,82): with the flag
// set, frames are fused at the given poses and ICP is skipped (mapping mode of the relocalisation experiments).
int xs_kinfu_set_gt_poses(xs_kinfu *k, const float *poses16, int n, int use_gt_pose) {
    if (!k || n < 0 || (n > 0 && !poses16)) return XS_ERR_ARG;
    k->gt_poses.assign(poses16, poses16 + (size_t) n * 16);
    k->use_gt_pose = use_gt_pose != 0;
    return XS_OK;
}

int xs_kinfu_get_pose_c2w(const xs_kinfu *k, float *out16) {
    if (!k || !out16) return XS_ERR_ARG;
    set_ctx(k);
    HMat4 c2w = hinverse(k->record.back());  // main.cpp:61
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) out16[i * 4 + j] = c2w.m[i][j].v;
    return XS_OK;
}

xs_volume *xs_kinfu_volume(xs_kinfu *k) { return k ? k->volume : nullptr; }

const float *xs_kinfu_map(const xs_kinfu *k, int which, int level, int *rows, int *cols, int *ncomp) {
    if (!k || level < 0 || level >= k->cfg.num_levels) return nullptr;
    if (rows) *rows = k->cfg.height >> level;
    if (cols) *cols = k->cfg.width >> level;
    if (ncomp) *ncomp = (which >= 3) ? k->ncomp : (which >= 1 && k->keep_curr_derivs ? k->batch.v.ncurr : 0);
    switch (which) {
        case 0: return k->depths[level];
        case 1: return k->vmaps_curr[level];
        case 2: return k->nmaps_curr[level];
        case 3: return k->vmaps_prev[level];
        case 4: return k->nmaps_prev[level];
    }
    return nullptr;
}

int xs_kinfu_get_times(const xs_kinfu *k, float *ms10) {
    if (!k || !ms10) return XS_ERR_ARG;
    for (int i = 0; i < 5; ++i) {
        ms10[i] = k->ms[i];
        ms10[5 + i] = (float) k->launches[i];
    }
    return XS_OK;
}

int xs_kinfu_take_icp_log(xs_kinfu *k, double *out, int max_iters) {
    if (!k) return 0;
    const size_t stride = (size_t) (1 + k->ncomp) * 42;
    int n = (int) (k->icp_log.size() / stride);
    if (n > max_iters) n = max_iters;
    if (out) std::memcpy(out, k->icp_log.data(), (size_t) n * stride * sizeof(double));
    k->icp_log.clear();
    return n;
}

int xs_kinfu_enable_icp_log(xs_kinfu *k, int on) {
    if (!k) return XS_ERR_ARG;
    k->log_icp = on != 0;
    return XS_OK;
}

int xs_kinfu_get_stats(const xs_kinfu *k, unsigned long long *out4) {
    if (!k || !out4) return XS_ERR_ARG;
    std::memcpy(out4, k->stats, sizeof(k->stats));
    return XS_OK;
}

// Algorithmic bytes of the last frame (following SURVEY.md §8d): FP32 storage, D = ncomp.
int xs_kinfu_get_algorithmic_bytes(const xs_kinfu *k, double *out4) {
    if (!k || !out4) return XS_ERR_ARG;
    const double D = k->ncomp, P0 = (double) k->cfg.width * k->cfg.height;
    double P = 0;
    for (int i = 0; i < k->cfg.num_levels; ++i) P += P0 / double(1 << (2 * i));
    out4[0] = 2 * P0 + 2 * 4 * P + 12 * P + 2 * 12 * P;  // surface measurement (real maps)
    double icp = 0;
    if (k->icp_iters_done > 0) {
        int done = 0;
        for (int level = k->cfg.num_levels - 1; level >= 0; --level)
            for (int it = 0; it < k->icp_iterations[level] && done < k->icp_iters_done; ++it, ++done)
                icp += P0 / double(1 << (2 * level)) * (48 + 24 * D) + 27 * 8 * (1 + D);
    }
    out4[1] = icp;
    // integration: RMW of value + weight for every updated voxel, RMW of the D derivative planes for the voxels whose
    // planes can be non-zero (truncation-band voxels and saturated voxels of live bricks; the rest are exactly zero)
    out4[2] = (double) k->stats[0] * 2 * (4 + 4) + (double) k->stats[2] * 2 * 4 * D + 2 * P0;
    // raycast: output maps + pyramid (march / hit gathers are data dependent and reported separately)
    out4[3] = 24 * (1 + D) * P0 + 2 * 12 * (1 + D) * 5 * (P - P0);
    return XS_OK;
}

float *xs_kinfu_pose_record_device(xs_kinfu *k) { return k ? k->d_record : nullptr; }

// se3Exp (KinectFusionReconstruction.h:176-219) on batched jets: the parameterisation of the reference's relocalisation /
// pose-set experiments.  xi: [(1 + ncomp)][6] = (v, omega), component 0 real, the rest h-scaled derivative components of the
// batch kind (comps, dirs, pairs as for xs_kinfu_create / _create_hessian); T_out: [(1 + ncomp)][16] row-major 4x4.
// Like the reference, |omega| < 1e-6 takes the first-order branch R = V = I + omega^ (the reference's norm includes the
// imaginary parts, which are h-sized; the branch here is decided on the real part).
int xs_se3_exp(const float *xi, int comps, int dirs, int npairs, const int *pairs, float *T_out) {
    if (!xi || !T_out || (comps != 1 && comps != 2 && comps != 3) || dirs < 0) return XS_ERR_ARG;
    std::vector<HPair> hp;
    if (comps == 2) {
        if (pairs)
            for (int k = 0; k < npairs; ++k) hp.push_back(HPair{pairs[2 * k], pairs[2 * k + 1]});
        else
            for (int i = 0; i < dirs; ++i)
                for (int j = i; j < dirs; ++j) hp.push_back(HPair{i, j});
    }
    const HJetCtx saved = hj_ctx();
    hj_ctx().comps = comps;
    hj_ctx().dirs = dirs;
    hj_ctx().npairs = (int) hp.size();
    hj_ctx().pairs = hp.empty() ? nullptr : hp.data();
    const int ncomp = hj_ctx().ncomp();
    if (ncomp > HJ_MAX) {
        hj_ctx() = saved;
        set_error("xs_se3_exp: at most 256 derivative components");
        return XS_ERR_ARG;
    }
    {
        HJet x[6];
        for (int e = 0; e < 6; ++e) {
            x[e] = HJet(xi[e]);
            for (int q = 0; q < ncomp; ++q) x[e].d[q] = xi[(size_t) (1 + q) * 6 + e];
        }
        const HJet *v = x, *w = x + 3;
        HMat3 W;  // omega^
        W.m[0][1] = -w[2], W.m[0][2] = w[1], W.m[1][2] = -w[0];
        W.m[1][0] = w[2], W.m[2][0] = -w[1], W.m[2][1] = w[0];
        HMat3 R, V;
        for (int i = 0; i < 3; ++i) R.m[i][i] = V.m[i][i] = HJet(1.f);
        const float nrm = std::sqrt(w[0].v * w[0].v + w[1].v * w[1].v + w[2].v * w[2].v);
        if (nrm < 1e-6f) {
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) R.m[i][j] = R.m[i][j] + W.m[i][j], V.m[i][j] = V.m[i][j] + W.m[i][j];
        } else {
            const HJet theta = hj_sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
            const HJet sn = hj_sin(theta), cs = hj_cos(theta);
            const HMat3 W2 = hmul(W, W);
            const HJet A = sn / theta, B = (HJet(1.f) - cs) / (theta * theta), C = (theta - sn) / (theta * theta * theta);
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) {
                    R.m[i][j] = R.m[i][j] + A * W.m[i][j] + B * W2.m[i][j];
                    V.m[i][j] = V.m[i][j] + B * W.m[i][j] + C * W2.m[i][j];
                }
        }
        HVec3 vv;
        for (int i = 0; i < 3; ++i) vv.v[i] = v[i];
        const HVec3 t = hmul(V, vv);
        for (int q = 0; q <= ncomp; ++q) {
            float *T = T_out + (size_t) q * 16;
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) T[i * 4 + j] = q == 0 ? R.m[i][j].v : R.m[i][j].d[q - 1];
                T[i * 4 + 3] = q == 0 ? t.v[i].v : t.v[i].d[q - 1];
            }
            T[12] = T[13] = T[14] = 0.f;
            T[15] = q == 0 ? 1.f : 0.f;
        }
    }
    hj_ctx() = saved;
    return XS_OK;
}

// Attaches a communicator (comm.cpp): from the next frame on, every rank's record - (1 + ncomp) 4x4 matrices, padded to
// record_floats, which must be the same on every rank and >= the largest (1 + ncomp) * 16 - is all-gathered after each
// processed frame.  The gather is queued by ProcessFrame on an internal stream behind the frame's record upload.
int xs_kinfu_set_comm(xs_kinfu *k, xs_comm *comm, int record_floats) {
    if (!k) return XS_ERR_ARG;
    if (k->stream_comm) cudaStreamSynchronize(k->stream_comm);
    k->gather_in_flight = false;
    k->comm = comm;
    if (!comm) return XS_OK;
    if (record_floats < (1 + k->ncomp) * 16) {
        set_error("xs_kinfu_set_comm: record_floats is smaller than this rank's record");
        k->comm = nullptr;
        return XS_ERR_ARG;
    }
    KCUDA(cudaStreamSynchronize(k->stream));
    if (!k->stream_comm) {
        KCUDA(cudaStreamCreateWithFlags(&k->stream_comm, cudaStreamNonBlocking));
        KCUDA(cudaEventCreateWithFlags(&k->ev_record, cudaEventDisableTiming));
        KCUDA(cudaEventCreateWithFlags(&k->ev_gather[0], cudaEventDisableTiming));
        KCUDA(cudaEventCreateWithFlags(&k->ev_gather[1], cudaEventDisableTiming));
    }
    // the send buffer is the record itself, re-allocated at the padded size (padding stays zero)
    cudaFree(k->d_record);
    cudaFree(k->d_gathered[0]);
    cudaFree(k->d_gathered[1]);
    k->d_record = k->d_gathered[0] = k->d_gathered[1] = nullptr;
    k->record_floats = record_floats;
    KCUDA(cudaMalloc((void **) &k->d_record, (size_t) record_floats * sizeof(float)));
    KCUDA(cudaMemset(k->d_record, 0, (size_t) record_floats * sizeof(float)));
    for (int b = 0; b < 2; ++b) {
        KCUDA(cudaMalloc((void **) &k->d_gathered[b], (size_t) xs_comm_world(comm) * record_floats * sizeof(float)));
        KCUDA(cudaMemset(k->d_gathered[b], 0, (size_t) xs_comm_world(comm) * record_floats * sizeof(float)));
    }
    return XS_OK;
}

// The gathered records of the last processed frame, [world][record_floats]: waits for that frame's all-gather only.
int xs_kinfu_get_gathered_records(xs_kinfu *k, float *host_out) { return xs_kinfu_get_gathered_records_lagged(k, 0, host_out); }
// lag = 1: the records of the frame BEFORE the last processed one.  Their all-gather ran beside the last frame's kernels, so a
// consumer that reads one frame late (process frame f, then read frame f - 1) never waits for a collective.
int xs_kinfu_get_gathered_records_lagged(xs_kinfu *k, int lag, float *host_out) {
    if (!k || !host_out || !k->comm || (lag != 0 && lag != 1)) return XS_ERR_ARG;
    const int slot = k->gather_slot ^ lag;
    if (k->gather_in_flight) KCUDA(cudaEventSynchronize(k->ev_gather[slot]));
    KCUDA(cudaMemcpy(host_out, k->d_gathered[slot], (size_t) xs_comm_world(k->comm) * k->record_floats * sizeof(float), cudaMemcpyDeviceToHost));
    return XS_OK;
}
const float *xs_kinfu_gathered_records_device(xs_kinfu *k) { return k ? k->d_gathered[k->gather_slot] : nullptr; }

void *xs_kinfu_stream(xs_kinfu *k) { return k ? (void *) k->stream : nullptr; }

}  // extern "C"
