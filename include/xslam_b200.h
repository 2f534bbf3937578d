/* xslam_b200.h — C-ABI of libxslam_b200.so (B200 / sm_100a).
 *
 * Drop-in boundary for the CSFD/DCSFD-differentiated KinectFusion frame loop of MisEty/X-SLAM.
 * The reference has no FFI layer; its seam is the set of C++ free functions that
 * XKinectFusion/src/KinectFusionReconstruction.cpp calls (aggregated by
 * XKinectFusion/include/CudaFunctions.h:4-8).  Every entry point below names the reference
 * interface it replaces (file:line relative to the reference tree).  C++ wrappers with the
 * reference's exact signatures live in include/xslam_b200.hpp and forward to these symbols.
 *
 * Conventions
 *  - plain pointers and sizes only; `stream` is a cudaStream_t passed as void* (NULL = default).
 *  - every function returns 0 on success or a negative xs_status; xs_last_error() gives text.
 *    There is NO CPU fallback: without a CUDA device every compute call returns XS_ERR_CUDA.
 *  - "ncomp" = number of derivative components carried besides the real value.  Batch kinds ("comps"):
 *        comps = 1  CSFD list: dirs independent first-order directions (eps), ncomp = dirs;
 *        comps = 3  DCSFD list: dirs independent bicomplex directions (eps1, eps2, eps1eps2), ncomp = 3 * dirs;
 *        comps = 2  Hessian batch over dirs = n parameters: n first-order components F_i followed by one second-order
 *                   component S_k per listed parameter pair (i <= j; default all n (n + 1) / 2), ncomp = n + pairs.  It
 *                   equals the DCSFD list of those pairs with (eps1, eps2, eps1eps2) = (F_i, F_j, S_ij), but stores and
 *                   moves every first-order plane once (65 planes instead of 165 for the Hessian of 10 parameters).
 *    Derivative components are stored h-scaled exactly like the reference's imaginary parts
 *    (Internal.h:33, H_ = 1e-7): eps = h*d/dtheta, eps1eps2 = h^2 * d2/dtheta1 dtheta2.
 *  - maps are packed SoA: float[(1+ncomp)][3][rows][cols]  (component 0 = real part; x|y|z planes),
 *    replacing the reference's interleaved pitched devComplex maps (Internal.h:31).
 *  - the TSDF volume lives behind an opaque handle in a brick-tiled layout (8x8x8 voxels):
 *    value[brick][512], weight[brick][512], deriv[brick][ncomp][512]; seam-compatible
 *    value/weight/grad planes (TsdfVolume.h:27-36) are obtained with xs_volume_export_planes.
 */
#ifndef XSLAM_B200_H
#define XSLAM_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    XS_OK = 0,
    XS_ERR_ARG = -1,
    XS_ERR_CUDA = -2,
    XS_ERR_ICP_DEGENERATE = -3, /* |det(Re A)| < 1e-15 or NaN, KinectFusionReconstruction.cpp:203-210 */
    XS_ERR_NCCL = -4
} xs_status;

const char *xs_last_error(void);
int xs_version(void);
/* number of CUDA kernels this library has launched in the calling process (bench.py's gpu_launches) */
long long xs_launch_count(void);

/* Intr, Internal.h:49-59 */
typedef struct {
    float fx, fy, cx, cy;
} xs_intr;

/* A rigid transform with batched derivative components (host memory).  Replaces the by-value
 * MatS33 + devComplex3 kernel arguments (Internal.h:63-65,146-148): R row-major. */
typedef struct {
    float R[9];
    float t[3];
    int ncomp;
    const float *dR; /* [ncomp][9] or NULL */
    const float *dt; /* [ncomp][3] or NULL */
} xs_pose;

/* ---------------------------------------------------------------- number types (a1-a3)
 * Packed-SoA bicomplex arrays: float[4][n] = value | eps1 | eps2 | eps1eps2 planes, the device
 * counterpart of DoubleComplex (DeviceArray/include/DoubleComplex.h:15-95) and d_complex<T>
 * (DeviceArray/include/cuda_double_complex.hpp:16-134).  ops follow DoubleComplex.cpp. */
typedef enum {
    XS_DC_ADD = 0, XS_DC_SUB, XS_DC_MUL, XS_DC_DIV, XS_DC_SQRT, XS_DC_EXP, XS_DC_LOG, XS_DC_SIN, XS_DC_COS,
    XS_DC_ATAN2, XS_DC_POW, XS_DC_ATAN
} xs_dc_op;
int xs_dc_apply(int op, const float *d_a, const float *d_b, float p, float *d_out, long n, void *stream);
/* Experiments/test_CSFD/main.cpp:194-205: loss = f1(t*t, sin t) with t seeded (t,h | h,0) */
int xs_dc_chain(const float *d_t, float h, float *d_out, long n, void *stream);

/* The host-side number type (include/xslam_dcomplex.hpp, the API surface of DoubleComplex.h:15-95) evaluated element-wise on host
 * AoS arrays, n x (re.re, re.im, im.re, im.im): host code like the reference's DoubleComplex.cpp, used by tests and bindings. */
int xs_dc_host_apply(int op, const float *a_aos, const float *b_aos, float p, float *out_aos, long n);

/* Owning packed-SoA bicomplex array (a4): the semantics of the reference's DeviceArray<T> (DeviceArray/include/device_array.hpp:25-134,
 * src/device_memory.cpp:72-178) for the new layout - resize is a no-op when the size is unchanged (create), upload / download are
 * blocking and synchronise, copy is a deep copy that (re)creates the destination (copyTo).  Host data is AoS, n x
 * (re.re, re.im, im.re, im.im), i.e. an array of the reference's DoubleComplex objects (DoubleComplex.h:15-19) or of
 * xslam_b200::DoubleComplex (include/xslam_dcomplex.hpp); the device side is float[4][n] as xs_dc_apply takes it. */
typedef struct xs_dc_array xs_dc_array;
xs_dc_array *xs_dc_array_create(long n);
void xs_dc_array_release(xs_dc_array *a);
int xs_dc_array_resize(xs_dc_array *a, long n);
long xs_dc_array_size(const xs_dc_array *a);
float *xs_dc_array_ptr(xs_dc_array *a);
int xs_dc_array_upload(xs_dc_array *a, const float *host_aos, long n);
int xs_dc_array_download(const xs_dc_array *a, float *host_aos);
int xs_dc_array_copy(const xs_dc_array *src, xs_dc_array *dst);

/* ---------------------------------------------------------------- surface measurement (a6) */
/* bilateralFilter, Map.h:16 / Map.cu:262.  out: float[rows][cols] (the reference's imaginary part is 0) */
int xs_bilateral_filter(const uint16_t *d_depth, size_t depth_step_bytes, int rows, int cols, float *d_out,
                        void *stream);
/* pyrDown, Map.h:22 / Map.cu:274 */
int xs_pyr_down(const float *d_src, int rows, int cols, float *d_dst, void *stream);
/* createVMap, Map.h:29 / Map.cu:73.  d_vmap: float[3][rows][cols]; invalid pixels get NaN in x, 0 in y,z */
int xs_create_vmap(xs_intr intr, const float *d_depth, int rows, int cols, float *d_vmap, void *stream);
/* createNMap, Map.h:35 / Map.cu:89 */
int xs_create_nmap(const float *d_vmap, int rows, int cols, float *d_nmap, void *stream);
/* resizeVMap / resizeNMap, Map.h:47,54 / Map.cu:252,257.  in: [(1+ncomp)][3][rows][cols] */
int xs_resize_vmap(const float *d_in, int rows, int cols, int comps, int dirs, float *d_out, void *stream);
int xs_resize_nmap(const float *d_in, int rows, int cols, int comps, int dirs, float *d_out, void *stream);

/* Seam layout <-> packed SoA (a4).  The reference's MapArr is a pitched DeviceArray2D<devComplex> (interleaved re, im;
 * Internal.h:31) with `nplanes` planes stacked by rows (3 for vertex / normal maps, 1 for depth).  Component 0 of the
 * SoA map [(1+ncomp)][nplanes][rows][cols] is the real part; the imaginary part maps to derivative component `comp`
 * (comp < 0: dropped on import / written as zero on export).  These are what the C++ wrappers with the reference's
 * signatures (include/xslam_b200.hpp) use at the boundary. */
int xs_map_complex_to_soa(const void *d_src, size_t step_bytes, int nplanes, int rows, int cols, float *d_soa, int ncomp,
                          int comp, void *stream);
int xs_map_soa_to_complex(const float *d_soa, int ncomp, int comp, int nplanes, int rows, int cols, void *d_dst,
                          size_t step_bytes, void *stream);

/* ---------------------------------------------------------------- TSDF volume (a9, a11) */
typedef struct xs_volume xs_volume;
/* TsdfVolume::TsdfVolume, TsdfVolume.cpp:11-29 (trunc = max(voxel*thres_range, 2.1*voxel)); res multiple of 8 */
xs_volume *xs_volume_create(const int res[3], float voxel_size, float thres_range, int comps, int dirs);
/* Hessian batch (comps = 2) with an explicit pair list: pairs = int[npairs][2], 0 <= i <= j < nparams, sorted by i (NULL: all) */
xs_volume *xs_volume_create_hessian(const int res[3], float voxel_size, float thres_range, int nparams, int npairs, const int *pairs);
/* parameters of a Hessian batch that move the intrinsics: dintr[nparams][4] = h d(fx, fy, cx, cy) / d theta_p (see xs_kinfu_set_intrinsic_seeds) */
int xs_volume_set_intrinsic_seeds(xs_volume *v, const float *dintr);
void xs_volume_destroy(xs_volume *v);
/* initVolume, TsdfVolume.h:16 / TsdfFusion.cu:34 */
int xs_volume_reset(xs_volume *v, void *stream);
float xs_volume_trunc_dist(const xs_volume *v);
size_t xs_volume_bytes(const xs_volume *v);
/* device duration (CUDA events on the launching stream) of the last integration kernel, in ms */
float xs_volume_last_integrate_ms(const xs_volume *v);
/* the same for the last raycast hit kernel, and the pixels with a valid vertex / valid normal it found (out2) */
float xs_volume_last_raycast_hit_ms(const xs_volume *v);
int xs_volume_raycast_stats(const xs_volume *v, unsigned long long *out2);
/* Seam views (TsdfVolume::value/weight/grad, TsdfVolume.h:46-49): dense x-fastest planes [z][y][x]
 * on the device; comp = derivative component index in [0, ncomp) for d_grad (ignored when NULL). */
int xs_volume_export_planes(const xs_volume *v, int comp, float *d_value, int *d_weight, float *d_grad, void *stream);
int xs_volume_import_planes(xs_volume *v, int comp, const float *d_value, const int *d_weight, const float *d_grad,
                            void *stream);
/* integrateTsdfVolume, TsdfFusion.h:40-45 / TsdfFusion.cu:173.  v2c = (Rv2c, tv2c) with ncomp derivative
 * components.  stats (optional, 4 values, device-written then copied): [0] updated voxels, [1] bricks surviving the
 * frustum cull, [2] updated voxels whose derivative planes were read and written (band voxels + saturated voxels of
 * bricks that hold non-zero derivatives; all other derivative planes are exactly zero and are not touched). */
int xs_integrate(xs_volume *v, const uint16_t *d_depth, size_t depth_step_bytes, int rows, int cols, xs_intr intr,
                 int max_weight, const xs_pose *v2c, float bilinear_threshold, unsigned long long *stats_host,
                 void *stream);
/* Frame-loop mode: with pipelined != 0, xs_integrate and xs_raycast only queue work (no host synchronisation, separate pose
 * staging slots); the caller synchronises the stream once per frame and then calls xs_volume_finish_frame to collect the
 * integration statistics / kernel time.  Off by default (the seam-level calls synchronise like the reference's). */
int xs_volume_set_pipelined(xs_volume *v, int on);
int xs_volume_finish_frame(xs_volume *v, unsigned long long *stats_host4);
/* raycast, RayCaster.h:21-25 / RayCaster.cu:327.  c2v = (Rc2v, tc2v), v2w = (Rv2w, tv2w).
 * outputs: [(1+ncomp)][3][rows][cols] world-frame vertex / normal maps. */
int xs_raycast(const xs_volume *v, xs_intr intr, const xs_pose *c2v, const xs_pose *v2w, int rows, int cols,
               float *d_vmap, float *d_nmap, void *stream);
/* ComputeLocalTsdf_hessian, TsdfFusion.h:55-60 / TsdfFusion.cu:286 (DCSFD volume loss, one bicomplex
 * direction): pose = (Rv2c, tv2c) with ncomp = 3 (eps1, eps2, eps1eps2).  d_gt: dense [z][y][x].
 * out4 = {sum loss, sum grad, sum hessian, count} (host). */
int xs_tsdf_hessian(const uint16_t *d_depth, size_t depth_step_bytes, int rows, int cols, xs_intr intr,
                    const int res[3], float voxel_size, const xs_pose *v2c, float trunc, const float *d_gt,
                    double *out4_host, void *stream);
/* The same loss for dirs bicomplex directions in one call (BASELINE.json configs[4]: derivatives w.r.t. a multi-frame pose
 * set): v2c carries ncomp = 3 * dirs components; the ground-truth volume is streamed once per 8 directions instead of once per
 * direction.  out_host: double[dirs][4]; every row is bit-identical to an xs_tsdf_hessian call with that direction. */
int xs_tsdf_hessian_batch(const uint16_t *d_depth, size_t depth_step_bytes, int rows, int cols, xs_intr intr,
                          const int res[3], float voxel_size, const xs_pose *v2c, float trunc, const float *d_gt,
                          double *out_host, void *stream);
/* ComputeLocalTsdf_loss, TsdfFusion.h:48-52 / TsdfFusion.cu:409 (real-only twin of the DCSFD volume loss, used with
 * se3Exp-parameterised pose sets in relocalisation-style optimisation): (Rv2c row-major, tv2c) are plain floats.
 * out2 = {sum loss, count} (host). */
int xs_tsdf_loss(const uint16_t *d_depth, size_t depth_step_bytes, int rows, int cols, xs_intr intr, const int res[3],
                 float voxel_size, const float Rv2c[9], const float tv2c[3], float trunc, const float *d_gt,
                 double *out2_host, void *stream);
/* extractPoints / extractNormals, ExtractPointCloud.h:19-23 (real-only output path) */
long xs_extract_points(const xs_volume *v, float *d_points_xyz, float *d_normals_xyz, long max_points, void *stream);
/* extractNormals alone, ExtractPointCloud.h:22-23 / ExtractPointCloud.cu:342-362: trilinear central differences at +-1 voxel
 * divided by the SQUARED norm (reference quirk, :305-306) for n given points (device, xyz interleaved). */
int xs_extract_normals(const xs_volume *v, const float *d_points_xyz, float *d_normals_xyz, long n, void *stream);

/* ---------------------------------------------------------------- ICP (a7) */
/* estimateCombined, ICP.h:24-31 / ICP.cu:365.  curr = (Rcurr, tcurr), prev = (Rprev_inv, tprev).
 * Current-frame maps are real (3 planes); previous maps carry ncomp components.
 * A_host: double[(1+ncomp)][36] column-major 6x6, b_host: double[(1+ncomp)][6]; component 0 is the real part. */
int xs_estimate_combined(const xs_pose *curr, const float *d_vmap_curr, const float *d_nmap_curr, const xs_pose *prev,
                         xs_intr intr, const float *d_vmap_g_prev, const float *d_nmap_g_prev, int rows, int cols,
                         int comps, int dirs, float dist_thres, float angle_thres, double *A_host, double *b_host,
                         void *stream);
/* computeOptimizeMatrix, ICP.h:34-40 / ICP.cu:431-489 (inputs of the second-order pose optimiser the header reserves for
 * PoseNewtonEstimate): same data association as estimateCombined, real parts only.  jacobi_host: double[3][4] row-major
 * (jacobi_host(i, j)); hessian_host: double[12][12], index i * 4 + j on both sides (hessian_host[i1][j1](i2, j2)).
 * Returns the number of correspondences, or -1 on error. */
long xs_compute_optimize_matrix(const xs_pose *curr, const float *d_vmap_curr, const float *d_nmap_curr, const xs_pose *prev,
                                xs_intr intr, const float *d_vmap_g_prev, const float *d_nmap_g_prev, int rows, int cols,
                                float dist_thres, float angle_thres, double *jacobi_host, double *hessian_host, void *stream);

/* Device durations (CUDA events on the launching stream) of the derivative-accumulation kernel launches of the last
 * xs_kinfu_pose_estimate / xs_estimate_combined, with the pixel count of each launch (roofline timing). */
int xs_icp_deriv_times(float *ms, int *npix, int max_n);

/* ---------------------------------------------------------------- pipeline (a4, a8, a13) */
/* The YAML keys read by KinectFusionReconstruction::SetYamlParameters (KinectFusionReconstruction.cpp:12-72) */
typedef struct {
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
    int frame_step; /* KinectFusionReconstruction.cpp:72,157: frame_id advances by it (gt_poses[frame_id], dataset index); <= 0 reads as 1 */
} xs_config;

typedef enum {
    XS_SOLVE_EIGEN_LLT = 0, /* Hermitian LLT of the complex-symmetric A, as KinectFusionReconstruction.cpp:211; with comps=3 each
                               first-order component (eps1, eps2) is solved that way - what a one-direction complex run of the
                               reference yields - and eps1eps2 is the truncated-algebra second derivative on top of them */
    XS_SOLVE_ANALYTIC = 1   /* derivative of x = A^-1 b by the truncated algebra */
} xs_solve_mode;

typedef struct xs_kinfu xs_kinfu;
/* KinectFusionReconstruction() + SetYamlParameters, KinectFusionReconstruction.cpp:4-73.
 * seeds: [dirs*comps][16] derivative components of the initial world2camera (row-major 4x4, h-scaled),
 * the generalisation of the commented seeding line KinectFusionReconstruction.cpp:22; NULL = zeros. */
xs_kinfu *xs_kinfu_create(const xs_config *cfg, int comps, int dirs, const float *seeds, int solve_mode);
/* Hessian batch (comps = 2) with an explicit pair list (multi-GPU: every rank carries the nparams first-order components and its
 * share of the pairs).  pairs = int[npairs][2] sorted by i (NULL: all pairs); seeds: [(nparams + npairs)][16]: h G_i for the
 * parameters, h^2 (G_i G_j + G_j G_i) / 2 for the pairs when the parameters are se3Exp coordinates. */
xs_kinfu *xs_kinfu_create_hessian(const xs_config *cfg, int nparams, int npairs, const int *pairs, const float *seeds, int solve_mode);
void xs_kinfu_destroy(xs_kinfu *k);
/* Intrinsic parameters (BASELINE.json configs[3]: Hessian w.r.t. pose + intrinsics; new behaviour, the reference's Intr is plain
 * floats, Internal.h:49-59).  Hessian batches only, before the first frame: dintr[nparams][4] = h d(fx, fy, cx, cy) / d theta_p for
 * every parameter (zero rows for pure pose parameters).  The current-frame vertex / normal maps then carry derivative components
 * for these parameters and their pairs (xs_kinfu_map reports them), the ICP and the raycast differentiate through them.  With
 * biInterpolate_threshold = 0 the TSDF integration does not depend on the intrinsics; the bilinear branch is rejected. */
int xs_kinfu_set_intrinsic_seeds(xs_kinfu *k, const float *dintr);
/* stores the derivative components of the current-frame vertex / normal maps (xs_kinfu_map, which = 1, 2) - the frame loop itself
 * forms them on the fly, so they are kept only on request; before the first frame, after xs_kinfu_set_intrinsic_seeds */
int xs_kinfu_keep_current_map_derivatives(xs_kinfu *k, int on);
/* ProcessFrame, KinectFusionReconstruction.cpp:147-159.  depth: 640x480 uint16 mm, dense; host pointer
 * unless depth_on_device != 0.  Returns 1 on success, 0 when frame alignment failed (as the reference). */
int xs_kinfu_process_frame(xs_kinfu *k, const uint16_t *depth, int depth_on_device);
/* Deferred mode (off by default; the reference's ProcessFrame is synchronous, KinectFusionReconstruction.cpp:147-159 with the
 * syncs of resizeVMap / resizeNMap, Map.cu:234-260).  With on != 0 xs_kinfu_process_frame returns once integration, raycast and
 * the pyramid are queued on the stream - pose and return status are final, the volume and the maps follow in stream order -
 * and the end-of-frame wait moves to the start of the next xs_kinfu_process_frame or to xs_kinfu_sync.  Statistics and stage
 * times then describe the last collected frame; a device-resident depth frame must stay valid until that point. */
int xs_kinfu_set_deferred(xs_kinfu *k, int on);
/* waits for all queued work of the pipeline and collects the statistics of a deferred frame */
int xs_kinfu_sync(xs_kinfu *k);
/* stage entry points, KinectFusionReconstruction.h:113-141 */
int xs_kinfu_surface_measure(xs_kinfu *k, const uint16_t *d_depth);
int xs_kinfu_pose_estimate(xs_kinfu *k);
int xs_kinfu_integrate_frame(xs_kinfu *k, const uint16_t *d_depth);
int xs_kinfu_calculate_point_cloud(xs_kinfu *k);
int xs_kinfu_frame_id(const xs_kinfu *k);
/* world2camera (KinectFusionReconstruction.h:31): out [(1+ncomp)][16] row-major */
int xs_kinfu_get_world2camera(const xs_kinfu *k, float *out);
/* sets world2camera (real part + all derivative components, [(1+ncomp)][16]) before the first frame */
int xs_kinfu_set_world2camera(xs_kinfu *k, const float *in);
/* camera-to-world of the last frame, real part, as written to frame-%06d.pose.txt (main.cpp:61) */
int xs_kinfu_get_pose_c2w(const xs_kinfu *k, float *out16);
xs_volume *xs_kinfu_volume(xs_kinfu *k);
/* which: 0 depth pyramid, 1 vmap_curr, 2 nmap_curr, 3 vmap_g_prev, 4 nmap_g_prev; device pointer + dims */
const float *xs_kinfu_map(const xs_kinfu *k, int which, int level, int *rows, int *cols, int *ncomp);
/* per-stage device times of the last collected frame (ms): surface, icp, integrate, raycast+resize, total (= their sum);
 * followed by per-stage kernel-launch counts (5 more floats).  icp / integrate / raycast are brackets on the pipeline's
 * stream; in deferred mode the surface measurement of a frame runs beside the previous frame's raycast (second stream), so its
 * bracket is not part of the frame's critical path. */
int xs_kinfu_get_times(const xs_kinfu *k, float *ms10);
/* ICP log of the last frame: iterations x (1+ncomp) x 42 doubles (A 36 column-major, b 6).  The Gauss-Newton loop
 * runs on the device without host round trips; the per-iteration normal equations are only downloaded when the log
 * has been enabled (off by default). */
int xs_kinfu_enable_icp_log(xs_kinfu *k, int on);
int xs_kinfu_take_icp_log(xs_kinfu *k, double *out, int max_iters);
/* updated-voxel count of the last integration (drives the algorithmic-bytes model) */
int xs_kinfu_get_stats(const xs_kinfu *k, unsigned long long *out4);
/* Algorithmic bytes of the last frame per stage (surface, icp, integrate, raycast+resize), DESIGN.md §5 */
int xs_kinfu_get_algorithmic_bytes(const xs_kinfu *k, double *out4);
/* device pointer where the per-frame derivative record (world2camera, all components) is kept,
 * laid out [(1+ncomp)][16] floats — the buffer the multi-GPU layer all-gathers. */
float *xs_kinfu_pose_record_device(xs_kinfu *k);
/* gt_poses + flag_use_gtPose (KinectFusionReconstruction.h:36,82 / .cpp:69,164-166,239-247): with use_gt_pose != 0 frames are
 * fused at the given camera-to-world poses (row-major 4x4 per frame, real) and ICP is skipped (mapping mode). */
int xs_kinfu_set_gt_poses(xs_kinfu *k, const float *poses16, int n, int use_gt_pose);
/* se3Exp (KinectFusionReconstruction.h:176-219), the parameterisation of the reference's pose-set / relocalisation experiments,
 * on batched jets (host): xi [(1 + ncomp)][6] = (v, omega), component 0 real, then the h-scaled derivative components of the
 * batch kind (comps, dirs, pairs as for xs_kinfu_create / xs_kinfu_create_hessian); T_out [(1 + ncomp)][16] row-major 4x4. */
int xs_se3_exp(const float *xi, int comps, int dirs, int npairs, const int *pairs, float *T_out);
/* the cudaStream_t of this pipeline object: every result a consumer can observe (maps, volume, pose record) is ordered on it
 * (for event timing and stream-ordered consumers).  Work that depends on no derivative component - the real chain of the ICP,
 * the head of a deferred frame - runs on an internal second stream and is joined back by events. */
void *xs_kinfu_stream(xs_kinfu *k);

/* ---------------------------------------------------------------- multi-GPU layer (e)
 * One process per GPU (the reference is single-GPU).  Perturbation directions are independent given the real state, which is
 * deterministic and identical on every rank, so ranks carry shares of the derivative components (Hessian batch: every
 * first-order component + a share of the pairs) and exchange only the per-frame pose records: one NCCL all-gather per frame,
 * queued by xs_kinfu_process_frame itself on an internal stream behind the frame's record upload.  NCCL is bound at run time
 * (libnccl.so.2).  The launcher distributes the 128-byte id of rank 0 (file, MPI, torch.distributed ...). */
typedef struct xs_comm xs_comm;
int xs_set_device(int device);
int xs_comm_unique_id(unsigned char id_out[128]);
xs_comm *xs_comm_create(int rank, int world, const unsigned char id[128]); /* on the current device */
void xs_comm_destroy(xs_comm *c);
int xs_comm_rank(const xs_comm *c);
int xs_comm_world(const xs_comm *c);
int xs_comm_all_gather(xs_comm *c, const float *d_send, float *d_recv, long floats, void *stream);
/* record_floats: floats per rank in the gather, the same on every rank and >= the largest (1 + ncomp) * 16 (records are padded) */
int xs_kinfu_set_comm(xs_kinfu *k, xs_comm *comm, int record_floats);
/* gathered records of the last processed frame, [world][record_floats] (waits for that frame's all-gather) */
int xs_kinfu_get_gathered_records(xs_kinfu *k, float *host_out);
/* lag = 0: as above; lag = 1: the records of the frame before the last processed one (two buffers alternate), whose all-gather
 * ran beside the last frame's kernels - the read of a consumer that runs one frame behind never waits for a collective */
int xs_kinfu_get_gathered_records_lagged(xs_kinfu *k, int lag, float *host_out);
const float *xs_kinfu_gathered_records_device(xs_kinfu *k);

/* ---------------------------------------------------------------- outputs & synthetic input (a13, f1) */
/* saveTxtMatrix, IOHelper.cpp:21-32 */
int xs_save_pose_txt(const char *path, const float *m16);
/* CPointCloud::exportPly, Visualization/src/CPointCloud.cpp:42-67 */
int xs_export_ply(const char *path, const float *points_xyz, const float *normals_xyz, long n);
/* Synthetic ICL-NUIM-shaped depth: analytic-SDF room sphere-traced to planar depth,
 * uint16 mm, invalid outside [200,5000] -> 0.  c2w: row-major 4x4 real.  Host-side (multi-threaded). */
int xs_synth_depth(const float *c2w16, xs_intr intr, int rows, int cols, uint16_t *out_host);
/* closed-form smooth trajectory, frame 0 = identity; <=1.5 cm and <=0.4 deg per frame */
int xs_synth_pose(int frame, float *c2w16_out);


/* ---------------------------------------------------------------- dataset readers (f2: the step before the path) */
/* Dataset / ICL_Dataset / seven_scenes_Dataset, Dataset.h:18-81 / Dataset.cpp:3-124 - host code, no OpenCV. */
typedef struct xs_dataset xs_dataset;
/* cv::imread(path, IMREAD_UNCHANGED) for non-interlaced 8/16-bit greyscale PNG (Dataset.cpp:7).  out_host may be NULL to
 * query the size. */
int xs_read_png16(const char *path, uint16_t *out_host, long capacity, int *rows, int *cols);
/* loadTxtMatrix, IOHelper.cpp:4-19 (row-major out[rows*cols]) */
int xs_load_txt_matrix(const char *path, int rows, int cols, float *out);
/* ICL_Dataset::readPoseFile, Dataset.cpp:90-124: lines [start, end) -> 3x4 top of a row-major 4x4, last row 0 0 0 1 */
int xs_icl_read_pose_file(const char *poses_path, int start, int end, float *pose16);
/* ICL_Dataset(dataset_dir, start_frame, end_frame, is_flip), Dataset.cpp:69-88: `depth/<i>.png` (raw / 5 = mm) and
 * `livingRoom1n.gt.sim`; frames start_frame..end_frame inclusive.  NULL on error (xs_last_error). */
xs_dataset *xs_dataset_open_icl(const char *dataset_dir, int start_frame, int end_frame, int is_flip);
/* seven_scenes_Dataset(dataset_dir, start_frames, end_frames, seq_names, is_flip), Dataset.cpp:13-39:
 * `<seq_name>frame-%06d.depth.png` / `.pose.txt`, seq_name as readInfo returns it ("seq-01/"). */
xs_dataset *xs_dataset_open_seven_scenes(const char *dataset_dir, const int *start_frames, const int *end_frames,
                                         const char *const *seq_names, int nseq, int is_flip);
/* seven_scenes_Dataset::readInfo, Dataset.cpp:41-67.  seq_names_out: max_seq x 16 chars.  Returns the sequence count. */
int xs_seven_scenes_read_info(const char *filename, int *start_frames, int *end_frames, char *seq_names_out, int max_seq);
int xs_dataset_size(const xs_dataset *d);
/* Dataset::getDepthData, Dataset.cpp:3-11: decode, divide by the dataset factor (rounded like cv::Mat /=), optional
 * horizontal flip; out_host: rows x cols uint16 millimetres, ready for xs_kinfu_process_frame. */
int xs_dataset_get_depth(const xs_dataset *d, int index, uint16_t *out_host, int rows, int cols);
int xs_dataset_get_pose(const xs_dataset *d, int index, float *pose16);       /* Dataset::getPose, row-major 4x4 */
int xs_dataset_set_pose(xs_dataset *d, int index, const float *pose16);       /* Dataset::setPose */
const char *xs_dataset_timestamp(const xs_dataset *d, int index);            /* Dataset::getTimestamp */
const char *xs_dataset_depth_filename(const xs_dataset *d, int index);
void xs_dataset_close(xs_dataset *d);

#ifdef __cplusplus
}
#endif
#endif /* XSLAM_B200_H */
