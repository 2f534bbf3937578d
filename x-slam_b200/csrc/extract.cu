// extract.cu — point-cloud extraction from the brick-tiled TSDF (real-only output path).
//
// Replaces extractKernel / extractPoints (XKinectFusion/src/ExtractPointCloud.cu:25-210) and
// extractNormalsKernel / extractNormals (ExtractPointCloud.cu:213-362): zero crossings along the +x, +y, +z
// voxel edges (both values < 0.99, opposite signs), point at the linear zero; normals from trilinear central
// differences at +-1 voxel, divided by the SQUARED norm exactly as the reference does (:305-306).
// One CTA scans one brick (x fastest, coalesced) and appends with a warp-aggregated atomic, so the order of
// points is not the reference's (which is itself scheduling dependent); parity is on the point set.
#include "xs_common.cuh"

namespace xs {

XS_DEV float fetch_value(const VolumeView &V, int x, int y, int z) { return __ldg(V.value + value_index(V, x, y, z)); }

__global__ void __launch_bounds__(512) extract_points_kernel(VolumeView V, float *__restrict__ out, long max_points,
                                                             unsigned long long *counter) {
    const int tid = threadIdx.x;
    const int nbricks = V.bx * V.by * V.bz;
    const float vs = V.voxel;
    for (int b = blockIdx.x; b < nbricks; b += gridDim.x) {
        const int bx = b % V.bx, by = (b / V.bx) % V.by, bz = b / (V.bx * V.by);
        const int x = bx * 8 + (tid & 7), y = by * 8 + ((tid >> 3) & 7), z = bz * 8 + (tid >> 6);
        float px[3], py[3], pz[3];
        int n = 0;
        // ExtractPointCloud.cu:61-66: z < res.z-1, x < res.x-1, y < res.y-1
        if (z < V.rz - 1 && x < V.rx - 1 && y < V.ry - 1) {
            const float F = V.value[(size_t) b * BRICK_VOX + tid];
            if (F < 0.99f) {
                const float Vx = (x + 0.5f) * vs, Vy = (y + 0.5f) * vs, Vz = (z + 0.5f) * vs;
                float Fn = fetch_value(V, x + 1, y, z);
                if (Fn < 0.99f && ((F > 0 && Fn < 0) || (F < 0 && Fn > 0))) {
                    px[n] = Vx - (F / (Fn - F)) * vs;
                    py[n] = Vy;
                    pz[n] = Vz;
                    ++n;
                }
                Fn = fetch_value(V, x, y + 1, z);
                if (Fn < 0.99f && ((F > 0 && Fn < 0) || (F < 0 && Fn > 0))) {
                    px[n] = Vx;
                    py[n] = Vy - (F / (Fn - F)) * vs;
                    pz[n] = Vz;
                    ++n;
                }
                Fn = fetch_value(V, x, y, z + 1);
                if (Fn < 0.99f && ((F > 0 && Fn < 0) || (F < 0 && Fn > 0))) {
                    px[n] = Vx;
                    py[n] = Vy;
                    pz[n] = Vz - (F / (Fn - F)) * vs;
                    ++n;
                }
            }
        }
        // warp-aggregated append
        int incl = n;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((tid & 31) >= o) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        unsigned long long base = 0;
        if ((tid & 31) == 0) base = atomicAdd(counter, (unsigned long long) total);
        base = __shfl_sync(0xffffffffu, base, 0);
        const unsigned long long at = base + (incl - n);
        for (int l = 0; l < n; ++l)
            if ((long) (at + l) < max_points) {
                float *p = out + 3 * (at + l);
                p[0] = px[l];
                p[1] = py[l];
                p[2] = pz[l];
            }
    }
}

// ExtractPointCloud.cu:311-337 (plain float arithmetic, no +1e-5 bias)
XS_DEV float trilinear_real(const VolumeView &V, float x, float y, float z) {
    const float vs = V.voxel;
    int gx = __float2int_rd(x / vs), gy = __float2int_rd(y / vs), gz = __float2int_rd(z / vs);
    const float vx = (gx + 0.5f) * vs, vy = (gy + 0.5f) * vs, vz = (gz + 0.5f) * vs;
    gx = (x < vx) ? (gx - 1) : gx;
    gy = (y < vy) ? (gy - 1) : gy;
    gz = (z < vz) ? (gz - 1) : gz;
    const float a = (x - (gx + 0.5f) * vs) / vs;
    const float b = (y - (gy + 0.5f) * vs) / vs;
    const float c = (z - (gz + 0.5f) * vs) / vs;
    return fetch_value(V, gx + 0, gy + 0, gz + 0) * (1 - a) * (1 - b) * (1 - c) +
           fetch_value(V, gx + 0, gy + 0, gz + 1) * (1 - a) * (1 - b) * c +
           fetch_value(V, gx + 0, gy + 1, gz + 0) * (1 - a) * b * (1 - c) +
           fetch_value(V, gx + 0, gy + 1, gz + 1) * (1 - a) * b * c +
           fetch_value(V, gx + 1, gy + 0, gz + 0) * a * (1 - b) * (1 - c) +
           fetch_value(V, gx + 1, gy + 0, gz + 1) * a * (1 - b) * c +
           fetch_value(V, gx + 1, gy + 1, gz + 0) * a * b * (1 - c) +
           fetch_value(V, gx + 1, gy + 1, gz + 1) * a * b * c;
}

__global__ void extract_normals_kernel(VolumeView V, const float *__restrict__ points, float *__restrict__ normals, long n) {
    const long idx = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const float vs = V.voxel;
    const float x = points[3 * idx], y = points[3 * idx + 1], z = points[3 * idx + 2];
    float nx = 0.f, ny = 0.f, nz = 0.f;
    const int gx = __float2int_rd(x / vs), gy = __float2int_rd(y / vs), gz = __float2int_rd(z / vs);
    if (gx > 1 && gy > 1 && gz > 1 && gx < V.rx - 2 && gy < V.ry - 2 && gz < V.rz - 2) {
        nx = trilinear_real(V, x + vs, y, z) - trilinear_real(V, x - vs, y, z);
        ny = trilinear_real(V, x, y + vs, z) - trilinear_real(V, x, y - vs, z);
        nz = trilinear_real(V, x, y, z + vs) - trilinear_real(V, x, y, z - vs);
        const float norm = nx * nx + ny * ny + nz * nz;  // squared norm: reference quirk, :305-306
        nx = nx / norm;
        ny = ny / norm;
        nz = nz / norm;
    }
    normals[3 * idx] = nx;
    normals[3 * idx + 1] = ny;
    normals[3 * idx + 2] = nz;
}

}  // namespace xs

using namespace xs;

extern "C" long xs_extract_points(const xs_volume *v, float *d_points_xyz, float *d_normals_xyz, long max_points,
                                  void *stream) {
    if (!v || !d_points_xyz || max_points <= 0) return XS_ERR_ARG;
    cudaStream_t s = (cudaStream_t) stream;
    // The point counter is a slot of its own ([3]; [0..2] are the integration's statistics, which a deferred frame may not
    // have collected yet).  A frame that is still integrating on another (non-blocking) stream must be complete before the
    // volume is scanned: the extraction is an end-of-sequence call, so a device-wide wait is the simple, safe order.
    if (cudaDeviceSynchronize() != cudaSuccess) return XS_ERR_CUDA;
    unsigned long long *d_count = v->d_stats + 3, *h_count = v->h_stats + 3;
    if (cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), s) != cudaSuccess) return XS_ERR_CUDA;
    const int nbricks = v->view.bx * v->view.by * v->view.bz;
    const int max_grid = sm_count() * 8;
    extract_points_kernel<<<nbricks < max_grid ? nbricks : max_grid, 512, 0, s>>>(v->view, d_points_xyz, max_points, d_count);
    ++g_launches;
    if (cudaGetLastError() != cudaSuccess) return XS_ERR_CUDA;
    if (cudaMemcpyAsync(h_count, d_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s) != cudaSuccess)
        return XS_ERR_CUDA;
    if (cudaStreamSynchronize(s) != cudaSuccess) return XS_ERR_CUDA;
    long n = (long) *h_count;
    if (n > max_points) n = max_points;  // output_count = min(size, global_count), :175
    if (d_normals_xyz && n > 0) {
        const int rc = xs_extract_normals(v, d_points_xyz, d_normals_xyz, n, stream);
        if (rc != XS_OK) return rc;
    }
    return n;
}

// extractNormals, ExtractPointCloud.h:22-23 / ExtractPointCloud.cu:342-362 (sync)
extern "C" int xs_extract_normals(const xs_volume *v, const float *d_points_xyz, float *d_normals_xyz, long n, void *stream) {
    if (!v || !d_points_xyz || !d_normals_xyz || n < 0) return XS_ERR_ARG;
    if (n == 0) return XS_OK;
    cudaStream_t s = (cudaStream_t) stream;
    extract_normals_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(v->view, d_points_xyz, d_normals_xyz, n);
    ++g_launches;
    if (cudaGetLastError() != cudaSuccess) return XS_ERR_CUDA;
    if (cudaStreamSynchronize(s) != cudaSuccess) return XS_ERR_CUDA;
    return XS_OK;
}
