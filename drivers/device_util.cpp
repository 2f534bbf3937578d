// device_util.cpp — the three CUDA runtime calls the drivers need (they are otherwise pure C-ABI clients).
#include <cuda_runtime_api.h>

#include <cstddef>

int xs_driver_device_alloc(float **p, size_t floats) { return cudaMalloc((void **) p, floats * sizeof(float)) == cudaSuccess ? 0 : -1; }
int xs_driver_device_download(float *dst, const float *src, size_t floats) {
    return cudaMemcpy(dst, src, floats * sizeof(float), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1;
}
int xs_driver_device_upload(float *dst, const float *src, size_t floats) {
    return cudaMemcpy(dst, src, floats * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess ? 0 : -1;
}
void xs_driver_device_free(float *p) { cudaFree(p); }
