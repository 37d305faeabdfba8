// tests/fake_cuda/fake_cudart.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The handful of CUDA runtime entry points libhvb's host side uses, implemented on host memory, so that the library's
// host code (context, pictures, pools, staging, the batch entry points) can be built with g++ and exercised in the
// CPU-only suite together with the emulated kernels (tests/host_build.py).  "Device" memory is malloc'ed and 256-byte
// aligned like cudaMalloc's; every copy is synchronous; streams and events are opaque tokens; one device that reports
// compute capability 10.0 with 148 SMs.  Signatures come from the real cuda_runtime_api.h, so a mismatch does not compile.
#include <cuda_runtime_api.h>
#include <cstdlib>
#include <cstring>

extern "C" {

cudaError_t cudaGetDeviceCount(int *count) { *count = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp *prop, int)
{
    memset(prop, 0, sizeof(*prop));
    strcpy(prop->name, "emulated B200");
    prop->major = 10;
    prop->minor = 0;
    prop->multiProcessorCount = 148;
    return cudaSuccess;
}
cudaError_t cudaDeviceGetAttribute(int *value, cudaDeviceAttr attr, int)
{
    *value = attr == cudaDevAttrMultiProcessorCount ? 148 : attr == cudaDevAttrComputeCapabilityMajor ? 10 : 0;
    return cudaSuccess;
}
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaSetDeviceFlags(unsigned int) { return cudaSuccess; }
// kernel attributes and occupancy: the launch code only sizes grids with them
cudaError_t cudaFuncSetAttribute(const void *, cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *blocks, const void *, int, size_t) { *blocks = 4; return cudaSuccess; }
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(int *blocks, const void *, int, size_t, unsigned) { *blocks = 4; return cudaSuccess; }
// a __device__ variable is an ordinary variable here, and its "symbol" is its address
cudaError_t cudaMemcpyToSymbol(const void *symbol, const void *src, size_t bytes, size_t offset, cudaMemcpyKind)
{
    memcpy(static_cast<char *>(const_cast<void *>(symbol)) + offset, src, bytes);
    return cudaSuccess;
}
const char *cudaGetErrorString(cudaError_t) { return "emulated runtime: no error text"; }
const char *cudaGetErrorName(cudaError_t) { return "cudaEmulated"; }

cudaError_t cudaMalloc(void **p, size_t bytes) { return posix_memalign(p, 256, bytes ? bytes : 1) ? cudaErrorMemoryAllocation : cudaSuccess; }
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void **p, size_t bytes) { return posix_memalign(p, 256, bytes ? bytes : 1) ? cudaErrorMemoryAllocation : cudaSuccess; }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaHostAlloc(void **p, size_t bytes, unsigned) { return posix_memalign(p, 256, bytes ? bytes : 1) ? cudaErrorMemoryAllocation : cudaSuccess; }
cudaError_t cudaHostGetDevicePointer(void **dev, void *host, unsigned) { *dev = host; return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *attr, const void *)
{
    memset(attr, 0, sizeof(*attr));
    attr->type = cudaMemoryTypeUnregistered; // every caller buffer is pageable host memory here
    return cudaSuccess;
}

cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind) { memmove(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind, cudaStream_t) { memmove(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaMemcpy2DAsync(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t)
{
    if (width > dpitch || width > spitch) return cudaErrorInvalidPitchValue;
    for (size_t y = 0; y < height; ++y) memmove(static_cast<char *>(dst) + y * dpitch, static_cast<const char *>(src) + y * spitch, width);
    return cudaSuccess;
}
cudaError_t cudaMemset(void *p, int value, size_t bytes) { memset(p, value, bytes); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *p, int value, size_t bytes, cudaStream_t) { memset(p, value, bytes); return cudaSuccess; }

cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = reinterpret_cast<cudaStream_t>(malloc(1)); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = reinterpret_cast<cudaEvent_t>(malloc(1)); return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = reinterpret_cast<cudaEvent_t>(malloc(1)); return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }

} // extern "C"
