// ABI version / architecture probe.
#include "common.cuh"
#include "../../include/spvipes_b200.h"

unsigned long long g_spv_launches = 0;

// number of kernels this library has launched (or captured into a CUDA graph) so far in this process
extern "C" long long spv_launch_count(void) { return (long long)g_spv_launches; }

extern "C" int spv_abi_version(void) { return SPV_ABI_VERSION; }

extern "C" int spv_arch_check(int dev) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return SPV_ERR_LAUNCH;
    return (prop.major == 10 && prop.minor == 0) ? SPV_OK : SPV_ERR_ARCH;
}

// host -> device copy of a column block of a row-major host matrix (pinned: asynchronous) into a dense device buffer: the
// scvi minibatch layout is [B, G0 + G1] per group while a group's kernels read only its own G_g columns
// (module/spVIPESmodule.py:428-430), so the plugin call moves just those.  width / pitches in bytes.
extern "C" int spv_copy2d_h2d(void* dst, long long dpitch, const void* src, long long spitch, long long width, long long rows,
                              void* stream) {
    if (!dst || !src || width <= 0 || rows <= 0 || dpitch < width || spitch < width) return SPV_ERR_ARG;
    cudaError_t e = cudaMemcpy2DAsync(dst, (size_t)dpitch, src, (size_t)spitch, (size_t)width, (size_t)rows, cudaMemcpyHostToDevice,
                                      reinterpret_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? SPV_OK : SPV_ERR_LAUNCH;
}
