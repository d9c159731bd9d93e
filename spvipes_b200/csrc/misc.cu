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
