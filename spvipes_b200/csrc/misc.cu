// ABI version / architecture probe.
#include "common.cuh"
#include "../../include/spvipes_b200.h"

extern "C" int spv_abi_version(void) { return SPV_ABI_VERSION; }

extern "C" int spv_arch_check(int dev) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return SPV_ERR_LAUNCH;
    return (prop.major == 10 && prop.minor == 0) ? SPV_OK : SPV_ERR_ARCH;
}
