// Library identification and error decoding for libglue_b200 (sm_100a).
#include <atomic>

#include "common.cuh"

static std::atomic<unsigned long long> g_launches{0};
extern "C" void glue_count_launch_(void) { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" unsigned long long glue_kernel_launches(void) { return g_launches.load(); }

extern "C" int glue_abi_version(void) { return GLUE_B200_ABI_VERSION; }

extern "C" const char* glue_arch(void) { return "sm_100a"; }

extern "C" const char* glue_error_string(int code) {
    switch (code) {
        case 0: return "success";
        case GLUE_E_BADARG: return "glue_b200: bad argument (NULL pointer or non-positive size)";
        case GLUE_E_UNSUPPORTED: return "glue_b200: unsupported configuration";
        case GLUE_E_WORKSPACE: return "glue_b200: workspace too small";
        case GLUE_E_ALIGN: return "glue_b200: misaligned pointer";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "glue_b200: unknown error";
}
