// Shared helpers for the e2e_asr_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace e2e {

void set_error(const char* fmt, ...);

#define E2E_CHECK_CUDA(expr)                                                            \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            e2e::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr,                 \
                           cudaGetErrorString(_e));                                     \
            return 1;                                                                   \
        }                                                                               \
    } while (0)

#define E2E_REQUIRE(cond, ...)                                                          \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            e2e::set_error(__VA_ARGS__);                                                \
            return 2;                                                                   \
        }                                                                               \
    } while (0)

// every kernel launch of this library goes through here: error check + launch counter
extern unsigned long long g_launches;
#define E2E_LAUNCH_CHECK()                 \
    do {                                   \
        ++e2e::g_launches;                 \
        E2E_CHECK_CUDA(cudaGetLastError()); \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

int sm_count();

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// acquire / release helpers for the per-group step barriers of the persistent kernels
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Spin until *ctr >= target; gives up after ~2^31 cycles (sets *err) so a lost
// arrival can never hang the GPU.  Called by ONE thread; follow with __syncthreads().
__device__ __forceinline__ bool spin_wait_ge(const unsigned* ctr, unsigned target, int* err) {
    long long t0 = clock64();
    while (ld_acquire_gpu(ctr) < target) {
        if (clock64() - t0 > (1ll << 31)) {
            atomicExch(err, 1);
            return false;
        }
        __nanosleep(20);
    }
    return true;
}

}  // namespace e2e
