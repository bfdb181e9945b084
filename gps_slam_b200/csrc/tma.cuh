// 1-D TMA bulk copy (cp.async.bulk, UBLKCP in SASS) + mbarrier helpers for sm_100a: global -> shared with byte-count completion.
#pragma once
#include <cuda_runtime.h>

namespace tma
{
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); // make the init visible to the async proxy
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT_LOOP;\nDONE:\n}" ::"r"(
                     smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}
// bytes: multiple of 16; both addresses 16-byte aligned
__device__ __forceinline__ void load_1d(void *dstSmem, const void *srcGmem, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dstSmem)),
                 "l"(srcGmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy (bytes: multiple of 16, 16-byte aligned); generic-proxy writes to the source need fence_proxy_async() first
__device__ __forceinline__ void store_1d(void *dstGmem, const void *srcSmem, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dstGmem), "r"(smem_u32(srcSmem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
} // namespace tma
