// 1-D bulk asynchronous copies (TMA engine, SASS UBLKCP) and the mbarrier calls they need,
// as inline PTX for sm_100a.  Used to stage whole tiles of move bytes into shared memory and
// whole tiles of sticker rows back to HBM without spending LSU wavefronts or ALU address
// arithmetic on the copies.  Sizes and both addresses must be multiples of 16 bytes.
#pragma once
#include <cstdint>

namespace bulk {

__device__ __forceinline__ uint32_t smem_addr(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");      // visible to the async proxy
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return done != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}

// one lane of the (converged) warp, chosen by the hardware: code under it is known to run in a single thread, which
// spares the copy instructions' uniform operands the compiler's "which active lane?" loop
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// global -> shared, completion counted on `bar`
__device__ __forceinline__ void load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

// hint: bring [src, src + bytes) into L2.  No architectural effect, so it may be issued BEFORE a dependency wait
// (griddepcontrol.wait): the data is only read after the wait, from wherever its latest copy is
__device__ __forceinline__ void prefetch_l2(const void* src_gmem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}

// global -> shared through a 2-D tensor map (box at element coordinates {c0, c1}; the map fixes the box,
// the swizzle and the byte count), completion counted on `bar`
__device__ __forceinline__ void load_tile_2d(void* dst_smem, const void* tensor_map, int c0, int c1, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_addr(dst_smem)), "l"(tensor_map), "r"(c0), "r"(c1), "r"(smem_addr(bar))
                 : "memory");
}

// shared -> global, tracked by the thread's bulk group
__device__ __forceinline__ void store(void* dst_gmem, const void* src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst_gmem), "r"(smem_addr(src_smem)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }

// all committed stores of this thread have finished READING shared memory
__device__ __forceinline__ void wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// order this thread's generic-proxy shared-memory writes before later async-proxy reads
__device__ __forceinline__ void fence_smem_writes() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace bulk
