// Tensor memory (TMEM, 256 KB per SM: 128 lanes x 512 columns x 32 bits) used as thread-private storage.
//
// With the 32x32b access shape, thread i of a warp reads/writes TMEM lane 32 * (warp % 4) + i, i.e. every thread
// owns one TMEM lane of 512 columns (2 KB) shared with the threads of the same index in the other warps of its
// quadrant. The cooperative kernel keeps there the history rows that only ONE thread ever touches (r^2 and
// r^alpha of the gravitational pair interaction, see fused.cuh): they stop competing for shared memory with
// the rows that the threads of a warp exchange, and more warps fit on an SM. tcgen05.ld / tcgen05.st are
// warp-wide (.sync.aligned): every thread of the warp must execute them, converged.
#ifndef HEYOKA_B200_CSRC_TMEM_CUH
#define HEYOKA_B200_CSRC_TMEM_CUH

#include <cstdint>

#include <cuda_runtime.h>

#include "recurrences.cuh"

namespace heyoka_b200::dev::tm
{

constexpr std::uint32_t n_cols = 512u;

// Allocation of all the columns by ONE warp; the base address is written to *smem_dst.
__device__ __forceinline__ void alloc_all(std::uint32_t *smem_dst)
{
    const std::uint32_t dst = static_cast<std::uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(n_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void dealloc_all(std::uint32_t taddr)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(n_cols) : "memory");
}
__device__ __forceinline__ void fence_before_sync()
{
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after_sync()
{
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- raw 32-bit column loads / stores (NW consecutive columns of this thread's lane) ----
template <int NW>
struct words {
    std::uint32_t w[NW];
};

__device__ __forceinline__ void ld(std::uint32_t taddr, words<2> &r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r.w[0]), "=r"(r.w[1]) : "r"(taddr));
}
__device__ __forceinline__ void ld(std::uint32_t taddr, words<4> &r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3])
                 : "r"(taddr));
}
__device__ __forceinline__ void ld(std::uint32_t taddr, words<8> &r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]),
                   "=r"(r.w[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void ld(std::uint32_t taddr, words<16> &r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
                 "%14, %15}, [%16];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]),
                   "=r"(r.w[7]), "=r"(r.w[8]), "=r"(r.w[9]), "=r"(r.w[10]), "=r"(r.w[11]), "=r"(r.w[12]),
                   "=r"(r.w[13]), "=r"(r.w[14]), "=r"(r.w[15])
                 : "r"(taddr));
}

// Wait for the outstanding loads. The loaded registers are passed through the statement so that no use of
// them can be scheduled before the wait.
__device__ __forceinline__ void wait_ld(words<2> &r)
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r.w[0]), "+r"(r.w[1])::"memory");
}
__device__ __forceinline__ void wait_ld(words<4> &r)
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r.w[0]), "+r"(r.w[1]), "+r"(r.w[2]), "+r"(r.w[3])::"memory");
}
__device__ __forceinline__ void wait_ld(words<8> &r)
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r.w[0]), "+r"(r.w[1]), "+r"(r.w[2]), "+r"(r.w[3]), "+r"(r.w[4]), "+r"(r.w[5]), "+r"(r.w[6]),
                   "+r"(r.w[7])::"memory");
}
__device__ __forceinline__ void wait_ld(words<16> &r)
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r.w[0]), "+r"(r.w[1]), "+r"(r.w[2]), "+r"(r.w[3]), "+r"(r.w[4]), "+r"(r.w[5]), "+r"(r.w[6]),
                   "+r"(r.w[7]), "+r"(r.w[8]), "+r"(r.w[9]), "+r"(r.w[10]), "+r"(r.w[11]), "+r"(r.w[12]),
                   "+r"(r.w[13]), "+r"(r.w[14]), "+r"(r.w[15])::"memory");
}

__device__ __forceinline__ void st(std::uint32_t taddr, const words<2> &r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(r.w[0]), "r"(r.w[1])
                 : "memory");
}
__device__ __forceinline__ void st(std::uint32_t taddr, const words<4> &r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r.w[0]),
                 "r"(r.w[1]), "r"(r.w[2]), "r"(r.w[3])
                 : "memory");
}
__device__ __forceinline__ void wait_st()
{
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ---- rows of vd<N> values: the order-o coefficient (N doubles) lives in columns [o * 2N, (o + 1) * 2N) ----
template <int N>
struct row {
    static constexpr std::uint32_t W = 2u * N; // columns per order
    std::uint32_t addr;                        // TMEM address of order 0 (lane field = the warp's quadrant)

    // C consecutive orders starting at order o: issue the load ...
    template <int C>
    __device__ __forceinline__ void issue(std::uint32_t o, words<2 * N * C> &r) const
    {
        ld(addr + o * W, r);
    }
    // ... and, after wait_ld(r), unpack.
    template <int C>
    __device__ __forceinline__ static void unpack(const words<2 * N * C> &r, vd<N> (&out)[C])
    {
#pragma unroll
        for (int c = 0; c < C; ++c) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                out[c].v[i] = __hiloint2double(static_cast<int>(r.w[(c * N + i) * 2 + 1]),
                                               static_cast<int>(r.w[(c * N + i) * 2]));
            }
        }
    }
    __device__ __forceinline__ vd<N> get(std::uint32_t o) const
    {
        words<2 * N> r;
        ld(addr + o * W, r);
        wait_ld(r);
        vd<N> out[1];
        unpack<1>(r, out);
        return out[0];
    }
    // Store + wait: the value is visible to the loads that follow.
    __device__ __forceinline__ void set(std::uint32_t o, const vd<N> &v) const
    {
        words<2 * N> r;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            r.w[2 * i] = static_cast<std::uint32_t>(__double2loint(v.v[i]));
            r.w[2 * i + 1] = static_cast<std::uint32_t>(__double2hiint(v.v[i]));
        }
        st(addr + o * W, r);
        wait_st();
    }
};

} // namespace heyoka_b200::dev::tm

#endif
