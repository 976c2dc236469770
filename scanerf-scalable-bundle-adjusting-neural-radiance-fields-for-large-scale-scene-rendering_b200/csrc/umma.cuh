// Blackwell (sm_100a) tensor-core plumbing used by the decoder kernels: tcgen05.mma with
// accumulators in TMEM, shared-memory operand descriptors, mbarrier completion, TMEM
// allocation and tcgen05.ld.  Inline PTX only -- no library code.
//
// Operand convention used throughout (all operand tiles are "rows of 128 bytes"):
//   a tile is R rows x 64 bf16 columns, row r at byte offset r*128, and the eight 16-byte
//   chunks of a row are stored XOR-swizzled: chunk c lives at position c ^ (r & 7)
//   (the 128-byte swizzle of the UMMA shared-memory descriptor; tile base 1024-byte aligned).
//   The same memory image can be read by the tensor core either
//     K-major  : rows index M (or N), columns are the contraction dimension   (Y = X W^T), or
//     MN-major : columns index M (or N), rows are the contraction dimension   (dX = G W, dW = G^T X)
//   so neither activations nor weights ever need a transposed copy.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

// ---------------------------------------------------------------- fences
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- TMEM allocation (one full warp)
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem)
{
    static_assert(COLS == 32 || COLS == 64 || COLS == 128 || COLS == 256 || COLS == 512, "power of two >= 32");
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "n"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_free(uint32_t base)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(COLS) : "memory");
}

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor, 128-byte swizzle, rows of 128 bytes, 8-row groups 1024 B apart.
// (start address >> 4 in bits [0,14), LBO >> 4 in [16,30), SBO >> 4 in [32,46), version 1 in
//  [46,48), layout type SWIZZLE_128B = 2 in [61,64).)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                 // leading byte offset: unused for a single 64-column atom
    d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset: next 8-row group
    d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}
// The start-address field sits in the low 14 bits and shared memory ends below 256 KB, so stepping an operand is a plain
// addition to the descriptor of the tile base -- which the compiler computes once per tile (the ~10 shift / mask / or
// instructions per descriptor in front of every tcgen05.mma were ~30 % of the one issuing thread's time in the backward).
// K-major operand: one MMA consumes 16 columns = 32 bytes of every row -> advance inside the row
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile_addr, int kstep) { return desc_sw128(tile_addr) + (uint64_t)(uint32_t)(kstep * 2); }
// MN-major operand: one MMA consumes 16 rows = 2048 bytes -> advance by whole row groups
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_addr, int kstep) { return desc_sw128(tile_addr) + (uint64_t)(uint32_t)(kstep * 128); }

// Instruction descriptor for kind::f16, bf16 x bf16 -> f32 (c_format F32 = 1 at [4,6), a/b format
// BF16 = 1 at [7,10)/[10,13), a_major at 15, b_major at 16 (0 = K-major, 1 = MN-major),
// N >> 3 at [17,23), M >> 4 at [24,29)).
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn_major, int b_mn_major)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// The same with the operand formats chosen independently (a_bf16 / b_bf16: 1 = BF16, 0 = F16): kind::f16 takes
// an fp16 operand against a bf16 one.
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int a_mn_major, int b_mn_major, int a_bf16, int b_bf16)
{
    return (1u << 4) | ((uint32_t)a_bf16 << 7) | ((uint32_t)b_bf16 << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------- MMA issue / completion (one thread)
// One elected lane of a fully converged warp (the same lane every time for a full member mask).  Issuing the MMAs under
// `if (warp-uniform condition && elect_one())` instead of `if (tid == leader)` lets ptxas keep descriptors and issue on the
// uniform datapath without wrapping every tcgen05.mma in an elect / branch loop.
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "elect.sync _|P1, 0xFFFFFFFF;\n"
        "selp.b32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
// warp index as a warp-uniform value
__device__ __forceinline__ int warp_uniform() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on `bar` once every MMA issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- TMEM -> registers
// 32 consecutive 32-bit columns of this thread's TMEM lane (lane = 32*(warp%4) + laneid).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v)
{
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v)
{
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v)
{
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
// TMEM address of (lane, column) relative to an allocation base
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, int lane, int col) { return base + ((uint32_t)lane << 16) + (uint32_t)col; }

// ---------------------------------------------------------------- operand tiles in shared memory
// byte offset of 16-byte chunk `chunk` (8 bf16 columns) of row `row` in a swizzled tile
__device__ __forceinline__ uint32_t tile_chunk_off(int row, int chunk) { return (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b)
{
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// (a -> low half, b -> high half); saturating: a value beyond +-65504 becomes +-65504, never inf
__device__ __forceinline__ uint32_t pack_f16(float a, float b)
{
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
// store 8 consecutive columns [8*chunk, 8*chunk+8) of row `row` as bf16
__device__ __forceinline__ void tile_store8(unsigned char* tile, int row, int chunk, const float* v)
{
    uint4 q;
    q.x = pack_bf16(v[0], v[1]); q.y = pack_bf16(v[2], v[3]); q.z = pack_bf16(v[4], v[5]); q.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(tile + tile_chunk_off(row, chunk)) = q;
}
// same slot, but 8 fp16 values
__device__ __forceinline__ void tile_store8_f16(unsigned char* tile, int row, int chunk, const float* v)
{
    uint4 q;
    q.x = pack_f16(v[0], v[1]); q.y = pack_f16(v[2], v[3]); q.z = pack_f16(v[4], v[5]); q.w = pack_f16(v[6], v[7]);
    *reinterpret_cast<uint4*>(tile + tile_chunk_off(row, chunk)) = q;
}
__device__ __forceinline__ void tile_zero8(unsigned char* tile, int row, int chunk)
{
    *reinterpret_cast<uint4*>(tile + tile_chunk_off(row, chunk)) = make_uint4(0u, 0u, 0u, 0u);
}

}  // namespace umma
