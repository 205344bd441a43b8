// tc.cuh — sm_100a primitives used by the tensor-core candidate kernels: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma.kind::tf32 / commit / ld) and the UMMA descriptors.  Inline PTX only, no library code.
//
// Descriptor encodings follow the PTX ISA "tcgen05 matrix descriptor" / "instruction descriptor" tables:
//   shared-memory matrix descriptor (64 bit): [0,14) start address >> 4, [16,30) leading byte offset >> 4,
//     [32,46) stride byte offset >> 4, [46,48) version = 1, [49,52) base offset, [61,64) layout (2 = SWIZZLE_128B)
//   K-major operand tile with 128-byte swizzle: rows are 128 B (32 fp32) apart, 8-row groups 1024 B apart
//     => LBO = 1 (unused for swizzled K-major), SBO = 1024 >> 4 = 64; the tile base must be 1024-byte aligned;
//     stepping K by one MMA (8 tf32 = 32 B) adds 32 >> 4 = 2 to the start-address field.
//   instruction descriptor (32 bit, kind::tf32): [4,6) D format = 1 (F32), [7,10) A format = 2 (TF32),
//     [10,13) B format = 2 (TF32), bit 15/16 A/B major = 0 (K-major), [17,23) N >> 3, [24,29) M >> 4
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace vers {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2D tile load: c0 = innermost coordinate (elements), c1 = row
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// ---------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // the same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_thread_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_thread_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__host__ __device__ constexpr uint64_t smem_desc_k_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t idesc_tf32(uint32_t M, uint32_t N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// kind::f16 with fp16 A and B (format code 0; bf16 would be 1), fp32 accumulator; one MMA covers K = 16
__host__ __device__ constexpr uint32_t idesc_f16(uint32_t M, uint32_t N) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, one elected thread issues
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
// fp16 operands (K = 16 per instruction), both from shared memory
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
// same with the A operand taken from tensor memory (lane = row, one 32-bit column per K element)
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
// fp16 operands with A from tensor memory: lane = row, two K elements per 32-bit column (even K in the low half),
// 8 columns per K = 16 step
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
// arrives on the mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// warp w may read TMEM lanes 32*(w%4) .. +31; thread t gets lane base+t, 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// one fp32 column per thread (lane base+t, column of taddr); issue only — pair with tmem_ld_wait()
__device__ __forceinline__ float tmem_ld_1_nowait(uint32_t taddr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
    return __uint_as_float(r);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// wait that carries the loaded registers as in/out operands, so no use of them can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait3(float& a, float& b, float& c) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(a), "+f"(b), "+f"(c)::"memory");
}
// 8 consecutive fp32 columns per thread, issue only — pair with tmem_ld_wait_16() over both halves
__device__ __forceinline__ void tmem_ld_8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait_16(uint32_t (&a)[8], uint32_t (&b)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
                   "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7])::"memory");
}
__device__ __forceinline__ void tmem_ld_wait_8(uint32_t (&a)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7])::"memory");
}
__device__ __forceinline__ void tmem_ld_wait_24(uint32_t (&a)[8], uint32_t (&b)[8], uint32_t (&c)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
                   "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]),
                   "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7])::"memory");
}
// 32 consecutive fp32 columns per thread, issue only — pair with tmem_ld_wait_32()
__device__ __forceinline__ void tmem_ld_32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait_32(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
                   "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                   "+r"(r[30]), "+r"(r[31])::"memory");
}
// plain (non-tensor) bulk copy global -> shared, completion on an mbarrier; 16-byte aligned, size multiple of 16
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// thread t of warp w writes 32 consecutive 32-bit columns of TMEM lane 32*(w%4)+t
__device__ __forceinline__ void tmem_st_32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// thread t of warp w writes 16 consecutive 32-bit columns of TMEM lane 32*(w%4)+t
__device__ __forceinline__ void tmem_st_16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

}  // namespace tc

__device__ __forceinline__ uint32_t round_tf32_bits(uint32_t u) {  // round to nearest even at bit 13
    return (u + 0xFFFu + ((u >> 13) & 1u)) & 0xFFFFE000u;
}

// A row table [n][ld] fp32 rounded to nearest tf32 (the tensor core would truncate; rounding halves the error term),
// written as the SHARED-MEMORY IMAGE of 64-row B tiles: per tile and K chunk one [64 rows x 32 floats] box in the
// 128-byte swizzle the UMMA descriptor expects (16-byte chunk c of row r at chunk position c ^ (r & 7)), the boxes of
// a tile back to back, rows past n and columns past ld zero.  A tile is then ONE linear cp.async.bulk of nk x 8 KB
// instead of nk TMA boxes of 64 separate 128-byte rows: the TMA row-request rate (~1 per 6-8 clk per SM), not L2 or HBM
// bandwidth, is what bounds a kernel that streams such boxes (measured: 52 % tensor-pipe activity in tc_assign1_kernel,
// 11 % in tc_flat_kernel with 2D TMA loads).
static __global__ void tile_image_tf32_kernel(const float* __restrict__ in, uint64_t n, uint32_t ld, uint32_t nk,
                                              float* __restrict__ out) {
    const uint64_t n4 = ((n + 63) / 64) * nk * 64 * 8;  // float4s of the image
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t c4 = (uint32_t)(i & 7u);        // chunk position inside the 128-byte row of the image
        const uint32_t r = (uint32_t)((i >> 3) & 63u);  // row of the box
        const uint64_t box = i >> 9;                   // (tile, K chunk)
        const uint32_t kc = (uint32_t)(box % nk);
        const uint64_t tile = box / nk;
        const uint32_t src_c4 = c4 ^ (r & 7u);         // the logical chunk stored at this position
        const uint64_t row = tile * 64 + r;
        const uint32_t col = kc * 32 + src_c4 * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < n && col < ld) v = *reinterpret_cast<const float4*>(in + row * ld + col);  // ld % 4 == 0
        v.x = __uint_as_float(round_tf32_bits(__float_as_uint(v.x)));
        v.y = __uint_as_float(round_tf32_bits(__float_as_uint(v.y)));
        v.z = __uint_as_float(round_tf32_bits(__float_as_uint(v.z)));
        v.w = __uint_as_float(round_tf32_bits(__float_as_uint(v.w)));
        reinterpret_cast<float4*>(out)[i] = v;
    }
}

// The same image with fp16 elements (kind::f16 has tf32's 11-bit significand at twice the MMA rate and half the
// operand bytes): per 64-row tile and K chunk of 64 halfs one [64 rows x 128 bytes] box in the 128-byte swizzle, values
// fp16(x * scale) (scale = a power of two chosen by the host so that nothing overflows; `bad` counts elements that did).
static __global__ void tile_image_f16_kernel(const float* __restrict__ in, uint64_t n, uint32_t ld, uint32_t nk16,
                                             float scale, uint4* __restrict__ out, uint32_t* __restrict__ bad) {
    const uint64_t n16 = ((n + 63) / 64) * nk16 * 64 * 8;  // 16-byte chunks of the image
    uint32_t nbad = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t c8 = (uint32_t)(i & 7u);         // chunk position inside the 128-byte row of the image
        const uint32_t r = (uint32_t)((i >> 3) & 63u);  // row of the box
        const uint64_t box = i >> 9;                    // (tile, K chunk)
        const uint32_t kc = (uint32_t)(box % nk16);
        const uint64_t tile = box / nk16;
        const uint32_t src_c8 = c8 ^ (r & 7u);          // the logical chunk stored at this position
        const uint64_t row = tile * 64 + r;
        const uint32_t col = kc * 64 + src_c8 * 8;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (row < n && col < ld) {  // ld % 4 == 0
            const float4 a = *reinterpret_cast<const float4*>(in + row * ld + col);
            v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w;
            if (col + 4 < ld) {
                const float4 b = *reinterpret_cast<const float4*>(in + row * ld + col + 4);
                v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
            }
        }
        __align__(16) __half h[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            h[e] = __float2half_rn(v[e] * scale);
            if (!(fabsf(__half2float(h[e])) <= 65504.0f)) ++nbad;  // inf or nan
        }
        out[i] = *reinterpret_cast<const uint4*>(h);
    }
    if (nbad) atomicAdd(bad, nbad);
}

}  // namespace vers
