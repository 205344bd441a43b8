// tma_stream_bench.cu — ceiling of the list-scan access pattern: 148 persistent CTAs stream a row-major fp32 matrix
// through a TMA ring with no math.  Patterns: 0 = 128-row x 32-float boxes, k-chunks inner (the scan kernel's order);
// 1 = contiguous 16 KB bulk copies (what a tile-major copy of the lists would allow); 2 = 64-row x 32-float boxes.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../vers_b200/csrc tma_stream_bench.cu -lcuda -o tma_stream_bench
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc.cuh"
using namespace vers;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int S, int STAGE_BYTES>
__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap tmap, const float* base, uint32_t ld,
                                                       uint64_t rows, uint32_t item_rows, int pattern, int box_rows,
                                                       unsigned long long* counter, unsigned long long* sink) {
    extern __shared__ uint8_t raw_smem[];
    const uint32_t raw = tc::smem_u32(raw_smem);
    uint8_t* smem = raw_smem + (((raw + 1023u) & ~1023u) - raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S * STAGE_BYTES);
    uint64_t* empty = full + S;
    __shared__ long long cur_item[2];
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        tc::fence_barrier_init();
    }
    __syncthreads();
    const uint64_t n_items = (rows + item_rows - 1) / item_rows;
    const uint32_t nk = ld / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // static round-robin items so both roles agree without a ring
    if (warp == 0 && lane == 0) {
        uint32_t stage = 0, phase = 0;
        for (uint64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
            uint64_t r0 = it * item_rows, r1 = min(rows, r0 + item_rows);
            for (uint64_t a0 = r0; a0 < r1; a0 += box_rows) {
                for (uint32_t kc = 0; kc < nk; ++kc) {
                    tc::mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* dst = smem + stage * STAGE_BYTES;
                    const uint32_t bytes = box_rows * 128;
                    tc::mbar_arrive_expect_tx(&full[stage], bytes);
                    if (pattern == 1) {
                        const float* src = base + (a0 * ld) + (uint64_t)kc * box_rows * 32;  // contiguous tile
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(tc::smem_u32(dst)), "l"(src), "r"(bytes), "r"(tc::smem_u32(&full[stage])) : "memory");
                    } else {
                        tc::tma_load_2d(dst, &tmap, &full[stage], (int32_t)(kc * 32), (int32_t)a0);
                    }
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        uint32_t stage = 0, phase = 0;
        unsigned long long acc = 0;
        for (uint64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
            uint64_t r0 = it * item_rows, r1 = min(rows, r0 + item_rows);
            for (uint64_t a0 = r0; a0 < r1; a0 += box_rows) {
                for (uint32_t kc = 0; kc < nk; ++kc) {
                    tc::mbar_wait(&full[stage], phase);
                    acc += *reinterpret_cast<volatile uint32_t*>(smem + stage * STAGE_BYTES);
                    tc::mbar_arrive(&empty[stage]);
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
        }
        if (acc == 0x1234567) *sink = acc;
    }
}

static CUtensorMap make_map(const float* base, uint64_t rows, uint32_t ld, uint32_t box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {ld, rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("tensormap encode failed %d\n", (int)r); exit(1); }
    return m;
}

template <int S, int SB>
static void run(const char* name, const float* d, uint64_t rows, uint32_t ld, int pattern, int box_rows, uint32_t item_rows,
                unsigned long long* d_ctr, int grid) {
    CUtensorMap m = make_map(d, rows, ld, box_rows);
    const int smem = 1024 + S * SB + 256;
    CK(cudaFuncSetAttribute(stream_kernel<S, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        stream_kernel<S, SB><<<grid, 64, smem>>>(m, d, ld, rows, item_rows, pattern, box_rows, d_ctr, d_ctr + 1);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    printf("%-44s stages %2d stage_bytes %6d grid %4d: %8.3f ms  %8.1f GB/s\n", name, S, SB, grid, best,
           (double)rows * ld * 4 / best / 1e6);
}

int main() {
    CK(cudaSetDevice(0));
    cuInit(0);
    const uint32_t ld = 768;
    const uint64_t rows = 9000000;  // 27.6 GB, the bench step's distinct rows
    float* d;
    CK(cudaMalloc(&d, rows * ld * 4));
    CK(cudaMemset(d, 0, rows * ld * 4));
    unsigned long long* d_ctr;
    CK(cudaMalloc(&d_ctr, 64));
    CK(cudaMemset(d_ctr, 0, 64));
    for (int grid : {148, 296}) {
        if (grid == 148) {
            run<10, 16384>("box 128x32 strided (scan kernel order)", d, rows, ld, 0, 128, 2432, d_ctr, grid);
            run<6, 16384>("box 128x32 strided (scan kernel order)", d, rows, ld, 0, 128, 2432, d_ctr, grid);
            run<13, 16384>("box 128x32 strided (scan kernel order)", d, rows, ld, 0, 128, 2432, d_ctr, grid);
            run<10, 16384>("bulk 16 KB contiguous (tile-major)", d, rows, ld, 1, 128, 2432, d_ctr, grid);
            run<6, 16384>("bulk 16 KB contiguous (tile-major)", d, rows, ld, 1, 128, 2432, d_ctr, grid);
            run<13, 16384>("bulk 16 KB contiguous (tile-major)", d, rows, ld, 1, 128, 2432, d_ctr, grid);
            run<20, 8192>("box 64x32 strided", d, rows, ld, 0, 64, 2432, d_ctr, grid);
            run<6, 32768>("box 256x32 strided", d, rows, ld, 0, 256, 2432, d_ctr, grid);
        } else {
            run<5, 16384>("box 128x32 strided, 2 CTA/SM", d, rows, ld, 0, 128, 2432, d_ctr, grid);
            run<5, 16384>("bulk 16 KB contiguous, 2 CTA/SM", d, rows, ld, 1, 128, 2432, d_ctr, grid);
        }
    }
    return 0;
}
