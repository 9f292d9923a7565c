// Microbenchmark behind the K-batched fused kernel: what does it cost to gather rows of C fp32 (128 B at C = 32, 256 B
// at C = 64) from an L2-resident table by index -- the T[idx] gather of K-4 -- with
//   mode 0  LDG.256, 4 lanes per 128-byte row (what k_fusion_tc does)
//   mode 1  LDG.128, 8 lanes per row
//   mode 2  LDG.64, 16 lanes per row
//   mode 3  LDG.32, 32 lanes per row (fully coalesced)
//   mode 4  cp.async 16 B (LDGSTS) into shared memory, 8 lanes per row
//   mode 5  cp.async.bulk, one bulk copy per row (every lane issues its own row), mbarrier complete_tx
//   mode 6  TMA tile::gather4 through a tensor map: 4 rows per instruction, one lane issues
// Indices have the locality of a KNN table (consecutive rows draw from a window of `win` table rows).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o gather_patterns gather_patterns.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t a, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t par)
{
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(a), "r"(par) : "memory");
    } while (!done);
}

constexpr int kRowsPerBatch = 128;   // rows gathered per "slot" (one neighbour slot of a 128-cell tile)

// modes 0-3: register gathers.  A warp walks batches of 128 rows; sums what it loads.
template <int MODE, int C>
__global__ void __launch_bounds__(256) k_ldg(const float *__restrict__ T, const int32_t *__restrict__ idx, int64_t batches, float *out)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float acc = 0.f;
    constexpr int LPR = MODE == 0 ? C / 8 : MODE == 1 ? C / 4 : MODE == 2 ? C / 2 : C;   // lanes per row
    constexpr int RPI = 32 / (LPR > 32 ? 32 : LPR);                                       // rows per warp instruction
    for (int64_t bt = (int64_t)blockIdx.x * nw + warp; bt < batches; bt += (int64_t)gridDim.x * nw) {
        const int32_t *ib = idx + bt * kRowsPerBatch;
#pragma unroll 4
        for (int r = 0; r < kRowsPerBatch; r += RPI) {
            const int32_t pr = __ldg(ib + r + lane / LPR);
            const float *src = T + (size_t)pr * C;
            if (MODE == 0) {
                float4 a, b;
                const float *p = src + (lane % LPR) * 8;
                asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
                acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
            } else if (MODE == 1) {
                const float4 a = __ldg(reinterpret_cast<const float4 *>(src) + lane % LPR);
                acc += a.x + a.y + a.z + a.w;
            } else if (MODE == 2) {
                const float2 a = __ldg(reinterpret_cast<const float2 *>(src) + lane % LPR);
                acc += a.x + a.y;
            } else {
#pragma unroll
                for (int c = 0; c < C; c += 32) acc += __ldg(src + c + lane);
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// modes 4-6: asynchronous gathers into a ring of NB shared-memory slot buffers (128 rows x C fp32 each); every warp of the
// CTA then reads the slot back with LDS.128 (the operand build's read) and sums it.
template <int MODE, int C>
__global__ void __launch_bounds__(256) k_async(const __grid_constant__ CUtensorMap tm, const float *__restrict__ T, const int32_t *__restrict__ idx,
                                               int64_t batches, int nb, float *out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[16];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr uint32_t kSlot = kRowsPerBatch * C * 4;
    if (tid == 0) {
        for (int i = 0; i < 16; ++i) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t n_mine = (batches - blockIdx.x + gridDim.x - 1) / gridDim.x;
    float acc = 0.f;
    auto issue = [&](int64_t i) {   // called by warp 0 only
        const int slot = (int)(i % nb);
        const int32_t *ib = idx + (blockIdx.x + i * gridDim.x) * kRowsPerBatch;
        const uint32_t dst = smem_u32(smem) + slot * kSlot, bar = smem_u32(&bars[slot]);
        if (MODE == 5) {
            if (lane == 0) mbar_expect(bar, kSlot);
            __syncwarp();
            for (int r = lane; r < kRowsPerBatch; r += 32) {
                const int32_t pr = __ldg(ib + r);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + r * C * 4),
                             "l"(T + (size_t)pr * C), "r"(C * 4), "r"(bar) : "memory");
            }
        } else if (MODE == 6) {
            // each lane holds the indices of 4 rows; lane 0 issues the 32 gather4 instructions of the slot
            const int4 my = __ldg(reinterpret_cast<const int4 *>(ib) + lane);
            if (lane == 0) mbar_expect(bar, kSlot);
            __syncwarp();
#pragma unroll 1
            for (int g = 0; g < 32; ++g) {
                const int r0 = __shfl_sync(0xffffffffu, my.x, g), r1 = __shfl_sync(0xffffffffu, my.y, g);
                const int r2 = __shfl_sync(0xffffffffu, my.z, g), r3 = __shfl_sync(0xffffffffu, my.w, g);
                if (lane == 0)
                    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                                 ::"r"(dst + g * 4 * C * 4), "l"(&tm), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
            }
        }
    };
    if (MODE == 4) {
        // LDGSTS: all 8 warps issue; each warp instruction moves 32 x 16 B; commit groups per slot
        constexpr int LPR = C / 4;
        auto issue4 = [&](int64_t i) {
            const int slot = (int)(i % nb);
            const int32_t *ib = idx + (blockIdx.x + i * gridDim.x) * kRowsPerBatch;
            const uint32_t dst = smem_u32(smem) + slot * kSlot;
            for (int u = tid; u < kRowsPerBatch * LPR; u += 256) {
                const int r = u / LPR, q = u % LPR;
                const int32_t pr = __ldg(ib + r);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (r * LPR + q) * 16), "l"(T + (size_t)pr * C + q * 4) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        for (int64_t i = 0; i < nb - 1 && i < n_mine; ++i) issue4(i);
        for (int64_t i = 0; i < n_mine; ++i) {
            if (i + nb - 1 < n_mine) issue4(i + nb - 1); else asm volatile("cp.async.commit_group;" ::: "memory");
            if (nb == 5) asm volatile("cp.async.wait_group 4;" ::: "memory");   // the oldest of the 5 groups in flight has landed
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            const float4 *s = reinterpret_cast<const float4 *>(smem + (i % nb) * kSlot);
            for (int u = tid; u < kRowsPerBatch * C / 4; u += 256) { const float4 a = s[u]; acc += a.x + a.y + a.z + a.w; }
            __syncthreads();
        }
    } else {
        if (warp == 0) for (int64_t i = 0; i < nb - 1 && i < n_mine; ++i) issue(i);
        for (int64_t i = 0; i < n_mine; ++i) {
            if (warp == 0 && i + nb - 1 < n_mine) issue(i + nb - 1);
            mbar_wait(smem_u32(&bars[i % nb]), (uint32_t)((i / nb) & 1));
            const float4 *s = reinterpret_cast<const float4 *>(smem + (i % nb) * kSlot);
            for (int u = tid; u < kRowsPerBatch * C / 4; u += 256) { const float4 a = s[u]; acc += a.x + a.y + a.z + a.w; }
            __syncthreads();   // slot (i % nb) may be refilled by the issue of the next iteration
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int C>
void run(int rows_table, int64_t batches, int win, int ctas_per_sm, int nb, int only)
{
    std::vector<float> hT((size_t)rows_table * C);
    for (size_t i = 0; i < hT.size(); ++i) hT[i] = (float)(i % 97) * 0.01f;
    std::vector<int32_t> hidx((size_t)batches * kRowsPerBatch);
    uint32_t s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s >> 8; };
    for (int64_t b = 0; b < batches; ++b) {
        const int base = (int)(rnd() % (uint32_t)(rows_table - win));
        for (int r = 0; r < kRowsPerBatch; ++r) hidx[b * kRowsPerBatch + r] = base + (int)(rnd() % (uint32_t)win);
    }
    float *dT, *dout;
    int32_t *didx;
    CK(cudaMalloc(&dT, hT.size() * 4));
    CK(cudaMalloc(&didx, hidx.size() * 4));
    CK(cudaMalloc(&dout, 148 * 8 * 256 * 4));
    CK(cudaMemcpy(dT, hT.data(), hT.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(didx, hidx.data(), hidx.size() * 4, cudaMemcpyHostToDevice));
    // tensor map: 2-D (C, rows), box (C, 1)
    CUtensorMap tm;
    {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows_table};
        const cuuint64_t strides[1] = {(cuuint64_t)C * 4};
        const cuuint32_t box[2] = {(cuuint32_t)C, 1};
        const cuuint32_t el[2] = {1, 1};
        CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dT, dims, strides, box, el, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
    }
    const int grid = 148 * ctas_per_sm;
    const size_t smem = (size_t)nb * kRowsPerBatch * C * 4;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    auto launch = [&](int mode) {
        switch (mode) {
        case 0: k_ldg<0, C><<<grid, 256>>>(dT, didx, batches, dout); break;
        case 1: k_ldg<1, C><<<grid, 256>>>(dT, didx, batches, dout); break;
        case 2: k_ldg<2, C><<<grid, 256>>>(dT, didx, batches, dout); break;
        case 3: k_ldg<3, C><<<grid, 256>>>(dT, didx, batches, dout); break;
        case 4: k_async<4, C><<<grid, 256, smem>>>(tm, dT, didx, batches, nb, dout); break;
        case 5: k_async<5, C><<<grid, 256, smem>>>(tm, dT, didx, batches, nb, dout); break;
        case 6: k_async<6, C><<<grid, 256, smem>>>(tm, dT, didx, batches, nb, dout); break;
        }
    };
    CK(cudaFuncSetAttribute(k_async<4, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_async<5, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_async<6, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    std::vector<float> hout(148 * 8 * 256);
    for (int mode = 0; mode <= 6; ++mode) {
        if (only >= 0 && mode != only) continue;
        CK(cudaMemset(dout, 0, 148 * 8 * 256 * 4));
        launch(mode);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hout.data(), dout, hout.size() * 4, cudaMemcpyDeviceToHost));
        double sum = 0;
        for (float v : hout) sum += v;
        CK(cudaEventRecord(e0));
        for (int it = 0; it < 5; ++it) launch(mode);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= 5;
        const double rows = (double)batches * kRowsPerBatch;
        printf("C=%d mode=%d ctas/sm=%d nb=%d win=%d: %.1f us  %.2f rows/ns  %.2f TB/s  checksum %.6e\n", C, mode, ctas_per_sm, nb, win, ms * 1e3,
               rows / (ms * 1e6), rows * C * 4 / (ms * 1e9), sum);
    }
    cudaFree(dT);
    cudaFree(didx);
    cudaFree(dout);
}

int main(int argc, char **argv)
{
    const int C = argc > 1 ? atoi(argv[1]) : 32;
    const int ctas = argc > 2 ? atoi(argv[2]) : 2;
    const int nb = argc > 3 ? atoi(argv[3]) : 5;
    const int win = argc > 4 ? atoi(argv[4]) : 256;
    const int only = argc > 5 ? atoi(argv[5]) : -1;
    const int64_t batches = argc > 6 ? atoll(argv[6]) : 31500;   // 31500 x 128 rows = 4.03 M rows = scale 1 of configs[1]
    if (C == 32) run<32>(80000, batches, win, ctas, nb, only);
    else run<64>(80000, batches, win, ctas, nb, only);
    return 0;
}
