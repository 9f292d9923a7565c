// Microbenchmark behind the strip pipeline of cf_fusion_strip.cu: how fast can ONE warp per CTA move (C rows x S cells)
// strips global -> shared -> global with cp.async.bulk (1-D bulk copies, one per channel row), as a function of the strip
// length S, the ring depth NB, CTAs per SM and the L2 cache hint.   nvcc -arch=sm_100a -O3 -o tma_strip_copy tma_strip_copy.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

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
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t pol, int hint)
{
    if (hint)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
    else
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, uint32_t src, uint32_t bytes, uint64_t pol, int hint)
{
    if (hint)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src), "r"(bytes), "l"(pol) : "memory");
    else
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

// mode 0: copy (load + store); mode 1: load only.   One warp per CTA does everything.
__global__ void __launch_bounds__(32) k_copy(const float *in, float *out, int C, int cells, int frames, int S, int nb, int hint, int mode)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[16];
    const int lane = threadIdx.x;
    const uint32_t strip_bytes = (uint32_t)C * S * 4;
    if (lane == 0) {
        for (int i = 0; i < nb; ++i) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const int spf = (cells + S - 1) / S, total = spf * frames, grid = gridDim.x;
    const int n_mine = (total - (int)blockIdx.x + grid - 1) / grid;
    int li = 0, si = 0;
    while (si < n_mine) {
        while (li < n_mine && li - si < nb) {
            const int slot = li % nb;
            if (li >= nb) {   // the store that last read this buffer: at most (li - si - ... ) groups may still be pending
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // conservative; see variant below
            }
            const int g = blockIdx.x + li * grid, b = g / spf, cell0 = (g - b * spf) * S, len = min(S, cells - cell0);
            const uint32_t bar = smem_u32(&bars[slot]);
            if (lane == 0) mbar_expect(bar, (uint32_t)C * len * 4);
            __syncwarp();
            for (int c = lane; c < C; c += 32)
                bulk_load(smem_u32(smem) + slot * strip_bytes + c * S * 4, in + ((size_t)b * C + c) * cells + cell0, len * 4, bar, pol, hint);
            ++li;
        }
        {
            const int slot = si % nb;
            mbar_wait(smem_u32(&bars[slot]), (si / nb) & 1);
            const int g = blockIdx.x + si * grid, b = g / spf, cell0 = (g - b * spf) * S, len = min(S, cells - cell0);
            if (mode == 0) {
                for (int c = lane; c < C; c += 32)
                    bulk_store(out + ((size_t)b * C + c) * cells + cell0, smem_u32(smem) + slot * strip_bytes + c * S * 4, len * 4, pol, hint);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            ++si;
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// variant with exact wait_group accounting: a buffer is reloaded when all but the (nb - 1) most recent stores have been read
template <int NBM1>
__global__ void __launch_bounds__(32) k_copy_exact(const float *in, float *out, int C, int cells, int frames, int S, int hint)
{
    constexpr int nb = NBM1 + 1;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[16];
    const int lane = threadIdx.x;
    const uint32_t strip_bytes = (uint32_t)C * S * 4;
    if (lane == 0) {
        for (int i = 0; i < nb; ++i) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const int spf = (cells + S - 1) / S, total = spf * frames, grid = gridDim.x;
    const int n_mine = (total - (int)blockIdx.x + grid - 1) / grid;
    auto load = [&](int i) {
        const int slot = i % nb;
        const int g = blockIdx.x + i * grid, b = g / spf, cell0 = (g - b * spf) * S, len = min(S, cells - cell0);
        const uint32_t bar = smem_u32(&bars[slot]);
        if (lane == 0) mbar_expect(bar, (uint32_t)C * len * 4);
        __syncwarp();
        for (int c = lane; c < C; c += 32)
            bulk_load(smem_u32(smem) + slot * strip_bytes + c * S * 4, in + ((size_t)b * C + c) * cells + cell0, len * 4, bar, pol, hint);
    };
    // prologue: nb - 1 loads in flight; steady state: wait strip i, store it, then refill the buffer of strip i - 1 (its store
    // is the second most recent group) with strip i + nb - 1
    for (int i = 0; i < nb - 1 && i < n_mine; ++i) load(i);
    for (int i = 0; i < n_mine; ++i) {
        const int slot = i % nb;
        mbar_wait(smem_u32(&bars[slot]), (i / nb) & 1);
        const int g = blockIdx.x + i * grid, b = g / spf, cell0 = (g - b * spf) * S, len = min(S, cells - cell0);
        for (int c = lane; c < C; c += 32)
            bulk_store(out + ((size_t)b * C + c) * cells + cell0, smem_u32(smem) + slot * strip_bytes + c * S * 4, len * 4, pol, hint);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (i + nb - 1 < n_mine) {
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // store i - 1 has left its buffer
            load(i + nb - 1);
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// tensor-map variant: ONE cp.async.bulk.tensor.2d per strip and direction (box = S cells x C rows), issued by one thread
__global__ void __launch_bounds__(32) k_copy_tmap(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out, int C,
                                                  int cells, int frames, int S, int nb, int mode)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[16];
    const int lane = threadIdx.x;
    const uint32_t strip_bytes = (uint32_t)C * S * 4;
    if (lane == 0) {
        for (int i = 0; i < nb; ++i) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane != 0) return;
    const int spf = (cells + S - 1) / S, total = spf * frames, grid = gridDim.x;
    const int n_mine = (total - (int)blockIdx.x + grid - 1) / grid;
    auto load = [&](int i) {
        const int slot = i % nb;
        const int g = blockIdx.x + i * grid, b = g / spf, cell0 = (g - b * spf) * S;
        const uint32_t bar = smem_u32(&bars[slot]);
        mbar_expect(bar, strip_bytes);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                         smem_u32(smem) + slot * strip_bytes),
                     "l"(&tm_in), "r"(cell0), "r"(b * C), "r"(bar)
                     : "memory");
    };
    for (int i = 0; i < nb - 1 && i < n_mine; ++i) load(i);
    for (int i = 0; i < n_mine; ++i) {
        const int slot = i % nb;
        mbar_wait(smem_u32(&bars[slot]), (i / nb) & 1);
        const int g = blockIdx.x + i * grid, b = g / spf, cell0 = (g - b * spf) * S;
        if (mode == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&tm_out), "r"(cell0), "r"(b * C),
                         "r"(smem_u32(smem) + slot * strip_bytes)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (i + nb - 1 < n_mine) {
            if (mode == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            load(i + nb - 1);
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    const int C = 32, cells = 560000, frames = 4;
    const size_t n = (size_t)frames * C * cells;
    float *in, *out;
    cudaMalloc(&in, n * 4);
    cudaMalloc(&out, n * 4);
    cudaMemset(in, 1, n * 4);
    cudaMemset(out, 0, n * 4);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(k_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(k_copy_exact<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(k_copy_exact<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto run = [&](const char *name, auto launch, double bytes) {
        for (int i = 0; i < 3; ++i) launch();
        cudaDeviceSynchronize();
        float best = 1e9f;
        for (int i = 0; i < 7; ++i) {
            cudaEventRecord(e0);
            launch();
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            best = ms < best ? ms : best;
        }
        cudaError_t err = cudaGetLastError();
        printf("%-58s %8.1f us  %7.0f GB/s  %s\n", name, best * 1e3, bytes / best * 1e-6, err == cudaSuccess ? "" : cudaGetErrorString(err));
    };
    const double copy_bytes = 2.0 * n * 4, load_bytes = 1.0 * n * 4;
    char name[128];
    for (int per_sm = 1; per_sm <= 4; per_sm *= 2)
        for (int S : {256, 512, 1024})
            for (int nb : {2, 4, 6}) {
                const int smem = nb * C * S * 4;
                if (smem * per_sm > 220 * 1024) continue;
                for (int hint = 0; hint < 2; ++hint) {
                    snprintf(name, sizeof name, "copy  ctas/sm %d S %4d nb %d hint %d (conservative wait)", per_sm, S, nb, hint);
                    run(name, [&] { k_copy<<<sms * per_sm, 32, smem>>>(in, out, C, cells, frames, S, nb, hint, 0); }, copy_bytes);
                }
                snprintf(name, sizeof name, "load  ctas/sm %d S %4d nb %d hint 1", per_sm, S, nb);
                run(name, [&] { k_copy<<<sms * per_sm, 32, smem>>>(in, out, C, cells, frames, S, nb, 1, 1); }, load_bytes);
            }
    for (int per_sm = 1; per_sm <= 2; ++per_sm)
        for (int S : {256, 512}) {
            if (4 * C * S * 4 * per_sm <= 220 * 1024) {
                snprintf(name, sizeof name, "copy  ctas/sm %d S %4d nb 4 hint 1 (exact wait)", per_sm, S);
                run(name, [&] { k_copy_exact<3><<<sms * per_sm, 32, 4 * C * S * 4>>>(in, out, C, cells, frames, S, 1); }, copy_bytes);
            }
            if (6 * C * S * 4 * per_sm <= 220 * 1024) {
                snprintf(name, sizeof name, "copy  ctas/sm %d S %4d nb 6 hint 1 (exact wait)", per_sm, S);
                run(name, [&] { k_copy_exact<5><<<sms * per_sm, 32, 6 * C * S * 4>>>(in, out, C, cells, frames, S, 1); }, copy_bytes);
            }
        }
    {
        EncodeFn encode = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres);
        if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
        cudaFuncSetAttribute(k_copy_tmap, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        for (int per_sm = 1; per_sm <= 2; ++per_sm)
            for (int S : {128, 256})
                for (int nb : {2, 3, 4, 6}) {
                    const int smem = nb * C * S * 4;
                    if (smem * per_sm > 220 * 1024) continue;
                    CUtensorMap tin, tout;
                    cuuint64_t dims[2] = {(cuuint64_t)cells, (cuuint64_t)frames * C};
                    cuuint64_t strides[1] = {(cuuint64_t)cells * 4};
                    cuuint32_t box[2] = {(cuuint32_t)S, (cuuint32_t)C};
                    cuuint32_t estr[2] = {1, 1};
                    CUresult r1 = encode(&tin, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    CUresult r2 = encode(&tout, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    if (r1 || r2) { printf("encode failed %d %d\n", (int)r1, (int)r2); continue; }
                    snprintf(name, sizeof name, "tmap copy ctas/sm %d S %4d nb %d", per_sm, S, nb);
                    run(name, [&] { k_copy_tmap<<<sms * per_sm, 32, smem>>>(tin, tout, C, cells, frames, S, nb, 0); }, copy_bytes);
                    snprintf(name, sizeof name, "tmap load ctas/sm %d S %4d nb %d", per_sm, S, nb);
                    run(name, [&] { k_copy_tmap<<<sms * per_sm, 32, smem>>>(tin, tout, C, cells, frames, S, nb, 1); }, load_bytes);
                }
        // check the copy
        cudaMemset(out, 0, n * 4);
        {
            CUtensorMap tin, tout;
            cuuint64_t dims[2] = {(cuuint64_t)cells, (cuuint64_t)frames * C};
            cuuint64_t strides[1] = {(cuuint64_t)cells * 4};
            cuuint32_t box[2] = {256, (cuuint32_t)C};
            cuuint32_t estr[2] = {1, 1};
            encode(&tin, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            encode(&tout, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            std::vector<float> h(n);
            for (size_t i = 0; i < n; ++i) h[i] = (float)(i % 1000003);
            cudaMemcpy(in, h.data(), n * 4, cudaMemcpyHostToDevice);
            k_copy_tmap<<<sms, 32, 4 * C * 256 * 4>>>(tin, tout, C, cells, frames, 256, 4, 0);
            std::vector<float> h2(n);
            cudaMemcpy(h2.data(), out, n * 4, cudaMemcpyDeviceToHost);
            size_t bad = 0;
            for (size_t i = 0; i < n; ++i) bad += h[i] != h2[i];
            printf("tmap copy check: %zu mismatches of %zu (%s)\n", bad, n, cudaGetErrorString(cudaGetLastError()));
        }
    }
    // reference: cudaMemcpy device to device
    run("cudaMemcpyAsync d2d", [&] { cudaMemcpyAsync(out, in, n * 4, cudaMemcpyDeviceToDevice); }, copy_bytes);
    return 0;
}
