// cf_bucket.cu -- K-1: counting sort of each frame's valid LiDAR points into a uniform BEV grid.
//   histogram (global atomics, one int per bucket) -> exclusive prefix (one CTA per frame, warp-shuffle
//   scan) -> scatter (atomic cursor).  The order of points inside a bucket is not deterministic; the
//   KNN result is, because the query orders candidates by the total order (d2, original index).
// Input layout: sample["pointcloud_raw"] / ["num_points_raw"], data_import_carla.py:261-267.
#include "cf_common.cuh"

namespace cf {

__global__ void __launch_bounds__(256) k_bucket_hist(const float *__restrict__ points,
                                                     const int64_t *__restrict__ num_points, int32_t N,
                                                     BucketGrid g, int32_t *__restrict__ counts)
{
    const int b = blockIdx.y;
    const int32_t n = valid_points(num_points, b, N);
    const int32_t G = g.nbx * g.nby;
    for (int32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const float *q = points + ((size_t)b * N + p) * 3;
        const int32_t bx = bucket_coord(q[0], g.gx0, g.inv_cell, g.nbx);
        const int32_t by = bucket_coord(q[1], g.gy0, g.inv_cell, g.nby);
        atomicAdd(&counts[(size_t)b * G + bx * g.nby + by], 1);
    }
}

// one CTA of 1024 threads per frame: start[g] = sum_{h<g} count[h]; start[G] = n; cursor := start
__global__ void __launch_bounds__(1024) k_bucket_scan(int32_t *__restrict__ counts_cursor, int32_t G,
                                                      int32_t *__restrict__ bucket_start)
{
    __shared__ int32_t warp_sums[32];
    const int b = blockIdx.x;
    int32_t *cnt = counts_cursor + (size_t)b * G;
    int32_t *start = bucket_start + (size_t)b * (G + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t per = (G + 1023) / 1024;
    const int32_t lo = tid * per, hi = min(lo + per, G);
    int32_t local = 0;
    for (int32_t i = lo; i < hi; ++i) local += cnt[i];
    int32_t incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int32_t w = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t v = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += v;
        }
        warp_sums[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    int32_t run = incl - local + (warp > 0 ? warp_sums[warp - 1] : 0);
    for (int32_t i = lo; i < hi; ++i) {
        const int32_t c = cnt[i];
        start[i] = run;
        cnt[i] = run;  // becomes the scatter cursor
        run += c;
    }
    if (tid == 1023) start[G] = warp_sums[31];
}

// The same scan with the frame's counts staged in shared memory: global memory is only touched with coalesced accesses
// (thread t moves elements t, t + 1024, ...), each thread scans its contiguous run out of shared memory (run length odd or
// not, consecutive threads start `per` words apart: at most 2-way bank conflicts) -- 24 -> ~6 us per launch on the step's
// critical path (bucket -> KNN -> everything else).  Needs G * 4 bytes of shared memory.
__global__ void __launch_bounds__(1024) k_bucket_scan_smem(int32_t *__restrict__ counts_cursor, int32_t G,
                                                           int32_t *__restrict__ bucket_start)
{
    extern __shared__ int32_t sc[];
    __shared__ int32_t warp_sums[32];
    const int b = blockIdx.x;
    int32_t *cnt = counts_cursor + (size_t)b * G;
    int32_t *start = bucket_start + (size_t)b * (G + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int32_t i = tid; i < G; i += 1024) sc[i] = cnt[i];
    __syncthreads();
    const int32_t per = (G + 1023) / 1024;
    const int32_t lo = min(tid * per, G), hi = min(lo + per, G);
    int32_t local = 0;
    for (int32_t i = lo; i < hi; ++i) local += sc[i];
    int32_t incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int32_t w = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t v = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += v;
        }
        warp_sums[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    int32_t run = incl - local + (warp > 0 ? warp_sums[warp - 1] : 0);
    for (int32_t i = lo; i < hi; ++i) {
        const int32_t c = sc[i];
        sc[i] = run;
        run += c;
    }
    __syncthreads();
    for (int32_t i = tid; i < G; i += 1024) {
        const int32_t v = sc[i];
        start[i] = v;
        cnt[i] = v;  // becomes the scatter cursor
    }
    if (tid == 1023) start[G] = warp_sums[31];
}

__global__ void __launch_bounds__(256) k_bucket_scatter(const float *__restrict__ points,
                                                        const int64_t *__restrict__ num_points, int32_t N,
                                                        BucketGrid g, int32_t *__restrict__ cursor,
                                                        float4 *__restrict__ sorted)
{
    const int b = blockIdx.y;
    const int32_t n = valid_points(num_points, b, N);
    const int32_t G = g.nbx * g.nby;
    for (int32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const float *q = points + ((size_t)b * N + p) * 3;
        const float x = q[0], y = q[1], z = q[2];
        const int32_t bx = bucket_coord(x, g.gx0, g.inv_cell, g.nbx);
        const int32_t by = bucket_coord(y, g.gy0, g.inv_cell, g.nby);
        const int32_t pos = atomicAdd(&cursor[(size_t)b * G + bx * g.nby + by], 1);
        sorted[(size_t)b * N + pos] = make_float4(x, y, z, __int_as_float(p));
    }
}

}  // namespace cf

extern "C" size_t cf_bucket_workspace_bytes(int32_t B, int32_t nbx, int32_t nby)
{
    if (B <= 0 || nbx <= 0 || nby <= 0) return 0;
    return (size_t)B * (size_t)nbx * (size_t)nby * sizeof(int32_t);
}

extern "C" int cf_bucket_points(const float *d_points, const int64_t *d_num_points, int32_t B, int32_t N,
                                float gx0, float gy0, float cell, int32_t nbx, int32_t nby,
                                int32_t *d_bucket_start, float *d_sorted, void *d_workspace, void *stream)
{
    using namespace cf;
    CF_TRY(require_sm100());
    CF_REQUIRE(d_points && d_num_points && d_bucket_start && d_sorted && d_workspace, CF_ERR_ARG,
               "cf_bucket_points: null pointer");
    CF_REQUIRE(B > 0 && N > 0 && nbx > 0 && nby > 0 && cell > 0.0f, CF_ERR_ARG,
               "cf_bucket_points: bad extents B=%d N=%d nbx=%d nby=%d cell=%g", B, N, nbx, nby, (double)cell);
    CF_REQUIRE((int64_t)nbx * nby < (1 << 24), CF_ERR_ARG, "cf_bucket_points: bucket grid too large");
    CF_REQUIRE(aligned16(d_sorted), CF_ERR_ALIGN, "cf_bucket_points: d_sorted must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int32_t G = nbx * nby;
    BucketGrid g{gx0, gy0, cell, 1.0f / cell, nbx, nby};
    int32_t *cursor = (int32_t *)d_workspace;
    CF_TRY(cuda_status(cudaMemsetAsync(cursor, 0, (size_t)B * G * sizeof(int32_t), st), "cf_bucket_points memset"));
    const int blocks = (int)std::min<int64_t>(ceil_div64(N, 256), 4096);
    k_bucket_hist<<<dim3(blocks, B), 256, 0, st>>>(d_points, d_num_points, N, g, cursor);
    const size_t scan_smem = (size_t)G * sizeof(int32_t);
    if (scan_smem <= 200 * 1024) {
        static size_t attr = 0;
        if (scan_smem > attr) {
            CF_TRY(cuda_status(cudaFuncSetAttribute(k_bucket_scan_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem),
                               "k_bucket_scan smem attribute"));
            attr = scan_smem;
        }
        k_bucket_scan_smem<<<B, 1024, scan_smem, st>>>(cursor, G, d_bucket_start);
    } else {
        k_bucket_scan<<<B, 1024, 0, st>>>(cursor, G, d_bucket_start);
    }
    k_bucket_scatter<<<dim3(blocks, B), 256, 0, st>>>(d_points, d_num_points, N, g, cursor, (float4 *)d_sorted);
    count_launches(3);
    return launch_status("cf_bucket_points");
}
