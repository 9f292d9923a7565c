// cf_knn.cu -- K-2: bounded-radius top-K nearest LiDAR points per BEV cell over the bucket grid.
//
// One thread per BEV cell.  The K best candidates live in registers as 64-bit keys
// (bits(d2) << 32 | index): d2 >= 0, so the IEEE bit pattern orders like the value and one unsigned
// compare implements the total order (d2, idx) of SURVEY Appendix A5.  Buckets are visited in
// Chebyshev rings around the cell's own bucket; inside a ring every bucket-row segment is one
// contiguous slice of the sorted point array.  The search stops as soon as the K-th best distance is
// strictly below a conservative lower bound of everything not yet visited, or when the bounding
// square of the radius disc is exhausted.  Distances use __fsub_rn/__fmul_rn/__fadd_rn so no FMA
// contraction can make d2 differ from the CPU oracle (Appendix A2).
#include "cf_common.cuh"

namespace cf {

struct KnnGeom {
    float x0, y0, dx, dy;  // cell centre: x0 + i*dx, y0 + j*dy
    float radius, r2;
    int32_t H, W;
};

template <int K>
__device__ __forceinline__ void knn_insert(unsigned long long (&best)[K], unsigned long long key)
{
    // precondition: key < best[K-1]
    best[K - 1] = key;
#pragma unroll
    for (int q = K - 1; q > 0; --q) {
        const unsigned long long a = best[q - 1], b = best[q];
        const bool sw = b < a;
        best[q - 1] = sw ? b : a;
        best[q] = sw ? a : b;
    }
}

template <int K>
__device__ __forceinline__ void knn_scan(const float4 *__restrict__ sp, int32_t s, int32_t e, float cx, float cy,
                                         float r2, unsigned long long (&best)[K])
{
    for (int32_t p = s; p < e; ++p) {
        const float4 q = __ldg(sp + p);
        const float ddx = __fsub_rn(q.x, cx);
        const float ddy = __fsub_rn(q.y, cy);
        const float d2 = __fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy));
        if (d2 <= r2) {
            const unsigned long long key =
                ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)__float_as_uint(q.w);
            if (key < best[K - 1]) knn_insert<K>(best, key);
        }
    }
}

template <int K>
__global__ void __launch_bounds__(128) k_knn_query(const int32_t *__restrict__ bucket_start,
                                                   const float4 *__restrict__ sorted, int32_t N, BucketGrid g,
                                                   KnnGeom q, int32_t *__restrict__ knn_idx)
{
    const int b = blockIdx.y;
    const int64_t cells = (int64_t)q.H * q.W;
    const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= cells) return;
    const int32_t i = (int32_t)(cell / q.W), j = (int32_t)(cell - (int64_t)i * q.W);
    const float cx = __fadd_rn(q.x0, __fmul_rn((float)i, q.dx));
    const float cy = __fadd_rn(q.y0, __fmul_rn((float)j, q.dy));

    const int32_t G = g.nbx * g.nby;
    const int32_t *__restrict__ bs = bucket_start + (size_t)b * (G + 1);
    const float4 *__restrict__ sp = sorted + (size_t)b * N;

    unsigned long long best[K];
#pragma unroll
    for (int k = 0; k < K; ++k) best[k] = ~0ull;

    // Conservative slack for every comparison between a bucket boundary and a distance: bucket
    // assignment and boundary coordinates carry a few ulp of fp32 error (<< 1e-4*cell for |coords| < 1e4).
    const float margin = 0.01f * g.cell;
    const float reach = q.radius + margin;
    const int32_t bxlo = bucket_coord(cx - reach, g.gx0, g.inv_cell, g.nbx);
    const int32_t bxhi = bucket_coord(cx + reach, g.gx0, g.inv_cell, g.nbx);
    const int32_t bylo = bucket_coord(cy - reach, g.gy0, g.inv_cell, g.nby);
    const int32_t byhi = bucket_coord(cy + reach, g.gy0, g.inv_cell, g.nby);
    const int32_t bcx = bucket_coord(cx, g.gx0, g.inv_cell, g.nbx);
    const int32_t bcy = bucket_coord(cy, g.gy0, g.inv_cell, g.nby);
    const int32_t rho_max = max(max(bcx - bxlo, bxhi - bcx), max(bcy - bylo, byhi - bcy));

    for (int32_t rho = 0; rho <= rho_max; ++rho) {
        const int32_t xa = bcx - rho, xb = bcx + rho;
        const int32_t ya = bcy - rho, yb = bcy + rho;
        const int32_t yac = max(ya, bylo), ybc = min(yb, byhi);
        if (yac <= ybc) {
            if (xa >= bxlo) {  // first row of the ring, full width
                const int32_t s = __ldg(bs + xa * g.nby + yac), e = __ldg(bs + xa * g.nby + ybc + 1);
                knn_scan<K>(sp, s, e, cx, cy, q.r2, best);
            }
            if (rho > 0 && xb <= bxhi) {  // last row
                const int32_t s = __ldg(bs + xb * g.nby + yac), e = __ldg(bs + xb * g.nby + ybc + 1);
                knn_scan<K>(sp, s, e, cx, cy, q.r2, best);
            }
        }
        if (rho > 0) {  // the two side columns of the rows in between
            const int32_t x_lo = max(xa + 1, bxlo), x_hi = min(xb - 1, bxhi);
            for (int32_t x = x_lo; x <= x_hi; ++x) {
                if (ya >= bylo) {
                    const int32_t s = __ldg(bs + x * g.nby + ya), e = __ldg(bs + x * g.nby + ya + 1);
                    knn_scan<K>(sp, s, e, cx, cy, q.r2, best);
                }
                if (yb <= byhi) {
                    const int32_t s = __ldg(bs + x * g.nby + yb), e = __ldg(bs + x * g.nby + yb + 1);
                    knn_scan<K>(sp, s, e, cx, cy, q.r2, best);
                }
            }
        }
        // early exit: everything not yet visited lies outside the (2rho+1)^2 block of buckets
        if (rho < rho_max && best[K - 1] != ~0ull) {
            float lb = 3.0e38f;
            if (xa > bxlo) lb = fminf(lb, cx - (g.gx0 + (float)xa * g.cell));
            if (xb < bxhi) lb = fminf(lb, (g.gx0 + (float)(xb + 1) * g.cell) - cx);
            if (ya > bylo) lb = fminf(lb, cy - (g.gy0 + (float)ya * g.cell));
            if (yb < byhi) lb = fminf(lb, (g.gy0 + (float)(yb + 1) * g.cell) - cy);
            const float lbs = lb - margin;
            const float kth = __uint_as_float((unsigned)(best[K - 1] >> 32));
            if (lbs > 0.0f && kth < lbs * lbs * 0.999999f) break;
        }
    }

    int32_t *out = knn_idx + ((size_t)b * cells + cell) * K;
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] = best[k] == ~0ull ? -1 : (int32_t)(unsigned)(best[k] & 0xffffffffull);
}

template <int K>
static void launch_knn(const int32_t *bs, const float4 *sp, int32_t B, int32_t N, const BucketGrid &g,
                       const KnnGeom &q, int32_t *out, cudaStream_t st)
{
    const int64_t cells = (int64_t)q.H * q.W;
    dim3 grid((unsigned)ceil_div64(cells, 128), (unsigned)B);
    k_knn_query<K><<<grid, 128, 0, st>>>(bs, sp, N, g, q, out);
}

// coarse[b,i,j,:] = fine[b, i*step, j*step, :]
__global__ void __launch_bounds__(256) k_knn_subsample(const int32_t *__restrict__ fine, int32_t Hf, int32_t Wf,
                                                       int32_t step, int32_t *__restrict__ coarse, int32_t Hc, int32_t Wc,
                                                       int32_t K)
{
    const int b = blockIdx.y;
    const int64_t n = (int64_t)Hc * Wc * K;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int32_t k = (int32_t)(t % K);
        const int64_t cell = t / K;
        const int32_t i = (int32_t)(cell / Wc), j = (int32_t)(cell - (int64_t)i * Wc);
        coarse[(size_t)b * n + t] = __ldg(fine + (((size_t)b * Hf + (size_t)i * step) * Wf + (size_t)j * step) * K + k);
    }
}

}  // namespace cf

extern "C" int cf_knn_subsample(const int32_t *d_fine, int32_t B, int32_t Hf, int32_t Wf, int32_t step,
                                int32_t *d_coarse, int32_t Hc, int32_t Wc, int32_t K, void *stream)
{
    using namespace cf;
    CF_TRY(require_sm100());
    CF_REQUIRE(d_fine && d_coarse, CF_ERR_ARG, "cf_knn_subsample: null pointer");
    CF_REQUIRE(B > 0 && B <= 65535 && step >= 1 && K >= 1 && K <= CF_MAX_K && Hc > 0 && Wc > 0, CF_ERR_ARG,
               "cf_knn_subsample: bad extents");
    CF_REQUIRE((int64_t)(Hc - 1) * step < Hf && (int64_t)(Wc - 1) * step < Wf, CF_ERR_ARG,
               "cf_knn_subsample: coarse grid (%d,%d) x step %d does not fit in the fine grid (%d,%d)", Hc, Wc, step, Hf, Wf);
    const int64_t n = (int64_t)Hc * Wc * K;
    const int blocks = (int)std::min<int64_t>(ceil_div64(n, 256), 148 * 8);
    k_knn_subsample<<<dim3(blocks, B), 256, 0, (cudaStream_t)stream>>>(d_fine, Hf, Wf, step, d_coarse, Hc, Wc, K);
    count_launches(1);
    return launch_status("cf_knn_subsample");
}

extern "C" int cf_knn_query(const int32_t *d_bucket_start, const float *d_sorted, int32_t B, int32_t N, float gx0,
                            float gy0, float cell, int32_t nbx, int32_t nby, int32_t H, int32_t W, float x0,
                            float y0, float dx, float dy, float radius, int32_t K, int32_t *d_knn_idx,
                            void *stream)
{
    using namespace cf;
    CF_TRY(require_sm100());
    CF_REQUIRE(d_bucket_start && d_sorted && d_knn_idx, CF_ERR_ARG, "cf_knn_query: null pointer");
    CF_REQUIRE(B > 0 && N > 0 && H > 0 && W > 0 && nbx > 0 && nby > 0 && cell > 0.0f, CF_ERR_ARG,
               "cf_knn_query: bad extents");
    CF_REQUIRE(K >= 1 && K <= CF_MAX_K, CF_ERR_ARG, "cf_knn_query: K=%d outside [1,%d]", K, CF_MAX_K);
    CF_REQUIRE(radius > 0.0f && radius < 1.0e4f, CF_ERR_ARG, "cf_knn_query: radius=%g", (double)radius);
    CF_REQUIRE(aligned16(d_sorted), CF_ERR_ALIGN, "cf_knn_query: d_sorted must be 16-byte aligned");
    CF_REQUIRE(B <= 65535, CF_ERR_ARG, "cf_knn_query: B too large");
    cudaStream_t st = (cudaStream_t)stream;
    BucketGrid g{gx0, gy0, cell, 1.0f / cell, nbx, nby};
    KnnGeom q{x0, y0, dx, dy, radius, radius * radius, H, W};
    const float4 *sp = (const float4 *)d_sorted;
    switch (K) {
#define CF_KNN_CASE(k) \
    case k: launch_knn<k>(d_bucket_start, sp, B, N, g, q, d_knn_idx, st); break;
        CF_KNN_CASE(1) CF_KNN_CASE(2) CF_KNN_CASE(3) CF_KNN_CASE(4) CF_KNN_CASE(5) CF_KNN_CASE(6)
        CF_KNN_CASE(7) CF_KNN_CASE(8) CF_KNN_CASE(9) CF_KNN_CASE(10) CF_KNN_CASE(11) CF_KNN_CASE(12)
        CF_KNN_CASE(13) CF_KNN_CASE(14) CF_KNN_CASE(15) CF_KNN_CASE(16)
#undef CF_KNN_CASE
    }
    count_launches(1);
    return launch_status("cf_knn_query");
}
