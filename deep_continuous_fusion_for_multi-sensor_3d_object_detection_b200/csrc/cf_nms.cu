// cf_nms.cu -- P-1..P-7 on device: get_bboxes compaction, SAT / IoU overlap masks, greedy scan, IoU matrix.
//
// Greedy NMS of test.py:142-175 ("keep box i iff it overlaps no previously kept box", input order) is
// split the classic way: (1) an embarrassingly parallel upper-triangular overlap bit-mask, one thread
// per (row, 64-column word), the 64 column boxes staged in shared memory; (2) a per-frame scan that
// walks the boxes in blocks of 64: one thread resolves the block against itself from registers, then
// all threads OR the rows of the newly kept boxes into the running `removed` bit-set.
#include "cf_boxgeom.cuh"

namespace cf {

constexpr int kNmsMaxWords = 256;  // cap <= 16384 boxes per frame

// ---------------------------------------------------------------------------------------- get_bboxes
// one CTA per frame; ordered compaction: anchor 0 cells row-major, then anchor 1 (test.py:97-107)
__global__ void __launch_bounds__(1024) k_get_bboxes(const float *__restrict__ cls, const float *__restrict__ box,
                                                     int32_t H, int32_t W, float thr, int32_t cap,
                                                     float *__restrict__ boxes, int32_t *__restrict__ counts,
                                                     int32_t *__restrict__ counts_raw)
{
    __shared__ int32_t warp_excl[32];
    __shared__ int32_t block_total;
    __shared__ int32_t base;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t plane = (int64_t)H * W;
    const float *cls_b = cls + (size_t)b * 4 * plane;
    const float *box_b = box + (size_t)b * 14 * plane;
    float *out = boxes + (size_t)b * cap * 7;
    if (tid == 0) base = 0;
    for (int a = 0; a < 2; ++a) {
        const float *score = cls_b + (size_t)(2 * a + 1) * plane;
        for (int64_t c0 = 0; c0 < plane; c0 += 1024) {
            const int64_t cell = c0 + tid;
            const bool flag = cell < plane && __ldg(score + cell) > thr;
            const unsigned bal = __ballot_sync(0xffffffffu, flag);
            const int32_t in_warp = __popc(bal & ((1u << lane) - 1u));
            if (lane == 0) warp_excl[warp] = __popc(bal);
            __syncthreads();
            if (warp == 0) {
                const int32_t w = warp_excl[lane];
                int32_t incl = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                warp_excl[lane] = incl - w;
                if (lane == 31) block_total = incl;
            }
            __syncthreads();
            const int32_t pos = base + warp_excl[warp] + in_warp;
            if (flag && pos < cap) {
#pragma unroll
                for (int c = 0; c < 7; ++c) out[(size_t)pos * 7 + c] = __ldg(box_b + (size_t)(7 * a + c) * plane + cell);
            }
            __syncthreads();
            if (tid == 0) base += block_total;
        }
    }
    __syncthreads();
    if (tid == 0) {
        counts[b] = base < cap ? base : cap;
        if (counts_raw) counts_raw[b] = base;
    }
}

// ---------------------------------------------------------------------------------------- masks
__global__ void __launch_bounds__(128) k_sat_prepare(const float *__restrict__ boxes, const int32_t *__restrict__ counts,
                                                     int32_t cap, SatBox *__restrict__ prep)
{
    const int b = blockIdx.y;
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= counts[b]) return;
    SatBox s;
    sat_prepare(boxes + ((size_t)b * cap + i) * 7, s);
    prep[(size_t)b * cap + i] = s;
}

// mask[b][i][w] bit t  <=>  j = 64w+t > i, j < n, boxes i and j overlap.  Only words w >= i/64 are written.
__global__ void __launch_bounds__(64) k_sat_mask(const SatBox *__restrict__ prep, const int32_t *__restrict__ counts,
                                                 int32_t cap, int32_t nw, unsigned long long *__restrict__ mask)
{
    const int b = blockIdx.z;
    const int32_t rb = blockIdx.y, cb = blockIdx.x;
    if (cb < rb) return;
    const int32_t n = counts[b];
    if (rb * 64 >= n) return;
    __shared__ SatBox cols[64];
    const int t = threadIdx.x;
    const int32_t jc = cb * 64 + t;
    if (jc < n) cols[t] = prep[(size_t)b * cap + jc];
    __syncthreads();
    const int32_t i = rb * 64 + t;
    if (i >= n) return;
    const SatBox A = prep[(size_t)b * cap + i];
    unsigned long long bits = 0;
    const int32_t jmax = min(64, n - cb * 64);
    for (int32_t jj = 0; jj < jmax; ++jj) {
        const int32_t j = cb * 64 + jj;
        if (j > i && sat_overlap(A, cols[jj])) bits |= 1ull << jj;
    }
    mask[((size_t)b * cap + i) * nw + cb] = bits;
}

__global__ void __launch_bounds__(128) k_iou_prepare(const float *__restrict__ boxes, const int32_t *__restrict__ counts,
                                                     int32_t cap, float nudge, IouBox *__restrict__ prep)
{
    const int b = blockIdx.y;
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= counts[b]) return;
    IouBox s;
    iou_prepare(boxes + ((size_t)b * cap + i) * 7, nudge, s);
    prep[(size_t)b * cap + i] = s;
}

// NMS_IOU predicate (test.py:122-134): candidate j (later, plain) against kept i (earlier, nudged)
__global__ void __launch_bounds__(64) k_iou_mask(const IouBox *__restrict__ plain, const IouBox *__restrict__ nudged,
                                                 const int32_t *__restrict__ counts, int32_t cap, int32_t nw,
                                                 double thr, unsigned long long *__restrict__ mask)
{
    const int b = blockIdx.z;
    const int32_t rb = blockIdx.y, cb = blockIdx.x;
    if (cb < rb) return;
    const int32_t n = counts[b];
    if (rb * 64 >= n) return;
    __shared__ IouBox cols[64];
    const int t = threadIdx.x;
    const int32_t jc = cb * 64 + t;
    if (jc < n) cols[t] = plain[(size_t)b * cap + jc];
    __syncthreads();
    const int32_t i = rb * 64 + t;
    if (i >= n) return;
    const IouBox Kept = nudged[(size_t)b * cap + i];
    unsigned long long bits = 0;
    const int32_t jmax = min(64, n - cb * 64);
    for (int32_t jj = 0; jj < jmax; ++jj) {
        const int32_t j = cb * 64 + jj;
        if (j <= i) continue;
        double i3, i2;
        iou_pair(cols[jj], Kept, i3, i2);
        if (i3 > thr) bits |= 1ull << jj;
    }
    mask[((size_t)b * cap + i) * nw + cb] = bits;
}

// ---------------------------------------------------------------------------------------- greedy scan
__global__ void __launch_bounds__(kNmsMaxWords) k_nms_scan(const unsigned long long *__restrict__ mask,
                                                           const int32_t *__restrict__ counts, int32_t cap, int32_t nw,
                                                           int32_t *__restrict__ keep_idx, int32_t *__restrict__ keep_count)
{
    __shared__ unsigned long long removed[kNmsMaxWords];
    __shared__ unsigned long long kept[kNmsMaxWords];
    __shared__ unsigned long long diag[64];
    __shared__ int32_t wsum[8];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int32_t n = counts[b];
    const unsigned long long *M = mask + (size_t)b * cap * nw;
    const int32_t nwa = (n + 63) >> 6;  // active words
    removed[tid] = 0ull;
    kept[tid] = 0ull;
    __syncthreads();
    for (int32_t blk = 0; blk < nwa; ++blk) {
        const int32_t i0 = blk * 64;
        if (tid < 64) diag[tid] = (i0 + tid < n) ? M[(size_t)(i0 + tid) * nw + blk] : 0ull;
        __syncthreads();
        if (tid == 0) {
            unsigned long long d[64];
#pragma unroll
            for (int t = 0; t < 64; ++t) d[t] = diag[t];
            unsigned long long rem = removed[blk], kb = 0ull;
            const int32_t lim = min(64, n - i0);
#pragma unroll
            for (int t = 0; t < 64; ++t) {
                const bool alive = t < lim && !((rem >> t) & 1ull);
                kb |= alive ? (1ull << t) : 0ull;
                rem |= alive ? d[t] : 0ull;
            }
            kept[blk] = kb;
        }
        __syncthreads();
        const int32_t w = blk + 1 + tid;
        if (w < nwa) {
            unsigned long long kb = kept[blk], acc = 0ull;
            while (kb) {
                const int t = __ffsll((long long)kb) - 1;
                kb &= kb - 1;
                acc |= M[(size_t)(i0 + t) * nw + w];
            }
            removed[w] |= acc;
        }
        __syncthreads();
    }
    // ordered compaction of the kept bits
    const int lane = tid & 31, warp = tid >> 5;
    const int32_t mine = tid < nwa ? __popcll(kept[tid]) : 0;
    int32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int32_t off = incl - mine;
    for (int w = 0; w < warp; ++w) off += wsum[w];
    int32_t total = 0;
    for (int w = 0; w < kNmsMaxWords / 32; ++w) total += wsum[w];
    int32_t *out = keep_idx + (size_t)b * cap;
    if (tid < nwa) {
        unsigned long long kb = kept[tid];
        while (kb) {
            const int t = __ffsll((long long)kb) - 1;
            kb &= kb - 1;
            out[off++] = tid * 64 + t;
        }
    }
    for (int32_t i = total + tid; i < cap; i += kNmsMaxWords) out[i] = -1;
    if (tid == 0) keep_count[b] = total;
}

// dense uint8 overlap matrix of one frame (diagnostics / mask-level parity)
__global__ void __launch_bounds__(128) k_sat_matrix(const float *__restrict__ boxes, int32_t n, uint8_t *__restrict__ m)
{
    const int32_t i = blockIdx.y;
    const int32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    SatBox A, Bx;
    sat_prepare(boxes + (size_t)i * 7, A);
    sat_prepare(boxes + (size_t)j * 7, Bx);
    m[(size_t)i * n + j] = sat_overlap(A, Bx) ? 1 : 0;
}

__global__ void __launch_bounds__(128) k_box_iou(const float *__restrict__ a, int32_t na, const float *__restrict__ bq,
                                                 int32_t nb, float nudge, double *__restrict__ iou3d,
                                                 double *__restrict__ iou2d)
{
    const int32_t i = blockIdx.y;
    const int32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nb) return;
    IouBox A, Bx;
    iou_prepare(a + (size_t)i * 7, 0.0f, A);
    iou_prepare(bq + (size_t)j * 7, nudge, Bx);
    double i3, i2;
    iou_pair(A, Bx, i3, i2);
    iou3d[(size_t)i * nb + j] = i3;
    iou2d[(size_t)i * nb + j] = i2;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace cf

extern "C" int cf_get_bboxes(const float *d_pred_cls, const float *d_pred_box, int32_t B, int32_t H, int32_t W,
                             float thr, int32_t cap, float *d_boxes, int32_t *d_counts, int32_t *d_counts_raw,
                             void *stream)
{
    using namespace cf;
    CF_TRY(require_sm100());
    CF_REQUIRE(d_pred_cls && d_pred_box && d_boxes && d_counts, CF_ERR_ARG, "cf_get_bboxes: null pointer");
    CF_REQUIRE(B > 0 && H > 0 && W > 0 && cap > 0, CF_ERR_ARG, "cf_get_bboxes: bad extents");
    k_get_bboxes<<<B, 1024, 0, (cudaStream_t)stream>>>(d_pred_cls, d_pred_box, H, W, thr, cap, d_boxes, d_counts,
                                                       d_counts_raw);
    count_launches(1);
    return launch_status("cf_get_bboxes");
}

extern "C" size_t cf_nms_workspace_bytes(int32_t B, int32_t cap)
{
    using namespace cf;
    if (B <= 0 || cap <= 0) return 0;
    const size_t nw = (size_t)(cap + 63) / 64;
    const size_t prep = align_up((size_t)B * cap * (sizeof(IouBox) > sizeof(SatBox) ? sizeof(IouBox) : sizeof(SatBox)), 256);
    return 2 * prep + (size_t)B * cap * nw * sizeof(unsigned long long);
}

static int nms_common_checks(const float *d_boxes, const int32_t *d_counts, int32_t B, int32_t cap,
                             int32_t *d_keep_idx, int32_t *d_keep_count, void *d_workspace, const char *who)
{
    using namespace cf;
    CF_TRY(require_sm100());
    CF_REQUIRE(d_boxes && d_counts && d_keep_idx && d_keep_count && d_workspace, CF_ERR_ARG, "%s: null pointer", who);
    CF_REQUIRE(B > 0 && B <= 65535 && cap > 0, CF_ERR_ARG, "%s: bad extents", who);
    CF_REQUIRE(cap <= 64 * kNmsMaxWords, CF_ERR_ARG, "%s: cap=%d exceeds %d boxes per frame", who, cap, 64 * kNmsMaxWords);
    CF_REQUIRE(aligned16(d_workspace), CF_ERR_ALIGN, "%s: workspace must be 16-byte aligned", who);
    return CF_OK;
}

extern "C" int cf_nms_sat(const float *d_boxes, const int32_t *d_counts, int32_t B, int32_t cap, int32_t *d_keep_idx,
                          int32_t *d_keep_count, void *d_workspace, void *stream)
{
    using namespace cf;
    CF_TRY(nms_common_checks(d_boxes, d_counts, B, cap, d_keep_idx, d_keep_count, d_workspace, "cf_nms_sat"));
    cudaStream_t st = (cudaStream_t)stream;
    const int32_t nw = (cap + 63) / 64;
    const size_t prep_bytes = align_up((size_t)B * cap * (sizeof(IouBox) > sizeof(SatBox) ? sizeof(IouBox) : sizeof(SatBox)), 256);
    SatBox *prep = (SatBox *)d_workspace;
    unsigned long long *mask = (unsigned long long *)((char *)d_workspace + 2 * prep_bytes);
    k_sat_prepare<<<dim3((cap + 127) / 128, B), 128, 0, st>>>(d_boxes, d_counts, cap, prep);
    k_sat_mask<<<dim3(nw, nw, B), 64, 0, st>>>(prep, d_counts, cap, nw, mask);
    k_nms_scan<<<B, kNmsMaxWords, 0, st>>>(mask, d_counts, cap, nw, d_keep_idx, d_keep_count);
    count_launches(3);
    return launch_status("cf_nms_sat");
}

extern "C" int cf_nms_iou(const float *d_boxes, const int32_t *d_counts, int32_t B, int32_t cap, float thr,
                          int32_t *d_keep_idx, int32_t *d_keep_count, void *d_workspace, void *stream)
{
    using namespace cf;
    CF_TRY(nms_common_checks(d_boxes, d_counts, B, cap, d_keep_idx, d_keep_count, d_workspace, "cf_nms_iou"));
    cudaStream_t st = (cudaStream_t)stream;
    const int32_t nw = (cap + 63) / 64;
    const size_t prep_bytes = align_up((size_t)B * cap * (sizeof(IouBox) > sizeof(SatBox) ? sizeof(IouBox) : sizeof(SatBox)), 256);
    IouBox *plain = (IouBox *)d_workspace;
    IouBox *nudged = (IouBox *)((char *)d_workspace + prep_bytes);
    unsigned long long *mask = (unsigned long long *)((char *)d_workspace + 2 * prep_bytes);
    k_iou_prepare<<<dim3((cap + 127) / 128, B), 128, 0, st>>>(d_boxes, d_counts, cap, 0.0f, plain);
    k_iou_prepare<<<dim3((cap + 127) / 128, B), 128, 0, st>>>(d_boxes, d_counts, cap, 0.0001f, nudged);
    k_iou_mask<<<dim3(nw, nw, B), 64, 0, st>>>(plain, nudged, d_counts, cap, nw, (double)thr, mask);
    k_nms_scan<<<B, kNmsMaxWords, 0, st>>>(mask, d_counts, cap, nw, d_keep_idx, d_keep_count);
    count_launches(4);
    return launch_status("cf_nms_iou");
}

extern "C" int cf_sat_matrix(const float *d_boxes, int32_t n, uint8_t *d_matrix, void *stream)
{
    using namespace cf;
    CF_TRY(require_sm100());
    CF_REQUIRE(d_boxes && d_matrix && n > 0 && n <= 65535, CF_ERR_ARG, "cf_sat_matrix: bad arguments");
    k_sat_matrix<<<dim3((n + 127) / 128, n), 128, 0, (cudaStream_t)stream>>>(d_boxes, n, d_matrix);
    count_launches(1);
    return launch_status("cf_sat_matrix");
}

extern "C" int cf_box_iou(const float *d_boxes_a, int32_t na, const float *d_boxes_b, int32_t nb, float nudge_b,
                          double *d_iou3d, double *d_iou2d, void *stream)
{
    using namespace cf;
    CF_TRY(require_sm100());
    CF_REQUIRE(d_boxes_a && d_boxes_b && d_iou3d && d_iou2d, CF_ERR_ARG, "cf_box_iou: null pointer");
    CF_REQUIRE(na > 0 && nb > 0 && na <= 65535, CF_ERR_ARG, "cf_box_iou: bad extents");
    k_box_iou<<<dim3((nb + 127) / 128, na), 128, 0, (cudaStream_t)stream>>>(d_boxes_a, na, d_boxes_b, nb, nudge_b,
                                                                           d_iou3d, d_iou2d);
    count_launches(1);
    return launch_status("cf_box_iou");
}
