// cf_loss.cu -- SURVEY 8(f-4): the target assignment of LossTotal (loss.py:74-127) on the device.
//
// Per frame the reference walks its (<= 20) ground-truth boxes in Python:
//   getPositionOfPositive (loss.py:74-110)  centre cell of the box in the prediction map
//        point_x = int((x * x_scale + x_offset) / reduced_scale)   (fp32 tensor arithmetic, truncated toward zero)
//        point_y = int((y * y_scale + y_offset) / reduced_scale)
//      boxes whose centre cell lies outside the map are skipped; the R x R window around it (clipped to the map) goes to
//      the positive list (duplicates of overlapping windows are kept), window order x-major; the regression list holds the
//      whole window (regress_type 0) or only the centre (regress_type 1); the positive list is shuffled
//      (np.random.shuffle) and cut to pos_sample_threshold entries.
//   getPositionOfNegative (loss.py:112-127)  rejection sampling: uniformly random cells that are not in the (cut) positive
//      list, until neg_sample_threshold + 1 of them have been drawn (the loop tests `sample > threshold` after counting).
// and then copies every list to the GPU with torch.tensor(...).cuda().
//
// Here one CTA per frame does the same with the random numbers supplied by the caller (the RNG contract that makes the
// result checkable against the reference run with the same draws):
//   shuffle     d_shuffle_keys (B, M * R * R) fp32: the positive list of length n is reordered by the STABLE ascending order of
//               its first n keys (entry j moves to position rank(j)); == np.random.shuffle replaced by list[argsort(keys[:n])]
//   negatives   d_candidates (B, L, 2) int32: the (x, y) draws of the rejection loop in order; the first
//               neg_threshold + 1 candidates that are not positive are kept (d_neg_count < neg_threshold + 1 tells the
//               caller that L draws were not enough)
// Everything stays on the device: no Python loop, no host list, no H2D copy, capturable in a CUDA graph.
#include "cf_common.cuh"

namespace cf {

namespace {

constexpr int kThreads = 256;
constexpr int kMaxEntries = 2048;   // M * R * R
constexpr int kMaxCells = 1 << 18;  // H * W (bitmap of 32 KB)

struct LossParams {
    const float *ref;           // (B, M, ref_stride) boxes, x = [0], y = [1]
    const int64_t *num_ref;     // (B)
    int32_t B, M, ref_stride, H, W;
    float x_scale, y_scale, x_offset, y_offset, reduced_scale;
    int32_t R, regress_type, pos_thr, neg_thr;
    const float *keys;          // (B, M * R * R)
    const int32_t *cand;        // (B, L, 2)
    int32_t L;
    int32_t *pos_cells;         // (B, pos_thr)        linear cell x * W + y, -1 padded
    int32_t *pos_count;         // (B)
    int32_t *neg_cells;         // (B, neg_thr + 1)
    int32_t *neg_count;         // (B)
    int32_t *reg_cells;         // (B, M, R * R)       regression cells of every box in window order, -1 = none
};

// exclusive prefix sum of one flag per thread over the CTA (ballot + warp totals); returns the thread's offset, *total = sum
__device__ __forceinline__ int block_scan(bool flag, int *warp_sums, int *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    __syncthreads();   // warp_sums may still be read from the previous call
    if (lane == 0) warp_sums[warp] = __popc(bal);
    __syncthreads();
    int base = 0, sum = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
        const int c = warp_sums[w];
        if (w < warp) base += c;
        sum += c;
    }
    *total = sum;
    return base + __popc(bal & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(kThreads) k_loss_targets(const LossParams p)
{
    extern __shared__ int32_t sm[];
    int32_t *list = sm;                                     // [M R R] positive cells in reference order
    uint32_t *bitmap = reinterpret_cast<uint32_t *>(sm + p.M * p.R * p.R);   // [(H W + 31) / 32]
    __shared__ int warp_sums[kThreads / 32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int RR = p.R * p.R, entries = p.M * RR, cells = p.H * p.W;
    const int64_t nr64 = p.num_ref[b];
    const int nref = (int)(nr64 < 0 ? 0 : nr64 > p.M ? p.M : nr64);

    for (int w = tid; w < (cells + 31) / 32; w += kThreads) bitmap[w] = 0u;

    // ---- positive windows, compacted in reference order (box-major, x offset, y offset) -------------------------------
    int n_pos = 0;
    for (int e0 = 0; e0 < entries; e0 += kThreads) {
        const int e = e0 + tid;
        bool valid = false;
        int32_t cell = -1;
        bool centre = false;
        if (e < entries) {
            const int i = e / RR, wdx = (e - i * RR) / p.R, wdy = e - i * RR - wdx * p.R;
            if (i < nref) {
                const float *rb = p.ref + ((size_t)b * p.M + i) * p.ref_stride;
                // loss.py:86-87: fp32 tensor arithmetic (multiply, add, divide -- separately rounded), int() truncates toward zero
                const int px = (int)__fdiv_rn(__fadd_rn(__fmul_rn(rb[0], p.x_scale), p.x_offset), p.reduced_scale);
                const int py = (int)__fdiv_rn(__fadd_rn(__fmul_rn(rb[1], p.y_scale), p.y_offset), p.reduced_scale);
                if (px >= 0 && px <= p.H - 1 && py >= 0 && py <= p.W - 1) {
                    const int x = px - p.R / 2 + wdx, y = py - p.R / 2 + wdy;
                    if (x >= 0 && x <= p.H - 1 && y >= 0 && y <= p.W - 1) {
                        valid = true;
                        cell = x * p.W + y;
                        centre = x == px && y == py;
                    }
                }
            }
            p.reg_cells[(size_t)b * entries + e] = valid && (p.regress_type == 0 || centre) ? cell : -1;
        }
        int chunk;
        const int off = block_scan(valid, warp_sums, &chunk);
        if (valid) list[n_pos + off] = cell;
        n_pos += chunk;
    }
    __syncthreads();

    // ---- shuffle = stable ascending order of the first n_pos keys; keep the first pos_thr ------------------------------
    const float *key = p.keys + (size_t)b * entries;
    const int n_keep = min(n_pos, p.pos_thr);
    for (int j = tid; j < n_pos; j += kThreads) {
        const float kj = key[j];
        int rank = 0;
        for (int k = 0; k < n_pos; ++k) {
            const float kk = key[k];
            rank += (kk < kj) || (kk == kj && k < j);
        }
        if (rank < n_keep) {
            const int32_t c = list[j];
            p.pos_cells[(size_t)b * p.pos_thr + rank] = c;
            atomicOr(&bitmap[c >> 5], 1u << (c & 31));
        }
    }
    for (int j = n_keep + tid; j < p.pos_thr; j += kThreads) p.pos_cells[(size_t)b * p.pos_thr + j] = -1;
    if (tid == 0) p.pos_count[b] = n_keep;
    __syncthreads();

    // ---- negatives: the first neg_thr + 1 candidates that are not positive --------------------------------------------
    const int want = p.neg_thr + 1;
    int n_neg = 0;
    for (int c0 = 0; c0 < p.L && n_neg < want; c0 += kThreads) {
        const int c = c0 + tid;
        bool ok = false;
        int32_t cell = -1;
        if (c < p.L) {
            const int x = p.cand[((size_t)b * p.L + c) * 2], y = p.cand[((size_t)b * p.L + c) * 2 + 1];
            if (x >= 0 && x < p.H && y >= 0 && y < p.W) {
                cell = x * p.W + y;
                ok = ((bitmap[cell >> 5] >> (cell & 31)) & 1u) == 0u;
            }
        }
        int chunk;
        const int off = block_scan(ok, warp_sums, &chunk);
        if (ok && n_neg + off < want) p.neg_cells[(size_t)b * want + n_neg + off] = cell;
        n_neg += chunk;
    }
    n_neg = min(n_neg, want);
    for (int j = n_neg + tid; j < want; j += kThreads) p.neg_cells[(size_t)b * want + j] = -1;
    if (tid == 0) p.neg_count[b] = n_neg;
}

}  // namespace
}  // namespace cf

extern "C" int cf_loss_targets(const float *d_ref_boxes, const int64_t *d_num_ref, int32_t B, int32_t M, int32_t ref_stride,
                               int32_t H, int32_t W, float x_scale, float y_scale, float x_offset, float y_offset,
                               float reduced_scale, int32_t positive_range, int32_t regress_type, int32_t pos_threshold,
                               int32_t neg_threshold, const float *d_shuffle_keys, const int32_t *d_candidates, int32_t L,
                               int32_t *d_pos_cells, int32_t *d_pos_count, int32_t *d_neg_cells, int32_t *d_neg_count,
                               int32_t *d_reg_cells, void *stream)
{
    using namespace cf;
    CF_TRY(require_sm100());
    CF_REQUIRE(d_ref_boxes && d_num_ref && d_shuffle_keys && d_candidates && d_pos_cells && d_pos_count && d_neg_cells &&
                   d_neg_count && d_reg_cells,
               CF_ERR_ARG, "cf_loss_targets: null pointer");
    CF_REQUIRE(B > 0 && B <= 65535 && M > 0 && ref_stride >= 2 && H > 0 && W > 0 && L > 0, CF_ERR_ARG, "cf_loss_targets: bad extents");
    CF_REQUIRE(positive_range >= 1 && positive_range <= 15 && M * positive_range * positive_range <= kMaxEntries, CF_ERR_ARG,
               "cf_loss_targets: %d boxes x %d^2 window cells exceed %d entries", M, positive_range, kMaxEntries);
    CF_REQUIRE((int64_t)H * W <= kMaxCells, CF_ERR_ARG, "cf_loss_targets: %d x %d map exceeds %d cells", H, W, kMaxCells);
    CF_REQUIRE(pos_threshold >= 1 && neg_threshold >= 0 && reduced_scale != 0.f, CF_ERR_ARG, "cf_loss_targets: bad thresholds / scale");
    LossParams p;
    p.ref = d_ref_boxes; p.num_ref = d_num_ref; p.B = B; p.M = M; p.ref_stride = ref_stride; p.H = H; p.W = W;
    p.x_scale = x_scale; p.y_scale = y_scale; p.x_offset = x_offset; p.y_offset = y_offset; p.reduced_scale = reduced_scale;
    p.R = positive_range; p.regress_type = regress_type; p.pos_thr = pos_threshold; p.neg_thr = neg_threshold;
    p.keys = d_shuffle_keys; p.cand = d_candidates; p.L = L;
    p.pos_cells = d_pos_cells; p.pos_count = d_pos_count; p.neg_cells = d_neg_cells; p.neg_count = d_neg_count; p.reg_cells = d_reg_cells;
    const size_t smem = (size_t)M * positive_range * positive_range * 4 + (((size_t)H * W + 31) / 32) * 4;
    if (smem > 48 * 1024)
        CF_TRY(cuda_status(cudaFuncSetAttribute(k_loss_targets, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "k_loss_targets smem attribute"));
    k_loss_targets<<<B, kThreads, smem, (cudaStream_t)stream>>>(p);
    count_launches(1);
    return launch_status("cf_loss_targets");
}
