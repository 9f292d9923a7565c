// cf_mlp_tc.cu -- K-4 on the 5th-generation tensor cores: per-neighbour MLP layer 2, K-sum-pool, layer 3 and
// the BEV add for one backbone scale, fused so that neither the gathered rows nor the hidden activations
// ever leave the SM.
//
// Tile = 128 consecutive BEV cells (= the M of a cta_group::1 UMMA, one TMEM lane per cell).  Per tile:
//   for each neighbour slot k:
//     A_k[128 x C]  = relu(T[idx_k] - e_cell)            built by CUDA cores straight into shared memory in the
//                                                         UMMA operand layout, bf16 (+ bf16 residual in fp32 mode)
//     acc[128 x C]  = A_k * W2^T                          tcgen05.mma, accumulator in TMEM
//     pooled       += valid ? relu(acc + b2) : 0          tcgen05.ld -> registers -> tcgen05.st (pooled lives in TMEM)
//   acc = pooled * W3^T                                   tcgen05.mma
//   out = bev + acc + n_valid * b3                        epilogue: coalesced NCHW read of bev, write of out
// Slots for which no cell of the tile has a neighbour are skipped (most far-field tiles skip everything).
//
// Precision modes
//   CF_MODE_BF16: operands rounded to bf16, fp32 accumulate (tolerance 1e-2, Appendix A12).
//   CF_MODE_FP32: every fp32 operand x is split into bf16 hi + bf16 lo (x ~ hi + lo to 2^-17); the product is
//                 hi*hi + hi*lo + lo*hi, three MMAs into the same fp32 accumulator: relative error ~2^-16 per
//                 product, which keeps the fused features within 1e-4 of the fp32 oracle.
//
// Weights are pre-packed once per call (k_pack_weights) into the shared-memory operand image, chunked along K;
// when a whole layer fits they stay resident in shared memory for the life of the CTA, otherwise (C >= 192 in
// fp32 mode, C = 256) chunks of 64 input channels are streamed from L2 per use.
#include "cf_common.cuh"
#include "cf_tcgen05.cuh"

namespace cf {

namespace {

constexpr int kTile = 128;  // cells per tile == UMMA M
constexpr int kThreads = 128;

struct TcParams {
    const float *bev;
    const float *T;
    const int32_t *knn;
    float *out;
    const uint8_t *wimg2;  // packed W2 image (chunked)
    const uint8_t *wimg3;
    const float *W1;       // for the offset columns
    const float *b2;
    const float *b3;
    int32_t B, N, H, W, K, Ci;
    float x0, y0, dx, dy;
    int64_t tiles_per_frame, tiles_total;
    const int32_t *cell_list;   // (B, cells) compacted indices of the cells that have at least one neighbour
    const int32_t *cell_count;  // (B)
};

__host__ __device__ constexpr int kc_for(int C, int NS)
{
    // resident when both layers' packed weights + the A tile fit in shared memory, else stream 64-wide chunks
    return (NS == 1 ? (C <= 192) : (C <= 128)) ? C : 64;
}

// neighbour slots processed per barrier round (each has its own A tile and TMEM accumulator); > 1 only where the
// A tiles and the resident weights still leave room for several CTAs per SM
// launch shape of the two small-C instantiations (measured on B200, BASELINE configs[1]: several small CTAs per SM --
// more independent tiles in flight -- beat fewer barrier rounds per tile; see profiles/README.md)
#ifndef CF_ABUILD_UNROLL
#define CF_ABUILD_UNROLL 2
#endif
constexpr int kAUnroll = CF_ABUILD_UNROLL;  // A-tile gather items in flight per warp
#ifndef CF_SB32
#define CF_SB32 1
#define CF_SB64 1
#define CF_G32 1
#define CF_G64 2
#define CF_EW32 16
#define CF_EW64 16
#define CF_MB32 6
#define CF_MB64 3
#endif
__host__ __device__ constexpr int slots_per_round(int C, int NS) { return C <= 32 ? CF_SB32 : (C <= 64 ? CF_SB64 : (NS == 1 && C <= 128 ? 2 : 1)); }

__host__ __device__ constexpr int tmem_cols_for(int cols)
{
    return cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
}

template <int C, int NS>
struct TcLayout {
    static constexpr int KC = kc_for(C, NS);
    static constexpr bool kResident = KC == C;
    static constexpr int kChunks = C / KC;
    static_assert(C % KC == 0 && KC % 32 == 0, "channel count must be a multiple of the K-chunk");
    static constexpr int kWChunkBytes = NS * C * KC * 2;            // one K-chunk of one layer, all splits
    static constexpr int kWBytes = kResident ? 2 * kWChunkBytes : kWChunkBytes;
    static constexpr int SB = kResident ? slots_per_round(C, NS) : 1;
    static constexpr int kASlotBytes = NS * kTile * KC * 2;         // one A tile (all splits)
    static constexpr int kABytes = SB * kASlotBytes;
    static constexpr int kOffA = kWBytes;
    static constexpr int kOffF = kOffA + kABytes;                   // floats: b2, b3, w1x, w1y
    static constexpr int kOffIdx = kOffF + 4 * C * 4;               // int32 [128][CF_MAX_K]
    static constexpr int kOffCtr = kOffIdx + kTile * CF_MAX_K * 4;  // float cx[128], cy[128]
    static constexpr int kOffCell = kOffCtr + 2 * kTile * 4;        // int32 cell index of each row
    static constexpr int kOffBar = kOffCell + kTile * 4;            // mbarrier (8 B) + tmem ptr (4 B)
    static constexpr int kSmemBytes = kOffBar + 16;
    static constexpr int kTmemCols = tmem_cols_for((SB + 1) * C);   // SB accumulators + the pooled sum
};

// ---------------------------------------------------------------------------------------------------------------
// fp32 (rows, cols) weights with row stride ld -> packed operand image:
//   [chunk][split][n/8][k/8 in chunk][n%8][k%8] bf16, chunk = KC consecutive input channels
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack_weights(const float *__restrict__ W, int32_t rows, int32_t cols, int32_t ld,
                                                      int32_t KC, int32_t NS, uint8_t *__restrict__ img)
{
    const int32_t units = rows * (cols / 8);  // one 16-byte unit = 8 consecutive k of one n
    const int32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= units) return;
    const int32_t n = u / (cols / 8), k8 = u - n * (cols / 8);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = W[(size_t)n * ld + k8 * 8 + i];
    uint4 hi, lo;
    tc::split_bf16x8(v, hi, lo, NS == 2);
    const int32_t kc_units = KC / 8;
    const int32_t chunk = k8 / kc_units, ku = k8 - chunk * kc_units;
    const size_t chunk_bytes = (size_t)rows * KC * 2;
    const size_t base = (size_t)chunk * NS * chunk_bytes + tc::unit_offset(n, ku, kc_units);
    *reinterpret_cast<uint4 *>(img + base) = hi;
    if (NS == 2) *reinterpret_cast<uint4 *>(img + base + chunk_bytes) = lo;
}

// linear copy of a packed weight chunk (global/L2 -> shared), 16 bytes per thread per step
template <int NT>
__device__ __forceinline__ void copy_chunk(uint8_t *dst, const uint8_t *__restrict__ src, int bytes)
{
    for (int o = threadIdx.x * 16; o < bytes; o += NT * 16)
        *reinterpret_cast<uint4 *>(dst + o) = __ldg(reinterpret_cast<const uint4 *>(src + o));
}

// issue the MMAs of one K-chunk: acc (+)= A[128 x KC] * Wchunk[C x KC]^T, all split products
template <int C, int NS, int KC>
__device__ __forceinline__ void issue_chunk(uint32_t a_addr, uint32_t w_addr, uint32_t tmem_acc, bool accumulate)
{
    constexpr uint32_t idesc = tc::make_idesc_bf16(kTile, C);
    constexpr uint32_t sbo = (KC / 8) * 128, lbo = 128;
    constexpr uint32_t a_split = kTile * KC * 2, w_split = C * KC * 2;
    uint32_t acc = accumulate ? 1u : 0u;
#pragma unroll
    for (int kk = 0; kk < KC / 16; ++kk) {
        const uint32_t koff = kk * 2 * lbo;  // 16 bf16 = two 16-byte k-units
        const uint64_t a_hi = tc::make_desc(a_addr + koff, lbo, sbo);
        const uint64_t w_hi = tc::make_desc(w_addr + koff, lbo, sbo);
        tc::mma_bf16(tmem_acc, a_hi, w_hi, idesc, acc);
        acc = 1u;
        if (NS == 2) {
            const uint64_t a_lo = tc::make_desc(a_addr + a_split + koff, lbo, sbo);
            const uint64_t w_lo = tc::make_desc(w_addr + w_split + koff, lbo, sbo);
            tc::mma_bf16(tmem_acc, a_hi, w_lo, idesc, 1u);
            tc::mma_bf16(tmem_acc, a_lo, w_hi, idesc, 1u);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Cell compaction.  Only ~1/3 of the BEV cells of a LiDAR frame have a point within the radius.  This pass
//   * appends the index of every cell with a neighbour to a per-frame list (block-local order, one atomicAdd per
//     block, so the list is a concatenation of ascending runs: neighbouring list entries are neighbouring cells and
//     the fused kernel's BEV accesses stay coalesced), and
//   * finishes the cells WITHOUT a neighbour right here: out = bev (skipped when the layer runs in place).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cell_compact(const int32_t *__restrict__ knn, int32_t K, int64_t cells, int32_t C,
                                                      const float *__restrict__ bev, float *__restrict__ out,
                                                      int32_t *__restrict__ list, int32_t *__restrict__ count)
{
    __shared__ int32_t warp_excl[8];
    __shared__ int32_t base;
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t cell = (int64_t)blockIdx.x * 256 + tid;
    const bool inside = cell < cells;
    const bool live = inside && __ldg(knn + ((size_t)b * cells + cell) * K) >= 0;
    const unsigned bal = __ballot_sync(0xffffffffu, live);
    if (lane == 0) warp_excl[warp] = __popc(bal);
    __syncthreads();
    if (tid == 0) {
        int32_t run = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const int32_t c = warp_excl[w];
            warp_excl[w] = run;
            run += c;
        }
        base = run ? atomicAdd(count + b, run) : 0;
    }
    __syncthreads();
    if (live) list[(size_t)b * cells + base + warp_excl[warp] + __popc(bal & ((1u << lane) - 1u))] = (int32_t)cell;
    if (inside && !live && out != bev) {
        const size_t o = (size_t)b * C * cells + cell;
#pragma unroll 8
        for (int c = 0; c < C; ++c) out[o + (size_t)c * cells] = __ldg(bev + o + (size_t)c * cells);
    }
}

// Thread layout: G groups of 128 threads.  Thread (row = tid % 128, grp = tid / 128) owns BEV cell `row` of the
// tile; the G threads of a row split the 16-byte operand units of the A tile and the EW-column chunks of the
// epilogues between them (warp w may only touch TMEM lanes 32*(w%4)..+31, which is exactly its rows).
template <int C>
struct TcShape {
    static constexpr int G = C <= 32 ? CF_G32 : C <= 64 ? CF_G64 : C == 96 || C == 192 ? 3 : 4;
    static constexpr int EW = C <= 32 ? CF_EW32 : C <= 64 ? CF_EW64 : 32;
    static constexpr int kMinBlocks = C <= 32 ? CF_MB32 : C <= 64 ? CF_MB64 : 1;
};

template <int C, int NS>
__global__ void __launch_bounds__(kTile * TcShape<C>::G, TcShape<C>::kMinBlocks) k_fusion_tc(const TcParams p)
{
    using L = TcLayout<C, NS>;
    constexpr int G = TcShape<C>::G, EW = TcShape<C>::EW;
    constexpr int NT = kTile * G;
    constexpr int KC = L::KC;
    constexpr int kc_units = KC / 8;
    constexpr int kChunksE = C / EW;  // epilogue chunks over all C columns
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sW = smem;
    uint8_t *sA = smem + L::kOffA;
    float *sb2 = reinterpret_cast<float *>(smem + L::kOffF);
    float *sb3 = sb2 + C;
    float *sw1x = sb3 + C;
    float *sw1y = sw1x + C;
    int32_t *sidx = reinterpret_cast<int32_t *>(smem + L::kOffIdx);
    float *scx = reinterpret_cast<float *>(smem + L::kOffCtr);
    float *scy = scx + kTile;
    int32_t *scell = reinterpret_cast<int32_t *>(smem + L::kOffCell);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + L::kOffBar);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::kOffBar + 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = tid & (kTile - 1), grp = tid / kTile;
    constexpr int NW = NT / 32;
    const int K = p.K;
    const int64_t cells = (int64_t)p.H * p.W;

    // ---- one-time setup -----------------------------------------------------------------------------------------
    if (tid == 0) {
        tc::mbar_init(bar, 1);
        tc::mbar_fence_init();
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(tmem_slot, L::kTmemCols);
    for (int c = tid; c < C; c += NT) {
        sb2[c] = __ldg(p.b2 + c);
        sb3[c] = __ldg(p.b3 + c);
        sw1x[c] = __ldg(p.W1 + (size_t)c * (p.Ci + 3) + p.Ci);
        sw1y[c] = __ldg(p.W1 + (size_t)c * (p.Ci + 3) + p.Ci + 1);
    }
    if (L::kResident) {
        copy_chunk<NT>(sW, p.wimg2, L::kWChunkBytes);
        copy_chunk<NT>(sW + L::kWChunkBytes, p.wimg3, L::kWChunkBytes);
        tc::fence_proxy_async();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    constexpr int SB = L::SB;
    const uint32_t tmem_acc = tmem_base;                                // SB accumulators: columns [s*C, (s+1)*C)
    const uint32_t tmem_pool = tmem_base + SB * C;                      // pooled sum: columns [SB*C, (SB+1)*C)
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;        // this warp's TMEM lanes == its rows
    const uint32_t sA_addr = tc::smem_u32(sA), sW_addr = tc::smem_u32(sW);
    uint32_t phase = 0;

    for (int64_t tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
        // tiles run over the COMPACTED list of cells that have a neighbour (k_cell_compact); cells without one were
        // already copied bev -> out by that pass, so every row built below is (almost always) useful work
        const int b = (int)(tile / p.tiles_per_frame);
        const int64_t e0 = (tile - (int64_t)b * p.tiles_per_frame) * kTile;
        const int32_t n_live = p.cell_list ? __ldg(p.cell_count + b) : (int32_t)cells;  // no list: every cell, in order
        if (e0 >= n_live) continue;  // uniform
        const bool in_range = e0 + row < n_live;
        const int32_t cell = !in_range ? 0 : p.cell_list ? __ldg(p.cell_list + (size_t)b * cells + e0 + row) : (int32_t)(e0 + row);
        if (tid < kTile) {
            float cx = 0.f, cy = 0.f;
            if (in_range) {
                const int32_t i = cell / p.W, j = cell - i * p.W;
                cx = __fadd_rn(p.x0, __fmul_rn((float)i, p.dx));
                cy = __fadd_rn(p.y0, __fmul_rn((float)j, p.dy));
            }
            scx[row] = cx;
            scy[row] = cy;
            scell[row] = in_range ? cell : -1;
        }
        __syncthreads();
        // neighbour indices of the tile's cells (K consecutive ints per cell; list entries are mostly consecutive cells)
        for (int i = tid; i < kTile * K; i += NT) {
            const int r = i / K, kk = i - r * K;
            const int32_t cr = scell[r];
            sidx[i] = cr >= 0 ? __ldg(p.knn + ((size_t)b * cells + cr) * K + kk) : -1;
        }
        __syncthreads();
        const float *Tb = p.T + (size_t)b * p.N * C;
        int n_valid = 0;
        for (int k = 0; k < K; ++k) n_valid += sidx[row * K + k] >= 0;
        bool pooled_live = false;  // uniform across the CTA

        // neighbour slots in rounds of SB: build SB A tiles, ONE barrier, SB MMA groups, ONE commit / wait, one epilogue
        for (int k0 = 0; k0 < K; k0 += SB) {
            const int nb = min(SB, K - k0);
            // slots are sorted (empty ones form a suffix): if nobody has a k0-th neighbour, nobody has a later one
            if (!__syncthreads_or(sidx[row * K + k0] >= 0)) break;

            for (int ch = 0; ch < L::kChunks; ++ch) {
                if (!L::kResident) copy_chunk<NT>(sW, p.wimg2 + (size_t)ch * L::kWChunkBytes, L::kWChunkBytes);
                // A chunk [128 rows x KC] per slot: a warp takes one 8-row group x 4 operand units (32 channels) per
                // step: lane (r8 = lane%8, u = lane/8).  Global side: the 4 lanes of a row read one contiguous
                // 128-byte segment of the point's T row; shared side: each quarter-warp (the unit of a 128-bit shared
                // access) writes the 8 rows of one unit = 128 contiguous bytes of the operand image: conflict free.
                // (lane/4, lane%4 would put 4 units x 128 B apart in one quarter-warp: a 4-way bank conflict.)
                // NW is a multiple of the unit-quads per row, so a warp always works on the same 32 channels: its 16
                // offset-weight values live in registers for the whole chunk (no shared-memory traffic per item)
                constexpr int kQuads = kc_units / 4;
                static_assert(NW % kQuads == 0, "warps per CTA must be a multiple of the unit quads per row");
                const int uq = warp % kQuads, ku = uq * 4 + (lane >> 3);
                const int c0 = ch * KC + ku * 8;
                // negated once, so that  T - (w1x cx + w1y cy)  is two FFMAs per channel
                float nx[8], ny[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    nx[i] = -sw1x[c0 + i];
                    ny[i] = -sw1y[c0 + i];
                }
#pragma unroll kAUnroll
                for (int it = warp / kQuads; it < nb * 16; it += NW / kQuads) {
                    const int sl = it >> 4, rg = it & 15;
                    const int r = rg * 8 + (lane & 7);
                    const int32_t pr = sidx[r * K + k0 + sl];
                    const bool ok = pr >= 0;
                    const float cx = scx[r], cy = scy[r];
                    const float4 *trow = reinterpret_cast<const float4 *>(Tb + (size_t)(ok ? pr : 0) * C + c0);
                    const float4 t0 = __ldg(trow), t1 = __ldg(trow + 1);
                    const float t[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = fmaxf(fmaf(nx[i], cx, fmaf(ny[i], cy, t[i])), 0.0f);
                    if (!__all_sync(0xffffffffu, ok)) {  // rare after compaction: rows without a k-th neighbour
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = ok ? v[i] : 0.0f;
                    }
                    uint4 hi, lo;
                    tc::split_bf16x8(v, hi, lo, NS == 2);
                    uint8_t *dst = sA + sl * L::kASlotBytes + tc::unit_offset(r, ku, kc_units);
                    *reinterpret_cast<uint4 *>(dst) = hi;
                    if (NS == 2) *reinterpret_cast<uint4 *>(dst + kTile * KC * 2) = lo;
                }
                tc::fence_proxy_async();
                tc::fence_before_sync();
                __syncthreads();
                if (tid == 0) {
                    tc::fence_after_sync();
                    for (int sl = 0; sl < nb; ++sl)
                        issue_chunk<C, NS, KC>(sA_addr + sl * L::kASlotBytes, sW_addr, tmem_acc + sl * C, ch > 0);
                    tc::commit(bar);
                }
                tc::mbar_wait(bar, phase);
                phase ^= 1u;
                tc::fence_after_sync();
            }
            // ---- epilogue of the round: pooled (+)= sum over its slots of valid ? relu(acc + b2) : 0 ------------------
            __syncwarp();
#pragma unroll 1
            for (int cc = grp; cc < kChunksE; cc += G) {
                float s[EW];
                if (pooled_live) {
                    tc::tmem_ld<EW>(tmem_pool + lane_off + cc * EW, s);
                } else {
#pragma unroll
                    for (int i = 0; i < EW; ++i) s[i] = 0.0f;
                }
                for (int sl = 0; sl < nb; ++sl) {
                    const bool valid = sidx[row * K + k0 + sl] >= 0;
                    float z[EW];
                    tc::tmem_ld<EW>(tmem_acc + sl * C + lane_off + cc * EW, z);
#pragma unroll
                    for (int i4 = 0; i4 < EW / 4; ++i4) {
                        const float4 bb = *reinterpret_cast<const float4 *>(sb2 + cc * EW + i4 * 4);
                        const float bq[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) z[i4 * 4 + i] = fmaxf(z[i4 * 4 + i] + bq[i], 0.0f);
                    }
                    if (__all_sync(0xffffffffu, valid)) {  // the common case after compaction: no mask needed
#pragma unroll
                        for (int i = 0; i < EW; ++i) s[i] += z[i];
                    } else {
#pragma unroll
                        for (int i = 0; i < EW; ++i) s[i] += valid ? z[i] : 0.0f;
                    }
                }
                tc::tmem_st<EW>(tmem_pool + lane_off + cc * EW, s);
            }
            pooled_live = true;
            tc::fence_before_sync();  // TMEM accesses above are ordered before the next MMA by the next barrier
        }

        if (pooled_live) {
            // ---- layer 3: acc = pooled * W3^T --------------------------------------------------------------------------
            for (int ch = 0; ch < L::kChunks; ++ch) {
                if (!L::kResident) copy_chunk<NT>(sW, p.wimg3 + (size_t)ch * L::kWChunkBytes, L::kWChunkBytes);
                __syncwarp();
#pragma unroll 1
                for (int cc = grp; cc < KC / EW; cc += G) {
                    float s[EW];
                    tc::tmem_ld<EW>(tmem_pool + lane_off + ch * KC + cc * EW, s);
#pragma unroll
                    for (int q = 0; q < EW / 8; ++q) {
                        float v[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = s[q * 8 + i];
                        uint4 hi, lo;
                        tc::split_bf16x8(v, hi, lo, NS == 2);
                        const uint32_t off = tc::unit_offset(row, cc * (EW / 8) + q, kc_units);
                        *reinterpret_cast<uint4 *>(sA + off) = hi;
                        if (NS == 2) *reinterpret_cast<uint4 *>(sA + kTile * KC * 2 + off) = lo;
                    }
                }
                tc::fence_proxy_async();
                tc::fence_before_sync();
                __syncthreads();
                if (tid == 0) {
                    tc::fence_after_sync();
                    issue_chunk<C, NS, KC>(sA_addr, sW_addr + (L::kResident ? L::kWChunkBytes : 0), tmem_acc, ch > 0);
                    tc::commit(bar);
                }
                tc::mbar_wait(bar, phase);
                phase ^= 1u;
                tc::fence_after_sync();
            }
        }
        // ---- final epilogue: out = bev + acc + n_valid * b3 (thread = cell: a warp touches 128 contiguous bytes) --
        const float nv = (float)n_valid;
        __syncwarp();
#pragma unroll 1
        for (int cc = grp; cc < kChunksE; cc += G) {
            float z[EW];
            if (pooled_live) tc::tmem_ld<EW>(tmem_acc + lane_off + cc * EW, z);
            if (in_range) {
                const size_t base = ((size_t)b * C + cc * EW) * cells + cell;
                float bv[EW];
#pragma unroll
                for (int i = 0; i < EW; ++i) bv[i] = __ldg(p.bev + base + (size_t)i * cells);
#pragma unroll
                for (int i = 0; i < EW; ++i) {
                    const float add = pooled_live ? z[i] + nv * sb3[cc * EW + i] : 0.0f;
                    p.out[base + (size_t)i * cells] = bv[i] + add;
                }
            }
        }
        tc::fence_before_sync();
        __syncthreads();  // sidx / A / TMEM are reused by the next tile
        tc::fence_after_sync();
    }

    __syncthreads();
    if (warp == 0) tc::tmem_free(tmem_base, L::kTmemCols);
}

// ---------------------------------------------------------------------------------------------------------------
// K-4a on tensor cores:  T[b, m, :] = feat[b, m, :] W1[:, :Ci]^T + W1[:, Ci:Ci+3] p_m + b1      (m < num_points[b])
// Tile = 128 points.  W1's image part stays resident in shared memory; the rank-3 offset term and the bias are
// added on CUDA cores in the epilogue (Ci+3 = 131 is not an MMA-friendly K).
// ---------------------------------------------------------------------------------------------------------------
struct Mlp1Params {
    const float *feat;
    const float *points;
    const int64_t *num_points;
    const uint8_t *wimg;
    const float *W1;
    const float *b1;
    float *T;
    int32_t B, N, Ci;
    int32_t tiles_per_frame;
};

template <int C>
struct Mlp1Shape {
    static constexpr int G = C <= 32 ? 1 : C <= 64 ? 2 : C == 96 || C == 192 ? 3 : 4;  // 128-thread groups per CTA
};

template <int C, int NS>
__global__ void __launch_bounds__(kTile * Mlp1Shape<C>::G) k_point_mlp1_tc(const Mlp1Params p)
{
    constexpr int G = Mlp1Shape<C>::G, NT = kTile * G, NW = NT / 32;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int Ci = p.Ci, kc_units = Ci / 8;
    const int w_bytes = NS * C * Ci * 2, a_bytes = NS * kTile * Ci * 2;
    uint8_t *sW = smem;
    uint8_t *sA = smem + w_bytes;
    float *sb1 = reinterpret_cast<float *>(smem + w_bytes + a_bytes);
    float *swx = sb1 + C, *swy = swx + C, *swz = swy + C;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = tid & (kTile - 1), grp = tid / kTile;
    constexpr int kCols = C <= 32 ? 32 : C <= 64 ? 64 : C <= 128 ? 128 : 256;

    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::mbar_fence_init();
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(&tmem_slot, kCols);
    for (int c = tid; c < C; c += NT) {
        const float *w = p.W1 + (size_t)c * (Ci + 3) + Ci;
        sb1[c] = __ldg(p.b1 + c);
        swx[c] = __ldg(w);
        swy[c] = __ldg(w + 1);
        swz[c] = __ldg(w + 2);
    }
    copy_chunk<NT>(sW, p.wimg, w_bytes);
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_acc = tmem_slot;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t sA_addr = tc::smem_u32(sA), sW_addr = tc::smem_u32(sW);
    const uint32_t idesc = tc::make_idesc_bf16(kTile, C);
    const uint32_t sbo = kc_units * 128, lbo = 128;
    uint32_t phase = 0;

    const int64_t tiles_total = (int64_t)p.tiles_per_frame * p.B;
    for (int64_t tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
        const int b = (int)(tile / p.tiles_per_frame);
        const int32_t m0 = (int32_t)(tile - (int64_t)b * p.tiles_per_frame) * kTile;
        const int32_t n_pts = valid_points(p.num_points, b, p.N);
        if (m0 >= n_pts) continue;  // uniform across the CTA
        // A tile: a warp takes one 8-row group x 4 operand units per step, lane (r8 = lane%8, u = lane/8): coalesced
        // 128-byte row segments in, and each quarter-warp writes 128 contiguous bytes of the operand image (conflict free)
        const float *fb = p.feat + ((size_t)b * p.N + m0) * Ci;
        for (int item = warp; item < 16 * (kc_units / 4); item += NW) {
            const int rg = item / (kc_units / 4), uq = item - rg * (kc_units / 4);
            const int r = rg * 8 + (lane & 7), ku = uq * 4 + (lane >> 3);
            float v[8];
            if (m0 + r < n_pts) {
                const float4 *src = reinterpret_cast<const float4 *>(fb + (size_t)r * Ci + ku * 8);
                const float4 t0 = __ldg(src), t1 = __ldg(src + 1);
                v[0] = t0.x; v[1] = t0.y; v[2] = t0.z; v[3] = t0.w; v[4] = t1.x; v[5] = t1.y; v[6] = t1.z; v[7] = t1.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = 0.0f;
            }
            uint4 hi, lo;
            tc::split_bf16x8(v, hi, lo, NS == 2);
            const uint32_t off = tc::unit_offset(r, ku, kc_units);
            *reinterpret_cast<uint4 *>(sA + off) = hi;
            if (NS == 2) *reinterpret_cast<uint4 *>(sA + kTile * Ci * 2 + off) = lo;
        }
        tc::fence_proxy_async();
        tc::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            uint32_t acc = 0;
            for (int kk = 0; kk < Ci / 16; ++kk) {
                const uint32_t koff = kk * 2 * lbo;
                const uint64_t a_hi = tc::make_desc(sA_addr + koff, lbo, sbo), w_hi = tc::make_desc(sW_addr + koff, lbo, sbo);
                tc::mma_bf16(tmem_acc, a_hi, w_hi, idesc, acc);
                acc = 1;
                if (NS == 2) {
                    const uint64_t a_lo = tc::make_desc(sA_addr + kTile * Ci * 2 + koff, lbo, sbo);
                    const uint64_t w_lo = tc::make_desc(sW_addr + C * Ci * 2 + koff, lbo, sbo);
                    tc::mma_bf16(tmem_acc, a_hi, w_lo, idesc, 1);
                    tc::mma_bf16(tmem_acc, a_lo, w_hi, idesc, 1);
                }
            }
            tc::commit(&bar);
        }
        tc::mbar_wait(&bar, phase);
        phase ^= 1u;
        tc::fence_after_sync();
        const int32_t m = m0 + row;
        const bool live = m < n_pts;
        float px = 0.f, py = 0.f, pz = 0.f;
        if (live) {
            const float *q = p.points + ((size_t)b * p.N + m) * 3;
            px = __ldg(q); py = __ldg(q + 1); pz = __ldg(q + 2);
        }
        __syncwarp();
#pragma unroll 1
        for (int cc = grp; cc < C / 32; cc += G) {
            float z[32];
            tc::tmem_ld32(tmem_acc + lane_off + cc * 32, z);
            if (live) {
                float4 *dst = reinterpret_cast<float4 *>(p.T + ((size_t)b * p.N + m) * C + cc * 32);
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) {
                    float o[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int c = cc * 32 + q4 * 4 + i;
                        o[i] = z[q4 * 4 + i] + (swx[c] * px + swy[c] * py + swz[c] * pz) + sb1[c];
                    }
                    dst[q4] = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
        }
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    }
    __syncthreads();
    if (warp == 0) tc::tmem_free(tmem_acc, kCols);
}

template <int C, int NS>
int launch_mlp1_tc(const Mlp1Params &p, cudaStream_t st)
{
    constexpr int NT = kTile * Mlp1Shape<C>::G;
    const int smem = NS * (C + kTile) * p.Ci * 2 + 4 * C * 4;
    CF_TRY(cuda_status(cudaFuncSetAttribute(k_point_mlp1_tc<C, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                       "k_point_mlp1_tc smem attribute"));
    constexpr int kCols = C <= 32 ? 32 : C <= 64 ? 64 : C <= 128 ? 128 : 256;
    const int per_sm = std::max(1, std::min(std::min((227 * 1024) / (smem + 1024), 512 / kCols), std::min(2048 / NT, 4)));
    const int64_t tiles = (int64_t)p.tiles_per_frame * p.B;
    const int64_t grid = std::min<int64_t>(tiles, (int64_t)sm_count() * per_sm);
    k_point_mlp1_tc<C, NS><<<(unsigned)grid, NT, smem, st>>>(p);
    return CF_OK;
}

// a standalone D[128 x N] = A[128 x Kd] * B[N x Kd]^T through exactly the same packing / descriptor / TMEM code,
// so operand-layout mistakes can be told apart from fusion-logic mistakes (cf_debug_umma_gemm).
template <int N, int NS>
__global__ void __launch_bounds__(kThreads) k_umma_selftest(const float *__restrict__ A, const float *__restrict__ Bm,
                                                            int32_t Kd, float *__restrict__ D)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int kc_units = Kd / 8;
    uint8_t *sA = smem;
    uint8_t *sB = smem + NS * kTile * Kd * 2;
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::mbar_fence_init();
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(&tmem_slot, N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : 256);
    for (int ku = 0; ku < kc_units; ++ku) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = A[(size_t)tid * Kd + ku * 8 + i];
        uint4 hi, lo;
        tc::split_bf16x8(v, hi, lo, NS == 2);
        *reinterpret_cast<uint4 *>(sA + tc::unit_offset(tid, ku, kc_units)) = hi;
        if (NS == 2) *reinterpret_cast<uint4 *>(sA + kTile * Kd * 2 + tc::unit_offset(tid, ku, kc_units)) = lo;
    }
    for (int u = tid; u < N * kc_units; u += kThreads) {
        const int n = u / kc_units, ku = u - n * kc_units;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = Bm[(size_t)n * Kd + ku * 8 + i];
        uint4 hi, lo;
        tc::split_bf16x8(v, hi, lo, NS == 2);
        *reinterpret_cast<uint4 *>(sB + tc::unit_offset(n, ku, kc_units)) = hi;
        if (NS == 2) *reinterpret_cast<uint4 *>(sB + N * Kd * 2 + tc::unit_offset(n, ku, kc_units)) = lo;
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_bf16(kTile, N);
        const uint32_t sbo = kc_units * 128, lbo = 128;
        const uint32_t a0 = tc::smem_u32(sA), b0 = tc::smem_u32(sB);
        uint32_t acc = 0;
        for (int kk = 0; kk < Kd / 16; ++kk) {
            const uint32_t koff = kk * 2 * lbo;
            const uint64_t a_hi = tc::make_desc(a0 + koff, lbo, sbo), b_hi = tc::make_desc(b0 + koff, lbo, sbo);
            tc::mma_bf16(tmem_base, a_hi, b_hi, idesc, acc);
            acc = 1;
            if (NS == 2) {
                const uint64_t a_lo = tc::make_desc(a0 + kTile * Kd * 2 + koff, lbo, sbo);
                const uint64_t b_lo = tc::make_desc(b0 + N * Kd * 2 + koff, lbo, sbo);
                tc::mma_bf16(tmem_base, a_hi, b_lo, idesc, 1);
                tc::mma_bf16(tmem_base, a_lo, b_hi, idesc, 1);
            }
        }
        tc::commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    for (int cc = 0; cc < N / 32; ++cc) {
        float z[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + cc * 32, z);
#pragma unroll
        for (int i = 0; i < 32; ++i) D[(size_t)tid * N + cc * 32 + i] = z[i];
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_free(tmem_base, N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : 256);
}

template <int C, int NS>
int launch_tc(const TcParams &p, cudaStream_t st)
{
    using L = TcLayout<C, NS>;
    constexpr int NT = kTile * TcShape<C>::G;
    static bool attr_set = false;
    if (!attr_set) {
        CF_TRY(cuda_status(cudaFuncSetAttribute(k_fusion_tc<C, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                L::kSmemBytes),
                           "k_fusion_tc smem attribute"));
        attr_set = true;
    }
    // CTAs per SM: limited by shared memory, TMEM columns (512 per SM, never oversubscribed) and threads
    const int by_smem = (227 * 1024) / (L::kSmemBytes + 1024);
    const int by_tmem = 512 / L::kTmemCols;  // (SB + 1) * C columns per CTA, rounded up to a power of two
    const int by_threads = 2048 / NT;
    const int per_sm = std::max(1, std::min(std::min(by_smem, by_tmem), std::min(by_threads, 8)));
    const int64_t grid = std::min<int64_t>(p.tiles_total, (int64_t)sm_count() * per_sm);
    k_fusion_tc<C, NS><<<(unsigned)grid, NT, L::kSmemBytes, st>>>(p);
    return CF_OK;
}

}  // namespace

static size_t tc_weight_bytes(int32_t C, int NS) { return ((size_t)2 * NS * C * C * 2 + 255) / 256 * 256; }

size_t fusion_tc_packed_bytes(int32_t C, int32_t mode) { return tc_weight_bytes(C, mode == CF_MODE_FP32 ? 2 : 1); }

int fusion_tc_pack(const float *d_W2, const float *d_W3, int32_t C, int32_t mode, void *d_packed, cudaStream_t st)
{
    CF_REQUIRE(C % 32 == 0 && C >= 32 && C <= 256, CF_ERR_UNSUPPORTED, "cf_fusion_pack_weights: C=%d", C);
    const int NS = mode == CF_MODE_FP32 ? 2 : 1;
    const int KC = kc_for(C, NS);
    const int pack_blocks = (C * (C / 8) + 255) / 256;
    k_pack_weights<<<pack_blocks, 256, 0, st>>>(d_W2, C, C, C, KC, NS, (uint8_t *)d_packed);
    k_pack_weights<<<pack_blocks, 256, 0, st>>>(d_W3, C, C, C, KC, NS, (uint8_t *)d_packed + (size_t)NS * C * C * 2);
    count_launches(2);
    return launch_status("cf_fusion_pack_weights");
}

size_t fusion_tc_workspace_bytes(int32_t C, int32_t mode, int32_t B, int32_t H, int32_t W)
{
    const int NS = mode == CF_MODE_FP32 ? 2 : 1;
    // packed W2/W3 images | per-frame live-cell counts | per-frame live-cell lists
    return tc_weight_bytes(C, NS) + 256 + (size_t)B * H * W * sizeof(int32_t);
}

int fusion_tc(const float *d_bev, const float *d_T, const int32_t *d_knn, int32_t B, int32_t N, int32_t C, int32_t H,
              int32_t W, int32_t K, float x0, float y0, float dx, float dy, const float *d_W1, int32_t Ci,
              const float *d_W2, const float *d_b2, const float *d_W3, const float *d_b3, float *d_out, int32_t mode,
              const void *d_packed, void *d_workspace, cudaStream_t st)
{
    CF_REQUIRE(C % 32 == 0, CF_ERR_UNSUPPORTED, "cf_fusion_fwd: tensor-core path needs C %% 32 == 0 (C=%d)", C);
    const int NS = mode == CF_MODE_FP32 ? 2 : 1;
    const int KC = kc_for(C, NS);
    // packed W2 | W3 operand images: the caller's cached copy (cf_fusion_pack_weights) or packed here
    const uint8_t *img2 = d_packed ? (const uint8_t *)d_packed : (const uint8_t *)d_workspace;
    const uint8_t *img3 = img2 + (size_t)NS * C * C * 2;
    int32_t *cell_count = (int32_t *)((uint8_t *)d_workspace + tc_weight_bytes(C, NS));
    int32_t *cell_list = cell_count + 64;
    CF_REQUIRE(B <= 64, CF_ERR_ARG, "cf_fusion_fwd: batch %d > 64 frames per call", B);
    const int64_t n_cells = (int64_t)H * W;
    // Compaction pays when there are many more tiles than SMs; on the small coarse scales it would only reduce the
    // number of CTAs that have work, so those run every tile (empty tiles just copy bev -> out).
    const bool compact = ceil_div64(n_cells, kTile) * B >= 4 * (int64_t)sm_count();
    if (compact) {
        CF_TRY(cuda_status(cudaMemsetAsync(cell_count, 0, 256, st), "cf_fusion_fwd memset"));
        k_cell_compact<<<dim3((unsigned)ceil_div64(n_cells, 256), (unsigned)B), 256, 0, st>>>(d_knn, K, n_cells, C, d_bev,
                                                                                            d_out, cell_list, cell_count);
        count_launches(1);
    }
    if (!d_packed) {
        const int pack_blocks = (C * (C / 8) + 255) / 256;
        k_pack_weights<<<pack_blocks, 256, 0, st>>>(d_W2, C, C, C, KC, NS, (uint8_t *)d_workspace);
        k_pack_weights<<<pack_blocks, 256, 0, st>>>(d_W3, C, C, C, KC, NS, (uint8_t *)d_workspace + (size_t)NS * C * C * 2);
        count_launches(2);
    }
    TcParams p;
    p.bev = d_bev; p.T = d_T; p.knn = d_knn; p.out = d_out; p.wimg2 = img2; p.wimg3 = img3; p.W1 = d_W1;
    p.b2 = d_b2; p.b3 = d_b3; p.B = B; p.N = N; p.H = H; p.W = W; p.K = K; p.Ci = Ci;
    p.x0 = x0; p.y0 = y0; p.dx = dx; p.dy = dy;
    p.tiles_per_frame = ceil_div64((int64_t)H * W, kTile);
    p.tiles_total = p.tiles_per_frame * B;
    p.cell_list = compact ? cell_list : nullptr;
    p.cell_count = cell_count;
    int rc = CF_ERR_UNSUPPORTED;
#define CF_TC_CASE(c)                                                          \
    case c:                                                                    \
        rc = NS == 2 ? launch_tc<c, 2>(p, st) : launch_tc<c, 1>(p, st);        \
        break;
    switch (C) {
        CF_TC_CASE(32) CF_TC_CASE(64) CF_TC_CASE(96) CF_TC_CASE(128) CF_TC_CASE(192) CF_TC_CASE(256)
        default:
            set_error("cf_fusion_fwd: unsupported C=%d on the tensor-core path", C);
            return CF_ERR_UNSUPPORTED;
    }
#undef CF_TC_CASE
    CF_TRY(rc);
    count_launches(1);
    return launch_status("cf_fusion_fwd (tcgen05)");
}


int point_mlp1_tc_pack(const float *d_W1, int32_t Ci, int32_t C, int32_t mode, void *d_packed, cudaStream_t st)
{
    CF_REQUIRE(Ci % 16 == 0 && Ci <= 256 && C % 32 == 0, CF_ERR_UNSUPPORTED, "cf_point_mlp1_pack_weights: Ci=%d C=%d", Ci, C);
    const int NS = mode == CF_MODE_FP32 ? 2 : 1;
    k_pack_weights<<<(C * (Ci / 8) + 255) / 256, 256, 0, st>>>(d_W1, C, Ci, Ci + 3, Ci, NS, (uint8_t *)d_packed);
    count_launches(1);
    return launch_status("cf_point_mlp1_pack_weights");
}

size_t point_mlp1_tc_workspace_bytes(int32_t Ci, int32_t C, int32_t mode)
{
    const int NS = mode == CF_MODE_FP32 ? 2 : 1;
    return (size_t)NS * C * Ci * 2 + 256;
}

// returns CF_ERR_UNSUPPORTED (without setting an error) when the shape has no tensor-core instantiation
int point_mlp1_tc(const float *d_feat, const float *d_points, const int64_t *d_num_points, int32_t B, int32_t N,
                  int32_t Ci, int32_t C, const float *d_W1, const float *d_b1, float *d_T, int32_t mode,
                  const void *d_packed, void *d_workspace, cudaStream_t st)
{
    const int NS = mode == CF_MODE_FP32 ? 2 : 1;
    if (Ci % 16 != 0 || Ci > 256 || C % 32 != 0) return CF_ERR_UNSUPPORTED;
    if ((size_t)NS * (C + kTile) * Ci * 2 + 4 * C * 4 > 220 * 1024) return CF_ERR_UNSUPPORTED;
    const uint8_t *img = d_packed ? (const uint8_t *)d_packed : (const uint8_t *)d_workspace;
    if (!d_packed) {
        k_pack_weights<<<(C * (Ci / 8) + 255) / 256, 256, 0, st>>>(d_W1, C, Ci, Ci + 3, Ci, NS, (uint8_t *)d_workspace);
        count_launches(1);
    }
    Mlp1Params p;
    p.feat = d_feat; p.points = d_points; p.num_points = d_num_points; p.wimg = img; p.W1 = d_W1; p.b1 = d_b1; p.T = d_T;
    p.B = B; p.N = N; p.Ci = Ci; p.tiles_per_frame = (N + kTile - 1) / kTile;
    int rc = CF_ERR_UNSUPPORTED;
#define CF_M1_CASE(c)                                                              \
    case c:                                                                        \
        rc = NS == 2 ? launch_mlp1_tc<c, 2>(p, st) : launch_mlp1_tc<c, 1>(p, st);  \
        break;
    switch (C) {
        CF_M1_CASE(32) CF_M1_CASE(64) CF_M1_CASE(96) CF_M1_CASE(128) CF_M1_CASE(192) CF_M1_CASE(256)
        default: return CF_ERR_UNSUPPORTED;
    }
#undef CF_M1_CASE
    CF_TRY(rc);
    count_launches(1);
    return launch_status("cf_point_mlp1 (tcgen05)");
}

int umma_selftest(const float *d_A, const float *d_B, int32_t N, int32_t Kd, int32_t split, float *d_D, cudaStream_t st)
{
    CF_REQUIRE(Kd % 16 == 0 && Kd >= 16 && Kd <= 256, CF_ERR_ARG, "cf_debug_umma_gemm: K=%d", Kd);
    const int NS = split ? 2 : 1;
    const size_t smem = (size_t)NS * (kTile + N) * Kd * 2;
    CF_REQUIRE(smem <= 200 * 1024, CF_ERR_ARG, "cf_debug_umma_gemm: operands do not fit in shared memory");
#define CF_ST_CASE(n)                                                                                                  \
    case n:                                                                                                            \
        if (NS == 2) {                                                                                                 \
            CF_TRY(cuda_status(cudaFuncSetAttribute(k_umma_selftest<n, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                    200 * 1024), "selftest attr"));                                   \
            k_umma_selftest<n, 2><<<1, kThreads, smem, st>>>(d_A, d_B, Kd, d_D);                                       \
        } else {                                                                                                       \
            CF_TRY(cuda_status(cudaFuncSetAttribute(k_umma_selftest<n, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                    200 * 1024), "selftest attr"));                                   \
            k_umma_selftest<n, 1><<<1, kThreads, smem, st>>>(d_A, d_B, Kd, d_D);                                       \
        }                                                                                                              \
        break;
    switch (N) {
        CF_ST_CASE(32) CF_ST_CASE(64) CF_ST_CASE(128) CF_ST_CASE(192) CF_ST_CASE(256)
        default:
            set_error("cf_debug_umma_gemm: N=%d not instantiated", N);
            return CF_ERR_ARG;
    }
#undef CF_ST_CASE
    count_launches(1);
    return launch_status("cf_debug_umma_gemm");
}

}  // namespace cf

// D (128,N) fp32 = A (128,K) * B (N,K)^T through the library's tcgen05 building blocks; split != 0 uses the
// bf16 hi/lo three-product scheme of CF_MODE_FP32.  Self-test entry point (tests/test_gpu_umma.py).
extern "C" int cf_debug_umma_gemm(const float *d_A, const float *d_B, int32_t N, int32_t K, int32_t split,
                                         float *d_D, void *stream)
{
    using namespace cf;
    CF_TRY(require_sm100());
    CF_REQUIRE(d_A && d_B && d_D, CF_ERR_ARG, "cf_debug_umma_gemm: null pointer");
    return umma_selftest(d_A, d_B, N, K, split, d_D, (cudaStream_t)stream);
}
