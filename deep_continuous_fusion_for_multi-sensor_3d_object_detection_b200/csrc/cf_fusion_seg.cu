// cf_fusion_seg.cu -- K-4 for the fine scales (C = 32 / 64): the fused MLP + K-sum-pool + BEV add on SEGMENT tiles with
// all neighbour slots of a tile going through the tensor pipe in ONE batch.
//
// Unit of work = a SEGMENT: 32 consecutive BEV cells (linear index), i.e. one 128-byte line of every channel plane.
// k_seg_compact splits the segments of a frame into those with at least one cell that has a neighbour (front of the list)
// and those without (back).  At BASELINE configs[1] 57 % of the 32-cell segments are entirely empty (90 % of the empty
// cells), 31 % entirely live, 12 % mixed.
//   * tile = 4 live segments = 128 rows = UMMA M; warp w of the epilogues owns segment w % 4 = TMEM lanes 32 (w % 4) .. +31,
//     so every BEV access of a warp is ONE aligned 128-byte line per channel (the cell-compacted kernel k_fusion_tc gathers
//     cells from a list: partial sectors, two lines per access).
//   * empty segments are copied bev -> out by whole warps with 128-bit accesses (8 lanes x 16 B per channel line), loads
//     issued before the layer-3 wait, stores after the final epilogue.
//   * the G neighbour slots of a batch (all K at K <= G) are built back to back into G operand buffers, every slot has its
//     own TMEM accumulator, the MMAs of all slots are issued under ONE barrier / commit / wait, and a single epilogue reads
//     the G accumulators, applies the ReLU and sums them in registers: the pooled sum never round-trips through TMEM and a
//     tile meets the tensor pipe twice (slots, layer 3) instead of K + 1 times.
//   * b2 rides on a constant K=16 step (a column of ones); rows without a k-th neighbour are masked in the epilogue (the
//     slots of a row are sorted, so slot k is valid iff k < n_valid): no per-slot flag operands.
// Arithmetic per element is that of k_fusion_tc (same operand split, same products, fp32 accumulation in TMEM); only the
// order of the K-pool additions differs (registers instead of TMEM read-modify-write: same order k = 0 .. K-1).
//
// Two CTAs of 256 threads per SM: while one waits for its MMAs or its BEV lines the other builds operands.
#include <stdio.h>
#include <stdlib.h>

#include "cf_common.cuh"
#include "cf_tcgen05.cuh"

namespace cf {

namespace {

constexpr int kTile = 128;       // rows per tile == UMMA M
constexpr int kSeg = 32;         // cells per segment
constexpr int kSegTile = 4;      // segments per tile
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

struct SegParams {
    const float *bev;
    const float *T;
    const int32_t *knn;
    float *out;
    const uint8_t *wimg2;
    const uint8_t *wimg3;
    const float *W1;
    const float *b2;
    const float *b3;
    int32_t B, N, W, K, Ci;
    int32_t cells, nseg;          // cells per frame, segments per frame = ceil(cells / 32)
    float x0, y0, dx, dy;
    const int32_t *seg_list;      // (B, nseg): live segments from the front, empty segments from the back
    const int32_t *seg_count;     // [b] live segments, [64 + b] empty segments
    int32_t copy_dead;            // out != bev: this kernel also copies the empty segments
    unsigned long long *prof;     // CF_SEG_PROF: per-phase cycle sums of thread 0 and thread 160 of every CTA (nullptr: off)
};

// the "row" a (cell, k) slot without a neighbour gathers: relu(-1e30 - e) = 0
#define CF_NEG8 -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f
__device__ __align__(32) float g_seg_neg_row[64] = {CF_NEG8, CF_NEG8, CF_NEG8, CF_NEG8, CF_NEG8, CF_NEG8, CF_NEG8, CF_NEG8};
#undef CF_NEG8

__host__ __device__ constexpr int seg_tmem_cols(int cols) { return cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512; }

template <int C, int NS, int G>
struct SegLayout {
    static constexpr int kWLayer = NS * C * C * 2;                  // one layer's packed image (resident)
    static constexpr int kOffA = 2 * kWLayer;                       // G operand buffers [hi | lo]
    static constexpr int kASlot = NS * kTile * C * 2;
    // bias operands of the extra K=16 step: only its first 16-byte k-unit carries data (A: column 0/1 = 1 resp. the row's
    // n_valid; B: (hi(b), lo(b))); the second k-unit of all four operands is ONE shared block of zeros, reached through the
    // descriptor's leading byte offset (16 row groups x 128 B, stride byte offset 128 like the data units)
    static constexpr int kOffOnes = kOffA + G * kASlot;             // A, layer 2: 128 rows x 16 B
    static constexpr int kOffCnt = kOffOnes + kTile * 16;           // A, layer 3
    static constexpr int kOffWb = kOffCnt + kTile * 16;             // B: layer 2 | layer 3, C rows x 16 B each
    static constexpr int kOffZero = kOffWb + 2 * C * 16;            // shared zero k-unit (2 KB), behind every data unit
    static constexpr int kOffCtr = kOffZero + kTile * 16;           // float2 (cx, cy) [2][128]
    static constexpr int kOffNv = kOffCtr + 2 * kTile * 8;          // int32 n_valid [2][128]
    static constexpr int kOffMisc = kOffNv + 2 * kTile * 4;         // mbarrier (8), tmem slot (4), pad (4), wmax[2][8] int32, seg[2][4] int32
    static constexpr int kOffCounts = kOffMisc + 128;               // int32 tiles[64] | dead units[64]
    static constexpr int kOffIdx = kOffCounts + 512;                // int32 [2][K][128]
    static __host__ __device__ constexpr int smem_bytes(int K) { return kOffIdx + 2 * K * kTile * 4; }
    static constexpr int kTmemCols = seg_tmem_cols(G * C);
    static_assert(G * C <= 512, "accumulators exceed the tensor memory");
};

// ---------------------------------------------------------------------------------------------------------------------
// Segment compaction: one warp per segment, 32 segments per block; block-local order, one atomicAdd per block and list.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_seg_compact(const int32_t *__restrict__ knn, int32_t K, int32_t cells, int32_t nseg,
                                                       int32_t *__restrict__ list, int32_t *__restrict__ count)
{
    __shared__ int32_t flag[32];
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t seg = blockIdx.x * 32 + warp;
    const int32_t cell = seg * kSeg + lane;
    const bool live = seg < nseg && cell < cells && __ldg(knn + ((size_t)b * cells + cell) * K) >= 0;
    const bool any = __any_sync(0xffffffffu, live);
    if (lane == 0) flag[warp] = seg < nseg ? (any ? 1 : 2) : 0;
    __syncthreads();
    if (warp == 0) {
        const int f = flag[lane];
        const unsigned bl = __ballot_sync(0xffffffffu, f == 1), bd = __ballot_sync(0xffffffffu, f == 2);
        int32_t base_l = 0, base_d = 0;
        if (lane == 0) {
            if (bl) base_l = atomicAdd(count + b, __popc(bl));
            if (bd) base_d = atomicAdd(count + 64 + b, __popc(bd));
        }
        base_l = __shfl_sync(0xffffffffu, base_l, 0);
        base_d = __shfl_sync(0xffffffffu, base_d, 0);
        const unsigned below = (1u << lane) - 1u;
        int32_t *fl = list + (size_t)b * nseg;
        const int32_t s = blockIdx.x * 32 + lane;
        if (f == 1) fl[base_l + __popc(bl & below)] = s;
        if (f == 2) fl[nseg - 1 - (base_d + __popc(bd & below))] = s;
    }
}

__device__ __forceinline__ float4 ldcs_f4(const float *p) { return __ldcs(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void stcs_f4(float *p, const float4 &v) { __stcs(reinterpret_cast<float4 *>(p), v); }

template <int C, int NS, int G>
__global__ void __launch_bounds__(kThreads, 2) k_fusion_seg(const SegParams p)
{
    using L = SegLayout<C, NS, G>;
    constexpr int kc_units = C / 8;
    constexpr int kQuads = C / 32;                 // 32-channel quads per row
    constexpr int kRG = 16 / (kWarps / kQuads);    // row groups (8 rows) per warp and slot
    constexpr int CH = C / 2;                      // channels per epilogue thread
    constexpr uint32_t idesc = tc::make_idesc_bf16(kTile, C);
    extern __shared__ __align__(1024) uint8_t smem[];
    float2 *sctr = reinterpret_cast<float2 *>(smem + L::kOffCtr);
    int32_t *snv = reinterpret_cast<int32_t *>(smem + L::kOffNv);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + L::kOffMisc);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::kOffMisc + 8);
    int32_t *swmax = reinterpret_cast<int32_t *>(smem + L::kOffMisc + 16);     // [2][8]
    int32_t *sseg = reinterpret_cast<int32_t *>(smem + L::kOffMisc + 80);      // [2][4]
    int32_t *stiles = reinterpret_cast<int32_t *>(smem + L::kOffCounts);       // [64] tiles per frame
    int32_t *sdead = stiles + 64;                                              // [64] empty-segment units per frame

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = p.K;
    const int32_t cells = p.cells, nseg = p.nseg;

    // ---- one-time setup ---------------------------------------------------------------------------------------------
    if (tid == 0) {
        tc::mbar_init(bar, 1);
        tc::mbar_fence_init();
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(tmem_slot, L::kTmemCols);
    for (int b = tid; b < 64; b += kThreads) {
        const int32_t nl = b < p.B ? __ldg(p.seg_count + b) : 0, nd = b < p.B && p.copy_dead ? __ldg(p.seg_count + 64 + b) : 0;
        stiles[b] = (nl + kSegTile - 1) / kSegTile;
        sdead[b] = (nd * kQuads + kWarps - 1) / kWarps;   // copy item = (empty segment, 32 channels), one per warp and unit
    }
    // bias operands: ones | count (A side), b2 | b3 as (hi, lo) pairs (B side); everything else in them stays zero
    for (int o = tid * 16; o < 3 * kTile * 16 + 2 * C * 16; o += kThreads * 16) *reinterpret_cast<uint4 *>(smem + L::kOffOnes + o) = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (tid < kTile) tc::sts_u32(tc::smem_u32(smem + L::kOffOnes) + tid * 16, 0x3F803F80u);
    for (int n = tid; n < 2 * C; n += kThreads) {
        const int layer = n / C, c = n - layer * C;
        const float bv = __ldg((layer ? p.b3 : p.b2) + c);
        const __nv_bfloat16 h = __float2bfloat16_rn(bv);
        const __nv_bfloat16 l = __float2bfloat16_rn(bv - __bfloat162float(h));
        const uint32_t packed = (uint32_t)__bfloat16_as_ushort(h) | ((uint32_t)__bfloat16_as_ushort(l) << 16);
        *reinterpret_cast<uint32_t *>(smem + L::kOffWb + layer * C * 16 + c * 16) = packed;
    }
    for (int o = tid * 16; o < L::kWLayer; o += kThreads * 16) {
        *reinterpret_cast<uint4 *>(smem + o) = __ldg(reinterpret_cast<const uint4 *>(p.wimg2 + o));
        *reinterpret_cast<uint4 *>(smem + L::kWLayer + o) = __ldg(reinterpret_cast<const uint4 *>(p.wimg3 + o));
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t s0 = tc::smem_u32(smem);
    const uint32_t sA_addr = s0 + L::kOffA, sidx_addr = s0 + L::kOffIdx, sctr_addr = s0 + L::kOffCtr, snv_addr = s0 + L::kOffNv;
    uint32_t phase = 0, iter = 0;
    // phase profiler (debug): thread 0 (the MMA issuer) and thread 160 (warp 5) add the cycles since their last mark
    const bool prof_on = p.prof != nullptr && (tid == 0 || tid == 160);
    long long prof_t = prof_on ? clock64() : 0;
    auto mark = [&](int ph) {
        if (prof_on) {
            const long long t = clock64();
            atomicAdd(p.prof + (tid == 0 ? 0 : 16) + ph, (unsigned long long)(t - prof_t));
            prof_t = t;
        }
    };

    // ---- operand build roles ------------------------------------------------------------------------------------------
    // warp w: quad = w % kQuads (32 channels), row groups rg0 .. rg0 + kRG - 1; lane (r8 = lane % 8, u = lane / 8)
    const int quad = warp % kQuads, rg0 = (warp / kQuads) * kRG;
    const int ku = quad * 4 + (lane >> 3), r8 = lane & 7;
    float2 nx[4], ny[4];   // -(w1x, w1y) of this lane's 8 channels
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = ku * 8 + 2 * i;
        nx[i] = make_float2(-__ldg(p.W1 + (size_t)c * (p.Ci + 3) + p.Ci), -__ldg(p.W1 + (size_t)(c + 1) * (p.Ci + 3) + p.Ci));
        ny[i] = make_float2(-__ldg(p.W1 + (size_t)c * (p.Ci + 3) + p.Ci + 1), -__ldg(p.W1 + (size_t)(c + 1) * (p.Ci + 3) + p.Ci + 1));
    }
    // epilogue roles: thread = (row, half of the channels)
    const int row = (warp & 3) * 32 + lane, half = warp >> 2;

    // ---- tile sequencing: position (frame, tile in frame), every gridDim.x-th position is this CTA's ---------------------
    auto advance = [&](const int32_t *cnt, int32_t &b, int32_t &q, int32_t step) {
        q += step;
        while (b < p.B && q >= cnt[b]) {
            q -= cnt[b];
            ++b;
        }
    };
    int32_t cb = 0, cq = 0, db = 0, dq = 0;
    advance(stiles, cb, cq, (int32_t)blockIdx.x);
    if (p.copy_dead) advance(sdead, db, dq, (int32_t)blockIdx.x); else db = p.B;

    // tile header (threads 0..127, thread = row): segment, cell, the K neighbour indices (cp.async), centre
    auto header_fill = [&](int32_t b, int32_t q, int par) {
        const int32_t e = q * kSegTile + (tid >> 5);
        int32_t seg = -1;
        if (e < __ldg(p.seg_count + b)) seg = __ldg(p.seg_list + (size_t)b * nseg + e);
        if (lane == 0) sseg[par * 4 + (tid >> 5)] = seg;
        const int32_t cell = seg >= 0 ? seg * kSeg + lane : -1;
        const bool inside = cell >= 0 && cell < cells;
        const uint32_t dst = sidx_addr + (uint32_t)((par * K * kTile + tid) * 4);
        float cx = 0.f, cy = 0.f;
        if (inside) {
            const int32_t *kr = p.knn + ((size_t)b * cells + cell) * K;
            for (int k = 0; k < K; ++k) tc::cp_async4(dst + k * kTile * 4, kr + k);
            const int32_t i = (int32_t)((uint32_t)cell / (uint32_t)p.W), j = cell - i * p.W;
            cx = __fadd_rn(p.x0, __fmul_rn((float)i, p.dx));
            cy = __fadd_rn(p.y0, __fmul_rn((float)j, p.dy));
        } else {
            for (int k = 0; k < K; ++k) tc::sts_u32(dst + k * kTile * 4, 0xFFFFFFFFu);
        }
        tc::cp_async_commit();
        sctr[par * kTile + tid] = make_float2(cx, cy);
    };
    if (cb < p.B && tid < kTile) header_fill(cb, cq, 0);

    // one empty segment per warp: bev -> out with 128-bit accesses (lane = (channel % 4, 16 bytes of the 128-byte line))
    // copy item = (empty segment, block of 32 channels): 8 float4 per lane
    constexpr int kCopyRegs = 8;
    auto dead_segment = [&](int32_t b, int32_t q, int32_t &cblock) -> int32_t {   // this warp's item of unit (b, q), -1: none
        const int32_t it = q * kWarps + warp, e = it / kQuads;
        cblock = it - e * kQuads;
        if (e >= __ldg(p.seg_count + 64 + b)) return -1;
        return __ldg(p.seg_list + (size_t)b * nseg + (nseg - 1 - e));
    };
    auto dead_offset = [&](int32_t b, int32_t seg, int32_t cblock) -> size_t {
        return ((size_t)b * C + cblock * 32 + (lane >> 3)) * cells + (size_t)seg * kSeg + (lane & 7) * 4;
    };

    while (cb < p.B || db < p.B) {
        if (cb < p.B) {
            const int par = iter & 1;
            ++iter;
            const int b = cb;
            int32_t nb = cb, nq = cq;
            advance(stiles, nb, nq, (int32_t)gridDim.x);
            const bool has_next = nb < p.B;
            // ---- header of this tile (prefetched) ----------------------------------------------------------------------
            if (tid < kTile) {
                tc::cp_async_wait_all();
                int nv = 0;
                for (int k = 0; k < K; ++k) nv += (int32_t)tc::lds_u32(sidx_addr + (uint32_t)(((par * K + k) * kTile + tid) * 4)) >= 0;
                snv[par * kTile + tid] = nv;
                const int wm = __reduce_max_sync(0xffffffffu, nv);
                if (lane == 0) swmax[par * 8 + warp] = wm;
            }
            mark(0);   // header: cp.async wait, n_valid
            __syncthreads();
            mark(1);   // barrier S0
            const int4 wm4 = *reinterpret_cast<const int4 *>(swmax + par * 8);
            const int R = max(max(wm4.x, wm4.y), max(wm4.z, wm4.w));
            if (has_next && tid < kTile) header_fill(nb, nq, par ^ 1);   // in flight during the whole tile

            const int32_t seg = sseg[par * 4 + (warp & 3)];
            const int32_t cell = seg >= 0 ? seg * kSeg + lane : -1;
            const bool in_range = cell >= 0 && cell < cells;
            const int nv_row = snv[par * kTile + row];
            const float *Tb = p.T + (size_t)b * p.N * C + ku * 8;
            const float *neg = g_seg_neg_row + ku * 8;

            float pooled[CH];
            for (int k0 = 0; k0 < R; k0 += G) {
                const int ns = min(G, R - k0);
                // ---- build the ns operand tiles of this batch ------------------------------------------------------------
#pragma unroll 1
                for (int h = 0; h < kRG; ++h) {
                    const int rg = rg0 + h;
                    const int rrow = rg * 8 + r8;
                    const uint32_t idx_a = sidx_addr + (uint32_t)(((par * K + k0) * kTile + rrow) * 4);
                    const int nvr = (int32_t)tc::lds_u32(snv_addr + (uint32_t)((par * kTile + rrow) * 4));
                    const int gmax = __reduce_max_sync(0xffffffffu, nvr) - k0;   // slots of this batch that any row of the group uses
                    float tv[G][8];
#pragma unroll
                    for (int j = 0; j < G; ++j) {
                        if (j < ns && j < gmax) {
                            const int32_t pr = (int32_t)tc::lds_u32(idx_a + j * kTile * 4);
                            tc::ldg_nc_f32x8(pr >= 0 ? Tb + (size_t)pr * C : neg, tv[j]);
                        }
                    }
                    const float2 ctr = tc::lds_f32x2(sctr_addr + (uint32_t)((par * kTile + rrow) * 8));
                    const float2 cxx = make_float2(ctr.x, ctr.x), cyy = make_float2(ctr.y, ctr.y);
                    const uint32_t dst0 = sA_addr + tc::unit_offset(rrow, ku, kc_units);
#pragma unroll
                    for (int j = 0; j < G; ++j) {
                        if (j < ns) {
                            uint4 hi = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
                            if (j < gmax) {
                                float2 v[4];
#pragma unroll
                                for (int i = 0; i < 4; ++i) v[i] = tc::ffma2(nx[i], cxx, tc::ffma2(ny[i], cyy, make_float2(tv[j][2 * i], tv[j][2 * i + 1])));
                                tc::relu_split_bf16x8(v, hi, lo, NS == 2);
                            }
                            tc::sts_u32x4(dst0 + j * L::kASlot, hi);
                            if (NS == 2) tc::sts_u32x4(dst0 + j * L::kASlot + kTile * C * 2, lo);
                        }
                    }
                }
                mark(2);   // gather + build
                tc::fence_proxy_async();
                tc::fence_before_sync();
                __syncthreads();
                mark(3);   // barrier S1
                if (tid == 0) {
                    tc::fence_after_sync();
                    for (int j = 0; j < ns; ++j) {
                        const uint32_t acc = tmem_base + j * C;
                        tc::mma_bf16(acc, tc::make_desc(s0 + L::kOffOnes, L::kOffZero - L::kOffOnes, 128),
                                     tc::make_desc(s0 + L::kOffWb, L::kOffZero - L::kOffWb, 128), idesc, 0u);
                        const uint32_t a = sA_addr + j * L::kASlot;
#pragma unroll
                        for (int kk = 0; kk < C / 16; ++kk) {
                            const uint32_t koff = kk * 256;
                            const uint64_t a_hi = tc::make_desc(a + koff, 128, kc_units * 128), w_hi = tc::make_desc(s0 + koff, 128, kc_units * 128);
                            tc::mma_bf16(acc, a_hi, w_hi, idesc, 1u);
                            if (NS == 2) {
                                const uint64_t a_lo = tc::make_desc(a + kTile * C * 2 + koff, 128, kc_units * 128);
                                const uint64_t w_lo = tc::make_desc(s0 + C * C * 2 + koff, 128, kc_units * 128);
                                tc::mma_bf16(acc, a_hi, w_lo, idesc, 1u);
                                tc::mma_bf16(acc, a_lo, w_hi, idesc, 1u);
                            }
                        }
                    }
                    tc::commit(bar);
                }
                mark(4);   // MMA issue
                tc::mbar_wait(bar, phase);
                phase ^= 1u;
                tc::fence_after_sync();
                mark(5);   // MMA wait
                // ---- pool: pooled += [k < n_valid] relu(acc_k) --------------------------------------------------------------
                __syncwarp();
#pragma unroll
                for (int cc = 0; cc < CH; cc += 16) {
#pragma unroll 1
                    for (int j = 0; j < ns; ++j) {
                        float z[16];
                        tc::tmem_ld16(tmem_base + lane_off + j * C + half * CH + cc, z);
                        const float f = k0 + j < nv_row ? 1.f : 0.f;
                        if (k0 == 0 && j == 0) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) pooled[cc + i] = fmaxf(z[i], 0.f) * f;
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i) pooled[cc + i] = fmaf(fmaxf(z[i], 0.f), f, pooled[cc + i]);
                        }
                    }
                }
                tc::fence_before_sync();
                mark(6);   // pool
            }
            const float *src_bev = p.bev + ((size_t)b * C + half * CH) * cells + cell;
            float *dst_out = p.out + ((size_t)b * C + half * CH) * cells + cell;

            // bev values of the final epilogue: in flight during layer 3
            float bv[CH];
            if (in_range) {
#pragma unroll
                for (int i = 0; i < CH; ++i) bv[i] = __ldcs(src_bev + (size_t)i * cells);
            }
            if (R > 0) {
                // ---- layer 3: acc = n_valid * b3 + pooled * W3^T (operand in buffer 0, accumulator in columns [0, C)) ------
#pragma unroll
                for (int q = 0; q < CH / 8; ++q) {
                    float2 v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[i] = make_float2(pooled[q * 8 + 2 * i], pooled[q * 8 + 2 * i + 1]);
                    uint4 hi, lo;
                    tc::relu_split_bf16x8(v, hi, lo, NS == 2);   // pooled >= 0: the ReLU is the identity
                    const uint32_t off = sA_addr + tc::unit_offset(row, half * (CH / 8) + q, kc_units);
                    tc::sts_u32x4(off, hi);
                    if (NS == 2) tc::sts_u32x4(off + kTile * C * 2, lo);
                }
                if (half == 0) {
                    const uint32_t nv16 = __float_as_uint((float)nv_row) >> 16;   // small integers are exact in bf16
                    tc::sts_u32(s0 + L::kOffCnt + row * 16, nv16 | (nv16 << 16));
                }
                mark(7);   // bev loads issued, layer-3 operand
                tc::fence_proxy_async();
                tc::fence_before_sync();
                __syncthreads();
                mark(8);   // barrier S2
                if (tid == 0) {
                    tc::fence_after_sync();
                    tc::mma_bf16(tmem_base, tc::make_desc(s0 + L::kOffCnt, L::kOffZero - L::kOffCnt, 128),
                                 tc::make_desc(s0 + L::kOffWb + C * 16, L::kOffZero - L::kOffWb - C * 16, 128), idesc, 0u);
#pragma unroll
                    for (int kk = 0; kk < C / 16; ++kk) {
                        const uint32_t koff = kk * 256;
                        const uint64_t a_hi = tc::make_desc(sA_addr + koff, 128, kc_units * 128);
                        const uint64_t w_hi = tc::make_desc(s0 + L::kWLayer + koff, 128, kc_units * 128);
                        tc::mma_bf16(tmem_base, a_hi, w_hi, idesc, 1u);
                        if (NS == 2) {
                            const uint64_t a_lo = tc::make_desc(sA_addr + kTile * C * 2 + koff, 128, kc_units * 128);
                            const uint64_t w_lo = tc::make_desc(s0 + L::kWLayer + C * C * 2 + koff, 128, kc_units * 128);
                            tc::mma_bf16(tmem_base, a_hi, w_lo, idesc, 1u);
                            tc::mma_bf16(tmem_base, a_lo, w_hi, idesc, 1u);
                        }
                    }
                    tc::commit(bar);
                }
            }
            if (R > 0) {
                tc::mbar_wait(bar, phase);
                phase ^= 1u;
                tc::fence_after_sync();
            }
            mark(9);   // layer-3 issue + wait
            // ---- final epilogue: out = bev + acc (a warp writes one aligned 128-byte line per channel) --------------------
            __syncwarp();
            if (R > 0 || p.out != p.bev) {
#pragma unroll
                for (int cc = 0; cc < CH; cc += 16) {
                    float z[16];
                    if (R > 0) tc::tmem_ld16(tmem_base + lane_off + half * CH + cc, z);
                    if (in_range) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) __stcs(dst_out + (size_t)(cc + i) * cells, R > 0 ? bv[cc + i] + z[i] : bv[cc + i]);
                    }
                }
            }
            tc::fence_before_sync();
            mark(10);   // final epilogue
            cb = nb;
            cq = nq;
        }
        // ---- one unit of empty segments (one copy item per warp) after every MLP tile, and the rest when the tiles run out ---
        if (db < p.B) {
            int32_t blk;
            const int32_t dseg = dead_segment(db, dq, blk);
            if (dseg >= 0 && dseg * kSeg + (lane & 7) * 4 < cells) {
                const size_t o = dead_offset(db, dseg, blk);
                float4 cp[kCopyRegs];
#pragma unroll
                for (int i = 0; i < kCopyRegs; ++i) cp[i] = ldcs_f4(p.bev + o + (size_t)i * 4 * cells);
#pragma unroll
                for (int i = 0; i < kCopyRegs; ++i) stcs_f4(p.out + o + (size_t)i * 4 * cells, cp[i]);
            }
            advance(sdead, db, dq, (int32_t)gridDim.x);
            mark(11);   // empty-segment copy
        }
    }

    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_free(tmem_base, L::kTmemCols);
}

template <int C, int NS, int G>
int launch_seg(const SegParams &p, int64_t tiles_max, cudaStream_t st)
{
    using L = SegLayout<C, NS, G>;
    const int smem = L::smem_bytes(p.K);
    if (smem > 227 * 1024) return CF_ERR_UNSUPPORTED;
    static int attr_bytes = 0;
    if (smem > attr_bytes) {
        CF_TRY(cuda_status(cudaFuncSetAttribute(k_fusion_seg<C, NS, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "k_fusion_seg smem attribute"));
        attr_bytes = smem;
    }
    static int regs = 0;
    if (!regs) {
        cudaFuncAttributes fa;
        CF_TRY(cuda_status(cudaFuncGetAttributes(&fa, k_fusion_seg<C, NS, G>), "k_fusion_seg attributes"));
        regs = (fa.numRegs + 7) / 8 * 8;
    }
    int per_sm = std::min(std::min((228 * 1024) / (smem + 1024), 65536 / (regs * kThreads)), 512 / L::kTmemCols);
    static const int cap = getenv("CF_SEG_CTAS") ? atoi(getenv("CF_SEG_CTAS")) : 0;
    if (cap > 0) per_sm = std::min(per_sm, cap);
    per_sm = std::max(1, per_sm);
    const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(tiles_max, (int64_t)sm_count() * per_sm));
    static const bool dbg = getenv("CF_DEBUG_LAUNCH") != nullptr;
    if (dbg) fprintf(stderr, "k_fusion_seg<%d,%d,%d>: %d CTAs/SM grid %lld smem %d regs %d\n", C, NS, G, per_sm, (long long)grid, smem, regs);
    k_fusion_seg<C, NS, G><<<(unsigned)grid, kThreads, smem, st>>>(p);
    return CF_OK;
}

}  // namespace

// workspace of the segment path: [counts: 128 int32][lists: B * nseg int32]
size_t fusion_seg_workspace_bytes(int32_t B, int32_t H, int32_t W) { return 512 + (size_t)B * ceil_div64((int64_t)H * W, kSeg) * 4; }

// returns CF_ERR_UNSUPPORTED (nothing launched, no error text) when the shape is not handled here
int fusion_seg(const float *d_bev, const float *d_T, const int32_t *d_knn, int32_t B, int32_t N, int32_t C, int32_t H, int32_t W,
               int32_t K, float x0, float y0, float dx, float dy, const float *d_W1, int32_t Ci, const float *d_b2, const float *d_b3,
               float *d_out, int32_t mode, const uint8_t *img2, const uint8_t *img3, void *d_ws, cudaStream_t st)
{
    const int64_t cells = (int64_t)H * W;
    if (C != 32 && C != 64) return CF_ERR_UNSUPPORTED;
    if (B > 64 || cells % 4 != 0 || cells >= (1ll << 30) || !aligned16(d_bev) || !aligned16(d_out)) return CF_ERR_UNSUPPORTED;
    const int32_t nseg = (int32_t)ceil_div64(cells, kSeg);
    int32_t *count = (int32_t *)d_ws, *list = count + 128;
    CF_TRY(cuda_status(cudaMemsetAsync(count, 0, 512, st), "cf_fusion_fwd memset"));
    k_seg_compact<<<dim3((unsigned)ceil_div64(nseg, 32), (unsigned)B), 1024, 0, st>>>(d_knn, K, (int32_t)cells, nseg, list, count);
    count_launches(1);
    SegParams p;
    p.bev = d_bev; p.T = d_T; p.knn = d_knn; p.out = d_out; p.wimg2 = img2; p.wimg3 = img3; p.W1 = d_W1; p.b2 = d_b2; p.b3 = d_b3;
    p.B = B; p.N = N; p.W = W; p.K = K; p.Ci = Ci; p.cells = (int32_t)cells; p.nseg = nseg;
    p.x0 = x0; p.y0 = y0; p.dx = dx; p.dy = dy;
    p.seg_list = list; p.seg_count = count; p.copy_dead = d_out != d_bev;
    p.prof = nullptr;
    if (getenv("CF_SEG_PROF")) {   // debug: 32 counters, printed (and the stream synchronised) after the launch
        static unsigned long long *d_prof = nullptr;
        if (!d_prof) cudaMalloc(&d_prof, 32 * 8);
        cudaMemsetAsync(d_prof, 0, 32 * 8, st);
        p.prof = d_prof;
    }
    const int64_t tiles_max = ceil_div64(nseg, kSegTile) * B;
    const int NS = mode == CF_MODE_FP32 ? 2 : 1;
    int rc;
    if (C == 32) rc = NS == 2 ? launch_seg<32, 2, 5>(p, tiles_max, st) : launch_seg<32, 1, 5>(p, tiles_max, st);
    else rc = NS == 2 ? launch_seg<64, 2, 2>(p, tiles_max, st) : launch_seg<64, 1, 4>(p, tiles_max, st);
    if (rc != CF_OK) return rc;
    count_launches(1);
    if (p.prof) {
        unsigned long long h[32];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, p.prof, sizeof(h), cudaMemcpyDeviceToHost);
        static const char *names[12] = {"header", "S0", "build", "S1", "mma issue", "mma wait", "pool", "bev+L3 operand", "S2", "L3 wait", "final", "copy"};
        for (int t = 0; t < 2; ++t) {
            unsigned long long sum = 0;
            for (int i = 0; i < 12; ++i) sum += h[t * 16 + i];
            fprintf(stderr, "seg prof thread %d:", t ? 160 : 0);
            for (int i = 0; i < 12; ++i) fprintf(stderr, " %s %.1f%%", names[i], 100.0 * h[t * 16 + i] / (double)std::max(sum, 1ull));
            fprintf(stderr, "  (total %.0f kcycles per CTA)\n", sum / 1e3 / std::max(1, std::min((int)tiles_max, sm_count() * 2)));
        }
    }
    return launch_status("cf_fusion_fwd (tcgen05, segment tiles)");
}

}  // namespace cf
