// cf_fusion_seg.cu -- EXPERIMENTAL variant of K-4 for the finest scale (C = 32), opt-in with CF_SEG=1 (the product path is
// k_fusion_tc in cf_mlp_tc.cu, which is faster: 256 us against 356 us out of place at BASELINE configs[1]; the measurements and
// what they showed are in profiles/README.md).  It is kept, with a full-size parity test (tests/test_gpu_fullsize.py), as the
// TMA / mbarrier-pipeline formulation of the layer: the fused MLP + K-sum-pool + BEV add on SEGMENT tiles, in which no warp
// waits for a CTA-wide barrier inside a tile and no BEV byte passes through a register.
//
// Unit of work = a SEGMENT: 32 consecutive BEV cells (linear index), i.e. one 128-byte line of every channel plane.
// k_seg_compact splits the segments of a frame into those with at least one cell that has a neighbour (front of the list)
// and those without (back).  At BASELINE configs[1] 57 % of the 32-cell segments are entirely empty (90 % of the empty
// cells), 31 % entirely live, 12 % mixed.  Tile = 4 live segments = 128 rows = UMMA M.
//
// Roles (256 threads = 8 warps, two CTAs per SM):
//   every warp builds operands and runs the epilogues: warp w owns rows 16 w .. 16 w + 15 in the operand build (lane = row % 8,
//              8 channels) and segment w % 4 / channel half w / 4 in the epilogues (TMEM lanes 32 (w % 4) .. +31).
//   warp 7     is also the driver: after its own part of a slot it probes (mbarrier.test_wait) whether the slot's operand
//              buffer is complete and then lets ONE elected lane issue the slot's tcgen05.mma (descriptors stay in uniform
//              registers, the UTCHMMAs go out back to back).  (A ninth, dedicated driver warp was measured first: with 18
//              warps per SM the register file gives 96 registers per thread and the gather prefetch spilled.)
// Data flow of a tile:
//   * the K neighbour slots are built one after the other into a RING of 3 operand buffers (bf16 hi | lo, 16 KB each); a slot
//     is handed to the driver through an mbarrier (8 arrivals: one per warp), its MMAs go into the slot's OWN TMEM accumulator
//     and commit the buffer back (tcgen05.commit -> mbarrier), so builds, MMAs and the gathers of later slots overlap; the T
//     rows of slot j + 1 are gathered (LDG.256 into registers) before slot j is built.
//   * after the last slot one commit tells the warps that all accumulators are complete; ONE epilogue reads them, applies the
//     ReLU and the valid mask and sums them in registers (the pooled sum never round-trips through TMEM), writes the pooled
//     tile as the layer-3 operand into the next ring buffer, and the driver issues layer 3 into its own accumulator.
//   * BEV: every warp owns one box (32 cells x 16 channels, 2 KB) of the tile.  The box is loaded by TMA (2-D tensor map over
//     the (B C, cells) planes) at the start of the tile, the final epilogue adds the layer-3 result to it in shared memory,
//     and a TMA store writes it to `out`: no bev value is ever held in a register across a wait.
//   * empty segments: every warp also owns one 2 KB staging box through which it copies empty segments bev -> out with a
//     TMA load / TMA store pair, polled (never waited for) twice per tile: the copy stream runs beside the MLP tiles and costs
//     two instructions of one lane per 2 KB (+28 us for 57 % of the map, where the thread copies of k_fusion_tc cost +70 us).
//   * b2 rides on a constant K=16 step (a column of ones); rows without a k-th neighbour are masked in the epilogue (the
//     slots of a row are sorted, so slot k is valid iff k < n_valid); n_valid * b3 rides on a K=16 step of layer 3.  The second
//     k-unit of all four bias operands is one shared block of zeros reached through the descriptors' leading byte offset.
// Arithmetic per element is that of k_fusion_tc (same operand split, same products, fp32 accumulation in TMEM); the K-pool
// additions run in the same order k = 0 .. K-1.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "cf_common.cuh"
#include "cf_tcgen05.cuh"

namespace cf {

namespace {

constexpr int kTile = 128;       // rows per tile == UMMA M
constexpr int kSeg = 32;         // cells per segment
constexpr int kSegTile = 4;      // segments per tile
constexpr int kWarps = 8;        // build / epilogue warps
constexpr int kThreads = kWarps * 32;
constexpr int kDriver = kWarps - 1;   // the warp whose elected lane also issues the MMAs
constexpr int kRing = 3;         // operand buffers
#ifndef CF_SEG_PROFILE
#define CF_SEG_PROFILE 0         // make EXTRA=-DCF_SEG_PROFILE=1 + CF_SEG_PROF=1 in the environment: per-phase cycle counters
#endif

struct SegParams {
    const float *T;
    const int32_t *knn;
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
    unsigned long long *prof;     // CF_SEG_PROF: per-phase cycle sums of warp 5's lane 0 in every CTA (nullptr: off)
};

// the "row" a (cell, k) slot without a neighbour gathers: relu(-1e30 - e) = 0
#define CF_NEG8 -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f
__device__ __align__(32) float g_seg_neg_row[64] = {CF_NEG8, CF_NEG8, CF_NEG8, CF_NEG8, CF_NEG8, CF_NEG8, CF_NEG8, CF_NEG8};
#undef CF_NEG8

__host__ __device__ constexpr int seg_tmem_cols(int cols) { return cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512; }

template <int C, int NS, int G>
struct SegLayout {
    static constexpr int CH = C / 2;                                // channels per box
    static constexpr int kBox = CH * kSeg * 4;                      // one TMA box: CH channel rows x 32 cells
    static constexpr int kWLayer = NS * C * C * 2;                  // one layer's packed image (resident)
    static constexpr int kOffA = 2 * kWLayer;                       // ring of operand buffers [hi | lo]
    static constexpr int kASlot = NS * kTile * C * 2;
    // bias operands of the extra K=16 step: only its first 16-byte k-unit carries data (A: column 0/1 = 1 resp. the row's
    // n_valid; B: (hi(b), lo(b))); the second k-unit of all four operands is ONE shared block of zeros, reached through the
    // descriptor's leading byte offset (16 row groups x 128 B, stride byte offset 128 like the data units)
    static constexpr int kOffOnes = kOffA + kRing * kASlot;         // A, layer 2: 128 rows x 16 B
    static constexpr int kOffCnt = kOffOnes + kTile * 16;           // A, layer 3
    static constexpr int kOffWb = kOffCnt + kTile * 16;             // B: layer 2 | layer 3, C rows x 16 B each
    static constexpr int kOffZero = kOffWb + 2 * C * 16;            // shared zero k-unit (2 KB), behind every data unit
    static constexpr int kOffBev = kOffZero + kTile * 16;           // BEV boxes of the tile, one per warp
    static constexpr int kOffDead = kOffBev + kWarps * kBox;        // staging boxes of the empty-segment copies, one per warp
    static constexpr int kOffCtr = kOffDead + kWarps * kBox;        // float2 (cx, cy) [2][128]
    static constexpr int kOffNv = kOffCtr + 2 * kTile * 8;          // int32 n_valid [2][128]
    static constexpr int kOffBar = kOffNv + 2 * kTile * 4;          // mbarriers: full[3] empty[3] acc l3 bev[8] dead[8] = 24 x 8 B
    static constexpr int kOffMisc = kOffBar + 24 * 8;               // tmem slot (4), pad (12), wmax[2][4] int32, seg[2][4] int32
    static constexpr int kOffW1 = kOffMisc + 80;                    // negated offset weights: per 8 channels -w1x[8] | -w1y[8]
    static constexpr int kOffCounts = kOffW1 + 2 * C * 4;           // int32 tiles[64] | dead units[64]
    static constexpr int kOffIdx = kOffCounts + 512;                // int32 [2][K][128]
    static __host__ __device__ constexpr int smem_bytes(int K) { return kOffIdx + 2 * K * kTile * 4; }
    static constexpr int kTmemCols = seg_tmem_cols((G + 1) * C);    // G slot accumulators + the layer-3 accumulator
    static_assert((G + 1) * C <= 512, "accumulators exceed the tensor memory");
};
constexpr int kBarFull = 0, kBarEmpty = 3, kBarAcc = 6, kBarL3 = 7, kBarBev = 8, kBarDead = 16;

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

template <int C, int NS, int G>
__global__ void __launch_bounds__(kThreads, 2) k_fusion_seg(const SegParams p, const __grid_constant__ CUtensorMap tm_bev,
                                                             const __grid_constant__ CUtensorMap tm_out)
{
    using L = SegLayout<C, NS, G>;
    static_assert(C == 32, "one 32-channel quad per row: warp w builds rows 16 w .. 16 w + 15");
    constexpr int kc_units = C / 8;
    constexpr int CH = L::CH;
    constexpr uint32_t idesc = tc::make_idesc_bf16(kTile, C);
    extern __shared__ __align__(1024) uint8_t smem[];
    float2 *sctr = reinterpret_cast<float2 *>(smem + L::kOffCtr);
    int32_t *snv = reinterpret_cast<int32_t *>(smem + L::kOffNv);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::kOffMisc);
    int32_t *swmax = reinterpret_cast<int32_t *>(smem + L::kOffMisc + 16);     // [2][4]
    int32_t *sseg = reinterpret_cast<int32_t *>(smem + L::kOffMisc + 48);      // [2][4]
    int32_t *stiles = reinterpret_cast<int32_t *>(smem + L::kOffCounts);       // [64] tiles per frame
    int32_t *sdead = stiles + 64;                                              // [64] empty-segment units per frame

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = p.K;
    const int32_t cells = p.cells, nseg = p.nseg;
    const uint32_t s0 = tc::smem_u32(smem);
    const uint32_t bar0 = s0 + L::kOffBar;
    auto bar_at = [&](int i) -> uint32_t { return bar0 + (uint32_t)i * 8; };

    // ---- one-time setup ---------------------------------------------------------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < kRing; ++i) {
            tc::mbar_init(reinterpret_cast<uint64_t *>(smem + L::kOffBar) + kBarFull + i, kWarps);
            tc::mbar_init(reinterpret_cast<uint64_t *>(smem + L::kOffBar) + kBarEmpty + i, 1);
        }
        tc::mbar_init(reinterpret_cast<uint64_t *>(smem + L::kOffBar) + kBarAcc, 1);
        tc::mbar_init(reinterpret_cast<uint64_t *>(smem + L::kOffBar) + kBarL3, 1);
        for (int i = 0; i < 2 * kWarps; ++i) tc::mbar_init(reinterpret_cast<uint64_t *>(smem + L::kOffBar) + kBarBev + i, 1);
        tc::mbar_fence_init();
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(tmem_slot, L::kTmemCols);
    for (int b = tid; b < 64; b += kThreads) {
        const int32_t nl = b < p.B ? __ldg(p.seg_count + b) : 0, nd = b < p.B && p.copy_dead ? __ldg(p.seg_count + 64 + b) : 0;
        stiles[b] = (nl + kSegTile - 1) / kSegTile;
        sdead[b] = (nd * 2 + kWarps - 1) / kWarps;   // copy item = (empty segment, channel half) = one box, one item per warp and unit
    }
    // bias operands: ones | count (A side), b2 | b3 as (hi, lo) pairs (B side), the shared zero unit
    for (int o = tid * 16; o < 3 * kTile * 16 + 2 * C * 16; o += kThreads * 16) *reinterpret_cast<uint4 *>(smem + L::kOffOnes + o) = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (tid < kTile) tc::sts_u32(s0 + L::kOffOnes + tid * 16, 0x3F803F80u);
    for (int n = tid; n < 2 * C; n += kThreads) {
        const int layer = n / C, c = n - layer * C;
        const float bv = __ldg((layer ? p.b3 : p.b2) + c);
        const __nv_bfloat16 h = __float2bfloat16_rn(bv);
        const __nv_bfloat16 l = __float2bfloat16_rn(bv - __bfloat162float(h));
        const uint32_t packed = (uint32_t)__bfloat16_as_ushort(h) | ((uint32_t)__bfloat16_as_ushort(l) << 16);
        *reinterpret_cast<uint32_t *>(smem + L::kOffWb + layer * C * 16 + c * 16) = packed;
    }
    for (int c = tid; c < C; c += kThreads) {
        float *swn = reinterpret_cast<float *>(smem + L::kOffW1);
        swn[(c >> 3) * 16 + (c & 7)] = -__ldg(p.W1 + (size_t)c * (p.Ci + 3) + p.Ci);
        swn[(c >> 3) * 16 + 8 + (c & 7)] = -__ldg(p.W1 + (size_t)c * (p.Ci + 3) + p.Ci + 1);
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
    const uint32_t sA_addr = s0 + L::kOffA;

    // ---- tile sequencing: position (frame, tile in frame), every gridDim.x-th position is this CTA's ---------------------
    auto advance = [&](const int32_t *cnt, int32_t &b, int32_t &q, int32_t step) {
        q += step;
        while (b < p.B && q >= cnt[b]) {
            q -= cnt[b];
            ++b;
        }
    };

    {
        // ============================================ build / epilogue warps =============================================
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t sidx_addr = s0 + L::kOffIdx, sctr_addr = s0 + L::kOffCtr, snv_addr = s0 + L::kOffNv;
        const uint32_t bev_box = s0 + L::kOffBev + warp * L::kBox, dead_box = s0 + L::kOffDead + warp * L::kBox;
        const uint32_t bev_bar = bar_at(kBarBev + warp), dead_bar = bar_at(kBarDead + warp);
        // phase profiler (debug): lane 0 of warp 5 adds the cycles since its last mark
#if CF_SEG_PROFILE
        const bool prof_on = p.prof != nullptr && tid == 160;   // lane 0 of warp 5
        long long prof_t = prof_on ? clock64() : 0;
        auto mark = [&](int ph) {
            if (prof_on) {
                const long long t = clock64();
                atomicAdd(p.prof + ph, (unsigned long long)(t - prof_t));
                prof_t = t;
            }
        };
#else
        auto mark = [](int) {};
#endif
        // operand build: rows 16 warp + 8 h + r8 (h = 0, 1), the lane's 8 channels ku * 8 ..
        const int ku = lane >> 3, r8 = lane & 7;
        const uint32_t swn_addr = s0 + L::kOffW1 + (uint32_t)(ku * 64);   // -(w1x[8] | w1y[8]) of this lane's 8 channels
        // epilogues: thread = (row, half of the channels)
        const int row = (warp & 3) * 32 + lane, half = warp >> 2;

        int32_t cb = 0, cq = 0, db = 0, dq = 0;
        advance(stiles, cb, cq, (int32_t)blockIdx.x);
        if (p.copy_dead) advance(sdead, db, dq, (int32_t)blockIdx.x); else db = p.B;
        uint32_t buf = 0, use = 0, iter = 0, acc_ph = 0, l3_ph = 0, bev_ph = 0, dead_ph = 0;
        auto next_buf = [&]() {
            if (++buf == kRing) {
                buf = 0;
                ++use;
            }
        };
        // ---- MMA issue (warp kDriver only): its own ring position and the number of slots of the current tile issued so far ---
        uint32_t dbuf = 0, duse = 0;
        int dj = 0;
        auto dnext_buf = [&]() {
            if (++dbuf == kRing) {
                dbuf = 0;
                ++duse;
            }
        };
        // issue the MMAs of every slot of this tile (R slots) whose operand buffer is complete; blocking: of all of them
        auto drive_slots = [&](int limit, int R, bool blocking) {   // slots [dj, limit) of a tile with R slots
            while (dj < limit) {
                if (blocking) tc::mbar_wait_a(bar_at(kBarFull + dbuf), duse & 1u);
                else if (!tc::mbar_poll(bar_at(kBarFull + dbuf), duse & 1u)) break;
                if (tc::elect_one()) {
                    tc::fence_after_sync();
                    const uint32_t acc = tmem_base + (uint32_t)((dj % G) * C);
                    const uint32_t a = sA_addr + dbuf * L::kASlot;
                    tc::mma_bf16(acc, tc::make_desc(s0 + L::kOffOnes, L::kOffZero - L::kOffOnes, 128),
                                 tc::make_desc(s0 + L::kOffWb, L::kOffZero - L::kOffWb, 128), idesc, 0u);
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
                    tc::commit(reinterpret_cast<uint64_t *>(smem + L::kOffBar) + kBarEmpty + dbuf);
                    if ((dj + 1) % G == 0 || dj + 1 == R) tc::commit(reinterpret_cast<uint64_t *>(smem + L::kOffBar) + kBarAcc);
                }
                __syncwarp();
                dnext_buf();
                ++dj;
            }
        };
        auto drive_l3 = [&]() {
            tc::mbar_wait_a(bar_at(kBarFull + dbuf), duse & 1u);
            if (tc::elect_one()) {
                tc::fence_after_sync();
                const uint32_t acc = tmem_base + (uint32_t)(G * C);
                const uint32_t a = sA_addr + dbuf * L::kASlot;
                tc::mma_bf16(acc, tc::make_desc(s0 + L::kOffCnt, L::kOffZero - L::kOffCnt, 128),
                             tc::make_desc(s0 + L::kOffWb + C * 16, L::kOffZero - L::kOffWb - C * 16, 128), idesc, 0u);
#pragma unroll
                for (int kk = 0; kk < C / 16; ++kk) {
                    const uint32_t koff = kk * 256;
                    const uint64_t a_hi = tc::make_desc(a + koff, 128, kc_units * 128);
                    const uint64_t w_hi = tc::make_desc(s0 + L::kWLayer + koff, 128, kc_units * 128);
                    tc::mma_bf16(acc, a_hi, w_hi, idesc, 1u);
                    if (NS == 2) {
                        const uint64_t a_lo = tc::make_desc(a + kTile * C * 2 + koff, 128, kc_units * 128);
                        const uint64_t w_lo = tc::make_desc(s0 + L::kWLayer + C * C * 2 + koff, 128, kc_units * 128);
                        tc::mma_bf16(acc, a_hi, w_lo, idesc, 1u);
                        tc::mma_bf16(acc, a_lo, w_hi, idesc, 1u);
                    }
                }
                tc::commit(reinterpret_cast<uint64_t *>(smem + L::kOffBar) + kBarEmpty + dbuf);
                tc::commit(reinterpret_cast<uint64_t *>(smem + L::kOffBar) + kBarL3);
            }
            __syncwarp();
            dnext_buf();
            dj = 0;
        };

        // tile header (threads 0..127, thread = row): segment, cell, the K neighbour indices (cp.async), centre
        auto header_fill = [&](int32_t b, int32_t q, int par) {
            const int32_t e = q * kSegTile + warp;
            int32_t seg = -1;
            if (e < __ldg(p.seg_count + b)) seg = __ldg(p.seg_list + (size_t)b * nseg + e);
            if (lane == 0) sseg[par * 4 + warp] = seg;
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

        // ---- empty-segment copy stream of this warp: one staging box, load -> store -> (smem read done) -> load ... -------
        int dstate = 0;                 // 0: box free, 1: load in flight, 2: store issued (shared memory still being read)
        int32_t dcx = 0, dcy = 0;       // box coordinates of the item in the staging box
        auto dead_poll = [&](bool blocking) {
            if (dstate == 1) {
                if (blocking) tc::mbar_wait_a(dead_bar, dead_ph);
                else if (!tc::mbar_poll(dead_bar, dead_ph)) return;
                dead_ph ^= 1u;
                if (lane == 0) {
                    tc::tma_store_2d(&tm_out, dcx, dcy, dead_box);
                    tc::bulk_commit();
                }
                dstate = 2;
            }
            if (dstate == 2) {
                if (lane == 0) tc::bulk_wait_read0();
                dstate = 0;
            }
            while (dstate == 0 && db < p.B) {
                const int32_t itx = dq * kWarps + warp, e = itx >> 1;
                int32_t seg = -1;
                if (e < __ldg(p.seg_count + 64 + db)) seg = __ldg(p.seg_list + (size_t)db * nseg + (nseg - 1 - e));
                if (seg >= 0) {
                    dcx = seg * kSeg;
                    dcy = db * C + (itx & 1) * CH;
                    if (lane == 0) {
                        tc::mbar_expect_tx(dead_bar, L::kBox);
                        tc::tma_load_2d(dead_box, &tm_bev, dcx, dcy, dead_bar);
                    }
                    dstate = 1;
                }
                advance(sdead, db, dq, (int32_t)gridDim.x);
            }
        };

        while (cb < p.B) {
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
                if (lane == 0) swmax[par * 4 + warp] = wm;
            }
            mark(0);
            tc::named_bar_sync(1, kWarps * 32);
            mark(1);
            const int4 wm4 = *reinterpret_cast<const int4 *>(swmax + par * 4);
            const int R = max(1, max(max(wm4.x, wm4.y), max(wm4.z, wm4.w)));
            if (has_next && tid < kTile) header_fill(nb, nq, par ^ 1);   // in flight during the whole tile

            const int32_t seg = sseg[par * 4 + (warp & 3)];
            const int nv_row = snv[par * kTile + row];
            // this warp's BEV box of the tile: TMA load, consumed by the final epilogue
            if (seg >= 0 && lane == 0) {
                tc::bulk_wait_read0();   // the previous tile's store has finished reading the box
                tc::mbar_expect_tx(bev_bar, L::kBox);
                tc::tma_load_2d(bev_box, &tm_bev, seg * kSeg, b * C + half * CH, bev_bar);
            }
            dead_poll(false);
            mark(2);

            const float *Tb = p.T + (size_t)b * p.N * C + ku * 8;
            const float *neg = g_seg_neg_row + ku * 8;
            // rows of this lane in the build, their n_valid (group maxima decide which slots a group needs at all)
            const int rrow0 = warp * 16 + r8, rrow1 = rrow0 + 8;
            const int gm0 = __reduce_max_sync(0xffffffffu, (int32_t)tc::lds_u32(snv_addr + (uint32_t)((par * kTile + rrow0) * 4)));
            const int gm1 = __reduce_max_sync(0xffffffffu, (int32_t)tc::lds_u32(snv_addr + (uint32_t)((par * kTile + rrow1) * 4)));
            const uint32_t ctr_a = sctr_addr + (uint32_t)((par * kTile + rrow0) * 8);
            const uint32_t idx0 = sidx_addr + (uint32_t)((par * K * kTile + rrow0) * 4);
            const uint32_t dstu0 = tc::unit_offset(rrow0, ku, kc_units);   // rows 8 further down: + kc_units * 128

            float pooled[CH];
            for (int k0 = 0; k0 < R; k0 += G) {
                const int ns = min(G, R - k0);
                float tv[G][2][8];
                auto gather = [&](int j) {   // the T rows of slot k0 + j for this lane's two rows
                    const uint32_t ia = idx0 + (uint32_t)((k0 + j) * kTile * 4);
                    if (k0 + j < gm0) {
                        const int32_t pr = (int32_t)tc::lds_u32(ia);
                        tc::ldg_nc_f32x8_pinned(pr >= 0 ? Tb + (size_t)pr * C : neg, tv[j][0]);
                    }
                    if (k0 + j < gm1) {
                        const int32_t pr = (int32_t)tc::lds_u32(ia + 32);
                        tc::ldg_nc_f32x8_pinned(pr >= 0 ? Tb + (size_t)pr * C : neg, tv[j][1]);
                    }
                };
                auto build_one = [&](const float *t, bool any, const float2 &ctr, uint32_t dst) {
                    uint4 hi = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
                    if (any) {
                        const float2 cxx = make_float2(ctr.x, ctr.x), cyy = make_float2(ctr.y, ctr.y);
                        const float4 x0 = tc::lds_f32x4(swn_addr), x1 = tc::lds_f32x4(swn_addr + 16), y0 = tc::lds_f32x4(swn_addr + 32), y1 = tc::lds_f32x4(swn_addr + 48);
                        float2 v[4];
                        v[0] = tc::ffma2(make_float2(x0.x, x0.y), cxx, tc::ffma2(make_float2(y0.x, y0.y), cyy, make_float2(t[0], t[1])));
                        v[1] = tc::ffma2(make_float2(x0.z, x0.w), cxx, tc::ffma2(make_float2(y0.z, y0.w), cyy, make_float2(t[2], t[3])));
                        v[2] = tc::ffma2(make_float2(x1.x, x1.y), cxx, tc::ffma2(make_float2(y1.x, y1.y), cyy, make_float2(t[4], t[5])));
                        v[3] = tc::ffma2(make_float2(x1.z, x1.w), cxx, tc::ffma2(make_float2(y1.z, y1.w), cyy, make_float2(t[6], t[7])));
                        tc::relu_split_bf16x8(v, hi, lo, NS == 2);
                    }
                    tc::sts_u32x4(dst, hi);
                    if (NS == 2) tc::sts_u32x4(dst + kTile * C * 2, lo);
                };
                gather(0);
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    if (j + 1 < G && j + 1 < ns) gather(j + 1);
                    if (j < ns) {
                        // (the driver warp must have issued the slot that last used this buffer before it may wait for it)
                        if (warp == kDriver && k0 + j >= kRing) drive_slots(k0 + j - kRing + 1, R, true);
                        if (use > 0) tc::mbar_wait_a(bar_at(kBarEmpty + buf), (use - 1) & 1u);   // the MMAs that read this buffer are complete
                        const uint32_t a = sA_addr + buf * L::kASlot;
                        build_one(tv[j][0], k0 + j < gm0, tc::lds_f32x2(ctr_a), a + dstu0);
                        build_one(tv[j][1], k0 + j < gm1, tc::lds_f32x2(ctr_a + 64), a + dstu0 + kc_units * 128);
                        tc::fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive_a(bar_at(kBarFull + buf));
                        next_buf();
                        if (warp == kDriver) drive_slots(R, R, false);
                    }
                }
                if (warp == kDriver) drive_slots(min(R, k0 + G), R, true);
                mark(3);
                // ---- pool: pooled += [k < n_valid] relu(acc_k) ---------------------------------------------------------------
                tc::mbar_wait_a(bar_at(kBarAcc), acc_ph);
                acc_ph ^= 1u;
                tc::fence_after_sync();
                mark(4);
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
                mark(5);
            }
            // ---- layer 3 operand: the pooled tile (>= 0: the ReLU of the split is the identity) into the next ring buffer -------
            {
                if (use > 0) tc::mbar_wait_a(bar_at(kBarEmpty + buf), (use - 1) & 1u);
                const uint32_t a = sA_addr + buf * L::kASlot;
#pragma unroll
                for (int q = 0; q < CH / 8; ++q) {
                    float2 v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[i] = make_float2(pooled[q * 8 + 2 * i], pooled[q * 8 + 2 * i + 1]);
                    uint4 hi, lo;
                    tc::relu_split_bf16x8(v, hi, lo, NS == 2);
                    const uint32_t off = a + tc::unit_offset(row, half * (CH / 8) + q, kc_units);
                    tc::sts_u32x4(off, hi);
                    if (NS == 2) tc::sts_u32x4(off + kTile * C * 2, lo);
                }
                if (half == 0) {
                    const uint32_t nv16 = __float_as_uint((float)nv_row) >> 16;   // small integers are exact in bf16
                    tc::sts_u32(s0 + L::kOffCnt + row * 16, nv16 | (nv16 << 16));
                }
                tc::fence_proxy_async();
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive_a(bar_at(kBarFull + buf));
                next_buf();
                if (warp == kDriver) drive_l3();
            }
            mark(6);
            dead_poll(false);
            // ---- final epilogue: box += layer-3 result, TMA store ------------------------------------------------------------
            tc::mbar_wait_a(bar_at(kBarL3), l3_ph);
            l3_ph ^= 1u;
            tc::fence_after_sync();
            mark(7);
            if (seg >= 0) {
                tc::mbar_wait_a(bev_bar, bev_ph);
                bev_ph ^= 1u;
                mark(8);
#pragma unroll
                for (int cc = 0; cc < CH; cc += 16) {
                    float z[16];
                    tc::tmem_ld16(tmem_base + lane_off + G * C + half * CH + cc, z);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const uint32_t a = bev_box + (uint32_t)((cc + i) * 128 + lane * 4);
                        tc::sts_u32(a, __float_as_uint(__uint_as_float(tc::lds_u32(a)) + z[i]));
                    }
                }
                tc::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tc::tma_store_2d(&tm_out, seg * kSeg, b * C + half * CH, bev_box);
                    tc::bulk_commit();
                }
            }
            tc::fence_before_sync();
            mark(9);
            cb = nb;
            cq = nq;
        }
        // the rest of the empty segments
        while (db < p.B || dstate != 0) dead_poll(true);
        mark(10);
        if (lane == 0) tc::bulk_wait0();
    }

    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_free(tmem_base, L::kTmemCols);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

// 2-D tensor map over the channel planes of a (B, C, H, W) fp32 map: x = cell (fastest), y = b * C + c; box = 32 cells x rows
int make_plane_map(CUtensorMap *tm, const float *base, int64_t cells, int64_t planes, int box_rows)
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        CF_TRY(cuda_status(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q), "cuTensorMapEncodeTiled entry point"));
        CF_REQUIRE(f != nullptr && q == cudaDriverEntryPointSuccess, CF_ERR_LAUNCH, "cuTensorMapEncodeTiled is not available in this driver");
        fn = (EncodeTiledFn)f;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)cells, (cuuint64_t)planes};
    const cuuint64_t strides[1] = {(cuuint64_t)cells * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kSeg, (cuuint32_t)box_rows};
    const cuuint32_t el[2] = {1, 1};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, el, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CF_REQUIRE(r == CUDA_SUCCESS, CF_ERR_LAUNCH, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return CF_OK;
}

namespace {

template <int C, int NS, int G>
int launch_seg(const SegParams &p, const CUtensorMap &tm_bev, const CUtensorMap &tm_out, int64_t tiles_max, cudaStream_t st)
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
    k_fusion_seg<C, NS, G><<<(unsigned)grid, kThreads, smem, st>>>(p, tm_bev, tm_out);
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
    if (C != 32) return CF_ERR_UNSUPPORTED;
    // TMA: plane stride (cells * 4 bytes) and base addresses must be multiples of 16 bytes
    if (B > 64 || cells % 4 != 0 || cells >= (1ll << 30) || !aligned16(d_bev) || !aligned16(d_out)) return CF_ERR_UNSUPPORTED;
    const int32_t nseg = (int32_t)ceil_div64(cells, kSeg);
    // tensor maps of the two planes: encoded on the host (a few hundred ns), cached for the last (bev, out, shape)
    struct MapCache {
        const float *bev = nullptr;
        const float *out = nullptr;
        int64_t cells = 0, planes = 0;
        CUtensorMap tm_bev, tm_out;
    };
    static thread_local MapCache mc;
    const int64_t planes = (int64_t)B * C;
    if (mc.bev != d_bev || mc.out != d_out || mc.cells != cells || mc.planes != planes) {
        CF_TRY(make_plane_map(&mc.tm_bev, d_bev, cells, planes, C / 2));
        CF_TRY(make_plane_map(&mc.tm_out, d_out, cells, planes, C / 2));
        mc.bev = d_bev; mc.out = d_out; mc.cells = cells; mc.planes = planes;
    }
    int32_t *count = (int32_t *)d_ws, *list = count + 128;
    CF_TRY(cuda_status(cudaMemsetAsync(count, 0, 512, st), "cf_fusion_fwd memset"));
    k_seg_compact<<<dim3((unsigned)ceil_div64(nseg, 32), (unsigned)B), 1024, 0, st>>>(d_knn, K, (int32_t)cells, nseg, list, count);
    count_launches(1);
    SegParams p;
    p.T = d_T; p.knn = d_knn; p.wimg2 = img2; p.wimg3 = img3; p.W1 = d_W1; p.b2 = d_b2; p.b3 = d_b3;
    p.B = B; p.N = N; p.W = W; p.K = K; p.Ci = Ci; p.cells = (int32_t)cells; p.nseg = nseg;
    p.x0 = x0; p.y0 = y0; p.dx = dx; p.dy = dy;
    p.seg_list = list; p.seg_count = count; p.copy_dead = d_out != d_bev;
    p.prof = nullptr;
    if (CF_SEG_PROFILE && getenv("CF_SEG_PROF")) {   // debug: 16 counters, printed (and the stream synchronised) after the launch
        static unsigned long long *d_prof = nullptr;
        if (!d_prof) cudaMalloc(&d_prof, 16 * 8);
        cudaMemsetAsync(d_prof, 0, 16 * 8, st);
        p.prof = d_prof;
    }
    const int64_t tiles_max = ceil_div64(nseg, kSegTile) * B;
    const int NS = mode == CF_MODE_FP32 ? 2 : 1;
    const int rc = NS == 2 ? launch_seg<32, 2, 5>(p, mc.tm_bev, mc.tm_out, tiles_max, st) : launch_seg<32, 1, 5>(p, mc.tm_bev, mc.tm_out, tiles_max, st);
    if (rc != CF_OK) return rc;
    count_launches(1);
    if (p.prof) {
        unsigned long long h[16];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, p.prof, sizeof(h), cudaMemcpyDeviceToHost);
        static const char *names[11] = {"header", "T0 barrier", "bev load + poll", "gather + build", "acc wait", "pool", "L3 operand", "L3 wait", "box wait", "final", "copy tail"};
        unsigned long long sum = 0;
        for (int i = 0; i < 11; ++i) sum += h[i];
        fprintf(stderr, "seg prof warp 5:");
        for (int i = 0; i < 11; ++i) fprintf(stderr, " %s %.1f%%", names[i], 100.0 * h[i] / (double)std::max(sum, 1ull));
        fprintf(stderr, "  (total %.0f kcycles per CTA)\n", sum / 1e3 / std::max(1, std::min((int)tiles_max, sm_count() * 2)));
    }
    return launch_status("cf_fusion_fwd (tcgen05, segment tiles)");
}

}  // namespace cf
