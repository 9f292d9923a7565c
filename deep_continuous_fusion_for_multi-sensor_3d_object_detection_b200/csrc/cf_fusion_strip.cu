// cf_fusion_strip.cu -- K-4 for the fine scales (C = 32 / 64): the fused MLP + K-sum-pool + BEV add as a STRIP
// pipeline.  The BEV map of a scale is walked in strips of S consecutive cells x all C channels.  A strip travels
//     global --(cp.async.bulk, one per channel row)--> shared memory --(+= MLP result of its live cells)--> global
// entirely through the bulk-copy engine: neither the cells without a neighbour (2/3 of scale 1) nor the BEV rows of the
// live cells ever pass through the LSU / register file, and there is no separate compaction pass -- the strip's
// neighbour indices arrive by bulk copy as well and are compacted in shared memory by one warp.
//
// Warp roles (18 warps, one CTA per SM, no __syncthreads in the steady state; everything is mbarrier based):
//   warp 17  scheduler: issues the strip loads (ring of NB strip buffers), compacts the live cells of every loaded
//            strip into tiles of <= 128 cells, publishes the GROUP STREAM (ring of 32-bit descriptors: a group = up to
//            G neighbour slots of one tile), issues the strip stores.
//   warp 16  issuer:    lane 0 issues the tcgen05.mma batches (descriptors precomputed: ~3 instructions per MMA).
//   warps 0-15 workers: per slot build one A operand tile (gather T rows -- prefetched a whole group ahead into
//            registers --, subtract the cell term, ReLU, split into bf16 hi / lo) straight into a ring of UMMA operand
//            buffers; every slot of a group has its OWN TMEM accumulator, so nothing waits for an MMA between slots.
//            One group later the workers read the group's accumulators back in one go: relu + K-pool in registers,
//            the pooled tile becomes the layer-3 operand (same ring), and after the next group's builds the layer-3
//            result is added into the strip buffer.
//
// Per tile the workers meet the tensor pipe twice (group read-back, layer-3 read-back) instead of once per neighbour
// slot, the gathers of the next group are in flight under the read-back of the previous one, and the pooled sum never
// round-trips through TMEM.  Arithmetic and accumulation order are those of k_fusion_tc (cf_mlp_tc.cu): results are
// bit-identical to it.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "cf_common.cuh"
#include "cf_tcgen05.cuh"

namespace cf {

namespace {

constexpr int kTile = 128;
#ifndef CF_STRIP_NW32
#define CF_STRIP_NW32 8
#endif
constexpr int strip_threads(int NW) { return (NW + 2) * 32; }   // NW worker warps + the issuer warp + the scheduler warp
constexpr int kRing = 256;      // group descriptors (power of two)
constexpr int kMaxNB = 4;       // strip buffers

// group descriptor
//   [1:0] kind   [4:2] strip slot   [5] tile in strip   [9:6] first neighbour slot k0   [12:10] slots - 1
//   [13] first group of the tile   [14] last group of the tile (its layer 3 follows)   [15] last tile of the strip
//   [22:16] rows - 1   [23] parity of the strip's load barrier
enum : uint32_t { kBubble = 0, kGroup = 1, kEnd = 3 };

struct StripParams {
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
    int32_t cells;
    int32_t strips_per_frame, strips_total;
    float x0, y0, dx, dy;
    int32_t inplace;
    int32_t nb;          // strip buffers in use (2 .. kMaxNB)
    int32_t strip_bytes; // bytes of one strip buffer: C * S * 4 + round16(S * K * 4)
    int32_t debug;       // CF_STRIP_DEBUG bits (timing experiments only; results are wrong when set)
};

// the "row" a (cell, k) slot without a neighbour gathers: relu(-1e30 - e) = 0
__device__ __align__(32) float g_strip_neg_row[64] = {
    -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f,
    -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f,
    -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f,
    -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f};

template <int C, int NS, int S, int G, int NA>
struct StripLayout {
    static constexpr int kWLayer = NS * C * C * 2;                   // one layer's packed image (K-chunk = C: resident)
    static constexpr int kOffWb = 2 * kWLayer;                       // bias B operands: layer 2 | layer 3, C rows x 32 B
    static constexpr int kWbBytes = C * 32;
    static constexpr int kOffAb = kOffWb + 2 * kWbBytes;             // bias A operands (128 rows x 16 bf16): ones | count[2]
    static constexpr int kAbBytes = kTile * 32;
    static constexpr int kOffA = kOffAb + 3 * kAbBytes;              // operand ring: NA x [hi | lo]
    static constexpr int kASlot = NS * kTile * C * 2;
    static constexpr int kOffW1 = kOffA + NA * kASlot;               // negated offset weights, 2 * C floats
    static constexpr int kOffRing = kOffW1 + 2 * C * 4;              // group descriptors
    static constexpr int kOffBar = kOffRing + kRing * 4;             // mbarriers + small state (512 B)
    static constexpr int kOffPos = kOffBar + 512;                    // uint16 pos[kMaxNB][S]
    static constexpr int kOffNv = kOffPos + kMaxNB * S * 2;          // uint8 nvalid[kMaxNB][S]
    static constexpr int kOffStrip = (kOffNv + kMaxNB * S + 127) / 128 * 128;
    static constexpr int kTmemCols = 2 * G * C + 2 * C <= 256 ? 256 : 512;   // two accumulator sets of G + two layer-3 accumulators
    static_assert(2 * G * C + 2 * C <= 512, "accumulators exceed the tensor memory");
    static __host__ __device__ constexpr int strip_bytes(int K) { return C * S * 4 + (S * K * 4 + 15) / 16 * 16; }
    static __host__ __device__ constexpr int smem_bytes(int K, int nb) { return kOffStrip + nb * strip_bytes(K); }
};

// barrier block (offsets in units of 8 bytes from kOffBar)
constexpr int kBarFull = 0;        // [8]  operand written            (16 arrivals: one per worker warp)
constexpr int kBarEmpty = 8;       // [8]  operand consumed           (tcgen05.commit)
constexpr int kBarGroup = 16;      // [2]  accumulator set complete   (tcgen05.commit)
constexpr int kBarAccFree = 18;    // [2]  accumulator set read back  (16 arrivals)
constexpr int kBarL3 = 20;         // [2]  layer-3 accumulator done   (tcgen05.commit)
constexpr int kBarLoad = 22;       // [4]  strip landed               (expect_tx)
constexpr int kBarStore = 26;      // [4]  strip finished             (16 arrivals)
constexpr int kStateTmem = 30 * 8;     // uint32 tmem base
constexpr int kStatePub = 30 * 8 + 4;  // uint32 groups published
constexpr int kStateInfo = 32 * 8;     // int4 sinfo[kMaxNB]: (frame, cell0, n_live, n_tiles)

// ---- small PTX helpers local to this kernel ------------------------------------------------------------------------
// A wait that lasts millions of polls is a protocol bug, not a slow step: report and trap instead of hanging the GPU.
__device__ __forceinline__ void spin_guard(uint32_t &it, const char *what)
{
    if (++it == (1u << 24)) {
        printf("k_fusion_strip: block %d warp %d stuck waiting for %s\n", blockIdx.x, (int)(threadIdx.x >> 5), what);
        __trap();
    }
}
__device__ __forceinline__ void mbar_wait_g(uint32_t addr, uint32_t parity, const char *what)
{
    uint32_t done, it = 0;
    for (;;) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) break;
        spin_guard(it, what);
    }
}
__device__ __forceinline__ bool mbar_test(uint32_t addr, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    return done != 0;
}
// warp-uniform variant: lane 0 tests, every lane gets its answer (a per-lane test may flip between lanes)
__device__ __forceinline__ bool mbar_test_warp(uint32_t addr, uint32_t parity)
{
    int done = 0;
    if ((threadIdx.x & 31) == 0) done = mbar_test(addr, parity);
    return __shfl_sync(0xffffffffu, done, 0) != 0;
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t addr)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t addr, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_init_a(uint32_t addr, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void commit_a(uint32_t addr)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(uint32_t addr, uint32_t v)
{
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// bulk copy global -> shared, completion on an mbarrier (bytes: multiple of 16, both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t pol)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void bulk_load_nohint(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
// bulk copy shared -> global, completion through the thread's bulk async-group
__device__ __forceinline__ void bulk_store(void *dst, uint32_t src, uint32_t bytes, uint64_t pol)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src), "r"(bytes),
                 "l"(pol)
                 : "memory");
}
// tensor-map TMA, 2-D tile (x = cell within the map's rows, y = row = frame * C + channel): ONE instruction moves a whole
// strip (box = S cells x C rows); rows that reach past the end of the map are clipped by the hardware
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, int32_t x, int32_t y, uint32_t bar, uint64_t pol)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst),
        "l"(tm), "r"(x), "r"(y), "r"(bar), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *tm, int32_t x, int32_t y, uint32_t src, uint64_t pol)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;" ::"l"(tm), "r"(x), "r"(y),
                 "r"(src), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ float lds_f32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr)
{
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// TMEM -> registers: n accumulators x CS consecutive fp32 columns of this warp's 32 lanes, one wait
__device__ __forceinline__ void tmem_ld8x1(uint32_t t0, uint32_t (&r)[8])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(t0)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8x2(uint32_t t0, uint32_t t1, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%16];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8, %9, %10, %11, %12, %13, %14, %15}, [%17];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(t0), "r"(t1)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8x3(uint32_t t0, uint32_t t1, uint32_t t2, uint32_t (&r)[24])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%24];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8, %9, %10, %11, %12, %13, %14, %15}, [%25];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%16, %17, %18, %19, %20, %21, %22, %23}, [%26];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23])
        : "r"(t0), "r"(t1), "r"(t2)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16x1(uint32_t t0, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(t0)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16x2(uint32_t t0, uint32_t t1, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%32];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%33];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(t0), "r"(t1)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16x3(uint32_t t0, uint32_t t1, uint32_t t2, uint32_t (&r)[48])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%48];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%49];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47}, [%50];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47])
        : "r"(t0), "r"(t1), "r"(t2)
        : "memory");
}

template <int C, int NS, int S, int G, int NA, int NW>
__global__ void __launch_bounds__(strip_threads(NW), 1) k_fusion_strip(const StripParams p, const __grid_constant__ CUtensorMap tm_in,
                                                                       const __grid_constant__ CUtensorMap tm_out)
{
    using L = StripLayout<C, NS, S, G, NA>;
    static_assert(NW == 8 || NW == 16, "worker warps");
    constexpr int kWorkers = NW, kIssuerWarp = NW, kSchedWarp = NW + 1, kThreads = strip_threads(NW);
    constexpr int RG = 16 / NW;          // 8-row groups per worker warp
    static_assert(C == 32 || C == 64, "strip kernel: C = 32 or 64");
    static_assert(C / (NW / 4) == 8 || C / (NW / 4) == 16, "accumulator columns per worker");
    static_assert(S % 32 == 0 && S <= 256 && G >= 1 && G <= 8 && NA >= 2 && NA <= 8, "strip shape");
    constexpr int kc_units = C / 8;
    constexpr int CS = C / (NW / 4);     // accumulator columns per worker in the read-backs
    constexpr int kItems = C / 32;       // (8 rows x 4 units) operand items per row group and slot
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = tc::smem_u32(smem);
    const uint32_t sW = sbase, sWb = sbase + L::kOffWb, sAb = sbase + L::kOffAb, sA = sbase + L::kOffA, sW1 = sbase + L::kOffW1;
    const uint32_t sRing = sbase + L::kOffRing, sBar = sbase + L::kOffBar, sPos = sbase + L::kOffPos, sNv = sbase + L::kOffNv;
    const uint32_t sStrip = sbase + L::kOffStrip;
    const uint32_t sPub = sBar + kStatePub, sInfo = sBar + kStateInfo;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = p.K;
    const uint32_t strip_bytes = (uint32_t)p.strip_bytes;
    constexpr uint32_t kBevBytes = C * S * 4;   // the index block of a strip buffer follows its BEV block

    // ---- one-time setup ------------------------------------------------------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < 8; ++i) {
            mbar_init_a(sBar + (kBarFull + i) * 8, kWorkers);
            mbar_init_a(sBar + (kBarEmpty + i) * 8, 1);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init_a(sBar + (kBarLoad + i) * 8, 1);
            mbar_init_a(sBar + (kBarStore + i) * 8, kWorkers);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init_a(sBar + (kBarGroup + i) * 8, 1);
            mbar_init_a(sBar + (kBarAccFree + i) * 8, kWorkers);
            mbar_init_a(sBar + (kBarL3 + i) * 8, 1);
        }
        tc::sts_u32(sPub, 0u);
        tc::mbar_fence_init();
    }
    __syncwarp();
    if (warp == kIssuerWarp) tc::tmem_alloc(reinterpret_cast<uint32_t *>(smem + L::kOffBar + kStateTmem), L::kTmemCols);
    for (int o = tid * 16; o < L::kWLayer; o += kThreads * 16) {
        *reinterpret_cast<uint4 *>(smem + o) = __ldg(reinterpret_cast<const uint4 *>(p.wimg2 + o));
        *reinterpret_cast<uint4 *>(smem + L::kWLayer + o) = __ldg(reinterpret_cast<const uint4 *>(p.wimg3 + o));
    }
    for (int o = tid * 16; o < 2 * L::kWbBytes; o += kThreads * 16) *reinterpret_cast<uint4 *>(smem + L::kOffWb + o) = make_uint4(0, 0, 0, 0);
    // bias A operands: column pair 0 of row r = (1, 1) for layer 2 (every row gets b2; slots without a neighbour are masked in
    // the read-back), = the row's neighbour count for layer 3 (two buffers, by tile parity); the other 14 columns stay 0
    for (int o = tid * 16; o < 3 * L::kAbBytes; o += kThreads * 16)
        *reinterpret_cast<uint4 *>(smem + L::kOffAb + o) = make_uint4(o < L::kAbBytes && (o & 128) == 0 ? 0x3F803F80u : 0u, 0, 0, 0);
    for (int c = tid; c < C; c += kThreads) {
        float *swn = reinterpret_cast<float *>(smem + L::kOffW1);
        swn[(c >> 3) * 16 + (c & 7)] = -__ldg(p.W1 + (size_t)c * (p.Ci + 3) + p.Ci);
        swn[(c >> 3) * 16 + 8 + (c & 7)] = -__ldg(p.W1 + (size_t)c * (p.Ci + 3) + p.Ci + 1);
    }
    __syncthreads();
    for (int n = tid; n < 2 * C; n += kThreads) {
        const int layer = n / C, c = n - layer * C;
        const float bv = __ldg((layer ? p.b3 : p.b2) + c);
        const __nv_bfloat16 h = __float2bfloat16_rn(bv);
        const __nv_bfloat16 l = __float2bfloat16_rn(bv - __bfloat162float(h));
        const uint32_t packed = (uint32_t)__bfloat16_as_ushort(h) | ((uint32_t)__bfloat16_as_ushort(l) << 16);
        *reinterpret_cast<uint32_t *>(smem + L::kOffWb + layer * L::kWbBytes + tc::unit_offset(c, 0, 2)) = packed;
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *reinterpret_cast<const uint32_t *>(smem + L::kOffBar + kStateTmem);
    constexpr uint32_t kTmemL3 = 2 * G * C;   // first column of the two layer-3 accumulators

    if (warp == kSchedWarp) {
        // =================================================================================================================
        // scheduler
        // =================================================================================================================
        const uint64_t pol = policy_evict_first();
        const int32_t grid = (int32_t)gridDim.x;
        const int32_t n_mine = ((int32_t)p.strips_total - (int32_t)blockIdx.x + grid - 1) / grid;
        const int nb = p.nb;
        int32_t li = 0, pi = 0, si = 0;
        uint32_t pub = 0, store_par = 0;
        uint32_t drain_to = 0;   // the workers finish group g in iteration g + 2: the stream must reach this index
        bool end_sent = false;
        auto put = [&](uint32_t desc) {   // lane 0
            tc::sts_u32(sRing + (pub & (kRing - 1)) * 4, desc);
            ++pub;
            if ((desc & 3u) == kGroup) drain_to = pub + 2;
        };
        auto strip_geometry = [&](int32_t i, int32_t &b, int32_t &cell0, int32_t &len) {
            const int32_t g = (int32_t)blockIdx.x + i * grid;
            b = g / p.strips_per_frame;
            cell0 = (g - b * p.strips_per_frame) * S;
            len = min(S, p.cells - cell0);
        };
        uint32_t idle = 0;
        while (si < n_mine || !end_sent) {
            bool progressed = false;
            // ---- (1) strip loads: buffer li % nb is free once strip li - nb has been stored and its store has read the buffer
            if (li < n_mine && li - si < nb) {
                const int slot = li % nb;
                if (li >= nb) {
                    const int allowed = si - (li - nb + 1);   // stores committed after the one that must be complete
                    if (allowed <= 0) bulk_wait_read<0>();
                    else if (allowed == 1) bulk_wait_read<1>();
                    else if (allowed == 2) bulk_wait_read<2>();
                    else bulk_wait_read<3>();
                }
                int32_t b, cell0, len;
                strip_geometry(li, b, cell0, len);
                const uint32_t bar = sBar + (kBarLoad + slot) * 8;
                const uint32_t dst = sStrip + slot * strip_bytes;
                const uint32_t idx_bytes = (uint32_t)len * K * 4u;
                if (lane == 0) {
                    mbar_expect_tx(bar, kBevBytes + idx_bytes);   // a clipped box still counts in full
                    tma_load_2d(dst, &tm_in, cell0, b * C, bar, pol);
                    bulk_load_nohint(dst + kBevBytes, p.knn + ((size_t)b * p.cells + cell0) * K, idx_bytes, bar);
                }
                ++li;
                progressed = true;
            }
            // ---- (2) compaction of the next loaded strip + its groups --------------------------------------------------------
            if (pi < li) {
                const int slot = pi % nb;
                const uint32_t par = (uint32_t)(pi / nb) & 1u;
                if (mbar_test_warp(sBar + (kBarLoad + slot) * 8, par)) {
                    int32_t b, cell0, len;
                    strip_geometry(pi, b, cell0, len);
                    const uint32_t idx0 = sStrip + slot * strip_bytes + kBevBytes;
                    int32_t n_live = 0, rmax0 = 0, rmax1 = 0;
                    // neighbour counts of this lane's S / 32 cells first (independent shared-memory loads), then the ordered
                    // compaction: ballot + prefix per 32 cells
                    int32_t nvs[S / 32];
#pragma unroll
                    for (int i = 0; i < S / 32; ++i) nvs[i] = 0;
#ifdef CF_STRIP_NOCOMPACT
                    for (int k = 0; k < 0; ++k) {
#else
                    for (int k = 0; k < K; ++k) {
#endif
#pragma unroll
                        for (int i = 0; i < S / 32; ++i) {
                            const int c = i * 32 + lane;
                            nvs[i] += (c < len && (int32_t)tc::lds_u32(idx0 + (uint32_t)(c * K + k) * 4u) >= 0) ? 1 : 0;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < S / 32; ++i) {
                        const int c = i * 32 + lane, nv = nvs[i];
                        const bool live = nv > 0;   // neighbours fill the slots from the front
                        const unsigned bal = __ballot_sync(0xffffffffu, live);
                        const int32_t e = n_live + __popc(bal & ((1u << lane) - 1u));
                        if (live) {
                            asm volatile("st.shared.u16 [%0], %1;" ::"r"(sPos + (uint32_t)(slot * S + e) * 2u), "h"((uint16_t)c) : "memory");
                            asm volatile("st.shared.u8 [%0], %1;" ::"r"(sNv + (uint32_t)(slot * S + e)), "r"(nv) : "memory");
                            if (e < kTile) rmax0 = max(rmax0, nv); else rmax1 = max(rmax1, nv);
                        }
                        n_live += __popc(bal);
                    }
                    rmax0 = __reduce_max_sync(0xffffffffu, rmax0);
                    rmax1 = __reduce_max_sync(0xffffffffu, rmax1);
                    const int32_t n_tiles = (n_live + kTile - 1) / kTile;
                    if (lane == 0) {
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sInfo + slot * 16), "r"(b), "r"(cell0), "r"(n_live),
                                     "r"(n_tiles)
                                     : "memory");
                    }
                    __threadfence_block();
                    __syncwarp();
                    if (lane == 0 && n_tiles > 0) {
                        for (int t = 0; t < n_tiles; ++t) {
                            const int32_t R = t ? rmax1 : rmax0, rows = min(kTile, n_live - t * kTile);
                            const uint32_t common = ((uint32_t)slot << 2) | ((uint32_t)t << 5) | ((uint32_t)(rows - 1) << 16) | (par << 23) |
                                                    (t == n_tiles - 1 ? 1u << 15 : 0u);
                            for (int k0 = 0; k0 < R; k0 += G) {
                                const int ns = min(G, R - k0);
                                put(kGroup | common | ((uint32_t)k0 << 6) | ((uint32_t)(ns - 1) << 10) | (k0 == 0 ? 1u << 13 : 0u) |
                                    (k0 + ns >= R ? 1u << 14 : 0u));
                            }
                        }
                        st_release(sPub, pub);
                    }
                    ++pi;
                    progressed = true;
                }
            }
            if (!progressed || pi == n_mine) {
                if (lane == 0 && !end_sent) {
                    if (pi == n_mine) {
                        put(kBubble);   // drive the deferred read-backs of the last group (pool, then layer 3)
                        put(kBubble);
                        put(kEnd);
                        st_release(sPub, pub);
                        end_sent = true;
                    } else if (pub < drain_to) {
                        // every strip buffer waits for read-backs that nothing new pushes along (the workers have caught up
                        // with the loads): a bubble drives them
                        put(kBubble);
                        st_release(sPub, pub);
                    }
                }
                end_sent = __shfl_sync(0xffffffffu, (int)end_sent, 0) != 0;
            }
            // ---- (3) strip stores ----------------------------------------------------------------------------------------------
            if (si < pi) {
                const int slot = si % nb;
                int32_t b, cell0, len;
                strip_geometry(si, b, cell0, len);
                int32_t n_tiles;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(n_tiles) : "r"(sInfo + slot * 16 + 12) : "memory");
                bool ready = n_tiles == 0;
                if (!ready) ready = mbar_test_warp(sBar + (kBarStore + slot) * 8, (store_par >> slot) & 1u);
                if (ready) {
                    if (n_tiles) store_par ^= 1u << slot;
                    if (lane == 0 && !(p.inplace && n_tiles == 0)) tma_store_2d(&tm_out, cell0, b * C, sStrip + slot * strip_bytes, pol);
                    bulk_commit();   // one group per strip (possibly empty): keeps the wait_group arithmetic uniform
                    ++si;
                    progressed = true;
                }
            }
            if (progressed) {
                idle = 0;
            } else if (++idle == (1u << 24)) {
                if (lane == 0)
                    printf("k_fusion_strip: block %d scheduler stuck: li %d pi %d si %d of %d, pub %u end %d store_par %x\n", blockIdx.x, li,
                           pi, si, n_mine, pub, (int)end_sent, store_par);
                __syncwarp();
                __trap();
            }
        }
        bulk_wait_all();
    } else if (warp == kIssuerWarp) {
        // =================================================================================================================
        // issuer: descriptors are built once; an MMA costs an add or two and the instruction itself
        // =================================================================================================================
        if (lane == 0) {
            constexpr uint32_t idesc = tc::make_idesc_bf16(kTile, C);
            constexpr uint32_t sbo = kc_units * 128;
            constexpr uint64_t kSlotStep = L::kASlot >> 4, kALo = (kTile * C * 2) >> 4, kWLo = (C * C * 2) >> 4;
            const uint64_t dA0 = tc::make_desc(sA, 128, sbo);
            const uint64_t dW2 = tc::make_desc(sW, 128, sbo), dW3 = tc::make_desc(sW + L::kWLayer, 128, sbo);
            const uint64_t dWb2 = tc::make_desc(sWb, 128, 256), dWb3 = tc::make_desc(sWb + L::kWbBytes, 128, 256);
            const uint64_t dAb1 = tc::make_desc(sAb, 128, 256), dAb3 = tc::make_desc(sAb + L::kAbBytes, 128, 256);
            constexpr uint64_t kAbStep = L::kAbBytes >> 4;
            uint32_t seen = 0, r = 0, full_par = 0, free_par = 0, gcount = 0, tcount = 0;
            bool prev_last = false;
            auto issue = [&](uint64_t dab, uint64_t dwb, uint64_t dw, uint32_t tmem_acc) {   // one operand of the ring: acc = bias + A * W^T
                mbar_wait_g(sBar + (kBarFull + r) * 8, (full_par >> r) & 1u, "full (issuer)");
                full_par ^= 1u << r;
                tc::fence_after_sync();
                const uint64_t da = dA0 + r * kSlotStep;
                if (!(p.debug & 1)) tc::mma_bf16(tmem_acc, dab, dwb, idesc, 0u);
#pragma unroll
                for (int kk = 0; kk < C / 16; ++kk) {
                    if (p.debug & 1) break;
                    const uint64_t koff = kk * 16;   // 16 bf16 = two 16-byte k-units = 256 bytes
                    tc::mma_bf16(tmem_acc, da + koff, dw + koff, idesc, 1u);
                    if (NS == 2) {
                        tc::mma_bf16(tmem_acc, da + koff, dw + kWLo + koff, idesc, 1u);
                        tc::mma_bf16(tmem_acc, da + kALo + koff, dw + koff, idesc, 1u);
                    }
                }
                commit_a(sBar + (kBarEmpty + r) * 8);
                r = r + 1 == (uint32_t)NA ? 0 : r + 1;
            };
            for (uint32_t g = 0;; ++g) {
                if (seen <= g) {
                    uint32_t it = 0;
                    while ((seen = ld_acquire(sPub)) <= g) {
                        __nanosleep(100);   // a polling warp must not take issue slots from the scheduler it waits for
                        spin_guard(it, "groups (issuer)");
                    }
                }
                const uint32_t desc = tc::lds_u32(sRing + (g & (kRing - 1)) * 4);
                const uint32_t kind = desc & 3u;
                if (kind == kEnd) break;
                bool cur_last = false;
                if (kind == kGroup) {
                    const uint32_t set = gcount & 1u, ns = ((desc >> 10) & 7u) + 1u;
                    if (gcount >= 2) {
                        mbar_wait_g(sBar + (kBarAccFree + set) * 8, (free_par >> set) & 1u, "acc_free (issuer)");
                        free_par ^= 1u << set;
                    }
                    for (uint32_t j = 0; j < ns; ++j) issue(dAb1, dWb2, dW2, tmem_base + set * (G * C) + j * C);
                    commit_a(sBar + (kBarGroup + set) * 8);
                    ++gcount;
                    cur_last = (desc & (1u << 14)) != 0;
                }
                if (prev_last) {   // layer 3 of the previous tile: the workers write its operand after this iteration's builds
                    issue(dAb3 + (tcount & 1u) * kAbStep, dWb3, dW3, tmem_base + kTmemL3 + (tcount & 1u) * C);
                    commit_a(sBar + (kBarL3 + (tcount & 1u)) * 8);
                    ++tcount;
                }
                prev_last = cur_last;
            }
        }
    } else {
        // =================================================================================================================
        // workers
        // =================================================================================================================
        const int q = warp & 3, cs = warp >> 2;
        const int erow = q * 32 + lane;                               // read-backs: this thread's row == its TMEM lane
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int col0 = cs * CS;
        const int r8 = lane & 7, u4 = lane >> 3;
        const int brow = warp * RG * 8 + r8;                          // builds: RG row groups per warp, this lane's first row
        const uint32_t a_row_off = (uint32_t)(warp * RG * kc_units * 128) + (uint32_t)(u4 * 128) + (uint32_t)(r8 * 16);
        float pooled[CS];
#pragma unroll
        for (int i = 0; i < CS; ++i) pooled[i] = 0.f;
        float cx[RG], cy[RG];
#pragma unroll
        for (int i = 0; i < RG; ++i) cx[i] = cy[i] = 0.f;
        uint32_t seen = 0, r = 0, items = 0;
        uint32_t empty_par = 0, group_par = 0, l3_par = 0, gcount = 0, pcount = 0, tcount = 0, ltile = 0;
        uint32_t pd = kBubble;

        auto peek_published = [&]() { return __shfl_sync(0xffffffffu, ld_acquire(sPub), 0); };   // warp-uniform
        auto wait_published = [&](uint32_t n) {
            if (seen <= n) {
                uint32_t it = 0;
                while ((seen = peek_published()) <= n) {
                    __nanosleep(100);   // a polling warp must not take issue slots from the scheduler it waits for
                    spin_guard(it, "groups (worker)");
                }
            }
        };
        auto desc_at = [&](uint32_t n) { return tc::lds_u32(sRing + (n & (kRing - 1)) * 4); };
        auto acquire = [&]() {   // the ring slot this operand goes into was last read by the MMAs of operand items - NA
            if (items >= (uint32_t)NA) {
                mbar_wait_g(sBar + (kBarEmpty + r) * 8, (empty_par >> r) & 1u, "empty (worker)");
                empty_par ^= 1u << r;
            }
        };
        auto publish = [&]() {
            tc::fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(sBar + (kBarFull + r) * 8);
            r = r + 1 == (uint32_t)NA ? 0 : r + 1;
            ++items;
        };

        float tv[G][RG * kItems][8];
        bool tv_valid = false;
        // all neighbour rows of a group: this thread's (row, 8 channels) segments
        auto gather = [&](uint32_t d) {
            const int slot = (d >> 2) & 7, t = (d >> 5) & 1, k0 = (d >> 6) & 15, ns = ((d >> 10) & 7) + 1, rows = ((d >> 16) & 127) + 1;
            mbar_wait_g(sBar + (kBarLoad + slot) * 8, (d >> 23) & 1u, "strip load (worker)");   // complete long ago: visibility
            int32_t b;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(b) : "r"(sInfo + slot * 16));
            const float *Tb = p.T + (size_t)b * p.N * C + u4 * 8;
#pragma unroll
            for (int rg = 0; rg < RG; ++rg) {
                if (brow + rg * 8 < rows) {
                    const uint32_t pos = lds_u16(sPos + (uint32_t)(slot * S + t * kTile + brow + rg * 8) * 2u);
                    const uint32_t ibase = sStrip + slot * strip_bytes + kBevBytes + (pos * K + k0) * 4u;
#pragma unroll
                    for (int j = 0; j < G; ++j) {
                        if (j < ns) {
                            const int32_t pr = (p.debug & 2) ? -1 : (int32_t)tc::lds_u32(ibase + j * 4);
                            const float *src = pr >= 0 ? Tb + (size_t)pr * C : g_strip_neg_row + u4 * 8;
#pragma unroll
                            for (int it = 0; it < kItems; ++it) tc::ldg_nc_f32x8(src + it * 32, tv[j][rg * kItems + it]);
                        }
                    }
                }
            }
        };
        auto build = [&](const float (&t)[RG * kItems][8], uint32_t a_addr, int rows) {
#pragma unroll
            for (int rg = 0; rg < RG; ++rg) {
                if (brow + rg * 8 < rows) {
                    const float2 cxx = make_float2(cx[rg], cx[rg]), cyy = make_float2(cy[rg], cy[rg]);
#pragma unroll
                    for (int it = 0; it < kItems; ++it) {
                        const uint32_t wa = sW1 + (uint32_t)((it * 4 + u4) * 64);
                        const float4 x0 = tc::lds_f32x4(wa), x1 = tc::lds_f32x4(wa + 16), y0 = tc::lds_f32x4(wa + 32), y1 = tc::lds_f32x4(wa + 48);
                        const float(&tt)[8] = t[rg * kItems + it];
                        float2 v[4];
                        v[0] = tc::ffma2(make_float2(x0.x, x0.y), cxx, tc::ffma2(make_float2(y0.x, y0.y), cyy, make_float2(tt[0], tt[1])));
                        v[1] = tc::ffma2(make_float2(x0.z, x0.w), cxx, tc::ffma2(make_float2(y0.z, y0.w), cyy, make_float2(tt[2], tt[3])));
                        v[2] = tc::ffma2(make_float2(x1.x, x1.y), cxx, tc::ffma2(make_float2(y1.x, y1.y), cyy, make_float2(tt[4], tt[5])));
                        v[3] = tc::ffma2(make_float2(x1.z, x1.w), cxx, tc::ffma2(make_float2(y1.z, y1.w), cyy, make_float2(tt[6], tt[7])));
                        uint4 hi, lo;
                        tc::relu_split_bf16x8(v, hi, lo, NS == 2);
                        const uint32_t dst = a_addr + a_row_off + (uint32_t)(rg * kc_units * 128) + (uint32_t)(it * 4 * 128);
                        tc::sts_u32x4(dst, hi);
                        if (NS == 2) tc::sts_u32x4(dst + kTile * C * 2, lo);
                    }
                }
            }
        };
        // pooled (+)= relu(acc) for the `n` accumulators of a group starting at column address `ta` (stride C columns);
        // slot k of a row counts only if the row has a k-th neighbour (k < nvp): the accumulate is predicated per lane
        auto pool = [&](uint32_t ta, int k0, int n, int nvp) {
            constexpr int kB = 2;   // accumulators read back per TMEM wait
            uint32_t z[kB * CS];
#pragma unroll 1
            for (int j = 0; j < n; j += kB) {
                const int m = min(kB, n - j);
                const uint32_t t0 = ta + j * C;
                if constexpr (CS == 8) {
                    if (m == 2) tmem_ld8x2(t0, t0 + C, z);
                    else tmem_ld8x1(t0, reinterpret_cast<uint32_t(&)[8]>(z));
                } else {
                    if (m == 2) tmem_ld16x2(t0, t0 + C, z);
                    else tmem_ld16x1(t0, reinterpret_cast<uint32_t(&)[16]>(z));
                }
#pragma unroll
                for (int a = 0; a < kB; ++a) {
                    if (a < m) {
                        if (k0 + j + a == 0) {   // every row of a tile has a first neighbour
#pragma unroll
                            for (int i = 0; i < CS; ++i) pooled[i] = fmaxf(__uint_as_float(z[a * CS + i]), 0.f);
                        } else if (k0 + j + a < nvp) {
#pragma unroll
                            for (int i = 0; i < CS; i += 2) {
                                const float2 s = tc::fadd2(make_float2(pooled[i], pooled[i + 1]),
                                                           make_float2(fmaxf(__uint_as_float(z[a * CS + i]), 0.f),
                                                                       fmaxf(__uint_as_float(z[a * CS + i + 1]), 0.f)));
                                pooled[i] = s.x;
                                pooled[i + 1] = s.y;
                            }
                        }
                    }
                }
            }
        };

        wait_published(0);
        {
            const uint32_t d0 = desc_at(0);
            if ((d0 & 3u) == kGroup) {
                gather(d0);
                tv_valid = true;
            }
        }
        uint32_t l3_desc = 0;   // a tile whose layer-3 operand was published in the previous iteration (0: none)
        for (uint32_t g = 0;; ++g) {
            wait_published(g);
            const uint32_t d = desc_at(g), kind = d & 3u;
            if (kind == kEnd) break;   // the stream ends bubble, bubble, end: nothing is pending here
            // ---- B. this group's operands ---------------------------------------------------------------------------------------
            if (kind == kGroup) {
                const int slot = (d >> 2) & 7, t = (d >> 5) & 1, ns = ((d >> 10) & 7) + 1, rows = ((d >> 16) & 127) + 1;
                if (!tv_valid) gather(d);
                if (d & (1u << 13)) {   // first group of a tile: the cell centres of this lane's rows
                    int32_t cell0;
                    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(cell0) : "r"(sInfo + slot * 16 + 4));
#pragma unroll
                    for (int rg = 0; rg < RG; ++rg) {
                        if (brow + rg * 8 < rows) {
                            const uint32_t pos = lds_u16(sPos + (uint32_t)(slot * S + t * kTile + brow + rg * 8) * 2u);
                            const int32_t cell = cell0 + (int32_t)pos;
                            const int32_t i = (int32_t)((uint32_t)cell / (uint32_t)p.W), j = cell - i * p.W;
                            cx[rg] = __fadd_rn(p.x0, __fmul_rn((float)i, p.dx));
                            cy[rg] = __fadd_rn(p.y0, __fmul_rn((float)j, p.dy));
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    if (j < ns) {
                        acquire();
                        if (brow < rows && !(p.debug & 4)) build(tv[j], sA + r * L::kASlot, rows);
                        publish();
                    }
                }
                ++gcount;
            }
            // ---- C. the next group's neighbour rows: in flight under D, A and the loop turn -----------------------------------------
            tv_valid = false;
            if (seen <= g + 1) seen = peek_published();
            if (seen > g + 1) {
                const uint32_t d1 = desc_at(g + 1);
                if ((d1 & 3u) == kGroup) {
                    gather(d1);
                    tv_valid = true;
                }
            }
            // ---- D. layer-3 read-back of the tile whose operand went out one iteration ago: strip += acc ---------------------------
            if (l3_desc) {
                const int slot = (l3_desc >> 2) & 7, t = (l3_desc >> 5) & 1, rows = ((l3_desc >> 16) & 127) + 1;
                const uint32_t par = tcount & 1u;
                mbar_wait_g(sBar + (kBarL3 + par) * 8, (l3_par >> par) & 1u, "layer 3 done (worker)");
                l3_par ^= 1u << par;
                tc::fence_after_sync();
                if (q * 32 < rows && !(p.debug & 16)) {
                    uint32_t z[CS];
                    if constexpr (CS == 8) tmem_ld8x1(tmem_base + lane_off + kTmemL3 + par * C + col0, z);
                    else tmem_ld16x1(tmem_base + lane_off + kTmemL3 + par * C + col0, z);
                    if (erow < rows) {
                        const uint32_t pos = lds_u16(sPos + (uint32_t)(slot * S + t * kTile + erow) * 2u);
                        const uint32_t base = sStrip + slot * strip_bytes + (uint32_t)(col0 * S) * 4u + pos * 4u;
#pragma unroll
                        for (int i = 0; i < CS; ++i) {
                            const uint32_t addr = base + (uint32_t)(i * S) * 4u;
                            sts_f32(addr, lds_f32(addr) + __uint_as_float(z[i]));
                        }
                    }
                }
                tc::fence_before_sync();
                const bool strip_done = (l3_desc & (1u << 15)) != 0;
                if (strip_done) tc::fence_proxy_async();   // the strip buffer goes back out through the bulk-copy engine
                __syncwarp();
                if (lane == 0 && strip_done) mbar_arrive_a(sBar + (kBarStore + slot) * 8);
                ++tcount;
                l3_desc = 0;
            }
            // ---- A. read-back of the PREVIOUS group (its MMAs ran under this iteration's builds): relu + pool; after a tile's last
            //         group the pooled tile becomes the layer-3 operand ------------------------------------------------------------
            if ((pd & 3u) == kGroup) {
                const int slot = (pd >> 2) & 7, t = (pd >> 5) & 1, k0 = (pd >> 6) & 15, ns = ((pd >> 10) & 7) + 1, rows = ((pd >> 16) & 127) + 1;
                const uint32_t set = pcount & 1u;
                ++pcount;
                mbar_wait_g(sBar + (kBarGroup + set) * 8, (group_par >> set) & 1u, "group done (worker)");
                group_par ^= 1u << set;
                tc::fence_after_sync();
                const int nvp = erow < rows ? (int)lds_u8(sNv + (uint32_t)(slot * S + t * kTile + erow)) : 0;
                if (q * 32 < rows && !(p.debug & 8)) pool(tmem_base + lane_off + set * (G * C) + col0, k0, ns, nvp);
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive_a(sBar + (kBarAccFree + set) * 8);
                if (pd & (1u << 14)) {
                    acquire();
                    const uint32_t a_addr = sA + r * L::kASlot;
                    if (q * 32 < rows) {
#pragma unroll
                        for (int u = 0; u < CS / 8; ++u) {   // pooled >= 0: the fused ReLU is the identity
                            float2 v[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) v[i] = make_float2(pooled[u * 8 + 2 * i], pooled[u * 8 + 2 * i + 1]);
                            uint4 hi, lo;
                            tc::relu_split_bf16x8(v, hi, lo, NS == 2);
                            const uint32_t dst = a_addr + tc::unit_offset(erow, col0 / 8 + u, kc_units);
                            tc::sts_u32x4(dst, hi);
                            if (NS == 2) tc::sts_u32x4(dst + kTile * C * 2, lo);
                        }
                    }
                    if (cs == 0) {   // bias column = the row's neighbour count (small integers are exact in bf16)
                        const uint32_t nv16 = __float_as_uint((float)nvp) >> 16;
                        tc::sts_u32(sAb + (1 + (ltile & 1u)) * L::kAbBytes + tc::unit_offset(erow, 0, 2), nv16 | (nv16 << 16));
                    }
                    ++ltile;
                    publish();
                    l3_desc = pd;
                }
            }
            pd = d;
        }
    }

    tc::fence_before_sync();
    __syncthreads();
    if (warp == kIssuerWarp) tc::tmem_free(tmem_base, L::kTmemCols);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) f = nullptr;
        (void)cudaGetLastError();
        return (EncodeTiledFn)f;
    }();
    return fn;
}
// (B * C rows) x (cells) fp32 view of a BEV map, box = S cells x C rows
int make_strip_map(CUtensorMap *tm, const float *base, int64_t rows, int64_t cells, int S, int C)
{
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return CF_ERR_UNSUPPORTED;
    const cuuint64_t dims[2] = {(cuuint64_t)cells, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cells * 4};
    const cuuint32_t box[2] = {(cuuint32_t)S, (cuuint32_t)C};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? CF_OK : CF_ERR_UNSUPPORTED;
}

template <int C, int NS, int S, int G, int NA, int NW>
int launch_strip(StripParams &p, cudaStream_t st)
{
    using L = StripLayout<C, NS, S, G, NA>;
    const int sb = L::strip_bytes(p.K);
    int nb = kMaxNB;
    while (nb > 2 && L::smem_bytes(p.K, nb) > 227 * 1024) --nb;
    const int smem = L::smem_bytes(p.K, nb);
    if (smem > 227 * 1024) return CF_ERR_UNSUPPORTED;
    static int attr_bytes = 0;
    if (smem > attr_bytes) {
        CF_TRY(cuda_status(cudaFuncSetAttribute(k_fusion_strip<C, NS, S, G, NA, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                           "k_fusion_strip smem attribute"));
        attr_bytes = smem;
    }
    p.nb = nb;
    p.strip_bytes = sb;
    p.strips_per_frame = (p.cells + S - 1) / S;
    p.strips_total = p.strips_per_frame * p.B;
    CUtensorMap tm_in, tm_out;
    CF_TRY(make_strip_map(&tm_in, p.bev, (int64_t)p.B * C, p.cells, S, C));
    CF_TRY(make_strip_map(&tm_out, p.out, (int64_t)p.B * C, p.cells, S, C));
    const int grid = (int)std::min<int64_t>(p.strips_total, sm_count());
    k_fusion_strip<C, NS, S, G, NA, NW><<<grid, strip_threads(NW), smem, st>>>(p, tm_in, tm_out);
    return CF_OK;
}

}  // namespace

// K-4 through the strip pipeline.  Returns CF_ERR_UNSUPPORTED (nothing launched, no error text) when the shape or the
// alignment does not fit: C not 32 / 64, H * W not a multiple of 4, a BEV / index pointer that is not 16-byte aligned.
int fusion_strip(const float *d_bev, const float *d_T, const int32_t *d_knn, int32_t B, int32_t N, int32_t C, int32_t H,
                 int32_t W, int32_t K, float x0, float y0, float dx, float dy, const float *d_W1, int32_t Ci,
                 const float *d_b2, const float *d_b3, float *d_out, int32_t mode, const uint8_t *img2, const uint8_t *img3,
                 cudaStream_t st)
{
    const int64_t cells = (int64_t)H * W;
    if ((C != 32 && C != 64) || cells % 4 != 0 || cells > (int64_t)1 << 30) return CF_ERR_UNSUPPORTED;
    if (!aligned16(d_bev) || !aligned16(d_out) || !aligned16(d_knn) || !aligned16(img2) || !aligned16(img3)) return CF_ERR_UNSUPPORTED;
    StripParams p;
    p.bev = d_bev; p.T = d_T; p.knn = d_knn; p.out = d_out; p.wimg2 = img2; p.wimg3 = img3; p.W1 = d_W1; p.b2 = d_b2; p.b3 = d_b3;
    p.B = B; p.N = N; p.W = W; p.K = K; p.Ci = Ci; p.cells = (int32_t)cells;
    p.x0 = x0; p.y0 = y0; p.dx = dx; p.dy = dy;
    p.inplace = d_out == d_bev;
    static const int dbg = getenv("CF_STRIP_DEBUG") ? atoi(getenv("CF_STRIP_DEBUG")) : 0;
    p.debug = dbg;
    const int NS = mode == CF_MODE_FP32 ? 2 : 1;
    int rc;
    if (C == 32) rc = NS == 2 ? launch_strip<32, 2, 256, 5, 5, CF_STRIP_NW32>(p, st) : launch_strip<32, 1, 256, 5, 5, CF_STRIP_NW32>(p, st);
    else rc = NS == 2 ? launch_strip<64, 2, 128, 3, 3, 16>(p, st) : launch_strip<64, 1, 128, 3, 3, 16>(p, st);
    if (rc != CF_OK) return rc;
    count_launches(1);
    return launch_status("cf_fusion_fwd (strip pipeline)");
}

}  // namespace cf
