// cf_fusion_strip.cu -- K-4 for the fine scales (C = 32 / 64): the fused MLP + K-sum-pool + BEV add as a STRIP
// pipeline.  The BEV map of a scale is walked in strips of S consecutive cells x all C channels.  A strip travels
//     global --(cp.async.bulk, one per channel row)--> shared memory --(+= MLP result of its live cells)--> global
// entirely through the bulk-copy engine: neither the cells without a neighbour (2/3 of scale 1) nor the BEV rows of the
// live cells ever pass through the LSU / register file, and there is no separate compaction pass -- the strip's
// neighbour indices arrive by bulk copy as well and are compacted in shared memory by one warp.
//
// Warp roles (18 warps, one CTA per SM, no __syncthreads in the steady state; everything is mbarrier based):
//   warp 17  scheduler: issues the strip loads (ring of NB strip buffers), compacts the live cells of every loaded
//            strip, publishes the POSITION STREAM (ring of 32-bit descriptors), issues the strip stores.
//   warp 16  issuer:    lane 0 issues the tcgen05.mma batch of every position and commits it to an mbarrier.
//   warps 0-15 workers: per position build one A operand tile (gather T rows, subtract the cell term, ReLU, split
//            into bf16 hi / lo) straight into the UMMA layout; D positions later read the accumulator back from
//            TMEM: relu + pool in registers (slot positions) or add into the strip buffer (layer-3 positions).
//
// Position stream of a tile of <= 128 live cells with R neighbour rounds:   S0 S1 .. S(R-1)  [D-1 positions of the
// next tile]  L3  ...   Every position owns one A-ring slot (n % D) and one TMEM accumulator (n % NACC); its MMAs
// run while the workers build the following positions, so neither the MMA latency nor the L2 latency of the
// gathers (prefetched two positions ahead into registers) sits on a critical path, and the pooled sum never
// round-trips through TMEM.  Arithmetic and accumulation order are those of k_fusion_tc (cf_mlp_tc.cu): results are
// bit-identical to it.
#include <stdio.h>

#include "cf_common.cuh"
#include "cf_tcgen05.cuh"

namespace cf {

namespace {

constexpr int kTile = 128;
constexpr int kWorkers = 16;
constexpr int kIssuerWarp = kWorkers, kSchedWarp = kWorkers + 1;
constexpr int kThreads = (kWorkers + 2) * 32;
constexpr int kRing = 256;      // position descriptors (power of two)
constexpr int kNacc = 4;        // TMEM accumulators
constexpr int kMaxNB = 4;       // strip buffers

// position descriptor
//   [1:0] kind   [4:2] strip slot   [5] tile in strip   [9:6] k   [10] first slot   [11] last slot / last tile of strip
//   [18:12] rows - 1   [19] parity of the strip's load barrier
enum : uint32_t { kBubble = 0, kSlot = 1, kL3 = 2, kEnd = 3 };

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
};

// the "row" a (cell, k) slot without a neighbour gathers: relu(-1e30 - e) = 0
__device__ __align__(32) float g_strip_neg_row[64] = {
    -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f,
    -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f,
    -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f,
    -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f};

template <int C, int NS, int S, int D>
struct StripLayout {
    static constexpr int kWLayer = NS * C * C * 2;                   // one layer's packed image (K-chunk = C: resident)
    static constexpr int kOffWb = 2 * kWLayer;                       // bias B operands: layer 2 | layer 3, C rows x 32 B
    static constexpr int kWbBytes = C * 32;
    static constexpr int kOffA = kOffWb + 2 * kWbBytes;              // A ring: D x [hi | lo | bias-flag operand]
    static constexpr int kAData = NS * kTile * C * 2;
    static constexpr int kASlot = kAData + kTile * 32;
    static constexpr int kOffW1 = kOffA + D * kASlot;                // negated offset weights, 2 * C floats
    static constexpr int kOffRing = kOffW1 + 2 * C * 4;              // position descriptors
    static constexpr int kOffBar = kOffRing + kRing * 4;             // mbarriers + small state (512 B)
    static constexpr int kOffPos = kOffBar + 512;                    // uint16 pos[kMaxNB][S]
    static constexpr int kOffNv = kOffPos + kMaxNB * S * 2;          // uint8 nvalid[kMaxNB][S]
    static constexpr int kOffStrip = (kOffNv + kMaxNB * S + 127) / 128 * 128;
    static __host__ __device__ constexpr int strip_bytes(int K) { return C * S * 4 + (S * K * 4 + 15) / 16 * 16; }
    static __host__ __device__ constexpr int smem_bytes(int K, int nb) { return kOffStrip + nb * strip_bytes(K); }
};

// barrier block (offsets in units of 8 bytes from kOffBar)
constexpr int kBarFull = 0;        // [4]  A slot written           (16 arrivals: one per worker warp)
constexpr int kBarDone = 4;        // [4]  MMAs of a position done   (tcgen05.commit)
constexpr int kBarAccFree = 8;     // [4]  accumulator read back     (16 arrivals)
constexpr int kBarLoad = 12;       // [4]  strip landed              (expect_tx)
constexpr int kBarStore = 16;      // [4]  strip finished            (16 arrivals)
constexpr int kStateTmem = 20 * 8;     // uint32 tmem base
constexpr int kStatePub = 20 * 8 + 4;  // uint32 positions published
constexpr int kStateInfo = 22 * 8;     // int4 sinfo[kMaxNB]: (frame, cell0, n_live, n_tiles)

// ---- small PTX helpers local to this kernel ------------------------------------------------------------------------
__device__ __forceinline__ void spin_guard(uint32_t &it, long long &t0, const char *what)
{
    ++it;
    if (it == 4096) t0 = clock64();
    if (it > 4096 && (it & 4095) == 0 && clock64() - t0 > 6000000000ll) {   // ~3 s: a protocol bug, not a slow step
        printf("k_fusion_strip: block %d warp %d stuck waiting for %s\n", blockIdx.x, (int)(threadIdx.x >> 5), what);
        __trap();
    }
}
__device__ __forceinline__ void mbar_wait_g(uint32_t addr, uint32_t parity, const char *what)
{
    uint32_t done, it = 0;
    long long t0 = 0;
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
        spin_guard(it, t0, what);
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

// TMEM -> registers, 32 lanes x 8 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8])
{
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
template <int CS>
__device__ __forceinline__ void tmem_ld_cs(uint32_t taddr, float (&v)[CS])
{
    if constexpr (CS == 8) {
        tmem_ld8(taddr, v);
    } else {
        static_assert(CS == 16, "columns per worker");
        tc::tmem_ld16(taddr, v);
    }
}

// the MMAs of one position: acc = flag * bias + A[128 x C] * W[C x C]^T (all split products), fresh accumulator
template <int C, int NS>
__device__ __forceinline__ void issue_position(uint32_t a_addr, uint32_t ab_addr, uint32_t w_addr, uint32_t wb_addr, uint32_t tmem_acc)
{
    constexpr uint32_t idesc = tc::make_idesc_bf16(kTile, C);
    constexpr uint32_t sbo = (C / 8) * 128, lbo = 128;
    constexpr uint32_t a_split = kTile * C * 2, w_split = C * C * 2;
    tc::mma_bf16(tmem_acc, tc::make_desc(ab_addr, 128, 256), tc::make_desc(wb_addr, 128, 256), idesc, 0u);
#pragma unroll
    for (int kk = 0; kk < C / 16; ++kk) {
        const uint32_t koff = kk * 2 * lbo;
        const uint64_t a_hi = tc::make_desc(a_addr + koff, lbo, sbo);
        const uint64_t w_hi = tc::make_desc(w_addr + koff, lbo, sbo);
        tc::mma_bf16(tmem_acc, a_hi, w_hi, idesc, 1u);
        if (NS == 2) {
            const uint64_t a_lo = tc::make_desc(a_addr + a_split + koff, lbo, sbo);
            const uint64_t w_lo = tc::make_desc(w_addr + w_split + koff, lbo, sbo);
            tc::mma_bf16(tmem_acc, a_hi, w_lo, idesc, 1u);
            tc::mma_bf16(tmem_acc, a_lo, w_hi, idesc, 1u);
        }
    }
}

template <int C, int NS, int S, int D>
__global__ void __launch_bounds__(kThreads, 1) k_fusion_strip(const StripParams p)
{
    using L = StripLayout<C, NS, S, D>;
    static_assert(C == 32 || C == 64, "strip kernel: C = 32 or 64");
    static_assert(S % 32 == 0 && S <= 256 && D >= 1 && D <= 3 && D <= kNacc, "strip shape");
    constexpr int kc_units = C / 8;
    constexpr int CS = C / 4;            // accumulator columns per worker in the epilogues
    constexpr int kItems = C / 32;       // (8 rows x 4 units) operand items per worker warp and position
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = tc::smem_u32(smem);
    const uint32_t sW = sbase, sWb = sbase + L::kOffWb, sA = sbase + L::kOffA, sW1 = sbase + L::kOffW1;
    const uint32_t sRing = sbase + L::kOffRing, sBar = sbase + L::kOffBar, sPos = sbase + L::kOffPos, sNv = sbase + L::kOffNv;
    const uint32_t sStrip = sbase + L::kOffStrip;
    const uint32_t sPub = sBar + kStatePub, sInfo = sBar + kStateInfo;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = p.K;
    const uint32_t strip_bytes = (uint32_t)p.strip_bytes;
    constexpr uint32_t kBevBytes = C * S * 4;   // the index block of a strip buffer follows its BEV block

    // ---- one-time setup ------------------------------------------------------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) {
            mbar_init_a(sBar + (kBarFull + i) * 8, kWorkers);
            mbar_init_a(sBar + (kBarDone + i) * 8, 1);
            mbar_init_a(sBar + (kBarAccFree + i) * 8, kWorkers);
            mbar_init_a(sBar + (kBarLoad + i) * 8, 1);
            mbar_init_a(sBar + (kBarStore + i) * 8, kWorkers);
        }
        tc::sts_u32(sPub, 0u);
        tc::mbar_fence_init();
    }
    __syncwarp();
    if (warp == kIssuerWarp) tc::tmem_alloc(reinterpret_cast<uint32_t *>(smem + L::kOffBar + kStateTmem), kNacc * C);
    for (int o = tid * 16; o < L::kWLayer; o += kThreads * 16) {
        *reinterpret_cast<uint4 *>(smem + o) = __ldg(reinterpret_cast<const uint4 *>(p.wimg2 + o));
        *reinterpret_cast<uint4 *>(smem + L::kWLayer + o) = __ldg(reinterpret_cast<const uint4 *>(p.wimg3 + o));
    }
    for (int o = tid * 16; o < 2 * L::kWbBytes; o += kThreads * 16) *reinterpret_cast<uint4 *>(smem + L::kOffWb + o) = make_uint4(0, 0, 0, 0);
    for (int o = tid * 16; o < D * kTile * 32; o += kThreads * 16) {   // bias-flag operands: only column pair 0 is ever rewritten
        const int slot = o / (kTile * 32), r = o - slot * (kTile * 32);
        *reinterpret_cast<uint4 *>(smem + L::kOffA + slot * L::kASlot + L::kAData + r) = make_uint4(0, 0, 0, 0);
    }
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
        // pending layer-3 positions (at most D - 1 <= 2): descriptor + positions still to pass
        uint32_t pd0 = 0, pd1 = 0;
        int pc0 = -1, pc1 = -1;
        bool end_sent = false;
        auto put = [&](uint32_t desc) {   // lane 0
            tc::sts_u32(sRing + (pub & (kRing - 1)) * 4, desc);
            ++pub;
        };
        auto tick = [&]() {   // one position has been emitted: age the pending layer-3 positions, emit the ones that are due
            if (pc0 >= 0) --pc0;
            if (pc1 >= 0) --pc1;
            while (pc0 == 0) {
                put(pd0);
                pd0 = pd1; pc0 = pc1; pc1 = -1;
                if (pc0 > 0) --pc0;
            }
        };
        auto emit = [&](uint32_t desc) {
            put(desc);
            tick();
        };
        auto add_pending = [&](uint32_t desc) {
            if (D == 1) {
                put(desc);
                return;
            }
            if (pc0 < 0) { pd0 = desc; pc0 = D - 1; }
            else { pd1 = desc; pc1 = D - 1; }
        };
        auto strip_geometry = [&](int32_t i, int32_t &b, int32_t &cell0, int32_t &len) {
            const int32_t g = (int32_t)blockIdx.x + i * grid;
            b = g / p.strips_per_frame;
            cell0 = (g - b * p.strips_per_frame) * S;
            len = min(S, p.cells - cell0);
        };
        uint32_t idle = 0;
        long long t_idle = 0;
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
                const uint32_t row_bytes = (uint32_t)len * 4u, idx_bytes = (uint32_t)len * K * 4u;
                if (lane == 0) mbar_expect_tx(bar, C * row_bytes + idx_bytes);
                __syncwarp();
                for (int c = lane; c < C; c += 32)
                    bulk_load(dst + c * S * 4, p.bev + ((size_t)b * C + c) * p.cells + cell0, row_bytes, bar, pol);
                if (lane == 0) bulk_load_nohint(dst + kBevBytes, p.knn + ((size_t)b * p.cells + cell0) * K, idx_bytes, bar);
                ++li;
                progressed = true;
            }
            // ---- (2) compaction of the next loaded strip + its positions ----------------------------------------------------
            if (pi < li) {
                const int slot = pi % nb;
                const uint32_t par = (uint32_t)(pi / nb) & 1u;
                if (mbar_test(sBar + (kBarLoad + slot) * 8, par)) {
                    int32_t b, cell0, len;
                    strip_geometry(pi, b, cell0, len);
                    const uint32_t idx0 = sStrip + slot * strip_bytes + kBevBytes;
                    int32_t n_live = 0, rmax0 = 0, rmax1 = 0;
#pragma unroll 1
                    for (int c0 = 0; c0 < S; c0 += 32) {
                        const int c = c0 + lane;
                        const bool live = c < len && (int32_t)tc::lds_u32(idx0 + (uint32_t)(c * K) * 4u) >= 0;
                        int32_t nv = 0;
                        if (live)
                            for (int k = 0; k < K; ++k) nv += (int32_t)tc::lds_u32(idx0 + (uint32_t)(c * K + k) * 4u) >= 0;
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
                    if (lane == 0) {
                        for (int t = 0; t < n_tiles; ++t) {
                            const int32_t R = t ? rmax1 : rmax0, rows = min(kTile, n_live - t * kTile);
                            const uint32_t common = ((uint32_t)slot << 2) | ((uint32_t)t << 5) | ((uint32_t)(rows - 1) << 12) | (par << 19);
                            for (int k = 0; k < R; ++k) {
                                if (k == R - 1) {   // the tile's layer-3 position follows D - 1 positions after its last slot
                                    put(kSlot | common | ((uint32_t)k << 6) | (k == 0 ? 1u << 10 : 0u) | (1u << 11));
                                    add_pending(kL3 | common | (t == n_tiles - 1 ? 1u << 11 : 0u));
                                    if (D > 1) tick();
                                } else {
                                    emit(kSlot | common | ((uint32_t)k << 6) | (k == 0 ? 1u << 10 : 0u));
                                }
                            }
                        }
                        st_release(sPub, pub);
                    }
                    ++pi;
                    progressed = true;
                }
            }
            // pending layer-3 positions are flushed with bubbles when nothing else can be published right now
            if (!progressed || pi == n_mine) {
                if (lane == 0) {
                    if (pc0 >= 0) {
                        emit(kBubble);
                        st_release(sPub, pub);
                    } else if (pi == n_mine && !end_sent) {
                        for (int i = 0; i < D; ++i) put(kBubble);   // drain the deferred epilogues of the last positions
                        put(kEnd);
                        st_release(sPub, pub);
                        end_sent = true;
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
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(n_tiles) : "r"(sInfo + slot * 16 + 12));
                bool ready = n_tiles == 0;
                if (!ready) ready = mbar_test(sBar + (kBarStore + slot) * 8, (store_par >> slot) & 1u);
                if (ready) {
                    if (n_tiles) store_par ^= 1u << slot;
                    if (!(p.inplace && n_tiles == 0)) {
                        const uint32_t src = sStrip + slot * strip_bytes;
                        for (int c = lane; c < C; c += 32)
                            bulk_store(p.out + ((size_t)b * C + c) * p.cells + cell0, src + c * S * 4, (uint32_t)len * 4u, pol);
                    }
                    bulk_commit();   // one group per strip and lane (possibly empty): keeps the wait_group arithmetic uniform
                    ++si;
                    progressed = true;
                }
            }
            if (progressed) idle = 0; else spin_guard(idle, t_idle, "strips (scheduler)");
        }
        bulk_wait_all();
    } else if (warp == kIssuerWarp) {
        // =================================================================================================================
        // issuer
        // =================================================================================================================
        if (lane == 0) {
            uint32_t seen = 0;
            uint32_t a = 0, acc = 0, full_par = 0, free_par = 0;   // ring slot, accumulator, their phase bits
            for (uint32_t n = 0;; ++n) {
                if (seen <= n) {
                    uint32_t it = 0;
                    long long t0 = 0;
                    while ((seen = ld_acquire(sPub)) <= n) spin_guard(it, t0, "positions (issuer)");
                }
                const uint32_t desc = tc::lds_u32(sRing + (n & (kRing - 1)) * 4);
                const uint32_t kind = desc & 3u;
                if (kind == kEnd) break;
                mbar_wait_g(sBar + (kBarFull + a) * 8, (full_par >> a) & 1u, "full (issuer)");
                full_par ^= 1u << a;
                if (n >= (uint32_t)kNacc) {
                    mbar_wait_g(sBar + (kBarAccFree + acc) * 8, (free_par >> acc) & 1u, "acc_free (issuer)");
                    free_par ^= 1u << acc;
                }
                tc::fence_after_sync();
                if (kind != kBubble) {
                    const uint32_t a_addr = sA + a * L::kASlot;
                    const bool l3 = kind == kL3;
                    issue_position<C, NS>(a_addr, a_addr + L::kAData, sW + (l3 ? L::kWLayer : 0), sWb + (l3 ? L::kWbBytes : 0),
                                          tmem_base + acc * C);
                }
                commit_a(sBar + (kBarDone + acc) * 8);
                a = a + 1 == (uint32_t)D ? 0 : a + 1;
                acc = (acc + 1) & (kNacc - 1);
            }
        }
    } else {
        // =================================================================================================================
        // workers
        // =================================================================================================================
        const int q = warp & 3, cs = warp >> 2;
        const int erow = q * 32 + lane;                               // epilogue: this thread's row == its TMEM lane
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int col0 = cs * CS;
        const int r8 = lane & 7, u4 = lane >> 3;
        const int brow = warp * 8 + r8;                               // build: row group = warp
        const uint32_t a_row_off = (uint32_t)(warp * kc_units * 128) + (uint32_t)(r8 * 16);
        float pooled[CS];
#pragma unroll
        for (int i = 0; i < CS; ++i) pooled[i] = 0.f;
        float cx = 0.f, cy = 0.f;
        uint32_t seen = 0;
        uint32_t done_par = 0;                                         // phase bits of the mma_done barriers
        uint32_t a = 0;                                                // A-ring slot of position n

        auto wait_published = [&](uint32_t n) {
            if (seen <= n) {
                uint32_t it = 0;
                long long t0 = 0;
                while ((seen = ld_acquire(sPub)) <= n) spin_guard(it, t0, "positions (worker)");
            }
        };
        auto desc_at = [&](uint32_t n) { return tc::lds_u32(sRing + (n & (kRing - 1)) * 4); };

        // gather of one slot position: this thread's (row, 8 channels) segments of the neighbour's T row
        auto gather = [&](uint32_t desc, float (&tv)[kItems][8]) {
            const int slot = (desc >> 2) & 7, t = (desc >> 5) & 1, k = (desc >> 6) & 15, rows = ((desc >> 12) & 127) + 1;
            mbar_wait_g(sBar + (kBarLoad + slot) * 8, (desc >> 19) & 1u, "strip load (worker)");   // already complete: visibility
            if (brow < rows) {
                const uint32_t pos = lds_u16(sPos + (uint32_t)(slot * S + t * kTile + brow) * 2u);
                const int32_t pr = (int32_t)tc::lds_u32(sStrip + slot * strip_bytes + kBevBytes + (pos * K + k) * 4u);
                int32_t b;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(b) : "r"(sInfo + slot * 16));
                const float *src = pr >= 0 ? p.T + ((size_t)b * p.N + pr) * C + u4 * 8 : g_strip_neg_row + u4 * 8;
#pragma unroll
                for (int it = 0; it < kItems; ++it) tc::ldg_nc_f32x8(src + it * 32, tv[it]);
            }
        };

        // deferred read-back of position m (its MMAs were committed D positions ago)
        auto epilogue = [&](uint32_t m) {
            const uint32_t pd = desc_at(m);
            const uint32_t kind = pd & 3u, acc = m & (kNacc - 1);
            mbar_wait_g(sBar + (kBarDone + acc) * 8, (done_par >> acc) & 1u, "mma_done (worker)");
            done_par ^= 1u << acc;
            tc::fence_after_sync();
            if (kind == kSlot) {
                float z[CS];
                tmem_ld_cs<CS>(tmem_base + lane_off + acc * C + col0, z);
                if (pd & (1u << 10)) {
#pragma unroll
                    for (int i = 0; i < CS; ++i) pooled[i] = fmaxf(z[i], 0.f);
                } else {
#pragma unroll
                    for (int i = 0; i < CS; i += 2) {
                        const float2 s = tc::fadd2(make_float2(pooled[i], pooled[i + 1]), make_float2(fmaxf(z[i], 0.f), fmaxf(z[i + 1], 0.f)));
                        pooled[i] = s.x;
                        pooled[i + 1] = s.y;
                    }
                }
            } else if (kind == kL3) {
                float z[CS];
                tmem_ld_cs<CS>(tmem_base + lane_off + acc * C + col0, z);
                const int slot = (pd >> 2) & 7, t = (pd >> 5) & 1, rows = ((pd >> 12) & 127) + 1;
                if (erow < rows) {
                    const uint32_t pos = lds_u16(sPos + (uint32_t)(slot * S + t * kTile + erow) * 2u);
                    const uint32_t base = sStrip + slot * strip_bytes + (uint32_t)(col0 * S) * 4u + pos * 4u;
#pragma unroll
                    for (int i = 0; i < CS; ++i) {
                        const uint32_t addr = base + (uint32_t)(i * S) * 4u;
                        sts_f32(addr, lds_f32(addr) + z[i]);
                    }
                }
            }
            tc::fence_before_sync();
            if (kind == kL3 && (pd & (1u << 11))) tc::fence_proxy_async();   // the strip buffer goes back through the bulk-copy engine
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_a(sBar + (kBarAccFree + acc) * 8);
                if (kind == kL3 && (pd & (1u << 11))) mbar_arrive_a(sBar + (kBarStore + ((pd >> 2) & 7)) * 8);
            }
        };

        // one position: deferred epilogue of n - D, build of n, prefetch of n + 2 into tv
        auto step = [&](uint32_t n, float (&tv)[kItems][8], bool &tv_valid) -> bool {
            wait_published(n);
            const uint32_t desc = desc_at(n);
            const uint32_t kind = desc & 3u;
            if (kind == kEnd) return false;
            if (n >= (uint32_t)D) epilogue(n - D);
            const uint32_t a_addr = sA + a * L::kASlot;
            if (kind == kSlot) {
                const int slot = (desc >> 2) & 7, t = (desc >> 5) & 1, k = (desc >> 6) & 15, rows = ((desc >> 12) & 127) + 1;
                if (!tv_valid) gather(desc, tv);
                if (desc & (1u << 10)) {   // first slot of a tile: this row's cell centre
                    if (brow < rows) {
                        const uint32_t pos = lds_u16(sPos + (uint32_t)(slot * S + t * kTile + brow) * 2u);
                        int32_t cell0;
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(cell0) : "r"(sInfo + slot * 16 + 4));
                        const int32_t cell = cell0 + (int32_t)pos;
                        const int32_t i = (int32_t)((uint32_t)cell / (uint32_t)p.W), j = cell - i * p.W;
                        cx = __fadd_rn(p.x0, __fmul_rn((float)i, p.dx));
                        cy = __fadd_rn(p.y0, __fmul_rn((float)j, p.dy));
                    }
                }
                if (brow < rows) {
                    const float2 cxx = make_float2(cx, cx), cyy = make_float2(cy, cy);
#pragma unroll
                    for (int it = 0; it < kItems; ++it) {
                        const int ku = it * 4 + u4;
                        const uint32_t wa = sW1 + (uint32_t)(ku * 64);
                        const float4 x0 = tc::lds_f32x4(wa), x1 = tc::lds_f32x4(wa + 16), y0 = tc::lds_f32x4(wa + 32), y1 = tc::lds_f32x4(wa + 48);
                        float2 v[4];
                        v[0] = tc::ffma2(make_float2(x0.x, x0.y), cxx, tc::ffma2(make_float2(y0.x, y0.y), cyy, make_float2(tv[it][0], tv[it][1])));
                        v[1] = tc::ffma2(make_float2(x0.z, x0.w), cxx, tc::ffma2(make_float2(y0.z, y0.w), cyy, make_float2(tv[it][2], tv[it][3])));
                        v[2] = tc::ffma2(make_float2(x1.x, x1.y), cxx, tc::ffma2(make_float2(y1.x, y1.y), cyy, make_float2(tv[it][4], tv[it][5])));
                        v[3] = tc::ffma2(make_float2(x1.z, x1.w), cxx, tc::ffma2(make_float2(y1.z, y1.w), cyy, make_float2(tv[it][6], tv[it][7])));
                        uint4 hi, lo;
                        tc::relu_split_bf16x8(v, hi, lo, NS == 2);
                        const uint32_t dst = a_addr + a_row_off + (uint32_t)(ku * 128);
                        tc::sts_u32x4(dst, hi);
                        if (NS == 2) tc::sts_u32x4(dst + kTile * C * 2, lo);
                    }
                }
                if (cs == 0) {   // bias flag of the row: bf16 (1, 1) if it has a k-th neighbour
                    const bool on = erow < rows && k < (int)lds_u8(sNv + (uint32_t)(slot * S + t * kTile + erow));
                    tc::sts_u32(a_addr + L::kAData + tc::unit_offset(erow, 0, 2), on ? 0x3F803F80u : 0u);
                }
            } else if (kind == kL3) {
                // pooled -> A operand (pooled >= 0: the fused ReLU is the identity); row = TMEM lane, CS channels per worker
                const int slot = (desc >> 2) & 7, t = (desc >> 5) & 1, rows = ((desc >> 12) & 127) + 1;
#pragma unroll
                for (int u = 0; u < CS / 8; ++u) {
                    float2 v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[i] = make_float2(pooled[u * 8 + 2 * i], pooled[u * 8 + 2 * i + 1]);
                    uint4 hi, lo;
                    tc::relu_split_bf16x8(v, hi, lo, NS == 2);
                    const uint32_t dst = a_addr + tc::unit_offset(erow, col0 / 8 + u, kc_units);
                    tc::sts_u32x4(dst, hi);
                    if (NS == 2) tc::sts_u32x4(dst + kTile * C * 2, lo);
                }
                if (cs == 0) {   // bias column = the row's neighbour count (small integers are exact in bf16)
                    const uint32_t nv = erow < rows ? lds_u8(sNv + (uint32_t)(slot * S + t * kTile + erow)) : 0u;
                    const uint32_t nv16 = __float_as_uint((float)nv) >> 16;
                    tc::sts_u32(a_addr + L::kAData + tc::unit_offset(erow, 0, 2), nv16 | (nv16 << 16));
                }
            }
            tc::fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(sBar + (kBarFull + a) * 8);
            a = a + 1 == (uint32_t)D ? 0 : a + 1;
            // prefetch: the neighbour rows of position n + 2 (same register buffer) if it is already published
            tv_valid = false;
            if (seen <= n + 2) seen = ld_acquire(sPub);
            if (seen > n + 2) {
                const uint32_t d2 = desc_at(n + 2);
                if ((d2 & 3u) == kSlot) {
                    gather(d2, tv);
                    tv_valid = true;
                }
            }
            return true;
        };

        float tvA[kItems][8], tvB[kItems][8];
        bool validA = false, validB = false;
        // warm-up of the register FIFO: positions 0 and 1
        wait_published(0);
        {
            const uint32_t d0 = desc_at(0);
            if ((d0 & 3u) == kSlot) { gather(d0, tvA); validA = true; }
            if ((d0 & 3u) != kEnd) {
                wait_published(1);
                const uint32_t d1 = desc_at(1);
                if ((d1 & 3u) == kSlot) { gather(d1, tvB); validB = true; }
            }
        }
        for (uint32_t n = 0;; n += 2) {
            if (!step(n, tvA, validA)) break;
            if (!step(n + 1, tvB, validB)) break;
        }
    }

    tc::fence_before_sync();
    __syncthreads();
    if (warp == kIssuerWarp) tc::tmem_free(tmem_base, kNacc * C);
}

template <int C, int NS, int S, int D>
int launch_strip(StripParams &p, cudaStream_t st)
{
    using L = StripLayout<C, NS, S, D>;
    const int sb = L::strip_bytes(p.K);
    int nb = kMaxNB;
    while (nb > 2 && L::smem_bytes(p.K, nb) > 227 * 1024) --nb;
    const int smem = L::smem_bytes(p.K, nb);
    if (smem > 227 * 1024) return CF_ERR_UNSUPPORTED;
    static int attr_bytes = 0;
    if (smem > attr_bytes) {
        CF_TRY(cuda_status(cudaFuncSetAttribute(k_fusion_strip<C, NS, S, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                           "k_fusion_strip smem attribute"));
        attr_bytes = smem;
    }
    p.nb = nb;
    p.strip_bytes = sb;
    p.strips_per_frame = (p.cells + S - 1) / S;
    p.strips_total = p.strips_per_frame * p.B;
    const int grid = std::min<int64_t>(p.strips_total, sm_count());
    k_fusion_strip<C, NS, S, D><<<grid, kThreads, smem, st>>>(p);
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
    const int NS = mode == CF_MODE_FP32 ? 2 : 1;
    int rc;
    if (C == 32) rc = NS == 2 ? launch_strip<32, 2, 256, 3>(p, st) : launch_strip<32, 1, 256, 3>(p, st);
    else rc = NS == 2 ? launch_strip<64, 2, 128, 2>(p, st) : launch_strip<64, 1, 128, 2>(p, st);
    if (rc != CF_OK) return rc;
    count_launches(1);
    return launch_status("cf_fusion_fwd (strip pipeline)");
}

}  // namespace cf
