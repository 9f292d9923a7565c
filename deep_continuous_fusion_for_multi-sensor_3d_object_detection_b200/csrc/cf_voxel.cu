// cf_voxel.cu -- dataset-side voxelisation + projection on the device: the work of
// CarlaDataset.Voxelization_Projection / .Projection (data_import_carla.py:196-267) for a batch of raw sweeps:
//   range filter -> trilinear splat into the (Z, X, Y) voxel grid -> camera projection + image filter -> zero-padded
//   pointcloud_raw / projected_loc_uv / num_points_raw, i.e. every tensor the LiDAR backbone and the fusion layer consume.
//
// Semantics follow the reference statement by statement (the CPU restatement in the test infrastructure is pinned against
// the reference's own tensors):
//   * survivors keep their input order (ordered block-scan compaction, one CTA per frame);
//   * `voxel[idx] += w` with advanced indexing is an index_put_ WITHOUT accumulation: within each of the 8 splat statements,
//     of the points that fall into the same voxel only the LAST one counts.  Reproduced deterministically with an owner
//     pass (atomicMax of a tag that grows with statement and point index) followed by an add pass in which only the owner
//     writes -- no floating-point atomics, bit-identical to the sequential reference;
//   * the nonzero()/3 and nonzero()/2 bookkeeping of the reference (:228-229, :206-207) is kept: the number of kept points
//     is floor(count_nonzero / 3) resp. floor(count_nonzero / 2) of the surviving coordinates.
// All arithmetic is fp32 with separately rounded multiplies and adds (as torch evaluates the expressions).
#include "cf_common.cuh"

namespace cf {

namespace {

struct VoxParams {
    const float *raw;          // (B, Nraw, 3)
    const int64_t *num_raw;    // (B)
    int32_t B, Nraw;
    float x_lo, x_hi, y_lo, y_hi, z_lo, z_hi;   // keep lo < v < hi
    float xs, ys, zs, xo, yo, zo;               // voxel index = coord * scale + offset
    int32_t Z, X, Y;
    float crt[12];                              // CRT_tensor (4,3) row-major: [x y z 1] @ CRT = (u w, v w, w)
    float u_hi, v_hi;                           // 0 < u < u_hi (image_height), 0 < v < v_hi (image_width)  [sic, :202-205]
    int32_t max_num_pc;
    float *voxel;              // (B, Z, X, Y)
    float *points;             // (B, max_num_pc, 3)
    float *uv;                 // (B, max_num_pc, 2)
    int64_t *num_points;       // (B)
    float *kept;               // workspace (B, Nraw, 3): range-filtered points in order
    int32_t *n_kept;           // workspace (B)
    uint32_t *owner;           // workspace (B, Z*X*Y)
};

constexpr int kScanThreads = 1024;

// ordered compaction step of one chunk: returns the output slot of this thread's element (if flag) given the running base;
// updates base for the next chunk.  All kScanThreads threads must call it.
__device__ __forceinline__ int32_t block_ordered_slot(bool flag, int32_t &base, int32_t *warp_tot)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0) {
        int32_t v = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        warp_tot[32 + lane] = v;   // inclusive prefix of the warp totals
    }
    __syncthreads();
    const int32_t before = warp ? warp_tot[32 + warp - 1] : 0;
    const int32_t slot = base + before + __popc(bal & ((1u << lane) - 1u));
    base += warp_tot[32 + 31];
    __syncthreads();
    return slot;
}

// range filter + ordered compaction (one CTA per frame)
__global__ void __launch_bounds__(kScanThreads) k_vox_filter(const VoxParams p)
{
    __shared__ int32_t warp_tot[64];
    __shared__ unsigned long long nz_total;
    const int b = blockIdx.x;
    int64_t n = p.num_raw[b];
    n = n < 0 ? 0 : (n > p.Nraw ? p.Nraw : n);
    if (threadIdx.x == 0) nz_total = 0ull;
    __syncthreads();
    const float *src = p.raw + (size_t)b * p.Nraw * 3;
    float *dst = p.kept + (size_t)b * p.Nraw * 3;
    int32_t base = 0;
    unsigned nz = 0;
    for (int64_t i0 = 0; i0 < n; i0 += kScanThreads) {
        const int64_t i = i0 + threadIdx.x;
        float x = 0.f, y = 0.f, z = 0.f;
        bool keep = false;
        if (i < n) {
            x = src[i * 3];
            y = src[i * 3 + 1];
            z = src[i * 3 + 2];
            keep = x > p.x_lo && x < p.x_hi && y > p.y_lo && y < p.y_hi && z > p.z_lo && z < p.z_hi;
        }
        const int32_t slot = block_ordered_slot(keep, base, warp_tot);
        if (keep) {
            dst[(size_t)slot * 3] = x;
            dst[(size_t)slot * 3 + 1] = y;
            dst[(size_t)slot * 3 + 2] = z;
            nz += (x != 0.f) + (y != 0.f) + (z != 0.f);
        }
    }
    atomicAdd(&nz_total, (unsigned long long)nz);
    __syncthreads();
    if (threadIdx.x == 0) {
        // nonzero()[: count / 3]: the first third of the entries are column indices of the first coordinate row
        const int64_t third = (int64_t)(nz_total / 3ull);
        p.n_kept[b] = (int32_t)(third < base ? third : base);
    }
}

__device__ __forceinline__ bool splat_target(const VoxParams &p, int b, int i, int s, size_t &lin, float &w)
{
    const float *q = p.kept + ((size_t)b * p.Nraw + i) * 3;
    const float fx = __fadd_rn(__fmul_rn(q[0], p.xs), p.xo), fy = __fadd_rn(__fmul_rn(q[1], p.ys), p.yo);
    const float fz = __fadd_rn(__fmul_rn(q[2], p.zs), p.zo);
    const int32_t xl = (int32_t)fx, yl = (int32_t)fy, zl = (int32_t)fz;   // .type(torch.long): truncation
    const float dx = __fsub_rn(fx, (float)xl), dy = __fsub_rn(fy, (float)yl), dz = __fsub_rn(fz, (float)zl);
    // statement order of data_import_carla.py:249-256: bit 0 = z upper, bit 1 = x upper, bit 2 = y upper
    const int zu = s & 1, xu = (s >> 1) & 1, yu = (s >> 2) & 1;
    const float wx = xu ? dx : __fsub_rn(1.0f, dx), wy = yu ? dy : __fsub_rn(1.0f, dy), wz = zu ? dz : __fsub_rn(1.0f, dz);
    w = __fmul_rn(__fmul_rn(wx, wy), wz);
    const int32_t zi = zl + zu, xi = xl + xu, yi = yl + yu;
    if (zi < 0 || zi >= p.Z || xi < 0 || xi >= p.X || yi < 0 || yi >= p.Y) return false;   // (the reference would raise)
    lin = ((size_t)zi * p.X + xi) * p.Y + yi;
    return true;
}

// owner pass of statement s: the last point (largest index) that targets a voxel wins the statement
__global__ void __launch_bounds__(256) k_vox_owner(const VoxParams p, int s)
{
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n_kept[b]) return;
    size_t lin;
    float w;
    if (!splat_target(p, b, i, s, lin, w)) return;
    atomicMax(p.owner + (size_t)b * p.Z * p.X * p.Y + lin, (uint32_t)s * (uint32_t)p.Nraw + (uint32_t)i + 1u);
}

__global__ void __launch_bounds__(256) k_vox_add(const VoxParams p, int s)
{
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n_kept[b]) return;
    size_t lin;
    float w;
    if (!splat_target(p, b, i, s, lin, w)) return;
    const size_t o = (size_t)b * p.Z * p.X * p.Y + lin;
    if (p.owner[o] == (uint32_t)s * (uint32_t)p.Nraw + (uint32_t)i + 1u) p.voxel[o] = __fadd_rn(p.voxel[o], w);
}

// projection + image filter + ordered compaction into the padded outputs (one CTA per frame)
__global__ void __launch_bounds__(kScanThreads) k_vox_project(const VoxParams p)
{
    __shared__ int32_t warp_tot[64];
    __shared__ unsigned long long nz_total;
    const int b = blockIdx.x;
    const int32_t n = p.n_kept[b];
    if (threadIdx.x == 0) nz_total = 0ull;
    __syncthreads();
    const float *src = p.kept + (size_t)b * p.Nraw * 3;
    int32_t base = 0;
    unsigned nz = 0;
    for (int32_t i0 = 0; i0 < n; i0 += kScanThreads) {
        const int32_t i = i0 + threadIdx.x;
        float x = 0.f, y = 0.f, z = 0.f, u = 0.f, v = 0.f;
        bool keep = false;
        if (i < n) {
            x = src[(size_t)i * 3];
            y = src[(size_t)i * 3 + 1];
            z = src[(size_t)i * 3 + 2];
            float q[3];
#pragma unroll
            for (int c = 0; c < 3; ++c)
                q[c] = __fadd_rn(__fadd_rn(__fmul_rn(x, p.crt[c]), __fmul_rn(y, p.crt[3 + c])),
                                 __fadd_rn(__fmul_rn(z, p.crt[6 + c]), p.crt[9 + c]));
            u = __fdiv_rn(q[0], q[2]);
            v = __fdiv_rn(q[1], q[2]);
            keep = u > 0.f && u < p.u_hi && v > 0.f && v < p.v_hi;
        }
        const int32_t slot = block_ordered_slot(keep, base, warp_tot);
        if (keep) {
            nz += (u != 0.f) + (v != 0.f);
            if (slot < p.max_num_pc) {
                float *dp = p.points + ((size_t)b * p.max_num_pc + slot) * 3;
                dp[0] = x; dp[1] = y; dp[2] = z;
                float *du = p.uv + ((size_t)b * p.max_num_pc + slot) * 2;
                du[0] = u; du[1] = v;
            }
        }
    }
    atomicAdd(&nz_total, (unsigned long long)nz);
    __syncthreads();
    // nonzero()[: count / 2]; rows beyond it stay zero padding (they are cleared below if they were written)
    const int64_t half = (int64_t)(nz_total / 2ull);
    int32_t num = (int32_t)(half < base ? half : base);
    num = num < p.max_num_pc ? num : p.max_num_pc;
    for (int32_t r = num + threadIdx.x; r < base && r < p.max_num_pc; r += kScanThreads) {
        float *dp = p.points + ((size_t)b * p.max_num_pc + r) * 3;
        dp[0] = dp[1] = dp[2] = 0.f;
        float *du = p.uv + ((size_t)b * p.max_num_pc + r) * 2;
        du[0] = du[1] = 0.f;
    }
    if (threadIdx.x == 0) p.num_points[b] = num;
}

}  // namespace

}  // namespace cf

extern "C" size_t cf_voxelize_workspace_bytes(int32_t B, int32_t Nraw, int32_t Z, int32_t X, int32_t Y)
{
    if (B <= 0 || Nraw <= 0 || Z <= 0 || X <= 0 || Y <= 0) return 0;
    return (size_t)B * Nraw * 3 * sizeof(float) + 256 + (size_t)B * Z * X * Y * sizeof(uint32_t);
}

extern "C" int cf_voxelize_project(const float *d_raw, const int64_t *d_num_raw, int32_t B, int32_t Nraw,
                                   const float *h_range, const float *h_vox, int32_t Z, int32_t X, int32_t Y,
                                   const float *h_calib, float u_hi, float v_hi, int32_t max_num_pc, float *d_voxel,
                                   float *d_points, float *d_uv, int64_t *d_num_points, void *d_workspace, void *stream)
{
    using namespace cf;
    CF_TRY(require_sm100());
    CF_REQUIRE(d_raw && d_num_raw && h_range && h_vox && h_calib && d_voxel && d_points && d_uv && d_num_points && d_workspace,
               CF_ERR_ARG, "cf_voxelize_project: null pointer");
    CF_REQUIRE(B > 0 && B <= 65535 && Nraw > 0 && Z > 0 && X > 0 && Y > 0 && max_num_pc > 0, CF_ERR_ARG,
               "cf_voxelize_project: bad extents");
    CF_REQUIRE((uint64_t)8 * (uint64_t)Nraw + 1 < (1ull << 32), CF_ERR_ARG, "cf_voxelize_project: Nraw=%d too large", Nraw);
    CF_REQUIRE(h_range[0] >= 0.0f, CF_ERR_UNSUPPORTED,
               "cf_voxelize_project: lidar_x_min < 0 (the reference's nonzero()/3 bookkeeping is only reproduced for x > 0)");
    CF_REQUIRE(aligned16(d_workspace), CF_ERR_ALIGN, "cf_voxelize_project: workspace must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    VoxParams p;
    p.raw = d_raw; p.num_raw = d_num_raw; p.B = B; p.Nraw = Nraw;
    p.x_lo = h_range[0]; p.x_hi = h_range[1]; p.y_lo = h_range[2]; p.y_hi = h_range[3]; p.z_lo = h_range[4]; p.z_hi = h_range[5];
    p.xs = h_vox[0]; p.ys = h_vox[1]; p.zs = h_vox[2]; p.xo = h_vox[3]; p.yo = h_vox[4]; p.zo = h_vox[5];
    p.Z = Z; p.X = X; p.Y = Y;
    for (int i = 0; i < 12; ++i) p.crt[i] = h_calib[i];
    p.u_hi = u_hi; p.v_hi = v_hi; p.max_num_pc = max_num_pc;
    p.voxel = d_voxel; p.points = d_points; p.uv = d_uv; p.num_points = d_num_points;
    uint8_t *ws = (uint8_t *)d_workspace;
    p.kept = (float *)ws;
    const size_t kept_bytes = ((size_t)B * Nraw * 3 * sizeof(float) + 15) / 16 * 16;
    p.n_kept = (int32_t *)(ws + kept_bytes);
    p.owner = (uint32_t *)(ws + kept_bytes + 256);
    CF_REQUIRE(B <= 64, CF_ERR_ARG, "cf_voxelize_project: batch %d > 64 frames per call", B);
    const size_t vox_elems = (size_t)B * Z * X * Y;
    CF_TRY(cuda_status(cudaMemsetAsync(d_voxel, 0, vox_elems * sizeof(float), st), "cf_voxelize_project memset"));
    CF_TRY(cuda_status(cudaMemsetAsync(p.owner, 0, vox_elems * sizeof(uint32_t), st), "cf_voxelize_project memset"));
    CF_TRY(cuda_status(cudaMemsetAsync(d_points, 0, (size_t)B * max_num_pc * 3 * sizeof(float), st), "cf_voxelize_project memset"));
    CF_TRY(cuda_status(cudaMemsetAsync(d_uv, 0, (size_t)B * max_num_pc * 2 * sizeof(float), st), "cf_voxelize_project memset"));
    k_vox_filter<<<B, kScanThreads, 0, st>>>(p);
    const dim3 grid((unsigned)((Nraw + 255) / 256), (unsigned)B);
    for (int s = 0; s < 8; ++s) {
        k_vox_owner<<<grid, 256, 0, st>>>(p, s);
        k_vox_add<<<grid, 256, 0, st>>>(p, s);
    }
    k_vox_project<<<B, kScanThreads, 0, st>>>(p);
    count_launches(18);
    return launch_status("cf_voxelize_project");
}
