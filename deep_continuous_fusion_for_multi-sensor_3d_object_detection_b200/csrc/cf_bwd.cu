// cf_bwd.cu -- K-4b: backward of the continuous-fusion layer (needed by the training configuration only).
//
// Forward (per scale, see cf_mlp_simt.cu / cf_mlp_tc.cu):
//   T_p   = W1[:, :Ci] f_p + W1[:, Ci:] p + b1                      per point
//   h1_r  = relu(T[j_r] - w1x*cx - w1y*cy)                          per row r = (cell, k) with a neighbour j_r
//   h2_r  = relu(W2 h1_r + b2);  pooled_c = sum_k h2_(c,k);  out_c = bev_c + W3 pooled_c + n_valid(c) b3
// Backward, given g = dL/dout (NCHW):   d bev = g   (returned by the host layer as the same tensor), and
//   G (cells x C)   = g transposed to cell-major
//   dW3 += G^T pooled        db3 += sum_c n_valid(c) G_c        dPooled = G W3
//   dZ2_r = dPooled_cell(r) * [h2_r > 0]      dW2 += dZ2^T H1      db2 += sum_r dZ2_r
//   dA_r  = (dZ2_r W2) * [h1_r > 0]           dT[j_r] += dA_r  (scatter-add, atomics)
//   dW1[:, Ci]   -= sum_r dA_r cx_r           dW1[:, Ci+1] -= sum_r dA_r cy_r
//   dW1[:, :Ci] += dT^T F     dW1[:, Ci:] += dT^T P     db1 += sum_p dT_p     dF += dT W1[:, :Ci]
// Nothing is saved by the forward except indices, inputs and weights: H1, H2 and pooled are recomputed here and
// materialised in the caller-provided workspace (dense (cell,k) rows, zero for empty slots).  All arithmetic is
// fp32 on CUDA cores (this is the round-1 "correct first" implementation; reductions over rows use split
// accumulation + atomics, so gradients are reproducible to rounding, not bit for bit).
#include "cf_common.cuh"

namespace cf {

namespace {

// ---------------------------------------------------------------------------------------------------------------
// C[R x N] (beta ? += : =) A[R x Kd] * B[Kd x N]; optional mask: C := 0 where mask[r,n] <= 0.   64x64x16 tiles.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sgemm_nn(const float *__restrict__ A, int64_t lda, const float *__restrict__ Bm,
                                                  int64_t ldb, float *__restrict__ Cm, int64_t ldc, int64_t R, int32_t N,
                                                  int32_t Kd, int beta, const float *__restrict__ mask, int64_t ldm)
{
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t r0 = (int64_t)blockIdx.x * 64;
    const int32_t n0 = blockIdx.y * 64;
    float acc[4][4] = {};
    for (int32_t k0 = 0; k0 < Kd; k0 += 16) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int idx = tid + t * 256;
            {   // A tile: 64 rows x 16 k
                const int rr = idx >> 4, kk = idx & 15;
                const int64_t r = r0 + rr;
                As[kk][rr] = (r < R && k0 + kk < Kd) ? __ldg(A + r * lda + k0 + kk) : 0.0f;
            }
            {   // B tile: 16 k x 64 n
                const int kk = idx >> 6, nn = idx & 63;
                Bs[kk][nn] = (k0 + kk < Kd && n0 + nn < N) ? __ldg(Bm + (int64_t)(k0 + kk) * ldb + n0 + nn) : 0.0f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const float4 av = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            const float a[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = r0 + ty * 4 + i;
        if (r >= R) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int32_t n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (mask && !(__ldg(mask + r * ldm + n) > 0.0f)) v = 0.0f;
            float *dst = Cm + r * ldc + n;
            *dst = beta ? *dst + v : v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// C[M x N] += scale * A[R x M]^T * B[R x N]   (reduction over the long dimension R, split across blockIdx.z, atomics)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sgemm_tn(const float *__restrict__ A, int64_t lda, const float *__restrict__ Bm,
                                                  int64_t ldb, float *__restrict__ Cm, int64_t ldc, int64_t R, int32_t M,
                                                  int32_t N, int64_t rows_per_split)
{
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int32_t m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    const int64_t r_begin = (int64_t)blockIdx.z * rows_per_split, r_end = min(R, r_begin + rows_per_split);
    float acc[4][4] = {};
    for (int64_t r0 = r_begin; r0 < r_end; r0 += 16) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int idx = tid + t * 256;
            const int rr = idx >> 6, cc = idx & 63;
            const int64_t r = r0 + rr;
            As[rr][cc] = (r < r_end && m0 + cc < M) ? __ldg(A + r * lda + m0 + cc) : 0.0f;
            Bs[rr][cc] = (r < r_end && n0 + cc < N) ? __ldg(Bm + r * ldb + n0 + cc) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const float4 av = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            const float a[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int32_t m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < M && n < N && acc[i][j] != 0.0f) atomicAdd(Cm + (int64_t)m * ldc + n, acc[i][j]);
        }
}

// out[n] += scale * sum_r w[r] * X[r, n]   (w == nullptr: plain column sum).
// grid (row slabs of 1024, column chunks of 32); a warp owns every 8th row of the slab, lane = column: 128-byte row
// segments, 4 independent partial sums per lane, then an 8-warp reduction in shared memory and one atomicAdd per column.
__global__ void __launch_bounds__(256) k_colsum(const float *__restrict__ X, int64_t ldx, int64_t R, int32_t N,
                                                const float *__restrict__ w, float scale, float *__restrict__ out,
                                                int64_t out_stride)
{
    __shared__ float part[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t n = blockIdx.y * 32 + lane;
    const int64_t r_begin = (int64_t)blockIdx.x * 1024, r_end = min(R, r_begin + 1024);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (n < N) {
        int64_t r = r_begin + warp;
        for (; r + 24 < r_end; r += 32) {
            const float x0 = __ldg(X + r * ldx + n), x1 = __ldg(X + (r + 8) * ldx + n);
            const float x2 = __ldg(X + (r + 16) * ldx + n), x3 = __ldg(X + (r + 24) * ldx + n);
            s0 = fmaf(w ? __ldg(w + r) : 1.0f, x0, s0);
            s1 = fmaf(w ? __ldg(w + r + 8) : 1.0f, x1, s1);
            s2 = fmaf(w ? __ldg(w + r + 16) : 1.0f, x2, s2);
            s3 = fmaf(w ? __ldg(w + r + 24) : 1.0f, x3, s3);
        }
        for (; r < r_end; r += 8) s0 = fmaf(w ? __ldg(w + r) : 1.0f, __ldg(X + r * ldx + n), s0);
    }
    part[warp][lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (warp == 0 && n < N) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += part[i][lane];
        if (s != 0.0f) atomicAdd(out + (int64_t)n * out_stride, s * scale);
    }
}

// (B, C, cells) -> (B, cells, C)
__global__ void __launch_bounds__(256) k_nchw_to_cell_major(const float *__restrict__ src, int32_t C, int64_t cells,
                                                            float *__restrict__ dst)
{
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int64_t p0 = (int64_t)blockIdx.x * 32;
    const int32_t c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int32_t c = c0 + r;
        const int64_t p = p0 + tx;
        tile[r][tx] = (c < C && p < cells) ? __ldg(src + ((size_t)b * C + c) * cells + p) : 0.0f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int64_t p = p0 + r;
        const int32_t c = c0 + tx;
        if (p < cells && c < C) dst[((size_t)b * cells + p) * C + c] = tile[tx][r];
    }
}

// Recompute H1 rows: H1[(b,cell,k), c] = relu(T[b, j, c] - w1x[c] cx - w1y[c] cy) (0 for an empty slot); also the
// per-row cell centre (for the offset-column gradients) and per-cell n_valid.
__global__ void __launch_bounds__(256) k_bwd_h1(const float *__restrict__ T, const int32_t *__restrict__ knn, int32_t N,
                                                int32_t C, int32_t H, int32_t W, int32_t K, float x0, float y0, float dx,
                                                float dy, const float *__restrict__ W1, int32_t Ci,
                                                float *__restrict__ H1, float *__restrict__ row_cx,
                                                float *__restrict__ row_cy, float *__restrict__ n_valid)
{
    const int b = blockIdx.y;
    const int64_t cells = (int64_t)H * W, rows = cells * K;
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
        const int64_t cell = r / K;
        const int32_t k = (int32_t)(r - cell * K);
        const int32_t i = (int32_t)(cell / W), j = (int32_t)(cell - (int64_t)i * W);
        const float cx = __fadd_rn(x0, __fmul_rn((float)i, dx)), cy = __fadd_rn(y0, __fmul_rn((float)j, dy));
        const int32_t p = __ldg(knn + ((size_t)b * cells + cell) * K + k);
        float *h = H1 + ((size_t)b * rows + r) * C;
        for (int32_t c = lane; c < C; c += 32) {
            float v = 0.0f;
            if (p >= 0) {
                const float *w = W1 + (size_t)c * (Ci + 3) + Ci;
                v = fmaxf(__ldg(T + ((size_t)b * N + p) * C + c) - (__ldg(w) * cx + __ldg(w + 1) * cy), 0.0f);
            }
            h[c] = v;
        }
        if (lane == 0) {
            row_cx[(size_t)b * rows + r] = p >= 0 ? cx : 0.0f;
            row_cy[(size_t)b * rows + r] = p >= 0 ? cy : 0.0f;
            if (k == 0) {
                int nv = 0;
                for (int kk = 0; kk < K; ++kk) nv += __ldg(knn + ((size_t)b * cells + cell) * K + kk) >= 0;
                n_valid[(size_t)b * cells + cell] = (float)nv;
            }
        }
    }
}

// H2[r, c] = valid(r) ? relu(Z[r, c] + b2[c]) : 0 (in place on Z);  pooled[cell, c] = sum_k H2
__global__ void __launch_bounds__(256) k_bwd_h2_pool(float *__restrict__ Z, const int32_t *__restrict__ knn_flat,
                                                     const float *__restrict__ b2, int64_t n_cells_total, int32_t K,
                                                     int32_t C, float *__restrict__ pooled)
{
    const int64_t total = n_cells_total * C;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t cell = t / C;
        const int32_t c = (int32_t)(t - cell * C);
        float s = 0.0f;
        for (int k = 0; k < K; ++k) {
            const int64_t r = cell * K + k;
            const bool valid = __ldg(knn_flat + r) >= 0;
            const float h = valid ? fmaxf(Z[r * C + c] + __ldg(b2 + c), 0.0f) : 0.0f;
            Z[r * C + c] = h;
            s += h;
        }
        pooled[t] = s;
    }
}

// dZ2[r, c] = dPooled[cell(r), c] * [H2[r, c] > 0]   (in place on H2)
__global__ void __launch_bounds__(256) k_bwd_dz2(float *__restrict__ H2, const float *__restrict__ dPooled, int64_t rows,
                                                 int32_t K, int32_t C)
{
    const int64_t total = rows * C;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / C;
        const int32_t c = (int32_t)(t - r * C);
        H2[t] = H2[t] > 0.0f ? __ldg(dPooled + (r / K) * C + c) : 0.0f;
    }
}

// dT[b, j_r, :] += dA[r, :]   (one warp per row, float atomics)
__global__ void __launch_bounds__(256) k_bwd_scatter(const float *__restrict__ dA, const int32_t *__restrict__ knn, int32_t N,
                                                     int32_t C, int64_t rows_per_frame, float *__restrict__ dT)
{
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows_per_frame; r += warps) {
        const int32_t p = __ldg(knn + (size_t)b * rows_per_frame + r);
        if (p < 0) continue;
        const float *src = dA + ((size_t)b * rows_per_frame + r) * C;
        float *dst = dT + ((size_t)b * N + p) * C;
        for (int32_t c = lane; c < C; c += 32) {
            const float v = __ldg(src + c);
            if (v != 0.0f) atomicAdd(dst + c, v);
        }
    }
}

// d img[b, c, y, x] += w_tap * dF[b, p, c] for the 4 bilinear taps of point p (adjoint of k_point_gather)
struct CalibB {
    float m[12];
};
__global__ void __launch_bounds__(256) k_point_gather_bwd(const float *__restrict__ dF, int64_t sb, int64_t sc, int64_t sh,
                                                          int64_t sw, int32_t Ci, int32_t Hf, int32_t Wf,
                                                          const float *__restrict__ points, const float *__restrict__ uv,
                                                          CalibB cal, int use_calib, const int64_t *__restrict__ num_points,
                                                          int32_t N, float sx, float sy, float *__restrict__ dimg)
{
    const int b = blockIdx.y;
    const int32_t n = valid_points(num_points, b, N);
    const int lane = threadIdx.x & 31;
    const int32_t warps = gridDim.x * (blockDim.x >> 5);
    for (int32_t p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); p < n; p += warps) {
        float u, v;
        if (use_calib) {
            const float *q = points + ((size_t)b * N + p) * 3;
            const float x = __ldg(q), y = __ldg(q + 1), z = __ldg(q + 2);
            float r[3];
#pragma unroll
            for (int c = 0; c < 3; ++c)
                r[c] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, cal.m[c]), __fmul_rn(y, cal.m[3 + c])),
                                           __fmul_rn(z, cal.m[6 + c])), cal.m[9 + c]);
            u = __fdiv_rn(r[0], r[2]);
            v = __fdiv_rn(r[1], r[2]);
        } else {
            u = __ldg(uv + ((size_t)b * N + p) * 2);
            v = __ldg(uv + ((size_t)b * N + p) * 2 + 1);
        }
        const float uf = __fsub_rn(__fmul_rn(__fadd_rn(u, 0.5f), sx), 0.5f);
        const float vf = __fsub_rn(__fmul_rn(__fadd_rn(v, 0.5f), sy), 0.5f);
        if (!(uf > -1.0f && uf < (float)Wf && vf > -1.0f && vf < (float)Hf)) continue;
        const float fx = floorf(uf), fy = floorf(vf);
        const int32_t ix = (int32_t)fx, iy = (int32_t)fy;
        const float wx1 = uf - fx, wy1 = vf - fy, wx0 = 1.0f - wx1, wy0 = 1.0f - wy1;
        const float wt[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};
        const int32_t tx[4] = {ix, ix + 1, ix, ix + 1}, ty[4] = {iy, iy, iy + 1, iy + 1};
        const float *src = dF + ((size_t)b * N + p) * Ci;
        for (int32_t c = lane; c < Ci; c += 32) {
            const float gval = __ldg(src + c);
            if (gval == 0.0f) continue;
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (tx[t] >= 0 && tx[t] < Wf && ty[t] >= 0 && ty[t] < Hf)
                    atomicAdd(dimg + b * sb + c * sc + ty[t] * sh + tx[t] * sw, wt[t] * gval);
        }
    }
}

static size_t up256(size_t v) { return (v + 255) / 256 * 256; }

struct BwdWs {
    float *H1, *H2, *pooled, *G, *dPooled, *row_cx, *row_cy, *n_valid, *T, *dT, *W2t;
    size_t bytes;
};

static BwdWs carve(void *base, int32_t B, int32_t N, int32_t C, int64_t cells, int32_t K)
{
    BwdWs w;
    size_t off = 0;
    auto take = [&](size_t n_floats) {
        float *p = base ? (float *)((char *)base + off) : nullptr;
        off += up256(n_floats * sizeof(float));
        return p;
    };
    const size_t rows = (size_t)B * cells * K, ncell = (size_t)B * cells;
    w.H1 = take(rows * C);
    w.H2 = take(rows * C);
    w.pooled = take(ncell * C);
    w.G = take(ncell * C);
    w.dPooled = take(ncell * C);
    w.row_cx = take(rows);
    w.row_cy = take(rows);
    w.n_valid = take(ncell);
    w.T = take((size_t)B * N * C);
    w.dT = take((size_t)B * N * C);
    w.W2t = take((size_t)C * C);
    w.bytes = off;
    return w;
}

__global__ void __launch_bounds__(256) k_transpose_sq_b(const float *__restrict__ W, int32_t C, float *__restrict__ Wt)
{
    const int32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= C * C) return;
    const int32_t o = idx / C, i = idx - o * C;
    Wt[(size_t)i * C + o] = W[idx];
}

static inline unsigned blocks_for(int64_t n, int per, int64_t cap = 148 * 16)
{
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(n, per), cap));
}

}  // namespace

int point_mlp1_simt(const float *d_feat, const float *d_points, const int64_t *d_num_points, int32_t B, int32_t N,
                    int32_t Ci, int32_t C, const float *d_W1, const float *d_b1, float *d_T, cudaStream_t st);

}  // namespace cf

extern "C" size_t cf_fusion_bwd_workspace_bytes(int32_t B, int32_t N, int32_t C, int32_t H, int32_t W, int32_t K)
{
    if (B <= 0 || N <= 0 || C <= 0 || H <= 0 || W <= 0 || K <= 0) return 0;
    return cf::carve(nullptr, B, N, C, (int64_t)H * W, K).bytes;
}

// Gradients are ACCUMULATED into d_gW1 (C,Ci+3), d_gb1, d_gW2 (C,C), d_gb2, d_gW3, d_gb3 and d_gfeat (B,N,Ci): the
// caller zero-initialises them (or passes buffers that already hold other scales' contributions).  d bev = d_gout.
extern "C" int cf_fusion_bwd(const float *d_gout, const float *d_feat, const float *d_points,
                             const int64_t *d_num_points, const int32_t *d_knn_idx, int32_t B, int32_t N, int32_t C,
                             int32_t H, int32_t W, int32_t K, float x0, float y0, float dx, float dy,
                             const float *d_W1, const float *d_b1, int32_t Ci, const float *d_W2, const float *d_b2,
                             const float *d_W3, float *d_gW1, float *d_gb1, float *d_gW2, float *d_gb2, float *d_gW3,
                             float *d_gb3, float *d_gfeat, void *d_workspace, void *stream)
{
    using namespace cf;
    CF_TRY(require_sm100());
    CF_REQUIRE(d_gout && d_feat && d_points && d_num_points && d_knn_idx && d_W1 && d_b1 && d_W2 && d_b2 && d_W3 &&
                   d_gW1 && d_gb1 && d_gW2 && d_gb2 && d_gW3 && d_gb3 && d_gfeat && d_workspace,
               CF_ERR_ARG, "cf_fusion_bwd: null pointer");
    CF_REQUIRE(B > 0 && B <= 65535 && N > 0 && C > 0 && H > 0 && W > 0 && K >= 1 && K <= CF_MAX_K && Ci > 0 && Ci % 4 == 0,
               CF_ERR_ARG, "cf_fusion_bwd: bad extents");
    CF_REQUIRE(aligned16(d_workspace), CF_ERR_ALIGN, "cf_fusion_bwd: workspace must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t cells = (int64_t)H * W, ncell = cells * B, rows_pf = cells * K, rows = rows_pf * B;
    BwdWs w = carve(d_workspace, B, N, C, cells, K);
    const int32_t ldw1 = Ci + 3;

    // ---- recompute the forward intermediates ------------------------------------------------------------------------
    CF_TRY(point_mlp1_simt(d_feat, d_points, d_num_points, B, N, Ci, C, d_W1, d_b1, w.T, st));
    k_bwd_h1<<<dim3(blocks_for(rows_pf, 8), B), 256, 0, st>>>(w.T, d_knn_idx, N, C, H, W, K, x0, y0, dx, dy, d_W1, Ci, w.H1,
                                                            w.row_cx, w.row_cy, w.n_valid);
    k_transpose_sq_b<<<(C * C + 255) / 256, 256, 0, st>>>(d_W2, C, w.W2t);
    // Z2 = H1 W2^T  (as  H1 [rows x C] * W2t [C x C])
    k_sgemm_nn<<<dim3((unsigned)ceil_div64(rows, 64), (unsigned)((C + 63) / 64)), 256, 0, st>>>(
        w.H1, C, w.W2t, C, w.H2, C, rows, C, C, 0, nullptr, 0);
    k_bwd_h2_pool<<<blocks_for(ncell * C, 256), 256, 0, st>>>(w.H2, d_knn_idx, d_b2, ncell, K, C, w.pooled);

    // ---- layer 3 ------------------------------------------------------------------------------------------------------
    k_nchw_to_cell_major<<<dim3((unsigned)ceil_div64(cells, 32), (unsigned)((C + 31) / 32), (unsigned)B), 256, 0, st>>>(
        d_gout, C, cells, w.G);
    const int64_t split3 = std::max<int64_t>(1024, ceil_div64(ncell, 296));
    k_sgemm_tn<<<dim3((unsigned)((C + 63) / 64), (unsigned)((C + 63) / 64), (unsigned)ceil_div64(ncell, split3)), 256, 0,
                 st>>>(w.G, C, w.pooled, C, d_gW3, C, ncell, C, C, split3);
    k_colsum<<<dim3((unsigned)ceil_div64(ncell, 1024), (unsigned)((C + 31) / 32)), 256, 0, st>>>(w.G, C, ncell, C, w.n_valid, 1.0f, d_gb3, 1);
    // dPooled = G W3   (W3 is (out, in) row-major: exactly the [K=out x N=in] operand)
    k_sgemm_nn<<<dim3((unsigned)ceil_div64(ncell, 64), (unsigned)((C + 63) / 64)), 256, 0, st>>>(
        w.G, C, d_W3, C, w.dPooled, C, ncell, C, C, 0, nullptr, 0);

    // ---- layer 2 ------------------------------------------------------------------------------------------------------
    k_bwd_dz2<<<blocks_for(rows * C, 256), 256, 0, st>>>(w.H2, w.dPooled, rows, K, C);  // H2 now holds dZ2
    const int64_t split2 = std::max<int64_t>(1024, ceil_div64(rows, 296));
    k_sgemm_tn<<<dim3((unsigned)((C + 63) / 64), (unsigned)((C + 63) / 64), (unsigned)ceil_div64(rows, split2)), 256, 0,
                 st>>>(w.H2, C, w.H1, C, d_gW2, C, rows, C, C, split2);
    k_colsum<<<dim3((unsigned)ceil_div64(rows, 1024), (unsigned)((C + 31) / 32)), 256, 0, st>>>(w.H2, C, rows, C, nullptr, 1.0f, d_gb2, 1);
    // dA = (dZ2 W2) * [H1 > 0]   written over H1 (the mask is read before the overwrite, element by element)
    k_sgemm_nn<<<dim3((unsigned)ceil_div64(rows, 64), (unsigned)((C + 63) / 64)), 256, 0, st>>>(
        w.H2, C, d_W2, C, w.H1, C, rows, C, C, 0, w.H1, C);

    // ---- layer 1: cell-side offset columns, scatter to points, point-side GEMMs -----------------------------------------
    k_colsum<<<dim3((unsigned)ceil_div64(rows, 1024), (unsigned)((C + 31) / 32)), 256, 0, st>>>(w.H1, C, rows, C, w.row_cx, -1.0f, d_gW1 + Ci, ldw1);
    k_colsum<<<dim3((unsigned)ceil_div64(rows, 1024), (unsigned)((C + 31) / 32)), 256, 0, st>>>(w.H1, C, rows, C, w.row_cy, -1.0f, d_gW1 + Ci + 1, ldw1);
    CF_TRY(cuda_status(cudaMemsetAsync(w.dT, 0, (size_t)B * N * C * sizeof(float), st), "cf_fusion_bwd memset"));
    k_bwd_scatter<<<dim3(blocks_for(rows_pf, 8), B), 256, 0, st>>>(w.H1, d_knn_idx, N, C, rows_pf, w.dT);
    const int64_t pts = (int64_t)B * N;
    const int64_t splitp = std::max<int64_t>(512, ceil_div64(pts, 148));
    // feat / points rows beyond num_points are never referenced by knn, so their dT rows are zero
    k_sgemm_tn<<<dim3((unsigned)((C + 63) / 64), (unsigned)((Ci + 63) / 64), (unsigned)ceil_div64(pts, splitp)), 256, 0,
                 st>>>(w.dT, C, d_feat, Ci, d_gW1, ldw1, pts, C, Ci, splitp);
    k_sgemm_tn<<<dim3((unsigned)((C + 63) / 64), 1, (unsigned)ceil_div64(pts, splitp)), 256, 0, st>>>(
        w.dT, C, d_points, 3, d_gW1 + Ci, ldw1, pts, C, 3, splitp);
    k_colsum<<<dim3((unsigned)ceil_div64(pts, 1024), (unsigned)((C + 31) / 32)), 256, 0, st>>>(w.dT, C, pts, C, nullptr, 1.0f, d_gb1, 1);
    // dF += dT W1[:, :Ci]
    k_sgemm_nn<<<dim3((unsigned)ceil_div64(pts, 64), (unsigned)((Ci + 63) / 64)), 256, 0, st>>>(
        w.dT, C, d_W1, ldw1, d_gfeat, Ci, pts, Ci, C, 1, nullptr, 0);
    count_launches(18);
    return launch_status("cf_fusion_bwd");
}

// adjoint of cf_point_gather: d_gimg (same logical shape / strides as the camera map) += scatter of d_gfeat (B,N,Ci)
extern "C" int cf_point_gather_bwd(const float *d_gfeat, float *d_gimg, int64_t sb, int64_t sc, int64_t sh, int64_t sw,
                                   int32_t B, int32_t Ci, int32_t Hf, int32_t Wf, const float *d_points,
                                   const float *d_uv, const float *h_calib, const int64_t *d_num_points, int32_t N,
                                   float img_w, float img_h, void *stream)
{
    using namespace cf;
    CF_TRY(require_sm100());
    CF_REQUIRE(d_gfeat && d_gimg && d_points && d_num_points, CF_ERR_ARG, "cf_point_gather_bwd: null pointer");
    CF_REQUIRE((d_uv != nullptr) != (h_calib != nullptr), CF_ERR_ARG, "cf_point_gather_bwd: pass exactly one of d_uv / h_calib");
    CF_REQUIRE(B > 0 && B <= 65535 && N > 0 && Ci > 0 && Hf > 0 && Wf > 0 && img_w > 0 && img_h > 0, CF_ERR_ARG,
               "cf_point_gather_bwd: bad extents");
    CalibB cal{};
    if (h_calib)
        for (int i = 0; i < 12; ++i) cal.m[i] = h_calib[i];
    const int blocks = (int)std::min<int64_t>(ceil_div64(N, 8), 148 * 16);
    k_point_gather_bwd<<<dim3(blocks, B), 256, 0, (cudaStream_t)stream>>>(d_gfeat, sb, sc, sh, sw, Ci, Hf, Wf, d_points, d_uv,
                                                                          cal, h_calib != nullptr, d_num_points, N,
                                                                          (float)Wf / img_w, (float)Hf / img_h, d_gimg);
    count_launches(1);
    return launch_status("cf_point_gather_bwd");
}
