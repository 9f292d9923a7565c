/*
 * cf_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the continuous-fusion hot path and of the rotated-box
 * post-process of Chanuk-Yang/Deep_Continuous_Fusion_for_Multi-Sensor_3D_Object_Detection.
 * Nothing under the product package may import, link or call this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Parity status
 *   - NMS (cfo_nms_sat), SAT, rectangle corners, rotated IoU, get_bboxes:
 *     PINNED against the reference's own Python code (test.py, separation_axis_theorem.py,
 *     IOU.py) run in the build container; golden vectors + generating script are in
 *     tests/golden/ and oracle/gen_golden.py.
 *   - Fusion layer (KNN, projection, bilinear gather, MLP, pool, add): PARITY UNPINNED.
 *     The reference leaves the layer as a TODO (model.py:199-203); this file restates the
 *     specification in SURVEY.md Appendix A (A1..A13), brute force, so that it is independent
 *     of the hashed/factorised algorithm the CUDA kernels use.
 *
 * Build: see oracle/Makefile  (-O2 -fopenmp -ffp-contract=off : no FMA contraction, so every
 * fp32 product and sum below is separately rounded, exactly like numpy / the __fmul_rn /
 * __fadd_rn sequence in the kernels).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CFO_API __attribute__((visibility("default")))
#define CFO_KMAX 64

CFO_API int cfo_version(void) { return 1; }

#ifdef _OPENMP
#include <omp.h>
/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline wants every host core, so set it explicitly. */
CFO_API int cfo_set_threads(int n) { if (n > 0) omp_set_num_threads(n); return omp_get_max_threads(); }
#else
CFO_API int cfo_set_threads(int n) { (void)n; return 1; }
#endif

/* ------------------------------------------------------------------------------------------
 * K-2  bounded-radius KNN, brute force.   SURVEY.md Appendix A1-A5.
 *   candidates: rows < n_valid of pts (N,3)                                   (A1)
 *   metric: d2 = dx*dx + dy*dy in the BEV plane, fp32, separately rounded     (A2)
 *   cell centre: cx = x0 + (float)i * dx_cell, cy = y0 + (float)j * dy_cell   (A3; i over X/H, j over Y/W)
 *   keep d2 <= r2, fewer than K -> -1                                         (A4)
 *   order: ascending (d2, idx)                                                (A5)
 * The input layout follows data_import_carla.py:263-266 (zero padded (max_num_pc,3) rows) and
 * num_points_raw (data_import_carla.py:262).
 * ---------------------------------------------------------------------------------------- */
CFO_API void cfo_knn_bruteforce(const float *pts, int32_t n_valid, int32_t H, int32_t W, float x0,
                                float y0, float dxc, float dyc, float r2, int32_t K, int32_t *idx_out,
                                int64_t cell_begin, int64_t cell_end)
{
    if (K > CFO_KMAX) K = CFO_KMAX;
    float *px = (float *)malloc(sizeof(float) * (size_t)(n_valid > 0 ? n_valid : 1));
    float *py = (float *)malloc(sizeof(float) * (size_t)(n_valid > 0 ? n_valid : 1));
    for (int32_t p = 0; p < n_valid; ++p) {
        px[p] = pts[3 * (size_t)p + 0];
        py[p] = pts[3 * (size_t)p + 1];
    }
    (void)H;
#pragma omp parallel
    {
        float dbuf[512];
#pragma omp for schedule(dynamic, 256)
        for (int64_t cell = cell_begin; cell < cell_end; ++cell) {
            const int32_t i = (int32_t)(cell / W), j = (int32_t)(cell % W);
            const float cx = x0 + (float)i * dxc;
            const float cy = y0 + (float)j * dyc;
            float bd[CFO_KMAX];
            int32_t bi[CFO_KMAX];
            int32_t cnt = 0;
            for (int32_t base = 0; base < n_valid; base += 512) {
                const int32_t m = (n_valid - base) < 512 ? (n_valid - base) : 512;
                for (int32_t q = 0; q < m; ++q) { /* vectorisable */
                    const float ddx = px[base + q] - cx;
                    const float ddy = py[base + q] - cy;
                    dbuf[q] = ddx * ddx + ddy * ddy;
                }
                for (int32_t q = 0; q < m; ++q) {
                    const float d2 = dbuf[q];
                    if (!(d2 <= r2)) continue;
                    /* points are visited in ascending index, so an equal d2 never displaces */
                    if (cnt == K && !(d2 < bd[K - 1])) continue;
                    int32_t pos = (cnt < K) ? cnt : K - 1;
                    while (pos > 0 && bd[pos - 1] > d2) {
                        bd[pos] = bd[pos - 1];
                        bi[pos] = bi[pos - 1];
                        --pos;
                    }
                    bd[pos] = d2;
                    bi[pos] = base + q;
                    if (cnt < K) ++cnt;
                }
            }
            int32_t *o = idx_out + (size_t)(cell - cell_begin) * (size_t)K;
            for (int32_t k = 0; k < K; ++k) o[k] = (k < cnt) ? bi[k] : -1;
        }
    }
    free(px);
    free(py);
}

/* ------------------------------------------------------------------------------------------
 * K-3  projection + bilinear gather, per point.   Appendix A6, A7.
 *   q = [x y z 1] @ CRT (4x3, row-vector convention of data_import_carla.py:197-201),
 *   u = q0/q2, v = q1/q2; or the dataset's precomputed projected_loc_uv (:265-266).
 *   uf = (u+0.5)*Wf/img_w - 0.5, vf = (v+0.5)*Hf/img_h - 0.5; 4 taps, outside reads 0
 *   ( == torch grid_sample(mode=bilinear, padding_mode=zeros, align_corners=False) ).
 * img is (Ci,Hf,Wf) channel-major fp32.  out is (n_valid, Ci).
 * ---------------------------------------------------------------------------------------- */
CFO_API void cfo_project_points(const float *pts, int32_t n, const float *calib, float *uv_out)
{
    for (int32_t p = 0; p < n; ++p) {
        const float x = pts[3 * (size_t)p], y = pts[3 * (size_t)p + 1], z = pts[3 * (size_t)p + 2];
        float q[3];
        for (int c = 0; c < 3; ++c)
            q[c] = ((x * calib[0 * 3 + c] + y * calib[1 * 3 + c]) + z * calib[2 * 3 + c]) + calib[3 * 3 + c];
        uv_out[2 * (size_t)p + 0] = q[0] / q[2];
        uv_out[2 * (size_t)p + 1] = q[1] / q[2];
    }
}

CFO_API void cfo_gather_points(const float *img, int32_t Ci, int32_t Hf, int32_t Wf, const float *uv,
                               int32_t n, float img_w, float img_h, float *out)
{
    const size_t plane = (size_t)Hf * (size_t)Wf;
#pragma omp parallel for schedule(static)
    for (int32_t p = 0; p < n; ++p) {
        const float u = uv[2 * (size_t)p], v = uv[2 * (size_t)p + 1];
        const float uf = (u + 0.5f) * ((float)Wf / img_w) - 0.5f;
        const float vf = (v + 0.5f) * ((float)Hf / img_h) - 0.5f;
        float *o = out + (size_t)p * (size_t)Ci;
        if (!(uf > -1.0f && uf < (float)Wf && vf > -1.0f && vf < (float)Hf)) { /* also rejects NaN */
            for (int32_t c = 0; c < Ci; ++c) o[c] = 0.0f;
            continue;
        }
        const float fx = floorf(uf), fy = floorf(vf);
        const int32_t ix = (int32_t)fx, iy = (int32_t)fy;
        const float wx1 = uf - fx, wy1 = vf - fy;
        const float wx0 = 1.0f - wx1, wy0 = 1.0f - wy1;
        const int okx0 = ix >= 0 && ix < Wf, okx1 = ix + 1 >= 0 && ix + 1 < Wf;
        const int oky0 = iy >= 0 && iy < Hf, oky1 = iy + 1 >= 0 && iy + 1 < Hf;
        const float w00 = (okx0 && oky0) ? wx0 * wy0 : 0.0f;
        const float w01 = (okx1 && oky0) ? wx1 * wy0 : 0.0f;
        const float w10 = (okx0 && oky1) ? wx0 * wy1 : 0.0f;
        const float w11 = (okx1 && oky1) ? wx1 * wy1 : 0.0f;
        const size_t o00 = (size_t)(oky0 ? iy : 0) * Wf + (size_t)(okx0 ? ix : 0);
        const size_t o01 = (size_t)(oky0 ? iy : 0) * Wf + (size_t)(okx1 ? ix + 1 : 0);
        const size_t o10 = (size_t)(oky1 ? iy + 1 : 0) * Wf + (size_t)(okx0 ? ix : 0);
        const size_t o11 = (size_t)(oky1 ? iy + 1 : 0) * Wf + (size_t)(okx1 ? ix + 1 : 0);
        for (int32_t c = 0; c < Ci; ++c) {
            const float *pl = img + (size_t)c * plane;
            o[c] = ((w00 * pl[o00] + w01 * pl[o01]) + w10 * pl[o10]) + w11 * pl[o11];
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * K-4  per-neighbour MLP + K-sum-pool + BEV add, naive formulation.   Appendix A8-A10.
 *   x   = [ f_j (Ci) , px-cx, py-cy, pz ]                                  (A8)
 *   h1  = relu(W1 x + b1); h2 = relu(W2 h1 + b2); y = W3 h2 + b3           (A9; W are (out,in) row-major
 *                                                                            like nn.Linear.weight)
 *   out = bev + sum_{k: idx != -1} y                                       (A10)
 * bev / out are (C,H,W) fp32 (the layout of model.py:73-79 feature maps, one frame).
 * feat is the (n_valid,Ci) table from cfo_gather_points (a pure per-point function, cached).
 * ---------------------------------------------------------------------------------------- */
CFO_API void cfo_fusion_mlp(const float *bev, int32_t C, int32_t H, int32_t W, const float *feat, int32_t Ci,
                            const float *pts, const int32_t *knn, int32_t K, float x0, float y0, float dxc,
                            float dyc, const float *W1, const float *b1, const float *W2, const float *b2,
                            const float *W3, const float *b3, float *out, int64_t cell_begin,
                            int64_t cell_end)
{
    const int32_t Cin = Ci + 3;
    const size_t plane = (size_t)H * (size_t)W;
#pragma omp parallel
    {
        float *x = (float *)malloc(sizeof(float) * (size_t)Cin);
        float *h1 = (float *)malloc(sizeof(float) * (size_t)C);
        float *h2 = (float *)malloc(sizeof(float) * (size_t)C);
        float *acc = (float *)malloc(sizeof(float) * (size_t)C);
#pragma omp for schedule(dynamic, 64)
        for (int64_t cell = cell_begin; cell < cell_end; ++cell) {
            const int32_t i = (int32_t)(cell / W), j = (int32_t)(cell % W);
            const float cx = x0 + (float)i * dxc;
            const float cy = y0 + (float)j * dyc;
            for (int32_t c = 0; c < C; ++c) acc[c] = 0.0f;
            for (int32_t k = 0; k < K; ++k) {
                const int32_t p = knn[(size_t)(cell - cell_begin) * (size_t)K + k];
                if (p < 0) continue;
                memcpy(x, feat + (size_t)p * (size_t)Ci, sizeof(float) * (size_t)Ci);
                x[Ci + 0] = pts[3 * (size_t)p + 0] - cx;
                x[Ci + 1] = pts[3 * (size_t)p + 1] - cy;
                x[Ci + 2] = pts[3 * (size_t)p + 2];
                for (int32_t o = 0; o < C; ++o) {
                    const float *w = W1 + (size_t)o * (size_t)Cin;
                    float s = 0.0f;
                    for (int32_t q = 0; q < Cin; ++q) s += w[q] * x[q];
                    s += b1[o];
                    h1[o] = s > 0.0f ? s : 0.0f;
                }
                for (int32_t o = 0; o < C; ++o) {
                    const float *w = W2 + (size_t)o * (size_t)C;
                    float s = 0.0f;
                    for (int32_t q = 0; q < C; ++q) s += w[q] * h1[q];
                    s += b2[o];
                    h2[o] = s > 0.0f ? s : 0.0f;
                }
                for (int32_t o = 0; o < C; ++o) {
                    const float *w = W3 + (size_t)o * (size_t)C;
                    float s = 0.0f;
                    for (int32_t q = 0; q < C; ++q) s += w[q] * h2[q];
                    acc[o] += s + b3[o];
                }
            }
            for (int32_t c = 0; c < C; ++c)
                out[(size_t)c * plane + (size_t)cell] = bev[(size_t)c * plane + (size_t)cell] + acc[c];
        }
        free(x);
        free(h1);
        free(h2);
        free(acc);
    }
}

/* ------------------------------------------------------------------------------------------
 * P-2  get_vertice_rect   (separation_axis_theorem.py:82-94), as executed under numpy >= 2:
 *      inputs are np.float32 scalars (test.py:157-159), math.cos/sin evaluate in fp64 and the
 *      Python float result is a *weak* scalar, so it is cast to fp32 and every product / sum
 *      is fp32.  L = size[0], W = size[1]  (:87-88); order v1,v2,v3,v4 (:89-93).
 * ---------------------------------------------------------------------------------------- */
typedef struct { float x[4], y[4]; } cfo_quad;

static void cfo_rect(const float *box, cfo_quad *q)
{
    const float cx = box[0], cy = box[1], L = box[3], Wd = box[4], yaw = box[6];
    const float c = (float)cos((double)yaw), s = (float)sin((double)yaw);
    const float Lh = L / 2.0f, Wh = Wd / 2.0f;
    const float a = Lh * c, b = Wh * s, e = Lh * s, f = Wh * c;
    q->x[0] = cx + (-a + b); q->y[0] = cy + (-e - f); /* vertex_1 :91 */
    q->x[1] = cx + (a + b);  q->y[1] = cy + (e - f);  /* vertex_2 :92 */
    q->x[2] = cx + (a - b);  q->y[2] = cy + (e + f);  /* vertex_3 :89 */
    q->x[3] = cx + (-a - b); q->y[3] = cy + (-e + f); /* vertex_4 :90 */
}

CFO_API void cfo_get_vertice_rect(const float *box7, float *xy8)
{
    cfo_quad q;
    cfo_rect(box7, &q);
    for (int i = 0; i < 4; ++i) { xy8[2 * i] = q.x[i]; xy8[2 * i + 1] = q.y[i]; }
}

/* P-3  separating_axis_theorem (separation_axis_theorem.py:66-80) with
 *      normalize (:26-28), orthogonal (:36-37), project (:43-45), overlap/contains (:47-64).
 *      sq_mode 0: v*v in fp32 (correctly rounded square; what the CUDA kernel does)
 *      sq_mode 1: powf(v,2) -- what numpy's np.float32.__pow__ calls; within 1 ulp of v*v and
 *                 libm-build dependent; kept to show the oracle reproduces the reference's
 *                 intermediates bit for bit on this box.                                      */
static int cfo_sq_mode = 0;
CFO_API void cfo_set_sq_mode(int m) { cfo_sq_mode = m; }

static inline float cfo_sq(float v) { return cfo_sq_mode ? powf(v, 2.0f) : v * v; }

static void cfo_axes(const cfo_quad *q, float *ax, float *ay)
{
    for (int i = 0; i < 4; ++i) {
        const int n = (i + 1) & 3;
        const float ex = q->x[n] - q->x[i], ey = q->y[n] - q->y[i]; /* edge_direction :33-34 */
        const float ox = ey, oy = -ex;                               /* orthogonal :36-37    */
        const float s = cfo_sq(ox) + cfo_sq(oy);
        const float norm = (float)sqrt((double)s); /* math.sqrt -> weak python float -> fp32 */
        ax[i] = ox / norm;
        ay[i] = oy / norm;
    }
}

static inline void cfo_project(const cfo_quad *q, float ax, float ay, float *lo, float *hi)
{
    float mn = 0, mx = 0;
    for (int v = 0; v < 4; ++v) {
        const float d = q->x[v] * ax + q->y[v] * ay; /* dot :30-31 */
        if (v == 0) { mn = mx = d; }
        else { if (d < mn) mn = d; if (d > mx) mx = d; } /* builtin min/max keep the first on ties */
    }
    *lo = mn; *hi = mx;
}

static inline int cfo_contains(float n, float a, float b)
{
    if (b < a) { float t = a; a = b; b = t; }
    return (n >= a) && (n <= b);
}

static int cfo_sat_quads(const cfo_quad *A, const cfo_quad *B)
{
    float ax[8], ay[8];
    cfo_axes(A, ax, ay);
    cfo_axes(B, ax + 4, ay + 4);
    for (int i = 0; i < 8; ++i) {
        float a0, a1, b0, b1;
        cfo_project(A, ax[i], ay[i], &a0, &a1);
        cfo_project(B, ax[i], ay[i], &b0, &b1);
        const int ov = cfo_contains(a0, b0, b1) || cfo_contains(a1, b0, b1) || cfo_contains(b0, a0, a1) ||
                       cfo_contains(b1, a0, a1);
        if (!ov) return 0;
    }
    return 1;
}

CFO_API int cfo_sat_overlap(const float *box_a, const float *box_b)
{
    cfo_quad A, B;
    cfo_rect(box_a, &A);
    cfo_rect(box_b, &B);
    return cfo_sat_quads(&A, &B);
}

CFO_API void cfo_sat_axes(const float *box7, float *axes8)
{
    cfo_quad q;
    float ax[4], ay[4];
    cfo_rect(box7, &q);
    cfo_axes(&q, ax, ay);
    for (int i = 0; i < 4; ++i) { axes8[2 * i] = ax[i]; axes8[2 * i + 1] = ay[i]; }
}

/* generic convex-polygon form of P-3 (what separation_axis_theorem.py:66-80 accepts); used to pin the
 * known answers printed by its main() (:98-105 -> True True True).  fp32 like the box path. */
CFO_API int cfo_sat_polygons(const float *va, int32_t na, const float *vb, int32_t nb)
{
    for (int pass = 0; pass < 2; ++pass) {
        const float *src = pass == 0 ? va : vb;
        const int32_t n = pass == 0 ? na : nb;
        for (int32_t i = 0; i < n; ++i) {
            const int32_t nx = (i + 1) % n;
            const float ex = src[2 * nx] - src[2 * i], ey = src[2 * nx + 1] - src[2 * i + 1];
            const float ox = ey, oy = -ex;
            const float norm = (float)sqrt((double)(cfo_sq(ox) + cfo_sq(oy)));
            const float ax = ox / norm, ay = oy / norm;
            float a0 = 0, a1 = 0, b0 = 0, b1 = 0;
            for (int32_t v = 0; v < na; ++v) {
                const float d = va[2 * v] * ax + va[2 * v + 1] * ay;
                if (v == 0) a0 = a1 = d; else { if (d < a0) a0 = d; if (d > a1) a1 = d; }
            }
            for (int32_t v = 0; v < nb; ++v) {
                const float d = vb[2 * v] * ax + vb[2 * v + 1] * ay;
                if (v == 0) b0 = b1 = d; else { if (d < b0) b0 = d; if (d > b1) b1 = d; }
            }
            const int ov = cfo_contains(a0, b0, b1) || cfo_contains(a1, b0, b1) || cfo_contains(b0, a0, a1) ||
                           cfo_contains(b1, a0, a1);
            if (!ov) return 0;
        }
    }
    return 1;
}

/* P-4  Test.NMS_SAT (test.py:142-175): greedy in input order; box i is kept iff it overlaps no
 *      previously kept box.  Returns the number kept; keep[] = ascending input indices.       */
CFO_API int32_t cfo_nms_sat(const float *boxes, int32_t n, int32_t *keep)
{
    cfo_quad *kq = (cfo_quad *)malloc(sizeof(cfo_quad) * (size_t)(n > 0 ? n : 1));
    int32_t nk = 0;
    for (int32_t i = 0; i < n; ++i) {
        cfo_quad q;
        cfo_rect(boxes + 7 * (size_t)i, &q);
        int hit = 0;
        for (int32_t j = 0; j < nk && !hit; ++j) hit = cfo_sat_quads(&q, &kq[j]);
        if (!hit) { kq[nk] = q; keep[nk] = i; ++nk; }
    }
    free(kq);
    return nk;
}

/* full overlap matrix row-major (n,n) uint8, for mask-level parity tests */
CFO_API void cfo_sat_matrix(const float *boxes, int32_t n, uint8_t *m)
{
    cfo_quad *q = (cfo_quad *)malloc(sizeof(cfo_quad) * (size_t)(n > 0 ? n : 1));
    for (int32_t i = 0; i < n; ++i) cfo_rect(boxes + 7 * (size_t)i, &q[i]);
#pragma omp parallel for schedule(dynamic, 8)
    for (int32_t i = 0; i < n; ++i)
        for (int32_t j = 0; j < n; ++j) m[(size_t)i * n + j] = (uint8_t)cfo_sat_quads(&q[i], &q[j]);
    free(q);
}

/* ------------------------------------------------------------------------------------------
 * P-5  get_3d_box (IOU.py:127-155) as called from test.py:122-125,185-188: centre / size are
 *      float32 arrays, heading a 0-d float32 array.  np.cos/np.sin of it are float32 (numpy's
 *      own float32 kernels, <= ~1.5 ulp; we use the correctly rounded value), the rotation
 *      matrix is promoted to float64 by the python-int entries, half sizes are float32, the
 *      product and everything after it is float64.
 *      Axis convention (reference behaviour, SURVEY 8a "convention trap"): rotation about axis 1,
 *      l,w,h along axes 0,2,1.
 * P-6  box3d_iou (IOU.py:91-120): polygon = corners 3,2,1,0 on (axis0, axis2); Sutherland-Hodgman
 *      clip (:9-56, strict '>' inside test); area of the clipped convex polygon (the reference
 *      asks Qhull for the hull "volume", :64-74 -- equal to the shoelace area of a convex polygon);
 *      height overlap on axis 1 (:112-115); volumes from edge lengths (:77-82).
 * ---------------------------------------------------------------------------------------- */
static void cfo_box_corners(const float *box, double c3[8][3])
{
    const float ang = box[6];
    const double c = (double)(float)cos((double)ang), s = (double)(float)sin((double)ang);
    const float l = box[3], w = box[4], h = box[5];
    const float lh = l / 2.0f, wh = w / 2.0f, hh = h / 2.0f;
    const float xs[8] = { lh, lh, -lh, -lh, lh, lh, -lh, -lh };
    const float ys[8] = { hh, hh, hh, hh, -hh, -hh, -hh, -hh };
    const float zs[8] = { wh, -wh, -wh, wh, wh, -wh, -wh, wh };
    for (int i = 0; i < 8; ++i) {
        const double X = (double)xs[i], Y = (double)ys[i], Z = (double)zs[i];
        c3[i][0] = (c * X + 0.0 * Y) + s * Z + (double)box[0];
        c3[i][1] = (0.0 * X + 1.0 * Y) + 0.0 * Z + (double)box[1];
        c3[i][2] = (-s * X + 0.0 * Y) + c * Z + (double)box[2];
    }
}

CFO_API void cfo_get_3d_box(const float *box7, double *corners24)
{
    double c3[8][3];
    cfo_box_corners(box7, c3);
    memcpy(corners24, c3, sizeof(c3));
}

static int cfo_clip(const double (*subj)[2], int ns, const double (*clip)[2], int nc, double (*out)[2])
{
    double bufA[16][2], bufB[16][2];
    double (*in)[2] = bufA, (*op)[2] = bufB;
    int n_in = ns;
    memcpy(in, subj, sizeof(double) * 2 * (size_t)ns);
    double cp1x = clip[nc - 1][0], cp1y = clip[nc - 1][1];
    for (int ci = 0; ci < nc; ++ci) {
        const double cp2x = clip[ci][0], cp2y = clip[ci][1];
        int n_out = 0;
        double sx = in[n_in - 1][0], sy = in[n_in - 1][1];
        for (int v = 0; v < n_in; ++v) {
            const double ex = in[v][0], ey = in[v][1];
            const int e_in = (cp2x - cp1x) * (ey - cp1y) > (cp2y - cp1y) * (ex - cp1x);
            const int s_in = (cp2x - cp1x) * (sy - cp1y) > (cp2y - cp1y) * (sx - cp1x);
            if (e_in != s_in) {
                const double dcx = cp1x - cp2x, dcy = cp1y - cp2y;
                const double dpx = sx - ex, dpy = sy - ey;
                const double n1 = cp1x * cp2y - cp1y * cp2x;
                const double n2 = sx * ey - sy * ex;
                const double n3 = 1.0 / (dcx * dpy - dcy * dpx);
                op[n_out][0] = (n1 * dpx - n2 * dcx) * n3;
                op[n_out][1] = (n1 * dpy - n2 * dcy) * n3;
                ++n_out;
            }
            if (e_in) { op[n_out][0] = ex; op[n_out][1] = ey; ++n_out; }
            sx = ex; sy = ey;
        }
        cp1x = cp2x; cp1y = cp2y;
        if (n_out == 0) return 0;
        double (*t)[2] = in; in = op; op = t;
        n_in = n_out;
    }
    memcpy(out, in, sizeof(double) * 2 * (size_t)n_in);
    return n_in;
}

static double cfo_shoelace(const double (*p)[2], int n)
{
    /* poly_area (IOU.py:59-61): 0.5*|dot(x, roll(y,1)) - dot(y, roll(x,1))| */
    double a = 0.0, b = 0.0;
    for (int i = 0; i < n; ++i) {
        const int pr = (i + n - 1) % n;
        a += p[i][0] * p[pr][1];
        b += p[i][1] * p[pr][0];
    }
    return 0.5 * fabs(a - b);
}

/* Area of the convex hull of a point set (Andrew monotone chain + shoelace).  This is what
 * scipy.spatial.ConvexHull(...).volume returns for 2-D input (IOU.py:71-72).  For a well-formed clip result
 * it equals the polygon's shoelace area; for degenerate clips (identical / edge-aligned boxes, where
 * computeIntersection divides by ~0) it is what keeps the oracle on the reference's value. */
static double cfo_hull_area(const double (*p)[2], int n)
{
    double q[16][2], h[34][2];
    if (n < 3) return 0.0;
    for (int i = 0; i < n; ++i) { q[i][0] = p[i][0]; q[i][1] = p[i][1]; }
    for (int i = 1; i < n; ++i) { /* insertion sort by (x, y) */
        const double x = q[i][0], y = q[i][1];
        int j = i - 1;
        while (j >= 0 && (q[j][0] > x || (q[j][0] == x && q[j][1] > y))) { q[j + 1][0] = q[j][0]; q[j + 1][1] = q[j][1]; --j; }
        q[j + 1][0] = x; q[j + 1][1] = y;
    }
    int k = 0;
    for (int i = 0; i < n; ++i) {
        while (k >= 2 && (h[k - 1][0] - h[k - 2][0]) * (q[i][1] - h[k - 2][1]) -
                                 (h[k - 1][1] - h[k - 2][1]) * (q[i][0] - h[k - 2][0]) <= 0.0) --k;
        h[k][0] = q[i][0]; h[k][1] = q[i][1]; ++k;
    }
    const int lower = k + 1;
    for (int i = n - 2; i >= 0; --i) {
        while (k >= lower && (h[k - 1][0] - h[k - 2][0]) * (q[i][1] - h[k - 2][1]) -
                                     (h[k - 1][1] - h[k - 2][1]) * (q[i][0] - h[k - 2][0]) <= 0.0) --k;
        h[k][0] = q[i][0]; h[k][1] = q[i][1]; ++k;
    }
    --k; /* last point repeats the first */
    if (k < 3) return 0.0;
    double a = 0.0;
    for (int i = 0; i < k; ++i) {
        const int nx = (i + 1) % k;
        a += h[i][0] * h[nx][1] - h[nx][0] * h[i][1];
    }
    return 0.5 * fabs(a);
}

static double cfo_dist3(const double *a, const double *b)
{
    const double d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
    return sqrt(d0 * d0 + d1 * d1 + d2 * d2);
}

/* P-6 proper: corners in, (iou3d, iou2d) out -- fed with reference-made corners it reproduces the
 * IOU.py:161-168 known answer to 1e-12. */
CFO_API void cfo_box3d_iou_corners(const double *corners1, const double *corners2, double *iou3d, double *iou2d)
{
    const double (*c1)[3] = (const double (*)[3])corners1;
    const double (*c2)[3] = (const double (*)[3])corners2;
    double r1[4][2], r2[4][2];
    for (int i = 0; i < 4; ++i) {
        r1[i][0] = c1[3 - i][0]; r1[i][1] = c1[3 - i][2];
        r2[i][0] = c2[3 - i][0]; r2[i][1] = c2[3 - i][2];
    }
    const double area1 = cfo_shoelace(r1, 4), area2 = cfo_shoelace(r2, 4);
    double inter[16][2];
    const int ni = cfo_clip(r1, 4, r2, 4, inter);
    const double inter_area = ni > 0 ? cfo_hull_area(inter, ni) : 0.0;
    *iou2d = inter_area / (area1 + area2 - inter_area);
    const double ymax = c1[0][1] < c2[0][1] ? c1[0][1] : c2[0][1];
    const double ymin = c1[4][1] > c2[4][1] ? c1[4][1] : c2[4][1];
    const double hov = (ymax - ymin) > 0.0 ? (ymax - ymin) : 0.0;
    const double inter_vol = inter_area * hov;
    const double vol1 = cfo_dist3(c1[0], c1[1]) * cfo_dist3(c1[1], c1[2]) * cfo_dist3(c1[0], c1[4]);
    const double vol2 = cfo_dist3(c2[0], c2[1]) * cfo_dist3(c2[1], c2[2]) * cfo_dist3(c2[0], c2[4]);
    *iou3d = inter_vol / (vol1 + vol2 - inter_vol);
}

/* nudge: test.py:129 adds 1e-4 (fp32) to the centre of box_b in NMS_IOU; 0 elsewhere. */
CFO_API void cfo_box3d_iou(const float *box_a, const float *box_b, float nudge_b, double *iou3d, double *iou2d)
{
    float bb[7];
    memcpy(bb, box_b, sizeof(bb));
    if (nudge_b != 0.0f) { bb[0] = bb[0] + nudge_b; bb[1] = bb[1] + nudge_b; bb[2] = bb[2] + nudge_b; }
    double c1[8][3], c2[8][3];
    cfo_box_corners(box_a, c1);
    cfo_box_corners(bb, c2);
    cfo_box3d_iou_corners(&c1[0][0], &c2[0][0], iou3d, iou2d);
}

CFO_API void cfo_box3d_iou_matrix(const float *boxes_a, int32_t na, const float *boxes_b, int32_t nb,
                                  float nudge_b, double *iou3d, double *iou2d)
{
#pragma omp parallel for schedule(dynamic, 4)
    for (int32_t i = 0; i < na; ++i)
        for (int32_t j = 0; j < nb; ++j)
            cfo_box3d_iou(boxes_a + 7 * (size_t)i, boxes_b + 7 * (size_t)j, nudge_b,
                          iou3d + (size_t)i * nb + j, iou2d + (size_t)i * nb + j);
}

/* P-7  Test.NMS_IOU (test.py:110-140): same greedy rule with predicate iou3d > thr, kept box
 *      centre nudged by +1e-4.                                                               */
CFO_API int32_t cfo_nms_iou(const float *boxes, int32_t n, float thr, int32_t *keep)
{
    int32_t nk = 0;
    for (int32_t i = 0; i < n; ++i) {
        int hit = 0;
        for (int32_t j = 0; j < nk && !hit; ++j) {
            double i3, i2;
            cfo_box3d_iou(boxes + 7 * (size_t)i, boxes + 7 * (size_t)keep[j], 0.0001f, &i3, &i2);
            hit = i3 > (double)thr;
        }
        if (!hit) keep[nk++] = i;
    }
    return nk;
}

/* P-1  Test.get_bboxes (test.py:88-108), one frame: for anchor a in {0,1}: cells (row-major)
 *      whose cls[2a+1] > thr, gather the 7 decoded channels [7a,7a+7).  Returns the count.   */
CFO_API int32_t cfo_get_bboxes(const float *cls4, const float *box14, int32_t H, int32_t W, float thr,
                               float *boxes_out)
{
    const size_t plane = (size_t)H * (size_t)W;
    int32_t n = 0;
    for (int a = 0; a < 2; ++a)
        for (size_t cell = 0; cell < plane; ++cell)
            if (cls4[(size_t)(2 * a + 1) * plane + cell] > thr) {
                for (int c = 0; c < 7; ++c) boxes_out[7 * (size_t)n + c] = box14[(size_t)(7 * a + c) * plane + cell];
                ++n;
            }
    return n;
}
