// cf_boxgeom.cuh -- device-side rotated-box geometry mirroring the reference's arithmetic.
//
//  * SAT path (separation_axis_theorem.py:26-94 as called from test.py:157-168): the reference runs on
//    np.float32 scalars under numpy >= 2, where Python floats are weak scalars: math.cos/sin/sqrt are
//    evaluated in fp64, their result is cast to fp32 and every product / sum / quotient is fp32.
//    We reproduce that sequence with explicitly rounded intrinsics (no FMA contraction).
//    One deliberate difference: `v ** 2` on an np.float32 calls libm powf, which is within 1 ulp of the
//    correctly rounded square and differs from it in ~0.07% of evaluations depending on the libm build;
//    we use the correctly rounded v*v ("SAT numerics").
//  * IoU path (IOU.py:9-155 as called from test.py:122-133,185-196): float32 half sizes and float32
//    cos/sin, everything downstream in fp64.
#pragma once
#include "cf_common.cuh"

namespace cf {

struct SatBox {
    float vx[4], vy[4];  // vertex_1..vertex_4            separation_axis_theorem.py:89-93
    float ax[4], ay[4];  // normalised edge normals        :67-72, :26-28, :36-37
    float lo[4], hi[4];  // own projection on own axes     :43-45
};

__device__ __forceinline__ void sat_project(const float (&vx)[4], const float (&vy)[4], float ax, float ay, float &lo,
                                            float &hi)
{
    float d = __fadd_rn(__fmul_rn(vx[0], ax), __fmul_rn(vy[0], ay));
    lo = d;
    hi = d;
#pragma unroll
    for (int v = 1; v < 4; ++v) {
        d = __fadd_rn(__fmul_rn(vx[v], ax), __fmul_rn(vy[v], ay));
        lo = d < lo ? d : lo;
        hi = d > hi ? d : hi;
    }
}

__device__ __forceinline__ void sat_prepare(const float *__restrict__ box, SatBox &s)
{
    const float cx = box[0], cy = box[1], L = box[3], Wd = box[4], yaw = box[6];
    const float c = (float)cos((double)yaw), sn = (float)sin((double)yaw);
    const float Lh = __fdiv_rn(L, 2.0f), Wh = __fdiv_rn(Wd, 2.0f);
    const float a = __fmul_rn(Lh, c), b = __fmul_rn(Wh, sn), e = __fmul_rn(Lh, sn), f = __fmul_rn(Wh, c);
    s.vx[0] = __fadd_rn(cx, __fadd_rn(-a, b));  s.vy[0] = __fadd_rn(cy, __fsub_rn(-e, f));
    s.vx[1] = __fadd_rn(cx, __fadd_rn(a, b));   s.vy[1] = __fadd_rn(cy, __fsub_rn(e, f));
    s.vx[2] = __fadd_rn(cx, __fsub_rn(a, b));   s.vy[2] = __fadd_rn(cy, __fadd_rn(e, f));
    s.vx[3] = __fadd_rn(cx, __fsub_rn(-a, b));  s.vy[3] = __fadd_rn(cy, __fadd_rn(-e, f));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = (i + 1) & 3;
        const float ex = __fsub_rn(s.vx[n], s.vx[i]), ey = __fsub_rn(s.vy[n], s.vy[i]);
        const float ox = ey, oy = -ex;
        const float sq = __fadd_rn(__fmul_rn(ox, ox), __fmul_rn(oy, oy));
        const float norm = (float)sqrt((double)sq);
        s.ax[i] = __fdiv_rn(ox, norm);
        s.ay[i] = __fdiv_rn(oy, norm);
        sat_project(s.vx, s.vy, s.ax[i], s.ay[i], s.lo[i], s.hi[i]);
    }
}

// closed-interval overlap == the four `contains` tests of separation_axis_theorem.py:47-64
__device__ __forceinline__ bool sat_overlap(const SatBox &A, const SatBox &B)
{
    bool ov = true;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float lo, hi;
        sat_project(B.vx, B.vy, A.ax[i], A.ay[i], lo, hi);
        ov = ov && (A.lo[i] <= hi) && (lo <= A.hi[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float lo, hi;
        sat_project(A.vx, A.vy, B.ax[i], B.ay[i], lo, hi);
        ov = ov && (lo <= B.hi[i]) && (B.lo[i] <= hi);
    }
    return ov;
}

// ----------------------------------------------------------------------------------------- IoU
struct IouBox {
    double rx[4], rz[4];  // corners 3,2,1,0 on axes (0,2)     IOU.py:104-105
    double ytop, ybot;    // corners[0,1], corners[4,1]        IOU.py:112-113
    double area, vol;     // poly_area / box3d_vol             IOU.py:59-61,77-82
};

__device__ __forceinline__ double shoelace(const double *x, const double *y, int n)
{
    double a = 0.0, b = 0.0;
    for (int i = 0; i < n; ++i) {
        const int pr = i == 0 ? n - 1 : i - 1;
        a += x[i] * y[pr];
        b += y[i] * x[pr];
    }
    return 0.5 * fabs(a - b);
}

__device__ __forceinline__ void iou_prepare(const float *__restrict__ box, float nudge, IouBox &o)
{
    const float bx = nudge != 0.0f ? __fadd_rn(box[0], nudge) : box[0];
    const float by = nudge != 0.0f ? __fadd_rn(box[1], nudge) : box[1];
    const float bz = nudge != 0.0f ? __fadd_rn(box[2], nudge) : box[2];
    const float ang = box[6];
    const double c = (double)(float)cos((double)ang), s = (double)(float)sin((double)ang);
    const float lh = __fdiv_rn(box[3], 2.0f), wh = __fdiv_rn(box[4], 2.0f), hh = __fdiv_rn(box[5], 2.0f);
    const float xs[4] = {lh, lh, -lh, -lh};   // corners 0..3 (IOU.py:147)
    const float zs[4] = {wh, -wh, -wh, wh};   // (IOU.py:149)
    double X[4], Z[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        X[i] = (c * (double)xs[i] + s * (double)zs[i]) + (double)bx;
        Z[i] = (-s * (double)xs[i] + c * (double)zs[i]) + (double)bz;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        o.rx[i] = X[3 - i];
        o.rz[i] = Z[3 - i];
    }
    o.ytop = (double)hh + (double)by;
    o.ybot = (double)(-hh) + (double)by;
    o.area = shoelace(o.rx, o.rz, 4);
    // box3d_vol: |c0-c1| * |c1-c2| * |c0-c4|
    const double e01 = sqrt((X[0] - X[1]) * (X[0] - X[1]) + (Z[0] - Z[1]) * (Z[0] - Z[1]));
    const double e12 = sqrt((X[1] - X[2]) * (X[1] - X[2]) + (Z[1] - Z[2]) * (Z[1] - Z[2]));
    const double e04 = sqrt((o.ytop - o.ybot) * (o.ytop - o.ybot));
    o.vol = e01 * e12 * e04;
}

// Area of the convex hull of <= 10 points (Andrew monotone chain + shoelace): the value the reference gets
// from scipy.spatial.ConvexHull(inter_p).volume (IOU.py:71-72).  Equal to the polygon area for a well-formed
// clip; keeps parity with the reference on degenerate clips (identical / edge-aligned boxes).
__device__ __forceinline__ double hull_area(const double *x, const double *y, int n)
{
    if (n < 3) return 0.0;
    double qx[10], qy[10], hx[22], hy[22];
    for (int i = 0; i < n; ++i) {
        qx[i] = x[i];
        qy[i] = y[i];
    }
    for (int i = 1; i < n; ++i) {
        const double vx = qx[i], vy = qy[i];
        int j = i - 1;
        while (j >= 0 && (qx[j] > vx || (qx[j] == vx && qy[j] > vy))) {
            qx[j + 1] = qx[j];
            qy[j + 1] = qy[j];
            --j;
        }
        qx[j + 1] = vx;
        qy[j + 1] = vy;
    }
    int k = 0;
    for (int i = 0; i < n; ++i) {
        while (k >= 2 && (hx[k - 1] - hx[k - 2]) * (qy[i] - hy[k - 2]) - (hy[k - 1] - hy[k - 2]) * (qx[i] - hx[k - 2]) <= 0.0) --k;
        hx[k] = qx[i];
        hy[k] = qy[i];
        ++k;
    }
    const int lower = k + 1;
    for (int i = n - 2; i >= 0; --i) {
        while (k >= lower && (hx[k - 1] - hx[k - 2]) * (qy[i] - hy[k - 2]) - (hy[k - 1] - hy[k - 2]) * (qx[i] - hx[k - 2]) <= 0.0) --k;
        hx[k] = qx[i];
        hy[k] = qy[i];
        ++k;
    }
    --k;
    if (k < 3) return 0.0;
    double a = 0.0;
    for (int i = 0; i < k; ++i) {
        const int nx = i + 1 == k ? 0 : i + 1;
        a += hx[i] * hy[nx] - hx[nx] * hy[i];
    }
    return 0.5 * fabs(a);
}

// Sutherland-Hodgman clip of quad A by quad B (IOU.py:9-56, strict '>' inside test); returns the area
// of the clipped polygon (the reference takes the Qhull hull volume of it, IOU.py:64-74).
__device__ __forceinline__ double clip_area(const IouBox &A, const IouBox &B)
{
    double px[10], py[10], qx[10], qy[10];
    int n = 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        px[i] = A.rx[i];
        py[i] = A.rz[i];
    }
    double c1x = B.rx[3], c1y = B.rz[3];
    for (int ci = 0; ci < 4; ++ci) {
        const double c2x = B.rx[ci], c2y = B.rz[ci];
        int m = 0;
        double sx = px[n - 1], sy = py[n - 1];
        for (int v = 0; v < n; ++v) {
            const double ex = px[v], ey = py[v];
            const bool e_in = (c2x - c1x) * (ey - c1y) > (c2y - c1y) * (ex - c1x);
            const bool s_in = (c2x - c1x) * (sy - c1y) > (c2y - c1y) * (sx - c1x);
            if (e_in != s_in) {
                const double dcx = c1x - c2x, dcy = c1y - c2y;
                const double dpx = sx - ex, dpy = sy - ey;
                const double n1 = c1x * c2y - c1y * c2x;
                const double n2 = sx * ey - sy * ex;
                const double n3 = 1.0 / (dcx * dpy - dcy * dpx);
                qx[m] = (n1 * dpx - n2 * dcx) * n3;
                qy[m] = (n1 * dpy - n2 * dcy) * n3;
                ++m;
            }
            if (e_in) {
                qx[m] = ex;
                qy[m] = ey;
                ++m;
            }
            sx = ex;
            sy = ey;
        }
        c1x = c2x;
        c1y = c2y;
        if (m == 0) return 0.0;
        n = m;
        for (int v = 0; v < n; ++v) {
            px[v] = qx[v];
            py[v] = qy[v];
        }
    }
    return hull_area(px, py, n);
}

__device__ __forceinline__ void iou_pair(const IouBox &A, const IouBox &B, double &iou3d, double &iou2d)
{
    const double inter = clip_area(A, B);
    iou2d = inter / (A.area + B.area - inter);
    const double ymax = fmin(A.ytop, B.ytop), ymin = fmax(A.ybot, B.ybot);
    const double inter_vol = inter * fmax(0.0, ymax - ymin);
    iou3d = inter_vol / (A.vol + B.vol - inter_vol);
}

}  // namespace cf
