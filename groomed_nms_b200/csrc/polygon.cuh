// Exact bird's-eye-view overlap of two rotated cuboids (reference lib/core.py:246-302, `iou3d`): the reference builds two
// shapely polygons from the bottom-face corners [7, 2, 3, 6] in the (x, z) plane and asks GEOS for the area of their
// intersection, one pair per call, on the host.  Here the intersection of the two convex quadrilaterals is clipped
// directly (Sutherland-Hodgman, fp64) and its area taken with the shoelace formula.
//
// The arithmetic is plain + - * / fmin fmax in a fixed order, and the functions are host+device: the GPU kernel
// (misc.cu) and the host shim the CPU tests compile (tests/native/polygon_host.cpp, -ffp-contract=off) run the same
// source and agree bit for bit (the library is built with -fmad=false).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define GNMS_HD __host__ __device__ __forceinline__
#else
#define GNMS_HD inline
#endif

namespace gnms {

struct ExactBox {
    double x[4], z[4];     // bottom face in the (x, z) plane, corner order 7, 2, 3, 6 (lib/core.py:290-292)
    double ymin, ymax;     // :281-284
    double vol;            // get_volume (:434-458): prod(max - min) over x, y, z
    double area;           // |shoelace| / 2 of the face
};

// twice the signed area
GNMS_HD double shoelace2(const double* x, const double* z, int n) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1 == n) ? 0 : i + 1;
        s += x[i] * z[j] - x[j] * z[i];
    }
    return s;
}

// corners: rows x (then) y (then) z, 8 values each (the reference's (3, 8) layout, :276-277 transposes it)
GNMS_HD void exact_box_from_corners(const double* c, ExactBox& b) {
    const int order[4] = {7, 2, 3, 6};
    for (int i = 0; i < 4; ++i) { b.x[i] = c[order[i]]; b.z[i] = c[16 + order[i]]; }
    double lo[3], hi[3];
    for (int r = 0; r < 3; ++r) {
        lo[r] = c[r * 8]; hi[r] = c[r * 8];
        for (int k = 1; k < 8; ++k) { lo[r] = fmin(lo[r], c[r * 8 + k]); hi[r] = fmax(hi[r], c[r * 8 + k]); }
    }
    b.ymin = lo[1]; b.ymax = hi[1];
    b.vol = ((hi[0] - lo[0]) * (hi[1] - lo[1])) * (hi[2] - lo[2]);
    b.area = fabs(shoelace2(b.x, b.z, 4)) * 0.5;
}

// Area of (convex quadrilateral a) n (convex quadrilateral b); either orientation.  Every clip adds at most one vertex:
// 4 -> at most 8.  A vertex exactly on a clip edge counts as inside (identical boxes give the full area).
GNMS_HD double convex_quad_intersection_area(const double* ax, const double* az, const double* bx, const double* bz) {
    const double sb = shoelace2(bx, bz, 4);
    if (!(sb != 0.0)) return 0.0;                       // degenerate clip polygon (or NaN)
    const double sgn = sb > 0.0 ? 1.0 : -1.0;
    double px[8], pz[8], qx[8], qz[8];
    int n = 4;
    for (int i = 0; i < 4; ++i) { px[i] = ax[i]; pz[i] = az[i]; }
    for (int e = 0; e < 4 && n > 0; ++e) {
        const double ex = bx[e], ez = bz[e];
        const double dx = bx[(e + 1) & 3] - ex, dz = bz[(e + 1) & 3] - ez;
        int m = 0;
        double prx = px[n - 1], prz = pz[n - 1];
        double sp = sgn * (dx * (prz - ez) - dz * (prx - ex));
        for (int i = 0; i < n; ++i) {
            const double cx = px[i], cz = pz[i];
            const double sc = sgn * (dx * (cz - ez) - dz * (cx - ex));
            const bool in_c = sc >= 0.0, in_p = sp >= 0.0;
            if (in_c != in_p) {                          // the segment crosses the clip line
                const double t = sp / (sp - sc);
                qx[m] = prx + t * (cx - prx);
                qz[m] = prz + t * (cz - prz);
                ++m;
            }
            if (in_c) { qx[m] = cx; qz[m] = cz; ++m; }
            prx = cx; prz = cz; sp = sc;
        }
        n = m;
        for (int i = 0; i < n; ++i) { px[i] = qx[i]; pz[i] = qz[i]; }
    }
    if (n < 3) return 0.0;
    return fabs(shoelace2(px, pz, n)) * 0.5;
}

// (iou_bev, iou_3d) of box b1 against box b2 (lib/core.py:286-302); `vol` = sum of the two volumes unless given (:278-279)
GNMS_HD void iou3d_exact_pair(const ExactBox& b1, const ExactBox& b2, bool has_vol, double vol, double& iou_bev, double& iou_3d) {
    const double y_int = fmax(0.0, fmin(b1.ymax, b2.ymax) - fmax(b1.ymin, b2.ymin));        // :285
    const double inter = convex_quad_intersection_area(b2.x, b2.z, b1.x, b1.z);             // :295
    const double i3d = y_int * inter;                                                        // :296
    if (!has_vol) vol = b1.vol + b2.vol;
    iou_bev = inter / ((b2.area + b1.area) - inter);                                         // :298
    iou_3d = i3d / (vol - i3d);                                                              // :299
}

}  // namespace gnms
