// Analytic backward of the two differentiable feeders of the 3D overlap, so that the mirrored functions keep the autograd the
// reference's torch composites have (lib/loss/rpn_3d.py:663-679 differentiates through both when the acceptance-probability
// target is not detached):
//   get_corners_of_cuboid   lib/math_3d.py:364-435   dL/d(x, y, z, w, h, l, ry) from dL/dcorners[N,3,8]
//   iou3d_approximate       lib/core.py:305-421      dL/dcorners_1, dL/dcorners_2 from dL/diou_bev and dL/diou_3d
// Sub-gradient conventions are torch's: clamp(x, 0) passes the gradient where x >= 0 (lib/core.py:210-218 intersect);
// binary min / max and max(0, x) (y overlap :376, hulls :430) split it evenly on ties; min / max over the corners of a box
// (index-returning reductions) send it to one corner -- the first extremal one here (any choice gives the same gradient wrt the
// 7 box parameters: tied corners of a y-rotated cuboid depend on them identically).
#include "common.cuh"

namespace gnms {

// ---------------------------------------------------------------------------------------------- corners backward
__global__ void __launch_bounds__(128) corners_bwd_kernel(const float* __restrict__ boxes7, int64_t ld, int N, int kitti_order,
                                                          const float* __restrict__ g, float* __restrict__ gb) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* p = boxes7 + (int64_t)n * ld;
    const float w = p[3], h = p[4], l = p[5], ry = p[6];
    const float cs = cosf(ry), sn = sinf(ry);
    const unsigned xmask = kitti_order ? 0x1Eu : 0x6Au, ymask = 0xCCu, zmask = kitti_order ? 0x78u : 0xF0u;
    const float* gx = g + (size_t)n * 24;
    float sx = 0.f, sy = 0.f, sz = 0.f, gl = 0.f, gh = 0.f, gw = 0.f, gr = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float a = gx[k], b = gx[8 + k], c = gx[16 + k];
        const float hx = ((xmask >> k) & 1) ? 0.5f : -0.5f, hy = ((ymask >> k) & 1) ? 0.5f : -0.5f, hz = ((zmask >> k) & 1) ? 0.5f : -0.5f;
        const float px = hx * l, pz = hz * w;
        sx += a; sy += b; sz += c;
        gl += hx * (cs * a - sn * c);
        gh += hy * b;
        gw += hz * (sn * a + cs * c);
        gr += a * (-sn * px + cs * pz) + c * (-cs * px - sn * pz);
    }
    float* o = gb + (size_t)n * 7;
    o[0] = sx; o[1] = sy; o[2] = sz; o[3] = gw; o[4] = gh; o[5] = gl; o[6] = gr;
}

// ---------------------------------------------------------------------------------------------- iou3d_approximate backward
struct Ext {                      // the 10 extents iou3d_approximate reads from a box, with the corner each one comes from
    float x8lo, x8hi, ylo, yhi, z8lo, z8hi, bx1, bx2, bz1, bz2;
    int ix8lo, ix8hi, iylo, iyhi, iz8lo, iz8hi, ibx1, ibx2, ibz1, ibz2;
};
__device__ __forceinline__ void argmin8(const float* v, float& m, int& im) { m = v[0]; im = 0; for (int k = 1; k < 8; ++k) if (v[k] < m) { m = v[k]; im = k; } }
__device__ __forceinline__ void argmax8(const float* v, float& m, int& im) { m = v[0]; im = 0; for (int k = 1; k < 8; ++k) if (v[k] > m) { m = v[k]; im = k; } }
__device__ __forceinline__ void argmin_bev(const float* v, float& m, int& im) {
    const int id[4] = {2, 3, 6, 7};
    m = v[2]; im = 2;
    for (int q = 1; q < 4; ++q) if (v[id[q]] < m) { m = v[id[q]]; im = id[q]; }
}
__device__ __forceinline__ void argmax_bev(const float* v, float& m, int& im) {
    const int id[4] = {2, 3, 6, 7};
    m = v[2]; im = 2;
    for (int q = 1; q < 4; ++q) if (v[id[q]] > m) { m = v[id[q]]; im = id[q]; }
}
__device__ __forceinline__ Ext extents_of(const float* c) {               // c = [3][8]
    Ext e;
    argmin8(c, e.x8lo, e.ix8lo); argmax8(c, e.x8hi, e.ix8hi);
    argmin8(c + 8, e.ylo, e.iylo); argmax8(c + 8, e.yhi, e.iyhi);
    argmin8(c + 16, e.z8lo, e.iz8lo); argmax8(c + 16, e.z8hi, e.iz8hi);
    argmin_bev(c, e.bx1, e.ibx1); argmax_bev(c, e.bx2, e.ibx2);
    argmin_bev(c + 16, e.bz1, e.ibz1); argmax_bev(c + 16, e.bz2, e.ibz2);
    return e;
}
struct ExtGrad { float x8lo, x8hi, ylo, yhi, z8lo, z8hi, bx1, bx2, bz1, bz2; };
__device__ __forceinline__ void zero(ExtGrad& g) { g.x8lo = g.x8hi = g.ylo = g.yhi = g.z8lo = g.z8hi = g.bx1 = g.bx2 = g.bz1 = g.bz2 = 0.f; }

__device__ __forceinline__ void split_min(float a, float b, float g, float& ga, float& gb) {     // d min(a,b)
    if (a < b) ga += g; else if (b < a) gb += g; else { ga += 0.5f * g; gb += 0.5f * g; }
}
__device__ __forceinline__ void split_max(float a, float b, float g, float& ga, float& gb) {
    if (a > b) ga += g; else if (b > a) gb += g; else { ga += 0.5f * g; gb += 0.5f * g; }
}
__device__ __forceinline__ float relu_grad_binary(float x) { return x > 0.f ? 1.f : (x == 0.f ? 0.5f : 0.f); }     // torch.max(zeros, x)
__device__ __forceinline__ float relu_grad_clamp(float x) { return x >= 0.f ? 1.f : 0.f; }                           // torch.clamp(x, 0)

// one pair: accumulates into ga (box a) and gb (box b).  list_tie: in "list" mode the y overlap uses an index-returning max over
// the concatenated pair (:374-375) -- ties go to the first argument there instead of being split
template <bool kGen>
__device__ __forceinline__ void pair_backward(const Ext& a, const Ext& b, float g_bev, float g_3d, ExtGrad& ga, ExtGrad& gb) {
    // forward
    const float mnx = fminf(a.bx2, b.bx2), mxx = fmaxf(a.bx1, b.bx1), dw = mnx - mxx, iw = fmaxf(dw, 0.f);
    const float mnz = fminf(a.bz2, b.bz2), mxz = fmaxf(a.bz1, b.bz1), dh = mnz - mxz, ih = fmaxf(dh, 0.f);
    const float ibev = iw * ih;
    const float Aa = (a.bx2 - a.bx1) * (a.bz2 - a.bz1), Ab = (b.bx2 - b.bx1) * (b.bz2 - b.bz1);
    const float U2 = Aa + Ab - ibev;
    const float ymn = fminf(a.yhi, b.yhi), ymx = fmaxf(a.ylo, b.ylo), dy = ymn - ymx, yint = fmaxf(dy, 0.f);
    const float va = (a.x8hi - a.x8lo) * (a.yhi - a.ylo) * (a.z8hi - a.z8lo), vb = (b.x8hi - b.x8lo) * (b.yhi - b.ylo) * (b.z8hi - b.z8lo);
    const float V = va + vb, i3 = ibev * yint, un = V - i3;
    // backward
    float g_un = -g_3d * i3 / (un * un), g_i3 = g_3d / un;
    float g_xh = 0.f, g_yh = 0.f, g_zh = 0.f, hx = 0.f, hy = 0.f, hz = 0.f;
    if (kGen) {
        hx = fmaxf(a.bx2, b.bx2) - fminf(a.bx1, b.bx1); hy = fmaxf(a.yhi, b.yhi) - fminf(a.ylo, b.ylo); hz = fmaxf(a.bz2, b.bz2) - fminf(a.bz1, b.bz1);
        const float xh = fmaxf(hx, 0.f), yh = fmaxf(hy, 0.f), zh = fmaxf(hz, 0.f), vh = xh * yh * zh;
        const float g_vh = -g_3d * un / (vh * vh);
        g_un += g_3d / vh;
        g_xh = g_vh * yh * zh; g_yh = g_vh * xh * zh; g_zh = g_vh * xh * yh;
    }
    const float g_V = g_un;
    g_i3 -= g_un;
    float g_ibev = g_i3 * yint + g_bev / U2;
    const float g_yint = g_i3 * ibev;
    const float g_U2 = -g_bev * ibev / (U2 * U2);
    g_ibev -= g_U2;
    const float g_dw = g_ibev * ih * relu_grad_clamp(dw), g_dh = g_ibev * iw * relu_grad_clamp(dh);
    split_min(a.bx2, b.bx2, g_dw, ga.bx2, gb.bx2); split_max(a.bx1, b.bx1, -g_dw, ga.bx1, gb.bx1);
    split_min(a.bz2, b.bz2, g_dh, ga.bz2, gb.bz2); split_max(a.bz1, b.bz1, -g_dh, ga.bz1, gb.bz1);
    // areas
    ga.bx2 += g_U2 * (a.bz2 - a.bz1); ga.bx1 -= g_U2 * (a.bz2 - a.bz1); ga.bz2 += g_U2 * (a.bx2 - a.bx1); ga.bz1 -= g_U2 * (a.bx2 - a.bx1);
    gb.bx2 += g_U2 * (b.bz2 - b.bz1); gb.bx1 -= g_U2 * (b.bz2 - b.bz1); gb.bz2 += g_U2 * (b.bx2 - b.bx1); gb.bz1 -= g_U2 * (b.bx2 - b.bx1);
    // y overlap
    const float g_dy = g_yint * relu_grad_binary(dy);
    split_min(a.yhi, b.yhi, g_dy, ga.yhi, gb.yhi); split_max(a.ylo, b.ylo, -g_dy, ga.ylo, gb.ylo);
    // volumes
    {
        const float dxa = a.x8hi - a.x8lo, dya = a.yhi - a.ylo, dza = a.z8hi - a.z8lo;
        ga.x8hi += g_V * dya * dza; ga.x8lo -= g_V * dya * dza; ga.yhi += g_V * dxa * dza; ga.ylo -= g_V * dxa * dza; ga.z8hi += g_V * dxa * dya; ga.z8lo -= g_V * dxa * dya;
        const float dxb = b.x8hi - b.x8lo, dyb = b.yhi - b.ylo, dzb = b.z8hi - b.z8lo;
        gb.x8hi += g_V * dyb * dzb; gb.x8lo -= g_V * dyb * dzb; gb.yhi += g_V * dxb * dzb; gb.ylo -= g_V * dxb * dzb; gb.z8hi += g_V * dxb * dyb; gb.z8lo -= g_V * dxb * dyb;
    }
    if (kGen) {
        const float gx = g_xh * relu_grad_binary(hx), gy = g_yh * relu_grad_binary(hy), gz = g_zh * relu_grad_binary(hz);
        split_max(a.bx2, b.bx2, gx, ga.bx2, gb.bx2); split_min(a.bx1, b.bx1, -gx, ga.bx1, gb.bx1);
        split_max(a.yhi, b.yhi, gy, ga.yhi, gb.yhi); split_min(a.ylo, b.ylo, -gy, ga.ylo, gb.ylo);
        split_max(a.bz2, b.bz2, gz, ga.bz2, gb.bz2); split_min(a.bz1, b.bz1, -gz, ga.bz1, gb.bz1);
    }
}

__device__ __forceinline__ void scatter_ext_grad(const Ext& e, const ExtGrad& g, float* out /* [3][8], zero-filled */) {
    out[e.ix8lo] += g.x8lo; out[e.ix8hi] += g.x8hi;
    out[8 + e.iylo] += g.ylo; out[8 + e.iyhi] += g.yhi;
    out[16 + e.iz8lo] += g.z8lo; out[16 + e.iz8hi] += g.z8hi;
    out[e.ibx1] += g.bx1; out[e.ibx2] += g.bx2;
    out[16 + e.ibz1] += g.bz1; out[16 + e.ibz2] += g.bz2;
}
__device__ __forceinline__ float warp_sum(float v) { for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; }
__device__ __forceinline__ void warp_sum(ExtGrad& g) {
    g.x8lo = warp_sum(g.x8lo); g.x8hi = warp_sum(g.x8hi); g.ylo = warp_sum(g.ylo); g.yhi = warp_sum(g.yhi); g.z8lo = warp_sum(g.z8lo);
    g.z8hi = warp_sum(g.z8hi); g.bx1 = warp_sum(g.bx1); g.bx2 = warp_sum(g.bx2); g.bz1 = warp_sum(g.bz1); g.bz2 = warp_sum(g.bz2);
}

// list mode: one thread per pair, both sides
template <bool kGen>
__global__ void __launch_bounds__(128) iou3d_bwd_list_kernel(const float* __restrict__ ca, const float* __restrict__ cb, int M,
                                                             const float* __restrict__ g_bev, const float* __restrict__ g_3d,
                                                             float* __restrict__ gca, float* __restrict__ gcb) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    float a[24], b[24];
    for (int k = 0; k < 24; ++k) { a[k] = ca[(size_t)m * 24 + k]; b[k] = cb[(size_t)m * 24 + k]; }
    const Ext ea = extents_of(a), eb = extents_of(b);
    ExtGrad ga, gb;
    zero(ga); zero(gb);
    pair_backward<kGen>(ea, eb, g_bev ? g_bev[m] : 0.f, g_3d ? g_3d[m] : 0.f, ga, gb);
    float oa[24], ob[24];
    for (int k = 0; k < 24; ++k) { oa[k] = 0.f; ob[k] = 0.f; }
    scatter_ext_grad(ea, ga, oa); scatter_ext_grad(eb, gb, ob);
    for (int k = 0; k < 24; ++k) { gca[(size_t)m * 24 + k] = oa[k]; gcb[(size_t)m * 24 + k] = ob[k]; }
}
// combinations: one warp per box of one side, lanes over the boxes of the other side; kSideB: this warp's box is a column box
template <bool kGen, bool kSideB>
__global__ void __launch_bounds__(256) iou3d_bwd_comb_kernel(const float* __restrict__ ca, int M, const float* __restrict__ cb, int N,
                                                             const float* __restrict__ g_bev, const float* __restrict__ g_3d,
                                                             float* __restrict__ gout) {
    const int me = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int n_me = kSideB ? N : M, n_other = kSideB ? M : N;
    if (me >= n_me) return;
    float mine[24];
    const float* src = (kSideB ? cb : ca) + (size_t)me * 24;
    for (int k = 0; k < 24; ++k) mine[k] = src[k];
    const Ext em = extents_of(mine);
    ExtGrad gm;
    zero(gm);
    for (int o = lane; o < n_other; o += 32) {
        float oth[24];
        const float* so = (kSideB ? ca : cb) + (size_t)o * 24;
        for (int k = 0; k < 24; ++k) oth[k] = so[k];
        const Ext eo = extents_of(oth);
        const size_t idx = kSideB ? (size_t)o * N + me : (size_t)me * N + o;          // outputs are [M, N]
        ExtGrad dump;
        zero(dump);
        if (kSideB) pair_backward<kGen>(eo, em, g_bev ? g_bev[idx] : 0.f, g_3d ? g_3d[idx] : 0.f, dump, gm);
        else pair_backward<kGen>(em, eo, g_bev ? g_bev[idx] : 0.f, g_3d ? g_3d[idx] : 0.f, gm, dump);
    }
    warp_sum(gm);
    if (lane == 0) {
        float out[24];
        for (int k = 0; k < 24; ++k) out[k] = 0.f;
        scatter_ext_grad(em, gm, out);
        for (int k = 0; k < 24; ++k) gout[(size_t)me * 24 + k] = out[k];
    }
}

}  // namespace gnms

using namespace gnms;

extern "C" int gnms_corners_backward_f32(const float* boxes7, int64_t ld, int N, int iou_3d_convention, const float* grad_corners,
                                         float* grad_boxes7, void* stream) {
    if (N < 0 || ld < 7) return GNMS_E_BADARG;
    if (N == 0) return 0;
    if (!boxes7 || !grad_corners || !grad_boxes7) return GNMS_E_BADARG;
    corners_bwd_kernel<<<gnms_div_up(N, 128), 128, 0, (cudaStream_t)stream>>>(boxes7, ld, N, iou_3d_convention ? 0 : 1, grad_corners, grad_boxes7);
    GNMS_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnms_iou3d_approx_backward_f32(const float* corners_a, int M, const float* corners_b, int N, int list_mode, int generalized,
                                              const float* g_bev, const float* g_3d, float* grad_corners_a, float* grad_corners_b,
                                              void* stream) {
    if (M < 0 || N < 0 || (list_mode && M != N)) return GNMS_E_BADARG;
    if (M == 0 || N == 0) return 0;
    if (!corners_a || !corners_b || !grad_corners_a || !grad_corners_b || (!g_bev && !g_3d)) return GNMS_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (list_mode) {
        if (generalized) iou3d_bwd_list_kernel<true><<<gnms_div_up(M, 128), 128, 0, s>>>(corners_a, corners_b, M, g_bev, g_3d, grad_corners_a, grad_corners_b);
        else iou3d_bwd_list_kernel<false><<<gnms_div_up(M, 128), 128, 0, s>>>(corners_a, corners_b, M, g_bev, g_3d, grad_corners_a, grad_corners_b);
    } else if (generalized) {
        iou3d_bwd_comb_kernel<true, false><<<gnms_div_up(M, 8), 256, 0, s>>>(corners_a, M, corners_b, N, g_bev, g_3d, grad_corners_a);
        iou3d_bwd_comb_kernel<true, true><<<gnms_div_up(N, 8), 256, 0, s>>>(corners_a, M, corners_b, N, g_bev, g_3d, grad_corners_b);
    } else {
        iou3d_bwd_comb_kernel<false, false><<<gnms_div_up(M, 8), 256, 0, s>>>(corners_a, M, corners_b, N, g_bev, g_3d, grad_corners_a);
        iou3d_bwd_comb_kernel<false, true><<<gnms_div_up(N, 8), 256, 0, s>>>(corners_a, M, corners_b, N, g_bev, g_3d, grad_corners_b);
    }
    GNMS_LAUNCH_CHECK();
    return 0;
}
