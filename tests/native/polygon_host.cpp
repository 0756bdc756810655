// Host build of the exact-overlap core the GPU kernel uses (groomed_nms_b200/csrc/polygon.cuh), for the CPU tests:
//   g++ -O2 -ffp-contract=off -shared -fPIC -o polygon_host.so polygon_host.cpp
// Test infrastructure: lets `-m "not gpu"` check the clipping logic against the oracle without a GPU, and the GPU test
// check that the kernel and this build agree bit for bit.  Not part of the product library.
#include "../../groomed_nms_b200/csrc/polygon.cuh"

extern "C" void polygon_host_iou3d(const double* corners_a, long ld_a, int M, const double* corners_b, long ld_b, int N,
                                   const double* vol, int list_mode, double* out_bev, double* out_3d) {
    const long total = list_mode ? M : (long)M * N;
    for (long p = 0; p < total; ++p) {
        const long i = list_mode ? p : p / N, j = list_mode ? p : p % N;
        gnms::ExactBox b1, b2;
        gnms::exact_box_from_corners(corners_a + i * ld_a, b1);
        gnms::exact_box_from_corners(corners_b + j * ld_b, b2);
        gnms::iou3d_exact_pair(b1, b2, vol != nullptr, vol ? vol[p] : 0.0, out_bev[p], out_3d[p]);
    }
}
