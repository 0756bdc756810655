/* Plain C99 consumer of the drop-in boundary: includes include/groomed_nms_b200.h, links libgroomed_b200.so and walks the
 * argument-validation paths that need no GPU.  Built and run by tests/test_abi_cpu.py:
 *   gcc -std=c99 -Wall -Wextra -Werror -pedantic -Iinclude examples/c_abi_check.c -Lgroomed_nms_b200 -lgroomed_b200 \
 *       -Wl,-rpath,$PWD/groomed_nms_b200 -o c_abi_check && ./c_abi_check
 * A maintainer binding the library from C / Cython / cgo sees exactly these signatures. */
#include <stdio.h>
#include <string.h>

#include "groomed_nms_b200.h"

#define EXPECT(cond) do { if (!(cond)) { fprintf(stderr, "FAILED line %d: %s\n", __LINE__, #cond); return 1; } } while (0)

int main(void) {
    gnms_params p;
    gnms_saved sv;
    memset(&p, 0, sizeof p);
    memset(&sv, 0, sizeof sv);
    EXPECT(gnms_version() == GNMS_VERSION);
    EXPECT(strcmp(gnms_error_string(0), "success") == 0);
    EXPECT(gnms_workspace_bytes(4096, 1) >= (size_t)4096 * 4096 / 8);
    EXPECT(gnms_workspace_bytes(4096, 2) > gnms_workspace_bytes(4096, 1));
    /* empty problems are no-ops, bad sizes are refused before anything touches the device */
    EXPECT(gnms_overlap2d_f32(NULL, 0, NULL, 4, NULL, 4, GNMS_KIND_IOU, NULL) == 0);
    EXPECT(gnms_overlap2d_f32(NULL, -1, NULL, 4, NULL, 4, GNMS_KIND_IOU, NULL) == GNMS_E_BADARG);
    EXPECT(gnms_forward_f32(NULL, NULL, GNMS_MAX_BOXES + 1, GNMS_MAX_BOXES + 1, 1, NULL, &p, NULL, NULL, NULL, NULL, sv,
                            NULL, NULL) == GNMS_E_TOOLARGE);
    EXPECT(gnms_forward_boxes_f32(NULL, NULL, GNMS_BOX_3D_REC, 1, 1, 0, 1, NULL, &p, NULL, NULL, NULL, NULL, NULL, sv, NULL,
                                  NULL) == 0);
    EXPECT(gnms_iou3d_exact_f64(NULL, 24, 0, NULL, 24, 5, NULL, 0, NULL, NULL, NULL) == 0);
    EXPECT(gnms_iou3d_exact_f64(NULL, 16, 3, NULL, 24, 5, NULL, 0, NULL, NULL, NULL) == GNMS_E_BADARG);
    EXPECT(gnms_iou3d_exact_f64(NULL, 24, 3, NULL, 24, 5, NULL, 1, NULL, NULL, NULL) == GNMS_E_BADARG);
    printf("c abi ok: version %d, workspace(4096, 1) = %lu bytes\n", gnms_version(), (unsigned long)gnms_workspace_bytes(4096, 1));
    return 0;
}
