/*
 * groomed_nms_b200 -- C-ABI of the B200-native GrooMeD-NMS hot path (libgroomed_b200.so).
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes (no torch types), returns an int
 * (0 = success, >0 = cudaError_t, <0 = GNMS_E_* argument error), never throws, never prints, keeps no mutable
 * process-wide state and is re-entrant (tuning knobs travel per call in a gnms_launch_opts, never in globals).  Unless stated otherwise all pointers are DEVICE pointers, outputs and workspaces
 * are caller-allocated, and work is enqueued on `stream` (a cudaStream_t passed as void*) without synchronising.
 *
 * Reference interfaces replaced (paths relative to abhi1kumar/groomed_nms @ ad10dbb):
 *   lib/core.py:178-243   intersect            -> gnms_overlap2d_f32 / gnms_overlap2d_list_f32 (kind = INTERSECT)
 *   lib/core.py:480-532   iou                  -> gnms_overlap2d_f32 / gnms_overlap2d_list_f32 (kind = IOU)
 *                         (float64 / mixed-dtype inputs of either: gnms_overlap2d_f64; autograd of iou: gnms_iou2d_backward_f32)
 *   lib/math_3d.py:364-435 get_corners_of_cuboid -> gnms_corners_from_boxes7_f32 (+ gnms_corners_backward_f32 for its autograd)
 *   lib/core.py:354-388,434-477 get_volume / remove_rotation_in_boxes / per-box min-max -> gnms_box3d_records_f32
 *   lib/core.py:305-421   iou3d_approximate    -> gnms_overlap3d_f32 / gnms_overlap3d_list_f32 (+ gnms_iou3d_approx_backward_f32)
 *   lib/groomed_nms.py:10-129 differentiable_nms (+ :167-189 pruning_function, :208-270 get_groups,
 *                         :272-336 indices_copy folded away) -> gnms_forward_f32 / gnms_backward_f32 and the
 *                         matrix-free gnms_forward_boxes_f32
 *   lib/nms/gpu_nms.hpp:1-2 `_nms` (the reference's only native ABI), lib/nms/nms_kernel.cu:24-144,
 *   lib/nms/cpu_nms.pyx:17-68, lib/nms/py_cpu_nms.py:10-38, lib/nms_others.py:119-150 girshick_nms
 *                                              -> gnms_hard_nms_f32 and the host-pointer `gnms_nms_host`
 *   lib/nms_others.py:6-116 navneeth_soft_nms  -> gnms_soft_nms_f64
 *   lib/loss/aploss.py:14-97 backpropAPLoss    -> gnms_aploss_f32
 *   lib/rpn_util.py:439-461 compute_targets overlaps (lib/core.py iou / iou_ign, numpy branch) -> gnms_targets_overlaps_f64
 *   lib/core.py:246-302   iou3d (exact polygon overlap, shapely in the reference) -> gnms_iou3d_exact_f64
 *   lib/groomed_nms.py:131-165 soft_sort        -> gnms_soft_sort_forward_f32 / gnms_soft_sort_backward_f32
 *   lib/loss/rpn_3d.py:722-768,801-825 (top-500 selection, best box per ground truth) -> gnms_masked_topk_f32 / gnms_best_box_per_gt_f32
 *   lib/core.py:68        nn.DataParallel gradient reduction of the acceptance head (models/...alpha.py:112-121,230)
 *                                              -> gnms_score_head_*_f32, gnms_score_head_backward_allreduce_f32 (+ gnms_peer_buffer_*)
 */
#ifndef GROOMED_NMS_B200_H_
#define GROOMED_NMS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GNMS_VERSION 100

/* argument errors (negative); positive return codes are cudaError_t values */
#define GNMS_E_BADARG   (-1)
#define GNMS_E_TOOLARGE (-2)   /* N above GNMS_MAX_BOXES */
#define GNMS_E_ALIGN    (-3)
#define GNMS_E_UNSUPPORTED (-4)

#define GNMS_MAX_BOXES 8192    /* per image; the greedy chain runs in one CTA's shared memory */

/* pruning_function, lib/groomed_nms.py:167-189 */
#define GNMS_PRUNE_LINEAR    0
#define GNMS_PRUNE_SIGMOIDAL 1
#define GNMS_PRUNE_SOFT_NMS  2

/* inversion mode, lib/groomed_nms.py:95-110 */
#define GNMS_MODE_GROUP_MASK   0   /* group_boxes=True,  mask_group_boxes=True  (shipped default) */
#define GNMS_MODE_GROUP_NOMASK 1   /* group_boxes=True,  mask_group_boxes=False: per-group (I+Phi_g)^-1 */
#define GNMS_MODE_NOGROUP      2   /* group_boxes=False: full (I+Phi)^-1 */
/* Mode GROUP_MASK with the lower-triangular cut of Phi (torch.tril, lib/groomed_nms.py:72) taken by INPUT position instead of by
 * score rank: a member that comes before its leader in the input keeps its score.  This is what the reference computes on its
 * sorting_method="soft" path (:42-45), where the scores it groups by (re-sorted inside get_groups, :213) need not be in the
 * order of the rows of the soft-sorted matrix.  `prob` is still returned in score-rank order. */
#define GNMS_MODE_GROUP_MASK_INPUT_TRIL 3

/* 2D overlap kinds */
#define GNMS_KIND_IOU       0      /* lib/core.py:480 */
#define GNMS_KIND_INTERSECT 1      /* lib/core.py:178 */

/* box sources of the matrix-free forward */
#define GNMS_BOX_2D       0        /* float[N,4] x1,y1,x2,y2 ; overlap = iou (lib/core.py:480) */
#define GNMS_BOX_3D_REC   1        /* float[N,8] records from gnms_box3d_records_f32 ; overlap = iou3d_approximate */

/* comparison used to mark a lower-scored box as suppressed by a kept one */
#define GNMS_CMP_GT   0            /* iou >  thr   lib/nms/nms_kernel.cu:71 (NaN is not suppressed) */
#define GNMS_CMP_GE   1            /* iou >= thr   lib/nms/cpu_nms.pyx:65 */
#define GNMS_CMP_NLE  2            /* !(iou <= thr) lib/nms/py_cpu_nms.py:35, lib/nms_others.py:146 (NaN is suppressed) */

typedef struct gnms_params {
    float nms_threshold;           /* group / prune threshold                  (lib/groomed_nms.py:10) */
    float temperature;             /* pruning temperature                       */
    float valid_box_prob_threshold;
    int32_t pruning_method;        /* GNMS_PRUNE_*                              */
    int32_t mode;                  /* GNMS_MODE_*                               */
    int32_t group_size;            /* groups keep group_size+1 boxes (:254-255) */
    int32_t thresholded_output;    /* 1: `prob` is the thresholded vector (group_boxes=False, :127) */
    int32_t sorted_output;         /* 1: return_sorted_prob=True (:116-119): prob is sorted descending */
} gnms_params;

int gnms_version(void);
const char* gnms_error_string(int rc);

/* ---------------------------------------------------------------- per-call launch options ---------- */
/* Optional tuning / profiling knobs of the `_ex` entry points.  Pass NULL for the library defaults (what the plain entry
 * points do).  Versioned by struct_size: set it to sizeof(gnms_launch_opts); a library built against a longer struct reads
 * only the fields the caller's size covers.  Every field's 0 means "library default".  Nothing here changes results: all
 * variants are bit-identical, only the schedule differs. */
#define GNMS_MATRIX_KERNEL_AUTO   0
#define GNMS_MATRIX_KERNEL_DIRECT 1   /* 256x64 symmetric tiles stored straight from registers (16-byte STG, direct + mirrored) */
#define GNMS_MATRIX_KERNEL_TMA    2   /* same tiles staged through swizzled shared memory, drained by 2-D TMA tensor stores */
#define GNMS_RANK_AUTO   0
#define GNMS_RANK_COUNT  1            /* rank by counting (small batches) */
#define GNMS_RANK_SORT   2            /* per-image radix sort in shared memory */
#define GNMS_ELECT_AUTO   0
#define GNMS_ELECT_DIRECT 1           /* leaders elected directly from the boxes, one per step (round-1 kernel, kept for comparison) */
#define GNMS_ELECT_MASK   2           /* always through the all-pairs suppression bitmask */
#define GNMS_ELECT_BATCHED 3          /* leaders elected from the boxes, up to 32 per step (what AUTO picks where applicable) */
/* stage bits of gnms_launch_opts.stage_mask (profiling: run only some kernels of a forward) */
#define GNMS_STAGE_RANK     1u
#define GNMS_STAGE_SPATIAL  2u
#define GNMS_STAGE_TILES    4u        /* tile kernel / matrix -> mask stream / matrix-only kernel */
#define GNMS_STAGE_EARLIER  8u
#define GNMS_STAGE_CHAIN   16u        /* chain kernel and triangular solves */
#define GNMS_STAGE_ELECT   32u
#define GNMS_OPT_SCALAR_MATH   1u     /* flags: scalar fp32 instead of packed fp32x2 arithmetic in the matrix-only kernel */
#define GNMS_OPT_INLINE_HITS   2u     /* flags: threshold hits handled inline instead of through the CTA-drained queue */
#define GNMS_OPT_ONE_PASS      4u     /* flags: matrix + suppression bits from one all-pairs pass (no separate matrix-only kernel) */
#define GNMS_OPT_SPLIT_CHAIN   8u     /* flags: batched election and chain as two launches (default: the chain runs at the end of the election kernel) */
typedef struct gnms_launch_opts {
    uint32_t struct_size;          /* sizeof(gnms_launch_opts) */
    int32_t  matrix_kernel;        /* GNMS_MATRIX_KERNEL_* : how the [N,N] overlap matrix is written */
    int32_t  tiles_per_cta;        /* matrix-only kernel: 0 = persistent CTAs (one wave), k > 0 = each CTA takes k 64x64-column
                                      tile units and retires (lets concurrent small kernels of another stream find SM slots) */
    int32_t  rank_method;          /* GNMS_RANK_* */
    int32_t  election;             /* GNMS_ELECT_* */
    uint32_t stage_mask;           /* 0 = all stages, else an OR of GNMS_STAGE_* */
    uint32_t flags;                /* OR of GNMS_OPT_* */
} gnms_launch_opts;

/* ---------------------------------------------------------------- pairwise overlaps ---------------- */
/* out[i*ld_out + j] = overlap(a_i, b_j), i<M, j<N ("combinations"); a,b are [.,4] row-major boxes.
 * kind=IOU gives the LOGICAL [M,N] result of lib/core.py:480 iou (the reference returns it as a transposed
 * view); kind=INTERSECT gives lib/core.py:178 intersect TRANSPOSED to [M,N] as well (the reference returns [N,M]).
 * fp32, separately rounded ops in the reference's order (bitwise equal to torch CPU). */
int gnms_overlap2d_f32(const float* a, int M, const float* b, int N, float* out, int64_t ld_out, int kind,
                       void* stream);
/* "list" mode: out[i] = overlap(a_i, b_i). */
int gnms_overlap2d_list_f32(const float* a, const float* b, int M, float* out, int kind, void* stream);

/* The same in fp64 for float64 inputs (the reference's functions keep the dtype of their inputs; its inference path compares
 * float64 IoUs with the NMS threshold, lib/rpn_util.py:1295): a[M,*] / b[N,*] with row strides ld_a / ld_b >= 4 doubles;
 * out is [M,N] row-major (combinations: out[i,j] = overlap(a_i, b_j)) or [M] (list_mode != 0, M == N).
 * area_f32: bit 0 / bit 1 set = side a / b was float32 in the caller's hands (mixed calls such as lib/rpn_util.py:448): its box
 * areas are then rounded to float32 as numpy / torch type promotion does, everything else is float64. */
int gnms_overlap2d_f64(const double* a, int64_t ld_a, int M, const double* b, int64_t ld_b, int N, int kind, int list_mode,
                       int area_f32, double* out, void* stream);

/* Backward of the IoU above (the reference's iou is a differentiable torch composite, lib/core.py:480-532, and
 * the detection loss differentiates its list mode, lib/loss/rpn_3d.py:620).  g is [M,N] (combinations, ld = N)
 * or [M] (list_mode!=0, then N must equal M); grad_a[M,4] / grad_b[N,4] are fully written.  min/max ties split the
 * gradient evenly and clamp(.,0) passes it on the closed side, as torch does. */
int gnms_iou2d_backward_f32(const float* a, int M, const float* b, int N, const float* g, int list_mode,
                            float* grad_a, float* grad_b, void* stream);

/* boxes7[N,7] = (x3d,y3d,z3d,w3d,h3d,l3d,ry3d) row-major with row stride `ld` floats -> corners[N,3,8]
 * (lib/math_3d.py:364-435, iou_3d_convention=True). */
int gnms_corners_from_boxes7_f32(const float* boxes7, int64_t ld, int N, float* corners, void* stream);
/* Same with the reference's `iou_3d_convention` argument: != 0 is the order above (lib/math_3d.py:379-403), 0 the other vertex
 * order of the reference (:405-426: length on x for corners {1,2,3,4}, height on y for {2,3,6,7}, width on z for {3,4,5,6}). */
int gnms_corners_from_boxes7_ex_f32(const float* boxes7, int64_t ld, int N, int iou_3d_convention, float* corners, void* stream);

/* Backward of the corners above (the reference's composite of cos / sin / bmm is differentiable and the acceptance-probability
 * target differentiates it, lib/loss/rpn_3d.py:663-679): grad_corners[N,3,8] -> grad_boxes7[N,7] (x,y,z,w,h,l,ry), fully written. */
int gnms_corners_backward_f32(const float* boxes7, int64_t ld, int N, int iou_3d_convention, const float* grad_corners,
                              float* grad_boxes7, void* stream);

/* lib/math_3d.py:47-72 project_3d_points_in_4D_format: out[4,n] = p2[4,4] @ [pts;1] (pts is [3,n] when
 * pad_ones!=0, else [4,n]); rows 0,1 divided by row 2 where |row 2| > 1e-2. */
int gnms_project_points_f32(const float* p2, const float* pts, int64_t n, int pad_ones, float* out, void* stream);

/* corners[N,3,8] -> rec[N,8] = (ymin,ymax,bx1,bx2,bz1,bz2,vol,area_bev) exactly as iou3d_approximate derives
 * them (lib/core.py:354-388).  mutate_input!=0 reproduces the reference's in-place Y<-Z write (:379-380). */
int gnms_box3d_records_f32(float* corners, int N, float* rec, int mutate_input, void* stream);
/* Both steps in one launch: boxes7[N,7] (row stride ld) -> rec[N,8], and corners[N,3,8] too if not NULL.  Same bits as
 * gnms_corners_from_boxes7_f32 followed by gnms_box3d_records_f32 (mutate_input = 0). */
int gnms_box3d_records_from_boxes7_f32(const float* boxes7, int64_t ld, int N, float* rec, float* corners, void* stream);

/* rec_a[M,8], rec_b[N,8] -> out_bev / out_3d [M,N] (either may be NULL).  generalized!=0: GIoU hull term
 * (lib/core.py:390-419).  affine!=0 additionally maps out_3d to 0.5*(1+x) (lib/loss/rpn_3d.py:781).
 * mul2d (NULL or [M,N] with ld_out): out_3d *= mul2d  (overlap_in_nms="product", lib/loss/rpn_3d.py:786). */
int gnms_overlap3d_f32(const float* rec_a, int M, const float* rec_b, int N, float* out_bev, float* out_3d,
                       int64_t ld_out, int generalized, int affine, const float* mul2d, void* stream);
/* Batched self-overlaps (one launch): boxes[B,N,4] -> out[B,N,N] IoU; rec[B,N,8] -> out_3d[B,N,N].  Symmetric 64x64
 * tiles, every unordered pair evaluated once, the tile and its mirror image stored straight from registers. */
int gnms_overlap2d_batched_f32(const float* boxes, int N, int batch, float* out, void* stream);
int gnms_overlap3d_batched_f32(const float* rec, int N, int batch, float* out_3d, int generalized, int affine,
                               void* stream);
int gnms_overlap3d_list_f32(const float* rec_a, const float* rec_b, int M, float* out_bev, float* out_3d,
                            int generalized, int affine, void* stream);

/* Backward of iou3d_approximate (lib/core.py:305-421) wrt both corner sets, from the UNMUTATED corners the forward read:
 * corners_a[M,3,8], corners_b[N,3,8]; g_bev / g_3d are dL/diou_bev and dL/diou_3d, [M,N] row-major (combinations) or [M]
 * (list_mode != 0, M == N), either may be NULL; grad_corners_a[M,3,8] / grad_corners_b[N,3,8] are fully written.
 * Sub-gradients follow torch: clamp(x, 0) passes on x >= 0, binary min / max and max(0, x) split ties evenly, the min / max
 * over the corners of one box send the gradient to the first extremal corner. */
int gnms_iou3d_approx_backward_f32(const float* corners_a, int M, const float* corners_b, int N, int list_mode, int generalized,
                                   const float* g_bev, const float* g_3d, float* grad_corners_a, float* grad_corners_b,
                                   void* stream);

/* The batched self-overlaps with per-call launch options (matrix_kernel, tiles_per_cta, flags). */
int gnms_overlap2d_batched_ex_f32(const float* boxes, int N, int batch, float* out, const gnms_launch_opts* opts, void* stream);
int gnms_overlap3d_batched_ex_f32(const float* rec, int N, int batch, float* out_3d, int generalized, int affine,
                                  const gnms_launch_opts* opts, void* stream);

/* ---------------------------------------------------------------- GrooMeD-NMS ---------------------- */
/* Bytes of scratch for a batch of `batch` images of up to N boxes each. */
size_t gnms_workspace_bytes(int N, int batch);

/* State saved by the forward for the backward (all device, caller-allocated, [batch,N] each):
 *   order  int32  sorted position -> input index            sorted_scores f32
 *   lead   int32  sorted position of the group leader (self for leaders, -1 if in no group)
 *   pval   f32    p(iou[m, leader])   dpval f32  p'(iou[m, leader])   pre f32  pre-clamp rescored value */
typedef struct gnms_saved {
    int32_t* order;
    float* sorted_scores;
    int32_t* lead;
    float* pval;
    float* dpval;
    float* pre;
} gnms_saved;

/* differentiable_nms forward from an overlap MATRIX (lib/groomed_nms.py:10-129, sorting_method="hard").
 *   scores[batch,N]; iou[batch,N,ld] row-major, unit column stride (row i = box i as the lower-scored box);
 *   n_per_image: NULL (all N) or device int32[batch] with the live box count of each image (<= N).
 * outputs: prob[batch,N] (score-sorted order unless sorted_output), valid_idx/invalid_idx int64[batch,N]
 * (input index space; valid by rescored prob descending, ties by sorted position; invalid by sorted position),
 * counts int32[batch,2] = (n_valid, n_invalid). */
int gnms_forward_f32(const float* scores, const float* iou, int64_t ld, int N, int batch,
                     const int32_t* n_per_image, const gnms_params* p, float* prob, int64_t* valid_idx,
                     int64_t* invalid_idx, int32_t* counts, gnms_saved saved, void* workspace, void* stream);

/* Forward from BOXES (the fused north-star path): no overlap matrix is read.  boxes: box_kind GNMS_BOX_2D
 * float[batch,N,4] or GNMS_BOX_3D_REC float[batch,N,8] records; overlap3d flags as in gnms_overlap3d_f32 (generalized,
 * affine).  Same outputs as gnms_forward_f32, bit for bit what it gives on the matrix of the same boxes.
 *   mode GROUP_MASK, N <= 4096: the leaders are elected directly from the boxes (the reference's greedy loop,
 *     lib/groomed_nms.py:247-262, one CTA per image, only the leaders' overlap columns are ever evaluated); an image
 *     that needs more than 384 leaders falls back, inside the same call, to
 *   the general route: every unordered pair of boxes is evaluated once, register-resident, by the symmetric tile
 *     kernel (spatially culled when no matrix is wanted), which emits the suppression bits the greedy stage consumes.
 * overlap_out (optional, [batch,N,N]): the API-visible overlap matrix is also streamed to HBM (written once, never
 * read back; bitwise equal to gnms_overlap2d_f32 / gnms_overlap3d_f32). */
int gnms_forward_boxes_f32(const float* scores, const float* boxes, int box_kind, int generalized, int affine,
                           int N, int batch, const int32_t* n_per_image, const gnms_params* p, float* overlap_out,
                           float* prob, int64_t* valid_idx, int64_t* invalid_idx, int32_t* counts, gnms_saved saved,
                           void* workspace, void* stream);

/* The two forwards with per-call launch options (NULL = defaults = the plain entry points). */
int gnms_forward_ex_f32(const float* scores, const float* iou, int64_t ld, int N, int batch,
                        const int32_t* n_per_image, const gnms_params* p, float* prob, int64_t* valid_idx,
                        int64_t* invalid_idx, int32_t* counts, gnms_saved saved, void* workspace,
                        const gnms_launch_opts* opts, void* stream);
int gnms_forward_boxes_ex_f32(const float* scores, const float* boxes, int box_kind, int generalized, int affine,
                              int N, int batch, const int32_t* n_per_image, const gnms_params* p, float* overlap_out,
                              float* prob, int64_t* valid_idx, int64_t* invalid_idx, int32_t* counts, gnms_saved saved,
                              void* workspace, const gnms_launch_opts* opts, void* stream);

/* Analytic backward.  grad_prob[batch,N] is dL/dprob for the prob the forward returned.
 * grad_scores[batch,N] (input order) is fully written.  grad_iou: NULL, or [batch,N,ld_gi] which must be
 * zero-filled by the caller for mode GROUP_MASK (only the <=N (member,leader) entries are written); modes
 * GROUP_NOMASK / NOGROUP write their lower-triangular blocks.  iou/ld are needed for modes NOMASK/NOGROUP only. */
int gnms_backward_f32(const float* grad_prob, const float* prob, const float* iou, int64_t ld, int N, int batch,
                      const int32_t* n_per_image, const gnms_params* p, gnms_saved saved, float* grad_scores,
                      float* grad_iou, int64_t ld_gi, void* workspace, void* stream);

/* Keep lists in transfer form: out[b, k] = (int32) valid_idx[b, k] for k < min(counts[b, 0], K), -1 beyond -- K << N entries
 * per image instead of the [batch, N] int64 array (the reference's `valid_boxes_index`, lib/groomed_nms.py:122, is that short
 * list; counts[b, 0] tells the caller whether K was enough). */
int gnms_pack_keep_i32(const int64_t* valid_idx, const int32_t* counts, int N, int batch, int K, int32_t* out, void* stream);

/* get_groups (lib/groomed_nms.py:208-270) from a matrix: group_id[N] int32 = index (in leader order) of the
 * group of INPUT box i or -1, group_rank[N] int32 = position inside the group (0 = leader), n_groups int32[1]. */
int gnms_get_groups_f32(const float* scores, const float* iou, int64_t ld, int N, float group_threshold,
                        int group_size, int32_t* group_id, int32_t* group_rank, int32_t* n_groups,
                        void* workspace, void* stream);

/* pruning_function (lib/groomed_nms.py:167-189), elementwise: out[i] = p(x[i]). */
int gnms_prune_f32(const float* x, int64_t n, int pruning_method, float nms_threshold, float temperature,
                   float* out, void* stream);

/* indices_copy (lib/groomed_nms.py:272-336): A[ra[k], ca[k], :] = B[rb[k], cb[k], :] for k < npairs, with
 * A [rowsA, colsA, C], B [rowsB, colsB, C] row-major.  If ca == NULL the pairs are the cartesian product of the
 * index list ra (length npairs) with itself and B is read densely: A[ra[i], ra[j], :] = B[i, j, :]. */
int gnms_indices_copy_f32(float* A, int64_t colsA, const float* B, int64_t colsB, int64_t C, const int64_t* ra,
                          const int64_t* ca, const int64_t* rb, const int64_t* cb, int64_t npairs, void* stream);

/* soft_sort (lib/groomed_nms.py:131-165; SoftSort, Prillo & Eisenschlos 2020), sorting_method="soft".
 * forward: s[N], optional M[N,N] (row stride ld_m) -> soft_scores[N] = P s, P[N,N] with P[i][j] = exp(-|s_j - h_i| / T) / S[j]
 * (h = s sorted descending, S[i] = sum_j exp(-|s_j - h_i| / T) + 1e-3; the division is column-wise by the ROW sums, as the
 * reference's broadcast at :155 does), soft_matrix[N,N] = P M (rows only, :164).  Saved for the backward: perm int32[N] (sorted
 * position -> input index), hsorted[N], S[N].  The N x N x N product is an fp32 FMA-pipe GEMM (1e-5 parity with the
 * reference's fp32 matmul rules out TF32 / BF16 tensor-core inputs).
 * backward: upstream g_scores[N], g_P[N,N], g_SM[N,N] (each optional) -> grad_s[N], grad_M[N,N] (optional); dP_scratch[N,N]
 * is caller-allocated scratch, workspace gnms_soft_sort_workspace_bytes(N). */
size_t gnms_soft_sort_workspace_bytes(int N);
int gnms_soft_sort_forward_f32(const float* s, int N, float temperature, const float* M, int64_t ld_m, float* soft_scores, float* P,
                               float* soft_matrix, int32_t* perm, float* hsorted, float* S, void* stream);
int gnms_soft_sort_backward_f32(const float* s, int N, float temperature, const float* M, int64_t ld_m, const float* P,
                                const int32_t* perm, const float* hsorted, const float* S, const float* g_scores, const float* g_P,
                                const float* g_SM, float* grad_s, float* grad_M, float* dP_scratch, void* workspace, void* stream);

/* ---------------------------------------------------------------- classical NMS -------------------- */
/* dets[N,5] = (x1,y1,x2,y2,score) device.  shift = the "+1" pixel convention (1.0 for lib/nms, any for
 * girshick_nms).  cmp = GNMS_CMP_*.  keep int32[N] = kept ORIGINAL indices in descending score order,
 * n_keep int32[1].  Replaces lib/nms/nms_kernel.cu:34-144 (device-resident, no malloc, no host reduce). */
int gnms_hard_nms_f32(const float* dets, int N, float thresh, float shift, int cmp, int32_t* keep,
                      int32_t* n_keep, void* workspace, void* stream);

/* Drop-in for the reference's only native symbol, `_nms` (lib/nms/gpu_nms.hpp:1-2): HOST pointers, boxes
 * already sorted by descending score, keep_out indexes into that sorted order, synchronous.  Unlike the
 * reference it returns an error code instead of printing. */
int gnms_nms_host(int* keep_out, int* num_out, const float* boxes_host, int boxes_num, int boxes_dim,
                  float nms_overlap_thresh, int device_id);

/* Soft-NMS (lib/nms_others.py:6-116), float64 like the reference's numpy arithmetic on a float64 array.
 * dets[N,5] device f64 (not modified).  keep int32[N] = the reference's `keep_orig[:N_final]` (original indices in
 * selection order), keep_scores f64[N] = their decayed scores (the reference leaves them in boxes[:N_final,4]),
 * n_keep int32[1].  method: 0 hard, 1 linear, 2 gaussian.  scratch: device f64[N] + int32[N] workspace
 * (gnms_soft_nms_workspace_bytes). */
int gnms_soft_nms_f64(const double* dets, int N, double sigma, double Nt, double threshold, int method,
                      double shift, int32_t* keep, double* keep_scores, int32_t* n_keep, void* workspace,
                      void* stream);
size_t gnms_soft_nms_workspace_bytes(int N);

/* ---------------------------------------------------------------- AP loss -------------------------- */
/* lib/loss/aploss.py:16-81: logits[n], targets[n] in {1,0,-1} -> loss[1], grad[n] (= ctx.grad). */
int gnms_aploss_f32(const float* logits, const float* targets, int n, float* loss, float* grad,
                    void* workspace, size_t workspace_bytes, void* stream);
size_t gnms_aploss_workspace_bytes(int n);

/* ---------------------------------------------------------------- loss-branch front / back end, batched */
/* lib/loss/rpn_3d.py:731-737, all images of a step in one launch: the (at most) K <= 1024 highest-scoring foreground anchors
 * of every image.  scores[batch, A] fp32, mask[batch, A] uint8 (non-zero = foreground).  out_idx int64[batch, K]: anchor ids
 * by descending score, ties by lower anchor id (slots past out_n[b] hold 0); out_n int32[batch] = min(#foreground, K).
 * Replaces a full sort of all A (= 69 120) scores per image and the host round trip of :739-744. */
int gnms_masked_topk_f32(const float* scores, const uint8_t* mask, int A, int batch, int K, int64_t* out_idx, int32_t* out_n,
                         void* stream);
/* lib/loss/rpn_3d.py:813-825, all ground truths of a step in one launch: for ground truth g of image gt_image[g], the
 * candidate i < n_per_image[b] maximising (0.5 * (1 + giou3d(rec[b,i], gt_rec[g]))) * iou2d(box2d[b,i], gt_2d[g]) (first maximum,
 * NaN counts as the maximum, like torch.max); if that score is > beta, targets[b, top[b, i]] = 1.  rec[batch,K,8] records and
 * box2d[batch,K,4] of the candidates, gt_rec[n_gt,8], gt_2d[n_gt,4], top int64[batch,K] = anchor id of candidate slot,
 * targets fp32[batch, A] (caller-zeroed).  best_slot int32[n_gt] / best_score fp32[n_gt]: optional. */
int gnms_best_box_per_gt_f32(const float* rec, const float* box2d, const int32_t* n_per_image, int K, const float* gt_rec,
                             const float* gt_2d, const int32_t* gt_image, int n_gt, float beta, const int64_t* top, int A,
                             float* targets, int32_t* best_slot, float* best_score, void* stream);

/* ---------------------------------------------------------------- score head (batched-image configuration) */
/* The head that produces the scores entering GrooMeD-NMS in the batched configuration (BASELINE.json configs[3]): the
 * reference's acceptance-probability head is a 1x1 convolution -- a per-anchor linear map of the proposal features --
 * followed by a sigmoid (models/densenet121_3d_dilate_decomp_alpha.py:112-121,230).  x[M,K] row-major features (K = 64),
 * wb[K+1] = weights then bias.  forward: scores[m] = sigmoid(x[m,:].w + b).  backward: grad_wb[K+1] = dL/d(w, b) given
 * dL/dscores (deterministic two-stage sum; this is the flat gradient bucket the step all-reduces over NCCL). */
size_t gnms_score_head_workspace_bytes(int K);
int gnms_score_head_forward_f32(const float* x, int64_t M, int K, const float* wb, float* scores, void* stream);
int gnms_score_head_backward_f32(const float* x, int64_t M, int K, const float* scores, const float* grad_scores,
                                 float* grad_wb, void* workspace, void* stream);

/* The same with the data-parallel all-reduce of the gradient inside the second launch (replaces the reference's nn.DataParallel
 * gradient reduction, lib/core.py:68, for this head): every rank owns a small exchange buffer that all peers map over NVLink
 * (CUDA IPC; gnms_peer_buffer_create on the owner, the 64-byte handle travels through any host channel, gnms_peer_buffer_open
 * on the peers); the kernel reduces this rank's block partials into its buffer, releases a flag into every peer's buffer, waits
 * for the flags of this step and sums the W ranks' 65 values in rank order.  grad_wb[K + 1] holds the ALL-REDUCED gradient on
 * return, bitwise the same on every rank.  All ranks must call it once per step; status (device int32, may be NULL) is set to 1
 * -- and grad_wb to NaN -- if a peer did not show up within ~2 s. */
typedef struct gnms_peers {
    void* buf[16];                 /* exchange buffer of every rank (buf[rank] is this rank's own), device pointers */
    int32_t rank, world;
} gnms_peers;
int gnms_peer_buffer_create(void** dev_ptr, unsigned char handle_out[64]);
int gnms_peer_buffer_open(const unsigned char handle[64], void** dev_ptr);
int gnms_peer_buffer_close(void* dev_ptr, int owned);
int gnms_score_head_backward_allreduce_f32(const float* x, int64_t M, int K, const float* scores, const float* grad_scores,
                                           float* grad_wb, void* workspace, const gnms_peers* peers, int32_t* status, void* stream);

/* ---------------------------------------------------------------- target-assignment overlaps (next to the path) */
/* lib/rpn_util.py:439-461 compute_targets: ols = iou(rois, gts) (kind GNMS_KIND_IOU, lib/core.py:480-513 numpy branch)
 * or iou_ign(rois, gts) (kind 1, lib/core.py:535-575), float64 like numpy's promotion at the call site (rois float32,
 * ground truths float64; area_f32 != 0 keeps the rois' own area in float32 arithmetic as numpy does there).
 * rois[M, ld_rois >= 4] and gts[G <= 64, 4] are device float64.  Outputs: ols[M,G] (optional), row_max[M] / row_arg[M]
 * (np.amax / np.argmax over axis 1: first maximum, NaN is the maximum), col_max[G] / col_arg[G] (axis 0). */
size_t gnms_targets_overlaps_workspace_bytes(int M, int G);
int gnms_targets_overlaps_f64(const double* rois, int64_t ld_rois, int M, const double* gts, int G, int kind, int area_f32,
                              double* ols, double* row_max, int64_t* row_arg, double* col_max, int64_t* col_arg,
                              void* workspace, void* stream);

/* lib/core.py:246-302 iou3d: exact overlap of rotated cuboids -- the bottom faces (corners 7, 2, 3, 6 in the x-z plane)
 * are intersected as polygons (the reference asks shapely/GEOS, one pair per host call; here the two convex quadrilaterals
 * are clipped in fp64, one thread per pair), the height overlap comes from the y range of the 8 corners.
 * corners_*: double[M][rows >= 3][8] (x row, y row, z row; ld_* = doubles between boxes, >= 24).  vol: NULL (the sum of
 * the two get_volume values, :278-279) or one volume sum per pair.  list_mode: pair i with i (M == N), else all M x N
 * combinations, a-major.  out_bev / out_3d: double[M*N] (or [M]); either may be NULL.  Zero-area boxes give NaN (0/0)
 * where the reference's Python floats raise ZeroDivisionError. */
int gnms_iou3d_exact_f64(const double* corners_a, int64_t ld_a, int M, const double* corners_b, int64_t ld_b, int N,
                         const double* vol, int list_mode, double* out_bev, double* out_3d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GROOMED_NMS_B200_H_ */
