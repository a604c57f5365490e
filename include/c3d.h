/*
 * c3d.h — C ABI of libc3d.so, the B200-native (sm_100a) kernels behind the Cube R-CNN hot path.
 *
 * The reference (facebookresearch/omni3d) is pure Python and has no FFI of its own; every entry
 * point below replaces the third-party native op the reference reaches at the cited call site.
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes; every pointer is DEVICE memory unless the name ends in _host;
 *   - the caller owns every buffer including the workspace (size from the matching
 *     *_workspace_bytes query); kernels never allocate, free or synchronise;
 *   - work is enqueued on `stream` (a cudaStream_t / CUstream handle passed as void*);
 *   - return 0 (C3D_OK) or a negative c3d_status; c3d_last_error() gives a thread-local string.
 */
#ifndef C3D_H_
#define C3D_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  C3D_OK = 0,
  C3D_EINVAL = -1,     /* bad argument (null pointer, negative size, misalignment) */
  C3D_EWORKSPACE = -2, /* workspace too small */
  C3D_ECUDA = -3       /* CUDA launch/runtime error, see c3d_last_error() */
} c3d_status;

const char* c3d_last_error(void);
/* library/ABI version, bumped when a signature changes */
int32_t c3d_abi_version(void);

/* ------------------------------------------------------------------------------------------
 * Oriented-box 3D IoU.
 * Replaces pytorch3d._C.iou_box3d as called at cubercnn/evaluation/omni3d_evaluation.py:155 and
 * the wrapper cubercnn/evaluation/omni3d_evaluation.py:106-166 (box3d_overlap) with its row
 * checks :65-104.  Boxes are (n,8,3) fp32 contiguous, corner order of DATA.md:109-131.
 * ------------------------------------------------------------------------------------------ */

/* workspace for n1 x n2 (cross) or n1 pairs (paired: pass n2 = 0) */
size_t c3d_iou_box3d_workspace_bytes(int64_t n1, int64_t n2);

/* == pytorch3d._C.iou_box3d(boxes1, boxes2) -> (vol, iou), both (n1, n2) fp32 row-major.
 * vol and nfaces may be NULL.  nfaces (int32) = number of triangles of the intersection
 * polyhedron per pair (debug/parity quantity; -1 if the pair exceeded every capacity). */
int32_t c3d_iou_box3d(const float* boxes1, int64_t n1, const float* boxes2, int64_t n2,
                      float* vol, float* iou, int32_t* nfaces,
                      void* workspace, size_t workspace_bytes, void* stream);

/* paired variant: pair k = (boxes1[k], boxes2[k]), outputs (n,) */
int32_t c3d_iou_box3d_paired(const float* boxes1, const float* boxes2, int64_t n,
                             float* vol, float* iou, int32_t* nfaces,
                             void* workspace, size_t workspace_bytes, void* stream);

/* == box3d_overlap(boxes_dt, boxes_gt, eps_coplanar, eps_nonzero) -> iou (n_dt, n_gt) with rows of
 * non-coplanar / zero-area dt boxes zeroed (the reference prints a warning instead of raising,
 * omni3d_evaluation.py:158-164).  n_bad (device int32[2], may be NULL) receives the number of
 * non-coplanar and zero-area dt boxes so the host mirror can print the same warnings. */
int32_t c3d_box3d_overlap(const float* boxes_dt, int64_t n_dt, const float* boxes_gt, int64_t n_gt,
                          float eps_coplanar, float eps_nonzero, float* iou, int32_t* n_bad,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * NHWC bf16 implicit-GEMM convolution on tcgen05 tensor cores (TMA-staged, fp32 accumulate in TMEM).
 * Replaces the cuDNN calls behind nn.Conv2d in cubercnn/modeling/backbone/dla.py:43-51,159-161,
 * 211-214,241-243,287-297, the detectron2 FPN convs built at dla.py:500-506 / resnet.py:88-95 and
 * the StandardRPNHead convs (configs/Base.yaml:49).  Data-gradient = the same entry point with
 * flipped/transposed weights; weight-gradient = c3d_conv2d_wgrad.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t N, H, W, Cin;      /* input  (N,H,W,Cin)  bf16, pixel stride x_pix_stride elements (0 => Cin) */
  int32_t Cout, KH, KW;      /* weight (Cout,KH,KW,Cin) bf16 contiguous */
  int32_t stride, pad;       /* stride 1 or 2 (same in h and w), symmetric zero padding */
  int32_t relu;              /* epilogue: max(.,0) after bias/addend */
  int32_t out_fp32;          /* output fp32 instead of bf16 */
  int32_t add_mode;          /* 0 none; 1 addend (N,Ho,Wo,Cout); 2 addend (N,Ho/2,Wo/2,Cout) nearest-up x2 (FPN) */
  int64_t x_pix_stride, y_pix_stride, add_pix_stride;   /* elements; 0 => dense */
} c3d_conv_desc;

/* number of 128-pixel output tiles (= rows of the BatchNorm partial-statistics buffer) and tile shape */
int32_t c3d_conv2d_tiles(const c3d_conv_desc* d, int32_t* tiles_m, int32_t* tile_h, int32_t* tile_w);

/* y = conv(x, w) [+ bias] [+ addend] [relu].  stats (may be NULL): fp32 [tiles_m][2][Cout] per-tile
 * partial (sum, sum of squares) of the raw fp32 conv output, for train-mode BatchNorm. */
int32_t c3d_conv2d_fwd(const c3d_conv_desc* d, const void* x, const void* w, const float* bias,
                       const void* addend, void* y, float* stats, void* stream);

/* dw[Cout][KH][KW][Cin] (fp32) += the weight gradient of the convolution described by d, from the
 * forward input x (N,H,W,Cin) and the output gradient dy (N,Ho,Wo,Cout), both bf16 NHWC.
 * Split-K over pixels with fp32 atomics: the caller zeroes (or pre-loads) dw. */
int32_t c3d_conv2d_wgrad(const c3d_conv_desc* d, const void* x, const void* dy, float* dw, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* C3D_H_ */
