/*
 * cellvit_b200.h -- C ABI of libcellvit_b200.so: the B200-native CellViT tile-inference hot path
 * (network forward + HoVer-Net post-processing).
 *
 * The reference (TIO-IKIM/CellViT) is pure Python with no FFI layer; the boundary it exposes for this path is the
 * Python API cited next to each entry point below. A binding only needs ctypes: every
 * argument is a plain pointer or integer, there are no torch types, no exceptions cross the boundary.
 *
 * Conventions
 *   - return value: 0 = OK, <0 = error class (below); cvb_last_error() gives the thread-local message.
 *   - all data pointers are DEVICE pointers unless the name says host; buffers are caller-allocated.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*) and performs no allocation and no
 *     synchronisation, so sequences of calls can be captured into a CUDA graph.
 *   - one cvb_model per device; a handle is not thread-safe, different handles are independent.
 */
#ifndef CELLVIT_B200_H
#define CELLVIT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVB_OK 0
#define CVB_EARG -1        /* bad argument (null pointer, missing parameter, ...)                           */
#define CVB_ESHAPE -2      /* unsupported shape (e.g. H or W not divisible by 16: cellvit.py:170-175)       */
#define CVB_ECUDA -3       /* CUDA runtime / driver error                                                   */
#define CVB_EWORKSPACE -4  /* workspace too small                                                           */
#define CVB_EOVERFLOW -5   /* instance table overflow: counts[] holds the number of rows that were needed  */

int cvb_version(void);        /* 200 = this header (100: round 1, before cvb_model_desc.shared_decoder / the arg-max and export entries) */
const char* cvb_last_error(void);

/* ------------------------------------------------------------------------------------------------ forward
 * Replaces  CellViT.forward / CellViT256.forward / CellViTSAM.forward
 *   models/segmentation/cell_segmentation/cellvit.py:153-210, :586-644  (+ encoder wrappers utils.py:149-233)
 */
typedef struct cvb_model cvb_model;

typedef struct {
    int sam;            /* 1: SAM ViTDet encoder (windowed + global attention, rel-pos); 0: ViT-S/16 with cls token */
    int embed_dim;      /* 1280 / 1024 / 768 (SAM-H/L/B), 384 (ViT-256)                   cellvit.py:660-665   */
    int depth;
    int num_heads;
    int window_size;    /* 14 for SAM (cellvit.py:563), ignored when sam == 0                                 */
    int n_global;       /* number of entries used in global_idx                                               */
    int global_idx[8];  /* blocks with global attention (cellvit.py:664)                                      */
    int extract[4];     /* 1-based block indices after which skips z1..z4 are taken (cellvit.py:665)          */
    int n_np_out;       /* 2, or 4 with regression_loss (cellvit.py:134-141)                                  */
    int n_nt;           /* num_nuclei_classes                                                                 */
    int n_tissue;       /* num_tissue_classes                                                                 */
    int skip11, skip12; /* decoder widths (cellvit.py:106-113)                                                */
    int bott_pad;       /* bottleneck width padded to a multiple of 64 (312 -> 320 for ViT-256)               */
    int shared_decoder; /* 1: the *Shared variants (cellvit_shared.py:147-231): ONE upsampling trunk (parameters "dec.*")
                           and a 1x1 head per output on its 64-channel feature map; 0: three branches            */
} cvb_model_desc;

int cvb_model_create(const cvb_model_desc* desc, cvb_model** out);
void cvb_model_destroy(cvb_model* m);
/* Per-handle options (no process-global state). "attention_tc": bit 0 = tcgen05 kernel for the global-attention blocks,
 * bit 1 = tcgen05 kernel for the 14 x 14 windows (default 3 = both); 0 = the mma.sync kernels (parity-test reference).
 * "dynamic_tiles": 1 (default) = persistent kernels claim tiles from a per-launch counter, 0 = static tile lists.
 * "square_canvas": 0 (default) = a non-native tile size extends each decoder dimension to its own canvas, 1 = to a square one. */
int cvb_model_set_option(cvb_model* m, const char* name, int value);

/* Registers one packed parameter tensor (device pointer; the caller keeps it alive). Names and layouts are
 * listed in DESIGN.md ("packed parameter table"); cellvit_b200/packing.py produces them from a reference
 * state_dict (what nn.Module.load_state_dict consumes, cell_detection.py:131-138). */
int cvb_model_set_param(cvb_model* m, const char* name, const void* dev_ptr);

int cvb_model_workspace_bytes(cvb_model* m, int B, int H, int W, size_t* out);

/* x [B,3,H,W] fp32 (already normalised, cell_detection.py:214-227). Outputs are raw logits, fp32 NCHW:
 * np_logits [B,n_np_out,H,W], hv [B,2,H,W], nt_logits [B,n_nt,H,W], tissue [B,n_tissue],
 * tokens [B,embed_dim,H/16,W/16] (nullable; the z4 skip, retrieve_tokens=True).
 * Shapes: any H, W divisible by 16 up to 1024 (cellvit.py:170-175, 603-608; CVB_ESHAPE otherwise), non-square included for the
 * ViT-S encoder; the SAM encoders take square tiles only, as in the reference (its pos_embed slice, utils.py:222-224, only
 * broadcasts for square token grids). Token grids of 16 / 32 / 64 per edge (256 / 512 / 1024 pixels) are the tile engine's
 * native decoder tilings; any other size runs the encoder on its real token grid and the decoder on the next larger
 * zero-extended canvas, cropped on output -- same results, the cost of the canvas size. */
int cvb_forward(cvb_model* m, const float* x, int B, int H, int W, float* np_logits, float* hv, float* nt_logits,
                float* tissue, float* tokens, void* workspace, size_t ws_bytes, void* stream);
/* The same forward that ALSO writes the arg-max planes the post-processing consumes (K12 fusion): np_argmax /
 * nt_argmax uint8 [B,H,W] (nullable) = torch.argmax over the two binary-map channels / the n_nt type channels, first maximum
 * wins (cellvit.py:369-375 takes the arg-max of the soft-maxed maps: same plane). Written by the fused head epilogue at no
 * extra pass; feed them to cvb_postproc_argmax. */
int cvb_forward_argmax(cvb_model* m, const float* x, int B, int H, int W, float* np_logits, float* hv, float* nt_logits,
                       float* tissue, float* tokens, uint8_t* np_argmax, uint8_t* nt_argmax, void* workspace, size_t ws_bytes,
                       void* stream);

/* ------------------------------------------------------------------------------------------------ post-processing
 * Replaces  DetectionCellPostProcessor.post_process_cell_segmentation  (cell_segmentation/utils/post_proc_cellvit.py:67-249)
 * as driven per tile by  CellViT.calculate_instance_map  (cellvit.py:332-383).
 */
typedef struct {
    int32_t id;                      /* instance id = marker label (non-contiguous)                            */
    int32_t rmin, cmin, rmax, cmax;  /* bbox, max exclusive (tools.py:24-34)                                   */
    int32_t area;
    int32_t type;                    /* majority vote with the "0 yields to runner-up" rule (:141-147)         */
    float type_prob_f;
    double cx, cy;                   /* centroid x,y (:117-125)                                                */
    double type_prob;                /* count / (area + 1e-6) (:149)                                           */
    int32_t hist[8];                 /* per-class pixel counts                                                 */
} cvb_inst_row;

int cvb_postproc_workspace_bytes(int B, int H, int W, size_t* out);

/* np_map [B,2,H,W] and nt_map [B,n_types,H,W] fp32 (probabilities or logits -- only the argmax is used,
 * cellvit.py:369-375; first maximum wins like torch.argmax), hv [B,2,H,W] fp32.
 * magnification 40 or 20 (post_proc_cellvit.py:55-62). Outputs: labels int32 [B,H,W]; table [B,max_rows];
 * counts int32 [B]. nt_map may be null (nr_types=None). */
int cvb_postproc(const float* np_map, const float* hv, const float* nt_map, int B, int H, int W, int n_types,
                 int magnification, int32_t* labels, cvb_inst_row* table, int32_t* counts, int max_rows,
                 void* workspace, size_t ws_bytes, void* stream);

/* cvb_postproc on arg-max planes (uint8 [B,H,W], e.g. from cvb_forward_argmax): np_argmax != 0 = nucleus, nt_argmax = class
 * (nullable). The planes are read in place, no preparation pass: 14 B/px of compulsory input instead of 44 B/px. */
int cvb_postproc_argmax(const uint8_t* np_argmax, const float* hv, const uint8_t* nt_argmax, int B, int H, int W, int n_types,
                        int magnification, int32_t* labels, cvb_inst_row* table, int32_t* counts, int max_rows, void* workspace,
                        size_t ws_bytes, void* stream);

/* Same, from already arg-maxed maps: np_bin uint8 [B,H,W] (0/1), type_map int32 [B,H,W] (nullable). Optional
 * debug outputs (nullable) expose the stage results the parity tests compare with the oracle:
 * blb uint8, dist float64, marker int32, all [B,H,W]. object_size / ksize are the resolved magnification
 * parameters (10,21 @x40; 3,11 @x20; 100,21 for gt). */
int cvb_postproc_maps(const uint8_t* np_bin, const float* hv, const int32_t* type_map, int B, int H, int W, int n_types,
                      int object_size, int ksize, int32_t* labels, cvb_inst_row* table, int32_t* counts, int max_rows,
                      uint8_t* dbg_blb, double* dbg_dist, int32_t* dbg_marker, void* workspace, size_t ws_bytes,
                      void* stream);

/* Contours: replaces  cv2.findContours(inst_map, RETR_TREE, CHAIN_APPROX_SIMPLE)[0][0]  per instance
 * (post_proc_cellvit.py:106-125) for the instances in table[b, :counts[b]]. pts int16 (x,y) [B,max_rows,max_pts] in
 * tile coordinates; npts int32 [B,max_rows] = number of points, or -1 when the host must resolve the instance with
 * cv2 (more than one 8-connected component -- cv2 then lists the last one first -- or more than max_pts points).
 * workspace == NULL: the caller vouches that every id is one 8-connected component (true for the label maps cvb_postproc
 * writes: markers are 4-connected components and the flood labels 4-neighbours only); the multi-component check is skipped. */
int cvb_contours_workspace_bytes(int B, int H, int W, size_t* out);
int cvb_contours(const int32_t* labels, const cvb_inst_row* table, const int32_t* counts, int B, int H, int W, int max_rows,
                 int max_pts, int16_t* pts, int32_t* npts, void* workspace, size_t ws_bytes, void* stream);

/* Cell tokens: replaces the per-cell loop  tokens[idx, :, floor(rmin/P):ceil(rmax/P), floor(cmin/P):ceil(cmax/P)]
 * -> mean over the window  (cell_segmentation/inference/cell_detection.py:397-409) for the instances in
 * table[b, :counts[b]]. tokens fp32 [B,D,th,tw] (the forward's `tokens` output), patch = 16 (model.patch_size);
 * out fp32 [B,max_rows,D]; rows >= counts[b] are left untouched. */
int cvb_cell_tokens(const float* tokens, const cvb_inst_row* table, const int32_t* counts, int B, int D, int th, int tw,
                    int patch, int max_rows, float* out, void* stream);

/* Polygon overlap for the WSI-level duplicate removal: replaces the shapely calls of
 * CellPostProcessor._remove_overlap (cell_segmentation/inference/cell_detection.py:687-767) --
 * Polygon(contour).area and query.intersection(other).area -- for a list of candidate pairs (envelope hits).
 * pts_xy fp64 [n_points,2] (all contours concatenated, global WSI coordinates), poly_off int32 [n_poly+1],
 * pairs int32 [n_pairs,2]. Outputs (device): poly_area fp64 [n_poly] (nullable), inter_area fp64 [n_pairs]; -1 marks
 * a pair the kernel cannot handle (a contour with more than 128 points or pathological crossing counts): the host
 * resolves it with the same algorithm. */
int cvb_polygon_overlap(const double* pts_xy, const int32_t* poly_off, int n_poly, const int32_t* pairs, int n_pairs,
                        double* poly_area, double* inter_area, void* stream);

/* ------------------------------------------------------------------------------------------------ WSI-level export (host only)
 * Replaces the per-cell record dicts and json dumps of process_wsi (cell_segmentation/inference/cell_detection.py:352-409 records,
 * :438-475 cells.json / cell_detection.json / *.geojson, :538-597 convert_geojson): the cells of a slide live in columns (HOST
 * pointers, global WSI coordinates) and the files are streamed from them. Format = Python's json.dumps(obj, indent) byte for byte
 * (floats as float.__repr__); indent < 0 = compact. */
typedef struct cvb_cell_columns {
    long long n;                /* cells                                                                              */
    const int64_t* bbox;        /* [n,2,2] [[rmin,cmin],[rmax,cmax]] + global offset (:357-358)                        */
    const double* centroid;     /* [n,2]                                                        (:359)                */
    const int64_t* contour_pts; /* [contour_off[n],2] (x, y) of all contours                    (:360)                */
    const int64_t* contour_off; /* [n+1]                                                                              */
    const double* type_prob;    /* [n]                                                                                */
    const int64_t* type;        /* [n]                                                                                */
    const int64_t* patch;       /* [n,2] tile (row, col)                                        (:366)                */
    const int64_t* status;      /* [n] cell_status 0..8                                         (:372-374)            */
    const int64_t* offset;      /* [n,2] offset_global of the tile                              (:343-350)            */
    const uint8_t* edge;        /* [n] edge_position                                            (:376-392)            */
    const int8_t* position;     /* [n,4] [top, right, down, left] border flags (read for edge cells; edge_patches follow
                                   from them and the tile coordinates, :877-902)                                     */
} cvb_cell_columns;
enum { CVB_JSON_CELLS = 0,      /* full records of cells.json                                                         */
       CVB_JSON_DETECTION = 1,  /* {"bbox", "centroid", "type"} records of cell_detection.json                        */
       CVB_JSON_POLYGONS = 2,   /* GeoJSON MultiPolygon coordinates: [[closed ring]] per cell   (:569-575)            */
       CVB_JSON_POINTS = 3 };   /* GeoJSON MultiPoint coordinates: centroids                                          */
/* One array rendered from the columns over the cells idx[0..n_idx), between two caller-rendered strings: the file is the
 * concatenation of head + array + tail over all sections. depth = nesting depth of the array (for indented output). */
typedef struct cvb_json_section {
    const char* head;
    const char* tail;
    const int64_t* idx;
    long long n_idx;
    int kind;
    int depth;
} cvb_json_section;
int cvb_export_json(const char* path, const cvb_cell_columns* cols, const cvb_json_section* sections, int n_sections, int indent);

/* ------------------------------------------------------------------------------------------------ operator level
 * The individual device operators, exposed for unit parity tests and for callers that want to compose them.
 * cvb_tc_epilogue mirrors TcEpilogue in cellvit_b200/csrc/tc_gemm.h (see that header for field semantics). */
struct TcEpilogue;
int cvb_tc_epilogue_bytes(void);
int cvb_op_gemm_f16(const void* A, int M, int K, long long lda, const void* W, int N, long long ldw, int block_n,
                    const struct TcEpilogue* epi, void* stream);
int cvb_op_conv3x3_f16(const void* src0, int C0, const void* src1, int C1, int NB, int H, int W, const void* Wp, int N,
                       int block_n, const struct TcEpilogue* epi, void* stream);
int cvb_op_layernorm_f16(const float* x, const float* gamma, const float* beta, float eps, int rows_dst, int D, void* out,
                         int map, int B, int tok_h, int tok_w, int ws, int g, void* stream);
/* Rh / Rw: fp16 rel-pos tables [2*gh-1, hd] / [2*gw-1, hd] (both null: no bias). */
int cvb_op_attention(const void* qkv, int Gb, int S, int heads, int hd, float scale, const void* Rh, const void* Rw,
                     int gh, int gw, void* out, void* stream);
/* tcgen05 attention for the SAM global-attention shape (head dim 80, 64-wide token grid, S % 128 == 0, rel-pos tables
 * required); workspace (1024-byte aligned) holds V^T and the rel-pos bias tables of the call. */
int cvb_op_attention_tc_workspace_bytes(int Gb, int S, int heads, size_t* out);
int cvb_op_attention_tc(const void* qkv, int Gb, int S, int heads, int hd, float scale, const void* Rh, const void* Rw,
                        int gh, int gw, void* out, void* workspace, size_t ws_bytes, void* stream);
/* tcgen05 attention for the 14 x 14 SAM windows (window_tc.cu; image_encoder.py:235-260 on window_partition'ed tokens):
 * qkv fp16 [n_items*196, 3*D] in window order, relcat fp16 [64, 80] (rel_h table rows at 0.., rel_w table rows at 32..),
 * out fp16 [n_items*196, D]. One kernel, no workspace. sched_counter (nullable): device int that is ZERO at launch; the
 * persistent CTAs then claim their (window, head) items from it instead of walking a static list. */
int cvb_op_window_attention_tc(const void* qkv, int n_items, int heads, int hd, float scale, const void* relcat, void* out,
                               int32_t* sched_counter, void* stream);
int cvb_op_patch_im2col(const float* x, int B, int H, int W, int P, void* out, void* stream);
/* Canvas helpers of cvb_forward for non-native tile sizes. copy_planes: dst[p][y][x] = (y < sH && x < sW) ? src[p][y][x] : 0 for
 * y < dH, x < dW over `planes` planes of elements of 1, 4 or 16 bytes (a crop when dst is smaller, a zero-extending embed when
 * larger). zero_margin: zero the pixels with y >= vH or x >= vW of NHWC planes [planes][H][W][bytes_per_px] (multiple of 16). */
int cvb_op_copy_planes(const void* src, int sH, int sW, void* dst, int dH, int dW, long long planes, int elem_bytes, void* stream);
int cvb_op_zero_margin(void* buf, long long planes, int H, int W, int bytes_per_px, int vH, int vW, void* stream);
int cvb_op_stem_conv(const float* x, int B, int H, int W, const float* w, const float* scale, const float* shift,
                     void* out, int cpad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CELLVIT_B200_H */
