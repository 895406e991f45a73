/*
 * pegasus_b200.h — C ABI of libpegasus_b200.so: the B200-native (sm_100a) replacement for the
 * compose -> forward-rasterize hot path of meyerls/PEGASUS.
 *
 * Every entry point takes plain pointers and sizes (device pointers unless stated), never a torch
 * type, and enqueues work on the caller's CUDA stream.  The library allocates no device memory and
 * keeps no state between calls except a thread-local error string (and the opt-in profiling events):
 * inputs, outputs and the workspace belong to the caller (the reference's binding hands the extension
 * three grow-only torch buffers the same way); everything that modifies how ONE call is issued
 * (a second stream for compositing, an event to record, the numerics mode) travels in that call's
 * pg_launch_opts.
 *
 * Which reference interface each entry replaces (paths relative to /root/reference,
 * GSP = submodules/gaussian-splatting-pegasus):
 *
 *   pg_rasterize_forward   <- diff_gaussian_rasterization._C.rasterize_gaussians, i.e. what
 *                             GaussianRasterizer.forward() calls (GSP/gaussian_renderer/__init__.py:14,
 *                             :38-53 settings, :87-95 call, 3-tuple return color/radii/depth).  The
 *                             extension source is the un-vendored submodule
 *                             meyerls/depth-diff-gaussian-rasterization @ 0062df97 (absent from the tree).
 *   pg_mark_visible        <- GaussianRasterizer.markVisible (upstream API; unused by PEGASUS).
 *   pg_pose_apply          <- GaussianModel.apply_transformation (src/gs/gaussian_model.py:579-582:
 *                             apply_transformation_on_xyz :494-497, apply_rotation_on_splats :499-505,
 *                             apply_rotation_on_sh :507-546) + merge_gaussians (:584-591) as called from
 *                             PegasusSetup.apply_transformation_on_gs (src/gs/pegasus_setup.py:195-207)
 *                             and the per-frame merge in pegasus.py:255-264.
 *   pg_render_composed     <- the K+3 passes of src/gs/render.py:14-129 (render_rgb_and_depth,
 *                             render_silhouette_mask, render_visib_mask,
 *                             render_semanticsegmentation_mask) driven by pegasus.py:295-332.
 *   pg_export_binning      <- test/debug only: the reference keeps these in its binningBuffer /
 *                             imgBuffer (point_list_keys, point_list, ranges).
 *   pg_pack_frame,         <- the host-side conversions in pegasus.py:340-358 (rgb*255 -> u8,
 *   pg_pack_masks             depth*1000 -> u16, masks -> 0/255 u8) done on the device before the D2H copy.
 *   pg_png_encode          <- the PNG encoding of every product by imageio inside the writer threads
 *                             (pegasus.py:346-358, src/tools/pegasus_working.py:407-438): the zlib stream of each
 *                             image is produced on the device; the host only frames it as a PNG file.
 */
#ifndef PEGASUS_B200_H
#define PEGASUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_OK 0
#define PG_ERR_INVALID (-1)   /* bad argument */
#define PG_ERR_CUDA (-2)      /* a CUDA call failed; see pg_last_error() */
#define PG_ERR_WORKSPACE (-3) /* workspace smaller than pg_workspace_bytes() */
#define PG_ERR_CAPACITY (-4)  /* pair capacity exceeded (reported by pg_read_status) */

#define PG_MAX_OBJECTS 32
#define PG_MAX_COLORS 64

typedef void* pg_stream_t; /* cudaStream_t */
typedef void* pg_event_t;  /* cudaEvent_t */

/* Mirrors GaussianRasterizationSettings (GSP/gaussian_renderer/__init__.py:38-51).
 * bg/viewmatrix/projmatrix/campos are DEVICE pointers, exactly as the reference passes CUDA tensors;
 * viewmatrix/projmatrix are the contiguous transposed 4x4 (flat index [4*col+row]). */
typedef struct pg_raster_settings {
    int32_t image_height;
    int32_t image_width;
    float tanfovx;
    float tanfovy;
    const float* bg;         /* [3] */
    float scale_modifier;
    const float* viewmatrix; /* [16] */
    const float* projmatrix; /* [16] */
    int32_t sh_degree;
    const float* campos;     /* [3] */
    int32_t prefiltered;
    int32_t debug;           /* bit 0: synchronise after every stage; bit 1: collect compositing statistics;
                              * bit 2: keep the reference's complete (tile, Gaussian) pair lists instead of only the
                              * pairs that can contribute (same images; for pg_export_binning / n_contrib parity) */
} pg_raster_settings;

/* Arguments of GaussianRasterizer.forward (GSP/gaussian_renderer/__init__.py:87-95). */
typedef struct pg_gaussians {
    int32_t P;
    const float* means3D;        /* [P,3] */
    const float* shs;            /* [P,M,3] or NULL */
    int32_t sh_coeffs;           /* M (16 for max degree 3) */
    const float* colors_precomp; /* [P,3] or NULL */
    const float* opacities;      /* [P] (the reference's [P,1]) */
    const float* scales;         /* [P,3] or NULL */
    const float* rotations;      /* [P,4] (w,x,y,z), unit, or NULL */
    const float* cov3D_precomp;  /* [P,6] or NULL */
} pg_gaussians;

/* The reference's 3-tuple plus optional side channels (NULL = not wanted). */
typedef struct pg_raster_outputs {
    float* color;        /* [3,H,W] */
    int32_t* radii;      /* [P] */
    float* depth;        /* [1,H,W] */
    float* final_T;      /* [H,W] optional: transmittance; alpha = 1 - final_T */
    uint32_t* n_contrib; /* [H,W] optional: position of the last contributor in the reference's tile list;
                          * requesting it implies debug bit 2 */
} pg_raster_outputs;

/* Objects of a composed scene: Gaussians [first[k], first[k+1]) belong to object k, everything
 * below first[0] is environment (merge order of src/gs/gaussian_model.py:584-591). HOST data. */
typedef struct pg_object_table {
    int32_t num_objects;                    /* K <= PG_MAX_OBJECTS */
    int32_t first[PG_MAX_OBJECTS + 1];
    int32_t color_index[PG_MAX_OBJECTS];    /* bullet_id - 1: row of `colors` (src/gs/render.py:46) */
    int32_t num_colors;                     /* rows of the colour set (<= PG_MAX_COLORS) */
    float colors[PG_MAX_COLORS][3];         /* generate_colors() (src/utility/graphic_utils.py:40-60) */
} pg_object_table;

/* Outputs of one composed frame = what the reference's K+3 passes produce. NULL = skip. */
typedef struct pg_frame_outputs {
    float* color;        /* [3,H,W] RGB pass */
    int32_t* radii;      /* [P] */
    float* depth;        /* [1,H,W] */
    float* final_T;      /* [H,W] optional */
    float* seg_color;    /* [3,H,W] objects-only flat-colour render (visib / sem-seg pass), optional */
    uint8_t* sem_seg;    /* [H,W,3] uint8(255*seg) (src/gs/render.py:129) */
    uint8_t* visible;    /* [num_colors,H,W] 0/1 (src/gs/render.py:89-93) */
    uint8_t* silhouette; /* [num_colors,H,W] 0/1 (src/gs/render.py:60-63) */
} pg_frame_outputs;

/* One rigid pose per object (412 bytes, all 4-byte fields): x' = R (x - pivot) + pivot + t; q' = q_R (x) normalize(q);
 * SH bands l=1..3 multiplied by D1/D2/D3 (row-major). */
typedef struct pg_pose {
    float R[9];
    float t[3];
    float pivot[3];
    float q[4];  /* (w,x,y,z) of R */
    float D1[9];
    float D2[25];
    float D3[49];
    int32_t rotate_sh; /* 0: copy canonical SH (what pegasus.py:262-263 ends up rendering) */
} pg_pose;

/* Canonical (un-posed) object clouds, concatenated in merge order. DEVICE pointers. */
typedef struct pg_canonical {
    int32_t n_total;
    const float* xyz;           /* [n,3] */
    const float* rotation;      /* [n,4] raw (w,x,y,z) */
    const float* features_rest; /* [n,15,3] */
} pg_canonical;

/* Destination: the composed scene's rasterizer-input arrays. DEVICE pointers. */
typedef struct pg_scene {
    int32_t P;
    float* means3D;   /* [P,3] */
    float* rotations; /* [P,4] */
    float* shs;       /* [P,16,3] */
} pg_scene;

typedef struct pg_status {
    uint32_t num_rendered; /* the reference's R: sum of the tile rectangles of all visible Gaussians */
    uint32_t overflow;     /* !=0: the stored pairs exceeded the workspace's pair capacity; outputs are invalid */
    uint32_t num_visible;
    uint32_t num_stored;   /* pairs actually stored and sorted (== R with debug bit 2), clamped to the capacity */
    /* sticky: accumulated over every forward that used this workspace since pg_workspace_init (a forward clears
     * the four fields above, never these two), so a caller that checks once after many frames misses nothing */
    uint32_t overflow_frames;  /* forwards whose stored pairs exceeded the pair capacity */
    uint32_t max_pairs_needed; /* largest stored-pair demand seen (what pair_capacity should have been), <= 2^30 */
} pg_status;

/* Numerics of the compositing stage. */
#define PG_NUMERICS_EXACT 0 /* every float op an individually rounded IEEE op in the reference's order, exp as an
                             * FMA polynomial: images bit-reproducible on the CPU oracle */
#define PG_NUMERICS_FAST 1  /* exp through MUFU ex2.approx (as the reference's own expf is) and the blend weight
                             * alpha*T formed once per Gaussian: within the reference tolerances (RGB 1e-3,
                             * depth 1e-4 relative), not bit-reproducible on a CPU */

/* Per-call options of pg_rasterize_forward / pg_render_composed (NULL = all defaults).  Plain data, read
 * during the call only.
 *  scene_read_event : recorded on `stream` right after the per-Gaussian stage, the last stage that reads the
 *                     scene arrays (means3D, shs, opacities, scales, rotations): a pg_pose_apply for the
 *                     following frame on another stream only has to wait for this event.
 *  composite_stream : the compositing kernel runs there instead of on `stream`; fork_event is recorded on
 *                     `stream` after the binning stages and waited for by composite_stream, join_event is
 *                     recorded on composite_stream after compositing and waited for by `stream`, so for the
 *                     caller the whole frame is still ordered on `stream`.  With `stream` created at a HIGHER
 *                     priority the latency-bound stages of the following frame (another stream pair) co-run
 *                     with this frame's compositing.  Needs both events.
 *  status_host      : PINNED host memory; the status block is copied there at the end of the frame (on
 *                     `stream`), so the caller can check every frame for overflow when it retires it. */
typedef struct pg_launch_opts {
    pg_event_t scene_read_event;
    pg_stream_t composite_stream;
    pg_event_t fork_event;
    pg_event_t join_event;
    pg_status* status_host;
    int32_t numerics; /* PG_NUMERICS_* */
} pg_launch_opts;

const char* pg_version(void);
const char* pg_last_error(void);

/* Bytes of workspace needed for P Gaussians, a WxH image and at most pair_capacity pairs. */
size_t pg_workspace_bytes(int32_t P, int32_t width, int32_t height, uint64_t pair_capacity);

/* Once per (re)allocated workspace, before its first use: clears the sticky status fields. */
int pg_workspace_init(void* workspace, size_t workspace_bytes, pg_stream_t stream);

int pg_rasterize_forward(const pg_raster_settings* settings, const pg_gaussians* g,
                         const pg_raster_outputs* out, void* workspace, size_t workspace_bytes,
                         uint64_t pair_capacity, const pg_launch_opts* opts, pg_stream_t stream);

int pg_render_composed(const pg_raster_settings* settings, const pg_gaussians* g,
                       const pg_object_table* objects, const pg_frame_outputs* out, void* workspace,
                       size_t workspace_bytes, uint64_t pair_capacity, const pg_launch_opts* opts,
                       pg_stream_t stream);

/* Asynchronously copies the workspace's status block to host_status (pinned host memory). */
int pg_read_status(const void* workspace, pg_status* host_status, pg_stream_t stream);

int pg_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, uint8_t* present,
                    pg_stream_t stream);

/* poses_dev is a DEVICE array of K pose packets (e.g. the NCCL receive buffer of the per-frame pose
 * broadcast); object k's Gaussians [first[k], first[k+1]) of the canonical arrays are written to
 * scene rows scene_offset + [first[k], first[k+1]). `first` is HOST data. */
int pg_pose_apply(int32_t num_objects, const int32_t* first, const pg_pose* poses_dev,
                  const pg_canonical* canon, int32_t scene_offset, const pg_scene* scene,
                  pg_stream_t stream);

/* Test/debug: rebuild the sorted 64-bit keys, point list and tile ranges of the last forward that
 * used `workspace` — the reference's when that forward ran with debug bit 2, else the stored
 * sub-lists.  keys/point_list hold >= num_stored entries. */
int pg_export_binning(const void* workspace, int32_t P, int32_t width, int32_t height,
                      uint64_t pair_capacity, uint64_t* keys, uint32_t* point_list,
                      uint32_t* ranges /*[tiles,2]*/, pg_stream_t stream);

/* ---- opt-in profiling used by bench.py / tests (no effect on results) ------------------------------
 * Stage boundaries of one forward: 0 clear, 1 preprocess, 2 depth sort, 3 emit, 4 tile scan,
 * 5 tile sort, 6 compositing.  pg_profile_enable(n) records CUDA events on the caller's stream for
 * the next n forwards; pg_profile_read(i, ms) returns frame i's per-stage milliseconds. */
#define PG_NUM_STAGES 7
int pg_profile_enable(int32_t max_frames);
int32_t pg_profile_frames(void);
int pg_profile_read(int32_t frame, float* stage_ms /*[PG_NUM_STAGES]*/);
uint64_t pg_launch_count(void); /* kernels this library has launched in this process */
/* settings.debug bit 1: {pixel-Gaussian pairs evaluated, reaching exp, blended, pixel slots walked,
 * warp-hits of environment entries, of object entries while a main chain lives, of object entries afterwards,
 * 32-entry cull passes afterwards} of the last forward */
int pg_read_stats(const void* workspace, uint64_t* host_stats8, pg_stream_t stream);

/* rgb [3,H,W] f32 -> [H,W,3] u8 ; depth [1,H,W] f32 metres -> [H,W] u16 millimetres. */
int pg_pack_frame(int32_t width, int32_t height, const float* color, const float* depth,
                  uint8_t* rgb_u8, uint16_t* depth_u16, pg_stream_t stream);

/* n_planes mask planes [n_planes,H,W] u8 (0 / non-zero: the `visible` / `silhouette` outputs) ->
 * [n_planes,H,ceil(W/8)] bytes, pixel x in bit x % 8 of byte x / 8 (numpy.unpackbits(..., bitorder="little")).
 * The masks are 2 x n_colours of the 3+2+3+2 x n_colours bytes per pixel a frame sends to the host; packed
 * they cross PCIe as one bit per pixel and are expanded by the writer thread that encodes the PNG. */
int pg_pack_masks(int32_t width, int32_t height, int32_t n_planes, const uint8_t* masks, uint8_t* bits,
                  pg_stream_t stream);

/* ---- PNG streams on the device -------------------------------------------------------------------------
 * One call encodes n_images images of the same width x height (a frame's RGB, depth, semantic map and mask
 * planes): per image the complete zlib stream of its PNG IDAT chunk — Sub-filtered scanlines, run-length matches,
 * one dynamic-Huffman deflate block, Adler-32 — is written to `out`; `result[0]` receives its length in bytes,
 * `result[1]` becomes non-zero when it did not fit `out_capacity` (the stream is then unusable: grow and encode
 * again).  The Huffman code is NOT derived per image: `table` is a u32[PG_PNG_TABLE_WORDS] device array built on the
 * host (pegasus_b200/png_codec.py: build_table) from the token histogram of sample images, which a call accumulates
 * into `hist` (u32[PG_PNG_HIST_WORDS], caller-zeroed) when that pointer is not NULL.  Any table encodes any image
 * correctly; the ratio depends on how typical the sample was.  `images` is a HOST array. */
#define PG_PNG_RGB8 0   /* src u8 [H][W][3]                         -> 8-bit RGB */
#define PG_PNG_GRAY16 1 /* src u16 [H][W]                           -> 16-bit grayscale (big-endian in the file) */
#define PG_PNG_MASK8 2  /* src u8 [H][W], non-zero = set            -> 8-bit grayscale 0 / 255 */
#define PG_PNG_TABLE_WORDS 610
#define PG_PNG_HIST_WORDS 286

typedef struct pg_png_image {
    const void* src;       /* device */
    int32_t kind;          /* PG_PNG_* */
    int32_t src_pitch;     /* bytes between source rows */
    uint8_t* out;          /* device, 16-byte aligned */
    uint32_t out_capacity; /* bytes, a multiple of 16; pg_png_worst_case_bytes() never overflows */
    const uint32_t* table; /* device */
    void* scratch;         /* device, pg_png_scratch_bytes(height), 16-byte aligned */
    uint32_t* hist;        /* device or NULL */
    uint32_t* result;      /* device u32[2] */
} pg_png_image;

size_t pg_png_scratch_bytes(int32_t height);
size_t pg_png_worst_case_bytes(int32_t kind, int32_t width, int32_t height);
int pg_png_encode(int32_t n_images, const pg_png_image* images, int32_t width, int32_t height, pg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PEGASUS_B200_H */
