/* pixelart_b200.h — C ABI of the B200-native pixel-art remaster path.
 *
 * Drop-in boundary for the hot path of marcoc2/pixel-art-remaster-gpu:
 *
 *   reference interface (file:line)                         replaced by
 *   ------------------------------------------------------  ---------------------------------------
 *   extern "C" Point* launch_kernel(...)  kernel.cu:286-288  launch_kernel() below — same symbol,
 *     declared by its only caller simpleVBO.cpp:57-60,        same signature, same ownership rules
 *     called simpleVBO.cpp:151-153
 *   the per-frame cudaMalloc/cudaFree + 8 launches            par_create() / par_remaster_device() /
 *     kernel.cu:313-523                                       par_remaster_host(): pooled context,
 *                                                             batches of frames, one stream
 *   graph_Kernel + trivial_cross_Kernel  kernel.cu:140-177   par_stage_similarity_graph()
 *   ambiguous_cross_Kernel               kernel.cu:180-189   par_stage_resolve_crossings()
 *   (dead) cc_Kernel / extractBorderPoints                    par_stage_cc_labels()  (new subsystem)
 *     cc_kernel_call.bkp:3-12, cc_functions.cu:348-503
 *   cells_Kernel + subdivision_Kernel    kernel.cu:192-261   par_stage_polygons()
 *   triangulate/color/position kernels + glDrawArrays         par_stage_raster()  (direct rasterizer)
 *     kernel.cu:264-282,67-137; simpleVBO.cpp:236-285
 *
 * Plain C: pointers and sizes only.  Every function returns a par_status (0 = success) unless it
 * says otherwise; par_last_error() gives the message of the last failure on that context.  There is
 * no CPU fallback: without a CUDA device par_create() fails with PAR_ERR_NO_DEVICE.
 *
 * Data conventions (identical to the reference):
 *   frame   BGR8, `widthstep` bytes per row, row 0 = BOTTOM scanline (main.cpp:59 flips on load)
 *   graph   1 byte per pixel, dense rows of `width` bytes; bit e <-> neighbour (di,dj):
 *           0:(-1,+1) 1:(0,+1) 2:(+1,+1) 3:(-1,0) 4:(+1,0) 5:(-1,-1) 6:(0,-1) 7:(+1,-1)
 *   labels  int32 per pixel = smallest row-major index (within the frame) of the pixel's component
 *   rgba    (scale*height) rows of (scale*width) RGBA8 pixels; row Y = pipeline row (0 = bottom)
 *           unless PAR_FLAG_FLIP_OUTPUT is set, in which case row 0 = top scanline
 *   polygons  PAR_CELL_SLOTS (x,y) float pairs per pixel in cell-local coordinates + int32 count
 * A batch is `n_frames` frames `frame_stride` bytes apart; all outputs are dense per frame.
 */
#ifndef PIXELART_B200_H
#define PIXELART_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PAR_CELL_SLOTS 45 /* kernel.cu:8 CELL_SIZE */

typedef enum par_status
{
    PAR_OK = 0,
    PAR_ERR_INVALID = 1,   /* bad argument */
    PAR_ERR_NO_DEVICE = 2, /* no usable CUDA device / wrong architecture */
    PAR_ERR_CUDA = 3,      /* a CUDA call failed; see par_last_error() */
    PAR_ERR_CAPACITY = 4   /* batch larger than the context was created for */
} par_status;

enum
{
    PAR_FLAG_SUBDIVIDE = 1u << 0,        /* stage E on (launch_kernel's `subdivide`, kernel.cu:288) */
    PAR_FLAG_FLIP_OUTPUT = 1u << 1,      /* image row 0 = top scanline (undo main.cpp:59's flip) */
    PAR_FLAG_NO_TMA = 1u << 2,           /* force the plain-load tile path (also taken automatically when pointers/strides
                                            are not 16-byte multiples) */
    PAR_FLAG_DEBUG_WIDE = 1u << 3,       /* test hook: rasterize every cell through the exact slow path that normally only
                                            handles cells reaching beyond their sample mask */
    PAR_FLAG_NO_SMOOTH_TABLES = 1u << 4, /* do not use the smoothing tables: every smoothed cell builds its polygon and
                                            rasterizes it (the geometric path; same image, for tests) */
    /* Anti-aliased output (the reference's GL_MULTISAMPLE toggle, simpleVBO.cpp:238-253, main.cpp:233): every output
     * pixel is the mean of 2x2 / 4x4 ordered-grid samples (the point-sampling rule at 2x / 4x the scale, averaged per
     * channel, halves rounded up).  scale x samples must be a supported scale (1..8): AA2 with scale 1..4; AA4 with 1,2. */
    PAR_FLAG_AA2 = 1u << 5,
    PAR_FLAG_AA4 = 1u << 6
};

/* Layout of the output image (par_job.out_format).  Every output pixel of the point-sampled image is the colour of a
 * source pixel (kernel.cu:98-101) or the black background (main.cpp:260), so the image has at most as many colours as
 * the frame + 1 and can be returned losslessly as palette indices: a quarter of the bytes of RGBA8. */
typedef enum par_out_format
{
    PAR_OUT_RGBA8 = 0,  /* 4 bytes per pixel: R, G, B, 255 (the colours of the reference's colour VBO, kernel.cu:98-101) */
    PAR_OUT_BGR8 = 1,   /* 3 bytes per pixel: B, G, R — the 3-channel image Image::saveImage writes (Image.cpp:64-71) */
    PAR_OUT_INDEX8 = 2  /* 1 byte per pixel: index into the frame's palette (par_job.palette: 256 RGBA8 entries per frame,
                           ascending by R<<16|G<<8|B, entry 0 is always black = the background; par_job.palette_count: entries
                           used).  A frame with more than 256 colours reports its count (> 256) and its index image is
                           undefined: ask for RGBA8 / BGR8 for such frames.  Not available with PAR_FLAG_AA2 / AA4. */
} par_out_format;

typedef struct par_context par_context;

/* One batch of frames and where its results go.  Pointers are DEVICE pointers for
 * par_remaster_device()/par_stage_*() and HOST pointers for par_remaster_host().  Any output may be
 * NULL (not produced / not copied). */
typedef struct par_job
{
    const uint8_t* bgr;   /* in : n_frames frames, BGR8                                  */
    int width, height;    /*      frame size in pixels                                   */
    int widthstep;        /*      bytes per row (>= 3*width)                             */
    size_t frame_stride;  /*      bytes between frames (>= widthstep*height); 0 = dense  */
    int n_frames;
    int scale;            /*      output magnification s: an integer 1..8 (the reference's viewer
                                  steps a float scale by 0.5, callbacks.cpp:162-166; the rasterizer
                                  here samples on an integer grid only)                          */
    unsigned flags;       /*      PAR_FLAG_*                                             */
    uint8_t* rgba;        /* out: the image, n_frames * (s*height) * (s*width) * {4,3,1} bytes
                                  (out_format below; RGBA8 unless set)                           */
    uint8_t* graph;       /* out: final similarity graph, n_frames * width*height bytes  */
    uint8_t* graph_aux;   /* out: graph after the trivial-crossing pass (kernel.cu:415)  */
    int32_t* labels;      /* out: connected-component labels                             */
    float* polygons;      /* out: n_frames * width*height * 45 * 2 floats                */
    int32_t* poly_count;  /* out: vertices per polygon (launch_kernel's edge_count_h)    */
    int out_format;       /*      par_out_format of `rgba` (0 = PAR_OUT_RGBA8)                   */
    uint32_t* palette;    /* out (INDEX8): n_frames * 256 RGBA8 words (R | G<<8 | B<<16 | 255<<24) */
    int32_t* palette_count; /* out (INDEX8): colours per frame incl. black; > 256: not representable */
} par_job;

/* Context: device, stream, pooled scratch (graph buffers for `max_frames` frames of up to
 * max_width x max_height pixels), the 4096-entry cell table.  Not thread-safe; one per host thread. */
int par_create( par_context** out, int device, int max_width, int max_height, int max_frames );
void par_destroy( par_context* ctx );
const char* par_last_error( const par_context* ctx ); /* ctx may be NULL: last par_create failure */
int par_device( const par_context* ctx );
/* Run all work on an existing CUDA stream (cudaStream_t passed as void*; NULL is the legacy default
 * stream).  par_use_own_stream() goes back to the non-blocking stream the context created. */
int par_set_stream( par_context* ctx, void* cuda_stream );
int par_use_own_stream( par_context* ctx );
int par_synchronize( par_context* ctx );
/* par_remaster_device() runs a batch in rounds of `frames` frames through all stages (0, the default: one launch per stage
 * over the whole batch).  Frames are independent units, so results do not depend on it; it only decides whether the
 * intermediates of a round (2 bytes per pixel) are still in L2 when the next stage reads them. */
int par_set_sub_batch( par_context* ctx, int frames );
/* Number of kernels this context has launched since creation (bench.py's `gpu_launches`). */
uint64_t par_launch_count( const par_context* ctx );

/* Per-stage device timing: when enabled, every stage launched by par_remaster_device()/par_stage_*()
 * is bracketed by CUDA events on the context's stream.  par_profile_read() synchronizes the stream,
 * adds up the event intervals recorded since the last read and returns, per stage, the total
 * milliseconds and the number of launches (arrays of PAR_N_STAGES). */
#define PAR_N_STAGES 6 /* 0 similarity graph, 1 resolve crossings, 2 cc labels, 3 polygons, 4 raster, 5 palette */
int par_profile_enable( par_context* ctx, int on );
int par_profile_read( par_context* ctx, double* total_ms, int* launches );

/* Smoothing statistics since par_create (synchronizes the stream): out2[0] = smoothed cells (stage E ran on
 * them), out2[1] = of which took the geometric path (polygon built and rasterized) instead of the
 * precomputed smoothing tables.  Both paths give the same mask; the tables only make it cheaper. */
int par_smooth_stats( par_context* ctx, uint64_t* out2 );

/* Border walk of every connected component: for each component the walk the reference's
 * extractBorderPoints (cc_functions.cu:348-503 — dead code there) makes from the component's first node in raster
 * order (= its label): clockwise along the outer face (nextNodeClockwise, :295-318) until the loop closes (:415-416);
 * dropped, as in the reference, when it steps on an interior node (== 90) or meets its start early (:425-438).  An
 * island (undefined in the reference) is a walk of one node.  All pointers are DEVICE pointers; asynchronous.
 *   graph, labels : n_frames * width*height (final graph, and its labels from par_stage_cc_labels / par_remaster_*)
 *   walk_len      : out, per pixel: nodes of the walk that starts there (0 for other pixels and dropped walks)
 *   walk_begin    : out, per pixel: where that walk starts in the frame's walk_nodes (walks are concatenated in raster
 *                   order of their start — the reference's CClist / CCsizes layout)
 *   walk_nodes    : out, n_frames * capacity_per_frame node indices
 *   total         : out, per frame: entries the frame's walks need; a frame with total > capacity_per_frame writes none */
int par_border_walks( par_context* ctx, const uint8_t* graph, const int32_t* labels, int width, int height, int n_frames, int32_t* walk_len,
                      int32_t* walk_begin, int32_t* walk_nodes, long long capacity_per_frame, long long* total );

/* Splines through the border walks (the Kopf-Lischinski stage the reference's walker, cc_functions.cu:348-503,
 * was written for and never reached): every walk is taken as a closed control polygon — the centres (x + 1/2, y + 1/2) of its
 * nodes, in source-pixel coordinates, row 0 = bottom — and sampled as the closed uniform quadratic B-spline over it, at
 * `samples_per_segment` (1, 2, 4 or 8) parameter values per node: sample (b + i) * samples + s belongs to node i of the walk
 * that begins at entry b of the frame's walk_nodes (segment i runs from the midpoint of nodes i-1, i to the midpoint of i, i+1).
 * walk_len, walk_begin, walk_nodes, total: the outputs of par_border_walks (DEVICE pointers); points: out, DEVICE,
 * n_frames * capacity_per_frame * samples_per_segment (x, y) float pairs; a frame with total > capacity_per_frame writes
 * none.  Results are exact dyadic rationals (no rounding).  Asynchronous. */
int par_walk_splines( par_context* ctx, const int32_t* walk_len, const int32_t* walk_begin, const int32_t* walk_nodes, const long long* total, int width,
                      int height, int n_frames, long long capacity_per_frame, int samples_per_segment, float* points );

/* Outlines of one frame, host in / host out: similarity graph, crossings, labels, the border walk of every component and
 * its spline (par_border_walks + par_walk_splines), gathered into one malloc'd block that par_outlines_free() releases.
 * Walks are listed in raster order of their start pixel; walk k starts at pixel `start[k]` (= its component's label; its
 * colour is the component's colour), has `count[k]` nodes and count[k] * samples curve points, stored back to back in
 * `points` (x, y pairs, source-pixel coordinates, row 0 = bottom).  What `remaster_cli --outlines out.svg` draws. */
typedef struct par_outlines
{
    int n_walks, samples;
    int32_t* start;     /* [n_walks] */
    int32_t* count;     /* [n_walks] */
    long long n_points; /* sum of count[k] * samples */
    float* points;      /* [n_points][2] */
} par_outlines;
int par_outlines_host( par_context* ctx, const uint8_t* bgr, int width, int height, int widthstep, int samples_per_segment, par_outlines* out );
void par_outlines_free( par_outlines* out );

/* Whole path on device-resident frames; asynchronous on the context's stream. */
int par_remaster_device( par_context* ctx, const par_job* job );
/* Whole path on host buffers: H2D of the frames, the kernels, D2H of every non-NULL output, then a
 * stream synchronize.  Host buffers may be pageable or pinned. */
int par_remaster_host( par_context* ctx, const par_job* job );

/* Single stages on device pointers (parity tests and embedding callers). */
int par_stage_similarity_graph( par_context* ctx, const par_job* job );  /* bgr -> graph_aux            */
int par_stage_resolve_crossings( par_context* ctx, const par_job* job ); /* graph_aux -> graph          */
int par_stage_cc_labels( par_context* ctx, const par_job* job );         /* graph -> labels             */
int par_stage_polygons( par_context* ctx, const par_job* job );          /* bgr, graph -> polygons,count */
int par_stage_raster( par_context* ctx, const par_job* job );            /* bgr, graph -> rgba          */

/* Cell table (stage D): vertices of the cell of pattern key = node | (left&4 ? 1<<8 : 0) |
 * (left&128 ? 1<<9 : 0) | (right&1 ? 1<<10 : 0) | (right&32 ? 1<<11 : 0); out_xy receives count+1
 * (x,y) pairs (first repeated last); returns count, or -1 for a bad key.  Host-side, no GPU needed. */
int par_cell_from_pattern( unsigned key, float* out_xy );
/* Packed YUV word of one colour as the device computes it (graph_functions.cu:80-98). Host-side. */
uint32_t par_yuv_word( int byte0, int byte1, int byte2 );

/* ---- multi-GPU: one very large image tiled into horizontal strips over several devices -------- */
typedef struct par_group par_group;
int par_group_create( par_group** out, const int* devices, int n_devices, int width, int height, int scale );
void par_group_destroy( par_group* grp );
const char* par_group_last_error( const par_group* grp );
/* Host image in, host outputs out (any may be NULL); strips exchange halo rows over P2P.  job->out_format may be
 * PAR_OUT_RGBA8 or PAR_OUT_BGR8 (a palette would be per strip). */
int par_group_remaster_host( par_group* grp, const par_job* job );
/* Device-resident tiled path (SURVEY §8(e), config 4 without the PCIe round trip).  Strip k owns image rows
 * [own_begin, own_end) and holds rows [load_begin, load_end) (its 40-row aprons included); all its buffers are dense rows
 * starting at row load_begin on `device`: bgr 3*width bytes per row (IN: the caller writes the strip's OWN rows, complete
 * before the call — e.g. cudaMemcpy, or a producer kernel followed by a synchronize), image scale*scale*{4,3} bytes per
 * pixel, graph / graph_aux 1, labels 4 (global pixel indices).  par_group_remaster_device() pulls the apron rows from the
 * neighbouring devices (peer copies over NVLink), runs the path on every strip, labels and stitches across the seams on
 * the devices and returns after synchronizing all strips; only the OWN rows of the outputs are meaningful. */
typedef struct par_strip
{
    int device;
    int own_begin, own_end, load_begin, load_end;
    uint8_t* bgr;
    uint8_t* image;
    uint8_t* graph;
    uint8_t* graph_aux;
    int32_t* labels;
} par_strip;
int par_group_n_strips( const par_group* grp );
int par_group_strip( const par_group* grp, int k, par_strip* out );
int par_group_remaster_device( par_group* grp, unsigned flags, int out_format, int want_image, int want_labels );
/* Duration of the last par_group_remaster_device(): host wall clock from the first enqueue to the last synchronize, and
 * the longest strip's own device time (CUDA events on its stream: waits for the neighbours included). */
int par_group_last_ms( const par_group* grp, double* wall_ms, double* device_ms );

/* ---- the reference's own entry point (kernel.cu:286-288), same symbol and signature ------------ */
#if defined( __CUDACC__ ) || defined( PAR_HAVE_CUDA_VECTOR_TYPES )
typedef struct par_point { float x, y; } par_point; /* point.cu:2-11 */
par_point* launch_kernel( float2* pos, uchar4* colorPos, float time, char* img_data, int img_width, int img_height,
                          int img_widthstep, int* edge_count_h, char* graph_h, bool subdivide );
#endif

#ifdef __cplusplus
}
#endif
#endif /* PIXELART_B200_H */
