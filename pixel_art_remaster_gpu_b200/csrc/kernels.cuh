// Launch interfaces of the sm_100a kernels (one translation unit per stage).
#pragma once
#include "../../include/pixelart_b200.h"
#include "common.cuh"
#include "polygon.cuh"
#include "smooth_table.h"
#include <string.h>

namespace par {

struct GraphArgs
{
    const uint8_t* bgr;  // frames, BGR8
    uint8_t* graph_aux;  // out, dense width*height per frame
    int width, height, widthstep, n_frames;
    size_t frame_stride;
};

struct CrossArgs
{
    const uint8_t* graph_aux; // in, dense
    uint8_t* graph;           // out, dense
    int width, height, n_frames;
};

struct LabelArgs
{
    const uint8_t* graph; // in, dense
    int32_t* labels;      // out, dense
    int width, height, n_frames;
};

// device pointers of the smoothing tables (smooth_table.h); cut / link are per scale
struct SmoothTablePtrs
{
    const uint2* head;      // [kCellKeys] link descriptors 0, 1 + flags, as this scale needs them (per scale)
    const uint2* head2;     // [kCellKeys] link descriptors 2, 3 (per scale)
    const uint2* pack;      // [kCellKeys] IDs of the key's neighbour records: x = directions 4..7 in bits [12,32), y = directions 0..3
    const uint64_t* cut;    // [kCellKeys][16] entries (per scale)
    const uint64_t* link;   // [link_entries] entries: kNbrIds per class (per scale)
};

struct RasterArgs
{
    const uint8_t* bgr;
    const uint8_t* graph;       // final graph, dense
    CellTablePtrs tables;       // 4096-entry cell tables (cell_table.h), device pointers
    const uint32_t* mask_lut;   // per-scale coverage masks of the 4096 plain hulls (raster only)
    SmoothTablePtrs smooth;     // smoothing tables (raster only); cut == null: every smoothed cell takes the geometric path
    unsigned long long* smooth_stats; // [0] smoothed cells, [1] of which took the geometric path (may be null)
    uint8_t* rgba;              // out (raster), may be null
    float* polygons;            // out (polygon export), may be null
    int32_t* poly_count;        // out (polygon export), may be null
    int width, height, widthstep, n_frames;
    size_t frame_stride;
    int scale;            // sampling scale S (= output scale x aa)
    int aa;               // samples per output pixel and axis (1 = off, 2, 4): A x A samples are averaged in the resolve step
    int subdivide;
    int flip_output;
    int debug_force_wide; // test hook: treat every cell as reaching beyond its mask (exercises the exact slow path)
    int out_format;       // par_out_format of `rgba`
    const uint32_t* pal_lut;  // INDEX8: per frame kPaletteSlots words index << 24 | colour (palette_kernels.cu)
    const int32_t* pal_count; // INDEX8: colours per frame (> 256: the frame's index image is undefined)
};

// per-frame colour -> palette index lookup table (open addressing, linear probing)
constexpr int kPaletteSlots = 1024;
struct PaletteArgs
{
    const uint8_t* bgr;
    int width, height, widthstep, n_frames;
    size_t frame_stride;
    uint32_t* lut;       // [n_frames][kPaletteSlots] scratch -> out: index << 24 | (R | G << 8 | B << 16), empty 0xFFFFFFFF
    int32_t* count;      // [n_frames] out: colours incl. black
    uint32_t* palette;   // [n_frames][256] out: RGBA8 words, ascending by B << 16 | G << 8 | R; may be null
};
// returns the number of kernels launched through *n_launches
cudaError_t launch_palette( const PaletteArgs& a, cudaStream_t stream, int* n_launches );

dim3 similarity_graph_grid( int width, int height, int n_frames );
void similarity_graph_tma_box( uint32_t box[ 3 ] );
cudaError_t launch_similarity_graph( const GraphArgs& a, const CUtensorMap* img_map, cudaStream_t stream );

void resolve_crossings_tma_box( uint32_t box[ 3 ] );
cudaError_t launch_resolve_crossings( const CrossArgs& a, const CUtensorMap* aux_map, cudaStream_t stream );

// labels: returns the number of kernels launched through *n_launches
cudaError_t launch_cc_labels( const LabelArgs& a, cudaStream_t stream, int* n_launches );

cudaError_t launch_border_walks( const uint8_t* graph, const int32_t* labels, int width, int height, int n_frames, int32_t* walk_len,
                                 int32_t* walk_begin, long long* total, int32_t* nodes, long long capacity, cudaStream_t stream );

cudaError_t launch_walk_splines( const int32_t* walk_len, const int32_t* walk_begin, const long long* total, const int32_t* nodes, int width, int height,
                                 int n_frames, long long capacity, int samples, float* points, cudaStream_t stream );

cudaError_t launch_polygons( const RasterArgs& a, cudaStream_t stream );
void raster_tma_box( int scale, uint32_t box[ 3 ] );
size_t mask_lut_words( int scale );
size_t smooth_entry_words( int scale ); // 64-bit words per CUT / LINK entry
// desc: the keys' generic descriptors (SmoothTables::desc) on the device; canon: n_canon x lo, then n_canon x span (bytes);
// head: [2][kCellKeys] out (descriptors 0, 1 then 2, 3); scratch: [256] bytes
cudaError_t launch_build_smooth_tables( int scale, const CellTablePtrs& tab, const LinkClass* d_classes, int n_classes, const uint4* desc, const uint8_t* canon,
                                        int n_canon, uint64_t* cut, uint64_t* link, uint2* head, uint8_t* scratch, cudaStream_t stream );
cudaError_t launch_build_mask_lut( int scale, const CellTablePtrs& tab, uint32_t* lut, cudaStream_t stream );
void raster_img_tma_box( int scale, uint32_t box[ 3 ] );
cudaError_t launch_raster( const RasterArgs& a, const CUtensorMap* graph_map, const CUtensorMap* img_map, cudaStream_t stream );
cudaError_t launch_raster_bgr8( const RasterArgs& a, const CUtensorMap* graph_map, const CUtensorMap* img_map, cudaStream_t stream );
cudaError_t launch_raster_index8( const RasterArgs& a, const CUtensorMap* graph_map, const CUtensorMap* img_map, cudaStream_t stream );
bool raster_scale_supported( int scale );
bool raster_aa_supported( int out_scale, int aa );

} // namespace par
