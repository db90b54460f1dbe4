// Stages D + E + direct rasterization in one kernel, and the polygon export kernel.
//
// Replaces cells_Kernel, subdivision_Kernel, triangulate_Kernel, color_kernel, position_kernel
// (kernel.cu:192-282, :67-137) and the OpenGL draw of the 45-slot VBOs (simpleVBO.cpp:236-285):
// instead of writing 1.1 kB of geometry per source pixel for a GL driver to rasterize, the cell
// polygons are sampled directly into the (s*W) x (s*H) RGBA8 image.
//
// Raster rule (restated GL point sampling; checked by tests/test_gpu_parity.py): output
// pixel (X,Y) samples ((X+1/2)/s, (Y+1/2)/s); it takes the colour of the HIGHEST-index source
// pixel whose polygon covers the sample (cells are drawn in row-major order without depth test),
// black if none; a sample exactly on an edge/vertex is resolved as if displaced by (+eps, -eps^2)
// (top-left rule).  Coverage is an even-odd crossing count in exact integer arithmetic: vertices
// are multiples of 1/64, samples odd multiples of 1/(2s); both are scaled to units of 1/(128 s).
//
// Design (B200): one CTA per 32x32 (s <= 4) / 32x16 tile of source pixels, 24.6 KB of shared memory, five / four CTAs
// per SM.  (1) The final-graph bytes and the BGR bytes of the tile (+halo) are staged in shared memory by two TMA
// bulk-tensor copies (zero fill outside the image; plain loads when rows are not 16-byte multiples) and turned, four
// pixels per thread, into RGBA words and two 32-bit cell words per cell: the 12-bit cell key and the IDs of the cell's
// eight neighbour records (one 8-byte gather per staged cell).  (2) Every cell of tile + 1 halo gets a coverage bitmask of
// the (s + 2h)^2 samples it can reach.  Cells whose polygon is their hull (interior nodes, or subdivision off) copy it
// from a per-scale 4096-entry mask table.  Smoothed cells are handled in tile order and do not build their polygon
// at all: even-odd coverage is XOR-linear in the polygon's edges, so the mask is the XOR of a few precomputed pieces
// (smooth_table.h) — one CUT entry for the cell's own key and kept corners, one LINK entry per shared edge with a
// blended end, indexed by the class's block and the 5-bit ID found in the neighbour's staged cell word (no neighbour
// record is gathered).  The tables are content-independent; the masks and the descriptor tables are built once per
// context and scale by raster_impl.cuh's own coverage code (classes whose masks are empty at a scale share one block per
// ID range).  The few cells the tables cannot express take the geometric
// path: one thread per cell, the polygon (hull from the cell table, corner cutting in exact 1/64-px integers) streamed
// into a small per-thread vertex buffer, then a converged loop over its edges toggling the sample
// rows each edge crosses.  Geometry never touches HBM.  (3) One thread per source pixel resolves its s x s output
// pixels with bit operations over the 3x3 neighbourhood's masks in priority order (for s <= 4 on whole-cell bit sets:
// the masks are kept in a "window form" whose fields are already positioned on the reader's pixels) and writes whole
// output-row segments with 128-bit streaming stores; with anti-aliasing on, A x A samples are averaged right there.
// Algorithmic HBM traffic: 3 B/px colour + 1 B/px graph in, 4*s*s B/px out.
//
// The kernel itself lives in raster_impl.cuh and is instantiated once per output format (par_out_format), one translation
// unit each so that they compile in parallel: this file (RGBA8, and the builders of the per-scale tables), raster_bgr8.cu
// (3 bytes per pixel, the image Image::saveImage takes) and raster_index8.cu (palette indices; palette_kernels.cu).  Only
// the stores of the resolve step differ.
#include "raster_impl.cuh"

namespace par {

namespace {

// ---- polygon export: one thread per pixel, everything from global memory ----------------------
struct GlobalEnv
{
    const uint8_t* graph;
    int width, height;
    FlatImage img;
    __device__ __forceinline__ uint32_t node( int i, int j ) const
    {
        return ( i >= 0 && j >= 0 && i < width && j < height ) ? __ldg( graph + ( size_t )j * width + i ) : 0u;
    }
    __device__ __forceinline__ uint32_t key( int i, int j ) const { return cell_key( node( i, j ), node( i - 1, j ), node( i + 1, j ) ); }
    __device__ __forceinline__ bool guard( int i, int j ) const { return img.guard( i, j ); }
    __device__ __forceinline__ bool keep_corner( int i, int j, Q2 p ) const { return img.keep_corner( i, j, p ); }
};

struct LocalSlots
{
    int16_t x[ kMaxVerts ], y[ kMaxVerts ];
    __device__ __forceinline__ void put( int slot, int x64, int y64 )
    {
        x[ slot ] = ( int16_t )x64;
        y[ slot ] = ( int16_t )y64;
    }
};

__global__ void __launch_bounds__( kThreads ) polygon_kernel( RasterArgs a )
{
    const int i = blockIdx.x * 32 + ( threadIdx.x & 31 ), j = blockIdx.y * 8 + ( threadIdx.x >> 5 ), f = blockIdx.z;
    if( i >= a.width || j >= a.height ) return;
    const size_t frame_px = ( size_t )a.width * a.height;
    GlobalEnv env;
    env.graph = a.graph + ( size_t )f * frame_px;
    env.width = a.width;
    env.height = a.height;
    env.img.frame = a.bgr + ( size_t )f * a.frame_stride;
    env.img.width = a.width;
    env.img.height = a.height;
    env.img.widthstep = a.widthstep;
    const size_t n = ( size_t )f * frame_px + ( size_t )j * a.width + i;
    LocalSlots slots;
    const CellPoly poly = build_cell_polygon( env, a.tables, i, j, env.key( i, j ), a.subdivide != 0, slots );
    float* out = a.polygons + n * 2 * 45;
    int m = 0;
    for( int t = 0; t < poly.n; t++ )
        for( int e = 0; e <= ( int )( ( poly.two >> t ) & 1u ); e++, m++ )
        {
            out[ 2 * m ] = ( float )slots.x[ 2 * t + e ] * 0.015625f;
            out[ 2 * m + 1 ] = ( float )slots.y[ 2 * t + e ] * 0.015625f;
        }
    for( int t = m; t < 45; t++ ) // unused slots are zeroed (they are undefined in the reference)
    {
        out[ 2 * t ] = 0.0f;
        out[ 2 * t + 1 ] = 0.0f;
    }
    if( a.poly_count ) a.poly_count[ n ] = m;
}

template< int S >
cudaError_t build_lut_s( const CellTablePtrs& tab, uint32_t* lut, cudaStream_t stream )
{
    build_mask_lut_kernel< S ><<< kCellKeys / 128, 128, 0, stream >>>( tab, lut );
    return cudaGetLastError();
}

} // namespace

bool raster_scale_supported( int scale ) { return scale >= 1 && scale <= 8; }
bool raster_aa_supported( int out_scale, int aa )
{
    return ( aa == 1 || aa == 2 || aa == 4 ) && out_scale >= 1 && raster_scale_supported( out_scale * aa );
}

size_t mask_lut_words( int scale )
{
#define PAR_WORDS( S ) return ( size_t )kCellKeys * ( Cfg< S >::PACK ? 2 : 8 ) /* window form: two words; rows form: an entry of four 64-bit words */
    PAR_FOR_SCALE( scale, PAR_WORDS )
#undef PAR_WORDS
    return 0;
}

cudaError_t launch_build_mask_lut( int scale, const CellTablePtrs& tab, uint32_t* lut, cudaStream_t stream )
{
#define PAR_BUILD( S ) return build_lut_s< S >( tab, lut, stream )
    PAR_FOR_SCALE( scale, PAR_BUILD )
#undef PAR_BUILD
    return cudaErrorInvalidValue;
}

size_t smooth_entry_words( int scale )
{
#define PAR_EW( S ) return ( size_t )Entry< S >::EW
    PAR_FOR_SCALE( scale, PAR_EW )
#undef PAR_EW
    return 0;
}

cudaError_t launch_build_smooth_tables( int scale, const CellTablePtrs& tab, const LinkClass* d_classes, int n_classes, const uint4* desc, const uint8_t* canon,
                                        int n_canon, uint64_t* cut, uint64_t* link, uint2* head, uint8_t* scratch, cudaStream_t stream )
{
#define PAR_BUILD_SMOOTH( S )                                                                        \
    build_cut_table_kernel< S ><<< kCellKeys * 16 / 128, 128, 0, stream >>>( tab, cut );             \
    cudaMemsetAsync( link, 0, kNbrIds * Entry< S >::EW * sizeof( uint64_t ), stream ); /* block 0: the all-zero block */ \
    build_link_table_kernel< S ><<< n_classes, kNbrIds, 0, stream >>>( d_classes, link );            \
    if( n_canon > 0 ) build_range_blocks_kernel< S ><<< n_canon, kNbrIds, 0, stream >>>( canon, canon + n_canon, ( uint32_t )n_classes + 1u, link ); \
    choose_class_blocks_kernel< S ><<< ( n_classes + 63 ) / 64, 64, 0, stream >>>( d_classes, n_classes, link, scratch ); \
    build_head_tables_kernel<<< kCellKeys / 128, 128, 0, stream >>>( desc, scratch, head, head + kCellKeys ); \
    return cudaGetLastError()
    PAR_FOR_SCALE( scale, PAR_BUILD_SMOOTH )
#undef PAR_BUILD_SMOOTH
    return cudaErrorInvalidValue;
}

void raster_tma_box( int scale, uint32_t box[ 3 ] )
{
    const int tw = 32;
    box[ 0 ] = ( uint32_t )( ( 16 + tw + 3 + 15 ) / 16 * 16 );
    box[ 1 ] = ( scale <= 4 ? 32 : 16 ) + 4;
    box[ 2 ] = 1;
}

void raster_img_tma_box( int scale, uint32_t box[ 3 ] )
{
    box[ 0 ] = 0;
    box[ 1 ] = ( scale <= 4 ? 32 : 16 ) + 4;
    box[ 2 ] = 1;
#define PAR_RAWP( S )                          \
    box[ 0 ] = ( uint32_t )Cfg< S >::RAWP;     \
    return
    PAR_FOR_SCALE( scale, PAR_RAWP )
#undef PAR_RAWP
}

cudaError_t launch_raster_rgba8( const RasterArgs& a, const CUtensorMap* graph_map, const CUtensorMap* img_map, cudaStream_t stream )
{
    return launch_raster_fmt< kFmtRgba8 >( a, graph_map, img_map, stream );
}

cudaError_t launch_raster( const RasterArgs& a, const CUtensorMap* graph_map, const CUtensorMap* img_map, cudaStream_t stream )
{
    switch( a.out_format ) // one translation unit per output format (they compile in parallel)
    {
        case PAR_OUT_RGBA8: return launch_raster_rgba8( a, graph_map, img_map, stream );
        case PAR_OUT_BGR8: return launch_raster_bgr8( a, graph_map, img_map, stream );
        case PAR_OUT_INDEX8: return launch_raster_index8( a, graph_map, img_map, stream );
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_polygons( const RasterArgs& a, cudaStream_t stream )
{
    dim3 grid( ( a.width + 31 ) / 32, ( a.height + 7 ) / 8, a.n_frames );
    polygon_kernel<<< grid, kThreads, 0, stream >>>( a );
    return cudaGetLastError();
}

} // namespace par
