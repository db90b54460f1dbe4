// C ABI of the remaster path (include/pixelart_b200.h): context, pooled buffers, stage dispatch.
//
// Replaces the host orchestration in launch_kernel (kernel.cu:286-526): instead of 10 cudaMalloc /
// cudaFree and 3 blocking D2H copies per frame, a context owns its scratch for a whole batch of
// frames, every stage is one launch over the batch (blockIdx.z = frame), and work is asynchronous
// on one stream.
#include "../../include/pixelart_b200.h"
#include "cell_table.h"
#include "kernels.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

using namespace par;

namespace {

std::string g_create_error;

// every entry point runs on its context's device and leaves the caller's current device as it found it
struct DeviceGuard
{
    int prev = -1;
    explicit DeviceGuard( int device )
    {
        if( cudaGetDevice( &prev ) != cudaSuccess ) prev = -1;
        if( prev != device ) cudaSetDevice( device );
        else prev = -1;
    }
    ~DeviceGuard()
    {
        if( prev >= 0 ) cudaSetDevice( prev );
    }
};

constexpr int kMaxFramesPerLaunch = 65535; // grid.z

typedef CUresult ( *EncodeTiledFn )( CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill );

EncodeTiledFn load_encode_tiled()
{
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if( cudaGetDriverEntryPoint( "cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q ) != cudaSuccess || q != cudaDriverEntryPointSuccess )
        return nullptr;
    return reinterpret_cast< EncodeTiledFn >( fn );
}

} // namespace

struct par_context
{
    int device = 0, n_sms = 0;
    int max_w = 0, max_h = 0, max_frames = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr; // copy_stream: D2H of par_remaster_host
    int sub_batch = 0; // frames per round of stage launches (0 = the whole batch per launch), par_set_sub_batch
    uint8_t *scratch_aux = nullptr, *scratch_graph = nullptr; // max_frames * max_w * max_h each
    CellRecord* d_tables = nullptr;        // kCellKeys 32-byte records
    uint32_t* d_mask_lut[ 9 ] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr }; // per scale, built on first use
    // smoothing tables (smooth_table.h): link descriptors + neighbour bytes + class list (scale-independent), CUT / LINK masks per scale
    uint32_t* d_smooth_words = nullptr; // desc (kCellKeys x 4 words) | pack (kCellKeys x 2 words) | 256 bytes of scratch
    LinkClass* d_link_classes = nullptr;
    int n_link_classes = 0;
    uint32_t link_entries = 0;
    uint64_t* d_cut[ 9 ] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
    uint64_t* d_link[ 9 ] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
    uint2* d_head[ 9 ] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
    int n_canon = 0;
    unsigned long long* d_smooth_stats = nullptr; // 2 counters
    CellTablePtrs tables() const { return CellTablePtrs{ d_tables }; }
    EncodeTiledFn encode = nullptr;
    uint64_t launches = 0;
    std::string error;
    // optional per-stage event timing (par_profile_enable)
    bool profiling = false;
    struct Span { int stage; cudaEvent_t a, b; };
    std::vector< Span > spans;
    std::vector< cudaEvent_t > free_events;
    cudaEvent_t get_event()
    {
        cudaEvent_t e = nullptr;
        if( !free_events.empty() )
        {
            e = free_events.back();
            free_events.pop_back();
        }
        else
            cudaEventCreate( &e );
        return e;
    }
    cudaEvent_t span_begin()
    {
        if( !profiling ) return nullptr;
        cudaEvent_t e = get_event();
        cudaEventRecord( e, stream );
        return e;
    }
    void span_end( int stage, cudaEvent_t a )
    {
        if( !a ) return;
        cudaEvent_t b = get_event();
        cudaEventRecord( b, stream );
        spans.push_back( Span{ stage, a, b } );
    }
    // staging for par_remaster_host (grown on demand)
    uint8_t* h_stage[ 10 ] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
    size_t h_stage_bytes[ 10 ] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    // palette scratch of par_remaster_device (INDEX8): lookup tables and counts, grown on demand
    uint32_t* d_pal_lut = nullptr;
    int32_t* d_pal_count = nullptr;
    int pal_frames = 0;

    int fail( par_status st, const char* fmt, ... )
    {
        char buf[ 512 ];
        va_list ap;
        va_start( ap, fmt );
        vsnprintf( buf, sizeof( buf ), fmt, ap );
        va_end( ap );
        error = buf;
        return st;
    }
    int cuda_fail( cudaError_t e, const char* what ) { return fail( PAR_ERR_CUDA, "%s: %s", what, cudaGetErrorString( e ) ); }

    // 3-D byte tensor (row bytes, rows, frames) -> tensor map; false when TMA's alignment rules do not hold
    bool make_map( CUtensorMap* map, const void* base, uint64_t row_bytes, uint64_t rows, uint64_t frames, uint64_t row_stride,
                   uint64_t frame_stride, const uint32_t box[ 3 ] )
    {
        if( !encode ) return false;
        if( ( reinterpret_cast< uintptr_t >( base ) & 15u ) || ( row_stride & 15u ) || ( frame_stride & 15u ) ) return false;
        if( box[ 0 ] > 256 || box[ 1 ] > 256 ) return false;
        cuuint64_t dims[ 3 ] = { row_bytes, rows, frames };
        cuuint64_t strides[ 2 ] = { row_stride, frame_stride };
        cuuint32_t bx[ 3 ] = { box[ 0 ], box[ 1 ], box[ 2 ] };
        cuuint32_t es[ 3 ] = { 1, 1, 1 };
        CUresult r = encode( map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast< void* >( base ), dims, strides, bx, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE );
        return r == CUDA_SUCCESS;
    }
};

namespace {

int check_job( par_context* c, const par_job* j, bool need_bgr )
{
    if( !c ) return PAR_ERR_INVALID;
    if( !j ) return c->fail( PAR_ERR_INVALID, "job is NULL" );
    if( j->width <= 0 || j->height <= 0 || j->n_frames <= 0 ) return c->fail( PAR_ERR_INVALID, "empty frame or batch (%dx%d x %d)", j->width, j->height, j->n_frames );
    if( need_bgr )
    {
        if( !j->bgr ) return c->fail( PAR_ERR_INVALID, "bgr is NULL" );
        if( j->widthstep < 3 * j->width ) return c->fail( PAR_ERR_INVALID, "widthstep %d < 3*width", j->widthstep );
        if( j->frame_stride != 0 && j->frame_stride < ( size_t )j->widthstep * j->height ) return c->fail( PAR_ERR_INVALID, "frame_stride too small" );
    }
    if( ( size_t )j->width * j->height > ( size_t )1 << 30 ) return c->fail( PAR_ERR_INVALID, "frame too large" );
    if( j->out_format < PAR_OUT_RGBA8 || j->out_format > PAR_OUT_INDEX8 ) return c->fail( PAR_ERR_INVALID, "unknown out_format %d", j->out_format );
    return PAR_OK;
}

int bytes_per_pixel( int out_format ) { return out_format == PAR_OUT_RGBA8 ? 4 : ( out_format == PAR_OUT_BGR8 ? 3 : 1 ); }

// what a raster request needs, checked before anything is launched
int check_raster( par_context* c, const par_job* j, bool device_pointers )
{
    const int aa = ( j->flags & PAR_FLAG_AA4 ) ? 4 : ( ( j->flags & PAR_FLAG_AA2 ) ? 2 : 1 );
    if( !raster_aa_supported( j->scale, aa ) )
        return c->fail( PAR_ERR_INVALID, "unsupported scale %d with %dx%d samples per pixel (scale x samples must be an integer 1..8)", j->scale, aa, aa );
    if( j->out_format == PAR_OUT_INDEX8 && aa != 1 ) return c->fail( PAR_ERR_INVALID, "PAR_OUT_INDEX8 cannot hold anti-aliased output (averaged samples are not palette colours)" );
    if( device_pointers && ( reinterpret_cast< uintptr_t >( j->rgba ) & 15u ) ) return c->fail( PAR_ERR_INVALID, "the output image must be 16-byte aligned" );
    return PAR_OK;
}

size_t frame_stride_of( const par_job* j ) { return j->frame_stride ? j->frame_stride : ( size_t )j->widthstep * j->height; }

// frames [f0, f0 + n) of a job as a job of its own (frames are independent units)
par_job slice_job( const par_job& j, int f0, int n )
{
    par_job d = j;
    const size_t fpx = ( size_t )j.width * j.height, fin = frame_stride_of( &j );
    const size_t fout = fpx * j.scale * j.scale * bytes_per_pixel( j.out_format );
    d.n_frames = n;
    d.frame_stride = fin;
    if( j.bgr ) d.bgr = j.bgr + ( size_t )f0 * fin;
    if( j.rgba ) d.rgba = j.rgba + ( size_t )f0 * fout;
    if( j.graph ) d.graph = j.graph + ( size_t )f0 * fpx;
    if( j.graph_aux ) d.graph_aux = j.graph_aux + ( size_t )f0 * fpx;
    if( j.labels ) d.labels = j.labels + ( size_t )f0 * fpx;
    if( j.polygons ) d.polygons = j.polygons + ( size_t )f0 * fpx * PAR_CELL_SLOTS * 2;
    if( j.poly_count ) d.poly_count = j.poly_count + ( size_t )f0 * fpx;
    if( j.palette ) d.palette = j.palette + ( size_t )f0 * 256;
    if( j.palette_count ) d.palette_count = j.palette_count + f0;
    return d;
}

// a stage over a batch of any size: one launch per kMaxFramesPerLaunch frames
template< class Run >
int for_launch_chunks( const par_job* j, Run run )
{
    for( int f0 = 0; f0 < j->n_frames; f0 += kMaxFramesPerLaunch )
    {
        const par_job d = slice_job( *j, f0, j->n_frames - f0 < kMaxFramesPerLaunch ? j->n_frames - f0 : kMaxFramesPerLaunch );
        const int st = run( &d );
        if( st ) return st;
    }
    return PAR_OK;
}

int check_capacity( par_context* c, const par_job* j )
{
    if( ( size_t )j->width * j->height * j->n_frames > ( size_t )c->max_w * c->max_h * c->max_frames )
        return c->fail( PAR_ERR_CAPACITY, "batch of %d %dx%d frames exceeds the context capacity (%d frames of %dx%d)", j->n_frames, j->width,
                        j->height, c->max_frames, c->max_w, c->max_h );
    return PAR_OK;
}

int run_similarity( par_context* c, const par_job* j, uint8_t* aux )
{
    GraphArgs a;
    a.bgr = j->bgr;
    a.graph_aux = aux;
    a.width = j->width;
    a.height = j->height;
    a.widthstep = j->widthstep;
    a.n_frames = j->n_frames;
    a.frame_stride = frame_stride_of( j );
    CUtensorMap map;
    uint32_t box[ 3 ];
    similarity_graph_tma_box( box );
    bool tma = !( j->flags & PAR_FLAG_NO_TMA ) &&
               c->make_map( &map, j->bgr, 3ull * j->width, j->height, j->n_frames, j->widthstep, a.frame_stride, box );
    cudaEvent_t t0 = c->span_begin();
    cudaError_t e = launch_similarity_graph( a, tma ? &map : nullptr, c->stream );
    c->span_end( 0, t0 );
    c->launches++;
    return e == cudaSuccess ? PAR_OK : c->cuda_fail( e, "similarity_graph" );
}

bool graph_map( par_context* c, const par_job* j, const uint8_t* g, const uint32_t box[ 3 ], CUtensorMap* map )
{
    return !( j->flags & PAR_FLAG_NO_TMA ) &&
           c->make_map( map, g, j->width, j->height, j->n_frames, j->width, ( uint64_t )j->width * j->height, box );
}

int run_crossings( par_context* c, const par_job* j, const uint8_t* aux, uint8_t* graph )
{
    CrossArgs a;
    a.graph_aux = aux;
    a.graph = graph;
    a.width = j->width;
    a.height = j->height;
    a.n_frames = j->n_frames;
    CUtensorMap map;
    uint32_t box[ 3 ];
    resolve_crossings_tma_box( box );
    bool tma = graph_map( c, j, aux, box, &map );
    cudaEvent_t t0 = c->span_begin();
    cudaError_t e = launch_resolve_crossings( a, tma ? &map : nullptr, c->stream );
    c->span_end( 1, t0 );
    c->launches++;
    return e == cudaSuccess ? PAR_OK : c->cuda_fail( e, "resolve_crossings" );
}

int run_labels( par_context* c, const par_job* j, const uint8_t* graph, int32_t* labels )
{
    LabelArgs a;
    a.graph = graph;
    a.labels = labels;
    a.width = j->width;
    a.height = j->height;
    a.n_frames = j->n_frames;
    int n = 0;
    cudaEvent_t t0 = c->span_begin();
    cudaError_t e = launch_cc_labels( a, c->stream, &n );
    c->span_end( 2, t0 );
    c->launches += n;
    return e == cudaSuccess ? PAR_OK : c->cuda_fail( e, "cc_labels" );
}

RasterArgs raster_args( par_context* c, const par_job* j, const uint8_t* graph )
{
    RasterArgs a;
    a.bgr = j->bgr;
    a.graph = graph;
    a.tables = c->tables();
    a.mask_lut = nullptr;
    a.rgba = j->rgba;
    a.polygons = j->polygons;
    a.poly_count = j->poly_count;
    a.width = j->width;
    a.height = j->height;
    a.widthstep = j->widthstep;
    a.n_frames = j->n_frames;
    a.frame_stride = frame_stride_of( j );
    a.aa = ( j->flags & PAR_FLAG_AA4 ) ? 4 : ( ( j->flags & PAR_FLAG_AA2 ) ? 2 : 1 );
    a.scale = j->scale * a.aa; // the kernels' sampling scale
    a.subdivide = ( j->flags & PAR_FLAG_SUBDIVIDE ) ? 1 : 0;
    a.flip_output = ( j->flags & PAR_FLAG_FLIP_OUTPUT ) ? 1 : 0;
    a.debug_force_wide = ( j->flags & PAR_FLAG_DEBUG_WIDE ) ? 1 : 0;
    a.out_format = j->out_format;
    a.pal_lut = nullptr;
    a.pal_count = nullptr;
    return a;
}

int run_polygons( par_context* c, const par_job* j, const uint8_t* graph )
{
    RasterArgs a = raster_args( c, j, graph );
    cudaEvent_t t0 = c->span_begin();
    cudaError_t e = launch_polygons( a, c->stream );
    c->span_end( 3, t0 );
    c->launches++;
    return e == cudaSuccess ? PAR_OK : c->cuda_fail( e, "polygons" );
}

// colour -> palette index tables of the job's frames (INDEX8) into lut / count; the palette itself into j->palette
int run_palette( par_context* c, const par_job* j, uint32_t* lut, int32_t* count )
{
    PaletteArgs p;
    p.bgr = j->bgr;
    p.width = j->width;
    p.height = j->height;
    p.widthstep = j->widthstep;
    p.n_frames = j->n_frames;
    p.frame_stride = frame_stride_of( j );
    p.lut = lut;
    p.count = count;
    p.palette = j->palette;
    int n = 0;
    cudaEvent_t t0 = c->span_begin();
    cudaError_t e = launch_palette( p, c->stream, &n );
    c->span_end( 5, t0 );
    c->launches += n;
    return e == cudaSuccess ? PAR_OK : c->cuda_fail( e, "palette" );
}

int run_raster( par_context* c, const par_job* j, const uint8_t* graph, const uint32_t* pal_lut = nullptr, const int32_t* pal_count = nullptr )
{
    RasterArgs a = raster_args( c, j, graph );
    int st = check_raster( c, j, true );
    if( st ) return st;
    if( j->out_format == PAR_OUT_INDEX8 && ( !pal_lut || !pal_count ) ) return c->fail( PAR_ERR_INVALID, "PAR_OUT_INDEX8 needs the palette stage" );
    a.pal_lut = pal_lut;
    a.pal_count = pal_count;
    const int S = a.scale; // the sampling scale: tables and tile shapes depend on it
    bool built = false;
    if( !c->d_mask_lut[ S ] )
    {
        // coverage masks of the 4096 plain hulls at this scale, computed once by the device's own coverage code
        cudaError_t le = cudaMalloc( &c->d_mask_lut[ S ], mask_lut_words( S ) * sizeof( uint32_t ) );
        if( le == cudaSuccess ) le = launch_build_mask_lut( S, c->tables(), c->d_mask_lut[ S ], c->stream );
        c->launches++;
        if( le != cudaSuccess ) return c->cuda_fail( le, "mask table" );
        built = true;
    }
    a.mask_lut = c->d_mask_lut[ S ];
    // (generic descriptors: kCellKeys x 4 words, then the ID words: kCellKeys x 2)
    a.smooth = SmoothTablePtrs{ nullptr, nullptr, reinterpret_cast< const uint2* >( c->d_smooth_words + 4 * kCellKeys ), nullptr, nullptr };
    a.smooth_stats = c->d_smooth_stats;
    if( a.subdivide && !( j->flags & PAR_FLAG_NO_SMOOTH_TABLES ) )
    {
        if( !c->d_cut[ S ] )
        {
            // CUT / LINK masks at this scale, computed once by the device's own coverage code
            const size_t ew = smooth_entry_words( S ) * sizeof( uint64_t );
            cudaError_t me = cudaMalloc( &c->d_cut[ S ], ( size_t )kCellKeys * 16 * ew );
            if( me == cudaSuccess ) me = cudaMalloc( &c->d_link[ S ], ( size_t )c->link_entries * ew );
            if( me == cudaSuccess ) me = cudaMalloc( &c->d_head[ S ], 2 * ( size_t )kCellKeys * sizeof( uint2 ) );
            if( me == cudaSuccess )
            {
                uint8_t* bytes = reinterpret_cast< uint8_t* >( c->d_smooth_words + 6 * kCellKeys ); // 256 scratch, then the shared ranges
                me = launch_build_smooth_tables( S, c->tables(), c->d_link_classes, c->n_link_classes, reinterpret_cast< const uint4* >( c->d_smooth_words ),
                                                 bytes + 256, c->n_canon, c->d_cut[ S ], c->d_link[ S ], c->d_head[ S ], bytes, c->stream );
            }
            c->launches += 5;
            if( me != cudaSuccess )
            {
                cudaFree( c->d_cut[ S ] );
                c->d_cut[ S ] = nullptr;
                return c->cuda_fail( me, "smoothing tables" );
            }
            built = true;
        }
        a.smooth.cut = c->d_cut[ S ];
        a.smooth.link = c->d_link[ S ];
        a.smooth.head = c->d_head[ S ];
        a.smooth.head2 = c->d_head[ S ] + kCellKeys;
    }
    if( built )
    {
        // the tables were built on the stream that is current NOW; a later par_set_stream must find them complete
        cudaError_t se = cudaStreamSynchronize( c->stream );
        if( se != cudaSuccess ) return c->cuda_fail( se, "table build" );
    }
    CUtensorMap map;
    uint32_t box[ 3 ];
    raster_tma_box( S, box );
    bool tma = graph_map( c, j, graph, box, &map );
    CUtensorMap img_map;
    raster_img_tma_box( S, box );
    tma = tma && c->make_map( &img_map, j->bgr, 3ull * j->width, j->height, j->n_frames, j->widthstep, a.frame_stride, box );
    cudaEvent_t t0 = c->span_begin();
    cudaError_t e = launch_raster( a, tma ? &map : nullptr, tma ? &img_map : nullptr, c->stream );
    c->span_end( 4, t0 );
    c->launches++;
    return e == cudaSuccess ? PAR_OK : c->cuda_fail( e, "raster" );
}

} // namespace

extern "C" {

int par_create( par_context** out, int device, int max_width, int max_height, int max_frames )
{
    if( !out || max_width <= 0 || max_height <= 0 || max_frames <= 0 )
    {
        g_create_error = "par_create: bad argument";
        return PAR_ERR_INVALID;
    }
    *out = nullptr;
    int n_dev = 0;
    if( cudaGetDeviceCount( &n_dev ) != cudaSuccess || n_dev == 0 || device < 0 || device >= n_dev )
    {
        cudaGetLastError();
        g_create_error = "par_create: no CUDA device (this library has no CPU fallback)";
        return PAR_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties( &prop, device );
    if( prop.major != 10 )
    {
        g_create_error = std::string( "par_create: device '" ) + prop.name + "' is not sm_100 (kernels are built for sm_100a only)";
        return PAR_ERR_NO_DEVICE;
    }
    par_context* c = new par_context();
    c->device = device;
    c->n_sms = prop.multiProcessorCount;
    c->max_w = max_width;
    c->max_h = max_height;
    c->max_frames = max_frames;
    DeviceGuard guard( device );
    cudaError_t e = cudaSuccess;
    if( e == cudaSuccess ) e = cudaStreamCreateWithFlags( &c->own_stream, cudaStreamNonBlocking );
    c->stream = c->own_stream;
    size_t px = ( size_t )max_width * max_height * max_frames;
    if( e == cudaSuccess ) e = cudaMalloc( &c->scratch_aux, px );
    if( e == cudaSuccess ) e = cudaMalloc( &c->scratch_graph, px );
    if( e == cudaSuccess ) e = cudaMalloc( &c->d_tables, sizeof( CellTables ) );
    if( e == cudaSuccess ) e = cudaMalloc( &c->d_smooth_stats, 2 * sizeof( unsigned long long ) );
    if( e == cudaSuccess ) e = cudaMemset( c->d_smooth_stats, 0, 2 * sizeof( unsigned long long ) );
    if( e == cudaSuccess )
    {
        static CellTables tables;
        static SmoothTables smooth;
        static std::once_flag once;
        std::call_once( once, [] {
            build_cell_tables( &tables );
            build_smooth_tables( tables, &smooth );
        } );
        static_assert( sizeof( CellTables ) == 32 * kCellKeys, "one 32-byte record per key" );
        e = cudaMemcpy( c->d_tables, &tables, sizeof( tables ), cudaMemcpyHostToDevice );
        c->n_link_classes = ( int )smooth.classes.size();
        c->link_entries = smooth.link_entries;
        const size_t kw = ( size_t )kCellKeys * sizeof( uint32_t );
        if( e == cudaSuccess ) e = cudaMalloc( &c->d_smooth_words, 6 * kw + 256 + 128 );
        c->n_canon = ( int )smooth.n_canon;
        if( e == cudaSuccess ) e = cudaMalloc( &c->d_link_classes, smooth.classes.size() * sizeof( LinkClass ) );
        if( e == cudaSuccess ) e = cudaMemcpy( c->d_smooth_words, smooth.desc, 4 * kw, cudaMemcpyHostToDevice );
        if( e == cudaSuccess ) e = cudaMemcpy( c->d_smooth_words + 4 * kCellKeys, smooth.pack, 2 * kw, cudaMemcpyHostToDevice );
        if( e == cudaSuccess && smooth.n_canon )
        {
            uint8_t ranges[ 128 ];
            memcpy( ranges, smooth.canon_lo, smooth.n_canon );
            memcpy( ranges + smooth.n_canon, smooth.canon_span, smooth.n_canon );
            e = cudaMemcpy( reinterpret_cast< uint8_t* >( c->d_smooth_words + 6 * kCellKeys ) + 256, ranges, 2 * smooth.n_canon, cudaMemcpyHostToDevice );
        }
        if( e == cudaSuccess )
            e = cudaMemcpy( c->d_link_classes, smooth.classes.data(), smooth.classes.size() * sizeof( LinkClass ), cudaMemcpyHostToDevice );
    }
    if( e != cudaSuccess )
    {
        g_create_error = std::string( "par_create: " ) + cudaGetErrorString( e );
        par_destroy( c );
        return PAR_ERR_CUDA;
    }
    c->encode = load_encode_tiled();
    *out = c;
    return PAR_OK;
}

void par_destroy( par_context* c )
{
    if( !c ) return;
    DeviceGuard guard( c->device );
    if( c->own_stream )
    {
        cudaStreamSynchronize( c->own_stream );
        cudaStreamDestroy( c->own_stream );
    }
    if( c->copy_stream ) cudaStreamDestroy( c->copy_stream );
    cudaFree( c->scratch_aux );
    cudaFree( c->scratch_graph );
    cudaFree( c->d_tables );
    for( int k = 0; k < 9; k++ ) cudaFree( c->d_mask_lut[ k ] );
    for( int k = 0; k < 9; k++ ) cudaFree( c->d_cut[ k ] );
    for( int k = 0; k < 9; k++ ) cudaFree( c->d_link[ k ] );
    for( int k = 0; k < 9; k++ ) cudaFree( c->d_head[ k ] );
    cudaFree( c->d_smooth_words );
    cudaFree( c->d_link_classes );
    cudaFree( c->d_smooth_stats );
    cudaFree( c->d_pal_lut );
    cudaFree( c->d_pal_count );
    for( int k = 0; k < 10; k++ ) cudaFree( c->h_stage[ k ] );
    for( auto& sp : c->spans )
    {
        cudaEventDestroy( sp.a );
        cudaEventDestroy( sp.b );
    }
    for( auto e : c->free_events ) cudaEventDestroy( e );
    delete c;
}

const char* par_last_error( const par_context* c ) { return c ? c->error.c_str() : g_create_error.c_str(); }
int par_device( const par_context* c ) { return c ? c->device : -1; }
uint64_t par_launch_count( const par_context* c ) { return c ? c->launches : 0; }

int par_set_stream( par_context* c, void* cuda_stream )
{
    if( !c ) return PAR_ERR_INVALID;
    c->stream = static_cast< cudaStream_t >( cuda_stream );
    return PAR_OK;
}

int par_use_own_stream( par_context* c )
{
    if( !c ) return PAR_ERR_INVALID;
    c->stream = c->own_stream;
    return PAR_OK;
}

int par_set_sub_batch( par_context* c, int frames )
{
    if( !c || frames < 0 ) return PAR_ERR_INVALID;
    c->sub_batch = frames;
    return PAR_OK;
}

int par_profile_enable( par_context* c, int on )
{
    if( !c ) return PAR_ERR_INVALID;
    c->profiling = on != 0;
    return PAR_OK;
}

int par_profile_read( par_context* c, double* total_ms, int* launches )
{
    if( !c ) return PAR_ERR_INVALID;
    DeviceGuard guard( c->device );
    cudaError_t e = cudaStreamSynchronize( c->stream );
    if( e != cudaSuccess ) return c->cuda_fail( e, "profile_read" );
    for( int k = 0; k < PAR_N_STAGES; k++ )
    {
        if( total_ms ) total_ms[ k ] = 0.0;
        if( launches ) launches[ k ] = 0;
    }
    for( auto& sp : c->spans )
    {
        float ms = 0.f;
        cudaEventElapsedTime( &ms, sp.a, sp.b );
        if( total_ms ) total_ms[ sp.stage ] += ms;
        if( launches ) launches[ sp.stage ]++;
        c->free_events.push_back( sp.a );
        c->free_events.push_back( sp.b );
    }
    c->spans.clear();
    return PAR_OK;
}

int par_smooth_stats( par_context* c, uint64_t* out2 )
{
    if( !c || !out2 ) return PAR_ERR_INVALID;
    DeviceGuard guard( c->device );
    cudaError_t e = cudaStreamSynchronize( c->stream );
    unsigned long long h[ 2 ] = { 0, 0 };
    if( e == cudaSuccess ) e = cudaMemcpy( h, c->d_smooth_stats, sizeof( h ), cudaMemcpyDeviceToHost );
    if( e != cudaSuccess ) return c->cuda_fail( e, "smooth_stats" );
    for( int k = 0; k < 2; k++ ) out2[ k ] = h[ k ];
    return PAR_OK;
}

int par_synchronize( par_context* c )
{
    if( !c ) return PAR_ERR_INVALID;
    DeviceGuard guard( c->device );
    cudaError_t e = cudaStreamSynchronize( c->stream );
    return e == cudaSuccess ? PAR_OK : c->cuda_fail( e, "synchronize" );
}

int par_stage_similarity_graph( par_context* c, const par_job* j )
{
    int st = check_job( c, j, true );
    if( st ) return st;
    if( !j->graph_aux ) return c->fail( PAR_ERR_INVALID, "graph_aux is NULL" );
    DeviceGuard guard( c->device );
    return for_launch_chunks( j, [ & ]( const par_job* d ) { return run_similarity( c, d, d->graph_aux ); } );
}

int par_stage_resolve_crossings( par_context* c, const par_job* j )
{
    int st = check_job( c, j, false );
    if( st ) return st;
    if( !j->graph_aux || !j->graph ) return c->fail( PAR_ERR_INVALID, "graph_aux / graph is NULL" );
    DeviceGuard guard( c->device );
    return for_launch_chunks( j, [ & ]( const par_job* d ) { return run_crossings( c, d, d->graph_aux, d->graph ); } );
}

int par_stage_cc_labels( par_context* c, const par_job* j )
{
    int st = check_job( c, j, false );
    if( st ) return st;
    if( !j->graph || !j->labels ) return c->fail( PAR_ERR_INVALID, "graph / labels is NULL" );
    DeviceGuard guard( c->device );
    return for_launch_chunks( j, [ & ]( const par_job* d ) { return run_labels( c, d, d->graph, d->labels ); } );
}

int par_stage_polygons( par_context* c, const par_job* j )
{
    int st = check_job( c, j, true );
    if( st ) return st;
    if( !j->graph || !j->polygons ) return c->fail( PAR_ERR_INVALID, "graph / polygons is NULL" );
    DeviceGuard guard( c->device );
    return for_launch_chunks( j, [ & ]( const par_job* d ) { return run_polygons( c, d, d->graph ); } );
}

// palette scratch for n frames (INDEX8)
static int ensure_palette_scratch( par_context* c, int n_frames )
{
    if( n_frames <= c->pal_frames ) return PAR_OK;
    cudaFree( c->d_pal_lut );
    cudaFree( c->d_pal_count );
    c->d_pal_lut = nullptr;
    c->d_pal_count = nullptr;
    c->pal_frames = 0;
    cudaError_t e = cudaMalloc( &c->d_pal_lut, ( size_t )n_frames * kPaletteSlots * sizeof( uint32_t ) );
    if( e == cudaSuccess ) e = cudaMalloc( &c->d_pal_count, ( size_t )n_frames * sizeof( int32_t ) );
    if( e != cudaSuccess ) return c->cuda_fail( e, "palette scratch" );
    c->pal_frames = n_frames;
    return PAR_OK;
}

int par_stage_raster( par_context* c, const par_job* j )
{
    int st = check_job( c, j, true );
    if( st ) return st;
    if( !j->graph || !j->rgba ) return c->fail( PAR_ERR_INVALID, "graph / rgba is NULL" );
    if( ( st = check_raster( c, j, true ) ) ) return st;
    DeviceGuard guard( c->device );
    if( j->out_format == PAR_OUT_INDEX8 && ( st = ensure_palette_scratch( c, j->n_frames < kMaxFramesPerLaunch ? j->n_frames : kMaxFramesPerLaunch ) ) ) return st;
    return for_launch_chunks( j, [ & ]( const par_job* d ) {
        if( d->out_format != PAR_OUT_INDEX8 ) return run_raster( c, d, d->graph );
        int32_t* count = d->palette_count ? d->palette_count : c->d_pal_count;
        int s2 = run_palette( c, d, c->d_pal_lut, count );
        return s2 ? s2 : run_raster( c, d, d->graph, c->d_pal_lut, count );
    } );
}

int par_border_walks( par_context* c, const uint8_t* graph, const int32_t* labels, int width, int height, int n_frames, int32_t* walk_len,
                      int32_t* walk_begin, int32_t* walk_nodes, long long capacity_per_frame, long long* total )
{
    if( !c ) return PAR_ERR_INVALID;
    if( !graph || !labels || !walk_len || !walk_begin || !walk_nodes || !total ) return c->fail( PAR_ERR_INVALID, "border_walks: NULL pointer" );
    if( width <= 0 || height <= 0 || n_frames <= 0 || capacity_per_frame <= 0 ) return c->fail( PAR_ERR_INVALID, "border_walks: empty frame, batch or capacity" );
    if( ( size_t )width * height > ( size_t )1 << 30 ) return c->fail( PAR_ERR_INVALID, "frame too large" );
    DeviceGuard guard( c->device );
    cudaError_t e = launch_border_walks( graph, labels, width, height, n_frames, walk_len, walk_begin, total, walk_nodes, capacity_per_frame, c->stream );
    c->launches += 3 * ( ( n_frames + 65534 ) / 65535 );
    return e == cudaSuccess ? PAR_OK : c->cuda_fail( e, "border_walks" );
}

int par_walk_splines( par_context* c, const int32_t* walk_len, const int32_t* walk_begin, const int32_t* walk_nodes, const long long* total, int width,
                      int height, int n_frames, long long capacity_per_frame, int samples_per_segment, float* points )
{
    if( !c ) return PAR_ERR_INVALID;
    if( !walk_len || !walk_begin || !walk_nodes || !total || !points ) return c->fail( PAR_ERR_INVALID, "walk_splines: NULL pointer" );
    if( width <= 0 || height <= 0 || n_frames <= 0 || capacity_per_frame <= 0 ) return c->fail( PAR_ERR_INVALID, "walk_splines: empty frame, batch or capacity" );
    if( samples_per_segment != 1 && samples_per_segment != 2 && samples_per_segment != 4 && samples_per_segment != 8 )
        return c->fail( PAR_ERR_INVALID, "walk_splines: samples_per_segment must be 1, 2, 4 or 8 (exact float results)" );
    DeviceGuard guard( c->device );
    cudaError_t e = launch_walk_splines( walk_len, walk_begin, total, walk_nodes, width, height, n_frames, capacity_per_frame, samples_per_segment, points,
                                         c->stream );
    c->launches += ( n_frames + 65534 ) / 65535;
    return e == cudaSuccess ? PAR_OK : c->cuda_fail( e, "walk_splines" );
}

int par_outlines_host( par_context* c, const uint8_t* bgr, int width, int height, int widthstep, int samples, par_outlines* out )
{
    if( !c ) return PAR_ERR_INVALID;
    if( !bgr || !out || width <= 0 || height <= 0 || widthstep < 3 * width ) return c->fail( PAR_ERR_INVALID, "outlines: bad argument" );
    memset( out, 0, sizeof( *out ) );
    DeviceGuard guard( c->device );
    const size_t px = ( size_t )width * height, in_bytes = ( size_t )widthstep * height;
    const long long capacity = 4ll * ( long long )px; // (a walk visits a node at most once per link direction)
    // one device block: frame | graph_aux | graph | labels | walk_len | walk_begin | total | nodes | points
    uint8_t* d = nullptr;
    auto up16 = []( size_t v ) { return ( v + 15 ) & ~( size_t )15; };
    const size_t o_aux = up16( in_bytes + 16 ), o_graph = o_aux + up16( px ), o_lab = o_graph + up16( px ), o_len = o_lab + px * 4, o_beg = o_len + px * 4,
                 o_tot = o_beg + px * 4, o_nodes = o_tot + 16, o_pts = o_nodes + ( size_t )capacity * 4, bytes = o_pts + ( size_t )capacity * samples * 8;
    cudaError_t e = cudaMalloc( &d, bytes );
    if( e != cudaSuccess ) return c->cuda_fail( e, "outlines cudaMalloc" );
    auto fail = [ & ]( int st ) {
        cudaFree( d );
        par_outlines_free( out );
        return st;
    };
    e = cudaMemcpyAsync( d, bgr, in_bytes, cudaMemcpyHostToDevice, c->stream );
    if( e != cudaSuccess ) return fail( c->cuda_fail( e, "outlines H2D" ) );
    par_job j = {};
    j.bgr = d;
    j.width = width;
    j.height = height;
    j.widthstep = widthstep;
    j.n_frames = 1;
    j.scale = 1;
    j.graph_aux = d + o_aux;
    j.graph = d + o_graph;
    j.labels = reinterpret_cast< int32_t* >( d + o_lab );
    int32_t *d_len = reinterpret_cast< int32_t* >( d + o_len ), *d_beg = reinterpret_cast< int32_t* >( d + o_beg ), *d_nodes = reinterpret_cast< int32_t* >( d + o_nodes );
    long long* d_tot = reinterpret_cast< long long* >( d + o_tot );
    float* d_pts = reinterpret_cast< float* >( d + o_pts );
    int st = par_remaster_device( c, &j );
    if( st == PAR_OK ) st = par_border_walks( c, j.graph, j.labels, width, height, 1, d_len, d_beg, d_nodes, capacity, d_tot );
    if( st == PAR_OK ) st = par_walk_splines( c, d_len, d_beg, d_nodes, d_tot, width, height, 1, capacity, samples, d_pts );
    if( st != PAR_OK ) return fail( st );
    std::vector< int32_t > len( px ), beg( px );
    long long total = 0;
    e = cudaMemcpyAsync( len.data(), d_len, px * 4, cudaMemcpyDeviceToHost, c->stream );
    if( e == cudaSuccess ) e = cudaMemcpyAsync( beg.data(), d_beg, px * 4, cudaMemcpyDeviceToHost, c->stream );
    if( e == cudaSuccess ) e = cudaMemcpyAsync( &total, d_tot, sizeof( total ), cudaMemcpyDeviceToHost, c->stream );
    if( e == cudaSuccess ) e = cudaStreamSynchronize( c->stream );
    if( e != cudaSuccess ) return fail( c->cuda_fail( e, "outlines D2H" ) );
    if( total > capacity ) return fail( c->fail( PAR_ERR_CAPACITY, "outlines: the walks need %lld entries, capacity %lld", total, capacity ) );
    int n_walks = 0;
    for( size_t n = 0; n < px; n++ ) n_walks += len[ n ] > 0;
    out->n_walks = n_walks;
    out->samples = samples;
    out->n_points = total * samples;
    out->start = static_cast< int32_t* >( malloc( sizeof( int32_t ) * ( n_walks ? n_walks : 1 ) ) );
    out->count = static_cast< int32_t* >( malloc( sizeof( int32_t ) * ( n_walks ? n_walks : 1 ) ) );
    out->points = static_cast< float* >( malloc( sizeof( float ) * 2 * ( size_t )( out->n_points ? out->n_points : 1 ) ) );
    if( !out->start || !out->count || !out->points ) return fail( c->fail( PAR_ERR_INVALID, "outlines: out of host memory" ) );
    // walks are stored in raster order of their start, so the points of all walks are one contiguous range
    e = cudaMemcpy( out->points, d_pts, sizeof( float ) * 2 * ( size_t )out->n_points, cudaMemcpyDeviceToHost );
    if( e != cudaSuccess ) return fail( c->cuda_fail( e, "outlines D2H" ) );
    int k = 0;
    for( size_t n = 0; n < px; n++ )
        if( len[ n ] > 0 )
        {
            out->start[ k ] = ( int32_t )n;
            out->count[ k ] = len[ n ];
            k++;
        }
    cudaFree( d );
    return PAR_OK;
}

void par_outlines_free( par_outlines* out )
{
    if( !out ) return;
    free( out->start );
    free( out->count );
    free( out->points );
    memset( out, 0, sizeof( *out ) );
}

int par_remaster_device( par_context* c, const par_job* j )
{
    int st = check_job( c, j, true );
    if( st ) return st;
    DeviceGuard guard( c->device );
    par_job job = *j;
    if( !job.graph_aux || !job.graph )
    {
        st = check_capacity( c, j );
        if( st ) return st;
        if( !job.graph_aux ) job.graph_aux = c->scratch_aux;
        if( !job.graph ) job.graph = c->scratch_graph;
    }
    if( job.rgba && ( st = check_raster( c, j, true ) ) ) return st; // refuse an impossible raster request before anything is launched
    // Frames are independent, so the batch may run in rounds of `sub_batch` frames through all stages: the 2 bytes per
    // pixel of intermediates (graph_aux, graph) of a round are then still in L2 when the next stage reads them.
    int round = c->sub_batch > 0 ? c->sub_batch : job.n_frames;
    if( round > kMaxFramesPerLaunch ) round = kMaxFramesPerLaunch;
    const bool indexed = job.rgba && job.out_format == PAR_OUT_INDEX8;
    if( indexed && ( st = ensure_palette_scratch( c, round < job.n_frames ? round : job.n_frames ) ) ) return st;
    for( int f0 = 0; f0 < job.n_frames; f0 += round )
    {
        const par_job d = slice_job( job, f0, job.n_frames - f0 < round ? job.n_frames - f0 : round );
        if( ( st = run_similarity( c, &d, d.graph_aux ) ) ) return st;
        if( ( st = run_crossings( c, &d, d.graph_aux, d.graph ) ) ) return st;
        if( d.labels && ( st = run_labels( c, &d, d.graph, d.labels ) ) ) return st;
        if( d.polygons && ( st = run_polygons( c, &d, d.graph ) ) ) return st;
        if( d.rgba )
        {
            int32_t* count = d.palette_count ? d.palette_count : c->d_pal_count;
            if( indexed && ( st = run_palette( c, &d, c->d_pal_lut, count ) ) ) return st;
            if( ( st = run_raster( c, &d, d.graph, indexed ? c->d_pal_lut : nullptr, indexed ? count : nullptr ) ) ) return st;
        }
    }
    return PAR_OK;
}

int par_remaster_host( par_context* c, const par_job* j )
{
    int st = check_job( c, j, true );
    if( st ) return st;
    if( j->rgba && ( st = check_raster( c, j, false ) ) ) return st; // (before the staging buffers are sized from the scale)
    DeviceGuard guard( c->device );
    const size_t px = ( size_t )j->width * j->height * j->n_frames;
    const size_t in_bytes = frame_stride_of( j ) * j->n_frames;
    const size_t out_px = px * j->scale * j->scale;
    const int bpp = bytes_per_pixel( j->out_format );
    const bool indexed = j->rgba && j->out_format == PAR_OUT_INDEX8;
    // device staging: 0 bgr, 1 image, 2 graph, 3 graph_aux, 4 labels, 5 polygons, 6 poly_count, 7 palette, 8 palette_count
    const size_t need[ 9 ] = { in_bytes + 16,
                               j->rgba ? out_px * bpp : 0,
                               px,
                               px,
                               j->labels ? px * 4 : 0,
                               j->polygons ? px * PAR_CELL_SLOTS * 2 * sizeof( float ) : 0,
                               ( j->polygons && j->poly_count ) ? px * 4 : 0,
                               indexed ? ( size_t )j->n_frames * 256 * 4 : 0,
                               indexed ? ( size_t )j->n_frames * 4 : 0 };
    for( int k = 0; k < 9; k++ )
        if( need[ k ] > c->h_stage_bytes[ k ] )
        {
            cudaFree( c->h_stage[ k ] );
            c->h_stage[ k ] = nullptr;
            c->h_stage_bytes[ k ] = 0;
            cudaError_t e = cudaMalloc( &c->h_stage[ k ], need[ k ] );
            if( e != cudaSuccess ) return c->cuda_fail( e, "staging cudaMalloc" );
            c->h_stage_bytes[ k ] = need[ k ];
        }
    // The batch runs in chunks so that the copies overlap the kernels and each other: chunk k's results go back
    // on a second stream (PCIe is full duplex) while chunk k+1 is uploaded and computed on the context's stream.
    // Frames are independent, so chunking cannot change a result.
    if( !c->copy_stream )
    {
        cudaError_t se = cudaStreamCreateWithFlags( &c->copy_stream, cudaStreamNonBlocking );
        if( se != cudaSuccess ) return c->cuda_fail( se, "copy stream" );
    }
    const size_t fpx = ( size_t )j->width * j->height, fin = frame_stride_of( j ), fout = fpx * j->scale * j->scale * bpp;
    const int n_chunks = j->n_frames >= 512 ? 16 : ( j->n_frames >= 64 ? 8 : 1 );
    std::vector< cudaEvent_t > done;
    cudaError_t e = cudaSuccess;
    for( int k = 0; k < n_chunks && e == cudaSuccess; k++ )
    {
        const int f0 = ( int )( ( long long )j->n_frames * k / n_chunks ), f1 = ( int )( ( long long )j->n_frames * ( k + 1 ) / n_chunks );
        if( f1 == f0 ) continue;
        const size_t nf = ( size_t )( f1 - f0 );
        // (the last frame may be shorter than frame_stride in the caller's buffer: copy exactly what check_job validated)
        const size_t in_n = ( f1 == j->n_frames && j->frame_stride ) ? ( nf - 1 ) * fin + ( size_t )j->widthstep * j->height : nf * fin;
        e = cudaMemcpyAsync( c->h_stage[ 0 ] + f0 * fin, j->bgr + f0 * fin, in_n, cudaMemcpyHostToDevice, c->stream );
        if( e != cudaSuccess ) return c->cuda_fail( e, "H2D" );
        par_job d = *j;
        d.n_frames = ( int )nf;
        d.frame_stride = fin;
        d.bgr = c->h_stage[ 0 ] + f0 * fin;
        d.rgba = j->rgba ? c->h_stage[ 1 ] + f0 * fout : nullptr;
        d.graph = c->h_stage[ 2 ] + f0 * fpx;
        d.graph_aux = c->h_stage[ 3 ] + f0 * fpx;
        d.labels = j->labels ? reinterpret_cast< int32_t* >( c->h_stage[ 4 ] ) + f0 * fpx : nullptr;
        d.polygons = j->polygons ? reinterpret_cast< float* >( c->h_stage[ 5 ] ) + f0 * fpx * PAR_CELL_SLOTS * 2 : nullptr;
        d.poly_count = ( j->polygons && j->poly_count ) ? reinterpret_cast< int32_t* >( c->h_stage[ 6 ] ) + f0 * fpx : nullptr;
        d.palette = indexed ? reinterpret_cast< uint32_t* >( c->h_stage[ 7 ] ) + ( size_t )f0 * 256 : nullptr;
        d.palette_count = indexed ? reinterpret_cast< int32_t* >( c->h_stage[ 8 ] ) + f0 : nullptr;
        if( ( st = par_remaster_device( c, &d ) ) ) return st;
        cudaEvent_t ev = c->get_event();
        done.push_back( ev );
        e = cudaEventRecord( ev, c->stream );
        if( e == cudaSuccess ) e = cudaStreamWaitEvent( c->copy_stream, ev, 0 );
        cudaStream_t cs = c->copy_stream;
        if( e == cudaSuccess && j->rgba ) e = cudaMemcpyAsync( j->rgba + f0 * fout, d.rgba, nf * fout, cudaMemcpyDeviceToHost, cs );
        if( e == cudaSuccess && j->graph ) e = cudaMemcpyAsync( j->graph + f0 * fpx, d.graph, nf * fpx, cudaMemcpyDeviceToHost, cs );
        if( e == cudaSuccess && j->graph_aux ) e = cudaMemcpyAsync( j->graph_aux + f0 * fpx, d.graph_aux, nf * fpx, cudaMemcpyDeviceToHost, cs );
        if( e == cudaSuccess && j->labels ) e = cudaMemcpyAsync( j->labels + f0 * fpx, d.labels, nf * fpx * 4, cudaMemcpyDeviceToHost, cs );
        if( e == cudaSuccess && j->polygons )
            e = cudaMemcpyAsync( j->polygons + f0 * fpx * PAR_CELL_SLOTS * 2, d.polygons, nf * fpx * PAR_CELL_SLOTS * 2 * sizeof( float ),
                                 cudaMemcpyDeviceToHost, cs );
        if( e == cudaSuccess && d.poly_count ) e = cudaMemcpyAsync( j->poly_count + f0 * fpx, d.poly_count, nf * fpx * 4, cudaMemcpyDeviceToHost, cs );
        if( e == cudaSuccess && indexed && j->palette ) e = cudaMemcpyAsync( j->palette + ( size_t )f0 * 256, d.palette, nf * 256 * 4, cudaMemcpyDeviceToHost, cs );
        if( e == cudaSuccess && indexed && j->palette_count ) e = cudaMemcpyAsync( j->palette_count + f0, d.palette_count, nf * 4, cudaMemcpyDeviceToHost, cs );
    }
    cudaError_t e2 = cudaStreamSynchronize( c->copy_stream );
    cudaError_t e3 = cudaStreamSynchronize( c->stream );
    for( auto ev : done ) c->free_events.push_back( ev );
    if( e != cudaSuccess ) return c->cuda_fail( e, "D2H" );
    if( e2 != cudaSuccess ) return c->cuda_fail( e2, "synchronize" );
    return e3 == cudaSuccess ? PAR_OK : c->cuda_fail( e3, "synchronize" );
}

int par_cell_from_pattern( unsigned key, float* out_xy )
{
    if( key >= ( unsigned )kCellKeys || !out_xy ) return -1;
    static CellTables tables;
    static std::once_flag once;
    std::call_once( once, [] { build_cell_tables( &tables ); } );
    uint64_t h = tables.rec[ key ].verts;
    int n = hull_count( tables.rec[ key ].info );
    for( int t = 0; t <= n; t++ )
    {
        out_xy[ 2 * t ] = 0.25f * ( float )hull_xq( h, t % n );
        out_xy[ 2 * t + 1 ] = 0.25f * ( float )hull_yq( h, t % n );
    }
    return n;
}

uint32_t par_yuv_word( int b0, int b1, int b2 ) { return yuv_word( b0 & 255, b1 & 255, b2 & 255 ); }

} // extern "C"
