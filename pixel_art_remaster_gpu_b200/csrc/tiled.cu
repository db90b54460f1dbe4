// One very large image over several GPUs (BASELINE config 4).
//
// New subsystem: the reference is single-GPU and allocates 1.5 kB per pixel per call
// (kernel.cu:313-388), so a 4096x4096 map needs ~25 GB on one device.  Here the image is cut into
// horizontal strips, one per device.  The exact dependency radius of the path is 37 source rows
// (34 for the crossing-heuristic walks, +2 subdivision, +1 raster gather) — a
// 1-pixel halo is NOT exact — so every strip is extended by an apron of kApron rows on each interior
// side.  Each device receives only its OWN rows from the host; the apron rows are then pulled from
// the neighbouring devices' buffers over NVLink (cudaMemcpyPeerAsync, peer access enabled where the
// topology allows).  Every device then runs the ordinary single-GPU path on its extended strip and
// returns its own rows.  Connected components are labelled per extended strip (labels are global
// pixel indices) and stitched exactly: on the last own row of every strip both neighbours have
// labelled the same pixels, which yields label equivalences; a tiny host union-find keeps the
// minimum of every class and a relabel kernel applies the map.  Because the label of a component is
// its minimum index, the result equals the single-GPU labelling bit for bit.
#include "../../include/pixelart_b200.h"
#include "kernels.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

namespace {

constexpr int kApron = 40; // >= 37, multiple of 8

std::string g_group_create_error;

struct Strip
{
    int device = 0;
    int own_b = 0, own_e = 0, load_b = 0, load_e = 0; // image rows
    par_context* ctx = nullptr;
    cudaStream_t stream = nullptr, copy_stream = nullptr; // copy_stream: image / graph rows back to the host while the labels are stitched
    uint8_t *d_in = nullptr, *d_rgba = nullptr, *d_graph = nullptr, *d_aux = nullptr;
    int32_t* d_labels = nullptr;
    int32_t *d_map_keys = nullptr, *d_map_vals = nullptr;
    size_t map_cap = 0;
    cudaEvent_t uploaded = nullptr, computed = nullptr;
};

__global__ void offset_labels_kernel( int32_t* lab, size_t n, int32_t offset )
{
    size_t t = ( size_t )blockIdx.x * blockDim.x + threadIdx.x;
    if( t < n ) lab[ t ] += offset;
}

// labels that appear in `keys` (sorted) are replaced by the class minimum in `vals`
__global__ void relabel_kernel( int32_t* lab, size_t n, const int32_t* __restrict__ keys, const int32_t* __restrict__ vals, int m )
{
    size_t t = ( size_t )blockIdx.x * blockDim.x + threadIdx.x;
    if( t >= n ) return;
    const int32_t v = lab[ t ];
    int lo = 0, hi = m - 1;
    while( lo <= hi )
    {
        int mid = ( lo + hi ) >> 1;
        int32_t k = __ldg( keys + mid );
        if( k == v )
        {
            lab[ t ] = __ldg( vals + mid );
            return;
        }
        if( k < v )
            lo = mid + 1;
        else
            hi = mid - 1;
    }
}

} // namespace

struct par_group
{
    int width = 0, height = 0, scale = 0;
    std::vector< Strip > strips;
    std::string error;
    int fail( int st, const char* fmt, ... )
    {
        char buf[ 512 ];
        va_list ap;
        va_start( ap, fmt );
        vsnprintf( buf, sizeof( buf ), fmt, ap );
        va_end( ap );
        error = buf;
        return st;
    }
};

extern "C" {

void par_group_destroy( par_group* g )
{
    if( !g ) return;
    for( auto& s : g->strips )
    {
        if( !s.ctx ) continue; // never created (par_group_create failed part-way): nothing on that device
        cudaSetDevice( s.device );
        if( s.stream ) cudaStreamSynchronize( s.stream );
        cudaFree( s.d_in );
        cudaFree( s.d_rgba );
        cudaFree( s.d_graph );
        cudaFree( s.d_aux );
        cudaFree( s.d_labels );
        cudaFree( s.d_map_keys );
        cudaFree( s.d_map_vals );
        if( s.uploaded ) cudaEventDestroy( s.uploaded );
        if( s.computed ) cudaEventDestroy( s.computed );
        if( s.copy_stream ) cudaStreamDestroy( s.copy_stream );
        if( s.ctx ) par_destroy( s.ctx );
        if( s.stream ) cudaStreamDestroy( s.stream );
    }
    delete g;
}

const char* par_group_last_error( const par_group* g ) { return g ? g->error.c_str() : g_group_create_error.c_str(); }

int par_group_create( par_group** out, const int* devices, int n_devices, int width, int height, int scale )
{
    if( !out || !devices || n_devices < 1 || width < 1 || height < 1 )
    {
        g_group_create_error = "par_group_create: bad argument";
        return PAR_ERR_INVALID;
    }
    *out = nullptr;
    if( !par::raster_scale_supported( scale ) )
    {
        g_group_create_error = "par_group_create: unsupported scale";
        return PAR_ERR_INVALID;
    }
    if( n_devices > 1 && height / n_devices < kApron )
    {
        g_group_create_error = "par_group_create: strips would be shorter than the 40-row apron; use fewer devices";
        return PAR_ERR_INVALID;
    }
    par_group* g = new par_group();
    g->width = width;
    g->height = height;
    g->scale = scale;
    g->strips.resize( n_devices );
    const int base = height / n_devices, extra = height % n_devices;
    for( int k = 0; k < n_devices; k++ )
    {
        Strip& s = g->strips[ k ];
        s.device = devices[ k ];
        s.own_b = k * base + std::min( k, extra );
        s.own_e = s.own_b + base + ( k < extra ? 1 : 0 );
        s.load_b = std::max( 0, s.own_b - kApron );
        s.load_e = std::min( height, s.own_e + kApron );
        const int rows = s.load_e - s.load_b;
        int st = par_create( &s.ctx, s.device, width, rows, 1 );
        if( st != PAR_OK )
        {
            g_group_create_error = std::string( "par_group_create: " ) + par_last_error( nullptr );
            par_group_destroy( g );
            return st;
        }
        cudaSetDevice( s.device );
        cudaError_t e = cudaStreamCreateWithFlags( &s.stream, cudaStreamNonBlocking );
        par_set_stream( s.ctx, s.stream );
        const size_t px = ( size_t )width * rows;
        if( e == cudaSuccess ) e = cudaMalloc( &s.d_in, px * 3 + 64 );
        if( e == cudaSuccess ) e = cudaMalloc( &s.d_rgba, px * scale * scale * 4 );
        if( e == cudaSuccess ) e = cudaMalloc( &s.d_graph, px );
        if( e == cudaSuccess ) e = cudaMalloc( &s.d_aux, px );
        if( e == cudaSuccess ) e = cudaMalloc( &s.d_labels, px * 4 );
        if( e == cudaSuccess ) e = cudaEventCreateWithFlags( &s.uploaded, cudaEventDisableTiming );
        if( e == cudaSuccess ) e = cudaEventCreateWithFlags( &s.computed, cudaEventDisableTiming );
        if( e == cudaSuccess ) e = cudaStreamCreateWithFlags( &s.copy_stream, cudaStreamNonBlocking );
        if( e != cudaSuccess )
        {
            g_group_create_error = std::string( "par_group_create: " ) + cudaGetErrorString( e );
            par_group_destroy( g );
            return PAR_ERR_CUDA;
        }
    }
    // direct NVLink/PCIe peer access between neighbouring strips where the topology allows it
    for( int k = 0; k + 1 < n_devices; k++ )
    {
        const int a = g->strips[ k ].device, b = g->strips[ k + 1 ].device;
        if( a == b ) continue;
        int ok = 0;
        if( cudaDeviceCanAccessPeer( &ok, a, b ) == cudaSuccess && ok )
        {
            cudaSetDevice( a );
            if( cudaDeviceEnablePeerAccess( b, 0 ) != cudaSuccess ) cudaGetLastError();
        }
        if( cudaDeviceCanAccessPeer( &ok, b, a ) == cudaSuccess && ok )
        {
            cudaSetDevice( b );
            if( cudaDeviceEnablePeerAccess( a, 0 ) != cudaSuccess ) cudaGetLastError();
        }
    }
    *out = g;
    return PAR_OK;
}

int par_group_remaster_host( par_group* g, const par_job* j )
{
    if( !g ) return PAR_ERR_INVALID;
    if( !j || !j->bgr ) return g->fail( PAR_ERR_INVALID, "job / bgr is NULL" );
    if( j->width != g->width || j->height != g->height || j->scale != g->scale || j->n_frames != 1 )
        return g->fail( PAR_ERR_INVALID, "job does not match the group (%dx%d s=%d, one frame)", g->width, g->height, g->scale );
    if( j->widthstep < 3 * j->width ) return g->fail( PAR_ERR_INVALID, "widthstep < 3*width" );
    if( j->polygons ) return g->fail( PAR_ERR_INVALID, "polygon export is not offered in tiled mode" );
    const int W = g->width, S = g->scale;
    const size_t row_in = ( size_t )3 * W; // strips are stored densely on the devices
    cudaError_t e = cudaSuccess;

    // (1) every device receives its OWN rows from the host
    for( auto& s : g->strips )
    {
        cudaSetDevice( s.device );
        uint8_t* dst = s.d_in + ( size_t )( s.own_b - s.load_b ) * row_in;
        e = cudaMemcpy2DAsync( dst, row_in, j->bgr + ( size_t )s.own_b * j->widthstep, j->widthstep, row_in, s.own_e - s.own_b,
                               cudaMemcpyHostToDevice, s.stream );
        if( e != cudaSuccess ) return g->fail( PAR_ERR_CUDA, "H2D: %s", cudaGetErrorString( e ) );
        cudaEventRecord( s.uploaded, s.stream );
    }
    // (2) apron rows come from the neighbouring devices (peer copies over NVLink), ordered after their upload
    for( size_t k = 0; k < g->strips.size(); k++ )
    {
        Strip& s = g->strips[ k ];
        cudaSetDevice( s.device );
        if( k > 0 && s.load_b < s.own_b )
        {
            Strip& n = g->strips[ k - 1 ];
            cudaStreamWaitEvent( s.stream, n.uploaded, 0 );
            const size_t bytes = ( size_t )( s.own_b - s.load_b ) * row_in;
            e = cudaMemcpyPeerAsync( s.d_in, s.device, n.d_in + ( size_t )( s.load_b - n.load_b ) * row_in, n.device, bytes, s.stream );
            if( e != cudaSuccess ) return g->fail( PAR_ERR_CUDA, "peer copy: %s", cudaGetErrorString( e ) );
        }
        if( k + 1 < g->strips.size() && s.load_e > s.own_e )
        {
            Strip& n = g->strips[ k + 1 ];
            cudaStreamWaitEvent( s.stream, n.uploaded, 0 );
            const size_t bytes = ( size_t )( s.load_e - s.own_e ) * row_in;
            e = cudaMemcpyPeerAsync( s.d_in + ( size_t )( s.own_e - s.load_b ) * row_in, s.device,
                                     n.d_in + ( size_t )( s.own_e - n.load_b ) * row_in, n.device, bytes, s.stream );
            if( e != cudaSuccess ) return g->fail( PAR_ERR_CUDA, "peer copy: %s", cudaGetErrorString( e ) );
        }
    }
    // (3) the ordinary path on every extended strip
    for( auto& s : g->strips )
    {
        cudaSetDevice( s.device );
        par_job d = *j;
        d.bgr = s.d_in;
        d.height = s.load_e - s.load_b;
        d.widthstep = 3 * W;
        d.frame_stride = 0;
        d.rgba = j->rgba ? s.d_rgba : nullptr;
        d.graph = s.d_graph;
        d.graph_aux = s.d_aux;
        d.labels = nullptr;
        d.polygons = nullptr;
        d.poly_count = nullptr;
        int st = par_remaster_device( s.ctx, &d );
        if( st != PAR_OK ) return g->fail( st, "strip on device %d: %s", s.device, par_last_error( s.ctx ) );
        cudaEventRecord( s.computed, s.stream );
        if( j->labels )
        {
            // Label only rows whose graph bytes are exact: the strip's own rows plus ONE row either side (needed
            // for the stitch).  The outer apron rows carry graph bytes computed without their full context and
            // must not be allowed to connect anything.  The window is labelled as an image of its own (links
            // leaving it are ignored) and then shifted to global pixel indices.
            const int win_b = std::max( s.load_b, s.own_b - 1 ), win_e = std::min( s.load_e, s.own_e + 1 );
            par_job l = d;
            l.height = win_e - win_b;
            l.graph = s.d_graph + ( size_t )( win_b - s.load_b ) * W;
            l.labels = s.d_labels + ( size_t )( win_b - s.load_b ) * W;
            st = par_stage_cc_labels( s.ctx, &l );
            if( st != PAR_OK ) return g->fail( st, "labels on device %d: %s", s.device, par_last_error( s.ctx ) );
            const size_t n = ( size_t )W * l.height;
            offset_labels_kernel<<< ( unsigned )( ( n + 255 ) / 256 ), 256, 0, s.stream >>>( l.labels, n, win_b * W );
        }
    }
    // (4) own rows of the image and the graphs go back to the host as soon as the strip has them — the big copies
    // overlap the label stitching below (the labels follow once they are final)
    const size_t out_row = ( size_t )W * S * 4;
    for( auto& s : g->strips )
    {
        cudaSetDevice( s.device );
        const size_t own_px = ( size_t )W * ( s.own_e - s.own_b ), off_px = ( size_t )W * ( s.own_b - s.load_b );
        cudaStream_t cs = s.copy_stream;
        cudaStreamWaitEvent( cs, s.computed, 0 );
        if( j->rgba )
        {
            const bool flip = ( j->flags & PAR_FLAG_FLIP_OUTPUT ) != 0;
            const size_t src_row = flip ? ( size_t )( s.load_e - s.own_e ) * S : ( size_t )( s.own_b - s.load_b ) * S;
            const size_t dst_row = flip ? ( size_t )( g->height - s.own_e ) * S : ( size_t )s.own_b * S;
            e = cudaMemcpyAsync( j->rgba + dst_row * out_row, s.d_rgba + src_row * out_row, ( size_t )( s.own_e - s.own_b ) * S * out_row,
                                 cudaMemcpyDeviceToHost, cs );
        }
        if( e == cudaSuccess && j->graph ) e = cudaMemcpyAsync( j->graph + ( size_t )W * s.own_b, s.d_graph + off_px, own_px, cudaMemcpyDeviceToHost, cs );
        if( e == cudaSuccess && j->graph_aux )
            e = cudaMemcpyAsync( j->graph_aux + ( size_t )W * s.own_b, s.d_aux + off_px, own_px, cudaMemcpyDeviceToHost, cs );
        if( e != cudaSuccess ) return g->fail( PAR_ERR_CUDA, "D2H: %s", cudaGetErrorString( e ) );
    }
    // (5) stitch the labels: on the last own row of strip k both k and k+1 labelled the same pixels.  All seam rows
    // are fetched at once; the equivalences (<= 7 x 4096 pairs) are closed by a union-find over the distinct labels.
    std::vector< int32_t > keys, vals; // relabelling map, alive until the final synchronize
    if( j->labels && g->strips.size() > 1 )
    {
        const size_t n_seams = g->strips.size() - 1;
        std::vector< int32_t > rows( 2 * n_seams * W );
        for( size_t k = 0; k < n_seams; k++ )
        {
            Strip &a = g->strips[ k ], &b = g->strips[ k + 1 ];
            const int row = a.own_e - 1;
            cudaSetDevice( a.device );
            cudaMemcpyAsync( rows.data() + ( 2 * k ) * W, a.d_labels + ( size_t )( row - a.load_b ) * W, ( size_t )W * 4, cudaMemcpyDeviceToHost, a.stream );
            cudaSetDevice( b.device );
            cudaMemcpyAsync( rows.data() + ( 2 * k + 1 ) * W, b.d_labels + ( size_t )( row - b.load_b ) * W, ( size_t )W * 4, cudaMemcpyDeviceToHost, b.stream );
        }
        for( auto& s : g->strips )
        {
            cudaSetDevice( s.device );
            e = cudaStreamSynchronize( s.stream );
            if( e != cudaSuccess ) return g->fail( PAR_ERR_CUDA, "label rows: %s", cudaGetErrorString( e ) );
        }
        std::vector< int32_t > ids( rows ); // the distinct labels on the seams, sorted: position = union-find node
        std::sort( ids.begin(), ids.end() );
        ids.erase( std::unique( ids.begin(), ids.end() ), ids.end() );
        std::vector< int32_t > parent( ids.size() );
        for( size_t k = 0; k < parent.size(); k++ ) parent[ k ] = ( int32_t )k;
        auto node = [ & ]( int32_t label ) { return ( int32_t )( std::lower_bound( ids.begin(), ids.end(), label ) - ids.begin() ); };
        auto find = [ & ]( int32_t x ) {
            while( parent[ x ] != x ) x = parent[ x ] = parent[ parent[ x ] ];
            return x;
        };
        for( size_t k = 0; k < n_seams; k++ )
            for( int x = 0; x < W; x++ )
            {
                const int32_t p = find( node( rows[ ( 2 * k ) * W + x ] ) ), q = find( node( rows[ ( 2 * k + 1 ) * W + x ] ) );
                if( p != q ) parent[ std::max( p, q ) ] = std::min( p, q ); // ids are sorted: the smaller node is the smaller label
            }
        for( size_t k = 0; k < ids.size(); k++ )
        {
            const int32_t r = find( ( int32_t )k );
            if( r != ( int32_t )k )
            {
                keys.push_back( ids[ k ] ); // ascending: the relabel kernel searches them
                vals.push_back( ids[ r ] );
            }
        }
        if( !keys.empty() )
            for( auto& s : g->strips )
            {
                cudaSetDevice( s.device );
                if( keys.size() > s.map_cap )
                {
                    cudaFree( s.d_map_keys );
                    cudaFree( s.d_map_vals );
                    s.map_cap = keys.size() * 2;
                    cudaMalloc( &s.d_map_keys, s.map_cap * 4 );
                    cudaMalloc( &s.d_map_vals, s.map_cap * 4 );
                }
                cudaMemcpyAsync( s.d_map_keys, keys.data(), keys.size() * 4, cudaMemcpyHostToDevice, s.stream );
                cudaMemcpyAsync( s.d_map_vals, vals.data(), vals.size() * 4, cudaMemcpyHostToDevice, s.stream );
                const size_t n = ( size_t )W * ( s.own_e - s.own_b );
                relabel_kernel<<< ( unsigned )( ( n + 255 ) / 256 ), 256, 0, s.stream >>>(
                    s.d_labels + ( size_t )( s.own_b - s.load_b ) * W, n, s.d_map_keys, s.d_map_vals, ( int )keys.size() );
            }
    }
    if( j->labels )
        for( auto& s : g->strips )
        {
            cudaSetDevice( s.device );
            const size_t own_px = ( size_t )W * ( s.own_e - s.own_b ), off_px = ( size_t )W * ( s.own_b - s.load_b );
            e = cudaMemcpyAsync( j->labels + ( size_t )W * s.own_b, s.d_labels + off_px, own_px * 4, cudaMemcpyDeviceToHost, s.stream );
            if( e != cudaSuccess ) return g->fail( PAR_ERR_CUDA, "D2H: %s", cudaGetErrorString( e ) );
        }
    for( auto& s : g->strips )
    {
        cudaSetDevice( s.device );
        e = cudaStreamSynchronize( s.copy_stream );
        if( e == cudaSuccess ) e = cudaStreamSynchronize( s.stream );
        if( e != cudaSuccess ) return g->fail( PAR_ERR_CUDA, "strip on device %d: %s", s.device, cudaGetErrorString( e ) );
    }
    for( auto& s : g->strips )
    {
        cudaSetDevice( s.device );
        e = cudaStreamSynchronize( s.stream );
        if( e != cudaSuccess ) return g->fail( PAR_ERR_CUDA, "strip on device %d: %s", s.device, cudaGetErrorString( e ) );
    }
    return PAR_OK;
}

} // extern "C"
