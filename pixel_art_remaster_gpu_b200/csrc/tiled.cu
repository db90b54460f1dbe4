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
// labelled the same pixels, which yields label equivalences; every device gathers the seam rows of all
// strips (peer copies), closes the equivalences with a one-CTA union-find that keeps the minimum of
// every class, and relabels its own rows — no host round trip.  Because the label of a component is
// its minimum index, the result equals the single-GPU labelling bit for bit.
// Two entries: host image in / host results out (par_group_remaster_host), and device-resident
// (par_group_remaster_device: the strips' own rows are already in their input buffers, results stay
// on the devices) — the halo exchange, the kernels and the stitch are the same stream-ordered code.
#include "../../include/pixelart_b200.h"
#include "kernels.cuh"

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
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
    int32_t* d_seams = nullptr;                    // [n_seams][2][W]: the label rows of every seam, gathered from all strips
    int32_t *d_tab_keys = nullptr, *d_tab_parent = nullptr; // the seam labels' union-find (open-addressing table)
    int32_t** d_seam_dst = nullptr;                // the d_seams of every strip (peer pointers), for push_seam_rows_kernel
    cudaEvent_t ready = nullptr, computed = nullptr, labelled = nullptr, t_begin = nullptr, t_end = nullptr;
};

__global__ void offset_labels_kernel( int32_t* lab, size_t n, int32_t offset )
{
    size_t t = ( size_t )blockIdx.x * blockDim.x + threadIdx.x;
    if( t < n ) lab[ t ] += offset;
}

// ---- label stitch on the device ---------------------------------------------------------------------------------
// On the last own row of strip k both k and k+1 labelled the same pixels: every column gives an equivalence between two
// labels.  The distinct seam labels are the nodes of a union-find held in an open-addressing table (key = label,
// parent = a label of the same class, roots point at themselves); link-by-minimum makes the root the class minimum,
// which is the canonical label.  Three small launches close the equivalences of ALL seams (<= 7 x 4096 pairs), every device
// does so redundantly on its own copy of the seam rows, then relabels its own rows through the table.
constexpr int32_t kNoKey = -1;
__device__ __forceinline__ uint32_t tab_hash( int32_t label, uint32_t mask ) { return ( ( uint32_t )label * 0x9E3779B1u ) >> 7 & mask; }

__device__ __forceinline__ int tab_find_slot( const int32_t* keys, uint32_t mask, int32_t label )
{
    uint32_t h = tab_hash( label, mask );
    for( ;; )
    {
        const int32_t k = keys[ h ];
        if( k == label ) return ( int )h;
        if( k == kNoKey ) return -1;
        h = ( h + 1u ) & mask;
    }
}

__device__ __forceinline__ int32_t tab_root( const int32_t* keys, const int32_t* parent, uint32_t mask, int32_t label )
{
    for( ;; )
    {
        const int32_t p = *( volatile const int32_t* )&parent[ tab_find_slot( keys, mask, label ) ];
        if( p == label ) return label;
        label = p;
    }
}

// (1) the distinct labels become nodes, each its own root
__global__ void stitch_insert_kernel( const int32_t* __restrict__ seams, int n_labels, int32_t* keys, int32_t* parent, uint32_t mask )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if( i >= n_labels ) return;
    const int32_t label = seams[ i ];
    uint32_t h = tab_hash( label, mask );
    for( ;; )
    {
        const int32_t k = atomicCAS( &keys[ h ], kNoKey, label );
        if( k == kNoKey ) parent[ h ] = label;
        if( k == kNoKey || k == label ) break;
        h = ( h + 1u ) & mask;
    }
}

// (2) one union per column of every seam: the larger root goes under the smaller one (lock-free, roots only decrease)
__global__ void stitch_unite_kernel( const int32_t* __restrict__ seams, int n_seams, int width, const int32_t* keys, int32_t* parent, uint32_t mask )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if( i >= n_seams * width ) return;
    const int k = i / width, x = i - k * width;
    int32_t a = seams[ ( 2 * k ) * width + x ], b = seams[ ( 2 * k + 1 ) * width + x ];
    for( ;; )
    {
        a = tab_root( keys, parent, mask, a );
        b = tab_root( keys, parent, mask, b );
        if( a == b ) break;
        if( a < b )
        {
            const int32_t t = a;
            a = b;
            b = t;
        }
        const int32_t old = atomicMin( &parent[ tab_find_slot( keys, mask, a ) ], b );
        if( old == a ) break;
        a = old;
    }
}

// (3) flatten: every node points at its class minimum
__global__ void stitch_flatten_kernel( const int32_t* keys, int32_t* parent, uint32_t mask )
{
    const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if( h <= mask && keys[ h ] != kNoKey ) parent[ h ] = tab_root( keys, parent, mask, keys[ h ] );
}

// A strip's two seam rows — its last own row (row A of the seam above it... of seam k) and the row below its first own row
// (row B of seam k-1) — written straight into the seam buffers of ALL strips: peer stores over NVLink (the buffers of the
// other devices are mapped by peer access), one launch instead of two copies per seam and device.
__global__ void push_seam_rows_kernel( const int32_t* __restrict__ row_a, int slot_a, const int32_t* __restrict__ row_b, int slot_b, int width,
                                       int32_t* const* __restrict__ dst )
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if( x >= width ) return;
    int32_t* out = dst[ blockIdx.y ];
    if( row_a ) out[ ( size_t )slot_a * width + x ] = row_a[ x ];
    if( row_b ) out[ ( size_t )slot_b * width + x ] = row_b[ x ];
}

// labels that are nodes of the seam union-find are replaced by their class minimum (most labels are not: one probe)
__global__ void relabel_kernel( int32_t* lab, size_t n, const int32_t* __restrict__ keys, const int32_t* __restrict__ parent, uint32_t mask )
{
    size_t t = ( size_t )blockIdx.x * blockDim.x + threadIdx.x;
    if( t >= n ) return;
    const int32_t v = lab[ t ];
    uint32_t h = tab_hash( v, mask );
    for( ;; )
    {
        const int32_t k = __ldg( keys + h );
        if( k == kNoKey ) return;
        if( k == v )
        {
            const int32_t r = __ldg( parent + h );
            if( r != v ) lab[ t ] = r;
            return;
        }
        h = ( h + 1u ) & mask;
    }
}

} // namespace

namespace {
struct StripWorkers;
}

struct par_group
{
    int width = 0, height = 0, scale = 0;
    std::vector< Strip > strips;
    StripWorkers* workers = nullptr; // one host thread per strip beyond the first (the caller's thread drives strip 0)
    uint32_t tab_mask = 0; // slots - 1 of the seam tables
    bool peer_stores = false; // every strip can store into every other strip's seam buffer (same device, or peer access enabled)
    double last_wall_ms = 0.0, last_device_ms = 0.0;
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

namespace {

// the caller's current device is restored when an entry point returns
struct GroupDeviceGuard
{
    int prev = -1;
    GroupDeviceGuard() { if( cudaGetDevice( &prev ) != cudaSuccess ) prev = -1; }
    ~GroupDeviceGuard() { if( prev >= 0 ) cudaSetDevice( prev ); }
};

int bytes_per_pixel_of( int out_format ) { return out_format == PAR_OUT_RGBA8 ? 4 : ( out_format == PAR_OUT_BGR8 ? 3 : 1 ); }

// One call of the tiled path: what every strip does, in three phases separated by "all strips have enqueued the phase"
// barriers (a cudaStreamWaitEvent must find the event recorded by THIS call).  The strips are driven by one host thread
// each (a call is ~30 enqueues per strip, serially 1 ms on eight devices against 0.1 ms of kernels per strip); everything
// enqueued is stream-ordered and asynchronous.
struct TiledCall
{
    par_group* g;
    unsigned flags;
    int out_format;
    bool want_image, want_labels, timed;
    const par_job* host; // host entry: the caller's job (image in, outputs out); nullptr for the device-resident entry
};

// phase A: the strip's own rows are on the device (host entry: uploaded here) -> `ready`
int strip_phase_a( const TiledCall& c, Strip& s, std::string& err )
{
    cudaSetDevice( s.device );
    if( c.timed ) cudaEventRecord( s.t_begin, s.stream );
    if( c.host )
    {
        const size_t row_in = ( size_t )3 * c.g->width;
        uint8_t* dst = s.d_in + ( size_t )( s.own_b - s.load_b ) * row_in;
        cudaError_t e = cudaMemcpy2DAsync( dst, row_in, c.host->bgr + ( size_t )s.own_b * c.host->widthstep, c.host->widthstep, row_in, s.own_e - s.own_b,
                                           cudaMemcpyHostToDevice, s.stream );
        if( e != cudaSuccess )
        {
            err = std::string( "H2D: " ) + cudaGetErrorString( e );
            return PAR_ERR_CUDA;
        }
    }
    cudaEventRecord( s.ready, s.stream );
    return PAR_OK;
}

// phase B: apron rows from the neighbouring devices (peer copies over NVLink, ordered after their `ready`), the ordinary
// path on the extended strip -> `computed`, and with labels the strip's own labelling -> `labelled`
int strip_phase_b( const TiledCall& c, size_t k, std::string& err )
{
    par_group* g = c.g;
    Strip& s = g->strips[ k ];
    const int W = g->width;
    const size_t row_in = ( size_t )3 * W; // strips are stored densely on the devices
    cudaSetDevice( s.device );
    cudaError_t e = cudaSuccess;
    if( c.want_labels && g->strips.size() > 1 ) e = cudaMemsetAsync( s.d_tab_keys, 0xFF, ( ( size_t )g->tab_mask + 1 ) * 4, s.stream ); // (off the critical path)
    if( e == cudaSuccess && k > 0 && s.load_b < s.own_b )
    {
        Strip& n = g->strips[ k - 1 ];
        cudaStreamWaitEvent( s.stream, n.ready, 0 );
        e = cudaMemcpyPeerAsync( s.d_in, s.device, n.d_in + ( size_t )( s.load_b - n.load_b ) * row_in, n.device, ( size_t )( s.own_b - s.load_b ) * row_in, s.stream );
    }
    if( e == cudaSuccess && k + 1 < g->strips.size() && s.load_e > s.own_e )
    {
        Strip& n = g->strips[ k + 1 ];
        cudaStreamWaitEvent( s.stream, n.ready, 0 );
        e = cudaMemcpyPeerAsync( s.d_in + ( size_t )( s.own_e - s.load_b ) * row_in, s.device, n.d_in + ( size_t )( s.own_e - n.load_b ) * row_in, n.device,
                                 ( size_t )( s.load_e - s.own_e ) * row_in, s.stream );
    }
    if( e != cudaSuccess )
    {
        err = std::string( "peer copy: " ) + cudaGetErrorString( e );
        return PAR_ERR_CUDA;
    }
    par_job d = {};
    d.bgr = s.d_in;
    d.width = W;
    d.height = s.load_e - s.load_b;
    d.widthstep = 3 * W;
    d.n_frames = 1;
    d.scale = g->scale;
    d.flags = c.flags;
    d.out_format = c.out_format;
    d.rgba = c.want_image ? s.d_rgba : nullptr;
    d.graph = s.d_graph;
    d.graph_aux = s.d_aux;
    int st = par_remaster_device( s.ctx, &d );
    if( st != PAR_OK )
    {
        err = std::string( "strip: " ) + par_last_error( s.ctx );
        return st;
    }
    cudaEventRecord( s.computed, s.stream );
    if( c.want_labels )
    {
        // Label only rows whose graph bytes are exact: the strip's own rows plus ONE row either side (needed
        // for the stitch).  The outer apron rows carry graph bytes computed without their full context and
        // must not be allowed to connect anything.  The window is labelled as an image of its own (links
        // leaving it are ignored) and then shifted to global pixel indices.
        const int win_b = std::max( s.load_b, s.own_b - 1 ), win_e = std::min( s.load_e, s.own_e + 1 );
        par_job l = d;
        l.rgba = nullptr;
        l.height = win_e - win_b;
        l.graph = s.d_graph + ( size_t )( win_b - s.load_b ) * W;
        l.labels = s.d_labels + ( size_t )( win_b - s.load_b ) * W;
        st = par_stage_cc_labels( s.ctx, &l );
        if( st != PAR_OK )
        {
            err = std::string( "labels: " ) + par_last_error( s.ctx );
            return st;
        }
        const size_t n = ( size_t )W * l.height;
        offset_labels_kernel<<< ( unsigned )( ( n + 255 ) / 256 ), 256, 0, s.stream >>>( l.labels, n, win_b * W );
        const size_t n_strips = g->strips.size();
        if( g->peer_stores && n_strips > 1 )
        {
            // this strip's rows of the seams above and below it, stored into every strip's seam buffer
            const int32_t* row_a = k + 1 < n_strips ? s.d_labels + ( size_t )( s.own_e - 1 - s.load_b ) * W : nullptr;
            const int32_t* row_b = k > 0 ? s.d_labels + ( size_t )( s.own_b - 1 - s.load_b ) * W : nullptr;
            push_seam_rows_kernel<<< dim3( ( unsigned )( ( W + 255 ) / 256 ), ( unsigned )n_strips ), 256, 0, s.stream >>>(
                row_a, ( int )( 2 * k ), row_b, ( int )( 2 * ( k - 1 ) + 1 ), W, s.d_seam_dst );
        }
        cudaEventRecord( s.labelled, s.stream );
    }
    return PAR_OK;
}

// phase C: the stitch — the two label rows of every seam gathered from all strips (peer copies, ordered after their
// `labelled`), the equivalences closed in one CTA, the strip's own rows relabelled — and, for the host entry, the results on
// their way back: image and graphs on a second stream as soon as the strip has them, the labels once they are final
int strip_phase_c( const TiledCall& c, size_t k, std::string& err )
{
    par_group* g = c.g;
    Strip& s = g->strips[ k ];
    const int W = g->width, S = g->scale;
    const size_t n_seams = g->strips.size() - 1;
    cudaSetDevice( s.device );
    cudaError_t e = cudaSuccess;
    if( c.want_labels && n_seams > 0 )
    {
        for( auto& o : g->strips )
            if( &o != &s ) cudaStreamWaitEvent( s.stream, o.labelled, 0 );
        if( !g->peer_stores ) // (no peer mapping between some pair of devices: fetch the rows with peer copies instead)
            for( size_t m = 0; m < n_seams && e == cudaSuccess; m++ )
            {
                Strip &a = g->strips[ m ], &b = g->strips[ m + 1 ];
                const int row = a.own_e - 1;
                e = cudaMemcpyPeerAsync( s.d_seams + ( 2 * m ) * W, s.device, a.d_labels + ( size_t )( row - a.load_b ) * W, a.device, ( size_t )W * 4, s.stream );
                if( e == cudaSuccess )
                    e = cudaMemcpyPeerAsync( s.d_seams + ( 2 * m + 1 ) * W, s.device, b.d_labels + ( size_t )( row - b.load_b ) * W, b.device, ( size_t )W * 4, s.stream );
            }
        if( e == cudaSuccess )
        {
            const int n_labels = ( int )( 2 * n_seams ) * W, n_pairs = ( int )n_seams * W;
            stitch_insert_kernel<<< ( n_labels + 255 ) / 256, 256, 0, s.stream >>>( s.d_seams, n_labels, s.d_tab_keys, s.d_tab_parent, g->tab_mask );
            stitch_unite_kernel<<< ( n_pairs + 255 ) / 256, 256, 0, s.stream >>>( s.d_seams, ( int )n_seams, W, s.d_tab_keys, s.d_tab_parent, g->tab_mask );
            stitch_flatten_kernel<<< ( g->tab_mask + 256 ) / 256, 256, 0, s.stream >>>( s.d_tab_keys, s.d_tab_parent, g->tab_mask );
            const size_t n = ( size_t )W * ( s.own_e - s.own_b );
            relabel_kernel<<< ( unsigned )( ( n + 255 ) / 256 ), 256, 0, s.stream >>>( s.d_labels + ( size_t )( s.own_b - s.load_b ) * W, n, s.d_tab_keys, s.d_tab_parent,
                                                                                       g->tab_mask );
            e = cudaGetLastError();
        }
        if( e != cudaSuccess )
        {
            err = std::string( "stitch: " ) + cudaGetErrorString( e );
            return PAR_ERR_CUDA;
        }
    }
    if( c.timed ) cudaEventRecord( s.t_end, s.stream );
    if( c.host )
    {
        const par_job* j = c.host;
        const size_t out_row = ( size_t )W * S * bytes_per_pixel_of( j->out_format );
        const size_t own_px = ( size_t )W * ( s.own_e - s.own_b ), off_px = ( size_t )W * ( s.own_b - s.load_b );
        cudaStream_t cs = s.copy_stream;
        cudaStreamWaitEvent( cs, s.computed, 0 );
        if( j->rgba )
        {
            const bool flip = ( j->flags & PAR_FLAG_FLIP_OUTPUT ) != 0;
            const size_t src_row = flip ? ( size_t )( s.load_e - s.own_e ) * S : ( size_t )( s.own_b - s.load_b ) * S;
            const size_t dst_row = flip ? ( size_t )( g->height - s.own_e ) * S : ( size_t )s.own_b * S;
            e = cudaMemcpyAsync( j->rgba + dst_row * out_row, s.d_rgba + src_row * out_row, ( size_t )( s.own_e - s.own_b ) * S * out_row, cudaMemcpyDeviceToHost, cs );
        }
        if( e == cudaSuccess && j->graph ) e = cudaMemcpyAsync( j->graph + ( size_t )W * s.own_b, s.d_graph + off_px, own_px, cudaMemcpyDeviceToHost, cs );
        if( e == cudaSuccess && j->graph_aux ) e = cudaMemcpyAsync( j->graph_aux + ( size_t )W * s.own_b, s.d_aux + off_px, own_px, cudaMemcpyDeviceToHost, cs );
        if( e == cudaSuccess && j->labels )
            e = cudaMemcpyAsync( j->labels + ( size_t )W * s.own_b, s.d_labels + off_px, own_px * 4, cudaMemcpyDeviceToHost, s.stream );
        if( e != cudaSuccess )
        {
            err = std::string( "D2H: " ) + cudaGetErrorString( e );
            return PAR_ERR_CUDA;
        }
    }
    return PAR_OK;
}

// a reusable barrier for the strips' host threads
struct PhaseBarrier
{
    std::mutex m;
    std::condition_variable cv;
    int waiting = 0, generation = 0;
    const int n;
    explicit PhaseBarrier( int n_ ) : n( n_ ) {}
    void arrive_and_wait()
    {
        std::unique_lock< std::mutex > lock( m );
        const int gen = generation;
        if( ++waiting == n )
        {
            waiting = 0;
            generation++;
            cv.notify_all();
        }
        else
            cv.wait( lock, [ & ] { return generation != gen; } );
    }
};

// The strips' host threads: created with the group, parked on a condition variable between calls (spawning seven threads
// per call would cost more than the call's kernels).
struct StripWorkers
{
    std::vector< std::thread > threads;
    std::mutex m;
    std::condition_variable wake, finished;
    const TiledCall* call = nullptr;
    std::vector< int >* status = nullptr;
    std::vector< std::string >* errors = nullptr;
    PhaseBarrier* barrier = nullptr;
    long generation = 0;
    int running = 0;
    bool stop = false;

    static void work( const TiledCall& c, size_t k, std::vector< int >& status, std::vector< std::string >& errors, PhaseBarrier& barrier )
    {
        // (a strip that failed still goes through the barriers, so that the others are not left waiting)
        status[ k ] = strip_phase_a( c, c.g->strips[ k ], errors[ k ] );
        barrier.arrive_and_wait();
        if( status[ k ] == PAR_OK ) status[ k ] = strip_phase_b( c, k, errors[ k ] );
        barrier.arrive_and_wait();
        if( status[ k ] == PAR_OK ) status[ k ] = strip_phase_c( c, k, errors[ k ] );
    }
    explicit StripWorkers( size_t n_strips )
    {
        for( size_t k = 1; k < n_strips; k++ )
            threads.emplace_back( [ this, k ] {
                long seen = 0;
                for( ;; )
                {
                    std::unique_lock< std::mutex > lock( m );
                    wake.wait( lock, [ & ] { return stop || generation != seen; } );
                    if( stop ) return;
                    seen = generation;
                    const TiledCall* c = call;
                    std::vector< int >* st = status;
                    std::vector< std::string >* er = errors;
                    PhaseBarrier* b = barrier;
                    lock.unlock();
                    work( *c, k, *st, *er, *b );
                    lock.lock();
                    if( --running == 0 ) finished.notify_one();
                }
            } );
    }
    ~StripWorkers()
    {
        {
            std::lock_guard< std::mutex > lock( m );
            stop = true;
        }
        wake.notify_all();
        for( auto& t : threads ) t.join();
    }
};

// enqueue one call on all strips; returns the first failure (its message in g->error)
int run_tiled_call( const TiledCall& c )
{
    par_group* g = c.g;
    const size_t n = g->strips.size();
    std::vector< int > status( n, PAR_OK );
    std::vector< std::string > errors( n );
    PhaseBarrier barrier( ( int )n );
    if( n > 1 )
    {
        StripWorkers& w = *g->workers;
        {
            std::lock_guard< std::mutex > lock( w.m );
            w.call = &c;
            w.status = &status;
            w.errors = &errors;
            w.barrier = &barrier;
            w.running = ( int )n - 1;
            w.generation++;
        }
        w.wake.notify_all();
        StripWorkers::work( c, 0, status, errors, barrier );
        std::unique_lock< std::mutex > lock( w.m );
        w.finished.wait( lock, [ & ] { return w.running == 0; } );
    }
    else
        StripWorkers::work( c, 0, status, errors, barrier );
    for( size_t k = 0; k < n; k++ )
        if( status[ k ] != PAR_OK ) return g->fail( status[ k ], "strip %d on device %d: %s", ( int )k, g->strips[ k ].device, errors[ k ].c_str() );
    return PAR_OK;
}

int check_group_job( par_group* g, unsigned flags, int out_format )
{
    if( out_format != PAR_OUT_RGBA8 && out_format != PAR_OUT_BGR8 )
        return g->fail( PAR_ERR_INVALID, "tiled mode offers PAR_OUT_RGBA8 and PAR_OUT_BGR8 (a palette would be per strip)" );
    if( flags & ( PAR_FLAG_AA2 | PAR_FLAG_AA4 ) )
    {
        const int aa = ( flags & PAR_FLAG_AA4 ) ? 4 : 2;
        if( !par::raster_aa_supported( g->scale, aa ) ) return g->fail( PAR_ERR_INVALID, "unsupported scale %d with %dx%d samples per pixel", g->scale, aa, aa );
    }
    return PAR_OK;
}

} // namespace

extern "C" {

void par_group_destroy( par_group* g )
{
    if( !g ) return;
    GroupDeviceGuard guard;
    delete g->workers;
    g->workers = nullptr;
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
        cudaFree( s.d_seams );
        cudaFree( s.d_tab_keys );
        cudaFree( s.d_tab_parent );
        cudaFree( s.d_seam_dst );
        for( cudaEvent_t ev : { s.ready, s.computed, s.labelled, s.t_begin, s.t_end } )
            if( ev ) cudaEventDestroy( ev );
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
    GroupDeviceGuard guard;
    par_group* g = new par_group();
    g->width = width;
    g->height = height;
    g->scale = scale;
    g->strips.resize( n_devices );
    const size_t seam_labels = ( size_t )2 * ( n_devices - 1 ) * width;
    uint32_t slots = 1024;
    while( slots < 4 * seam_labels ) slots <<= 1;
    g->tab_mask = slots - 1;
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
        if( e == cudaSuccess && n_devices > 1 ) e = cudaMalloc( &s.d_seams, seam_labels * 4 );
        if( e == cudaSuccess && n_devices > 1 ) e = cudaMalloc( &s.d_tab_keys, ( size_t )slots * 4 );
        if( e == cudaSuccess && n_devices > 1 ) e = cudaMalloc( &s.d_tab_parent, ( size_t )slots * 4 );
        if( e == cudaSuccess ) e = cudaEventCreateWithFlags( &s.ready, cudaEventDisableTiming );
        if( e == cudaSuccess ) e = cudaEventCreateWithFlags( &s.computed, cudaEventDisableTiming );
        if( e == cudaSuccess ) e = cudaEventCreateWithFlags( &s.labelled, cudaEventDisableTiming );
        if( e == cudaSuccess ) e = cudaEventCreate( &s.t_begin );
        if( e == cudaSuccess ) e = cudaEventCreate( &s.t_end );
        if( e == cudaSuccess ) e = cudaStreamCreateWithFlags( &s.copy_stream, cudaStreamNonBlocking );
        if( e != cudaSuccess )
        {
            g_group_create_error = std::string( "par_group_create: " ) + cudaGetErrorString( e );
            par_group_destroy( g );
            return PAR_ERR_CUDA;
        }
    }
    // direct NVLink/PCIe peer access between all pairs of devices where the topology allows it (aprons travel between
    // neighbours, seam label rows between all strips)
    g->peer_stores = true;
    for( int p = 0; p < n_devices; p++ )
        for( int q = 0; q < n_devices; q++ )
        {
            const int a = g->strips[ p ].device, b = g->strips[ q ].device;
            if( a == b ) continue;
            int ok = 0;
            if( cudaDeviceCanAccessPeer( &ok, a, b ) == cudaSuccess && ok )
            {
                cudaSetDevice( a );
                const cudaError_t pe = cudaDeviceEnablePeerAccess( b, 0 );
                if( pe != cudaSuccess ) cudaGetLastError();
                if( pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled ) g->peer_stores = false;
            }
            else
                g->peer_stores = false;
        }
    if( g->peer_stores && n_devices > 1 ) // every strip's table of all seam buffers
    {
        std::vector< int32_t* > all( n_devices );
        for( int k = 0; k < n_devices; k++ ) all[ k ] = g->strips[ k ].d_seams;
        for( auto& s : g->strips )
        {
            cudaSetDevice( s.device );
            cudaError_t e = cudaMalloc( &s.d_seam_dst, n_devices * sizeof( int32_t* ) );
            if( e == cudaSuccess ) e = cudaMemcpy( s.d_seam_dst, all.data(), n_devices * sizeof( int32_t* ), cudaMemcpyHostToDevice );
            if( e != cudaSuccess )
            {
                g_group_create_error = std::string( "par_group_create: " ) + cudaGetErrorString( e );
                par_group_destroy( g );
                return PAR_ERR_CUDA;
            }
        }
    }
    if( n_devices > 1 ) g->workers = new StripWorkers( ( size_t )n_devices );
    *out = g;
    return PAR_OK;
}

int par_group_n_strips( const par_group* g ) { return g ? ( int )g->strips.size() : 0; }

int par_group_strip( const par_group* g, int k, par_strip* out )
{
    if( !g || !out || k < 0 || k >= ( int )g->strips.size() ) return PAR_ERR_INVALID;
    const Strip& s = g->strips[ k ];
    out->device = s.device;
    out->own_begin = s.own_b;
    out->own_end = s.own_e;
    out->load_begin = s.load_b;
    out->load_end = s.load_e;
    out->bgr = s.d_in;
    out->image = s.d_rgba;
    out->graph = s.d_graph;
    out->graph_aux = s.d_aux;
    out->labels = s.d_labels;
    return PAR_OK;
}

int par_group_remaster_device( par_group* g, unsigned flags, int out_format, int want_image, int want_labels )
{
    if( !g ) return PAR_ERR_INVALID;
    int st = check_group_job( g, flags, out_format );
    if( st ) return st;
    GroupDeviceGuard guard;
    const auto t0 = std::chrono::steady_clock::now();
    const TiledCall call{ g, flags, out_format, want_image != 0, want_labels != 0, true, nullptr }; // (the caller has put every strip's own rows in place)
    st = run_tiled_call( call );
    if( st ) return st;
    g->last_device_ms = 0.0;
    for( auto& s : g->strips )
    {
        cudaSetDevice( s.device );
        cudaError_t e = cudaStreamSynchronize( s.stream );
        if( e != cudaSuccess ) return g->fail( PAR_ERR_CUDA, "strip on device %d: %s", s.device, cudaGetErrorString( e ) );
        float ms = 0.f;
        cudaEventElapsedTime( &ms, s.t_begin, s.t_end );
        g->last_device_ms = std::max( g->last_device_ms, ( double )ms );
    }
    g->last_wall_ms = std::chrono::duration< double, std::milli >( std::chrono::steady_clock::now() - t0 ).count();
    return PAR_OK;
}

int par_group_last_ms( const par_group* g, double* wall_ms, double* device_ms )
{
    if( !g ) return PAR_ERR_INVALID;
    if( wall_ms ) *wall_ms = g->last_wall_ms;
    if( device_ms ) *device_ms = g->last_device_ms;
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
    int st = check_group_job( g, j->flags, j->out_format );
    if( st ) return st;
    GroupDeviceGuard guard;
    // own rows up, halo exchange, kernels, label stitch, results down: all enqueued by the strips' threads
    const TiledCall call{ g, j->flags, j->out_format, j->rgba != nullptr, j->labels != nullptr, false, j };
    st = run_tiled_call( call );
    if( st ) return st;
    cudaError_t e = cudaSuccess;
    for( auto& s : g->strips )
    {
        cudaSetDevice( s.device );
        e = cudaStreamSynchronize( s.copy_stream );
        if( e == cudaSuccess ) e = cudaStreamSynchronize( s.stream );
        if( e != cudaSuccess ) return g->fail( PAR_ERR_CUDA, "strip on device %d: %s", s.device, cudaGetErrorString( e ) );
    }
    return PAR_OK;
}

} // extern "C"
