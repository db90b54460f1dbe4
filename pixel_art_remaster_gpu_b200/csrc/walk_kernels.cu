// Border walk of every connected component: the input of the spline stage the reference's author had
// started (extractBorderPoints, cc_functions.cu:348-503 — dead code in the reference, a serial single-thread walker).
//
// What is reproduced: for every component, the FIRST walk the reference makes for it — the one that starts at the
// component's first node in raster order, i.e. at its label.  For that walk the reference's scan-order bookkeeping
// (listed / discarded, :394-444) reduces to pure functions of the graph: the walk follows the outer face clockwise
// (nextNodeClockwise, :295-318) until it is about to re-enter the start through the edge it would have arrived by
// (:415-416); it is dropped when it steps on an interior node (== 90, :425) or meets its start early (:426).  Walks the
// reference starts later from other nodes of the same component depend on its serial state and are not produced.
// An island (no links) is undefined in the reference (edge = -1 falls off c_neighbor_index's switch, :215-246): here it
// is a walk of one node.
//
// Design: components are independent, so one thread per label pixel walks its component; a counting pass, a per-frame
// exclusive scan of the lengths, and a writing pass that repeats the walk into its slot — the output is the reference's
// CClist / CCsizes layout (walks concatenated in raster order of their start) without a serial dependency.
#include "kernels.cuh"

namespace par {

namespace {

constexpr int kThreads = 256;

// clock position 0..7 (clockwise from the up-left neighbour) <-> graph bit (getRealLinkIndex :68-99, getClockLinkIndex :153-184)
__device__ __forceinline__ int clock_to_bit( int c ) { return ( int )( ( 0x35674210u >> ( 4 * c ) ) & 15u ); }
__device__ __forceinline__ int bit_to_clock( int b ) { return ( int )( ( 0x45637210u >> ( 4 * b ) ) & 15u ); }

// first link clockwise strictly after clock position c0 (c0+1 .. c0+7); the node has at least one other link
__device__ __forceinline__ int next_clockwise( uint32_t node, int c0 )
{
    for( int c = c0 + 1; c <= c0 + 7; c++ )
    {
        const int b = clock_to_bit( c & 7 );
        if( node & ( 1u << b ) ) return b;
    }
    return clock_to_bit( c0 & 7 );
}

// Walks the component that starts at `start`; calls sink( k, node index ) for the k-th node.  Returns the number of
// nodes, 0 when the reference drops the walk.
template< class Sink >
__device__ __forceinline__ int walk_component( const uint8_t* __restrict__ g, int start, int width, long n_px, Sink sink )
{
    const uint32_t first = __ldg( g + start );
    sink( 0, start );
    if( first == 0u ) return 1; // island
    int edge = -1;              // getFirstLink (:107-118)
    for( int c = 0; c < 8 && edge < 0; c++ )
        if( first & ( 1u << clock_to_bit( c ) ) ) edge = clock_to_bit( c );
    int arrival = edge; // nextEdgeCounterClockwise (:262-285): first link counter-clockwise after `edge`
    if( __popc( first ) != 1 )
    {
        const int c0 = bit_to_clock( edge );
        for( int c = c0 + 7; c >= c0 + 1; c-- )
        {
            const int b = clock_to_bit( c & 7 );
            if( first & ( 1u << b ) )
            {
                arrival = b;
                break;
            }
        }
    }
    int index = start, count = 1;
    const long limit = 16 * n_px; // (a face of a finite graph closes long before; guards a malformed caller-supplied graph)
    for( long step = 0; step < limit; step++ )
    {
        const int next = index + edge_dj( edge ) * width + edge_di( edge ); // c_neighbor_index (:215-246)
        if( next == start && 7 - edge == arrival ) return count;             // the loop closes (:415-416)
        if( ( unsigned long )next >= ( unsigned long )n_px ) return 0;
        index = next;
        const uint32_t node = __ldg( g + index );
        edge = __popc( node ) == 1 ? 7 - edge : next_clockwise( node, bit_to_clock( 7 - edge ) ); // nextNodeClockwise (:295-318)
        if( node == 90u || index == start ) return 0; // interior node / back at the start too early: dropped (:425-438)
        sink( count, index );
        count++;
    }
    return 0;
}

__global__ void __launch_bounds__( kThreads ) walk_count_kernel( const uint8_t* graph, const int32_t* labels, int width, int height, int32_t* walk_len )
{
    const long n_px = ( long )width * height;
    const long n = ( long )blockIdx.x * kThreads + threadIdx.x;
    if( n >= n_px ) return;
    const size_t f = blockIdx.y;
    int len = 0;
    if( labels[ f * n_px + n ] == ( int32_t )n ) len = walk_component( graph + f * n_px, ( int )n, width, n_px, []( int, int ) {} );
    walk_len[ f * n_px + n ] = len;
}

// per frame: exclusive scan of walk_len in raster order -> walk_begin, total; one CTA per frame
__global__ void __launch_bounds__( 1024 ) walk_scan_kernel( const int32_t* walk_len, long n_px, int32_t* walk_begin, long long* total )
{
    __shared__ long long s_sum[ 1024 ];
    const size_t f = blockIdx.x;
    const int32_t* len = walk_len + f * n_px;
    int32_t* begin = walk_begin + f * n_px;
    const long chunk = ( n_px + 1023 ) / 1024, lo = threadIdx.x * chunk, hi = min( lo + chunk, n_px );
    long long mine = 0;
    for( long k = lo; k < hi; k++ ) mine += len[ k ];
    s_sum[ threadIdx.x ] = mine;
    __syncthreads();
    for( int d = 1; d < 1024; d <<= 1 ) // inclusive scan of the chunk sums
    {
        const long long v = threadIdx.x >= d ? s_sum[ threadIdx.x - d ] : 0;
        __syncthreads();
        s_sum[ threadIdx.x ] += v;
        __syncthreads();
    }
    long long at = s_sum[ threadIdx.x ] - mine;
    for( long k = lo; k < hi; k++ )
    {
        begin[ k ] = ( int32_t )min( at, ( long long )0x7FFFFFFF );
        at += len[ k ];
    }
    if( threadIdx.x == 1023 ) total[ f ] = s_sum[ 1023 ];
}

__global__ void __launch_bounds__( kThreads ) walk_write_kernel( const uint8_t* graph, int width, int height, const int32_t* walk_len, const int32_t* walk_begin,
                                                               const long long* total, int32_t* nodes, long long capacity )
{
    const long n_px = ( long )width * height;
    const long n = ( long )blockIdx.x * kThreads + threadIdx.x;
    if( n >= n_px ) return;
    const size_t f = blockIdx.y;
    if( walk_len[ f * n_px + n ] == 0 || total[ f ] > capacity ) return; // (a frame whose walks do not fit is reported through `total`)
    int32_t* out = nodes + f * capacity + walk_begin[ f * n_px + n ];
    walk_component( graph + f * n_px, ( int )n, width, n_px, [ out ]( int k, int node ) { out[ k ] = node; } );
}

// ---- splines through the walks -------------------------------------------------------------------------------------
// The stage the reference's author had started the walker for (Kopf-Lischinski: quadratic B-splines along the region
// outlines; the reference stops at the walk, cc_functions.cu:348-503).  Every walk is a closed control polygon — the centres
// (x + 1/2, y + 1/2) of its nodes in walk order — and the curve is the closed uniform quadratic B-spline over it: segment i
// runs from the midpoint of P(i-1) P(i) to the midpoint of P(i) P(i+1),
//     B_i(t) = 1/2 [ (1-t)^2 P(i-1) + (-2t^2 + 2t + 1) P(i) + t^2 P(i+1) ],   t in [0, 1),
// sampled at t = s / k, s = 0 .. k-1, for k a power of two: every weight is a multiple of 1/(2 k^2) and every coordinate a
// multiple of 1/2, so the float results are exact and equal any other evaluation order bit for bit.
// One thread per walk start walks its own nodes: sample (b + i) * k + s of the frame belongs to node i of the walk that
// begins at entry b — the layout of walk_nodes, times k.
__global__ void __launch_bounds__( kThreads ) walk_spline_kernel( const int32_t* walk_len, const int32_t* walk_begin, const long long* total, const int32_t* nodes,
                                                                int width, long n_px, long long capacity, int k, float2* points )
{
    const long n = ( long )blockIdx.x * kThreads + threadIdx.x;
    if( n >= n_px ) return;
    const size_t f = blockIdx.y;
    const int len = walk_len[ f * n_px + n ];
    if( len == 0 || total[ f ] > capacity ) return;
    const long long b = walk_begin[ f * n_px + n ];
    const int32_t* w = nodes + f * capacity + b;
    float2* out = points + ( f * capacity + b ) * k;
    auto centre = [ width ]( int node ) { return make_float2( ( float )( node % width ) + 0.5f, ( float )( node / width ) + 0.5f ); };
    float2 prev = centre( w[ len - 1 ] ), cur = centre( w[ 0 ] );
    const float inv_k = 1.0f / ( float )k;
    for( int i = 0; i < len; i++ )
    {
        const float2 next = centre( w[ i + 1 == len ? 0 : i + 1 ] );
        for( int s = 0; s < k; s++ )
        {
            const float t = ( float )s * inv_k;
            const float w0 = 0.5f * ( 1.0f - t ) * ( 1.0f - t ), w2 = 0.5f * t * t, w1 = 1.0f - w0 - w2;
            out[ ( size_t )i * k + s ] = make_float2( w0 * prev.x + w1 * cur.x + w2 * next.x, w0 * prev.y + w1 * cur.y + w2 * next.y );
        }
        prev = cur;
        cur = next;
    }
}

} // namespace

cudaError_t launch_border_walks( const uint8_t* graph, const int32_t* labels, int width, int height, int n_frames, int32_t* walk_len,
                                 int32_t* walk_begin, long long* total, int32_t* nodes, long long capacity, cudaStream_t stream )
{
    const long n_px = ( long )width * height;
    for( int f0 = 0; f0 < n_frames; f0 += 65535 ) // (grid.y limit)
    {
        const int nf = n_frames - f0 < 65535 ? n_frames - f0 : 65535;
        const size_t off = ( size_t )f0 * n_px;
        const dim3 grid( ( unsigned )( ( n_px + kThreads - 1 ) / kThreads ), nf );
        walk_count_kernel<<< grid, kThreads, 0, stream >>>( graph + off, labels + off, width, height, walk_len + off );
        walk_scan_kernel<<< nf, 1024, 0, stream >>>( walk_len + off, n_px, walk_begin + off, total + f0 );
        walk_write_kernel<<< grid, kThreads, 0, stream >>>( graph + off, width, height, walk_len + off, walk_begin + off, total + f0,
                                                            nodes + ( size_t )f0 * capacity, capacity );
    }
    return cudaGetLastError();
}

} // namespace par

namespace par {

cudaError_t launch_walk_splines( const int32_t* walk_len, const int32_t* walk_begin, const long long* total, const int32_t* nodes, int width, int height,
                                 int n_frames, long long capacity, int samples, float* points, cudaStream_t stream )
{
    const long n_px = ( long )width * height;
    for( int f0 = 0; f0 < n_frames; f0 += 65535 ) // (grid.y limit)
    {
        const int nf = n_frames - f0 < 65535 ? n_frames - f0 : 65535;
        const size_t off = ( size_t )f0 * n_px;
        const dim3 grid( ( unsigned )( ( n_px + kThreads - 1 ) / kThreads ), nf );
        walk_spline_kernel<<< grid, kThreads, 0, stream >>>( walk_len + off, walk_begin + off, total + f0, nodes + ( size_t )f0 * capacity, width, n_px, capacity,
                                                             samples, reinterpret_cast< float2* >( points ) + ( size_t )f0 * capacity * samples );
    }
    return cudaGetLastError();
}

} // namespace par
