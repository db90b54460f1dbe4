// Connected components of the final similarity graph: label[n] = smallest row-major index of n's
// component (8-connectivity as given by the graph bytes, interior nodes included).
//
// New subsystem: the reference's cc_functions.cu (extractBorderPoints, :348-503) is a serial
// single-thread border walker that is not part of its build (cc_kernel_call.bkp:3-12); what it
// defines is a component's identity — its first node in raster order (:394-413) — which is the
// canonical minimum-index label produced here, so any correct labeller is bit-exact.
//
// Design: lock-free union-find with link-by-minimum-index over the label array itself.
// (Tried and dropped in round 2: one CTA per frame of <= 65536 pixels with the whole forest as 16-bit labels in 168 KB of
// shared memory — no seams, no global atomics, no flatten kernel, but one CTA per SM, two CTA-wide barriers per 1024 pixels
// and CAS loops for the 16-bit minimum: 2.08 ms per 1024 frames against 0.91 ms for the three kernels below,
// profiles/r2e_stage_times_*.)
//   (1) tile pass: one CTA per 64x16 tile builds the tile's forest in SHARED memory (horizontal runs by
//       warp ballot, the few unions that can still connect two runs queued and executed one per lane,
//       flatten) and writes tile-local roots as global indices;
//   (2) seam pass: one thread per seam pixel unions across the seams with global atomicMin;
//   (3) flatten pass: every pixel replaces its label by its root.
// Roots only ever point to smaller indices, so the root of a set is its minimum index.
// Algorithmic HBM traffic: 1 B/px in + 4 B/px out.
#include "kernels.cuh"

namespace par {

namespace {

constexpr int kTW = 64, kTH = 16, kThreads = 256;

__device__ __forceinline__ int find_root( const int* lab, int x )
{
    int p = lab[ x ];
    while( p != x )
    {
        x = p;
        p = lab[ x ];
    }
    return x;
}

// find with path halving for the shared-memory forest: every step re-points x at its grandparent.  The write is an
// atomicMin: parents only ever decrease, so concurrent unions and other halving writes can only be improved on, never
// undone (and a plain store here would be a benign but reportable race).
__device__ __forceinline__ int find_root_halving( int* lab, int x )
{
    int p = lab[ x ];
    while( p != x )
    {
        const int gp = lab[ p ];
        if( gp != p ) atomicMin( &lab[ x ], gp );
        x = p;
        p = gp;
    }
    return x;
}

// volatile-free lock-free union: hang the larger root under the smaller one
__device__ __forceinline__ void unite( int* lab, int a, int b )
{
    for( ;; )
    {
        a = find_root( lab, a );
        b = find_root( lab, b );
        if( a == b ) return;
        if( a < b )
        {
            int t = a;
            a = b;
            b = t;
        }
        int old = atomicMin( &lab[ a ], b );
        if( old == a ) return;
        a = old;
    }
}

// Tile pass.  Two things keep the shared-memory forest shallow and the number of unions small:
//  * horizontal runs first: a warp owns 32 consecutive pixels of a row, one ballot gives the "linked to the right"
//    bits and every pixel starts with the index of the first pixel of its run as its label — no union at all;
//  * a link to the row above is only united when it can connect two runs that are not already connected through
//    the pixel to the left (same run below, same run above, also linked upwards), and a diagonal link only when its
//    target is not in the run of an orthogonal link that is united anyway.
// What remains is roughly one union per pair of touching runs.
__device__ __forceinline__ void unite_halving( int* lab, int a, int b )
{
    for( ;; )
    {
        a = find_root_halving( lab, a );
        b = find_root_halving( lab, b );
        if( a == b ) return;
        if( a < b )
        {
            int t = a;
            a = b;
            b = t;
        }
        int old = atomicMin( &lab[ a ], b );
        if( old == a ) return;
        a = old;
    }
}

__global__ void __launch_bounds__( kThreads ) cc_tile_kernel( LabelArgs a )
{
    __shared__ int s_lab[ kTW * kTH ];
    __shared__ uint8_t s_g[ kTW * kTH ];       // node bytes with the links that leave the tile (or the image) removed
    __shared__ uint32_t s_req[ 3 * kTW * kTH + kTW * kTH / 32 ]; // queued unions, a << 16 | b (three per pixel, one more for a warp's last lane)
    __shared__ int s_n;
    if( threadIdx.x == 0 ) s_n = 0;
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH, f = blockIdx.z;
    const size_t frame_px = ( size_t )a.width * a.height;
    const uint8_t* g = a.graph + ( size_t )f * frame_px;
    int* out = a.labels + ( size_t )f * frame_px;
    const uint32_t lane = threadIdx.x & 31u;
    static_assert( kTW % 32 == 0 && ( kTW * kTH ) % kThreads == 0, "a warp owns 32 consecutive pixels of one tile row" );
    for( int idx = threadIdx.x; idx < kTW * kTH; idx += kThreads )
    {
        const int ly = idx / kTW, lx = idx - ly * kTW, gx = x0 + lx, gy = y0 + ly;
        uint32_t node = 0u;
        if( gx < a.width && gy < a.height )
        {
            node = g[ ( size_t )gy * a.width + gx ];
            const bool right_ok = lx + 1 < kTW && gx + 1 < a.width, up_ok = ly + 1 < kTH && gy + 1 < a.height;
            if( !right_ok ) node &= ~( 16u | 4u );
            if( !up_ok ) node &= ~( 1u | 2u | 4u );
            if( lx == 0 ) node &= ~1u;
        }
        s_g[ idx ] = ( uint8_t )node;
        // run start inside the warp's 32 pixels: the lane after the last lane below me that is NOT linked to its right
        const uint32_t linked = __ballot_sync( 0xFFFFFFFFu, ( node & 16u ) != 0u );
        const uint32_t breaks = ~linked & ( ( 1u << lane ) - 1u );
        const int start = breaks ? 32 - __clz( ( int )breaks ) : 0;
        s_lab[ idx ] = idx - ( int )lane + start;
    }
    __syncthreads();
    // the unions that are left are queued first and then executed one per lane: a warp that unites as it goes spends
    // the time of its slowest lane three times per pixel (up, up-left, up-right), mostly on idle lanes
    for( int idx = threadIdx.x; idx < kTW * kTH; idx += kThreads )
    {
        const int lx = idx & ( kTW - 1 );
        const uint32_t node = s_g[ idx ];
        uint32_t want = ( ( node & 16u ) && lane == 31u ) ? 8u : 0u; // a run that continues into the next warp's pixels
        if( node & 7u )
        {
            const uint32_t left = lx > 0 ? s_g[ idx - 1 ] : 0u, right = lx + 1 < kTW ? s_g[ idx + 1 ] : 0u;
            const uint32_t up = s_g[ idx + kTW ], up_left = lx > 0 ? s_g[ idx + kTW - 1 ] : 0u; // (a link says the row above exists)
            const bool same_run_left = lane > 0u && ( left & 16u );                             // in the same run as the pixel to the left
            if( ( node & 2u ) && !( same_run_left && ( left & 2u ) && ( up_left & 16u ) ) ) want |= 2u;
            if( ( node & 1u ) && !( ( ( node & 2u ) && ( up_left & 16u ) ) || ( same_run_left && ( left & 2u ) ) ) ) want |= 1u;
            if( ( node & 4u ) && !( ( ( node & 2u ) && ( up & 16u ) ) || ( lane < 31u && ( node & 16u ) && ( right & 2u ) ) ) ) want |= 4u;
        }
        // slots for this warp's requests: one shared atomic per warp
        const int mine = __popc( want );
        int before = mine;
#pragma unroll
        for( int d = 1; d < 32; d <<= 1 )
        {
            const int v = __shfl_up_sync( 0xFFFFFFFFu, before, d );
            if( ( int )lane >= d ) before += v;
        }
        int base = 0;
        if( lane == 31u && before ) base = atomicAdd( &s_n, before );
        base = __shfl_sync( 0xFFFFFFFFu, base, 31 );
        int slot = base + before - mine;
        if( want & 8u ) s_req[ slot++ ] = ( uint32_t )idx << 16 | ( uint32_t )( idx + 1 );
        if( want & 2u ) s_req[ slot++ ] = ( uint32_t )idx << 16 | ( uint32_t )( idx + kTW );
        if( want & 1u ) s_req[ slot++ ] = ( uint32_t )idx << 16 | ( uint32_t )( idx + kTW - 1 );
        if( want & 4u ) s_req[ slot++ ] = ( uint32_t )idx << 16 | ( uint32_t )( idx + kTW + 1 );
    }
    __syncthreads();
    {
        const int n = s_n;
        for( int w = threadIdx.x; w < n; w += kThreads )
        {
            const uint32_t r = s_req[ w ];
            unite_halving( s_lab, ( int )( r >> 16 ), ( int )( r & 0xFFFFu ) );
        }
    }
    __syncthreads();
    for( int idx = threadIdx.x; idx < kTW * kTH; idx += kThreads )
    {
        int ly = idx / kTW, lx = idx - ly * kTW, gx = x0 + lx, gy = y0 + ly;
        if( gx >= a.width || gy >= a.height ) continue;
        int r = find_root( s_lab, idx );
        int ry = r / kTW, rx = r - ry * kTW;
        out[ ( size_t )gy * a.width + gx ] = ( y0 + ry ) * a.width + ( x0 + rx ); // local min index == global min index within a tile
    }
}

// Seam pass: one thread per pixel that sits on a tile seam (top row of a tile, first / last column of a tile),
// enumerated directly — the launch does not visit the other 90 % of the pixels.
__global__ void __launch_bounds__( kThreads ) cc_seam_kernel( LabelArgs a )
{
    const int f = blockIdx.y;
    const int tile_rows = ( a.height + kTH - 1 ) / kTH, tile_cols = ( a.width + kTW - 1 ) / kTW;
    const int n_top = tile_rows * a.width, n_side = 2 * tile_cols * a.height;
    const int t = blockIdx.x * kThreads + threadIdx.x;
    int gx, gy;
    if( t < n_top )
    {
        gy = ( t / a.width ) * kTH + kTH - 1;
        gx = t - ( t / a.width ) * a.width;
    }
    else if( t < n_top + n_side )
    {
        const int u = t - n_top, col = u / a.height;
        gy = u - col * a.height;
        gx = ( col >> 1 ) * kTW + ( ( col & 1 ) ? kTW - 1 : 0 );
        if( gy % kTH == kTH - 1 ) return; // (the top-row threads own the corner pixels)
    }
    else
        return;
    if( gx >= a.width || gy >= a.height ) return;
    const int lx = gx % kTW, ly = gy % kTH;
    const bool right_seam = lx == kTW - 1, left_seam = lx == 0, top_seam = ly == kTH - 1;
    const size_t frame_px = ( size_t )a.width * a.height;
    const uint8_t* g = a.graph + ( size_t )f * frame_px;
    int* lab = a.labels + ( size_t )f * frame_px;
    const int n = gy * a.width + gx;
    const uint32_t node = g[ n ];
    if( ( node & 16u ) && right_seam && gx + 1 < a.width ) unite( lab, n, n + 1 );
    if( gy + 1 < a.height )
    {
        if( ( node & 2u ) && top_seam ) unite( lab, n, n + a.width );
        if( ( node & 1u ) && gx > 0 && ( top_seam || left_seam ) ) unite( lab, n, n + a.width - 1 );
        if( ( node & 4u ) && gx + 1 < a.width && ( top_seam || right_seam ) ) unite( lab, n, n + a.width + 1 );
    }
}

__global__ void __launch_bounds__( kThreads ) cc_flatten_kernel( LabelArgs a )
{
    const size_t frame_px = ( size_t )a.width * a.height;
    const size_t n = ( size_t )blockIdx.x * kThreads + threadIdx.x;
    if( n >= frame_px ) return;
    int* lab = a.labels + ( size_t )blockIdx.y * frame_px;
    lab[ n ] = find_root( lab, ( int )n );
}

} // namespace

cudaError_t launch_cc_labels( const LabelArgs& a, cudaStream_t stream, int* n_launches )
{
    dim3 tiles( ( a.width + kTW - 1 ) / kTW, ( a.height + kTH - 1 ) / kTH, a.n_frames );
    cc_tile_kernel<<< tiles, kThreads, 0, stream >>>( a );
    const int seam_px = ( ( a.height + kTH - 1 ) / kTH ) * a.width + 2 * ( ( a.width + kTW - 1 ) / kTW ) * a.height;
    for( int f0 = 0; f0 < a.n_frames; f0 += 65535 ) // (grid.y limit)
    {
        LabelArgs part = a;
        part.graph = a.graph + ( size_t )f0 * a.width * a.height;
        part.labels = a.labels + ( size_t )f0 * a.width * a.height;
        part.n_frames = a.n_frames - f0 < 65535 ? a.n_frames - f0 : 65535;
        cc_seam_kernel<<< dim3( ( seam_px + kThreads - 1 ) / kThreads, part.n_frames ), kThreads, 0, stream >>>( part );
    }
    for( int f0 = 0; f0 < a.n_frames; f0 += 65535 )
    {
        LabelArgs part = a;
        part.labels = a.labels + ( size_t )f0 * a.width * a.height;
        part.n_frames = a.n_frames - f0 < 65535 ? a.n_frames - f0 : 65535;
        cc_flatten_kernel<<< dim3( ( unsigned )( ( ( size_t )a.width * a.height + kThreads - 1 ) / kThreads ), part.n_frames ), kThreads, 0, stream >>>( part );
    }
    if( n_launches ) *n_launches = 3;
    return cudaGetLastError();
}

} // namespace par
