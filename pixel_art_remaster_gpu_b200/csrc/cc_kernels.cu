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
//   (1) tile pass: one CTA per 64x32 tile builds the tile's forest in SHARED memory (a warp scans a band of rows with
//       the labels in registers: horizontal runs by ballot, labels inherited from the row below by shuffle, a union only
//       where a run joins two components; flatten) and writes tile-local roots as global indices;
//   (2) seam pass: one thread per seam pixel unions across the seams with global atomicMin;
//   (3) flatten pass: every pixel replaces its label by its root.
// Roots only ever point to smaller indices, so the root of a set is its minimum index.
// Algorithmic HBM traffic: 1 B/px in + 4 B/px out.
#include "kernels.cuh"

namespace par {

namespace {

#ifndef PAR_CC_TH
#define PAR_CC_TH 32
#endif
constexpr int kTW = 64, kTH = PAR_CC_TH, kThreads = 256;

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

// Tile pass (round 3).  A warp owns a band of 32 columns x 8 rows of the 64 x 32 tile and scans its rows bottom-up, a
// lane per column, labels in registers:
//  * horizontal runs from one ballot of the "linked to the right" bits (as before);
//  * a pixel INHERITS the label of a pixel it is linked to in the row below (that row's labels and node bytes travel
//    between the lanes by shuffle), and a run takes the minimum over its pixels (a segmented min-scan: five shuffle steps)
//    — or, with no link downwards, the index of its first pixel.  Labels only ever point to a smaller index of the same
//    component, so the array is a union-find forest from the start and almost every pixel already holds its root;
//  * a union is only needed where a pixel inherits a label that is not its run's (it joins two components of the rows below:
//    0.18 per pixel on the busy bench frames); those, and the links the warps cannot see in their registers — across the band
//    boundary (columns 31 | 32) and across the row groups (rows 7 | 8, 15 | 16, 23 | 24) — are appended to the warp's own list (a ballot
//    and two population counts per call, the count in a register; one list for the tile cost a shared-memory atomic and a
//    shuffle per call) and executed afterwards by the same warp with full lanes (executing them inside the row step cost every
//    warp the union's loops for three active lanes, row after row).
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

constexpr int kGroupRows = kTH / 4; // rows a warp scans: kTH / ( warps per band )
constexpr int kWarpReq = 3 * 32 * kGroupRows + 3 * 32 + 2 * kGroupRows; // capacity of a warp's union list

__global__ void __launch_bounds__( kThreads ) cc_tile_kernel( LabelArgs a )
{
    __shared__ int s_lab[ kTW * kTH ];
    // unions to make, a << 16 | b (tile-local indices), one list per warp (no atomics: the warp's count lives in a register):
    // at most three inherited labels per pixel, three links per pixel of the row group's top row, and two fixed slots per row
    // for the band boundary (written by the one lane that sits on it; 0 = nothing to do)
    __shared__ uint32_t s_req[ ( kThreads / 32 ) * kWarpReq ];
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH, f = blockIdx.z;
    const size_t frame_px = ( size_t )a.width * a.height;
    const uint8_t* g = a.graph + ( size_t )f * frame_px;
    int* out = a.labels + ( size_t )f * frame_px;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    static_assert( kTW == 64 && kThreads == 256 && kGroupRows * ( kThreads / 32 ) * 32 == kTW * kTH, "2 bands x 4 row groups" );
    const int band = warp & 1, lx = band * 32 + lane, row0 = ( warp >> 1 ) * kGroupRows;
    const int gx = x0 + lx;
    const uint32_t lanes_below = ( 1u << lane ) - 1u;
    uint32_t* req = s_req + warp * kWarpReq;
    uint32_t* req_edge = req + kWarpReq - 2 * kGroupRows;
    int n_req = 0; // (warp-uniform)
    auto request = [ & ]( bool want, int ia, int ib ) {
        const uint32_t votes = __ballot_sync( 0xFFFFFFFFu, want );
        if( want ) req[ n_req + __popc( votes & lanes_below ) ] = ( uint32_t )ia << 16 | ( uint32_t )ib;
        n_req += __popc( votes );
    };
    const bool edge_lane = band == 0 ? lane == 31 : lane == 0;
    uint32_t below = 0u; // label (tile-local index) << 8 | node byte of the pixel below this lane's; node 0 = no row below in this group
#pragma unroll
    for( int k = 0; k < kGroupRows; k++ )
    {
        const int ly = row0 + k, gy = y0 + ly, idx = ly * kTW + lx;
        uint32_t node = 0u;
        if( gx < a.width && gy < a.height )
        {
            node = g[ ( size_t )gy * a.width + gx ];
            const bool right_ok = lx + 1 < kTW && gx + 1 < a.width, up_ok = ly + 1 < kTH && gy + 1 < a.height;
            if( !right_ok ) node &= ~( 16u | 4u );
            if( !up_ok ) node &= ~( 1u | 2u | 4u );
            if( lx == 0 ) node &= ~1u;
        }
        // horizontal runs inside the band: the run starts after the last lane below me that is NOT linked to its right
        const uint32_t linked = __ballot_sync( 0xFFFFFFFFu, ( node & 16u ) != 0u && lane < 31 );
        const uint32_t breaks_below = ~linked & lanes_below;
        const int start = breaks_below ? 32 - __clz( ( int )breaks_below ) : 0;
        const int end = lane + __ffs( ( int )( ~linked >> lane ) ) - 1; // (bit 31 of ~linked is always set)
        // labels inherited from the row below: its pixel under me links up (bit 1), the one to the right up-left (bit 0), the
        // one to the left up-right (bit 2)
        const uint32_t bl = __shfl_up_sync( 0xFFFFFFFFu, below, 1 ), br = __shfl_down_sync( 0xFFFFFFFFu, below, 1 );
        constexpr int kNone = 0x7FFFFFFF;
        const int c0 = ( below & 2u ) ? ( int )( below >> 8 ) : kNone;
        const int c1 = ( lane > 0 && ( bl & 4u ) ) ? ( int )( bl >> 8 ) : kNone;
        const int c2 = ( lane < 31 && ( br & 1u ) ) ? ( int )( br >> 8 ) : kNone;
        const int cand = min( c0, min( c1, c2 ) );
        // minimum over the run (segmented inclusive min-scan, then the value of the run's last lane)
        int v = cand;
#pragma unroll
        for( int d = 1; d < 32; d <<= 1 )
        {
            const int t = __shfl_up_sync( 0xFFFFFFFFu, v, d );
            if( lane - d >= start ) v = min( v, t );
        }
        v = __shfl_sync( 0xFFFFFFFFu, v, end );
        const int label = min( v, ly * kTW + ( lx - lane + start ) ); // (no link downwards: the run's first pixel)
        s_lab[ idx ] = label;
        // a pixel that inherits another label than its run's joins two components of the rows below
        request( c0 != kNone && c0 != label, c0, label );
        request( c1 != kNone && c1 != label, c1, label );
        request( c2 != kNone && c2 != label, c2, label );
        // the links the scans do not see: upwards out of the row group (its top row; the tile's top row has none left), and
        // across the band boundary (columns 31 | 32)
        uint32_t edge_diag = 0u;
        if( k == kGroupRows - 1 )
        {
            request( ( node & 2u ) != 0u, idx, idx + kTW );
            request( ( node & 1u ) != 0u, idx, idx + kTW - 1 );
            request( ( node & 4u ) != 0u, idx, idx + kTW + 1 );
        }
        else if( band == 0 ? ( node & 4u ) != 0u : ( node & 1u ) != 0u )
            edge_diag = ( uint32_t )idx << 16 | ( uint32_t )( band == 0 ? idx + kTW + 1 : idx + kTW - 1 );
        if( edge_lane )
        {
            req_edge[ 2 * k ] = edge_diag;
            req_edge[ 2 * k + 1 ] = ( band == 0 && ( node & 16u ) ) ? ( uint32_t )idx << 16 | ( uint32_t )( idx + 1 ) : 0u;
        }
        below = ( uint32_t )label << 8 | node;
    }
    __syncthreads();
    // every warp executes its own list (the lists of a tile's eight warps are about equally long), then its edge slots
    for( int w = lane; w < n_req + 2 * kGroupRows; w += 32 )
    {
        const uint32_t r = w < n_req ? req[ w ] : req_edge[ w - n_req ];
        unite_halving( s_lab, ( int )( r >> 16 ), ( int )( r & 0xFFFFu ) );
    }
    __syncthreads();
    for( int idx = threadIdx.x; idx < kTW * kTH; idx += kThreads )
    {
        int ly = idx / kTW, lxx = idx - ly * kTW, px = x0 + lxx, py = y0 + ly;
        if( px >= a.width || py >= a.height ) continue;
        int r = find_root( s_lab, idx );
        int ry = r / kTW, rx = r - ry * kTW;
        out[ ( size_t )py * a.width + px ] = ( y0 + ry ) * a.width + ( x0 + rx ); // local min index == global min index within a tile
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

// Flatten pass: every pixel replaces its label by its root.  Four pixels per thread (one 128-bit load, four independent
// root walks in flight) when the frame size allows; most pixels already hold their root after the seam pass — only the
// labels of components that were joined across a seam change — so a label is only written back when it differs.
template< bool kVec4 >
__global__ void __launch_bounds__( kThreads ) cc_flatten_kernel( LabelArgs a )
{
    const size_t frame_px = ( size_t )a.width * a.height;
    int* lab = a.labels + ( size_t )blockIdx.y * frame_px;
    if( kVec4 )
    {
        const size_t q = ( size_t )blockIdx.x * kThreads + threadIdx.x;
        if( 4 * q >= frame_px ) return;
        const int4 l = *reinterpret_cast< const int4* >( lab + 4 * q );
        const int r0 = find_root( lab, l.x ), r1 = find_root( lab, l.y ), r2 = find_root( lab, l.z ), r3 = find_root( lab, l.w );
        if( r0 != l.x || r1 != l.y || r2 != l.z || r3 != l.w ) *reinterpret_cast< int4* >( lab + 4 * q ) = make_int4( r0, r1, r2, r3 );
    }
    else
    {
        const size_t n = ( size_t )blockIdx.x * kThreads + threadIdx.x;
        if( n >= frame_px ) return;
        const int l = lab[ n ], r = find_root( lab, l );
        if( r != l ) lab[ n ] = r;
    }
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
        const size_t px = ( size_t )a.width * a.height;
        if( px % 4 == 0 && ( reinterpret_cast< uintptr_t >( part.labels ) & 15u ) == 0 )
            cc_flatten_kernel< true ><<< dim3( ( unsigned )( ( px / 4 + kThreads - 1 ) / kThreads ), part.n_frames ), kThreads, 0, stream >>>( part );
        else
            cc_flatten_kernel< false ><<< dim3( ( unsigned )( ( px + kThreads - 1 ) / kThreads ), part.n_frames ), kThreads, 0, stream >>>( part );
    }
    if( n_launches ) *n_launches = 3;
    return cudaGetLastError();
}

} // namespace par
