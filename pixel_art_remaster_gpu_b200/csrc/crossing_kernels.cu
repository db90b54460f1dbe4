// Stage C: ambiguous diagonal crossings -> final similarity graph.
//
// Replaces ambiguous_cross_Kernel (kernel.cu:180-189): crossCheck_Heuristics
// (graph_functions.cu:1433-1518) -> processHeuristics2 (:761-849) -> checkValence2Edge/Vertex
// (:421-479) and the recursive calcVal2PathSize (:541-573).
//
// Design (B200): one CTA per 64x32-pixel tile.  The graph_aux bytes of the tile + 1-pixel halo are
// staged in shared memory (TMA bulk-tensor copy, zero fill outside the image; plain loads when the
// rows are not 16-byte multiples).  Every 2x2 block is decided ONCE (the reference decides it four
// times, once per corner pixel): the four local rules straight from the staged bytes; blocks that
// fall through to the curve-length rule are compacted into a shared-memory work list and walked
// warp-cooperatively — four lanes per block, one valence-2 chain each (the two chains of a diagonal
// only share a saturating sum, so they are independent), combined with shuffles.  Chains are
// followed through global memory (graph_aux is L2-resident: it was just written by stage A+B).
// Algorithmic HBM traffic: 1 B/px in + 1 B/px out.
#include "kernels.cuh"

namespace par {

namespace {

constexpr int kTW = 64, kTH = 32;
constexpr int kAW = kTW + 2, kAH = kTH + 2; // staged pixels (halo 1)
constexpr int kAuxOff = 15;                 // TMA needs a 16-byte aligned start: rows begin at column x0 - 16
constexpr int kAuxPitch = 96;               // 15 + 66 rounded up to 16 (TMA box row)
constexpr int kBW = kTW + 1, kBH = kTH + 1; // blocks decided per tile
constexpr int kDecPitch = 68;
constexpr int kThreads = 256;

enum : uint8_t { kNone = 0, kSlashDies = 1, kBackslashDies = 2, kPending = 3 };

struct __align__( 128 ) CrossSmem
{
    uint8_t aux[ kAH * kAuxPitch ];
    uint8_t dec[ kBH * kDecPitch ];
    uint16_t work[ kBH * kBW ];
    int n_work;
    uint64_t bar;
};

// links of a node other than edge e (the loops at graph_functions.cu:427-442, :462-469)
__device__ __forceinline__ int others( uint32_t node, int e ) { return __popc( node & 0xFFu & ~( 1u << e ) ); }

// length (<= 31) of the valence-2 chain leaving node n through the edges other than e
// (graph_functions.cu:541-573; the recursion only ever carries a counter, so it is a loop)
__device__ __forceinline__ int chain_length( const uint8_t* __restrict__ g, int n, int e, int width, int n_px )
{
    int len = 0;
#pragma unroll 1
    while( len < 31 && ( unsigned )n < ( unsigned )n_px ) // the bound only matters for a malformed caller-supplied graph
    {
        uint32_t m = ( uint32_t )__ldg( g + n ) & 0xFFu & ~( 1u << e );
        if( __popc( m ) != 1 ) break;
        int k = __ffs( ( int )m ) - 1;
        len++;
        n += edge_dj( k ) * width + edge_di( k ); // calc_index, graph_functions.cu:46-76
        e = 7 - k;                                // conected_edge, :26-28
    }
    return len;
}

template< bool kUseTma >
__global__ void __launch_bounds__( kThreads ) resolve_crossings_kernel( const __grid_constant__ CUtensorMap aux_map, CrossArgs a )
{
    __shared__ CrossSmem s;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH, f = blockIdx.z;
    const size_t frame_px = ( size_t )a.width * a.height;
    const uint8_t* aux_g = a.graph_aux + ( size_t )f * frame_px;

    if( tid == 0 ) s.n_work = 0;
    if( kUseTma )
    {
        if( tid == 0 )
        {
            mbar_init( &s.bar, 1 );
            fence_barrier_init();
        }
        __syncthreads();
        if( tid == 0 )
        {
            mbar_expect_tx( &s.bar, kAH * kAuxPitch );
            tma_load_3d( s.aux, &aux_map, &s.bar, x0 - 16, y0 - 1, f );
        }
        mbar_wait( &s.bar, 0 );
    }
    else
    {
        for( int idx = tid; idx < kAH * kAuxPitch; idx += kThreads )
        {
            int r = idx / kAuxPitch, c = idx - r * kAuxPitch;
            int gx = x0 - 16 + c, gy = y0 - 1 + r;
            uint8_t v = 0;
            if( gx >= 0 && gy >= 0 && gx < a.width && gy < a.height ) v = aux_g[ ( size_t )gy * a.width + gx ];
            s.aux[ idx ] = v;
        }
        __syncthreads();
    }

    // one decision per 2x2 block; block (c,r) has its lower-left pixel at staged position (c,r)
    for( int idx = tid; idx < kBH * kBW; idx += kThreads )
    {
        int r = idx / kBW, c = idx - r * kBW;
        const uint8_t* row0 = &s.aux[ r * kAuxPitch + kAuxOff + c ];
        uint32_t i1 = row0[ 0 ], i3 = row0[ 1 ];
        uint32_t i2 = row0[ kAuxPitch ], i4 = row0[ kAuxPitch + 1 ];
        uint8_t d = kNone;
        if( ( i1 & 4u ) && ( i2 & 128u ) && ( i3 & 1u ) && ( i4 & 32u ) ) // graph_functions.cu:1476-1512 guards
        {
            int o1 = others( i1, 2 ), o4 = others( i4, 5 ), o3 = others( i3, 0 ), o2 = others( i2, 7 );
            if( o1 == 1 && o4 == 1 ) d = kBackslashDies;                                      // :781-785
            else if( o3 == 1 && o2 == 1 ) d = kSlashDies;                                     // :791-795
            else if( ( o1 == 0 || o4 == 0 ) && o3 != 0 && o2 != 0 ) d = kBackslashDies;       // :798-802
            else if( o3 == 0 || ( o2 == 0 && o1 != 0 && o4 != 0 ) ) d = kSlashDies;           // :805-809
            else
            {
                d = kPending;
                s.work[ atomicAdd( &s.n_work, 1 ) ] = ( uint16_t )idx;
            }
        }
        s.dec[ r * kDecPitch + c ] = d;
    }
    __syncthreads();

    // curve-length rule (graph_functions.cu:815-839): 4 lanes per pending block, one chain each
    {
        const int n_work = s.n_work;
        const int lane = tid & 31, quad = tid >> 2, role = tid & 3;
        for( int base = 0; base < n_work; base += kThreads / 4 )
        {
            int w = base + quad;
            bool active = w < n_work;
            int len = 0, r = 0, c = 0;
            if( active )
            {
                int idx = s.work[ w ];
                r = idx / kBW;
                c = idx - r * kBW;
                int n1 = ( y0 - 1 + r ) * a.width + ( x0 - 1 + c ); // i1
                // role 0: (i1, edge 2)  1: (i4, edge 5)  2: (i3, edge 0)  3: (i2, edge 7)
                int n = n1 + ( role == 1 ? a.width + 1 : ( role == 2 ? 1 : ( role == 3 ? a.width : 0 ) ) );
                int e = role == 0 ? 2 : ( role == 1 ? 5 : ( role == 2 ? 0 : 7 ) );
                len = chain_length( aux_g, n, e, a.width, ( int )frame_px );
            }
            int pair = len + __shfl_xor_sync( 0xFFFFFFFFu, len, 1 ); // s (roles 0,1) or s2 (roles 2,3)
            pair = min( pair, 31 );
            int s_slash = __shfl_sync( 0xFFFFFFFFu, pair, ( lane & ~3 ) );
            int s_back = __shfl_sync( 0xFFFFFFFFu, pair, ( lane & ~3 ) + 2 );
            if( active && role == 0 ) s.dec[ r * kDecPitch + c ] = ( s_back < s_slash ) ? kBackslashDies : kSlashDies; // tie removes "/"
        }
    }
    __syncthreads();

    // final byte of 4 adjacent pixels per thread (the four SET_BITs of graph_functions.cu:1476-1518)
    uint8_t* out = a.graph + ( size_t )f * frame_px;
    const bool word_ok = ( a.width & 3 ) == 0;
    for( int idx = tid; idx < ( kTW / 4 ) * kTH; idx += kThreads )
    {
        int ly = idx / ( kTW / 4 ), lx = ( idx - ly * ( kTW / 4 ) ) * 4;
        int gx = x0 + lx, gy = y0 + ly;
        if( gx >= a.width || gy >= a.height ) continue;
        const uint8_t* up = &s.dec[ ( ly + 1 ) * kDecPitch + lx ];
        const uint8_t* dn = &s.dec[ ly * kDecPitch + lx ];
        const uint8_t* px = &s.aux[ ( ly + 1 ) * kAuxPitch + kAuxOff + lx + 1 ];
        uint32_t bytes = 0;
#pragma unroll
        for( int k = 0; k < 4; k++ )
        {
            uint32_t b = px[ k ];
            if( up[ k ] == kBackslashDies ) b &= ~1u;       // pixel is i3 of its up-left block
            if( up[ k + 1 ] == kSlashDies ) b &= ~4u;       // pixel is i1 of its up-right block
            if( dn[ k ] == kSlashDies ) b &= ~32u;          // pixel is i4 of its down-left block
            if( dn[ k + 1 ] == kBackslashDies ) b &= ~128u; // pixel is i2 of its down-right block
            bytes |= b << ( 8 * k );
        }
        size_t o = ( size_t )gy * a.width + gx;
        if( word_ok && gx + 3 < a.width )
            *reinterpret_cast< uint32_t* >( out + o ) = bytes;
        else
            for( int k = 0; k < 4 && gx + k < a.width; k++ ) out[ o + k ] = ( uint8_t )( bytes >> ( 8 * k ) );
    }
}

} // namespace

void resolve_crossings_tma_box( uint32_t box[ 3 ] )
{
    box[ 0 ] = kAuxPitch;
    box[ 1 ] = kAH;
    box[ 2 ] = 1;
}

cudaError_t launch_resolve_crossings( const CrossArgs& a, const CUtensorMap* aux_map, cudaStream_t stream )
{
    dim3 grid( ( a.width + kTW - 1 ) / kTW, ( a.height + kTH - 1 ) / kTH, a.n_frames );
    if( aux_map )
        resolve_crossings_kernel< true ><<< grid, kThreads, 0, stream >>>( *aux_map, a );
    else
    {
        CUtensorMap dummy;
        memset( &dummy, 0, sizeof( dummy ) );
        resolve_crossings_kernel< false ><<< grid, kThreads, 0, stream >>>( dummy, a );
    }
    return cudaGetLastError();
}

} // namespace par
