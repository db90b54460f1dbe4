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

#ifndef PAR_K2_TH
#define PAR_K2_TH 32
#endif
constexpr int kTW = 64, kTH = PAR_K2_TH;
constexpr int kAW = kTW + 2, kAH = kTH + 2; // staged pixels (halo 1)
constexpr int kAuxOff = 15;                 // TMA needs a 16-byte aligned start: rows begin at column x0 - 16
constexpr int kAuxPitch = 96;               // 15 + 66 rounded up to 16 (TMA box row)
constexpr int kBW = kTW + 1, kBH = kTH + 1; // blocks decided per tile
constexpr int kDecPitch = 68;
// 128 threads per 64x32 tile, not 256: the kernel is bound by each tile's critical path (TMA round trip -> block scan ->
// barrier -> rules -> barrier -> chain walks through L1/L2 -> barrier -> output), not by instruction issue — a sparse form that
// removed the scan and the output pass (45 % of the instructions) was no faster (profiles/r4h_*) — so more tiles in flight per
// SM (15 instead of 8) is what helps: 0.408 -> 0.328 ms per 4096 frames (64 / 192 / 512 threads: 0.393 / 0.375 / 0.658; 64x16
// tiles with 128 / 64 threads: 0.398 / 0.378; profiles/r4i_*, r4j_*, r4k_*)
#ifndef PAR_K2_THREADS
#define PAR_K2_THREADS 128
#endif
constexpr int kThreads = PAR_K2_THREADS;

enum : uint8_t { kNone = 0, kSlashDies = 1, kBackslashDies = 2, kPending = 3 };

struct __align__( 128 ) CrossSmem
{
    uint8_t aux[ kAH * kAuxPitch ];
    alignas( 4 ) uint8_t dec[ kBH * kDecPitch ];
    uint16_t amb[ kBH * kDecPitch ]; // ambiguous blocks (index into dec)
    uint16_t work[ kBH * kBW ];      // ... of which need the curve-length walks (index r * kBW + c)
    int n_amb, n_work;
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

    if( tid == 0 )
    {
        s.n_amb = 0;
        s.n_work = 0;
    }
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

    // one decision per 2x2 block; block (c,r) has its lower-left pixel at staged position (c,r).  Four blocks per
    // step: a block is ambiguous iff both of its diagonals survived stage B — "/" is bit 2 of i1 and bit 5 of i4,
    // "\" is bit 7 of i2 and bit 0 of i3 (graph_functions.cu:1476-1512 guards) — which is one AND over byte-shifted
    // words.  Ambiguous blocks are queued so that the rules below run with full warps.
    static_assert( kAuxOff == 15 && kAuxPitch % 4 == 0 && kDecPitch % 4 == 0, "word layout of the block pass" );
    for( int idx = tid; idx < kBH * ( kDecPitch / 4 ); idx += kThreads )
    {
        const int r = idx / ( kDecPitch / 4 ), c0 = ( idx - r * ( kDecPitch / 4 ) ) * 4;
        const uint32_t* w0 = reinterpret_cast< const uint32_t* >( &s.aux[ r * kAuxPitch + 12 + c0 ] ); // pixel c0 is byte 3 of this word
        const uint32_t* w1 = reinterpret_cast< const uint32_t* >( &s.aux[ ( r + 1 ) * kAuxPitch + 12 + c0 ] );
        const uint32_t a1 = w0[ 1 ], b1 = w1[ 1 ];                                                 // pixels c0+1 .. c0+4 of both rows
        const uint32_t a0 = __byte_perm( w0[ 0 ], a1, 0x6543 ), b0 = __byte_perm( w1[ 0 ], b1, 0x6543 ); // pixels c0 .. c0+3
        uint32_t amb = ( a0 >> 2 ) & ( b0 >> 7 ) & a1 & ( b1 >> 5 ) & 0x01010101u;
        if( c0 + 3 >= kBW ) amb &= 0xFFFFFFFFu >> ( 8 * ( c0 + 4 - kBW ) ); // the last group holds fewer than four blocks
        *reinterpret_cast< uint32_t* >( &s.dec[ r * kDecPitch + c0 ] ) = 0u;
        if( amb )
        {
            const int n = __popc( amb );
            int slot = atomicAdd( &s.n_amb, n );
#pragma unroll 1
            for( ; amb; amb &= amb - 1u ) s.amb[ slot++ ] = ( uint16_t )( r * kDecPitch + c0 + ( ( __ffs( ( int )amb ) - 1 ) >> 3 ) );
        }
    }
    __syncthreads();
    // the four local rules of processHeuristics2 (graph_functions.cu:781-809); what falls through is queued for the walks
    {
        const int n_amb = s.n_amb;
        for( int w = tid; w < n_amb; w += kThreads )
        {
            const int at = s.amb[ w ], r = at / kDecPitch, c = at - r * kDecPitch;
            const uint8_t* row0 = &s.aux[ r * kAuxPitch + kAuxOff + c ];
            const uint32_t i1 = row0[ 0 ], i3 = row0[ 1 ], i2 = row0[ kAuxPitch ], i4 = row0[ kAuxPitch + 1 ];
            const int o1 = others( i1, 2 ), o4 = others( i4, 5 ), o3 = others( i3, 0 ), o2 = others( i2, 7 );
            uint8_t d;
            if( o1 == 1 && o4 == 1 ) d = kBackslashDies;                                      // :781-785
            else if( o3 == 1 && o2 == 1 ) d = kSlashDies;                                     // :791-795
            else if( ( o1 == 0 || o4 == 0 ) && o3 != 0 && o2 != 0 ) d = kBackslashDies;       // :798-802
            else if( o3 == 0 || ( o2 == 0 && o1 != 0 && o4 != 0 ) ) d = kSlashDies;           // :805-809
            else
            {
                d = kPending;
                s.work[ atomicAdd( &s.n_work, 1 ) ] = ( uint16_t )( r * kBW + c );
            }
            s.dec[ at ] = d;
        }
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
        // decisions are 0 (none), 1 ("/" dies) or 2 ("\" dies) by now: bit 0 / bit 1 of the decision byte, tested on all
        // four pixels at once.  Up-left / down-left blocks of pixel lx+k are decision columns lx+k, up-right / down-right lx+k+1.
        const uint32_t* upw = reinterpret_cast< const uint32_t* >( up );
        const uint32_t* dnw = reinterpret_cast< const uint32_t* >( dn );
        const uint32_t ul = upw[ 0 ], dl = dnw[ 0 ];
        const uint32_t ur = __byte_perm( ul, upw[ 1 ], 0x4321 ), dr = __byte_perm( dl, dnw[ 1 ], 0x4321 );
        const uint32_t dead = ( ( ul >> 1 ) & 0x01010101u )    // "\" of the up-left block dies: pixel is its i3, bit 0
                              | ( ( ur << 2 ) & 0x04040404u )  // "/" of the up-right block dies: pixel is its i1, bit 2
                              | ( ( dl << 5 ) & 0x20202020u )  // "/" of the down-left block dies: pixel is its i4, bit 5
                              | ( ( dr << 6 ) & 0x80808080u ); // "\" of the down-right block dies: pixel is its i2, bit 7
        const uint32_t bytes = *reinterpret_cast< const uint32_t* >( px ) & ~dead; // (staged column of pixel lx is 16 + lx: word aligned)
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
