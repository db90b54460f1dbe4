// Shared device/host definitions of the B200 remaster path: graph-byte conventions, the exact
// packed-YUV conversion, TMA (cp.async.bulk.tensor) + mbarrier wrappers for sm_100a.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace par {

// graph bit e <-> neighbour offset (reference: graph_functions.cu:162-171, calc_index :46-76)
//   e        0      1      2      3      4      5      6      7
//   (di,dj) (-1,+1) (0,+1) (+1,+1) (-1,0) (+1,0) (-1,-1) (0,-1) (+1,-1)
// (two bits per edge, value + 1, packed: di+1 = 0,1,2,0,2,0,1,2 and dj+1 = 2,2,2,1,1,0,0,0)
__host__ __device__ __forceinline__ int edge_di( int e ) { return ( int )( ( 0x9224u >> ( 2 * e ) ) & 3u ) - 1; }
__host__ __device__ __forceinline__ int edge_dj( int e ) { return ( int )( ( 0x016Au >> ( 2 * e ) ) & 3u ) - 1; }

// thresholds on the packed fields (graph_functions.cu:14-19)
constexpr int kThrY = 0x00050000;
constexpr int kThrU = 0x00000700;
constexpr int kThrV = 0x00000006;

// Packed YUV word of the colour whose bytes in memory are (b0,b1,b2) = OpenCV (B,G,R).
// Restates graph_functions.cu:80-98 as the reference's DEVICE build evaluates it: Y is the double
// expression 0.299*b0 + 0.587*b1 + 0.114*b2 with nvcc's FMA contraction:
//     y = trunc( fma(0.114, b2, fma(0.299, b0, 0.587*b1)) )
// Evaluated here without FP64 on the common path: T = 299*b0 + 587*b1 + 114*b2 is the exact value
// times 1000; when T is not a multiple of 1000 the double expression is within 1e-12 of T/1000,
// far from any integer, so trunc() equals T/1000 in integer arithmetic.  Only exact multiples
// (e.g. every grey) need the real FMA chain, whose rounding decides between y and y-1.
// U and V are float products truncated toward zero (graph_functions.cu:93-94).
// The pieces (the graph kernel calls them one by one so that the rare case costs its four pixels ONE branch; T is passed
// in because the kernel forms it with two dot-product instructions):
//   T * M with M = ceil(2^32 / 1000) = 4294968:  high word = T / 1000 exactly for T <= 255000 (M * 1000 - 2^32 = 704 and
//   704 T < 2^32); low word = 704 (T / 1000) + (T mod 1000) M without wrapping (999 M + 704 * 255 < 2^32), i.e. below M
//   exactly when T is a multiple of 1000.
constexpr uint32_t kDiv1000 = 4294968u;
__host__ __device__ __forceinline__ int luma_div( int T )
{
#ifdef __CUDA_ARCH__
    return ( int )__umulhi( ( unsigned )T, kDiv1000 );
#else
    return ( int )( ( ( unsigned long long )( unsigned )T * kDiv1000 ) >> 32 );
#endif
}
// T is a non-zero multiple of 1000: rounding decides (black stays 0: the chain of three zeros is exact)
__host__ __device__ __forceinline__ bool luma_needs_chain( int T ) { return ( uint32_t )T * kDiv1000 - 1u < kDiv1000 - 1u; }
// y for such a colour, given y = T / 1000
__host__ __device__ __forceinline__ int luma_chain( int y, int b0, int b1, int b2 )
{
    if( b0 == b1 && b1 == b2 )
    {
        // greys: T = 1000 v exactly, and for 75 of the 256 values the rounded chain lands just below v (SURVEY
        // App. B-1).  Bit v of this 256-bit table says so; it is checked against the chain for every grey
        // (tests/test_golden.py) and, like every colour, against the reference's device code on the GPU.
        const int word = b0 >> 5;
        const uint32_t w = word == 0 ? 0x8c212116u : word == 1 ? 0xea528481u : word == 2 ? 0xc8324001u : word == 3 ? 0xd4449104u :
                           word == 4 ? 0x18200881u : word == 5 ? 0x50c08704u : word == 6 ? 0xc1030a18u : 0x53129890u;
        return y - ( int )( ( w >> ( b0 & 31 ) ) & 1u );
    }
#ifdef __CUDA_ARCH__
    return __double2int_rz( __fma_rn( 0.114, ( double )b2, __fma_rn( 0.299, ( double )b0, __dmul_rn( 0.587, ( double )b1 ) ) ) );
#else
    return ( int )__builtin_fma( 0.114, ( double )b2, __builtin_fma( 0.299, ( double )b0, 0.587 * ( double )b1 ) );
#endif
}
// U and V from y, the packed word (graph_functions.cu:93-97): float products truncated toward zero, added into the word as
// signed terms (a negative U or V borrows from the field above it, as in the reference).  d = b2 - y resp. b0 - y.
__host__ __device__ __forceinline__ uint32_t yuv_u_term( int d )
{
#ifdef __CUDA_ARCH__
    return ( uint32_t )( __float2int_rz( __fmul_rn( ( float )d, 0.492f ) ) * 256 );
#else
    return ( uint32_t )( ( int )( ( float )d * 0.492f ) * 256 );
#endif
}
__host__ __device__ __forceinline__ uint32_t yuv_v_term( int d )
{
#ifdef __CUDA_ARCH__
    return ( uint32_t )__float2int_rz( __fmul_rn( ( float )d, 0.877f ) );
#else
    return ( uint32_t )( int )( ( float )d * 0.877f );
#endif
}
__host__ __device__ __forceinline__ uint32_t yuv_pack( int y, int b0, int b2 ) { return ( uint32_t )( y << 16 ) + yuv_u_term( b2 - y ) + yuv_v_term( b0 - y ); }
__host__ __device__ __forceinline__ uint32_t yuv_word_t( int T, int b0, int b1, int b2 )
{
    int y = luma_div( T );
    if( luma_needs_chain( T ) ) y = luma_chain( y, b0, b1, b2 );
    return yuv_pack( y, b0, b2 );
}
__host__ __device__ __forceinline__ uint32_t yuv_word( int b0, int b1, int b2 ) { return yuv_word_t( 299 * b0 + 587 * b1 + 114 * b2, b0, b1, b2 ); }

// 1 when the two packed words are similar (graph_functions.cu:291-293 negated)
__host__ __device__ __forceinline__ bool yuv_similar( uint32_t p, uint32_t q )
{
    int dy = ( int )( ( p & 0x00FF0000u ) - ( q & 0x00FF0000u ) );
    int du = ( int )( ( p & 0x0000FF00u ) - ( q & 0x0000FF00u ) );
    int dv = ( int )( ( p & 0x000000FFu ) - ( q & 0x000000FFu ) );
    dy = dy < 0 ? -dy : dy;
    du = du < 0 ? -du : du;
    dv = dv < 0 ? -dv : dv;
    return dy <= kThrY && du <= kThrU && dv <= kThrV;
}

#ifdef __CUDACC__
// ---- mbarrier + TMA (sm_100a) ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32( const void* p ) { return ( uint32_t )__cvta_generic_to_shared( p ); }

__device__ __forceinline__ void mbar_init( uint64_t* bar, uint32_t count )
{
    asm volatile( "mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"( smem_u32( bar ) ), "r"( count ) : "memory" );
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" ); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile( "fence.proxy.async.shared::cta;" ::: "memory" ); }
__device__ __forceinline__ void mbar_expect_tx( uint64_t* bar, uint32_t bytes )
{
    asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"( smem_u32( bar ) ), "r"( bytes ) : "memory" );
}
__device__ __forceinline__ void mbar_wait( uint64_t* bar, uint32_t parity )
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"( smem_u32( bar ) ),
        "r"( parity )
        : "memory" );
}
// 3-D tiled bulk tensor load global -> shared, completion counted on `bar` (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_3d( void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2 )
{
    asm volatile( "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                      smem_u32( smem_dst ) ),
                  "l"( map ), "r"( smem_u32( bar ) ), "r"( c0 ), "r"( c1 ), "r"( c2 )
                  : "memory" );
}
__device__ __forceinline__ void tma_prefetch_desc( const CUtensorMap* map )
{
    asm volatile( "prefetch.tensormap [%0];" ::"l"( map ) : "memory" );
}
// streaming 128-bit store that does not allocate in L1 (output is written once, never re-read)
__device__ __forceinline__ void st_stream_v4( void* p, uint4 v )
{
    asm volatile( "st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"( p ), "r"( v.x ), "r"( v.y ), "r"( v.z ), "r"( v.w ) : "memory" );
}
#endif

} // namespace par
