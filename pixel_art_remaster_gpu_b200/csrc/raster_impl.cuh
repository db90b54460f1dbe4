// Implementation of the raster kernel (stages D + E + direct rasterization), shared by the translation units that
// instantiate it per output format (raster_kernels.cu: RGBA8 + the table builders, raster_bgr8.cu, raster_index8.cu).
// See raster_kernels.cu for the design notes.
#pragma once
#include "kernels.cuh"
#include "polygon.cuh"

namespace par {

namespace {

constexpr int kThreads = 256;
constexpr int kMaxVerts = 16;    // 8 hull vertices, each replaced by at most two

template< int S >
struct Cfg
{
    // A subdivided cell normally reaches at most 22/64 pixel outside its own square (hull vertices reach
    // 1/4; blending with a neighbour's cut point adds a little).  H = number of samples per side within
    // that reach: the smallest H whose next sample offset (2H+1)/(2S) exceeds 22/64.  A cell that reaches
    // further ("wide": only possible when the reference's getPointIndex falls back to vertex 0,
    // subdivision_functions.cu:537) is flagged and handled exactly by the slow path of the resolve step.
    static constexpr int H = 11 * S >= 16 ? ( 11 * S - 16 ) / 32 + 1 : 0;
    static constexpr int R = S + 2 * H;                   // samples per axis covered by a cell's mask
    // integer units: when S divides 32 a vertex (multiple of 1/64 px) and a sample (odd multiple of
    // 1/(2S) px) are both integers in units of 1/64 px; otherwise everything is scaled by 2S.
    static constexpr bool POW2 = ( S & ( S - 1 ) ) == 0 && S <= 32;
    static constexpr int VM = POW2 ? 1 : 2 * S;           // vertex multiplier
    static constexpr int SSP = POW2 ? 64 / S : 128;       // distance between samples
    static constexpr int SSP_LOG2 = SSP == 128 ? 7 : ( SSP == 64 ? 6 : ( SSP == 32 ? 5 : ( SSP == 16 ? 4 : ( SSP == 8 ? 3 : ( SSP == 4 ? 2 : 1 ) ) ) ) );
    static constexpr int S_FIRST = SSP / 2 - H * SSP;     // coordinate of mask sample 0 (local output index -H)
    static constexpr int SQUARE = 64 * VM;                // the cell's own square is [0, SQUARE]
    static constexpr int REACH = H * SSP + SSP / 2;       // offset of the first sample NOT covered by the mask
    static constexpr bool PACK = R * R <= 63;             // whole mask in one 64-bit word (bit 63 = wide flag)
    static constexpr int MW = PACK ? 2 : R;               // 32-bit words per mask
    static constexpr int TW = 32, TH = S <= 4 ? 32 : 16;
    static constexpr int CW = TW + 2, CH = TH + 2;        // cells whose masks are needed (halo 1)
    static constexpr int KW = TW + 4, KH = TH + 4;        // cells whose keys and colours are needed (halo 2)
    static constexpr int GOFF = 16;                       // staged rows begin at column x0 - 16 (TMA: 16-byte aligned start)
    static constexpr int GP = ( GOFF + TW + 3 + 15 ) / 16 * 16; // staged row pitch (TMA box row), covers x0-3 .. x0+TW+2
    // staged BGR row (TMA box row): bytes 3*x0-16 .. ; 16 bytes more than needed, because a 128-byte pitch would put every
    // row on the same banks and the staging pass (a warp reads four rows at once) would take 4-way conflicts
    static constexpr int RAWP = ( 16 + 3 * ( TW + 2 ) + 15 ) / 16 * 16 + ( ( 16 + 3 * ( TW + 2 ) + 15 ) / 16 * 16 % 128 == 0 ? 16 : 0 );
    static constexpr int RAWOFF = 16 - 6;                 // byte offset of pixel x0-2 in a staged row
    static constexpr int NC = CW * CH;
    static constexpr uint32_t FULL = ( 1u << S ) - 1u;
    static constexpr uint32_t ROWMASK = ( 1u << R ) - 1u;
    static constexpr uint32_t WIDE = PACK ? 0x40000000u : 0x80000000u; // flag in the high word (PACK, window form) / in row 0 (rows)
    // PACK masks live in shared memory and in the tables in WINDOW form (to_window below): four 16-bit fields, each
    // already positioned on the S x S output pixels (bit S*y + x) of the cell whose resolver will read it:
    //   lo[0,16)  F0 the cell's own S x S samples
    //   lo[16,32) F1 bit S*y+S-1 <- halo sample (-1, y): it lies in the LEFT neighbour's square, last column there;
    //                bit S*y     <- halo sample (S, y): first column of the RIGHT neighbour's square
    //   hi[0,16)  F2 bit S*(S-1)+x <- halo sample (x, -1): top row of the square BELOW; bit x <- (x, S): bottom row ABOVE
    //   hi[16,32) F3 the four corner halo samples: bit S*S-1 <- (-1,-1), bit S*(S-1) <- (S,-1), bit S-1 <- (-1,S),
    //                bit 0 <- (S,S); bit 14 = WIDE
    static constexpr uint32_t ALL = PACK ? ( 1u << ( S * S ) ) - 1u : 0u;
    static constexpr uint32_t M_COL0 = S == 1 ? 1u : ( S == 2 ? 0x5u : ( S == 3 ? 0x49u : 0x1111u ) ); // bits S*y, y < S
    static constexpr uint32_t M_LEFTCOL = M_COL0;                       // my column 0   <- F1 of the cell to the left
    static constexpr uint32_t M_RIGHTCOL = M_COL0 << ( S - 1 );         // my column S-1 <- F1 of the cell to the right
    static constexpr uint32_t M_BOTROW = ( 1u << S ) - 1u;              // my row 0      <- F2 of the cell below
    static constexpr uint32_t M_TOPROW = M_BOTROW << ( S * ( S - 1 ) ); // my row S-1    <- F2 of the cell above
    // Shared memory carve-up (bytes).  The mask array has two lives: until the staging pass has turned them into cell
    // words and colours it holds the staged graph and BGR rows (the TMA destinations); the masks are written after
    // that.  Keeping the CTA under 25.4 KB lets five of them share an SM with 124 KB left as L1 for the tables — the kernel
    // is sensitive to both.
    // two 32-bit words per cell (smooth_table.h): x = 12-bit key | IDs of directions 4..7 << 12, y = IDs of directions 0..3;
    // row r of the tile is KW x-words followed by KW y-words (consecutive cells on consecutive banks, and every
    // neighbour's word of either kind within a signed byte of word offsets)
    static constexpr int KP = 2 * KW;                      // words per row of cell words
    static constexpr int off_keys = 0;
    static constexpr int off_col = off_keys + KW * KH * 8;
    static constexpr int off_mask = ( off_col + KW * KH * 4 + 127 ) / 128 * 128;
    static constexpr int off_graph = off_mask;
    static constexpr int off_raw = off_graph + ( KH * GP + 127 ) / 128 * 128;
    static constexpr int sz_stage = off_raw - off_graph + KH * RAWP;
    static_assert( sz_stage <= NC * MW * 4, "the staged rows fit in the mask array" );
    static constexpr int off_bar = ( off_mask + NC * MW * 4 + 15 ) / 16 * 16;
    static constexpr int off_nwork = off_bar + 16; // four counters / flags
    static constexpr int smem_bytes = off_nwork + 16 + 2 * 4 * ( ( NC + kThreads - 1 ) / kThreads ) * ( kThreads / 32 ); // counters, then the two ballot arrays of the mask pass
};

// number of sample columns c in [0,N) with F - c*G > 0, i.e. clamp(ceil(F/G), 0, N), G > 0
template< int S, int N >
__device__ __forceinline__ int columns_left_of( int F, int G, float rcpG )
{
    if( Cfg< S >::POW2 )
    {
        // (F - 1/2)/G is never an integer and stays >= 1/(2G) >= 4e-5 away from one, far more than the error of
        // the approximate reciprocal on a quotient of magnitude <= N, so the floor is exact (G <= 64*154).
        const int q = __float2int_rd( ( ( float )F - 0.5f ) * rcpG ) + 1;
        return min( max( q, 0 ), N );
    }
    int cnt = 0;
#pragma unroll
    for( int c = 0; c < N; c++ ) cnt += ( F > c * G ) ? 1 : 0;
    return cnt;
}

// Toggle, on every sample row an edge crosses, the samples strictly left of the crossing (even-odd rule with
// the (+eps, -eps^2) displacement).  Sample (c, r) sits at (ox + SSP c, oy + SSP r).  Row r is crossed iff
// min(y0,y1) < y_r <= max(y0,y1), which is the (y0 < y) != (y1 < y) rule.
template< int S, int N, class Toggle >
__device__ __forceinline__ void cover_edge( int ox, int oy, int x0, int y0, int x1, int y1, Toggle& toggle )
{
    typedef Cfg< S > C;
    const int dy = y1 - y0;
    if( dy == 0 ) return;
    const int dx = x1 - x0;
    const int r_lo = max( ( ( min( y0, y1 ) - oy ) >> C::SSP_LOG2 ) + 1, 0 );
    const int r_hi = min( ( max( y0, y1 ) - oy ) >> C::SSP_LOG2, N - 1 );
    const int G = C::SSP * ( dy < 0 ? -dy : dy );                 // F decreases by G per sample column
    const int step = dy < 0 ? -C::SSP * dx : C::SSP * dx;          // F increases by step per sample row
    int F = dx * ( oy - y0 ) - ( ox - x0 ) * dy;                   // edge function at sample (0, 0) ...
    F = ( dy < 0 ? -F : F ) + r_lo * step;                         // ... oriented, at row r_lo
    const float rcpG = __fdividef( 1.0f, ( float )G );
#pragma unroll 1
    for( int r = r_lo; r <= r_hi; r++, F += step ) toggle( r, columns_left_of< S, N >( F, G, rcpG ) );
}

template< int R >
struct PackedToggle // R x R mask in one 64-bit register, row r at bits [R r, R r + R)
{
    uint64_t m;
    __device__ __forceinline__ void operator()( int r, int cnt ) { m ^= ( uint64_t )( ( 1u << cnt ) - 1u ) << ( R * r ); }
};
struct RowToggle // rows in memory (shared memory on the fast path), row r at rows[r * stride]
{
    uint32_t* rows;
    int stride;
    __device__ __forceinline__ void operator()( int r, int cnt ) { rows[ r * stride ] ^= ( 1u << cnt ) - 1u; }
};

// R-stride packed mask (bit R*r + c = sample (c - H, r - H)) -> window form (see Cfg)
template< int S >
__device__ __forceinline__ uint2 to_window( uint64_t m )
{
    typedef Cfg< S > C;
    constexpr int R = C::R, H = C::H;
    uint32_t f0 = 0u, f1 = 0u, f2 = 0u, f3 = 0u;
#pragma unroll
    for( int y = 0; y < S; y++ )
    {
        const uint32_t row = ( uint32_t )( m >> ( R * ( y + H ) ) );
        f0 |= ( ( row >> H ) & C::FULL ) << ( S * y );
        if( H > 0 )
        {
            f1 |= ( row & 1u ) << ( S * y + S - 1 );
            f1 |= ( ( row >> ( S + H ) ) & 1u ) << ( S * y );
        }
    }
    if( H > 0 )
    {
        const uint32_t bot = ( uint32_t )m, top = ( uint32_t )( m >> ( R * ( S + H ) ) );
        f2 = ( ( ( bot >> H ) & C::FULL ) << ( S * ( S - 1 ) ) ) | ( ( top >> H ) & C::FULL );
        f3 = ( ( bot & 1u ) << ( S * S - 1 ) ) | ( ( ( bot >> ( S + H ) ) & 1u ) << ( S * ( S - 1 ) ) ) | ( ( top & 1u ) << ( S - 1 ) ) |
             ( ( top >> ( S + H ) ) & 1u );
    }
    return make_uint2( f0 | f1 << 16, f2 | f3 << 16 );
}

// Slots of build_cell_polygon in a small buffer (slot k at buf[k * stride]), packed as (x64 + 64) << 8 | (y64 + 64)
struct PackedSlots
{
    uint16_t* buf;
    int stride;
    __device__ __forceinline__ void put( int slot, int x64, int y64 ) { buf[ slot * stride ] = ( uint16_t )( ( ( x64 + 64 ) << 8 ) | ( y64 + 64 ) ); }
};

// Coverage of the polygon held in `slots` over an N x N sample window whose sample (0,0) sits at (ox, oy).
// Returns the polygon's coordinate range through lo/hi (vertex units, i.e. already multiplied by VM).
template< int S, int N, class Toggle >
__device__ __forceinline__ void cover_polygon( const uint16_t* buf, int stride, CellPoly poly, int ox, int oy, Toggle& toggle, int& lo, int& hi )
{
    constexpr int VM = Cfg< S >::VM;
    uint32_t v = buf[ 0 ];
    int x0 = ( ( int )( v >> 8 ) - 64 ) * VM, y0 = ( ( int )( v & 255u ) - 64 ) * VM;
    const int fx = x0, fy = y0;
    lo = min( x0, y0 );
    hi = max( x0, y0 );
    const int last = 2 * poly.n - 1 - ( int )( ( ~poly.two >> ( poly.n - 1 ) ) & 1u ); // last occupied slot
    int slot = 0;
#pragma unroll 1
    while( true )
    {
        // next occupied slot: 2t+1 follows 2t only when hull vertex t was split
        const bool done = slot == last;
        slot = ( ( slot & 1 ) || !( ( poly.two >> ( slot >> 1 ) ) & 1u ) ) ? ( slot | 1 ) + 1 : slot + 1;
        int x1 = fx, y1 = fy;
        if( !done )
        {
            v = buf[ slot * stride ];
            x1 = ( ( int )( v >> 8 ) - 64 ) * VM;
            y1 = ( ( int )( v & 255u ) - 64 ) * VM;
            lo = min( lo, min( x1, y1 ) );
            hi = max( hi, max( x1, y1 ) );
        }
        cover_edge< S, N >( ox, oy, x0, y0, x1, y1, toggle );
        if( done ) break;
        x0 = x1;
        y0 = y1;
    }
}

// 1 when a, b, c are not all equal, else 0: one three-input logic operation ((a ^ b) | (b ^ c)) and one minimum.  (As
// inline PTX: written in C the compiler turns it back into two comparisons and a select.)
__device__ __forceinline__ uint32_t not_one_colour( uint32_t a, uint32_t b, uint32_t c )
{
    uint32_t t;
    asm( "{ .reg .b32 x; lop3.b32 x, %1, %2, %3, 0x7E; min.u32 %0, x, 1; }" : "=r"( t ) : "r"( a ), "r"( b ), "r"( c ) );
    return t;
}

// a rows-form table entry (four 64-bit words, 32-byte aligned) by ONE 256-bit read-only load (sm_100: LDG.256): a gather costs
// the LSU a wavefront per instruction and line, so four 64-bit loads of the same entry cost four times what this does
__device__ __forceinline__ void ldg_entry4( const uint64_t* e, uint64_t* m )
{
    asm( "ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"( m[ 0 ] ), "=l"( m[ 1 ] ), "=l"( m[ 2 ] ), "=l"( m[ 3 ] ) : "l"( e ) );
}
// ... XORed into m when `flag` is not zero (no branch); returns word 0 of the entry (0 when not loaded)
__device__ __forceinline__ uint64_t ldg_entry4_xor_if( const uint64_t* e, uint32_t flag, uint64_t* m )
{
    uint64_t v0 = 0ull, v1 = 0ull, v2 = 0ull, v3 = 0ull;
    asm( "{ .reg .pred q; setp.ne.u32 q, %5, 0; @q ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4]; }"
         : "+l"( v0 ), "+l"( v1 ), "+l"( v2 ), "+l"( v3 )
         : "l"( e ), "r"( flag ) );
    m[ 0 ] ^= v0;
    m[ 1 ] ^= v1;
    m[ 2 ] ^= v2;
    m[ 3 ] ^= v3;
    return v0;
}

// 64-bit read-only load when `flag` is not zero, 0 otherwise (no branch: the compiler will not speculate a load on its own)
__device__ __forceinline__ uint64_t ldg_u64_if( const uint64_t* p, uint32_t flag )
{
    uint64_t v = 0ull;
    asm( "{ .reg .pred q; setp.ne.u32 q, %2, 0; @q ld.global.nc.u64 %0, [%1]; }" : "+l"( v ) : "l"( p ), "r"( flag ) );
    return v;
}

// Output formats (par_out_format): how the S words of a row segment leave the kernel.  In INDEX8 mode a "colour word" carries
// the palette index of its colour in its top byte (where RGBA8 has the constant alpha), see the staging pass.
constexpr int kFmtRgba8 = 0, kFmtBgr8 = 1, kFmtIndex8 = 2;
template< int FMT >
struct Fmt
{
    static constexpr int BPP = FMT == kFmtRgba8 ? 4 : ( FMT == kFmtBgr8 ? 3 : 1 );       // bytes per output pixel
    static constexpr uint32_t BACKGROUND = FMT == kFmtIndex8 ? 0u : 0xFF000000u;       // black (main.cpp:260); palette entry 0 is black
};

// one output row segment of a source pixel: N pixels (colour words px[0..N)), widest stores the alignment allows
template< int N, int FMT >
__device__ __forceinline__ void store_row( uint8_t* dst, const uint32_t* px, bool align32 = false )
{
    if constexpr( FMT == kFmtRgba8 )
    {
        if( N == 8 && align32 )
        {
            // 8x: a lane's row segment is 32 bytes — one 256-bit store (sm_100: STG.256), so that a warp's store instruction
            // writes 1024 contiguous bytes; as two 128-bit stores each instruction touches the same eight lines half-filled
            // (twice the LSU wavefronts)
            asm volatile( "st.global.cs.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"( dst ), "r"( px[ 0 ] ), "r"( px[ 1 % N ] ), "r"( px[ 2 % N ] ),
                          "r"( px[ 3 % N ] ), "r"( px[ 4 % N ] ), "r"( px[ 5 % N ] ), "r"( px[ 6 % N ] ), "r"( px[ 7 % N ] )
                          : "memory" );
        }
        else if( N % 4 == 0 )
        {
#pragma unroll
            for( int k = 0; k < N; k += 4 ) st_stream_v4( dst + 4 * k, make_uint4( px[ k ], px[ k + 1 ], px[ k + 2 ], px[ k + 3 ] ) );
        }
        else if( N % 2 == 0 )
        {
#pragma unroll
            for( int k = 0; k < N; k += 2 ) *reinterpret_cast< uint2* >( dst + 4 * k ) = make_uint2( px[ k ], px[ k + 1 ] );
        }
        else
        {
#pragma unroll
            for( int k = 0; k < N; k++ ) *reinterpret_cast< uint32_t* >( dst + 4 * k ) = px[ k ];
        }
    }
    else if constexpr( FMT == kFmtBgr8 )
    {
        // B, G, R per pixel (Image.cpp:64-71 writes 3-channel images); a word is R | G << 8 | B << 16 | A << 24
        if constexpr( N % 4 == 0 )
        {
#pragma unroll
            for( int k = 0; k < N; k += 4 )
            {
                uint32_t* d = reinterpret_cast< uint32_t* >( dst + 3 * k );
                d[ 0 ] = __byte_perm( px[ k ], px[ k + 1 ], 0x6012 );     // B0 G0 R0 B1
                d[ 1 ] = __byte_perm( px[ k + 1 ], px[ k + 2 ], 0x5601 ); // G1 R1 B2 G2
                d[ 2 ] = __byte_perm( px[ k + 2 ], px[ k + 3 ], 0x4560 ); // R2 B3 G3 R3
            }
        }
        else if constexpr( N % 2 == 0 )
        {
#pragma unroll
            for( int k = 0; k < N; k += 2 )
            {
                uint16_t* d = reinterpret_cast< uint16_t* >( dst + 3 * k );
                d[ 0 ] = ( uint16_t )__byte_perm( px[ k ], 0u, 0x4412 );         // B0 G0
                d[ 1 ] = ( uint16_t )__byte_perm( px[ k ], px[ k + 1 ], 0x4460 ); // R0 B1
                d[ 2 ] = ( uint16_t )__byte_perm( px[ k + 1 ], 0u, 0x4401 );     // G1 R1
            }
        }
        else
        {
#pragma unroll
            for( int k = 0; k < N; k++ )
            {
                dst[ 3 * k ] = ( uint8_t )( px[ k ] >> 16 );
                dst[ 3 * k + 1 ] = ( uint8_t )( px[ k ] >> 8 );
                dst[ 3 * k + 2 ] = ( uint8_t )px[ k ];
            }
        }
    }
    else
    {
        // palette indices: the top byte of every word
        if constexpr( N % 4 == 0 )
        {
#pragma unroll
            for( int k = 0; k < N; k += 4 )
                *reinterpret_cast< uint32_t* >( dst + k ) =
                    __byte_perm( __byte_perm( px[ k ], px[ k + 1 ], 0x7373 ), __byte_perm( px[ k + 2 ], px[ k + 3 ], 0x7373 ), 0x5410 );
        }
        else if constexpr( N % 2 == 0 )
        {
#pragma unroll
            for( int k = 0; k < N; k += 2 ) *reinterpret_cast< uint16_t* >( dst + k ) = ( uint16_t )__byte_perm( px[ k ], px[ k + 1 ], 0x4473 );
        }
        else
        {
#pragma unroll
            for( int k = 0; k < N; k++ ) dst[ k ] = ( uint8_t )( px[ k ] >> 24 );
        }
    }
}

// Anti-aliased output (the reference's GL_MULTISAMPLE toggle, simpleVBO.cpp:238-253, as ordered-grid
// supersampling): the kernel samples at S = A x the output scale and averages A x A samples per output pixel
// right in the resolve step — the supersampled image never exists in memory.  RGBA sums on two pairs of 16-bit lanes.
struct ColourSum
{
    uint32_t rb = 0u, ga = 0u;
    __device__ __forceinline__ void add( uint32_t w )
    {
        rb += w & 0x00FF00FFu;
        ga += ( w >> 8 ) & 0x00FF00FFu;
    }
    template< int N > // mean of N = 4 or 16 samples per channel, rounded to nearest (halves up)
    __device__ __forceinline__ uint32_t mean() const
    {
        constexpr int SH = N == 4 ? 2 : 4;
        constexpr uint32_t BIAS = ( N / 2 ) * 0x00010001u;
        return ( ( ( rb + BIAS ) >> SH ) & 0x00FF00FFu ) | ( ( ( ( ga + BIAS ) >> SH ) & 0x00FF00FFu ) << 8 );
    }
};

// S x S sample colours of one source pixel (row-major) -> its (S/A) x (S/A) output pixels, written as whole row segments
template< int S, int A, int FMT >
__device__ __forceinline__ void store_cell( uint8_t* dst, ptrdiff_t row_step, const uint32_t* px )
{
    constexpr int O = S / A;
    if( A == 1 )
    {
#pragma unroll
        for( int b = 0; b < S; b++ ) store_row< S, FMT >( dst + ( ptrdiff_t )b * row_step, px + S * b );
    }
    else
    {
#pragma unroll
        for( int oy = 0; oy < O; oy++ )
        {
            uint32_t row[ O ];
#pragma unroll
            for( int ox = 0; ox < O; ox++ )
            {
                ColourSum sum;
#pragma unroll
                for( int j = 0; j < A; j++ )
#pragma unroll
                    for( int i = 0; i < A; i++ ) sum.add( px[ ( oy * A + j ) * S + ox * A + i ] );
                row[ ox ] = sum.template mean< A * A >();
            }
            store_row< O, FMT >( dst + ( ptrdiff_t )oy * row_step, row );
        }
    }
}

template< int S >
struct TileEnv
{
    const uint32_t* keys; // KH rows of cell words (Cfg::KP words each; key = low 12 bits of the x-word), origin (x0-2, y0-2)
    const uint32_t* cols; // KW x KH RGBA words, same origin; rows at or above the image height hold colour 0
    int x0, y0;
    FlatImage img;
    __device__ __forceinline__ uint32_t key( int i, int j ) const { return keys[ ( j - y0 + 2 ) * Cfg< S >::KP + ( i - x0 + 2 ) ] & 0xFFFu; }
    // checkTJunction (subdivision_functions.cu:170-242).  Away from the first/last column the flat byte
    // offsets the reference uses (idx +- widthstep +- 3) are exactly the 2-D neighbours, which are staged
    // in shared memory; at i = 0 / W-1 they wrap to the adjacent rows, so those cells take the flat path.
    __device__ __forceinline__ bool guard( int i, int j ) const { return img.guard( i, j ); }
    __device__ __forceinline__ bool keep_corner( int i, int j, Q2 p ) const
    {
        if( i < 1 || i > img.width - 2 ) return img.keep_corner( i, j, p );
        const bool px0 = p.x == 0, px1 = p.x == 4, py0 = p.y == 0, py1 = p.y == 4;
        if( !( ( px0 || px1 ) && ( py0 || py1 ) ) ) return false;
        const int sx = px1 ? 1 : -1, sy = py1 ? 1 : -1; // the corner's quadrant
        const uint32_t* c = cols + ( j - y0 + 2 ) * Cfg< S >::KW + ( i - x0 + 2 );
        const uint32_t side = c[ sx ], diag = c[ sy * Cfg< S >::KW + sx ], vert = c[ sy * Cfg< S >::KW ];
        return side != diag || diag != vert; // the three other pixels around the corner are not one colour
    }
};

// slow path of the resolve step: coverage of the S x S samples of target cell (ti,tj) by the polygon of
// cell (ci,cj) = (ti+di, tj+dj), recomputed from scratch (exact for any reach < 1 pixel)
template< int S >
__device__ __noinline__ void window_coverage( const uint32_t* keys, const uint32_t* cols, int x0, int y0, const uint8_t* frame, int width, int height,
                                              int widthstep, const CellRecord* rec, int ci, int cj, int di, int dj, bool subdivide, uint32_t* win )
{
    typedef Cfg< S > C;
    // (everything by value: a reference to the caller's TileEnv would force it into local memory on the fast path too)
    TileEnv< S > env;
    env.keys = keys;
    env.cols = cols;
    env.x0 = x0;
    env.y0 = y0;
    env.img.frame = frame;
    env.img.width = width;
    env.img.height = height;
    env.img.widthstep = widthstep;
    const CellTablePtrs tab{ rec };
    uint16_t verts[ kMaxVerts ];
    PackedSlots slots{ verts, 1 };
    const CellPoly poly = build_cell_polygon( env, tab, ci, cj, env.key( ci, cj ), subdivide, slots );
    for( int r = 0; r < S; r++ ) win[ r ] = 0u;
    RowToggle tg{ win, 1 };
    int lo, hi;
    cover_polygon< S, S >( verts, 1, poly, C::SSP / 2 - di * C::SQUARE, C::SSP / 2 - dj * C::SQUARE, tg, lo, hi );
}

// Exact resolve of a whole tile: every candidate's coverage of every pixel recomputed from its polygon.  Only runs
// for tiles that contain a cell reaching beyond its mask, or under PAR_FLAG_DEBUG_WIDE.
template< int S, int A, int FMT >
__device__ __noinline__ void resolve_tile_exact( const uint32_t* keys, const uint32_t* cols, int x0, int y0, const uint8_t* frame, int width, int height,
                                                 int widthstep, const CellRecord* rec, bool subdivide, uint8_t* out, bool flip )
{
    typedef Cfg< S > C;
    constexpr int O = S / A;
    const size_t out_w = ( size_t )width * O, out_h = ( size_t )height * O;
    for( int idx = threadIdx.x; idx < C::TW * C::TH; idx += kThreads )
    {
        const int ly = idx / C::TW, lx = idx - ly * C::TW, gx = x0 + lx, gy = y0 + ly;
        if( gx >= width || gy >= height ) continue;
        const uint32_t* col = cols + ( ly + 2 ) * C::KW + ( lx + 2 );
        uint32_t px[ S * S ], rem[ S ];
        for( int k = 0; k < S * S; k++ ) px[ k ] = Fmt< FMT >::BACKGROUND; // background (main.cpp:260)
        for( int b = 0; b < S; b++ ) rem[ b ] = C::FULL;
        // candidates in DESCENDING node index
        for( int dj = 1; dj >= -1; dj-- )
            for( int di = 1; di >= -1; di-- )
            {
                const int ci = gx + di, cj = gy + dj;
                if( ci < 0 || cj < 0 || ci >= width || cj >= height ) continue;
                uint32_t win[ S ];
                window_coverage< S >( keys, cols, x0, y0, frame, width, height, widthstep, rec, ci, cj, di, dj, subdivide, win );
                const uint32_t cw = col[ dj * C::KW + di ];
                for( int b = 0; b < S; b++ )
                {
                    const uint32_t take = win[ b ] & rem[ b ];
                    rem[ b ] &= ~take;
                    for( int k = 0; k < S; k++ )
                        if( ( take >> k ) & 1u ) px[ S * b + k ] = cw;
                }
            }
        constexpr int BPP = Fmt< FMT >::BPP;
        const ptrdiff_t row_step = flip ? -( ptrdiff_t )( out_w * BPP ) : ( ptrdiff_t )( out_w * BPP );
        store_cell< S, A, FMT >( out + ( ( flip ? out_h - 1 - ( size_t )gy * O : ( size_t )gy * O ) * out_w + ( size_t )gx * O ) * BPP, row_step, px );
    }
}

// General path of the mask pass: one thread per queued cell (a bit per cell in `workbits`, a word per warp and round of
// the mask pass); polygon -> per-thread vertex buffer -> edge loop.  Only runs for the few cells the smoothing tables
// cannot express (or all smoothed cells under PAR_FLAG_NO_SMOOTH_TABLES).  Out of line, with its vertex buffer in local
// memory: the common path pays neither its registers nor its code nor shared memory.
template< int S >
__device__ __noinline__ void geometric_cells( const uint32_t* keys, const uint32_t* cols, uint32_t* s_mask, const uint32_t* workbits, int* s_nwork, int x0, int y0,
                                              const uint8_t* frame, int width, int height, int widthstep, const CellRecord* rec, uint32_t force_wide )
{
    typedef Cfg< S > C;
    TileEnv< S > env;
    env.keys = keys;
    env.cols = cols;
    env.x0 = x0;
    env.y0 = y0;
    env.img.frame = frame;
    env.img.width = width;
    env.img.height = height;
    env.img.widthstep = widthstep;
    const CellTablePtrs tab{ rec };
    uint16_t verts[ kMaxVerts ];
    for( int idx = threadIdx.x; idx < C::NC; idx += kThreads )
    {
        if( !( ( workbits[ idx >> 5 ] >> ( idx & 31 ) ) & 1u ) ) continue;
        int cy = idx / C::CW, cx = idx - cy * C::CW;
        int gx = x0 - 1 + cx, gy = y0 - 1 + cy;
        PackedSlots slots{ verts, 1 };
        const CellPoly poly = build_cell_polygon( env, tab, gx, gy, keys[ ( cy + 1 ) * C::KP + cx + 1 ] & 0xFFFu, true, slots );
        int lo, hi;
        if constexpr( C::PACK )
        {
            PackedToggle< C::R > tg{ 0ull };
            cover_polygon< S, C::R >( verts, 1, poly, C::S_FIRST, C::S_FIRST, tg, lo, hi );
            // reach check: every sample outside the mask must be strictly outside the polygon's bounding box
            const uint32_t wide = ( lo <= -C::REACH || hi >= C::SQUARE + C::REACH ) ? C::WIDE : force_wide;
            uint2 wm = to_window< S >( tg.m );
            wm.y |= wide;
            s_mask[ idx ] = wm.x;
            s_mask[ C::NC + idx ] = wm.y;
            if( wide ) s_nwork[ 2 ] = 1;
        }
        else
        {
#pragma unroll
            for( int r = 0; r < C::R; r++ ) s_mask[ r * C::NC + idx ] = 0u;
            RowToggle tg{ s_mask + idx, C::NC };
            cover_polygon< S, C::R >( verts, 1, poly, C::S_FIRST, C::S_FIRST, tg, lo, hi );
            const uint32_t wide = ( lo <= -C::REACH || hi >= C::SQUARE + C::REACH ) ? C::WIDE : force_wide;
            s_mask[ idx ] |= wide;
            if( wide ) s_nwork[ 2 ] = 1;
        }
    }
}

// mask of a cell whose polygon is its plain hull, for every key: the per-scale table the raster kernel copies from
struct NoEnv
{
    __device__ __forceinline__ uint32_t key( int, int ) const { return 0u; }
    __device__ __forceinline__ bool guard( int, int ) const { return true; }
    __device__ __forceinline__ bool keep_corner( int, int, Q2 ) const { return true; }
};

template< int S >
__global__ void build_mask_lut_kernel( CellTablePtrs tab, uint32_t* lut )
{
    typedef Cfg< S > C;
    const int key = blockIdx.x * blockDim.x + threadIdx.x;
    if( key >= kCellKeys ) return;
    uint16_t verts[ kMaxVerts ];
    PackedSlots slots{ verts, 1 };
    NoEnv env;
    const CellPoly poly = build_cell_polygon( env, tab, 0, 0, ( uint32_t )key, false, slots );
    int lo, hi;
    if constexpr( C::PACK )
    {
        PackedToggle< C::R > tg{ 0ull };
        cover_polygon< S, C::R >( verts, 1, poly, C::S_FIRST, C::S_FIRST, tg, lo, hi );
        const uint2 w = to_window< S >( tg.m );
        lut[ 2 * key ] = w.x;
        lut[ 2 * key + 1 ] = w.y;
    }
    else
    {
        uint32_t rows[ C::R ];
        for( int r = 0; r < C::R; r++ ) rows[ r ] = 0u;
        RowToggle tg{ rows, 1 };
        cover_polygon< S, C::R >( verts, 1, poly, C::S_FIRST, C::S_FIRST, tg, lo, hi );
        // rows form: the table holds entries (Entry< S >: four rows per 64-bit word, four words per key)
        uint64_t* e = reinterpret_cast< uint64_t* >( lut ) + ( size_t )key * 4;
        for( int w = 0; w < 4; w++ ) e[ w ] = 0ull;
        for( int r = 0; r < C::R; r++ ) e[ r >> 2 ] |= ( uint64_t )( rows[ r ] & 0x7FFFu ) << ( 16 * ( r & 3 ) );
    }
}

// ---- smoothing tables (smooth_table.h) -------------------------------------------------------------
// Table entry = the coverage mask as 64-bit words: PACK -> one word in window form (wide flag = Cfg::WIDE of the
// high half); rows -> R 16-bit rows, four per word, wide flag in bit 15 of row 0.
template< int S >
struct Entry
{
    // (rows form: always four words — 32 bytes, so that an entry is two aligned 128-bit loads; at 5x .. 7x the fourth word is padding)
    static constexpr int EW = Cfg< S >::PACK ? 1 : 4;
    static_assert( Cfg< S >::PACK || ( Cfg< S >::R * 2 + 7 ) / 8 <= 4, "sixteen rows at most" );
    static constexpr uint64_t FLAG = Cfg< S >::PACK ? ( ( uint64_t )Cfg< S >::WIDE << 32 ) : ( 1ull << 15 );
    // link table only: the neighbour's record does not fit the class (its end / start vertex is not the blended vertex, or it
    // has no edge in that direction) — the cell takes the geometric path.  A bit no mask uses: bit 13 of the corner field
    // (window form) / bit 14 of row 0 (rows are at most 14 samples wide).
    static constexpr uint64_t MISMATCH = Cfg< S >::PACK ? ( 1ull << 61 ) : ( 1ull << 14 );
    static_assert( Cfg< S >::PACK || Cfg< S >::R <= 14, "bit 14 of a row is free" );
};

// coverage of the closed polygon (xs[k], ys[k]), k < m (1/64 px, cell-local), as a table entry
template< int S >
__device__ void cover_to_entry( const int* xs, const int* ys, int m, uint64_t* out )
{
    typedef Cfg< S > C;
    int lo = 1 << 30, hi = -( 1 << 30 );
    for( int k = 0; k < m; k++ )
    {
        lo = min( lo, min( xs[ k ], ys[ k ] ) * C::VM );
        hi = max( hi, max( xs[ k ], ys[ k ] ) * C::VM );
    }
    const bool wide = lo <= -C::REACH || hi >= C::SQUARE + C::REACH;
    if constexpr( C::PACK )
    {
        PackedToggle< C::R > tg{ 0ull };
        for( int k = 0; k < m; k++ )
        {
            const int k1 = k + 1 == m ? 0 : k + 1;
            cover_edge< S, C::R >( C::S_FIRST, C::S_FIRST, xs[ k ] * C::VM, ys[ k ] * C::VM, xs[ k1 ] * C::VM, ys[ k1 ] * C::VM, tg );
        }
        const uint2 w = to_window< S >( tg.m );
        out[ 0 ] = ( ( uint64_t )w.y << 32 | w.x ) | ( wide ? Entry< S >::FLAG : 0ull );
    }
    else
    {
        uint32_t rows[ C::R ];
        for( int r = 0; r < C::R; r++ ) rows[ r ] = 0u;
        RowToggle tg{ rows, 1 };
        for( int k = 0; k < m; k++ )
        {
            const int k1 = k + 1 == m ? 0 : k + 1;
            cover_edge< S, C::R >( C::S_FIRST, C::S_FIRST, xs[ k ] * C::VM, ys[ k ] * C::VM, xs[ k1 ] * C::VM, ys[ k1 ] * C::VM, tg );
        }
        for( int w = 0; w < Entry< S >::EW; w++ ) out[ w ] = 0ull;
        for( int r = 0; r < C::R; r++ ) out[ r >> 2 ] |= ( uint64_t )( rows[ r ] & 0x7FFFu ) << ( 16 * ( r & 3 ) );
        if( wide ) out[ 0 ] |= Entry< S >::FLAG;
    }
}

// CUT[key][kept]: the hull with its cut vertices replaced by R(prev), Q(cur) (subdivision_functions.cu:583-598),
// except the square corners (0,0) (1,0) (1,1) (0,1) whose bit in `kept` is set
template< int S >
__global__ void build_cut_table_kernel( CellTablePtrs tab, uint64_t* cut )
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if( id >= kCellKeys * 16 ) return;
    const uint32_t key = ( uint32_t )id >> 4, kept = ( uint32_t )id & 15u;
    uint64_t h, info;
    load_hull( tab, key, h, info );
    const int n = hull_count( info );
    const VertexClasses cls = classify_vertices( info );
    const uint32_t cv = hull_corner_vertices( info );
    uint32_t keptv = 0u;
    for( int c = 0; c < 4; c++ )
    {
        const uint32_t v = ( cv >> ( 4 * c ) ) & 15u;
        if( ( ( kept >> c ) & 1u ) && v != 15u ) keptv |= 1u << v;
    }
    int xs[ kMaxVerts ], ys[ kMaxVerts ], m = 0;
    for( int t = 0; t < n; t++ )
    {
        const Q2 p = hull_vertex( h, t );
        if( ( ( cls.cut >> t ) & 1u ) && !( ( keptv >> t ) & 1u ) )
        {
            cut_toward( p, hull_vertex( h, t == 0 ? n - 1 : t - 1 ), xs[ m ], ys[ m ] );
            m++;
            cut_toward( p, hull_vertex( h, t + 1 == n ? 0 : t + 1 ), xs[ m ], ys[ m ] );
            m++;
        }
        else
        {
            xs[ m ] = 16 * p.x;
            ys[ m ] = 16 * p.y;
            m++;
        }
    }
    cover_to_entry< S >( xs, ys, m, cut + ( size_t )id * Entry< S >::EW );
}

// the point with a given code (inverse of point_code, cell_table.h), quarter pixels
__device__ __forceinline__ Q2 point_of_code( int code )
{
    Q2 q{ 0, 0 };
    int seen = 0;
    for( int pos = 0; pos < 49; pos++ )
        if( ( kValidPoints >> pos ) & 1ull )
        {
            if( seen == code )
            {
                q.x = pos % 7 - 1;
                q.y = pos / 7 - 1;
            }
            seen++;
        }
    return q;
}

// LINK[class][ID]: the loop between the hull path and the smoothed path around one shared edge
// (subdivision_functions.cu:603-647 for the two blended vertices, see smooth_table.h), for the neighbour record the ID
// names; MISMATCH when that record's end / start vertex is not the class's blended vertex (getPointIndex's fallback, :527-538)
template< int S >
__global__ void build_link_table_kernel( const LinkClass* classes, uint64_t* link )
{
    const LinkClass& c = classes[ blockIdx.x ];
    const int id = threadIdx.x;
    uint64_t* entry = link + ( size_t )( c.block * ( uint32_t )kNbrIds + id ) * Entry< S >::EW;
    const uint32_t r = c.nrec[ id ];
    if( id == 0 || r == 0xFFFFu || ( c.hasA && ( ( r >> 8 ) & 15u ) != c.codeA ) || ( c.hasB && ( ( r >> 12 ) & 15u ) != c.codeB ) )
    {
        for( int w = 0; w < Entry< S >::EW; w++ ) entry[ w ] = w == 0 ? Entry< S >::MISMATCH : 0ull;
        return;
    }
    const int a = c.after[ r & 3u ], b = c.before[ ( r >> 4 ) & 3u ]; // (a class with one blended end ignores the other rank)
    const int di = edge_di( c.e ), dj = edge_dj( c.e );
    const Q2 P0{ c.px[ 0 ], c.py[ 0 ] }, P1{ c.px[ 1 ], c.py[ 1 ] }, P2{ c.px[ 2 ], c.py[ 2 ] }, P3{ c.px[ 3 ], c.py[ 3 ] };
    int xs[ 6 ], ys[ 6 ], m = 0;
    if( c.hasA )
    {
        int rx, ry, qx, qy;
        cut_toward( P1, P0, rx, ry );                                                   // own R on the border edge arriving at P1
        cut_toward( Q2{ P1.x - 4 * di, P1.y - 4 * dj }, point_of_code( a ), qx, qy );  // neighbour's Q on the edge leaving P1
        xs[ m ] = rx;
        ys[ m ] = ry;
        m++;
        xs[ m ] = ( rx + qx + 64 * di ) >> 1;
        ys[ m ] = ( ry + qy + 64 * dj ) >> 1;
        m++;
    }
    else
    {
        xs[ m ] = 16 * P1.x;
        ys[ m ] = 16 * P1.y;
        m++;
    }
    if( c.hasB )
    {
        int qx, qy, rx, ry;
        cut_toward( P2, P3, qx, qy );                                                   // own Q on the border edge leaving P2
        cut_toward( Q2{ P2.x - 4 * di, P2.y - 4 * dj }, point_of_code( b ), rx, ry );  // neighbour's R on the edge arriving at P2
        xs[ m ] = ( qx + rx + 64 * di ) >> 1;
        ys[ m ] = ( qy + ry + 64 * dj ) >> 1;
        m++;
        xs[ m ] = qx;
        ys[ m ] = qy;
        m++;
    }
    xs[ m ] = 16 * P2.x;
    ys[ m ] = 16 * P2.y;
    m++;
    if( c.hasA )
    {
        xs[ m ] = 16 * P1.x;
        ys[ m ] = 16 * P1.y;
        m++;
    }
    cover_to_entry< S >( xs, ys, m, entry );
}

// The blocks shared by the classes that only say which IDs do not fit them: entry = MISMATCH outside lo .. lo + span, else 0.
template< int S >
__global__ void build_range_blocks_kernel( const uint8_t* lo, const uint8_t* span, uint32_t first_block, uint64_t* link )
{
    const uint32_t id = threadIdx.x;
    uint64_t* entry = link + ( size_t )( ( first_block + blockIdx.x ) * ( uint32_t )kNbrIds + id ) * Entry< S >::EW;
    for( int w = 0; w < Entry< S >::EW; w++ ) entry[ w ] = 0ull;
    if( id - lo[ blockIdx.x ] > span[ blockIdx.x ] ) entry[ 0 ] = Entry< S >::MISMATCH;
}

// Which classes have something to XOR at this scale — a mask bit or the WIDE flag in some entry — and so keep their own
// block; the others are served by the shared block of their ID range.  block_of[ own block ] = the block to use.
template< int S >
__global__ void choose_class_blocks_kernel( const LinkClass* classes, int n_classes, const uint64_t* link, uint8_t* block_of )
{
    const int ci = blockIdx.x * blockDim.x + threadIdx.x;
    if( ci == 0 ) block_of[ 0 ] = 0;
    if( ci >= n_classes ) return;
    const LinkClass& c = classes[ ci ];
    bool any = c.canon == 0; // (no shared block for its range: keeps its own)
    for( int id = 0; id < kNbrIds; id++ )
        for( int w = 0; w < Entry< S >::EW; w++ )
        {
            const uint64_t v = link[ ( size_t )( c.block * ( uint32_t )kNbrIds + id ) * Entry< S >::EW + w ];
            any = any || ( w == 0 ? ( v & ~Entry< S >::MISMATCH ) : v ) != 0ull;
        }
    block_of[ c.block ] = ( uint8_t )( any ? c.block : c.canon );
}

// The kernel's descriptors at this scale (smooth_table.h): the key's descriptors with their blocks as chosen above, those
// whose class keeps its own block first (stable), then the ones on shared blocks; slots 0, 1 -> head01, slots 2, 3 -> head23.
__global__ void build_head_tables_kernel( const uint4* desc, const uint8_t* block_of, uint2* head01, uint2* head23 )
{
    const int key = blockIdx.x * blockDim.x + threadIdx.x;
    if( key >= kCellKeys ) return;
    const uint4 in = desc[ key ];
    const uint32_t d[ 4 ] = { in.x, in.y, in.z, in.w };
    uint32_t out[ 4 ] = { 0u, 0u, 0u, 0u };
    int n = 0;
    for( int pass = 0; pass < 2; pass++ )
        for( int k = 0; k < 4; k++ )
        {
            if( !( d[ k ] & kDescUsed ) ) continue;
            const uint32_t own = ( d[ k ] >> 13 ) & 255u, use = block_of[ own ];
#ifdef PAR_WHATIF_DROPSHARED
            if( use != own ) continue; // (what-if: as if a descriptor on a shared block could never mismatch)
#endif
            if( ( use == own ) == ( pass == 0 ) ) out[ n++ ] = ( d[ k ] & ~( 255u << 13 ) & ~( kDescSlow | kDescMore ) ) | use << 13;
        }
    out[ 0 ] |= ( d[ 0 ] & kDescSlow ) | ( out[ 2 ] ? kDescMore : 0u );
    head01[ key ] = make_uint2( out[ 0 ], out[ 1 ] );
    head23[ key ] = make_uint2( out[ 2 ], out[ 3 ] );
}

// NS link descriptors of a smoothed cell (smooth_table.h: one word each, ready to use), without branches: XORs their LINK
// entries into m and ORs word 0 of the entries into `flags` (WIDE and MISMATCH travel there).  The neighbour's record is
// named by a 5-bit ID that sits in the neighbour's cell word: one shared-memory load at the descriptor's offset from the
// cell's own word, one shift by the descriptor's amount — no neighbour record is gathered, and a neighbour without an edge
// in that direction has ID 0, whose entry is MISMATCH.
template< int S, int NS >
__device__ __forceinline__ void link_slots( const SmoothTablePtrs& st, const uint32_t* words_at_cell, const uint32_t* desc, uint64_t* m, uint64_t& flags )
{
    typedef Cfg< S > C;
    typedef Entry< S > E;
    static_assert( C::KP == kHeadRowWords && kNbrIds == 32, "descriptor word offsets; block << 5 | id" );
    const uint32_t* base = words_at_cell - ( kHeadRowWords + 1 ); // (the offsets are biased)
#pragma unroll
    for( int k = 0; k < NS; k++ )
    {
        const uint32_t d = desc[ k ];
        const uint32_t used = d & kDescUsed;
        const uint32_t id = ( base[ d & 255u ] >> ( ( d >> 8 ) & 31u ) ) & 31u;
        const uint64_t* le = st.link + ( size_t )( ( ( d >> 8 ) & ( 255u << 5 ) ) | id ) * E::EW;
        if constexpr( E::EW == 4 )
            flags |= ldg_entry4_xor_if( le, used, m );
        else
        {
#pragma unroll
            for( int w = 0; w < E::EW; w++ )
            {
                const uint64_t v = ldg_u64_if( le + w, used ); // (an unused slot loads nothing)
                if( w == 0 ) flags |= v;
                m[ w ] ^= v;
            }
        }
    }
}

// (Round 2, measured and dropped — the numbers are under profiles/r2b_*, r2c_*, r2d_*, r2h_*, r2j_*:
//  * a persistent kernel, one CTA of four 256-thread groups per SM, with the first two link descriptors, the neighbour
//    records and the LINK entries in 120 KB of shared memory (bank-aware layouts): the LSU data pipe went from 87 % to 64 %,
//    but with 28 KB of L1 left for the CUT entries (hit rate 64 % -> 20 %) and 32 warps per SM instead of 40 the kernel
//    became latency-bound: 5.45 ms against 4.49 ms per 4096 frames;
//  * a per-scale 16-byte "head" record per key (first two descriptors + CUT[key][0] in one gather, so that most cells skip
//    the CUT gather) and colour loads predicated on the corners the hull has a cut vertex at: one gather in six and most
//    of the eight shared loads per cell less, no change in time (4.42 ms) — the kernel is bound by instruction issue and
//    load latency together, not by LSU wavefronts alone;
//  * the corner flags computed after the descriptor load instead of before it: 1 % slower;
//  * for the scales above 4, the window form on four 64-bit fields per cell (nine 64-bit shared loads per pixel instead of
//    a word per candidate and sample row, an early exit for pixels that are all their own colour): bit-exact, and slower at
//    every scale (8x: 2.32 against 2.14 ms per 512 frames, 6x: 1.89 / 1.41, 5x: 1.95 / 1.29, 7x: 3.86 / 2.03) — 64-bit shifts
//    and selects cost more instructions than the shared loads they replace;
//  * the output rows through shared memory and out by bulk copies (cp.async.bulk shared -> global, SASS UBLKCP.G.S; two
//    512-byte row buffers per warp, one copy per warp and output row) instead of STG.128, so that the 64 bytes per pixel do
//    not cross the LSU's 32-byte-per-clock path to the crossbar (the stores are a third of the kernel's LSU data-pipe
//    cycles): bit-exact, 4.93 against 4.37 ms (and 3.20 against 2.60 with subdivision off) — a wait, two warp barriers, a
//    proxy fence and a command flush per 512 bytes cost more than the path they relieve;
//  * the same with ONE bulk-tensor store per warp and tile row (cp.async.bulk.tensor.3d shared -> global, SASS UTMASTG.3D: a
//    2 KB band of 4 output rows x 128 pixels per warp, the tensor unit clipping at the image edge): bit-exact, 5.13 against
//    4.37 ms (2.86 against 2.60 with subdivision off) — 16 KB more shared memory per CTA means four CTAs per SM and 92 KB
//    instead of 124 KB of L1 for the tables.)
// Mask of a smoothed cell from the tables, FIRST pass: the CUT entry and the first two link descriptors (nine cells in
// ten have no more).  `more` is set when the key has a third descriptor: the caller marks the cell for
// smooth_lookup_more.  Returns 0 when the mask is complete as it stands (nearly always); otherwise the rare cases as flag
// bits, with m[0] still carrying whatever the XOR left in the flag positions: E::MISMATCH — a blended vertex is not a vertex
// of the neighbour's hull (the reference's getPointIndex fallback), or the key is not in the tables: the caller takes the
// geometric path; E::FLAG — some piece reaches beyond the mask (wide).
template< int S >
__device__ __forceinline__ uint64_t smooth_lookup( const SmoothTablePtrs& st, const uint32_t* mask_lut, const uint32_t* keys_at_cell, uint32_t key, uint2 head,
                                                   uint32_t cflags, uint64_t* m, bool& more )
{
    typedef Cfg< S > C;
    typedef Entry< S > E;
    uint64_t flags = 0ull;
    if( cflags & 16u ) // checkTJunction's early exit keeps every cut vertex: the plain hull
    {
        if constexpr( C::PACK )
        {
            const uint2 v = __ldg( reinterpret_cast< const uint2* >( mask_lut ) + key );
            m[ 0 ] = ( uint64_t )v.y << 32 | v.x;
        }
        else
        {
            ldg_entry4( reinterpret_cast< const uint64_t* >( mask_lut ) + ( size_t )key * E::EW, m ); // (rows form: the table holds entries)
        }
    }
    else
    {
        // (the bits of corners without a cut vertex do not matter, and the 16 entries of a key are one 128-byte line at s <= 4:
        // not masking them off lets the gather start before the descriptors have arrived — 1.2 % on the bench frames)
        const uint64_t* e = st.cut + ( size_t )( key * 16u + ( cflags & 15u ) ) * E::EW;
        if constexpr( E::EW == 4 )
            ldg_entry4( e, m );
        else
        {
#pragma unroll
            for( int w = 0; w < E::EW; w++ ) m[ w ] = __ldg( e + w );
        }
    }
    more = ( int32_t )head.x < 0; // kDescMore
#ifndef PAR_WHATIF_NOLINKS
    const uint32_t desc[ 2 ] = { head.x, head.y };
    link_slots< S, 2 >( st, keys_at_cell, desc, m, flags ); // (a key that always takes the geometric path has no descriptors)
#else
    more = false;
#endif
    // the flag bits of word 0 were XORed along with the masks: take them from the OR (the CUT entry has none)
    if constexpr( C::PACK )
    {
        // (both flags live in the high word; kDescSlow is MISMATCH's bit there, and bit 30 of a descriptor is never set)
        static_assert( ( E::MISMATCH >> 32 ) == kDescSlow && ( E::FLAG >> 32 ) == ( 1u << 30 ) && kDescMore == ( 1u << 31 ), "flag positions" );
        const uint32_t rare_hi = ( ( uint32_t )( flags >> 32 ) | head.x ) & ( uint32_t )( ( E::FLAG | E::MISMATCH ) >> 32 );
        return ( uint64_t )rare_hi << 32;
    }
    return ( flags & ( E::FLAG | E::MISMATCH ) ) | ( ( head.x & kDescSlow ) ? E::MISMATCH : 0ull );
}

// SECOND pass, for the cells whose key has three or four link descriptors: the XOR of the remaining LINK entries; returns
// the rare cases as above.
template< int S >
__device__ __forceinline__ uint64_t smooth_lookup_more( const SmoothTablePtrs& st, const uint32_t* keys_at_cell, uint32_t key, uint64_t* m )
{
    typedef Entry< S > E;
    const uint2 head2 = __ldg( st.head2 + key );
    const uint32_t desc[ 2 ] = { head2.x, head2.y };
    uint64_t flags = 0ull;
#pragma unroll
    for( int w = 0; w < E::EW; w++ ) m[ w ] = 0ull;
    link_slots< S, 2 >( st, keys_at_cell, desc, m, flags );
    return flags & ( E::FLAG | E::MISMATCH );
}

// palette index of a colour (low 24 bits of a colour word) in the frame's lookup table (palette_kernels.cu): 1024 slots of
// index << 24 | colour, linear probing, empty = 0xFFFFFFFF.  Every colour of the frame is in the table; a colour that is not
// (row padding read through checkTJunction's flat offsets) gets 255 — it can only ever be compared, never written.
__device__ __forceinline__ uint32_t palette_index( const uint32_t* __restrict__ lut, uint32_t c24 )
{
    uint32_t h = ( c24 * 0x9E3779B1u ) >> 22;
#pragma unroll 1
    for( int n = 0; n < 1024; n++ )
    {
        const uint32_t w = __ldg( lut + h );
        if( ( w & 0x00FFFFFFu ) == c24 ) return w >> 24;
        if( w == 0xFFFFFFFFu ) break;
        h = ( h + 1u ) & 1023u;
    }
    return 255u;
}

template< int S, int A, bool kUseTma, int FMT >
__global__ void __launch_bounds__( kThreads, S <= 4 ? 5 : 4 ) raster_kernel( const __grid_constant__ CUtensorMap graph_map, const __grid_constant__ CUtensorMap img_map,
                                                           RasterArgs a )
{
    typedef Cfg< S > C;
    extern __shared__ __align__( 128 ) uint8_t smem[];
    uint8_t* s_graph = smem + C::off_graph;                                   // (staged rows live in the mask array)
    uint32_t* s_keys = reinterpret_cast< uint32_t* >( smem + C::off_keys );   // cell words, KP per row: x = key | IDs << 12, then y = IDs
    uint32_t* s_col = reinterpret_cast< uint32_t* >( smem + C::off_col );
    uint32_t* s_mask = reinterpret_cast< uint32_t* >( smem + C::off_mask );   // PACK: [2][NC]; rows: [R][NC]
    int* s_nwork = reinterpret_cast< int* >( smem + C::off_nwork );
    uint64_t* s_bar = reinterpret_cast< uint64_t* >( smem + C::off_bar );

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * C::TW, y0 = blockIdx.y * C::TH, f = blockIdx.z;
    const size_t frame_px = ( size_t )a.width * a.height;
    const uint8_t* frame = a.bgr + ( size_t )f * a.frame_stride;
    const uint8_t* graph = a.graph + ( size_t )f * frame_px;

    // (1) stage graph bytes: rows y0-2 .. y0+TH+1, columns x0-16 .. x0-16+GP-1; zero outside the image
    constexpr int kRounds = ( C::NC + kThreads - 1 ) / kThreads;     // rounds of the mask pass
    uint32_t* s_more = reinterpret_cast< uint32_t* >( s_nwork + 4 ); // [kRounds][8 warps]: the cells of a round with a third link, as ballots
    uint32_t* s_geo = s_more + kRounds * ( kThreads / 32 );          // a bit per cell (idx): queued for the geometric path
    if( tid < 3 ) s_nwork[ tid ] = 0; // [0] geometric work items, [1] smoothed cells (statistics), [2] some cell of the tile is wide
    if( tid < kRounds * ( kThreads / 32 ) ) s_geo[ tid ] = 0u;
    if( kUseTma )
    {
        if( tid == 0 )
        {
            mbar_init( s_bar, 1 );
            fence_barrier_init();
        }
        __syncthreads();
        if( tid == 0 )
        {
            mbar_expect_tx( s_bar, C::KH * C::GP + C::KH * C::RAWP );
            tma_load_3d( s_graph, &graph_map, s_bar, x0 - C::GOFF, y0 - 2, f );
            tma_load_3d( smem + C::off_raw, &img_map, s_bar, 3 * x0 - 16, y0 - 2, f ); // BGR bytes of tile + halo 2
        }
    }
    else
    {
        for( int idx = tid; idx < C::KH * C::GP; idx += kThreads )
        {
            int r = idx / C::GP, c = idx - r * C::GP;
            int gx = x0 - C::GOFF + c, gy = y0 - 2 + r;
            uint8_t v = 0;
            if( gx >= 0 && gy >= 0 && gx < a.width && gy < a.height ) v = graph[ ( size_t )gy * a.width + gx ];
            s_graph[ idx ] = v;
        }
    }
    // colours of the tile + halo 2 as RGBA words (kernel.cu:98-101: R = byte 2, G = byte 1, B = byte 0);
    // pixels outside the image hold colour 0 (the reference's reads beyond the last row see zeros), except the two
    // virtual columns x = -1 and x = width: checkTJunction addresses the pixels around a corner as FLAT byte
    // offsets idx +- widthstep +- 3 (subdivision_functions.cu:195-202), so for the first / last pixel of a row "the
    // pixel to the left / right" is the three bytes just before / after the row (the end of the previous row, the
    // start of the next one, or row padding; zero beyond the end of the image, SURVEY App. B-3 contract).
    auto virtual_colour = [ & ]( int gx, int gy ) -> uint32_t {
        const long end = ( long )a.height * a.widthstep;
        const long at = ( long )gy * a.widthstep + 3L * gx;
        uint32_t b[ 3 ];
#pragma unroll
        for( int k = 0; k < 3; k++ ) b[ k ] = ( at + k >= 0 && at + k < end ) ? ( uint32_t )__ldg( frame + at + k ) : 0u;
        return b[ 2 ] | b[ 1 ] << 8 | b[ 0 ] << 16 | 0xFF000000u;
    };
    // INDEX8: the top byte of a colour word is the colour's palette index instead of the constant alpha (equal colours
    // still compare equal, and the resolve step moves whole words, so only the stores differ)
    const uint32_t* pal_lut = nullptr;
    if constexpr( FMT == kFmtIndex8 )
        if( a.pal_count[ f ] <= 256 ) pal_lut = a.pal_lut + ( size_t )f * 1024;
    auto indexed = [ & ]( uint32_t w ) -> uint32_t {
        if constexpr( FMT != kFmtIndex8 ) return w;
        w &= 0x00FFFFFFu;
        return pal_lut ? ( w | palette_index( pal_lut, w ) << 24 ) : w;
    };
    const uint2* pack_tab = ( a.smooth.cut != nullptr && !a.debug_force_wide ) ? a.smooth.pack : nullptr;
    if( kUseTma )
    {
        // Four pixels per thread from aligned 32-bit words of the staged rows: byte permutes build the RGBA words
        // and the cell keys (left/right neighbour bits, kernel.cu:204-207), 128- / 64-bit stores.  TMA zero-filled
        // everything outside the image (row padding included), which is exactly "colour 0" / "no links".
        mbar_wait( s_bar, 0 );
        static_assert( C::KW % 4 == 0 && C::RAWOFF == 10 && C::GOFF == 16, "group layout of the vectorised staging pass" );
        constexpr int QW = C::KW / 4;
        const int vl = x0 == 0 ? 1 : -1, vr = a.width - x0 + 2; // tile columns of the virtual colour columns x = -1, x = width
        for( int idx = tid; idx < QW * C::KH; idx += kThreads )
        {
            const int cy = idx / QW, q = idx - cy * QW;
            const uint32_t* rw = reinterpret_cast< const uint32_t* >( smem + C::off_raw + cy * C::RAWP + 8 + 12 * q ); // pixel 4q starts at byte 10 + 12 q
            const uint32_t w0 = rw[ 0 ], w1 = rw[ 1 ], w2 = rw[ 2 ], w3 = rw[ 3 ];
            uint32_t c[ 4 ] = { __byte_perm( w0, w1, 0x2234 ) | 0xFF000000u, __byte_perm( w1, w1, 0x1123 ) | 0xFF000000u,
                                __byte_perm( w2, w2, 0x0012 ) | 0xFF000000u, __byte_perm( w2, w3, 0x3345 ) | 0xFF000000u };
            if( ( vl >> 2 ) == q || ( vr >> 2 ) == q )
            {
#pragma unroll
                for( int k = 0; k < 4; k++ )
                    if( 4 * q + k == vl || 4 * q + k == vr ) c[ k ] = virtual_colour( x0 - 2 + 4 * q + k, y0 - 2 + cy );
            }
            if constexpr( FMT == kFmtIndex8 )
            {
#pragma unroll
                for( int k = 0; k < 4; k++ ) c[ k ] = indexed( c[ k ] );
            }
            *reinterpret_cast< uint4* >( s_col + cy * C::KW + 4 * q ) = make_uint4( c[ 0 ], c[ 1 ], c[ 2 ], c[ 3 ] );
            const uint32_t* gw = reinterpret_cast< const uint32_t* >( s_graph + cy * C::GP + C::GOFF - 4 + 4 * q ); // cell 4q sits at staged column 14 + 4 q
            const uint32_t g0 = gw[ 0 ], g1 = gw[ 1 ];
            const uint32_t node = __byte_perm( g0, g1, 0x5432 ), left = __byte_perm( g0, g1, 0x4321 ), right = __byte_perm( g0, g1, 0x6543 );
            const uint32_t high = ( ( left >> 2 ) & 0x01010101u ) | ( ( left >> 6 ) & 0x02020202u ) | ( ( right << 2 ) & 0x04040404u ) |
                                  ( ( right >> 2 ) & 0x08080808u ); // cell_key's bits 8..11, one byte per cell
            const uint32_t k01 = __byte_perm( node, high, 0x5140 ), k23 = __byte_perm( node, high, 0x7362 );
            uint32_t kw[ 4 ] = { k01 & 0xFFFFu, k01 >> 16, k23 & 0xFFFFu, k23 >> 16 };
            uint32_t* kd = s_keys + cy * C::KP + 4 * q;
            if( pack_tab ) // the IDs of the cell's records, for the neighbours that will read these words (smooth_table.h)
            {
                uint2 cw[ 4 ];
#pragma unroll
                for( int k = 0; k < 4; k++ ) cw[ k ] = __ldg( pack_tab + kw[ k ] );
#pragma unroll
                for( int k = 0; k < 4; k++ ) kw[ k ] |= cw[ k ].x;
                *reinterpret_cast< uint4* >( kd + C::KW ) = make_uint4( cw[ 0 ].y, cw[ 1 ].y, cw[ 2 ].y, cw[ 3 ].y );
            }
            *reinterpret_cast< uint4* >( kd ) = make_uint4( kw[ 0 ], kw[ 1 ], kw[ 2 ], kw[ 3 ] );
        }
    }
    else
    {
        for( int idx = tid; idx < C::KW * C::KH; idx += kThreads )
        {
            int cy = idx / C::KW, cx = idx - cy * C::KW;
            int gx = x0 - 2 + cx, gy = y0 - 2 + cy;
            uint32_t w = 0xFF000000u;
            if( gx >= 0 && gy >= 0 && gx < a.width && gy < a.height )
            {
                const uint8_t* p = frame + ( size_t )gy * a.widthstep + 3 * gx;
                w = ( uint32_t )__ldg( p + 2 ) | ( uint32_t )__ldg( p + 1 ) << 8 | ( uint32_t )__ldg( p ) << 16 | 0xFF000000u;
            }
            else if( gx == -1 || gx == a.width )
                w = virtual_colour( gx, gy );
            s_col[ idx ] = indexed( w );
        }
        __syncthreads();
        // cell keys for tile + halo 2 (left/right neighbour bits, kernel.cu:204-207; zero outside the row)
        for( int idx = tid; idx < C::KW * C::KH; idx += kThreads )
        {
            int ky = idx / C::KW, kx = idx - ky * C::KW;
            const uint8_t* g = s_graph + ky * C::GP + kx + C::GOFF - 2; // column x0-2+kx sits at staged column kx+GOFF-2
            const uint32_t key = cell_key( g[ 0 ], g[ -1 ], g[ 1 ] );
            const uint2 ids = pack_tab ? __ldg( pack_tab + key ) : make_uint2( 0u, 0u );
            s_keys[ ky * C::KP + kx ] = ids.x | key;
            s_keys[ ky * C::KP + C::KW + kx ] = ids.y;
        }
    }
    __syncthreads();

    TileEnv< S > env;
    env.keys = s_keys;
    env.cols = s_col;
    env.x0 = x0;
    env.y0 = y0;
    env.img.frame = frame;
    env.img.width = a.width;
    env.img.height = a.height;
    env.img.widthstep = a.widthstep;
    const bool subdivide = a.subdivide != 0;
    const CellTablePtrs tab = a.tables;
    const uint32_t force_wide = a.debug_force_wide ? C::WIDE : 0u;
    const bool use_tables = a.smooth.cut != nullptr && !a.debug_force_wide;

    // (2a) Every cell of tile + halo 1 gets its mask, in tile order: a warp's cells are neighbours, so the key, colour
    // and mask accesses are conflict-free and nothing is queued.  Cells whose polygon is their plain hull copy the mask
    // from the per-scale table; smoothed cells assemble it from the smoothing tables (one CUT entry + one LINK entry per
    // shared edge with a blended end); the rare cell the tables cannot express is queued for the geometric path.
    // (Compacting the smoothed cells into a list first, so that the lookups run with full warps, was 3 % slower on the
    // busy frames of the bench — 83 % of the cells are smoothed — and 9 % faster on frames of flat 4 x 4 blocks.)
    int n_smoothed = 0;
    const int warp = tid >> 5, lane = tid & 31;
    // checkTJunction's early exit only ever holds on the image's rows 0, 1 and height - 1 (FlatImage::guard): tiles away
    // from them skip the test
    const bool tile_may_guard = y0 <= 2 || y0 + C::TH + 1 >= a.height - 1;
#pragma unroll 1
    for( int round = 0; round < kRounds; round++ ) // (whole warps: the vote at the end needs every lane)
    {
        const int idx = round * kThreads + tid;
        bool third_link = false;
        int cy = idx / C::CW, cx = idx - cy * C::CW;
        int gx = x0 - 1 + cx, gy = y0 - 1 + cy;
        const bool inside = ( unsigned )gx < ( unsigned )a.width && ( unsigned )gy < ( unsigned )a.height;
        const uint32_t* kc = s_keys + ( cy + 1 ) * C::KP + cx + 1;
        const uint32_t key = idx < C::NC ? ( *kc & 0xFFFu ) : 90u;
        const bool plain = !subdivide || ( key & 0xFFu ) == 90u;
        if( idx >= C::NC )
            ;
        else if( inside && !plain )
        {
            n_smoothed++;
            // the first two link descriptors + flags of the key: requested before the corner test below, which hides the gather
            const uint2 head = use_tables ? __ldg( a.smooth.head + key ) : make_uint2( 0u, 0u );
            // checkTJunction for the four corners of the pixel square (bit c: corner c stays), 16 = its early exit
            uint32_t cf = 16u;
#ifdef PAR_WHATIF_CF
            cf = key & 15u;
            if( false )
#else
            if( !( tile_may_guard && env.guard( gx, gy ) ) )
#endif
            {
                const uint32_t* c = s_col + ( cy + 1 ) * C::KW + ( cx + 1 );
                const uint32_t l = c[ -1 ], r = c[ 1 ], d = c[ -C::KW ], u = c[ C::KW ];
                const uint32_t dl = c[ -C::KW - 1 ], dr = c[ -C::KW + 1 ], ul = c[ C::KW - 1 ], ur = c[ C::KW + 1 ];
                // corner c stays when the three other pixels around it are not one colour: one three-input logic operation and one
                // minimum per corner (the comparisons as predicates took seven instructions more per cell)
                cf = not_one_colour( l, dl, d ) + 2u * not_one_colour( r, dr, d ) + 4u * not_one_colour( r, ur, u ) + 8u * not_one_colour( l, ul, u );
            }
            typedef Entry< S > E;
            uint64_t mw[ E::EW ];
            bool more = false;
            const uint64_t rare = use_tables ? smooth_lookup< S >( a.smooth, a.mask_lut, kc, key, head, cf, mw, more ) : E::MISMATCH;
            if( rare & E::MISMATCH ) // (rare: queued for the geometric path, a bit per cell)
            {
                atomicOr( &s_geo[ idx >> 5 ], 1u << ( idx & 31 ) );
                atomicAdd( s_nwork, 1 );
            }
            else
            {
                third_link = more;
                uint32_t wide = 0u;
                if( rare ) // (rare: some piece reaches beyond the mask)
                {
                    mw[ 0 ] &= ~( E::FLAG | E::MISMATCH );
                    wide = C::WIDE;
                    s_nwork[ 2 ] = 1;
                }
                if( C::PACK )
                {
                    s_mask[ idx ] = ( uint32_t )mw[ 0 ];
                    s_mask[ C::NC + idx ] = ( uint32_t )( mw[ 0 ] >> 32 ) | wide;
                }
                else
                {
#pragma unroll
                    for( int r = 0; r < C::R; r++ )
                    {
                        const uint32_t row = ( uint32_t )( mw[ r >> 2 ] >> ( 16 * ( r & 3 ) ) ) & 0xFFFFu;
                        s_mask[ r * C::NC + idx ] = row | ( r == 0 ? wide : 0u );
                    }
                }
            }
        }
        else if( C::PACK )
        {
            uint2 m = make_uint2( 0u, 0u );
            if( inside )
            {
                m = __ldg( reinterpret_cast< const uint2* >( a.mask_lut ) + key );
                m.y |= force_wide;
            }
            s_mask[ idx ] = m.x; // (PACK: the two halves live in separate arrays, conflict-free 32-bit accesses)
            s_mask[ C::NC + idx ] = m.y;
        }
        else
        {
            uint64_t hm[ 4 ] = { 0ull, 0ull, 0ull, 0ull };
            if( inside ) ldg_entry4( reinterpret_cast< const uint64_t* >( a.mask_lut ) + ( size_t )key * 4, hm );
#pragma unroll
            for( int r = 0; r < C::R; r++ )
                s_mask[ r * C::NC + idx ] = ( ( uint32_t )( hm[ r >> 2 ] >> ( 16 * ( r & 3 ) ) ) & 0xFFFFu ) | ( r == 0 && inside ? force_wide : 0u );
        }
        const uint32_t vote = __ballot_sync( 0xFFFFFFFFu, third_link );
        if( lane == 0 ) s_more[ round * ( kThreads / 32 ) + warp ] = vote;
    }
    // (2a') cells with three or four link descriptors (one in ten): the remaining LINK entries are XORed in.  (All four
    // slots in the first pass would cost every warp of it the instructions of the two slots that nine cells in ten do
    // not use.)  Each warp takes the cells of its own rounds, found in the ballots it left above: no list, no atomics and
    // no barrier — the dependent gathers of these few cells run under the other warps' first pass.
    __syncwarp();
    {
        uint32_t votes[ kRounds ];
        int total = 0;
#pragma unroll
        for( int r = 0; r < kRounds; r++ )
        {
            votes[ r ] = s_more[ r * ( kThreads / 32 ) + warp ];
            total += __popc( votes[ r ] );
        }
#ifdef PAR_WHATIF_NOSECOND
        total = 0;
#endif
        for( int j = lane; j < total; j += 32 ) // lane j takes the j-th marked cell
        {
            int skip = j, round = 0;
            uint32_t word = votes[ 0 ];
#pragma unroll
            for( int r = 0; r + 1 < kRounds; r++ )
            {
                const int c = __popc( votes[ r ] );
                if( round == r && skip >= c )
                {
                    skip -= c;
                    round = r + 1;
                    word = votes[ r + 1 ];
                }
            }
            // the skip-th set bit of the ballot (skip < popc), by halving
            int bitpos = 0;
#pragma unroll
            for( int h = 16; h >= 1; h >>= 1 )
            {
                const int c = __popc( ( word >> bitpos ) & ( ( 1u << h ) - 1u ) );
                if( skip >= c )
                {
                    skip -= c;
                    bitpos += h;
                }
            }
            const int idx = round * kThreads + warp * 32 + bitpos;
            const int cy = idx / C::CW, cx = idx - cy * C::CW;
            const uint32_t* kc = s_keys + ( cy + 1 ) * C::KP + cx + 1;
            typedef Entry< S > E;
            uint64_t mw[ E::EW ];
            const uint64_t rare = smooth_lookup_more< S >( a.smooth, kc, *kc & 0xFFFu, mw );
            uint32_t wide = 0u;
            if( rare )
            {
                mw[ 0 ] &= ~( E::FLAG | E::MISMATCH );
                if( rare & E::FLAG )
                {
                    wide = C::WIDE;
                    s_nwork[ 2 ] = 1;
                }
                if( rare & E::MISMATCH ) // the geometric path rebuilds the whole mask
                {
                    atomicOr( &s_geo[ idx >> 5 ], 1u << ( idx & 31 ) );
                    atomicAdd( s_nwork, 1 );
                }
            }
            if( C::PACK )
            {
                if( mw[ 0 ] | wide ) // (at small scales the third and fourth descriptors mostly sit on shared blocks: nothing to XOR)
                {
                    s_mask[ idx ] ^= ( uint32_t )mw[ 0 ];
                    const uint32_t hi = s_mask[ C::NC + idx ] ^ ( uint32_t )( mw[ 0 ] >> 32 );
                    s_mask[ C::NC + idx ] = hi | wide;
                }
            }
            else
            {
#pragma unroll
                for( int r = 0; r < C::R; r++ ) s_mask[ r * C::NC + idx ] ^= ( uint32_t )( mw[ r >> 2 ] >> ( 16 * ( r & 3 ) ) ) & 0xFFFFu;
                if( wide ) s_mask[ idx ] |= wide;
            }
        }
    }
    if( a.smooth_stats ) // (statistics for the bench line)
    {
        n_smoothed = __reduce_add_sync( 0xFFFFFFFFu, n_smoothed );
        if( lane == 0 ) atomicAdd( s_nwork + 1, n_smoothed );
    }
    __syncthreads();

    // (2b) general path, out of line (rarely runs: it costs the common path neither registers nor code)
    if( *s_nwork != 0 )
    {
        geometric_cells< S >( s_keys, s_col, s_mask, s_geo, s_nwork, x0, y0, frame, a.width, a.height, a.widthstep, tab.rec, force_wide );
        __syncthreads(); // (uniform: the counter is final since the barrier before this pass)
    }
    if( a.smooth_stats && tid == 0 )
    {
        atomicAdd( a.smooth_stats, ( unsigned long long )s_nwork[ 1 ] );     // smoothed cells
        atomicAdd( a.smooth_stats + 1, ( unsigned long long )s_nwork[ 0 ] ); // ... of which took the geometric path
    }

    // (3) resolve and write: one thread per source pixel, S output rows of S pixels each
    constexpr int O = S / A; // output pixels per source pixel and axis (A > 1: A x A samples are averaged per output pixel)
    const size_t out_w = ( size_t )a.width * O, out_h = ( size_t )a.height * O;
    constexpr int BPP = Fmt< FMT >::BPP;
    uint8_t* out = a.rgba + ( size_t )f * out_w * out_h * BPP;
    const ptrdiff_t row_step = a.flip_output ? -( ptrdiff_t )( out_w * BPP ) : ( ptrdiff_t )( out_w * BPP ); // bytes from one output row to the next
    if( s_nwork[ 2 ] != 0 || a.debug_force_wide )
    {
        // some cell of this tile reaches beyond its mask (never seen on real frames): the whole tile is resolved by
        // the exact path, kept out of line so that it costs the common path neither registers nor code
        resolve_tile_exact< S, A, FMT >( s_keys, s_col, x0, y0, frame, a.width, a.height, a.widthstep, tab.rec, subdivide, out, a.flip_output != 0 );
        return;
    }
    if constexpr( C::PACK )
    {
        // Window form: every candidate's coverage of my S x S output pixels is one masked 16-bit field of its
        // mask, so the priority resolve runs once on whole-cell bit sets instead of once per output row.
        // (a thread keeps its column; its output pointer advances by a constant per round instead of being rebuilt)
        static_assert( C::TW == 32 && kThreads % C::TW == 0, "a warp is one row of the tile" );
        constexpr int kRowsPerRound = kThreads / C::TW;
        const int lx = tid & 31, gx = x0 + lx;
        const int gy0 = y0 + ( tid >> 5 );
        uint8_t* dst = out + ( ( size_t )( a.flip_output ? out_h - 1 - ( size_t )gy0 * O : ( size_t )gy0 * O ) * out_w + ( size_t )gx * O ) * BPP;
        const ptrdiff_t round_step = row_step * ( kRowsPerRound * O );
        for( int ly = tid >> 5; ly < C::TH; ly += kRowsPerRound, dst += round_step )
        {
            const int gy = y0 + ly;
            if( gx >= a.width || gy >= a.height ) continue;
            const uint32_t* mlo = s_mask + ( ly + 1 ) * C::CW + ( lx + 1 ); // F0 | F1 << 16
            const uint32_t* mhi = mlo + C::NC;                             // F2 | F3 << 16
            const uint32_t* col = s_col + ( ly + 2 ) * C::KW + ( lx + 2 );
            uint32_t px[ S * S ];
            const uint32_t own = col[ 0 ];
#pragma unroll
            for( int k = 0; k < S * S; k++ ) px[ k ] = own;
            // candidates in DESCENDING node index: (dj,di) = (+1,+1) (+1,0) (+1,-1) (0,+1) (0,0) (0,-1) (-1,+1) (-1,0) (-1,-1);
            // the four drawn after this cell, then the cell itself: most pixels are settled by those
            uint32_t cov[ 9 ];
            cov[ 0 ] = ( mhi[ C::CW + 1 ] >> 16 ) & ( 1u << ( S * S - 1 ) );
            cov[ 1 ] = mhi[ C::CW ] & C::M_TOPROW;
            cov[ 2 ] = ( mhi[ C::CW - 1 ] >> 16 ) & ( 1u << ( S * ( S - 1 ) ) );
            cov[ 3 ] = ( mlo[ 1 ] >> 16 ) & C::M_RIGHTCOL;
            cov[ 4 ] = mlo[ 0 ] & C::ALL;
            if( C::H == 0 ) cov[ 0 ] = cov[ 1 ] = cov[ 2 ] = cov[ 3 ] = 0u; // no halo samples at this scale
#ifdef PAR_WHATIF_FASTRES
            if( ( cov[ 4 ] & ~( cov[ 0 ] | cov[ 1 ] | cov[ 2 ] | cov[ 3 ] ) ) == 0x12345u )
#else
            if( ( cov[ 4 ] & ~( cov[ 0 ] | cov[ 1 ] | cov[ 2 ] | cov[ 3 ] ) ) != C::ALL )
#endif
            {
                cov[ 5 ] = ( mlo[ -1 ] >> 16 ) & C::M_LEFTCOL;
                cov[ 6 ] = ( mhi[ -C::CW + 1 ] >> 16 ) & ( 1u << ( S - 1 ) );
                cov[ 7 ] = mhi[ -C::CW ] & C::M_BOTROW;
                cov[ 8 ] = ( mhi[ -C::CW - 1 ] >> 16 ) & 1u;
                if( C::H == 0 ) cov[ 5 ] = cov[ 6 ] = cov[ 7 ] = cov[ 8 ] = 0u;
                uint32_t rem = C::ALL;
#pragma unroll
                for( int k = 0; k < 9; k++ )
                {
                    const uint32_t take = cov[ k ] & rem;
                    rem &= ~cov[ k ];
                    if( k == 4 ) continue;
                    const int dj = 1 - k / 3, di = 1 - k % 3;
                    const uint32_t cw = take ? col[ dj * C::KW + di ] : 0u; // (predicated load: no branch per candidate)
                    // a candidate can only hold pixels of its own window
#pragma unroll
                    for( int bit = 0; bit < S * S; bit++ )
                    {
                        const int bx = bit % S, by = bit / S;
                        const bool in_window = ( di == 0 || bx == ( di > 0 ? S - 1 : 0 ) ) && ( dj == 0 || by == ( dj > 0 ? S - 1 : 0 ) );
                        if( in_window && ( ( take >> bit ) & 1u ) ) px[ bit ] = cw;
                    }
                }
                if( rem ) // nobody covers these: background (main.cpp:260)
                {
#pragma unroll
                    for( int bit = 0; bit < S * S; bit++ )
                        if( ( rem >> bit ) & 1u ) px[ bit ] = Fmt< FMT >::BACKGROUND;
                }
            }
#ifdef PAR_WHATIF_NOSTORE
            uint32_t acc = 0u;
#pragma unroll
            for( int k = 0; k < S * S; k++ ) acc += px[ k ] * ( k + 1 );
            if( acc == 0x12345u )
#endif
            store_cell< S, A, FMT >( dst, row_step, px );
        }
    }
    else
    {
        const bool out_align32 = O == 8 && ( ( reinterpret_cast< uintptr_t >( out ) | ( uintptr_t )( out_w * BPP ) ) & 31u ) == 0u; // (256-bit stores)
        for( int idx = tid; idx < C::TW * C::TH; idx += kThreads )
        {
            int ly = idx / C::TW, lx = idx - ly * C::TW;
            int gx = x0 + lx, gy = y0 + ly;
            if( gx >= a.width || gy >= a.height ) continue;
            const int cell = ( ly + 1 ) * C::CW + ( lx + 1 );
            const uint32_t* col = s_col + ( ly + 2 ) * C::KW + ( lx + 2 );
            // the 3x3 neighbourhood's masks; candidates are visited in DESCENDING node index:
            // (dj,di) = (+1,+1) (+1,0) (+1,-1) (0,+1) (0,0) (0,-1) (-1,+1) (-1,0) (-1,-1)
            uint8_t* dst = out + ( ( size_t )( a.flip_output ? out_h - 1 - ( size_t )gy * O : ( size_t )gy * O ) * out_w + ( size_t )gx * O ) * BPP;
#pragma unroll 1
            for( int ob = 0; ob < O; ob++ ) // one output row = A sample rows
            {
                ColourSum sum[ O ];
                uint32_t px[ S ];
#pragma unroll 1
                for( int r = 0; r < A; r++ )
                {
                    const int b = ob * A + r;
#pragma unroll
                    for( int k = 0; k < S; k++ ) px[ k ] = Fmt< FMT >::BACKGROUND; // background (main.cpp:260)
                    uint32_t rem = C::FULL;
#pragma unroll
                    for( int dj = 1; dj >= -1; dj-- )
                    {
                        const int ky = b - dj * S + C::H;
                        if( ky < 0 || ky >= C::R ) continue;
#pragma unroll
                        for( int di = 1; di >= -1; di-- )
                        {
                            const uint32_t m = s_mask[ ky * C::NC + cell + dj * C::CW + di ] & ~C::WIDE;
                            const uint32_t field = di == 0 ? ( m >> C::H ) : ( di > 0 ? ( m << ( S - C::H ) ) : ( m >> ( S + C::H ) ) );
                            const uint32_t take = field & rem;
                            if( take )
                            {
                                const uint32_t cw = col[ dj * C::KW + di ];
#pragma unroll
                                for( int k = 0; k < S; k++ )
                                    if( ( take >> k ) & 1u ) px[ k ] = cw;
                                rem &= ~take;
                            }
                        }
                    }
                    if( A > 1 )
                    {
#pragma unroll
                        for( int k = 0; k < S; k++ ) sum[ k / A ].add( px[ k ] );
                    }
                }
                if( A > 1 )
                {
#pragma unroll
                    for( int k = 0; k < O; k++ ) px[ k ] = sum[ k ].template mean< A * A >();
                }
                store_row< O, FMT >( dst + ( ptrdiff_t )ob * row_step, px, out_align32 );
            }
        }
    }
}

template< int S, int A, int FMT >
cudaError_t launch_raster_sa( const RasterArgs& a, const CUtensorMap* graph_map, const CUtensorMap* img_map, cudaStream_t stream )
{
    typedef Cfg< S > C;
    cudaError_t e;
    // (grid.z is limited to 65535: the callers in context.cu split larger batches)
    dim3 grid( ( a.width + C::TW - 1 ) / C::TW, ( a.height + C::TH - 1 ) / C::TH, a.n_frames );
    if( graph_map && img_map )
    {
        e = cudaFuncSetAttribute( raster_kernel< S, A, true, FMT >, cudaFuncAttributeMaxDynamicSharedMemorySize, C::smem_bytes );
        if( e != cudaSuccess ) return e;
        raster_kernel< S, A, true, FMT ><<< grid, kThreads, C::smem_bytes, stream >>>( *graph_map, *img_map, a );
    }
    else
    {
        CUtensorMap dummy;
        memset( &dummy, 0, sizeof( dummy ) );
        e = cudaFuncSetAttribute( raster_kernel< S, A, false, FMT >, cudaFuncAttributeMaxDynamicSharedMemorySize, C::smem_bytes );
        if( e != cudaSuccess ) return e;
        raster_kernel< S, A, false, FMT ><<< grid, kThreads, C::smem_bytes, stream >>>( dummy, dummy, a );
    }
    return cudaGetLastError();
}

// a.scale is the SAMPLING scale S; a.aa = A (1, 2 or 4) samples per output pixel and axis, S % A == 0
template< int S, int FMT >
cudaError_t launch_raster_s( const RasterArgs& a, const CUtensorMap* graph_map, const CUtensorMap* img_map, cudaStream_t stream )
{
    if( a.aa == 1 ) return launch_raster_sa< S, 1, FMT >( a, graph_map, img_map, stream );
    if constexpr( FMT != kFmtIndex8 ) // (averaged samples are not palette colours)
    {
        if constexpr( S % 2 == 0 )
            if( a.aa == 2 ) return launch_raster_sa< S, 2, FMT >( a, graph_map, img_map, stream );
        if constexpr( S % 4 == 0 )
            if( a.aa == 4 ) return launch_raster_sa< S, 4, FMT >( a, graph_map, img_map, stream );
    }
    return cudaErrorInvalidValue;
}

#define PAR_FOR_SCALE( scale, CALL )  \
    switch( scale )                    \
    {                                  \
        case 1: { CALL( 1 ); }         \
        case 2: { CALL( 2 ); }         \
        case 3: { CALL( 3 ); }         \
        case 4: { CALL( 4 ); }         \
        case 5: { CALL( 5 ); }         \
        case 6: { CALL( 6 ); }         \
        case 7: { CALL( 7 ); }         \
        case 8: { CALL( 8 ); }         \
        default: break;                \
    }

template< int FMT >
cudaError_t launch_raster_fmt( const RasterArgs& a, const CUtensorMap* graph_map, const CUtensorMap* img_map, cudaStream_t stream )
{
#define PAR_RASTER( S ) return launch_raster_s< S, FMT >( a, graph_map, img_map, stream )
    PAR_FOR_SCALE( a.scale, PAR_RASTER )
#undef PAR_RASTER
    return cudaErrorInvalidValue;
}

} // namespace

} // namespace par
