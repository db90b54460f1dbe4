// Stage A + B in one kernel: BGR8 frame tile -> similarity byte with trivial crossings removed
// (what the reference calls graph_aux, kernel.cu:402-415).
//
// Replaces graph_Kernel (kernel.cu:140-159; diff, graph_functions.cu:147-311) and
// trivial_cross_Kernel (kernel.cu:162-177; crossCheck_4, graph_functions.cu:1211-1274).
//
// Design (B200): one CTA per 64x32-pixel tile of one frame.  The BGR bytes of the tile plus a
// 1-pixel halo are brought into shared memory by ONE TMA bulk-tensor copy (out-of-image bytes are
// zero-filled by the TMA unit); a plain-load path fills the same layout when the frame pointer or
// strides are not 16-byte multiples.  Each pixel is converted to its packed YUV word once.  The
// graph is undirected, so instead of 8 comparisons per pixel the kernel evaluates each 2x2 block
// once (lower-left pixel owns it): the four side edges and the two diagonals, and — since a block
// whose four sides are all linked loses both diagonals (crossCheck_4) — stage B is folded into the
// same step.  The pixel byte is then assembled from the four blocks that touch the pixel.
// Algorithmic HBM traffic: 3 B/px in + 1 B/px out.
#include "kernels.cuh"

namespace par {

namespace {

constexpr int kTW = 64, kTH = 32;            // pixels per tile
constexpr int kYW = kTW + 2, kYH = kTH + 2;  // pixels whose colour is needed (halo 1); pixel column c = x - (x0 - 1)
constexpr int kRawOff = 13;                  // TMA needs a 16-byte aligned start: rows begin at byte 3*x0 - 16, pixel c at 13 + 3c
constexpr int kRawPitch = 224;               // bytes per staged row: 13 + 3*66 = 211 rounded up to 16
// Everything after staging works on groups of FOUR pixel columns c = 4g-3 .. 4g (g = 0..17), stored at index c + 3
// so that a group is one aligned 16-byte (YUV words) / 4-byte (block bytes) unit; the group's 12 raw bytes start
// at byte 13 + 3(4g-3) = 4 + 12g, which is word aligned as well.
constexpr int kGroups = 18;
constexpr int kYuvPitch = 4 * kGroups;       // words
constexpr int kBH = kTH + 1;                 // 2x2 block rows per tile (block (c,r): lower-left pixel (c,r), c = 0..64)
constexpr int kBlkGroups = 17;               // block columns 4g-3 .. 4g, g = 0..16
constexpr int kBlkPitch = 4 * kGroups;       // bytes
constexpr int kThreads = 256;
constexpr uint32_t kInvalid = 0x80000000u;   // pixel outside the image

static_assert( kRawPitch >= kRawOff + 3 * kYW && kRawPitch % 16 == 0, "TMA box rows are multiples of 16 bytes" );
static_assert( 4 + 12 * kGroups <= kRawPitch, "the last group's raw bytes are inside the staged row" );

struct __align__( 128 ) GraphSmem
{
    uint8_t raw[ kYH * kRawPitch ];
    alignas( 16 ) uint32_t yuv[ kYH * kYuvPitch ];
    alignas( 16 ) uint8_t blk[ kBH * kBlkPitch ];
    uint64_t bar;
};

// Similarity of two staged words.  A staged word is the packed YUV word's low three bytes (V, U, Y — the
// fields graph_functions.cu:291-293 masks out and compares) with a top byte of 0x00 for a pixel inside the
// image and 0x80 for one outside.  One VABSDIFF4 gives the four per-byte absolute differences; a byte
// exceeds its threshold (top 0, Y 5, U 7, V 6) iff its bit 7 is set or its low 7 bits plus (0x7F - threshold)
// carry into bit 7.  In-image vs out-of-image differs by 0x80 in the top byte, so it is never similar.
__device__ __forceinline__ uint32_t sim( uint32_t p, uint32_t q )
{
    const uint32_t d = __vabsdiffu4( p, q );
    const uint32_t s = ( d & 0x7F7F7F7Fu ) + 0x7F7A7879u;
    return ( ( ( s | d ) & 0x80808080u ) == 0u ) ? 1u : 0u;
}

template< bool kUseTma >
__global__ void __launch_bounds__( kThreads ) similarity_graph_kernel( const __grid_constant__ CUtensorMap img_map, GraphArgs a )
{
    __shared__ GraphSmem s;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH, f = blockIdx.z;
    const uint8_t* frame = a.bgr + ( size_t )f * a.frame_stride;

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
            mbar_expect_tx( &s.bar, kYH * kRawPitch );
            tma_load_3d( s.raw, &img_map, &s.bar, 3 * x0 - 16, y0 - 1, f ); // innermost coordinate must be a multiple of 16 bytes
        }
        mbar_wait( &s.bar, 0 );
    }
    else
    {
        // same layout, plain loads; bytes outside the frame's rows/row-bytes are zero
        const int row_bytes = 3 * a.width;
        for( int idx = tid; idx < kYH * kRawPitch; idx += kThreads )
        {
            int r = idx / kRawPitch, c = idx - r * kRawPitch;
            int gy = y0 - 1 + r, gb = 3 * x0 - 16 + c;
            uint8_t v = 0;
            if( gy >= 0 && gy < a.height && gb >= 0 && gb < row_bytes ) v = frame[ ( size_t )gy * a.widthstep + gb ];
            s.raw[ idx ] = v;
        }
        __syncthreads();
    }

    // packed YUV word per pixel, once; four pixels (three raw words) per step, one 128-bit store
    for( int idx = tid; idx < kYH * kGroups; idx += kThreads )
    {
        const int r = idx / kGroups, g = idx - r * kGroups;
        const uint32_t* rw = reinterpret_cast< const uint32_t* >( &s.raw[ r * kRawPitch + 4 + 12 * g ] );
        const uint32_t w0 = rw[ 0 ], w1 = rw[ 1 ], w2 = rw[ 2 ];
        uint32_t y[ 4 ];
        y[ 0 ] = yuv_word( w0 & 255u, ( w0 >> 8 ) & 255u, ( w0 >> 16 ) & 255u ) & 0x00FFFFFFu;
        y[ 1 ] = yuv_word( w0 >> 24, w1 & 255u, ( w1 >> 8 ) & 255u ) & 0x00FFFFFFu;
        y[ 2 ] = yuv_word( ( w1 >> 16 ) & 255u, w1 >> 24, w2 & 255u ) & 0x00FFFFFFu;
        y[ 3 ] = yuv_word( ( w2 >> 8 ) & 255u, ( w2 >> 16 ) & 255u, w2 >> 24 ) & 0x00FFFFFFu;
        const int gx = x0 - 4 + 4 * g, gy = y0 - 1 + r; // image column of the group's first pixel (c = 4g - 3)
        if( gy < 0 || gy >= a.height )
            y[ 0 ] = y[ 1 ] = y[ 2 ] = y[ 3 ] = kInvalid;
        else if( gx < 0 || gx + 3 >= a.width )
        {
#pragma unroll
            for( int k = 0; k < 4; k++ )
                if( gx + k < 0 || gx + k >= a.width ) y[ k ] = kInvalid;
        }
        *reinterpret_cast< uint4* >( &s.yuv[ r * kYuvPitch + 4 * g ] ) = make_uint4( y[ 0 ], y[ 1 ], y[ 2 ], y[ 3 ] );
    }
    __syncthreads();

    // four 2x2 blocks per step (block columns 4g-3 .. 4g): bit0 = bottom side, bit1 = left side, bit2 = "/" diagonal,
    // bit3 = "\" diagonal, diagonals already cleared when all four sides are linked (stage B)
    for( int idx = tid; idx < kBH * kBlkGroups; idx += kThreads )
    {
        const int r = idx / kBlkGroups, g = idx - r * kBlkGroups;
        const uint32_t* lo = &s.yuv[ r * kYuvPitch + 4 * g ];
        const uint32_t* hi = lo + kYuvPitch;
        const uint4 l4 = *reinterpret_cast< const uint4* >( lo ), h4 = *reinterpret_cast< const uint4* >( hi );
        const uint32_t p[ 5 ] = { l4.x, l4.y, l4.z, l4.w, lo[ 4 ] }, q[ 5 ] = { h4.x, h4.y, h4.z, h4.w, hi[ 4 ] };
        uint32_t v[ 5 ];
#pragma unroll
        for( int k = 0; k < 5; k++ ) v[ k ] = sim( p[ k ], q[ k ] ); // vertical sides
        uint32_t word = 0u;
#pragma unroll
        for( int k = 0; k < 4; k++ )
        {
            const uint32_t hb = sim( p[ k ], p[ k + 1 ] ), ht = sim( q[ k ], q[ k + 1 ] );
            const uint32_t d1 = sim( p[ k ], q[ k + 1 ] ), d2 = sim( p[ k + 1 ], q[ k ] );
            const uint32_t keep = ( hb & ht & v[ k ] & v[ k + 1 ] ) ^ 1u;
            word |= ( hb | ( v[ k ] << 1 ) | ( ( d1 & keep ) << 2 ) | ( ( d2 & keep ) << 3 ) ) << ( 8 * k );
        }
        *reinterpret_cast< uint32_t* >( &s.blk[ r * kBlkPitch + 4 * g ] ) = word;
    }
    __syncthreads();

    // assemble 4 horizontally adjacent pixel bytes per thread, all four at once on byte lanes.  Pixel (lx+k, ly) is
    // column c = lx+k+1 of staged row ly+1; around it: UL = block (c-1, ly+1), UR = (c, ly+1), DL = (c-1, ly), DR = (c, ly).
    uint8_t* out = a.graph_aux + ( size_t )f * a.width * a.height;
    const bool word_ok = ( a.width & 3 ) == 0;
    for( int idx = tid; idx < ( kTW / 4 ) * kTH; idx += kThreads )
    {
        int ly = idx / ( kTW / 4 ), lx = ( idx - ly * ( kTW / 4 ) ) * 4;
        int gx = x0 + lx, gy = y0 + ly;
        if( gx >= a.width || gy >= a.height ) continue;
        const uint32_t* up = reinterpret_cast< const uint32_t* >( &s.blk[ ( ly + 1 ) * kBlkPitch + lx ] ); // block column lx-3 .. : bytes lx ..
        const uint32_t* dn = reinterpret_cast< const uint32_t* >( &s.blk[ ly * kBlkPitch + lx ] );
        const uint32_t ur = up[ 1 ], dr = dn[ 1 ];                                                    // block columns lx+1 .. lx+4
        const uint32_t ul = __byte_perm( up[ 0 ], ur, 0x6543 ), dl = __byte_perm( dn[ 0 ], dr, 0x6543 ); // block columns lx .. lx+3
        const uint32_t bytes = ( ( ul >> 3 ) & 0x01010101u )    // bit 0: "\" of the up-left block
                               | ( ur & 0x02020202u )           // bit 1: up (left side of the up-right block)
                               | ( ur & 0x04040404u )           // bit 2: "/" of the up-right block
                               | ( ( ul << 3 ) & 0x08080808u )  // bit 3: left (bottom side of the up-left block)
                               | ( ( ur << 4 ) & 0x10101010u )  // bit 4: right (bottom side of the up-right block)
                               | ( ( dl << 3 ) & 0x20202020u )  // bit 5: "/" of the down-left block
                               | ( ( dr << 5 ) & 0x40404040u )  // bit 6: down (left side of the down-right block)
                               | ( ( dr << 4 ) & 0x80808080u ); // bit 7: "\" of the down-right block
        size_t o = ( size_t )gy * a.width + gx;
        if( word_ok && gx + 3 < a.width )
            *reinterpret_cast< uint32_t* >( out + o ) = bytes;
        else
            for( int k = 0; k < 4 && gx + k < a.width; k++ ) out[ o + k ] = ( uint8_t )( bytes >> ( 8 * k ) );
    }
}

} // namespace

dim3 similarity_graph_grid( int width, int height, int n_frames )
{
    return dim3( ( width + kTW - 1 ) / kTW, ( height + kTH - 1 ) / kTH, n_frames );
}

void similarity_graph_tma_box( uint32_t box[ 3 ] )
{
    box[ 0 ] = kRawPitch;
    box[ 1 ] = kYH;
    box[ 2 ] = 1;
}

cudaError_t launch_similarity_graph( const GraphArgs& a, const CUtensorMap* img_map, cudaStream_t stream )
{
    dim3 grid = similarity_graph_grid( a.width, a.height, a.n_frames );
    if( img_map )
        similarity_graph_kernel< true ><<< grid, kThreads, 0, stream >>>( *img_map, a );
    else
    {
        CUtensorMap dummy;
        memset( &dummy, 0, sizeof( dummy ) );
        similarity_graph_kernel< false ><<< grid, kThreads, 0, stream >>>( dummy, a );
    }
    return cudaGetLastError();
}

} // namespace par
