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
constexpr int kYW = kTW + 2, kYH = kTH + 2;  // pixels whose colour is needed (halo 1)
constexpr int kRawOff = 13;                  // TMA needs a 16-byte aligned start: rows begin at byte 3*x0 - 16
constexpr int kRawPitch = 224;               // bytes per staged row: 13 + 3*66 = 211 rounded up to 16
constexpr int kYuvPitch = kYW + 1;           // words
constexpr int kBW = kTW + 1, kBH = kTH + 1;  // 2x2 blocks per tile (one extra column/row at left/bottom)
constexpr int kBlkPitch = 68;                // bytes
constexpr int kThreads = 256;
constexpr uint32_t kInvalid = 0x80000000u;   // pixel outside the image

static_assert( kRawPitch >= kRawOff + 3 * kYW && kRawPitch % 16 == 0, "TMA box rows are multiples of 16 bytes" );

struct __align__( 128 ) GraphSmem
{
    uint8_t raw[ kYH * kRawPitch ];
    uint32_t yuv[ kYH * kYuvPitch ];
    uint8_t blk[ kBH * kBlkPitch ];
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

    // packed YUV word per pixel, once
    for( int idx = tid, r = tid / kYW, c = tid % kYW; idx < kYH * kYW; idx += kThreads )
    {
        int gx = x0 - 1 + c, gy = y0 - 1 + r;
        const uint8_t* p = &s.raw[ r * kRawPitch + kRawOff + 3 * c ];
        uint32_t w = yuv_word( p[ 0 ], p[ 1 ], p[ 2 ] ) & 0x00FFFFFFu;
        if( gx < 0 || gy < 0 || gx >= a.width || gy >= a.height ) w = kInvalid;
        s.yuv[ r * kYuvPitch + c ] = w;
        c += kThreads % kYW; // next element of this thread, without a division
        r += kThreads / kYW;
        if( c >= kYW )
        {
            c -= kYW;
            r++;
        }
    }
    __syncthreads();

    // one 2x2 block per step: bit0 = bottom side, bit1 = left side, bit2 = "/" diagonal, bit3 = "\" diagonal,
    // diagonals already cleared when all four sides are linked (stage B)
    for( int idx = tid, r = tid / kBW, c = tid % kBW; idx < kBH * kBW; idx += kThreads )
    {
        uint32_t p00 = s.yuv[ r * kYuvPitch + c ], p10 = s.yuv[ r * kYuvPitch + c + 1 ];
        uint32_t p01 = s.yuv[ ( r + 1 ) * kYuvPitch + c ], p11 = s.yuv[ ( r + 1 ) * kYuvPitch + c + 1 ];
        uint32_t hb = sim( p00, p10 ), ht = sim( p01, p11 ), vl = sim( p00, p01 ), vr = sim( p10, p11 );
        uint32_t d1 = sim( p00, p11 ), d2 = sim( p10, p01 );
        uint32_t keep = ( hb & ht & vl & vr ) ^ 1u;
        s.blk[ r * kBlkPitch + c ] = ( uint8_t )( hb | ( vl << 1 ) | ( ( d1 & keep ) << 2 ) | ( ( d2 & keep ) << 3 ) );
        c += kThreads % kBW;
        r += kThreads / kBW;
        if( c >= kBW )
        {
            c -= kBW;
            r++;
        }
    }
    __syncthreads();

    // assemble 4 horizontally adjacent pixel bytes per thread
    uint8_t* out = a.graph_aux + ( size_t )f * a.width * a.height;
    const bool word_ok = ( a.width & 3 ) == 0;
    for( int idx = tid; idx < ( kTW / 4 ) * kTH; idx += kThreads )
    {
        int ly = idx / ( kTW / 4 ), lx = ( idx - ly * ( kTW / 4 ) ) * 4;
        int gx = x0 + lx, gy = y0 + ly;
        if( gx >= a.width || gy >= a.height ) continue;
        const uint8_t* up = &s.blk[ ( ly + 1 ) * kBlkPitch + lx ]; // blocks whose lower-left pixel is (lx-1.., ly)
        const uint8_t* dn = &s.blk[ ly * kBlkPitch + lx ];         // ... and (lx-1.., ly-1)
        uint32_t bytes = 0;
#pragma unroll
        for( int k = 0; k < 4; k++ )
        {
            uint32_t e10 = up[ k ], e00 = up[ k + 1 ], e11 = dn[ k ], e01 = dn[ k + 1 ];
            uint32_t b = ( ( e10 >> 3 ) & 1u )          // bit 0: "\" of the up-left block
                         | ( ( ( e00 >> 1 ) & 1u ) << 1 ) // bit 1: up
                         | ( ( ( e00 >> 2 ) & 1u ) << 2 ) // bit 2: "/" of the up-right block
                         | ( ( e10 & 1u ) << 3 )          // bit 3: left
                         | ( ( e00 & 1u ) << 4 )          // bit 4: right
                         | ( ( ( e11 >> 2 ) & 1u ) << 5 ) // bit 5: "/" of the down-left block
                         | ( ( ( e01 >> 1 ) & 1u ) << 6 ) // bit 6: down
                         | ( ( ( e01 >> 3 ) & 1u ) << 7 ); // bit 7: "\" of the down-right block
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
