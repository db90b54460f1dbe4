// Stage A + B in one kernel: BGR8 frame tile -> similarity byte with trivial crossings removed
// (what the reference calls graph_aux, kernel.cu:402-415).
//
// Replaces graph_Kernel (kernel.cu:140-159; diff, graph_functions.cu:147-311) and
// trivial_cross_Kernel (kernel.cu:162-177; crossCheck_4, graph_functions.cu:1211-1274).
//
// Design (B200): one CTA per 64x32-pixel tile of one frame.  The BGR bytes of the tile plus a
// 1-pixel halo are brought into shared memory by ONE TMA bulk-tensor copy (out-of-image bytes are
// zero-filled by the TMA unit); a plain-load path fills the same layout when the frame pointer or
// strides are not 16-byte multiples.  Each pixel is converted to its packed YUV word once, and the
// words of four neighbouring pixels are stored PLANAR (one 32-bit word of four V bytes, one of U, one of
// Y), so that one VABSDIFF4 compares a field of four pixel pairs at once.  The graph is undirected, so
// instead of 8 comparisons per pixel the kernel evaluates each 2x2 block once (lower-left pixel owns it):
// the four side edges and the two diagonals — six tests of four pairs each per four blocks — and, since a
// block whose four sides are all linked loses both diagonals (crossCheck_4), stage B is folded into the
// same step.  The pixel byte is then assembled from the four blocks that touch the pixel.  Pixels outside
// the image are staged as zeros (black) and never marked: the links that point out of the image are
// cleared when the bytes are assembled, and a block across the image border can only differ from the
// reference's in its diagonals, which all point out of the image.
// Algorithmic HBM traffic: 3 B/px in + 1 B/px out.
#include "kernels.cuh"

namespace par {

namespace {

#ifndef PAR_K1_TH
#define PAR_K1_TH 32
#endif
constexpr int kTW = 64, kTH = PAR_K1_TH;     // pixels per tile
constexpr int kYW = kTW + 2, kYH = kTH + 2;  // pixels whose colour is needed (halo 1); pixel column c = x - (x0 - 1)
constexpr int kRawOff = 13;                  // TMA needs a 16-byte aligned start: rows begin at byte 3*x0 - 16, pixel c at 13 + 3c
constexpr int kRawPitch = 224;               // bytes per staged row: 13 + 3*66 = 211 rounded up to 16
// Everything after staging works on groups of FOUR pixel columns c = 4g-3 .. 4g (g = 0..17), stored at index c + 3
// so that a group is one aligned 16-byte (YUV words) / 4-byte (block bytes) unit; the group's 12 raw bytes start
// at byte 13 + 3(4g-3) = 4 + 12g, which is word aligned as well.
constexpr int kGroups = 18;
constexpr int kBH = kTH + 1;                 // 2x2 block rows per tile (block (c,r): lower-left pixel (c,r), c = 0..64)
constexpr int kBlkGroups = 17;               // block columns 4g-3 .. 4g, g = 0..16
constexpr int kBlkPitch = 4 * kGroups;       // bytes
// 128 threads per 64x32 tile, not 256: a tile's passes are separated by barriers and start behind a TMA round trip, and with
// half the threads per CTA an SM holds 11 tiles in flight instead of 8 (0.647 -> 0.609 ms per 4096 frames; 512 threads: 0.742;
// 64x16 tiles with 128 / 64 threads: 0.687 / 0.677 — the halo; profiles/r4j_*, r4k_*)
#ifndef PAR_K1_THREADS
#define PAR_K1_THREADS 128
#endif
constexpr int kThreads = PAR_K1_THREADS;

static_assert( kRawPitch >= kRawOff + 3 * kYW && kRawPitch % 16 == 0, "TMA box rows are multiples of 16 bytes" );
static_assert( 4 + 12 * kGroups <= kRawPitch, "the last group's raw bytes are inside the staged row" );

struct __align__( 128 ) GraphSmem
{
    uint8_t raw[ kYH * kRawPitch ];
    alignas( 16 ) uint4 planes[ kYH * kGroups ]; // per group of four pixels: x = their V bytes, y = U bytes, z = Y bytes (w unused)
    alignas( 16 ) uint8_t blk[ kBH * kBlkPitch ];
    uint64_t bar;
};

// Four pixel pairs at once: p and q hold the V, U, Y bytes of four pixels each (one word per field).  Returns 0x80 in
// byte k when pair k is NOT similar: a field's absolute difference (one VABSDIFF4 for the four pairs) exceeds its
// threshold (graph_functions.cu:14-19, 291-293: Y 5, U 7, V 6) iff its bit 7 is set or its low 7 bits plus
// (0x7F - threshold) carry into bit 7.
__device__ __forceinline__ uint32_t dissimilar4( const uint4& p, const uint4& q )
{
    const uint32_t dv = __vabsdiffu4( p.x, q.x ), du = __vabsdiffu4( p.y, q.y ), dy = __vabsdiffu4( p.z, q.z );
    const uint32_t tv = ( dv & 0x7F7F7F7Fu ) + 0x79797979u, tu = ( du & 0x7F7F7F7Fu ) + 0x78787878u, ty = ( dy & 0x7F7F7F7Fu ) + 0x7A7A7A7Au;
    return ( tv | dv | tu | du | ty | dy ) & 0x80808080u;
}

// the same four fields one pixel further right: bytes 1..3 of this group and byte 0 of the next
__device__ __forceinline__ uint4 shifted( const uint4& g, const uint4& next )
{
    return make_uint4( __byte_perm( g.x, next.x, 0x4321 ), __byte_perm( g.y, next.y, 0x4321 ), __byte_perm( g.z, next.z, 0x4321 ), 0u );
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

    // packed YUV word per pixel, once; four pixels (three raw words) per step.  T = 299 b0 + 587 b1 + 114 b2 of a
    // pixel is two 16 x 8-bit dot products (IDP.2A) on the raw words, wherever its three bytes sit in them.
    for( int idx = tid; idx < kYH * kGroups; idx += kThreads )
    {
        const int r = idx / kGroups, g = idx - r * kGroups;
        const uint32_t* rw = reinterpret_cast< const uint32_t* >( &s.raw[ r * kRawPitch + 4 + 12 * g ] );
        const uint32_t w0 = rw[ 0 ], w1 = rw[ 1 ], w2 = rw[ 2 ];
        constexpr uint32_t k299_587 = 299u | 587u << 16, k114_0 = 114u, k0_299 = 299u << 16, k587_114 = 587u | 114u << 16;
        const uint32_t t0 = __dp2a_hi( k114_0, w0, __dp2a_lo( k299_587, w0, 0u ) );   // bytes 0 1 2 of w0
        const uint32_t t1 = __dp2a_lo( k587_114, w1, __dp2a_hi( k0_299, w0, 0u ) );   // byte 3 of w0, bytes 0 1 of w1
        const uint32_t t2 = __dp2a_lo( k114_0, w2, __dp2a_hi( k299_587, w1, 0u ) );   // bytes 2 3 of w1, byte 0 of w2
        const uint32_t t3 = __dp2a_hi( k587_114, w2, __dp2a_lo( k0_299, w2, 0u ) );   // bytes 1 2 3 of w2
        // y = T / 1000; the colours whose rounding needs the reference's FP64 chain (non-zero multiples of 1000: rare)
        // are fixed up behind one branch for the four pixels
        int l0 = luma_div( ( int )t0 ), l1 = luma_div( ( int )t1 ), l2 = luma_div( ( int )t2 ), l3 = luma_div( ( int )t3 );
        const bool c0 = luma_needs_chain( ( int )t0 ), c1 = luma_needs_chain( ( int )t1 ), c2 = luma_needs_chain( ( int )t2 ), c3 = luma_needs_chain( ( int )t3 );
        if( c0 || c1 || c2 || c3 )
        {
            if( c0 ) l0 = luma_chain( l0, w0 & 255u, ( w0 >> 8 ) & 255u, ( w0 >> 16 ) & 255u );
            if( c1 ) l1 = luma_chain( l1, w0 >> 24, w1 & 255u, ( w1 >> 8 ) & 255u );
            if( c2 ) l2 = luma_chain( l2, ( w1 >> 16 ) & 255u, w1 >> 24, w2 & 255u );
            if( c3 ) l3 = luma_chain( l3, ( w2 >> 8 ) & 255u, ( w2 >> 16 ) & 255u, w2 >> 24 );
        }
        const uint32_t y0w = yuv_pack( l0, w0 & 255u, ( w0 >> 16 ) & 255u ), y1w = yuv_pack( l1, w0 >> 24, ( w1 >> 8 ) & 255u );
        const uint32_t y2w = yuv_pack( l2, ( w1 >> 16 ) & 255u, w2 & 255u ), y3w = yuv_pack( l3, ( w2 >> 8 ) & 255u, w2 >> 24 );
        // 4 x 3 byte transpose: word k = (V, U, Y, -) of pixel k  ->  (V0 V1 V2 V3), (U0 ..), (Y0 ..)
        const uint32_t vu01 = __byte_perm( y0w, y1w, 0x5140 ), vu23 = __byte_perm( y2w, y3w, 0x5140 ); // V0 V1 U0 U1
        const uint32_t yy01 = __byte_perm( y0w, y1w, 0x7362 ), yy23 = __byte_perm( y2w, y3w, 0x7362 ); // Y0 Y1 .  .
        s.planes[ idx ] = make_uint4( __byte_perm( vu01, vu23, 0x5410 ), __byte_perm( vu01, vu23, 0x7632 ), __byte_perm( yy01, yy23, 0x5410 ), 0u );
    }
    __syncthreads();

    // four 2x2 blocks per step (block columns 4g-3 .. 4g, lower-left pixels = group g of row r): bit0 = bottom side,
    // bit1 = left side, bit2 = "/" diagonal, bit3 = "\" diagonal, diagonals already cleared when all four sides are
    // linked (stage B)
    for( int idx = tid; idx < kBH * kBlkGroups; idx += kThreads )
    {
        const int r = idx / kBlkGroups, g = idx - r * kBlkGroups;
        const uint4* lo = &s.planes[ r * kGroups + g ];
        const uint4 p = lo[ 0 ], q = lo[ kGroups ];                                      // pixels (c, r), (c, r+1)
        const uint4 ps = shifted( p, lo[ 1 ] ), qs = shifted( q, lo[ kGroups + 1 ] );    // pixels (c+1, r), (c+1, r+1)
        const uint32_t hb = dissimilar4( p, ps ), ht = dissimilar4( q, qs ), vl = dissimilar4( p, q ), vr = dissimilar4( ps, qs );
        const uint32_t d1 = dissimilar4( p, qs ), d2 = dissimilar4( ps, q );
        const uint32_t keep = hb | ht | vl | vr; // 0x80: some side is missing, the diagonals stay
        const uint32_t word = ( ( hb ^ 0x80808080u ) >> 7 ) | ( ( vl ^ 0x80808080u ) >> 6 ) | ( ( ~d1 & keep ) >> 5 ) | ( ( ~d2 & keep ) >> 4 );
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
        uint32_t bytes = ( ( ul >> 3 ) & 0x01010101u )    // bit 0: "\" of the up-left block
                         | ( ur & 0x02020202u )           // bit 1: up (left side of the up-right block)
                         | ( ur & 0x04040404u )           // bit 2: "/" of the up-right block
                         | ( ( ul << 3 ) & 0x08080808u )  // bit 3: left (bottom side of the up-left block)
                         | ( ( ur << 4 ) & 0x10101010u )  // bit 4: right (bottom side of the up-right block)
                         | ( ( dl << 3 ) & 0x20202020u )  // bit 5: "/" of the down-left block
                         | ( ( dr << 5 ) & 0x40404040u )  // bit 6: down (left side of the down-right block)
                         | ( ( dr << 4 ) & 0x80808080u ); // bit 7: "\" of the down-right block
        // no links out of the image (graph_functions.cu:162-171 test the neighbour's coordinates): bits 5 6 7 point
        // down, 0 1 2 up, 0 3 5 left, 2 4 7 right
        if( gy == 0 ) bytes &= 0x1F1F1F1Fu;
        if( gy == a.height - 1 ) bytes &= 0xF8F8F8F8u;
        if( gx == 0 ) bytes &= 0xFFFFFFD6u;
        const int last = a.width - 1 - gx; // byte lane of the image's last column
        if( last < 4 ) bytes &= ~( 0x94u << ( 8 * last ) );
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
