// Per-frame palette for PAR_OUT_INDEX8: the distinct colours of a frame (+ black, the background, main.cpp:260) in ascending
// order, and a lookup table colour -> index for the raster kernel's staging pass.
//
// New subsystem (no reference counterpart): the reference hands 4 bytes per vertex to OpenGL (color_kernel, kernel.cu:67-104);
// here the point-sampled image can leave the device as one byte per pixel, because each of its pixels is the colour of a
// source pixel (kernel.cu:98-101) or the background.
//
// Design: (1) one CTA per 64 K-pixel chunk of a frame collects the chunk's distinct colours in a shared-memory hash set (a
// pixel whose colour is already there costs one shared load) and merges them into the frame's table in global memory, counting
// new entries; (2) one CTA per frame ranks the table's entries (<= 256, else the frame is not representable), writes the
// palette and turns the table into index << 24 | colour.  Colours are the low 24 bits of the raster kernel's colour words
// (R | G << 8 | B << 16).
// Algorithmic HBM traffic: 3 B/px in (the frame is read once more by the raster kernel), 5 KB per frame out.
#include "kernels.cuh"

namespace par {

namespace {

constexpr int kThreads = 1024;
constexpr int kChunkPx = 65536;
constexpr uint32_t kEmpty = 0xFFFFFFFFu;
constexpr int kMaxColours = 512; // beyond this the frame is hopeless anyway: stop inserting (the tables never fill up)

__device__ __forceinline__ uint32_t slot_of( uint32_t c ) { return ( c * 0x9E3779B1u ) >> 22; }
static_assert( kPaletteSlots == 1024, "slot_of keeps 10 bits" );

// insert into an open-addressing set; returns true when the colour was not there.  `n` counts the entries.
__device__ __forceinline__ bool set_insert( uint32_t* tab, int* n, uint32_t c )
{
    uint32_t h = slot_of( c );
    for( int probes = 0; probes < kPaletteSlots; probes++ )
    {
        const uint32_t k = tab[ h ];
        if( k == c ) return false;
        if( k == kEmpty )
        {
            if( *( volatile int* )n >= kMaxColours ) return false;
            const uint32_t old = atomicCAS( &tab[ h ], kEmpty, c );
            if( old == kEmpty )
            {
                atomicAdd( n, 1 );
                return true;
            }
            if( old == c ) return false;
        }
        h = ( h + 1u ) & ( kPaletteSlots - 1u );
    }
    return false;
}

__global__ void __launch_bounds__( kThreads ) palette_collect_kernel( PaletteArgs a )
{
    __shared__ uint32_t s_tab[ kPaletteSlots ];
    __shared__ int s_n;
    const int f = blockIdx.y;
    const size_t frame_px = ( size_t )a.width * a.height;
    const size_t p0 = ( size_t )blockIdx.x * kChunkPx, p1 = min( p0 + ( size_t )kChunkPx, frame_px );
    const uint8_t* frame = a.bgr + ( size_t )f * a.frame_stride;
    for( int k = threadIdx.x; k < kPaletteSlots; k += kThreads ) s_tab[ k ] = kEmpty;
    if( threadIdx.x == 0 ) s_n = 0;
    __syncthreads();
    if( blockIdx.x == 0 && threadIdx.x == 0 ) set_insert( s_tab, &s_n, 0u ); // black is always entry 0
    uint32_t last = kEmpty; // (runs of one colour are the rule in pixel art: skip the probe)
    if( ( a.width & 3 ) == 0 && ( a.widthstep & 3 ) == 0 && ( reinterpret_cast< uintptr_t >( frame ) & 3u ) == 0 )
    {
        // four pixels = three aligned words per step (rows are whole groups)
        const int groups_per_row = a.width >> 2;
        for( size_t g = ( p0 >> 2 ) + threadIdx.x; g < ( p1 >> 2 ); g += kThreads )
        {
            const int row = ( int )( g / groups_per_row ), q = ( int )( g - ( size_t )row * groups_per_row );
            const uint32_t* w = reinterpret_cast< const uint32_t* >( frame + ( size_t )row * a.widthstep + 12 * q );
            const uint32_t w0 = __ldg( w ), w1 = __ldg( w + 1 ), w2 = __ldg( w + 2 );
            // bytes in memory: B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3  ->  words R | G << 8 | B << 16
            const uint32_t c[ 4 ] = { __byte_perm( w0, 0u, 0x4012 ), __byte_perm( w0, w1, 0x0345 ) & 0x00FFFFFFu, __byte_perm( w1, w2, 0x0234 ) & 0x00FFFFFFu,
                                      __byte_perm( w2, 0u, 0x4123 ) };
#pragma unroll
            for( int k = 0; k < 4; k++ )
                if( c[ k ] != last )
                {
                    set_insert( s_tab, &s_n, c[ k ] );
                    last = c[ k ];
                }
        }
    }
    else
    {
        for( size_t p = p0 + threadIdx.x; p < p1; p += kThreads )
        {
            const int row = ( int )( p / a.width ), col = ( int )( p - ( size_t )row * a.width );
            const uint8_t* px = frame + ( size_t )row * a.widthstep + 3 * col;
            const uint32_t c = ( uint32_t )__ldg( px + 2 ) | ( uint32_t )__ldg( px + 1 ) << 8 | ( uint32_t )__ldg( px ) << 16;
            if( c != last )
            {
                set_insert( s_tab, &s_n, c );
                last = c;
            }
        }
    }
    __syncthreads();
    // merge into the frame's table
    uint32_t* g_tab = a.lut + ( size_t )f * kPaletteSlots;
    for( int k = threadIdx.x; k < kPaletteSlots; k += kThreads )
    {
        const uint32_t c = s_tab[ k ];
        if( c != kEmpty ) set_insert( g_tab, a.count + f, c );
    }
    if( threadIdx.x == 0 && s_n >= kMaxColours ) atomicMax( a.count + f, kMaxColours ); // the chunk alone has too many
}

__global__ void __launch_bounds__( kThreads ) palette_rank_kernel( PaletteArgs a )
{
    __shared__ uint32_t s_col[ 256 ];
    __shared__ int s_n;
    const int f = blockIdx.x, t = threadIdx.x;
    uint32_t* tab = a.lut + ( size_t )f * kPaletteSlots;
    const int n = a.count[ f ];
    if( t == 0 ) s_n = 0;
    __syncthreads();
    const uint32_t c = tab[ t ]; // (kThreads == kPaletteSlots)
    if( n <= 256 && c != kEmpty ) s_col[ atomicAdd( &s_n, 1 ) ] = c;
    __syncthreads();
    if( n <= 256 && c != kEmpty )
    {
        int rank = 0;
        for( int k = 0; k < n; k++ ) rank += s_col[ k ] < c ? 1 : 0;
        tab[ t ] = ( uint32_t )rank << 24 | c;
        if( a.palette ) a.palette[ ( size_t )f * 256 + rank ] = c | 0xFF000000u;
    }
    if( a.palette && t < 256 && ( n > 256 || t >= n ) ) a.palette[ ( size_t )f * 256 + t ] = 0xFF000000u; // unused entries: black
}
static_assert( kThreads == kPaletteSlots, "one thread per slot in the rank kernel" );

} // namespace

cudaError_t launch_palette( const PaletteArgs& a, cudaStream_t stream, int* n_launches )
{
    cudaError_t e = cudaMemsetAsync( a.lut, 0xFF, ( size_t )a.n_frames * kPaletteSlots * sizeof( uint32_t ), stream );
    if( e == cudaSuccess ) e = cudaMemsetAsync( a.count, 0, ( size_t )a.n_frames * sizeof( int32_t ), stream );
    if( e != cudaSuccess ) return e;
    const size_t frame_px = ( size_t )a.width * a.height;
    const unsigned chunks = ( unsigned )( ( frame_px + kChunkPx - 1 ) / kChunkPx );
    palette_collect_kernel<<< dim3( chunks, a.n_frames ), kThreads, 0, stream >>>( a ); // (callers keep n_frames <= 65535)
    palette_rank_kernel<<< a.n_frames, kThreads, 0, stream >>>( a );
    if( n_launches ) *n_launches = 2;
    return cudaGetLastError();
}

} // namespace par
