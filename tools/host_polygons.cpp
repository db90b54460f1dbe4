// Host build of the product's polygon logic (polygon.cuh is __host__ __device__): lets the CPU test suite
// check stage D+E of the product against the oracle without a GPU.  Not part of the shipped library.
#include <cstdint>
#include <cstring>
#define __restrict__
#include "../pixel_art_remaster_gpu_b200/csrc/polygon.cuh"
using namespace par;
namespace {
struct HostEnv
{
    const uint8_t* graph;
    int width, height;
    FlatImage img;
    uint32_t node( int i, int j ) const { return ( i >= 0 && j >= 0 && i < width && j < height ) ? graph[ ( size_t )j * width + i ] : 0u; }
    uint32_t key( int i, int j ) const { return cell_key( node( i, j ), node( i - 1, j ), node( i + 1, j ) ); }
    bool guard( int i, int j ) const { return img.guard( i, j ); }
    bool keep_corner( int i, int j, Q2 p ) const { return img.keep_corner( i, j, p ); }
};
struct HostSlots
{
    int x[ 16 ], y[ 16 ];
    void put( int slot, int x64, int y64 ) { x[ slot ] = x64; y[ slot ] = y64; }
};
} // namespace
extern "C" void host_polygons( const uint8_t* img_zero_tailed, const uint8_t* graph, int W, int H, int ws, int subdivide, float* poly /* N*45*2 */, int* count )
{
    static CellTables T;
    static bool built = false;
    if( !built ) { build_cell_tables( &T ); built = true; }
    HostEnv env;
    env.graph = graph; env.width = W; env.height = H;
    env.img.frame = img_zero_tailed; env.img.width = W; env.img.height = H; env.img.widthstep = ws;
    CellTablePtrs tab{ T.rec };
    memset( poly, 0, sizeof( float ) * 90 * ( size_t )W * H );
    for( int j = 0; j < H; j++ )
        for( int i = 0; i < W; i++ )
        {
            HostSlots s;
            CellPoly p = build_cell_polygon( env, tab, i, j, env.key( i, j ), subdivide != 0, s );
            float* o = poly + ( size_t )( j * W + i ) * 90;
            int m = 0;
            for( int t = 0; t < p.n; t++ )
                for( int e = 0; e <= ( int )( ( p.two >> t ) & 1u ); e++, m++ ) { o[ 2 * m ] = s.x[ 2 * t + e ] / 64.0f; o[ 2 * m + 1 ] = s.y[ 2 * t + e ] / 64.0f; }
            count[ j * W + i ] = m;
        }
}
