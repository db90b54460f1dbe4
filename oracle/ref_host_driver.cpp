/* TEST INFRASTRUCTURE ONLY (oracle/).  Never linked into, imported by, or executed from the product
 * path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load the library this file builds.
 *
 * ref_host: the reference's own per-pixel routines, compiled UNMODIFIED as serial host C++.
 * The five reference source files are #include'd from where they lie under /root/reference
 * (never copied into this repo); the build recipe is oracle/Makefile and the output goes to
 * oracle/_ref/ only.  This file contributes nothing but the loop nest that kernel.cu's
 * launch_kernel performs with CUDA launches (kernel.cu:402-475), guard bands for the reference's
 * out-of-bounds reads, and a C ABI so the tests can look at every stage.
 *
 * Two builds are made from this one file:
 *   libref_host.so      -ffp-contract=off            ("plain host" arithmetic)
 *   libref_host_fma.so  -ffp-contract=fast -mfma     (lets g++ fuse graph_functions.cu:90 the way
 *                                                     nvcc does on the device)
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <ctime>
#include <time.h>
#include <vector>

#define __device__
#define __global__
#define __host__
#define __align__( n ) __attribute__( ( aligned( n ) ) )
#define CELL_SIZE 45
static inline void __syncthreads() {}
using std::sqrt;

#include "point.cu"
#include "graph_functions.cu"
#include "diagram_functions.cu"
#include "triangulate_functions.cu"
#include "subdivision_functions.cu"

extern "C" {

/* packed YUV word of one colour, c = byte0 | byte1<<8 | byte2<<16 (graph_functions.cu:80-98) */
unsigned int ref_host_rgb_to_yuv( int c ) { return RGBtoYUV( c ); }

void ref_host_yuv_all( unsigned int* out )
{
    for( int c = 0; c < ( 1 << 24 ); c++ ) out[ c ] = RGBtoYUV( c );
}

/* one cell from a pattern (diagram_functions.cu:319); out = 45 (x,y) pairs, returns vertex count */
int ref_host_cell( int node, int node_left, int node_right, float* out )
{
    Point cell[ CELL_SIZE ];
    memset( cell, 0, sizeof( cell ) );
    int n = createCellFromPattern< Point >( ( char )node, ( char )node_left, ( char )node_right, cell, 0, 1, 1 );
    for( int t = 0; t < CELL_SIZE; t++ )
    {
        out[ 2 * t ] = cell[ t ].x;
        out[ 2 * t + 1 ] = cell[ t ].y;
    }
    return n;
}

/* Whole pipeline in kernel.cu:402-475 order.  Any output pointer may be NULL.
 *   graph_aux_out, graph_out : N bytes            (after trivial_cross / after ambiguous_cross)
 *   hull_out, poly_out, tri_out : N*45*2 floats   (after cells / after subdivision / after triangulate)
 *   hull_count_out, poly_count_out : N ints
 * stage_ms (may be NULL): 5 doubles = graph, crossings, cells, subdivision, triangulation wall ms. */
int ref_host_pipeline( const unsigned char* img, int W, int H, int ws, int subdivide,
                       unsigned char* graph_aux_out, unsigned char* graph_out,
                       float* hull_out, int* hull_count_out,
                       float* poly_out, int* poly_count_out,
                       float* tri_out, double* stage_ms )
{
    const int N = W * H;
    const int guard = W + 8;
    /* image followed by zeros: checkTJunction reads past the end (subdivision_functions.cu:187,195-197) */
    std::vector< char > image( ( size_t )ws * H + 2 * ( size_t )ws + 64, 0 );
    memcpy( image.data(), img, ( size_t )ws * H );
    /* zeroed guard bands either side of both graph buffers (kernel.cu:205, graph_functions.cu:1215-1222) */
    std::vector< char > gbuf( N + 2 * guard, 0 ), abuf( N + 2 * guard, 0 );
    char* graph = gbuf.data() + guard;
    char* graph_aux = abuf.data() + guard;
    std::vector< Point > diagram( ( size_t )N * CELL_SIZE ), diagram_aux( ( size_t )N * CELL_SIZE );
    memset( diagram.data(), 0, diagram.size() * sizeof( Point ) );
    std::vector< int > edge_count( N, 0 ), edge_count_aux( N, 0 );
    std::vector< char > edge_status( ( size_t )N * CELL_SIZE, 0 );
    std::vector< int > link_index( ( size_t )N * CELL_SIZE, 0 );
    struct timespec t0, t1;
#define TICK clock_gettime( CLOCK_MONOTONIC, &t0 )
#define TOCK( k ) do { clock_gettime( CLOCK_MONOTONIC, &t1 ); if( stage_ms ) stage_ms[ k ] = ( t1.tv_sec - t0.tv_sec ) * 1e3 + ( t1.tv_nsec - t0.tv_nsec ) * 1e-6; } while( 0 )

    /* graph_Kernel, kernel.cu:140-159 */
    TICK;
    for( int j = 0; j < H; j++ )
        for( int i = 0; i < W; i++ )
            for( int e = 0; e < 8; e++ )
            {
                if( diff( i, j, e, W, H, ws, image.data() ) )
                {
                    SET_BIT( graph[ j * W + i ], e, 0 );
                }
                else
                {
                    SET_BIT( graph[ j * W + i ], e, 1 );
                }
            }
    TOCK( 0 );
    /* trivial_cross_Kernel, kernel.cu:162-177 */
    TICK;
    for( int j = 0; j < H; j++ )
        for( int i = 0; i < W; i++ )
            crossCheck_4( graph, W, H, i, j );
    memcpy( graph_aux, graph, N ); /* kernel.cu:415 */
    if( graph_aux_out ) memcpy( graph_aux_out, graph_aux, N );
    /* ambiguous_cross_Kernel, kernel.cu:180-189 */
    for( int j = 0; j < H; j++ )
        for( int i = 0; i < W; i++ )
            crossCheck_Heuristics( graph, graph_aux, W, H, i, j, ( bool* )0 );
    TOCK( 1 );
    if( graph_out ) memcpy( graph_out, graph, N );
    /* cells_Kernel, kernel.cu:192-213 */
    TICK;
    for( int j = 0; j < H; j++ )
        for( int i = 0; i < W; i++ )
        {
            int n = j * W + i;
            edge_count[ n ] = createCellFromPattern< Point >( graph[ n ], graph[ n - 1 ], graph[ n + 1 ],
                                                              diagram.data(), n * CELL_SIZE, i, j );
        }
    TOCK( 2 );
    diagram_aux = diagram;       /* kernel.cu:434-437 */
    edge_count_aux = edge_count;
    if( hull_out ) memcpy( hull_out, diagram.data(), diagram.size() * sizeof( Point ) );
    if( hull_count_out ) memcpy( hull_count_out, edge_count.data(), N * sizeof( int ) );
    /* subdivision_Kernel, kernel.cu:216-261 */
    TICK;
    if( subdivide )
    {
        for( int j = 0; j < H; j++ )
            for( int i = 0; i < W; i++ )
            {
                int n = j * W + i;
                if( graph[ n ] == 90 ) continue;
                edge_count[ n ] = subdivision< Point >( image.data(), &diagram[ ( size_t )n * CELL_SIZE ], diagram_aux.data(),
                                                        edge_count_aux.data(), n, W, ws, H, i, j, graph[ n ],
                                                        ( bool* )&edge_status[ ( size_t )n * CELL_SIZE ],
                                                        &link_index[ ( size_t )n * CELL_SIZE ] );
            }
    }
    TOCK( 3 );
    if( poly_out ) memcpy( poly_out, diagram.data(), diagram.size() * sizeof( Point ) );
    if( poly_count_out ) memcpy( poly_count_out, edge_count.data(), N * sizeof( int ) );
    /* triangulate_Kernel, kernel.cu:264-282 */
    TICK;
    if( tri_out )
    {
        for( int j = 0; j < H; j++ )
            for( int i = 0; i < W; i++ )
            {
                int n = j * W + i;
                triangulate_polygon< Point >( &diagram[ ( size_t )n * CELL_SIZE ], edge_count.data(), n, CELL_SIZE, i, j );
            }
        memcpy( tri_out, diagram.data(), diagram.size() * sizeof( Point ) );
    }
    TOCK( 4 );
    return 0;
}

} /* extern "C" */
