/* TEST INFRASTRUCTURE ONLY (oracle/).  The reference's DEAD-CODE border walker, cc_functions.cu
 * (extractBorderPoints, :348-503), #include'd unmodified from /root/reference and compiled as
 * serial host C++ (it cannot share a translation unit with subdivision_functions.cu: duplicate
 * c_neighbor_index, cc_functions.cu:215).  It is the only thing in the reference that defines what a
 * connected component IS (named by its first node in raster order, :394-413); the product's
 * union-find labeller is checked against it on the reference's own 24x24 fixture (alex_png.txt). */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <fcntl.h>

#define __device__
#define __global__
#define __host__
#define __align__( n ) __attribute__( ( aligned( n ) ) )
#define CELL_SIZE 45

#include "point.cu"
#include "graph_functions.cu"
#include "cc_functions.cu"

extern "C" {

/* graph: W*H bytes.  cc_list / cc_sizes: caller-allocated, capacity entries each.  Returns the number
 * of border walks the reference produced (entries of cc_sizes filled). */
int ref_host_border_walks( const char* graph, int W, int H, int* cc_list, int* cc_sizes, int capacity )
{
    char* g = ( char* )calloc( ( size_t )W * H + 2 * ( W + 8 ), 1 );
    memcpy( g + W + 8, graph, ( size_t )W * H );
    for( int k = 0; k < capacity; k++ ) cc_sizes[ k ] = -1;
    fflush( stdout );
    int saved = dup( 1 ), nul = open( "/dev/null", O_WRONLY ); /* the walker printf()s every step */
    dup2( nul, 1 );
    close( nul );
    extractBorderPoints< Point >( g + W + 8, W, H, ( Point* )0, ( int* )0, cc_list, cc_sizes );
    fflush( stdout );
    dup2( saved, 1 );
    close( saved );
    free( g );
    int n = 0;
    while( n < capacity && cc_sizes[ n ] >= 0 ) n++;
    return n;
}

} /* extern "C" */
