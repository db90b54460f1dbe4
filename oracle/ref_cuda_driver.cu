/* TEST INFRASTRUCTURE ONLY (oracle/).  Never linked into, imported by, or executed from the product
 * path; only tests/ and bench.py's reported-baseline leg may load the library this file builds.
 *
 * ref_cuda: the reference's CUDA translation unit kernel.cu, #include'd UNMODIFIED from where it
 * lies under /root/reference and compiled for sm_100a with the reference's own flags
 * (-use_fast_math -fno-strict-aliasing, depixel_gpu.pro:47-50,84).  The only thing kernel.cu lacks
 * in this image is the CUDA-samples header helper_timer.h (kernel.cu:4); oracle/shim/ supplies the
 * six calls it uses.  This file adds a headless driver (the reference's only caller maps OpenGL
 * VBOs, simpleVBO.cpp:146-153; here pos/colorPos are plain cudaMalloc buffers) and two probe
 * kernels that call the reference's own __device__ functions directly.
 */
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <unistd.h>
#include <fcntl.h>

#include "kernel.cu" /* the reference, as is: defines extern "C" launch_kernel and every __device__ routine */

namespace {

/* the reference printf()s three lines per call (kernel.cu:396-397,486); park stdout meanwhile */
struct QuietStdout
{
    int saved;
    QuietStdout()
    {
        fflush( stdout );
        saved = dup( 1 );
        int nul = open( "/dev/null", O_WRONLY );
        dup2( nul, 1 );
        close( nul );
    }
    ~QuietStdout()
    {
        fflush( stdout );
        dup2( saved, 1 );
        close( saved );
    }
};

__global__ void probe_yuv_all( unsigned int* out )
{
    unsigned int c = blockIdx.x * blockDim.x + threadIdx.x;
    if( c < ( 1u << 24 ) ) out[ c ] = RGBtoYUV( ( int )c ); /* graph_functions.cu:80 */
}

/* stage A + B only, through the reference's own kernels, to expose graph_aux (kernel.cu:402-415) */
} // namespace

extern "C" {

/* packed YUV word of all 2^24 colours as the DEVICE build of graph_functions.cu:80-98 computes it */
int ref_cuda_yuv_all( unsigned int* out_host )
{
    unsigned int* d = nullptr;
    if( cudaMalloc( &d, sizeof( unsigned int ) << 24 ) != cudaSuccess ) return 1;
    probe_yuv_all<<< ( 1u << 24 ) / 256, 256 >>>( d );
    if( cudaDeviceSynchronize() != cudaSuccess ) return 2;
    cudaMemcpy( out_host, d, sizeof( unsigned int ) << 24, cudaMemcpyDeviceToHost );
    cudaFree( d );
    return 0;
}

/* graph after trivial_cross_Kernel (what the reference copies into graph_aux, kernel.cu:415) */
int ref_cuda_graph_aux( const char* img, int W, int H, int ws, char* graph_aux_host )
{
    char *img_d = nullptr, *g_d = nullptr;
    size_t img_size = ( size_t )ws * H;
    cudaMalloc( &img_d, img_size );
    cudaMalloc( &g_d, ( size_t )W * H );
    cudaMemset( g_d, 0, ( size_t )W * H );
    cudaMemcpy( img_d, img, img_size, cudaMemcpyHostToDevice );
    dim3 tpb( 2, 2 ), nb( ( W + 1 ) / 2, ( H + 1 ) / 2 ); /* kernel.cu:392-394 */
    graph_Kernel<<< nb, tpb >>>( img_d, ( int )img_size, W, H, ws, g_d );
    trivial_cross_Kernel<<< nb, tpb >>>( W, H, g_d );
    int rc = cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
    cudaMemcpy( graph_aux_host, g_d, ( size_t )W * H, cudaMemcpyDeviceToHost );
    cudaFree( img_d );
    cudaFree( g_d );
    return rc;
}

/* One launch_kernel call exactly as simpleVBO.cpp:151-153 makes it.  Outputs (host, may be NULL):
 * graph N bytes, edge_count N ints, diagram N*45 Points (triangle list, the returned pointer's
 * content), pos N*45 float2, color N*45 uchar4.  wall_ms = host wall time of the call itself. */
int ref_cuda_launch( const char* img, int W, int H, int ws, int subdivide, char* graph_out, int* edge_count_out,
                     float* diagram_out, float* pos_out, unsigned char* color_out, double* wall_ms )
{
    size_t N = ( size_t )W * H;
    float2* pos = nullptr;
    uchar4* col = nullptr;
    if( cudaMalloc( &pos, N * CELL_SIZE * sizeof( float2 ) ) != cudaSuccess ) return 1;
    if( cudaMalloc( &col, N * CELL_SIZE * sizeof( uchar4 ) ) != cudaSuccess ) return 1;
    char* graph_h = ( char* )calloc( N, 1 );
    int* ec_h = ( int* )calloc( N, sizeof( int ) );
    Point* diagram_h;
    struct timespec t0, t1;
    {
        QuietStdout q;
        clock_gettime( CLOCK_MONOTONIC, &t0 );
        diagram_h = launch_kernel( pos, col, 0.0f, ( char* )img, W, H, ws, ec_h, graph_h, subdivide != 0 );
        clock_gettime( CLOCK_MONOTONIC, &t1 );
    }
    if( wall_ms ) *wall_ms = ( t1.tv_sec - t0.tv_sec ) * 1e3 + ( t1.tv_nsec - t0.tv_nsec ) * 1e-6;
    if( graph_out ) memcpy( graph_out, graph_h, N );
    if( edge_count_out ) memcpy( edge_count_out, ec_h, N * sizeof( int ) );
    if( diagram_out ) memcpy( diagram_out, diagram_h, N * CELL_SIZE * sizeof( Point ) );
    if( pos_out ) cudaMemcpy( pos_out, pos, N * CELL_SIZE * sizeof( float2 ), cudaMemcpyDeviceToHost );
    if( color_out ) cudaMemcpy( color_out, col, N * CELL_SIZE * sizeof( uchar4 ), cudaMemcpyDeviceToHost );
    free( diagram_h ); /* caller frees, simpleVBO.cpp:437 */
    free( graph_h );
    free( ec_h );
    cudaFree( pos );
    cudaFree( col );
    return 0;
}

/* frames/s of back-to-back launch_kernel calls on one frame (the reference's design: alloc + H2D +
 * kernels + D2H + free every call).  Returns average ms per call. */
double ref_cuda_time_calls( const char* img, int W, int H, int ws, int subdivide, int calls )
{
    size_t N = ( size_t )W * H;
    float2* pos = nullptr;
    uchar4* col = nullptr;
    cudaMalloc( &pos, N * CELL_SIZE * sizeof( float2 ) );
    cudaMalloc( &col, N * CELL_SIZE * sizeof( uchar4 ) );
    char* graph_h = ( char* )calloc( N, 1 );
    int* ec_h = ( int* )calloc( N, sizeof( int ) );
    struct timespec t0, t1;
    {
        QuietStdout q;
        free( launch_kernel( pos, col, 0.0f, ( char* )img, W, H, ws, ec_h, graph_h, subdivide != 0 ) ); /* warm-up */
        clock_gettime( CLOCK_MONOTONIC, &t0 );
        for( int c = 0; c < calls; c++ )
            free( launch_kernel( pos, col, 0.0f, ( char* )img, W, H, ws, ec_h, graph_h, subdivide != 0 ) );
        clock_gettime( CLOCK_MONOTONIC, &t1 );
    }
    free( graph_h );
    free( ec_h );
    cudaFree( pos );
    cudaFree( col );
    return ( ( t1.tv_sec - t0.tv_sec ) * 1e3 + ( t1.tv_nsec - t0.tv_nsec ) * 1e-6 ) / calls;
}

} /* extern "C" */
