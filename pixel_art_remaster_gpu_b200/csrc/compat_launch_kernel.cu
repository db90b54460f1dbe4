// The reference's own entry point, same symbol and signature (kernel.cu:286-288):
//
//   extern "C" Point* launch_kernel( float2* pos, uchar4* colorPos, float time, char* img_data,
//                                    int img_width, int img_height, int img_widthstep,
//                                    int* edge_count_h, char* graph_h, bool subdivide )
//
// so that a GL consumer such as the reference's runCuda() (simpleVBO.cpp:131-160) links against this
// library unchanged.  Semantics kept: pos / colorPos are DEVICE arrays of 45 entries
// per pixel (the mapped VBOs); img_data is the host BGR8 frame, row 0 = bottom; edge_count_h receives
// the polygon vertex count per pixel; graph_h the final similarity graph; the return value is a
// malloc()ed host array of 45 Points per pixel holding the triangle list in global coordinates, which
// the caller free()s (simpleVBO.cpp:437); any CUDA error prints to stderr and exit(-1)s
// (kernel.cu:24-62).  `time` is unused, as in the reference.
//
// What differs is everything behind it: one cached context instead of 10 cudaMalloc/cudaFree per
// call, the graph/crossing/polygon kernels of this library, and an ear-clipping kernel that works in
// exact 1/64-pixel integers (the reference's float cross products are exact on these coordinates, so
// the triangle lists are identical: triangulate_functions.cu:108-217, :70-105, :41-67, :6-34).
// Slots past the triangle list are (-100,-100) in pos (position_kernel, kernel.cu:125-134) and zero in
// the returned array (uninitialised stack in the reference, triangulate_functions.cu:222,262-266).
#define PAR_HAVE_CUDA_VECTOR_TYPES 1
#include "../../include/pixelart_b200.h"
#include "kernels.cuh"

#include <cstdio>
#include <cstdlib>

namespace {

constexpr int kSlots = PAR_CELL_SLOTS;

// ear clipping of one polygon (vertices in 1/64 pixel, cell-local), triangles out as vertex indices
__device__ int ear_clip( const int* px, const int* py, int n, uint8_t* tri /* 3 per triangle */ )
{
    if( n < 3 ) return 0;
    int V[ 16 ];
    long twice_area = 0;
    for( int p = n - 1, q = 0; q < n; p = q++ ) twice_area += ( long )px[ p ] * py[ q ] - ( long )px[ q ] * py[ p ];
    for( int v = 0; v < n; v++ ) V[ v ] = twice_area > 0 ? v : n - 1 - v; // counter-clockwise order (:128-141)
    int nv = n, made = 0, guard = 2 * nv;
    for( int v = nv - 1; nv > 2; )
    {
        if( guard-- <= 0 ) break; // "probably a non-simple polygon" (:153-156)
        int u = v;
        if( nv <= u ) u = 0;
        v = u + 1;
        if( nv <= v ) v = 0;
        int w = v + 1;
        if( nv <= w ) w = 0;
        const int ax = px[ V[ u ] ], ay = py[ V[ u ] ], bx = px[ V[ v ] ], by = py[ V[ v ] ], cx = px[ V[ w ] ], cy = py[ V[ w ] ];
        bool ear = ( bx - ax ) * ( cy - ay ) - ( by - ay ) * ( cx - ax ) > 0; // EPSILON test (:88-91): exact values are multiples of 1/4096
        for( int p = 0; ear && p < nv; p++ )
        {
            if( p == u || p == v || p == w ) continue;
            const int qx = px[ V[ p ] ], qy = py[ V[ p ] ];
            const int e0 = ( cx - bx ) * ( qy - by ) - ( cy - by ) * ( qx - bx );
            const int e1 = ( bx - ax ) * ( qy - ay ) - ( by - ay ) * ( qx - ax );
            const int e2 = ( ax - cx ) * ( qy - cy ) - ( ay - cy ) * ( qx - cx );
            if( e0 >= 0 && e1 >= 0 && e2 >= 0 ) ear = false; // a vertex inside or on the ear blocks it (:41-67)
        }
        if( !ear ) continue;
        tri[ 3 * made ] = ( uint8_t )V[ u ];
        tri[ 3 * made + 1 ] = ( uint8_t )V[ v ];
        tri[ 3 * made + 2 ] = ( uint8_t )V[ w ];
        made++;
        for( int s = v, t = v + 1; t < nv; s++, t++ ) V[ s ] = V[ t ];
        nv--;
        guard = 2 * nv;
    }
    return made;
}

// polygons (45 float pairs, local) -> triangle list in place (global coordinates) + the two VBO arrays
__global__ void __launch_bounds__( 128 ) triangulate_kernel( float* diagram, const int32_t* count, float2* pos, uchar4* col, const uint8_t* bgr,
                                                             int width, int height, int widthstep )
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if( n >= width * height ) return;
    const int i = n % width, j = n / width;
    float* cell = diagram + ( size_t )n * 2 * kSlots;
    const int cnt = min( count[ n ], 16 );
    int px[ 16 ], py[ 16 ];
    for( int t = 0; t < cnt; t++ )
    {
        px[ t ] = __float2int_rn( cell[ 2 * t ] * 64.0f );
        py[ t ] = __float2int_rn( cell[ 2 * t + 1 ] * 64.0f );
    }
    uint8_t tri[ 3 * 14 ];
    const int made = ear_clip( px, py, cnt, tri );
    const uint8_t* c = bgr + ( size_t )j * widthstep + 3 * i;
    const uchar4 rgba = make_uchar4( c[ 2 ], c[ 1 ], c[ 0 ], 255 ); // color_kernel, kernel.cu:98-101
    const int shown = ( count[ n ] - 2 ) * 3;                         // position_kernel exposes (count-2)*3 slots
    for( int t = 0; t < kSlots; t++ )
    {
        float x = 0.0f, y = 0.0f;
        if( t < 3 * made )
        {
            x = ( float )px[ tri[ t ] ] * 0.015625f + ( float )i; // + pixel offset (triangulate_functions.cu:262-266)
            y = ( float )py[ tri[ t ] ] * 0.015625f + ( float )j;
        }
        cell[ 2 * t ] = x;
        cell[ 2 * t + 1 ] = y;
        pos[ ( size_t )n * kSlots + t ] = t < shown ? make_float2( x, y ) : make_float2( -100.0f, -100.0f );
        col[ ( size_t )n * kSlots + t ] = rgba;
    }
}

void die( const char* what, const char* msg )
{
    fprintf( stderr, "launch_kernel: %s: %s\n", what, msg );
    exit( -1 );
}

struct Cache
{
    par_context* ctx = nullptr;
    int w = 0, h = 0, device = -1;
    uint8_t *d_img = nullptr, *d_graph = nullptr;
    float* d_diagram = nullptr;
    int32_t* d_count = nullptr;
    size_t img_bytes = 0;
} g_cache;

} // namespace

extern "C" par_point* launch_kernel( float2* pos, uchar4* colorPos, float /*time*/, char* img_data, int img_width, int img_height,
                                     int img_widthstep, int* edge_count_h, char* graph_h, bool subdivide )
{
    const size_t N = ( size_t )img_width * img_height;
    const size_t img_bytes = ( size_t )img_widthstep * img_height;
    Cache& c = g_cache;
    int dev = 0;
    if( cudaGetDevice( &dev ) != cudaSuccess ) die( "cudaGetDevice", "no CUDA device" );
    if( !c.ctx || c.w != img_width || c.h != img_height || c.img_bytes != img_bytes || c.device != dev )
    {
        if( c.ctx )
        {
            cudaSetDevice( c.device ); // (its buffers live there)
            par_destroy( c.ctx );
            cudaFree( c.d_img );
            cudaFree( c.d_graph );
            cudaFree( c.d_diagram );
            cudaFree( c.d_count );
            c = Cache();
            cudaSetDevice( dev );
        }
        if( par_create( &c.ctx, dev, img_width, img_height, 1 ) != PAR_OK ) die( "par_create", par_last_error( nullptr ) );
        par_set_stream( c.ctx, nullptr ); // the reference runs on the default stream (kernel.cu:402-475)
        cudaError_t e = cudaMalloc( &c.d_img, img_bytes + 16 );
        if( e == cudaSuccess ) e = cudaMalloc( &c.d_graph, N );
        if( e == cudaSuccess ) e = cudaMalloc( &c.d_diagram, N * kSlots * sizeof( float2 ) );
        if( e == cudaSuccess ) e = cudaMalloc( &c.d_count, N * sizeof( int32_t ) );
        if( e != cudaSuccess ) die( "cudaMalloc", cudaGetErrorString( e ) );
        c.w = img_width;
        c.h = img_height;
        c.img_bytes = img_bytes;
        c.device = dev;
    }
    cudaError_t e = cudaMemcpy( c.d_img, img_data, img_bytes, cudaMemcpyHostToDevice ); // kernel.cu:319
    if( e != cudaSuccess ) die( "cudaMemcpy H2D", cudaGetErrorString( e ) );
    par_job j = {};
    j.bgr = c.d_img;
    j.width = img_width;
    j.height = img_height;
    j.widthstep = img_widthstep;
    j.n_frames = 1;
    j.scale = 1;
    j.flags = subdivide ? PAR_FLAG_SUBDIVIDE : 0;
    j.graph = c.d_graph;
    j.polygons = c.d_diagram;
    j.poly_count = c.d_count;
    if( par_remaster_device( c.ctx, &j ) != PAR_OK ) die( "par_remaster_device", par_last_error( c.ctx ) );
    triangulate_kernel<<< ( unsigned )( ( N + 127 ) / 128 ), 128 >>>( c.d_diagram, c.d_count, pos, colorPos, c.d_img, img_width, img_height,
                                                                        img_widthstep );
    e = cudaDeviceSynchronize(); // CudaCheckError, kernel.cu:478
    if( e != cudaSuccess ) die( "kernels", cudaGetErrorString( e ) );
    par_point* diagram_h = static_cast< par_point* >( malloc( N * kSlots * sizeof( par_point ) ) );
    if( !diagram_h ) die( "malloc", "out of host memory" );
    e = cudaMemcpy( edge_count_h, c.d_count, N * sizeof( int ), cudaMemcpyDeviceToHost ); // kernel.cu:488-493
    if( e == cudaSuccess ) e = cudaMemcpy( diagram_h, c.d_diagram, N * kSlots * sizeof( par_point ), cudaMemcpyDeviceToHost );
    if( e == cudaSuccess ) e = cudaMemcpy( graph_h, c.d_graph, N, cudaMemcpyDeviceToHost );
    if( e != cudaSuccess ) die( "cudaMemcpy D2H", cudaGetErrorString( e ) );
    return diagram_h;
}
