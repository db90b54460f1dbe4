// remaster_cli — the reference's program entry point (main.cpp:166-194) without the GLUT window:
// image in, remastered image out.
//
//   remaster_cli <input image> [-o out.png] [-s scale] [--aa 2|4] [--no-subdivide] [--graph g.pgm] [--labels l.pgm]
//                [--graph-image g.png] [--draw-graph g.pgm] [--strips N] [--device D] [--convert-only]
//   remaster_cli <image that gives the frame size> --raw-video in.bin --raw-out out.bin [--skip K] [--frames M] [-s scale] ...
//
// --raw-video is the stream the reference's author had wired in and commented out (main.cpp:67-75, simpleVBO.cpp:154): a
// file of consecutive frames of img_height * img_widthstep bytes each, in the layout launch_kernel is given (BGR8, row 0 =
// bottom scanline), the image in argv[1] only supplying the size; --skip frames are passed over (the reference seeks past
// 40), at most --frames are read (2000 there).  The frames go through the batch entry point 64 at a time; --raw-out
// receives the remastered frames back to back, (s*height) x (s*width) RGBA8, rows in the same bottom-up order.
//
// --graph-image draws the similarity graph the way the reference's debug view does (printToImage / display_graph,
// main.cpp:79-163): a white image, scale_graph = 20 pixels per source pixel, a black stroke from each pixel's centre
// half-way towards every linked neighbour.  --draw-graph takes the graph from a plane written by --graph instead of
// computing it (no GPU needed).
//
// Kept from the reference: argv[1] is the input path (main.cpp:172-173); load_image() loads in colour,
// flips the image vertically so that row 0 is the bottom scanline, and publishes img_data / img_width /
// img_height / img_nchannels / img_widthstep (main.cpp:50-64); allocate_graph() makes the host graph
// (main.cpp:141-147).  Where the reference enters glutMainLoop and draws 45 vertices per pixel through
// OpenGL, this program calls the C ABI once and writes the rasterized image (un-flipped again).
#include "Image.h"
#include "pixelart_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

// the reference's globals (main.cpp:31-38)
Image* img = nullptr;
char* img_data = nullptr;
int img_width = 0, img_height = 0, img_nchannels = 0, img_widthstep = 0;
char* graph = nullptr;

static bool load_image( const char* path )
{
    img = new Image();
    img->loadImage( path, CV_LOAD_IMAGE_COLOR );
    if( !img->ok() )
    {
        fprintf( stderr, "remaster_cli: %s\n", img->error().c_str() );
        return false;
    }
    img->reverses(); // main.cpp:59
    img_data = img->getImageData();
    img_width = img->getWidth();
    img_height = img->getHeight();
    img_nchannels = img->getNchannels();
    img_widthstep = img->getWidthStep();
    return true;
}

static void allocate_graph() { graph = ( char* )calloc( ( size_t )img_width * img_height, 1 ); } // main.cpp:141-147

int scale_graph = 20;       // main.cpp:24
Image* graph_img = nullptr; // main.cpp:28

// black 1-pixel stroke (the reference draws with cvLine(..., CV_AA): OpenCV's anti-aliased line is outside this path,
// a plain Bresenham line stands in for it)
static void draw_line( Image* im, int x0, int y0, int x1, int y1 )
{
    const int dx = abs( x1 - x0 ), dy = -abs( y1 - y0 ), sx = x0 < x1 ? 1 : -1, sy = y0 < y1 ? 1 : -1;
    int err = dx + dy;
    for( ;; )
    {
        if( x0 >= 0 && y0 >= 0 && x0 < im->getWidth() && y0 < im->getHeight() )
        {
            char* p = im->getImageData() + ( size_t )y0 * im->getWidthStep() + 3 * x0;
            p[ 0 ] = p[ 1 ] = p[ 2 ] = 0;
        }
        if( x0 == x1 && y0 == y1 ) break;
        const int e2 = 2 * err;
        if( e2 >= dy ) { err += dy; x0 += sx; }
        if( e2 <= dx ) { err += dx; y0 += sy; }
    }
}

/* Make a image representation of the graph (main.cpp:79-139): bit e of a node <-> neighbour (di,dj) of graph_functions.cu:162-171 */
void printToImage( char* graph, Image* src, Image* img_out )
{
    static const int di[ 8 ] = { -1, 0, 1, -1, 1, -1, 0, 1 }, dj[ 8 ] = { 1, 1, 1, 0, 0, -1, -1, -1 };
    const int half_sg = scale_graph / 2;
    for( int j = 0; j < src->getHeight(); j++ )
        for( int i = 0; i < src->getWidth(); i++ )
        {
            const int index = j * src->getWidth() + i;
            const int n_j = j * scale_graph + half_sg - 1, n_i = i * scale_graph + half_sg - 1;
            for( int e = 0; e < 8; e++ )
                if( CHECK_BIT( graph[ index ], e ) ) draw_line( img_out, n_i + di[ e ] * half_sg, n_j + dj[ e ] * half_sg, n_i, n_j );
        }
}

// display_graph (main.cpp:152-163) into a file: row 0 of the drawing is the bottom scanline (the frame is flipped,
// glDrawPixels shows it upright), so it is flipped back before it is stored
static bool save_graph_image( const char* path )
{
    graph_img = new Image();
    graph_img->createImage( img->getWidth() * scale_graph, img->getHeight() * scale_graph, IPL_DEPTH_8U, 3 );
    graph_img->setAllPixels( 255, 255, 255 );
    printToImage( graph, img, graph_img );
    graph_img->reverses();
    graph_img->saveImage( path );
    const bool ok = graph_img->error().empty();
    if( !ok ) fprintf( stderr, "remaster_cli: %s\n", graph_img->error().c_str() );
    delete graph_img;
    graph_img = nullptr;
    return ok;
}

static bool save_plane( const char* path, const void* data, int w, int h, int bytes_per_px )
{
    // 1 byte/px -> PGM/PNG grey; 4 byte labels -> RGBA PNG whose R,G,B,A bytes are the little-endian bytes of the int32
    // (lossless).  A 4-channel Image holds B,G,R,A (saveImage swaps channels 0 and 2 on the way out), so the label bytes
    // go in pre-swapped.
    Image out;
    out.createImage( w, h, IPL_DEPTH_8U, bytes_per_px == 1 ? 1 : 4 );
    for( int y = 0; y < h; y++ ) // stored top scanline first
    {
        char* o = out.getImageData() + ( size_t )y * out.getWidthStep();
        const char* s = ( const char* )data + ( size_t )( h - 1 - y ) * w * bytes_per_px;
        if( bytes_per_px == 1 )
            memcpy( o, s, ( size_t )w );
        else
            for( int x = 0; x < w; x++ ) { o[ 4 * x ] = s[ 4 * x + 2 ]; o[ 4 * x + 1 ] = s[ 4 * x + 1 ]; o[ 4 * x + 2 ] = s[ 4 * x ]; o[ 4 * x + 3 ] = s[ 4 * x + 3 ]; }
    }
    out.saveImage( path );
    return out.error().empty();
}

// --outlines: the spline outline of every connected component (par_outlines_host: graph, labels, border walks, closed
// quadratic B-splines over the walks — the stage the reference's extractBorderPoints, cc_functions.cu:348-503, was written
// for) as an SVG, one filled path per component in its own colour, coordinates in output pixels, top scanline first.
static bool save_outlines_svg( const char* path, int device, int scale )
{
    par_context* ctx = nullptr;
    if( par_create( &ctx, device, img_width, img_height, 1 ) != PAR_OK )
    {
        fprintf( stderr, "remaster_cli: %s\n", par_last_error( nullptr ) );
        return false;
    }
    par_outlines o;
    if( par_outlines_host( ctx, reinterpret_cast< const uint8_t* >( img_data ), img_width, img_height, img_widthstep, 4, &o ) != PAR_OK )
    {
        fprintf( stderr, "remaster_cli: %s\n", par_last_error( ctx ) );
        par_destroy( ctx );
        return false;
    }
    FILE* f = fopen( path, "w" );
    bool ok = f != nullptr;
    if( ok )
    {
        fprintf( f, "<svg xmlns=\"http://www.w3.org/2000/svg\" width=\"%d\" height=\"%d\" viewBox=\"0 0 %d %d\">\n", img_width * scale, img_height * scale,
                 img_width * scale, img_height * scale );
        const float* p = o.points;
        for( int k = 0; k < o.n_walks; k++ )
        {
            const int n = o.start[ k ], x = n % img_width, y = n / img_width;
            const unsigned char* c = reinterpret_cast< const unsigned char* >( img_data ) + ( size_t )y * img_widthstep + 3 * x; // B, G, R
            fprintf( f, "<path fill=\"#%02x%02x%02x\" d=\"", c[ 2 ], c[ 1 ], c[ 0 ] );
            const long long m = ( long long )o.count[ k ] * o.samples;
            for( long long t = 0; t < m; t++, p += 2 )
                fprintf( f, "%c%.8g %.8g", t ? 'L' : 'M', p[ 0 ] * scale, ( img_height - p[ 1 ] ) * scale ); // (row 0 of the path is the bottom scanline)
            fprintf( f, "Z\"/>\n" );
        }
        fprintf( f, "</svg>\n" );
        ok = fclose( f ) == 0;
    }
    if( !ok ) fprintf( stderr, "remaster_cli: cannot write %s\n", path );
    par_outlines_free( &o );
    par_destroy( ctx );
    return ok;
}

int main( int argc, char** argv )
{
    if( argc < 2 )
    {
        fprintf( stderr, "usage: %s <input image> [-o out.png] [-s scale] [--aa 2|4] [--no-subdivide] [--graph g.pgm] [--labels l.png] [--graph-image g.png] [--outlines o.svg] [--draw-graph g.pgm] [--strips N] [--device D] [--convert-only] [--raw-video in.bin --raw-out out.bin [--skip K] [--frames M]]\n", argv[ 0 ] );
        return 2;
    }
    std::string out_path = "remastered.png", graph_path, labels_path, graph_image_path, draw_graph_path, raw_video_path, raw_out_path, outlines_path;
    long skip_frames = 0, max_video_frames = 2000;
    int scale = 4, strips = 0, device = 0, aa = 1; // aa: anti-aliasing samples per axis (the reference's GL_MULTISAMPLE toggle, simpleVBO.cpp:238-253)
    bool subdivide = true, convert_only = false;
    for( int k = 2; k < argc; k++ )
    {
        std::string a = argv[ k ];
        if( a == "-o" && k + 1 < argc ) out_path = argv[ ++k ];
        else if( a == "-s" && k + 1 < argc ) scale = atoi( argv[ ++k ] );
        else if( a == "--aa" && k + 1 < argc ) aa = atoi( argv[ ++k ] );
        else if( a == "--no-subdivide" ) subdivide = false;
        else if( a == "--graph" && k + 1 < argc ) graph_path = argv[ ++k ];
        else if( a == "--labels" && k + 1 < argc ) labels_path = argv[ ++k ];
        else if( a == "--graph-image" && k + 1 < argc ) graph_image_path = argv[ ++k ];
        else if( a == "--outlines" && k + 1 < argc ) outlines_path = argv[ ++k ];
        else if( a == "--draw-graph" && k + 1 < argc ) draw_graph_path = argv[ ++k ];
        else if( a == "--raw-video" && k + 1 < argc ) raw_video_path = argv[ ++k ];
        else if( a == "--raw-out" && k + 1 < argc ) raw_out_path = argv[ ++k ];
        else if( a == "--skip" && k + 1 < argc ) skip_frames = atol( argv[ ++k ] );
        else if( a == "--frames" && k + 1 < argc ) max_video_frames = atol( argv[ ++k ] );
        else if( a == "--strips" && k + 1 < argc ) strips = atoi( argv[ ++k ] );
        else if( a == "--device" && k + 1 < argc ) device = atoi( argv[ ++k ] );
        else if( a == "--convert-only" ) convert_only = true;
        else { fprintf( stderr, "remaster_cli: unknown option %s\n", a.c_str() ); return 2; }
    }
    if( ( aa != 1 && aa != 2 && aa != 4 ) || scale < 1 || scale * aa > 8 ) // (before anything is sized from the scale)
    {
        fprintf( stderr, "remaster_cli: unsupported scale %d with --aa %d (scale x aa must be an integer 1..8, aa 1, 2 or 4)\n", scale, aa );
        return 2;
    }
    if( !load_image( argv[ 1 ] ) ) return 1;
    allocate_graph();
    if( !draw_graph_path.empty() ) // the graph of an earlier run (--graph) as a picture; no GPU
    {
        Image plane;
        plane.loadImage( draw_graph_path.c_str(), CV_LOAD_IMAGE_GRAYSCALE );
        if( !plane.ok() || plane.getWidth() != img_width || plane.getHeight() != img_height || plane.getNchannels() != 1 )
        {
            fprintf( stderr, "remaster_cli: %s is not a %dx%d graph plane\n", draw_graph_path.c_str(), img_width, img_height );
            return 1;
        }
        for( int y = 0; y < img_height; y++ ) // planes are stored top scanline first, the graph's row 0 is the bottom one
            memcpy( graph + ( size_t )( img_height - 1 - y ) * img_width, plane.getImageData() + ( size_t )y * plane.getWidthStep(), ( size_t )img_width );
        return save_graph_image( graph_image_path.empty() ? out_path.c_str() : graph_image_path.c_str() ) ? 0 : 1;
    }
    if( convert_only ) // image I/O round trip only (no GPU): load -> flip -> flip back -> save
    {
        img->reverses();
        img->saveImage( out_path.c_str() );
        if( !img->error().empty() ) { fprintf( stderr, "remaster_cli: %s\n", img->error().c_str() ); return 1; }
        return 0;
    }
    if( !raw_video_path.empty() ) // a stream of raw frames of the image's size through the batch entry point
    {
        if( raw_out_path.empty() ) { fprintf( stderr, "remaster_cli: --raw-video needs --raw-out\n" ); return 2; }
        FILE* in = fopen( raw_video_path.c_str(), "rb" );
        FILE* outf = in ? fopen( raw_out_path.c_str(), "wb" ) : nullptr;
        if( !in || !outf ) { fprintf( stderr, "remaster_cli: cannot open %s\n", in ? raw_out_path.c_str() : raw_video_path.c_str() ); return 1; }
        const size_t frame_bytes = ( size_t )img_height * img_widthstep, out_bytes = ( size_t )img_width * img_height * scale * scale * 4;
        const int batch = 64;
        if( fseek( in, ( long )( skip_frames * ( long )frame_bytes ), SEEK_SET ) != 0 ) { fprintf( stderr, "remaster_cli: cannot seek in %s\n", raw_video_path.c_str() ); return 1; }
        par_context* ctx = nullptr;
        if( par_create( &ctx, device, img_width, img_height, batch ) != PAR_OK ) { fprintf( stderr, "remaster_cli: %s\n", par_last_error( nullptr ) ); return 1; }
        std::vector< uint8_t > frames( frame_bytes * batch ), result( out_bytes * batch );
        long done = 0;
        while( done < max_video_frames )
        {
            const long want = max_video_frames - done < batch ? max_video_frames - done : batch;
            const int n = ( int )( fread( frames.data(), 1, frame_bytes * ( size_t )want, in ) / frame_bytes ); // (a trailing partial frame is dropped)
            if( n == 0 ) break;
            par_job vj;
            memset( &vj, 0, sizeof( vj ) );
            vj.bgr = frames.data();
            vj.width = img_width;
            vj.height = img_height;
            vj.widthstep = img_widthstep;
            vj.n_frames = n;
            vj.scale = scale;
            vj.flags = subdivide ? PAR_FLAG_SUBDIVIDE : 0u;
            if( aa == 2 ) vj.flags |= PAR_FLAG_AA2;
            if( aa == 4 ) vj.flags |= PAR_FLAG_AA4;
            vj.rgba = result.data();
            if( par_remaster_host( ctx, &vj ) != PAR_OK ) { fprintf( stderr, "remaster_cli: %s\n", par_last_error( ctx ) ); return 1; }
            if( fwrite( result.data(), 1, out_bytes * ( size_t )n, outf ) != out_bytes * ( size_t )n ) { fprintf( stderr, "remaster_cli: cannot write %s\n", raw_out_path.c_str() ); return 1; }
            done += n;
        }
        par_destroy( ctx );
        fclose( in );
        if( fclose( outf ) != 0 ) { fprintf( stderr, "remaster_cli: cannot write %s\n", raw_out_path.c_str() ); return 1; }
        printf( "%s: %ld frames of %dx%d -> %dx%d RGBA (scale %d, subdivide %s) -> %s\n", raw_video_path.c_str(), done, img_width, img_height, img_width * scale,
                img_height * scale, scale, subdivide ? "on" : "off", raw_out_path.c_str() );
        return 0;
    }
    const size_t N = ( size_t )img_width * img_height;
    // the result arrives as the 3-channel B,G,R image Image::saveImage takes (Image.cpp:64-71): PAR_OUT_BGR8, dense rows
    std::vector< uint8_t > bgr_out( N * scale * scale * 3 );
    std::vector< int32_t > labels( labels_path.empty() ? 0 : N );
    par_job job;
    memset( &job, 0, sizeof( job ) );
    job.bgr = reinterpret_cast< const uint8_t* >( img_data );
    job.width = img_width;
    job.height = img_height;
    job.widthstep = img_widthstep;
    job.n_frames = 1;
    job.scale = scale;
    job.flags = ( subdivide ? PAR_FLAG_SUBDIVIDE : 0u ) | PAR_FLAG_FLIP_OUTPUT; // output rows top scanline first
    if( aa == 2 ) job.flags |= PAR_FLAG_AA2;
    if( aa == 4 ) job.flags |= PAR_FLAG_AA4;
    job.rgba = bgr_out.data();
    job.out_format = PAR_OUT_BGR8;
    job.graph = reinterpret_cast< uint8_t* >( graph );
    job.labels = labels.empty() ? nullptr : labels.data();
    if( strips > 0 )
    {
        int n_dev = 0;
        std::vector< int > devs( strips );
        for( int k = 0; k < strips; k++ ) devs[ k ] = device + k; // consecutive devices, wrapped below by the library's own check
        par_group* grp = nullptr;
        if( par_group_create( &grp, devs.data(), strips, img_width, img_height, scale ) != PAR_OK )
        {
            // fewer devices than strips: put several strips on the devices that exist
            ( void )n_dev;
            for( int k = 0; k < strips; k++ ) devs[ k ] = device;
            if( par_group_create( &grp, devs.data(), strips, img_width, img_height, scale ) != PAR_OK )
            {
                fprintf( stderr, "remaster_cli: %s\n", par_group_last_error( nullptr ) );
                return 1;
            }
        }
        if( par_group_remaster_host( grp, &job ) != PAR_OK )
        {
            fprintf( stderr, "remaster_cli: %s\n", par_group_last_error( grp ) );
            return 1;
        }
        par_group_destroy( grp );
    }
    else
    {
        par_context* ctx = nullptr;
        if( par_create( &ctx, device, img_width, img_height, 1 ) != PAR_OK )
        {
            fprintf( stderr, "remaster_cli: %s\n", par_last_error( nullptr ) );
            return 1;
        }
        if( par_remaster_host( ctx, &job ) != PAR_OK )
        {
            fprintf( stderr, "remaster_cli: %s\n", par_last_error( ctx ) );
            return 1;
        }
        par_destroy( ctx );
    }
    Image out;
    out.createImage( img_width * scale, img_height * scale, IPL_DEPTH_8U, 3 );
    for( int y = 0; y < img_height * scale; y++ ) // (an Image's rows are padded to 4 bytes, IplImage style)
        memcpy( out.getImageData() + ( size_t )y * out.getWidthStep(), &bgr_out[ ( size_t )y * img_width * scale * 3 ], ( size_t )img_width * scale * 3 );
    out.saveImage( out_path.c_str() );
    if( !out.error().empty() ) { fprintf( stderr, "remaster_cli: %s\n", out.error().c_str() ); return 1; }
    if( !graph_path.empty() && !save_plane( graph_path.c_str(), graph, img_width, img_height, 1 ) ) return 1;
    if( !labels_path.empty() && !save_plane( labels_path.c_str(), labels.data(), img_width, img_height, 4 ) ) return 1;
    if( !graph_image_path.empty() && !save_graph_image( graph_image_path.c_str() ) ) return 1;
    if( !outlines_path.empty() && !save_outlines_svg( outlines_path.c_str(), device, scale ) ) return 1;
    printf( "%s: %dx%d -> %dx%d (scale %d, subdivide %s) -> %s\n", argv[ 1 ], img_width, img_height, img_width * scale, img_height * scale, scale,
            subdivide ? "on" : "off", out_path.c_str() );
    free( graph );
    delete img;
    return 0;
}
