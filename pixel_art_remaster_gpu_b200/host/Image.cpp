// See Image.h.  Interface of the reference's Image (Image.cpp:21-223) over a plain byte vector.
#include "Image.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <zlib.h>

namespace {

int align4( int v ) { return ( v + 3 ) & ~3; } // IplImage row alignment

bool ends_with( const std::string& s, const char* suf )
{
    std::string t( suf );
    if( s.size() < t.size() ) return false;
    std::string tail = s.substr( s.size() - t.size() );
    std::transform( tail.begin(), tail.end(), tail.begin(), ::tolower );
    return tail == t;
}

bool read_file( const char* path, std::vector< unsigned char >& out )
{
    FILE* f = fopen( path, "rb" );
    if( !f ) return false;
    fseek( f, 0, SEEK_END );
    long n = ftell( f );
    fseek( f, 0, SEEK_SET );
    out.resize( n > 0 ? ( size_t )n : 0 );
    bool ok = n >= 0 && fread( out.data(), 1, out.size(), f ) == out.size();
    fclose( f );
    return ok;
}

uint32_t be32( const unsigned char* p ) { return ( uint32_t )p[ 0 ] << 24 | ( uint32_t )p[ 1 ] << 16 | ( uint32_t )p[ 2 ] << 8 | p[ 3 ]; }

// decoded image as top-down rows of `ch` interleaved channels in R,G,B[,A] / grey order
struct Decoded { int w = 0, h = 0, ch = 0; std::vector< unsigned char > px; };

// Header fields come from untrusted files: sizes are bounded before anything is allocated or indexed.
constexpr int kMaxSide = 1 << 15;               // 32768 pixels per side
constexpr size_t kMaxPixels = ( size_t )1 << 28; // 268 M pixels
bool sane_size( long long w, long long h ) { return w > 0 && h > 0 && w <= kMaxSide && h <= kMaxSide && ( size_t )w * ( size_t )h <= kMaxPixels; }

bool decode_png( const std::vector< unsigned char >& file, Decoded& out, std::string& err )
{
    static const unsigned char sig[ 8 ] = { 137, 80, 78, 71, 13, 10, 26, 10 };
    if( file.size() < 33 || memcmp( file.data(), sig, 8 ) ) { err = "not a PNG"; return false; }
    size_t pos = 8;
    int w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
    std::vector< unsigned char > idat, plte, trns;
    while( pos + 12 <= file.size() )
    {
        uint32_t len = be32( &file[ pos ] );
        const unsigned char* type = &file[ pos + 4 ];
        const unsigned char* data = &file[ pos + 8 ];
        if( pos + 12 + len > file.size() ) { err = "truncated PNG"; return false; }
        if( !memcmp( type, "IHDR", 4 ) )
        {
            if( len != 13 ) { err = "bad PNG header"; return false; }
            if( be32( data ) > ( uint32_t )kMaxSide || be32( data + 4 ) > ( uint32_t )kMaxSide ) { err = "PNG too large"; return false; }
            w = ( int )be32( data );
            h = ( int )be32( data + 4 );
            depth = data[ 8 ];
            ctype = data[ 9 ];
            interlace = data[ 12 ];
        }
        else if( !memcmp( type, "PLTE", 4 ) ) plte.assign( data, data + len );
        else if( !memcmp( type, "tRNS", 4 ) ) trns.assign( data, data + len );
        else if( !memcmp( type, "IDAT", 4 ) ) idat.insert( idat.end(), data, data + len );
        else if( !memcmp( type, "IEND", 4 ) ) break;
        pos += 12 + len;
    }
    if( !sane_size( w, h ) || interlace != 0 ) { err = "unsupported PNG (interlaced, empty or too large)"; return false; }
    int samples = ctype == 0 ? 1 : ( ctype == 2 ? 3 : ( ctype == 3 ? 1 : ( ctype == 4 ? 2 : ( ctype == 6 ? 4 : 0 ) ) ) );
    if( !samples || ( depth != 8 && !( ctype == 3 && ( depth == 1 || depth == 2 || depth == 4 ) ) ) ) { err = "unsupported PNG bit depth / colour type"; return false; }
    const int bpp = std::max( 1, samples * depth / 8 );
    const size_t stride = ( ( size_t )w * samples * depth + 7 ) / 8;
    std::vector< unsigned char > raw( ( stride + 1 ) * h );
    uLongf raw_len = ( uLongf )raw.size();
    if( uncompress( raw.data(), &raw_len, idat.data(), ( uLong )idat.size() ) != Z_OK || raw_len != raw.size() ) { err = "PNG inflate failed"; return false; }
    std::vector< unsigned char > cur( stride ), prev( stride, 0 );
    out.w = w;
    out.h = h;
    out.ch = ( ctype == 0 ) ? 1 : ( ( ctype == 4 || ctype == 6 || ( ctype == 3 && !trns.empty() ) ) ? 4 : 3 );
    if( ctype == 4 ) out.ch = 4;
    out.px.assign( ( size_t )w * h * out.ch, 255 );
    for( int y = 0; y < h; y++ )
    {
        const unsigned char* line = &raw[ ( stride + 1 ) * y ];
        const int ft = line[ 0 ];
        for( size_t i = 0; i < stride; i++ )
        {
            int a = i >= ( size_t )bpp ? cur[ i - bpp ] : 0, b = prev[ i ], c = i >= ( size_t )bpp ? prev[ i - bpp ] : 0, x = line[ 1 + i ];
            int v;
            switch( ft )
            {
                case 0: v = x; break;
                case 1: v = x + a; break;
                case 2: v = x + b; break;
                case 3: v = x + ( ( a + b ) >> 1 ); break;
                case 4: { int p = a + b - c, pa = std::abs( p - a ), pb = std::abs( p - b ), pc = std::abs( p - c ); v = x + ( pa <= pb && pa <= pc ? a : ( pb <= pc ? b : c ) ); break; }
                default: err = "bad PNG filter"; return false;
            }
            cur[ i ] = ( unsigned char )v;
        }
        unsigned char* o = &out.px[ ( size_t )y * w * out.ch ];
        for( int x = 0; x < w; x++ )
        {
            if( ctype == 3 )
            {
                int idx = depth == 8 ? cur[ x ] : ( cur[ ( x * depth ) >> 3 ] >> ( 8 - depth - ( ( x * depth ) & 7 ) ) ) & ( ( 1 << depth ) - 1 );
                for( int k = 0; k < 3; k++ ) o[ x * out.ch + k ] = ( size_t )( idx * 3 + k ) < plte.size() ? plte[ idx * 3 + k ] : 0;
                if( out.ch == 4 ) o[ x * 4 + 3 ] = ( size_t )idx < trns.size() ? trns[ idx ] : 255;
            }
            else if( ctype == 4 ) { o[ x * 4 ] = o[ x * 4 + 1 ] = o[ x * 4 + 2 ] = cur[ 2 * x ]; o[ x * 4 + 3 ] = cur[ 2 * x + 1 ]; }
            else
                for( int k = 0; k < samples; k++ ) o[ x * out.ch + k ] = cur[ x * samples + k ];
        }
        prev.swap( cur );
    }
    return true;
}

void put_be32( std::vector< unsigned char >& v, uint32_t x ) { for( int s = 24; s >= 0; s -= 8 ) v.push_back( ( unsigned char )( x >> s ) ); }

void png_chunk( std::vector< unsigned char >& f, const char* type, const std::vector< unsigned char >& data )
{
    put_be32( f, ( uint32_t )data.size() );
    size_t start = f.size();
    f.insert( f.end(), type, type + 4 );
    f.insert( f.end(), data.begin(), data.end() );
    put_be32( f, ( uint32_t )crc32( 0, &f[ start ], ( uInt )( f.size() - start ) ) );
}

// rows: top-down, `ch` channels (1 grey, 3 RGB, 4 RGBA)
bool encode_png( const char* path, const unsigned char* rows, int w, int h, int ch, size_t row_stride )
{
    std::vector< unsigned char > raw( ( ( size_t )w * ch + 1 ) * h );
    for( int y = 0; y < h; y++ )
    {
        raw[ ( ( size_t )w * ch + 1 ) * y ] = 0;
        memcpy( &raw[ ( ( size_t )w * ch + 1 ) * y + 1 ], rows + row_stride * y, ( size_t )w * ch );
    }
    uLongf zl = compressBound( ( uLong )raw.size() );
    std::vector< unsigned char > z( zl );
    if( compress2( z.data(), &zl, raw.data(), ( uLong )raw.size(), 6 ) != Z_OK ) return false;
    z.resize( zl );
    std::vector< unsigned char > f = { 137, 80, 78, 71, 13, 10, 26, 10 }, ihdr;
    put_be32( ihdr, ( uint32_t )w );
    put_be32( ihdr, ( uint32_t )h );
    ihdr.push_back( 8 );
    ihdr.push_back( ch == 1 ? 0 : ( ch == 3 ? 2 : 6 ) );
    ihdr.push_back( 0 ); ihdr.push_back( 0 ); ihdr.push_back( 0 );
    png_chunk( f, "IHDR", ihdr );
    png_chunk( f, "IDAT", z );
    png_chunk( f, "IEND", {} );
    FILE* o = fopen( path, "wb" );
    if( !o ) return false;
    bool ok = fwrite( f.data(), 1, f.size(), o ) == f.size();
    fclose( o );
    return ok;
}

bool decode_pnm( const std::vector< unsigned char >& file, Decoded& out, std::string& err )
{
    if( file.size() < 7 || file[ 0 ] != 'P' || ( file[ 1 ] != '6' && file[ 1 ] != '5' ) ) { err = "not a binary PPM/PGM"; return false; }
    size_t pos = 2;
    long long vals[ 3 ] = { 0, 0, 0 };
    int got = 0;
    while( got < 3 && pos < file.size() )
    {
        if( file[ pos ] == '#' ) { while( pos < file.size() && file[ pos ] != '\n' ) pos++; continue; }
        if( isspace( file[ pos ] ) ) { pos++; continue; }
        if( !isdigit( file[ pos ] ) ) { err = "bad PNM header"; return false; }
        long long v = 0;
        while( pos < file.size() && isdigit( file[ pos ] ) )
        {
            v = v * 10 + ( file[ pos++ ] - '0' );
            if( v > 1000000 ) { err = "bad PNM header"; return false; }
        }
        vals[ got++ ] = v;
    }
    if( got < 3 || vals[ 2 ] != 255 || !sane_size( vals[ 0 ], vals[ 1 ] ) ) { err = "unsupported or truncated PNM"; return false; }
    pos++; // single whitespace after maxval
    out.w = ( int )vals[ 0 ];
    out.h = ( int )vals[ 1 ];
    out.ch = file[ 1 ] == '6' ? 3 : 1;
    if( pos > file.size() || ( size_t )out.w * out.h * out.ch > file.size() - pos ) { err = "unsupported or truncated PNM"; return false; }
    out.px.assign( file.begin() + pos, file.begin() + pos + ( size_t )out.w * out.h * out.ch );
    return true;
}

bool decode_bmp( const std::vector< unsigned char >& f, Decoded& out, std::string& err )
{
    if( f.size() < 54 || f[ 0 ] != 'B' || f[ 1 ] != 'M' ) { err = "not a BMP"; return false; }
    auto le32 = [ & ]( size_t o ) { return ( int32_t )( f[ o ] | f[ o + 1 ] << 8 | f[ o + 2 ] << 16 | ( uint32_t )f[ o + 3 ] << 24 ); };
    const long long off = le32( 10 ), w = le32( 18 ), h_signed = le32( 22 );
    const int bits = f[ 28 ] | f[ 29 ] << 8, comp = le32( 30 );
    const bool top_down = h_signed < 0;
    const long long h = top_down ? -h_signed : h_signed; // (64-bit: INT_MIN has no int negation)
    if( ( bits != 24 && bits != 32 ) || comp != 0 || !sane_size( w, h ) ) { err = "unsupported BMP (need uncompressed 24/32 bit of a sane size)"; return false; }
    const size_t stride = ( ( size_t )w * bits / 8 + 3 ) & ~( size_t )3;
    if( off < 54 || ( size_t )off > f.size() || stride * ( size_t )h > f.size() - ( size_t )off ) { err = "truncated BMP"; return false; }
    out.w = ( int )w; out.h = ( int )h; out.ch = 3;
    out.px.resize( ( size_t )w * h * 3 );
    for( int y = 0; y < h; y++ )
    {
        const unsigned char* src = &f[ ( size_t )off + stride * ( size_t )( top_down ? y : h - 1 - y ) ];
        for( int x = 0; x < w; x++ )
            for( int k = 0; k < 3; k++ ) out.px[ ( ( size_t )y * w + x ) * 3 + k ] = src[ x * ( bits / 8 ) + 2 - k ]; // BGR -> RGB
    }
    return true;
}

} // namespace

Image::Image() : width( 0 ), height( 0 ), widthStep( 0 ), n_channels( 0 ) {}
Image::~Image() {}

void Image::reverses()
{
    std::vector< unsigned char > row( widthStep );
    for( int i = 0; i < height / 2; i++ )
    {
        unsigned char *a = &pixels[ ( size_t )i * widthStep ], *b = &pixels[ ( size_t )( height - 1 - i ) * widthStep ];
        memcpy( row.data(), a, widthStep );
        memcpy( a, b, widthStep );
        memcpy( b, row.data(), widthStep );
    }
}

void Image::loadImage( const char* path, int colorness )
{
    pixels.clear();
    width = height = widthStep = n_channels = 0;
    std::vector< unsigned char > file;
    if( !read_file( path, file ) ) { last_error = std::string( "cannot read " ) + path; return; }
    Decoded d;
    bool ok = false;
    if( file.size() > 8 && file[ 0 ] == 137 ) ok = decode_png( file, d, last_error );
    else if( file.size() > 2 && file[ 0 ] == 'P' ) ok = decode_pnm( file, d, last_error );
    else if( file.size() > 2 && file[ 0 ] == 'B' ) ok = decode_bmp( file, d, last_error );
    else last_error = "unknown image format (PNG, PPM/PGM, BMP are supported)";
    if( !ok ) return;
    // cvLoadImage( path, CV_LOAD_IMAGE_COLOR ): always 3 channels, B,G,R order, alpha dropped
    const bool color = colorness != CV_LOAD_IMAGE_GRAYSCALE;
    createImage( d.w, d.h, IPL_DEPTH_8U, color ? 3 : 1 );
    for( int y = 0; y < d.h; y++ )
        for( int x = 0; x < d.w; x++ )
        {
            const unsigned char* s = &d.px[ ( ( size_t )y * d.w + x ) * d.ch ];
            const int r = s[ 0 ], g = d.ch >= 3 ? s[ 1 ] : s[ 0 ], b = d.ch >= 3 ? s[ 2 ] : s[ 0 ];
            unsigned char* o = &pixels[ ( size_t )y * widthStep + ( size_t )x * n_channels ];
            if( color ) { o[ 0 ] = ( unsigned char )b; o[ 1 ] = ( unsigned char )g; o[ 2 ] = ( unsigned char )r; }
            else o[ 0 ] = ( unsigned char )( ( 299 * r + 587 * g + 114 * b + 500 ) / 1000 );
        }
}

void Image::createImage( int w, int h, int /*depth*/, int nch )
{
    width = w;
    height = h;
    n_channels = nch;
    widthStep = align4( w * nch );
    pixels.assign( ( size_t )widthStep * h, 0 );
}

void Image::saveImage( const char* file_name )
{
    if( pixels.empty() ) { last_error = "saveImage: empty image"; return; }
    const std::string name( file_name );
    std::vector< unsigned char > rgb( ( size_t )width * height * n_channels );
    for( int y = 0; y < height; y++ )
        for( int x = 0; x < width; x++ )
        {
            const unsigned char* s = &pixels[ ( size_t )y * widthStep + ( size_t )x * n_channels ];
            unsigned char* o = &rgb[ ( ( size_t )y * width + x ) * n_channels ];
            if( n_channels >= 3 ) { o[ 0 ] = s[ 2 ]; o[ 1 ] = s[ 1 ]; o[ 2 ] = s[ 0 ]; if( n_channels == 4 ) o[ 3 ] = s[ 3 ]; }
            else o[ 0 ] = s[ 0 ];
        }
    bool ok;
    if( ends_with( name, ".ppm" ) || ends_with( name, ".pgm" ) )
    {
        FILE* f = fopen( file_name, "wb" );
        ok = f != nullptr;
        if( ok )
        {
            fprintf( f, "P%c\n%d %d\n255\n", n_channels == 1 ? '5' : '6', width, height );
            for( int y = 0; y < height && ok; y++ )
                for( int x = 0; x < width; x++ ) ok = fwrite( &rgb[ ( ( size_t )y * width + x ) * n_channels ], 1, n_channels == 1 ? 1 : 3, f ) > 0 && ok;
            fclose( f );
        }
    }
    else
        ok = encode_png( file_name, rgb.data(), width, height, n_channels, ( size_t )width * n_channels );
    if( !ok ) last_error = std::string( "cannot write " ) + file_name;
}

void Image::copy( Image* dst )
{
    if( !dst ) return;
    if( dst->width != width || dst->height != height || dst->n_channels != n_channels ) dst->createImage( width, height, IPL_DEPTH_8U, n_channels );
    dst->pixels = pixels;
}

void Image::resizeImage( Image* dst )
{
    // bicubic (a = -0.75, the kernel cvResize's CV_INTER_CUBIC uses), edge-clamped
    if( !dst || dst->pixels.empty() || pixels.empty() ) return;
    auto wgt = []( float t ) {
        const float a = -0.75f;
        t = std::fabs( t );
        return t <= 1 ? ( ( a + 2 ) * t - ( a + 3 ) ) * t * t + 1 : ( t < 2 ? ( ( a * t - 5 * a ) * t + 8 * a ) * t - 4 * a : 0.0f );
    };
    const float sx = ( float )width / dst->width, sy = ( float )height / dst->height;
    for( int y = 0; y < dst->height; y++ )
        for( int x = 0; x < dst->width; x++ )
        {
            const float fx = ( x + 0.5f ) * sx - 0.5f, fy = ( y + 0.5f ) * sy - 0.5f;
            const int ix = ( int )std::floor( fx ), iy = ( int )std::floor( fy );
            for( int c = 0; c < std::min( n_channels, dst->n_channels ); c++ )
            {
                float acc = 0;
                for( int m = -1; m <= 2; m++ )
                    for( int n = -1; n <= 2; n++ )
                    {
                        const int px = std::min( std::max( ix + n, 0 ), width - 1 ), py = std::min( std::max( iy + m, 0 ), height - 1 );
                        acc += wgt( fx - ( ix + n ) ) * wgt( fy - ( iy + m ) ) * pixels[ ( size_t )py * widthStep + ( size_t )px * n_channels + c ];
                    }
                dst->pixels[ ( size_t )y * dst->widthStep + ( size_t )x * dst->n_channels + c ] = ( unsigned char )std::min( 255.0f, std::max( 0.0f, acc + 0.5f ) );
            }
        }
}

void Image::setAllPixels( int c1, int c2, int c3 )
{
    // cvSet( img, CV_RGB( c1, c2, c3 ) ): CV_RGB(r,g,b) = cvScalar(b,g,r)
    for( int y = 0; y < height; y++ )
        for( int x = 0; x < width; x++ )
        {
            unsigned char* o = &pixels[ ( size_t )y * widthStep + ( size_t )x * n_channels ];
            if( n_channels >= 3 ) { o[ 0 ] = ( unsigned char )c3; o[ 1 ] = ( unsigned char )c2; o[ 2 ] = ( unsigned char )c1; }
            else o[ 0 ] = ( unsigned char )c3;
        }
}

char* Image::accessPixel( int x, int y )
{
    char* pixel = new char[ 3 ];
    for( int k = 0; k < 3; k++ ) pixel[ k ] = ( char )pixels[ ( size_t )y * widthStep + ( size_t )x * n_channels + std::min( k, n_channels - 1 ) ];
    return pixel;
}

void Image::setImageData( char* data )
{
    if( data && !pixels.empty() ) memcpy( pixels.data(), data, pixels.size() );
}
