// Image — the reference's image wrapper (Image.h:22-152, Image.cpp:21-223), re-implemented without
// OpenCV: same class name, same methods, same data layout.  The reference wraps an OpenCV-1.x IplImage;
// the hot path only ever sees what this class publishes — `getImageData()` = interleaved 8-bit BGR rows,
// `getWidthStep()` bytes apart (IplImage aligns rows to 4 bytes), row 0 = TOP scanline until
// `reverses()` flips it (main.cpp:59) — so that layout is what is kept.
//
// File formats: PNG (8-bit grey / RGB / RGBA / palette, non-interlaced; zlib is the only dependency),
// binary PPM/PGM (P6/P5) and 24/32-bit BMP.  Decoding parity with OpenCV is out of the path's scope.
#pragma once
#include <string>
#include <vector>

#define CHECK_BIT( var, pos ) ( ( var ) & ( 1 << ( pos ) ) ) /* Image.h:15 */

#ifndef CV_LOAD_IMAGE_COLOR
#define CV_LOAD_IMAGE_COLOR 1
#define CV_LOAD_IMAGE_GRAYSCALE 0
#define IPL_DEPTH_8U 8
#endif

class Image
{
public:
    Image();
    virtual ~Image();

    void reverses();                                                      // Image.cpp:21-38  vertical flip in place
    void loadImage( const char* path, int colorness );                    // Image.cpp:42-49
    void createImage( int width, int height, int depth, int n_channels ); // Image.cpp:53-60
    void saveImage( const char* file_name );                              // Image.cpp:64-71  (PNG unless the suffix says .ppm/.pgm/.bmp)
    void copy( Image* dst );                                              // Image.cpp:75-78
    void resizeImage( Image* dst );                                       // Image.cpp:82-86  bicubic to dst's size
    void setAllPixels( int c1, int c2, int c3 );                          // Image.cpp:90-93  CV_RGB(c1,c2,c3): stored as (c3,c2,c1)
    char* accessPixel( int x, int y );                                    // Image.cpp:97-103 new char[3], caller deletes

    void setWidth( int value ) { width = value; }
    int getWidth() { return width; }
    void setHeight( int value ) { height = value; }
    int getHeight() { return height; }
    void setWidthStep( int value ) { widthStep = value; }
    int getWidthStep() { return widthStep; }
    void setImageData( char* data );  // adopts nothing: copies height*widthStep bytes
    char* getImageData() { return pixels.empty() ? nullptr : reinterpret_cast< char* >( pixels.data() ); }
    int getNchannels() { return n_channels; }
    bool ok() const { return !pixels.empty(); }
    const std::string& error() const { return last_error; }

private:
    std::vector< unsigned char > pixels; // height * widthStep bytes (what IplImage::imageData points at)
    int width, height, widthStep, n_channels;
    std::string last_error;
};
