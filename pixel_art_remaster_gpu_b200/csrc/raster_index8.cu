// The raster kernel instantiated for PAR_OUT_INDEX8 (1 byte per pixel: index into the frame's palette, palette_kernels.cu).
// Every output pixel of the point-sampled image is a source colour (kernel.cu:98-101) or the background (main.cpp:260), so
// this is the same image in a quarter of the bytes.  Same kernel as raster_kernels.cu: a colour word carries its palette
// index in the top byte from the staging pass on, and the stores of the resolve step write that byte.
#include "raster_impl.cuh"

namespace par {

cudaError_t launch_raster_index8( const RasterArgs& a, const CUtensorMap* graph_map, const CUtensorMap* img_map, cudaStream_t stream )
{
    return launch_raster_fmt< kFmtIndex8 >( a, graph_map, img_map, stream );
}

} // namespace par
