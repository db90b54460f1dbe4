// The raster kernel instantiated for PAR_OUT_BGR8 (3 bytes per pixel, B G R: the image Image::saveImage writes,
// Image.cpp:64-71).  Same kernel as raster_kernels.cu; only the stores of the resolve step differ.
#include "raster_impl.cuh"

namespace par {

cudaError_t launch_raster_bgr8( const RasterArgs& a, const CUtensorMap* graph_map, const CUtensorMap* img_map, cudaStream_t stream )
{
    return launch_raster_fmt< kFmtBgr8 >( a, graph_map, img_map, stream );
}

} // namespace par
