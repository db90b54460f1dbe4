// Stage D as a table: the cell of a pixel is a pure function of 12 bits (its own graph byte plus the
// four "corner-cutting" diagonals of its left/right neighbours), so the 4096 hulls, the link
// classification of their edges and a point->vertex index map are computed once on the host and kept
// in HBM/L2 (128 KB as one 32-byte record per key, so that a lookup touches ONE sector; the keys a frame
// actually uses sit in L1).
//
// Replaces createCellFromPattern / convex_hull / sort (diagram_functions.cu:319-535, :238-316,
// :82-129), isLinkedEdge (subdivision_functions.cu:245-424) and the linear search of getPointIndex
// (:527-538).
#pragma once
#include <stdint.h>

#if defined( __CUDACC__ )
#define PAR_TAB_HD __host__ __device__ __forceinline__
#else
#define PAR_TAB_HD inline
#endif

namespace par {

constexpr int kCellKeys = 4096;

// key = node | left.bit2 << 8 | left.bit7 << 9 | right.bit0 << 10 | right.bit5 << 11
PAR_TAB_HD unsigned cell_key( unsigned node, unsigned left, unsigned right )
{
    return ( node & 0xFFu ) | ( ( left >> 2 ) & 1u ) << 8 | ( ( left >> 7 ) & 1u ) << 9 | ( right & 1u ) << 10 | ( ( right >> 5 ) & 1u ) << 11;
}

// Per key, one 32-byte record of four 64-bit words:
//   verts : byte t = vertex t of the hull, (4x+1) | (4y+1) << 4, quarter-pixel units, x,y in [-1/4, 5/4];
//           counter-clockwise from the lexicographically smallest vertex, no closing duplicate
//   info  : bits [0,32)  4 bits per edge t (vertex t -> t+1 mod n): the graph edge 0..7 the polygon edge
//                        is shared through, or 15 for a border edge
//           bits [32,36) vertex count n (4..8)
//           bits [36,44) bit t set when edge t is a border edge
//           bits [44,60) 4 bits per square corner (0,0) (1,0) (1,1) (0,1): the hull vertex sitting there, 15 = none
//   index : 4 bits per point code (see point_code): the index of the hull vertex at that point, 0 when
//           the hull has no vertex there (what getPointIndex returns for "not found")
//   aux   : one byte per vertex t: when t is a "blend" vertex (exactly one adjacent edge shared), the point
//           code of vertex t seen from the neighbour cell across that shared edge; 0xFF = not a hull point
struct CellRecord
{
    uint64_t verts, info, index, aux;
};
struct CellTables
{
    CellRecord rec[ kCellKeys ];
};

void build_cell_tables( CellTables* t );

// The 16 points a hull vertex can sit on (square corners, cut corners, the eight diagonal tips), as a
// code 0..15 = rank of the point in row-major order of the 7x7 quarter-pixel grid [-1,5]^2; -1 for any
// other point.  (x,y) in quarter pixels.
constexpr uint64_t kValidPoints = 0x511550155114ull; // bit (y+1)*7 + (x+1)
#if defined( __CUDA_ARCH__ )
#define PAR_POPCLL( v ) __popcll( v )
#else
#define PAR_POPCLL( v ) __builtin_popcountll( v )
#endif
PAR_TAB_HD int point_code( int x, int y )
{
    if( ( unsigned )( x + 1 ) > 6u || ( unsigned )( y + 1 ) > 6u ) return -1;
    const int pos = ( y + 1 ) * 7 + ( x + 1 );
    if( !( ( kValidPoints >> pos ) & 1u ) ) return -1;
    return ( int )PAR_POPCLL( kValidPoints & ( ( 1ull << pos ) - 1ull ) );
}

PAR_TAB_HD int hull_count( uint64_t info ) { return ( int )( ( info >> 32 ) & 15u ); }
PAR_TAB_HD uint32_t hull_border_mask( uint64_t info ) { return ( uint32_t )( info >> 36 ) & 255u; }
PAR_TAB_HD uint32_t hull_corner_vertices( uint64_t info ) { return ( uint32_t )( info >> 44 ) & 0xFFFFu; }
PAR_TAB_HD int hull_xq( uint64_t verts, int t ) { return ( int )( ( verts >> ( 8 * t ) ) & 15u ) - 1; }
PAR_TAB_HD int hull_yq( uint64_t verts, int t ) { return ( int )( ( verts >> ( 8 * t + 4 ) ) & 15u ) - 1; }

} // namespace par
