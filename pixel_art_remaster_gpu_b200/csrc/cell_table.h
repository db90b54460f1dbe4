// Stage D as a table: the cell of a pixel is a pure function of 12 bits (its own graph byte plus the
// four "corner-cutting" diagonals of its left/right neighbours), so the 4096 hulls and the link
// classification of their edges are computed once on the host and kept in HBM/L2 (48 KB).
//
// Replaces createCellFromPattern / convex_hull / sort (diagram_functions.cu:319-535, :238-316,
// :82-129) and isLinkedEdge (subdivision_functions.cu:245-424).
#pragma once
#include <stdint.h>

namespace par {

constexpr int kCellKeys = 4096;

// key = node | left.bit2 << 8 | left.bit7 << 9 | right.bit0 << 10 | right.bit5 << 11
#if defined( __CUDACC__ )
__host__ __device__
#endif
inline unsigned cell_key( unsigned node, unsigned left, unsigned right )
{
    return ( node & 0xFFu ) | ( ( left >> 2 ) & 1u ) << 8 | ( ( left >> 7 ) & 1u ) << 9 | ( right & 1u ) << 10 | ( ( right >> 5 ) & 1u ) << 11;
}

// Packed hull: bits [0,4) = vertex count n (4..8); vertex t at bits [4+6t, 10+6t): low 3 bits =
// 4*x + 1, high 3 bits = 4*y + 1 (quarter-pixel units, x,y in [-1/4, 5/4]).  Counter-clockwise,
// starting at the lexicographically smallest vertex, no closing duplicate.
// Packed links: 4 bits per edge t (vertex t -> t+1 mod n): the graph edge 0..7 the polygon edge is
// shared through, or 15 for a border edge.
struct CellTables
{
    uint64_t hull[ kCellKeys ];
    uint32_t link[ kCellKeys ];
};

void build_cell_tables( CellTables* t );

#if defined( __CUDACC__ )
__host__ __device__
#endif
inline int hull_count( uint64_t h ) { return ( int )( h & 15u ); }
#if defined( __CUDACC__ )
__host__ __device__
#endif
inline int hull_xq( uint64_t h, int t ) { return ( int )( ( h >> ( 4 + 6 * t ) ) & 7u ) - 1; }
#if defined( __CUDACC__ )
__host__ __device__
#endif
inline int hull_yq( uint64_t h, int t ) { return ( int )( ( h >> ( 7 + 6 * t ) ) & 7u ) - 1; }

} // namespace par
