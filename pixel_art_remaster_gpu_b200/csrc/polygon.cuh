// Stage D + E on the device: the (optionally corner-cut) polygon of one pixel's cell, produced as a
// stream of vertices in EXACT integer arithmetic (units of 1/64 pixel, cell-local coordinates).
//
// Replaces cells_Kernel (kernel.cu:192-213; via the tables of cell_table.h) and subdivision_Kernel
// (kernel.cu:216-261): subdivision (subdivision_functions.cu:564-669), getQ_i/getR_i (:42-122),
// get{R,Q}_i_from_linked_cell (:125-154), getOppositePoint/Coord (:427-524), getPointIndex
// (:527-538), midPoint (:554), checkTJunction (:170-242).  Every coordinate the reference computes
// here is a dyadic rational with denominator <= 64 (hull vertices are quarter-pixels; the cut
// points are 3/4-1/4 or 7/8-1/8 blends and one halving), so integers reproduce its floats exactly.
#pragma once
#include "cell_table.h"
#include "common.cuh"

// The polygon logic is also compiled for the host (tools/extent_probe.cpp): plain loads there.
#if defined( __CUDA_ARCH__ )
#define PAR_LDG( p ) __ldg( p )
#else
#define PAR_LDG( p ) ( *( p ) )
#endif
#if defined( __CUDACC__ )
#define PAR_HD __host__ __device__ __forceinline__
#else
#define PAR_HD inline
#endif

namespace par {

struct CellTablePtrs
{
    const uint64_t* verts;
    const uint64_t* info;
    const uint64_t* index;
};

// vertex t (0..7) of a packed hull in quarter-pixel units
struct Q2 { int x, y; };
PAR_HD Q2 hull_vertex( uint64_t verts, int t )
{
#if defined( __CUDA_ARCH__ )
    const uint32_t b = __byte_perm( ( uint32_t )verts, ( uint32_t )( verts >> 32 ), ( uint32_t )t ) & 0xFFu; // PRMT: byte t of 8
#else
    const uint32_t b = ( uint32_t )( verts >> ( 8 * t ) ) & 0xFFu;
#endif
    Q2 q;
    q.x = ( int )( b & 15u ) - 1;
    q.y = ( int )( b >> 4 ) - 1;
    return q;
}

// points 1/4 (1/8 when the edge is longer than one pixel) from either end of edge a->b, in 1/64 px
// (getQ_i / getR_i, subdivision_functions.cu:42-122; "lenght <= 1.0" <=> dx^2+dy^2 <= 16 quarter^2):
// Q = a + w*(b-a), R = b - w*(b-a) with w = 1/4 or 1/8; in 1/64 px a quarter unit is 16.
PAR_HD void cut_points( Q2 a, Q2 b, int& qx, int& qy, int& rx, int& ry )
{
    const int dx = b.x - a.x, dy = b.y - a.y;
    const int w = dx * dx + dy * dy > 16 ? 2 : 4; // 16 * (1/8) or 16 * (1/4)
    const int ox = dx * w, oy = dy * w;
    qx = 16 * a.x + ox;
    qy = 16 * a.y + oy;
    rx = 16 * b.x - ox;
    ry = 16 * b.y - oy;
}

// Env must provide:
//   uint32_t key( int i, int j )            cell key of an in-image pixel
//   bool keep_corner( int i, int j, Q2 p )  checkTJunction (subdivision_functions.cu:170-242)
// Sink must provide:  void vertex( int x64, int y64 )
// Every hull vertex yields one or two polygon vertices.  The case analysis only computes coordinates;
// the sink is fed from ONE place so that the threads of a warp stay converged on its code.
template< class Env, class Sink >
PAR_HD void emit_cell_polygon( const Env& env, const CellTablePtrs& tab, int i, int j, uint32_t key, bool subdivide, Sink& sink )
{
    const uint64_t h = PAR_LDG( tab.verts + key );
    const uint64_t info = PAR_LDG( tab.info + key );
    const int n = hull_count( info );
    const bool plain = !subdivide || ( key & 0xFFu ) == 90u; // interior nodes are not smoothed (kernel.cu:231)
    const uint32_t links = ( uint32_t )info;
    int prev_link = ( int )( ( links >> ( 4 * ( n - 1 ) ) ) & 15u );
    Q2 p_prev = hull_vertex( h, n - 1 );
    Q2 p_cur = hull_vertex( h, 0 );
    for( int t = 0; t < n; t++ )
    {
        const int cur_link = ( int )( ( links >> ( 4 * t ) ) & 15u );
        const Q2 p_next = hull_vertex( h, t + 1 == n ? 0 : t + 1 );
        const bool cur_border = cur_link == 15, prev_border = prev_link == 15;
        int ax = 16 * p_cur.x, ay = 16 * p_cur.y, bx = 0, by = 0;
        bool two = false;
        if( !plain && ( cur_border || prev_border ) ) // two shared edges: the vertex stays (:651-655)
        {
            int qx, qy, rx, ry, ux, uy;
            cut_points( p_cur, p_next, qx, qy, ux, uy ); // Q of the current edge
            cut_points( p_prev, p_cur, ux, uy, rx, ry ); // R of the previous edge
            if( cur_border && prev_border )
            {
                if( !env.keep_corner( i, j, p_cur ) ) // else the corner stays (:583-588)
                {
                    ax = rx; ay = ry; // :590-598
                    bx = qx; by = qy;
                    two = true;
                }
            }
            else
            {
                // exactly one of the two edges is shared with a neighbour cell: blend with that cell's cut
                // point so that both cells meet on the same curve (:603-647)
                const int L = cur_border ? prev_link : cur_link;
                const int di = edge_di( L ), dj = edge_dj( L );
                const uint32_t nkey = env.key( i + di, j + dj );
                const uint64_t hn = PAR_LDG( tab.verts + nkey );
                const int nn = hull_count( PAR_LDG( tab.info + nkey ) );
                // this vertex in the neighbour's frame -> which of the neighbour's vertices it is
                // (first match, 0 when absent: getPointIndex :527-538, here one table lookup)
                const int code = point_code( p_cur.x - 4 * di, p_cur.y - 4 * dj );
                const int op = code < 0 ? 0 : ( int )( ( PAR_LDG( tab.index + nkey ) >> ( 4 * code ) ) & 15u );
                int aqx, aqy, arx, ary;
                two = true;
                if( cur_border )
                {
                    // neighbour's R on the edge that ENDS at the shared vertex (:125-138)
                    cut_points( hull_vertex( hn, op == 0 ? nn - 1 : op - 1 ), hull_vertex( hn, op ), aqx, aqy, arx, ary );
                    ax = ( qx + arx + 64 * di ) >> 1;
                    ay = ( qy + ary + 64 * dj ) >> 1;
                    bx = qx; by = qy;
                }
                else
                {
                    // neighbour's Q on the edge that STARTS at the shared vertex (:141-154)
                    cut_points( hull_vertex( hn, op ), hull_vertex( hn, op + 1 == nn ? 0 : op + 1 ), aqx, aqy, arx, ary );
                    ax = rx; ay = ry;
                    bx = ( rx + aqx + 64 * di ) >> 1;
                    by = ( ry + aqy + 64 * dj ) >> 1;
                }
            }
        }
#if defined( __CUDA_ARCH__ )
#pragma unroll 1
#endif
        for( int e = 0; e < ( two ? 2 : 1 ); e++ ) sink.vertex( e ? bx : ax, e ? by : ay );
        prev_link = cur_link;
        p_prev = p_cur;
        p_cur = p_next;
    }
}

// checkTJunction (subdivision_functions.cu:170-242) on the frame as the flat byte array the
// reference indexes; bytes at or beyond height*widthstep read as zero (SURVEY App. B-3 contract).
struct FlatImage
{
    const uint8_t* frame;
    int width, height, widthstep;
    PAR_HD uint32_t colour( long idx ) const
    {
        const long end = ( long )height * widthstep;
        uint32_t b0 = idx < end ? PAR_LDG( frame + idx ) : 0u;
        uint32_t b1 = idx + 1 < end ? PAR_LDG( frame + idx + 1 ) : 0u;
        uint32_t b2 = idx + 2 < end ? PAR_LDG( frame + idx + 2 ) : 0u;
        return b0 | b1 << 8 | b2 << 16;
    }
    PAR_HD bool guard( int i, int j ) const
    {
        long idx = ( long )j * widthstep + 3L * i;
        return idx - widthstep - 1 < 0 || idx + width + 1 > ( long )height * widthstep - 1; // :187, "width" as written
    }
    PAR_HD bool keep_corner( int i, int j, Q2 p ) const
    {
        if( guard( i, j ) ) return true;
        const bool x0 = p.x == 0, x1 = p.x == 4, y0 = p.y == 0, y1 = p.y == 4;
        if( !( ( x0 || x1 ) && ( y0 || y1 ) ) ) return false;
        const long idx = ( long )j * widthstep + 3L * i;
        const long ws = widthstep;
        // the three OTHER pixels around that corner, as flat byte offsets (:195-202, wraps at row ends)
        long a, b, c;
        if( x0 && y0 ) { a = idx - 3; b = idx - ws - 3; c = idx - ws; }           // c3, c5, c6
        else if( x1 && y0 ) { a = idx + 3; b = idx - ws + 3; c = idx - ws; }      // c4, c7, c6
        else if( x1 && y1 ) { a = idx + ws; b = idx + ws + 3; c = idx + 3; }      // c1, c2, c4
        else { a = idx + ws - 3; b = idx + ws; c = idx - 3; }                     // c0, c1, c3
        const uint32_t cb = colour( b );
        return colour( a ) != cb || cb != colour( c );
    }
};

} // namespace par
