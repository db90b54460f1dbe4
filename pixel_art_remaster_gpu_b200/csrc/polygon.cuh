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
    const CellRecord* rec;
};

// (verts, info) of a key with one 16-byte load; index / aux are the other half of the same 32-byte sector
PAR_HD void load_hull( const CellTablePtrs& tab, uint32_t key, uint64_t& verts, uint64_t& info )
{
#if defined( __CUDA_ARCH__ )
    const ulonglong2 v = __ldg( reinterpret_cast< const ulonglong2* >( tab.rec + key ) );
    verts = v.x;
    info = v.y;
#else
    verts = tab.rec[ key ].verts;
    info = tab.rec[ key ].info;
#endif
}
PAR_HD uint64_t load_index( const CellTablePtrs& tab, uint32_t key ) { return PAR_LDG( &tab.rec[ key ].index ); }
PAR_HD uint64_t load_aux( const CellTablePtrs& tab, uint32_t key ) { return PAR_LDG( &tab.rec[ key ].aux ); }

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

// the point 1/4 (1/8 when the edge is longer than one pixel) of the way from p toward x, in 1/64 px:
// getQ_i is cut_toward( P_i, P_i+1 ), getR_i is cut_toward( P_i+1, P_i ) (subdivision_functions.cu:42-122;
// "lenght <= 1.0" <=> dx^2+dy^2 <= 16 in quarter-pixel units; a quarter pixel is 16/64)
PAR_HD void cut_toward( Q2 p, Q2 x, int& ox, int& oy )
{
    const int dx = x.x - p.x, dy = x.y - p.y;
    const int w = dx * dx + dy * dy > 16 ? 2 : 4; // 16 * (1/8) or 16 * (1/4)
    ox = 16 * p.x + dx * w;
    oy = 16 * p.y + dy * w;
}

// Env must provide:
//   uint32_t key( int i, int j )            cell key of an in-image pixel
//   bool guard( int i, int j )              checkTJunction's early "keep everything" test (:187)
//   bool keep_corner( int i, int j, Q2 p )  the rest of checkTJunction (subdivision_functions.cu:170-242)
//
// The hull vertices fall into three classes that are a pure function of the key (border mask of the hull
// edges): both adjacent edges shared -> the vertex stays (:651-655); both border -> corner cut unless it is
// a T-junction (:583-598); exactly one shared -> blend with the neighbour cell's cut point so that both cells
// meet on the same curve (:603-647).
struct VertexClasses
{
    uint32_t cur_border; // bit t: the edge leaving vertex t is a border edge
    uint32_t blend, cut; // bit t: vertex t is of that class
};
PAR_HD VertexClasses classify_vertices( uint64_t info )
{
    const int n = hull_count( info );
    VertexClasses c;
    c.cur_border = hull_border_mask( info );
    const uint32_t prev_border = ( ( c.cur_border << 1 ) | ( c.cur_border >> ( n - 1 ) ) ) & ( ( 1u << n ) - 1u ); // edge arriving at t
    c.blend = c.cur_border ^ prev_border;
    c.cut = c.cur_border & prev_border;
    return c;
}

// "blend" vertex t of cell (i,j): the two vertices that replace it, in emission order
template< class Env >
PAR_HD void blend_vertex( const Env& env, const CellTablePtrs& tab, int i, int j, uint64_t h, uint64_t info, uint64_t aux, int t, int& ax, int& ay,
                          int& bx, int& by )
{
    const int n = hull_count( info );
    const bool cur_border = ( hull_border_mask( info ) >> t ) & 1u; // else the arriving edge is the border one
    const int tp = t == 0 ? n - 1 : t - 1, tn = t + 1 == n ? 0 : t + 1;
    const Q2 p = hull_vertex( h, t );
    // own cut point on the BORDER edge: Q of the current edge, or R of the previous one
    int ownx, owny;
    cut_toward( p, hull_vertex( h, cur_border ? tn : tp ), ownx, owny );
    // the neighbour across the SHARED edge
    const int L = ( int )( ( ( uint32_t )info >> ( 4 * ( cur_border ? tp : t ) ) ) & 15u );
    const int di = edge_di( L ), dj = edge_dj( L );
    const uint32_t nkey = env.key( i + di, j + dj );
    uint64_t hn, info_n;
    load_hull( tab, nkey, hn, info_n );
    const int nn = hull_count( info_n );
    // this vertex in the neighbour's frame -> which of the neighbour's vertices it is
    // (first match, 0 when absent: getPointIndex :527-538, here two table lookups)
    const int code = ( int )( ( aux >> ( 8 * t ) ) & 255u );
    const int op = code == 255 ? 0 : ( int )( ( load_index( tab, nkey ) >> ( 4 * code ) ) & 15u );
    // neighbour's R on the edge that ENDS at that vertex (:125-138) / its Q on the edge that STARTS there (:141-154)
    const int other = cur_border ? ( op == 0 ? nn - 1 : op - 1 ) : ( op + 1 == nn ? 0 : op + 1 );
    int nbx, nby;
    cut_toward( hull_vertex( hn, op ), hull_vertex( hn, other ), nbx, nby );
    const int mx = ( ownx + nbx + 64 * di ) >> 1, my = ( owny + nby + 64 * dj ) >> 1; // midPoint (:554); sums are even
    ax = cur_border ? mx : ownx;
    ay = cur_border ? my : owny;
    bx = cur_border ? ownx : mx;
    by = cur_border ? owny : my;
}

// "cut" vertex t of a cell that passed guard(): false when the corner stays (T-junction), else R then Q
template< class Env >
PAR_HD bool cut_vertex( const Env& env, int i, int j, uint64_t h, int n, int t, int& rx, int& ry, int& qx, int& qy )
{
    const Q2 p = hull_vertex( h, t );
    if( env.keep_corner( i, j, p ) ) return false; // :583-588
    cut_toward( p, hull_vertex( h, t == 0 ? n - 1 : t - 1 ), rx, ry ); // R of the previous edge
    cut_toward( p, hull_vertex( h, t + 1 == n ? 0 : t + 1 ), qx, qy ); // Q of the current edge
    return true;
}

// The polygon of a cell as up to two vertices per hull vertex: slot 2t (always) and slot 2t+1 (when bit t
// of `two` is set), to be read in slot order.
struct CellPoly
{
    int n;        // hull vertices
    uint32_t two; // bit t: hull vertex t was replaced by two vertices
};

PAR_HD int lowest_bit( uint32_t m )
{
#if defined( __CUDA_ARCH__ )
    return __ffs( ( int )m ) - 1;
#else
    return __builtin_ctz( m );
#endif
}

// Slots must provide:  void put( int slot, int x64, int y64 ).  Each class is handled in its own loop so that
// the threads of a warp run the same code together instead of taking turns through a per-vertex switch.
template< class Env, class Slots >
PAR_HD CellPoly build_cell_polygon( const Env& env, const CellTablePtrs& tab, int i, int j, uint32_t key, bool subdivide, Slots& slots )
{
    uint64_t h, info;
    load_hull( tab, key, h, info );
    CellPoly poly;
    poly.n = hull_count( info );
    poly.two = 0u;
    for( int t = 0; t < poly.n; t++ )
    {
        const Q2 p = hull_vertex( h, t );
        slots.put( 2 * t, 16 * p.x, 16 * p.y );
    }
    if( !subdivide || ( key & 0xFFu ) == 90u ) return poly; // interior nodes are not smoothed (kernel.cu:231)
    VertexClasses c = classify_vertices( info );
    const uint64_t aux = c.blend ? load_aux( tab, key ) : 0ull;
    while( c.blend )
    {
        const int t = lowest_bit( c.blend );
        c.blend &= c.blend - 1u;
        int ax, ay, bx, by;
        blend_vertex( env, tab, i, j, h, info, aux, t, ax, ay, bx, by );
        slots.put( 2 * t, ax, ay );
        slots.put( 2 * t + 1, bx, by );
        poly.two |= 1u << t;
    }
    if( c.cut && !env.guard( i, j ) )
        while( c.cut )
        {
            const int t = lowest_bit( c.cut );
            c.cut &= c.cut - 1u;
            int rx, ry, qx, qy;
            if( !cut_vertex( env, i, j, h, poly.n, t, rx, ry, qx, qy ) ) continue;
            slots.put( 2 * t, rx, ry );
            slots.put( 2 * t + 1, qx, qy );
            poly.two |= 1u << t;
        }
    return poly;
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
    // checkTJunction's early exit (:187): with idx = j*widthstep + 3i it returns "keep" when
    //     idx - widthstep - 1 < 0   ||   idx + width + 1 > height*widthstep - 1      ("width", as written).
    // For widthstep >= 3*width the first test holds exactly on row 0 and at pixel (0,1); the second can only
    // hold on the top row, where it reads 3i + width + 2 > widthstep.  Same truth table, 32-bit arithmetic.
    PAR_HD bool guard( int i, int j ) const
    {
        return j == 0 || ( j == 1 && i == 0 ) || ( j == height - 1 && 3 * i + width + 2 > widthstep );
    }
    // the corner tests of checkTJunction, for a cell that passed guard()
    PAR_HD bool keep_corner( int i, int j, Q2 p ) const
    {
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
