// Host-side construction of the 4096-entry cell tables (see cell_table.h).
//
// The candidate points of a pattern follow createCellFromPattern (diagram_functions.cu:319-535):
// walking the eight neighbours clockwise from the up-left one, a linked diagonal pushes the cell
// corner out by a quarter pixel (one or two points, depending on the adjacent orthogonal links), an
// unlinked diagonal keeps the square corner unless the neighbouring pixels' diagonal cuts it; an
// orthogonal link contributes the two square corners of that side.  (The interior "dummy" points
// the reference adds for unlinked sides never reach the hull and are skipped.)  The hull is the
// convex hull of those points with collinear points dropped, counter-clockwise from the
// lexicographically smallest vertex — which is what the reference's sort + monotone chain
// (diagram_functions.cu:82-129, :238-316) produces for every one of the 4096 patterns
// (tests/test_cell_table.py compares all of them with the reference build and the CPU checker).
#include "cell_table.h"
#include <algorithm>
#include <utility>
#include <vector>

namespace par {

namespace {

typedef std::pair< int, int > Q; // quarter-pixel units

long turn( const Q& o, const Q& a, const Q& b ) { return ( long )( a.first - o.first ) * ( b.second - o.second ) - ( long )( a.second - o.second ) * ( b.first - o.first ); }

std::vector< Q > candidates( unsigned key )
{
    const bool n0 = key & 1, n1 = key & 2, n2 = key & 4, n3 = key & 8, n4 = key & 16, n5 = key & 32, n6 = key & 64, n7 = key & 128;
    const bool cut_ul = key & 256, cut_dl = key & 512, cut_ur = key & 1024, cut_dr = key & 2048;
    std::vector< Q > p;
    // up-left corner (0,1)
    if( n0 )
    {
        if( !( n3 && !n1 ) ) p.push_back( Q( -1, 3 ) );
        if( !( n1 && !n3 ) ) p.push_back( Q( 1, 5 ) );
    }
    else
        p.push_back( cut_ul ? Q( 1, 3 ) : Q( 0, 4 ) );
    // up-right corner (1,1)
    if( n2 )
    {
        if( !( n1 && !n4 ) ) p.push_back( Q( 3, 5 ) );
        if( !( n4 && !n1 ) ) p.push_back( Q( 5, 3 ) );
    }
    else
        p.push_back( cut_ur ? Q( 3, 3 ) : Q( 4, 4 ) );
    // down-right corner (1,0)
    if( n7 )
    {
        if( !( n4 && !n6 ) ) p.push_back( Q( 5, 1 ) );
        if( !( n6 && !n4 ) ) p.push_back( Q( 3, -1 ) );
    }
    else
        p.push_back( cut_dr ? Q( 3, 1 ) : Q( 4, 0 ) );
    // down-left corner (0,0)
    if( n5 )
    {
        if( !( n3 && !n6 ) ) p.push_back( Q( -1, 1 ) );
        if( !( n6 && !n3 ) ) p.push_back( Q( 1, -1 ) );
    }
    else
        p.push_back( cut_dl ? Q( 1, 1 ) : Q( 0, 0 ) );
    // sides: an orthogonal link keeps both square corners of that side
    if( n1 ) { p.push_back( Q( 0, 4 ) ); p.push_back( Q( 4, 4 ) ); }
    if( n4 ) { p.push_back( Q( 4, 4 ) ); p.push_back( Q( 4, 0 ) ); }
    if( n6 ) { p.push_back( Q( 4, 0 ) ); p.push_back( Q( 0, 0 ) ); }
    if( n3 ) { p.push_back( Q( 0, 0 ) ); p.push_back( Q( 0, 4 ) ); }
    return p;
}

std::vector< Q > convex_hull_ccw( std::vector< Q > p )
{
    std::sort( p.begin(), p.end() );
    p.erase( std::unique( p.begin(), p.end() ), p.end() );
    std::vector< Q > h( 2 * p.size() );
    size_t k = 0;
    for( size_t i = 0; i < p.size(); i++ )
    {
        while( k >= 2 && turn( h[ k - 2 ], h[ k - 1 ], p[ i ] ) <= 0 ) k--;
        h[ k++ ] = p[ i ];
    }
    for( size_t i = p.size() - 1, t = k + 1; i-- > 0; )
    {
        while( k >= t && turn( h[ k - 2 ], h[ k - 1 ], p[ i ] ) <= 0 ) k--;
        h[ k++ ] = p[ i ];
    }
    h.resize( k - 1 );
    return h;
}

// which graph edge a hull edge a->b is shared through (subdivision_functions.cu:245-424, first
// match wins, in the reference's order), 15 = border.  Quarter units: 1/2 pixel = 2.
unsigned classify( const Q& a, const Q& b, unsigned node )
{
    const int dx = b.first - a.first, dy = b.second - a.second;
    if( dy == 0 && a.second > 2 && ( node & 2 ) ) return 1;
    if( dx == 0 && a.first > 2 && ( node & 16 ) ) return 4;
    if( dy == 0 && a.second < 2 && ( node & 64 ) ) return 6;
    if( dx == 0 && a.first < 2 && ( node & 8 ) ) return 3;
    const bool up = dx != 0 && dy == dx, down = dx != 0 && dy == -dx; // slope +1 / -1
    const int twice_mid_y = a.second + b.second;                        // compare with 2 * (1/2 pixel) = 4
    if( up && twice_mid_y > 4 && ( node & 1 ) ) return 0;
    if( down && twice_mid_y > 4 && ( node & 4 ) ) return 2;
    if( up && twice_mid_y < 4 && ( node & 128 ) ) return 7;
    if( down && twice_mid_y < 4 && ( node & 32 ) ) return 5;
    return 15;
}

} // namespace

void build_cell_tables( CellTables* t )
{
    for( unsigned key = 0; key < ( unsigned )kCellKeys; key++ )
    {
        std::vector< Q > h = convex_hull_ccw( candidates( key ) );
        uint64_t verts = 0, info = ( uint64_t )h.size() << 32, index = 0;
        for( size_t v = h.size(); v-- > 0; ) // descending, so that the FIRST vertex at a point wins (they are distinct anyway)
        {
            verts |= ( uint64_t )( ( unsigned )( h[ v ].first + 1 ) | ( unsigned )( h[ v ].second + 1 ) << 4 ) << ( 8 * v );
            const unsigned link = classify( h[ v ], h[ ( v + 1 ) % h.size() ], key & 0xFFu );
            info |= ( uint64_t )link << ( 4 * v );
            if( link == 15 ) info |= 1ull << ( 36 + v );
            const int code = point_code( h[ v ].first, h[ v ].second );
            if( code >= 0 ) index = ( index & ~( 15ull << ( 4 * code ) ) ) | ( uint64_t )v << ( 4 * code );
        }
        // which hull vertex sits on each corner of the pixel square (checkTJunction only looks at those, :205-239)
        static const int corner_x[ 4 ] = { 0, 4, 4, 0 }, corner_y[ 4 ] = { 0, 0, 4, 4 };
        for( int c = 0; c < 4; c++ )
        {
            uint64_t at = 15;
            for( size_t v = 0; v < h.size(); v++ )
                if( h[ v ].first == corner_x[ c ] && h[ v ].second == corner_y[ c ] ) at = v;
            info |= at << ( 44 + 4 * c );
        }
        // blend vertices: the vertex as the neighbour across the shared edge sees it (subdivision_functions.cu:427-474)
        const int n = ( int )h.size();
        uint64_t aux = 0;
        for( int v = 0; v < n; v++ )
        {
            const unsigned cur = ( unsigned )( info >> ( 4 * v ) ) & 15u, prev = ( unsigned )( info >> ( 4 * ( ( v + n - 1 ) % n ) ) ) & 15u;
            unsigned code = 0xFF; // all 16 four-bit codes are real points, so "none" needs a byte
            if( ( cur == 15 ) != ( prev == 15 ) )
            {
                const unsigned L = cur == 15 ? prev : cur; // the shared edge
                static const int di[ 8 ] = { -1, 0, 1, -1, 1, -1, 0, 1 }, dj[ 8 ] = { 1, 1, 1, 0, 0, -1, -1, -1 };
                const int c = point_code( h[ v ].first - 4 * di[ L ], h[ v ].second - 4 * dj[ L ] );
                code = c < 0 ? 0xFFu : ( unsigned )c;
            }
            aux |= ( uint64_t )code << ( 8 * v );
        }
        t->rec[ key ].verts = verts;
        t->rec[ key ].info = info;
        t->rec[ key ].index = index;
        t->rec[ key ].aux = aux;
    }
}

} // namespace par
