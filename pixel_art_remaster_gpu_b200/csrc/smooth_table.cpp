// Host-side construction of the smoothing tables (see smooth_table.h): per key the link descriptors, per
// (key, link direction) the neighbour record, and the list of link classes the device turns into masks.
#include "smooth_table.h"
#include <map>
#include <string.h>
#include <tuple>

namespace par {

namespace {

struct Hull
{
    int n;
    int x[ 8 ], y[ 8 ], link[ 8 ];
    bool border[ 8 ];
};

Hull hull_of( const CellRecord& r )
{
    Hull h;
    h.n = hull_count( r.info );
    for( int t = 0; t < h.n; t++ )
    {
        h.x[ t ] = hull_xq( r.verts, t );
        h.y[ t ] = hull_yq( r.verts, t );
        h.link[ t ] = ( int )( ( r.info >> ( 4 * t ) ) & 15u );
        h.border[ t ] = h.link[ t ] == 15;
    }
    return h;
}

} // namespace

void build_smooth_tables( const CellTables& cells, SmoothTables* out )
{
    memset( out->rec, 0, sizeof( out->rec ) );
    out->classes.clear();
    out->link_entries = 0;
    out->slow_keys = 0;
    typedef std::tuple< int, int, int, int, int, int, int, int, int, int, int > ClassKey; // e, hasA, hasB, 4 x (x, y)
    std::map< ClassKey, uint32_t > class_first;
    for( unsigned key = 0; key < ( unsigned )kCellKeys; key++ )
    {
        const CellRecord& r = cells.rec[ key ];
        const Hull h = hull_of( r );
        // neighbour records: for every link direction the hull edge shared through it, and the vertices around it
        for( int e = 0; e < 8; e++ ) out->nbr[ key ][ e ] = 0; // start == end: no such edge
        for( int t = 0; t < h.n; t++ )
        {
            if( h.border[ t ] ) continue;
            const int tn = ( t + 1 ) % h.n, tnn = ( t + 2 ) % h.n, tp = ( t + h.n - 1 ) % h.n;
            const int start = point_code( h.x[ t ], h.y[ t ] ), end = point_code( h.x[ tn ], h.y[ tn ] );
            const int after = point_code( h.x[ tnn ], h.y[ tnn ] ), before = point_code( h.x[ tp ], h.y[ tp ] );
            out->nbr[ key ][ h.link[ t ] ] = ( uint16_t )( after | before << 4 | end << 8 | start << 12 );
        }
        // link descriptors
        int n_links = 0;
        bool slow = false;
        uint32_t links[ 8 ];
        for( int t = 0; t < h.n && !slow; t++ )
        {
            if( h.border[ t ] ) continue;
            const int tp = ( t + h.n - 1 ) % h.n, tn = ( t + 1 ) % h.n, tnn = ( t + 2 ) % h.n;
            const bool hasA = h.border[ tp ], hasB = h.border[ tn ];
            if( !hasA && !hasB ) continue;
            const unsigned codeA = ( unsigned )( r.aux >> ( 8 * t ) ) & 255u, codeB = ( unsigned )( r.aux >> ( 8 * tn ) ) & 255u;
            if( ( hasA && codeA == 255u ) || ( hasB && codeB == 255u ) )
            {
                slow = true; // the vertex is not one of the 16 hull points in the neighbour's frame: never found there
                break;
            }
            const ClassKey ck( h.link[ t ], hasA, hasB, hasA ? h.x[ tp ] : 0, hasA ? h.y[ tp ] : 0, h.x[ t ], h.y[ t ], h.x[ tn ], h.y[ tn ],
                               hasB ? h.x[ tnn ] : 0, hasB ? h.y[ tnn ] : 0 );
            auto it = class_first.find( ck );
            if( it == class_first.end() )
            {
                LinkClass c;
                memset( &c, 0, sizeof( c ) );
                c.e = ( int8_t )h.link[ t ];
                c.hasA = hasA;
                c.hasB = hasB;
                const int idx[ 4 ] = { tp, t, tn, tnn };
                for( int k = 0; k < 4; k++ )
                {
                    c.px[ k ] = ( int8_t )h.x[ idx[ k ] ];
                    c.py[ k ] = ( int8_t )h.y[ idx[ k ] ];
                }
                c.first = out->link_entries;
                c.count = ( hasA && hasB ) ? 256u : 16u;
                out->link_entries += c.count;
                out->classes.push_back( c );
                it = class_first.insert( std::make_pair( ck, c.first ) ).first;
            }
            links[ n_links++ ] = ( uint32_t )h.link[ t ] | ( hasA ? 8u : 0u ) | ( hasB ? 16u : 0u ) | ( hasA ? codeA << 5 : 0u ) |
                                 ( hasB ? codeB << 9 : 0u ) | it->second << 13;
        }
        if( slow || n_links > kMaxLinks )
        {
            out->rec[ key ].link[ 0 ] = kSmoothSlow;
            out->slow_keys++;
            continue;
        }
        for( int k = 0; k < n_links; k++ ) out->rec[ key ].link[ k ] = links[ k ];
    }
}

} // namespace par
