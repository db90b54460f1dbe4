// Host-side construction of the smoothing tables (see smooth_table.h): per key the link descriptors, per
// (key, link direction) the neighbour record, and the list of link classes the device turns into masks.
#include "smooth_table.h"
#include <algorithm>
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
    out->link_entries = kNbrIds; // block 0: the all-zero block
    out->slow_keys = 0;
    out->n_canon = 0;
    typedef std::tuple< int, int, int, int, int, int, int, int, int, int, int > ClassKey; // e, hasA, hasB, 4 x (x, y)
    std::map< ClassKey, uint32_t > class_block;
    uint32_t expected[ 8 ] = { 0, 0, 0, 0, 0, 0, 0, 0 }; // per link direction e: point codes some cell expects at the neighbour's edge ends
    static bool has_edge[ kCellKeys ][ 8 ];
    // per direction, the points that occur after the end / before the start of the hull edge shared through it, ranked
    int after_rank[ 8 ][ 16 ], before_rank[ 8 ][ 16 ], after_code[ 8 ][ 4 ], before_code[ 8 ][ 4 ], n_after[ 8 ], n_before[ 8 ];
    bool ranks_ok = true;
    for( int e = 0; e < 8; e++ )
    {
        n_after[ e ] = n_before[ e ] = 0;
        for( int c = 0; c < 16; c++ ) after_rank[ e ][ c ] = before_rank[ e ][ c ] = -1;
        for( int k = 0; k < 4; k++ ) after_code[ e ][ k ] = before_code[ e ][ k ] = 0;
    }
    for( unsigned key = 0; key < ( unsigned )kCellKeys; key++ )
    {
        const Hull h = hull_of( cells.rec[ key ] );
        for( int t = 0; t < h.n; t++ )
        {
            if( h.border[ t ] ) continue;
            const int e = h.link[ t ], tnn = ( t + 2 ) % h.n, tp = ( t + h.n - 1 ) % h.n;
            const int after = point_code( h.x[ tnn ], h.y[ tnn ] ), before = point_code( h.x[ tp ], h.y[ tp ] );
            if( after_rank[ e ][ after ] < 0 )
            {
                if( n_after[ e ] == 4 ) { ranks_ok = false; continue; }
                after_code[ e ][ n_after[ e ] ] = after;
                after_rank[ e ][ after ] = n_after[ e ]++;
            }
            if( before_rank[ e ][ before ] < 0 )
            {
                if( n_before[ e ] == 4 ) { ranks_ok = false; continue; }
                before_code[ e ][ n_before[ e ] ] = before;
                before_rank[ e ][ before ] = n_before[ e ]++;
            }
        }
    }
    for( unsigned key = 0; key < ( unsigned )kCellKeys; key++ )
    {
        const CellRecord& r = cells.rec[ key ];
        const Hull h = hull_of( r );
        // neighbour records: for every link direction the hull edge shared through it, and the vertices around it
        for( int e = 0; e < 8; e++ )
        {
            out->rec[ key ].nbr[ e ] = 0;
            has_edge[ key ][ e ] = false;
        }
        for( int t = 0; t < h.n; t++ )
        {
            if( h.border[ t ] ) continue;
            const int tn = ( t + 1 ) % h.n, tnn = ( t + 2 ) % h.n, tp = ( t + h.n - 1 ) % h.n;
            const int start = point_code( h.x[ t ], h.y[ t ] ), end = point_code( h.x[ tn ], h.y[ tn ] );
            const int after = point_code( h.x[ tnn ], h.y[ tnn ] ), before = point_code( h.x[ tp ], h.y[ tp ] );
            out->rec[ key ].nbr[ h.link[ t ] ] =
                ( uint16_t )( ( after_rank[ h.link[ t ] ][ after ] & 15 ) | ( before_rank[ h.link[ t ] ][ before ] & 15 ) << 4 | end << 8 | start << 12 );
            has_edge[ key ][ h.link[ t ] ] = true;
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
            if( hasA ) expected[ h.link[ t ] ] |= 1u << codeA;
            if( hasB ) expected[ h.link[ t ] ] |= 1u << codeB;
            const ClassKey ck( h.link[ t ], hasA, hasB, hasA ? h.x[ tp ] : 0, hasA ? h.y[ tp ] : 0, h.x[ t ], h.y[ t ], h.x[ tn ], h.y[ tn ],
                               hasB ? h.x[ tnn ] : 0, hasB ? h.y[ tnn ] : 0 );
            auto it = class_block.find( ck );
            if( it == class_block.end() )
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
                for( int k = 0; k < 4; k++ ) // the neighbour's edge is shared through ITS graph edge 7 - e
                {
                    c.after[ k ] = ( int8_t )after_code[ 7 - h.link[ t ] ][ k ];
                    c.before[ k ] = ( int8_t )before_code[ 7 - h.link[ t ] ][ k ];
                }
                c.codeA = ( uint8_t )( hasA ? codeA : 0u );
                c.codeB = ( uint8_t )( hasB ? codeB : 0u );
                c.block = out->link_entries / kNbrIds;
                out->link_entries += kNbrIds;
                out->classes.push_back( c );
                it = class_block.insert( std::make_pair( ck, c.block ) ).first;
            }
            links[ n_links++ ] = ( uint32_t )h.link[ t ] | ( ( hasA ? codeA : 0u ) | ( hasB ? codeB << 4 : 0u ) ) << 8 |
                                 ( ( hasA ? 0x0Fu : 0u ) | ( hasB ? 0xF0u : 0u ) ) << 16 | it->second << 24;
        }
        if( slow || !ranks_ok || n_links > kMaxLinks || out->link_entries / kNbrIds > 255 )
        {
            out->rec[ key ].link[ 0 ] = kSmoothSlow;
            out->slow_keys++;
            continue;
        }
        for( int k = 0; k < n_links; k++ ) out->rec[ key ].link[ k ] = links[ k ];
        // square corners that hold a cut vertex (both adjacent hull edges are border edges)
        uint32_t corners = 0;
        for( int c = 0; c < 4; c++ )
        {
            const int v = ( int )( ( r.info >> ( 44 + 4 * c ) ) & 15u );
            if( v != 15 && h.border[ v ] && h.border[ ( v + h.n - 1 ) % h.n ] ) corners |= 1u << c;
        }
        out->rec[ key ].link[ 0 ] |= corners << 4;
    }
    // a neighbour without an edge for direction e' is asked through the cell's link e = 7 - e': give it end/start codes
    // that nobody expects there, so the comparison fails without a separate validity test
    for( int e = 0; e < 8; e++ )
    {
        int never = -1;
        for( int code = 0; code < 16; code++ )
            if( !( ( expected[ e ] >> code ) & 1u ) ) never = code;
        for( unsigned key = 0; key < ( unsigned )kCellKeys; key++ )
        {
            if( has_edge[ key ][ 7 - e ] ) continue;
            if( never < 0 )
            {
                out->slow_keys = kCellKeys; // (cannot happen: a direction's shared edges touch only a few of the 16 points)
                continue;
            }
            out->rec[ key ].nbr[ 7 - e ] = ( uint16_t )( never << 8 | never << 12 );
        }
    }
    if( out->slow_keys == ( uint32_t )kCellKeys )
        for( unsigned key = 0; key < ( unsigned )kCellKeys; key++ ) out->rec[ key ].link[ 0 ] = kSmoothSlow;

    // ---- the kernel's view: IDs of the neighbour records, HEAD / HEAD2 / PACK words, per-class ID -> record lists ----
    uint16_t id_rec[ 8 ][ kNbrIds ];
    int n_ids[ 8 ];
    bool ids_ok = true;
    for( int e = 0; e < 8; e++ )
    {
        n_ids[ e ] = 1; // ID 0: no hull edge through this direction
        for( int k = 0; k < kNbrIds; k++ ) id_rec[ e ][ k ] = 0xFFFFu;
        // the direction's different records in ascending order: the end / start codes are the high bits, so the records a
        // class accepts (those with its blended vertices) are neighbours in the class's block of the link table
        std::vector< uint16_t > recs;
        for( unsigned key = 0; key < ( unsigned )kCellKeys; key++ )
            if( has_edge[ key ][ e ] ) recs.push_back( out->rec[ key ].nbr[ e ] );
        std::sort( recs.begin(), recs.end() );
        recs.erase( std::unique( recs.begin(), recs.end() ), recs.end() );
        if( recs.size() >= ( size_t )kNbrIds )
        {
            ids_ok = false;
            recs.resize( kNbrIds - 1 );
        }
        for( size_t k = 0; k < recs.size(); k++ ) id_rec[ e ][ n_ids[ e ]++ ] = recs[ k ];
        for( unsigned key = 0; key < ( unsigned )kCellKeys; key++ )
        {
            out->nbr_id[ key ][ e ] = 0;
            if( !has_edge[ key ][ e ] ) continue;
            for( int k = 1; k < n_ids[ e ]; k++ )
                if( id_rec[ e ][ k ] == out->rec[ key ].nbr[ e ] ) out->nbr_id[ key ][ e ] = ( uint8_t )k;
            if( !out->nbr_id[ key ][ e ] ) ids_ok = false;
        }
    }
    for( LinkClass& c : out->classes )
    {
        for( int k = 0; k < kNbrIds; k++ ) c.nrec[ k ] = id_rec[ 7 - c.e ][ k ];
        // the IDs whose record holds the class's blended vertices
        int lo = kNbrIds, hi = -1, n_fit = 0;
        for( int k = 1; k < kNbrIds; k++ )
        {
            const uint32_t r = c.nrec[ k ];
            if( r == 0xFFFFu || ( c.hasA && ( ( r >> 8 ) & 15u ) != c.codeA ) || ( c.hasB && ( ( r >> 12 ) & 15u ) != c.codeB ) ) continue;
            lo = k < lo ? k : lo;
            hi = k > hi ? k : hi;
            n_fit++;
        }
        c.exact = n_fit > 0 && hi - lo + 1 == n_fit;
        c.id_lo = ( uint8_t )( c.exact ? lo : 0 );
        c.id_span = ( uint8_t )( c.exact ? hi - lo : kNbrIds - 1 );
        c.canon = 0;
        if( c.exact )
        {
            uint32_t k = 0;
            while( k < out->n_canon && ( out->canon_lo[ k ] != c.id_lo || out->canon_span[ k ] != c.id_span ) ) k++;
            if( k == out->n_canon && out->n_canon < 64 && out->link_entries / kNbrIds < 255 )
            {
                out->canon_lo[ k ] = c.id_lo;
                out->canon_span[ k ] = c.id_span;
                out->n_canon++;
                out->link_entries += kNbrIds;
            }
            if( k < out->n_canon ) c.canon = ( uint16_t )( out->classes.size() + 1 + k );
        }
    }
    for( unsigned key = 0; key < ( unsigned )kCellKeys; key++ )
    {
        const SmoothRecord& r = out->rec[ key ];
        out->pack[ key ][ 0 ] = out->pack[ key ][ 1 ] = 0u;
        for( int e = 0; e < 4; e++ )
        {
            out->pack[ key ][ 0 ] |= ( uint32_t )out->nbr_id[ key ][ 4 + e ] << ( 12 + 5 * e );
            out->pack[ key ][ 1 ] |= ( uint32_t )out->nbr_id[ key ][ e ] << ( 5 * e );
        }
        uint32_t* w = out->desc[ key ];
        w[ 0 ] = w[ 1 ] = w[ 2 ] = w[ 3 ] = 0u;
        if( r.link[ 0 ] == kSmoothSlow || !ids_ok )
        {
            w[ 0 ] = kDescSlow;
            continue;
        }
        for( int k = 0; k < kMaxLinks; k++ )
        {
            if( !( r.link[ k ] >> 16 ) ) continue;
            const int e = ( int )( r.link[ k ] & 7u );
            static const int di[ 8 ] = { -1, 0, 1, -1, 1, -1, 0, 1 }, dj[ 8 ] = { 1, 1, 1, 0, 0, -1, -1, -1 };
            // the neighbour's word that holds the ID for its direction 7 - e: directions 4..7 (e < 4) in its x-word at
            // bits [12 + 5 (3 - e)), directions 0..3 (e >= 4) in its y-word, half a row further, at bits [5 (7 - e))
            const int woff = dj[ e ] * kHeadRowWords + di[ e ] + ( e >= 4 ? kHeadRowWords / 2 : 0 ) + kHeadRowWords + 1;
            const int shift = e < 4 ? 12 + 5 * ( 3 - e ) : 5 * ( 7 - e );
            w[ k ] = ( uint32_t )woff | ( uint32_t )shift << 8 | ( r.link[ k ] >> 24 ) << 13 | kDescUsed;
        }
        if( w[ 2 ] ) w[ 0 ] |= kDescMore;
    }
}

} // namespace par
