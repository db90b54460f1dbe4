// Host build of the product's polygon logic (polygon.cuh is __host__ __device__): lets the CPU test suite
// check stage D+E of the product against the oracle without a GPU.  Not part of the shipped library.
#include <cstdint>
#include <cstring>
#define __restrict__
#include "../pixel_art_remaster_gpu_b200/csrc/polygon.cuh"
using namespace par;
namespace {
struct HostEnv
{
    const uint8_t* graph;
    int width, height;
    FlatImage img;
    uint32_t node( int i, int j ) const { return ( i >= 0 && j >= 0 && i < width && j < height ) ? graph[ ( size_t )j * width + i ] : 0u; }
    uint32_t key( int i, int j ) const { return cell_key( node( i, j ), node( i - 1, j ), node( i + 1, j ) ); }
    bool guard( int i, int j ) const { return img.guard( i, j ); }
    bool keep_corner( int i, int j, Q2 p ) const { return img.keep_corner( i, j, p ); }
};
struct HostSlots
{
    int x[ 16 ], y[ 16 ];
    void put( int slot, int x64, int y64 ) { x[ slot ] = x64; y[ slot ] = y64; }
};
} // namespace
extern "C" void host_polygons( const uint8_t* img_zero_tailed, const uint8_t* graph, int W, int H, int ws, int subdivide, float* poly /* N*45*2 */, int* count )
{
    static CellTables T;
    static bool built = false;
    if( !built ) { build_cell_tables( &T ); built = true; }
    HostEnv env;
    env.graph = graph; env.width = W; env.height = H;
    env.img.frame = img_zero_tailed; env.img.width = W; env.img.height = H; env.img.widthstep = ws;
    CellTablePtrs tab{ T.rec };
    memset( poly, 0, sizeof( float ) * 90 * ( size_t )W * H );
    for( int j = 0; j < H; j++ )
        for( int i = 0; i < W; i++ )
        {
            HostSlots s;
            CellPoly p = build_cell_polygon( env, tab, i, j, env.key( i, j ), subdivide != 0, s );
            float* o = poly + ( size_t )( j * W + i ) * 90;
            int m = 0;
            for( int t = 0; t < p.n; t++ )
                for( int e = 0; e <= ( int )( ( p.two >> t ) & 1u ); e++, m++ ) { o[ 2 * m ] = s.x[ 2 * t + e ] / 64.0f; o[ 2 * m + 1 ] = s.y[ 2 * t + e ] / 64.0f; }
            count[ j * W + i ] = m;
        }
}

// ---- smoothing tables without a GPU ------------------------------------------------------------------------------
// The raster kernel does not build the polygon of a smoothed cell: it XORs precomputed coverage pieces
// (csrc/smooth_table.h).  This check rebuilds those pieces on the host from the same host-side tables and descriptors
// the kernel reads (SmoothTables) and compares their XOR with the coverage of the polygon build_cell_polygon produces,
// sample by sample, on a window of (S + 2 halo)^2 samples around the cell.  Even-odd rule, exact integers: vertices are
// multiples of 1/64 px, samples odd multiples of 1/(2S) px, both scaled to units of 1/(128 S) px.
#include "../pixel_art_remaster_gpu_b200/csrc/smooth_table.h"
#include <vector>
namespace {
struct Pt { long x, y; }; // 1/64 px, cell-local
typedef std::vector< unsigned char > Mask;
void cover( const std::vector< Pt >& poly, int S, int halo, Mask& m )
{
    const int N = S + 2 * halo;
    for( size_t k = 0; k < poly.size(); k++ )
    {
        const Pt a = poly[ k ], b = poly[ ( k + 1 ) % poly.size() ];
        const long x0 = a.x * 2 * S, y0 = a.y * 2 * S, x1 = b.x * 2 * S, y1 = b.y * 2 * S, dx = x1 - x0, dy = y1 - y0;
        if( dy == 0 ) continue;
        for( int r = 0; r < N; r++ )
        {
            const long ys = ( 2 * ( r - halo ) + 1 ) * 64;
            if( ( y0 < ys ) == ( y1 < ys ) ) continue; // the edge does not cross this sample row
            for( int c = 0; c < N; c++ )
            {
                const long xs = ( 2 * ( c - halo ) + 1 ) * 64;
                const bool left = dy > 0 ? ( xs - x0 ) * dy < ( ys - y0 ) * dx : ( xs - x0 ) * dy > ( ys - y0 ) * dx; // strictly left of the crossing
                if( left ) m[ r * N + c ] ^= 1;
            }
        }
    }
}
Pt cut( Q2 p, Q2 x ) { int ox, oy; cut_toward( p, x, ox, oy ); return Pt{ ox, oy }; }
Q2 point_of( int code ) // inverse of point_code
{
    int seen = 0;
    for( int pos = 0; pos < 49; pos++ )
        if( ( kValidPoints >> pos ) & 1ull )
        {
            if( seen == code ) return Q2{ pos % 7 - 1, pos / 7 - 1 };
            seen++;
        }
    return Q2{ 0, 0 };
}
} // namespace
// out[0] = smoothed cells checked, out[1] = of which the tables could not express (would take the geometric path),
// out[2] = cells whose XOR of pieces differs from the polygon's coverage (must be 0), out[3] = link classes
extern "C" void host_smooth_check( const uint8_t* img_zero_tailed, const uint8_t* graph, int W, int H, int ws, int S, int halo, long* out )
{
    static CellTables T;
    static SmoothTables ST;
    static bool built = false;
    if( !built ) { build_cell_tables( &T ); build_smooth_tables( T, &ST ); built = true; }
    HostEnv env;
    env.graph = graph; env.width = W; env.height = H;
    env.img.frame = img_zero_tailed; env.img.width = W; env.img.height = H; env.img.widthstep = ws;
    CellTablePtrs tab{ T.rec };
    const int N = S + 2 * halo;
    out[ 0 ] = out[ 1 ] = out[ 2 ] = 0;
    out[ 3 ] = ( long )ST.classes.size();
    for( int j = 0; j < H; j++ )
        for( int i = 0; i < W; i++ )
        {
            const uint32_t key = env.key( i, j );
            if( ( key & 0xFFu ) == 90u ) continue;
            out[ 0 ]++;
            // the polygon itself
            HostSlots s;
            const CellPoly p = build_cell_polygon( env, tab, i, j, key, true, s );
            std::vector< Pt > poly;
            for( int t = 0; t < p.n; t++ )
                for( int e = 0; e <= ( int )( ( p.two >> t ) & 1u ); e++ ) poly.push_back( Pt{ s.x[ 2 * t + e ], s.y[ 2 * t + e ] } );
            Mask direct( N * N, 0 ), pieces( N * N, 0 );
            cover( poly, S, halo, direct );
            // the pieces, as the kernel assembles them
            const SmoothRecord& rec = ST.rec[ key ];
            if( ( rec.link[ 0 ] == kSmoothSlow ) != ( ( ST.desc[ key ][ 0 ] & kDescSlow ) != 0u ) ) out[ 2 ]++;
            if( rec.link[ 0 ] == kSmoothSlow ) { out[ 1 ]++; continue; }
            const uint64_t h = T.rec[ key ].verts, info = T.rec[ key ].info;
            const int n = hull_count( info );
            const VertexClasses cls = classify_vertices( info );
            uint32_t cf = 16u;
            if( !env.guard( i, j ) )
            {
                const Q2 corner[ 4 ] = { { 0, 0 }, { 4, 0 }, { 4, 4 }, { 0, 4 } };
                cf = 0u;
                for( int c = 0; c < 4; c++ )
                    if( env.keep_corner( i, j, corner[ c ] ) ) cf |= 1u << c;
            }
            {   // CUT[key][kept] (every cut vertex stays under checkTJunction's early exit)
                const uint32_t kept = cf & ( rec.link[ 0 ] >> 4 ) & 15u, cv = hull_corner_vertices( info );
                uint32_t keptv = 0u;
                for( int c = 0; c < 4; c++ )
                {
                    const uint32_t v = ( cv >> ( 4 * c ) ) & 15u;
                    if( ( ( kept >> c ) & 1u ) && v != 15u ) keptv |= 1u << v;
                }
                std::vector< Pt > q;
                for( int t = 0; t < n; t++ )
                {
                    const Q2 v = hull_vertex( h, t );
                    if( !( cf & 16u ) && ( ( cls.cut >> t ) & 1u ) && !( ( keptv >> t ) & 1u ) )
                    {
                        q.push_back( cut( v, hull_vertex( h, t == 0 ? n - 1 : t - 1 ) ) );
                        q.push_back( cut( v, hull_vertex( h, t + 1 == n ? 0 : t + 1 ) ) );
                    }
                    else
                        q.push_back( Pt{ 16 * v.x, 16 * v.y } );
                }
                cover( q, S, halo, pieces );
            }
            bool ok = true;
            for( int k = 0; k < kMaxLinks && ok; k++ )
            {
                const uint32_t d = rec.link[ k ];
                if( !( d >> 16 ) ) continue;
                const int e = ( int )( d & 7u );
                const uint32_t nkey = env.key( i + edge_di( e ), j + edge_dj( e ) );
                const uint32_t r = ST.rec[ nkey ].nbr[ e ^ 7 ];
                const uint32_t ends = ( d >> 16 ) & 255u;
                const LinkClass& c = ST.classes[ ( d >> 24 ) - 1 ];
                {   // the kernel's view of the same lookup (HEAD / PACK / NBR_ID / per-class ID lists) must agree with the records
                    // (a tile row of cell words: 36 x-words = key | IDs of directions 4..7 << 12, then 36 y-words = IDs of 0..3)
                    const uint32_t hd = ST.desc[ key ][ k ];
                    if( ( ( hd >> 13 ) & 255u ) != ( d >> 24 ) || !( hd & kDescUsed ) ) out[ 2 ]++;
                    const int woff = ( int )( hd & 255u ) - ( kHeadRowWords + 1 ); // from the cell's own x-word
                    const int row = ( woff + kHeadRowWords / 4 + 4 * kHeadRowWords ) / kHeadRowWords - 4; // (floor)
                    const int rest = woff - row * kHeadRowWords; // column offset, + 36 for the y-word
                    const bool yword = rest > kHeadRowWords / 4;
                    const int col = yword ? rest - kHeadRowWords / 2 : rest;
                    if( row != edge_dj( e ) || col != edge_di( e ) || yword != ( e >= 4 ) ) out[ 2 ]++;
                    const uint32_t half = yword ? ST.pack[ nkey ][ 1 ] : ( nkey | ST.pack[ nkey ][ 0 ] );
                    const uint32_t id = ( half >> ( ( hd >> 8 ) & 31u ) ) & 31u;
                    if( id != ST.nbr_id[ nkey ][ e ^ 7 ] ) out[ 2 ]++;
                    const uint32_t r2 = c.nrec[ id ];
                    // the class's own block says MISMATCH for exactly the records that do not hold its blended vertices, and so does
                    // the block shared by the classes with its ID range, when it has one
                    const bool entry_mismatch = id == 0u || r2 == 0xFFFFu || ( c.hasA && ( ( r2 >> 8 ) & 15u ) != c.codeA ) || ( c.hasB && ( ( r2 >> 12 ) & 15u ) != c.codeB );
                    const bool mis = ( ( ( r ^ d ) >> 8 ) & ends ) != 0u;
                    if( entry_mismatch != mis ) out[ 2 ]++;
                    if( c.canon )
                    {
                        const uint32_t cn = c.canon - ( uint32_t )ST.classes.size() - 1u;
                        if( cn >= ST.n_canon || ( id - ST.canon_lo[ cn ] > ( uint32_t )ST.canon_span[ cn ] ) != mis ) out[ 2 ]++;
                    }
                    if( !mis && ( ( r2 ^ r ) & ends ) != 0u ) out[ 2 ]++;
                }
                if( ( ( r ^ d ) >> 8 ) & ends ) { ok = false; break; }
                const uint32_t sub = r & ends;
                const int di = edge_di( c.e ), dj = edge_dj( c.e );
                const Q2 P0{ c.px[ 0 ], c.py[ 0 ] }, P1{ c.px[ 1 ], c.py[ 1 ] }, P2{ c.px[ 2 ], c.py[ 2 ] }, P3{ c.px[ 3 ], c.py[ 3 ] };
                std::vector< Pt > q;
                if( c.hasA )
                {
                    const Pt R = cut( P1, P0 ), Q = cut( Q2{ P1.x - 4 * di, P1.y - 4 * dj }, point_of( c.after[ sub & 15u ] ) );
                    q.push_back( R );
                    q.push_back( Pt{ ( R.x + Q.x + 64 * di ) >> 1, ( R.y + Q.y + 64 * dj ) >> 1 } );
                }
                else
                    q.push_back( Pt{ 16 * P1.x, 16 * P1.y } );
                if( c.hasB )
                {
                    const Pt Q = cut( P2, P3 ), R = cut( Q2{ P2.x - 4 * di, P2.y - 4 * dj }, point_of( c.before[ sub >> 4 ] ) );
                    q.push_back( Pt{ ( Q.x + R.x + 64 * di ) >> 1, ( Q.y + R.y + 64 * dj ) >> 1 } );
                    q.push_back( Q );
                }
                q.push_back( Pt{ 16 * P2.x, 16 * P2.y } );
                if( c.hasA ) q.push_back( Pt{ 16 * P1.x, 16 * P1.y } );
                cover( q, S, halo, pieces );
            }
            if( !ok ) { out[ 1 ]++; continue; }
            if( direct != pieces ) out[ 2 ]++;
        }
}
