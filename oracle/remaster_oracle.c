/* TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * CPU restatement, in plain C, of the per-frame pixel-art remaster path of
 * marcoc2/pixel-art-remaster-gpu.  It exists to CHECK the CUDA product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.  The
 * product never calls it and has no CPU fallback.
 *
 * Every function cites the reference file:line whose behaviour it restates.  Nothing here is copied
 * from the reference; the algorithms are re-expressed from SURVEY.md Appendix A.  The restatement
 * is PINNED (tests/test_oracle_vs_reference.py, tests/test_golden.py):
 *   - against the reference's own code compiled as host C++ (oracle/_ref/libref_host*.so, built
 *     from /root/reference by oracle/Makefile) on seeded synthetic frames, stage by stage;
 *   - against the reference's embedded graph dumps (kernel.cu:291-296) and alex_png.txt;
 *   - against golden vectors generated from that reference build and committed under tests/golden/.
 * Parts with no runnable reference counterpart are pinned by definition only and say so:
 *   - connected-component labels (reference labeller is dead code, cc_functions.cu): canonical
 *     min-row-major-index definition + the alex_png.txt fixture;
 *   - rasterization (done by the OpenGL driver in the reference, simpleVBO.cpp:281): a restated
 *     point-sampling rule over the reference's triangle list, "parity unpinned".
 *
 * Conventions: pixel (i,j), i = column, j = row, row 0 = bottom scanline;
 * node index n = j*W + i; colour bytes img[j*ws + 3*i + {0,1,2}]; graph bit e <-> neighbour
 *   e        0      1      2      3      4      5      6      7
 *   (di,dj) (-1,+1) (0,+1) (+1,+1) (-1,0) (+1,0) (-1,-1) (0,-1) (+1,-1)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_SLOTS 45 /* kernel.cu:8 CELL_SIZE */

static const int ORC_DI[ 8 ] = { -1, 0, 1, -1, 1, -1, 0, 1 };
static const int ORC_DJ[ 8 ] = { 1, 1, 1, 0, 0, -1, -1, -1 };

/* ------------------------------------------------------------------------------------------- */
/* Stage A: similarity graph.  graph_functions.cu:80-98 (RGBtoYUV), :102-131 (DATAtoINT),        */
/* :147-311 (diff); kernel.cu:140-159 (graph_Kernel).                                            */
/* ------------------------------------------------------------------------------------------- */

/* Packed YUV word of the colour whose bytes in memory are (b0,b1,b2).  fused != 0 evaluates Y with
 * the FMA nesting nvcc emits for graph_functions.cu:90 on the device:
 *   y = trunc( fma(0.114, b2, fma(0.299, b0, 0.587*b1)) );   fused == 0 is the plain host order. */
uint32_t orc_yuv_word( int b0, int b1, int b2, int fused )
{
    double r = ( double )( float )b0, g = ( double )( float )b1, b = ( double )( float )b2;
    int y;
    if( fused )
        y = ( int )fma( 0.114, b, fma( 0.299, r, 0.587 * g ) );
    else
    {
        double s = 0.299 * r;
        double s2 = 0.587 * g;
        double s3 = 0.114 * b;
        y = ( int )( ( s + s2 ) + s3 );
    }
    int u = ( int )( ( float )( b2 - y ) * 0.492f ); /* graph_functions.cu:93 */
    int v = ( int )( ( float )( b0 - y ) * 0.877f ); /* graph_functions.cu:94 */
    return ( uint32_t )( y * 65536 ) + ( uint32_t )( u * 256 ) + ( uint32_t )v; /* :97, two's complement */
}

/* all 2^24 colours at once (c = byte0 | byte1<<8 | byte2<<16), for the exhaustive tests */
void orc_yuv_all( int fused, uint32_t* out )
{
    for( uint32_t c = 0; c < ( 1u << 24 ); c++ ) out[ c ] = orc_yuv_word( c & 255, ( c >> 8 ) & 255, c >> 16, fused );
}

static int orc_abs_i32( uint32_t d )
{
    int a = ( int )d;
    return a < 0 ? -a : a;
}

/* 1 when the two packed words are dissimilar (graph_functions.cu:291-293; thresholds :14-19) */
int orc_dissimilar( uint32_t p, uint32_t q )
{
    if( orc_abs_i32( ( p & 0x00FF0000u ) - ( q & 0x00FF0000u ) ) > 0x00050000 ) return 1;
    if( orc_abs_i32( ( p & 0x0000FF00u ) - ( q & 0x0000FF00u ) ) > 0x00000700 ) return 1;
    if( orc_abs_i32( ( p & 0x000000FFu ) - ( q & 0x000000FFu ) ) > 0x00000006 ) return 1;
    return 0;
}

void orc_similarity_graph( const uint8_t* img, int W, int H, int ws, int fused, uint8_t* graph )
{
    uint32_t* yuv = ( uint32_t* )malloc( sizeof( uint32_t ) * ( size_t )W * H );
    for( int j = 0; j < H; j++ )
        for( int i = 0; i < W; i++ )
        {
            const uint8_t* p = img + ( size_t )j * ws + 3 * i;
            yuv[ j * W + i ] = orc_yuv_word( p[ 0 ], p[ 1 ], p[ 2 ], fused );
        }
    for( int j = 0; j < H; j++ )
        for( int i = 0; i < W; i++ )
        {
            unsigned byte = 0;
            for( int e = 0; e < 8; e++ )
            {
                int ni = i + ORC_DI[ e ], nj = j + ORC_DJ[ e ];
                if( ni < 0 || nj < 0 || ni >= W || nj >= H ) continue; /* graph_functions.cu:245-247 */
                if( !orc_dissimilar( yuv[ j * W + i ], yuv[ nj * W + ni ] ) ) byte |= 1u << e;
            }
            graph[ j * W + i ] = ( uint8_t )byte;
        }
    free( yuv );
}

/* ------------------------------------------------------------------------------------------- */
/* Stage B: trivial crossings.  graph_functions.cu:1211-1274 (crossCheck_4).                     */
/* ------------------------------------------------------------------------------------------- */
static unsigned orc_g( const uint8_t* g, int W, int H, int i, int j )
{
    return ( i < 0 || j < 0 || i >= W || j >= H ) ? 0u : g[ j * W + i ];
}

/* Only diagonal bits are cleared and only orthogonal bits are read, so this may run in place the
 * way the reference kernel does (kernel.cu:162-177); written out-of-place here for clarity. */
void orc_trivial_crossings( const uint8_t* in, int W, int H, uint8_t* out )
{
    for( int j = 0; j < H; j++ )
        for( int i = 0; i < W; i++ )
        {
            unsigned c = in[ j * W + i ];
            unsigned l = orc_g( in, W, H, i - 1, j ), r = orc_g( in, W, H, i + 1, j );
            unsigned u = orc_g( in, W, H, i, j + 1 ), d = orc_g( in, W, H, i, j - 1 );
            unsigned ul = orc_g( in, W, H, i - 1, j + 1 ), ur = orc_g( in, W, H, i + 1, j + 1 );
            unsigned dl = orc_g( in, W, H, i - 1, j - 1 ), dr = orc_g( in, W, H, i + 1, j - 1 );
            unsigned o = c;
            /* up-left block: sides l-ul (l bit1), l-c (l bit4), ul-u (ul bit4, ul bit6), c (3,1), u (6,3)  :1239-1245 */
            if( ( l & 2 ) && ( l & 16 ) && ( ul & 64 ) && ( ul & 16 ) && ( c & 8 ) && ( c & 2 ) && ( u & 64 ) && ( u & 8 ) ) o &= ~1u;
            /* up-right block  :1248-1254 */
            if( ( c & 2 ) && ( c & 16 ) && ( u & 64 ) && ( u & 16 ) && ( r & 8 ) && ( r & 2 ) && ( ur & 64 ) && ( ur & 8 ) ) o &= ~4u;
            /* down-left block  :1257-1263 */
            if( ( dl & 2 ) && ( dl & 16 ) && ( l & 64 ) && ( l & 16 ) && ( d & 8 ) && ( d & 2 ) && ( c & 64 ) && ( c & 8 ) ) o &= ~32u;
            /* down-right block  :1266-1272 */
            if( ( d & 2 ) && ( d & 16 ) && ( c & 64 ) && ( c & 16 ) && ( dr & 8 ) && ( dr & 2 ) && ( r & 64 ) && ( r & 8 ) ) o &= ~128u;
            out[ j * W + i ] = ( uint8_t )o;
        }
}

/* ------------------------------------------------------------------------------------------- */
/* Stage C: ambiguous crossings.  graph_functions.cu:1433-1518 (crossCheck_Heuristics),          */
/* :761-849 (processHeuristics2), :421-479 (valence tests), :541-573 (calcVal2PathSize).         */
/* ------------------------------------------------------------------------------------------- */
static int orc_popcount8( unsigned v )
{
    int c = 0;
    for( v &= 0xFFu; v; v &= v - 1 ) c++;
    return c;
}

/* number of links of node n other than edge e (the loops at graph_functions.cu:427-442, :462-469) */
static int orc_others( const uint8_t* aux, int n, int e ) { return orc_popcount8( aux[ n ] & ~( 1u << e ) ); }

/* valence-2 chain length from node n entered through edge e; r carries across calls and is capped
 * so that it never exceeds 31 (graph_functions.cu:541-573; the recursion is unrolled into a loop) */
static void orc_chain( const uint8_t* aux, int W, int n, int e, int* r )
{
    for( ;; )
    {
        unsigned m = aux[ n ] & ~( 1u << e ) & 0xFFu;
        if( *r > 30 ) return;
        if( orc_popcount8( m ) != 1 ) return;
        int k = 0;
        while( !( m & ( 1u << k ) ) ) k++;
        ( *r )++;
        n += ORC_DJ[ k ] * W + ORC_DI[ k ]; /* calc_index, graph_functions.cu:46-76 */
        e = 7 - k;                          /* conected_edge, :26-28 */
    }
}

/* decision for the 2x2 block whose lower-left pixel is (bi,bj): 1 = the "/" diagonal dies,
 * 0 = the "\" diagonal dies.  *steps (may be NULL) receives the two chain totals when rule 5 ran. */
int orc_block_decision( const uint8_t* aux, int W, int bi, int bj, int* steps )
{
    int i1 = bj * W + bi, i2 = i1 + W, i3 = i1 + 1, i4 = i2 + 1;
    int o1 = orc_others( aux, i1, 2 ), o4 = orc_others( aux, i4, 5 );
    int o3 = orc_others( aux, i3, 0 ), o2 = orc_others( aux, i2, 7 );
    if( steps ) steps[ 0 ] = steps[ 1 ] = -1;
    if( o1 == 1 && o4 == 1 ) return 0;                               /* :781-785 */
    if( o3 == 1 && o2 == 1 ) return 1;                               /* :791-795 */
    if( ( o1 == 0 || o4 == 0 ) && o3 != 0 && o2 != 0 ) return 0;     /* :798-802 */
    if( o3 == 0 || ( o2 == 0 && o1 != 0 && o4 != 0 ) ) return 1;     /* :805-809, C precedence kept */
    int s = 0, s2 = 0;
    orc_chain( aux, W, i1, 2, &s );
    orc_chain( aux, W, i4, 5, &s );
    orc_chain( aux, W, i3, 0, &s2 );
    orc_chain( aux, W, i2, 7, &s2 );
    if( steps ) { steps[ 0 ] = s; steps[ 1 ] = s2; }
    return !( s2 < s );                                              /* :830-839, tie removes "/" */
}

/* aux = stage-B output; out = final graph.  n_ambiguous (may be NULL) counts ambiguous blocks. */
void orc_resolve_crossings( const uint8_t* aux, int W, int H, uint8_t* out, int* n_ambiguous )
{
    memcpy( out, aux, ( size_t )W * H );
    int cnt = 0;
    for( int bj = 0; bj + 1 < H; bj++ )
        for( int bi = 0; bi + 1 < W; bi++ )
        {
            int i1 = bj * W + bi, i2 = i1 + W, i3 = i1 + 1, i4 = i2 + 1;
            if( !( ( aux[ i1 ] & 4 ) && ( aux[ i2 ] & 128 ) && ( aux[ i3 ] & 1 ) && ( aux[ i4 ] & 32 ) ) ) continue;
            cnt++;
            if( orc_block_decision( aux, W, bi, bj, NULL ) )
            {
                out[ i1 ] &= ( uint8_t )~4u;   /* :1509-1518 */
                out[ i4 ] &= ( uint8_t )~32u;  /* :1487-1496 */
            }
            else
            {
                out[ i2 ] &= ( uint8_t )~128u; /* :1498-1507 */
                out[ i3 ] &= ( uint8_t )~1u;   /* :1476-1485 */
            }
        }
    if( n_ambiguous ) *n_ambiguous = cnt;
}

/* ------------------------------------------------------------------------------------------- */
/* Stage D: cell of a pattern.  diagram_functions.cu:319-535 (createCellFromPattern),            */
/* :238-316 (convex_hull), :82-129 (sort), :36-80 (sort_y), :225-230 (cross).                    */
/* Coordinates are handled in quarter units (value*4), which is exact for every candidate point. */
/* ------------------------------------------------------------------------------------------- */
typedef struct { int x, y; } orc_q4; /* quarter units */

static void orc_swap_q4( orc_q4* a, orc_q4* b ) { orc_q4 t = *a; *a = *b; *b = t; }

/* the reference's one-pass tie ordering (diagram_functions.cu:36-80) */
static void orc_tie_pass_y( orc_q4* P, int start, int end )
{
    int best = 99 * 4, ties = 0, where[ ORC_SLOTS ];
    for( int i = start; i <= end; i++ )
    {
        if( P[ i ].y < best ) { best = P[ i ].y; ties = 0; where[ 0 ] = i; }
        else if( P[ i ].y == best ) { ties++; where[ ties ] = i; }
    }
    if( ties > 0 )
        for( int i = 0; i <= ties; i++ ) orc_swap_q4( &P[ start + i ], &P[ where[ i ] ] );
    else
        orc_swap_q4( &P[ start ], &P[ where[ 0 ] ] );
}

/* the reference's selection "sort" (diagram_functions.cu:82-129): NOT a full lexicographic sort
 * when more than two points share an x; reproduced as is because the hull depends on it */
static void orc_quirky_sort( orc_q4* P, int n )
{
    int done = 0;
    while( done < n )
    {
        int best = 99 * 4, ties = 0, where[ ORC_SLOTS ];
        for( int i = done; i < n; i++ )
        {
            if( P[ i ].x < best ) { best = P[ i ].x; ties = 0; where[ 0 ] = i; }
            else if( P[ i ].x == best ) { ties++; where[ ties ] = i; }
        }
        if( ties > 0 )
        {
            for( int i = 0; i <= ties; i++ ) orc_swap_q4( &P[ done + i ], &P[ where[ i ] ] );
            orc_tie_pass_y( P, done, done + ties );
        }
        else
            orc_swap_q4( &P[ done ], &P[ where[ 0 ] ] );
        done += ties + 1;
    }
}

/* sign-compatible with diagram_functions.cu:225-230 evaluated on coordinates*100 (all exact) */
static int orc_turn( orc_q4 o, orc_q4 a, orc_q4 b ) { return ( a.x - o.x ) * ( b.y - o.y ) - ( a.y - o.y ) * ( b.x - o.x ); }

/* candidate points of a pattern in the reference's emission order (diagram_functions.cu:322-533) */
static int orc_cell_candidates( unsigned node, unsigned left, unsigned right, orc_q4* c )
{
    int n = 0;
#define PUT( X, Y ) do { c[ n ].x = ( X ); c[ n ].y = ( Y ); n++; } while( 0 )
    int b0 = node & 1, b1 = node & 2, b2 = node & 4, b3 = node & 8, b4 = node & 16, b5 = node & 32, b6 = node & 64, b7 = node & 128;
    /* corner toward neighbour 0 (:326-360) */
    if( b0 ) { if( b1 && !b3 ) PUT( -1, 3 ); else if( b3 && !b1 ) PUT( 1, 5 ); else { PUT( -1, 3 ); PUT( 1, 5 ); } }
    else { if( left & 4 ) PUT( 1, 3 ); else PUT( 0, 4 ); }
    /* side toward neighbour 1 (:363-374) */
    if( b1 ) { PUT( 0, 4 ); PUT( 4, 4 ); } else PUT( 2, 3 );
    /* corner toward neighbour 2 (:376-409) */
    if( b2 ) { if( b1 && !b4 ) PUT( 5, 3 ); else if( b4 && !b1 ) PUT( 3, 5 ); else { PUT( 3, 5 ); PUT( 5, 3 ); } }
    else { if( right & 1 ) PUT( 3, 3 ); else PUT( 4, 4 ); }
    /* side toward neighbour 4 (:412-423) */
    if( b4 ) { PUT( 4, 4 ); PUT( 4, 0 ); } else PUT( 3, 3 );
    /* corner toward neighbour 7 (:425-458) */
    if( b7 ) { if( b4 && !b6 ) PUT( 3, -1 ); else if( b6 && !b4 ) PUT( 5, 1 ); else { PUT( 5, 1 ); PUT( 3, -1 ); } }
    else { if( right & 32 ) PUT( 3, 1 ); else PUT( 4, 0 ); }
    /* side toward neighbour 6 (:461-472) */
    if( b6 ) { PUT( 4, 0 ); PUT( 0, 0 ); } else PUT( 1, 3 );
    /* corner toward neighbour 5 (:474-507) */
    if( b5 ) { if( b3 && !b6 ) PUT( 1, -1 ); else if( b6 && !b3 ) PUT( -1, 1 ); else { PUT( 1, -1 ); PUT( -1, 1 ); } }
    else { if( left & 128 ) PUT( 1, 1 ); else PUT( 0, 0 ); }
    /* side toward neighbour 3 (:510-521) */
    if( b3 ) { PUT( 0, 0 ); PUT( 0, 4 ); } else PUT( 1, 2 );
#undef PUT
    return n;
}

/* hull of a pattern: out_xy receives count+1 points (the first repeated last) as floats in pixel
 * units, counter-clockwise; returns count (diagram_functions.cu:315 "return --k") */
int orc_cell_hull( unsigned node, unsigned left, unsigned right, float* out_xy )
{
    orc_q4 P[ ORC_SLOTS ], Hh[ 2 * ORC_SLOTS ];
    int n = orc_cell_candidates( node, left, right, P );
    orc_quirky_sort( P, n );
    int k = 0;
    for( int i = 0; i < n; i++ ) /* lower chain, :268-278 */
    {
        while( k >= 2 && orc_turn( Hh[ k - 2 ], Hh[ k - 1 ], P[ i ] ) <= 0 ) k--;
        Hh[ k++ ] = P[ i ];
    }
    for( int i = n - 2, t = k + 1; i >= 0; i-- ) /* upper chain, :287-292 */
    {
        while( k >= t && orc_turn( Hh[ k - 2 ], Hh[ k - 1 ], P[ i ] ) <= 0 ) k--;
        Hh[ k++ ] = P[ i ];
    }
    for( int i = 0; i < k; i++ )
    {
        out_xy[ 2 * i ] = ( float )Hh[ i ].x * 0.25f;
        out_xy[ 2 * i + 1 ] = ( float )Hh[ i ].y * 0.25f;
    }
    return k - 1;
}

/* all cells of a frame; hull = N*45*2 floats (unused slots zero), count = N ints.
 * Out-of-image neighbour bytes are 0 (SURVEY App. B-3 resolution; kernel.cu:204-207 reads the
 * linear neighbours n-1 / n+1, whose relevant bits are provably 0 at row ends). */
void orc_cells( const uint8_t* graph, int W, int H, float* hull, int* count )
{
    int N = W * H;
    memset( hull, 0, sizeof( float ) * 2 * ORC_SLOTS * ( size_t )N );
    for( int n = 0; n < N; n++ )
    {
        unsigned left = n > 0 ? graph[ n - 1 ] : 0u, right = n + 1 < N ? graph[ n + 1 ] : 0u;
        count[ n ] = orc_cell_hull( graph[ n ], left, right, hull + ( size_t )n * 2 * ORC_SLOTS );
    }
}

/* ------------------------------------------------------------------------------------------- */
/* Stage E: one level of corner cutting.  subdivision_functions.cu:564-669 (subdivision),        */
/* :245-424 (isLinkedEdge), :170-242 (checkTJunction), :42-122 (getQ_i/getR_i),                  */
/* :125-154 (linked-cell lookups), :427-538 (frame shifts, getPointIndex), :554 (midPoint).      */
/* ------------------------------------------------------------------------------------------- */
typedef struct { float x, y; } orc_pt;

static int orc_wrap( int i, int n ) { return ( n + ( i % n ) ) % n; } /* :3-6 */

/* which link (0..7) the polygon edge t lies on, -1 for a border edge (:245-424, first match wins) */
static int orc_edge_link( const orc_pt* P, int cnt, int t, unsigned node )
{
    orc_pt a = P[ t ], b = P[ orc_wrap( t + 1, cnt ) ];
    float slope = ( b.y - a.y ) / ( b.x - a.x );
    float mid_y = ( float )( ( ( double )a.y + ( double )b.y ) / 2.0 );
    if( a.y == b.y && a.y > 0.5 && ( node & 2 ) ) return 1;
    if( a.x == b.x && a.x > 0.5 && ( node & 16 ) ) return 4;
    if( a.y == b.y && a.y < 0.5 && ( node & 64 ) ) return 6;
    if( a.x == b.x && a.x < 0.5 && ( node & 8 ) ) return 3;
    if( slope == 1 && mid_y > 0.5 && ( node & 1 ) ) return 0;
    if( slope == -1 && mid_y > 0.5 && ( node & 4 ) ) return 2;
    if( slope == 1 && mid_y < 0.5 && ( node & 128 ) ) return 7;
    if( slope == -1 && mid_y < 0.5 && ( node & 32 ) ) return 5;
    return -1;
}

/* point 1/4 (or 1/8 on an edge longer than 1) of the way along edge i, from its start (:42-83) */
static orc_pt orc_q_point( const orc_pt* P, int cnt, int i )
{
    int pi = i % cnt;
    if( pi < 0 ) pi += cnt;
    orc_pt p = P[ pi ], q = P[ ( i + 1 ) % cnt ], o;
    float len = sqrtf( ( q.x - p.x ) * ( q.x - p.x ) + ( q.y - p.y ) * ( q.y - p.y ) );
    if( len <= 1.0 )
    {
        o.x = ( float )( p.x / 2.0 + ( ( double )p.x + q.x ) / 4.0 );
        o.y = ( float )( p.y / 2.0 + ( ( double )p.y + q.y ) / 4.0 );
    }
    else
    {
        o.x = ( float )( ( 7.0 * p.x ) / 8.0 + q.x / 8.0 );
        o.y = ( float )( ( 7.0 * p.y ) / 8.0 + q.y / 8.0 );
    }
    return o;
}

/* same from the far end (:86-122) */
static orc_pt orc_r_point( const orc_pt* P, int cnt, int i )
{
    int pi = i % cnt;
    if( pi < 0 ) pi += cnt;
    orc_pt p = P[ pi ], q = P[ ( i + 1 ) % cnt ], o;
    float len = sqrtf( ( q.x - p.x ) * ( q.x - p.x ) + ( q.y - p.y ) * ( q.y - p.y ) );
    if( len <= 1.0 )
    {
        o.x = ( float )( q.x / 2.0 + ( ( double )p.x + q.x ) / 4.0 );
        o.y = ( float )( q.y / 2.0 + ( ( double )p.y + q.y ) / 4.0 );
    }
    else
    {
        o.x = ( float )( ( 1.0 * p.x ) / 8.0 + ( 7.0 * q.x ) / 8.0 );
        o.y = ( float )( ( 1.0 * p.y ) / 8.0 + ( 7.0 * q.y ) / 8.0 );
    }
    return o;
}

/* first hull vertex equal to p, 0 when absent (:527-538) */
static int orc_find_vertex( orc_pt p, const orc_pt* P, int cnt )
{
    for( int i = 0; i < cnt; i++ )
        if( p.x == P[ i ].x && p.y == P[ i ].y ) return i;
    return 0;
}

/* keep-the-corner test (:170-242).  img must be followed by >= ws+8 readable zero bytes. */
static int orc_t_junction( const uint8_t* img, int W, int ws, int H, int i, int j, orc_pt p )
{
    long idx = ( long )j * ws + 3L * i;
    if( idx - ws - 1 < 0 || idx + W + 1 > ( long )H * ws - 1 ) return 1; /* :187, "width" as written */
    const uint8_t *c0 = img + idx + ws - 3, *c1 = img + idx + ws, *c2 = img + idx + ws + 3;
    const uint8_t *c3 = img + idx - 3, *c4 = img + idx + 3;
    const uint8_t *c5 = img + idx - ws - 3, *c6 = img + idx - ws, *c7 = img + idx - ws + 3;
#define SAME( a, b ) ( ( a )[ 0 ] == ( b )[ 0 ] && ( a )[ 1 ] == ( b )[ 1 ] && ( a )[ 2 ] == ( b )[ 2 ] )
    if( p.x == 0.0 && p.y == 0.0 && ( !SAME( c3, c5 ) || !SAME( c5, c6 ) ) ) return 1;
    if( p.x == 1.0 && p.y == 0.0 && ( !SAME( c4, c7 ) || !SAME( c7, c6 ) ) ) return 1;
    if( p.x == 1.0 && p.y == 1.0 && ( !SAME( c1, c2 ) || !SAME( c2, c4 ) ) ) return 1;
    if( p.x == 0.0 && p.y == 1.0 && ( !SAME( c0, c1 ) || !SAME( c1, c3 ) ) ) return 1;
#undef SAME
    return 0;
}

/* subdivide the cell of node n.  hull/hcount = PRE-subdivision hulls of the whole frame
 * (kernel.cu:434-452); out receives the new polygon; returns its vertex count. */
static int orc_subdivide_cell( const uint8_t* img_z, int W, int ws, int H, const float* hull, const int* hcount,
                               int n, unsigned node, orc_pt* out )
{
    const orc_pt* P = ( const orc_pt* )( hull + ( size_t )n * 2 * ORC_SLOTS );
    int cnt = hcount[ n ], i_ = n % W, j_ = n / W, m = 0;
    int link[ ORC_SLOTS ];
    for( int t = 0; t < cnt; t++ ) link[ t ] = orc_edge_link( P, cnt, t, node );
    for( int t = 0; t < cnt; t++ )
    {
        int prev = orc_wrap( t - 1, cnt );
        int cur_linked = link[ t ] >= 0, prev_linked = link[ prev ] >= 0;
        if( !cur_linked && !prev_linked )
        {
            if( orc_t_junction( img_z, W, ws, H, i_, j_, P[ t ] ) ) out[ m++ ] = P[ t ];
            else
            {
                orc_pt q = orc_q_point( P, cnt, t ), r = orc_r_point( P, cnt, t - 1 );
                out[ m++ ] = r;
                out[ m++ ] = q;
            }
        }
        else if( !cur_linked && prev_linked )
        {
            int L = link[ prev ];
            int nb = n + ORC_DJ[ L ] * W + ORC_DI[ L ];
            const orc_pt* Q = ( const orc_pt* )( hull + ( size_t )nb * 2 * ORC_SLOTS );
            orc_pt opp = { P[ t ].x - ( float )ORC_DI[ L ], P[ t ].y - ( float )ORC_DJ[ L ] }; /* :427-474 */
            int op = orc_find_vertex( opp, Q, hcount[ nb ] );
            orc_pt r = orc_r_point( Q, hcount[ nb ], op - 1 );
            r.x += ( float )ORC_DI[ L ];                                                        /* :477-524 */
            r.y += ( float )ORC_DJ[ L ];
            orc_pt q = orc_q_point( P, cnt, t ), mid;
            mid.x = ( float )( ( ( double )q.x + r.x ) / 2.0 );
            mid.y = ( float )( ( ( double )q.y + r.y ) / 2.0 );
            out[ m++ ] = mid;
            out[ m++ ] = q;
        }
        else if( cur_linked && !prev_linked )
        {
            int L = link[ t ];
            int nb = n + ORC_DJ[ L ] * W + ORC_DI[ L ];
            const orc_pt* Q = ( const orc_pt* )( hull + ( size_t )nb * 2 * ORC_SLOTS );
            orc_pt opp = { P[ t ].x - ( float )ORC_DI[ L ], P[ t ].y - ( float )ORC_DJ[ L ] };
            int op = orc_find_vertex( opp, Q, hcount[ nb ] );
            orc_pt qa = orc_q_point( Q, hcount[ nb ], op );
            qa.x += ( float )ORC_DI[ L ];
            qa.y += ( float )ORC_DJ[ L ];
            orc_pt r = orc_r_point( P, cnt, t - 1 ), mid;
            mid.x = ( float )( ( ( double )r.x + qa.x ) / 2.0 );
            mid.y = ( float )( ( ( double )r.y + qa.y ) / 2.0 );
            out[ m++ ] = r;
            out[ m++ ] = mid;
        }
        else
            out[ m++ ] = P[ t ];
    }
    return m;
}

/* poly/pcount start as copies of hull/hcount; cells whose node byte is 90 are left alone
 * (kernel.cu:231).  img must hold H*ws bytes; a zero tail is appended internally. */
void orc_subdivide( const uint8_t* img, int W, int H, int ws, const uint8_t* graph, const float* hull, const int* hcount,
                    float* poly, int* pcount )
{
    int N = W * H;
    uint8_t* z = ( uint8_t* )calloc( ( size_t )ws * H + 2 * ( size_t )ws + 64, 1 );
    memcpy( z, img, ( size_t )ws * H );
    memcpy( poly, hull, sizeof( float ) * 2 * ORC_SLOTS * ( size_t )N );
    memcpy( pcount, hcount, sizeof( int ) * ( size_t )N );
    for( int n = 0; n < N; n++ )
    {
        if( graph[ n ] == 90 ) continue;
        orc_pt tmp[ ORC_SLOTS ];
        int m = orc_subdivide_cell( z, W, ws, H, hull, hcount, n, graph[ n ], tmp );
        memcpy( poly + ( size_t )n * 2 * ORC_SLOTS, tmp, sizeof( orc_pt ) * m );
        pcount[ n ] = m;
    }
    free( z );
}

/* ------------------------------------------------------------------------------------------- */
/* Stage F: ear clipping.  triangulate_functions.cu:108-217 (process), :70-105 (snip),           */
/* :41-67 (insideTriangle), :6-34 (area), :220-276 (triangulate_polygon, adds the pixel offset). */
/* ------------------------------------------------------------------------------------------- */
static int orc_blocks_ear( orc_pt a, orc_pt b, orc_pt c, orc_pt p )
{
    float e0 = ( c.x - b.x ) * ( p.y - b.y ) - ( c.y - b.y ) * ( p.x - b.x );
    float e1 = ( b.x - a.x ) * ( p.y - a.y ) - ( b.y - a.y ) * ( p.x - a.x );
    float e2 = ( a.x - c.x ) * ( p.y - c.y ) - ( a.y - c.y ) * ( p.x - c.x );
    return e0 >= 0.0f && e2 >= 0.0f && e1 >= 0.0f;
}

/* tri receives 3 points per triangle in LOCAL coordinates; returns the number of triangles, or the
 * negative of the number produced before the reference's loop guard gave up (:153-156) */
int orc_ear_clip( const float* poly_xy, int cnt, float* tri_xy )
{
    const orc_pt* C = ( const orc_pt* )poly_xy;
    orc_pt* T = ( orc_pt* )tri_xy;
    if( cnt < 3 ) return 0;
    int V[ ORC_SLOTS ], nv = cnt, made = 0;
    float twice_area = 0.0f;
    for( int p = cnt - 1, q = 0; q < cnt; p = q++ ) twice_area += C[ p ].x * C[ q ].y - C[ q ].x * C[ p ].y;
    for( int v = 0; v < cnt; v++ ) V[ v ] = ( 0.0f < twice_area * 0.5f ) ? v : cnt - 1 - v;
    int guard = 2 * nv;
    for( int v = nv - 1; nv > 2; )
    {
        if( guard-- <= 0 ) return -made;
        int u = v;
        if( nv <= u ) u = 0;
        v = u + 1;
        if( nv <= v ) v = 0;
        int w = v + 1;
        if( nv <= w ) w = 0;
        orc_pt a = C[ V[ u ] ], b = C[ V[ v ] ], c = C[ V[ w ] ];
        int ear = !( 0.0000000001f > ( b.x - a.x ) * ( c.y - a.y ) - ( b.y - a.y ) * ( c.x - a.x ) );
        for( int p = 0; ear && p < nv; p++ )
        {
            if( p == u || p == v || p == w ) continue;
            if( orc_blocks_ear( a, b, c, C[ V[ p ] ] ) ) ear = 0;
        }
        if( !ear ) continue;
        T[ 3 * made ] = a;
        T[ 3 * made + 1 ] = b;
        T[ 3 * made + 2 ] = c;
        made++;
        for( int s = v, t = v + 1; t < nv; s++, t++ ) V[ s ] = V[ t ];
        nv--;
        guard = 2 * nv;
    }
    return made;
}

/* triangle list of the whole frame in the reference's slot layout: tri = N*45*2 floats, GLOBAL
 * coordinates (local + (i,j)); slots >= 3*ntri[n] are zero here (uninitialised in the reference) */
void orc_triangulate( const float* poly, const int* pcount, int W, int H, float* tri, int* ntri )
{
    int N = W * H;
    memset( tri, 0, sizeof( float ) * 2 * ORC_SLOTS * ( size_t )N );
    for( int n = 0; n < N; n++ )
    {
        float local[ 2 * ORC_SLOTS ];
        int k = orc_ear_clip( poly + ( size_t )n * 2 * ORC_SLOTS, pcount[ n ], local );
        ntri[ n ] = k;
        if( k < 0 ) k = -k;
        float* o = tri + ( size_t )n * 2 * ORC_SLOTS;
        for( int t = 0; t < 3 * k; t++ )
        {
            o[ 2 * t ] = local[ 2 * t ] + ( float )( n % W );
            o[ 2 * t + 1 ] = local[ 2 * t + 1 ] + ( float )( n / W );
        }
    }
}

/* ------------------------------------------------------------------------------------------- */
/* Connected components (NEW subsystem; the reference's cc_functions.cu is dead code).           */
/* label[n] = smallest row-major index in n's component of the given graph (its identity rule is  */
/* cc_functions.cu:394-413: a component is named by its first node in raster order).             */
/* ------------------------------------------------------------------------------------------- */
void orc_cc_labels( const uint8_t* graph, int W, int H, int32_t* label )
{
    int N = W * H;
    int32_t* stack = ( int32_t* )malloc( sizeof( int32_t ) * ( size_t )N );
    for( int n = 0; n < N; n++ ) label[ n ] = -1;
    for( int root = 0; root < N; root++ )
    {
        if( label[ root ] >= 0 ) continue;
        int top = 0;
        stack[ top++ ] = root;
        label[ root ] = root;
        while( top )
        {
            int n = stack[ --top ], i = n % W, j = n / W;
            for( int e = 0; e < 8; e++ )
            {
                if( !( graph[ n ] & ( 1u << e ) ) ) continue;
                int ni = i + ORC_DI[ e ], nj = j + ORC_DJ[ e ];
                if( ni < 0 || nj < 0 || ni >= W || nj >= H ) continue;
                int m = nj * W + ni;
                if( label[ m ] < 0 ) { label[ m ] = root; stack[ top++ ] = m; }
            }
        }
    }
    free( stack );
}

/* ------------------------------------------------------------------------------------------- */
/* Border walk of every connected component: the FIRST walk the reference's      */
/* extractBorderPoints (cc_functions.cu:348-503, dead code) makes for a component — the one that   */
/* starts at the component's first node in raster order, which is its label.  For that walk the   */
/* reference's serial bookkeeping (listed / discarded, :394-444) reduces to pure functions of the  */
/* graph: the walk is dropped when it steps on an interior node (== 90, :425) or comes back to its */
/* start node before the loop closes (the start is the only node of this component in `listed`,    */
/* :426), otherwise it is the node sequence below.  Later walks of the same component (the         */
/* reference starts one from every node not yet listed) depend on the scan state and are not       */
/* restated.  An island (node == 0) is undefined in the reference (getFirstLink returns -1 and     */
/* c_neighbor_index( ., -1, . ) falls off its switch, :215-246): here it is a walk of one node.    */
/* walk_len[n] = nodes in the walk that starts at n (0 for every other pixel and for dropped       */
/* walks); nodes = the walks concatenated in raster order of their start (capacity entries);       */
/* returns the number of entries written, or -1 when capacity is too small.                        */
/* ------------------------------------------------------------------------------------------- */
static const int ORC_CLOCK_TO_BIT[ 8 ] = { 0, 1, 2, 4, 7, 6, 5, 3 }; /* getRealLinkIndex, :68-99 */
static const int ORC_BIT_TO_CLOCK[ 8 ] = { 0, 1, 2, 7, 3, 6, 5, 4 }; /* getClockLinkIndex, :153-184 */
long orc_border_walks( const uint8_t* graph, const int32_t* label, int W, int H, int32_t* walk_len, int32_t* nodes, long capacity )
{
    const int N = W * H;
    long used = 0;
    for( int n = 0; n < N; n++ ) walk_len[ n ] = 0;
    for( int start = 0; start < N; start++ )
    {
        if( label[ start ] != start ) continue;
        const unsigned first = graph[ start ];
        if( first == 0u ) /* island */
        {
            if( used + 1 > capacity ) return -1;
            nodes[ used++ ] = start;
            walk_len[ start ] = 1;
            continue;
        }
        int edge = -1; /* getFirstLink, :107-118 */
        for( int c = 0; c < 8 && edge < 0; c++ )
            if( first & ( 1u << ORC_CLOCK_TO_BIT[ c ] ) ) edge = ORC_CLOCK_TO_BIT[ c ];
        int arrival = edge; /* nextEdgeCounterClockwise, :262-285 */
        if( orc_popcount8( first ) != 1 )
        {
            const int c0 = ORC_BIT_TO_CLOCK[ edge ];
            for( int c = c0 + 7; c >= c0 + 1; c-- )
                if( first & ( 1u << ORC_CLOCK_TO_BIT[ c % 8 ] ) ) { arrival = ORC_CLOCK_TO_BIT[ c % 8 ]; break; }
        }
        const long begin = used;
        int index = start, ok = 1;
        if( used + 1 > capacity ) return -1;
        nodes[ used++ ] = start;
        long guard = 0;
        while( index + ORC_DJ[ edge ] * W + ORC_DI[ edge ] != start || 7 - edge != arrival ) /* :415-416 */
        {
            /* nextNodeClockwise, :295-318 */
            index += ORC_DJ[ edge ] * W + ORC_DI[ edge ];
            const unsigned node = graph[ index ];
            if( orc_popcount8( node ) == 1 )
                edge = 7 - edge;
            else
            {
                const int c0 = ORC_BIT_TO_CLOCK[ 7 - edge ];
                for( int c = c0 + 1; c <= c0 + 7; c++ )
                    if( node & ( 1u << ORC_CLOCK_TO_BIT[ c % 8 ] ) ) { edge = ORC_CLOCK_TO_BIT[ c % 8 ]; break; }
            }
            if( node == 90u || index == start || ++guard > 16L * N ) { ok = 0; break; } /* :425-438 */
            if( used + 1 > capacity ) return -1;
            nodes[ used++ ] = index;
        }
        if( ok )
            walk_len[ start ] = ( int32_t )( used - begin );
        else
            used = begin;
    }
    return used;
}

/* ------------------------------------------------------------------------------------------- */
/* Direct rasterization (NEW; replaces glDrawArrays(GL_TRIANGLES), simpleVBO.cpp:281).            */
/* PARITY UNPINNED by the reference: the reference rasterizes inside the OpenGL driver.  Restated */
/* rule: output (s*W)x(s*H) RGBA8, row Y = pipeline row (0 = bottom); output    */
/* pixel (X,Y) samples the point ((X+1/2)/s, (Y+1/2)/s); cells are painted in node order, later   */
/* cells overwrite earlier ones (no depth test, main.cpp:261), colour = (byte2,byte1,byte0,255)   */
/* (kernel.cu:98-101), background (0,0,0,255) (main.cpp:260).  A sample exactly on an edge or      */
/* vertex is resolved as if it were displaced by (+eps, -eps^2): the top-left fill rule in the     */
/* displayed orientation.  All arithmetic is exact: vertices are multiples of 1/64, samples odd    */
/* multiples of 1/(2s); both are scaled to integers in units of 1/(128 s).                         */
/* ------------------------------------------------------------------------------------------- */
static long long orc_fix( float v, int s ) { return llround( ( double )v * 64.0 ) * 2 * s; }

/* 1 when every coordinate is a multiple of 1/64 (precondition of the exact rasterizers) */
int orc_all_dyadic64( const float* xy, long count )
{
    for( long t = 0; t < count; t++ )
    {
        double v = ( double )xy[ t ] * 64.0;
        if( v != floor( v ) ) return 0;
    }
    return 1;
}

static void orc_clear( uint8_t* out, long npix )
{
    for( long t = 0; t < npix; t++ ) { out[ 4 * t ] = 0; out[ 4 * t + 1 ] = 0; out[ 4 * t + 2 ] = 0; out[ 4 * t + 3 ] = 255; }
}

/* (a) faithful: paint the reference's TRIANGLE list, triangle by triangle, in slot order */
void orc_raster_triangles( const uint8_t* img, int W, int H, int ws, int s, const float* tri, const int* ntri, uint8_t* out )
{
    int OW = s * W, OH = s * H;
    orc_clear( out, ( long )OW * OH );
    for( int n = 0; n < W * H; n++ )
    {
        const uint8_t* col = img + ( size_t )( n / W ) * ws + 3 * ( n % W );
        int k = ntri[ n ] < 0 ? -ntri[ n ] : ntri[ n ];
        for( int t = 0; t < k; t++ )
        {
            const float* v = tri + ( ( size_t )n * ORC_SLOTS + 3 * t ) * 2;
            long long x[ 3 ], y[ 3 ];
            for( int c = 0; c < 3; c++ ) { x[ c ] = orc_fix( v[ 2 * c ], s ); y[ c ] = orc_fix( v[ 2 * c + 1 ], s ); }
            long long area = ( x[ 1 ] - x[ 0 ] ) * ( y[ 2 ] - y[ 0 ] ) - ( y[ 1 ] - y[ 0 ] ) * ( x[ 2 ] - x[ 0 ] );
            if( area == 0 ) continue;
            if( area < 0 ) { long long tx = x[ 1 ], ty = y[ 1 ]; x[ 1 ] = x[ 2 ]; y[ 1 ] = y[ 2 ]; x[ 2 ] = tx; y[ 2 ] = ty; }
            long long minx = x[ 0 ], maxx = x[ 0 ], miny = y[ 0 ], maxy = y[ 0 ];
            for( int c = 1; c < 3; c++ )
            {
                if( x[ c ] < minx ) minx = x[ c ];
                if( x[ c ] > maxx ) maxx = x[ c ];
                if( y[ c ] < miny ) miny = y[ c ];
                if( y[ c ] > maxy ) maxy = y[ c ];
            }
            /* sample X sits at (2X+1)*64 */
            long long X0 = ( minx - 64 ) / 128 - 1, X1 = ( maxx - 64 ) / 128 + 1;
            long long Y0 = ( miny - 64 ) / 128 - 1, Y1 = ( maxy - 64 ) / 128 + 1;
            if( X0 < 0 ) X0 = 0;
            if( Y0 < 0 ) Y0 = 0;
            if( X1 > OW - 1 ) X1 = OW - 1;
            if( Y1 > OH - 1 ) Y1 = OH - 1;
            for( long long Y = Y0; Y <= Y1; Y++ )
                for( long long X = X0; X <= X1; X++ )
                {
                    long long px = ( 2 * X + 1 ) * 64, py = ( 2 * Y + 1 ) * 64;
                    int in = 1;
                    for( int c = 0; c < 3 && in; c++ )
                    {
                        int d = ( c + 1 ) % 3;
                        long long ex = x[ d ] - x[ c ], ey = y[ d ] - y[ c ];
                        long long E = ex * ( py - y[ c ] ) - ey * ( px - x[ c ] );
                        if( E < 0 ) in = 0;
                        else if( E == 0 ) in = ( ey < 0 ) || ( ey == 0 && ex < 0 ); /* left edge, or top edge */
                    }
                    if( in )
                    {
                        uint8_t* o = out + 4 * ( ( size_t )Y * OW + X );
                        o[ 0 ] = col[ 2 ]; o[ 1 ] = col[ 1 ]; o[ 2 ] = col[ 0 ]; o[ 3 ] = 255;
                    }
                }
        }
    }
}

/* (b) polygon form of the same rule (what the CUDA rasterizer implements): even-odd crossing
 * count of the displaced sample against the polygon outline.  Equal to (a) whenever the ear
 * clipping partitions the polygon, which tests/ check on every frame they use. */
void orc_raster_polygons( const uint8_t* img, int W, int H, int ws, int s, const float* poly, const int* pcount, uint8_t* out )
{
    int OW = s * W, OH = s * H;
    orc_clear( out, ( long )OW * OH );
    for( int n = 0; n < W * H; n++ )
    {
        const uint8_t* col = img + ( size_t )( n / W ) * ws + 3 * ( n % W );
        const float* v = poly + ( size_t )n * 2 * ORC_SLOTS;
        int cnt = pcount[ n ], i = n % W, j = n / W;
        long long x[ ORC_SLOTS ], y[ ORC_SLOTS ];
        for( int c = 0; c < cnt; c++ ) { x[ c ] = orc_fix( v[ 2 * c ], s ); y[ c ] = orc_fix( v[ 2 * c + 1 ], s ); }
        int h = ( s + 2 ) / 4 + 1;
        for( int Y = j * s - h; Y < ( j + 1 ) * s + h; Y++ )
            for( int X = i * s - h; X < ( i + 1 ) * s + h; X++ )
            {
                if( X < 0 || Y < 0 || X >= OW || Y >= OH ) continue;
                long long px = ( 2LL * ( X - i * s ) + 1 ) * 64, py = ( 2LL * ( Y - j * s ) + 1 ) * 64;
                int in = 0;
                for( int c = 0; c < cnt; c++ )
                {
                    int d = ( c + 1 ) % cnt;
                    if( ( y[ c ] < py ) == ( y[ d ] < py ) ) continue;
                    long long D = ( x[ d ] - x[ c ] ) * ( py - y[ c ] ) - ( px - x[ c ] ) * ( y[ d ] - y[ c ] );
                    if( ( y[ d ] > y[ c ] ) ? ( D > 0 ) : ( D < 0 ) ) in ^= 1;
                }
                if( in )
                {
                    uint8_t* o = out + 4 * ( ( size_t )Y * OW + X );
                    o[ 0 ] = col[ 2 ]; o[ 1 ] = col[ 1 ]; o[ 2 ] = col[ 0 ]; o[ 3 ] = 255;
                }
            }
    }
}

/* ------------------------------------------------------------------------------------------- */
/* Whole path, kernel.cu:402-475 order.  Any output pointer may be NULL.                         */
/* ------------------------------------------------------------------------------------------- */
int orc_pipeline( const uint8_t* img, int W, int H, int ws, int subdivide, int fused_y, int scale,
                  uint8_t* graph_aux_out, uint8_t* graph_out, int32_t* labels_out,
                  float* hull_out, int* hull_count_out, float* poly_out, int* poly_count_out,
                  float* tri_out, int* ntri_out, uint8_t* raster_out )
{
    size_t N = ( size_t )W * H;
    uint8_t* g0 = ( uint8_t* )malloc( N );
    uint8_t* aux = ( uint8_t* )malloc( N );
    uint8_t* g = ( uint8_t* )malloc( N );
    orc_similarity_graph( img, W, H, ws, fused_y, g0 );
    orc_trivial_crossings( g0, W, H, aux );
    orc_resolve_crossings( aux, W, H, g, NULL );
    if( graph_aux_out ) memcpy( graph_aux_out, aux, N );
    if( graph_out ) memcpy( graph_out, g, N );
    if( labels_out ) orc_cc_labels( g, W, H, labels_out );
    if( hull_out || hull_count_out || poly_out || poly_count_out || tri_out || ntri_out || raster_out )
    {
        float* hull = ( float* )malloc( sizeof( float ) * 2 * ORC_SLOTS * N );
        float* poly = ( float* )malloc( sizeof( float ) * 2 * ORC_SLOTS * N );
        int* hc = ( int* )malloc( sizeof( int ) * N );
        int* pc = ( int* )malloc( sizeof( int ) * N );
        orc_cells( g, W, H, hull, hc );
        if( subdivide )
            orc_subdivide( img, W, H, ws, g, hull, hc, poly, pc );
        else
        {
            memcpy( poly, hull, sizeof( float ) * 2 * ORC_SLOTS * N );
            memcpy( pc, hc, sizeof( int ) * N );
        }
        if( hull_out ) memcpy( hull_out, hull, sizeof( float ) * 2 * ORC_SLOTS * N );
        if( hull_count_out ) memcpy( hull_count_out, hc, sizeof( int ) * N );
        if( poly_out ) memcpy( poly_out, poly, sizeof( float ) * 2 * ORC_SLOTS * N );
        if( poly_count_out ) memcpy( poly_count_out, pc, sizeof( int ) * N );
        if( tri_out || ntri_out || raster_out )
        {
            float* tri = ( float* )malloc( sizeof( float ) * 2 * ORC_SLOTS * N );
            int* nt = ( int* )malloc( sizeof( int ) * N );
            orc_triangulate( poly, pc, W, H, tri, nt );
            if( raster_out ) orc_raster_triangles( img, W, H, ws, scale, tri, nt, raster_out );
            if( tri_out ) memcpy( tri_out, tri, sizeof( float ) * 2 * ORC_SLOTS * N );
            if( ntri_out ) memcpy( ntri_out, nt, sizeof( int ) * N );
            free( tri );
            free( nt );
        }
        free( hull ); free( poly ); free( hc ); free( pc );
    }
    free( g0 ); free( aux ); free( g );
    return 0;
}
