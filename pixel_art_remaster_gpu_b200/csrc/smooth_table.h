// Stage E as tables: the coverage mask of a smoothed cell, assembled from precomputed pieces.
//
// Coverage under the even-odd rule is XOR-linear in the polygon's edges, so the mask of the corner-cut
// polygon (subdivision_functions.cu:564-669) splits into pieces that each depend on very little:
//
//     mask(cell) = CUT[key][kept]  ^  XOR over the cell's "link pieces"  LINK[class][a][b]
//
// * CUT[key][kept] : the hull with every "cut" vertex (both adjacent edges border, :583-598) replaced by
//   R(prev), Q(cur), except square corners whose bit in `kept` says checkTJunction keeps them
//   (:170-242).  Blended vertices stay at their hull position here.
// * a link piece belongs to one SHARED hull edge t that has a blended vertex at one or both ends
//   (:603-647).  It is the closed loop  S -> [blend points] -> T -> back along the hull to S, i.e. the
//   difference between the hull path and the smoothed path around that edge.  Besides the cell's own hull
//   (the "class": a pure function of key and t) it depends only on ONE neighbour hull vertex per blended end:
//   the vertex after (end A, neighbour's Q, :141-154) / before (end B, neighbour's R, :125-138) the shared
//   vertex in the neighbour across the edge, given as a 4-bit point code a / b.
//
// The neighbour side is 16 bits per (neighbour key, link direction): NBR[key][e'] describes the neighbour's hull
// edge that is shared through ITS graph edge e' = 7 - e (a hull has at most one per direction):
//     bits [0,4)  a = point code of the vertex AFTER the edge's end      (what end A needs)
//     bits [4,8)  b = point code of the vertex BEFORE the edge's start   (what end B needs)
//     bits [8,12) point code of the edge's end, bits [12,16) of its start; start == end: no such edge.
// The cell's blended vertex must BE that end (A) / start (B) — the reference finds it by coordinates
// (getPointIndex, :527-538) and then takes the vertex after / before it, which is the same thing whenever it
// is there.  When it is not (codes differ, or no such edge: getPointIndex's "not found -> 0" fallback and a few
// consistent-but-unusual hull pairs) the cell takes the exact geometric path instead.
//
// Everything here is content-independent and built once per context (the masks per scale, on the device,
// by the same coverage code the geometric path uses).
#pragma once
#include "cell_table.h"
#include <vector>

namespace par {

// Link descriptor (32 bits), up to kMaxLinks per key; 0 = unused slot.
//   [0,3)   e     graph edge the hull edge is shared through = direction of the neighbour
//   [3]     hasA  the vertex the edge starts at is blended (edge before it is a border edge)
//   [4]     hasB  the vertex the edge ends at is blended (edge after it is a border edge)
//   [5,9)   codeA the start vertex as a point code in the neighbour's frame
//   [9,13)  codeB the end vertex as a point code in the neighbour's frame
//   [13,32) first entry of the class in the link table; entry = first + (hasA && hasB ? a + 16 b : hasA ? a : b)
constexpr int kMaxLinks = 4;
constexpr uint32_t kSmoothSlow = 0xFFFFFFFFu; // in link[0]: this key always takes the geometric path

struct SmoothRecord
{
    uint32_t link[ kMaxLinks ];
};

// geometry of one class, consumed by the device table builder (quarter-pixel units)
struct LinkClass
{
    int8_t e, hasA, hasB, pad;
    int8_t px[ 4 ], py[ 4 ]; // hull vertices t-1, t, t+1, t+2
    uint32_t first;          // first entry in the link table
    uint32_t count;          // 16 or 256
};

struct SmoothTables
{
    SmoothRecord rec[ kCellKeys ];
    uint16_t nbr[ kCellKeys ][ 8 ];
    std::vector< LinkClass > classes;
    uint32_t link_entries = 0; // total entries of the link table
    uint32_t slow_keys = 0;    // keys that always take the geometric path
};

void build_smooth_tables( const CellTables& cells, SmoothTables* out );

} // namespace par
