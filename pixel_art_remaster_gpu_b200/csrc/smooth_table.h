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
//     bits [0,4)  a = which vertex comes AFTER the edge's end      (what end A needs)
//     bits [4,8)  b = which vertex comes BEFORE the edge's start   (what end B needs)
//       both as a rank 0..3: over all 4096 hulls only four different points occur in either role for a given
//       direction, so a class's masks are 4 x 4 entries and each a-row is one 32-byte sector
//     bits [8,12) point code of the edge's end, bits [12,16) of its start.  A direction without such an edge
//     holds a code pair that no cell ever expects for that direction, so it can never compare equal.
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

// Link descriptor (32 bits), up to kMaxLinks per key, laid out so that the kernel decodes it with a few shifts:
//   [0,3)   e       graph edge the hull edge is shared through = direction of the neighbour
//   [4,8)   corners (slot 0 only) which square corners (0,0) (1,0) (1,1) (0,1) hold a cut vertex of the hull: the
//                   `kept` bits of the others do not matter and are masked off before CUT is indexed
//   [8,16)  expect  codeA | codeB << 4: the edge's start / end vertex as point codes in the neighbour's frame,
//                   to be compared with bits [8,16) of the neighbour record
//   [16,24) ends    0x0F: the start vertex is blended (end A), 0xF0: the end vertex is (end B), 0xFF: both.
//                   Masks the comparison above AND selects the ranks: entry = 16 block + (rank a | rank b << 2)
//   [24,32) block   the class's 16-entry block of the link table: 4 x 4 ranks, one 128-byte line at s <= 4 (block 0 is all zero)
// An unused slot is 0: it compares nothing, selects entry 0 of block 0 and so contributes an empty mask.
constexpr int kMaxLinks = 4;
constexpr uint32_t kSmoothSlow = 0xFFFFFFFFu; // in link[0]: this key always takes the geometric path

// One 32-byte record per key (one sector): the key's own link descriptors and, for its neighbours, its records.
struct SmoothRecord
{
    uint32_t link[ kMaxLinks ];
    uint16_t nbr[ 8 ];
};

// What the raster kernel reads (round 3: the records above are only the host-side source of these).
//
// * DESC[key][4] (32 bits per link descriptor): the descriptor reduced to what the kernel still needs, ready to use —
//       [0,8)   where the neighbour's ID sits: word offset from the cell's own word in the staged tile (kHeadRowWords words
//               per tile row, see PACK below) to the neighbour's x-word (e < 4) / y-word (e >= 4), biased by
//               kHeadRowWords + 1 (so that it is not negative)
//       [8,13)  shift of the ID field for direction 7 - e in that word
//       [13,21) the class's block of the link table   [21] used
//       slot 0 only: [31] the key has a third descriptor, [29] the key always takes the geometric path.
//   Per SCALE the kernel reads a copy of this (built on the device next to the link table, raster_kernels.cu): at 4x more
//   than half of the classes have an empty mask for every ID that fits them — the loop between the hull and the smoothed
//   outline around a blended vertex is a sliver that often holds no sample — so all their block says is which IDs do NOT
//   fit (MISMATCH).  The IDs of a direction are numbered in the order of their records, codes in the high bits, so the
//   fitting ones are a range lo .. lo + span, and classes with the same range share ONE block (LinkClass::canon): three
//   quarters of the lookups of the bench frames then fall on a dozen hot blocks.  The descriptors whose class keeps its
//   own block come first.
// * a neighbour record is one of at most 30 different 16-bit values per direction, so it is named by a 5-bit ID
//   (1..30; 0 = the neighbour has no hull edge shared through that direction).  NBR_ID[key][e'] is that ID, and
//   PACK[key] = { IDs of directions 4..7 in bits [12 + 5 (e' - 4)), IDs of directions 0..3 in bits [5 e') } is what the
//   raster kernel keeps per staged cell (the low word ORed onto the 12-bit key): a cell that blends across its link e finds
//   the ID of its neighbour's record for direction 7 - e in the shared-memory word it would read the neighbour's key from —
//   no neighbour record is gathered from global memory, and nothing depends on how many links the neighbour has.
// * the link table has 32 entries per class, indexed by the neighbour's ID: entry = the loop mask for the ranks the ID's
//   record gives, or the MISMATCH flag when the record's end / start vertex is not the class's blended vertex (the
//   comparison the kernel used to make per cell is made once, when the table is built).
constexpr int kNbrIds = 32;
constexpr int kHeadRowWords = 72; // words per row of staged cell words: 36 x-words, then 36 y-words (raster_impl.cuh Cfg::KP)
// (kDescSlow sits where MISMATCH sits in the high word of a window-form entry, kDescMore in the sign bit: the kernel folds
// the first into its flag test with one logic operation and tests the second with one comparison)
constexpr uint32_t kDescSlow = 1u << 29, kDescMore = 1u << 31, kDescUsed = 1u << 21;

// geometry of one class, consumed by the device table builder (quarter-pixel units)
struct LinkClass
{
    int8_t e, hasA, hasB, pad;
    int8_t px[ 4 ], py[ 4 ]; // hull vertices t-1, t, t+1, t+2
    uint32_t block;          // its 16-entry block of the link table (>= 1)
    int8_t after[ 4 ], before[ 4 ]; // rank -> point code of the neighbour's vertex after the edge end / before its start
    uint8_t codeA, codeB;           // the blended vertices as point codes in the neighbour's frame (what its record must hold)
    uint8_t id_lo, id_span;         // the IDs whose record holds them: lo .. lo + span (0, 31 and exact = 0 when they are not one range)
    uint16_t nrec[ kNbrIds ];       // ID -> the neighbour's 16-bit record for direction 7 - e (0xFFFF: no such ID)
    uint16_t exact;                 // the ID range says exactly which records fit
    uint16_t canon;                 // the block (after the classes' own) shared by the classes with this range
};

struct SmoothTables
{
    SmoothRecord rec[ kCellKeys ];
    uint32_t desc[ kCellKeys ][ 4 ], pack[ kCellKeys ][ 2 ];
    uint8_t nbr_id[ kCellKeys ][ 8 ];
    std::vector< LinkClass > classes;
    uint32_t link_entries = 0; // total entries of the link table (kNbrIds per class + the zero block + the shared range blocks)
    uint32_t n_canon = 0;      // shared range blocks
    uint8_t canon_lo[ 64 ], canon_span[ 64 ];
    uint32_t slow_keys = 0;    // keys that always take the geometric path
};

void build_smooth_tables( const CellTables& cells, SmoothTables* out );

} // namespace par
