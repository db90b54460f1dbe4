"""Host-side smoothing tables of the raster kernel (csrc/smooth_table.h): structural invariants the kernel relies on, checked
without a GPU through a small host program (the tables' semantics — pieces XOR to the polygon's coverage, IDs name the
neighbour records, shared ID ranges say exactly which records fit a class — are checked by tests/test_host_polygons.py)."""
import os
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "pixel_art_remaster_gpu_b200", "csrc")

PROGRAM = r"""
#include "smooth_table.h"
#include <stdio.h>
using namespace par;
int main()
{
    static CellTables ct;
    static SmoothTables st;
    build_cell_tables( &ct );
    build_smooth_tables( ct, &st );
    int exact = 0, with_canon = 0, bad_desc = 0, max_id = 0, slow = 0;
    for( const LinkClass& c : st.classes )
    {
        exact += c.exact ? 1 : 0;
        with_canon += c.canon ? 1 : 0;
        if( c.canon && ( c.canon <= st.classes.size() || c.canon > st.classes.size() + st.n_canon ) ) bad_desc++;
    }
    for( int key = 0; key < kCellKeys; key++ )
    {
        if( st.desc[ key ][ 0 ] & kDescSlow ) slow++;
        for( int e = 0; e < 8; e++ ) max_id = st.nbr_id[ key ][ e ] > max_id ? st.nbr_id[ key ][ e ] : max_id;
        for( int k = 0; k < 4; k++ )
        {
            const uint32_t d = st.desc[ key ][ k ];
            if( !( d & kDescUsed ) ) continue;
            const uint32_t woff = d & 255u, shift = ( d >> 8 ) & 31u, block = ( d >> 13 ) & 255u;
            // the word offset stays inside a 3-row neighbourhood of cell words, the 5-bit field inside its word, the block is a class's
            if( woff > 2u * kHeadRowWords + 2u || shift > 27u || block == 0u || block > st.classes.size() ) bad_desc++;
            if( ( d >> 22 ) & 0x7Fu ) bad_desc++; // bits 22..28 are free (29: slow, 30: never set, 31: more — slot 0 only)
            if( k > 0 && ( d & ( kDescSlow | kDescMore ) ) ) bad_desc++;
        }
        if( ( ( st.desc[ key ][ 0 ] & kDescMore ) != 0u ) != ( ( st.desc[ key ][ 2 ] & kDescUsed ) != 0u ) ) bad_desc++;
    }
    printf( "%zu %d %d %u %u %d %d %d\n", st.classes.size(), exact, with_canon, st.n_canon, st.link_entries, bad_desc, max_id, slow );
    return 0;
}
"""


@pytest.fixture(scope="module")
def table_stats(tmp_path_factory):
    d = tmp_path_factory.mktemp("smooth_tables")
    src = d / "tables.cpp"
    src.write_text(PROGRAM)
    exe = d / "tables"
    subprocess.run(["g++", "-O1", "-I", CSRC, str(src), os.path.join(CSRC, "cell_table.cpp"), os.path.join(CSRC, "smooth_table.cpp"), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    keys = ("classes", "exact", "with_canon", "n_canon", "link_entries", "bad_desc", "max_id", "slow")
    return dict(zip(keys, map(int, out)))


def test_descriptors_are_well_formed(table_stats):
    assert table_stats["bad_desc"] == 0
    assert table_stats["slow"] == 0                     # every key is expressible by the tables


def test_ids_fit_five_bits_and_ranges_are_exact(table_stats):
    assert table_stats["classes"] == 188
    assert 1 <= table_stats["max_id"] <= 30             # 0 = no edge in that direction; 31 entries per block hold every ID
    assert table_stats["exact"] == table_stats["classes"]   # the fitting IDs of every class are one range ...
    assert table_stats["with_canon"] == table_stats["classes"]  # ... and every class has a shared block for it
    assert 1 <= table_stats["n_canon"] <= 16


def test_link_table_has_a_block_per_class_and_range(table_stats):
    blocks = table_stats["link_entries"] // 32
    assert table_stats["link_entries"] % 32 == 0
    assert blocks == 1 + table_stats["classes"] + table_stats["n_canon"] <= 255   # block numbers are 8 bits in a descriptor
