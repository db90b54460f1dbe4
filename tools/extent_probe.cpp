// development probe: how far outside its own pixel square can a subdivided cell reach?
// Enumeration over own key x neighbour key, with increasing consistency constraints between the two.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define __restrict__
#include "../pixel_art_remaster_gpu_b200/csrc/polygon.cuh"
using namespace par;
static CellTables T;
// neighbour key may depend on the link direction asked for: we enumerate per (key, L) separately
struct Env { uint32_t nb; int wantL; mutable bool used; uint32_t key(int i, int j) const { used = true; return nb; } bool guard(int, int) const { return false; } bool keep_corner(int, int, Q2) const { return false; } };
struct Sink { int lo, hi; void put(int, int x, int y) { if (x < lo) lo = x; if (y < lo) lo = y; if (x > hi) hi = x; if (y > hi) hi = y; } };
int main()
{
    build_cell_tables(&T);
    for (int level = 0; level < 2; level++) {
        int lo = 0, hi = 64; uint32_t klo = 0, khi = 0, nlo = 0, nhi = 0;
        for (uint32_t key = 0; key < 4096; key++) {
            // which links does this cell consult?  (every non-border edge id)
            uint32_t links = (uint32_t)T.rec[key].info; int n = hull_count(T.rec[key].info);
            for (int L = 0; L < 8; L++) {
                bool has = false; for (int t = 0; t < n; t++) if (((links >> (4 * t)) & 15u) == (uint32_t)L) has = true;
                if (!has) continue;
                for (uint32_t nb = 0; nb < 4096; nb++) {
                    if (level >= 1 && !((nb >> (7 - L)) & 1u)) continue; // neighbour must hold the reciprocal link
                    Env e{nb, L, false}; Sink s{0, 64};
                    build_cell_polygon(e, CellTablePtrs{T.rec}, 1, 1, key, true, s);
                    if (s.lo < lo) { lo = s.lo; klo = key; nlo = nb; }
                    if (s.hi > hi) { hi = s.hi; khi = key; nhi = nb; }
                }
            }
        }
        printf("level %d: min coord %d/64 (key %u nb %u)  max coord %d/64 (key %u nb %u)\n", level, lo, klo, nlo, hi, khi, nhi);
    }
    return 0;
}
