// Host check of the window-stack work split (transformerupscaler_b200/csrc/tc/stack_split.cuh, compiled as plain C++):
// every (tile, block) unit is run exactly once; a tile is cut at most once; its leading blocks are the FIRST segment of a CTA
// and its remaining blocks the LAST segment of the NEXT CTA (so a wait only ever points at a CTA dispatched earlier, at work
// that CTA does before anything else); shares differ by at most the rounding of units_per_cta.
#include <cstdio>
#include <vector>

#include "../../transformerupscaler_b200/csrc/tc/stack_split.cuh"

static int check(int n_tiles, int nb, int sms) {
    const int grid = n_tiles < sms ? n_tiles : sms;
    const bool split = n_tiles % grid != 0;
    const int upc = (n_tiles * nb + grid - 1) / grid;
    std::vector<int> owner(n_tiles * nb, -1), first_cta(n_tiles, -1), cont_cta(n_tiles, -1);
    for (int c = 0; c < grid; ++c) {
        tu::Seg s;
        int nseg = 0, units = 0;
        while (tu::seg_of(c, grid, n_tiles, nb, split, upc, nseg, s)) ++nseg;
        for (int k = 0; k < nseg; ++k) {
            tu::seg_of(c, grid, n_tiles, nb, split, upc, k, s);
            if (s.tile < 0 || s.tile >= n_tiles || s.lo < 0 || s.hi > nb || s.lo >= s.hi) return 1;
            for (int b = s.lo; b < s.hi; ++b) {
                if (owner[s.tile * nb + b] != -1) return 2;        // a unit run twice
                owner[s.tile * nb + b] = c;
            }
            units += s.hi - s.lo;
            if (s.lo == 0 && s.hi < nb) {                           // leading part: must be this CTA's first segment
                if (k != 0) return 3;
                first_cta[s.tile] = c;
            }
            if (s.lo > 0) {                                         // continuation: must be this CTA's last segment, and run to the end
                if (k != nseg - 1 || s.hi != nb) return 4;
                cont_cta[s.tile] = c;
            }
        }
        if (split && units > upc) return 5;
    }
    for (int u = 0; u < n_tiles * nb; ++u)
        if (owner[u] == -1) return 6;                               // a unit never run
    for (int t = 0; t < n_tiles; ++t) {
        if ((first_cta[t] == -1) != (cont_cta[t] == -1)) return 7;
        if (first_cta[t] != -1 && cont_cta[t] != first_cta[t] + 1) return 8;
    }
    return 0;
}

int main() {
    const int cases[][3] = {{240, 8, 148}, {150, 8, 148}, {480, 6, 148}, {149, 8, 148}, {295, 8, 148}, {296, 8, 148}, {1000, 8, 148},
                            {3, 8, 148}, {148, 8, 148}, {240, 1, 148}, {241, 3, 7}, {17, 5, 4}};
    for (const auto &c : cases) {
        const int rc = check(c[0], c[1], c[2]);
        if (rc) { std::printf("FAIL n_tiles=%d n_blocks=%d sms=%d: rule %d\n", c[0], c[1], c[2], rc); return 1; }
    }
    std::printf("OK %zu cases\n", sizeof(cases) / sizeof(cases[0]));
    return 0;
}
