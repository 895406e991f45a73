// Host build of pegasus_b200/csrc/tile_cull.h for the CPU tests (same IEEE operations as the kernel).
#include "../../pegasus_b200/csrc/tile_cull.h"

extern "C" {
// runs[(ty - ry0) * 2 + {0,1}] = ta, tb for every tile row of the rectangle; returns the pair count
int cull_runs(float gx, float gy, float qa, float qb, float qc, float cut, int W, int H, int rx0, int ry0,
              int rx1, int ry1, int* runs) {
    pg::CullGauss c = pg::cull_setup(gx, gy, qa, qb, qc, cut);
    int total = 0;
    for (int ty = ry0; ty < ry1; ++ty) {
        int ta, tb;
        total += pg::cull_row_run(c, ty, rx0, rx1, W, H, &ta, &tb);
        runs[(ty - ry0) * 2] = ta;
        runs[(ty - ry0) * 2 + 1] = tb;
    }
    return total;
}
}
