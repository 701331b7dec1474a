// tma_maps.h — the tensor maps (TMA descriptors) of a session's plane buffers.
#pragma once
#include <cuda.h>

namespace j2p {

// Per plane: [0] / [1] the two iterate buffers (x_k and x_{k-1} swap roles every iteration), [2] the
// gradient, [3] the DCT-distance term gp (coefficient-grid geometry; 1x1 planes only, written by the
// projection's TMA stores).  Every map describes its buffer as a 2-D fp32 tensor rows x pitch, box
// 32 x 8 (or 16) elements, 128-byte swizzle.  Passed to the kernels by value (__grid_constant__),
// 64-byte aligned.
struct alignas(64) TileMaps {
    CUtensorMap m[3][4];
};

// 0 on success.  `base` is local row 0 of the buffer, `rows` the rows it holds (strip sessions: owned
// rows plus halo rows), W the row pitch in elements; box_rows = 8 (1x1 planes) or 16 (2x2 planes).
int encode_plane_map(CUtensorMap *out, const float *base, int W, int rows, int box_rows);

}  // namespace j2p
