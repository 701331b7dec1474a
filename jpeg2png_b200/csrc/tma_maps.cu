// tma_maps.cu — host side of the TMA path: cuTensorMapEncodeTiled through the runtime's driver entry
// point (the library does not link libcuda).
#include "tma_maps.h"

#include <cuda_runtime.h>

#include <mutex>

namespace j2p {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    });
    return fn;
}

int encode_plane_map(CUtensorMap *out, const float *base, int W, int rows, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return -1;
    const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)W * sizeof(float)};          // bytes between rows; W is a multiple of 8 => of 16 bytes
    const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};                   // 128 bytes x box_rows: one 128-byte-swizzle atom per row
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult rc = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return rc == CUDA_SUCCESS ? 0 : -1;
}

}  // namespace j2p
