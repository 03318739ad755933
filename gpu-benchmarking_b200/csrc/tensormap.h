// tensormap.h -- host side of the tiled-TMA (cp.async.bulk.tensor) gather of the interleaved layout
//     x[(e/32)*32*len + 32*idx + e%32]
// seen as a rank-3 tensor {32 elements of a group (innermost), len indices, groups}: a box of {EL, rows, 1} is the
// [idx][e] tile the coa-pipe kernel (sumfac_coapipe.cuh) wants in shared memory, dense, in one instruction per <= 256
// rows.  cuTensorMapEncodeTiled is a driver entry point; it is looked up through the runtime
// (cudaGetDriverEntryPoint) so that the library keeps linking against cudart only and still loads on a machine without
// a driver (this container): the lookup happens at the first call that needs a map.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

namespace b200fe
{

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn tensor_map_encoder()
{
    static EncodeTiledFn fn = [] {
        void *p                                 = nullptr;
        cudaDriverEntryPointQueryResult status = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &status) != cudaSuccess ||
            status != cudaDriverEntryPointSuccess)
        {
            (void)cudaGetLastError();
            p = nullptr;
        }
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// map of an interleaved array of `ngroups` groups of 32 elements with `len` values each; box = {el, rows, 1}.
// false when the driver entry point is missing or the encode fails (the caller then takes the cp.async gather)
template <typename T>
inline bool make_coa_tensor_map(CUtensorMap *map, const T *base, unsigned len, unsigned ngroups, unsigned el, unsigned rows)
{
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc)
        return false;
    const cuuint64_t dims[3]    = {32, len, ngroups};
    const cuuint64_t strides[2] = {32 * sizeof(T), (cuuint64_t)32 * len * sizeof(T)}; // bytes, dimensions 1 and 2
    const cuuint32_t box[3]     = {el, rows, 1};
    const cuuint32_t estr[3]    = {1, 1, 1};
    const CUtensorMapDataType dt = sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    return enc(map, dt, 3, const_cast<T *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

} // namespace b200fe
