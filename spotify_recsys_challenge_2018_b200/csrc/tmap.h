// Host side of TMA: tensor maps for the row-major bf16 operands (shared by the tensor-core translation units).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

namespace dae {

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

static inline PFN_tmapEncodeTiled tmap_encoder() {
    static PFN_tmapEncodeTiled fn = []() -> PFN_tmapEncodeTiled {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess || p == nullptr) {
            fprintf(stderr, "dae_b200: cuTensorMapEncodeTiled not available from the driver\n");
            abort();
        }
        return reinterpret_cast<PFN_tmapEncodeTiled>(p);
    }();
    return fn;
}

// Row-major bf16 matrix [outer, inner]; box = [box_outer rows, 64 elements (128 B)], SWIZZLE_128B,
// out-of-bounds elements read as zero.
// `ld` = row stride in elements (0: rows are dense, ld = inner).
static inline CUtensorMap make_map_bf16(const void* ptr, uint64_t inner, uint64_t outer, uint32_t box_outer,
                                        uint64_t ld = 0) {
    CUtensorMap m;
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {(ld ? ld : inner) * sizeof(__nv_bfloat16)};
    cuuint32_t box[2] = {64, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = tmap_encoder()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box,
                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "dae_b200: cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu box=%u\n", (int)r,
                (unsigned long long)inner, (unsigned long long)outer, box_outer);
        abort();
    }
    return m;
}

}  // namespace dae
