// Host side of the TMA paths: tensor maps through the runtime's driver entry point (cuTensorMapEncodeTiled) - no link-time
// libcuda dependency, so the library still loads on a box without a driver and fails with PGN_E_CUDA there.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

// bf16 [layers][rows][cols] view (row pitch ld elements, layer pitch layer_stride_elems), boxes of 64 columns x box_rows
// rows, SWIZZLE_128B: a box is box_rows rows of 128 bytes with the 16-byte chunks XOR-swizzled by the row - the canonical
// K-major (rows = M/N) or MN-major (rows = K) SWIZZLE_128B UMMA operand.  Out-of-range rows / columns are zero-filled on
// loads and clipped on stores.
static inline cudaError_t pgn_make_map_bf16(CUtensorMap* map, const void* base, long long cols, long long ld, long long rows,
                                            long long layers, long long layer_stride_elems, int box_rows) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess) return e;
    if (!fn || q != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
    encode = (EncodeTiledFn)fn;
  }
  const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)layers};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(layers > 1 ? layer_stride_elems : ld * rows) * 2};
  const cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}
