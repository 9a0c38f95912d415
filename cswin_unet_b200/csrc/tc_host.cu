// tc_host.cu — host-side TMA tensor-map encoding through the driver entry point (the library does not link libcuda).
#include <mutex>

#include "common.cuh"
#include "tc_common.cuh"

namespace cswin {
namespace tc {

EncodeTiledFn encode_tiled_fn() {
  static std::once_flag once;
  static EncodeTiledFn fn = nullptr;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

bool make_tensor_map_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle,
                          const uint32_t* elem_strides) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return false; }
  cuuint64_t gdim[5]; cuuint64_t gstr[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = elem_strides ? elem_strides[i] : 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return false; }
  return true;
}

}  // namespace tc
}  // namespace cswin
