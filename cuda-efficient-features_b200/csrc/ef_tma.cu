// ef_tma.cu -- host side of ef_tma.cuh: tensor-map encoding through the driver entry point (no link-time dependency on libcuda).
#include "ef_tma.cuh"

#include <cstring>

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
} // namespace

bool ef_tma_encode_u8(CUtensorMap* map, const void* base, int w, int h, int nframes, size_t pitch, size_t frame_stride, int box_w, int box_h)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn || !base || w <= 0 || h <= 0 || nframes <= 0) return false;
    if (frame_stride == 0 || nframes == 1) frame_stride = (pitch * (size_t)h + 15) & ~(size_t)15;   // one frame: any legal stride
    // TMA rules: 16-byte aligned base, strides multiples of 16 bytes, inner box extent a multiple of 16 bytes, box edges <= 256
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (pitch & 15) || (frame_stride & 15) || (box_w & 15) || box_w > 256 || box_h > 256) return false;
    const cuuint64_t dims[3] = { (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)nframes };
    const cuuint64_t strides[2] = { (cuuint64_t)pitch, (cuuint64_t)frame_stride };
    const cuuint32_t box[3] = { (cuuint32_t)box_w, (cuuint32_t)box_h, 1u };
    const cuuint32_t estr[3] = { 1u, 1u, 1u };
    std::memset(map, 0, sizeof(*map));
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
