// ef_tma.cuh -- Tensor Memory Accelerator plumbing of the image kernels (sm_100a): tensor maps over the pitched u8 pyramid levels
// of a batch (dimensions x, y, frame), encoded on the host per call (cuTensorMapEncodeTiled through cudaGetDriverEntryPoint: no link
// against libcuda) and passed to the kernels as a __grid_constant__ parameter; on the device one elected thread issues
// cp.async.bulk.tensor.3d (SASS: UTMALDG) onto an mbarrier and the CTA waits on its phase.  Out-of-bounds box elements are zero
// filled by the hardware -- the kernels keep their own REFLECT_101 / clamp handling for the tiles that touch an image edge.
// Measured on the B200 (tools/probe/tma_probe.cu): the innermost start coordinate times the element size must be a multiple of
// 16 bytes (negative is fine); an unaligned one raises an illegal-instruction trap, it is not rounded.
#pragma once

#include <cuda.h>          // CUtensorMap (types and enums only)
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ef_b200.h"

#define EF_BLUR_BOX_W 96   // bytes per staged row: columns x0-16 .. x0+79 (inner box extent AND inner start coordinate must be multiples of 16 bytes)
#define EF_BLUR_BOX_H 70   // rows y0-3 .. y0+66

#define EF_RS_BOX_W 176    // source window of one 128 x 64 resize tile: <= 176 columns from a 16-byte aligned start ...
#define EF_RS_BOX_H 78     // ... and <= 78 rows (ratios <= 1.21 x 1.22)

struct alignas(64) EfTmaMaps {
    CUtensorMap blur_src[EF_MAX_LEVELS];   // level images, box EF_BLUR_BOX_W x EF_BLUR_BOX_H x 1 (input of the Gaussian blur)
    CUtensorMap resize_src[EF_MAX_LEVELS]; // [l] = level l-1 image, box EF_RS_BOX_W x EF_RS_BOX_H x 1 (source of pyramid level l)
    unsigned blur_src_ok;                  // bit l: blur_src[l] is valid (base / strides 16-byte aligned)
    unsigned resize_src_ok;                // bit l: resize_src[l] is valid AND level l never clamps a tap (see ef_api.cu)
    unsigned pad_[14];
};

// host: encode a map over `nframes` pitched u8 images; false when the layout does not meet the TMA alignment rules
bool ef_tma_encode_u8(CUtensorMap* map, const void* base, int w, int h, int nframes, size_t pitch, size_t frame_stride, int box_w, int box_h);

#ifdef __CUDACC__
__device__ __forceinline__ unsigned ef_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ef_mbar_init(unsigned mbar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void ef_mbar_expect_tx(unsigned mbar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
// box at (x, y, frame) -> shared memory (128-byte aligned), completion counted in bytes on `mbar`
__device__ __forceinline__ void ef_tma_load_3d(unsigned dst, const CUtensorMap* map, int x, int y, int z, unsigned mbar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(x), "r"(y), "r"(z), "r"(mbar) : "memory");
}
// bounded spin (a lost transaction traps instead of hanging the GPU)
__device__ __forceinline__ void ef_mbar_wait(unsigned mbar, unsigned parity)
{
    unsigned done = 0;
    for (unsigned spin = 0; !done; spin++) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
        if (spin > (1u << 24)) __trap();
    }
}
#endif
